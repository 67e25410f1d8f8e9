"""debug helper: find the first launch that overwrites the unpack job table."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmmfods_b200 import _lib, config as cfgmod, synthetic
from dmmfods_b200.model import Dense_U_Net_lidar

c2, cb = int(sys.argv[1]), int(sys.argv[2])
c = cfgmod.get_config("/nonexistent")
for k, v in dict(growth_rate=16, block_config=(2, 2, 2, 2), num_init_features=32, bn_size=2, stream_2_in_channels=c2,
                 concat_before_block_num=cb).items():
    setattr(c.model, k, v)
torch.manual_seed(0)
m = Dense_U_Net_lidar(c).cuda().train()
B, H, W = 2, 64, 96
x1 = torch.from_numpy(synthetic.rgb_image(B, H, W)).cuda()
x2 = torch.from_numpy(synthetic.lidar_image(B, H, W)).cuda()
eng = m.engine(B, H, W)
ref_u = eng._unpack_tab.clone()
ref_p = eng._pack_tab.clone()
torch.cuda.synchronize()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(name):
    torch.cuda.synchronize()
    bad_u = not torch.equal(ref_u, eng._unpack_tab)
    bad_p = not torch.equal(ref_p, eng._pack_tab)
    if bad_u or bad_p:
        d = (ref_u != eng._unpack_tab).nonzero().flatten()
        print("CORRUPTED after", name, "unpack" if bad_u else "", "pack" if bad_p else "", "first/last byte", d[:1].tolist(), d[-1:].tolist(), "table ptr %x" % eng._unpack_tab.data_ptr(), flush=True)
        return True
    return False


eng.in1.copy_(x1)
if eng.c2:
    eng.in2.copy_(x2)
eng._stats.zero_used()
check("zero")
_lib.check(eng.lib.dmm_pack_weights_batched(C.c_void_p(eng._pack_tab.data_ptr()), eng._n_pack, stream), "pack")
check("pack")
for prog in (eng.fwd, None, eng.bwd):
    if prog is None:
        tgt = torch.from_numpy(synthetic.target_maps(B, H, W)).cuda()
        eng.loss(tgt)
        eng._sums.zero_used()
        eng._dw.zero_()
        if check("loss/zero"):
            break
        continue
    stop = False
    for op in prog:
        rc = op.fn(C.byref(op.arg), stream) if op.arg is not None else op.fn(None, stream)
        assert rc == 0, (op.name, _lib.last_error())
        if check(op.name):
            a = op.arg
            if a is not None:
                print({f[0]: getattr(a, f[0]) for f in a._fields_ if not hasattr(getattr(a, f[0]), "_fields_") and not hasattr(getattr(a, f[0]), "_length_")})
            stop = True
            break
    if stop:
        break
print("done")
