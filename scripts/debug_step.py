"""debug helper: one training step of a tiny model with a synchronize after every launch."""
import os
import sys
os.environ.setdefault("DMM_DEBUG_SYNC", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmmfods_b200 import config as cfgmod, synthetic
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
print("engine built: fwd ops", len(eng.fwd), "bwd ops", len(eng.bwd), "mem MB", eng.mem_bytes / 1e6)
torch.cuda.synchronize()
out = eng.forward(x1, x2)
torch.cuda.synchronize()
print("forward ok", out.float().abs().mean().item(), torch.isfinite(out).all().item())
tgt = torch.from_numpy(synthetic.target_maps(B, H, W)).cuda()
print("loss", eng.loss(tgt).tolist())
import ctypes as C
from dmmfods_b200 import _lib
eng._sums.zero_used()
eng._dw.zero_()
torch.cuda.synchronize(); print("zeroed", flush=True)
eng._run(eng.bwd)
torch.cuda.synchronize(); print("bwd program ok", flush=True)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
print("unpack jobs", eng._n_unpack, eng._unpack_tab.shape, eng._unpack_tab.device, flush=True)
import numpy as np
from dmmfods_b200.engine import _UNPACK_DT
uj = np.frombuffer(eng._unpack_tab.cpu().numpy().tobytes(), dtype=_UNPACK_DT)
base = eng._dw.data_ptr(); gb = eng.gflat.data_ptr()
print("dw base %x n %d ; g base %x n %d" % (base, eng._dw.numel(), gb, eng.gflat.numel()))
for i in range(eng._n_unpack):
    u = uj[i]
    print(i, "dw_off", (int(u["dw"]) - base) // 4, "g_off", (int(u["grad"]) - gb) // 4, u["M"], u["Mld"], u["N"], u["T"], u["ldw"], u["sn"], u["sc"], list(u["tap_off"][:u["T"]]), flush=True)
    rc = eng.lib.dmm_unpack_wgrad_batched(C.c_void_p(eng._unpack_tab.data_ptr() + i * _UNPACK_DT.itemsize), 1, stream)
    torch.cuda.synchronize()
print("all unpack ok")
