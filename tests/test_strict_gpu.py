"""-m gpu: the STRICT arithmetic modes (fp32 storage, tensor-core convolutions with fp32 accumulation, fp64 BatchNorm
statistics) against the fp64 golden vectors of the UNMODIFIED reference (tests/golden/*.npz) - the gate of SURVEY 8(c)(3):
forward logits, summed loss and BatchNorm running statistics after one training-mode forward.

  Engine(precision="tf32")    one tcgen05.mma kind::tf32 per product.  The tensor core TRUNCATES fp32 operands to 10 mantissa
                              bits (measured: 7.7e-4 per convolution, what two truncated operands give), and the network
                              amplifies that to 4-6e-3 in the logits of the small networks and 1.7e-2 at 640x960 - tf32 alone
                              does NOT meet the 1e-3 the north star quotes for "fp32/tf32" on this network.
  Engine(precision="tf32x3")  3xTF32: operands split into the tf32 head the hardware reads and the exact remainder, three MMAs
                              per product.  Gates (asserted below): convolution relL2 <= 5e-5 vs fp64, network logits <= 1e-3,
                              summed loss <= 1e-5, running statistics <= 1e-5 - the numbers SURVEY 8(c)(3) sets.
The measured values are printed next to the reference's own fp32 and bf16-autocast errors stored in the goldens."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dmmfods_b200 import config as cfgmod
from dmmfods_b200 import ops, synthetic
from dmmfods_b200.model import Dense_U_Net_lidar
from gpu_util import rel_l2
from test_oracle_golden import load_tiny

pytestmark = pytest.mark.gpu


def _f32mat(x, ld=None):
    """(B,C,H,W) float64 CPU -> fp32 pixel-major Mat on the GPU."""
    B, C_, H, W = x.shape
    ld = ld or ops.ceil_to(C_, 8)
    t = torch.zeros(B * H * W, ld, dtype=torch.float32)
    t[:, :C_] = x.permute(0, 2, 3, 1).reshape(-1, C_).float()
    return ops.Mat(t.cuda(), B, H, W)


def _from(m, C_):
    return m.t[:, :C_].double().cpu().reshape(m.B, m.H, m.W, C_).permute(0, 3, 1, 2)


def _pack(w, taps_off, Cin, sn, sc, kwidth=32, split=False):
    """w (Cout,Cin,kh,kw) -> [n_rows][T*Kp] fp32 (split: [2][n_rows][T*Kp] = values, tf32 remainders), like dmm_pack_weights_work_f32."""
    Cout = w.shape[0]
    T = len(taps_off)
    Kp = ops.ceil_to(Cin, kwidth)
    n_tile = ops.pick_n_tile(Cout)
    n_rows = ops.ceil_to(Cout, n_tile)
    out = torch.zeros(n_rows, T * Kp, dtype=torch.float32)
    flat = w.reshape(-1).float()
    for n in range(Cout):
        for t, off in enumerate(taps_off):
            idx = n * sn + torch.arange(Cin) * sc + off
            out[n, t * Kp:t * Kp + Cin] = flat[idx]
    if split:
        hi = (out.view(torch.int32) & -8192).view(torch.float32)          # 0xFFFFE000: the 19 bits the tensor core reads
        out = torch.cat([out, out - hi], 0)
    return out.cuda(), n_tile, n_rows, T * Kp


@pytest.mark.parametrize("dtype", [1, 2])
@pytest.mark.parametrize("Cin,Cout,K,H,W", [(40, 128, 1, 20, 30), (128, 32, 3, 16, 24), (136, 64, 3, 12, 20), (64, 3, 5, 16, 16), (200, 300, 1, 8, 12)])
def test_tf32_convolution_vs_fp64(Cin, Cout, K, H, W, dtype):
    torch.manual_seed(Cin + K)
    B = 2
    x = torch.randn(B, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cout, Cin, K, K, dtype=torch.float64) * 0.1
    ref = F.conv2d(x, w, padding=(K - 1) // 2)
    a = _f32mat(x)
    taps, _, off = ops.conv_taps(K, (K - 1) // 2)
    wp, n_tile, n_rows, ktot = _pack(w, off, Cin, Cin * K * K, K * K, split=dtype == 2)
    st = torch.zeros(ops.Stats.size(ops.ceil_to(Cout, 8)), dtype=torch.float64, device="cuda")
    if Cout <= 16:
        out = torch.zeros(B, Cout, H, W, device="cuda")
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, ktot, n_rows, W, H, B, Cout, out.data_ptr(), 0, out_mode=1, n_tile=n_tile, kwidth=32)
        d.dtype = dtype
        ops.run_igemm(d)
        got = out.double().cpu()
    else:
        o = ops.Mat(torch.zeros(B * H * W, ops.ceil_to(Cout, 8), dtype=torch.float32, device="cuda"), B, H, W)
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, ktot, n_rows, W, H, B, Cout, o.ptr(), o.ld, stats=ops.Stats(st, 0, o.ld),
                           n_tile=n_tile, kwidth=32)
        d.dtype = dtype
        ops.run_igemm(d)
        got = _from(o, Cout)
        s1, s2 = ops.Stats(st, 0, o.ld).totals()
        assert rel_l2(s1[:Cout].cpu(), got.sum(dim=(0, 2, 3))) < 1e-6
        assert rel_l2(s2[:Cout].cpu(), (got ** 2).sum(dim=(0, 2, 3))) < 1e-6
    e = rel_l2(got, ref)
    print("\n[%s conv %dx%d %d->%d] relL2 vs fp64 %.3e" % ("tf32" if dtype == 1 else "3xTF32", K, K, Cin, Cout, e))
    assert e < (1e-3 if dtype == 1 else 5e-5)      # fp32 accumulation over K up to 1600


def _cfg_from(mc):
    c = cfgmod.get_config("/nonexistent")
    for k, v in mc.items():
        setattr(c.model, k, v)
    return c


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
@pytest.mark.parametrize("name", ["no", "early", "mid", "mid_large"])
def test_strict_forward_matches_reference_golden(name, precision):
    g, mc, sd, x1, x2, tgt = load_tiny(name)
    model = Dense_U_Net_lidar(_cfg_from(mc))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    B, _, H, W = x1.shape
    eng = model.engine(B, H, W, training=True, precision=precision)
    logits = eng.forward(x1.cuda(), x2.cuda()).clone()
    sums = eng.loss(tgt.cuda()).clone()
    torch.cuda.synchronize()
    ref_logits = torch.from_numpy(g["logits64"])
    e_logits = rel_l2(logits.cpu(), ref_logits)
    e_loss = abs(sums.sum().item() - g["loss64_sum"][0]) / abs(g["loss64_sum"][0])
    e_lpc = float(np.abs(sums.cpu().numpy() - g["loss64_per_class"]).max() / np.abs(g["loss64_per_class"]).max())
    worst = 0.0
    new = model.state_dict()
    for k in g.files:
        if not k.startswith("new/"):
            continue
        ref = torch.from_numpy(g[k])
        if k.endswith("num_batches_tracked"):
            assert int(new[k[4:]]) == int(ref)
        else:
            worst = max(worst, rel_l2(new[k[4:]].cpu(), ref))
    fp32_err, bf16_err = float(g["ref_fp32_err"][0]), float(g["ref_bf16_autocast_err"][0])
    print("\n[strict %s, %s] logits relL2 %.3e vs the reference's fp64 run (its own fp32 run %.3e, its bf16-autocast run %.3e); "
          "summed loss rel %.3e (per class %.3e); worst running-statistics relL2 %.3e" % (precision, name, e_logits, fp32_err, bf16_err, e_loss, e_lpc, worst))
    if precision == "tf32x3":          # the gate of SURVEY 8(c)(3)
        assert e_logits < 1e-3
        assert e_loss < 1e-5
        assert worst < 1e-5
    else:                              # plain tf32: an order of magnitude better than bf16, short of 1e-3 (see the module docstring)
        assert e_logits < 1e-2
        assert e_loss < 1e-3
        assert worst < 5e-3


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
def test_strict_forward_full_resolution_vs_reference_golden(precision):
    """BASELINE config 3 at 640x960, batch 1: strict forward vs the fp64 golden crop / loss of the unmodified reference."""
    from test_fullsize_gpu import GOLD, MC, H, W, _model_and_state
    model, sd = _model_and_state()
    x1 = torch.from_numpy(synthetic.rgb_image(1, H, W, seed=11))
    x2 = torch.from_numpy(synthetic.lidar_image(1, H, W, seed=12))
    tgt = torch.from_numpy(synthetic.target_maps(1, H, W, seed=13))
    model = model.cuda().train()
    eng = model.engine(1, H, W, training=True, precision=precision)
    logits = eng.forward(x1.cuda(), x2.cuda()).clone()
    sums = eng.loss(tgt.cuda()).clone()
    torch.cuda.synchronize()
    e_crop = rel_l2(logits[:, :, 300:340, 400:480].cpu(), torch.from_numpy(GOLD["logits64_crop"]))
    gl = GOLD["loss64_per_class"]
    e_lpc = float(np.abs(sums.cpu().numpy() - gl).max() / np.abs(gl).max())
    e_norm = abs(logits.double().norm().item() - float(GOLD["logits64_norm"][0])) / float(GOLD["logits64_norm"][0])
    print("\n[strict %s, 640x960 DenseNet-121 mid-fusion] logits crop relL2 %.3e vs the reference's fp64 run (reference fp32 %.3e, "
          "reference bf16-autocast %.3e); loss per class rel %.3e; |logits| rel %.3e"
          % (precision, e_crop, float(GOLD["ref_fp32_err"][0]), float(GOLD["ref_bf16_autocast_err"][0]), e_lpc, e_norm))
    if precision == "tf32x3":
        assert e_crop < 1e-3 and e_lpc < 1e-5
    else:      # 121 BatchNorm-coupled layers amplify the truncation of tf32 operands; bf16 sits at 1.5e-1
        assert e_crop < 3e-2 and e_lpc < 1e-3
