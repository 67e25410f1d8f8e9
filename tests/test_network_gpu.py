"""End-to-end parity of the CUDA Dense-U-Net (forward, BCE loss, backward) with
 (i) golden vectors produced by the UNMODIFIED reference (tests/golden/tiny_unet_*.npz: fp32 and fp64 runs), and
 (ii) the CPU oracle (oracle/dense_unet_oracle.py) on other shapes / fusion settings.

Arithmetic: bf16 storage, fp32 tensor-core accumulation, fp32/fp64 BatchNorm statistics.
Tolerances (stated per BASELINE.json north_star: 2e-2 for bf16):
  logits relL2 vs fp64 reference  <= 2e-2
  summed loss                     <= 2e-3 relative
  BN running statistics           <= 2e-2 relL2
  gradients: global vector relL2 vs fp64 <= 5e-2, and no worse than 1.5x the error of the REFERENCE ITSELF
  run under torch.autocast(bfloat16) on CPU (the yard-stick of SURVEY.md section 8(c) item 5, committed in the goldens).
"""
import os

import numpy as np
import pytest
import torch

from dmmfods_b200 import config as cfgmod
from dmmfods_b200.model import Dense_U_Net_lidar, FusedBCEWithLogits
from gpu_util import rel_l2
from oracle import dense_unet_oracle as du

pytestmark = pytest.mark.gpu
GDIR = os.path.join(os.path.dirname(__file__), "golden")


def _cfg_from(mc):
    c = cfgmod.get_config("/nonexistent")
    for k, v in mc.items():
        setattr(c.model, k, v)
    return c


def _load_tiny(fusion):
    g = np.load(os.path.join(GDIR, "tiny_unet_%s.npz" % fusion))
    gr, b0, b1, b2, b3, nif, bns, c2, cb = (int(v) for v in g["model_cfg"])
    mc = {"growth_rate": gr, "block_config": (b0, b1, b2, b3), "num_init_features": nif, "bn_size": bns,
          "stream_1_in_channels": 3, "stream_2_in_channels": c2, "concat_before_block_num": cb,
          "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}
    return g, mc, sd


def _global_err(grads, ref):
    num = sum(((grads[k].double().cpu() - ref[k].double()) ** 2).sum().item() for k in ref)
    den = sum((ref[k].double() ** 2).sum().item() for k in ref)
    return (num / den) ** 0.5


@pytest.mark.parametrize("fusion", ["no", "early", "mid"])
def test_train_step_matches_reference_golden(fusion):
    g, mc, sd = _load_tiny(fusion)
    model = Dense_U_Net_lidar(_cfg_from(mc))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    x1, x2, tgt = (torch.from_numpy(g[k]).cuda() for k in ("x1", "x2", "target"))
    logits = model(x1, x2)
    loss = FusedBCEWithLogits()(logits, tgt)
    loss.backward(torch.ones_like(loss))
    torch.cuda.synchronize()

    ref_logits = torch.from_numpy(g["logits64"])
    e_logits = rel_l2(logits.detach().cpu(), ref_logits)
    e_ref_bf16 = rel_l2(torch.from_numpy(g["logits_bf16_autocast"]), ref_logits)
    e_ref_fp32 = rel_l2(torch.from_numpy(g["logits32"]), ref_logits)
    loss_sum = loss.detach().double().sum().item()
    e_loss = abs(loss_sum - g["loss64_sum"][0]) / abs(g["loss64_sum"][0])

    ref64 = {k[7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("grad64/")}
    refbf = {k[9:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("gradbf16/")}
    ref32 = {k[7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("grad32/")}
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert set(grads) == set(ref64)
    assert all(v is not None for v in grads.values())
    e_grad = _global_err(grads, ref64)
    e_grad_refbf = _global_err(refbf, ref64)
    e_grad_ref32 = _global_err(ref32, ref64)
    per = sorted(((rel_l2(grads[k].cpu(), ref64[k]), k) for k in ref64), reverse=True)
    med = per[len(per) // 2][0]
    print("\n[%s] logits relL2 %.3e (reference bf16-autocast %.3e, reference fp32 %.3e); loss rel %.3e; "
          "grad global %.3e (reference bf16-autocast %.3e, fp32 %.3e); per-tensor median %.3e worst %.3e %s"
          % (fusion, e_logits, e_ref_bf16, e_ref_fp32, e_loss, e_grad, e_grad_refbf, e_grad_ref32, med, per[0][0], per[0][1]))
    assert e_logits < 2e-2
    assert e_loss < 2e-3
    assert e_grad < 5e-2
    assert e_grad < 1.5 * e_grad_refbf

    new = model.state_dict()
    for k in g.files:
        if not k.startswith("new/"):
            continue
        name = k[4:]
        ref = torch.from_numpy(g[k])
        if name.endswith("num_batches_tracked"):
            assert int(new[name]) == int(ref)
        else:
            e = rel_l2(new[name].cpu(), ref)
            assert e < 2e-2, "%s relL2 %.3e" % (name, e)


def test_eval_mode_forward_matches_golden():
    g, mc, sd = _load_tiny("mid")
    for k in g.files:
        if k.startswith("new/"):
            sd[k[4:]] = torch.from_numpy(g[k])
    model = Dense_U_Net_lidar(_cfg_from(mc))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(torch.from_numpy(g["x1"]).cuda(), torch.from_numpy(g["x2"]).cuda())
    e = rel_l2(out.cpu(), torch.from_numpy(g["eval_logits64"]))
    print("\neval logits relL2 %.3e" % e)
    assert e < 2e-2


@pytest.mark.parametrize("c2,cb,B,H,W", [(1, 2, 1, 64, 64), (1, 4, 3, 32, 64), (0, 1, 2, 32, 32)])
def test_train_step_matches_oracle_other_shapes(c2, cb, B, H, W):
    """fresh seeded weights/inputs, CUDA path vs the CPU oracle evaluated in fp64 on the same box."""
    mc = {"growth_rate": 16, "block_config": (2, 3, 2, 2), "num_init_features": 32, "bn_size": 2,
          "stream_1_in_channels": 3, "stream_2_in_channels": c2, "concat_before_block_num": cb,
          "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}
    torch.manual_seed(100 + cb)
    model = Dense_U_Net_lidar(_cfg_from(mc))
    gen = torch.Generator().manual_seed(5)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=gen) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.2
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    from dmmfods_b200 import synthetic
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=1))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=2))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W, seed=3))
    ref = du.oracle_train_step(sd, mc, x1, x2, tgt, dtype=torch.float64)
    model = model.cuda().train()
    logits = model(x1.cuda(), x2.cuda())
    loss = FusedBCEWithLogits()(logits, tgt.cuda())
    loss.backward(torch.ones_like(loss))
    torch.cuda.synchronize()
    e_logits = rel_l2(logits.detach().cpu(), ref["logits"])
    grads = {k: p.grad for k, p in model.named_parameters()}
    e_grad = _global_err(grads, ref["grads"])
    print("\n[c2=%d cb=%d %dx%dx%d] logits relL2 %.3e grad global %.3e" % (c2, cb, B, H, W, e_logits, e_grad))
    assert e_logits < 2e-2
    assert e_grad < 5e-2


def test_repeatable_and_shape_switching():
    """two forwards of the same input give identical logits; engines for two shapes coexist."""
    g, mc, sd = _load_tiny("mid")
    model = Dense_U_Net_lidar(_cfg_from(mc))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    x1 = torch.from_numpy(g["x1"]).cuda()
    x2 = torch.from_numpy(g["x2"]).cuda()
    with torch.no_grad():
        a = model(x1, x2)
        b = model(x1[:1, :, :32, :64].contiguous(), x2[:1, :, :32, :64].contiguous())
        c = model(x1, x2)
    assert torch.equal(a, c)
    assert b.shape == (1, 3, 32, 64)


def test_shape_errors_like_reference():
    g, mc, sd = _load_tiny("mid")
    model = Dense_U_Net_lidar(_cfg_from(mc)).cuda()
    with pytest.raises(AssertionError):
        model(torch.zeros(1, 3, 64, 96).cuda(), torch.zeros(1, 1, 32, 96).cuda())
    with pytest.raises((ValueError, RuntimeError)):
        model(torch.zeros(1, 3, 72, 96).cuda(), torch.zeros(1, 1, 72, 96).cuda())
