"""Parity at the FULL resolution of BASELINE config 3 (mid-fusion DenseNet-121, 640x960 RGB + LiDAR image):
(a) one training step (forward + BCE + backward) of the CUDA path against the CPU oracle on the same seeded inputs at
    batch 1 (the oracle needs a few seconds there), and
(b) a size-independent property that covers the batch dimension: a batch made of k copies of a sample has the SAME
    BatchNorm batch statistics, so every copy reproduces the same logits and the (sum-reduced) loss and all gradients
    scale by k (Agent.py:247-264: the loss is summed over samples).
(c) the BENCHED engine shape (batch 32, 640x960: other tile plans, > 2^31-element buffers) through the same property against
    the batch-1 oracle result.
Tolerances are written in the tests.  The bf16 yard-stick of this 121-layer network at random initialisation is the error of
the UNMODIFIED reference run under torch.autocast('cpu', bfloat16) against its own fp64 run on these very inputs (SURVEY
8(c)(5)), stored with the fp64 golden values in tests/golden/fullsize_yardstick.npz by tests/golden/make_golden.py
(logits 1.49e-1, gradients 6.7e-2; the reference's fp32 run: 1.3e-5 / 8.2e-4)."""
import os

import numpy as np
import pytest
import torch

from dmmfods_b200 import config as cfgmod
from dmmfods_b200 import synthetic
from dmmfods_b200.model import Dense_U_Net_lidar, FusedBCEWithLogits
from gpu_util import rel_l2
from oracle import dense_unet_oracle as du

pytestmark = pytest.mark.gpu
H, W = 640, 960
MC = {"growth_rate": 32, "block_config": (6, 12, 24, 16), "num_init_features": 64, "bn_size": 4,
      "stream_1_in_channels": 3, "stream_2_in_channels": 1, "concat_before_block_num": 3,
      "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}


def _model_and_state():
    c = cfgmod.get_config("/nonexistent")
    for k, v in MC.items():
        setattr(c.model, k, v)
    torch.manual_seed(123)
    model = Dense_U_Net_lidar(c)
    gen = torch.Generator().manual_seed(7)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=gen) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.2
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    return model, sd


def _step(model, x1, x2, tgt):
    model.zero_grad(set_to_none=True)
    logits = model(x1.cuda(), x2.cuda())
    loss = FusedBCEWithLogits()(logits, tgt.cuda())
    loss.backward(torch.ones_like(loss))
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return logits.detach().clone(), loss.detach().double().sum().item(), grads


def _global(grads, ref, scale=1.0):
    num = sum(((grads[k].double().cpu() - scale * ref[k].double().cpu()) ** 2).sum().item() for k in ref)
    den = sum(((scale * ref[k].double()) ** 2).sum().item() for k in ref)
    return (num / den) ** 0.5


GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_yardstick.npz"))


@pytest.fixture(scope="module")
def full_res():
    model, sd = _model_and_state()
    x1 = torch.from_numpy(synthetic.rgb_image(1, H, W, seed=11))
    x2 = torch.from_numpy(synthetic.lidar_image(1, H, W, seed=12))
    tgt = torch.from_numpy(synthetic.target_maps(1, H, W, seed=13))
    model = model.cuda().train()
    return model, sd, x1, x2, tgt


@pytest.fixture(scope="module")
def oracle_b1(full_res):
    """the CPU oracle at batch 1 (bf16-emulated and exact fp32), computed once for the tests of this module."""
    _, sd, x1, x2, tgt = full_res
    emu = du.oracle_train_step(sd, MC, x1, x2, tgt, dtype=torch.float32, emulate_bf16=True)
    ref = du.oracle_train_step(sd, MC, x1, x2, tgt, dtype=torch.float32)
    return emu, ref


def test_state_and_inputs_are_the_golden_ones(full_res):
    """the fixture reproduces the state the reference was run on when the golden file was made (same init sequence)."""
    _, sd, x1, x2, tgt = full_res
    chk = sum(v.double().sum().item() for v in sd.values() if v.is_floating_point())
    assert abs(chk - float(GOLD["param_checksum"][0])) <= 1e-6 * abs(chk)
    assert tuple(GOLD["seeds"]) == (11, 12, 13) and tuple(GOLD["shape"]) == (1, H, W)


def test_full_resolution_train_step_matches_oracle(full_res, oracle_b1):
    """vs the bf16-EMULATED fp32 oracle (same rounding points: the check of the implementation), vs the exact fp32 oracle and
    vs the fp64 golden values of the unmodified reference; yard-stick = the reference's own torch.autocast(bf16) error."""
    model, sd, x1, x2, tgt = full_res
    model.load_state_dict(sd, strict=True)
    logits, loss_sum, grads = _step(model, x1, x2, tgt)
    emu, ref = oracle_b1
    yard_logits, yard_grad = (float(v) for v in GOLD["ref_bf16_autocast_err"])
    own_yard = rel_l2(emu["logits"], ref["logits"])
    # golden values of the reference's fp64 run: per-class loss sums, a logits crop, two gradient tensors, all gradient norms
    gl = GOLD["loss64_per_class"]
    lpc = FusedBCEWithLogits()(logits.cuda(), tgt.cuda()).double().sum(dim=(0, 2, 3)).cpu().numpy()
    e_lpc = float(np.abs(lpc - gl).max() / np.abs(gl).max())
    crop = torch.from_numpy(GOLD["logits64_crop"])
    e_crop = rel_l2(logits[:, :, 300:340, 400:480].cpu(), crop)
    e_g_r1 = rel_l2(grads["dec_out_to_heat_maps.refine1.weight"].cpu(), torch.from_numpy(GOLD["grad64_refine1"]))
    e_g_c0 = rel_l2(grads["features.conv0.weight"].cpu(), torch.from_numpy(GOLD["grad64_conv0"]))
    gn = dict(zip([str(n) for n in GOLD["grad64_names"]], GOLD["grad64_norm"]))
    e_norm = np.median([abs(grads[k].double().norm().item() - gn[k]) / gn[k] for k in gn if gn[k] > 0])
    print("\n[640x960 B=1 vs the reference's fp64 golden] loss per class rel %.3e; logits crop relL2 %.3e; d refine1.weight %.3e; "
          "d features.conv0.weight %.3e; median |grad| error %.3e (reference autocast(bf16): logits %.3e grads %.3e; own emulation %.3e)"
          % (e_lpc, e_crop, e_g_r1, e_g_c0, e_norm, yard_logits, yard_grad, own_yard))
    assert e_lpc < 2e-3
    assert e_crop < 1.5 * yard_logits + 2e-2
    # the head's gradient is well conditioned; d features.conv0.weight - 121 BatchNorm-coupled layers away from the loss - is not
    # (the FP32 reference is already 9e-2 off fp64 on the stem at small sizes; under bf16 it is O(1)): printed only
    assert e_g_r1 < 1.5 * yard_grad + 2e-2
    assert e_norm < 5e-2
    e_logits, e_logits_exact = rel_l2(logits.cpu(), emu["logits"]), rel_l2(logits.cpu(), ref["logits"])
    ref_loss = ref["loss"].double().sum().item()
    e_loss = abs(loss_sum - ref_loss) / abs(ref_loss)
    e_grad, e_grad_exact = _global(grads, emu["grads"]), _global(grads, ref["grads"])
    print("\n[640x960 DenseNet-121 mid-fusion, B=1] logits relL2 %.3e vs bf16-emulated oracle, %.3e vs exact fp32 (yard-stick "
          "emulated vs exact %.3e); loss rel %.3e; grad global relL2 %.3e vs emulated, %.3e vs exact (yard-stick %.3e)"
          % (e_logits, e_logits_exact, yard_logits, e_loss, e_grad, e_grad_exact, yard_grad))
    assert e_loss < 2e-3
    assert e_logits < max(4e-2, 0.75 * yard_logits)         # measured 0.49 x; two uncorrelated realisations would sit at 1.41 x
    assert e_logits_exact < 1.5 * yard_logits + 2e-2
    assert e_grad < 1.5e-1
    assert e_grad_exact < 1.5 * yard_grad + 2e-2


def test_benched_batch32_engine_against_the_batch1_oracle(full_res, oracle_b1):
    """the engine plan bench.py times (B = 32 at 640x960: msub / tile choices, 64-bit offsets beyond 2^31 elements): 32 copies
    of the golden sample share its BatchNorm statistics, so every copy must reproduce the batch-1 oracle logits, the summed
    loss is 32 x and every gradient 32 x the oracle's.  Same tolerances as the batch-1 test."""
    model, sd, x1, x2, tgt = full_res
    emu, ref = oracle_b1
    yard_logits, yard_grad = (float(v) for v in GOLD["ref_bf16_autocast_err"])
    k = 32
    model.load_state_dict(sd, strict=True)
    lk, sk, gk = _step(model, x1.repeat(k, 1, 1, 1), x2.repeat(k, 1, 1, 1), tgt.repeat(k, 1, 1, 1))
    e_first = rel_l2(lk[0:1].cpu(), emu["logits"])
    e_last = rel_l2(lk[k - 1:k].cpu(), emu["logits"])
    e_copy = max(rel_l2(lk[i:i + 1].cpu(), lk[0:1].cpu()) for i in (1, 7, 16, 31))
    ref_loss = ref["loss"].double().sum().item()
    e_loss = abs(sk - k * ref_loss) / abs(k * ref_loss)
    e_grad = _global(gk, emu["grads"], scale=float(k))
    e_grad_exact = _global(gk, ref["grads"], scale=float(k))
    print("\n[B=32 at 640x960, the benched plan] logits relL2 vs bf16-emulated oracle %.3e (copy 0) %.3e (copy 31); copy-vs-copy %.3e; "
          "loss rel %.3e; gradient relL2 %.3e vs emulated, %.3e vs exact fp32 (reference autocast(bf16) %.3e / %.3e)"
          % (e_first, e_last, e_copy, e_loss, e_grad, e_grad_exact, yard_logits, yard_grad))
    assert e_copy < 1e-5
    assert max(e_first, e_last) < max(4e-2, 0.75 * yard_logits)
    assert e_loss < 2e-3
    assert e_grad < 1.5e-1
    assert e_grad_exact < 1.5 * yard_grad + 2e-2
    model._engines = {}                      # release the 70 GB batch-32 engine (reference cycles: needs the collector)
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def test_batch_replication_property_at_full_resolution(full_res):
    """k copies of one sample share its BatchNorm statistics: inside the replicated batch every copy must produce the SAME
    logits (same tiles, same arithmetic per image: 1e-5), and against the single-sample run the copies, the summed loss / k
    and the gradients / k agree up to the distance between two bf16 realisations of this network (different accumulation
    orders flip bf16 roundings that 121 layers amplify; the CUDA path sits 7e-2 from the bf16-emulated oracle)."""
    model, sd, x1, x2, tgt = full_res
    model.load_state_dict(sd, strict=True)
    l1, s1, g1 = _step(model, x1, x2, tgt)
    model.load_state_dict(sd, strict=True)                      # same running statistics as before the first step
    k = 2
    lk, sk, gk = _step(model, x1.repeat(k, 1, 1, 1), x2.repeat(k, 1, 1, 1), tgt.repeat(k, 1, 1, 1))
    e_copy = rel_l2(lk[1:2].cpu(), lk[0:1].cpu())
    e_single = rel_l2(lk[0:1].cpu(), l1.cpu())
    e_loss = abs(sk - k * s1) / abs(k * s1)
    e_grad = _global(gk, g1, scale=float(k))
    print("\n[replication x%d at 640x960] copy-vs-copy logits relL2 %.3e; copy vs single run %.3e; loss rel %.3e; gradient relL2 %.3e"
          % (k, e_copy, e_single, e_loss, e_grad))
    assert e_copy < 1e-5
    assert e_single < 1e-1
    assert e_loss < 1e-3
    assert e_grad < 1.5e-1
