"""CPU tests: the oracle (oracle/) is pinned against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py), and - when /root/reference is present - against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import dense_unet_oracle as du
from oracle import lidar_heatmap_oracle as orc
from oracle import ref_shim

GDIR = os.path.join(os.path.dirname(__file__), "golden")
GOLD = np.load(os.path.join(GDIR, "lidar_heatmap.npz"))


def _labels(boxes):
    return {str(i): {"type": int(b[0]), "x": int(b[1]), "y": int(b[2]), "width": int(b[3]), "height": int(b[4])}
            for i, b in enumerate(boxes)}


@pytest.mark.parametrize("name", ["small", "edge"])
def test_scatter_oracle_matches_reference_golden(name):
    shape = tuple(int(v) for v in GOLD["%s_shape" % name])
    img = orc.lidar_array_to_image(GOLD["%s_points" % name], shape, 5)
    assert np.array_equal(img, GOLD["%s_img" % name])
    assert np.array_equal(orc.pool_lidar(img), GOLD["%s_pooled" % name])
    maps = orc.create_ground_truth_maps(_labels(GOLD["%s_boxes" % name]), shape[2], shape[1])
    assert np.array_equal(maps, GOLD["%s_maps" % name])


def test_scatter_oracle_full_resolution_golden():
    img = orc.lidar_array_to_image(GOLD["full_points"], (1, 1280, 1920), 5)
    assert np.array_equal(img[:, 600:700, 900:1100], GOLD["full_img_crop"])
    assert img.astype(np.float64).sum() == GOLD["full_img_sum"][0]
    assert np.array_equal(orc.pool_lidar(img), GOLD["full_pooled"])
    maps = orc.create_ground_truth_maps(_labels(GOLD["full_boxes"]))
    assert maps.astype(np.float64).sum() == GOLD["full_maps_sum"][0]
    assert np.array_equal(orc.maxpool(maps, 10), GOLD["full_maps_pooled"])


def load_tiny(name):
    """golden of the small Dense_U_Net_lidar; inputs are regenerated from the stored seeds and checked
    against the stored checksums; the large case shares the state_dict of tiny_unet_mid.npz."""
    from dmmfods_b200 import synthetic
    g = np.load(os.path.join(GDIR, "tiny_unet_%s.npz" % name))
    gr, b0, b1, b2, b3, nif, bns, c2, cb = (int(v) for v in g["model_cfg"])
    cfg = {"growth_rate": gr, "block_config": (b0, b1, b2, b3), "num_init_features": nif, "bn_size": bns,
           "stream_1_in_channels": 3, "stream_2_in_channels": c2, "concat_before_block_num": cb,
           "num_layers_before_blocks": 4, "drop_rate": 0, "num_classes": 3, "memory_efficient": False}
    gs = g if any(k.startswith("sd/") for k in g.files) else np.load(os.path.join(GDIR, "tiny_unet_mid.npz"))
    sd = {k[3:]: torch.from_numpy(gs[k]) for k in gs.files if k.startswith("sd/")}
    B, H, W = (int(v) for v in g["shape"])
    s1, s2, s3 = (int(v) for v in g["seeds"])
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=s1))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=s2))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W, seed=s3))
    chk = np.array([x1.double().sum().item(), x2.double().sum().item(), tgt.double().sum().item()])
    assert np.array_equal(chk, g["input_checksum"]), "synthetic generator drifted from the golden inputs"
    return g, cfg, sd, x1, x2, tgt


@pytest.mark.parametrize("fusion", ["no", "early", "mid"])
def test_unet_oracle_matches_reference_golden(fusion):
    g, cfg, sd, x1, x2, tgt = load_tiny(fusion)
    r = du.oracle_train_step(sd, cfg, x1, x2, tgt, dtype=torch.float32)
    assert torch.allclose(r["logits"], torch.from_numpy(g["logits32"]), rtol=1e-4, atol=1e-4)
    assert torch.allclose(r["loss"], torch.from_numpy(g["loss32"]), rtol=1e-4, atol=1e-4)
    r64 = du.oracle_train_step(sd, cfg, x1, x2, tgt, dtype=torch.float64)
    assert torch.allclose(r64["logits"].float(), torch.from_numpy(g["logits64"]), rtol=1e-5, atol=1e-5)
    assert abs(r64["loss"].sum().item() - g["loss64_sum"][0]) <= 1e-9 * abs(g["loss64_sum"][0])
    for k, v in r64["grads"].items():
        ref = torch.from_numpy(g["grad64/" + k]).double()
        err = (v - ref).norm() / (ref.norm() + 1e-30)
        assert err < 1e-5, (k, err.item())
    for k, v in r64["new_stats"].items():
        assert torch.allclose(v.float(), torch.from_numpy(g["new/" + k]).float(), rtol=1e-5, atol=1e-6), k
    # eval mode with the updated running statistics
    sd_eval = dict(sd)
    sd_eval.update({k: v for k, v in r64["new_stats"].items()})
    full = {k: (v.double() if v.is_floating_point() else v) for k, v in sd_eval.items()}
    lg, _ = du.oracle_forward(full, cfg, x1.double(), x2.double(), train=False)
    assert torch.allclose(lg.float(), torch.from_numpy(g["eval_logits64"]), rtol=1e-4, atol=1e-4)


def test_unet_oracle_bf16_emulation_is_a_small_perturbation():
    """the storage-precision emulation used to check the CUDA engine tightly stays within the reference's own
    bf16-autocast error level of the exact arithmetic."""
    g, cfg, sd, x1, x2, tgt = load_tiny("mid")
    r = du.oracle_train_step(sd, cfg, x1, x2, tgt, dtype=torch.float64, emulate_bf16=True)
    ref = torch.from_numpy(g["logits64"]).double()
    e = ((r["logits"] - ref).norm() / ref.norm()).item()
    assert 1e-3 < e < 1.5 * g["ref_bf16_autocast_err"][0], e


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_unet_oracle_matches_live_reference():
    model_mod, _ = ref_shim.ref_modules()
    cfg = ref_shim.ref_config(stream_2_in_channels=1, concat_before_block_num=2, growth_rate=8, block_config=(1, 2, 1, 1),
                              num_init_features=16, bn_size=2)
    torch.manual_seed(0)
    model = model_mod.Dense_U_Net_lidar(cfg).double()
    x1 = torch.rand(1, 3, 32, 64, dtype=torch.float64) * 255
    x2 = torch.rand(1, 1, 32, 64, dtype=torch.float64) * 255
    ref = model(x1, x2)
    got, _ = du.oracle_forward(model.state_dict(), dict(cfg.model), x1, x2, train=True)
    assert torch.allclose(got, ref.detach(), rtol=1e-9, atol=1e-9)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_metric_oracle_matches_live_reference():
    """numpy restatement of compute_IoU_whole_img_batch / compute_accuracy (helper:311-401) vs the reference itself."""
    import numpy as np
    from oracle import lidar_heatmap_oracle as orc
    _, helper_mod = ref_shim.ref_modules()
    rng = np.random.default_rng(7)
    gt = rng.choice(np.array([0, 0.3, 0.5, 0.75, 1.0], dtype=np.float32), size=(3, 3, 24, 40), p=[0.6, 0.1, 0.1, 0.1, 0.1])
    gt[1, 2] = 0                                        # a class without any box: IoU 0/0 = nan in the reference
    pred = (rng.standard_normal((3, 3, 24, 40)) * 1.5).astype(np.float32)
    pred[1, 2] = -5
    ref_iou = helper_mod.compute_IoU_whole_img_batch(torch.from_numpy(gt), torch.from_numpy(pred), 0.7).numpy()
    got = orc.iou_whole_img_batch(gt, pred, 0.7)
    assert np.array_equal(np.isnan(ref_iou), np.isnan(got)) and np.isnan(got[1, 2])
    assert np.array_equal(np.nan_to_num(ref_iou), np.nan_to_num(got))
    ref_acc = helper_mod.compute_accuracy(torch.from_numpy(gt), torch.from_numpy(pred), 0.7).numpy()
    assert np.allclose(ref_acc, orc.accuracy(gt, pred, 0.7), rtol=0, atol=1e-7)
    ref_acc1 = helper_mod.compute_accuracy(torch.from_numpy(gt[0]), torch.from_numpy(pred[0]), 0.7).numpy()
    assert np.allclose(ref_acc1, orc.accuracy(gt[0], pred[0], 0.7), rtol=0, atol=1e-7)
