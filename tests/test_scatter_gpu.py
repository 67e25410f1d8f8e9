"""Bit-exact parity of the integer-scatter kernels (LiDAR splat, range transform + pool, heat-map masks,
10x10 pooling) with (i) the golden vectors produced by the unmodified reference and (ii) the numpy oracle
on fresh seeded inputs, through the host mirror dmmfods_b200.helper (reference function names)."""
import os

import numpy as np
import pytest
import torch

from dmmfods_b200 import helper, synthetic
from oracle import lidar_heatmap_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lidar_heatmap.npz"))


def _labels(boxes):
    return {str(i): {"type": int(b[0]), "x": int(b[1]), "y": int(b[2]), "width": int(b[3]), "height": int(b[4])}
            for i, b in enumerate(boxes)}


@pytest.mark.parametrize("name", ["small", "edge"])
def test_golden_small_cases(name):
    shape = tuple(int(v) for v in GOLD["%s_shape" % name])
    img = helper.lidar_array_to_image_like_tensor(GOLD["%s_points" % name], shape=shape, kernel_size=5)
    assert np.array_equal(img.cpu().numpy(), GOLD["%s_img" % name])
    pooled = helper.pool_lidar_tensor(img)
    assert np.array_equal(pooled.cpu().numpy(), GOLD["%s_pooled" % name])
    maps = helper.create_ground_truth_maps(_labels(GOLD["%s_boxes" % name]), width_img=shape[2], height_img=shape[1])
    assert np.array_equal(maps.cpu().numpy(), GOLD["%s_maps" % name])


def test_golden_full_resolution():
    """Waymo FRONT camera resolution (1,1280,1920), 30 000 points (BASELINE config 4 pre-processing)."""
    pts = GOLD["full_points"]
    img = helper.lidar_array_to_image_like_tensor(pts)
    a = img.cpu().numpy()
    assert np.array_equal(a[:, 600:700, 900:1100], GOLD["full_img_crop"])
    assert a.astype(np.float64).sum() == GOLD["full_img_sum"][0]
    assert (a.astype(np.float64) ** 2).sum() == GOLD["full_img_sum"][1]
    assert np.array_equal(helper.pool_lidar_tensor(img).cpu().numpy(), GOLD["full_pooled"])
    maps = helper.create_ground_truth_maps(_labels(GOLD["full_boxes"]))
    m = maps.cpu().numpy()
    assert m.astype(np.float64).sum() == GOLD["full_maps_sum"][0]
    assert np.array_equal(helper.maxpool_tensor(maps).cpu().numpy(), GOLD["full_maps_pooled"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_against_oracle_random(seed):
    H, W = 200 + 10 * seed, 300 + 7 * seed
    pts = synthetic.lidar_points(5000, H, W, seed=seed, out_of_range=0.05)
    img = helper.lidar_array_to_image_like_tensor(pts, shape=(1, H, W))
    ref = orc.lidar_array_to_image(pts, (1, H, W), 5)
    assert np.array_equal(img.cpu().numpy(), ref)
    assert np.array_equal(helper.pool_lidar_tensor(img).cpu().numpy(), orc.pool_lidar(ref))
    labels = synthetic.boxes(40, H, W, seed=seed)
    maps = helper.create_ground_truth_maps(labels, width_img=W, height_img=H)
    refm = orc.create_ground_truth_maps(labels, W, H)
    assert np.array_equal(maps.cpu().numpy(), refm)
    assert np.array_equal(helper.maxpool_tensor(maps).cpu().numpy(), orc.maxpool(refm, 10))
    x = torch.from_numpy(synthetic.rgb_image(1, H, W, seed=seed)[0])
    got = helper.avgpool_tensor(x).cpu()
    assert torch.equal(got, torch.nn.AvgPool2d(10, stride=10)(x))


def test_kernel_size_and_empty_inputs():
    H, W = 50, 70
    pts = synthetic.lidar_points(300, H, W, seed=9)
    for k in (1, 3, 7):
        got = helper.lidar_array_to_image_like_tensor(pts, shape=(1, H, W), kernel_size=k).cpu().numpy()
        assert np.array_equal(got, orc.lidar_array_to_image(pts, (1, H, W), k))
    empty = helper.lidar_array_to_image_like_tensor(np.zeros((0, 3), np.float32), shape=(1, H, W)).cpu().numpy()
    assert (empty == -1).all()
    maps = helper.create_ground_truth_maps({}, width_img=W, height_img=H).cpu().numpy()
    assert (maps == 0).all()
    with pytest.raises(ValueError):
        helper.create_ground_truth_maps({"0": {"type": 1, "x": W - 2, "y": 0, "width": 5, "height": 5}}, W, H)


def test_idempotent_and_order_dependent():
    """size-independent properties: repainting the same list is a no-op; reversing it changes overlaps."""
    H, W = 1280, 1920
    pts = synthetic.lidar_points(30000, H, W, seed=77)
    a = helper.lidar_array_to_image_like_tensor(pts)
    b = helper.lidar_array_to_image_like_tensor(np.concatenate([pts, pts], 0))
    assert torch.equal(a, b)
    c = helper.lidar_array_to_image_like_tensor(pts[::-1].copy())
    assert not torch.equal(a, c)
    assert torch.equal(a != -1, c != -1)


@pytest.mark.parametrize("seed,shape", [(1, (2, 3, 128, 192)), (2, (3, 3, 37, 53)), (3, (1, 3, 640, 960))])
def test_step_metrics_match_oracle(seed, shape):
    """fused IoU / accuracy counters (dmm_step_metrics) == the numpy restatement of helper:311-401, incl. nan for 0/0."""
    rng = np.random.default_rng(seed)
    gt = rng.choice(np.array([0, 0.3, 0.5, 0.75, 1.0], dtype=np.float32), size=shape, p=[0.7, 0.05, 0.05, 0.1, 0.1])
    pred = (rng.standard_normal(shape) * 2).astype(np.float32)
    gt[0, 1] = 0
    pred[0, 1] = -3.0                                   # empty union -> nan
    iou = helper.compute_IoU_whole_img_batch(torch.from_numpy(gt), torch.from_numpy(pred), 0.7).cpu().numpy()
    want = orc.iou_whole_img_batch(gt, pred, 0.7)
    assert np.array_equal(np.isnan(iou), np.isnan(want)) and np.isnan(iou[0, 1])
    assert np.array_equal(np.nan_to_num(iou), np.nan_to_num(want))
    acc = helper.compute_accuracy(torch.from_numpy(gt), torch.from_numpy(pred), 0.7).cpu().numpy()
    assert np.allclose(acc, orc.accuracy(gt, pred, 0.7), rtol=0, atol=1e-7)
    acc1 = helper.compute_accuracy(torch.from_numpy(gt[0]), torch.from_numpy(pred[0]), 0.7).cpu().numpy()
    assert np.allclose(acc1, orc.accuracy(gt[0], pred[0], 0.7), rtol=0, atol=1e-7)
    last = shape[0] - 1
    iou1 = helper.compute_IoU_whole_img_per_class(torch.from_numpy(gt[last]), torch.from_numpy(pred[last]), 0.5).cpu().numpy()
    assert np.array_equal(np.nan_to_num(iou1), np.nan_to_num(orc.iou_whole_img_batch(gt[last:], pred[last:], 0.5)[0]))


def test_batched_single_pass_forms_equal_the_per_frame_reference():
    """BASELINE config 4 pre-processing: dmm_lidar_splat_batched / dmm_heatmap_boxes_batched (one launch per batch, every pixel
    written once) against the golden full-resolution frame of the unmodified reference and against the numpy oracle on
    ragged frames: an empty frame, a frame whose boxes exceed one shared-memory chunk (> 512), out-of-range points."""
    H, W = 1280, 1920
    pts0, boxes0 = GOLD["full_points"], GOLD["full_boxes"]
    frames_p = [pts0, np.zeros((0, 3), np.float32), synthetic.lidar_points(12000, H, W, seed=5, out_of_range=0.05), pts0[:7]]
    many = synthetic.boxes(700, H, W, seed=9)
    frames_b = [_labels(boxes0), {}, many, synthetic.boxes(3, H, W, seed=10)]
    pre = helper.BatchPreprocessor(4, H, W, max_points=40000, max_boxes=800, transform=False)
    pre.load(frames_p, frames_b)
    img = torch.full((4, 1, H, W), 123.0, device="cuda")
    maps = torch.full((4, 3, H, W), 123.0, device="cuda")
    pre.run(img, maps)
    a, m = img.cpu().numpy(), maps.cpu().numpy()
    # frame 0 = the golden frame of the unmodified reference
    assert np.array_equal(a[0][:, 600:700, 900:1100], GOLD["full_img_crop"])
    assert a[0].astype(np.float64).sum() == GOLD["full_img_sum"][0] and (a[0].astype(np.float64) ** 2).sum() == GOLD["full_img_sum"][1]
    assert m[0].astype(np.float64).sum() == GOLD["full_maps_sum"][0]
    assert np.array_equal(helper.maxpool_tensor(maps[0]).cpu().numpy(), GOLD["full_maps_pooled"])
    for b in range(4):
        assert np.array_equal(a[b], orc.lidar_array_to_image(frames_p[b], (1, H, W), 5)), "lidar frame %d" % b
        assert np.array_equal(m[b], orc.create_ground_truth_maps(frames_b[b], W, H)), "heat-map frame %d" % b
    # mode 1: the range transform of pool_lidar_tensor at full resolution (network input), negatives -> 0
    pre_t = helper.BatchPreprocessor(4, H, W, max_points=40000, max_boxes=800, transform=True)
    pre_t.load(frames_p, frames_b)
    pre_t.run(img, None)
    want = np.stack([np.maximum(orc.lidar_value_transform(orc.lidar_array_to_image(p, (1, H, W), 5)), 0.0) for p in frames_p])
    assert np.array_equal(img.cpu().numpy(), want.astype(np.float32))
    with pytest.raises(ValueError):
        pre.load([np.zeros((50000, 3), np.float32)] * 4, frames_b)


def test_batched_forms_small_odd_shapes():
    for seed, (H, W) in enumerate([(40, 50), (65, 257), (130, 300)]):
        B = 3
        fp = [synthetic.lidar_points(400 + 100 * i, H, W, seed=seed * 10 + i, out_of_range=0.1) for i in range(B)]
        fb = [synthetic.boxes(10 + 5 * i, H, W, seed=seed * 10 + i) for i in range(B)]
        pre = helper.BatchPreprocessor(B, H, W, max_points=1000, max_boxes=64, transform=False)
        pre.load(fp, fb)
        img = torch.empty((B, 1, H, W), device="cuda")
        maps = torch.empty((B, 3, H, W), device="cuda")
        pre.run(img, maps)
        for b in range(B):
            assert np.array_equal(img[b].cpu().numpy(), orc.lidar_array_to_image(fp[b], (1, H, W), 5))
            assert np.array_equal(maps[b].cpu().numpy(), orc.create_ground_truth_maps(fb[b], W, H))
