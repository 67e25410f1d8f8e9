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
