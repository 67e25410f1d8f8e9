"""CPU oracle (numpy) for the two integer pre-processing scatters of the hot path.

TEST INFRASTRUCTURE - NOT PRODUCT CODE (see oracle/dense_unet_oracle.py header).

Parity status: the reference has no tests for these; the oracle is pinned against the
unmodified reference functions executed in the authoring container
(tests/golden/make_golden.py -> tests/golden/lidar_heatmap.npz, tests/test_oracle_golden.py).

Restates, in vectorised numpy with exactly the reference's sequential semantics:
  lidar_array_to_image_like_tensor   Dense_U_Net_lidar_helper.py:493-515
  pool_lidar_tensor                  Dense_U_Net_lidar_helper.py:446-491
  create_ground_truth_maps (+ _bb*)  Dense_U_Net_lidar_helper.py:233-305
  maxpool_tensor                     Dense_U_Net_lidar_helper.py:438-444
"""
import numpy as np


def _py_slice(lo, hi, n):
    """Bounds of python slice lo:hi on an axis of length n (negative indices wrap, then clip)."""
    if lo < 0:
        lo = max(lo + n, 0)
    if hi < 0:
        hi = max(hi + n, 0)
    lo = min(lo, n)
    hi = min(hi, n)
    return lo, max(hi, lo)


def splat_rect(x, y, H, W, kernel_size=5):
    """Index rectangle painted by one LiDAR point (helper:499-513): int() truncates toward
    zero, lower bound clamped to 0, upper bound clamped to size-1 and EXCLUSIVE (so the last
    row / column is never written), then python slice semantics (a negative upper bound wraps)."""
    shift = (kernel_size - 1) // 2
    min_y = int(y - shift)
    if min_y < 0:
        min_y = 0
    max_y = int(y + shift + 1)
    if max_y > H - 1:
        max_y = H - 1
    min_x = int(x - shift)
    if min_x < 0:
        min_x = 0
    max_x = int(x + shift + 1)
    if max_x > W - 1:
        max_x = W - 1
    y0, y1 = _py_slice(min_y, max_y, H)
    x0, x1 = _py_slice(min_x, max_x, W)
    return y0, y1, x0, x1


def lidar_array_to_image(lidar_array, shape=(1, 1280, 1920), kernel_size=5):
    """helper:493-515.  lidar_array: (N,3) float32 rows [x, y, d]; later points overwrite earlier."""
    lidar_array = np.asarray(lidar_array, dtype=np.float32).reshape(-1, 3)
    _, H, W = shape
    img = np.full(shape, -1.0, dtype=np.float32)
    for x, y, d in lidar_array:
        y0, y1, x0, x1 = splat_rect(np.float32(x), np.float32(y), H, W, kernel_size)
        img[0, y0:y1, x0:x1] = d
    return img


def lidar_value_transform(v):
    """Element-wise part of pool_lidar_tensor (helper:472-481), float32, sequential masks."""
    v = np.array(v, dtype=np.float32, copy=True)
    v[v > np.float32(75.0)] = np.float32(75.0)
    v[v == np.float32(-1.0)] = np.float32(76.0)
    m = v <= np.float32(25)
    v[m] = (v[m] * np.float32(-6.2)) + np.float32(255)          # two separately rounded fp32 ops
    m = (v > np.float32(25)) & (v <= np.float32(76.0))
    v[m] = (v[m] * np.float32(-2)) + np.float32(150)
    return v


def pool_lidar(lidar_img):
    """helper:446-491: value transform, MaxPool2d((20,10), stride (10,10)), replicate-pad one
    row at the bottom, negatives -> 0.  (1,H,W) -> (1, (H-20)//10+2, (W-10)//10+1)."""
    v = lidar_value_transform(lidar_img)
    _, H, W = v.shape
    oh = (H - 20) // 10 + 1
    ow = (W - 10) // 10 + 1
    out = np.empty((1, oh + 1, ow), dtype=np.float32)
    for i in range(oh):
        rows = v[0, i * 10:i * 10 + 20, :ow * 10]
        out[0, i] = rows.reshape(20, ow, 10).max(axis=(0, 2))
    out[0, oh] = out[0, oh - 1]
    out[out < 0] = 0
    return out


# float32 template constants (reference builds float64 maps then torch.Tensor() -> float32)
UNLIKELY = np.float32(0.3)
UNCERTAIN = np.float32(0.5)
HALF_CERTAIN = np.float32(0.75)


def box_template(object_class, width, height):
    """helper:233-274 (_create_ground_truth_bb*).  Rules applied sequentially on a ones box."""
    box = np.ones((height, width), dtype=np.float64)
    if object_class == 2:
        hf = height // 5
        wf = width // 4
        box[0:hf, :wf] = 0.3
        box[0:hf, wf * 3:] = 0.3
        box[hf * 3:, :wf] = 0.5
        box[hf * 3:, wf * 3:] = 0.5
        box[hf * 3:, wf:wf * 3] = 0.75
    elif object_class in (1, 4):
        pass
    else:
        raise TypeError("the ground truth label class does not exist")
    return box


def create_ground_truth_maps(ground_truth, width_img=1920, height_img=1280):
    """helper:276-305.  ground_truth: dict of dicts {type,x,y,width,height}; dict order =
    paint order; later boxes overwrite earlier ones inside the same class channel.
    Out-of-bounds boxes raise ValueError (numpy broadcast failure in the reference)."""
    maps = np.zeros((3, height_img, width_img), dtype=np.float64)
    for elem in ground_truth.values():
        c = elem["type"]
        if c == 1 or c == 2 or c == 4:
            w, h, x, y = elem["width"], elem["height"], elem["x"], elem["y"]
            idx = (c == 1) * 0 + (c == 2) * 1 + (c == 4) * 2
            maps[idx, y:y + h, x:x + w] = box_template(c, w, h)
    return maps.astype(np.float32)


def maxpool(img, k=10):
    """helper:438-444: MaxPool2d(k, stride=k) on (C,H,W)."""
    C, H, W = img.shape
    oh, ow = H // k, W // k
    return img[:, :oh * k, :ow * k].reshape(C, oh, k, ow, k).max(axis=(2, 4))


def avgpool(img, k=10):
    """helper:430-436: AvgPool2d(k, stride=k) on (C,H,W) (used on the camera image)."""
    C, H, W = img.shape
    oh, ow = H // k, W // k
    return img[:, :oh * k, :ow * k].reshape(C, oh, k, ow, k).mean(axis=(2, 4), dtype=np.float32)


# ---- step metrics (helper:311-401) --------------------------------------------------------------------------------
def iou_whole_img_batch(ground_truth_map_batch, estimated_heat_map_batch, threshold=0.7):
    """compute_IoU_whole_img_batch (helper:345-367) over compute_IoU_whole_img_per_class (:311-343): both maps are
    thresholded with >=, IoU = |and| / |or| per (sample, class) as float32, nan where the union is empty."""
    gt = np.asarray(ground_truth_map_batch) >= threshold
    est = np.asarray(estimated_heat_map_batch) >= threshold
    inter = (gt & est).sum(axis=(2, 3)).astype(np.float32)
    union = (gt | est).sum(axis=(2, 3)).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / union


def accuracy(ground_truth, prediction, threshold=0.7):
    """compute_accuracy (helper:369-401): class-wise fraction of positions whose thresholded values agree; (C,H,W) or
    (B,C,H,W) inputs."""
    gt = np.asarray(ground_truth)
    pr = np.asarray(prediction)
    if gt.ndim == 3:
        axes, ncls = (1, 2), gt.shape[0]
    elif gt.ndim == 4:
        axes, ncls = (0, 2, 3), gt.shape[1]
    else:
        raise ValueError('Number of dimensions must be either 3 or 4, you gave ' + str(gt.ndim))
    return ((pr >= threshold) == (gt >= threshold)).sum(axis=axes) / (gt.size / ncls)
