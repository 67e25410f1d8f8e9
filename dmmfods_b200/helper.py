"""Host-side mirror of the pre-processing helpers of dmmfods/utils/Dense_U_Net_lidar_helper.py that sit on
the hot path (same names, argument meaning and error behaviour), backed by the CUDA integer-scatter
kernels in csrc/scatter.cu.  Results are bit-identical to the reference's CPU functions.

    lidar_array_to_image_like_tensor   helper:493-515
    pool_lidar_tensor                  helper:446-491
    create_ground_truth_maps           helper:276-305 (+ templates :233-274)
    maxpool_tensor / avgpool_tensor    helper:430-444
    compute_IoU_whole_img_per_class / compute_IoU_whole_img_batch / compute_accuracy   helper:311-401 (fused counters)

Inputs may be numpy arrays / python dicts (as in the reference) or CUDA tensors; outputs are CUDA
float32 tensors.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .config import (EasyDict, create_config, get_config, load_config, load_json_file, save_config,  # noqa: F401
                     save_json_file, set_current_run)
from .ops import _ptr, _stream, require_device


def _dev():
    require_device()
    return torch.device("cuda", torch.cuda.current_device())


def lidar_array_to_image_like_tensor(lidar_array, shape=(1, 1280, 1920), kernel_size=5):
    """(N,3) rows [x, y, d] -> (1,H,W) float32 image, -1 where no point, kernel_size^2 splat, the LAST
    point in array order wins (helper:493-515, incl. its clamping and slice semantics)."""
    dev = _dev()
    if isinstance(lidar_array, torch.Tensor):
        pts = lidar_array.to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
    else:
        pts = torch.from_numpy(np.ascontiguousarray(np.asarray(lidar_array, dtype=np.float32).reshape(-1, 3))).to(dev)
    _, H, W = shape
    img = torch.empty(shape, dtype=torch.float32, device=dev)
    scratch = torch.empty(H * W, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().dmm_lidar_splat(_ptr(pts) if pts.numel() else None, pts.shape[0], H, W, kernel_size,
                                           _ptr(scratch), _ptr(img), _stream()), "dmm_lidar_splat")
    return img


def pool_lidar_tensor(lidar_tensor):
    """(1,H,W) range image -> inverted [0,255] intensities, MaxPool2d((20,10), stride 10), one replicated
    bottom row, negatives -> 0 (helper:446-491).  Like the reference this leaves the input transformed
    semantics aside: the input tensor is NOT modified here."""
    dev = _dev()
    x = lidar_tensor.to(device=dev, dtype=torch.float32).contiguous()
    assert x.dim() == 3 and x.shape[0] == 1
    _, H, W = x.shape
    out = torch.empty((1, (H - 20) // 10 + 2, (W - 10) // 10 + 1), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().dmm_lidar_pool(_ptr(x), H, W, _ptr(out), _stream()), "dmm_lidar_pool")
    return out


def boxes_from_ground_truth(ground_truth, width_img, height_img):
    """dict of {type,x,y,width,height} -> int32 (N,5) rows [type,x,y,w,h] in dict (= paint) order.
    Raises like the reference does: TypeError never (unknown classes are skipped, helper:295),
    ValueError for boxes that do not lie inside the image (numpy broadcast error in the reference)."""
    rows = []
    for elem in ground_truth.values():
        c = elem["type"]
        if c == 1 or c == 2 or c == 4:
            w, h, x, y = int(elem["width"]), int(elem["height"]), int(elem["x"]), int(elem["y"])
            if x < 0 or y < 0 or w < 0 or h < 0 or x + w > width_img or y + h > height_img:
                raise ValueError("could not broadcast input array from shape (%d,%d) into the image" % (h, w))
            rows.append((c, x, y, w, h))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 5)


def create_ground_truth_maps(ground_truth, width_img=1920, height_img=1280):
    """label dict (or int32 (N,5) box tensor) -> (3,H,W) float32 class heat maps; vehicles / cyclists are
    filled rectangles, pedestrians the 0.3/0.5/0.75/1 silhouette; later boxes overwrite earlier ones
    of the same class (helper:233-305)."""
    dev = _dev()
    if isinstance(ground_truth, torch.Tensor):
        boxes = ground_truth.to(device=dev, dtype=torch.int32).reshape(-1, 5).contiguous()
    else:
        boxes = torch.from_numpy(boxes_from_ground_truth(ground_truth, width_img, height_img)).to(dev)
    maps = torch.empty((3, height_img, width_img), dtype=torch.float32, device=dev)
    scratch = torch.empty(3 * height_img * width_img, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().dmm_heatmap_boxes(_ptr(boxes) if boxes.numel() else None, boxes.shape[0], height_img,
                                             width_img, _ptr(scratch), _ptr(maps), _stream()), "dmm_heatmap_boxes")
    return maps


class BatchPreprocessor:
    """Per-step on-GPU pre-processing of a whole batch (BASELINE config 4): the LiDAR point lists of B frames -> (B,1,H,W)
    network input (helper:493-515 splat + the range transform of helper:472-481 at full resolution) and the label boxes of
    B frames -> (B,3,H,W) heat-map targets (helper:276-305), ONE single-pass launch each (dmm_lidar_splat_batched,
    dmm_heatmap_boxes_batched).  Static device buffers of fixed capacity + device-resident frame offsets: the two launches can
    sit inside a captured CUDA graph while the number of points / boxes changes from step to step.

        pre = BatchPreprocessor(B, H, W, max_points=40000, max_boxes=256)
        pre.load(points_per_frame, labels_per_frame)     # host -> static device buffers (async copies)
        pre.run(lidar_out, target_out)                   # two kernel launches on the current stream
    """

    def __init__(self, B, H, W, max_points=40000, max_boxes=256, kernel_size=5, transform=True):
        dev = _dev()
        self.B, self.H, self.W, self.kernel_size, self.mode = B, H, W, int(kernel_size), 1 if transform else 0
        self.max_points, self.max_boxes = int(max_points), int(max_boxes)
        self.points = torch.zeros((B * self.max_points, 3), dtype=torch.float32, device=dev)
        self.boxes = torch.zeros((B * self.max_boxes, 5), dtype=torch.int32, device=dev)
        self.p_off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        self.b_off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        self._h_points = torch.zeros((B * self.max_points, 3), dtype=torch.float32).pin_memory()
        self._h_boxes = torch.zeros((B * self.max_boxes, 5), dtype=torch.int32).pin_memory()
        self._h_poff = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
        self._h_boff = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
        self.h2d_bytes = 0

    def load(self, points_per_frame, labels_per_frame):
        """points_per_frame: B arrays (N_b,3) float32; labels_per_frame: B label dicts (helper:625-640) or int32 (N_b,5) arrays."""
        assert len(points_per_frame) == self.B and len(labels_per_frame) == self.B
        np_, nb_ = 0, 0
        hp, hb = self._h_points.numpy(), self._h_boxes.numpy()
        for i in range(self.B):
            pts = np.asarray(points_per_frame[i], dtype=np.float32).reshape(-1, 3)
            lab = labels_per_frame[i]
            bx = boxes_from_ground_truth(lab, self.W, self.H) if isinstance(lab, dict) else np.asarray(lab, dtype=np.int32).reshape(-1, 5)
            if pts.shape[0] > self.max_points or bx.shape[0] > self.max_boxes:
                raise ValueError("frame %d: %d points / %d boxes exceed the static capacity (%d / %d)"
                                 % (i, pts.shape[0], bx.shape[0], self.max_points, self.max_boxes))
            self._h_poff[i], self._h_boff[i] = np_, nb_
            hp[np_:np_ + pts.shape[0]] = pts
            hb[nb_:nb_ + bx.shape[0]] = bx
            np_ += pts.shape[0]
            nb_ += bx.shape[0]
        self._h_poff[self.B], self._h_boff[self.B] = np_, nb_
        self.points[:np_].copy_(self._h_points[:np_], non_blocking=True)
        self.boxes[:nb_].copy_(self._h_boxes[:nb_], non_blocking=True)
        self.p_off.copy_(self._h_poff, non_blocking=True)
        self.b_off.copy_(self._h_boff, non_blocking=True)
        self.h2d_bytes = np_ * 12 + nb_ * 20 + 8 * (self.B + 1)
        return np_, nb_

    def run(self, lidar_out, target_out):
        """lidar_out: (B,1,H,W) float32 CUDA tensor, target_out: (B,3,H,W); either may be None."""
        lib = _lib.load()
        if lidar_out is not None:
            assert lidar_out.is_contiguous() and tuple(lidar_out.shape) == (self.B, 1, self.H, self.W)
            _lib.check(lib.dmm_lidar_splat_batched(_ptr(self.points), _ptr(self.p_off), self.B, self.H, self.W, self.kernel_size,
                                                   self.mode, _ptr(lidar_out), _stream()), "dmm_lidar_splat_batched")
        if target_out is not None:
            assert target_out.is_contiguous() and tuple(target_out.shape) == (self.B, 3, self.H, self.W)
            _lib.check(lib.dmm_heatmap_boxes_batched(_ptr(self.boxes), _ptr(self.b_off), self.B, self.H, self.W, _ptr(target_out),
                                                     _stream()), "dmm_heatmap_boxes_batched")


def _pool(img_tensor, k, is_max):
    dev = _dev()
    x = img_tensor.to(device=dev, dtype=torch.float32).contiguous()
    assert x.dim() == 3
    Cc, H, W = x.shape
    out = torch.empty((Cc, H // k, W // k), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().dmm_pool_kxk(_ptr(x), Cc, H, W, k, 1 if is_max else 0, _ptr(out), _stream()), "dmm_pool_kxk")
    return out


def maxpool_tensor(img_tensor):
    """torch.nn.MaxPool2d(10, stride=10) (helper:438-444)."""
    return _pool(img_tensor, 10, True)


def avgpool_tensor(img_tensor):
    """torch.nn.AvgPool2d(10, stride=10) (helper:430-436)."""
    return _pool(img_tensor, 10, False)


def _metric_counts(ground_truth, prediction, threshold):
    """int64 [planes, 3] = (intersection, union, equal) of the thresholded maps, one fused pass (dmm_step_metrics)."""
    dev = _dev()
    gt = ground_truth.to(device=dev, dtype=torch.float32).contiguous()
    pr = prediction.to(device=dev, dtype=torch.float32).contiguous()
    if gt.shape != pr.shape:
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s" % (tuple(pr.shape), tuple(gt.shape)))
    H, W = gt.shape[-2:]
    planes = gt.numel() // (H * W) if H * W else 0
    counts = torch.zeros((planes, 3), dtype=torch.int64, device=dev)
    _lib.check(_lib.load().dmm_step_metrics(_ptr(pr), _ptr(gt), planes, H * W, float(threshold), _ptr(counts), _stream()),
               "dmm_step_metrics")
    return counts


def compute_IoU_whole_img_per_class(ground_truth_map, estimated_heat_map, threshold):
    """(C,H,W) maps -> IoU per class, nan where the union is empty (helper:311-343)."""
    c = _metric_counts(ground_truth_map, estimated_heat_map, threshold)
    return c[:, 0].float() / c[:, 1].float()


def compute_IoU_whole_img_batch(ground_truth_map_batch, estimated_heat_map_batch, threshold=0.7):
    """(B,C,H,W) batches -> whole-image IoU per sample and class, nan for 0/0 (helper:345-367), without the reference's
    per-sample Python loop."""
    B, Cc = ground_truth_map_batch.shape[:2]
    c = _metric_counts(ground_truth_map_batch, estimated_heat_map_batch, threshold).view(B, Cc, 3)
    # the reference allocates its result with torch.zeros(...) on the HOST and the agent feeds it to np.nanmean (Agent.py:251-252)
    return (c[:, :, 0].float() / c[:, :, 1].float()).cpu()


def compute_accuracy(ground_truth, prediction, threshold=0.7):
    """class-wise (TP+TN)/all of the thresholded maps for one sample (C,H,W) or a batch (B,C,H,W) (helper:369-401)."""
    if ground_truth.dim() == 3:
        Cc = ground_truth.shape[0]
        c = _metric_counts(ground_truth, prediction, threshold)[:, 2]
    elif ground_truth.dim() == 4:
        B, Cc = ground_truth.shape[:2]
        c = _metric_counts(ground_truth, prediction, threshold).view(B, Cc, 3)[:, :, 2].sum(0)
    else:
        raise ValueError('Number of dimensions must be either 3 or 4, you gave ' + str(ground_truth.dim()))
    return c / (ground_truth.numel() / Cc)
