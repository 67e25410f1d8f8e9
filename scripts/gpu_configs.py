"""robustness run of the other BASELINE configs (not bench lines): config 4 geometry (1280x1920, B=8) training step through the
Trainer, config 2 (early fusion, B=16, 640x960), config 5 (DenseNet-201 mid-fusion, eval forward, batch sweep)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmmfods_b200 import config as cfgmod, synthetic
from dmmfods_b200.model import densenet121_u_lidar, densenet201_u_lidar
from dmmfods_b200.trainer import Trainer

def cfg(c2, cb):
    c = cfgmod.get_config("/nonexistent")
    c.model.stream_2_in_channels, c.model.concat_before_block_num = c2, cb
    return c

def train_case(name, c2, cb, B, H, W, steps=3):
    torch.manual_seed(123)
    m = densenet121_u_lidar(pretrained=False, config=cfg(c2, cb)).cuda()
    tr = Trainer(m, B, H, W, use_graph=True)
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=1)).cuda()
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=2)).cuda()
    tg = torch.from_numpy(synthetic.target_maps(B, H, W, seed=3)).cuda()
    for _ in range(3):
        cs = tr.step(x1, x2, tg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        cs = tr.step(x1, x2, tg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert torch.isfinite(cs).all()
    print("%-40s B=%d %dx%d  %.1f ms/step  %.1f img/s  loss sums %s" % (name, B, H, W, ms, B / ms * 1e3, [round(float(v), 1) for v in cs]), flush=True)
    del tr, m
    torch.cuda.empty_cache()

def eval_case(B, H=640, W=960):
    torch.manual_seed(123)
    m = densenet201_u_lidar(pretrained=False, config=cfg(1, 3)).cuda().eval()
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=1)).cuda()
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=2)).cuda()
    with torch.no_grad():
        for _ in range(2):
            y = m(x1, x2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = m(x1, x2)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    assert torch.isfinite(y).all()
    print("config 5 DenseNet-201 mid eval forward     B=%d %dx%d  %.1f ms  %.1f img/s" % (B, H, W, ms, B / ms * 1e3), flush=True)
    del m
    torch.cuda.empty_cache()

if __name__ == "__main__":
    train_case("config 4 geometry: mid-fusion 1280x1920", 1, 3, 8, 1280, 1920)
    train_case("config 2: early fusion", 1, 1, 16, 640, 960)
    train_case("config 1 geometry: no fusion", 0, 1, 2, 256, 384)
    for B in (1, 8, 32):
        eval_case(B)
