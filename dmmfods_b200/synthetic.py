"""Deterministic synthetic Waymo-shaped inputs: seeded numpy generators so that the
CUDA path, the oracle and the golden-vector script all see the same bytes.  Seed convention:
config.agent.seed (123, helper:179) + rank."""
import numpy as np

SEED = 123


def rgb_image(B, H, W, seed=SEED, channels=3):
    """un-normalised camera images, U[0,255) float32 (the dataset stores 0..255 floats, helper:604-607)."""
    rng = np.random.default_rng(seed)
    return (rng.random((B, channels, H, W), dtype=np.float32) * np.float32(255.0)).astype(np.float32)


def lidar_image(B, H, W, seed=SEED + 1000):
    """pooled LiDAR channel: 20 % of the pixels U[0,255), rest 0 (range of pool_lidar_tensor, helper:472-488)."""
    rng = np.random.default_rng(seed)
    v = rng.random((B, 1, H, W), dtype=np.float32) * np.float32(255.0)
    keep = rng.random((B, 1, H, W), dtype=np.float32) < np.float32(0.2)
    return np.where(keep, v, np.float32(0.0)).astype(np.float32)


def lidar_points(n=30000, H=1280, W=1920, seed=SEED + 2000, out_of_range=0.01):
    """(n,3) float32 rows [x, y, d]: integer-valued pixel coordinates, d = clip(Exp(20 m), 0.5, 80) so that
    some ranges exceed the 75 m clamp; `out_of_range` of the points lie outside the image; duplicates allowed."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, W, n).astype(np.float32)
    y = rng.integers(0, H, n).astype(np.float32)
    d = np.clip(rng.exponential(20.0, n), 0.5, 80.0).astype(np.float32)
    m = rng.random(n) < out_of_range
    k = int(m.sum())
    if k:
        x[m] = rng.integers(-40, W + 40, k).astype(np.float32)
        y[m] = rng.integers(-40, H + 40, k).astype(np.float32)
    return np.stack([x, y, d], 1).astype(np.float32)


def boxes(n=None, H=1280, W=1920, seed=SEED + 3000):
    """label dict like the dataset's (helper:625-640): {i: {type,x,y,width,height}} with types drawn from
    {1: .65, 2: .25, 4: .05, other: .05}, log-uniform sizes, a few degenerate (w<4 / h<5) boxes, all in-bounds."""
    rng = np.random.default_rng(seed)
    if n is None:
        n = int(rng.integers(5, 61))
    out = {}
    for i in range(n):
        t = int(rng.choice([1, 2, 4, 0, 3], p=[0.65, 0.25, 0.05, 0.025, 0.025]))
        wmax, hmax = min(400, W), min(300, H)
        w = int(np.exp(rng.uniform(np.log(4), np.log(wmax))))
        h = int(np.exp(rng.uniform(np.log(5), np.log(hmax))))
        if rng.random() < 0.1:
            w = int(rng.integers(1, 4))
        if rng.random() < 0.1:
            h = int(rng.integers(1, 5))
        w, h = min(w, W), min(h, H)
        x = int(rng.integers(0, W - w + 1))
        y = int(rng.integers(0, H - h + 1))
        out[str(i)] = {"type": t, "x": x, "y": y, "width": w, "height": h}
    return out


def target_maps(B, H, W, seed=SEED + 4000):
    """(B,3,H,W) float32 heat-map targets in {0,.3,.5,.75,1}: the reference's mask semantics painted at the
    training resolution (vectorised numpy restatement used only to SYNTHESISE targets, not as a checker)."""
    out = np.zeros((B, 3, H, W), dtype=np.float32)
    for b in range(B):
        for e in boxes(None, H, W, seed + b).values():
            c = e["type"]
            if c not in (1, 2, 4):
                continue
            idx = {1: 0, 2: 1, 4: 2}[c]
            w, h, x, y = e["width"], e["height"], e["x"], e["y"]
            box = np.ones((h, w), dtype=np.float32)
            if c == 2:
                hf, wf = h // 5, w // 4
                box[0:hf, :wf] = 0.3
                box[0:hf, wf * 3:] = 0.3
                box[hf * 3:, :wf] = 0.5
                box[hf * 3:, wf * 3:] = 0.5
                box[hf * 3:, wf:wf * 3] = 0.75
            out[b, idx, y:y + h, x:x + w] = box
    return out
