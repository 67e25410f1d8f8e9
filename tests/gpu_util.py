"""Helpers shared by the -m gpu parity tests: NCHW float <-> pixel-major bf16 Mat, error metrics."""
import torch

from dmmfods_b200 import ops


def bf16_round(x):
    return x.to(torch.bfloat16).to(x.dtype)


def to_mat(x, ld=None, c0=0):
    """x: (B,C,H,W) float (CPU) -> Mat on cuda with channels placed at [c0, c0+C), rest zero."""
    B, C, H, W = x.shape
    ld = ld or C
    t = torch.zeros(B * H * W, ld, dtype=torch.bfloat16)
    t[:, c0:c0 + C] = x.permute(0, 2, 3, 1).reshape(-1, C).to(torch.bfloat16)
    return ops.Mat(t.cuda(), B, H, W)


def from_mat(m, c0=0, C=None):
    """Mat -> (B,C,H,W) float64 CPU."""
    C = m.ld - c0 if C is None else C
    t = m.t[:, c0:c0 + C].float().cpu().double()
    return t.reshape(m.B, m.H, m.W, C).permute(0, 3, 1, 2).contiguous()


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    d = (a - b).norm()
    n = b.norm()
    return (d / n).item() if n > 0 else d.item()


def new_stats(ld):
    buf = torch.zeros(ops.Stats.size(ld), dtype=torch.float64, device="cuda")
    return ops.Stats(buf, 0, ld)
