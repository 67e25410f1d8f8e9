"""micro-benchmark of dmm_conv_wgrad on layer shapes of BASELINE config 3."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmmfods_b200 import ops

CASES = {   # name: (B, H, W, Cin(M), Cout(N), K)
    "b1_conv2": (32, 160, 240, 128, 32, 3),
    "b2_conv2": (32, 80, 120, 128, 32, 3),
    "b3_conv2": (32, 40, 60, 128, 32, 3),
    "b1_conv1_k160": (32, 160, 240, 160, 128, 1),
    "b2_conv1_k320": (32, 80, 120, 320, 128, 1),
    "b3_conv1_k640": (32, 40, 60, 640, 128, 1),
    "b4_conv1_k768": (32, 20, 30, 768, 128, 1),
    "refine0": (32, 640, 960, 132, 64, 3),
    "reduce4": (32, 160, 240, 512, 128, 1),
    "reduce1": (32, 20, 30, 1024, 1024, 1),
    "b3_conv1_k1024": (32, 40, 60, 1024, 128, 1),
    "b4_conv1_k1024": (32, 20, 30, 1024, 128, 1),
    "b2_conv1_k512": (32, 80, 120, 512, 128, 1),
}

def run(name, reps=5):
    B, H, W, M, N, K = CASES[name]
    torch.manual_seed(0)
    x = ops.Mat((torch.randn(B * H * W, ops.ceil_to(M, 8), device="cuda") * 0.5).to(torch.bfloat16), B, H, W)
    g = ops.Mat((torch.randn(B * H * W, ops.ceil_to(N, 8), device="cuda") * 0.5).to(torch.bfloat16), B, H, W)
    taps = ops.conv_taps(K, (K - 1) // 2)[0]
    plan = ops.plan_conv_wgrad(x.view(0, M), [g.view(0, N)], taps, M, N)
    dw = torch.zeros(plan["rows"] * plan["ld"], dtype=torch.float32, device="cuda")
    ds = [ops.make_wgrad(W=W, H=H, B=B, dw=dw, ld=plan["ld"], **kw) for kw in plan["launches"]]
    def go():
        for d in ds:
            ops.run_wgrad(d)
    go(); torch.cuda.synchronize()
    ts = []
    inner = int(os.environ.get("WG_INNER", "20"))       # back-to-back launches per timing: keeps the clocks up, hides launch latency
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _i in range(inner):
            go()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * B * H * W * M * N * K * K
    by = B * H * W * (M + N) * 2
    d0 = ds[0]
    print("%-16s %8.3f ms %8.1f TF/s %8.0f GB/s  (launches %d, n_tile %d, num_a %d num_b %d kpx %d)" %
          (name, ms, fl / ms / 1e9, by / ms / 1e6, len(ds), d0.n_tile, d0.num_a, d0.num_b, d0.kpx), flush=True)

if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run(n)
