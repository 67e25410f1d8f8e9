"""micro-benchmark of dmm_conv_igemm on the layer shapes of BASELINE config 3 (CUDA events, L2 flushed by size)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dmmfods_b200 import ops

CASES = {   # name: (B, H, W, Cin, Cout, K, out_mode)
    "refine0": (32, 640, 960, 132, 64, 3, 0),
    "refine1": (32, 640, 960, 64, 3, 5, 1),
    "refine0_dgrad": (32, 640, 960, 64, 132, 3, 0),
    "b1_conv2": (32, 160, 240, 128, 32, 3, 0),
    "b1_conv2_dgrad": (32, 160, 240, 32, 128, 3, 0),
    "b1_conv1_k160": (32, 160, 240, 160, 128, 1, 0),
    "b2_conv1_k320": (32, 80, 120, 320, 128, 1, 0),
    "b2_conv2": (32, 80, 120, 128, 32, 3, 0),
    "b3_conv2": (32, 40, 60, 128, 32, 3, 0),
    "b3_conv1_k640": (32, 40, 60, 640, 128, 1, 0),
    "reduce4": (32, 160, 240, 512, 128, 1, 0),
    "convT4_phase11": (32, 160, 240, 128, 128, 2, 0),
    "refine1_dgrad": (32, 640, 960, 3, 64, 5, 0),
    "b1_conv2_fold": (32, 160, 240, 128, 96, -3, 4),           # out_mode 4 here = C-ABI out_mode 3 (3 kernel columns folded, bf16 + stats)
    "refine0_fold": (32, 640, 960, 136, 192, -3, 4),
    "refine1_fold": (32, 640, 960, 64, 15, -5, 3),             # out_mode 3 here = folded kernel columns (out_mode 2 of the C-ABI)
    "refine1_dgrad_v": (32, 640, 960, 16, 64, -5, 0),          # K < 0: |K| vertical taps over the horizontally unfolded d(logits)
    "b1_conv1_k160_pro": (32, 160, 240, 160, 128, 1, 2),      # out_mode 2 here = BN-ReLU prologue on a [P, 256] block buffer
    "b2_conv1_k320_pro": (32, 80, 120, 320, 128, 1, 2),
    "b1_conv1_k64_pro": (32, 160, 240, 64, 128, 1, 2),
    "b3_conv1_k640_pro": (32, 40, 60, 640, 128, 1, 2),
    "b3_conv1_k992_pro": (32, 40, 60, 992, 128, 1, 2),
    "b4_conv1_k768_pro": (32, 20, 30, 768, 128, 1, 2),
    "b1_conv1_dgrad_n160": (32, 160, 240, 128, 160, 1, 0),      # conv1 data gradient: K = 128 -> N = Ci
    "b2_conv1_dgrad_n320": (32, 80, 120, 128, 320, 1, 0),
    "b3_conv1_dgrad_n640": (32, 40, 60, 128, 640, 1, 0),
    "b3_conv1_dgrad_n992": (32, 40, 60, 128, 992, 1, 0),
    "b4_conv1_dgrad_n768": (32, 20, 30, 128, 768, 1, 0),
    # growth convolution data gradient as the engine launches it: two 32-channel taps per 64-wide K block (kwidth 32) and the
    # reduce pass of norm2's backward fused into the epilogue (out_mode 5 here)
    "b1_conv2_dgrad_bnb": (32, 160, 240, 32, 128, 3, 5),
    "b2_conv2_dgrad_bnb": (32, 80, 120, 32, 128, 3, 5),
    "b3_conv2_dgrad_bnb": (32, 40, 60, 32, 128, 3, 5),
    "b1_conv2_dgrad_k32": (32, 160, 240, 32, 128, 3, 6),        # same without the fused reduce
    # deep decoder stages: few pixels, K and N in the thousands (weights streamed from L2 per tile)
    "convT1_phase11": (32, 20, 30, 1024, 1024, 2, 0),
    "convT2_phase11": (32, 40, 60, 512, 512, 2, 0),
    "convT3_phase11": (32, 80, 120, 256, 256, 2, 0),
    "reduce1": (32, 20, 30, 1024, 1024, 1, 0),
    "reduce2": (32, 40, 60, 2048, 512, 1, 0),
    "reduce3": (32, 80, 120, 1024, 256, 1, 0),
}

def run(name, reps=5):
    B, H, W, Cin, Cout, K, om = CASES[name]
    torch.manual_seed(0)
    pro = om == 2
    bnb, k32 = om == 5, om in (5, 6)
    if pro or k32:
        om = 0
    ld = ops.ceil_to(Cin, 8) if not pro else (256 if Cin <= 256 else (512 if Cin <= 512 else 1024))
    a = ops.Mat((torch.randn(B * H * W, ld, device="cuda") * 0.5).to(torch.bfloat16), B, H, W)
    if K < 0:
        taps = [(0, -K // 2 - kh, 0) for kh in range(-K)]
    elif K == 2:
        taps = [(0, dy, dx) for dy in (0, 1) for dx in (0, 1)]
    else:
        taps = ops.conv_taps(K, (K - 1) // 2)[0]
    T = len(taps)
    kwidth = 16 if (Cin <= 16 and T > 1) else (32 if k32 else 64)
    Kp = ops.ceil_to(Cin, kwidth)
    n_tile = ops.pick_n_tile(Cout)
    n_rows = ops.ceil_to(Cout, n_tile)
    wp = (torch.randn(n_rows, T * Kp, device="cuda") * 0.05).to(torch.bfloat16)
    if om == 0:
        out = ops.new_mat(B, H, W, ops.ceil_to(Cout, 8))
        st = torch.zeros(ops.Stats.size(out.ld), dtype=torch.float64, device="cuda")
        # data gradients of the network carry no statistics: IG_NOSTATS=1 (or a "dgrad" case name) drops them
        nostats = os.environ.get("IG_NOSTATS", "0") != "0" or "dgrad" in name
        if bnb:
            xr = ops.Mat((torch.randn(B * H * W, out.ld, device="cuda") * 0.5).to(torch.bfloat16), B, H, W)
            sums = torch.zeros(ops.Stats.size(out.ld), dtype=torch.float64, device="cuda")
            g_, b2_, sm_, si_ = (torch.rand(out.ld, device="cuda") for _ in range(4))
            bb = ops.make_bn_bwd(ops.Stats(sums, 0, out.ld), 0, B * H * W, g_, b2_, sm_, si_)
            keep2 = (xr, sums, g_, b2_, sm_, si_, bb)
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, T * Kp, n_rows, W, H, B, Cout, out.ptr(), out.ld,
                           stats=None if nostats else ops.Stats(st, 0, out.ld), n_tile=n_tile, kwidth=kwidth)
        if bnb:
            ops.fuse_bn_bwd_reduce(d, xr, 0, bb)
        if pro:
            bst = torch.rand(ops.Stats.size(ld), dtype=torch.float64, device="cuda") * 1000 + 5000
            g, b_ = torch.ones(ld, device="cuda"), torch.zeros(ld, device="cuda")
            rm, rv, sm, si = (torch.zeros(ld, device="cuda") for _ in range(4))
            d.pro_enable = 1
            d.pro_bn = ops.make_bn(ops.Stats(bst, 0, ld), 0, B * H * W, g, b_, rm, rv, sm, si, training=True)
            keep = (bst, g, b_, rm, rv, sm, si)
    elif om == 4:
        out = ops.new_mat(B, H, W, Cout // 3)
        st = torch.zeros(ops.Stats.size(out.ld), dtype=torch.float64, device="cuda")
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, T * Kp, n_rows, W, H, B, Cout, out.ptr(), out.ld, out_mode=3,
                           stats=ops.Stats(st, 0, out.ld), n_tile=Cout, fold_kw=3, tile_w=32)
    elif om == 3:
        out = torch.zeros(B, Cout // 5, H, W, device="cuda")
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, T * Kp, n_rows, W, H, B, Cout, out.data_ptr(), 0, out_mode=2, n_tile=n_tile,
                           fold_kw=5, tile_w=32)
    else:
        out = torch.zeros(B, Cout, H, W, device="cuda")
        d = ops.make_igemm([a.view(0, Cin)], taps, wp, T * Kp, n_rows, W, H, B, Cout, out.data_ptr(), 0, out_mode=1, n_tile=n_tile)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.run_igemm(d); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * B * H * W * Cin * Cout * T
    by = B * H * W * (Cin + Cout) * 2
    print("%-16s %8.3f ms %8.1f TF/s %8.0f GB/s" % (name, ms, fl / ms / 1e9, by / ms / 1e6), flush=True)

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run(n)
