"""time the dense-block BN backward kernels (contrib, gather) at the training shapes; env toggles select kernel variants.
usage: python scripts/bench_bn.py [reps]"""
import ctypes as C
import sys

import torch

from dmmfods_b200 import _lib, ops

lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn):
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def contrib_case(name, B, H, W, Cbuf, Ci, gw=32):
    P = B * H * W
    x = ops.new_mat(B, H, W, Cbuf)
    x.t.normal_()
    g = ops.new_mat(B, H, W, Ci)
    g.t.normal_()
    gamma, beta = torch.rand(Ci, device="cuda") + 0.5, torch.randn(Ci, device="cuda")
    mean, invstd = torch.zeros(Ci, device="cuda"), torch.ones(Ci, device="cuda")
    sums = ops.Stats(torch.zeros(ops.Stats.size(Ci), dtype=torch.float64, device="cuda"), 0, Ci)
    dg, db = torch.zeros(Ci, device="cuda"), torch.zeros(Ci, device="cuda")
    bnb = ops.make_bn_bwd(sums, 0, P, gamma, beta, mean, invstd, dg, db)
    slab = torch.zeros(P * Ci, dtype=torch.bfloat16, device="cuda")
    for planar in (1, 0):
        d = ops.make_bn_bwd_args(x, 0, Ci, g.ptr(), g.ld, bnb, slab.data_ptr(), Ci, 0)
        if planar:
            d.out_gw, d.out_plane = gw, P * gw
        ms = timeit(lambda: _lib.check(lib.dmm_bn_relu_bwd_contrib(C.byref(d), None), "contrib"))
        print("contrib %-14s planar=%d  %7.3f ms  %6.0f GB/s" % (name, planar, ms, P * Ci * 6 / ms / 1e6), flush=True)
    # gather of the last k channels over nsrc consumers (planar)
    return x, slab


def gather_case(name, B, H, W, Cbuf, nsrc, Cg=32, gw=32):
    P = B * H * W
    x = ops.new_mat(B, H, W, Cbuf)
    x.t.normal_()
    slabs = [torch.zeros(P * gw, dtype=torch.bfloat16, device="cuda").normal_() for _ in range(nsrc)]
    kv = torch.zeros(2 * Cbuf, device="cuda")
    mean = torch.zeros(Cbuf, device="cuda")
    out = ops.new_mat(B, H, W, Cg)
    g = _lib.GradGather()
    for i, s in enumerate(slabs):
        g.src[i], g.ld[i], g.plane[i] = s.data_ptr(), gw, P * gw
        g.k1[i], g.k2[i] = kv.data_ptr(), kv.data_ptr() + 4 * Cbuf
    g.gw, g.nsrc, g.nk = gw, nsrc, nsrc
    g.mean, g.x, g.ldx = mean.data_ptr(), x.ptr(0).value, x.ld
    g.rows, g.C, g.out, g.ldo = P, Cg, out.ptr().value, out.ld
    ms = timeit(lambda: _lib.check(lib.dmm_grad_gather(C.byref(g), None), "gather"))
    print("gather  %-14s nsrc=%2d   %7.3f ms  %6.0f GB/s" % (name, nsrc, ms, P * Cg * 2 * (nsrc + 2) / ms / 1e6), flush=True)


if __name__ == "__main__":
    contrib_case("b1_l6", 32, 160, 240, 256, 224)
    contrib_case("b1_l2", 32, 160, 240, 256, 96)
    contrib_case("b2_l12", 32, 80, 120, 512, 480)
    contrib_case("b2_l4", 32, 80, 120, 512, 224)
    contrib_case("b3_l24", 32, 40, 60, 1280, 1248)
    contrib_case("b3_l8", 32, 40, 60, 1280, 736)
    gather_case("b1_l1", 32, 160, 240, 256, 5)
    gather_case("b2_l1", 32, 80, 120, 512, 11)
    gather_case("b3_l1", 32, 40, 60, 1280, 23)
    gather_case("b3_l12", 32, 40, 60, 1280, 12)
