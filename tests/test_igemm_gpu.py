"""Parity of the tcgen05 implicit-GEMM convolution (dmm_conv_igemm) with the torch CPU ops the
reference executes (nn.Conv2d / nn.ConvTranspose2d and their data gradients) on identical bf16-rounded
inputs.  Tolerance: relL2 <= 1e-2 vs the fp64 evaluation (bf16 storage of the result; the per-kernel
bf16 bound of BASELINE.json is 2e-2)."""
import pytest
import torch
import torch.nn.functional as F

from dmmfods_b200 import ops
from gpu_util import bf16_round, from_mat, new_stats, rel_l2, to_mat

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _pack(w, n_valid, C, T, tap_off, sn, sc, kwidth=64):
    n_rows = ops.ceil_to(n_valid, 16)
    Kp = ops.ceil_to(C, kwidth)
    dst = torch.empty(n_rows, T * Kp, dtype=torch.bfloat16, device="cuda")
    ops.pack_weights(w.cuda().contiguous(), dst, n_valid, n_rows, C, T, tap_off, sn, sc, kwidth)
    return dst, T * Kp, n_rows


def _conv_case(B, Cin, Cout, H, W, K, seed, coff=0, extra=0, with_stats=True, kwidth=64, cin_ld=None):
    torch.manual_seed(seed)
    pad = (K - 1) // 2
    x = bf16_round(torch.randn(B, Cin, H, W))
    w = bf16_round(torch.randn(Cout, Cin, K, K) / (Cin * K * K) ** 0.5)
    ref = F.conv2d(x.double(), w.double(), padding=pad)
    a = to_mat(x, ld=cin_ld)
    fwd, _, off = ops.conv_taps(K, pad)
    wp, ktot, n_rows = _pack(w, Cout, Cin, K * K, off, Cin * K * K, K * K, kwidth)
    ldo = coff + Cout + extra
    out = ops.new_mat(B, H, W, ldo, zero=True)
    st = new_stats(ldo) if with_stats else None
    d = ops.make_igemm([a.view(0, Cin)], fwd, wp, ktot, n_rows, W, H, B, Cout, out.ptr(), ldo, coff=coff, stats=st,
                       stats_off=coff, kwidth=kwidth)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    got = from_mat(out, coff, Cout)
    err = rel_l2(got, ref)
    assert err < TOL, "conv %dx%d Cin=%d Cout=%d %dx%d: relL2 %.3e" % (K, K, Cin, Cout, H, W, err)
    if coff:
        assert float(from_mat(out, 0, coff).abs().max()) == 0.0, "wrote outside the channel slice"
    if extra:
        assert float(from_mat(out, coff + Cout, extra).abs().max()) == 0.0, "wrote outside the channel slice"
    if with_stats:
        s1, s2 = st.totals()
        g = got.permute(1, 0, 2, 3).reshape(Cout, -1)
        assert torch.allclose(s1[coff:coff + Cout].cpu(), g.sum(1), rtol=1e-4, atol=1e-3 * g.abs().sum(1).max().item())
        assert torch.allclose(s2[coff:coff + Cout].cpu(), (g * g).sum(1), rtol=1e-4, atol=1e-6)
    return err


@pytest.mark.parametrize("Cin,Cout,H,W", [(64, 128, 16, 24), (96, 128, 8, 12), (256, 128, 20, 30), (1024, 512, 4, 6),
                                          (160, 128, 5, 7), (2048, 512, 3, 5), (64, 32, 16, 16)])
def test_conv1x1(Cin, Cout, H, W):
    _conv_case(2, Cin, Cout, H, W, 1, seed=Cin + Cout)


def test_conv1x1_channel_offset_into_block_buffer():
    _conv_case(2, 128, 32, 12, 20, 1, seed=5, coff=96, extra=32)


@pytest.mark.parametrize("Cin,Cout,H,W", [(128, 32, 16, 24), (128, 32, 7, 9), (132, 64, 12, 16), (64, 64, 33, 17),
                                          (128, 32, 20, 30), (32, 128, 40, 60)])      # the last two: x / y roles swapped by the launcher
def test_conv3x3(Cin, Cout, H, W):
    _conv_case(2, Cin, Cout, H, W, 3, seed=Cin + H, cin_ld=ops.ceil_to(Cin, 8))


def test_conv3x3_growth_slice():
    _conv_case(1, 128, 32, 10, 14, 3, seed=9, coff=64, extra=64)


def test_conv_kwidth16():
    _conv_case(2, 16, 64, 12, 16, 3, seed=77, kwidth=16)
    _conv_case(1, 48, 32, 9, 11, 1, seed=78, kwidth=16)


def test_conv5x5_logits_fp32_nchw():
    torch.manual_seed(3)
    B, Cin, Cout, H, W = 2, 64, 3, 16, 24
    x = bf16_round(torch.randn(B, Cin, H, W))
    w = bf16_round(torch.randn(Cout, Cin, 5, 5) / 40.0)
    ref = F.conv2d(x.double(), w.double(), padding=2)
    a = to_mat(x)
    fwd, _, off = ops.conv_taps(5, 2)
    wp, ktot, n_rows = _pack(w, Cout, Cin, 25, off, Cin * 25, 25)
    out = torch.zeros(B, Cout, H, W, dtype=torch.float32, device="cuda")
    d = ops.make_igemm([a.view()], fwd, wp, ktot, n_rows, W, H, B, Cout, out.data_ptr(), 0, out_mode=1)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(out.cpu(), ref)
    assert err < 1e-5, "5x5 fp32 logits relL2 %.3e" % err    # fp32 accumulate of bf16 products, fp32 store


@pytest.mark.parametrize("H,W,tile_w", [(16, 24, 16), (40, 70, 32), (9, 13, 8), (33, 64, 32), (20, 28, 32)])
def test_conv5x5_logits_folded_kernel_columns(H, W, tile_w):
    """out_mode 2: the five kernel columns ride in the GEMM's N (weight row kw*C + n), five kernel-row taps, the epilogue sums
    the horizontal neighbours; tiles overlap by four columns.  Same result as the 25-tap launch (Dense_U_Net_lidar.py:130-131)."""
    torch.manual_seed(4)
    B, Cin, Cout = 2, 64, 3
    x = bf16_round(torch.randn(B, Cin, H, W))
    w = bf16_round(torch.randn(Cout, Cin, 5, 5) / 40.0)
    ref = F.conv2d(x.double(), w.double(), padding=2)
    a = to_mat(x)
    wp = torch.zeros(16, 5 * Cin, dtype=torch.bfloat16)
    wp[:15] = w.permute(3, 0, 2, 1).reshape(15, 5 * Cin).to(torch.bfloat16)      # (kw, n) x (kh, ci)
    wp = wp.cuda()
    out = torch.full((B, Cout, H, W), float("nan"), dtype=torch.float32, device="cuda")
    taps = [(0, kh - 2, 0) for kh in range(5)]
    d = ops.make_igemm([a.view()], taps, wp, 5 * Cin, 16, W, H, B, 15, out.data_ptr(), 0, out_mode=2, fold_kw=5, tile_w=tile_w)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all()), "folded 5x5: output pixels left unwritten"
    err = rel_l2(out.cpu(), ref)
    assert err < 1e-5, "folded 5x5 fp32 logits relL2 %.3e" % err


@pytest.mark.parametrize("C,H,W,OH,OW", [(128, 6, 8, 12, 16), (256, 5, 7, 10, 14), (128, 5, 7, 9, 13), (64, 4, 4, 8, 7),
                                         (128, 20, 30, 40, 60), (64, 20, 30, 39, 59)])   # 20 x 30: x / y roles swapped by the launcher
def test_conv_transpose_phases(C, H, W, OH, OW):
    """nn.ConvTranspose2d(C, C, 3, stride=2, padding=1)(x, output_size=(OH, OW)) as 4 sub-pixel GEMMs."""
    torch.manual_seed(C + OH)
    B = 2
    x = bf16_round(torch.randn(B, C, H, W))
    w = bf16_round(torch.randn(C, C, 3, 3) / (C * 2.25) ** 0.5)   # (Cin, Cout, kh, kw)
    oph, opw = OH - ((H - 1) * 2 - 2 + 3), OW - ((W - 1) * 2 - 2 + 3)
    ref = F.conv_transpose2d(x.double(), w.double(), stride=2, padding=1, output_padding=(oph, opw))
    a = to_mat(x)
    out = ops.new_mat(B, OH, OW, C, zero=True)
    st = new_stats(C)
    for py in range(2):
        for px in range(2):
            taps, off = ops.convt_phase_taps(py, px)
            wp, ktot, n_rows = _pack(w, C, C, len(taps), off, 9, C * 9)
            d = ops.make_igemm([a.view()], taps, wp, ktot, n_rows, W, H, B, C, out.ptr(), C, stats=st,
                               out_stride=(2, 2), out_phase=(py, px), out_hw=(OH, OW))
            ops.run_igemm(d)
    torch.cuda.synchronize()
    got = from_mat(out)
    err = rel_l2(got, ref)
    assert err < TOL, "convT C=%d %dx%d->%dx%d relL2 %.3e" % (C, H, W, OH, OW, err)
    s1, _ = st.totals()
    assert torch.allclose(s1.cpu(), got.permute(1, 0, 2, 3).reshape(C, -1).sum(1), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("K,Cin,Cout", [(3, 128, 32), (1, 160, 128), (5, 64, 3), (3, 132, 64)])
def test_conv_dgrad(K, Cin, Cout):
    """data gradient of Conv2d = igemm with flipped taps and (n, c)-swapped weight packing."""
    torch.manual_seed(K * 100 + Cin)
    B, H, W = 2, 10, 14
    pad = (K - 1) // 2
    w = bf16_round(torch.randn(Cout, Cin, K, K) / (Cin * K * K) ** 0.5)
    g = bf16_round(torch.randn(B, Cout, H, W))
    x = torch.zeros(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    F.conv2d(x, w.double(), padding=pad).backward(g.double())
    ref = x.grad
    gld = ops.ceil_to(Cout, 16)
    gm = to_mat(g, ld=gld)
    _, dg, off = ops.conv_taps(K, pad)
    wp, ktot, n_rows = _pack(w, Cin, Cout, K * K, off, K * K, Cin * K * K)
    ldo = ops.ceil_to(Cin, 8)
    out = ops.new_mat(B, H, W, ldo, zero=True)
    d = ops.make_igemm([gm.view(0, gld)], dg, wp, ktot, n_rows, W, H, B, Cin, out.ptr(), ldo)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out, 0, Cin), ref)
    assert err < TOL, "dgrad K=%d relL2 %.3e" % (K, err)


@pytest.mark.parametrize("C,H,W,OH,OW", [(128, 6, 8, 12, 16), (64, 5, 7, 9, 13)])
def test_conv_transpose_dgrad(C, H, W, OH, OW):
    torch.manual_seed(C + OW)
    B = 2
    w = bf16_round(torch.randn(C, C, 3, 3) / (C * 2.25) ** 0.5)
    g = bf16_round(torch.randn(B, C, OH, OW))
    oph, opw = OH - ((H - 1) * 2 - 2 + 3), OW - ((W - 1) * 2 - 2 + 3)
    x = torch.zeros(B, C, H, W, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(x, w.double(), stride=2, padding=1, output_padding=(oph, opw)).backward(g.double())
    ref = x.grad
    gm = to_mat(g)
    taps, off = ops.convt_dgrad_taps()
    wp, ktot, n_rows = _pack(w, C, C, 9, off, C * 9, 9)     # n = ci, c = co
    srcs = [gm.phase_view(py, px) for py in range(2) for px in range(2)]
    out = ops.new_mat(B, H, W, C, zero=True)
    d = ops.make_igemm(srcs, taps, wp, ktot, n_rows, W, H, B, C, out.ptr(), C)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), ref)
    assert err < TOL, "convT dgrad relL2 %.3e" % err


def test_igemm_large_pixel_count():
    """many tiles / 2 CTAs per SM, n_tile 128: exercises the pipeline for real."""
    _conv_case(4, 256, 128, 64, 96, 1, seed=11)
    _conv_case(2, 128, 32, 64, 96, 3, seed=12)


@pytest.mark.parametrize("K,Cin,Cout,H,W", [(1, 128, 160, 20, 30), (3, 32, 128, 16, 24), (1, 128, 288, 9, 13)])
def test_fused_bn_backward_reduce_matches_separate_pass(K, Cin, Cout, H, W):
    """the epilogue-fused BatchNorm-ReLU backward reduce (dmm_igemm_t.bnb_*) must produce the sums of
    dmm_bn_relu_bwd_reduce run on the stored bf16 output."""
    torch.manual_seed(K * 7 + Cout)
    B = 2
    pad = (K - 1) // 2
    g_in = bf16_round(torch.randn(B, Cin, H, W))
    w = bf16_round(torch.randn(Cout, Cin, K, K) / (Cin * K * K) ** 0.5)
    a = to_mat(g_in)
    fwd, _, off = ops.conv_taps(K, pad)
    wp, ktot, n_rows = _pack(w, Cout, Cin, K * K, off, Cin * K * K, K * K)
    x = to_mat(bf16_round(torch.randn(B, Cout, H, W) * 2 + 0.3))          # raw BN input of the consumer
    gamma = (torch.rand(Cout) + 0.5).cuda()
    beta = (torch.randn(Cout) * 0.3).cuda()
    mean = (torch.randn(Cout) * 0.2 + 0.3).cuda()
    invstd = (torch.rand(Cout) * 0.5 + 0.4).cuda()
    out = ops.new_mat(B, H, W, Cout, zero=True)
    fused, sep = new_stats(Cout), new_stats(Cout)
    bn_f = ops.make_bn_bwd(fused, 0, B * H * W, gamma, beta, mean, invstd)
    d = ops.make_igemm([a.view(0, Cin)], fwd, wp, ktot, n_rows, W, H, B, Cout, out.ptr(), Cout)
    ops.fuse_bn_bwd_reduce(d, x, 0, bn_f)
    ops.run_igemm(d)
    bn_s = ops.make_bn_bwd(sep, 0, B * H * W, gamma, beta, mean, invstd)
    tmp = ops.new_mat(B, H, W, Cout)
    args = ops.make_bn_bwd_args(x, 0, Cout, out.ptr(), out.ld, bn_s, tmp.ptr(), tmp.ld, 0)
    ops.check(ops._lib.load().dmm_bn_relu_bwd_reduce(ops.C.byref(args), ops._stream()), "reduce")
    torch.cuda.synchronize()
    f1, f2 = fused.totals()
    s1, s2 = sep.totals()
    scale1, scale2 = s1.abs().max().item() + 1e-6, s2.abs().max().item() + 1e-6
    assert (f1 - s1).abs().max().item() < 2e-4 * scale1 + 1e-3
    assert (f2 - s2).abs().max().item() < 2e-4 * scale2 + 1e-3


def _bn_setup(x, C, seed):
    """batch statistics of a raw (B,C,H,W) tensor in the engine's double[slots][2][C] layout + random affine."""
    g = torch.Generator().manual_seed(seed)
    gamma = (torch.rand(C, generator=g) + 0.5).cuda()
    beta = (torch.randn(C, generator=g) * 0.3).cuda()
    st = new_stats(C)
    flat = x.permute(1, 0, 2, 3).reshape(C, -1).double()
    st.buf[:C] = flat.sum(1).cuda()                      # slot 0: sums
    st.buf[C:2 * C] = (flat * flat).sum(1).cuda()        # slot 0: sums of squares
    mean = flat.mean(1)
    var = flat.var(1, unbiased=False)
    sc = gamma.cpu().double() / torch.sqrt(var + 1e-5)
    sh = beta.cpu().double() - mean * sc
    act = torch.relu(x.double() * sc.view(1, C, 1, 1) + sh.view(1, C, 1, 1))
    return gamma, beta, st, mean, var, act


@pytest.mark.parametrize("Cin,Cout,H,W", [(64, 128, 16, 24), (96, 128, 9, 13), (352, 128, 20, 30)])
def test_conv1x1_bn_relu_prologue(Cin, Cout, H, W):
    """dmm_igemm_t.pro_*: conv1x1(relu(bn(x))) with the BatchNorm-ReLU applied to the operand tiles in shared memory ==
    nn.Conv2d(nn.ReLU(nn.BatchNorm2d(x))) of the reference (tv:47-50), incl. saved mean / invstd and running statistics."""
    torch.manual_seed(Cin + H)
    B = 2
    x = bf16_round(torch.randn(B, Cin, H, W) * 1.7 + 0.4)
    w = bf16_round(torch.randn(Cout, Cin, 1, 1) / Cin ** 0.5)
    gamma, beta, st, mean, var, act = _bn_setup(x, Cin, 5)
    ref = F.conv2d(bf16_round(act.float()).double(), w.double())
    a = to_mat(x)
    fwd, _, off = ops.conv_taps(1, 0)
    wp, ktot, n_rows = _pack(w, Cout, Cin, 1, off, Cin, 1)
    out = ops.new_mat(B, H, W, Cout, zero=True)
    rm, rv = torch.zeros(Cin, device="cuda"), torch.ones(Cin, device="cuda")
    sm, si = torch.empty(Cin, device="cuda"), torch.empty(Cin, device="cuda")
    d = ops.make_igemm([a.view(0, Cin)], fwd, wp, ktot, n_rows, W, H, B, Cout, out.ptr(), Cout)
    d.pro_enable = 1
    d.pro_bn = ops.make_bn(st, 0, B * H * W, gamma, beta, rm, rv, sm, si, training=True)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), ref)
    assert err < TOL, "prologue conv relL2 %.3e" % err
    n = B * H * W
    assert torch.allclose(sm.cpu().double(), mean, atol=1e-5) and torch.allclose(si.cpu().double(), 1 / torch.sqrt(var + 1e-5), rtol=1e-5)
    assert torch.allclose(rm.cpu().double(), 0.1 * mean, atol=1e-5)
    assert torch.allclose(rv.cpu().double(), 0.9 + 0.1 * var * n / (n - 1), rtol=1e-5)


@pytest.mark.parametrize("Cin,Cout,H,W", [(128, 32, 12, 20), (128, 32, 17, 9), (64, 48, 33, 40)])
def test_conv3x3_bn_relu_prologue(Cin, Cout, H, W):
    """K x K prologue: conv3x3(relu(bn(x)), padding=1) from the RAW x - the halo patch is activated in shared memory and the
    patch pixels outside the image stay zero (the reference pads the ACTIVATED tensor, tv:51-53 / norm2-relu2-conv2)."""
    torch.manual_seed(Cin + W)
    B = 2
    x = bf16_round(torch.randn(B, Cin, H, W) * 1.7 + 0.4)
    w = bf16_round(torch.randn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5)
    gamma, beta, st, mean, var, act = _bn_setup(x, Cin, 6)
    ref = F.conv2d(bf16_round(act.float()).double(), w.double(), padding=1)
    a = to_mat(x)
    fwd, _, off = ops.conv_taps(3, 1)
    wp, ktot, n_rows = _pack(w, Cout, Cin, 9, off, Cin * 9, 9)
    out = ops.new_mat(B, H, W, Cout, zero=True)
    rm, rv = torch.zeros(Cin, device="cuda"), torch.ones(Cin, device="cuda")
    sm, si = torch.empty(Cin, device="cuda"), torch.empty(Cin, device="cuda")
    d = ops.make_igemm([a.view(0, Cin)], fwd, wp, ktot, n_rows, W, H, B, Cout, out.ptr(), Cout)
    d.pro_enable = 1
    d.pro_bn = ops.make_bn(st, 0, B * H * W, gamma, beta, rm, rv, sm, si, training=True)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), ref)
    assert err < TOL, "3x3 prologue conv relL2 %.3e" % err
    assert torch.allclose(sm.cpu().double(), mean, atol=1e-5)


@pytest.mark.parametrize("H,W,coff,extra,Cin,Cout", [(12, 30, 0, 0, 128, 32), (16, 40, 64, 32, 128, 32), (9, 13, 32, 0, 128, 32),
                                                     (33, 64, 0, 32, 128, 32), (5, 91, 32, 32, 128, 32), (18, 61, 0, 0, 136, 64),
                                                     (7, 30, 64, 0, 64, 64)])
def test_conv3x3_growth_folded_kernel_columns(H, W, coff, extra, Cin, Cout):
    """out_mode 3: conv2 of a dense layer (3x3, 128 -> 32) with the three kernel columns folded into N = 96 (weight row
    kw*32 + n), three kernel-row taps, neighbours summed with warp shuffles; bf16 slice of a block buffer + BN statistics
    exactly like the 9-tap launch (tv:51-53)."""
    torch.manual_seed(H * W)
    B = 2
    x = bf16_round(torch.randn(B, Cin, H, W))
    w = bf16_round(torch.randn(Cout, Cin, 3, 3) / (Cin * 9) ** 0.5)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    a = to_mat(x)
    Kp = ops.ceil_to(Cin, 64)
    wp = torch.zeros(3 * Cout, 3, Kp, dtype=torch.bfloat16)
    wp[:, :, :Cin] = w.permute(3, 0, 2, 1).reshape(3 * Cout, 3, Cin).to(torch.bfloat16)                # (kw, n) x (kh, ci)
    wp = wp.reshape(3 * Cout, 3 * Kp).contiguous().cuda()
    ldo = coff + Cout + extra
    out = ops.new_mat(B, H, W, ldo, zero=True)
    st = new_stats(ldo)
    taps = [(0, kh - 1, 0) for kh in range(3)]
    d = ops.make_igemm([a.view(0, Cin)], taps, wp, 3 * Kp, 3 * Cout, W, H, B, 3 * Cout, out.ptr(), ldo, coff=coff, out_mode=3,
                       stats=st, stats_off=coff, n_tile=3 * Cout, fold_kw=3, tile_w=32)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    got = from_mat(out, coff, Cout)
    err = rel_l2(got, ref)
    assert err < TOL, "folded 3x3 growth conv %dx%d: relL2 %.3e" % (H, W, err)
    if coff:
        assert float(from_mat(out, 0, coff).abs().max()) == 0.0, "wrote outside the channel slice"
    if extra:
        assert float(from_mat(out, coff + Cout, extra).abs().max()) == 0.0, "wrote outside the channel slice"
    s1, s2 = st.totals()
    g = got.permute(1, 0, 2, 3).reshape(Cout, -1)
    assert torch.allclose(s1[coff:coff + Cout].cpu(), g.sum(1), rtol=1e-4, atol=1e-3 * g.abs().sum(1).max().item())
    assert torch.allclose(s2[coff:coff + Cout].cpu(), (g * g).sum(1), rtol=1e-4, atol=1e-6)
    other = torch.ones(ldo, dtype=torch.bool)
    other[coff:coff + Cout] = False
    if bool(other.any()):
        assert float(s1.cpu()[other].abs().max()) == 0.0 and float(s2.cpu()[other].abs().max()) == 0.0, "statistics outside the slice"
