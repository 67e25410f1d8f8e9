"""Parity of the HBM-bound kernels (BN-ReLU(+pool) apply, BN-ReLU backward, stem im2col, head input
fwd/bwd, BCE loss+grad, layout converters, Adam) with the torch CPU ops of the reference, fp64 on
identical inputs.  Outputs stored as bf16: relL2 <= 5e-3 (one bf16 rounding); fp32 outputs 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from dmmfods_b200 import ops
from gpu_util import bf16_round, from_mat, new_stats, rel_l2, to_mat

pytestmark = pytest.mark.gpu
TOL_BF16 = 5e-3


def _col_stats(x, st, off=0):
    """fill a Stats row with the exact column sums of x (B,C,H,W) - what a producing kernel would leave."""
    C = x.shape[1]
    v = x.double().permute(1, 0, 2, 3).reshape(C, -1)
    view = st.buf.view(ops.STATS_SLOTS, 2, st.ld)
    view[0, 0, off:off + C] = v.sum(1).cuda()
    view[0, 1, off:off + C] = (v * v).sum(1).cuda()


def _bn_setup(C, seed):
    torch.manual_seed(seed)
    gamma = (torch.rand(C) + 0.5).cuda()
    beta = (torch.randn(C) * 0.3).cuda()
    rm = torch.randn(C).cuda()
    rv = (torch.rand(C) + 0.5).cuda()
    sm = torch.zeros(C).cuda()
    si = torch.zeros(C).cuda()
    return gamma, beta, rm, rv, sm, si


@pytest.mark.parametrize("C,H,W,pool", [(64, 16, 24, 0), (96, 9, 7, 0), (1024, 4, 6, 0), (256, 8, 12, 1),
                                        (104, 7, 9, 1), (64, 16, 24, 2), (64, 9, 11, 2)])
def test_bn_relu_apply(C, H, W, pool):
    B = 2
    gamma, beta, rm, rv, sm, si = _bn_setup(C, C + H)
    x = bf16_round(torch.randn(B, C, H, W) * 2 + 0.5)
    rm0, rv0 = rm.clone(), rv.clone()
    ref = F.relu(F.batch_norm(x.double(), rm0.cpu().double(), rv0.cpu().double(), gamma.cpu().double(), beta.cpu().double(),
                              training=True, momentum=0.1, eps=1e-5))
    ref_rm, ref_rv = rm0.cpu().double(), rv0.cpu().double()
    F.batch_norm(x.double(), ref_rm, ref_rv, None, None, training=True, momentum=0.1, eps=1e-5)
    if pool == 1:
        ref = F.avg_pool2d(ref, 2, 2)
    elif pool == 2:
        ref = F.max_pool2d(ref, 3, 2, 1)
    xm = to_mat(x, ld=C + 8, c0=8)
    st = new_stats(C + 8)
    _col_stats(x, st, 8)
    OH, OW = ref.shape[2], ref.shape[3]
    y = ops.new_mat(B, OH, OW, C + 16, zero=True)
    yst = new_stats(C + 16)
    bn = ops.make_bn(st, 8, B * H * W, gamma, beta, rm, rv, sm, si, training=True)
    d = ops.make_bn_apply(xm, 8, C, bn, y, 16, pool=pool, ystats=yst, ystats_off=16)
    ops.run_bn_apply(d)
    torch.cuda.synchronize()
    got = from_mat(y, 16, C)
    err = rel_l2(got, ref)
    assert err < TOL_BF16, "bn_relu_apply pool=%d relL2 %.3e" % (pool, err)
    assert float(from_mat(y, 0, 16).abs().max()) == 0.0
    assert torch.allclose(rm.cpu().double(), ref_rm, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.cpu().double(), ref_rv, rtol=1e-5, atol=1e-6)
    v = x.double().permute(1, 0, 2, 3).reshape(C, -1)
    assert torch.allclose(sm.cpu().double(), v.mean(1), rtol=1e-5, atol=1e-6)
    assert torch.allclose(si.cpu().double(), 1.0 / torch.sqrt(v.var(1, unbiased=False) + 1e-5), rtol=1e-5)
    s1, s2 = yst.totals()
    g = got.permute(1, 0, 2, 3).reshape(C, -1)
    assert torch.allclose(s1[16:].cpu(), g.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s2[16:].cpu(), (g * g).sum(1), rtol=1e-4, atol=1e-2)


def test_bn_relu_apply_eval_mode():
    B, C, H, W = 2, 64, 6, 10
    gamma, beta, rm, rv, sm, si = _bn_setup(C, 4)
    x = bf16_round(torch.randn(B, C, H, W))
    ref = F.relu(F.batch_norm(x.double(), rm.cpu().double(), rv.cpu().double(), gamma.cpu().double(), beta.cpu().double(),
                              training=False, eps=1e-5))
    xm = to_mat(x)
    y = ops.new_mat(B, H, W, C)
    bn = ops.make_bn(None, 0, B * H * W, gamma, beta, rm, rv, None, None, training=False)
    ops.run_bn_apply(ops.make_bn_apply(xm, 0, C, bn, y, 0))
    torch.cuda.synchronize()
    assert rel_l2(from_mat(y), ref) < TOL_BF16


@pytest.mark.parametrize("C,H,W,gmode,gf32,out_mode", [(64, 12, 16, 0, False, 0), (96, 7, 9, 0, False, 2),
                                                       (256, 8, 12, 1, False, 1), (104, 7, 9, 1, False, 2),
                                                       (64, 16, 24, 2, True, 0), (64, 9, 11, 2, True, 0),
                                                       (128, 6, 10, 0, True, 1)])
def test_bn_relu_backward(C, H, W, gmode, gf32, out_mode):
    """d/dx of pool(relu(bn(x))) for a given upstream gradient, + dgamma / dbeta."""
    B = 2
    gamma, beta, _, _, sm, si = _bn_setup(C, C + W)
    x = bf16_round(torch.randn(B, C, H, W) * 1.5 + 0.2)
    xd = x.double().requires_grad_(True)
    gd = gamma.cpu().double().requires_grad_(True)
    bd = beta.cpu().double().requires_grad_(True)
    a = F.relu(F.batch_norm(xd, None, None, gd, bd, training=True, eps=1e-5))
    if gmode == 1:
        a = F.avg_pool2d(a, 2, 2)
    elif gmode == 2:
        a = F.max_pool2d(a, 3, 2, 1)
    g = torch.randn_like(a)
    g = g.float() if gf32 else bf16_round(g.float())
    a.backward(g.double())
    v = x.double().permute(1, 0, 2, 3).reshape(C, -1)
    sm.copy_(v.mean(1).float())
    si.copy_((1.0 / torch.sqrt(v.var(1, unbiased=False) + 1e-5)).float())
    xm = to_mat(x)
    OH, OW = g.shape[2], g.shape[3]
    if gf32:
        gt = g.permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda()
        gptr, ldg = gt.data_ptr(), C
    else:
        gm = to_mat(g)
        gptr, ldg = gm.ptr(), gm.ld
    sums = new_stats(C)
    dgam = torch.zeros(C, device="cuda")
    dbet = torch.zeros(C, device="cuda")
    bnb = ops.make_bn_bwd(sums, 0, B * H * W, gamma, beta, sm, si, dgam, dbet)
    prev = torch.randn(B * H * W, C) if out_mode == 2 else torch.zeros(B * H * W, C)
    if out_mode == 0:
        out = ops.new_mat(B, H, W, C, zero=True)
        optr = out.ptr()
    else:
        out = prev.clone().cuda()
        optr = out.data_ptr()
    d = ops.make_bn_bwd_args(xm, 0, C, gptr, ldg, bnb, optr, C, out_mode, gmode=gmode, g_is_f32=gf32)
    ops.run_bn_bwd(d)
    torch.cuda.synchronize()
    if out_mode == 0:
        got = from_mat(out)
        tol = TOL_BF16
    else:
        got = (out.cpu().double() - (prev.double() if out_mode == 2 else 0)).reshape(B, H, W, C).permute(0, 3, 1, 2)
        tol = 2e-5
    err = rel_l2(got, xd.grad)
    assert err < tol, "bn bwd gmode=%d out_mode=%d relL2 %.3e" % (gmode, out_mode, err)
    assert rel_l2(dgam.cpu(), gd.grad) < 1e-4
    assert rel_l2(dbet.cpu(), bd.grad) < 1e-4


@pytest.mark.parametrize("C1,C2,H,W", [(3, 0, 16, 24), (3, 1, 18, 14), (1, 0, 9, 13)])
def test_stem_im2col_conv0(C1, C2, H, W):
    """conv0 7x7 stride 2 padding 3 = im2col + 1x1 igemm."""
    torch.manual_seed(C1 * 10 + C2)
    B = 2
    x1 = (torch.rand(B, C1, H, W) * 255).float()
    x2 = (torch.rand(B, C2, H, W) * 255).float() if C2 else None
    Cin = C1 + C2
    w = bf16_round(torch.randn(64, Cin, 7, 7) / (Cin * 49) ** 0.5)
    xin = x1 if x2 is None else torch.cat((x1, x2), 1)
    ref = F.conv2d(bf16_round(xin).double(), w.double(), stride=2, padding=3)
    OH, OW = ref.shape[2], ref.shape[3]
    kpad = ops.ceil_to(Cin * 49, 8)
    col = ops.new_mat(B, OH, OW, kpad)
    ops.im2col_7x7s2(x1.cuda(), None if x2 is None else x2.cuda(), col)
    Kp = ops.ceil_to(Cin * 49, 64)
    wp = torch.empty(64, Kp, dtype=torch.bfloat16, device="cuda")
    ops.pack_weights(w.cuda(), wp, 64, 64, Cin * 49, 1, [0], Cin * 49, 1)
    out = ops.new_mat(B, OH, OW, 64)
    d = ops.make_igemm([col.view(0, kpad)], [(0, 0, 0)], wp, Kp, 64, OW, OH, B, 64, out.ptr(), 64)
    ops.run_igemm(d)
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), ref)
    assert err < 1e-2, "conv0 relL2 %.3e" % err


def _head_ref(u, x1, x2, gamma, beta):
    up = F.interpolate(u, scale_factor=2, mode="nearest")
    parts = [up, x1] + ([x2] if x2 is not None else [])
    cat = torch.cat(parts, 1)
    return F.relu(F.batch_norm(cat, None, None, gamma, beta, training=True, eps=1e-5)), cat


@pytest.mark.parametrize("C1,C2", [(3, 1), (3, 0)])
def test_head_input_forward_backward(C1, C2):
    torch.manual_seed(40 + C2)
    B, Cu, H, W = 2, 128, 12, 16
    Ct = Cu + C1 + C2
    ldo = ops.ceil_to(Ct, 8)
    u = bf16_round(torch.randn(B, Cu, H // 2, W // 2))
    x1 = (torch.rand(B, C1, H, W) * 255).float()
    x2 = (torch.rand(B, C2, H, W) * 255).float() if C2 else None
    gamma = (torch.rand(Ct) + 0.5).cuda()
    beta = (torch.randn(Ct) * 0.3).cuda()
    ud = u.double().requires_grad_(True)
    gd = gamma.cpu().double().requires_grad_(True)
    bd = beta.cpu().double().requires_grad_(True)
    ref, cat = _head_ref(ud, x1.double(), None if x2 is None else x2.double(), gd, bd)
    g = bf16_round(torch.randn(B, Ct, H, W))
    ref.backward(g.double())

    um = to_mat(u)
    ust = new_stats(Cu)
    _col_stats(u, ust)
    xst = new_stats(8)
    x1c = x1.cuda()
    x2c = None if x2 is None else x2.cuda()
    ops.nchw_stats(x1c, xst, 0)
    if x2c is not None:
        ops.nchw_stats(x2c, xst, C1)
    rm = torch.zeros(Ct, device="cuda")
    rv = torch.ones(Ct, device="cuda")
    sm = torch.zeros(Ct, device="cuda")
    si = torch.zeros(Ct, device="cuda")
    out = ops.new_mat(B, H, W, ldo)
    d = ops.Head()
    d.u, d.ldu, d.Cu = um.ptr().value, um.ld, Cu
    d.x1, d.C1 = x1c.data_ptr(), C1
    d.x2, d.C2 = (x2c.data_ptr() if x2c is not None else None), C2
    d.B, d.H, d.W = B, H, W
    d.bn_u = ops.make_bn(ust, 0, B * (H // 2) * (W // 2), gamma, beta, rm, rv, sm, si, rep=4.0)
    d.bn_x = ops.make_bn(xst, 0, B * H * W, gamma, beta, rm, rv, sm, si, c0=Cu)
    d.out, d.ldo = out.ptr().value, ldo
    from dmmfods_b200 import _lib
    import ctypes
    _lib.check(_lib.load().dmm_head_input(ctypes.byref(d), ops._stream()), "dmm_head_input")
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out, 0, Ct), ref.detach())
    assert err < TOL_BF16, "head_input relL2 %.3e" % err
    if ldo > Ct:
        assert float(from_mat(out, Ct, ldo - Ct).abs().max()) == 0.0
    # running stats: same as nn.BatchNorm2d over the concatenated (up-sampled) tensor
    ref_rm, ref_rv = torch.zeros(Ct, dtype=torch.float64), torch.ones(Ct, dtype=torch.float64)
    F.batch_norm(cat.detach(), ref_rm, ref_rv, None, None, training=True, momentum=0.1, eps=1e-5)
    assert torch.allclose(rm.cpu().double(), ref_rm, rtol=1e-4, atol=1e-4)
    assert torch.allclose(rv.cpu().double(), ref_rv, rtol=1e-4, atol=1e-4)

    gm = to_mat(g, ld=ldo)
    sums_u = new_stats(Cu)
    sums_x = new_stats(8)
    dgam = torch.zeros(Ct, device="cuda")
    dbet = torch.zeros(Ct, device="cuda")
    du = ops.new_mat(B, H // 2, W // 2, Cu)
    hb = ops.HeadBwd()
    hb.u, hb.ldu, hb.Cu = um.ptr().value, um.ld, Cu
    hb.x1, hb.C1 = x1c.data_ptr(), C1
    hb.x2, hb.C2 = (x2c.data_ptr() if x2c is not None else None), C2
    hb.B, hb.H, hb.W = B, H, W
    hb.g, hb.ldg = gm.ptr().value, gm.ld
    hb.bn_u = ops.make_bn_bwd(sums_u, 0, B * H * W, gamma, beta, sm, si, dgam, dbet)
    hb.bn_x = ops.make_bn_bwd(sums_x, 0, B * H * W, gamma, beta, sm, si, dgam, dbet, c0=Cu)
    hb.du, hb.lddu = du.ptr().value, du.ld
    lib = _lib.load()
    _lib.check(lib.dmm_head_input_bwd_reduce(ctypes.byref(hb), ops._stream()), "head bwd reduce")
    _lib.check(lib.dmm_head_input_bwd_apply(ctypes.byref(hb), ops._stream()), "head bwd apply")
    torch.cuda.synchronize()
    err = rel_l2(from_mat(du), ud.grad)
    assert err < TOL_BF16, "head_input bwd relL2 %.3e" % err
    assert rel_l2(dgam.cpu(), gd.grad) < 1e-4
    assert rel_l2(dbet.cpu(), bd.grad) < 1e-4


def test_bce_logits_loss_and_grad():
    torch.manual_seed(7)
    B, C, H, W = 2, 3, 16, 24
    x = (torch.randn(B, C, H, W) * 6).float()
    x[0, 0, 0, :4] = torch.tensor([0.0, 100.0, -100.0, 1e-8])
    t = torch.tensor([0.0, 0.3, 0.5, 0.75, 1.0])[torch.randint(0, 5, (B, C, H, W))]
    xd = x.double().requires_grad_(True)
    ref = F.binary_cross_entropy_with_logits(xd, t.double(), reduction="none")
    ref.backward(torch.ones_like(ref))
    xc, tc = x.cuda(), t.cuda()
    loss = torch.empty_like(xc)
    grad = torch.empty_like(xc)
    cs = torch.zeros(C, dtype=torch.float64, device="cuda")
    ops.bce_logits(xc, tc, loss, grad, cs)
    torch.cuda.synchronize()
    assert torch.allclose(loss.cpu().double(), ref.detach(), rtol=2e-6, atol=1e-7)
    assert torch.allclose(grad.cpu().double(), xd.grad, rtol=2e-6, atol=1e-7)
    assert torch.allclose(cs.cpu(), ref.detach().sum(dim=(0, 2, 3)), rtol=1e-6)
    # the fp32 torch op the reference runs (Agent.py:54): agreement to fp32 round-off
    ref32 = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    assert torch.allclose(loss.cpu(), ref32, rtol=1e-5, atol=2e-6)   # |x| ~ 30: fp32 cancellation in x - x*t


def test_layout_converters():
    torch.manual_seed(8)
    B, C, H, W = 2, 3, 10, 12
    x = torch.randn(B, C, H, W)
    out = ops.new_mat(B, H, W, 16)
    ops.nchw_to_nhwc_bf16(x.cuda(), out)
    torch.cuda.synchronize()
    assert torch.equal(from_mat(out, 0, C), bf16_round(x).double())
    assert float(from_mat(out, C, 16 - C).abs().max()) == 0.0
    src = torch.randn(B * H * W, 96).cuda()
    dst = ops.new_mat(B, H, W, 32)
    ops.rows_f32_to_bf16(src, 64, 32, dst)
    torch.cuda.synchronize()
    assert torch.equal(dst.t.float().cpu(), bf16_round(src[:, 64:96].cpu()))


def test_adam_flat_matches_torch_adam():
    torch.manual_seed(9)
    n = 10007
    p0 = torch.randn(n)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    p = p0.clone().cuda()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n)
        ref_p.grad = g.clone()
        opt.step()
        ops.adam_flat(p, g.cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    torch.cuda.synchronize()
    assert torch.allclose(p.cpu(), ref_p.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("H,W,gf32", [(18, 22, True), (18, 22, False), (17, 21, False), (9, 12, False)])
def test_maxpool_backward_with_recorded_argmax(H, W, gf32):
    """stem path of the engine: forward records each window's winner, backward routes through the codes and the reduce
    pass stores dz so that the apply pass runs with gmode 0 (bf16 gradients take the 2x2-quad kernel)."""
    torch.manual_seed(51)
    B, C = 2, 64
    gamma, beta, rm, rv, sm, si = _bn_setup(C, 3)
    x = bf16_round(torch.randn(B, C, H, W) * 1.5 + 0.2)
    x[:, :, 4:8, 4:8] = x[:, :, 4:5, 4:5]                      # exact ties inside windows
    xd = x.double().requires_grad_(True)
    gd = gamma.cpu().double().requires_grad_(True)
    bd = beta.cpu().double().requires_grad_(True)
    a = F.max_pool2d(F.relu(F.batch_norm(xd, None, None, gd, bd, training=True, eps=1e-5)), 3, 2, 1)
    g = torch.randn_like(a).float()
    if not gf32:
        g = bf16_round(g)
    a.backward(g.double())
    OH, OW = a.shape[2], a.shape[3]
    xm = to_mat(x)
    st = new_stats(C)
    _col_stats(x, st)
    y = ops.new_mat(B, OH, OW, C)
    amax = torch.zeros(B * OH * OW, C, dtype=torch.uint8, device="cuda")
    bn = ops.make_bn(st, 0, B * H * W, gamma, beta, rm, rv, sm, si, training=True)
    d = ops.make_bn_apply(xm, 0, C, bn, y, 0, pool=2)
    d.argmax, d.ldarg = amax.data_ptr(), C
    ops.run_bn_apply(d)
    torch.cuda.synchronize()
    assert rel_l2(from_mat(y), a.detach()) < TOL_BF16
    gt = g.permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda()
    if not gf32:
        gt = gt.to(torch.bfloat16)
    sums = new_stats(C)
    dgam = torch.zeros(C, device="cuda")
    dbet = torch.zeros(C, device="cuda")
    bnb = ops.make_bn_bwd(sums, 0, B * H * W, gamma, beta, sm, si, dgam, dbet)
    dz = ops.new_mat(B, H, W, C)
    out = ops.new_mat(B, H, W, C)
    d1 = ops.make_bn_bwd_args(xm, 0, C, gt.data_ptr(), C, bnb, out.ptr(), C, 0, gmode=2, g_is_f32=gf32)
    d1.argmax, d1.ldarg = amax.data_ptr(), C
    d1.dz_out, d1.lddz = dz.ptr().value, C
    d2 = ops.make_bn_bwd_args(xm, 0, C, dz.ptr(), C, bnb, out.ptr(), C, 0, gmode=0)
    import ctypes
    from dmmfods_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.dmm_bn_relu_bwd_reduce(ctypes.byref(d1), ops._stream()), "reduce")
    _lib.check(lib.dmm_bn_relu_bwd_apply(ctypes.byref(d2), ops._stream()), "apply")
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), xd.grad)
    assert err < 1e-2, "max-pool bwd via argmax codes relL2 %.3e" % err       # dz stored as bf16 in between
    assert rel_l2(dgam.cpu(), gd.grad) < 5e-3
    assert rel_l2(dbet.cpu(), bd.grad) < 5e-3


def test_dlogits_im2col():
    torch.manual_seed(52)
    B, C, H, W, K = 2, 3, 9, 12, 5
    dl = torch.randn(B, C, H, W)
    ld = 80
    out = ops.new_mat(B, H, W, ld)
    import ctypes
    from dmmfods_b200 import _lib
    _lib.check(_lib.load().dmm_dlogits_im2col(ctypes.c_void_p(dl.cuda().data_ptr()), B, C, H, W, K, out.ptr(), ld, ops._stream()),
               "dlogits_im2col")
    torch.cuda.synchronize()
    got = from_mat(out)                                        # (B, ld, H, W)
    pad = K // 2
    dp = F.pad(bf16_round(dl).double(), (pad, pad, pad, pad))
    for t in range(K * K):
        kh, kw = t // K, t % K
        ref = dp[:, :, 2 * pad - kh:2 * pad - kh + H, 2 * pad - kw:2 * pad - kw + W]      # dl(y - (kh-pad), x - (kw-pad))
        assert torch.equal(got[:, t * C:(t + 1) * C], ref), t
    assert float(got[:, K * K * C:].abs().max()) == 0.0


@pytest.mark.parametrize("C1,C2,H,W,ld", [(3, 0, 10, 24, 32), (3, 1, 7, 13, 32), (1, 0, 9, 18, 16)])
def test_stem_unfold_w7s2(C1, C2, H, W, ld):
    """out[(b, iy, ox)][kw*C + c] = x[c](iy, 2 ox + kw - 3): the operand that makes conv0 (7x7, stride 2, padding 3,
    Dense_U_Net_lidar.py:73-74) a 7-tap vertical convolution over even / odd input rows."""
    import ctypes
    from dmmfods_b200 import _lib
    torch.manual_seed(61)
    B, Cc = 2, C1 + C2
    x1 = torch.randn(B, C1, H, W)
    x2 = torch.randn(B, C2, H, W) if C2 else None
    OW = (W - 1) // 2 + 1
    out = torch.full((B * H * OW, ld), float("nan"), dtype=torch.bfloat16, device="cuda")
    x1c, x2c = x1.cuda(), (x2.cuda() if C2 else None)
    _lib.check(_lib.load().dmm_unfold_w7s2(ctypes.c_void_p(x1c.data_ptr()), C1, ctypes.c_void_p(x2c.data_ptr()) if C2 else None, C2,
                                           B, H, W, ctypes.c_void_p(out.data_ptr()), ld, ops._stream()), "unfold_w7s2")
    torch.cuda.synchronize()
    x = torch.cat([x1, x2], 1) if C2 else x1
    xp = F.pad(bf16_round(x), (3, 3 + 2, 0, 0))
    got = out.float().cpu().view(B, H, OW, ld)
    for kw in range(7):
        ref = xp[:, :, :, kw:kw + 2 * OW:2].permute(0, 2, 3, 1)              # (B, H, OW, C)
        assert torch.equal(got[..., kw * Cc:(kw + 1) * Cc], ref), kw
    assert float(got[..., 7 * Cc:].abs().max()) == 0.0


@pytest.mark.parametrize("W,ld", [(13, 16), (16, 16), (24, 24)])
def test_dlogits_unfold_w(W, ld):
    """column kw*C + n = dlogits[n](y, x - (kw - K/2)): the operand that turns refine1's data gradient into a 5-tap vertical
    convolution (Dense_U_Net_lidar.py:130-131 backward)."""
    torch.manual_seed(53)
    B, C, H, K = 2, 3, 7, 5
    dl = torch.randn(B, C, H, W)
    out = ops.new_mat(B, H, W, ld)
    import ctypes
    from dmmfods_b200 import _lib
    _lib.check(_lib.load().dmm_dlogits_unfold_w(ctypes.c_void_p(dl.cuda().data_ptr()), B, C, H, W, K, out.ptr(), ld, ops._stream()),
               "dlogits_unfold_w")
    torch.cuda.synchronize()
    got = from_mat(out)
    pad = K // 2
    dp = F.pad(bf16_round(dl).double(), (pad, pad, 0, 0))
    for kw in range(K):
        ref = dp[:, :, :, 2 * pad - kw:2 * pad - kw + W]
        assert torch.equal(got[:, kw * C:(kw + 1) * C], ref), kw
    assert float(got[:, K * C:].abs().max()) == 0.0


@pytest.mark.parametrize("planar", [False, True])
def test_dense_block_deferred_bn_backward(planar):
    """dense-block backward algebra: every consumer i of a block buffer writes the slab A_i*dz_i and its sums
    (dmm_bn_relu_bwd_contrib), dmm_bn_bwd_finalize turns the sums into the per-channel correction vectors, and
    dmm_grad_gather forms sum_i slab_i - sum_i (k1_i + k2_i (x - mean)) for a channel range.  Must equal autograd through
    sum_i <g_i, relu(bn_i(x[:, :C_i]))> (tv:47-50,96-104 backward) on the same bf16-rounded inputs."""
    import ctypes as C
    from dmmfods_b200 import _lib
    lib = _lib.load()
    B, H, W, Cbuf, gw = 2, 9, 11, 96, 32
    consumers = [64, 96, 96]
    c0, Cg = 32, 32
    torch.manual_seed(5)
    x = bf16_round(torch.randn(B, Cbuf, H, W) * 1.3 - 0.1)
    xd = x.double().requires_grad_(True)
    xm = to_mat(x)
    P = B * H * W
    v = x.double().permute(1, 0, 2, 3).reshape(Cbuf, -1)
    mean = v.mean(1).float().cuda()
    invstd = (1.0 / torch.sqrt(v.var(1, unbiased=False) + 1e-5)).float().cuda()
    keep, loss = [], 0.0
    g = _lib.GradGather()
    refs = []
    for i, Ci in enumerate(consumers):
        gamma, beta, _, _, _, _ = _bn_setup(Ci, 40 + i)
        gd = gamma.cpu().double().requires_grad_(True)
        bd = beta.cpu().double().requires_grad_(True)
        gi = bf16_round(torch.randn(B, Ci, H, W))
        loss = loss + (F.relu(F.batch_norm(xd[:, :Ci], None, None, gd, bd, training=True, eps=1e-5)) * gi.double()).sum()
        gm = to_mat(gi)
        sums = new_stats(Ci)
        dgam, dbet = torch.zeros(Ci, device="cuda"), torch.zeros(Ci, device="cuda")
        bnb = ops.make_bn_bwd(sums, 0, P, gamma, beta, mean, invstd, dgam, dbet)
        slab = torch.zeros(P * Ci, dtype=torch.bfloat16, device="cuda")
        d = ops.make_bn_bwd_args(xm, 0, Ci, gm.ptr(), gm.ld, bnb, slab.data_ptr(), Ci, 0)
        if planar:
            d.out_gw, d.out_plane = gw, P * gw
        _lib.check(lib.dmm_bn_relu_bwd_contrib(C.byref(d), None), "contrib")
        kvec = torch.zeros(2 * Ci, device="cuda")
        _lib.check(lib.dmm_bn_bwd_finalize(C.byref(bnb), Ci, C.c_void_p(kvec.data_ptr()), None), "finalize")
        if planar:
            g.src[i], g.ld[i], g.plane[i] = slab.data_ptr() + 2 * (c0 // gw) * P * gw, gw, P * gw
        else:
            g.src[i], g.ld[i], g.plane[i] = slab.data_ptr() + 2 * c0, Ci, 0
        g.k1[i], g.k2[i] = kvec.data_ptr() + 4 * c0, kvec.data_ptr() + 4 * (Ci + c0)
        keep += [gamma, beta, gm, sums, dgam, dbet, slab, kvec, bnb]
        refs.append((dgam, dbet, gd, bd))
    loss.backward()
    g.gw = gw if planar else 0
    g.nsrc = g.nk = len(consumers)
    g.mean = mean.data_ptr() + 4 * c0
    g.x, g.ldx = xm.ptr(c0).value, xm.ld
    out = ops.new_mat(B, H, W, Cg, zero=True)
    g.rows, g.C, g.out, g.ldo = P, Cg, out.ptr().value, out.ld
    _lib.check(lib.dmm_grad_gather(C.byref(g), None), "gather")
    torch.cuda.synchronize()
    err = rel_l2(from_mat(out), xd.grad[:, c0:c0 + Cg])
    assert err < 1e-2, "deferred BN backward relL2 %.3e" % err          # three bf16 slabs summed
    for dgam, dbet, gd, bd in refs:
        assert rel_l2(dgam.cpu(), gd.grad) < 1e-4 and rel_l2(dbet.cpu(), bd.grad) < 1e-4
