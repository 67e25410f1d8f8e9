"""Parity of the tcgen05 weight-gradient GEMM (dmm_conv_wgrad + dmm_unpack_wgrad) with the weight
gradients torch autograd computes for nn.Conv2d / nn.ConvTranspose2d (aten::convolution_backward),
fp64 on identical bf16-rounded inputs.  fp32 accumulation: relL2 <= 1e-4."""
import pytest
import torch
import torch.nn.functional as F

from dmmfods_b200 import ops
from gpu_util import bf16_round, rel_l2, to_mat

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _run_plan(xv, yvs, taps, M, N, W, H, B, splits=0):
    plan = ops.plan_conv_wgrad(xv, yvs, taps, M, N)
    dw = torch.zeros(plan["rows"] * plan["ld"], dtype=torch.float32, device="cuda")
    for kw in plan["launches"]:
        ops.run_wgrad(ops.make_wgrad(W=W, H=H, B=B, dw=dw, ld=plan["ld"], splits=splits, **kw))
    return plan, dw


def _conv_wgrad_case(B, Cin, Cout, H, W, K, seed, n_tile=None, splits=0):
    torch.manual_seed(seed)
    pad = (K - 1) // 2
    x = bf16_round(torch.randn(B, Cin, H, W))
    g = bf16_round(torch.randn(B, Cout, H, W))
    w = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, padding=pad).backward(g.double())
    ref = w.grad
    xm = to_mat(x, ld=ops.ceil_to(Cin, 8))
    N = ops.ceil_to(Cout, 16)
    gm = to_mat(g, ld=N)
    fwd, _, off = ops.conv_taps(K, pad)
    T = K * K
    plan, dw = _run_plan(xm.view(0, Cin), [gm.view(0, N)], fwd, Cin, N, W, H, B, splits)
    grad = torch.full((Cout, Cin, K, K), float("nan"), dtype=torch.float32, device="cuda")
    off = [off[t] for t in plan["tap_order"]]
    ops.unpack_wgrad(dw, plan["dt"], plan["dm"], plan["dn"], Cin, Cout, grad, T, off, Cin * K * K, K * K)
    torch.cuda.synchronize()
    err = rel_l2(grad.cpu(), ref)
    assert err < TOL, "wgrad K=%d Cin=%d Cout=%d: relL2 %.3e (%d launches)" % (K, Cin, Cout, err, len(plan["launches"]))


@pytest.mark.parametrize("Cin,Cout", [(128, 128), (256, 128), (96, 128), (64, 64), (1024, 256), (136, 64), (992, 128),
                                      (2048, 512), (512, 256), (160, 32), (64, 16)])
def test_wgrad_1x1(Cin, Cout):
    _conv_wgrad_case(2, Cin, Cout, 12, 20, 1, seed=Cin + Cout)


@pytest.mark.parametrize("Cout", [32, 16, 48])
def test_wgrad_3x3_growth(Cout):
    """conv2 of a dense layer: activation tile loaded once, 9 shifted copies of the narrow output gradient."""
    _conv_wgrad_case(2, 128, Cout, 12, 20, 3, seed=21 + Cout)


def test_wgrad_3x3_odd_sizes_and_head():
    _conv_wgrad_case(1, 128, 32, 7, 9, 3, seed=22)
    _conv_wgrad_case(2, 132, 64, 10, 14, 3, seed=23)       # refine0: roles swapped, two launches
    _conv_wgrad_case(2, 68, 32, 10, 14, 3, seed=27)


def test_wgrad_5x5_three_classes():
    _conv_wgrad_case(2, 64, 3, 12, 16, 5, seed=24)


@pytest.mark.parametrize("H,W", [(12, 16), (33, 41), (64, 96)])
def test_wgrad_5x5_three_classes_unfolded_columns(H, W):
    """refine1 through the horizontally unfolded d(logits) (dmm_dlogits_unfold_w): five kernel-row taps merged into ONE
    N = 80 MMA per k-step (vertical family, atom stride = one patch row); gradient column kw*C + n -> dW[n, :, kh, kw]."""
    import ctypes
    from dmmfods_b200 import _lib
    torch.manual_seed(29)
    B, Cin, Cout, K = 2, 64, 3, 5
    x = bf16_round(torch.randn(B, Cin, H, W))
    g = torch.randn(B, Cout, H, W)
    w = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, padding=2).backward(bf16_round(g).double())
    xm = to_mat(x)
    gl = ops.new_mat(B, H, W, 16)
    _lib.check(_lib.load().dmm_dlogits_unfold_w(ctypes.c_void_p(g.cuda().data_ptr()), B, Cout, H, W, K, gl.ptr(), 16, ops._stream()),
               "unfold")
    taps = [(0, kh - 2, 0) for kh in range(K)]
    plan, dw = _run_plan(xm.view(0, Cin), [gl.view(0, 16)], taps, Cin, 16, W, H, B)
    assert len(plan["launches"]) == 1
    grad = torch.full((Cout, Cin, K, K), float("nan"), dtype=torch.float32, device="cuda")
    dwv = dw.view(-1)
    for i, t in enumerate(plan["tap_order"]):            # (tap, m, n) -> dW[n % C, m, kh, n // C]
        kh = t
        blk = torch.stack([dwv[i * plan["dt"] + torch.arange(Cin, device="cuda") * plan["dm"] + n * plan["dn"]] for n in range(K * Cout)], 1)
        grad[:, :, kh, :] = blk.view(Cin, K, Cout).permute(2, 0, 1)
    torch.cuda.synchronize()
    err = rel_l2(grad.cpu(), w.grad)
    assert err < TOL, "unfolded 5x5 wgrad relL2 %.3e" % err


def test_wgrad_many_tiles_split():
    _conv_wgrad_case(4, 256, 128, 64, 96, 1, seed=25, splits=0)
    _conv_wgrad_case(4, 128, 32, 32, 48, 3, seed=26, splits=7)
    _conv_wgrad_case(2, 64, 3, 64, 96, 5, seed=28)


@pytest.mark.parametrize("C,H,W,OH,OW", [(128, 6, 8, 12, 16), (64, 5, 7, 9, 13), (256, 6, 8, 12, 16), (512, 4, 6, 8, 12)])
def test_wgrad_conv_transpose(C, H, W, OH, OW):
    torch.manual_seed(C + OH)
    B = 2
    x = bf16_round(torch.randn(B, C, H, W))
    g = bf16_round(torch.randn(B, C, OH, OW))
    oph, opw = OH - ((H - 1) * 2 - 2 + 3), OW - ((W - 1) * 2 - 2 + 3)
    w = torch.zeros(C, C, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(x.double(), w, stride=2, padding=1, output_padding=(oph, opw)).backward(g.double())
    ref = w.grad
    xm = to_mat(x)
    gm = to_mat(g)
    taps, off = ops.convt_wgrad_taps()
    ys = [gm.phase_view(py, px) for py in range(2) for px in range(2)]
    plan, dw = _run_plan(xm.view(), ys, taps, C, C, W, H, B)
    grad = torch.full((C, C, 3, 3), float("nan"), dtype=torch.float32, device="cuda")
    off = [off[t] for t in plan["tap_order"]]
    ops.unpack_wgrad(dw, plan["dt"], plan["dm"], plan["dn"], C, C, grad, 9, off, 9, C * 9)      # weight (Cin=m, Cout=n, kh, kw)
    torch.cuda.synchronize()
    err = rel_l2(grad.cpu(), ref)
    assert err < TOL, "convT wgrad relL2 %.3e" % err


@pytest.mark.parametrize("Cin,Cout", [(96, 128), (352, 128), (64, 128)])
def test_wgrad_1x1_bn_relu_prologue(Cin, Cout):
    """dmm_wgrad_t.pro_*: the weight gradient of conv1x1(relu(bn(x))) computed from the RAW x (activated in shared memory)."""
    torch.manual_seed(Cin)
    B, H, W = 2, 12, 20
    x = bf16_round(torch.randn(B, Cin, H, W) * 1.5 + 0.2)
    g = bf16_round(torch.randn(B, Cout, H, W))
    gen = torch.Generator().manual_seed(3)
    gamma = (torch.rand(Cin, generator=gen) + 0.5)
    beta = torch.randn(Cin, generator=gen) * 0.3
    mean = torch.randn(Cin, generator=gen) * 0.2 + 0.2
    invstd = torch.rand(Cin, generator=gen) * 0.5 + 0.4
    sc = gamma * invstd
    act = bf16_round(torch.relu(x * sc.view(1, -1, 1, 1) + (beta - mean * sc).view(1, -1, 1, 1)))
    w = torch.zeros(Cout, Cin, 1, 1, dtype=torch.float64, requires_grad=True)
    F.conv2d(act.double(), w).backward(g.double())
    ref = w.grad
    xm = to_mat(x, ld=ops.ceil_to(Cin, 8))
    gm = to_mat(g)
    fwd, _, off = ops.conv_taps(1, 0)
    plan = ops.plan_conv_wgrad(xm.view(0, Cin), [gm.view(0, Cout)], fwd, Cin, Cout)
    dw = torch.zeros(plan["rows"] * plan["ld"], dtype=torch.float32, device="cuda")
    dev = [t.cuda() for t in (gamma, beta, mean, invstd)]
    for kw in plan["launches"]:
        d = ops.make_wgrad(W=W, H=H, B=B, dw=dw, ld=plan["ld"], **kw)
        d.pro_enable = 1
        d.pro_gamma, d.pro_beta, d.pro_mean, d.pro_invstd = (t.data_ptr() for t in dev)
        ops.run_wgrad(d)
    grad = torch.full((Cout, Cin, 1, 1), float("nan"), dtype=torch.float32, device="cuda")
    ops.unpack_wgrad(dw, plan["dt"], plan["dm"], plan["dn"], Cin, Cout, grad, 1, [off[t] for t in plan["tap_order"]], Cin, 1)
    torch.cuda.synchronize()
    err = rel_l2(grad.cpu(), ref)
    assert err < 2e-3, "prologue wgrad relL2 %.3e" % err      # the in-kernel fma may round a few activations differently


@pytest.mark.parametrize("H,W", [(12, 20), (17, 9), (33, 40)])
def test_wgrad_3x3_bn_relu_prologue(H, W):
    """weight gradient of conv3x3(relu(bn(x)), padding 1) from the RAW x (family mode: the gradient patch is shifted, so activation
    rows outside the image must be written as zeros - partial tiles at the right / bottom edges)."""
    torch.manual_seed(H + W)
    B, Cin, Cout = 2, 128, 32
    x = bf16_round(torch.randn(B, Cin, H, W) * 1.5 + 0.2)
    g = bf16_round(torch.randn(B, Cout, H, W))
    gen = torch.Generator().manual_seed(4)
    gamma = (torch.rand(Cin, generator=gen) + 0.5)
    beta = torch.randn(Cin, generator=gen) * 0.3 + 0.2       # mostly positive shifts: relu(shift) != 0 outside the image
    mean = torch.randn(Cin, generator=gen) * 0.2 + 0.2
    invstd = torch.rand(Cin, generator=gen) * 0.5 + 0.4
    sc = gamma * invstd
    act = bf16_round(torch.relu(x * sc.view(1, -1, 1, 1) + (beta - mean * sc).view(1, -1, 1, 1)))
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(act.double(), w, padding=1).backward(g.double())
    xm = to_mat(x)
    gm = to_mat(g)
    fwd, _, off = ops.conv_taps(3, 1)
    plan = ops.plan_conv_wgrad(xm.view(0, Cin), [gm.view(0, Cout)], fwd, Cin, Cout)
    dw = torch.zeros(plan["rows"] * plan["ld"], dtype=torch.float32, device="cuda")
    dev = [t.cuda() for t in (gamma, beta, mean, invstd)]
    for kw in plan["launches"]:
        d = ops.make_wgrad(W=W, H=H, B=B, dw=dw, ld=plan["ld"], **kw)
        d.pro_enable = 1
        d.pro_gamma, d.pro_beta, d.pro_mean, d.pro_invstd = (t.data_ptr() for t in dev)
        ops.run_wgrad(d)
    grad = torch.full((Cout, Cin, 3, 3), float("nan"), dtype=torch.float32, device="cuda")
    ops.unpack_wgrad(dw, plan["dt"], plan["dm"], plan["dn"], Cin, Cout, grad, 9, [off[t] for t in plan["tap_order"]], Cin * 9, 9)
    torch.cuda.synchronize()
    err = rel_l2(grad.cpu(), w.grad)
    assert err < 2e-3, "3x3 prologue wgrad relL2 %.3e" % err
