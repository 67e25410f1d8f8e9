"""Thin host-side wrappers: torch tensors (device memory, streams) -> C-ABI calls.

Every function only enqueues kernels on torch's current CUDA stream.  Activations are pixel-major
bf16 matrices `[P = B*H*W, ld]` (`Mat`).  Nothing here computes on the CPU or falls back to torch ops.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (MAX_SRC, MAX_TAPS, STATS_SLOTS, Bn, BnApply, BnBwd, BnBwdArgs, Head, HeadBwd, Igemm, View,
                   Wgrad, check)

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
KWIDTH = 64


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, byte_offset=0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + byte_offset)


def require_device():
    lib = _lib.load()
    if not torch.cuda.is_available() or not lib.dmm_device_ok():
        raise RuntimeError("dmmfods_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")
    return lib


def ceil_to(v, m):
    return (v + m - 1) // m * m


class Mat:
    """pixel-major activation matrix [B*H*W, ld] living in `t` (2-D tensor): bf16, or fp32 in the strict (tf32) mode."""

    def __init__(self, t, B, H, W):
        assert t.dtype in (torch.bfloat16, torch.float32) and t.dim() == 2 and t.is_contiguous()
        assert t.shape[0] == B * H * W, (t.shape, B, H, W)
        self.t, self.B, self.H, self.W = t, B, H, W
        self.ld = t.shape[1]
        self.esize = t.element_size()

    @property
    def P(self):
        return self.B * self.H * self.W

    def ptr(self, c0=0):
        return C.c_void_p(self.t.data_ptr() + self.esize * c0)

    def view(self, c0=0, C_=None):
        """dmm_view_t of channels [c0, c0+C_)."""
        C_ = self.ld - c0 if C_ is None else C_
        assert c0 % 8 == 0 and c0 + C_ <= self.ld
        v = View()
        v.ptr = self.t.data_ptr() + self.esize * c0
        v.C, v.W, v.H, v.B = C_, self.W, self.H, self.B
        v.sw, v.sh, v.sb = self.ld, self.ld * self.W, self.ld * self.W * self.H
        return v

    def row_phase_view(self, py, c0=0, C_=None):
        """view of the image rows 2i+py (all columns) - even / odd input rows of a stride-2 convolution."""
        C_ = self.ld - c0 if C_ is None else C_
        v = View()
        v.ptr = self.t.data_ptr() + self.esize * (py * self.W * self.ld + c0)
        v.C, v.W, v.H, v.B = C_, self.W, (self.H - py + 1) // 2, self.B
        v.sw, v.sh, v.sb = self.ld, 2 * self.ld * self.W, self.ld * self.W * self.H
        return v

    def phase_view(self, py, px, c0=0, C_=None):
        """view of the pixels (2i+py, 2j+px) - sub-pixel phase of a stride-2 transposed conv output."""
        C_ = self.ld - c0 if C_ is None else C_
        v = View()
        v.ptr = self.t.data_ptr() + self.esize * ((py * self.W + px) * self.ld + c0)
        v.C, v.W, v.H, v.B = C_, (self.W - px + 1) // 2, (self.H - py + 1) // 2, self.B
        v.sw, v.sh, v.sb = 2 * self.ld, 2 * self.ld * self.W, self.ld * self.W * self.H
        return v


def new_mat(B, H, W, ld, device="cuda", zero=False):
    f = torch.zeros if zero else torch.empty
    return Mat(f((B * H * W, ld), dtype=torch.bfloat16, device=device), B, H, W)


class Stats:
    """double[STATS_SLOTS][2][ld] accumulator rows living inside a flat float64 tensor."""

    def __init__(self, buf, offset, ld):
        self.buf, self.offset, self.ld = buf, offset, ld

    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 8 * self.offset)

    @staticmethod
    def size(ld):
        return STATS_SLOTS * 2 * ld

    def totals(self):
        """(sum, sumsq) per channel as float64 tensors (debug / tests)."""
        v = self.buf[self.offset:self.offset + self.size(self.ld)].view(STATS_SLOTS, 2, self.ld).sum(0)
        return v[0], v[1]


def pick_tile_w(W, H, pixels):
    """tile_w in {pixels, pixels/2, ..., 8} minimising the zero-padded area (ties: the squarest tile)."""
    best = None
    tw = pixels
    while tw >= 8:
        th = pixels // tw
        area = ceil_to(W, tw) * ceil_to(H, th)
        key = (area, abs(tw - th))
        if best is None or key < best[0]:
            best = (key, tw)
        tw //= 2
    return best[1]


def pick_n_tile(N):
    cap = int(os.environ.get("DMM_NTILE_MAX", "256"))      # experiment knob
    if N <= cap:
        return ceil_to(N, 16)
    if cap < 256:
        return cap
    return 256 if N % 256 == 0 or N > 512 else 128


# ------------------------------------------------------------------------------------------------
# convolution tap tables
# ------------------------------------------------------------------------------------------------
class ConvGeom:
    """Tap tables + weight (un)packing strides for one reference convolution.

    fwd_taps / dgrad_taps: lists of (src, dy, dx); *_off: offset of the tap inside the kh*kw plane.
    Weight tensor layouts: Conv2d (Cout, Cin, K, K); ConvTranspose2d (Cin, Cout, 3, 3)."""


def conv_taps(K, pad):
    """stride-1 Conv2d: forward taps, data-gradient taps (flipped), offsets kh*K+kw."""
    fwd, dg, off = [], [], []
    for kh in range(K):
        for kw in range(K):
            fwd.append((0, kh - pad, kw - pad))
            dg.append((0, pad - kh, pad - kw))
            off.append(kh * K + kw)
    return fwd, dg, off


def convt_phase_taps(py, px):
    """ConvTranspose2d(3, stride 2, padding 1): taps feeding output phase (py, px).
    out[2i+py, 2j+px] += x[i+dy, j+dx] * w[kh, kw] with kh = py+1-2dy, kw = px+1-2dx."""
    taps, off = [], []
    for dy in ((0,) if py == 0 else (0, 1)):
        for dx in ((0,) if px == 0 else (0, 1)):
            kh, kw = py + 1 - 2 * dy, px + 1 - 2 * dx
            taps.append((0, dy, dx))
            off.append(kh * 3 + kw)
    return taps, off


def convt_dgrad_taps():
    """dx[i,j] = sum_{kh,kw} dout[2i-1+kh, 2j-1+kw] * w[kh,kw]: source = phase view (py*2+px), shifted."""
    taps, off = [], []
    for kh in range(3):
        for kw in range(3):
            py, dy = (0, 0) if kh == 1 else (1, -1 if kh == 0 else 0)
            px, dx = (0, 0) if kw == 1 else (1, -1 if kw == 0 else 0)
            taps.append((py * 2 + px, dy, dx))
            off.append(kh * 3 + kw)
    return taps, off


def convt_wgrad_taps():
    """dw[kh,kw] = sum_pix X(pix + shift) * dout_phase(pix): (ysrc, dy, dx) with the shift applied to X."""
    taps, off = [], []
    for kh in range(3):
        for kw in range(3):
            py, dy = (0, 0) if kh == 1 else (1, 1 if kh == 0 else 0)
            px, dx = (0, 0) if kw == 1 else (1, 1 if kw == 0 else 0)
            taps.append((py * 2 + px, dy, dx))
            off.append(kh * 3 + kw)
    return taps, off


# ------------------------------------------------------------------------------------------------
# descriptor builders (the returned ctypes structs can be re-launched any number of times)
# ------------------------------------------------------------------------------------------------
def make_igemm(srcs, taps, weights, ktot, n_rows, W, H, B, N, out_ptr, ldo, coff=0, out_mode=0, stats=None,
               stats_off=0, out_stride=(1, 1), out_phase=(0, 0), out_hw=None, n_tile=None, tile_w=None,
               kwidth=KWIDTH, fold_kw=0):
    d = Igemm()
    d.fold_kw = fold_kw
    assert 1 <= len(srcs) <= MAX_SRC and 1 <= len(taps) <= MAX_TAPS
    for i, v in enumerate(srcs):
        d.src[i] = v
    d.num_src = len(srcs)
    d.num_taps = len(taps)
    for i, (s, dy, dx) in enumerate(taps):
        d.tap_src[i], d.tap_dy[i], d.tap_dx[i] = s, dy, dx
    d.weights = weights.data_ptr() if isinstance(weights, torch.Tensor) else weights
    d.ktot, d.n_rows, d.kwidth = ktot, n_rows, kwidth
    d.W, d.H, d.B = W, H, B
    d.tile_w = tile_w or pick_tile_w(W, H, 128)
    d.N = N
    d.n_tile = n_tile or pick_n_tile(N)
    d.out = out_ptr.value if isinstance(out_ptr, C.c_void_p) else out_ptr
    d.out_mode, d.ldo, d.coff = out_mode, ldo, coff
    d.out_sy, d.out_sx = out_stride
    d.out_py, d.out_px = out_phase
    d.OH, d.OW = out_hw if out_hw is not None else (H, W)
    if stats is not None:
        d.stats = stats.ptr().value
        d.stats_ld, d.stats_off = stats.ld, stats_off
    return d


def fuse_bn_bwd_reduce(d, x, xc0, bn_bwd):
    """make the data-gradient launch `d` also perform the reduce pass of the BatchNorm-ReLU backward that consumes its output
    (x: Mat holding the raw BN input, channel n of the output <-> column xc0 + n; bn_bwd: the BnBwd of make_bn_bwd)."""
    d.bnb_x, d.bnb_ldx = x.ptr(xc0).value, x.ld
    d.bnb_gamma, d.bnb_beta = bn_bwd.gamma, bn_bwd.beta
    d.bnb_mean, d.bnb_invstd = bn_bwd.save_mean, bn_bwd.save_invstd
    d.bnb_sums, d.bnb_sums_ld, d.bnb_sums_off = bn_bwd.sums, bn_bwd.sums_ld, bn_bwd.sums_off
    return d


def run_igemm(d):
    check(_lib.load().dmm_conv_igemm(C.byref(d), _stream()), "dmm_conv_igemm")


def make_wgrad(a_srcs, b_srcs, a_slots, b_slots, n_tile, W, H, B, dw, ld, ya=1, yb=1, a_step=0, b_step=0, kpx=None,
               tile_w=None, splits=0):
    """dmm_wgrad_t.  a_slots / b_slots: lists of (src, dy, dx, ch0, out0)."""
    d = Wgrad()
    assert 1 <= len(a_srcs) <= MAX_SRC and 1 <= len(b_srcs) <= MAX_SRC
    assert 1 <= len(a_slots) <= _lib.WG_MAX_A and 1 <= len(b_slots) <= _lib.WG_MAX_B
    fam = (len(b_slots) > 1 and n_tile <= 64 and yb == 1 and len({(s[0], s[3]) for s in b_slots}) == 1)
    if fam and tile_w is None and os.environ.get("DMM_WGRAD_SWAP_XY", "1") != "0":
        # family launches tile the image 8 (x) by kpx / 8 (y) pixels: a 30 x 20 / 60 x 40 image wastes up to 60 % of every tile in y.
        # The sum over pixels does not care which image axis is called x: swap the roles (views with exchanged extents / strides,
        # shifts with exchanged offsets) when that pads less.
        th = 16
        plain = ceil_to(W, 8) * ceil_to(H, th)
        swp = ceil_to(H, 8) * ceil_to(W, th)
        if swp * 20 < plain * 19:
            def sv(v):
                t = View()
                t.ptr, t.C, t.B, t.sb = v.ptr, v.C, v.B, v.sb
                t.W, t.H, t.sw, t.sh = v.H, v.W, v.sh, v.sw
                return t
            a_srcs, b_srcs = [sv(v) for v in a_srcs], [sv(v) for v in b_srcs]
            a_slots = [(s_, dx, dy, c0, o0) for (s_, dy, dx, c0, o0) in a_slots]
            b_slots = [(s_, dx, dy, c0, o0) for (s_, dy, dx, c0, o0) in b_slots]
            W, H = H, W
    for i, v in enumerate(a_srcs):
        d.a_src[i] = v
    for i, v in enumerate(b_srcs):
        d.b_src[i] = v
    d.num_a_src, d.num_b_src = len(a_srcs), len(b_srcs)
    for i, (s, dy, dx, ch0, out0) in enumerate(a_slots):
        d.a[i].src, d.a[i].dy, d.a[i].dx, d.a[i].ch0, d.a[i].out0 = s, dy, dx, ch0, out0
    for i, (s, dy, dx, ch0, out0) in enumerate(b_slots):
        d.b[i].src, d.b[i].dy, d.b[i].dx, d.b[i].ch0, d.b[i].out0 = s, dy, dx, ch0, out0
    d.num_a, d.num_b = len(a_slots), len(b_slots)
    d.n_tile = n_tile
    d.ya, d.yb, d.a_step, d.b_step = ya, yb, a_step, b_step
    d.W, d.H, d.B = W, H, B
    na = (len(a_slots) + 1) // 2
    bw = 64 if n_tile >= 64 else n_tile
    # "family" (see wgrad.cu): every B group is a shifted view of one source -> one halo patch per stage, tile_w = 8
    family = (len(b_slots) > 1 and n_tile <= 64 and yb == 1 and len({(s[0], s[3]) for s in b_slots}) == 1)
    hx = max(s[2] for s in b_slots) - min(s[2] for s in b_slots)
    hy = max(s[1] for s in b_slots) - min(s[1] for s in b_slots)

    def stage_bytes(k):
        a = 2 * na * k * 128
        if family:
            return a + (k // 8 + hy) * (8 + hx) * bw * 2
        return a + len(b_slots) * ceil_to(n_tile, bw) * k * 2
    if kpx is None:      # the MMA warp pays a few hundred cycles per stage hand-over: big stages, but >= 3 (else 2) of them
        kpx = int(os.environ.get("DMM_WGRAD_KPX", "0")) or None
    if kpx is None:
        # one tensor load costs the producer thread a few hundred cycles whatever its size: the largest k-block that still
        # double-buffers wins (measured: K=512 conv_reduce wgrad 0.41 ms at kpx 32 x 5 stages, 0.25 ms at kpx 64 x 2 stages)
        for k in (128, 64, 32):
            if stage_bytes(k) * 2 <= 200 * 1024:
                kpx = k
                break
        kpx = kpx or 32
    d.kpx = kpx
    d.tile_w = tile_w or (8 if family else pick_tile_w(W, H, kpx))
    d.splits = splits
    d.dw = dw.data_ptr() if isinstance(dw, torch.Tensor) else dw
    d.ld = ld
    return d


def run_wgrad(d):
    check(_lib.load().dmm_conv_wgrad(C.byref(d), _stream()), "dmm_conv_wgrad")


def plan_conv_wgrad(x, ys, taps, M, N):
    """Launch plan for  dw[t][m][n] = sum_pix X(pix + tap_t)[m] * Y[ysrc_t](pix)[n]   (t < T, m < M, n < N).

    x: View with M channels; ys: list of Views with N channels; taps: list of (ysrc, dy, dx).
    Returns dict(launches=[kwargs for make_wgrad without dw], rows, ld, dt, dm, dn, tap_order): the scratch matrix has
    `rows` x `ld` fp32 entries and element (t', m, n) lives at dw[t'*dt + m*dm + n*dn], where t' is the position of tap
    tap_order[t'] of the caller's list (the plan may reorder taps so that horizontally adjacent ones can share one wide
    MMA; permute any per-tap table, e.g. the unpack offsets, with tap_order)."""
    T = len(taps)
    mch = (M + 63) // 64                     # 64-channel chunks of X
    if N <= 64:
        nt = 16 if N <= 16 else (32 if N <= 32 else 64)
        if T * nt <= 512:
            # A = activation (loaded once), B = one group per tap: the output gradient shifted by -tap
            na = max(1, min(4, 512 // (T * nt), (mch + 1) // 2))
            n_slots = min(2 * na, mch)
            ya = (mch + 2 * na - 1) // (2 * na)
            ld = T * nt
            a_slots = [(0, 0, 0, 64 * i, 64 * i) for i in range(n_slots)]
            # B shift = -tap; ascending (source, row, x) order puts the taps of one kernel row next to each other
            order = sorted(range(T), key=lambda t: (taps[t][0], -taps[t][1], -taps[t][2]))
            b_slots = [(taps[t][0], -taps[t][1], -taps[t][2], 0, i * nt) for i, t in enumerate(order)]
            launch = dict(a_srcs=[x], b_srcs=list(ys), a_slots=a_slots, b_slots=b_slots, n_tile=nt, ya=ya, yb=1,
                          a_step=128 * na, b_step=0)
            return dict(launches=[launch], rows=mch * 64, ld=ld, dt=nt, dm=ld, dn=1, tap_order=order)
        kw_max = max(sum(1 for t in taps if (t[0], t[1]) == r) for r in {(t[0], t[1]) for t in taps})
        if kw_max * nt <= 512:
            # same, but too many taps for the 512 TMEM columns: one launch per group of kernel rows
            na = max(1, min(4, 512 // (kw_max * nt), (mch + 1) // 2))
            n_slots = min(2 * na, mch)
            ya = (mch + 2 * na - 1) // (2 * na)
            order = sorted(range(T), key=lambda t: (taps[t][0], -taps[t][1], -taps[t][2]))
            a_slots = [(0, 0, 0, 64 * i, 64 * i) for i in range(n_slots)]
            launches, cur, cur_rows = [], [], []
            for i, t in enumerate(order):
                row = (taps[t][0], taps[t][1])
                if row not in cur_rows and na * (len(cur) + kw_max) * nt > 512 and cur:
                    launches.append(cur)
                    cur, cur_rows = [], []
                if row not in cur_rows:
                    cur_rows.append(row)
                cur.append((taps[t][0], -taps[t][1], -taps[t][2], 0, i * nt))
            launches.append(cur)
            launches = [dict(a_srcs=[x], b_srcs=list(ys), a_slots=a_slots, b_slots=bs, n_tile=nt, ya=ya, yb=1, a_step=128 * na,
                             b_step=0) for bs in launches]
            return dict(launches=launches, rows=mch * 64, ld=T * nt, dt=nt, dm=T * nt, dn=1, tap_order=order)
        if M <= 256:
            # roles swapped: A = shifted output-gradient chunks (rows = (tap, n)), B = the activation (columns = m)
            nt = 16 if M <= 16 else (32 if M <= 32 else (64 if M <= 64 else ceil_to(M, 16)))
            nch = (N + 63) // 64
            npad = nch * 64
            max_chunks = min(_lib.WG_MAX_A, 2 * (512 // nt))
            taps_per = max(1, max_chunks // nch)
            launches = []
            for t0 in range(0, T, taps_per):
                a_slots = []
                for t in range(t0, min(T, t0 + taps_per)):
                    ys_i, dy, dx = taps[t]
                    for j in range(nch):
                        a_slots.append((ys_i, -dy, -dx, 64 * j, t * npad + 64 * j))
                launches.append(dict(a_srcs=list(ys), b_srcs=[x], a_slots=a_slots, b_slots=[(0, 0, 0, 0, 0)], n_tile=nt,
                                     ya=1, yb=1, a_step=0, b_step=0))
            return dict(launches=launches, rows=T * npad, ld=nt, dt=npad * nt, dm=1, dn=nt, tap_order=list(range(T)))
    if len(ys) > 1 and N <= 128 and mch <= 2 and os.environ.get("DMM_WGRAD_GROUP_TAPS", "1") != "0":
        # several output-gradient sources (the four sub-pixel phases of a ConvTranspose) and one m-tile per tap: ONE launch per
        # source, A = the activation at the shifts of all taps that read this source (up to 4 taps = 4 accumulators of 128 columns),
        # B = the source once.  A launch per tap (below) streams every phase once per tap and the activation nine times:
        # 3.6x the algorithmic DRAM traffic on Transposed_Convolution_4 (ncu), and these launches are bound by the bytes the SM
        # takes in.  dw rows = (tap, input channel), columns = output channel.
        mpad, npad = mch * 64, ceil_to(N, 128)
        launches = []
        for y in sorted({t[0] for t in taps}):
            group = [t for t in range(T) if taps[t][0] == y]
            per = max(1, _lib.WG_MAX_A // mch)
            for g0 in range(0, len(group), per):
                a_slots = [(0, taps[t][1], taps[t][2], 64 * i, t * mpad + 64 * i) for t in group[g0:g0 + per] for i in range(mch)]
                launches.append(dict(a_srcs=[x], b_srcs=list(ys), a_slots=a_slots, b_slots=[(y, 0, 0, 0, 0)], n_tile=128, ya=1, yb=1,
                                     a_step=0, b_step=0))
        return dict(launches=launches, rows=T * mpad, ld=npad, dt=mpad * npad, dm=npad, dn=1, tap_order=list(range(T)))
    # general: one launch per tap, A = activation shifted by +tap, B = output gradient in 128-channel groups
    nb = 1 if N <= 128 else 2
    na = max(1, min(4 // nb, (mch + 1) // 2))
    # m-tiles per CTA.  Few pixels (dense blocks 2-4, the deep decoder stages): ONE m-tile per CTA and more replicas along the input
    # channels instead - every CTA then runs a longer K loop and the fp32 reduce-add epilogue (the fixed cost of these launches)
    # shrinks with the accumulator (measured back to back, 19 200 px x 1024 channels: 24 -> 15 us; 76 800 px x 640: 39 -> 27 us)
    P = int(x.W) * int(x.H) * int(x.B)
    na_cap = int(os.environ.get("DMM_WGRAD_NA", "0")) or (1 if P <= int(os.environ.get("DMM_WGRAD_NA1_PIXELS", "400000")) else 4)
    na = max(1, min(na, na_cap))
    n_slots = min(2 * na, mch)
    npad = ceil_to(N, 128)
    ya = (mch + 2 * na - 1) // (2 * na)
    yb = (N + 128 * nb - 1) // (128 * nb)
    ld = T * npad
    launches = []
    for t, (ys_i, dy, dx) in enumerate(taps):
        a_slots = [(0, dy, dx, 64 * i, 64 * i) for i in range(n_slots)]
        b_slots = [(ys_i, 0, 0, 128 * g, t * npad + 128 * g) for g in range(nb)]
        launches.append(dict(a_srcs=[x], b_srcs=list(ys), a_slots=a_slots, b_slots=b_slots, n_tile=128, ya=ya, yb=yb,
                             a_step=128 * na, b_step=128 * nb))
    return dict(launches=launches, rows=mch * 64, ld=ld, dt=npad, dm=ld, dn=1, tap_order=list(range(T)))


def _i32arr(vals):
    return (C.c_int32 * len(vals))(*vals)


def pack_weights(w, dst, n_valid, n_rows, C_, T, tap_off, sn, sc, kwidth=KWIDTH):
    check(_lib.load().dmm_pack_weights(_ptr(w), _ptr(dst), n_valid, n_rows, C_, kwidth, T, _i32arr(tap_off), sn, sc,
                                       _stream()), "dmm_pack_weights")


def unpack_wgrad(dw, dt, dm, dn, M, N, grad, T, tap_off, sn, sc, accumulate=False):
    check(_lib.load().dmm_unpack_wgrad(_ptr(dw), dt, dm, dn, M, N, _ptr(grad), T, _i32arr(tap_off), sn, sc,
                                       1 if accumulate else 0, _stream()), "dmm_unpack_wgrad")


def make_bn(stats, stats_off, count, gamma, beta, running_mean=None, running_var=None, save_mean=None,
            save_invstd=None, training=True, rep=1.0, c0=0, eps=BN_EPS, momentum=BN_MOMENTUM):
    """dmm_bn_t for channels [c0, c0+C) of one BatchNorm2d; tensors are the module's full fp32 vectors."""
    b = Bn()
    if stats is not None:
        b.stats = stats.ptr().value
        b.stats_ld, b.stats_off = stats.ld, stats_off
    b.count, b.rep = float(count), float(rep)
    o = 4 * c0
    b.gamma = gamma.data_ptr() + o
    b.beta = beta.data_ptr() + o
    if running_mean is not None:
        b.running_mean = running_mean.data_ptr() + o
        b.running_var = running_var.data_ptr() + o
    if save_mean is not None:
        b.save_mean = save_mean.data_ptr() + o
        b.save_invstd = save_invstd.data_ptr() + o
    b.eps, b.momentum, b.training = eps, momentum, 1 if training else 0
    return b


def make_bn_apply(x, c0, C_, bn, y, yc0, pool=0, ystats=None, ystats_off=0):
    """x, y: Mat.  Normalises channels [c0, c0+C_) of x into channels [yc0, ...) of y."""
    d = BnApply()
    d.x, d.ldx = x.ptr(c0).value, x.ld
    d.B, d.H, d.W, d.C = x.B, x.H, x.W, C_
    d.bn = bn
    d.pool = pool
    d.y, d.ldy = y.ptr(yc0).value, y.ld
    if ystats is not None:
        d.ystats = ystats.ptr().value
        d.ystats_ld, d.ystats_off = ystats.ld, ystats_off
    return d


def run_bn_apply(d):
    check(_lib.load().dmm_bn_relu_apply(C.byref(d), _stream()), "dmm_bn_relu_apply")


def make_bn_bwd(sums, sums_off, count, gamma, beta, save_mean, save_invstd, dgamma=None, dbeta=None, c0=0):
    b = BnBwd()
    b.sums = sums.ptr().value
    b.sums_ld, b.sums_off = sums.ld, sums_off
    b.count = float(count)
    o = 4 * c0
    b.gamma = gamma.data_ptr() + o
    b.beta = beta.data_ptr() + o
    b.save_mean = save_mean.data_ptr() + o
    b.save_invstd = save_invstd.data_ptr() + o
    if dgamma is not None:
        b.dgamma = dgamma.data_ptr() + o
        b.dbeta = dbeta.data_ptr() + o
    return b


def make_bn_bwd_args(x, c0, C_, g_ptr, ldg, bn, out_ptr, ldo, out_mode, gmode=0, g_is_f32=False):
    d = BnBwdArgs()
    d.x, d.ldx = x.ptr(c0).value, x.ld
    d.g = g_ptr.value if isinstance(g_ptr, C.c_void_p) else g_ptr
    d.ldg = ldg
    d.g_is_f32 = 1 if g_is_f32 else 0
    d.gmode = gmode
    d.B, d.H, d.W, d.C = x.B, x.H, x.W, C_
    d.bn = bn
    d.out = (out_ptr.value if isinstance(out_ptr, C.c_void_p) else out_ptr) if out_ptr is not None else None
    d.ldo, d.out_mode = ldo, out_mode
    return d


def run_bn_bwd(d):
    lib = _lib.load()
    check(lib.dmm_bn_relu_bwd_reduce(C.byref(d), _stream()), "dmm_bn_relu_bwd_reduce")
    check(lib.dmm_bn_relu_bwd_apply(C.byref(d), _stream()), "dmm_bn_relu_bwd_apply")


def im2col_7x7s2(x1, x2, out):
    """x1 (B,C1,H,W) fp32, x2 optional; out: Mat at half resolution with ld = kpad."""
    B, C1, H, W = x1.shape
    C2 = 0 if x2 is None else x2.shape[1]
    check(_lib.load().dmm_im2col_7x7s2(_ptr(x1), C1, _ptr(x2), C2, B, H, W, out.ptr(), out.ld, _stream()),
          "dmm_im2col_7x7s2")


def nchw_stats(x, stats, off):
    B, C_, H, W = x.shape
    check(_lib.load().dmm_nchw_stats(_ptr(x), B, C_, H * W, stats.ptr(), stats.ld, off, _stream()), "dmm_nchw_stats")


def nchw_to_nhwc_bf16(x, out):
    B, C_, H, W = x.shape
    check(_lib.load().dmm_nchw_to_nhwc_bf16(_ptr(x), B, C_, H, W, out.ptr(), out.ld, _stream()), "dmm_nchw_to_nhwc_bf16")


def rows_f32_to_bf16(src, c0, C_, dst):
    """src: fp32 [P, lds] tensor, channels [c0, c0+C_) -> dst Mat columns [0, C_)."""
    check(_lib.load().dmm_rows_f32_to_bf16(C.c_void_p(src.data_ptr() + 4 * c0), src.shape[1], dst.ptr(), dst.ld,
                                           src.shape[0], C_, _stream()), "dmm_rows_f32_to_bf16")


def bce_logits(logits, target, loss=None, grad=None, class_sums=None):
    B, C_, H, W = logits.shape
    assert logits.is_contiguous() and target.is_contiguous() and target.shape == logits.shape
    check(_lib.load().dmm_bce_logits(_ptr(logits), _ptr(target), logits.numel(), C_, H * W, _ptr(loss), _ptr(grad),
                                     _ptr(class_sums), _stream()), "dmm_bce_logits")


def adam_flat(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step):
    check(_lib.load().dmm_adam_flat(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), lr, beta1,
                                    beta2, eps, weight_decay, step, _stream()), "dmm_adam_flat")
