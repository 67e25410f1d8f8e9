"""Execution engine of the Dense-U-Net hot path: turns the reference graph
(Dense_U_Net_lidar.forward, dmmfods/graphs/models/Dense_U_Net_lidar.py:210-267, with torchvision's
_DenseLayer/_DenseBlock/_Transition, tv:31-133) into two static programs of C-ABI kernel launches
(forward, backward) over preallocated HBM buffers.

Data layout in HBM (all per engine = per (B, H, W)):
  * dense-block buffers: one bf16 matrix [P_b, C_total_b] per block; every layer's conv2 writes its
    `growth_rate` channels in place at its channel offset (no torch.cat); the same buffer is the U-Net skip.
  * per dense layer: a1 = relu(norm1(concat)) [P, C_i], z1 = conv1 output [P, 4k], a2 = relu(norm2(z1))
    [P, 4k] - kept for backward (11 + 6 + 6 GB at BASELINE config 3; 180 GB HBM makes recompute unnecessary).
  * batch statistics: double[8 slots][2][C] rows accumulated by the producing kernel's epilogue; a raw
    channel's statistics are computed ONCE and shared by every later BatchNorm that normalises it.
  * gradients w.r.t. block buffers: fp32 [P_b, C_total_b] (accumulated over all consuming layers);
    every other activation gradient is a bf16 matrix; weight gradients: fp32 split-K accumulators that
    are scattered into a flat fp32 gradient buffer in parameter layout.
All launches go to torch's current stream, there is no host synchronisation and no allocation after
construction, so a whole step can be captured in a CUDA graph.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib, ops
from ._lib import Head, HeadBwd
from .ops import Mat, Stats, ceil_to


# fuse the reduce pass of a BatchNorm-ReLU backward into the epilogue of the data-gradient convolution that produces its input
FUSE_BN_BWD_REDUCE = os.environ.get("DMM_FUSE_BN_BWD_REDUCE", "1") != "0"
# run the weight-gradient kernels on a second stream beside the data-gradient / BatchNorm-backward chain
WGRAD_SIDE_STREAM = os.environ.get("DMM_WGRAD_SIDE_STREAM", "1") != "0"
# dense-layer norm1 + relu1 as a PROLOGUE of conv1 (and of its weight gradient): relu(bn(x)) is applied to the operand tiles in
# shared memory, the activated tensor is never written to HBM (tv:47-50, north-star "BN-ReLU prologues fused into the conv loads")
FUSE_BN_PROLOGUE = os.environ.get("DMM_FUSE_BN_PROLOGUE", "1") != "0"
# the same for norm2 + relu2 in front of the 3x3 growth convolution (halo patches activated in shared memory, out-of-image
# pixels kept at zero) and its weight gradient.  OFF by default: measured +0.85 ms per step (the 3x3 kernels are shared-memory
# bound, the extra in-place pass costs conv2 fprop +1.5 ms and its weight gradient +1.75 ms against 2.4 ms of saved BN-ReLU)
FUSE_BN_PROLOGUE_KXK = FUSE_BN_PROLOGUE and os.environ.get("DMM_FUSE_BN_PROLOGUE_KXK", "0") != "0"


# data gradients of the 3x3 growth convolutions (32 gradient channels): two taps share one 64-wide K block of the packed weights
PACK32 = os.environ.get("DMM_DGRAD_PACK32", "1") != "0"
# the reduce pass of norm2's backward inside the growth convolutions' data gradient (igemm epilogue) or as its own launch
# (measured r02: the separate launch wins, 82.2 vs 83.3 ms per step - the growth data gradient is epilogue-bound, and a chunk
# with the fused statistics costs 3-4x a plain one)
CONV2_DGRAD_FUSED = os.environ.get("DMM_CONV2_DGRAD_FUSED", "0") != "0"
FUSE_MIN_PIXELS = int(os.environ.get("DMM_FUSE_MIN_PIXELS", "1000000"))
DA1_ALIGN = int(os.environ.get("DMM_DA1_ALIGN", "8"))       # 64 (128-byte aligned rows) measured = (88.3 vs 87.9 ms): dense pitch kept
# inference (eval-mode engines): fold every BatchNorm whose only producer is one convolution into that convolution - scale into
# the packed weight rows, shift + ReLU into the igemm epilogue (Agent.py:337-352 validation / notebook inference)
FOLD_EVAL_BN = os.environ.get("DMM_FOLD_EVAL_BN", "1") != "0"
# dense-layer norm1 backward: dmm_bn_bwd_finalize fused into the contribution launch (76 launches fewer per step on DenseNet-121)
FUSE_FINALIZE = os.environ.get("DMM_FUSE_FINALIZE", "1") != "0"
# weight pack / gradient unpack as load-balanced (job, chunk) launches (DMM_BALANCED_PACK=0: 32 blocks per job as in round 1)
BALANCED_PACK = os.environ.get("DMM_BALANCED_PACK", "1") != "0"
WORK_CHUNK = 8192


def _work_table(sizes, dev):
    """int32 [n, 2] device table of (job, chunk) pairs covering `sizes[job]` elements in chunks of WORK_CHUNK."""
    rows = [(j, c) for j, n in enumerate(sizes) for c in range((int(n) + WORK_CHUNK - 1) // WORK_CHUNK)]
    t = torch.tensor(rows, dtype=torch.int32).view(-1, 2) if rows else torch.zeros((0, 2), dtype=torch.int32)
    return t.to(dev)


class Op:
    """one C-ABI launch.  kind / flops / bytes: kernel family and ALGORITHMIC work of the launch (each distinct
    input element read once + each output written once; 2*M*N*K of the reference operator) for roofline reports."""
    __slots__ = ("fn", "arg", "gbuf", "name", "kind", "flops", "bytes", "heavy")

    def __init__(self, fn, arg, name, gbuf=None, kind="misc", flops=0.0, nbytes=0.0):
        self.fn, self.arg, self.name, self.gbuf = fn, arg, name, gbuf
        self.kind, self.flops, self.bytes = kind, float(flops), float(nbytes)
        self.heavy = False


class _Arena:
    """bump allocator over one torch tensor (zeroed in one memset per step)."""

    def __init__(self, numel, dtype, device):
        self.buf = torch.zeros(numel, dtype=dtype, device=device)
        self.used = 0

    def take(self, n, align=64):
        off = ceil_to(self.used, align)
        if off + n > self.buf.numel():
            raise RuntimeError("dmmfods_b200: arena exhausted (%d + %d > %d)" % (off, n, self.buf.numel()))
        self.used = off + n
        return off

    def zero_used(self):
        self.buf[:self.used].zero_()


class _BNInfo:
    def __init__(self, eng, prefix, C_):
        self.prefix, self.C = prefix, C_
        self.gamma = eng.p[prefix + ".weight"]
        self.beta = eng.p[prefix + ".bias"]
        self.rm = eng.p[prefix + ".running_mean"]
        self.rv = eng.p[prefix + ".running_var"]
        off = eng._save.take(2 * C_, 4)
        self.save_mean = eng._save.buf[off:off + C_]
        self.save_invstd = eng._save.buf[off + C_:off + 2 * C_]
        self.dgamma = eng.grad[prefix + ".weight"]
        self.dbeta = eng.grad[prefix + ".bias"]


class Engine:
    def __init__(self, params, model_cfg, B, H, W, training=True, need_backward=True, plan_only=False,
                 bucket_bytes=32 << 20, _order_only=False, precision="bf16"):
        """params: dict name -> CUDA tensor with the reference's state_dict keys (parameters fp32, BN buffers);
        model_cfg: mapping with the keys of helper:111-123.  plan_only=True builds the launch programs without a
        GPU (host-logic tests); such an engine cannot run.
        precision: "bf16" (production: bf16 storage, kind::f16 MMAs), or a STRICT forward mode with fp32 storage of every
        activation and packed weight, fp32 accumulation, unfused plan (forward + loss only): "tf32" = one kind::tf32 MMA per
        product (operands truncated to 10 mantissa bits), "tf32x3" = 3xTF32 error-compensated products (fp32-grade)."""
        self.lib = _lib.load() if plan_only else ops.require_device()
        self.plan_only = plan_only
        if precision not in ("bf16", "tf32", "tf32x3"):
            raise ValueError("precision must be 'bf16', 'tf32' or 'tf32x3'")
        self.f32 = precision in ("tf32", "tf32x3")
        self.split3 = precision == "tf32x3"
        self.precision = precision
        if self.f32 and need_backward and training:
            raise NotImplementedError("dmmfods_b200: the strict tf32 mode covers forward + loss (need_backward=False)")
        self.kwidth = 32 if self.f32 else ops.KWIDTH
        self.adt = torch.float32 if self.f32 else torch.bfloat16
        self.p = params
        self.training = training
        self.need_backward = need_backward and training
        any_p = next(iter(params.values()))
        self.dev = any_p.device
        self.B, self.H, self.W = B, H, W
        m = model_cfg
        self.k = int(m["growth_rate"])
        self.block_config = tuple(int(v) for v in m["block_config"])
        self.nif = int(m["num_init_features"])
        self.bnk = int(m["bn_size"]) * self.k
        self.c1 = int(m["stream_1_in_channels"])
        self.c2 = int(m["stream_2_in_channels"])
        self.cb = int(m["concat_before_block_num"])
        self.ncls = int(m["num_classes"])
        nb = len(self.block_config)
        if self.cb == 1 and self.c2 == 0:
            self.fusion = "no"
        elif self.cb == 1 and self.c2 > 0:
            self.fusion = "early"
        elif 1 < self.cb <= nb:
            self.fusion = "mid"
        else:
            raise AttributeError("invalid fusion configuration")
        if m.get("drop_rate", 0):
            raise NotImplementedError("drop_rate > 0 is not supported by the CUDA path (reference default is 0)")
        for v in (self.k, self.nif, self.bnk):
            if v % 8:
                raise ValueError("channel counts must be multiples of 8 for the bf16 pixel-major layout (got %d)" % v)
        if H % 2 or W % 2:
            raise ValueError("input height/width must be even (the reference's final torch.cat fails otherwise)")

        # ---- flat gradient buffer (fp32): parameters ordered by the backward stage that finalises their
        # gradient, so that gradient buckets (all-reduce overlap) are contiguous ranges --------------------
        float_names = [k for k, v in params.items() if v.is_floating_point() and "running_" not in k]
        self.nbt_keys = [k for k in params if k.endswith("num_batches_tracked")]
        self.bucket_bytes = int(bucket_bytes)
        self._dry = True
        self._alloc_dev = torch.device("meta")
        self.gflat = None
        self.grad = {k: params[k] for k in float_names}       # placeholders for the dry pass (addresses unused)
        self._reset_plan_state()
        self._plan()
        order, seen = [], set()
        for st in reversed(self._bwd_stages):
            for n in self._stage_params.get(id(st), []):
                if n not in seen:
                    seen.add(n)
                    order.append(n)
        if not self.need_backward:
            order = list(float_names)
        missing = [k for k in float_names if k not in set(order)]
        if missing:
            raise RuntimeError("dmmfods_b200: no backward stage produces the gradient of %s" % missing[:4])
        self.param_names = order
        if _order_only:          # gradient_order(): the dry (meta-device) planning pass is all that is needed
            return
        total = sum(params[k].numel() for k in order)
        self.gflat = torch.zeros(total, dtype=torch.float32, device=self.dev)
        self.grad, self.grad_offset = {}, {}
        off = 0
        for k in order:
            n = params[k].numel()
            self.grad[k] = self.gflat[off:off + n].view(params[k].shape)
            self.grad_offset[k] = off
            off += n

        self._dry = False
        self._alloc_dev = self.dev
        self._reset_plan_state()
        self.in1 = torch.zeros(B, self.c1, H, W, dtype=torch.float32, device=self.dev)
        self.in2 = torch.zeros(B, max(self.c2, 1), H, W, dtype=torch.float32, device=self.dev)
        self.logits = torch.zeros(B, self.ncls, H, W, dtype=torch.float32, device=self.dev)
        self.dlogits = torch.zeros(B, self.ncls, H, W, dtype=torch.float32, device=self.dev)
        self.class_sums = torch.zeros(self.ncls, dtype=torch.float64, device=self.dev)
        self._plan()
        self._finalize()

    @staticmethod
    def gradient_order(params, model_cfg, B, H, W):
        """parameter names in the order of the engine's flat gradient buffer (= backward completion order)."""
        e = Engine.__new__(Engine)
        Engine.__init__(e, params, model_cfg, B, H, W, plan_only=True, _order_only=True)
        return list(e.param_names)

    def _reset_plan_state(self):
        if self._dry:      # meta arenas: only offsets matter
            self._stats = _Arena(1 << 40, torch.float64, "meta")
            self._sums = _Arena(1 << 40, torch.float64, "meta")
            self._save = _Arena(1 << 34, torch.float32, "meta")
        else:
            self._stats = _Arena(4 << 20, torch.float64, self.dev)
            self._sums = _Arena(4 << 20, torch.float64, self.dev)
            self._save = _Arena(1 << 20, torch.float32, self.dev)
        self._dw = None          # sized after planning
        self._wpk = None
        self._dw_req = []
        self._wpk_req = []
        self._pack_jobs = []
        self._fin_req = []       # contribution launches with the fused finalize: ticket counters assigned in _finalize
        self._fold_jobs = []     # eval-mode BatchNorms folded into their producing convolution: (bn, c0, C, scale, shift)
        self._unpack_jobs = []
        self._stage_params = {}
        self._tmp = {}
        self._keep = []
        self.named = {}          # name -> Mat of the main raw activations (debugging / per-stage parity tests)
        self.fwd = []
        self._bwd_stages = []
        self.mem_bytes = 0
        if self._dry:
            z = torch.empty(1, device="meta")
            self.in1 = self.in2 = self.logits = self.dlogits = self.class_sums = z

    # ------------------------------------------------------------------------------------------------
    # small allocation helpers
    # ------------------------------------------------------------------------------------------------
    def _mat(self, B, H, W, ld):
        self.mem_bytes += B * H * W * ld * (4 if self.f32 else 2)
        m = Mat(torch.empty((B * H * W, ld), dtype=self.adt, device=self._alloc_dev), B, H, W)
        self._keep.append(m)      # launch descriptors hold raw addresses only
        return m

    def _tmpmat(self, tag, B, H, W, ld):
        key = (tag, B, H, W, ld)
        if key not in self._tmp:
            self._tmp[key] = self._mat(B, H, W, ld)
        return self._tmp[key]

    def _f32(self, rows, ld):
        self.mem_bytes += rows * ld * 4
        t = torch.empty((rows, ld), dtype=torch.float32, device=self._alloc_dev)
        self._keep.append(t)
        return t

    def _new_stats(self, ld):
        return Stats(self._stats.buf, self._stats.take(Stats.size(ld)), ld)

    def _new_sums(self, ld):
        return Stats(self._sums.buf, self._sums.take(Stats.size(ld)), ld)

    # deferred arenas: record requests, resolve pointers in _finalize()
    def _req_wpk(self, n_rows, ktot):
        self._wpk_req.append(n_rows * ktot * (2 if getattr(self, "split3", False) else 1))
        return len(self._wpk_req) - 1

    def _req_dw(self, numel):
        self._dw_req.append(numel)
        return len(self._dw_req) - 1

    # ------------------------------------------------------------------------------------------------
    # op emitters
    # ------------------------------------------------------------------------------------------------
    def _emit(self, lst, fn, arg, name, gbuf=None, kind="misc", flops=0.0, nbytes=0.0):
        lst.append(Op(fn, arg, name, gbuf, kind, flops, nbytes))

    def _conv_fwd(self, lst, name, wname, srcs, taps, tap_off, Cin, Cout, sn, sc, W, H, B, out, coff, stats, stats_off,
                  out_stride=(1, 1), out_phase=(0, 0), out_hw=None, out_mode=0, out_ptr=None, fold_kw=0, tile_w=None, cdiv=0, sc2=0,
                  fold=None):
        """emit pack job + igemm launch for a forward convolution (or one ConvTranspose phase).
        fold_kw (out_mode 2): `taps` / `tap_off` are the kernel ROWS; the kernel columns are folded into the GEMM's N (packed
        weight row kw*Cout + n)."""
        T = len(taps)
        Kp = ceil_to(Cin, self.kwidth)
        N = Cout * fold_kw if fold_kw else Cout
        n_tile = ops.pick_n_tile(N)
        n_rows = ceil_to(N, n_tile)
        if not self.training:
            stats = None
        assert not (self.f32 and (fold_kw or out_mode >= 2)), "strict mode: unfolded convolutions only"
        wid = self._req_wpk(n_rows, T * Kp)
        job = dict(w=self.p[wname], wid=wid, n_valid=N, n_rows=n_rows, C=Cin, T=T, tap_off=tap_off, sn=sn, sc=sc, cdiv=cdiv, sc2=sc2,
                   kwidth=self.kwidth)
        if fold_kw:
            job.update(ndiv=Cout, sn=1, sn2=sn)        # row kw*Cout + n <- w[n, :, kh, kw]
        self._pack_jobs.append(job)
        d = ops.make_igemm(srcs, taps, 0, T * Kp, n_rows, W, H, B, N,
                           out_ptr if out_ptr is not None else out.ptr(), 0 if out is None else out.ld, coff=coff,
                           out_mode=out_mode, stats=stats, stats_off=stats_off, out_stride=out_stride,
                           out_phase=out_phase, out_hw=out_hw, n_tile=n_tile, fold_kw=fold_kw, tile_w=tile_w, kwidth=self.kwidth)
        if self.f32:
            d.dtype = 2 if self.split3 else 1
        if fold is not None:
            # (scale, shift) of an eval-mode BatchNorm over this convolution's output: relu(bn(conv(x))) in ONE launch
            assert out_mode == 0 and not fold_kw
            job["rscale"] = fold[0]
            d.epi_bias, d.epi_relu = fold[1].data_ptr(), 1
        P = B * H * W
        osz = 4 if out_mode in (1, 2) else 2
        kk = fold_kw if fold_kw else 1
        self._emit(lst, self.lib.dmm_conv_igemm, d, name, kind="igemm_fprop", flops=2.0 * P * Cout * Cin * T * kk,
                   nbytes=P * (Cin * 2 + Cout * osz) + Cout * Cin * T * kk * 2)
        self._fix_w.append((d, wid))
        return d

    def _conv_dgrad(self, lst, name, wname, srcs, taps, tap_off, Cg, Cin, sn, sc, W, H, B, out, cdiv=0, sc2=0, heavy=None):
        """data gradient: igemm over the output-gradient views `srcs` (Cg channels) -> out [P, >=Cin]."""
        T = len(taps)
        # a narrow (<= 16 channels) single gradient source: four taps share one 64-wide K block of the packed weights
        kwidth = 16 if (Cg <= 16 and len(srcs) == 1 and T > 1) else ops.KWIDTH
        if 16 < Cg <= 32 and len(srcs) == 1 and T > 1 and PACK32:
            kwidth = 32          # growth-rate-wide gradient sources: two taps per 64-wide K block (half the weight bytes per tile)
        Kp = ceil_to(Cg, kwidth)
        n_tile = ops.pick_n_tile(Cin)
        n_rows = ceil_to(Cin, n_tile)
        wid = self._req_wpk(n_rows, T * Kp)
        self._pack_jobs.append(dict(w=self.p[wname], wid=wid, n_valid=Cin, n_rows=n_rows, C=Cg, T=T, tap_off=tap_off,
                                    sn=sn, sc=sc, kwidth=kwidth, cdiv=cdiv, sc2=sc2))
        d = ops.make_igemm(srcs, taps, 0, T * Kp, n_rows, W, H, B, Cin, out.ptr(), out.ld, n_tile=n_tile, kwidth=kwidth)
        P = B * H * W
        self._emit(lst, self.lib.dmm_conv_igemm, d, name, kind="igemm_dgrad", flops=2.0 * P * Cin * Cg * T,
                   nbytes=P * (Cg * 2 * (4 if len(srcs) == 4 else 1) + Cin * 2) + Cg * Cin * T * 2)
        self._fix_w.append((d, wid))
        # "heavy": enough MMA work per output chunk that a longer epilogue stays hidden (KxK data gradients); the 1x1 data
        # gradients of the dense layers are epilogue/HBM-bound and lose more than the separate reduce pass costs
        # ... and few-pixel launches (deep decoder stages) lose more ring depth to the two x-tile buffers per team than the separate
        # reduce pass over P x Cin elements costs (ncu r02: three such launches, 1.57 ms for 1.6 GB)
        lst[-1].heavy = (T * ceil_to(Cg, 64) >= 512 and P >= FUSE_MIN_PIXELS) if heavy is None else heavy
        return lst[-1]

    def _conv_wgrad(self, lst, name, wname, x, ys, taps, tap_off, M, N, Mvalid, Nvalid, sn, sc, W, H, B, pro=None, n_off=0, ndiv=0, sn2=0):
        """weight gradient: wgrad launches into a zeroed fp32 scratch matrix + unpack job into the flat gradient
        buffer.  x: View with M channels, ys: Views with N channels, taps: (ysrc, dy, dx) applied to x.
        pro (_BNInfo): x is the RAW input of that BatchNorm; relu(bn(x)) is applied to the operand tiles on the fly."""
        T = len(taps)
        plan = ops.plan_conv_wgrad(x, ys, taps, M, N)
        tap_off = [tap_off[t] for t in plan["tap_order"]]
        did = self._req_dw(plan["rows"] * plan["ld"])
        P = B * H * W
        nl = len(plan["launches"])
        for i, kw in enumerate(plan["launches"]):
            d = ops.make_wgrad(W=W, H=H, B=B, dw=0, ld=plan["ld"], **kw)
            if pro is not None:
                d.pro_enable = 1
                d.pro_gamma, d.pro_beta = pro.gamma.data_ptr(), pro.beta.data_ptr()
                d.pro_mean, d.pro_invstd = pro.save_mean.data_ptr(), pro.save_invstd.data_ptr()
            self._emit(lst, self.lib.dmm_conv_wgrad, d, name + ("[%d]" % i if nl > 1 else ""), kind="wgrad",
                       flops=2.0 * P * Mvalid * Nvalid * T / nl, nbytes=(P * (M * 2 + N * 2 * len(ys)) + T * M * N * 4) / nl)
            self._fix_dw.append((d, did))
        self._unpack_jobs.append(dict(wname=wname, did=did, grad=self.grad[wname], dt=plan["dt"], dm=plan["dm"], dn=plan["dn"], M=Mvalid,
                                      N=Nvalid, T=T, tap_off=tap_off, sn=sn, sc=sc, stage=id(lst), dw_off=n_off * plan["dn"], ndiv=ndiv, sn2=sn2))
        self._stage_params.setdefault(id(lst), []).append(wname)

    def _conv_wgrad_tail(self, lst, name, wname, xtail, c_off, r, y, taps, tap_off, N, Nvalid, sn, sc, W, H, B):
        """weight gradient of the LAST r (<= 16) input channels of a KxK convolution with few (<= 64) output channels, roles
        swapped: A = the output gradient y (rows = output channels), B = one halo patch of the r-channel activation slice
        xtail, every tap a shifted view of it (wgrad.cu family mode).  Keeps the odd channels of e.g. the head's 128+3+1
        input out of the 128-row m-tiles of the main launch."""
        T = len(taps)
        order = sorted(range(T), key=lambda t: (taps[t][1], taps[t][2]))
        nch = (N + 63) // 64
        a_slots = [(0, 0, 0, 64 * j, 64 * j) for j in range(nch)]
        b_slots = [(0, taps[t][1], taps[t][2], 0, i * 16) for i, t in enumerate(order)]
        ld = T * 16
        did = self._req_dw(nch * 64 * ld)
        d = ops.make_wgrad(a_srcs=[y], b_srcs=[xtail], a_slots=a_slots, b_slots=b_slots, n_tile=16, W=W, H=H, B=B, dw=0, ld=ld)
        P = B * H * W
        self._emit(lst, self.lib.dmm_conv_wgrad, d, name, kind="wgrad", flops=2.0 * P * r * Nvalid * T, nbytes=P * (N + 8) * 2)
        self._fix_dw.append((d, did))
        g = self.grad[wname].view(-1)[c_off * sc:]
        self._unpack_jobs.append(dict(wname=wname, did=did, grad=g, dt=16, dm=1, dn=ld, M=r, N=Nvalid, T=T, tap_off=[tap_off[t] for t in order],
                                      sn=sn, sc=sc, stage=id(lst)))
        self._stage_params.setdefault(id(lst), []).append(wname)

    def _fold(self, bn, C_, c0=0):
        """eval mode: (scale, shift) vectors of channels [c0, c0+C_) of `bn` (filled by dmm_bn_fold_batched at the head of every
        forward), padded to a multiple of 256 entries so that an n-tile may read past C_."""
        n = ceil_to(C_, 256)
        off = self._save.take(2 * n, 64)
        scale, shift = self._save.buf[off:off + n], self._save.buf[off + n:off + 2 * n]
        self._fold_jobs.append((bn, c0, C_, scale, shift))
        return scale, shift

    @property
    def fold_eval(self):
        return (not self.training) and (not self.f32) and FOLD_EVAL_BN

    def _bn_fwd(self, bn, stats, stats_off, count, c0=0, rep=1.0):
        return ops.make_bn(stats, stats_off, count, bn.gamma, bn.beta, bn.rm, bn.rv, bn.save_mean, bn.save_invstd,
                           training=self.training, rep=rep, c0=c0)

    def _apply(self, lst, name, bn, x, xc0, C_, stats, stats_off, y, yc0, pool=0, ystats=None, ystats_off=0, bn_c0=0,
               count=None):
        b = self._bn_fwd(bn, stats, stats_off, x.P if count is None else count, c0=bn_c0)
        if not self.training:
            ystats = None
        d = ops.make_bn_apply(x, xc0, C_, b, y, yc0, pool=pool, ystats=ystats, ystats_off=ystats_off)
        fn = self.lib.dmm_bn_relu_apply_f32 if self.f32 else self.lib.dmm_bn_relu_apply
        self._emit(lst, fn, d, name, kind="bn_relu_apply", nbytes=(x.P + y.P) * C_ * x.esize)

    def _bn_bwd(self, lst, name, bn, x, xc0, C_, g_ptr, ldg, out_ptr, ldo, out_mode, gmode=0, g_is_f32=False, bn_c0=0,
                gbuf=None, dz_tmp=None, producer=None):
        """BN-ReLU backward = reduce pass + apply pass.  dz_tmp (Mat): the reduce pass stores the masked, pool-routed
        gradient there and the apply pass reads it back with gmode 0 (max-pool routing is evaluated once)."""
        self._stage_params.setdefault(id(lst), []).extend([bn.prefix + ".weight", bn.prefix + ".bias"])
        sums = self._new_sums(C_)
        b = ops.make_bn_bwd(sums, 0, x.P, bn.gamma, bn.beta, bn.save_mean, bn.save_invstd, bn.dgamma, bn.dbeta, c0=bn_c0)
        d = ops.make_bn_bwd_args(x, xc0, C_, g_ptr, ldg, b, out_ptr, ldo, out_mode, gmode=gmode, g_is_f32=g_is_f32)
        pg = x.P if gmode == 0 else (x.P // 4)
        rd = x.P * C_ * 2 + pg * C_ * (4 if g_is_f32 else 2)
        wr = x.P * C_ * (2 if out_mode == 0 else (4 if gbuf is None else 8))
        d2 = d
        if dz_tmp is not None:
            d.dz_out, d.lddz = dz_tmp.ptr().value, dz_tmp.ld
            d2 = ops.make_bn_bwd_args(x, xc0, C_, dz_tmp.ptr(), dz_tmp.ld, b, out_ptr, ldo, out_mode, gmode=0, g_is_f32=False)
            rd2 = x.P * C_ * 4
        else:
            rd2 = rd
        if producer is not None and gmode == 0 and not g_is_f32 and dz_tmp is None and FUSE_BN_BWD_REDUCE and producer.heavy:
            # the data-gradient launch that produces g also accumulates (sum dz, sum dz*xhat): no separate reduce pass
            ops.fuse_bn_bwd_reduce(producer.arg, x, xc0, b)
            producer.bytes += x.P * C_ * 2
        else:
            self._emit(lst, self.lib.dmm_bn_relu_bwd_reduce, d, name + ".reduce", kind="bn_relu_bwd_reduce",
                       nbytes=rd + (x.P * C_ * 2 if dz_tmp is not None else 0))
        self._emit(lst, self.lib.dmm_bn_relu_bwd_apply, d2, name + ".apply", gbuf=gbuf, kind="bn_relu_bwd_apply",
                   nbytes=rd2 + wr)

    def _slab_bwd(self, lst, name, bn, blk, C_, g_ptr, ldg, gmode=0, bn_c0=0):
        """BatchNorm-ReLU backward of a consumer of dense-block buffer `blk` (transition / decoder skip) : the usual two
        passes, dx stored as its own bf16 slab that the gradient gathers of the block sum up (no fp32 read-modify-write)."""
        slab = self._mat(self.B, blk.H, blk.W, C_)
        self._bn_bwd(lst, name, bn, blk.buf, 0, C_, g_ptr, ldg, slab.ptr(), slab.ld, 0, gmode=gmode, bn_c0=bn_c0)
        blk.contribs.append(dict(mat=slab, C=C_, k=None, mean=None))

    def _contrib_bwd(self, lst, name, bn, blk, C_, g):
        """dense-layer norm1 backward in ONE pass (dmm_bn_relu_bwd_contrib): slab A*dz + sums; the per-channel correction
        vectors k (dmm_bn_bwd_finalize) are applied when the block's channel gradients are gathered."""
        self._stage_params.setdefault(id(lst), []).extend([bn.prefix + ".weight", bn.prefix + ".bias"])
        sums = self._new_sums(C_)
        b = ops.make_bn_bwd(sums, 0, blk.buf.P, bn.gamma, bn.beta, bn.save_mean, bn.save_invstd, bn.dgamma, bn.dbeta)
        slab = self._mat(self.B, blk.H, blk.W, C_)
        d = ops.make_bn_bwd_args(blk.buf, 0, C_, g.ptr(), g.ld, b, slab.ptr(), slab.ld, 0)
        # planar slab: one contiguous [P, k] matrix per growth-rate-wide channel group, so that the gather of a layer's k
        # channels reads whole cache lines (row-major slabs cost 1.5x the DRAM traffic: 64-byte pieces of wide rows)
        gw = self.k if (self.k % 8 == 0 and C_ % self.k == 0 and blk.C0 % self.k == 0) else 0
        if gw:
            d.out_gw, d.out_plane = gw, blk.buf.P * gw
        self._emit(lst, self.lib.dmm_bn_relu_bwd_contrib, d, name + ".contrib", kind="bn_relu_bwd_apply", nbytes=blk.buf.P * C_ * 6)
        off = self._save.take(2 * C_, 4)
        kvec = self._save.buf[off:off + 2 * C_]
        if FUSE_FINALIZE:
            # the last block of the contribution launch finalises dgamma / dbeta / k itself (ticket counters, zeroed per backward)
            d.fin_k = kvec.data_ptr()
            self._fin_req.append(d)
        else:
            def run_fin(_a, stream, b=b, C_=C_, kvec=kvec, lib=self.lib):
                return lib.dmm_bn_bwd_finalize(C.byref(b), C_, C.c_void_p(kvec.data_ptr()), stream)
            self._emit(lst, run_fin, None, name + ".finalize", kind="bn_finalize", nbytes=C_ * 150)
        self._keep.append(b)
        blk.contribs.append(dict(mat=slab, C=C_, k=kvec, mean=bn.save_mean, gw=gw))

    def _gather(self, lst, name, blk, c0, C_, dst):
        """gradient of channels [c0, c0+C_) of a block buffer = sum of the slabs of every consumer of those channels (all of
        them run earlier in the backward program) minus the deferred BatchNorm corrections.  Sources are resolved in _finalize
        (one dmm_grad_gather launch, or a chain of launches that accumulate into dst when a block has more than
        DMM_GATHER_MAX consumers)."""
        descs = []

        def run_gather(_a, stream, descs=descs, lib=self.lib):
            for d in descs:
                rc = lib.dmm_grad_gather(C.byref(d), stream)
                if rc:
                    return rc
            return 0
        self._emit(lst, run_gather, None, name, kind="grad_gather", nbytes=0.0)
        self._gathers.append((lst[-1], blk, c0, C_, dst, descs))

    def _cast(self, lst, name, src, c0, C_, dst):
        def run(_arg, stream, src=src, c0=c0, C_=C_, dst=dst, lib=self.lib):
            return lib.dmm_rows_f32_to_bf16(C.c_void_p(src.data_ptr() + 4 * c0), src.shape[1], dst.ptr(), dst.ld,
                                            src.shape[0], C_, stream)
        self._emit(lst, run, None, name, kind="cast_f32_bf16", nbytes=src.shape[0] * C_ * 6)

    # ------------------------------------------------------------------------------------------------
    # the plan
    # ------------------------------------------------------------------------------------------------
    def _plan(self):
        B, H, W = self.B, self.H, self.W
        k, bnk = self.k, self.bnk
        nb = len(self.block_config)
        self._fix_w, self._fix_dw = [], []
        self._gathers = []
        fwd = self.fwd
        H2, W2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        H4, W4 = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
        res = [(H4, W4)]
        for _ in range(nb - 1):
            res.append((res[-1][0] // 2, res[-1][1] // 2))
        if min(res[-1]) < 1:
            raise ValueError("input %dx%d is too small for %d dense blocks" % (H, W, nb))
        for b in range(nb - 1):
            if res[b][0] % 2 or res[b][1] % 2:
                raise ValueError("feature map %s of dense block %d is odd: the reference's ConvTranspose2d "
                                 "output_size check fails for this input size" % (res[b], b + 1))

        # channel bookkeeping (Dense_U_Net_lidar.py:81-100)
        cin_blk, ctot_blk = [], []
        nf = self.nif
        for b, L in enumerate(self.block_config):
            cin_blk.append(nf)
            nf += L * k
            ctot_blk.append(nf)
            if b != nb - 1:
                nf //= 2

        for c in cin_blk + ctot_blk + [v // 2 for v in ctot_blk[:-1]]:
            if c % 8:
                raise ValueError("dense-block channel count %d is not a multiple of 8 (bf16 pixel-major layout needs "
                                 "16-byte channel groups); choose growth_rate / block_config accordingly" % c)

        class Blk:
            pass

        def new_block(b):
            o = Blk()
            o.H, o.W = res[b]
            o.C0, o.Ct = cin_blk[b], ctot_blk[b]
            o.buf = self._mat(B, o.H, o.W, o.Ct)
            o.stats = self._new_stats(o.Ct)
            o.contribs = []      # backward: slabs written by the consumers of this buffer (see _gather)
            return o

        self._blk_cls = Blk

        conv1x1 = ops.conv_taps(1, 0)
        conv3x3 = ops.conv_taps(3, 1)

        # ---------------- stem (features.conv0/norm0/relu0/pool0) ----------------
        def stem(prefix, x1, c1, x2, c2, blk):
            cin = c1 + c2
            z0 = self._mat(B, H2, W2, self.nif)
            z0s = self._new_stats(self.nif)
            self.named[prefix + ".conv0"] = z0
            # (a single-channel stream keeps the small im2col matrix: its 7 narrow tap loads make the weight gradient slower)
            unfold = 7 * cin <= 64 and cin >= int(os.environ.get("DMM_STEM_UNFOLD_MIN_C", "2")) and not self.f32
            if unfold:
                # horizontal unfold only: xw[(b, iy, ox)][kw*cin + c] = x[c](iy, 2 ox + kw - 3) for ALL input rows; conv0 is then
                # a 7-tap vertical convolution over the even-row / odd-row views of xw (input row 2 oy + kh - 3 = 2 (oy + dy) + p)
                cw = ceil_to(7 * cin, 16)
                xw = self._mat(B, H, W2, cw)

                def run_unfold(_a, stream, x1=x1, x2=x2, c1=c1, c2=c2, xw=xw, lib=self.lib):
                    return lib.dmm_unfold_w7s2(C.c_void_p(x1.data_ptr()), c1, C.c_void_p(x2.data_ptr()) if c2 else None, c2,
                                               B, H, W, xw.ptr(), xw.ld, stream)
                self._emit(fwd, run_unfold, None, prefix + ".im2col", kind="im2col", nbytes=B * H * W * cin * 4 + B * H * W2 * cw * 2)
                phases = [xw.row_phase_view(0), xw.row_phase_view(1)]
                vtaps = [((kh - 3) % 2, (kh - 3 - (kh - 3) % 2) // 2, 0) for kh in range(7)]        # (phase, dy, 0)
                self._conv_fwd(fwd, prefix + ".conv0", prefix + ".conv0.weight", phases, vtaps, [7 * kh for kh in range(7)],
                               7 * cin, self.nif, cin * 49, 1, W2, H2, B, z0, 0, z0s, 0, cdiv=cin, sc2=49)
            else:
                kpad = ceil_to(cin * 49, int(os.environ.get("DMM_COL_ALIGN", "16")))   # 32-byte aligned rows: whole-sector stores
                col = self._mat(B, H2, W2, kpad)

                def run_im2col(_a, stream, x1=x1, x2=x2, c1=c1, c2=c2, col=col, lib=self.lib):
                    fn = lib.dmm_im2col_7x7s2_f32 if self.f32 else lib.dmm_im2col_7x7s2
                    return fn(C.c_void_p(x1.data_ptr()), c1, C.c_void_p(x2.data_ptr()) if c2 else None, c2,
                              B, H, W, col.ptr(), col.ld, stream)
                self._emit(fwd, run_im2col, None, prefix + ".im2col", kind="im2col",
                           nbytes=B * H * W * cin * 4 + B * H2 * W2 * kpad * 2)
                self._conv_fwd(fwd, prefix + ".conv0", prefix + ".conv0.weight", [col.view(0, kpad)], [(0, 0, 0)], [0],
                               cin * 49, self.nif, cin * 49, 1, W2, H2, B, z0, 0, z0s, 0)
            bn0 = _BNInfo(self, prefix + ".norm0", self.nif)
            amax = torch.empty((B * blk.H * blk.W, self.nif), dtype=torch.uint8, device=self._alloc_dev)
            self._keep.append(amax)
            self.mem_bytes += amax.numel()
            self._apply(fwd, prefix + ".norm0+pool0", bn0, z0, 0, self.nif, z0s, 0, blk.buf, 0, pool=2, ystats=blk.stats)
            fwd[-1].arg.argmax, fwd[-1].arg.ldarg = amax.data_ptr(), self.nif
            if self.need_backward:
                st = []
                dz0 = self._tmpmat("dz0", B, H2, W2, self.nif)
                dzr = self._tmpmat("dz0_routed", B, H2, W2, self.nif)
                g0 = self._tmpmat("g_blockin", B, blk.H, blk.W, self.nif)
                self._gather(st, prefix + ".gout", blk, 0, self.nif, g0)
                self._bn_bwd(st, prefix + ".norm0.bwd", bn0, z0, 0, self.nif, g0.ptr(), g0.ld, dz0.ptr(), dz0.ld, 0,
                             gmode=2, dz_tmp=dzr)
                st[-2].arg.argmax, st[-2].arg.ldarg = amax.data_ptr(), self.nif
                if unfold:
                    # dW[n, c, kh, kw] = sum_p dz0(p)[n] * xw_phase(p + dy)[kw*cin + c]: roles of "x" and "y" swapped so that the
                    # two row-phase views are the tap sources; gradient column kw*cin + c -> w[n, c, kh, kw]
                    self._conv_wgrad(st, prefix + ".conv0.wgrad", prefix + ".conv0.weight", dz0.view(), phases,
                                     [(p_, -dy_, 0) for (p_, dy_, _) in vtaps], [7 * kh for kh in range(7)], self.nif, cw,
                                     self.nif, 7 * cin, 49 if cin > 1 else 1, cin * 49, W2, H2, B, ndiv=cin if cin > 1 else 0, sn2=1)
                else:
                    self._conv_wgrad(st, prefix + ".conv0.wgrad", prefix + ".conv0.weight", col.view(0, kpad), [dz0.view()],
                                     [(0, 0, 0)], [0], kpad, self.nif, cin * 49, self.nif, cin * 49, 1, W2, H2, B)
                self._bwd_stages.append(st)

        # ---------------- dense block ----------------
        def dense_block(prefix, blk, L):
            Hb, Wb = blk.H, blk.W
            for i in range(L):
                Ci = blk.C0 + i * k
                lp = "%s.denselayer%d" % (prefix, i + 1)
                bn1 = _BNInfo(self, lp + ".norm1", Ci)
                bn2 = _BNInfo(self, lp + ".norm2", bnk)
                fold2 = self.fold_eval and FUSE_BN_PROLOGUE
                a2f = self._mat(B, Hb, Wb, bnk) if fold2 else None
                z1 = None if fold2 else self._mat(B, Hb, Wb, bnk)
                z1s = self._new_stats(bnk)
                if FUSE_BN_PROLOGUE and not self.f32:
                    # norm1 + relu1 run inside conv1: its A tiles are the raw block-buffer channels, activated in shared memory
                    # (inference: norm2 + relu2 are folded into the same launch - weights scaled, shift + ReLU in the epilogue)
                    a1 = None
                    d1 = self._conv_fwd(fwd, lp + ".conv1", lp + ".conv1.weight", [blk.buf.view(0, Ci)], conv1x1[0], conv1x1[2],
                                        Ci, bnk, Ci, 1, Wb, Hb, B, a2f if fold2 else z1, 0, z1s, 0,
                                        fold=self._fold(bn2, bnk) if fold2 else None)
                    d1.pro_enable = 1
                    d1.pro_bn = self._bn_fwd(bn1, blk.stats, 0, blk.buf.P)
                else:
                    a1 = self._mat(B, Hb, Wb, Ci)
                    self._apply(fwd, lp + ".norm1", bn1, blk.buf, 0, Ci, blk.stats, 0, a1, 0)
                    self._conv_fwd(fwd, lp + ".conv1", lp + ".conv1.weight", [a1.view()], conv1x1[0], conv1x1[2], Ci, bnk, Ci, 1,
                                   Wb, Hb, B, z1, 0, z1s, 0)
                if FUSE_BN_PROLOGUE_KXK and not self.f32:
                    a2 = None
                    d2 = self._conv_fwd(fwd, lp + ".conv2", lp + ".conv2.weight", [z1.view()], conv3x3[0], conv3x3[2], bnk, k,
                                        bnk * 9, 9, Wb, Hb, B, blk.buf, Ci, blk.stats, Ci)
                    d2.pro_enable = 1
                    d2.pro_bn = self._bn_fwd(bn2, z1s, 0, z1.P)
                else:
                    if fold2:
                        a2 = a2f
                    else:
                        a2 = self._mat(B, Hb, Wb, bnk)
                        self._apply(fwd, lp + ".norm2", bn2, z1, 0, bnk, z1s, 0, a2, 0)
                    if k == 32 and os.environ.get("DMM_CONV2_FOLD", "1") != "0" and not self.f32:
                        # kernel columns folded into N: 3 taps of N = 96 instead of 9 taps of N = 32 (2.9x fewer MMA cycles),
                        # the epilogue adds the three horizontal neighbours with warp shuffles (igemm out_mode 3)
                        self._conv_fwd(fwd, lp + ".conv2", lp + ".conv2.weight", [a2.view()], [(0, kh - 1, 0) for kh in range(3)],
                                       [3 * kh for kh in range(3)], bnk, k, bnk * 9, 9, Wb, Hb, B, blk.buf, Ci, blk.stats, Ci,
                                       out_mode=3, fold_kw=3, tile_w=32)
                    else:
                        self._conv_fwd(fwd, lp + ".conv2", lp + ".conv2.weight", [a2.view()], conv3x3[0], conv3x3[2], bnk, k,
                                       bnk * 9, 9, Wb, Hb, B, blk.buf, Ci, blk.stats, Ci)
                if self.need_backward:
                    st = []
                    kk = ceil_to(k, 8)
                    go = self._tmpmat("go", B, Hb, Wb, kk)
                    da2 = self._tmpmat("da2", B, Hb, Wb, bnk)
                    dz1 = self._tmpmat("dz1", B, Hb, Wb, bnk)
                    # shared scratch of the widest layer, re-pitched to this layer's channel count: dense rows for the
                    # data-gradient store and the contribution pass (a [P, Ct] pitch would leave holes in every DRAM page)
                    da1_full = self._tmpmat("da1", B, Hb, Wb, blk.Ct)
                    # (DMM_DA1_ALIGN=64 rounds the pitch up so that every 128-byte TMA-store row starts on a 128-byte boundary)
                    pitch = min(ceil_to(Ci, DA1_ALIGN), blk.Ct)
                    da1 = Mat(da1_full.t.view(-1)[:B * Hb * Wb * pitch].view(B * Hb * Wb, pitch), B, Hb, Wb)
                    self._gather(st, lp + ".gout", blk, Ci, k, go)
                    self._conv_wgrad(st, lp + ".conv2.wgrad", lp + ".conv2.weight", (z1 if a2 is None else a2).view(),
                                     [go.view(0, k)], conv3x3[0], conv3x3[2], bnk, k, bnk, k, bnk * 9, 9, Wb, Hb, B,
                                     pro=bn2 if a2 is None else None)
                    dg2 = self._conv_dgrad(st, lp + ".conv2.dgrad", lp + ".conv2.weight", [go.view(0, k)], conv3x3[1], conv3x3[2],
                                           k, bnk, 9, bnk * 9, Wb, Hb, B, da2, heavy=CONV2_DGRAD_FUSED)
                    self._bn_bwd(st, lp + ".norm2.bwd", bn2, z1, 0, bnk, da2.ptr(), da2.ld, dz1.ptr(), dz1.ld, 0, producer=dg2)
                    if a1 is None:
                        self._conv_wgrad(st, lp + ".conv1.wgrad", lp + ".conv1.weight", blk.buf.view(0, Ci), [dz1.view()],
                                         conv1x1[0], conv1x1[2], Ci, bnk, Ci, bnk, Ci, 1, Wb, Hb, B, pro=bn1)
                    else:
                        self._conv_wgrad(st, lp + ".conv1.wgrad", lp + ".conv1.weight", a1.view(), [dz1.view()], conv1x1[0],
                                         conv1x1[2], Ci, bnk, Ci, bnk, Ci, 1, Wb, Hb, B)
                    dg1 = self._conv_dgrad(st, lp + ".conv1.dgrad", lp + ".conv1.weight", [dz1.view()], conv1x1[1], conv1x1[2],
                                           bnk, Ci, 1, Ci, Wb, Hb, B, da1)
                    self._contrib_bwd(st, lp + ".norm1.bwd", bn1, blk, Ci, da1)
                    self._bwd_stages.append(st)

        # ---------------- transition: BN-ReLU -> (avg-pool first) -> 1x1 conv ----------------
        def transition(prefix, blk, out, out_stats, out_G, out_is_block):
            """out: Mat receiving Ct/2 channels at column 0 (next block buffer or a concat-module input)."""
            Ct, Co = blk.Ct, blk.Ct // 2
            bn = _BNInfo(self, prefix + ".norm", Ct)
            ap = self._mat(B, out.H, out.W, Ct)
            self._apply(fwd, prefix + ".norm+pool", bn, blk.buf, 0, Ct, blk.stats, 0, ap, 0, pool=1)
            self._conv_fwd(fwd, prefix + ".conv", prefix + ".conv.weight", [ap.view()], conv1x1[0], conv1x1[2], Ct, Co, Ct, 1,
                           out.W, out.H, B, out, 0, out_stats, 0)
            if self.need_backward:
                st = []
                if out_is_block:          # gradient lives in the fp32 block gradient buffer
                    gt = self._tmpmat("gt", B, out.H, out.W, Co)
                    self._gather(st, prefix + ".gout", out_G, 0, Co, gt)
                else:                     # bf16 gradient matrix written by the concat module backward
                    gt = out_G
                dap = self._tmpmat("dap", B, out.H, out.W, Ct)
                self._conv_wgrad(st, prefix + ".conv.wgrad", prefix + ".conv.weight", ap.view(), [gt.view(0, Co)], conv1x1[0],
                                 conv1x1[2], Ct, Co, Ct, Co, Ct, 1, out.W, out.H, B)
                self._conv_dgrad(st, prefix + ".conv.dgrad", prefix + ".conv.weight", [gt.view(0, Co)], conv1x1[1], conv1x1[2],
                                 Co, Ct, 1, Ct, out.W, out.H, B, dap)
                self._slab_bwd(st, prefix + ".norm.bwd", bn, blk, Ct, dap.ptr(), dap.ld, gmode=1)
                self._bwd_stages.append(st)

        # ================= encoder =================
        s1_blocks = [new_block(b) for b in range(nb)]
        for b, o in enumerate(s1_blocks):
            self.named["features.denseblock%d" % (b + 1)] = o.buf
        x2 = self.in2 if self.fusion == "early" else None
        stem("features", self.in1, self.c1, x2, self.c2 if self.fusion == "early" else 0, s1_blocks[0])

        s2_t = None
        if self.fusion == "mid":
            s2_blocks = [new_block(b) for b in range(self.cb - 1)]
            for b, o in enumerate(s2_blocks):
                self.named["stream_2_features.denseblock%d" % (b + 1)] = o.buf
            stem("stream_2_features", self.in2, self.c2, None, 0, s2_blocks[0])
            Cc = ctot_blk[self.cb - 2] // 2
            hc, wc = res[self.cb - 1]
            t2 = self._mat(B, hc, wc, Cc)
            t2s = self._new_stats(Cc)
            dt2 = self._tmpmat("dt2", B, hc, wc, Cc) if self.need_backward else None
            for b in range(self.cb - 1):
                dense_block("stream_2_features.denseblock%d" % (b + 1), s2_blocks[b], self.block_config[b])
                last = b == self.cb - 2
                if last:
                    transition("stream_2_features.transition%d" % (b + 1), s2_blocks[b], t2, t2s, dt2, False)
                else:
                    nxt = s2_blocks[b + 1]
                    transition("stream_2_features.transition%d" % (b + 1), s2_blocks[b], nxt.buf, nxt.stats, nxt, True)
            s2_t = (t2, t2s, dt2)

        for b in range(nb):
            blk = s1_blocks[b]
            dense_block("features.denseblock%d" % (b + 1), blk, self.block_config[b])
            if b == nb - 1:
                break
            nxt = s1_blocks[b + 1]
            if self.fusion == "mid" and b + 1 == self.cb - 1:
                Cc = blk.Ct // 2
                t1 = self._mat(B, nxt.H, nxt.W, Cc)
                t1s = self._new_stats(Cc)
                dt1 = self._tmpmat("dt1", B, nxt.H, nxt.W, Cc) if self.need_backward else None
                transition("features.transition%d" % (b + 1), blk, t1, t1s, dt1, False)
                # concat_module: BN(2C) -> ReLU -> 1x1 conv 2C -> C over cat(stream_1, stream_2) (:186-192, :242-245)
                t2, t2s, dt2 = s2_t
                bn = _BNInfo(self, "concat_module.norm", 2 * Cc)
                acat = self._mat(B, nxt.H, nxt.W, 2 * Cc)
                self._apply(fwd, "concat_module.norm[s1]", bn, t1, 0, Cc, t1s, 0, acat, 0)
                self._apply(fwd, "concat_module.norm[s2]", bn, t2, 0, Cc, t2s, 0, acat, Cc, bn_c0=Cc)
                self._conv_fwd(fwd, "concat_module.conv", "concat_module.conv.weight", [acat.view()], conv1x1[0], conv1x1[2],
                               2 * Cc, Cc, 2 * Cc, 1, nxt.W, nxt.H, B, nxt.buf, 0, nxt.stats, 0)
                if self.need_backward:
                    st = []
                    gt = self._tmpmat("gt", B, nxt.H, nxt.W, Cc)
                    dacat = self._tmpmat("dacat", B, nxt.H, nxt.W, 2 * Cc)
                    self._gather(st, "concat_module.gout", nxt, 0, Cc, gt)
                    self._conv_wgrad(st, "concat_module.conv.wgrad", "concat_module.conv.weight", acat.view(), [gt.view()],
                                     conv1x1[0], conv1x1[2], 2 * Cc, Cc, 2 * Cc, Cc, 2 * Cc, 1, nxt.W, nxt.H, B)
                    self._conv_dgrad(st, "concat_module.conv.dgrad", "concat_module.conv.weight", [gt.view()], conv1x1[1],
                                     conv1x1[2], Cc, 2 * Cc, 1, 2 * Cc, nxt.W, nxt.H, B, dacat)
                    self._bn_bwd(st, "concat_module.norm.bwd[s1]", bn, t1, 0, Cc, dacat.ptr(0), dacat.ld, dt1.ptr(), dt1.ld, 0)
                    self._bn_bwd(st, "concat_module.norm.bwd[s2]", bn, t2, 0, Cc, dacat.ptr(Cc), dacat.ld, dt2.ptr(), dt2.ld, 0,
                                 bn_c0=Cc)
                    self._bwd_stages.append(st)
            else:
                transition("features.transition%d" % (b + 1), blk, nxt.buf, nxt.stats, nxt, True)

        # ================= decoder (:105-120, :255-261) =================
        fstack = [self.nif + 2 * k] + ctot_blk            # feature_size_stack (:81-82,95)
        sizes = [(H2, W2)] + [res[b] for b in range(nb - 1)]
        cur = s1_blocks[nb - 1]                            # raw input of Sequence_1 = block4 buffer
        cur_raw, cur_stats, cur_C = cur.buf, cur.stats, cur.Ct
        num_in = fstack.pop()
        up_prev = None                                     # (u Mat, stats, du Mat) of the previous ConvTranspose
        for kdec in range(1, nb + 1):
            num_f = fstack.pop()
            sp = "decoder.Transposed_Convolution_Sequence_%d" % kdec
            cp = "decoder.Transposed_Convolution_%d" % kdec
            bn0 = _BNInfo(self, sp + ".norm0", num_in)
            bn1 = _BNInfo(self, sp + ".norm1", num_f)
            if kdec == 1:
                hk, wk = cur.H, cur.W
                a = self._mat(B, hk, wk, num_in)
                self._apply(fwd, sp + ".norm0", bn0, cur_raw, 0, num_in, cur_stats, 0, a, 0)
                skip_blk = None
            else:
                u, us, du = up_prev
                skip_blk = s1_blocks[nb - kdec]
                hk, wk = skip_blk.H, skip_blk.W
                Cu = u.ld
                assert Cu + skip_blk.Ct == num_in
                a = self._mat(B, hk, wk, num_in)
                self._apply(fwd, sp + ".norm0[up]", bn0, u, 0, Cu, us, 0, a, 0)
                self._apply(fwd, sp + ".norm0[skip]", bn0, skip_blk.buf, 0, skip_blk.Ct, skip_blk.stats, 0, a, Cu, bn_c0=Cu)
            rs = self._new_stats(num_f)
            a1 = self._mat(B, hk, wk, num_f)
            if self.fold_eval:
                r = a1       # inference: norm1 + relu1 folded into conv_reduce
                self._conv_fwd(fwd, sp + ".conv_reduce", sp + ".conv_reduce.weight", [a.view()], conv1x1[0], conv1x1[2], num_in,
                               num_f, num_in, 1, wk, hk, B, a1, 0, rs, 0, fold=self._fold(bn1, num_f))
            else:
                r = self._mat(B, hk, wk, num_f)
                self._conv_fwd(fwd, sp + ".conv_reduce", sp + ".conv_reduce.weight", [a.view()], conv1x1[0], conv1x1[2], num_in,
                               num_f, num_in, 1, wk, hk, B, r, 0, rs, 0)
                self._apply(fwd, sp + ".norm1", bn1, r, 0, num_f, rs, 0, a1, 0)
            oh, ow = sizes.pop()
            for dim_in, dim_out in ((hk, oh), (wk, ow)):
                if not (2 * dim_in - 1 <= dim_out <= 2 * dim_in):
                    raise ValueError("requested an output size of %s, but valid sizes range from %d to %d"
                                     % ((oh, ow), 2 * dim_in - 1, 2 * dim_in))
            unew = self._mat(B, oh, ow, num_f)
            unews = self._new_stats(num_f)
            self.named["decoder.%d" % kdec] = unew
            self.named["decoder.%d.reduce" % kdec] = r
            for py in range(2):
                for px in range(2):
                    taps, off = ops.convt_phase_taps(py, px)
                    self._conv_fwd(fwd, cp + "[%d%d]" % (py, px), cp + ".weight", [a1.view()], taps, off, num_f, num_f, 9,
                                   num_f * 9, wk, hk, B, unew, 0, unews, 0, out_stride=(2, 2), out_phase=(py, px),
                                   out_hw=(oh, ow))
            dunew = self._mat(B, oh, ow, num_f) if self.need_backward else None
            if self.need_backward:
                st = []
                da1 = self._tmpmat("dec_da1", B, hk, wk, num_f)
                dr = self._tmpmat("dec_dr", B, hk, wk, num_f)
                da = self._tmpmat("dec_da", B, hk, wk, num_in)
                wt, woff = ops.convt_wgrad_taps()
                dt, doff = ops.convt_dgrad_taps()
                phases = [dunew.phase_view(py, px) for py in range(2) for px in range(2)]
                self._conv_wgrad(st, cp + ".wgrad", cp + ".weight", a1.view(), phases, wt, woff, num_f, num_f, num_f, num_f, 9,
                                 num_f * 9, wk, hk, B)
                dgt = self._conv_dgrad(st, cp + ".dgrad", cp + ".weight", phases, dt, doff, num_f, num_f, num_f * 9, 9, wk, hk, B, da1)
                self._bn_bwd(st, sp + ".norm1.bwd", bn1, r, 0, num_f, da1.ptr(), da1.ld, dr.ptr(), dr.ld, 0, producer=dgt)
                self._conv_wgrad(st, sp + ".conv_reduce.wgrad", sp + ".conv_reduce.weight", a.view(), [dr.view()], conv1x1[0],
                                 conv1x1[2], num_in, num_f, num_in, num_f, num_in, 1, wk, hk, B)
                self._conv_dgrad(st, sp + ".conv_reduce.dgrad", sp + ".conv_reduce.weight", [dr.view()], conv1x1[1], conv1x1[2],
                                 num_f, num_in, 1, num_in, wk, hk, B, da)
                if kdec == 1:
                    self._slab_bwd(st, sp + ".norm0.bwd", bn0, cur, num_in, da.ptr(), da.ld)
                else:
                    self._bn_bwd(st, sp + ".norm0.bwd[up]", bn0, u, 0, Cu, da.ptr(0), da.ld, du.ptr(), du.ld, 0)
                    self._slab_bwd(st, sp + ".norm0.bwd[skip]", bn0, skip_blk, skip_blk.Ct, da.ptr(Cu), da.ld, bn_c0=Cu)
                self._bwd_stages.append(st)
            up_prev = (unew, unews, dunew)
            num_in = num_f * 2

        # ================= head (:120-132, :264-265) =================
        u, us, du = up_prev
        Cu = u.ld
        cx = self.c1 + self.c2
        Ct = Cu + cx
        if (2 * u.H, 2 * u.W) != (H, W):
            raise RuntimeError("Sizes of tensors must match except in dimension 1 (decoder output %dx%d vs input %dx%d)"
                               % (2 * u.H, 2 * u.W, H, W))
        hp = "dec_out_to_heat_maps"
        bn0 = _BNInfo(self, hp + ".norm0", Ct)
        nf2 = Cu // 2
        bn1 = _BNInfo(self, hp + ".norm1", nf2)
        xst = self._new_stats(8 if cx <= 8 else ceil_to(cx, 8))
        # row pitch of the head activations: a multiple of 16 channels keeps every pixel row 32-byte (sector) aligned, so neither
        # the 16-byte stores of head_input nor the 128-byte TMA-store rows of refine0's data gradient leave partial sectors
        ld0 = ceil_to(Ct, int(os.environ.get("DMM_HEAD_LD_ALIGN", "16")))
        a0 = self._mat(B, H, W, ld0)

        def run_xstats(_a, stream, lib=self.lib):
            rc = lib.dmm_nchw_stats(C.c_void_p(self.in1.data_ptr()), B, self.c1, H * W, xst.ptr(), xst.ld, 0, stream)
            if rc == 0 and self.c2:
                rc = lib.dmm_nchw_stats(C.c_void_p(self.in2.data_ptr()), B, self.c2, H * W, xst.ptr(), xst.ld, self.c1, stream)
            return rc
        if self.training:
            self._emit(fwd, run_xstats, None, hp + ".input_stats", kind="nchw_stats", nbytes=B * H * W * cx * 4)
        hd = Head()
        hd.u, hd.ldu, hd.Cu = u.ptr().value, u.ld, Cu
        hd.x1, hd.C1 = self.in1.data_ptr(), self.c1
        hd.x2, hd.C2 = (self.in2.data_ptr() if self.c2 else None), self.c2
        hd.B, hd.H, hd.W = B, H, W
        hd.bn_u = self._bn_fwd(bn0, us, 0, u.P, rep=4.0)
        hd.bn_x = self._bn_fwd(bn0, xst, 0, B * H * W, c0=Cu)
        hd.out, hd.ldo = a0.ptr().value, a0.ld
        self._emit(fwd, self.lib.dmm_head_input_f32 if self.f32 else self.lib.dmm_head_input, hd, hp + ".upsample+cat+norm0", kind="head_input",
                   nbytes=u.P * Cu * 2 + B * H * W * (cx * 4 + ld0 * 2))
        r0 = self._mat(B, H, W, nf2)
        r0s = self._new_stats(nf2)
        self.named["head.a0"] = a0
        self.named["head.refine0"] = r0
        if nf2 == 64 and os.environ.get("DMM_HEAD_FOLD0", "0") != "0" and not self.f32:
            # kernel columns folded into N = 192 (igemm out_mode 3): 3 taps at 96 cycles per MMA instead of 9 taps of N = 64.
            # Parity-tested but OFF: only one 128-pixel sub-tile fits in TMEM at N = 192, so every tile re-streams all weights
            # and the launch is L2-bound at the same 3.9 ms as the 9-tap form (measured)
            self._conv_fwd(fwd, hp + ".refine0", hp + ".refine0.weight", [a0.view(0, Ct)], [(0, kh - 1, 0) for kh in range(3)],
                           [3 * kh for kh in range(3)], Ct, nf2, Ct * 9, 9, W, H, B, r0, 0, r0s, 0, out_mode=3, fold_kw=3, tile_w=32)
            a1h = self._mat(B, H, W, nf2)
            self._apply(fwd, hp + ".norm1", bn1, r0, 0, nf2, r0s, 0, a1h, 0)
        elif self.fold_eval:
            a1h = r0         # inference: norm1 + relu1 folded into refine0
            self._conv_fwd(fwd, hp + ".refine0", hp + ".refine0.weight", [a0.view(0, Ct)], conv3x3[0], conv3x3[2], Ct, nf2, Ct * 9, 9,
                           W, H, B, r0, 0, r0s, 0, fold=self._fold(bn1, nf2))
        else:
            self._conv_fwd(fwd, hp + ".refine0", hp + ".refine0.weight", [a0.view(0, Ct)], conv3x3[0], conv3x3[2], Ct, nf2, Ct * 9, 9,
                           W, H, B, r0, 0, r0s, 0)
            a1h = self._mat(B, H, W, nf2)
            self._apply(fwd, hp + ".norm1", bn1, r0, 0, nf2, r0s, 0, a1h, 0)
        conv5 = ops.conv_taps(5, 2)
        if 5 * self.ncls <= 16 and os.environ.get("DMM_HEAD_FOLD", "1") != "0" and not self.f32:
            # kernel columns folded into N: 5 (kernel rows) instead of 25 MMAs per pixel tile, horizontal sum in the epilogue
            self._conv_fwd(fwd, hp + ".refine1", hp + ".refine1.weight", [a1h.view()], [(0, kh - 2, 0) for kh in range(5)],
                           [5 * kh for kh in range(5)], nf2, self.ncls, nf2 * 25, 25, W, H, B, None, 0, None, 0, out_mode=2,
                           out_ptr=self.logits.data_ptr(), fold_kw=5, tile_w=32 if W >= 28 else (16 if W >= 12 else 8))
        else:
            self._conv_fwd(fwd, hp + ".refine1", hp + ".refine1.weight", [a1h.view()], conv5[0], conv5[2], nf2, self.ncls, nf2 * 25,
                           25, W, H, B, None, 0, None, 0, out_mode=1, out_ptr=self.logits.data_ptr())
        if self.need_backward:
            st = []
            # d(logits) as a pixel-major bf16 matrix.  With 5 * num_classes <= 16 the five HORIZONTAL shifts of every class
            # share the 16 columns (column kw*ncls + n = dlogits[n](y, x - (kw - 2))): the data gradient of refine1 is then a
            # 5-tap vertical convolution (5 instead of 25 MMAs per tile); columns 2*ncls.. hold the unshifted gradient.
            dl = self._mat(B, H, W, 16)
            unfold = 5 * self.ncls <= 16

            def run_dl(_a, stream, lib=self.lib, dl=dl, unfold=unfold):
                if unfold:
                    return lib.dmm_dlogits_unfold_w(C.c_void_p(self.dlogits.data_ptr()), B, self.ncls, H, W, 5, dl.ptr(), dl.ld, stream)
                return lib.dmm_nchw_to_nhwc_bf16(C.c_void_p(self.dlogits.data_ptr()), B, self.ncls, H, W, dl.ptr(), dl.ld, stream)
            self._emit(st, run_dl, None, hp + ".dlogits_nhwc", kind="nchw_to_nhwc", nbytes=B * H * W * (self.ncls * 4 + 16 * 2))
            da1h = self._tmpmat("head_da1", B, H, W, nf2)
            dr0 = self._tmpmat("head_dr0", B, H, W, nf2)
            da0 = self._tmpmat("head_da0", B, H, W, ld0)
            # refine1 (5x5, 64 -> num_classes): weight gradient = activation x 25 shifted views of d(logits) (one halo patch,
            # five wide MMAs per k-step); data gradient = a 25-tap convolution over the 16-channel d(logits) matrix
            if unfold:
                # dW[n, ci, kh, kw] = sum_p a1h(p + (kh - 2, 0))[ci] * dl(p)[kw*ncls + n]: five kernel-row taps, the kernel columns
                # ride in the 16 gradient columns (one N = 80 MMA per k-step instead of five)
                self._conv_wgrad(st, hp + ".refine1.wgrad", hp + ".refine1.weight", a1h.view(), [dl.view(0, 16)],
                                 [(0, kh - 2, 0) for kh in range(5)], [5 * kh for kh in range(5)], nf2, 16, nf2, 5 * self.ncls,
                                 nf2 * 25, 25, W, H, B, ndiv=self.ncls, sn2=1)
            else:
                self._conv_wgrad(st, hp + ".refine1.wgrad", hp + ".refine1.weight", a1h.view(), [dl.view(0, 16)], conv5[0], conv5[2],
                                 nf2, 16, nf2, self.ncls, nf2 * 25, 25, W, H, B)
            if unfold:
                # packed weight column (tap kh, channel c = kw*ncls + n) <- w[n, ci, kh, kw]
                dgh = self._conv_dgrad(st, hp + ".refine1.dgrad", hp + ".refine1.weight", [dl.view(0, 16)],
                                       [(0, 2 - kh, 0) for kh in range(5)], [5 * kh for kh in range(5)], 5 * self.ncls, nf2, 25, 1,
                                       W, H, B, da1h, cdiv=self.ncls, sc2=nf2 * 25,
                                       heavy=os.environ.get("DMM_HEAD_DGRAD_FUSED", "1") != "0")
            else:
                dgh = self._conv_dgrad(st, hp + ".refine1.dgrad", hp + ".refine1.weight", [dl.view(0, self.ncls)], conv5[1],
                                       conv5[2], self.ncls, nf2, 25, nf2 * 25, W, H, B, da1h)
            self._bn_bwd(st, hp + ".norm1.bwd", bn1, r0, 0, nf2, da1h.ptr(), da1h.ld, dr0.ptr(), dr0.ld, 0, producer=dgh)
            if Cu % 64 == 0 and 0 < cx <= 16:
                # the 128 decoder channels fill one 128-row m-tile exactly; the few raw input channels go through the tail path
                self._conv_wgrad(st, hp + ".refine0.wgrad", hp + ".refine0.weight", a0.view(0, Cu), [dr0.view()], conv3x3[0],
                                 conv3x3[2], Cu, nf2, Cu, nf2, Ct * 9, 9, W, H, B)
                self._conv_wgrad_tail(st, hp + ".refine0.wgrad[tail]", hp + ".refine0.weight", a0.view(Cu, min(16, ld0 - Cu)), Cu, cx,
                                      dr0.view(), conv3x3[0], conv3x3[2], nf2, nf2, Ct * 9, 9, W, H, B)
            else:
                self._conv_wgrad(st, hp + ".refine0.wgrad", hp + ".refine0.weight", a0.view(0, Ct), [dr0.view()], conv3x3[0],
                                 conv3x3[2], Ct, nf2, Ct, nf2, Ct * 9, 9, W, H, B)
            self._conv_dgrad(st, hp + ".refine0.dgrad", hp + ".refine0.weight", [dr0.view()], conv3x3[1], conv3x3[2], nf2, Ct, 9,
                             Ct * 9, W, H, B, da0)
            hb = HeadBwd()
            hb.u, hb.ldu, hb.Cu = u.ptr().value, u.ld, Cu
            hb.x1, hb.C1 = self.in1.data_ptr(), self.c1
            hb.x2, hb.C2 = (self.in2.data_ptr() if self.c2 else None), self.c2
            hb.B, hb.H, hb.W = B, H, W
            hb.g, hb.ldg = da0.ptr().value, da0.ld
            su, sx = self._new_sums(Cu), self._new_sums(8 if cx <= 8 else ceil_to(cx, 8))
            hb.bn_u = ops.make_bn_bwd(su, 0, B * H * W, bn0.gamma, bn0.beta, bn0.save_mean, bn0.save_invstd, bn0.dgamma,
                                      bn0.dbeta)
            hb.bn_x = ops.make_bn_bwd(sx, 0, B * H * W, bn0.gamma, bn0.beta, bn0.save_mean, bn0.save_invstd, bn0.dgamma,
                                      bn0.dbeta, c0=Cu)
            hb.du, hb.lddu = du.ptr().value, du.ld
            self._stage_params.setdefault(id(st), []).extend([hp + ".norm0.weight", hp + ".norm0.bias"])
            hb_rd = u.P * Cu * 2 + B * H * W * (cx * 4 + ld0 * 2)
            self._emit(st, self.lib.dmm_head_input_bwd_reduce, hb, hp + ".norm0.bwd.reduce", kind="head_input_bwd",
                       nbytes=hb_rd)
            self._emit(st, self.lib.dmm_head_input_bwd_apply, hb, hp + ".norm0.bwd.apply", kind="head_input_bwd",
                       nbytes=u.P * Cu * 4 + B * H * W * Cu * 2)
            self._bwd_stages.append(st)

    # ------------------------------------------------------------------------------------------------
    def _finalize(self):
        dev = self.dev
        # packed weights arena
        offs, tot = [], 0
        for n in self._wpk_req:
            offs.append(tot)
            tot += ceil_to(n, 512)      # 1 KiB alignment for TMA
        wsz = 4 if self.f32 else 2
        self._wpk = torch.zeros(max(tot, 1), dtype=self.adt, device=dev)
        self.mem_bytes += tot * wsz
        base = self._wpk.data_ptr()
        for d, wid in self._fix_w:
            d.weights = base + wsz * offs[wid]
        pj = np.zeros(len(self._pack_jobs), dtype=_PACK_DT)
        for i, j in enumerate(self._pack_jobs):
            pj[i]["w"] = j["w"].data_ptr()
            pj[i]["dst"] = base + wsz * offs[j["wid"]]
            pj[i]["n_valid"], pj[i]["n_rows"], pj[i]["C"], pj[i]["T"] = j["n_valid"], j["n_rows"], j["C"], j["T"]
            pj[i]["kwidth"] = j.get("kwidth", ops.KWIDTH)
            pj[i]["tap_off"][:j["T"]] = j["tap_off"]
            pj[i]["sn"], pj[i]["sc"] = j["sn"], j["sc"]
            pj[i]["sc2"], pj[i]["cdiv"] = j.get("sc2", 0), j.get("cdiv", 0)
            pj[i]["sn2"], pj[i]["ndiv"] = j.get("sn2", 0), j.get("ndiv", 0)
            pj[i]["rscale"] = j["rscale"].data_ptr() if j.get("rscale") is not None else 0
        fj = np.zeros(len(self._fold_jobs), dtype=_FOLD_DT)
        for i, (bn, c0, C_, scale, shift) in enumerate(self._fold_jobs):
            fj[i]["gamma"], fj[i]["beta"] = bn.gamma.data_ptr() + 4 * c0, bn.beta.data_ptr() + 4 * c0
            fj[i]["rm"], fj[i]["rv"] = bn.rm.data_ptr() + 4 * c0, bn.rv.data_ptr() + 4 * c0
            fj[i]["scale"], fj[i]["shift"] = scale.data_ptr(), shift.data_ptr()
            fj[i]["C"], fj[i]["eps"] = C_, ops.BN_EPS
        self._fold_tab = torch.from_numpy(fj.view(np.uint8).copy()).to(dev) if len(fj) else None
        self._pack_tab = torch.from_numpy(pj.view(np.uint8).copy()).to(dev)
        self._n_pack = len(self._pack_jobs)
        # load-balanced launch: one block per WORK_CHUNK consecutive packed elements of a job
        self._pack_work = _work_table([j["n_rows"] * j["T"] * ceil_to(j["C"], j.get("kwidth", ops.KWIDTH)) for j in self._pack_jobs], dev)
        self._param_ptrs = [j["w"].data_ptr() for j in self._pack_jobs]
        self._pack_src = [j["w"] for j in self._pack_jobs]
        # backward program = stages in reverse forward order, cut into SEGMENTS whose parameter gradients form one
        # contiguous range of the flat gradient buffer (a gradient bucket, >= bucket_bytes) each
        stages = list(reversed(self._bwd_stages))
        stage_rank = {id(st): i for i, st in enumerate(stages)}
        self._unpack_jobs.sort(key=lambda j: stage_rank[j["stage"]])
        # weight-gradient scratch arena
        offs, tot = [], 0
        for n in self._dw_req:
            offs.append(tot)
            tot += ceil_to(n, 64)
        self._dw = torch.zeros(max(tot, 1), dtype=torch.float32, device=dev)
        self.mem_bytes += tot * 4
        base = self._dw.data_ptr()
        for d, did in self._fix_dw:
            d.dw = base + 4 * offs[did]
        uj = np.zeros(len(self._unpack_jobs), dtype=_UNPACK_DT)
        for i, j in enumerate(self._unpack_jobs):
            uj[i]["dw"] = base + 4 * (offs[j["did"]] + j.get("dw_off", 0))
            uj[i]["grad"] = j["grad"].data_ptr()
            uj[i]["dt"], uj[i]["dm"], uj[i]["dn"] = j["dt"], j["dm"], j["dn"]
            uj[i]["M"], uj[i]["N"], uj[i]["T"] = j["M"], j["N"], j["T"]
            uj[i]["accumulate"] = 0
            uj[i]["tap_off"][:j["T"]] = j["tap_off"]
            uj[i]["sn"], uj[i]["sc"] = j["sn"], j["sc"]
            uj[i]["sn2"], uj[i]["ndiv"] = j.get("sn2", 0), j.get("ndiv", 0)
        self._unpack_tab = torch.from_numpy(uj.view(np.uint8).copy()).to(dev)
        self._n_unpack = len(self._unpack_jobs)
        self._unpack_sizes = [j["T"] * j["M"] * j["N"] for j in self._unpack_jobs]
        self.bwd = []
        self.segments = []       # (ops, first unpack job, number of unpack jobs, flat_lo, flat_hi)
        self._seg_work = []      # per segment: (job, chunk) work table of its unpack jobs (indices relative to job_lo)
        seg_ops, seg_names, job_lo, job_i = [], [], 0, 0
        done = set()
        for si, st in enumerate(stages):
            marker = Op(None, None, "stage%d" % si, kind="stage_begin")
            seg_ops.append(marker)
            seg_ops.extend(st)
            self.bwd.extend(st)
            for n in self._stage_params.get(id(st), []):
                if n not in done:
                    done.add(n)
                    seg_names.append(n)
            while job_i < len(self._unpack_jobs) and stage_rank[self._unpack_jobs[job_i]["stage"]] <= si:
                job_i += 1
            nbytes = 4 * sum(self.p[n].numel() for n in seg_names)
            if nbytes >= self.bucket_bytes or si == len(stages) - 1:
                if seg_names:
                    lo = self.grad_offset[seg_names[0]]
                    hi = self.grad_offset[seg_names[-1]] + self.p[seg_names[-1]].numel()
                else:
                    lo = hi = 0
                self.segments.append((seg_ops, job_lo, job_i - job_lo, lo, hi))
                self._seg_work.append(_work_table(self._unpack_sizes[job_lo:job_i], dev))
                seg_ops, seg_names, job_lo = [], [], job_i
        self._fin_ctr = torch.zeros(max(8 * len(self._fin_req), 8), dtype=torch.int32, device=dev)
        for i, d in enumerate(self._fin_req):
            d.fin_ctr = self._fin_ctr.data_ptr() + 4 * 8 * i
        self.gather_launches = 0
        for op, blk, c0, C_, dst, descs in self._gathers:
            srcs = [c for c in blk.contribs if c["C"] >= c0 + C_]
            if not srcs:
                raise RuntimeError("dmmfods_b200: no gradient source for channels [%d, %d) of a block buffer (%s)"
                                   % (c0, c0 + C_, op.name))
            gws = {c.get("gw", 0) for c in srcs} - {0}
            if len(gws) > 1 or any(c0 % g_ for g_ in gws):
                raise RuntimeError("dmmfods_b200: inconsistent planar slab groups for %s" % op.name)
            gw = gws.pop() if gws else 0
            # more consumers than one launch takes: chain, every further launch re-reads dst as its first source
            groups, i, first = [], 0, True
            while i < len(srcs):
                n = _lib.GATHER_MAX if first else _lib.GATHER_MAX - 1
                groups.append(srcs[i:i + n])
                i += n
                first = False
            any_k = False
            for gi, grp in enumerate(groups):
                d = _lib.GradGather()
                d.rows, d.C = blk.buf.P, C_
                d.out, d.ldo = dst.ptr().value, dst.ld
                d.gw = gw
                ns = nk = 0
                if gi > 0:
                    d.src[0], d.ld[0], d.plane[0] = dst.ptr().value, dst.ld, 0
                    ns = 1
                for c in grp:
                    if c.get("gw", 0):    # planar: group c0 / gw starts at that plane; the kernel adds further groups itself
                        plane = blk.buf.P * c["gw"]
                        d.src[ns] = c["mat"].t.data_ptr() + 2 * (c0 // c["gw"]) * plane
                        d.ld[ns] = c["gw"]
                        d.plane[ns] = plane
                    else:
                        d.src[ns] = c["mat"].ptr(c0).value
                        d.ld[ns] = c["mat"].ld
                        d.plane[ns] = 0
                    ns += 1
                    if c["k"] is not None:
                        d.k1[nk] = c["k"].data_ptr() + 4 * c0
                        d.k2[nk] = c["k"].data_ptr() + 4 * (c["C"] + c0)
                        d.mean = c["mean"].data_ptr() + 4 * c0
                        nk += 1
                d.nsrc, d.nk = ns, nk
                if nk:
                    d.x, d.ldx = blk.buf.ptr(c0).value, blk.buf.ld
                    any_k = True
                descs.append(d)
            self.gather_launches += len(descs)
            op.bytes = float(blk.buf.P * C_ * 2 * (len(srcs) + 1 + (1 if any_k else 0) + 2 * (len(descs) - 1)))
        seen = set()
        for op in self.bwd:
            if op.gbuf is not None:
                op.arg.out_mode = 2 if id(op.gbuf) in seen else 1
                seen.add(id(op.gbuf))
        self.nbt = [self.p[k] for k in self.nbt_keys]

    # ------------------------------------------------------------------------------------------------
    def _run(self, program):
        if self.plan_only:
            raise RuntimeError("dmmfods_b200: plan-only engine cannot execute (no CUDA device)")
        main = torch.cuda.current_stream()
        stream = C.c_void_p(main.cuda_stream)
        byref = C.byref
        debug = bool(os.environ.get("DMM_DEBUG_SYNC"))
        side = self._side_stream() if (WGRAD_SIDE_STREAM and not debug) else None
        for op in program:
            if op.kind == "stage_begin":
                # temporaries (go, dz1, ...) are re-written from here on: weight-gradient kernels of the previous stage that
                # still read them on the side stream must have finished
                self._join_side(main)
                continue
            if side is not None and op.kind == "wgrad":
                # weight gradients feed nothing but the final unpack: they run beside the dgrad -> BN-backward chain
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                rc = op.fn(byref(op.arg), C.c_void_p(side.cuda_stream))
                self._side_busy = True
                if rc != 0:
                    raise RuntimeError("dmmfods_b200: %s failed (rc=%d): %s" % (op.name, rc, _lib.last_error()))
                continue
            rc = op.fn(byref(op.arg), stream) if op.arg is not None else op.fn(None, stream)
            if rc != 0:
                raise RuntimeError("dmmfods_b200: %s failed (rc=%d): %s" % (op.name, rc, _lib.last_error()))
            if debug:
                try:
                    torch.cuda.synchronize()
                except Exception as e:      # noqa: BLE001
                    raise RuntimeError("dmmfods_b200: kernel of op %s faulted: %s" % (op.name, e))

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.dev)
            self._side_busy = False
        return self._side

    def _join_side(self, main):
        if getattr(self, "_side", None) is not None and self._side_busy:
            ev = torch.cuda.Event()
            ev.record(self._side)
            main.wait_event(ev)
            self._side_busy = False

    def check_param_pointers(self):
        for t, ptr in zip(self._pack_src, self._param_ptrs):
            if t.data_ptr() != ptr:
                raise RuntimeError("dmmfods_b200: a parameter tensor was re-allocated after the engine was built")

    def forward(self, x1, x2):
        """logits (B, num_classes, H, W) fp32 - engine-owned buffer, valid until the next forward()."""
        if self.plan_only:
            raise RuntimeError("dmmfods_b200: plan-only engine cannot execute (no CUDA device)")
        if x1 is not self.in1:
            self.in1.copy_(x1, non_blocking=True)
        if self.c2 and x2 is not self.in2:
            self.in2.copy_(x2, non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.training:
            self._stats.zero_used()
        if self._fold_tab is not None:
            _lib.check(self.lib.dmm_bn_fold_batched(C.c_void_p(self._fold_tab.data_ptr()), len(self._fold_jobs), stream),
                       "dmm_bn_fold_batched")
        if self.f32:
            _lib.check(self.lib.dmm_pack_weights_work_f32(C.c_void_p(self._pack_tab.data_ptr()), C.c_void_p(self._pack_work.data_ptr()),
                                                          self._pack_work.shape[0], WORK_CHUNK, 1 if self.split3 else 0, stream),
                       "dmm_pack_weights_work_f32")
        elif BALANCED_PACK:
            _lib.check(self.lib.dmm_pack_weights_work(C.c_void_p(self._pack_tab.data_ptr()), C.c_void_p(self._pack_work.data_ptr()),
                                                      self._pack_work.shape[0], WORK_CHUNK, stream), "dmm_pack_weights_work")
        else:
            _lib.check(self.lib.dmm_pack_weights_batched(C.c_void_p(self._pack_tab.data_ptr()), self._n_pack, stream),
                       "dmm_pack_weights_batched")
        self._run(self.fwd)
        if self.training and self.nbt:
            torch._foreach_add_(self.nbt, 1)
        return self.logits

    def backward(self, dlogits=None, on_bucket=None):
        """gradient of all parameters for the cotangent dlogits (defaults to the engine's own dlogits buffer, as
        filled by loss()).  Results land in self.grad[name] (views of the flat fp32 buffer self.gflat).
        on_bucket(i, flat_range) is called as soon as the i-th gradient bucket is final (stream-ordered)."""
        if not self.need_backward:
            raise RuntimeError("engine was built without backward support")
        if dlogits is not None:
            self.dlogits.copy_(dlogits)
        self.backward_begin()
        for i in range(len(self.segments)):
            flat = self.backward_segment(i)
            if on_bucket is not None and flat is not None:
                on_bucket(i, flat)
        return self.grad

    def backward_begin(self):
        """zero the backward accumulators (first half of backward(); separate so that a caller can capture the segments of
        the backward program in individual CUDA graphs with the gradient all-reduce between them)."""
        self._sums.zero_used()
        self._dw.zero_()
        if self._fin_req:
            self._fin_ctr.zero_()

    def backward_segment(self, i):
        """run backward segment i (stages whose parameter gradients form gradient bucket i); returns the bucket's range
        of the flat gradient buffer, final once the enqueued work completes, or None if the segment owns no parameters."""
        seg_ops, job_lo, njobs, lo, hi = self.segments[i]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._run(seg_ops)
        self._join_side(torch.cuda.current_stream())
        if njobs:
            tab = self._unpack_tab.data_ptr() + job_lo * _UNPACK_DT.itemsize
            if BALANCED_PACK:
                work = self._seg_work[i]
                _lib.check(self.lib.dmm_unpack_wgrad_work(C.c_void_p(tab), C.c_void_p(work.data_ptr()), work.shape[0], WORK_CHUNK,
                                                          stream), "dmm_unpack_wgrad_work")
            else:
                _lib.check(self.lib.dmm_unpack_wgrad_batched(C.c_void_p(tab), njobs, stream), "dmm_unpack_wgrad_batched")
        return self.gflat[lo:hi] if hi > lo else None

    def loss(self, target, loss_out=None):
        """BCEWithLogits(reduction='none') of the current logits; fills dlogits (= d sum(loss) / d logits) and
        the per-class loss sums (Agent.py:247-249)."""
        self.class_sums.zero_()
        ops.bce_logits(self.logits, target, loss_out, self.dlogits, self.class_sums)
        return self.class_sums


_PACK_DT = np.dtype([("w", np.uint64), ("dst", np.uint64), ("n_valid", np.int32), ("n_rows", np.int32), ("C", np.int32),
                     ("kwidth", np.int32), ("T", np.int32), ("tap_off", np.int32, (32,)), ("sn", np.int64),
                     ("sc", np.int64), ("sc2", np.int64), ("cdiv", np.int32), ("ndiv", np.int32), ("sn2", np.int64),
                     ("rscale", np.uint64)], align=True)
_FOLD_DT = np.dtype([("gamma", np.uint64), ("beta", np.uint64), ("rm", np.uint64), ("rv", np.uint64), ("scale", np.uint64),
                     ("shift", np.uint64), ("C", np.int32), ("eps", np.float32)], align=True)
_UNPACK_DT = np.dtype([("dw", np.uint64), ("grad", np.uint64), ("dt", np.int64), ("dm", np.int64), ("dn", np.int64),
                       ("M", np.int32), ("N", np.int32), ("T", np.int32), ("accumulate", np.int32),
                       ("tap_off", np.int32, (32,)), ("sn", np.int64), ("sc", np.int64), ("sn2", np.int64), ("ndiv", np.int32),
                       ("pad_", np.int32)], align=True)
