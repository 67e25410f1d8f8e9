"""Training step driver: forward + BCE heat-map loss + backward (+ bucketed NCCL gradient all-reduce overlapped
with backward) + fused flat Adam - the per-iteration work of Dense_U_Net_lidar_Agent.train_one_epoch
(Agent.py:244-265) without its logging.  One process per GPU; BatchNorm statistics stay per process (the reference
uses plain nn.BatchNorm2d), gradients are SUMMED over ranks (the reference back-propagates the sum over all samples,
Agent.py:264).
"""
import ctypes as C

import torch

from . import _lib, ops
from .engine import Engine


class BucketReducer:
    """SUM all-reduce of the gradient buckets as they become final during backward.

    CUDA: every bucket is reduced on a side stream that waits for the event recorded on the compute stream when
    the bucket was finalised, so NCCL overlaps the rest of backward; finish() makes the compute stream wait for
    the side stream.  CPU tensors (gloo; host-logic tests): asynchronous work handles, finish() waits for them."""

    def __init__(self, dist, device):
        self.dist = dist
        self.cuda = torch.device(device).type == "cuda"
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self.pending = []
        self.ranges = []

    def __call__(self, i, flat):
        self.ranges.append((i, flat.numel()))
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM)
        else:
            self.pending.append(self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)
        for w in self.pending:
            w.wait()
        self.pending = []


class Trainer:
    def __init__(self, model, B, H, W, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, distributed=None,
                 bucket_bytes=32 << 20, use_graph=False):
        import torch.distributed as dist
        self.model = model.train()
        self.dist = dist if (distributed if distributed is not None else (dist.is_available() and dist.is_initialized()
                                                                          and dist.get_world_size() > 1)) else None
        named = dict(model.named_parameters())
        sd = model.state_dict(keep_vars=True)
        params = {k: (v.data if isinstance(v, torch.nn.Parameter) else v) for k, v in sd.items()}
        dev = next(iter(params.values())).device
        if dev.type != "cuda":
            raise RuntimeError("dmmfods_b200.Trainer: the model must live on a CUDA device (no CPU path)")
        # flatten the parameters in the engine's gradient order so that Adam and the all-reduce see flat ranges
        order = Engine.gradient_order(params, model.model_cfg(), B, H, W)
        total = sum(named[n].numel() for n in order)
        self.pflat = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for n in order:
            p = named[n]
            view = self.pflat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.eng = model.engine(B, H, W)
        assert self.eng.param_names == order
        self.eng.bucket_bytes = bucket_bytes
        self.exp_avg = torch.zeros_like(self.pflat)
        self.exp_avg_sq = torch.zeros_like(self.pflat)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.steps = 0
        self.reducer = BucketReducer(self.dist, dev) if self.dist is not None else None
        # single GPU: one CUDA graph for forward + loss + backward.  Data parallel: one graph for forward + loss and one per
        # backward segment, the bucket all-reduces are issued between them (NCCL calls stay outside the graphs)
        self.use_graph = bool(use_graph)
        self.graph = None
        self.seg_graphs = None
        self._static_target = None
        self._copy_stream, self._stage, self._staged, self._stage_ready, self._stage_free = None, None, None, None, None

    # ------------------------------------------------------------------------------------------------
    def _fwd_loss_bwd(self, target):
        eng = self.eng
        eng.loss(target)
        eng.backward(on_bucket=self.reducer)
        if self.reducer is not None:
            self.reducer.finish()

    # ------------------------------------------------------------------------------------------------
    def prefetch(self, x1, x2, target):
        """start the host -> device copy of the NEXT step's (pinned) inputs on a copy stream while the current step
        computes (the data-loader prefetch of a training loop).  step() called with the same tensor objects then only
        does a device-to-device hand-over."""
        dev = self.pflat.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [torch.empty(t.shape, dtype=torch.float32, device=dev) for t in (x1, x2, target)]
        cs = self._copy_stream
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)          # the previous hand-over has finished reading the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage, (x1, x2, target)):
                dst.copy_(src, non_blocking=True)
            self._stage_ready = torch.cuda.Event()
            self._stage_ready.record(cs)
        self._staged = (x1, x2, target)

    def _take_staged(self, x1, x2, target):
        """if (x1, x2, target) are the tensors handed to prefetch(): wait for that copy and return the device copies."""
        st = self._staged
        if st is None or st[0] is not x1 or st[1] is not x2 or st[2] is not target:
            return None
        torch.cuda.current_stream().wait_event(self._stage_ready)
        self._staged = None
        return self._stage

    def step(self, x1, x2, target, prefetch_next=None):
        """one optimisation step; x1/x2/target: CUDA tensors or pinned host tensors (copied asynchronously).
        prefetch_next: optional (x1, x2, target) of the following step (see prefetch()).
        Returns the per-class loss sums (float64 CUDA tensor of num_classes entries, Agent.py:248)."""
        eng = self.eng
        staged = self._take_staged(x1, x2, target)
        if staged is not None:
            # hand-over: device-to-device into the engine-owned buffers, then the staging buffers are free again and the
            # next prefetch overlaps with this step's kernels
            if self._static_target is None:
                self._static_target = torch.empty(target.shape, dtype=torch.float32, device=self.pflat.device)
            eng.in1.copy_(staged[0], non_blocking=True)
            if eng.c2:
                eng.in2.copy_(staged[1], non_blocking=True)
            self._static_target.copy_(staged[2], non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
            x1, x2, target = eng.in1, eng.in2, self._static_target
        if not target.is_cuda:
            if self._static_target is None:
                self._static_target = torch.empty(target.shape, dtype=torch.float32, device=self.pflat.device)
            self._static_target.copy_(target, non_blocking=True)
            target = self._static_target
        if self.use_graph:
            if self._static_target is None:
                self._static_target = torch.empty(target.shape, dtype=torch.float32, device=self.pflat.device)
            if target is not self._static_target:
                self._static_target.copy_(target, non_blocking=True)
            if x1 is not eng.in1:
                eng.in1.copy_(x1, non_blocking=True)
            if eng.c2 and x2 is not eng.in2:
                eng.in2.copy_(x2, non_blocking=True)
            if self.graph is None:
                self._capture()
            self.graph.replay()
            if self.seg_graphs is not None:
                for i, (g, flat) in enumerate(self.seg_graphs):
                    g.replay()
                    if flat is not None:
                        self.reducer(i, flat)
                self.reducer.finish()
        else:
            eng.forward(x1, x2)
            self._fwd_loss_bwd(target)
        if prefetch_next is not None:
            self.prefetch(*prefetch_next)
        self.steps += 1
        ops.adam_flat(self.pflat, eng.gflat, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.steps)
        return eng.class_sums

    def _capture(self):
        """CUDA graph of forward + loss + backward over the engine's static buffers (launch-bound otherwise)."""
        eng = self.eng
        # warm-up on a side stream as torch's capture rules require, then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eng.forward(eng.in1, eng.in2)
            self._fwd_loss_bwd(self._static_target)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        if self.dist is None:
            with torch.cuda.graph(g):
                eng.forward(eng.in1, eng.in2)
                self._fwd_loss_bwd(self._static_target)
        else:
            with torch.cuda.graph(g):
                eng.forward(eng.in1, eng.in2)
                eng.loss(self._static_target)
                eng.backward_begin()
            self.seg_graphs = []
            for i in range(len(eng.segments)):
                gi = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gi, pool=g.pool()):
                    flat = eng.backward_segment(i)
                self.seg_graphs.append((gi, flat))
        self.graph = g

    def launches_per_step(self):
        """number of this library's kernel launches in one step."""
        eng = self.eng
        n = len(eng.fwd) + len(eng.bwd) + 1 + len([s for s in eng.segments if s[2]]) + 1 + 1   # + pack, unpacks, bce, adam
        n += sum(1 for op in eng.fwd if op.kind == "nchw_stats" and eng.c2)                     # second input tensor
        return n


class Evaluator:
    """Validation / inference side of the hot path (SURVEY 8(f) N1 + N4): the body of `Dense_U_Net_lidar_Agent.validate`
    (Agent.py:337-352) - eval-mode forward (running-statistics BatchNorm, no backward buffers), per-class BCE loss sums,
    whole-image IoU per instance and class, class-wise accuracy - as one replayable CUDA graph over static buffers.

        ev = Evaluator(model, B, H, W)
        out = ev.step(image, lidar, ht_map)      # dict: loss_per_class, iou_per_instance_per_class (nan = empty union),
                                                 #       iou_nans, acc_per_class   (CUDA tensors, no host sync)
        maps = ev.heat_maps()                    # sigmoid(logits) of the last step (notebook visual check)
    """

    def __init__(self, model, B, H, W, iou_threshold=0.7, use_graph=True):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dmmfods_b200.Evaluator: the model must live on a CUDA device (no CPU path)")
        self.model = model
        self.eng = model.engine(B, H, W, training=False, need_backward=False)
        self.threshold = float(iou_threshold)
        self.use_graph = bool(use_graph)
        self.graph = None
        ncls = self.eng.ncls
        self.B, self.H, self.W, self.ncls = B, H, W, ncls
        self._target = torch.zeros(B, ncls, H, W, dtype=torch.float32, device=dev)
        self._counts = torch.zeros(B * ncls, 3, dtype=torch.int64, device=dev)

    def _run(self):
        eng = self.eng
        eng.forward(eng.in1, eng.in2)
        eng.loss(self._target)
        self._counts.zero_()
        _lib.check(_lib.load().dmm_step_metrics(C.c_void_p(eng.logits.data_ptr()), C.c_void_p(self._target.data_ptr()),
                                                self.B * self.ncls, self.H * self.W, self.threshold,
                                                C.c_void_p(self._counts.data_ptr()), ops._stream()), "dmm_step_metrics")

    def step(self, x1, x2, target):
        eng = self.eng
        eng.check_param_pointers()
        if x1 is not eng.in1:
            eng.in1.copy_(x1, non_blocking=True)
        if eng.c2 and x2 is not eng.in2:
            eng.in2.copy_(x2, non_blocking=True)
        self._target.copy_(target, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._run()
                torch.cuda.current_stream().wait_stream(s)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._run()
            self.graph.replay()
        else:
            self._run()
        c = self._counts.view(self.B, self.ncls, 3)
        iou = c[:, :, 0].float() / c[:, :, 1].float()                      # nan where the union is empty (helper:311-343)
        acc = c[:, :, 2].sum(0).double() / float(self.B * self.H * self.W)    # (TP + TN) / all, per class (helper:369-401)
        return dict(loss_per_class=eng.class_sums, iou_per_instance_per_class=iou, iou_nans=torch.isnan(iou).sum(0),
                    acc_per_class=acc, logits=eng.logits)

    def heat_maps(self):
        return torch.sigmoid(self.eng.logits)
