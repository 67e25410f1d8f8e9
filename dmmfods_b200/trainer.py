"""Training step driver: forward + BCE heat-map loss + backward (+ bucketed NCCL gradient all-reduce overlapped
with backward) + fused flat Adam - the per-iteration work of Dense_U_Net_lidar_Agent.train_one_epoch
(Agent.py:244-265) without its logging.  One process per GPU; BatchNorm statistics stay per process (the reference
uses plain nn.BatchNorm2d), gradients are SUMMED over ranks (the reference back-propagates the sum over all samples,
Agent.py:264).
"""
import ctypes as C
import os

import torch

from . import _lib, ops
from .data import mark_async_read
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


class StepLR:
    """torch.optim.lr_scheduler.StepLR semantics for the flat fused Adam (Agent.py:63-67, stepped once per epoch at
    Agent.py:297-298): lr = base_lr * gamma ** (epoch // step_size)."""

    def __init__(self, trainer, step_size, gamma=0.1):
        self.trainer, self.step_size, self.gamma = trainer, int(step_size), float(gamma)
        self.base_lr = trainer.lr
        self.last_epoch = 0

    def step(self):
        self.last_epoch += 1
        self.trainer.lr = self.base_lr * self.gamma ** (self.last_epoch // self.step_size)

    def get_last_lr(self):
        return [self.trainer.lr]

    def state_dict(self):
        return {"step_size": self.step_size, "gamma": self.gamma, "base_lr": self.base_lr, "last_epoch": self.last_epoch}

    def load_state_dict(self, sd):
        self.step_size, self.gamma = int(sd["step_size"]), float(sd["gamma"])
        self.base_lr, self.last_epoch = float(sd["base_lr"]), int(sd["last_epoch"])
        self.trainer.lr = self.base_lr * self.gamma ** (self.last_epoch // self.step_size)


class Trainer:
    """forward + loss + backward + (all-reduce) + Adam over static buffers.

        tr = Trainer(model, B, H, W, lr=..., use_graph=True)
        loss_per_class = tr.step(image, lidar, ht_map)          # CUDA or pinned host tensors

    The returned per-class loss sums are the ENGINE'S STATIC buffer (float64 CUDA tensor): the next step() overwrites it;
    `.clone()` it (or copy it to the host) to keep the value."""

    def __init__(self, model, B, H, W, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 distributed=None, bucket_bytes=32 << 20, use_graph=False, graph_nccl=None, preprocess=None):
        import torch.distributed as dist
        if amsgrad:
            raise NotImplementedError("dmmfods_b200.Trainer: amsgrad=True is not supported by the fused flat Adam "
                                      "(the reference default is False, helper:151)")
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
        order = Engine.gradient_order(params, model.model_cfg(), B, H, W)      # (independent of the bucket size)
        total = sum(named[n].numel() for n in order)
        self.pflat = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for n in order:
            p = named[n]
            view = self.pflat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.eng = model.engine(B, H, W, bucket_bytes=bucket_bytes)
        assert self.eng.param_names == order
        self.exp_avg = torch.zeros_like(self.pflat)
        self.exp_avg_sq = torch.zeros_like(self.pflat)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.steps = 0
        self.reducer = BucketReducer(self.dist, dev) if self.dist is not None else None
        # one CUDA graph for forward + loss + backward; data parallel: one graph for forward + loss and one per backward segment
        # with the bucket all-reduces (NCCL, side stream) issued between them (default), or - graph_nccl / DMM_GRAPH_NCCL=1,
        # experimental - the all-reduces captured INSIDE one step graph
        self.use_graph = bool(use_graph)
        self.graph_nccl = (os.environ.get("DMM_GRAPH_NCCL", "0") != "0") if graph_nccl is None else bool(graph_nccl)
        self.graph = None
        self.seg_graphs = None
        # preprocess(lidar_buffer, target_buffer): optional on-GPU pre-processing launched at the head of every step (inside the
        # step graph): fills the engine's LiDAR input and the heat-map targets from raw point lists / label boxes
        # (helper.BatchPreprocessor.run - BASELINE config 4); step() is then called with x2 = target = None
        self.preprocess = preprocess
        self._static_target = torch.empty((B, self.eng.ncls, H, W), dtype=torch.float32, device=dev)
        self._copy_stream, self._stage, self._staged, self._stage_ready, self._stage_free = None, None, None, None, None

    # ------------------------------------------------------------------------------------------------
    def _forward(self):
        eng = self.eng
        if self.preprocess is not None:
            self.preprocess(eng.in2 if eng.c2 else None, self._static_target)
        eng.forward(eng.in1, eng.in2)

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
            self._stage = [None if t is None else torch.empty(t.shape, dtype=torch.float32, device=dev) for t in (x1, x2, target)]
        cs = self._copy_stream
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)          # the previous hand-over has finished reading the staging buffers
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage, (x1, x2, target)):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
            self._stage_ready = torch.cuda.Event()
            self._stage_ready.record(cs)
        # the pinned sources must not be refilled before this copy has run (BatchFileRing waits for the event)
        mark_async_read((x1, x2, target), self._stage_ready)
        self._staged = (x1, x2, target)

    def _take_staged(self, x1, x2, target):
        """if (x1, x2, target) are the tensors handed to prefetch(): wait for that copy and return the device copies."""
        st = self._staged
        if st is None or st[0] is not x1 or st[1] is not x2 or st[2] is not target:
            return None
        torch.cuda.current_stream().wait_event(self._stage_ready)
        self._staged = None
        return self._stage

    def _load_inputs(self, x1, x2, target):
        """bring this step's inputs into the engine-owned static buffers (in1, in2, target)."""
        eng = self.eng
        staged = self._take_staged(x1, x2, target)
        if staged is not None:
            # hand-over: device-to-device into the engine-owned buffers, then the staging buffers are free again and the
            # next prefetch overlaps with this step's kernels
            eng.in1.copy_(staged[0], non_blocking=True)
            if eng.c2 and staged[1] is not None:
                eng.in2.copy_(staged[1], non_blocking=True)
            if staged[2] is not None:
                self._static_target.copy_(staged[2], non_blocking=True)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
            return
        host = []
        for dst, src in ((eng.in1, x1), (eng.in2 if eng.c2 else None, x2), (self._static_target, target)):
            if dst is None or src is dst or src is None:
                continue
            dst.copy_(src, non_blocking=True)
            if not src.is_cuda:
                host.append(src)
        if host:
            ev = torch.cuda.Event()
            ev.record()
            mark_async_read(host, ev)

    def step(self, x1, x2, target, prefetch_next=None):
        """one optimisation step; x1/x2/target: CUDA tensors or pinned host tensors (copied asynchronously).
        prefetch_next: optional (x1, x2, target) of the following step (see prefetch()).
        Returns the per-class loss sums (float64 CUDA tensor of num_classes entries, Agent.py:248) - the engine's static
        buffer, overwritten by the next step."""
        eng = self.eng
        self._load_inputs(x1, x2, target)
        if self.use_graph:
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
            self._forward()
            self._fwd_loss_bwd(self._static_target)
        if prefetch_next is not None:
            self.prefetch(*prefetch_next)
        self.steps += 1
        ops.adam_flat(self.pflat, eng.gflat, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.steps)
        return eng.class_sums

    def _bn_state(self):
        return [v for k, v in self.eng.p.items() if "running_" in k or k.endswith("num_batches_tracked")]

    def _capture(self):
        """CUDA graph of forward + loss + backward over the engine's static buffers (launch-bound otherwise).
        The warm-up pass torch's capture rules require is a real forward + backward: the BatchNorm running statistics and
        num_batches_tracked it updates are restored afterwards, so the first replayed step is the ONLY momentum update of this
        batch (state_dict parity with the un-graphed path and with the reference)."""
        eng = self.eng
        state = self._bn_state()
        saved = [t.clone() for t in state]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._forward()
            self._fwd_loss_bwd(self._static_target)
        torch.cuda.current_stream().wait_stream(s)
        for t, v in zip(state, saved):
            t.copy_(v)
        g = torch.cuda.CUDAGraph()
        if self.dist is None:
            with torch.cuda.graph(g):
                self._forward()
                self._fwd_loss_bwd(self._static_target)
            self.graph = g
            return
        if self.graph_nccl:
            try:
                with torch.cuda.graph(g):
                    self._forward()
                    self._fwd_loss_bwd(self._static_target)
                self.graph = g
                return
            except Exception as e:      # noqa: BLE001  (NCCL build without capture support: say so, use the segmented form)
                import sys
                sys.stderr.write("dmmfods_b200.Trainer: capturing the NCCL all-reduces failed (%s); using per-segment graphs\n" % e)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._forward()
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
        n += eng.gather_launches - sum(1 for op in eng.bwd if op.kind == "grad_gather")          # chained gathers
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
