#!/usr/bin/env python
"""bench.py - train images/s of the mid-fusion Dense-U-Net hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mid|cfg4|cfg5] [--api trainer|module]
    torchrun ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

  --workload mid   (default) BASELINE configs[2], the configuration the metric is quoted on
             cfg4  configs[3]: 1280x1920, batch 8 per GPU, per-step on-GPU LiDAR splat + bbox heat-map masks inside the step graph
             cfg5  configs[4]: DenseNet-201 mid-fusion eval-mode heat-map throughput sweep, batch 1..128 (1 GPU)
  --api trainer    (default) dmmfods_b200.trainer.Trainer.step - the replacement of the agent's loop body (Agent.py:244-265)
        module     the LITERAL loop body on the drop-in nn.Module: model(image, lidar) -> FusedBCEWithLogits -> backward(ones)
                   -> torch.optim.Adam.step()

One step = forward + per-pixel BCE heat-map loss + backward + (bucketed gradient all-reduce) + fused Adam
over ONE batch of synthetic Waymo-shaped RGB+LiDAR tensors (BASELINE configs[2]: DenseNet-121 encoder,
mid-fusion concat before denseblock3, batch 32 per GPU, 640x960; weak scaling).  Prints ONE JSON line.

  value     images/s with the batch already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       images/s through the public API (dmmfods_b200.trainer.Trainer.step) fed from PINNED HOST tensors:
            the H2D copy of the inputs/targets and a D2H read of the loss are inside the timed region
  roofline  dominant kernel family, algorithmic FLOPs (or bytes) / CUDA-event time of its launches in one
            instrumented step, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference algorithm on the host cores (oracle port, torch CPU ops),
            bounded sample of the same workload.  /root/reference is never read here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2] - the configuration the metric is quoted on
    "mid": dict(name="DenseNet-121 mid-fusion (concat before denseblock3), 640x960, fwd+loss+bwd+Adam",
                c2=1, cb=3, batch=32, H=640, W=960),
    # configs[3]: full Waymo FRONT resolution, per-step on-GPU pre-processing (30 000 LiDAR points and 5-60 boxes per frame)
    "cfg4": dict(name="DenseNet-121 mid-fusion, 1280x1920, on-GPU LiDAR splat + heat-map masks + fwd+loss+bwd+Adam",
                 c2=1, cb=3, batch=8, H=1280, W=1920, points=30000),
    # configs[4]: inference-only sweep
    "cfg5": dict(name="DenseNet-201 mid-fusion eval-mode heat maps (forward + loss + IoU/accuracy counters), 640x960, batch sweep",
                 c2=1, cb=3, batch=32, H=640, W=960, sweep=(1, 2, 4, 8, 16, 32, 64, 128)),
}


def model_cfg(c2, cb):
    from dmmfods_b200 import config as cfgmod
    c = cfgmod.get_config("/nonexistent")
    c.model.stream_2_in_channels = c2
    c.model.concat_before_block_num = cb
    return c


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """number of samples read so far (the sampler is started BEFORE the warm-up: nvidia-smi needs ~0.5 s to produce its
        first row, longer than the default timed region)."""
        return len(self.rows)

    def stop(self, first=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows[first:] if len(self.rows) > first else self.rows[-4:]      # the samples of the timed region
        self.rows = rows
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's algorithm on the host cores (oracle port)
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_step(wl, sample_batch, steps, warmup, threads):
    """times forward + BCE(none) + backward(ones) (Agent.py:244-264) of the oracle port in fp32 on `threads` cores.
    Returns (images/s, seconds per step)."""
    import torch
    from dmmfods_b200 import synthetic
    from dmmfods_b200.model import densenet121_u_lidar
    from oracle import dense_unet_oracle as du
    torch.set_num_threads(threads)
    torch.manual_seed(123)
    model = densenet121_u_lidar(pretrained=False, config=model_cfg(wl["c2"], wl["cb"]))   # parameter container only
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mc = model.model_cfg()
    B, H, W = sample_batch, wl["H"], wl["W"]
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W))
    x2 = torch.from_numpy(synthetic.lidar_image(B, H, W))
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        du.oracle_train_step(sd, mc, x1, x2, tgt, dtype=torch.float32)
        times.append(time.perf_counter() - t0)
    t = sum(times[warmup:]) / max(1, steps)
    return B / t, t


def cpu_model_string():
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_config1(threads, steps=10, warmup=3):
    """BASELINE configs[0] exactly - DenseNet-121 no-fusion, B = 2, U[0,255) RGB 3x256x384, forward + BCE(none)
    + backward(ones), fp32, 3 warm-up + 10 timed steps: min / median seconds per step on all host cores."""
    import torch
    from dmmfods_b200 import synthetic
    from dmmfods_b200.model import densenet121_u_lidar
    from oracle import dense_unet_oracle as du
    torch.set_num_threads(threads)
    torch.manual_seed(123)
    model = densenet121_u_lidar(pretrained=False, config=model_cfg(0, 1))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mc = model.model_cfg()
    B, H, W = 2, 256, 384
    x1 = torch.from_numpy(synthetic.rgb_image(B, H, W))
    x2 = torch.zeros(B, 1, H, W)
    tgt = torch.from_numpy(synthetic.target_maps(B, H, W))
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        du.oracle_train_step(sd, mc, x1, x2, tgt, dtype=torch.float32)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    return {"workload": "BASELINE configs[0]: DenseNet-121 no-fusion, B=2, 3x256x384, fwd+BCE+bwd, fp32", "steps": steps, "warmup": warmup,
            "s_per_step_min": ts[0], "s_per_step_median": ts[len(ts) // 2], "images_per_s_median": B / ts[len(ts) // 2],
            "images_per_s_best": B / ts[0]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS["mid"]
    threads = os.cpu_count() or 1
    sample = args.cpu_batch
    ips, t = cpu_reference_step(wl, sample, args.steps, args.warmup, threads)
    c1 = cpu_config1(threads)
    line = {
        "impl": "reference", "metric": "train images/sec (mid-fusion Dense-U-Net, DenseNet-121, 640x960)", "value": ips,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_step": sample, "H": wl["H"], "W": wl["W"],
                   "note": "reference algorithm on host cores (oracle port of Dense_U_Net_lidar fwd+BCE+bwd, torch CPU ops)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_string(),
                         "sample": "%d step(s) of batch %d (of the workload's 32) at %dx%d, fp32, torch CPU ops on all host cores; "
                                   "/root/reference does not exist on the GPU box, so the oracle port stands in for it"
                                   % (args.steps, sample, wl["H"], wl["W"]), "config1": c1},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
def kernel_breakdown(eng, x1, x2, tgt, dump=None, pre=None, backward=True):
    """one instrumented step: CUDA events around every launch, grouped by kernel family.  The stream is first held busy by
    a device-side spin so that the host enqueues the whole program ahead of the GPU: the event pairs then bracket kernels
    that run back to back (un-held, the small launches of blocks 3/4 are host-bound and their brackets include the gap)."""
    import ctypes as C
    import torch
    recs = []
    orig_run = eng._run

    def hold():
        torch.cuda._sleep(int(os.environ.get("DMM_BENCH_HOLD_CYCLES", "40000000")))

    def timed_run(program):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        hold()
        for op in program:
            if op.kind == "stage_begin":
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = op.fn(C.byref(op.arg), stream) if op.arg is not None else op.fn(None, stream)
            e1.record()
            assert rc == 0, op.name
            recs.append((op, e0, e1))
    eng._run = timed_run
    try:
        if pre is not None:
            pre()
        eng.forward(x1, x2)
        eng.loss(tgt)
        if backward:
            eng.backward()
        torch.cuda.synchronize()
    finally:
        eng._run = orig_run
    if dump:
        rows = [dict(name=op.name, kind=op.kind, ms=e0.elapsed_time(e1), flops=op.flops, bytes=op.bytes) for op, e0, e1 in recs]
        with open(dump, "w") as fh:
            json.dump(rows, fh)
    fam = {}
    for op, e0, e1 in recs:
        f = fam.setdefault(op.kind, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        f["ms"] += e0.elapsed_time(e1)
        f["flops"] += op.flops
        f["bytes"] += op.bytes
        f["launches"] += 1
    return fam


def roofline_from(fam, peaks):
    tot_ms = sum(f["ms"] for f in fam.values())
    kernels = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        kernels[k] = {"ms": round(f["ms"], 3), "share": round(f["ms"] / tot_ms, 4), "launches": f["launches"],
                      "tflops": round(f["flops"] / (f["ms"] * 1e9), 2) if f["flops"] else None,
                      "gbs": round(f["bytes"] / (f["ms"] * 1e6), 1)}
    dom = next(iter(kernels))
    tensor_bound = dom in ("igemm_fprop", "igemm_dgrad", "wgrad")
    f = fam[dom]
    # roofline of the dominant kernel family: algorithmic work of all its launches in one step / their summed CUDA-event time
    # (= per-launch work / average launch duration); "traffic" = average DRAM bytes per launch from the committed ncu capture
    if tensor_bound:
        ach = f["flops"] / (f["ms"] * 1e9)
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_sustained"], "traffic": None, "peak_source": peaks["source"] + " (sustained bf16)"}
    else:
        ach = f["bytes"] / (f["ms"] * 1e6)
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                "traffic": None, "peak_source": peaks["source"]}
    roof["launches"] = f["launches"]
    roof["algorithmic_bytes_per_launch"] = f["bytes"] / f["launches"]
    roof["algorithmic_flops_per_launch"] = f["flops"] / f["launches"]
    roof["avg_launch_ms"] = f["ms"] / f["launches"]
    # the HBM-bound families, for the >= 70 % of HBM peak target of BASELINE.json
    roof["hbm_families"] = {k: round(v["bytes"] / (v["ms"] * 1e6) / peaks["hbm"], 3) for k, v in fam.items()
                            if not v["flops"] and v["ms"] > 0.5}
    roof["tensor_families"] = {k: round(v["flops"] / (v["ms"] * 1e9) / peaks["tf_sustained"], 3) for k, v in fam.items() if v["flops"]}
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tfile):
        try:
            t = json.load(open(tfile))
            roof["traffic"] = t.get(dom)
            roof["traffic_source"] = t.get("_source", "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch)")
        except Exception:
            pass
    return roof, kernels, sum(v["flops"] for v in fam.values())


def cpu_baseline_obj(wl, args):
    threads = os.cpu_count() or 1
    ips, t = cpu_reference_step(wl, args.cpu_batch, 2, 1, threads)
    return {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_string(),
            "sample": "2 timed steps (after 1 warm-up) of batch %d (of the workload's %d) at %dx%d, fp32, oracle port of the reference "
                      "fwd+BCE+bwd on all host cores" % (args.cpu_batch, wl["batch"], wl["H"], wl["W"]),
            "config1": cpu_config1(threads)}


def run_cfg5(args, wl):
    """BASELINE configs[4]: DenseNet-201 mid-fusion, eval mode (running-statistics BatchNorm), heat-map throughput for batch
    1 ... 128 on one GPU through the graph-captured Evaluator (forward + per-class loss + IoU / accuracy counters)."""
    import gc
    import torch
    from dmmfods_b200 import synthetic
    from dmmfods_b200.model import densenet201_u_lidar
    from dmmfods_b200.trainer import Evaluator
    torch.cuda.set_device(0)
    torch.manual_seed(123)
    model = densenet201_u_lidar(pretrained=False, config=model_cfg(wl["c2"], wl["cb"])).cuda().eval()
    H, W = wl["H"], wl["W"]
    peaks = load_peaks()
    sampler = ClockSampler(0)
    sampler.start()
    sweep, best = [], None
    launches = 0
    for B in wl["sweep"]:
        x1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=7)).cuda()
        x2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=8)).cuda()
        tg = torch.from_numpy(synthetic.target_maps(B, H, W, seed=9)).cuda()
        hx1, hx2, htg = x1.cpu().pin_memory(), x2.cpu().pin_memory(), tg.cpu().pin_memory()
        ev = Evaluator(model, B, H, W)
        for _ in range(max(args.warmup, 3)):
            ev.step(x1, x2, tg)
        torch.cuda.synchronize()
        steps = max(args.steps, 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ev.step(x1, x2, tg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        host = torch.empty(3, dtype=torch.float64).pin_memory()
        e0.record()
        for _ in range(steps):
            out = ev.step(hx1, hx2, htg)
            host.copy_(out["loss_per_class"])
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / steps
        flops = sum(op.flops for op in ev.eng.fwd)
        nl = len(ev.eng.fwd) + 3
        launches += nl * 2 * steps
        row = {"batch": B, "ms_per_step": round(ms, 3), "images_per_s": round(B / ms * 1e3, 1), "e2e_images_per_s": round(B / ms_e2e * 1e3, 1),
               "tflops": round(flops / ms / 1e9, 1), "frac_of_sustained_bf16_peak": round(flops / ms / 1e9 / peaks["tf_sustained"], 3),
               "launches_per_step": nl, "activations_gb": round(ev.eng.mem_bytes / 1e9, 2)}
        sweep.append(row)
        if best is None or row["images_per_s"] > best["images_per_s"]:
            best = row
        if B == wl["sweep"][-1] or B == 32:
            fam = kernel_breakdown(ev.eng, x1, x2, tg, backward=False) if B == 32 else None
            if fam is not None:
                roof, kernels, _ = roofline_from({k: v for k, v in fam.items()}, peaks)
        del ev, x1, x2, tg, hx1, hx2, htg
        model._engines = {}
        gc.collect()              # an engine's launch closures reference its own buffers: cycles, freed by the collector only
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    line = {"metric": "inference heat-map images/sec (DenseNet-201 mid-fusion, eval mode, 640x960)", "value": best["images_per_s"],
            "unit": "images/s", "n_gpus": 1, "steps": max(args.steps, 3), "warmup": max(args.warmup, 3), "ms_per_step": best["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "H": H, "W": W, "best_batch": best["batch"], "cuda_graph": True,
                       "l2": "working set >> 126 MB L2 from batch 2 on (activations column of the sweep)"},
            "e2e": {"value": best["e2e_images_per_s"], "unit": "images/s", "h2d_bytes_per_step": best["batch"] * 7 * H * W * 4,
                    "d2h_bytes_per_step": 24},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernels": kernels, "sweep": sweep, "cpu_baseline": None}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mid", choices=sorted(WORKLOADS))
    ap.add_argument("--api", default="trainer", choices=["trainer", "module"])
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch (default: workload's)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="capture fwd+loss+bwd in a CUDA graph")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch CUDA-event timings of one instrumented step (json)")
    ap.add_argument("--strict-loss-check", action="store_true",
                    help="fail if the per-class loss sums deviate from profiles/bench_loss_golden.json (default: report loss_check.ok)")
    ap.add_argument("--profile-step", action="store_true",
                    help="after the timed region run ONE more step between cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    wl = dict(WORKLOADS[args.workload])
    if args.workload == "cfg5":
        return run_cfg5(args, wl)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dmmfods_b200 import helper, synthetic
    from dmmfods_b200.model import FusedBCEWithLogits, densenet121_u_lidar
    from dmmfods_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.batch:
        wl["batch"] = args.batch
    if args.height:
        wl["H"] = args.height
    if args.width:
        wl["W"] = args.width
    B, H, W = wl["batch"], wl["H"], wl["W"]
    cfg4 = args.workload == "cfg4"
    module_api = args.api == "module"
    if module_api and (cfg4 or world > 1):
        raise SystemExit("--api module is a single-GPU measurement of the mid workload")

    torch.manual_seed(123)
    model = densenet121_u_lidar(pretrained=False, config=model_cfg(wl["c2"], wl["cb"])).cuda()
    seed = 123 + rank
    hx1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=seed)).pin_memory()
    pre = None
    if cfg4:
        # raw per-frame inputs of the step: LiDAR point lists and label boxes; the LiDAR image and the heat-map targets are
        # produced on the GPU inside the step (helper.BatchPreprocessor: dmm_lidar_splat_batched, dmm_heatmap_boxes_batched)
        frames_p = [synthetic.lidar_points(wl["points"], H, W, seed=seed + 2000 + b) for b in range(B)]
        frames_b = [synthetic.boxes(None, H, W, seed=seed + 3000 + b) for b in range(B)]
        pre = helper.BatchPreprocessor(B, H, W, max_points=wl["points"] + 2000, max_boxes=128)
        n_pts, n_box = pre.load(frames_p, frames_b)
        hx2 = htg = None
        trainer = Trainer(model, B, H, W, lr=1e-3, use_graph=bool(args.graph), preprocess=pre.run)
        x1, x2, tg = hx1.cuda(), None, None
    else:
        hx2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=seed + 1000)).pin_memory()
        htg = torch.from_numpy(synthetic.target_maps(B, H, W, seed=seed + 4000)).pin_memory()
        x1, x2, tg = hx1.cuda(), hx2.cuda(), htg.cuda()
        trainer = None if module_api else Trainer(model, B, H, W, lr=1e-3, use_graph=bool(args.graph))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    loss_host = torch.empty(3, dtype=torch.float64).pin_memory()
    if module_api:
        # Agent.py:244-265, literally: the drop-in module, the fused loss module, torch's own Adam
        loss_fn = FusedBCEWithLogits()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)
        model.train()

        def agent_step(image, lidar, ht_map, read_loss=False):
            image, lidar, ht_map = (t.cuda(non_blocking=True) for t in (image, lidar, ht_map))
            prediction = model(image, lidar)
            current_loss = loss_fn(prediction, ht_map)
            loss_per_class = torch.sum(current_loss.detach(), dim=(0, 2, 3))
            opt.zero_grad()
            current_loss.backward(torch.ones_like(current_loss.detach()))
            opt.step()
            if read_loss:
                loss_host.copy_(loss_per_class.double())
        step_dev = lambda: agent_step(x1, x2, tg)
        step_e2e = lambda: agent_step(hx1, hx2, htg, read_loss=True)
        graph_ok = False
    else:
        step_dev = lambda: trainer.step(x1, x2, tg)

        def step_e2e():
            # every step: this step's inputs arrive over PCIe (copy stream, overlapped with the previous step's compute - the
            # prefetch of a training loop), are handed to the engine, and the per-class loss sums are read back
            if cfg4:
                pre.load(frames_p, frames_b)          # host packing of the point lists / boxes + their H2D copies
            cs = trainer.step(hx1, hx2, htg, prefetch_next=(hx1, hx2, htg))
            loss_host.copy_(cs, non_blocking=False)
        graph_ok = trainer.use_graph

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # the loss of the very FIRST optimiser step (seeded parameters and inputs: deterministic up to kernel round-off) is the
    # regression guard of the benched configuration; it also is one more warm-up step
    loss_first = None
    try:
        step_e2e()
        torch.cuda.synchronize()
        loss_first = [float(v) for v in loss_host]
        for _ in range(max(args.warmup, 3)):
            step_dev()
        torch.cuda.synchronize()
    except Exception as e:      # noqa: BLE001  (graph capture problems only: fall back to plain launches, say so)
        if module_api or not trainer.use_graph:
            raise
        sys.stderr.write("CUDA graph capture failed (%s); using plain launches\n" % e)
        trainer.use_graph, trainer.graph, graph_ok = False, None, False
        if loss_first is None:
            step_e2e()
            torch.cuda.synchronize()
            loss_first = [float(v) for v in loss_host]
        for _ in range(max(args.warmup, 3)):
            step_dev()
        torch.cuda.synchronize()

    first = sampler.mark()
    ms = timed(step_dev, args.steps)
    clocks = sampler.stop(first) if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # end to end: pinned host inputs -> H2D -> step -> D2H of the per-class loss sums
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    if cfg4:
        h2d = hx1.numel() * 4 + pre.h2d_bytes
    else:
        h2d = hx1.numel() * 4 + hx2.numel() * 4 + htg.numel() * 4
    loss_per_class = [float(v) for v in loss_host]

    eng = trainer.eng if trainer is not None else model.engine(B, H, W)
    tgt_dev = trainer._static_target if cfg4 else tg
    fam = kernel_breakdown(eng, x1, eng.in2 if cfg4 else x2, tgt_dev, dump=args.dump_ops if rank == 0 else None)
    scatter = None
    if cfg4 and rank == 0:
        # the two pre-processing launches against their compulsory traffic: 4*H*W + 12*N bytes per LiDAR frame,
        # 12*H*W + 20*N_box per heat-map frame; CUDA events, L2 flushed between repetitions by a 256 MB write
        flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
        ts = {"lidar": [], "heat": []}
        for _ in range(7):
            for key, a, b in (("lidar", eng.in2, None), ("heat", None, trainer._static_target)):
                flush.fill_(0.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pre.run(a, b)
                e1.record()
                torch.cuda.synchronize()
                ts[key].append(e0.elapsed_time(e1))
        peaks0 = load_peaks()
        bl, bh = B * 4.0 * H * W + 12.0 * n_pts, B * 12.0 * H * W + 20.0 * n_box
        tl, th = sorted(ts["lidar"])[3], sorted(ts["heat"])[3]
        scatter = {"lidar_splat_batched": {"ms": tl, "algorithmic_bytes": bl, "gbs": bl / tl / 1e6, "frac_of_hbm_peak": bl / tl / 1e6 / peaks0["hbm"],
                                           "points": n_pts},
                   "heatmap_boxes_batched": {"ms": th, "algorithmic_bytes": bh, "gbs": bh / th / 1e6, "frac_of_hbm_peak": bh / th / 1e6 / peaks0["hbm"],
                                             "boxes": n_box}}
    if args.profile_step and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if cfg4:
            pre.run(eng.in2, trainer._static_target)
        eng.forward(x1, eng.in2 if cfg4 else x2)
        eng.loss(tgt_dev)
        eng.backward()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    roof, kernels, total_flops = roofline_from(fam, peaks)
    cpu = cpu_baseline_obj(wl, args) if (world == 1 and not args.no_cpu_baseline and not cfg4) else None
    if module_api:
        launches = (len(eng.fwd) + len(eng.bwd) + 2 + len([s_ for s_ in eng.segments if s_[2]]) + 2) * args.steps
    else:
        launches = (trainer.launches_per_step() + (2 if cfg4 else 0)) * args.steps
    # regression guard on the numerics of the benched configuration (profiles/bench_loss_golden.json, written from a green run):
    #  * "first": per-class loss sums of the FIRST optimiser step - deterministic up to kernel round-off, checked to 2 %
    #    (--strict-loss-check turns it into an assertion);
    #  * the sums after all steps of the run are only reported next to the stored ones: a few dozen Adam steps of this bf16 network
    #    are chaotic (two runs on ONE box: class-0 sums 651 003 and 682 416; different kernel builds: up to 35 % apart after 13 steps)
    loss_check = None
    gfile = os.path.join(ROOT, "profiles", "bench_loss_golden.json")
    key = "%s/%s/B%d/%dx%d/steps%d/warmup%d/gpus%d" % (args.workload, args.api, B, H, W, args.steps, max(args.warmup, 3), world)
    if os.path.isfile(gfile):
        gj = json.load(open(gfile))
        gold, gold_first = gj.get(key), gj.get("%s/%s/B%d/%dx%d/first" % (args.workload, args.api, B, H, W))
        loss_check = {}
        if gold_first is not None:
            rel = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(loss_first, gold_first))
            loss_check.update({"first_step_golden": gold_first, "first_step_max_rel_diff": rel, "ok": rel < 2e-2})
            if args.strict_loss_check:
                assert rel < 2e-2, "first-step per-class loss %s deviates from the stored value %s (rel %.3e)" % (loss_first, gold_first, rel)
        if gold is not None:
            loss_check.update({"last_step_golden": gold,
                               "last_step_max_rel_diff": max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(loss_per_class, gold))})
    metric = {"mid": "train images/sec (mid-fusion Dense-U-Net, DenseNet-121, 640x960)",
              "cfg4": "train images/sec (mid-fusion Dense-U-Net, DenseNet-121, 1280x1920, incl. on-GPU LiDAR projection + heat-map masks)"}[args.workload]
    line = {
        "metric": metric, "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["name"], "api": args.api, "batch_per_gpu": B, "global_batch": B * world, "H": H, "W": W,
                   "parallelism": "dp%d" % world, "l2": "working set >> 126 MB L2 (activations %.1f GB per step)"
                   % (eng.mem_bytes / 1e9), "cuda_graph": bool(graph_ok),
                   "grad_reduce": "sum, %d buckets%s" % (len(eng.segments), ", NCCL captured in the step graph"
                                                         if (trainer is not None and world > 1 and trainer.seg_graphs is None and graph_ok) else "")},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 24,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "kernels": kernels,
        "model_tflops": total_flops / (ms / args.steps * 1e9),
        "loss_per_class": loss_per_class,
        "loss_first_step": loss_first,
        "loss_check": loss_check,
        "cpu_baseline": cpu,
    }
    if scatter is not None:
        line["scatter"] = scatter
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
