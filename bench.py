#!/usr/bin/env python
"""bench.py - train images/s of the mid-fusion Dense-U-Net hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun ... bench.py --gpus N ...        (N > 1: one rank per GPU, NCCL)

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

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS["mid"]
    threads = os.cpu_count() or 1
    sample = args.cpu_batch
    ips, t = cpu_reference_step(wl, sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "train images/sec (mid-fusion Dense-U-Net, DenseNet-121, 640x960)", "value": ips,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_step": sample, "H": wl["H"], "W": wl["W"],
                   "note": "reference algorithm on host cores (oracle port of Dense_U_Net_lidar fwd+BCE+bwd, torch CPU ops)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d step(s) of batch %d at %dx%d" % (args.steps, sample, wl["H"], wl["W"])},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
def kernel_breakdown(trainer, x1, x2, tgt, dump=None):
    """one instrumented step: CUDA events around every launch, grouped by kernel family."""
    import torch
    eng = trainer.eng
    recs = []
    orig_run = eng._run

    def timed_run(program):
        import ctypes as C
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
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
        eng.forward(x1, x2)
        eng.loss(tgt)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="override the per-GPU batch (default: workload's)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=4, help="batch of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="capture fwd+loss+bwd in a CUDA graph (single GPU)")
    ap.add_argument("--dump-ops", default=None, help="write the per-launch CUDA-event timings of one instrumented step (json)")
    ap.add_argument("--profile-step", action="store_true",
                    help="after the timed region run ONE more step between cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dmmfods_b200 import synthetic
    from dmmfods_b200.model import densenet121_u_lidar
    from dmmfods_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = dict(WORKLOADS["mid"])
    if args.batch:
        wl["batch"] = args.batch
    if args.height:
        wl["H"] = args.height
    if args.width:
        wl["W"] = args.width
    B, H, W = wl["batch"], wl["H"], wl["W"]

    torch.manual_seed(123)
    model = densenet121_u_lidar(pretrained=False, config=model_cfg(wl["c2"], wl["cb"])).cuda()
    trainer = Trainer(model, B, H, W, lr=1e-3, use_graph=bool(args.graph))
    seed = 123 + rank
    hx1 = torch.from_numpy(synthetic.rgb_image(B, H, W, seed=seed)).pin_memory()
    hx2 = torch.from_numpy(synthetic.lidar_image(B, H, W, seed=seed + 1000)).pin_memory()
    htg = torch.from_numpy(synthetic.target_maps(B, H, W, seed=seed + 4000)).pin_memory()
    x1, x2, tg = hx1.cuda(), hx2.cuda(), htg.cuda()

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

    graph_ok = trainer.use_graph
    try:
        for _ in range(max(args.warmup, 3)):
            trainer.step(x1, x2, tg)
        torch.cuda.synchronize()
    except Exception as e:      # noqa: BLE001  (graph capture problems only: fall back to plain launches, say so)
        if not trainer.use_graph:
            raise
        sys.stderr.write("CUDA graph capture failed (%s); using plain launches\n" % e)
        trainer.use_graph, trainer.graph, graph_ok = False, None, False
        for _ in range(max(args.warmup, 3)):
            trainer.step(x1, x2, tg)
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: trainer.step(x1, x2, tg), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    # end to end: pinned host inputs -> H2D -> step -> D2H of the per-class loss sums
    loss_host = torch.empty(3, dtype=torch.float64).pin_memory()

    def e2e_step():
        # every step: this step's inputs arrive over PCIe (copy stream, overlapped with the previous step's compute - the
        # prefetch of a training loop), are handed to the engine, and the per-class loss sums are read back
        cs = trainer.step(hx1, hx2, htg, prefetch_next=(hx1, hx2, htg))
        loss_host.copy_(cs, non_blocking=False)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = hx1.numel() * 4 + hx2.numel() * 4 + htg.numel() * 4
    loss_per_class = [float(v) for v in loss_host]

    fam = kernel_breakdown(trainer, x1, x2, tg, dump=args.dump_ops if rank == 0 else None)
    if args.profile_step and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer.eng.forward(x1, x2)
        trainer.eng.loss(tg)
        trainer.eng.backward()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
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
            roof["traffic"] = json.load(open(tfile)).get(dom)
        except Exception:
            pass
    total_flops = sum(v["flops"] for v in fam.values())
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ips, t = cpu_reference_step(wl, args.cpu_batch, 2, 1, threads)
        cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": "2 timed steps (after 1 warm-up) of batch %d at %dx%d, fp32, oracle port of the reference fwd+BCE+bwd"
                         % (args.cpu_batch, H, W)}
    line = {
        "metric": "train images/sec (mid-fusion Dense-U-Net, DenseNet-121, 640x960)", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_gpu": B, "global_batch": B * world, "H": H, "W": W,
                   "parallelism": "dp%d" % world, "l2": "working set >> 126 MB L2 (activations %.1f GB per step)"
                   % (trainer.eng.mem_bytes / 1e9), "cuda_graph": bool(graph_ok), "grad_reduce": "sum, %d buckets" % len(trainer.eng.segments)},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 24,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": trainer.launches_per_step() * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "kernels": kernels,
        "model_tflops": total_flops / (ms / args.steps * 1e9),
        "loss_per_class": loss_per_class,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
