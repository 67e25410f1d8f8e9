"""Turn the raw ncu output under gpurun_out/ into the small text/json summaries committed under profiles/.
usage: python scripts/summarize_profiles.py <tag>      (reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_*.ncu-rep)"""
import csv, glob, io, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

def short(name):
    n = name.split("(")[0].replace("void ", "").replace("dmm::", "")
    return n

lc = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.isfile(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("==")) if r]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
        t = tot[short(r[ki])]
        t[0] += us; t[1] += 1
    total = sum(t[0] for t in tot.values())
    with open(os.path.join(out_dir, tag + "_launch_shares.txt"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none over ONE training step (fwd+loss+bwd, config 3, B=32 640x960)\n")
        fh.write("# cold-cache, serialised launches: compare SHARES with bench.py's CUDA-event shares, not absolutes\n")
        fh.write("# total %.2f ms over %d launches\n" % (total / 1e3, sum(t[1] for t in tot.values())))
        for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            fh.write("%-60s launches %5d  %10.3f ms  share %6.2f%%\n" % (k[:60], t[1], t[0] / 1e3, 100 * t[0] / total))
    print(open(os.path.join(out_dir, tag + "_launch_shares.txt")).read())

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
summ = {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag + "_*.ncu-rep"))):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                d[k] = r[h.index(k)] + " " + units[h.index(k)]
        summ.setdefault(os.path.basename(rep), []).append(d)
with open(os.path.join(out_dir, tag + "_ncu_summary.json"), "w") as fh:
    json.dump(summ, fh, indent=1)
print(json.dumps(summ, indent=1)[:3000])
