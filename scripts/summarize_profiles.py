"""Turn the raw ncu output under gpurun_out/ into the small text/json summaries committed under profiles/.
usage: python scripts/summarize_profiles.py <tag>
  gpurun_out/<tag>_launches.csv  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over ONE training step
  gpurun_out/<tag>_ops.json      bench.py --dump-ops of the same command (op name / kernel family per launch, in launch order)
  gpurun_out/<tag>_*.ncu-rep     ncu --set full captures"""
import csv, glob, io, json, os, subprocess, sys, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    return name.split("(")[0].replace("void ", "").replace("dmm::", "")


def to_us(v, u):
    return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6)


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


lc = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
launches = collections.OrderedDict()     # id -> dict(name, us, rd, wr)
if os.path.isfile(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("==")) if r]
    h = rows[0]
    ii, ki, mi, vi, ui = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        d = launches.setdefault(int(r[ii]), dict(name=short(r[ki]), us=0.0, rd=0.0, wr=0.0))
        if r[mi] == "gpu__time_duration.sum":
            d["us"] = to_us(v, r[ui])
        elif r[mi] == "dram__bytes_read.sum":
            d["rd"] = to_bytes(v, r[ui])
        elif r[mi] == "dram__bytes_write.sum":
            d["wr"] = to_bytes(v, r[ui])
    tot = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for d in launches.values():
        t = tot[d["name"]]
        t[0] += d["us"]; t[1] += 1; t[2] += d["rd"] + d["wr"]
    total = sum(t[0] for t in tot.values())
    with open(os.path.join(out_dir, tag + "_launch_shares.txt"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_{read,write}.sum --clock-control none over ONE training step\n")
        fh.write("# (fwd + loss + bwd, BASELINE config 3: mid-fusion DenseNet-121, B=32, 640x960); cold-cache, serialised launches:\n")
        fh.write("# compare SHARES with the CUDA-event shares of bench.py (\"kernels\" object), not absolutes\n")
        fh.write("# total %.2f ms over %d launches, %.1f GB of DRAM traffic\n" % (total / 1e3, sum(t[1] for t in tot.values()),
                                                                                 sum(t[2] for t in tot.values()) / 1e9))
        for k, t in sorted(tot.items(), key=lambda kv: -kv[1][0]):
            fh.write("%-58s launches %5d  %9.3f ms  share %6.2f%%  dram %8.2f GB  %6.0f GB/s\n"
                     % (k[:58], t[1], t[0] / 1e3, 100 * t[0] / total, t[2] / 1e9, t[2] / max(t[0], 1e-9) / 1e3))
    print(open(os.path.join(out_dir, tag + "_launch_shares.txt")).read())

# ---- per kernel family: measured DRAM traffic vs the algorithmic bytes bench.py uses for the roofline --------------------------
oj = os.path.join(ROOT, "gpurun_out", tag + "_ops.json")
if launches and os.path.isfile(oj):
    ops = json.load(open(oj))
    ours = ("igemm", "wgrad_kernel", "bn_", "head_", "im2col", "nchw_", "rows_f32", "grad_gather", "dlogits_", "unfold_")
    skip = ("pack_weights", "unpack_wgrad", "bce_logits", "adam", "at::")
    seq = [d for d in launches.values() if d["name"].startswith(ours) and not d["name"].startswith(skip)]
    # an op = one launch, except the head input statistics (two nchw_stats launches for two input tensors)
    fam = collections.defaultdict(lambda: dict(launches=0, us=0.0, dram=0.0, alg_bytes=0.0, flops=0.0))
    i = 0
    ok = True
    for op in ops:
        # ops that launch two kernels: the head input statistics (two input tensors) and the head BN backward reduce
        n = 2 if (op["kind"] == "nchw_stats" or (op["kind"] == "head_input_bwd" and op["name"].endswith("reduce"))) else 1
        for _ in range(n):
            if i >= len(seq):
                ok = False
                break
            f = fam[op["kind"]]
            f["launches"] += 1; f["us"] += seq[i]["us"]; f["dram"] += seq[i]["rd"] + seq[i]["wr"]
            i += 1
        f = fam[op["kind"]]
        f["alg_bytes"] += op["bytes"]; f["flops"] += op["flops"]
    if ok and i == len(seq):
        traffic = {k: v["dram"] / v["launches"] for k, v in fam.items()}
        json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
        with open(os.path.join(out_dir, tag + "_family_traffic.txt"), "w") as fh:
            fh.write("# per kernel family of one training step: ncu DRAM traffic (read+write) vs the algorithmic bytes of bench.py\n")
            fh.write("%-22s %8s %10s %12s %12s %8s\n" % ("family", "launches", "ncu ms", "dram GB", "algorithmic GB", "ratio"))
            for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
                fh.write("%-22s %8d %10.2f %12.2f %12.2f %8.2f\n" % (k, v["launches"], v["us"] / 1e3, v["dram"] / 1e9, v["alg_bytes"] / 1e9,
                                                                   v["dram"] / max(v["alg_bytes"], 1.0)))
        print(open(os.path.join(out_dir, tag + "_family_traffic.txt")).read())
    else:
        print("launch list (%d of ours) does not line up with the op list (%d ops): traffic.json not written" % (len(seq), len(ops)))

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
for k, v in summ.items():
    print(k, {a: b for a, b in v[0].items() if a in ("kernel", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                                     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum")})
