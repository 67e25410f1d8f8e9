"""profiles/sass_summary.txt: per-kernel counts of the Blackwell instructions that matter (cuobjdump -sass of the in-tree library).
usage: python scripts/sass_summary.py [tag]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "dmmfods_b200", "libdmmfods_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "ELECT"]
cur, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
rows = []
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[x] for x in KEYS):
        rows.append((demangle(k), c))
path = os.path.join(ROOT, "profiles", "%s_sass_summary.txt" % tag)
with open(path, "w") as fh:
    fh.write("# cuobjdump -sass dmmfods_b200/libdmmfods_b200.so (%s): instruction counts per kernel (kernels without any of these omitted)\n" % ", ".join(sorted(arch)))
    fh.write("# UTCHMMA = tcgen05.mma kind::f16 / tf32, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UTCBAR = tcgen05.commit\n")
    fh.write("%-78s %8s %s\n" % ("kernel", "instrs", " ".join("%9s" % k for k in KEYS)))
    for n, c in sorted(rows, key=lambda r: r[0]):
        fh.write("%-78s %8d %s\n" % (n[:78], c["_total"], " ".join("%9d" % c[k] for k in KEYS)))
    fh.write("%-78s %8d %s\n" % ("TOTAL (%d kernels in the library)" % len(counts), tot["_total"], " ".join("%9d" % tot[k] for k in KEYS)))
print(open(path).read()[-1500:])
