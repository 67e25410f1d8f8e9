"""summarise a --dump-ops json: totals per kernel family and the slowest launches."""
import json
import sys
rows = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tot = sum(r["ms"] for r in rows)
print("total %.2f ms over %d launches" % (tot, len(rows)))
fam = {}
for r in rows:
    f = fam.setdefault(r["kind"], [0.0, 0.0, 0.0, 0])
    f[0] += r["ms"]; f[1] += r["flops"]; f[2] += r["bytes"]; f[3] += 1
for k, f in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print("  %-20s n=%4d %8.2f ms %5.1f%%  %7.1f TF/s %7.0f GB/s" % (k, f[3], f[0], 100 * f[0] / tot, f[1] / (f[0] * 1e9), f[2] / (f[0] * 1e6)))
rows.sort(key=lambda r: -r["ms"])
for r in rows[:top]:
    print("%-62s %-18s %8.3f ms  %7.1f TF/s %8.1f GB/s" % (r["name"][:62], r["kind"], r["ms"], r["flops"] / (r["ms"] * 1e9) if r["flops"] else 0, r["bytes"] / (r["ms"] * 1e6)))
