"""Per-kernel share of a step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki]
    short = re.sub(r"^.*::", "", name.split("(")[0])
    if "gemm_bf16_kernel" in name:
        m = re.search(r"gemm_bf16_kernel<(\d+), *(\w+)>", name)
        short = "gemm_bf16_kernel<epi %s, cta_pair %s>" % (m.group(1), m.group(2)) if m else "gemm_bf16_kernel"
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
    agg[short][0] += 1
    agg[short][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.1f us (cold-cache, serialised: shares are meaningful, absolutes are not)" % (sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-56s n=%4d  %9.1f us  %5.1f%%  avg %7.1f us" % (k[:56], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
