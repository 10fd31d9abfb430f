"""Summarise gpurun_out: native test verdicts, pytest groups, bench line, ncu launch shares."""
import collections, csv, json, re, sys
from pathlib import Path
out = Path("gpurun_out")
if (out / "native.log").exists():
    for l in (out / "native.log").read_text().splitlines():
        if re.search(r"FAIL|TFLOP|GB/s|exit [^0]", l): print(l)
if (out / "gpu_tests.log").exists():
    for l in (out / "gpu_tests.log").read_text().splitlines():
        if re.search(r"passed|failed|\[exit|^FAILED", l): print(l)
if (out / "bench.json").exists():
    try:
        d = json.loads((out / "bench.json").read_text().strip().splitlines()[-1])
        for k in ["value", "ms_per_step", "e2e", "gpu_launches", "clocks", "roofline", "attention", "model_tflops", "cpu_baseline"]:
            print(k, ":", d.get(k))
    except Exception as e:
        print("bench.json unreadable", e, (out / "bench.err").read_text()[-2000:])
if (out / "launches.csv").exists() and "--launches" in sys.argv:
    rows = list(csv.reader(open(out / "launches.csv")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; ix = {n: i for i, n in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:60]
        v = float(r[ix["Metric Value"]].replace(",", "")); unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("total us", round(tot, 1), "launches", sum(v[0] for v in agg.values()))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        print("%6d %10.1f us %5.1f%%  avg %7.1f  %s" % (v[0], v[1], 100 * v[1] / tot, v[1] / v[0], k))
