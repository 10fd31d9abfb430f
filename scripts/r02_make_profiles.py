"""Turn the round-2 ncu captures in gpurun_out/ into the tracked text summaries under profiles/ (offline, no GPU).

    python scripts/r02_make_profiles.py

  prof_r02_*.ncu-rep      -> profiles/r02_ncu_<class>.txt   (per launch: duration, DRAM bytes, DRAM %, tensor pipe %,
                             issue %, registers, occupancy limits; top stall sites of the first launch of every kernel)
  r02_launches.csv        -> profiles/r02_launches.txt      (per-kernel share of the bench step)
  all captures            -> profiles/ncu_traffic.json      (dram bytes per launch, read by bench.py)
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

OUT = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("profiles")
SRC = Path("gpurun_out")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("void ", "").replace("cb::(anonymous namespace)::", "").replace("cb::<unnamed>::", "").replace("unnamed>::", "")
    m = re.match(r"([\w:<>, ]+?)\(", name)
    return (m.group(1) if m else name)[:70]


def summarize(rep: Path, traffic: dict) -> str:
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return f"{rep.name}: empty\n"
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep.name}: `ncu --set full --clock-control none --import-source on` (single launches, cold caches, serialised)"]
    for r in rows[2:]:
        name = short(r[ix["Kernel Name"]])
        vals = {}
        for w in WANT:
            if w in ix:
                vals[w] = (r[ix[w]], units[ix[w]])
        def num(key):
            v, u = vals.get(key, ("0", ""))
            try:
                return float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                return 0.0
        t = num("gpu__time_duration.sum")
        t_us = t if vals["gpu__time_duration.sum"][1] in ("us", "usecond") else t / 1e3
        dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        e = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0, "source": rep.name})
        e["launches"] += 1
        e["dram_bytes"] += dram
        e["time_us"] += t_us
        lines.append(f"{name}\n    time {t_us:8.1f} us | DRAM {dram / 1e6:8.1f} MB ({dram / max(t_us, 1e-9) / 1e3:6.0f} GB/s, "
                     f"{vals.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', ('?', ''))[0]} % of peak) | tensor pipe "
                     f"{vals.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', ('?', ''))[0]} % | issue "
                     f"{vals.get('smsp__issue_active.avg.pct_of_peak_sustained_active', ('?', ''))[0]} % | warps active "
                     f"{vals.get('sm__warps_active.avg.pct_of_peak_sustained_active', ('?', ''))[0]} % | regs "
                     f"{vals.get('launch__registers_per_thread', ('?', ''))[0]} | grid {vals.get('launch__grid_size', ('?', ''))[0]} x "
                     f"{vals.get('launch__block_size', ('?', ''))[0]} | L2 hit {vals.get('lts__t_sector_hit_rate.pct', ('?', ''))[0]} % | "
                     f"smem bank conflicts {vals.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', ('?', ''))[0]}")
    # top stall sites per distinct kernel (source page)
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    ks, cur = [], None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = [r]
            ks.append(cur)
        elif cur is not None:
            cur.append(r)
    seen = set()
    f = lambda x: float(x) if x not in ("",) else 0.0
    for k in ks:
        kname = short(k[0][1]) if len(k[0]) > 1 else "?"
        if kname in seen or len(k) < 3:
            continue
        seen.add(kname)
        h = k[1]
        hx = {c: i for i, c in enumerate(h)}
        if "# Samples" not in hx:
            continue
        data = [r for r in k[2:] if len(r) >= len(h)]
        tot = sum(f(r[hx["# Samples"]]) for r in data)
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = {s_: sum(f(r[hx[s_]]) for r in data) for s_ in stalls}
        lines.append(f"  stalls of {kname} ({int(tot)} samples): " + ", ".join(f"{s_[6:]} {int(v)}" for s_, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
        for r in sorted(data, key=lambda r: -f(r[hx["# Samples"]]))[:8]:
            top = max(((s_, f(r[hx[s_]])) for s_ in stalls), key=lambda kv: kv[1])
            lines.append(f"      {int(f(r[hx['# Samples']])):5d}  {r[hx['Source']][:70]:70s} {top[0][6:]}")
    return "\n".join(lines) + "\n"


def launch_shares(path: Path) -> str:
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki]
        s_ = re.sub(r"^.*::", "", name.split("(")[0])
        if "gemm_bf16_kernel" in name:
            m = re.search(r"gemm_bf16_kernel<(\d+), *(\w+)>", name)
            s_ = "gemm_bf16_kernel<epi %s, cta_pair %s>" % (m.group(1), m.group(2)) if m else "gemm_bf16_kernel"
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg[s_][0] += 1
        agg[s_][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["# per-kernel share of the bench step: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of",
           "# `python bench.py --steps 2 --warmup 3` (a window of launches; cold-cache and serialised: shares are meaningful, absolutes are not)",
           "launches %d, total %.1f us" % (sum(v[0] for v in agg.values()), tot)]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
        out.append("%-58s n=%4d  %9.1f us  %5.1f%%  avg %7.1f us" % (k[:58], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    return "\n".join(out) + "\n"


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    traffic = {}
    for rep in sorted(SRC.glob("prof_r02_*.ncu-rep")):
        (OUT / ("r02_ncu_" + rep.stem.replace("prof_r02_", "") + ".txt")).write_text(summarize(rep, traffic))
        print("wrote", rep.stem)
    for e in traffic.values():
        e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
        e["time_us_per_launch"] = e.pop("time_us") / e["launches"]
    if traffic:
        (OUT / "ncu_traffic.json").write_text(json.dumps(traffic, indent=1))
    if (SRC / "r02_launches.csv").exists():
        (OUT / "r02_launches.txt").write_text(launch_shares(SRC / "r02_launches.csv"))
        print("wrote launch shares")


if __name__ == "__main__":
    sys.exit(main())
