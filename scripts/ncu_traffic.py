"""Extract DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the
committed `ncu --set full` captures into profiles/ncu_traffic.json (read by bench.py)."""
import csv, io, json, subprocess, sys
from pathlib import Path
out = {}
unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for rep in sorted(Path("gpurun_out").glob("prof_*.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
        rd = float(r[ix["dram__bytes_read.sum"]]) * unit.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = float(r[ix["dram__bytes_write.sum"]]) * unit.get(units[ix["dram__bytes_write.sum"]], 1.0)
        t = float(r[ix["gpu__time_duration.sum"]])
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "time_us": 0.0, "source": rep.name})
        e["launches"] += 1; e["dram_bytes"] += rd + wr
        e["time_us"] += t if units[ix["gpu__time_duration.sum"]] in ("us", "usecond") else t / 1e3
for e in out.values():
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / e["launches"]
    e["time_us_per_launch"] = e.pop("time_us") / e["launches"]
Path("profiles/ncu_traffic.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
