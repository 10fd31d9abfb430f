"""Print the headline metrics and the top stall sites of an .ncu-rep (offline, no GPU)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [i for i, h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print({hdr[i][:48]: r[i][:70] for i in idx})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
ks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = [r]; ks.append(cur)
    elif cur is not None:
        cur.append(r)
f = lambda x: float(x) if x not in ("",) else 0.0
for k in ks[:1]:
    hdr = k[1]; ix = {h: i for i, h in enumerate(hdr)}; n = len(hdr)
    data = [r for r in k[2:] if len(r) >= n]
    tot = sum(f(r[ix["# Samples"]]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(f(r[ix[s]]) for r in data) for s in stalls}
    print("samples", tot, [(s, int(v)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]])
    for r in sorted(data, key=lambda r: -f(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
        st = sorted(((s, f(r[ix[s]])) for s in stalls), key=lambda kv: -kv[1])[:1]
        print("   ", int(f(r[ix["# Samples"]])), r[ix["Source"]][:80], st)
