"""Print the key numbers of one or more bench.py JSON lines: python scripts/summarize.py file.json [...]"""
import json
import sys
from pathlib import Path

KEYS = ["value", "ms_per_step", "e2e", "gpu_launches", "host_enqueue_ms_per_step", "clocks", "roofline", "attention",
        "attention_bwd", "byte_movers", "gpu_time_share", "kernel_table_cupti", "model_tflops_per_gpu", "gpu_reference", "speedup_vs_gpu_reference",
        "cpu_baseline"]
for f in sys.argv[1:]:
    try:
        d = json.loads(Path(f).read_text().strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    print("==", f, "|", d.get("config", {}).get("baseline_config"), d.get("config", {}).get("runtime"), "n_gpus", d.get("n_gpus"))
    for k in KEYS:
        if k in d:
            print("  ", k, ":", d[k])
