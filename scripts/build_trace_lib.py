#!/usr/bin/env python
"""Build build_trace/libcream_b200_trace.so: the library with -DCREAM_TRACE in attention_bwd.cu (clock64 stamps of the
rows kernel's phases, printed by cream_attn_bwd when CREAM_ATTN_TRACE is set).  Debug tool; the product library is
never built with it.   CREAM_B200_LIB=build_trace/libcream_b200_trace.so CREAM_ATTN_TRACE=1 python scripts/time_attention.py"""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cream_b200 import build as b

abl = int(sys.argv[1]) if len(sys.argv) > 1 else 0      # CREAM_ABL bits (timing ablations; results are wrong)
out = b.ROOT / "build_trace"
out.mkdir(exist_ok=True)
b.build_library()
objs = []
for src in b._sources():
    if src.stem.startswith("attention"):
        o = out / (src.stem + f"_{abl}.o")
        subprocess.run([b.NVCC, *b.NVCC_FLAGS, *b.GENCODE, "-DCREAM_TRACE", f"-DCREAM_ABL={abl}", "-I", str(b.ROOT / "include"), "-c", str(src), "-o", str(o)], check=True)
        objs.append(o)
    else:
        objs.append(b.BUILD / (src.stem + ".o"))
lib = out / ("libcream_b200_trace.so" if abl == 0 else f"libcream_b200_abl{abl}.so")
subprocess.run([b.NVCC, "-shared", *b.GENCODE, "-o", str(lib), *map(str, objs), "-Xlinker", "--no-undefined"], check=True)
print(lib)
