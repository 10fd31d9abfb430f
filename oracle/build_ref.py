"""Build recipe for oracle/_ref: the reference's OWN rpe_index CPU operator.

Compiles /root/reference/iRPE/DeiT-with-iRPE/rpe_ops/rpe_index.cpp where it lies (a single
source file, pybind11 + ATen; no reference build system is run, no reference source is
copied) into oracle/_ref/rpe_index_cpp<ext>.so.  The output directory is git-ignored but
travels to the GPU box; there it is only ever *loaded* (tests + bench cpu_baseline).

Also compiles oracle/c/*.c (the plain-C restatement) into oracle/_build/liboracle_c.so.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("CREAM_REFERENCE", "/root/reference")) / "iRPE" / "DeiT-with-iRPE" / "rpe_ops" / "rpe_index.cpp"


def build_c_oracle() -> Path:
    out_dir = HERE / "_build"
    out_dir.mkdir(exist_ok=True)
    srcs = sorted((HERE / "c").glob("*.c"))
    out = out_dir / "liboracle_c.so"
    if out.exists() and all(out.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return out
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", str(out), *map(str, srcs), "-lm"]
    subprocess.run(cmd, check=True)
    return out


def build_reference_rpe_index() -> Path | None:
    """g++ on the reference source directly; returns None when /root/reference is absent
    (GPU box: the prebuilt .so is used)."""
    out_dir = HERE / "_ref"
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = out_dir / f"rpe_index_cpp{ext}"
    if not REF_SRC.exists():
        return out if out.exists() else None
    if out.exists() and out.stat().st_mtime >= REF_SRC.stat().st_mtime:
        return out
    out_dir.mkdir(exist_ok=True)
    from torch.utils import cpp_extension as ce
    import torch

    incs = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = Path(torch.__file__).parent / "lib"
    cmd = ["g++", "-O3", "-fopenmp", "-std=c++17", "-fPIC", "-shared",
           "-DTORCH_EXTENSION_NAME=rpe_index_cpp", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *incs, str(REF_SRC), "-o", str(out),
           f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building the reference rpe_index.cpp failed")
    return out


if __name__ == "__main__":
    print(build_c_oracle())
    print(build_reference_rpe_index())
