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


def build_reference_rpe_index_cuda() -> Path | None:
    """The reference's rpe_ops WITH its CUDA kernels (rpe_index.cpp + rpe_index_cuda.cu, -DWITH_CUDA),
    compiled where the sources lie by nvcc/g++ directly (what rpe_ops/setup.py would do on a GPU box)
    into oracle/_ref/cuda/rpe_index_cpp<ext>.so.  It is the GPU-side reference bench.py times the
    library's rpe_index / fused kernels against; nothing in the product path loads it."""
    src_dir = REF_SRC.parent
    out_dir = HERE / "_ref" / "cuda"
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = out_dir / f"rpe_index_cpp{ext}"
    cu = src_dir / "rpe_index_cuda.cu"
    if not cu.exists():
        return out if out.exists() else None
    if out.exists() and out.stat().st_mtime >= max(cu.stat().st_mtime, REF_SRC.stat().st_mtime):
        return out
    out_dir.mkdir(parents=True, exist_ok=True)
    from torch.utils import cpp_extension as ce
    import torch

    incs = [f"-I{p}" for p in ce.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = Path(torch.__file__).parent / "lib"
    defs = ["-DTORCH_EXTENSION_NAME=rpe_index_cpp", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DWITH_CUDA",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj_cu, obj_cpp = out_dir / "rpe_index_cuda.o", out_dir / "rpe_index.o"
    cmds = [
        [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", *defs, *incs, "-c", str(cu), "-o", str(obj_cu)],
        ["g++", "-O3", "-fopenmp", "-std=c++17", "-fPIC", *defs, *incs, "-c", str(REF_SRC), "-o", str(obj_cpp)],
        ["g++", "-shared", "-fopenmp", str(obj_cu), str(obj_cpp), "-o", str(out), f"-L{libdir}", "-ltorch", "-ltorch_cpu",
         "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart",
         f"-Wl,-rpath,{libdir}"],
    ]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building the reference rpe_ops CUDA extension failed")
    obj_cu.unlink(missing_ok=True)
    obj_cpp.unlink(missing_ok=True)
    return out


if __name__ == "__main__":
    print(build_c_oracle())
    print(build_reference_rpe_index())
    print(build_reference_rpe_index_cuda())
