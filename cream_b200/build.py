"""In-tree nvcc build of libcream_b200.so (sm_100a only) and the helper binaries.

The shared library is the drop-in boundary (C ABI, see include/cream_b200.h); it is
built here on CPU (nvcc cross-compiles) and travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "cream_b200" / "csrc"
BUILD = ROOT / "build"
LIB = ROOT / "cream_b200" / "libcream_b200.so"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(path: Path, extra: str) -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    h.update(path.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + [ROOT / "include" / "cream_b200.h"]):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile_one(src: Path, verbose: bool) -> Path:
    BUILD.mkdir(exist_ok=True)
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha")
    flags = NVCC_FLAGS + GENCODE + os.environ.get("CREAM_B200_EXTRA_NVCC_FLAGS", "").split()
    dig = _digest(src, " ".join(flags))
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [NVCC, *flags, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        sys.stderr.write(r.stderr)
    stamp.write_text(dig)
    return obj


def build_library(verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link cream_b200/libcream_b200.so."""
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    cmd = [NVCC, "-shared", *GENCODE, "-o", str(LIB), *map(str, objs), "-Xlinker", "--no-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


def build_native_tests() -> list[Path]:
    """Stand-alone C++ drivers over the C ABI (no torch): tests/native/*.cu."""
    lib = build_library()
    outs = []
    for src in sorted((ROOT / "tests" / "native").glob("*.cu")):
        exe = BUILD / src.stem
        if exe.exists() and exe.stat().st_mtime >= max(src.stat().st_mtime, lib.stat().st_mtime):
            outs.append(exe)
            continue
        cmd = [NVCC, "-O2", "-std=c++17", *GENCODE, "-I", str(ROOT / "include"), str(src), "-o", str(exe),
               "-L", str(lib.parent), "-lcream_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../cream_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed for {src.name}")
        outs.append(exe)
    return outs


if __name__ == "__main__":
    v = "-v" in sys.argv
    print(build_library(verbose=v))
    if "--native-tests" in sys.argv:
        for e in build_native_tests():
            print(e)
