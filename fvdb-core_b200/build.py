"""Builds libfvdbconv.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python fvdb-core_b200/build.py [--force]

The shared library lands next to the Python package (fvdb-core_b200/fvdb/libfvdbconv.so) so that the
gpurun snapshot carries it to the GPU box.  Object files are cached under fvdb-core_b200/build/.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "fvdb" / "libfvdbconv.so"
OBJ = HERE / "build"
SOURCES = ["abi.cu", "grid_build.cu", "kmap.cu", "weights.cu", "conv.cu", "conv_simt.cu", "conv_tc.cu", "conv_tc_wgrad.cu", "conv_tc_bwd.cu", "conv_tc_ts.cu", "norm.cu", "pool.cu", "grid_morph.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "fvdbconv.h"]

    def compile_one(name: str):
        src, obj = CSRC / name, OBJ / (name + ".o")
        if force or _stale(obj, [src, *headers]):
            cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    if force or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", str(OUT), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
