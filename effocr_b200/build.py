"""In-tree nvcc build of libeffocr_b200.so (sm_100a only).

`python -m effocr_b200.build` or `__graft_entry__.build()`.  Objects are rebuilt only when a
source or header is newer; translation units compile in parallel.  The .so is git-ignored but
travels with the tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
OBJDIR = PKG / "build"
LIB = PKG / "libeffocr_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-DEFFOCR_BUILDING=1",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; effocr_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _newest_header() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    OBJDIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    hdr_time = _newest_header()
    jobs = []
    objs = []
    for src in sources:
        obj = OBJDIR / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        # EFFOCR_NVCC_EXTRA: extra nvcc flags for experiments, e.g. "-DEFFOCR_ATT_ABLATION" (tools/att_ablate.py); rebuild with --force
        # EFFOCR_AB=1: also compile the A/B kernel variants (earlier attention kernels, direct-store GEMM epilogues, ...)
        ab = ["-DEFFOCR_AB"] if os.environ.get("EFFOCR_AB") == "1" else []
        cmd = [nvcc, *NVCC_FLAGS, *ab, *os.environ.get("EFFOCR_NVCC_EXTRA", "").split(), "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"compiled {name}", flush=True)
    if jobs or not LIB.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
