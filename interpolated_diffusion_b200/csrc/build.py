"""Build libidb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

    python -m interpolated_diffusion_b200.csrc.build [-v] [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(PKG, "libidb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]
# files whose fp32 arithmetic must match the reference's op-by-op rounding: never contract to FMA
NO_FMAD = {"keyframes.cu", "elementwise.cu"}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(path: str, flags) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    for dep in [path] + sorted(
        os.path.join(d, f) for d in (HERE, os.path.join(ROOT, "include")) for f in os.listdir(d) if f.endswith((".cuh", ".h"))
    ):
        with open(dep, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    sources = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))
    objs, jobs = [], []
    extra = os.environ.get("IDB200_NVCC_EXTRA", "").split()      # dev: e.g. -DIDB200_WAIT_LIMIT_CYCLES=400000000LL
    for src in sources:
        flags = list(NVCC_FLAGS) + extra + (["-fmad=false"] if src in NO_FMAD else [])
        path = os.path.join(HERE, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest(path, flags)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((src, [nvcc] + flags + ["-c", path, "-o", obj], stamp, dig))

    def run(job):
        src, cmd, stamp, dig = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as fh:
            fh.write(dig)
        with open(os.path.join(OBJ, src[:-3] + ".ptxas.txt"), "w") as fh:
            fh.write(r.stderr)
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, log in ex.map(run, jobs):
            if verbose:
                print(f"--- {src}\n{log}")
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
