"""Build libpydvs_b200.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

The translation units under ``csrc/`` are compiled in parallel (the fused tile pass is instantiated
per frame dtype / adaptive flag in its own TU) and linked into one shared library with a static CUDA
runtime, so the ``.so`` only needs the driver at load time and shares the process's primary context
(and therefore device pointers and streams) with PyTorch.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libpydvs_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
STAMP = LIB + ".sha256"  # next to the library: travels with it to the GPU box

UNITS = ["dvs_b200.cu", "tile_pass_i16.cu", "tile_pass_i16_adpt.cu", "tile_pass_u8.cu",
         "tile_pass_u8_adpt.cu", "tile_pass_planes.cu", "tile_fast_i16.cu", "tile_fast_u8.cu",
         "tile_fast_i16_adpt.cu", "tile_fast_u8_adpt.cu", "tile_fast_i16_s8.cu", "tile_fast_u8_s8.cu",
         "tile_fast_i16_adpt_s8.cu", "tile_fast_u8_adpt_s8.cu", "dvs_next_rows.cu", "dvs_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "pydvs_b200.h")]
    for name in files:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        with open(path, "rb") as f:
            h.update(os.path.basename(name).encode() + b"\0" + f.read())
    # flags without the absolute include paths: the same tree built elsewhere (the GPU box) has the same digest
    h.update(" ".join(f for f in NVCC_FLAGS if not os.path.isabs(f)).encode())
    return h.hexdigest()


def up_to_date():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=True, ptxas_verbose=False):
    if up_to_date() and not force:
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if ptxas_verbose else []

    def compile_one(unit):
        obj = os.path.join(OBJ_DIR, os.path.splitext(unit)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, unit), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, UNITS))
    if ptxas_verbose:
        for _, err in results:
            sys.stderr.write(err)
    objs = [o for o, _ in results]
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(STAMP, "w") as f:
        f.write(_digest())
    if verbose:
        print("[pydvs_b200] built", LIB)
    return LIB


def ensure_built(verbose=True):
    """Build the library if it is missing or older than its sources (used by the tests, smoke() and
    bench.py so that a fresh checkout works on a box that has nvcc)."""
    if not up_to_date():
        if verbose:
            print("[pydvs_b200] libpydvs_b200.so missing or stale: building with nvcc ...", flush=True)
        build(verbose=verbose)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, ptxas_verbose="-v" in sys.argv)
