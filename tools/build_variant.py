#!/usr/bin/env python
"""Build a tuning variant of the CUDA library next to the real one:

    python tools/build_variant.py <name> [-DMACRO=value ...]   ->  pydvs_b200/libpydvs_b200_<name>.so
    PYDVS_B200_LIB=pydvs_b200/libpydvs_b200_<name>.so python bench.py ...

Same translation units and flags as pydvs_b200/build.py plus the given macros (FAST2_MINB, FAST2_MINB_ADPT, ...).
Variants are experiments: they are git-ignored and never loaded unless PYDVS_B200_LIB names them."""
import importlib.util
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "pydvs_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)


def main():
    name, extra = sys.argv[1], sys.argv[2:]
    obj_dir = os.path.join(b.OBJ_DIR, "variant_" + name)
    os.makedirs(obj_dir, exist_ok=True)
    out = os.path.join(b.HERE, "libpydvs_b200_%s.so" % name)

    def one(unit):
        obj = os.path.join(obj_dir, os.path.splitext(unit)[0] + ".o")
        r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + extra + ["-c", os.path.join(b.CSRC, unit), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr)
        return obj
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        objs = list(pool.map(one, b.UNITS))
    r = subprocess.run([b._nvcc(), "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs,
                       capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)
    import shutil
    shutil.rmtree(obj_dir, ignore_errors=True)   # objects are large (-lineinfo) and everything under the repo travels to the GPU box
    print("built", out)


if __name__ == "__main__":
    main()
