#!/usr/bin/env python
"""Generate tests/golden/golden_float_v1.npz by RUNNING THE REFERENCE's own ``DVSOperator::operator()``
(``pydvs/cpp/dvs_op.cpp:4-33``, compiled unchanged by oracle/build_ref.py against the stand-in OpenCV headers of
oracle/opencv_stub/): seeded frame sequences in, the diff / ref / thr planes after every 10th frame out.

State initialisation follows ``pydvs/cpp/dvs_emu.hpp:91-104`` (``ref = 0``, ``thr = thr_init``), parameters
``pydvs/cpp/test.cpp:40-43`` (relax 0.999, up 1.5, down 0.99, thr 10) plus two other settings.

    python tests/golden/make_golden_float.py     # needs /root/reference; the .npz is committed
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_loader  # noqa: E402

CASES = {
    # name: (H, W, frames, thr_init, relax, up, down, kind)
    "test_cpp_defaults": (12, 20, 100, 10.0, 0.999, 1.5, 0.99, "uniform"),
    "leaky_fast": (8, 16, 60, 4.0, 0.9, 1.25, 0.9, "bar"),
    "no_relax": (9, 13, 60, 25.0, 1.0, 2.0, 0.5, "uniform"),   # thresholds under- and overflow towards 0 / large
}


def frames_of(kind, H, W, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(0, 255, (n, H, W)).astype(np.float32)
    out = np.full((n, H, W), 64.0, np.float32)
    for t in range(n):
        out[t, :, (t * 2) % W:(t * 2) % W + 3] = 200.0
    return out + rng.integers(-8, 9, (n, H, W)).astype(np.float32)


def main():
    if not build_ref.build_dvs_op():
        raise SystemExit("reference sources not available")
    op = ref_loader.load_dvs_op()
    out = {}
    for i, (name, (H, W, n, thr0, relax, up, down, kind)) in enumerate(CASES.items()):
        frames = frames_of(kind, H, W, n, 100 + i)
        ref = np.zeros((H, W), np.float32)
        thr = np.full((H, W), thr0, np.float32)
        diff = np.zeros((H, W), np.float32)
        snaps = []
        for t in range(n):
            op(np.ascontiguousarray(frames[t]), diff, ref, thr, relax, up, down)
            if (t + 1) % 10 == 0:
                snaps.append(np.stack([diff, ref, thr]))
        out[name + "/frames"] = frames
        out[name + "/params"] = np.array([thr0, relax, up, down], np.float32)
        out[name + "/snaps"] = np.stack(snaps)   # [n / 10, 3, H, W]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_float_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
