#!/usr/bin/env python
"""Generate tests/golden/golden_callers_v1.npz by RUNNING THE REFERENCE's ``stand_alone.py`` unchanged on the compiled,
unmodified reference module (oracle/_ref) with a synthetic webcam (tests/_callers.py): per frame the checksum of what
the script displayed (its render_frame output) and of its spike lists, plus the final reference plane and the last
list of lists.  The GPU test runs the same byte-compiled script on pydvs_b200 under ``install_as_pydvs()``.

    python tests/golden/make_golden_callers.py      # needs /root/reference; the .npz is committed
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_loader  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _callers  # noqa: E402

N_FRAMES = 60


def main():
    assert build_ref.build_callers(), "reference sources not available"
    gs = ref_loader.load()[0]
    out = {}
    for res in (128, 64):
        r = _callers.run_stand_alone(build_ref.load_caller_code("stand_alone"), gs, N_FRAMES, res=res)
        assert r["n"] == N_FRAMES and len(r["shown"]) == N_FRAMES
        lens = np.array([len(s) for s in r["spike_lists"]], dtype=np.int64)
        out["res%d/shown" % res], out["res%d/lists" % res], out["res%d/ref" % res] = r["shown"], r["lists"], r["ref"]
        out["res%d/last_lens" % res] = lens
        out["res%d/last_keys" % res] = np.array([k for s in r["spike_lists"] for k in s], dtype=np.int64)
        print("res", res, "events in the last frame:", int(lens.sum()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_callers_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
