#!/usr/bin/env python
"""Generate tests/golden/golden_vcam_v1.npz by RUNNING THE REFERENCE's own ``VirtualCam`` class (unmodified, from
/root/reference/pydvs/virtual_cam.py, see tests/_vcam_ref.py) on its ``mnist/t10k`` fixtures with a virtual clock:
per behaviour the images it loaded, the frames it served, the image centres and the random draws it made, so that the
GPU ``pydvs_b200.virtual_cam.VirtualCam`` can be replayed against it where /root/reference does not exist.

    python tests/golden/make_golden_vcam.py      # needs /root/reference; the .npz is committed
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _vcam_ref as R  # noqa: E402
from oracle.virtual_cam_oracle import VirtualCamOracle  # noqa: E402


def main():
    out = {}
    for behaviour, masked, n in (("SACCADE", True, 60), ("SACCADE", False, 40), ("TRAVERSE", True, 60), ("FADE", False, 60),
                                 ("ATTENTION", True, 60), ("NONE", False, 25)):
        kw = {}
        if behaviour == "ATTENTION":
            rng = np.random.default_rng(3)
            kw["ref_planes"] = [rng.integers(0, 256, (32, 32)).astype(np.int16) for _ in range(n)]
        run = R.run_reference(behaviour, n, with_mask=masked, **kw)
        # replay through the restatement with a mirror of the reference's generator to record the draws it made
        clock = R.Clock()
        a = run["args"]
        mirror = R.MirrorRng(clock, 12345)
        ora = VirtualCamOracle(run["images"], behaviour, a["fps"], a["image_on_time_ms"], a["inter_off_time_ms"],
                               a["max_saccade_distance"], a["frames_per_microsaccade"], a["max_cycles"], run["mask"],
                               a["background_gray"], 0, mirror, frames_per_saccade=a["frames_per_saccade"])
        tied_at = []
        for i in range(n):
            clock.now = ora.now
            before = ora.tied_frames
            ora.read(kw["ref_planes"][i] if kw else None)
            if ora.tied_frames != before:
                tied_at.append(i)
            if (ora.cx, ora.cy) != run["centers"][i] and ora.showing:
                ora.cx, ora.cy = run["centers"][i]
                ora.current = run["frames"][i].copy()
        name = behaviour + ("_mask" if masked else "")
        out[name + "/frames"] = np.stack(run["frames"])
        out[name + "/valid"] = np.array(run["valid"])
        out[name + "/centers"] = np.array(run["centers"], dtype=np.int32)
        out[name + "/steps"] = np.array(mirror.steps, dtype=np.int32).reshape(-1, 2)
        out[name + "/picks"] = np.array(mirror.picks, dtype=np.int32)
        out[name + "/tied_at"] = np.array(tied_at, dtype=np.int32)
        if kw:
            out[name + "/ref_planes"] = np.stack(kw["ref_planes"])
    out["images"] = np.stack(run["images"])
    out["mask"] = R.run_reference("SACCADE", 1, with_mask=True)["mask"]
    out["args"] = np.array([3000, 4, 1, 2, 2, 5, 1, 0], dtype=np.int64)  # fps, on_ms, off_ms, max_dist, fp_usacc, fp_sacc, cycles, bg
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_vcam_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
