"""CPU: the scalar VirtualCam restatement (oracle/virtual_cam_oracle.py) against the reference's OWN ``VirtualCam``
class, run unmodified from /root/reference on its ``mnist/t10k`` fixtures with a virtual clock (tests/_vcam_ref.py).
This is what pins the restatement the GPU ``pydvs_b200.virtual_cam.VirtualCam`` is compared with.  Skipped where the
reference tree is absent (the GPU box): there ``tests/golden/golden_vcam_v1.npz`` (recorded from these runs) stands in."""
import os

import numpy as np
import pytest

from oracle.virtual_cam_oracle import VirtualCamOracle, attention_target
import _vcam_ref as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_vcam_v1.npz")


def _oracle_for(run, rng):
    a = run["args"]
    return VirtualCamOracle(run["images"], a["behaviour"], a["fps"], a["image_on_time_ms"], a["inter_off_time_ms"],
                            a["max_saccade_distance"], a["frames_per_microsaccade"], a["max_cycles"], run["mask"],
                            a["background_gray"], 0, rng, frames_per_saccade=a["frames_per_saccade"])


@pytest.fixture(scope="module")
def ref_ok():
    if not R.available():
        pytest.skip("reference tree / oracle/_ref not available")


@pytest.mark.parametrize("behaviour", ["SACCADE", "TRAVERSE", "FADE", "NONE"])
@pytest.mark.parametrize("with_mask", [False, True])
def test_restated_schedule_matches_the_reference_class(ref_ok, behaviour, with_mask):
    n = 40 if behaviour == "NONE" else 190       # ten images x (12 frames on + 3 off) = one full cycle and the end of stream
    run = R.run_reference(behaviour, n, with_mask=with_mask)
    clock = R.Clock()
    ora = _oracle_for(run, R.MirrorRng(clock, 12345))
    for i in range(n):
        clock.now = ora.now
        ok, img = ora.read()
        assert ok == run["valid"][i], (behaviour, i)
        assert np.array_equal(img, run["frames"][i]), (behaviour, with_mask, i)
    if behaviour != "NONE":
        assert not run["valid"][-1] and sum(run["valid"]) > 100     # end of stream reached (max_cycles = 1)


def test_restated_attention_matches_the_reference_class(ref_ok):
    """ATTENTION consumes the DVS reference plane.  Frames on which the five largest changes are tied have no defined
    order in the reference (unstable np.argsort): the restatement is re-synchronised there and the frame not counted."""
    n = 150
    rng = np.random.default_rng(3)
    ref_planes = [rng.integers(0, 256, (32, 32)).astype(np.int16) for _ in range(n)]
    run = R.run_reference("ATTENTION", n, with_mask=True, ref_planes=ref_planes)
    clock = R.Clock()
    ora = _oracle_for(run, R.MirrorRng(clock, 12345))
    compared = saccades = 0
    for i in range(n):
        clock.now = ora.now
        tied_before = ora.tied_frames
        ok, img = ora.read(ref_planes[i])
        assert ok == run["valid"][i]
        if ora.tied_frames != tied_before and (ora.cx, ora.cy) != run["centers"][i]:
            ora.cx, ora.cy = run["centers"][i]      # unspecified order in the reference: follow it and move on
            ora.current = run["frames"][i].copy()
            continue
        assert (ora.cx, ora.cy) == run["centers"][i] or not ok or not ora.showing, i
        assert np.array_equal(img, run["frames"][i]), i
        compared += 1
        saccades += int(ora.showing and ok and (ora.frame_number - 1) % 5 == 0)
    assert compared > 120 and saccades > 15


def test_area_halving_is_cv2_inter_area(ref_ok):
    """The CUDA search takes (sum of the 2x2 block + 2) >> 2 for cv2.resize(INTER_AREA) on int16 planes."""
    import cv2
    rng = np.random.default_rng(0)
    for trial in range(60):
        w = int(rng.choice([8, 16, 32, 64, 128, 30]))
        lo, hi = [(0, 256), (-300, 300), (-32768, 32768), (0, 4)][trial % 4]
        a = rng.integers(lo, hi, (w, w)).astype(np.int16)
        q = a.astype(np.int32).reshape(w // 2, 2, w // 2, 2).sum(axis=(1, 3))
        assert np.array_equal(cv2.resize(a, (w // 2, w // 2), interpolation=cv2.INTER_AREA), ((q + 2) >> 2).astype(np.int16))


def test_attention_target_matches_the_compiled_attention_image(ref_ok):
    """oracle attention_target against the compiled reference function itself (np.random seeded for the pick)."""
    import warnings
    from oracle import ref_loader
    gs = ref_loader.load()[0]
    rng = np.random.default_rng(9)
    untied = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for trial in range(120):
            w = int(rng.choice([16, 32, 64]))
            orig = rng.integers(0, 256, (w, w)).astype(np.int16)
            prev = rng.integers(0, 256, (w, w)).astype(np.int16)
            ref = rng.integers(0, 256, (w, w)).astype(np.int16)
            np.random.seed(trial)
            img, cx, cy = gs.attention_image(orig, prev, ref, 0, 1, 5, 2, 0, 0, 0)
            pick = int(np.random.RandomState(trial).randint(5))
            ecx, ecy, tied = attention_target(prev, ref, pick)
            if tied:
                continue
            untied += 1
            assert (int(cx), int(cy)) == (ecx, ecy), trial
    assert untied > 60


def test_golden_vcam_file_is_current(ref_ok):
    """The committed golden frames are what the reference class produces today."""
    g = np.load(GOLD)
    for name in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        behaviour, masked = name.split("_")[0], name.endswith("_mask")
        n = len(g[name + "/frames"])
        kw = {}
        if behaviour == "ATTENTION":
            kw["ref_planes"] = list(g[name + "/ref_planes"])
        run = R.run_reference(behaviour, n, with_mask=masked, **kw)
        assert np.array_equal(np.stack(run["frames"]), g[name + "/frames"]), name
