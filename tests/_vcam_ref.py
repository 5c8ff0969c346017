"""Test helper: run the REFERENCE's own ``VirtualCam`` class (``/root/reference/pydvs/virtual_cam.py``, loaded
unmodified) deterministically in this container.

What the reference needs and does not find here is supplied from outside, nothing in it is edited:
  * ``glob2`` (not installed): a stand-in module whose ``glob`` is the standard library's;
  * ``pydvs.generate_spikes``: the compiled, unmodified reference module of ``oracle/_ref`` (``oracle/build_ref.py``);
  * the wall clock: ``virtual_cam.get_time`` and the ``time`` module global of the compiled ``generate_spikes`` (read by
    ``usaccade_image`` to reseed NumPy, ``generate_spikes.pyx:1238``) both return a virtual clock that the caller
    advances by ``1 / fps`` after every ``read()``.
``MirrorRng`` replays the reference's use of the global NumPy generator (seed from the clock, two draws per
micro-saccade; saccade picks without reseeding) so that a restatement can be fed the same numbers.
"""
import glob
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PYDVS_REFERENCE", "/root/reference")
VCAM_PATH = os.path.join(REF_ROOT, "pydvs", "virtual_cam.py")
MNIST_DIR = os.path.join(REF_ROOT, "mnist", "t10k")


class Clock:
    def __init__(self):
        self.now = 0.0

    def time(self):
        return self.now


def available():
    from oracle import ref_loader
    return os.path.exists(VCAM_PATH) and os.path.isdir(MNIST_DIR) and ref_loader.load() is not None


def load(clock):
    """-> (reference virtual_cam module, compiled generate_spikes) with the clock patched in."""
    from oracle import ref_loader
    gs = ref_loader.load()[0]
    stub_glob2 = types.ModuleType("glob2")
    stub_glob2.glob = glob.glob
    pkg = types.ModuleType("pydvs")
    pkg.generate_spikes = gs
    saved = {k: sys.modules.get(k) for k in ("glob2", "pydvs", "pydvs.generate_spikes")}
    sys.modules.update({"glob2": stub_glob2, "pydvs": pkg, "pydvs.generate_spikes": gs})
    try:
        spec = importlib.util.spec_from_file_location("_ref_virtual_cam", VCAM_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod.get_time = clock.time
    fake_time = types.ModuleType("time")
    fake_time.time = clock.time
    gs.time = fake_time          # the `time` global usaccade_image reads (gs:1238)
    return mod, gs


class MirrorRng:
    """The global NumPy generator as the reference drives it, on a private ``RandomState``."""

    def __init__(self, clock, seed0):
        self.clock, self.rs, self.first = clock, np.random.RandomState(seed0), True
        self.steps, self.picks = [], []

    def randint(self, a, b=None):
        if b is None:
            v = int(self.rs.randint(a))
            self.picks.append(v)
            return v
        if self.first:
            with np.errstate(all="ignore"):
                self.rs.seed(np.uint32(self.clock.now * 10000000000))   # generate_spikes.pyx:1238
            self.steps.append([0, 0])
        v = int(self.rs.randint(a, b))
        self.steps[-1][0 if self.first else 1] = v
        self.first = not self.first
        return v


def run_reference(behaviour, n_reads, fps=3000, with_mask=True, ref_planes=None, seed0=12345, **kw):
    """Run the reference class for ``n_reads`` frames.  -> dict(frames, valid, images, mask, centers, ...).

    The virtual clock runs in milliseconds-scale units (3000 fps, image on for 4 ms = 12 frames, off for 1 ms = 3 frames):
    ``usaccade_image`` seeds NumPy with ``np.uint32(time.time() * 1e10)`` (gs:1238), which NumPy >= 2 refuses for values
    above 2**32 - 1 -- with the real wall clock the reference's micro-saccades raise OverflowError on this NumPy; a
    clock that stays below 0.43 s keeps that statement executable as written."""
    import warnings
    clock = Clock()
    mod, gs = load(clock)
    np.random.seed(seed0)
    args = dict(behaviour=behaviour, fps=fps, resolution=32, image_on_time_ms=4, inter_off_time_ms=1,
                max_saccade_distance=2, frames_per_microsaccade=2, frames_per_saccade=5, max_cycles=1,
                fade_with_mask=with_mask, background_gray=0)
    args.update(kw)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):     # the class prints every file name
        cam = mod.VirtualCam(MNIST_DIR, **args)
    images = [a.copy() for a in cam.image_buffer[cam.current_buffer]]
    mask = cam.fade_mask.copy() if with_mask else None
    frames, valid, centers = [], [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(n_reads):
            ref = ref_planes[i] if ref_planes is not None else np.zeros((32, 32), np.int16)
            ok, img = cam.read(ref)
            frames.append(np.array(img, dtype=np.int16, copy=True))
            valid.append(bool(ok))
            centers.append((int(cam.center_x), int(cam.center_y)))
            clock.now += 1.0 / fps
    cam.stop()
    return dict(frames=frames, valid=valid, images=images, mask=mask, centers=centers, args=args, clock=clock)
