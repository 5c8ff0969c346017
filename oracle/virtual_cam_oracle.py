"""Scalar NumPy restatement of ``VirtualCam.read`` (``pydvs/virtual_cam.py:217-318, 387-424``) for ONE camera with an
injected clock and injected random draws.  TEST INFRASTRUCTURE ONLY.

PINNED: ``tests/test_virtual_cam_reference.py`` runs the reference's own ``VirtualCam`` class (loaded from
``/root/reference/pydvs/virtual_cam.py`` with a stand-in ``glob2``, the compiled reference as ``pydvs.generate_spikes``,
a virtual clock and the reference's own wall-clock seeding made deterministic) on the ``mnist/t10k`` fixtures and
compares it frame by frame with this restatement; ``tests/golden/golden_vcam_v1.npz`` holds frames recorded from those
runs for the GPU box.  The frame animators are the pinned ones of ``oracle/frame_oracle.py``.  ATTENTION's saccade
search uses ``cv2.resize(INTER_AREA)`` like the reference and a STABLE argsort: the reference's ``np.argsort`` leaves the
order of equal values unspecified, so comparisons against it hold on frames whose five largest changes are untied.
"""
import numpy as np

from . import frame_oracle as fo


def attention_target(previous, reference, pick):
    """The saccade search of ``attention_image`` (``generate_spikes.pyx:1271-1295``) -> (cx, cy, tied): ``tied`` tells
    that the five largest values are not separated from each other / from the sixth (unspecified order in NumPy)."""
    import cv2
    new_w = previous.shape[1] // 2
    tiny_prev = cv2.resize(previous, (new_w, new_w), interpolation=cv2.INTER_AREA).astype(np.int16)
    tiny_ref = cv2.resize(reference, (new_w, new_w), interpolation=cv2.INTER_AREA).astype(np.int16)
    diff = np.abs(tiny_prev - tiny_ref).astype(np.int16).reshape(new_w * new_w)
    order = np.argsort(diff, kind="stable")
    top_n = order[-5:]
    top6 = diff[order[-6:]]
    tied = bool(len(np.unique(top6)) != len(top6))
    max_idx = int(top_n[pick])
    row, col = (max_idx // new_w) * 2, (max_idx % new_w) * 2
    cx, cy = new_w - col, new_w - row
    if cx > new_w or cx < -new_w or cy > new_w or cy < -new_w:
        cx = cy = 0
    return cx, cy, tied


class ReplayDraws:
    """The vectorised draws of ``pydvs_b200.virtual_cam.StreamDraws`` (one ``RandomState``, one call per frame for all
    S cameras) handed out one camera at a time: ``stream(s)`` is an object with ``randint`` for ``VirtualCamOracle``."""

    def __init__(self, seed, S):
        self.rs, self.S, self.calls = np.random.RandomState(seed), S, []

    def call(self, k, kind):
        while len(self.calls) <= k:
            if kind[0] == "pick":
                self.calls.append(self.rs.randint(5, size=self.S))
            else:
                self.calls.append(self.rs.randint(kind[1], kind[2], size=(self.S, 2)))
        return self.calls[k]

    def stream(self, s):
        parent = self

        class _View:
            def __init__(self):
                self.k, self.pending = 0, None

            def randint(self, a, b=None):
                if b is None:
                    v = parent.call(self.k, ("pick",))
                    self.k += 1
                    return int(v[s])
                if self.pending is None:
                    self.pending = parent.call(self.k, ("off", a, b))
                    self.k += 1
                    return int(self.pending[s, 0])
                v, self.pending = self.pending[s, 1], None
                return int(v)
        return _View()


class VirtualCamOracle:
    def __init__(self, images, behaviour, fps, image_on_time_ms, inter_off_time_ms, max_saccade_distance,
                 frames_per_microsaccade, max_cycles, fade_mask, background_gray, start_index, rng,
                 frames_per_saccade=29):
        self.images = images
        self.n = len(images)
        self.h, self.w = images[0].shape
        self.behaviour = behaviour
        self.period = 1. / fps
        self.half_frame = int((fps * (image_on_time_ms / 1000.)) / 2.)
        self.on_time, self.off_time = image_on_time_ms / 1000., inter_off_time_ms / 1000.
        self.max_dist, self.fp_usacc = max_saccade_distance, frames_per_microsaccade
        self.fp_sacc = frames_per_saccade
        self.tied_frames = 0
        self.speed = (self.w * 2.) // (self.on_time * fps)
        self.max_cycles, self.num_cycles = max_cycles, 0
        self.mask, self.bg = fade_mask, background_gray
        self.start, self.idx = start_index, 0
        self.rng = rng
        self.cx = self.cy = 0
        self.frame_number = 0
        self.showing = True
        self.now = 0.0
        self.t0 = 0.0
        # virtual_cam.py:84-85, 112: `original_image` starts as ZEROS and is only filled when an off-period ends
        # (:284), while `current_image` starts as the first image -- so the whole first on-period animates an all-zero
        # picture and image 0 itself is never shown moved (pinned by tests/test_virtual_cam_reference.py)
        self.original = np.zeros((self.h, self.w), np.int16)
        self.current = self._orig().copy()
        if behaviour in ("FADE", "TRAVERSE"):
            self.current[:] = self.bg

    def _orig(self):
        return self.images[(self.start + self.idx) % self.n]

    def read(self, ref=None):
        run_time = self.now - self.t0
        if self.behaviour == "NONE":
            self.current = self._orig().copy()
            self.idx = (self.idx + 1) % self.n
            self.now += self.period
            return True, self.current
        if self.num_cycles == self.max_cycles:
            self.current = np.full((self.h, self.w), self.bg, np.int16)
            return False, self.current
        moving = self.behaviour in ("TRAVERSE", "FADE")
        if not self.showing:
            if run_time >= self.off_time:
                self.showing = True
                self.t0 = self.now
                self.idx += 1
                if self.idx == self.n:
                    self.num_cycles += 1
                    if self.num_cycles == self.max_cycles:
                        self.current = np.full((self.h, self.w), self.bg, np.int16)
                        return False, self.current
                    self.idx = 0
                self.frame_number = 0
                self.original = self._orig()
                self.current = self.original.copy()
                if moving:
                    self.current[:] = self.bg
                else:
                    self.cx = self.cy = 0
            else:
                self.current = np.full((self.h, self.w), self.bg, np.int16)
        else:
            if run_time >= self.on_time:
                self.showing = False
                self.t0 = self.now
                self.frame_number = 0
            else:
                if self.behaviour == "TRAVERSE":
                    img = fo.traverse_image(self.original, self.frame_number, self.speed, self.bg)
                elif self.behaviour == "FADE":
                    img = fo.fade_image(self.original, self.frame_number, self.half_frame, self.bg)
                else:
                    if self.behaviour == "ATTENTION" and self.frame_number % self.fp_sacc == 0:
                        self.cx, self.cy, tied = attention_target(self.current, ref, self.rng.randint(5))
                        self.tied_frames += tied
                    elif self.frame_number % self.fp_usacc == 0:
                        self.cx += self.rng.randint(-self.max_dist, self.max_dist + 1)
                        self.cy += self.rng.randint(-self.max_dist, self.max_dist + 1)
                    img = fo.move_image(self.original, self.cx, self.cy, self.bg)
                if self.mask is not None:
                    img = fo.mask_image(img, self.mask)
                self.current = img
        self.frame_number += 1
        self.now += self.period
        return True, self.current
