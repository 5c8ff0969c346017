"""Scalar NumPy restatement of ``VirtualCam.read`` (``pydvs/virtual_cam.py:217-318, 387-424``) for ONE camera with an
injected clock and injected random draws.  TEST INFRASTRUCTURE ONLY.

The reference class cannot run here as shipped (``glob2`` is not installed, it spins threads and reads the wall
clock), so the schedule is restated; the frame animators it calls are the pinned ones of ``oracle/frame_oracle.py``
(checked against the compiled reference).  ATTENTION's saccade search is not covered (its tie order inside
``np.argsort`` and its unseeded pick make it non-reproducible in the reference itself).
"""
import numpy as np

from . import frame_oracle as fo


class VirtualCamOracle:
    def __init__(self, images, behaviour, fps, image_on_time_ms, inter_off_time_ms, max_saccade_distance,
                 frames_per_microsaccade, max_cycles, fade_mask, background_gray, start_index, rng):
        self.images = images
        self.n = len(images)
        self.h, self.w = images[0].shape
        self.behaviour = behaviour
        self.period = 1. / fps
        self.half_frame = int((fps * (image_on_time_ms / 1000.)) / 2.)
        self.on_time, self.off_time = image_on_time_ms / 1000., inter_off_time_ms / 1000.
        self.max_dist, self.fp_usacc = max_saccade_distance, frames_per_microsaccade
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
        self.original = self._orig()
        self.current = self.original.copy()
        if behaviour in ("FADE", "TRAVERSE"):
            self.current[:] = self.bg

    def _orig(self):
        return self.images[(self.start + self.idx) % self.n]

    def read(self):
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
                self.original = self._orig()
                if self.behaviour == "TRAVERSE":
                    img = fo.traverse_image(self.original, self.frame_number, self.speed, self.bg)
                elif self.behaviour == "FADE":
                    img = fo.fade_image(self.original, self.frame_number, self.half_frame, self.bg)
                else:
                    if self.frame_number % self.fp_usacc == 0:
                        self.cx += self.rng.randint(-self.max_dist, self.max_dist + 1)
                        self.cy += self.rng.randint(-self.max_dist, self.max_dist + 1)
                    img = fo.move_image(self.original, self.cx, self.cy, self.bg)
                if self.mask is not None:
                    img = fo.mask_image(img, self.mask)
                self.current = img
        self.frame_number += 1
        self.now += self.period
        return True, self.current
