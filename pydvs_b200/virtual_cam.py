"""Deterministic, batched counterpart of ``pydvs/virtual_cam.py``.

The reference's ``VirtualCam`` paces itself with the wall clock (``virtual_cam.py:232-233, 270-303``), reseeds NumPy
from it on every micro-saccade (``generate_spikes.pyx:1238``), loads images from disk with a double-buffering thread
and serves one camera.  Here the schedule of ``read()`` is kept statement for statement, but

* time is injected: every ``read()`` advances a virtual clock by ``1 / fps`` (or by what ``clock`` returns),
* the random draws come from one seeded generator, one vectorised call per frame for all streams (``StreamDraws``),
* the images are a device-resident ``int16 [N, H, W]`` stack and ``S`` cameras are served at once: stream ``s``
  starts at image ``start_index[s]`` and they all follow the same on/off schedule, so one call returns the
  ``[S, H, W]`` batch a ``DVSStreams.step`` consumes -- no host round trip per frame.

Frames are produced by the batched kernels of ``pydvs_b200.frame_sources`` (``move_images`` + fade mask,
``traverse_images``, ``fade_images``); ``ATTENTION`` adds the saccade target search of ``attention_image``
(``generate_spikes.pyx:1246-1298``: 2x2 area means of the previous frame and the DVS reference, arg-top-5 of their
absolute difference, a random pick) as one kernel launch for all cameras (``dvs_attention_targets_i16``).  Nothing in
``read()`` loops over streams on the host: the image centres live on the device and the random draws of a frame are one
vectorised call, so a 4096-camera source costs the same handful of launches as a single camera.
"""
import ctypes as C

import numpy as np
import torch

from . import frame_sources as fs
from ._lib import lib, check


class StreamDraws:
    """The random draws of ``S`` cameras from one seeded generator, one vectorised call per frame (the reference
    reseeds the global NumPy generator from the wall clock inside ``usaccade_image``, ``generate_spikes.pyx:1238``)."""

    def __init__(self, seed=0):
        self.rs = np.random.RandomState(seed)

    def offsets(self, S, max_delta):
        """int [S, 2]: the (x, y) micro-saccade steps, uniform in [-max_delta, max_delta] (gs:1239-1240)."""
        return self.rs.randint(-max_delta, max_delta + 1, size=(S, 2))

    def picks(self, S):
        """int [S]: which of the five largest changes a saccade jumps to (gs:1283)."""
        return self.rs.randint(5, size=S)


class VirtualCam:
    BEHAVE_MICROSACCADE = "SACCADE"
    BEHAVE_ATTENTION = "ATTENTION"
    BEHAVE_TRAVERSE = "TRAVERSE"
    BEHAVE_FADE = "FADE"
    BEHAVE_NONE = "NONE"

    def __init__(self, images, behaviour="SACCADE", fps=90, image_on_time_ms=1000, inter_off_time_ms=100,
                 max_saccade_distance=1, frames_per_microsaccade=1, frames_per_saccade=29, max_cycles=3,
                 fade_mask=None, background_gray=0, streams=1, start_index=None, seed=0, clock=None, draws=None):
        if images.dtype != torch.int16 or images.dim() != 3 or not images.is_cuda:
            raise ValueError("images must be a CUDA int16 tensor [N, H, W]")
        # image 0 of the device stack is an all-zero picture: the reference's `original_image` starts as zeros and is only
        # filled when an off-period ends (virtual_cam.py:84, 284), so its whole first on-period animates zeros
        self.images = torch.cat([torch.zeros_like(images[:1]), images.contiguous()], dim=0)
        self.total_images, self.height, self.width = images.shape
        self._original_loaded = False
        self.shape = (self.height, self.width)
        self.S = int(streams)
        dev = images.device
        self.behaviour = behaviour
        self.fps = fps
        self.time_period = 1. / fps
        self.half_frame = int((fps * (image_on_time_ms / 1000.)) / 2.)           # virtual_cam.py:44
        self.image_on_time = image_on_time_ms / 1000.
        self.inter_off_time = inter_off_time_ms / 1000.
        self.max_saccade_distance = max_saccade_distance
        self.traverse_speed = (self.width * 2.) // (self.image_on_time * self.fps)  # virtual_cam.py:50
        self.frames_per_microsaccade = frames_per_microsaccade
        self.frames_per_saccade = frames_per_saccade
        self.background_gray = background_gray
        self.max_cycles = max_cycles
        self.num_cycles = 0
        self.fade_mask = None if fade_mask is None else torch.as_tensor(fade_mask, dtype=torch.float64, device=dev)
        start = np.zeros(self.S, dtype=np.int64) if start_index is None else np.asarray(start_index, dtype=np.int64)
        if start.shape != (self.S,):
            raise ValueError("start_index needs one entry per stream")
        self._start = start % self.total_images
        self.global_image_idx = 0           # position in the image sequence, shared by all streams
        self.cx = torch.zeros(self.S, dtype=torch.int32, device=dev)   # image centre per stream, on the device
        self.cy = torch.zeros(self.S, dtype=torch.int32, device=dev)
        self.frame_number = 0
        self.showing_img = True
        self.draws = draws if draws is not None else StreamDraws(seed)
        self._clock = clock
        self._now = 0.0
        self.on_off_start_time = self._time()
        self._bg = torch.full((self.S,) + self.shape, int(background_gray), dtype=torch.int16, device=dev)
        self.current_image = self._originals().clone()
        if behaviour in (self.BEHAVE_FADE, self.BEHAVE_TRAVERSE):                 # virtual_cam.py:120-122
            self.current_image = self._bg.clone()

    # -- time --------------------------------------------------------------------------------------------
    def _time(self):
        return self._clock() if self._clock is not None else self._now

    def _tick(self):
        if self._clock is None:
            self._now += self.time_period

    # -- images --------------------------------------------------------------------------------------------
    def image_index(self):
        """int32 [S]: which image every stream currently shows."""
        return ((self._start + self.global_image_idx) % self.total_images).astype(np.int32)

    def _stack_index(self):
        """Index into the device stack of the picture the animators work on (`original_image`): the zero picture until
        the first off-period has ended."""
        if not self._original_loaded:
            return np.zeros(self.S, dtype=np.int32)
        return self.image_index() + 1

    def _originals(self):
        idx = torch.from_numpy((self.image_index() + 1).astype(np.int64)).to(self.images.device)
        return self.images.index_select(0, idx)

    # -- cv2.VideoCapture look-alikes (virtual_cam.py:171-189) --------------------------------------------------
    def isOpened(self):
        return True

    def get(self, prop):
        import cv2
        return {cv2.CAP_PROP_FRAME_WIDTH: self.width, cv2.CAP_PROP_FRAME_HEIGHT: self.height,
                cv2.CAP_PROP_FPS: self.fps}.get(prop, False)

    def release(self):
        pass

    stop = release

    # -- the frame loop: virtual_cam.py:217-318 ----------------------------------------------------------------
    def read(self, ref=None):
        """-> (valid, frames int16 CUDA [S, H, W]).  ``ref``: the DVS reference planes [S, H, W] (ATTENTION only)."""
        now = self._time()
        run_time = now - self.on_off_start_time
        if self.behaviour == self.BEHAVE_NONE:
            self.current_image = self._originals()
            self.global_image_idx += 1
            if self.global_image_idx == self.total_images:
                self.global_image_idx = 0
            self._tick()
            return True, self.current_image
        if self.num_cycles == self.max_cycles:
            self.current_image = self._bg.clone()
            return False, self.current_image
        moving = self.behaviour in (self.BEHAVE_TRAVERSE, self.BEHAVE_FADE)
        if not self.showing_img:
            if run_time >= self.inter_off_time:
                self.showing_img = True
                self.on_off_start_time = self._time()
                self.global_image_idx += 1
                if self.global_image_idx == self.total_images:
                    self.num_cycles += 1
                    if self.num_cycles == self.max_cycles:
                        self.current_image = self._bg.clone()
                        return False, self.current_image
                    self.global_image_idx = 0
                self.frame_number = 0
                self._original_loaded = True                                      # virtual_cam.py:284
                self.current_image = self._bg.clone() if moving else self._originals().clone()
                if not moving:
                    self.cx.zero_()
                    self.cy.zero_()
            else:
                self.current_image = self._bg.clone()
        else:
            if run_time >= self.image_on_time:
                self.showing_img = False
                self.on_off_start_time = self._time()
                self.frame_number = 0
            else:
                self.current_image = self._move_image(ref)
        self.frame_number += 1
        self._tick()
        return True, self.current_image

    # -- virtual_cam.py:387-424 -----------------------------------------------------------------------------------
    @property
    def center(self):
        """(cx, cy) of every stream as a host array [S, 2] (synchronises; for inspection and tests)."""
        return torch.stack([self.cx, self.cy], dim=1).cpu().numpy().astype(np.int64)

    def _draw_offsets(self):
        # two draws per stream, x then y (generate_spikes.pyx:1239-1240): one vectorised call, one small upload
        step = torch.from_numpy(np.ascontiguousarray(self.draws.offsets(self.S, self.max_saccade_distance), dtype=np.int32))
        step = step.to(self.cx.device, non_blocking=True)
        self.cx += step[:, 0]
        self.cy += step[:, 1]

    def _saccade_targets(self, ref):
        """attention_image gs:1276-1295 for every stream: new (cx, cy) from the largest change between the previous
        frame and the DVS reference at half resolution (one launch, ``dvs_attention_targets_i16``)."""
        if self.height != self.width or self.width % 2:
            raise NotImplementedError("ATTENTION needs square frames of even size (the reference resizes to (w/2, w/2))")
        if ref is None:
            raise ValueError("ATTENTION needs the DVS reference planes")
        ref = ref.to(self.current_image.device)
        if ref.dtype != torch.int16 or ref.numel() != self.current_image.numel():
            raise ValueError("ref must be int16 [S, H, W]")
        picks = torch.from_numpy(np.ascontiguousarray(self.draws.picks(self.S), dtype=np.int32)).to(self.cx.device)
        prev = self.current_image.contiguous()
        check(lib.dvs_attention_targets_i16(C.c_void_p(prev.data_ptr()), C.c_void_p(ref.contiguous().data_ptr()),
                                            C.c_void_p(picks.data_ptr()), C.c_void_p(self.cx.data_ptr()),
                                            C.c_void_p(self.cy.data_ptr()), self.S, self.width,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def _move_image(self, ref):
        fn, idx = self.frame_number, self._stack_index()
        if self.behaviour == self.BEHAVE_TRAVERSE:
            out = fs.traverse_images(self.images, idx, fn, self.traverse_speed, self.background_gray)
        elif self.behaviour == self.BEHAVE_FADE:
            out = fs.fade_images(self.images, idx, fn, self.half_frame, self.background_gray)
        else:
            if self.behaviour == self.BEHAVE_ATTENTION and fn % self.frames_per_saccade == 0:
                self._saccade_targets(ref)
            elif fn % self.frames_per_microsaccade == 0:
                self._draw_offsets()
            return fs.move_images(self.images, idx, self.cx, self.cy, self.background_gray, self.fade_mask)
        if self.fade_mask is not None:   # virtual_cam.py:309-311: mask after the animator
            zero = torch.zeros(self.S, dtype=torch.int32, device=self.cx.device)
            out = fs.move_images(out, torch.arange(self.S, dtype=torch.int32, device=self.cx.device), zero, zero, 0,
                                 self.fade_mask)
        return out
