"""On-GPU frame sources -- the step BEFORE the hot path.

``move_images`` is the batched form of the reference's ``move_image`` (``generate_spikes.pyx:1184-1215``, the
core of ``usaccade_image`` / ``attention_image``) fused with ``mask_image`` (``:1124-1127``).
``SaccadeSource`` is a deterministic, batched counterpart of the frame loop of ``mnist_to_spikes.py:236-260``
(what ``VirtualCam`` does in its MICROSACCADE behaviour, ``virtual_cam.py:391-424``): the reference reseeds NumPy
from the wall clock on every call (``generate_spikes.pyx:1238``), here the per-stream offsets come from a seeded
generator so that sequences are reproducible and can be fed to the CPU oracle.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def move_images(images, img_index, dx, dy, bg=0, mask=None, out=None):
    """images: CUDA int16 [N, H, W]; img_index / dx / dy: int32 [S] (CUDA); mask: float64 [H, W] or None.
    Returns int16 [S, H, W]: stream s = images[img_index[s]] moved right by dx[s] and up by dy[s]."""
    if images.dtype != torch.int16 or images.dim() != 3 or not images.is_cuda:
        raise ValueError("images must be a CUDA int16 tensor [N, H, W]")
    n, H, W = images.shape
    dev = images.device
    idx, dx, dy = (torch.as_tensor(v, dtype=torch.int32, device=dev).contiguous() for v in (img_index, dx, dy))
    S = idx.numel()
    if dx.numel() != S or dy.numel() != S:
        raise ValueError("img_index, dx and dy must have one entry per stream")
    if mask is not None:
        mask = torch.as_tensor(mask, dtype=torch.float64, device=dev).contiguous()
        if tuple(mask.shape) != (H, W):
            raise ValueError("mask must be [H, W]")
    if out is None:
        out = torch.empty((S, H, W), dtype=torch.int16, device=dev)
    images = images.contiguous()
    check(lib.dvs_move_mask_i16(C.c_void_p(images.data_ptr()), n, C.c_void_p(idx.data_ptr()), C.c_void_p(dx.data_ptr()),
                                C.c_void_p(dy.data_ptr()), int(bg), C.c_void_p(mask.data_ptr()) if mask is not None else None,
                                C.c_void_p(out.data_ptr()), S, H, W, _stream()))
    return out


def _animate(mode, images, img_index, frame_number, speed, half_frame, bg, out):
    if images.dtype != torch.int16 or images.dim() != 3 or not images.is_cuda:
        raise ValueError("images must be a CUDA int16 tensor [N, H, W]")
    n, H, W = images.shape
    dev = images.device
    idx = torch.as_tensor(img_index, dtype=torch.int32, device=dev).contiguous()
    fn = torch.as_tensor(frame_number, dtype=torch.int32, device=dev).contiguous()
    S = idx.numel()
    if fn.numel() == 1 and S > 1:
        fn = fn.expand(S).contiguous()
    if fn.numel() != S:
        raise ValueError("frame_number must be a scalar or have one entry per stream")
    if out is None:
        out = torch.empty((S, H, W), dtype=torch.int16, device=dev)
    images = images.contiguous()
    check(lib.dvs_animate_i16(mode, C.c_void_p(images.data_ptr()), n, C.c_void_p(idx.data_ptr()), C.c_void_p(fn.data_ptr()),
                              float(speed), int(half_frame), int(bg), C.c_void_p(out.data_ptr()), S, H, W, _stream()))
    return out


def traverse_images(images, img_index, frame_number, speed, bg=0, out=None):
    """Batched ``traverse_image`` (``generate_spikes.pyx:1128-1152``): stream s shows images[img_index[s]] sliding in
    from the left / out to the right at ``speed`` pixels per frame, at its own ``frame_number[s]``."""
    return _animate(0, images, img_index, frame_number, speed, 0, bg, out)


def fade_images(images, img_index, frame_number, half_frame, bg=0, out=None):
    """Batched ``fade_image`` (``generate_spikes.pyx:1156-1182``): ``int16(original * alpha)`` with the triangular
    alpha profile peaking at ``half_frame``."""
    return _animate(1, images, img_index, frame_number, 0.0, half_frame, bg, out)


def saccade_offsets(n_streams, n_frames, frames_per_microsaccade=1, max_dist=1, seed=0):
    """Per-frame (cx, cy) of every stream following mnist_to_spikes.py:246-258: frame 0 shows the image
    unmoved; on frames that are multiples of `frames_per_microsaccade` both offsets take a uniform step in
    [-max_dist, max_dist]; after each frame `if abs(cx) > 1: cx = 0 elif abs(cy) > 1: cy = 0`.
    Returns int32 [n_frames, n_streams, 2] (host)."""
    rng = np.random.RandomState(seed)
    steps = rng.randint(-max_dist, max_dist + 1, size=(n_frames, n_streams, 2))
    cx = np.zeros(n_streams, dtype=np.int64)
    cy = np.zeros(n_streams, dtype=np.int64)
    out = np.zeros((n_frames, n_streams, 2), dtype=np.int32)
    for t in range(1, n_frames):
        if t % frames_per_microsaccade == 0:
            cx = cx + steps[t, :, 0]
            cy = cy + steps[t, :, 1]
        out[t, :, 0], out[t, :, 1] = cx, cy
        reset_x = np.abs(cx) > 1
        reset_y = (~reset_x) & (np.abs(cy) > 1)
        cx = np.where(reset_x, 0, cx)
        cy = np.where(reset_y, 0, cy)
    return out


class SaccadeSource:
    """Batched micro-saccade frame source: stream s shows images[img_index[s]] for `frames_per_image` frames."""

    def __init__(self, images, img_index, frames_per_image=10, frames_per_microsaccade=1, max_dist=1, bg=0,
                 fade_mask=None, seed=0):
        self.images = images
        self.index = torch.as_tensor(img_index, dtype=torch.int32, device=images.device)
        self.n_frames, self.bg = frames_per_image, bg
        self.mask = None if fade_mask is None else torch.as_tensor(fade_mask, dtype=torch.float64, device=images.device)
        self.offsets_host = saccade_offsets(self.index.numel(), frames_per_image, frames_per_microsaccade, max_dist, seed)
        self.offsets = torch.from_numpy(self.offsets_host).to(images.device)
        self._out = None

    def frame(self, t):
        off = self.offsets[t]
        self._out = move_images(self.images, self.index, off[:, 0].contiguous(), off[:, 1].contiguous(), self.bg,
                                self.mask, out=self._out)
        return self._out
