"""The float32 NVS variants of ``pydvs/numba_nvs.py`` as one fused sm_100a pass.

Same function names and positional arguments as the reference module; every function works in place on
the caller's ``reference`` / ``thresholds`` / ``spikes`` planes (any shape -- ``[H, W]`` like the
reference or ``[S, H, W]`` for a batch of streams).  CUDA float32 tensors are used as they are; NumPy
arrays are uploaded, processed on the GPU and written back (an I/O adapter, not a CPU path).

Precision is what numba gives the reference (pinned by ``tests/test_oracle_vs_reference.py``):
float32 for ``thresholded_difference`` (``numba_nvs.py:6-20``; ``//`` is NumPy's floor_divide) and for
array-only expressions, float64 rounded once wherever a Python-float scalar enters.  The noisy variants
draw ``uniform(0, 1) <= p`` from numba's private unseeded generator (``numba_nvs.py:85,119,133``); here
the draw is either supplied (``leak_mask=``, 1 = the pixel leaks this frame) or made on the device from
``generator=`` -- statistically the same process, reproducible when seeded.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Planes:
    """Upload NumPy planes / pass CUDA tensors through; copy results back to NumPy callers."""

    def __init__(self):
        self.back = []

    def f32(self, x, name, like=None, writable=False):
        if isinstance(x, torch.Tensor):
            if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous():
                raise TypeError("%s must be a contiguous float32 CUDA tensor (or a NumPy float32 array)" % name)
            t = x
        else:
            if not isinstance(x, np.ndarray) or x.dtype != np.float32:
                raise TypeError("%s: expected a float32 array, got %r" % (name, getattr(x, "dtype", type(x))))
            t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
            if writable:
                self.back.append((x, t))
        if like is not None and t.numel() != like.numel():
            raise ValueError("%s has %d elements, frame has %d" % (name, t.numel(), like.numel()))
        return t

    def finish(self):
        for host, dev in self.back:
            host[...] = dev.cpu().numpy().reshape(host.shape)


def _mask(mask, probability, like, generator):
    if mask is not None:  # any non-zero byte = "this pixel leaks"; uint8 / bool tensors are used as they are
        m = mask if isinstance(mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(mask))
        if m.dtype == torch.bool:
            m = m.view(torch.uint8)
        elif m.dtype != torch.uint8:
            m = (m != 0).view(torch.uint8)
        m = m.to(like.device).contiguous()
        if m.numel() != like.numel():
            raise ValueError("leak mask has %d elements, frame has %d" % (m.numel(), like.numel()))
        return m
    u = torch.rand(like.shape, device=like.device, dtype=torch.float32, generator=generator)
    return (u <= float(probability)).view(torch.uint8)


def _run(mode, frame, reference, thresholds, spikes, reference_leak=1.0, threshold_base=0.0,
         threshold_mult_incr=0.0, threshold_leak=1.0, leak_mask=None, reference_off=None, thresholds_off=None,
         target_off=0.0, leak_mask_off=None, spikes_off=None, want_active=True):
    io = _Planes()
    f = io.f32(frame, "frame")
    r = io.f32(reference, "reference", f, True)
    t = io.f32(thresholds, "thresholds", f, mode >= _lib.NVS_LEAKY_ADAPTIVE)
    s = io.f32(spikes, "spikes", f, True)
    active = torch.empty(f.shape, dtype=torch.uint8, device=f.device) if want_active else None
    p = _lib.NvsParams()
    p.mode, p.n_px = mode, f.numel()
    p.frame, p.ref, p.thr, p.spikes = f.data_ptr(), r.data_ptr(), t.data_ptr(), s.data_ptr()
    p.active = active.data_ptr() if active is not None else None
    p.leak_mask = leak_mask.data_ptr() if leak_mask is not None else None
    keep = [f, r, t, s, active, leak_mask, leak_mask_off]
    if mode == _lib.NVS_BI_NOISY_LEAKY_ADAPTIVE:
        ro = io.f32(reference_off, "reference_off", f, True)
        to = io.f32(thresholds_off, "thresholds_off", f, True)
        so = io.f32(spikes_off, "spikes_off", f, True) if spikes_off is not None else None
        p.ref_off, p.thr_off = ro.data_ptr(), to.data_ptr()
        p.spikes_off = so.data_ptr() if so is not None else None
        p.leak_mask_off = leak_mask_off.data_ptr() if leak_mask_off is not None else None
        keep += [ro, to, so]
    p.reference_leak, p.threshold_base = float(reference_leak), float(threshold_base)
    p.threshold_mult_incr, p.threshold_leak, p.target_off = float(threshold_mult_incr), float(threshold_leak), float(target_off)
    check(lib.dvs_nvs_step_f32(C.byref(p), _stream()))
    io.finish()
    if active is None:
        return None
    return active if isinstance(frame, torch.Tensor) else active.cpu().numpy().reshape(frame.shape)


def thresholded_difference(frame, reference, thresholds, spikes):
    """``numba_nvs.py:6-20``: ``spikes <- ((frame - reference) * active // thresholds) * thresholds`` with
    ``active = |frame - reference| >= thresholds``; returns ``active`` (uint8).  ``reference`` is not modified."""
    io = _Planes()
    f = io.f32(frame, "frame")
    r = io.f32(reference, "reference", f).clone()  # nvs_0 on a scratch copy of the reference = the difference pass
    t = io.f32(thresholds, "thresholds", f)
    s = io.f32(spikes, "spikes", f, True)
    out = _run(_lib.NVS_0, f, r, t, s)
    io.finish()
    return out if isinstance(frame, torch.Tensor) else out.cpu().numpy().reshape(frame.shape)


def nvs_0(frame, reference, thresholds, spikes):
    """``numba_nvs.py:23-29``: ``reference += spikes``."""
    _run(_lib.NVS_0, frame, reference, thresholds, spikes, want_active=False)


def nvs_leaky(frame, reference, thresholds, spikes, reference_leak):
    """``numba_nvs.py:32-40``: ``reference = reference * reference_leak + spikes``."""
    _run(_lib.NVS_LEAKY, frame, reference, thresholds, spikes, reference_leak, want_active=False)


def nvs_leaky_adaptive(frame, reference, thresholds, spikes, reference_leak, threshold_base, threshold_mult_incr,
                       threshold_leak):
    """``numba_nvs.py:43-60``: thresholds and reference are leaky integrators."""
    _run(_lib.NVS_LEAKY_ADAPTIVE, frame, reference, thresholds, spikes, reference_leak, threshold_base,
         threshold_mult_incr, threshold_leak, want_active=False)


def nvs_noisy_leaky_adaptive(frame, reference, thresholds, spikes, reference_leak, threshold_base,
                             threshold_mult_incr, threshold_leak, reference_leak_probability, leak_mask=None,
                             generator=None):
    """``numba_nvs.py:64-90``: the reference leaks only where ``uniform(0,1) <= reference_leak_probability``."""
    like = frame if isinstance(frame, torch.Tensor) else torch.empty(frame.shape, device="cuda")
    m = _mask(leak_mask, reference_leak_probability, like, generator)
    _run(_lib.NVS_NOISY_LEAKY_ADAPTIVE, frame, reference, thresholds, spikes, reference_leak, threshold_base,
         threshold_mult_incr, threshold_leak, leak_mask=m, want_active=False)


def nvs_bi_noisy_leaky_adaptive(frame, reference, thresholds, spikes, reference_leak, threshold_base,
                                threshold_mult_incr, threshold_leak, reference_leak_probability, reference_off,
                                thresholds_off, target_off, leak_mask=None, leak_mask_off=None, generator=None,
                                spikes_off=None):
    """``numba_nvs.py:93-147``: an ON and an OFF cell; the OFF reference relaxes towards ``target_off`` from the
    updated ON reference (``numba_nvs.py:133-135``).  ``spikes_off`` (a local in the reference) is optional."""
    like = frame if isinstance(frame, torch.Tensor) else torch.empty(frame.shape, device="cuda")
    m = _mask(leak_mask, reference_leak_probability, like, generator)
    mo = _mask(leak_mask_off, reference_leak_probability, like, generator)
    _run(_lib.NVS_BI_NOISY_LEAKY_ADAPTIVE, frame, reference, thresholds, spikes, reference_leak, threshold_base,
         threshold_mult_incr, threshold_leak, leak_mask=m, reference_off=reference_off,
         thresholds_off=thresholds_off, target_off=target_off, leak_mask_off=mo, spikes_off=spikes_off,
         want_active=False)
