"""float32 mode: the reference's only native fused kernel, ``DVSOperator::operator()``
(``pydvs/cpp/dvs_op.cpp:4-33``), and the state-owning object around it, ``PyDVS``
(``pydvs/cpp/dvs_emu.hpp:89-108``, ``dvs_emu.cpp:126-143``), on B200.

Per pixel, all float32 and evaluated in this order with no FMA contraction:
``d = src - ref; t = (d < -thr) or (d > thr); d = d * t; ref = relax * ref + d; thr *= up if t else down``.
"""
import ctypes as C

import torch

from ._lib import lib, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dvs_op(src, diff, ref, thr, relax, up, down):
    """In place on ``diff`` (may be None), ``ref`` and ``thr``; any shape, float32 CUDA tensors."""
    for name, t in (("src", src), ("ref", ref), ("thr", thr)) + ((("diff", diff),) if diff is not None else ()):
        if not isinstance(t, torch.Tensor) or t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("%s must be a contiguous float32 CUDA tensor" % name)
        if t.numel() != src.numel():
            raise ValueError("%s has %d elements, src has %d" % (name, t.numel(), src.numel()))
    check(lib.dvs_step_f32(C.c_void_p(src.data_ptr()), C.c_void_p(diff.data_ptr()) if diff is not None else None,
                           C.c_void_p(ref.data_ptr()), C.c_void_p(thr.data_ptr()), float(relax), float(up),
                           float(down), src.numel(), _stream()))


class DVSOperatorF32:
    """Owns ``ref`` / ``thr`` / ``diff`` planes for ``S`` streams like ``PyDVS`` does for one camera
    (``dvs_emu.hpp:91-104``: ``ref = 0``, ``thr = max(thr_init, base_thresh)``)."""

    def __init__(self, height, width, streams=1, thr_init=10.0, relax=0.999, up=1.5, down=0.99, base_thresh=0.0,
                 device="cuda", keep_diff=True):
        shape = (streams, height, width)
        self.relax, self.up, self.down = float(relax), float(up), float(down)
        self.ref = torch.zeros(shape, dtype=torch.float32, device=device)
        self.thr = torch.full(shape, max(float(thr_init), float(base_thresh)), dtype=torch.float32, device=device)
        self.diff = torch.zeros(shape, dtype=torch.float32, device=device) if keep_diff else None

    def update(self, frame):
        """``frame``: float32 CUDA tensor [S, H, W] (grayscale converted to CV_32F, dvs_emu.cpp:133)."""
        dvs_op(frame, self.diff, self.ref, self.thr, self.relax, self.up, self.down)
        return self.diff
