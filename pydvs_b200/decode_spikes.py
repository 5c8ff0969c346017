"""Receiver side: ``pydvs/decode_spikes.pyx`` on the GPU.

``key_to_spike`` and the three ``update_ref_spike_list_*`` functions keep the reference's names,
positional arguments and return values (new ``ref``; ``time_last`` / ``val_last`` mutated in place and
returned).  NumPy in -> NumPy out through an upload/download adapter, CUDA tensors in -> CUDA tensors
out; all arithmetic runs in ``libpydvs_b200.so``.  ``decode_rate_batch`` is the batched form used for
encode -> decode round trips over many streams.

Reference behaviour kept: the decoder reads the *uncoded* key layout and maps
flag bit 0 to +1; keys are ``int16`` and a negative key raises ``OverflowError``; a ``val_now`` that is
nan raises ``ValueError``, one that does not fit ``int16`` raises ``OverflowError``.  A key whose
(row, col) lies outside the plane is undefined behaviour in the reference (bounds checks are off) and
raises ``IndexError`` here.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

DTYPE = np.int16
DTYPE_IDX = np.int64
DTYPE_U8 = np.uint8
DTYPE_U16 = np.uint16
DTYPE_U32 = np.uint32
DTYPE_FLOAT = float

_ERRORS = {
    _lib.DECODE_KEY_NEGATIVE: (OverflowError, "can't convert negative value to npy_uint16"),
    _lib.DECODE_INDEX: (IndexError, "decoded (row, col) lies outside the reference plane"),
    _lib.DECODE_VALUE_NAN: (ValueError, "cannot convert float NaN to integer"),
    _lib.DECODE_VALUE_RANGE: (OverflowError, "value too large to convert to npy_int16"),
}
_log2_lut = {}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _u8(v, name):
    v = int(v)
    if not 0 <= v <= 255:
        raise OverflowError("value too large to convert to npy_uint8" if v > 255 else
                            "can't convert negative value to npy_uint8")
    return v


def _i16(v, name):
    v = int(v)
    if not -32768 <= v <= 32767:
        raise OverflowError("value too large to convert to npy_int16")
    return v


def key_to_spike(key, flag_shift, data_shift, data_mask):
    """``decode_spikes.pyx:46-67``: ``(row, col, sign)`` of one 16-bit key."""
    key = int(key)
    if key < 0:
        raise OverflowError("can't convert negative value to npy_uint16")
    if key > 0xffff:
        raise OverflowError("value too large to convert to npy_uint16")
    flag_shift, data_shift, data_mask = _u8(flag_shift, "flag_shift"), _u8(data_shift, "data_shift"), _u8(data_mask, "data_mask")
    row = (key >> data_shift) & data_mask
    col = key & data_mask
    return row, col, (1 if ((key >> flag_shift) & 1) == 0 else -1)


def _plane(x, name, ndim=2):
    """-> (CUDA int16 tensor, was_numpy)."""
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.int16:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % x.dtype)
        if x.dim() != ndim:
            raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, x.dim()))
        return x.contiguous().cuda(), False
    a = np.asarray(x)
    if a.dtype != np.int16:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % a.dtype)
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, a.ndim))
    return torch.from_numpy(np.ascontiguousarray(a)).cuda(), True


def _key_list(keys, offsets, S, H, W, flag_shift, data_shift, data_mask):
    kl = _lib.KeyList()
    kl.keys, kl.stream_offsets, kl.n = keys.data_ptr(), offsets.data_ptr(), keys.numel()
    kl.S, kl.H, kl.W = S, H, W
    kl.flag_shift, kl.data_shift, kl.data_mask = flag_shift, data_shift, data_mask
    return kl


def _raise_first(first_error):
    v = int(first_error.item())
    if v != _lib.DECODE_NO_ERROR:
        exc, msg = _ERRORS[v & 3]
        raise exc("%s (key %d)" % (msg, v >> 2))


def _new_error_word(device):
    return torch.full((1,), _lib.DECODE_NO_ERROR, dtype=torch.int32, device=device)


def decode_rate_batch(ref, keys, stream_offsets, threshold, history_weight, flag_shift, data_shift, data_mask):
    """``update_ref_spike_list_rate`` for ``S`` streams at once: ``ref`` int16 CUDA ``[S, H, W]``, ``keys`` int16
    CUDA (all streams concatenated), ``stream_offsets`` int64 CUDA ``[S + 1]``.  Returns the new planes."""
    S, H, W = ref.shape
    out = torch.empty_like(ref)
    check(lib.dvs_decay_i16(C.c_void_p(ref.data_ptr()), C.c_void_p(out.data_ptr()), float(history_weight),
                            ref.numel(), _stream()))
    err = _new_error_word(ref.device)
    kl = _key_list(keys, stream_offsets, S, H, W, flag_shift, data_shift, data_mask)
    check(lib.dvs_decode_rate_i16(C.byref(kl), C.c_void_p(out.data_ptr()), int(threshold),
                                  C.c_void_p(err.data_ptr()), _stream()))
    _raise_first(err)
    return out


def update_ref_spike_list_rate(ref, spikes, threshold, history_weight, flag_shift, data_shift, data_mask):
    """``decode_spikes.pyx:71-94``: ``ref' = int16(ref * history_weight)``, then ``ref'[row, col] += threshold * sign``
    for every key (duplicates accumulate).  Returns a new plane."""
    r, was_np = _plane(ref, "ref")
    k, _ = _plane(spikes, "spikes", 1)
    threshold = _i16(threshold, "threshold")
    fs, ds, dm = _u8(flag_shift, "flag_shift"), _u8(data_shift, "data_shift"), _u8(data_mask, "data_mask")
    offsets = torch.tensor([0, k.numel()], dtype=torch.int64, device=r.device)
    out = decode_rate_batch(r.unsqueeze(0), k, offsets, threshold, history_weight, fs, ds, dm)[0]
    return out.cpu().numpy() if was_np else out


def _log2_table(device):
    key = str(device)
    if key not in _log2_lut:
        with np.errstate(divide="ignore"):
            t = np.log2(np.arange(32768, dtype=np.int64))  # NumPy's own values, as the reference computes them
        _log2_lut[key] = torch.from_numpy(t).to(device)
    return _log2_lut[key]


def _decode_time(binary, ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now, threshold,
                 history_weight, flag_shift, data_shift, data_mask):
    r, was_np = _plane(ref, "ref")
    tl, tl_np = _plane(time_last, "time_last")
    vl, vl_np = _plane(val_last, "val_last")
    k, _ = _plane(spikes, "spikes", 1)
    num_bins, time_period, threshold = _i16(num_bins, "num_bins"), _i16(time_period, "time_period"), _i16(threshold, "threshold")
    time_now = int(time_now)
    if not 0 <= time_now <= 0xffffffff:
        raise OverflowError("value too large to convert to npy_uint32" if time_now > 0 else
                            "can't convert negative value to npy_uint32")
    fs, ds, dm = _u8(flag_shift, "flag_shift"), _u8(data_shift, "data_shift"), _u8(data_mask, "data_mask")
    H, W = r.shape
    dev = r.device
    out = torch.empty_like(r)
    check(lib.dvs_decay_i16(C.c_void_p(r.data_ptr()), C.c_void_p(out.data_ptr()), float(history_weight), r.numel(),
                            _stream()))
    n = k.numel()
    done = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
    claim = torch.full((H * W,), _lib.DECODE_NO_ERROR, dtype=torch.int32, device=dev)
    remaining = torch.zeros(1, dtype=torch.int32, device=dev)
    err = _new_error_word(dev)
    offsets = torch.tensor([0, n], dtype=torch.int64, device=dev)
    kl = _key_list(k, offsets, 1, H, W, fs, ds, dm)
    lut = _log2_table(dev) if binary else None
    while True:  # one round unless a pixel repeats inside the list
        check(lib.dvs_decode_time_round_i16(C.byref(kl), int(binary), C.c_void_p(out.data_ptr()),
                                            C.c_void_p(tl.data_ptr()), C.c_void_p(vl.data_ptr()),
                                            C.c_void_p(lut.data_ptr()) if lut is not None else None, num_bins,
                                            float(bin_width), time_period, time_now, threshold,
                                            C.c_void_p(done.data_ptr()), C.c_void_p(claim.data_ptr()),
                                            C.c_void_p(remaining.data_ptr()), C.c_void_p(err.data_ptr()), _stream()))
        left = int(remaining.item())
        _raise_first(err)
        if left == 0:
            break
    if tl_np:
        time_last[...] = tl.cpu().numpy()
    elif tl.data_ptr() != time_last.data_ptr():
        time_last.copy_(tl)
    if vl_np:
        val_last[...] = vl.cpu().numpy()
    elif vl.data_ptr() != val_last.data_ptr():
        val_last.copy_(vl)
    return (out.cpu().numpy() if was_np else out), time_last, val_last


def update_ref_spike_list_time(ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now,
                               threshold, history_weight, flag_shift, data_shift, data_mask):
    """``decode_spikes.pyx:98-145``.  Keys are applied in list order per pixel, as the reference's loop does."""
    return _decode_time(0, ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now, threshold,
                        history_weight, flag_shift, data_shift, data_mask)


def update_ref_spike_list_time_bin(ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now,
                                   threshold, history_weight, flag_shift, data_shift, data_mask):
    """``decode_spikes.pyx:149-197``."""
    return _decode_time(1, ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now, threshold,
                        history_weight, flag_shift, data_shift, data_mask)
