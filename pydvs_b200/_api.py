"""Host side of the drop-in module API (``pydvs.generate_spikes`` / ``pydvs.generate_spikes_uncoded``).

``build_api(uncoded)`` returns the module-level callables with the reference's names and positional
signatures; ``generate_spikes.py`` and ``generate_spikes_uncoded.py`` export
them.  Every function validates like the Cython buffer declarations do, allocates outputs with torch
on the CUDA device and calls the ``extern "C"`` kernels of ``libpydvs_b200.so`` on the current stream.

Inputs may be CUDA tensors (``[H, W]`` like the reference, or ``[S, H, W]`` batches) -- outputs are then
CUDA tensors -- or NumPy arrays, which are uploaded, processed on the GPU and downloaded (an I/O
adapter so that ``stand_alone*.py`` run unchanged; there is no CPU arithmetic path).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .spike_lists import SpikeLists

DTYPE = np.int16
DTYPE_IDX = np.int64
DTYPE_U8 = np.uint8
DTYPE_U16 = np.uint16
DTYPE_FLOAT = float

UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY = 0, 1, 2, 3
KEY_SPINNAKER, KEY_XYP = 0, 1
_CHUNK = 1024  # records per tile when encoding from (rows; cols; vals) lists

_NP2T = {np.dtype(np.int16): torch.int16, np.dtype(np.uint16): torch.uint16, np.dtype(np.uint8): torch.uint8,
         np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}
_CY_NAME = {torch.int16: "short", torch.uint16: "unsigned short", torch.uint8: "unsigned char",
            torch.float64: "double", torch.float32: "float", torch.int32: "int", torch.int64: "long"}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pydvs_b200 needs a CUDA device: the drop-in module has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class _Arg:
    """A validated array argument on the device + how to hand results back."""

    def __init__(self, x, name, dtype, ndims, expected="DTYPE_t"):
        self.src = x
        self.is_numpy = isinstance(x, np.ndarray)
        if self.is_numpy:
            tdt = _NP2T.get(x.dtype)
            if tdt is None or tdt != dtype:
                raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (expected, x.dtype.name))
            nd = x.ndim
        elif isinstance(x, torch.Tensor):
            if x.dtype != dtype:
                raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'"
                                 % (expected, _CY_NAME.get(x.dtype, str(x.dtype))))
            nd = x.dim()
        else:
            raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray or torch.Tensor, got %s)"
                            % (name, type(x).__name__))
        if nd not in ndims:
            raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndims[0], nd))
        if self.is_numpy:
            arr = np.ascontiguousarray(x)
            if dtype == torch.uint16:
                t = torch.from_numpy(arr.view(np.int16)).to(_device()).view(torch.uint16)
            else:
                t = torch.from_numpy(arr).to(_device())
        else:
            if not x.is_cuda:
                raise ValueError("pydvs_b200: torch inputs must be CUDA tensors (argument '%s')" % name)
            t = x if x.is_contiguous() else x.contiguous()
        self.t = t


def _back(t, like_numpy):
    if not like_numpy:
        return t
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _c_int(v, name, lo, hi, cname):
    """Scalar conversion with the OverflowError / TypeError behaviour of Cython's C integer args."""
    if isinstance(v, (torch.Tensor, np.ndarray)) and getattr(v, "ndim", 0) == 0:
        v = v.item()
    if isinstance(v, (float, np.floating, str, bytes)) or v is None:
        raise TypeError("an integer is required")
    try:
        iv = v.__index__() if hasattr(v, "__index__") else int(v)
    except (TypeError, ValueError):
        raise TypeError("an integer is required")
    if iv < lo or iv > hi:
        raise OverflowError("value too large to convert to %s" % cname)
    return iv


def _i16(v, name="value"):
    return _c_int(v, name, -32768, 32767, "npy_int16")


def _u8(v, name="value"):
    if isinstance(v, (int, np.integer)) and int(v) < 0:
        raise OverflowError("can't convert negative value to npy_uint8")
    return _c_int(v, name, 0, 255, "npy_uint8")


def _same_shape(*args):
    shp = tuple(args[0].t.shape)
    for a in args[1:]:
        if tuple(a.t.shape) != shp:
            raise ValueError("operands could not be broadcast together with shapes %s %s" % (shp, tuple(a.t.shape)))
    return shp


def _geom(shape):
    if len(shape) == 2:
        return 1, shape[0], shape[1]
    return shape[0], shape[1], shape[2]


def _lut_arg(log2_table):
    if isinstance(log2_table, np.ndarray):
        if log2_table.dtype != np.uint8:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_U8_t' but got '%s'" % log2_table.dtype.name)
        if log2_table.ndim != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % log2_table.ndim)
        return torch.from_numpy(np.ascontiguousarray(log2_table)).to(_device())
    if isinstance(log2_table, torch.Tensor):
        if log2_table.dtype != torch.uint8:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_U8_t' but got '%s'" % log2_table.dtype)
        if log2_table.dim() != 1:
            raise ValueError("Buffer has wrong number of dimensions (expected 1, got %d)" % log2_table.dim())
        return log2_table.to(_device()).contiguous()
    raise TypeError("Argument 'log2_table' has incorrect type")


def _raise_status(status):
    st = int(status.item())
    if st & (_lib.STATUS_LUT_INDEX | _lib.STATUS_INDEX):
        raise IndexError("index out of bounds for the log2 table / noise sample")
    if st & _lib.STATUS_SLOT_RANGE:
        raise IndexError("list index out of range")
    if st & _lib.STATUS_EVENT_OVERFLOW:
        raise OverflowError("event buffer too small")


def generate_log2_table(max_active_bits, bit_resolution):
    """``generate_spikes.pyx:1105-1118`` (+ ``encode_time_n_bits_single`` ``:477-500``): uint8 table
    ``[max_active_bits, 2**bit_resolution]``; row r keeps the top ``max(r, 1)`` set bits of the value
    (the uint8 cast wraps bits above 7 away, as ``DTYPE_U8(np.power(2, max_pow))`` does).
    One-off host-side setup; the row a caller selects is uploaded when it is first used."""
    n = 1 << int(bit_resolution)
    values = np.arange(n, dtype=np.int64)
    table = np.zeros((int(max_active_bits), n), dtype=np.uint8)
    for row in range(int(max_active_bits)):
        keep = max(row, 1)
        rest = values.copy()
        acc = np.zeros(n, dtype=np.int64)
        for _ in range(keep):
            nz = rest > 0
            top = np.zeros(n, dtype=np.int64)
            top[nz] = 1 << np.floor(np.log2(rest[nz])).astype(np.int64)
            acc |= top
            rest = values - (acc & 0xff)  # the reference subtracts the uint8-truncated accumulator
            rest[~nz] = 0
        table[row] = (acc & 0xff).astype(np.uint8)
    return table


def generate_inh_coords(width, height, inh_width):
    """``generate_spikes.pyx:155-173``: the permutation that lays every ``k x k`` area out as a row.
    Kept for signature compatibility -- the CUDA inhibition kernels index areas directly."""
    width, height, k = _i16(width), _i16(height), _i16(inh_width)
    new_w, new_h, ratio = k * k, (height * width) // (k * k), width // k
    u = np.repeat(np.arange(new_h, dtype=np.int64), new_w)
    v = np.tile(np.arange(new_w, dtype=np.int64), new_h)
    return np.stack([k * (u // ratio) + v // k, k * (u % ratio) + v % k], axis=1).astype(DTYPE)


def build_api(uncoded):
    """Module namespace of the coded (False) or uncoded (True) drop-in module."""
    UNC = int(bool(uncoded))
    api = {}

    def export(fn):
        api[fn.__name__] = fn
        return fn

    # ---- thresholded_difference, gs:49-78 / gs:83-113 ------------------------------------------
    def _thresholded(curr_frame, ref_frame, thr_plane, thr_scalar):
        c = _Arg(curr_frame, "curr_frame", torch.int16, (2, 3))
        r = _Arg(ref_frame, "ref_frame", torch.int16, (2, 3))
        args = [c, r]
        tp = None
        if thr_plane is not None:
            tp = _Arg(thr_plane, "threshold", torch.int16, (2, 3))
            args.append(tp)
        _same_shape(*args)
        diff, abs_diff, spikes = (torch.empty_like(c.t) for _ in range(3))
        check(lib.dvs_thresholded_difference_i16(_ptr(c.t), _ptr(r.t), _ptr(tp.t) if tp else None, thr_scalar,
                                                 _ptr(diff), _ptr(abs_diff), _ptr(spikes), c.t.numel(), _stream()))
        np_out = c.is_numpy
        return _back(diff, np_out), _back(abs_diff, np_out), _back(spikes, np_out)

    @export
    def thresholded_difference(curr_frame, ref_frame, threshold):
        return _thresholded(curr_frame, ref_frame, None, _i16(threshold, "threshold"))

    @export
    def thresholded_difference_adpt(curr_frame, ref_frame, threshold, min_threshold, max_threshold):
        _i16(min_threshold), _i16(max_threshold)  # accepted and unused, like gs:83-113
        return _thresholded(curr_frame, ref_frame, threshold, 0)

    api["generate_inh_coords"] = generate_inh_coords
    api["generate_log2_table"] = generate_log2_table

    # ---- local_inhibition, gs:177-221: mutates `spikes`, returns the same object ---------------
    @export
    def local_inhibition(spikes, abs_diff, inh_coords, width, height, inh_width):
        s = _Arg(spikes, "spikes", torch.int16, (2, 3))
        a = _Arg(abs_diff, "abs_diff", torch.int16, (2, 3))
        if isinstance(inh_coords, (np.ndarray, torch.Tensor)):
            if (inh_coords.ndim if isinstance(inh_coords, np.ndarray) else inh_coords.dim()) != 2:
                raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % inh_coords.ndim)
        width, height, k = _i16(width), _i16(height), _i16(inh_width)
        shp = _same_shape(s, a)
        S, H, W = _geom(shp)
        if (H, W) != (height, width):
            raise IndexError("index out of bounds: frame is %dx%d, arguments say %dx%d" % (W, H, width, height))
        if k <= 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        if W % k or H % k:
            raise IndexError("inhibition areas of width %d do not tile a %dx%d frame" % (k, W, H))
        work = s.t  # the caller's tensor itself when it is a contiguous CUDA tensor
        check(lib.dvs_local_inhibition_i16(_ptr(work), _ptr(a.t), S, H, W, k, _stream()))
        if s.is_numpy:
            spikes[...] = work.cpu().numpy()
        elif work is not spikes:
            spikes.copy_(work)
        return spikes

    # ---- update_reference_*, gs:225-473 ---------------------------------------------------------
    def _update(mode, abs_diff, spikes, ref_frame, threshold_scalar, max_time_ms, history_weight,
                log2_table=None, thr_plane=None, adpt=None):
        a = _Arg(abs_diff, "abs_diff", torch.int16, (2, 3))
        s = _Arg(spikes, "spikes", torch.int16, (2, 3))
        r = _Arg(ref_frame, "ref_frame", torch.int16, (2, 3))
        _same_shape(a, s, r)
        hw = float(history_weight)
        lut = _lut_arg(log2_table) if log2_table is not None else None
        ref_out = torch.empty_like(r.t)
        status = torch.zeros(1, dtype=torch.int32, device=r.t.device)
        thr_t = thr_out = None
        mn = mx = dn = up = 0
        if mode == _lib.UPD_RATE_ADPT:
            th = _Arg(thr_plane, "threshold", torch.int16, (2, 3))
            _same_shape(a, th)
            mn, mx, dn, up = adpt
            thr_t = th.t  # the caller's tensor itself when it is a contiguous CUDA tensor
            thr_out = thr_t if UNC else torch.empty_like(thr_t)
        check(lib.dvs_update_reference_i16(mode, UNC, _ptr(a.t), _ptr(s.t), _ptr(r.t), _ptr(ref_out), _ptr(thr_t),
                                           _ptr(thr_out), threshold_scalar, mn, mx, dn, up, max_time_ms, hw,
                                           _ptr(lut), lut.numel() if lut is not None else 0, r.t.numel(),
                                           _ptr(status), _stream()))
        if lut is not None:
            _raise_status(status)
        np_out = r.is_numpy
        if mode != _lib.UPD_RATE_ADPT:
            return _back(ref_out, np_out)
        # the caller's threshold array is left holding the in-place result (gs:292-295)
        if isinstance(thr_plane, np.ndarray):
            thr_plane[...] = thr_t.cpu().numpy()
            ret_thr = thr_plane if UNC else _back(thr_out, True)
        else:
            if thr_t is not thr_plane:
                thr_plane.copy_(thr_t)
            ret_thr = thr_plane if UNC else thr_out
        return _back(ref_out, np_out), ret_thr

    @export
    def update_reference_rate(abs_diff, spikes, ref_frame, threshold, max_time_ms, history_weight):
        return _update(_lib.UPD_RATE, abs_diff, spikes, ref_frame, _i16(threshold), _i16(max_time_ms), history_weight)

    @export
    def update_reference_rate_adpt(abs_diff, spikes, ref_frame, threshold, min_threshold, max_threshold,
                                   down_threshold_change, up_threshold_change, max_time_ms, history_weight):
        adpt = (_i16(min_threshold), _i16(max_threshold), _i16(down_threshold_change), _i16(up_threshold_change))
        return _update(_lib.UPD_RATE_ADPT, abs_diff, spikes, ref_frame, 0, _i16(max_time_ms), history_weight,
                       thr_plane=threshold, adpt=adpt)

    @export
    def update_reference_time_thresh(abs_diff, spikes, ref_frame, threshold, max_time_ms, history_weight):
        return _update(_lib.UPD_RATE, abs_diff, spikes, ref_frame, _i16(threshold), _i16(max_time_ms), history_weight)

    @export
    def update_reference_time_binary_raw(abs_diff, spikes, ref_frame, threshold, max_time_ms, num_spikes,
                                         history_weight, log2_table):
        _i16(num_spikes)
        return _update(_lib.UPD_TIME_BIN_RAW, abs_diff, spikes, ref_frame, _i16(threshold), _i16(max_time_ms),
                       history_weight, log2_table=log2_table)

    @export
    def update_reference_time_binary_thresh(abs_diff, spikes, ref_frame, threshold, max_time_ms, num_spikes,
                                            history_weight, log2_table):
        _i16(num_spikes)
        return _update(_lib.UPD_TIME_BIN_THR, abs_diff, spikes, ref_frame, _i16(threshold), _i16(max_time_ms),
                       history_weight, log2_table=log2_table)

    def _noise(thresh_variant, abs_diff, spikes, ref_frame, width, height, threshold, max_time_ms, history_weight,
               log2_table, noise_probability, sample=None):
        a = _Arg(abs_diff, "abs_diff", torch.int16, (2,))
        s = _Arg(spikes, "spikes", torch.int16, (2,))
        r = _Arg(ref_frame, "ref_frame", torch.int16, (2,))
        _same_shape(a, s, r)
        width, height = _i16(width), _i16(height)
        H, W = r.t.shape
        if (H, W) != (height, width):
            raise IndexError("frame is %dx%d, arguments say %dx%d" % (W, H, width, height))
        if sample is None:
            # same draw as gs:441-442 / gs:464-465 from the global NumPy RNG, int16-cast
            num = int((float(_i16(noise_probability)) / 100.) * (width * height))
            sample = np.random.randint(0, height * width, size=num).astype(DTYPE)
        sample_t = torch.from_numpy(np.ascontiguousarray(sample, dtype=np.int16)).to(r.t.device)
        lut = _lut_arg(log2_table) if log2_table is not None else None
        ref_out = torch.empty_like(r.t)
        flags = torch.empty(H * W, dtype=torch.uint8, device=r.t.device)
        status = torch.zeros(1, dtype=torch.int32, device=r.t.device)
        check(lib.dvs_update_reference_noise_i16(thresh_variant, _ptr(a.t), _ptr(s.t), _ptr(r.t), _ptr(ref_out), W, H,
                                                 _i16(threshold), _i16(max_time_ms), float(history_weight), _ptr(lut),
                                                 lut.numel() if lut is not None else 0, _ptr(sample_t),
                                                 sample_t.numel(), _ptr(flags), _ptr(status), _stream()))
        _raise_status(status)
        return _back(ref_out, r.is_numpy)

    @export
    def update_reference_time_binary_raw_noise(abs_diff, spikes, ref_frame, width, height, threshold, max_time_ms,
                                               num_spikes, history_weight, log2_table, noise_probability,
                                               _sample=None):
        _i16(num_spikes)
        return _noise(0, abs_diff, spikes, ref_frame, width, height, threshold, max_time_ms, history_weight,
                      log2_table, noise_probability, _sample)

    if not UNC:  # coded module only, gs:453-473
        @export
        def update_reference_time_thresh_noise(abs_diff, spikes, ref_frame, width, height, threshold, max_time_ms,
                                               history_weight, noise_probability, _sample=None):
            return _noise(1, abs_diff, spikes, ref_frame, width, height, threshold, max_time_ms, history_weight,
                          None, noise_probability, _sample)

    # ---- split_spikes, gs:651-719 / gsu:623-672 ---------------------------------------------------
    @export
    def split_spikes(spikes, abs_diff, polarity):
        s = _Arg(spikes, "spikes", torch.int16, (2,))
        a = _Arg(abs_diff, "abs_diff", torch.int16, (2,))
        polarity = _u8(polarity, "polarity")
        _same_shape(s, a)
        H, W = s.t.shape
        dev = s.t.device
        out_dt = torch.uint16 if UNC else torch.int16
        if H * W == 0:
            e = torch.empty((3, 0), dtype=out_dt, device=dev)
            return _back(e, s.is_numpy), _back(e.clone(), s.is_numpy), 0
        tr = lib.dvs_choose_tile_rows(1, H, W, 1)
        if tr <= 0:
            raise NotImplementedError("frame too wide (width > 32768)")
        tps = (H + tr - 1) // tr
        tile_px = tr * W
        tiling = _lib.Tiling(1, H, W, tr)
        records = torch.empty(tps * tile_px, dtype=torch.int64, device=dev)
        counts = torch.empty(2 * tps, dtype=torch.int32, device=dev)
        excl = torch.empty_like(counts)
        totals = torch.empty(2, dtype=torch.int32, device=dev)
        gmax = torch.full((2,), -2 ** 31, dtype=torch.int32, device=dev)
        enc = _lib.EncParams()
        enc.mode, enc.uncoded = _lib.ENC_NONE, UNC
        st = _stream()
        check(lib.dvs_split_tiles(C.byref(tiling), _ptr(s.t), _ptr(a.t), polarity, C.byref(enc), _ptr(records),
                                  _ptr(counts), _ptr(gmax), None, st))
        check(lib.dvs_scan_tiles(1, tps, 0, _ptr(counts), _ptr(excl), _ptr(totals), None, st))
        host = torch.cat([totals, gmax]).cpu().numpy()  # one synchronising copy
        n_pos, n_neg, mx_pos, mx_neg = (int(v) for v in host)
        pos = torch.empty((3, n_pos), dtype=torch.int16, device=dev)
        neg = torch.empty((3, n_neg), dtype=torch.int16, device=dev)
        check(lib.dvs_gather_split(0, tps, tile_px, 0, _ptr(records), _ptr(counts), _ptr(excl), _ptr(pos),
                                   max(n_pos, 1), _ptr(neg), max(n_neg, 1), st))
        # global_max, gs:679-713 / gsu:650-665
        have_p, have_n = n_pos > 0, n_neg > 0
        g = 0
        if not UNC or polarity == MERGED_POLARITY:
            if have_p and have_n:
                g = max(mx_pos, mx_neg)
            elif have_n:
                g = mx_neg
            elif have_p:
                g = mx_pos
        elif polarity == UP_POLARITY:
            g = mx_pos if have_p else 0
        else:
            g = mx_neg if have_n else 0
        if UNC:
            pos, neg = pos.view(torch.uint16), neg.view(torch.uint16)
        return _back(neg, s.is_numpy), _back(pos, s.is_numpy), g

    # ---- make_spike_lists_*, gs:783-1099 / gsu:710-998 ---------------------------------------------
    def _encode(mode, pos_spikes, neg_spikes, threshold, n_slots, flag_shift, data_shift, data_mask, log2_table,
                key_coding):
        u16_in = bool(UNC and mode in (_lib.ENC_RATE, _lib.ENC_TIME))
        in_dt = torch.uint16 if u16_in else torch.int16
        exp = "DTYPE_U16_t" if u16_in else "DTYPE_t"
        pos = _Arg(pos_spikes, "pos_spikes", in_dt, (2,), exp)
        neg = _Arg(neg_spikes, "neg_spikes", in_dt, (2,), exp)
        flag_shift, data_shift, data_mask = _u8(flag_shift), _u8(data_shift), _u8(data_mask)
        if not UNC and _u8(key_coding) != KEY_SPINNAKER:
            # gs:731-733 assigns an ndarray to a C int16: the reference raises TypeError
            raise TypeError("only length-1 arrays can be converted to Python scalars")
        if n_slots < 0:
            raise OverflowError("can't convert negative value to unsigned int")
        if n_slots > _lib.limits().max_slots:
            raise NotImplementedError("more than %d time slots" % _lib.limits().max_slots)
        if pos.t.shape[0] < 3 or neg.t.shape[0] < 3:
            raise IndexError("spike lists must be [3, N] (rows; cols; vals)")
        n_pos, n_neg = int(pos.t.shape[1]), int(neg.t.shape[1])
        triples = bool(UNC and mode in (_lib.ENC_RATE, _lib.ENC_TIME))
        np_out = pos.is_numpy
        dev = pos.t.device
        if mode != _lib.ENC_TIME_BIN and threshold == 0 and (n_pos + n_neg) > 0:
            raise ZeroDivisionError("integer division or modulo by zero")
        if n_slots == 0 or (n_pos + n_neg) == 0:
            if n_slots == 0 and (n_pos + n_neg) > 0 and mode != _lib.ENC_RATE:
                raise IndexError("list index out of range")
            empty = [[] for _ in range(n_slots)]
            return empty if np_out else SpikeLists(np.zeros(n_slots + 1, dtype=np.int64),
                                                   torch.empty((0, 3) if triples else (0,), dtype=torch.int32,
                                                               device=dev))
        lut = _lut_arg(log2_table) if log2_table is not None else None
        enc = _lib.EncParams()
        enc.mode, enc.uncoded, enc.threshold, enc.n_slots, enc.u16_vals = mode, UNC, threshold or 1, n_slots, int(u16_in)
        if mode == _lib.ENC_TIME_BIN:
            enc.threshold = 1
        enc.lut_len = int(lut.numel()) if lut is not None else 0
        enc.lut = lut.data_ptr() if lut is not None else None
        tiles = max(1, (n_pos + _CHUNK - 1) // _CHUNK + (n_neg + _CHUNK - 1) // _CHUNK)
        Cn = 2 + 2 * n_slots
        records = torch.empty(tiles * _CHUNK, dtype=torch.int64, device=dev)
        counts = torch.empty(Cn * tiles, dtype=torch.int32, device=dev)
        excl = torch.empty_like(counts)
        totals = torch.empty(Cn, dtype=torch.int32, device=dev)
        seg = torch.empty(2 * n_slots + 1, dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        st = _stream()
        pos16, neg16 = pos.t.view(torch.int16), neg.t.view(torch.int16)
        check(lib.dvs_lists_to_tiles(_ptr(pos16), n_pos, pos16.stride(0), _ptr(neg16), n_neg, neg16.stride(0), _CHUNK,
                                     C.byref(enc), _ptr(records), _ptr(counts), _ptr(status), st))
        check(lib.dvs_scan_tiles(1, tiles, n_slots, _ptr(counts), _ptr(excl), _ptr(totals), _ptr(seg), st))
        seg_host = torch.cat([seg, status.to(torch.int64)]).cpu().numpy()  # one synchronising copy
        if seg_host[-1]:
            _raise_status(status)
        seg_host = seg_host[:-1]
        n_ev = int(seg_host[-1])
        events = torch.empty(max(n_ev, 1), dtype=torch.int64, device=dev)
        check(lib.dvs_aer_expand(1, tiles, _CHUNK, C.byref(enc), _ptr(records), _ptr(counts), _ptr(excl), _ptr(seg),
                                 _ptr(events), n_ev, _ptr(status), st))
        slot_off = seg_host[::2].copy()
        if triples:
            ev = events[:n_ev].view(torch.int16).view(-1, 4)
            x = ev[:, 0].to(torch.int32) & 0xffff
            y = ev[:, 1].to(torch.int32) & 0xffff
            payload = torch.stack([y, x, ev[:, 3].to(torch.int32)], dim=1)
        else:
            payload = torch.empty(n_ev, dtype=torch.int16, device=dev)
            check(lib.dvs_keys_from_events(_ptr(events), n_ev, UNC, flag_shift, data_shift, data_mask, _ptr(payload), st))
        out = SpikeLists(slot_off, payload)
        return out.tolist() if np_out else out

    if not UNC:
        @export
        def make_spike_lists_rate(pos_spikes, neg_spikes, global_max, threshold, flag_shift, data_shift, data_mask,
                                  max_time_ms, key_coding=KEY_SPINNAKER):
            _i16(global_max)
            return _encode(_lib.ENC_RATE, pos_spikes, neg_spikes, _i16(threshold), _i16(max_time_ms), flag_shift,
                           data_shift, data_mask, None, key_coding)

        @export
        def make_spike_lists_time(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask, num_bins,
                                  max_time_ms, min_threshold, max_threshold, key_coding=KEY_SPINNAKER):
            _i16(global_max), _i16(max_time_ms), _i16(max_threshold)
            return _encode(_lib.ENC_TIME, pos_spikes, neg_spikes, _i16(min_threshold), _i16(num_bins), flag_shift,
                           data_shift, data_mask, None, key_coding)

        @export
        def make_spike_lists_time_bin(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask,
                                      max_time_ms, min_threshold, max_threshold, num_bins, log2_table,
                                      key_coding=KEY_SPINNAKER):
            _i16(global_max), _i16(max_time_ms), _i16(min_threshold), _i16(max_threshold)
            return _encode(_lib.ENC_TIME_BIN, pos_spikes, neg_spikes, 1, _i16(num_bins), flag_shift, data_shift,
                           data_mask, log2_table, key_coding)

        @export
        def make_spike_lists_time_bin_thr(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask,
                                          max_time_ms, min_threshold, max_threshold, num_bins, log2_table,
                                          key_coding=KEY_SPINNAKER):
            _i16(global_max), _i16(max_time_ms), _i16(max_threshold)
            return _encode(_lib.ENC_TIME_BIN_THR, pos_spikes, neg_spikes, _i16(min_threshold), _i16(num_bins),
                           flag_shift, data_shift, data_mask, log2_table, key_coding)
    else:
        @export
        def make_spike_lists_rate(pos_spikes, neg_spikes, global_max, threshold, flag_shift, data_shift, data_mask,
                                  max_time_ms):
            _i16(global_max)
            return _encode(_lib.ENC_RATE, pos_spikes, neg_spikes, _i16(threshold), _i16(max_time_ms), flag_shift,
                           data_shift, data_mask, None, KEY_SPINNAKER)

        @export
        def make_spike_lists_time(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask, num_bins,
                                  max_time_ms, min_threshold, max_threshold):
            _i16(global_max), _i16(max_time_ms), _i16(max_threshold)
            return _encode(_lib.ENC_TIME, pos_spikes, neg_spikes, _i16(min_threshold), _i16(num_bins), flag_shift,
                           data_shift, data_mask, None, KEY_SPINNAKER)

        @export
        def make_spike_lists_time_bin(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask,
                                      max_time_ms, min_threshold, max_threshold, num_bins, log2_table):
            _i16(global_max), _i16(max_time_ms), _i16(min_threshold), _i16(max_threshold)
            return _encode(_lib.ENC_TIME_BIN, pos_spikes, neg_spikes, 1, _i16(num_bins), flag_shift, data_shift,
                           data_mask, log2_table, KEY_SPINNAKER)

        @export
        def make_spike_lists_time_bin_thr(pos_spikes, neg_spikes, global_max, flag_shift, data_shift, data_mask,
                                          max_time_ms, min_threshold, max_threshold, num_bins, log2_table):
            _i16(global_max), _i16(max_time_ms), _i16(max_threshold)
            return _encode(_lib.ENC_TIME_BIN_THR, pos_spikes, neg_spikes, _i16(min_threshold), _i16(num_bins),
                           flag_shift, data_shift, data_mask, log2_table, KEY_SPINNAKER)

    return api
