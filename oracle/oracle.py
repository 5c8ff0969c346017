"""ctypes front-end of ``oracle/dvs_oracle.c`` -- the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``pydvs_b200/`` does.

``Oracle(uncoded=False)`` exposes the same function names and positional signatures as the reference's
``pydvs.generate_spikes`` (``uncoded=True``: ``pydvs.generate_spikes_uncoded``), NumPy in / NumPy (or
list-of-lists) out, so parity tests read like calls into the reference.  Parity of this oracle itself
is pinned against the compiled reference (``oracle/_ref``) and the golden vectors in ``tests/golden``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "dvs_oracle.c")
_LIB = os.path.join(_HERE, "libdvs_oracle.so")

ERRORS = {1: IndexError, 2: MemoryError, 3: ZeroDivisionError, 4: IndexError, 5: OverflowError,
          6: ValueError}
UP, DOWN, MERGED, RECTIFIED = 0, 1, 2, 3
ENC_RATE, ENC_TIME, ENC_TIME_BIN, ENC_TIME_BIN_THR = 0, 1, 2, 3


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fwrapv",
                               "-o", _LIB, _SRC, "-lm"])
    return _LIB


class EncParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("uncoded", C.c_int), ("u16_in", C.c_int), ("threshold", C.c_int),
                ("n_slots", C.c_int), ("flag_shift", C.c_int), ("data_shift", C.c_int),
                ("data_mask", C.c_int), ("lut", C.c_void_p), ("lut_len", C.c_int)]


class PipeParams(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("adaptive", C.c_int), ("update_mode", C.c_int),
                ("threshold", C.c_int), ("min_thr", C.c_int), ("max_thr", C.c_int), ("down", C.c_int),
                ("up", C.c_int), ("max_time_ms", C.c_int), ("hw", C.c_double), ("inh_k", C.c_int),
                ("polarity", C.c_int), ("uncoded", C.c_int), ("enc", EncParams)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_spike_to_key.restype = C.c_int16
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i16(a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.int16:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % a.dtype)
    return a


def _check(rc):
    if rc:
        raise ERRORS.get(rc, RuntimeError)("oracle error %d" % rc)


def generate_log2_table(max_active_bits, bit_resolution):
    t = np.zeros((max_active_bits, 2 ** bit_resolution), dtype=np.uint8)
    lib().orc_generate_log2_table(C.c_int(max_active_bits), C.c_int(bit_resolution), _p(t))
    return t


def dvs_op_f32(src, diff, ref, thr, relax, up, down):
    """In place on diff/ref/thr, like DVSOperator::operator() (cpp/dvs_op.cpp:4-33)."""
    for a in (src, diff, ref, thr):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().orc_dvs_op_f32(_p(src), _p(diff), _p(ref), _p(thr), C.c_float(relax), C.c_float(up),
                         C.c_float(down), C.c_long(src.size))


class Oracle:
    """Function-for-function mirror of the reference module API on the C restatement."""

    def __init__(self, uncoded=False):
        self.uncoded = bool(uncoded)
        self.L = lib()

    # -- A.1 ------------------------------------------------------------------------------------
    def thresholded_difference(self, curr_frame, ref_frame, threshold):
        c, r = _i16(curr_frame), _i16(ref_frame)
        d, a, s = (np.empty_like(c) for _ in range(3))
        self.L.orc_thresholded_difference(_p(c), _p(r), None, C.c_int(int(threshold)), _p(d), _p(a),
                                          _p(s), C.c_long(c.size))
        return d, a, s

    def thresholded_difference_adpt(self, curr_frame, ref_frame, threshold, min_threshold,
                                    max_threshold):
        c, r, t = _i16(curr_frame), _i16(ref_frame), _i16(threshold)
        d, a, s = (np.empty_like(c) for _ in range(3))
        self.L.orc_thresholded_difference(_p(c), _p(r), _p(t), C.c_int(0), _p(d), _p(a), _p(s),
                                          C.c_long(c.size))
        return d, a, s

    # -- A.2 ------------------------------------------------------------------------------------
    @staticmethod
    def generate_inh_coords(width, height, inh_width):
        k = int(inh_width)
        new_w, new_h, ratio = k * k, (height * width) // (k * k), width // k
        u, v = np.meshgrid(np.arange(new_h), np.arange(new_w), indexing="ij")
        rows = k * (u // ratio) + v // k
        cols = k * (u % ratio) + v % k
        return np.stack([rows.ravel(), cols.ravel()], axis=1).astype(np.int16)

    def local_inhibition(self, spikes, abs_diff, inh_coords, width, height, inh_width):
        assert spikes.dtype == np.int16 and spikes.flags.c_contiguous
        a = _i16(abs_diff)
        _check(self.L.orc_local_inhibition(_p(spikes), _p(a), C.c_int(width), C.c_int(height),
                                           C.c_int(inh_width)))
        return spikes

    # -- A.3 ------------------------------------------------------------------------------------
    def update_reference_rate(self, abs_diff, spikes, ref_frame, threshold, max_time_ms,
                              history_weight):
        a, s, r = _i16(abs_diff), _i16(spikes), _i16(ref_frame)
        out = np.empty_like(r)
        self.L.orc_update_reference_rate(_p(a), _p(s), _p(r), C.c_int(threshold), C.c_int(max_time_ms),
                                         C.c_double(history_weight), _p(out), C.c_long(r.size))
        return out

    update_reference_time_thresh = update_reference_rate

    def update_reference_rate_adpt(self, abs_diff, spikes, ref_frame, threshold, min_threshold,
                                   max_threshold, down_threshold_change, up_threshold_change,
                                   max_time_ms, history_weight):
        a, s, r = _i16(abs_diff), _i16(spikes), _i16(ref_frame)
        assert threshold.dtype == np.int16 and threshold.flags.c_contiguous
        out, thr_out = np.empty_like(r), np.empty_like(r)
        self.L.orc_update_reference_rate_adpt(
            _p(a), _p(s), _p(r), _p(threshold), C.c_int(min_threshold), C.c_int(max_threshold),
            C.c_int(down_threshold_change), C.c_int(up_threshold_change), C.c_double(history_weight),
            C.c_int(self.uncoded), _p(out), _p(thr_out), C.c_long(r.size))
        return out, (threshold if self.uncoded else thr_out)

    def _time_binary(self, thresh, abs_diff, spikes, ref_frame, threshold, history_weight, log2_table):
        a, s, r = _i16(abs_diff), _i16(spikes), _i16(ref_frame)
        lut = np.ascontiguousarray(log2_table, dtype=np.uint8)
        out = np.empty_like(r)
        _check(self.L.orc_update_reference_time_binary(
            _p(a), _p(s), _p(r), C.c_int(threshold), C.c_double(history_weight), _p(lut),
            C.c_int(lut.size), C.c_int(thresh), _p(out), C.c_long(r.size)))
        return out

    def update_reference_time_binary_raw(self, abs_diff, spikes, ref_frame, threshold, max_time_ms,
                                         num_spikes, history_weight, log2_table):
        return self._time_binary(0, abs_diff, spikes, ref_frame, threshold, history_weight, log2_table)

    def update_reference_time_binary_thresh(self, abs_diff, spikes, ref_frame, threshold, max_time_ms,
                                            num_spikes, history_weight, log2_table):
        return self._time_binary(1, abs_diff, spikes, ref_frame, threshold, history_weight, log2_table)

    def update_reference_noise(self, thresh_variant, abs_diff, spikes, ref_frame, width, height,
                               threshold, max_time_ms, history_weight, log2_table, sample):
        """Noise variants with the random flat indices injected (already int16-cast)."""
        a, s, r = _i16(abs_diff), _i16(spikes), _i16(ref_frame)
        lut = np.ascontiguousarray(log2_table if log2_table is not None else np.zeros(1), np.uint8)
        smp = np.ascontiguousarray(sample, dtype=np.int16)
        out = np.empty_like(r)
        _check(self.L.orc_update_reference_noise(
            _p(a), _p(s), _p(r), C.c_int(width), C.c_int(height), C.c_int(threshold),
            C.c_int(max_time_ms), C.c_double(history_weight), _p(lut), C.c_int(lut.size),
            C.c_int(thresh_variant), _p(smp), C.c_long(smp.size), _p(out)))
        return out

    generate_log2_table = staticmethod(generate_log2_table)

    # -- A.5 ------------------------------------------------------------------------------------
    def split_spikes(self, spikes, abs_diff, polarity):
        s, a = _i16(spikes), _i16(abs_diff)
        H, W = s.shape
        cap = max(1, s.size)
        neg = np.empty((3, cap), dtype=np.int16)
        pos = np.empty((3, cap), dtype=np.int16)
        nn, npos, g = C.c_long(0), C.c_long(0), C.c_int(0)
        _check(self.L.orc_split_spikes(_p(s), _p(a), C.c_int(W), C.c_int(H), C.c_int(polarity),
                                       C.c_int(self.uncoded), _p(neg), _p(pos), C.c_long(cap),
                                       C.byref(nn), C.byref(npos), C.byref(g)))
        dt = np.uint16 if self.uncoded else np.int16
        return (np.ascontiguousarray(neg[:, :nn.value]).view(dt),
                np.ascontiguousarray(pos[:, :npos.value]).view(dt), g.value)

    # -- A.6 ------------------------------------------------------------------------------------
    def spike_to_key(self, row, col, flag_shift, data_shift, data_mask, is_pos):
        return int(self.L.orc_spike_to_key(C.c_int(row), C.c_int(col), C.c_int(flag_shift),
                                           C.c_int(data_shift), C.c_int(data_mask), C.c_int(is_pos),
                                           C.c_int(self.uncoded)))

    # -- A.7 / A.8 ------------------------------------------------------------------------------
    def _enc_params(self, mode, threshold, n_slots, flag_shift, data_shift, data_mask, lut, u16_in):
        p = EncParams()
        p.mode, p.uncoded, p.u16_in = mode, int(self.uncoded), int(u16_in)
        p.threshold, p.n_slots = int(threshold), int(n_slots)
        p.flag_shift, p.data_shift, p.data_mask = int(flag_shift), int(data_shift), int(data_mask)
        if lut is not None:
            self._lut_keep = np.ascontiguousarray(lut, dtype=np.uint8)
            p.lut, p.lut_len = self._lut_keep.ctypes.data, self._lut_keep.size
        else:
            p.lut, p.lut_len = None, 0
        return p

    def _encode(self, p, pos, neg):
        pos = np.ascontiguousarray(pos)
        neg = np.ascontiguousarray(neg)
        n_pos, n_neg = pos.shape[1], neg.shape[1]
        triples = self.uncoded and p.mode in (ENC_RATE, ENC_TIME)
        per_px = max(p.n_slots, 8)
        cap = max(1, (n_pos + n_neg) * per_px)
        out = np.empty((cap, 3 if triples else 1), dtype=np.int32)
        counts = np.zeros(max(1, p.n_slots), dtype=np.int64)
        n_ev = C.c_long(0)
        _check(self.L.orc_make_spike_lists(
            C.byref(p), _p(pos.view(np.int16)), C.c_long(n_pos), C.c_long(max(1, n_pos)),
            _p(neg.view(np.int16)), C.c_long(n_neg), C.c_long(max(1, n_neg)), _p(counts), _p(out),
            C.c_long(cap), C.byref(n_ev)))
        lists, o = [], 0
        flat = out[:n_ev.value]
        for j in range(p.n_slots):
            c = int(counts[j])
            chunk = flat[o:o + c]
            lists.append(chunk.tolist() if triples else chunk[:, 0].tolist())
            o += c
        return lists

    def make_spike_lists_rate(self, pos_spikes, neg_spikes, global_max, threshold, flag_shift,
                              data_shift, data_mask, max_time_ms, key_coding=0):
        p = self._enc_params(ENC_RATE, threshold, max_time_ms, flag_shift, data_shift, data_mask, None,
                             self.uncoded)
        return self._encode(p, pos_spikes, neg_spikes)

    def make_spike_lists_time(self, pos_spikes, neg_spikes, global_max, flag_shift, data_shift,
                              data_mask, num_bins, max_time_ms, min_threshold, max_threshold,
                              key_coding=0):
        p = self._enc_params(ENC_TIME, min_threshold, num_bins, flag_shift, data_shift, data_mask, None,
                             self.uncoded)
        return self._encode(p, pos_spikes, neg_spikes)

    def make_spike_lists_time_bin(self, pos_spikes, neg_spikes, global_max, flag_shift, data_shift,
                                  data_mask, max_time_ms, min_threshold, max_threshold, num_bins,
                                  log2_table, key_coding=0):
        p = self._enc_params(ENC_TIME_BIN, min_threshold, num_bins, flag_shift, data_shift, data_mask,
                             log2_table, False)
        return self._encode(p, pos_spikes, neg_spikes)

    def make_spike_lists_time_bin_thr(self, pos_spikes, neg_spikes, global_max, flag_shift, data_shift,
                                      data_mask, max_time_ms, min_threshold, max_threshold, num_bins,
                                      log2_table, key_coding=0):
        p = self._enc_params(ENC_TIME_BIN_THR, min_threshold, num_bins, flag_shift, data_shift,
                             data_mask, log2_table, False)
        return self._encode(p, pos_spikes, neg_spikes)


# -- numba_nvs.py float variants (in place, like the reference) -------------------
NVS_0, NVS_LEAKY, NVS_LEAKY_ADAPTIVE, NVS_NOISY_LEAKY_ADAPTIVE = 0, 1, 2, 3


def _f32(*arrs):
    for a in arrs:
        assert a.dtype == np.float32 and a.flags.c_contiguous


def nvs_step(mode, frame, reference, thresholds, spikes, reference_leak=1.0, threshold_base=0.0,
             threshold_mult_incr=0.0, threshold_leak=1.0, leak_mask=None):
    """nvs_0 / nvs_leaky / nvs_leaky_adaptive / nvs_noisy_leaky_adaptive (numba_nvs.py:23-90); returns
    the activity plane (uint8) that thresholded_difference (numba_nvs.py:6-20) returns."""
    _f32(frame, reference, thresholds, spikes)
    active = np.zeros(frame.shape, np.uint8)
    if leak_mask is not None:
        leak_mask = np.ascontiguousarray(leak_mask, dtype=np.uint8)
    lib().orc_nvs_f32(C.c_int(mode), _p(frame), _p(reference), _p(thresholds), _p(spikes), _p(active),
                      _p(leak_mask) if leak_mask is not None else None, C.c_double(reference_leak),
                      C.c_double(threshold_base), C.c_double(threshold_mult_incr),
                      C.c_double(threshold_leak), C.c_long(frame.size))
    return active


def nvs_bi_step(frame, reference, thresholds, spikes, reference_leak, threshold_base, threshold_mult_incr,
                threshold_leak, reference_off, thresholds_off, target_off, leak_mask=None,
                leak_mask_off=None, spikes_off=None):
    """nvs_bi_noisy_leaky_adaptive (numba_nvs.py:93-147) with the two random draws injected."""
    _f32(frame, reference, thresholds, spikes, reference_off, thresholds_off)
    masks = [None if m is None else np.ascontiguousarray(m, dtype=np.uint8) for m in (leak_mask, leak_mask_off)]
    lib().orc_nvs_bi_f32(_p(frame), _p(reference), _p(thresholds), _p(spikes), _p(reference_off),
                         _p(thresholds_off), _p(spikes_off) if spikes_off is not None else None,
                         _p(masks[0]) if masks[0] is not None else None,
                         _p(masks[1]) if masks[1] is not None else None, C.c_double(reference_leak),
                         C.c_double(threshold_base), C.c_double(threshold_mult_incr),
                         C.c_double(threshold_leak), C.c_double(target_off), C.c_long(frame.size))


# -- decode_spikes.pyx ---------------------------------------------------------------
def key_to_spike(key, flag_shift, data_shift, data_mask):
    """decode_spikes.pyx:46-67."""
    row = (key >> data_shift) & data_mask
    col = key & data_mask
    return row, col, (1 if ((key >> flag_shift) & 1) == 0 else -1)


def update_ref_spike_list_rate(ref, spikes, threshold, history_weight, flag_shift, data_shift, data_mask):
    """decode_spikes.pyx:71-94; returns a new reference plane."""
    ref = _i16(ref)
    spikes = _i16(spikes)
    out = np.empty_like(ref)
    _check(lib().orc_decode_rate(_p(ref), _p(out), C.c_int(ref.shape[0]), C.c_int(ref.shape[1]), _p(spikes),
                                 C.c_long(spikes.size), C.c_int(threshold), C.c_double(history_weight),
                                 C.c_int(flag_shift), C.c_int(data_shift), C.c_int(data_mask)))
    return out


def _decode_time(binary, ref, time_last, val_last, spikes, num_bins, bin_width, time_period, time_now,
                 threshold, history_weight, flag_shift, data_shift, data_mask):
    ref = _i16(ref)
    spikes = _i16(spikes)
    assert time_last.dtype == np.int16 and val_last.dtype == np.int16
    assert time_last.flags.c_contiguous and val_last.flags.c_contiguous
    out = np.empty_like(ref)
    _check(lib().orc_decode_time(C.c_int(binary), _p(ref), _p(out), _p(time_last), _p(val_last),
                                 C.c_int(ref.shape[0]), C.c_int(ref.shape[1]), _p(spikes), C.c_long(spikes.size),
                                 C.c_int(num_bins), C.c_double(bin_width), C.c_int(time_period),
                                 C.c_uint32(time_now), C.c_int(threshold), C.c_double(history_weight),
                                 C.c_int(flag_shift), C.c_int(data_shift), C.c_int(data_mask)))
    return out, time_last, val_last


def update_ref_spike_list_time(*args):
    """decode_spikes.pyx:98-145; mutates and returns time_last / val_last like the reference."""
    return _decode_time(0, *args)


def update_ref_spike_list_time_bin(*args):
    """decode_spikes.pyx:149-197."""
    return _decode_time(1, *args)


class Pipeline:
    """One stream's frame loop in C (the reference's call sequence, stand_alone.py:160-185):
    threshold -> inhibition -> reference update -> split -> encode; state (ref, thr) carried."""

    def __init__(self, W, H, *, mode=ENC_RATE, threshold=12, max_time_ms=8, num_bins=None,
                 history_weight=1.0, inh_k=0, polarity=MERGED, adaptive=False, min_thr=6, max_thr=168,
                 down=2, up=12, lut=None, uncoded=False, flag_shift=0, data_shift=0, data_mask=0,
                 ref0=128, thr0=None, update_mode=None):
        self.L = lib()
        self.W, self.H, self.P = W, H, W * H
        pp = PipeParams()
        pp.W, pp.H, pp.adaptive = W, H, int(adaptive)
        if update_mode is None:
            update_mode = {ENC_RATE: 0, ENC_TIME: 0, ENC_TIME_BIN: 2, ENC_TIME_BIN_THR: 3}[mode]
        pp.update_mode = update_mode
        pp.threshold, pp.min_thr, pp.max_thr, pp.down, pp.up = threshold, min_thr, max_thr, down, up
        pp.max_time_ms, pp.hw, pp.inh_k, pp.polarity, pp.uncoded = (max_time_ms, history_weight, inh_k,
                                                                     polarity, int(uncoded))
        n_slots = max_time_ms if mode == ENC_RATE else (num_bins if num_bins is not None else max_time_ms)
        enc_thr = threshold if not adaptive else min_thr
        self._lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.uint8)
        e = pp.enc
        e.mode, e.uncoded = mode, int(uncoded)
        e.u16_in = int(uncoded and mode in (ENC_RATE, ENC_TIME))
        e.threshold, e.n_slots = enc_thr, n_slots
        e.flag_shift, e.data_shift, e.data_mask = flag_shift, data_shift, data_mask
        if self._lut is not None:
            e.lut, e.lut_len = self._lut.ctypes.data, self._lut.size
        self.pp, self.n_slots = pp, n_slots
        self.triples = uncoded and mode in (ENC_RATE, ENC_TIME)
        self.ref = np.full((H, W), ref0, dtype=np.int16)
        self.thr = np.full((H, W), threshold if thr0 is None else thr0, dtype=np.int16)
        self.scratch = np.empty(11 * self.P + 16, dtype=np.int16)
        self.cap = max(1, self.P * max(n_slots, 8))
        self.out = np.empty((self.cap, 3 if self.triples else 1), dtype=np.int32)
        self.counts = np.zeros(max(1, n_slots), dtype=np.int64)

    def step_raw(self, frame):
        """Returns (slot_counts copy, flat events view) without building Python lists."""
        f = _i16(frame)
        n_ev = C.c_long(0)
        _check(self.L.orc_pipeline_frame(C.byref(self.pp), _p(f), _p(self.ref), _p(self.thr),
                                         _p(self.scratch), _p(self.counts), _p(self.out),
                                         C.c_long(self.cap), C.byref(n_ev)))
        return self.counts.copy(), self.out[:n_ev.value]

    def step(self, frame):
        counts, flat = self.step_raw(frame)
        lists, o = [], 0
        for j in range(self.n_slots):
            c = int(counts[j])
            chunk = flat[o:o + c]
            lists.append(chunk.tolist() if self.triples else chunk[:, 0].tolist())
            o += c
        return lists
