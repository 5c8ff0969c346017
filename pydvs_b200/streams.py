"""``DVSStreams`` -- the batched, state-resident form of the spike-generation path.

One object owns the reference (and adaptive-threshold) planes of ``S`` independent camera streams in
HBM and advances all of them by one frame per ``step()`` with three launches on the current CUDA
stream (fused tile pass -> scans -> AER expand; ``include/pydvs_b200.h``).  It is the native
counterpart of the loop every reference caller writes by hand
(``stand_alone.py:160-185``, ``external_dvs_emulator_device.py:549-683``):

    thresholded_difference[_adpt] -> local_inhibition -> update_reference_* -> split_spikes
    -> make_spike_lists_*

and of the C++ ``PyDVS`` object that owns ``_ref/_thr`` (``pydvs/cpp/dvs_emu.hpp:89-108``).
PyTorch is used for device memory and streams only.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import lib, check
from .spike_lists import SpikeLists

UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY = 0, 1, 2, 3
OUTPUT_RATE, OUTPUT_TIME, OUTPUT_TIME_BIN, OUTPUT_TIME_BIN_THR = "RATE", "TIME", "TIME_BIN", "TIME_BIN_THR"

_ENC = {OUTPUT_RATE: _lib.ENC_RATE, OUTPUT_TIME: _lib.ENC_TIME, OUTPUT_TIME_BIN: _lib.ENC_TIME_BIN,
        OUTPUT_TIME_BIN_THR: _lib.ENC_TIME_BIN_THR}
# reference-update rule per output type, external_dvs_emulator_device.py:594-632
_UPD = {OUTPUT_RATE: _lib.UPD_RATE, OUTPUT_TIME: _lib.UPD_RATE, OUTPUT_TIME_BIN: _lib.UPD_TIME_BIN_RAW,
        OUTPUT_TIME_BIN_THR: _lib.UPD_TIME_BIN_THR}


def current_stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@dataclass
class DVSConfig:
    """Mirrors the constructor arguments of ``ExternalDvsEmulatorDevice``
    (``external_dvs_emulator_device.py:73-84``) plus the geometry / batching the CUDA path needs."""
    width: int
    height: int
    streams: int = 1
    output_type: str = OUTPUT_RATE
    polarity: int = MERGED_POLARITY
    threshold: int = 12
    adaptive_threshold: bool = False
    min_threshold: int = 6
    max_threshold: int = 168
    threshold_delta_down: int = 2
    threshold_delta_up: int = 12
    history_weight: float = 1.0
    fps: Optional[float] = None
    max_time_ms: Optional[int] = None          # int(1000 / fps) when None (stand_alone.py:138)
    num_bins: Optional[int] = None             # 8 TIME_BIN, 6 TIME_BIN_THR, else max_time_ms
    num_bits_per_spike: int = 4                # log2 table row = num_bits_per_spike - 1
    log2_table: Optional[object] = None        # explicit uint8 LUT row overrides the generated one
    inh_area_width: int = 0                    # <= 1: no local inhibition
    uncoded: bool = False                      # rules of generate_spikes_uncoded.pyx
    encoder_threshold: Optional[int] = None    # threshold given to make_spike_lists_*
    ref0: int = 128                            # external_dvs_emulator_device.py:133
    threshold0: Optional[int] = None           # initial adaptive threshold plane (device: zeros)
    frame_dtype: str = "int16"                 # "int16" (API parity) or "uint8" (camera-native)
    tile_rows: Optional[int] = None

    def resolved(self):
        c = DVSConfig(**self.__dict__)
        if c.max_time_ms is None:
            c.max_time_ms = int(1000.0 / c.fps) if c.fps else 8
        if c.num_bins is None:  # external_dvs_emulator_device.py:167-172
            c.num_bins = {OUTPUT_TIME_BIN: 8, OUTPUT_TIME_BIN_THR: 6}.get(c.output_type, c.max_time_ms)
        if c.encoder_threshold is None:
            c.encoder_threshold = c.min_threshold if c.adaptive_threshold else c.threshold
        if c.output_type not in _ENC:
            raise ValueError("unknown output_type %r" % (c.output_type,))
        return c


def log2_table_row(num_active_bits, bit_resolution):
    """Row ``num_active_bits - 1`` of ``generate_log2_table`` (``generate_spikes.pyx:1105-1118``):
    keep the top ``max(row, 1)`` set bits of every value in ``[0, 2**bit_resolution)``."""
    from .generate_spikes import generate_log2_table
    return generate_log2_table(num_active_bits, bit_resolution)[num_active_bits - 1]


class StepResult:
    """Device-resident AER output of one ``DVSStreams.step``.

    ``events[seg_offsets[(s * n_slots + slot) * 2 + pol] : seg_offsets[... + 1]]`` are the events of
    stream ``s``, time slot ``slot``, list ``pol`` (0 = pos, 1 = neg) in the reference's order
    (SURVEY.md Appendix A.7).  Each event is 8 bytes: uint16 x, y, t and int16 p.

    Lifetime: the result ALIASES the stream object's step buffers (no copy) and fetches the row pointers lazily,
    so it describes its step only until the next ``step()`` on the same object (``pipeline_depth`` steps when
    pipelined).  Copy ``events()`` / ``offsets_host()`` first if an older step must outlive a newer one.
    """

    def __init__(self, owner, events, seg_offsets, status, ready=None):
        self._owner, self.events_raw, self.seg_offsets, self._status = owner, events, seg_offsets, status
        self._off_host = None
        self._ready = ready   # CUDA event of the expand launch when it ran on the side stream (pipeline_depth 2)

    def wait(self):
        """Make the current stream wait for this step's expand (no-op unless the step was pipelined)."""
        if self._ready is not None:
            torch.cuda.current_stream().wait_event(self._ready)
        return self

    # -- host-side accessors (synchronise) ----------------------------------------------------
    def offsets_host(self):
        if self._off_host is None:
            self.wait()
            self._off_host = self.seg_offsets.cpu().numpy()
            self._owner.check_status()
        return self._off_host

    def total_events(self):
        return int(self.offsets_host()[-1])

    def events_per_stream(self):
        o = self.offsets_host()
        per = 2 * self._owner.n_slots
        return np.diff(o[::per]) if per else np.zeros(self._owner.S, dtype=np.int64)

    def events(self):
        """int16 view [N, 4] of (x, y, t, p); x, y, t are uint16 bit patterns."""
        n = self.total_events()
        self.wait()
        return self.events_raw[:n].view(torch.int16).view(-1, 4)

    def keys(self, flag_shift, data_shift, data_mask, stream=None):
        """The reference's int16 AER keys (``spike_to_key``) for all events or one stream's."""
        o = self.offsets_host()
        per = 2 * self._owner.n_slots
        lo, hi = (0, int(o[-1])) if stream is None else (int(o[stream * per]), int(o[(stream + 1) * per]))
        keys = torch.empty(max(hi - lo, 0), dtype=torch.int16, device=self.events_raw.device)
        if hi > lo:
            check(lib.dvs_keys_from_events(_ptr(self.events_raw[lo:hi]), hi - lo, int(self._owner.cfg.uncoded),
                                           int(flag_shift), int(data_shift), int(data_mask), _ptr(keys),
                                           current_stream_ptr()))
        return keys

    def spike_lists(self, stream, flag_shift, data_shift, data_mask):
        """What ``make_spike_lists_*`` returns for one stream: a list of ``n_slots`` lists."""
        o = self.offsets_host()
        n_slots = self._owner.n_slots
        per = 2 * n_slots
        slot_off = o[stream * per:(stream + 1) * per + 1:2].copy()
        cfg = self._owner.cfg
        if cfg.uncoded and cfg.output_type in (OUTPUT_RATE, OUTPUT_TIME):
            lo, hi = int(slot_off[0]), int(slot_off[-1])
            ev = self.events_raw[lo:hi].view(torch.int16).view(-1, 4)
            x = ev[:, 0].to(torch.int32) & 0xffff
            y = ev[:, 1].to(torch.int32) & 0xffff
            payload = torch.stack([y, x, ev[:, 3].to(torch.int32)], dim=1)
        else:
            payload = self.keys(flag_shift, data_shift, data_mask, stream=stream)
        return SpikeLists(slot_off, payload)


class DVSStreams:
    # up to this many pixels per step (all streams) a step is replayed as a CUDA graph once its frame buffer has been
    # seen: launch-bound for small problems (128x128: 0.06 -> 0.014 ms), and still 1-5 % at 32-256 streams of 4K (the
    # gaps between the tile pass, the redo launch, the scan and the expand: 256 x 4K 1.505 -> 1.486 ms per step)
    AUTO_GRAPH_MAX_PX = 1 << 32
    GRAPH_CACHE = 16          # distinct frame buffers (by address) that get their own captured graph

    def __init__(self, cfg: DVSConfig, device=None, event_capacity: Optional[int] = None,
                 debug_planes: bool = False, force_generic: bool = False, pipeline_depth: int = 1,
                 split_lists: bool = True, graph="auto", state_dtype: str = "auto"):
        if not torch.cuda.is_available():
            raise RuntimeError("pydvs_b200 needs a CUDA device (there is no CPU path)")
        self.cfg = cfg = cfg.resolved()
        self.device = torch.device(device if device is not None else "cuda")
        self.S, self.H, self.W = cfg.streams, cfg.height, cfg.width
        self.P = self.H * self.W
        k = cfg.inh_area_width if cfg.inh_area_width and cfg.inh_area_width > 1 else 1
        self.inh_k = k
        if k > 1 and (self.W % k or self.H % k):
            raise IndexError("local inhibition needs width and height divisible by inh_area_width")
        # Row pitch of the resident planes.  The register-only inhibition kernels need rows that start on 16-byte
        # boundaries (W % 8 == 0); a sensor such as the DAVIS346 (346 x 260) with 2x2 inhibition would otherwise take the
        # generic body at a twentieth of the speed.  The state planes are ours, so they are kept with rows padded to a
        # multiple of 8 pixels and every frame is copied into a padded buffer first (one extra pass over the frame):
        # pad pixels hold 0 in frame and reference, never exceed a threshold >= 0, never spike, never reach an event.
        t0_ = cfg.threshold if cfg.threshold0 is None else cfg.threshold0
        pad_ok = (k in (2, 4) and self.W % 8 != 0 and not debug_planes and not force_generic and cfg.tile_rows is None
                  and int(cfg.threshold) >= 0 and (not cfg.adaptive_threshold or (int(cfg.min_threshold) >= 0 and int(t0_) >= 0)))
        self.Wp = (self.W + 7) // 8 * 8 if pad_ok else self.W
        self.enc_mode = _ENC[cfg.output_type]
        self.upd_mode = _lib.UPD_RATE_ADPT if cfg.adaptive_threshold else _UPD[cfg.output_type]
        self.n_slots = cfg.max_time_ms if cfg.output_type == OUTPUT_RATE else cfg.num_bins
        if cfg.encoder_threshold == 0 and self.enc_mode != _lib.ENC_TIME_BIN:
            raise ZeroDivisionError("integer division or modulo by zero")
        with torch.cuda.device(self.device):
            tr = cfg.tile_rows or lib.dvs_choose_tile_rows(self.S, self.H, self.Wp, k)
        if tr <= 0:
            raise NotImplementedError("frame too wide for one tile (width * inh_area_width > 32768)")
        self.tile_rows = tr
        self.tps = (self.H + tr - 1) // tr
        self.tile_px = tr * self.Wp
        # records reserved per tile: with k x k inhibition at most one record per area (tile_px / k^2)
        self.rec_stride = (tr // k) * (self.Wp // k) if k > 1 else self.tile_px
        self.C = 2 + 2 * self.n_slots
        dev = self.device
        # state.  "int16": the planes as the reference's API holds them (``ref`` / ``thr`` ARE the resident tensors).
        # "uint8": one byte per pixel -- every update rule clips the reference to [0, 255] (gs:251, 297, 378, 425) and
        # the coded threshold rule to [min, max] (gs:297), so for a start inside [0, 255] the bytes hold the reference's
        # own values and a step moves half the state bytes (``ref`` / ``thr`` then return int16 COPIES: read them
        # freely, write with set_state()).  "auto" (the default): uint8 whenever that is exact for this configuration,
        # int16 otherwise (ref0 / thresholds outside a byte, the uncoded threshold rule, debug planes).
        t0 = cfg.threshold if cfg.threshold0 is None else cfg.threshold0
        compact_ok = (0 <= int(cfg.ref0) <= 255 and not debug_planes and
                      (not cfg.adaptive_threshold or
                       (not cfg.uncoded and 0 <= int(cfg.min_threshold) <= int(cfg.max_threshold) <= 255
                        and 0 <= int(t0) <= 255)))
        if state_dtype == "auto":
            state_dtype = "uint8" if compact_ok else "int16"
        if state_dtype not in ("int16", "uint8"):
            raise ValueError("state_dtype must be 'int16', 'uint8' or 'auto'")
        if state_dtype == "uint8" and not compact_ok:
            raise ValueError("one-byte state needs ref0 (and threshold0, min/max_threshold of the coded adaptive rule) in "
                             "[0, 255] and no debug planes")
        self.state_dtype = state_dtype
        sdt = torch.uint8 if state_dtype == "uint8" else torch.int16
        self._ref = torch.full((self.S, self.H, self.Wp), cfg.ref0, dtype=sdt, device=dev)
        self._thr = None
        if cfg.adaptive_threshold:
            self._thr = torch.full((self.S, self.H, self.Wp), t0, dtype=sdt, device=dev)
        self._fpad = {}            # frame dtype -> padded frame buffer (rows of Wp pixels, pad columns zero)
        if self.Wp != self.W:
            self._ref[:, :, self.W:] = 0
        # LUT (time-binary modes)
        self.lut = None
        if self.enc_mode in (_lib.ENC_TIME_BIN, _lib.ENC_TIME_BIN_THR) or self.upd_mode in (
                _lib.UPD_TIME_BIN_RAW, _lib.UPD_TIME_BIN_THR):
            row = cfg.log2_table
            if row is None:
                row = log2_table_row(cfg.num_bits_per_spike, 8)
            if isinstance(row, torch.Tensor):
                row = row.detach().cpu().numpy()
            self.lut = torch.from_numpy(np.ascontiguousarray(row, dtype=np.uint8)).to(dev)
        # work buffers.  pipeline_depth 2 double-buffers everything K3 reads or writes, so that the expand of
        # step n runs on a side stream while the tile pass of step n+1 already runs on the caller's stream.
        n_tiles = self.S * self.tps
        if event_capacity is None:
            # capacity policy: one event per two pixels (4 B/px of HBM) covers the dense start-up frame of the
            # reference's defaults; a step that overflows sets a status bit, check_status() raises OverflowError and
            # grow_events() re-expands the kept records.  Callers that know their event volume pass a small
            # event_capacity and size it with reserve_events() (bench.py: measured steady state x 1.5)
            event_capacity = max(1024, (self.S * self.P) // 2)
        self.event_capacity = int(event_capacity)
        self.pipeline_depth = max(1, int(pipeline_depth))

        def make_set():
            counts = torch.zeros(self.S * self.C * self.tps, dtype=torch.int32, device=dev)
            return {"records": torch.empty(n_tiles * self.rec_stride, dtype=torch.int64, device=dev),
                    "tile_counts": counts, "tile_excl": torch.zeros_like(counts),
                    "stream_totals": torch.zeros(self.S * self.C, dtype=torch.int32, device=dev),
                    "seg_offsets": torch.zeros(self.S * self.n_slots * 2 + 1, dtype=torch.int64, device=dev),
                    "ticket": torch.zeros(1, dtype=torch.int32, device=dev),
                    "events": torch.empty(self.event_capacity, dtype=torch.int64, device=dev), "k3_done": None}

        self._sets = [make_set() for _ in range(self.pipeline_depth)]
        self._cur = 0
        self._side = torch.cuda.Stream(device=dev) if self.pipeline_depth > 1 else None
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.redo = torch.zeros(2 + n_tiles, dtype=torch.int32, device=dev)   # count, CTA ticket, declined tiles
        self.debug = None
        if debug_planes:
            self.debug = {n: torch.zeros((self.S, self.H, self.W), dtype=torch.int16, device=dev)
                          for n in ("diff", "abs_diff", "spikes")}
        self.force_generic = int(force_generic)  # 0/1; 3 = persistent staged fast kernel (see dvs_step_params)
        # split_lists=False: the per-tile record lists only feed the AER expand, so spikes that emit no event
        # (coded rate: k == 0) need not be recorded; events are identical, split_spikes() is unavailable.
        self.split_lists = bool(split_lists)
        self._enc = self._make_enc()
        self._params = self._make_params()
        self.frames_done = 0
        # CUDA-graph replay of launch-bound steps: "auto" = small problems only, True = always, False = never
        if graph == "auto":
            graph = self.S * self.P <= self.AUTO_GRAPH_MAX_PX
        self.use_graph = bool(graph) and self.pipeline_depth == 1
        self._graphs = {}          # frame buffer address -> captured step
        self._graph_misses = 0

    # state planes as int16 (the reference's dtype): the resident tensors themselves, or copies of the one-byte state
    # (with padded rows: the [S, H, W] part of the planes)
    def _plane(self, t):
        if t is None:
            return None
        t = t if self.Wp == self.W else t[:, :, :self.W]
        return t if t.dtype == torch.int16 else t.to(torch.int16)

    ref = property(lambda self: self._plane(self._ref))
    thr = property(lambda self: self._plane(self._thr))

    def ref_plane(self, s):
        """Reference plane of stream ``s`` as int16 [H, W] (a copy when the state is held in bytes)."""
        return self._ref[s, :, :self.W].to(torch.int16)

    def thr_plane(self, s):
        return None if self._thr is None else self._thr[s, :, :self.W].to(torch.int16)

    def set_state(self, ref=None, thr=None):
        """Overwrite the reference and / or threshold planes (int16 values; with one-byte state they must lie in
        [0, 255], which is where every update rule leaves them)."""
        for dst, src, name in ((self._ref, ref, "ref"), (self._thr, thr, "thr")):
            if src is None:
                continue
            if dst is None:
                raise ValueError("no %s plane in this configuration" % name)
            src = torch.as_tensor(src, device=self.device)
            if dst.dtype == torch.uint8 and src.dtype != torch.uint8 and (int(src.min()) < 0 or int(src.max()) > 255):
                raise ValueError("%s values outside [0, 255] cannot be held in one-byte state" % name)
            dst[:, :, :self.W].copy_(src.reshape(self.S, self.H, self.W))

    # the buffer set of the most recent step
    records = property(lambda self: self._sets[self._cur]["records"])
    tile_counts = property(lambda self: self._sets[self._cur]["tile_counts"])
    tile_excl = property(lambda self: self._sets[self._cur]["tile_excl"])
    stream_totals = property(lambda self: self._sets[self._cur]["stream_totals"])
    seg_offsets = property(lambda self: self._sets[self._cur]["seg_offsets"])
    events = property(lambda self: self._sets[self._cur]["events"])

    # -- parameter blocks ---------------------------------------------------------------------
    def _make_enc(self):
        e = _lib.EncParams()
        cfg = self.cfg
        e.mode, e.uncoded, e.threshold, e.n_slots = self.enc_mode, int(cfg.uncoded), int(cfg.encoder_threshold), self.n_slots
        e.u16_vals = int(cfg.uncoded and self.enc_mode in (_lib.ENC_RATE, _lib.ENC_TIME))
        e.lut_len = int(self.lut.numel()) if self.lut is not None else 0
        e.lut = self.lut.data_ptr() if self.lut is not None else None
        return e

    def _make_params(self):
        cfg, p = self.cfg, _lib.StepParams()
        p.tiling = _lib.Tiling(self.S, self.H, self.Wp, self.tile_rows)
        p.frame_dtype = _lib.FRAME_U8 if cfg.frame_dtype == "uint8" else _lib.FRAME_I16
        p.adaptive, p.update_mode, p.uncoded = int(cfg.adaptive_threshold), self.upd_mode, int(cfg.uncoded)
        p.threshold, p.min_thr, p.max_thr = int(cfg.threshold), int(cfg.min_threshold), int(cfg.max_threshold)
        p.thr_down, p.thr_up, p.max_time_ms = int(cfg.threshold_delta_down), int(cfg.threshold_delta_up), int(cfg.max_time_ms)
        p.inh_k, p.polarity, p.history_weight = self.inh_k, int(cfg.polarity), float(cfg.history_weight)
        p.force_generic = int(self.force_generic)
        p.drop_silent = int(not self.split_lists)
        p.record_stride = self.rec_stride
        p.enc = self._enc
        p.state_dtype = _lib.STATE_U8 if self.state_dtype == "uint8" else _lib.STATE_I16
        p.ref = self._ref.data_ptr()
        p.thr = self._thr.data_ptr() if self._thr is not None else None
        if self.debug:
            p.diff_out, p.abs_diff_out, p.spikes_out = (self.debug[n].data_ptr() for n in ("diff", "abs_diff", "spikes"))
        p.status, p.redo = self.status.data_ptr(), self.redo.data_ptr()   # records / tile_counts: set per step
        return p

    # -- the hot path -------------------------------------------------------------------------
    def step(self, frames: torch.Tensor, expand: bool = True) -> StepResult:
        """Advance every stream by one frame.  ``frames``: CUDA tensor [S, H, W] (or [H, W] when
        S == 1) of the configured frame dtype.  Asynchronous: three kernel launches, no host sync."""
        # (an int16 configuration also takes uint8 frames -- the same values in half the bytes, see HostFeed)
        ok = (torch.uint8,) if self.cfg.frame_dtype == "uint8" else (torch.int16, torch.uint8)
        if frames.dtype not in ok or not frames.is_cuda:
            raise ValueError("frames must be a CUDA tensor of dtype %s" % (ok[0],))
        if frames.numel() != self.S * self.P or not frames.is_contiguous():
            raise ValueError("frames must be contiguous with shape [%d, %d, %d]" % (self.S, self.H, self.W))
        # (the first step runs eagerly: module load, attributes, the padded frame buffer of this dtype)
        if self.use_graph and expand and self.frames_done > 0 and (self.Wp == self.W or frames.dtype in self._fpad):
            g = self._graphs.get(frames.data_ptr())
            if g is None and len(self._graphs) < self.GRAPH_CACHE:
                g = self._capture_step(frames)
            if g is not None:
                self._cur = 0
                self.frames_done += 1
                g[0].replay()
                return StepResult(self, self.events, self.seg_offsets, self.status)
            self._graph_misses += 1
            if self._graph_misses > 4 * self.GRAPH_CACHE:   # the caller hands a fresh buffer every frame: stop trying
                self.use_graph = False
        self.begin_step()
        self.tile_pass(frames)
        self.scan()
        ready = self.expand_async() if expand else None
        return StepResult(self, self.events, self.seg_offsets, self.status, ready)

    def _capture_step(self, frames):
        """Record the launches of one step reading the buffer at ``frames.data_ptr()`` into a CUDA graph.  The graph is
        only replayed for a validated frames tensor at that same address, so no reference to ``frames`` is kept."""
        if self.Wp != self.W and frames.dtype not in self._fpad:   # (allocated outside the capture)
            self._fpad[frames.dtype] = torch.zeros((self.S, self.H, self.Wp), dtype=frames.dtype, device=self.device)
        graph = torch.cuda.CUDAGraph()
        self._cur = 0
        with torch.cuda.graph(graph):
            self.tile_pass(frames)
            self.scan()
            self.expand()
        entry = (graph,)
        self._graphs[frames.data_ptr()] = entry
        return entry

    def reserve_events(self, capacity):
        """Make room for ``capacity`` events per step (re-allocates; drops captured graphs)."""
        if int(capacity) <= self.event_capacity:
            return
        self.drain()
        self.event_capacity = int(capacity)
        for st in self._sets:
            st["events"] = torch.empty(self.event_capacity, dtype=torch.int64, device=self.device)
        self._graphs.clear()

    # the launches of a step, individually (bench.py times K1 on its own)
    def begin_step(self):
        """Select the buffer set of the new step; when pipelined, wait for the expand that last used it."""
        self._cur = self.frames_done % self.pipeline_depth
        self.frames_done += 1
        done = self._sets[self._cur]["k3_done"]
        if done is not None:
            torch.cuda.current_stream().wait_event(done)

    def expand_async(self):
        """K3 on the side stream when pipeline_depth > 1 (returns the event to wait on), else in line."""
        if self._side is None:
            self.expand()
            return None
        fork = torch.cuda.Event()
        fork.record()
        self._side.wait_event(fork)
        with torch.cuda.stream(self._side):
            self.expand()
            done = torch.cuda.Event()
            done.record()
        self._sets[self._cur]["k3_done"] = done
        return done

    def drain(self):
        """Make the current stream wait for every outstanding side-stream expand."""
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    def tile_pass(self, frames):
        """K1: fused threshold / inhibition / state update / record compaction / slot counts."""
        if self.Wp != self.W:   # rows padded to a multiple of 8 pixels (see __init__)
            buf = self._fpad.get(frames.dtype)
            if buf is None:
                buf = self._fpad[frames.dtype] = torch.zeros((self.S, self.H, self.Wp), dtype=frames.dtype, device=self.device)
            check(lib.dvs_pad_rows(frames.data_ptr(), buf.data_ptr(), self.S * self.H, self.W, self.Wp, frames.element_size(),
                                   current_stream_ptr()))
            frames = buf
        self._params.curr = frames.data_ptr()
        self._params.frame_dtype = _lib.FRAME_U8 if frames.dtype == torch.uint8 else _lib.FRAME_I16
        self._params.records, self._params.tile_counts = self.records.data_ptr(), self.tile_counts.data_ptr()
        check(lib.dvs_step_fused(C.byref(self._params), current_stream_ptr()))
        self.kernel_path = lib.dvs_last_step_path()   # _lib.PATH_*: which tile kernel ran (generic body or a fast path)

    def scan(self):
        """K2: per-tile counters -> exclusive prefixes and CSR row pointers (one launch: the last CTA of the tile scan
        builds the row pointers)."""
        check(lib.dvs_scan_tiles_ticket(self.S, self.tps, self.n_slots, _ptr(self.tile_counts), _ptr(self.tile_excl),
                                        _ptr(self.stream_totals), _ptr(self.seg_offsets),
                                        _ptr(self._sets[self._cur]["ticket"]), current_stream_ptr()))

    # -- CUDA-graph replay for launch-bound (small) configurations ---------------------------------
    def capture(self, frames: torch.Tensor):
        """Capture one step (all launches) reading from the static buffer ``frames`` into a CUDA graph.
        Copy each new frame into ``frames`` and call ``replay()``: one graph launch instead of six
        launches from Python (128x128 / 640x480 single streams are launch-bound)."""
        if self.pipeline_depth != 1:
            raise ValueError("graph capture uses the single-buffered pipeline (pipeline_depth=1)")
        if self.frames_done == 0:
            # warm-up outside the capture (function attributes, lazy module load) on a snapshot of the state, so
            # that capturing leaves ref / thr / frames_done exactly as they were: replaying the first frame
            # afterwards applies it once, not twice
            keep = (self._ref.clone(), None if self._thr is None else self._thr.clone())
            use_graph, self.use_graph = self.use_graph, False
            self.step(frames)
            self.use_graph = use_graph
            torch.cuda.synchronize()
            self._ref.copy_(keep[0])
            if self._thr is not None:
                self._thr.copy_(keep[1])
            self.status.zero_()
            self.frames_done = 0
        self._graph, self._graph_frames = self._capture_step(frames)[0], frames
        return self._graph

    def replay(self) -> StepResult:
        """Replay the step captured by :meth:`capture` (the result aliases the step buffers: valid until the next step)."""
        self._cur = 0
        self._graph.replay()
        self.frames_done += 1
        return StepResult(self, self.events, self.seg_offsets, self.status)

    def expand(self):
        check(lib.dvs_aer_expand(self.S, self.tps, self.rec_stride, C.byref(self._enc), _ptr(self.records),
                                 _ptr(self.tile_counts), _ptr(self.tile_excl), _ptr(self.seg_offsets),
                                 _ptr(self.events), self.event_capacity, _ptr(self.status), current_stream_ptr()))

    def grow_events(self, capacity):
        """Re-allocate the event buffer and redo the expand of the last step (records are kept)."""
        self.drain()
        self.event_capacity = int(capacity)
        for st in self._sets:
            st["events"] = torch.empty(self.event_capacity, dtype=torch.int64, device=self.device)
        self._graphs.clear()   # captured steps point at the old buffer
        self.status.zero_()
        self.expand()
        return StepResult(self, self.events, self.seg_offsets, self.status)

    def check_status(self):
        """Synchronises; raises what the reference would have raised for data-dependent faults."""
        st = int(self.status.item())
        if st:
            self.status.zero_()
        if st & _lib.STATUS_EVENT_OVERFLOW:
            raise OverflowError("event buffer too small: call grow_events(capacity)")
        if st & (_lib.STATUS_LUT_INDEX | _lib.STATUS_INDEX):
            raise IndexError("index out of range for the log2 table")
        if st & _lib.STATUS_SLOT_RANGE:
            raise IndexError("time slot outside the list of lists (undefined in the reference)")

    # -- split_spikes view of the last step ---------------------------------------------------
    def split_spikes(self, stream=0):
        """(neg[3, Nn], pos[3, Np]) of the last step for one stream, as ``split_spikes`` returns
        them (``generate_spikes.pyx:716-719``)."""
        if not self.split_lists:
            raise RuntimeError("split_spikes() needs DVSStreams(..., split_lists=True): silent spikes were not recorded")
        tot = self.stream_totals.view(self.S, self.C)[stream, :2].cpu().numpy()
        n_pos, n_neg = int(tot[0]), int(tot[1])
        pos = torch.empty((3, n_pos), dtype=torch.int16, device=self.device)
        neg = torch.empty((3, n_neg), dtype=torch.int16, device=self.device)
        check(lib.dvs_gather_split(stream, self.tps, self.rec_stride, self.n_slots, _ptr(self.records),
                                   _ptr(self.tile_counts), _ptr(self.tile_excl), _ptr(pos), max(n_pos, 1),
                                   _ptr(neg), max(n_neg, 1), current_stream_ptr()))
        return neg, pos
