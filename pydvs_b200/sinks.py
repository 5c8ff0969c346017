"""AER sinks -- the step AFTER the hot path: the text spike-file format of
``mnist_to_spikes.py:290-310`` / ``sequence_to_spikes.py:292-311`` (one ``"<key> <time_ms>"`` line per event,
slot j of a frame stamped ``t_frame + j * t_bin_ms``) and the per-slot pacing rule of
``ExternalDvsEmulatorDevice.calculate_time_per_pack`` (``external_dvs_emulator_device.py:700-715``).

Host-side formatting over the CSR result of a step (one D2H copy of keys + offsets per stream); works on a
``SpikeLists`` / ``StepResult`` from the CUDA path as well as on plain lists of lists.
"""
import io

import numpy as np

from .spike_lists import SpikeLists
from .streams import OUTPUT_RATE, OUTPUT_TIME, OUTPUT_TIME_BIN


def _keys_and_offsets(lists):
    if isinstance(lists, SpikeLists):
        keys = lists.payload
        keys = keys.cpu().numpy() if hasattr(keys, "cpu") else np.asarray(keys)
        off = lists.slot_offsets
        off = off.cpu().numpy() if hasattr(off, "cpu") else np.asarray(off)
        return keys, np.asarray(off, dtype=np.int64) - int(off[0])
    off = np.cumsum([0] + [len(x) for x in lists]).astype(np.int64)
    keys = np.array([k for sl in lists for k in sl], dtype=np.int64)
    return keys, off


def spike_file_lines(lists, t_frame_ms, t_bin_ms):
    """The lines ``mnist_to_spikes.py:290-298`` appends for one frame: ``"%s %f" % (key, t + t_idx)``."""
    keys, off = _keys_and_offsets(lists)
    if keys.ndim != 1:
        raise ValueError("spike files hold 16-bit keys; the uncoded rate/time encoders emit (row, col, p) triples")
    slots = np.repeat(np.arange(len(off) - 1), np.diff(off))
    times = float(t_frame_ms) + slots.astype(np.float64) * float(t_bin_ms)
    return ["%s %f" % (int(k), t) for k, t in zip(keys.tolist(), times.tolist())]


class SpikeFileWriter:
    """Accumulates frames of one stream like the reference's ``write_buff`` and writes them joined by newlines
    (no trailing newline), ``mnist_to_spikes.py:308-310``."""

    def __init__(self, frame_time_ms, t_bin_ms):
        self.frame_time_ms, self.t_bin_ms = frame_time_ms, t_bin_ms
        self.t = 0
        self.lines = []

    def add_frame(self, lists):
        self.lines += spike_file_lines(lists, self.t, self.t_bin_ms)
        self.t += self.frame_time_ms

    def add_lines(self, lines):
        """Lines already formatted for this frame's time stamp (``spike_file_lines_all_streams`` with ``self.t``)."""
        self.lines += lines
        self.t += self.frame_time_ms

    def getvalue(self):
        return "\n".join(self.lines)

    def write(self, path):
        with io.open(path, "w") as f:
            f.write(self.getvalue())


def calculate_time_per_pack(output_type, max_time_ms, min_threshold, max_threshold, num_bits_per_spike):
    """Milliseconds between two slot packets, ``external_dvs_emulator_device.py:700-715``."""
    if output_type == OUTPUT_RATE:
        return 1
    if output_type == OUTPUT_TIME:
        return max_time_ms // min(max_time_ms, max_threshold // min_threshold + 1)
    if output_type == OUTPUT_TIME_BIN:
        return max_time_ms // 8
    return max_time_ms // (num_bits_per_spike + 1)


def emit_spikes(sender, lists, time_per_spike_pack_ms, label, clock=None, sleep=None):
    """Slot-paced emission, ``external_dvs_emulator_device.py:532-545``: one ``sender.send_spikes(label, pack,
    send_full_keys=False)`` per time slot, then sleep for the rest of ``time_per_spike_pack_ms``.  ``lists`` is
    what ``make_spike_lists_*`` / ``StepResult.spike_lists`` return; ``clock`` / ``sleep`` default to
    ``time.time`` / ``time.sleep`` and can be injected (tests, simulated time).  Returns the number of packs sent."""
    import time
    clock = clock or time.time
    sleep = sleep or time.sleep
    max_time_s = time_per_spike_pack_ms / 1000.
    sent = 0
    if lists is not None:
        for spike_pack in lists:
            start = clock()
            sender.send_spikes(label, spike_pack, send_full_keys=False)
            elapsed = clock() - start
            if elapsed < max_time_s:
                sleep(max_time_s - elapsed)
            sent += 1
    return sent


def spike_file_lines_all_streams(result, flag_shift, data_shift, data_mask, t_frame_ms, t_bin_ms):
    """``spike_file_lines`` for every stream of a ``StepResult`` at once: one key kernel launch and two device-to-host
    copies (keys, CSR row pointers) for the whole batch instead of one round trip per stream.  Returns a list with one
    list of lines per stream."""
    owner = result._owner
    if owner.cfg.uncoded and owner.cfg.output_type in (OUTPUT_RATE, OUTPUT_TIME):
        raise ValueError("spike files hold 16-bit keys; the uncoded rate/time encoders emit (row, col, p) triples")
    off = np.asarray(result.offsets_host(), dtype=np.int64)
    keys = result.keys(flag_shift, data_shift, data_mask).cpu().numpy()
    n_slots, S = owner.n_slots, owner.S
    per = 2 * n_slots
    # events are ordered stream -> slot -> (pos, neg): the slot of event e is the segment it falls in, halved
    seg_len = np.diff(off)
    slot_of_seg = (np.arange(S * per) % per) // 2
    times = float(t_frame_ms) + np.repeat(slot_of_seg, seg_len).astype(np.float64) * float(t_bin_ms)
    bounds = off[::per] if per else np.zeros(S + 1, dtype=np.int64)
    out = []
    for s in range(S):
        lo, hi = int(bounds[s]), int(bounds[s + 1])
        out.append(["%s %f" % (int(k), t) for k, t in zip(keys[lo:hi].tolist(), times[lo:hi].tolist())])
    return out


class AsyncKeySink:
    """Streams the AER result of every step -- the reference's 16-bit keys (``spike_to_key``, ``generate_spikes.pyx:744-778``)
    and the CSR row pointers of all streams -- to pinned host memory WITHOUT stalling the device pipeline.

    ``push(res)`` right after ``dvs.step(...)`` queues, on the current stream, the key pass (event count read on the
    device, ``dvs_keys_from_events_devn``) and two asynchronous copies into one of ``depth`` host slots; ``pop()`` hands
    back the oldest slot as ``(offsets, keys)`` NumPy views once its copies have landed -- typically ``depth - 1`` steps
    later, so the host formats / sends frame n while the device works on frame n + 1 (the reference's three-process
    pipeline, ``external_dvs_emulator_device.py:549-683``, on streams and events).  The key copy is sized from the
    previous steps (twice their largest count); a step that emits more is completed by a second copy in ``pop()``.
    The views stay valid until ``depth`` further pushes.
    """

    def __init__(self, dvs, flag_shift, data_shift, data_mask, depth=2, initial_keys=1 << 16):
        import torch
        self.dvs, self.depth = dvs, max(1, int(depth))
        self.params = (int(flag_shift), int(data_shift), int(data_mask))
        self.cap = int(dvs.event_capacity)
        dev = dvs.device
        n_off = dvs.seg_offsets.numel()
        self._keys_dev = [torch.empty(self.cap, dtype=torch.int16, device=dev) for _ in range(self.depth)]
        self._keys_host = [torch.empty(self.cap, dtype=torch.int16).pin_memory() for _ in range(self.depth)]
        self._off_host = [torch.empty(n_off, dtype=torch.int64).pin_memory() for _ in range(self.depth)]
        self._done = [None] * self.depth
        self._copied = [0] * self.depth
        self._guess = max(1, min(self.cap, int(initial_keys)))
        self._pushed = self._popped = 0

    def push(self, res):
        import ctypes as C
        import torch
        from ._lib import check, lib
        from .streams import current_stream_ptr
        if self._pushed - self._popped >= self.depth:
            raise RuntimeError("every host slot holds an unread step: pop() first")
        b = self._pushed % self.depth
        self._pushed += 1
        fs, ds, dm = self.params
        n_dev = res.seg_offsets[-1:]                       # the step's event count, on the device
        check(lib.dvs_keys_from_events_devn(C.c_void_p(res.events_raw.data_ptr()), C.c_void_p(n_dev.data_ptr()), self.cap,
                                            int(self.dvs.cfg.uncoded), fs, ds, dm, C.c_void_p(self._keys_dev[b].data_ptr()),
                                            current_stream_ptr()))
        self._off_host[b].copy_(res.seg_offsets, non_blocking=True)
        m = self._copied[b] = min(self.cap, self._guess)
        self._keys_host[b][:m].copy_(self._keys_dev[b][:m], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._done[b] = ev

    def pop(self):
        """(offsets int64 [S * n_slots * 2 + 1], keys int16 [n]) of the oldest pushed step."""
        if self._popped >= self._pushed:
            raise RuntimeError("nothing pushed")
        b = self._popped % self.depth
        self._popped += 1
        self._done[b].synchronize()
        off = self._off_host[b].numpy()
        n = int(off[-1])
        if n > self.cap:
            raise OverflowError("event buffer too small: call grow_events(capacity)")
        if n > self._copied[b]:   # more than the sized copy brought over: fetch the rest now (and size the next ones up)
            self._keys_host[b][self._copied[b]:n].copy_(self._keys_dev[b][self._copied[b]:n])
        self._guess = max(self._guess, min(self.cap, 2 * n))
        return off, self._keys_host[b].numpy()[:n]
