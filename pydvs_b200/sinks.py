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
