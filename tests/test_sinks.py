"""CPU: the text spike-file sink against a literal replay of mnist_to_spikes.py:285-310 on golden lists."""
import os

import numpy as np
import torch

from pydvs_b200 import SpikeLists
from pydvs_b200.sinks import SpikeFileWriter, calculate_time_per_pack, spike_file_lines

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def golden_lists(name, t):
    lens = G[name + "_list_vals_len"]
    start = int(lens[:t].sum())
    vals = G[name + "_list_vals"][start:start + int(lens[t])]
    off = G[name + "_list_off"][t]
    return [vals[int(off[j]):int(off[j + 1])].tolist() for j in range(len(off) - 1)], vals, off


def test_spike_file_matches_the_reference_writer_loop(tmp_path):
    frame_time_ms, t_bin_ms = 8, 1
    write_buff, t = [], 0                     # mnist_to_spikes.py:285-303, verbatim control flow
    writer = SpikeFileWriter(frame_time_ms, t_bin_ms)
    for frame in range(G["cfg1_frames"].shape[0]):
        lists, vals, off = golden_lists("cfg1", frame)
        t_idx = 0
        for spk_list in lists:
            for spk in spk_list:
                write_buff.append("%s %f" % (spk, t + t_idx))
            t_idx += t_bin_ms
        t += frame_time_ms
        if frame % 2:                          # feed the CSR form and the list form alternately
            writer.add_frame(SpikeLists(off.astype(np.int64), torch.from_numpy(vals.astype(np.int16))))
        else:
            writer.add_frame(lists)
    assert writer.getvalue() == "\n".join(write_buff)
    path = tmp_path / "spikes.txt"
    writer.write(str(path))
    assert path.read_text() == "\n".join(write_buff)
    assert spike_file_lines([[], [5], []], 16, 2) == ["5 18.000000"]


def test_time_per_pack_rule():
    assert calculate_time_per_pack("RATE", 16, 6, 168, 4) == 1
    assert calculate_time_per_pack("TIME", 16, 6, 168, 4) == 16 // min(16, 168 // 6 + 1)
    assert calculate_time_per_pack("TIME_BIN", 16, 6, 168, 4) == 2
    assert calculate_time_per_pack("TIME_BIN_THR", 16, 6, 168, 4) == 3


def test_emit_spikes_is_paced_per_slot():
    from pydvs_b200.sinks import emit_spikes

    class Sender:
        def __init__(self):
            self.calls = []

        def send_spikes(self, label, pack, send_full_keys=True):
            self.calls.append((label, list(pack), send_full_keys))
            now[0] += cost.pop(0)

    now, slept = [100.0], []
    cost = [0.0002, 0.0030, 0.0005]          # the second pack takes longer than its 2 ms budget

    def sleep(dt):
        slept.append(round(dt, 6))
        now[0] += dt

    s = Sender()
    n = emit_spikes(s, [[1, 2], [], [7]], 2, "retina", clock=lambda: now[0], sleep=sleep)
    assert n == 3 and s.calls == [("retina", [1, 2], False), ("retina", [], False), ("retina", [7], False)]
    assert slept == [0.0018, 0.0015]
    assert emit_spikes(s, None, 2, "retina", clock=lambda: now[0], sleep=sleep) == 0
