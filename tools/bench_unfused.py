#!/usr/bin/env python
"""Per-function device time of the unfused drop-in calls (pydvs_b200.generate_spikes on CUDA tensors) against the
algorithmic bytes of SURVEY 8(d): thresholded_difference 10 B/px, local_inhibition 6 B/px, update_reference_rate
8 B/px, split_spikes + make_spike_lists_rate 4 B/px + 8 B/event.  One JSON line per function."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydvs_b200.generate_spikes as gs  # noqa: E402
from pydvs_b200 import synth  # noqa: E402

PEAK = 6552.3
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f).get("hbm_gbs", PEAK))
except (OSError, ValueError):
    pass


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"function": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                      "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3), "note": note}))


def main():
    S, H, W = 16, 2160, 3840
    n = S * H * W
    frames = synth.moving_bar(S, H, W, 3, noise=8, seed=1)
    ref = synth.moving_bar(S, H, W, 2, noise=8, seed=1)
    report("thresholded_difference", 10 * n, timed(lambda: gs.thresholded_difference(frames, ref, 12)),
           "includes the allocation of the three result planes")
    diff, abs_diff, spikes = gs.thresholded_difference(frames, ref, 12)
    copies = [spikes.clone() for _ in range(12)]   # in place: every timed call gets planes that were not inhibited yet
    it = iter(copies)
    report("local_inhibition (k = 2, in place)", 6 * n, timed(lambda: gs.local_inhibition(next(it), abs_diff, None, W, H, 2)))
    sp = copies[0]
    del copies, it
    report("update_reference_rate", 8 * n, timed(lambda: gs.update_reference_rate(abs_diff, sp, ref, 12, 8, 1.0)),
           "includes the allocation of the result plane")
    # split + lists on one 4K stream (the reference API is per frame)
    s1, a1 = sp[0].contiguous(), abs_diff[0].contiguous()

    def split_and_lists():
        neg, pos, gmax = gs.split_spikes(s1, a1, 2)
        return gs.make_spike_lists_rate(pos, neg, gmax, 12, 15, 12, 255, 8)

    lists = split_and_lists()
    n_ev = sum(len(x) for x in lists)
    report("split_spikes + make_spike_lists_rate (one 4K frame)", 4 * H * W + 8 * n_ev, timed(split_and_lists, reps=5, warm=1),
           "%d events; includes host synchronisation for the list sizes" % n_ev)


if __name__ == "__main__":
    main()
