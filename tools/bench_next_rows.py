#!/usr/bin/env python
"""Device-time of the kernels next to the hot path against their HBM roofline.
Prints one JSON line per kernel: algorithmic bytes / CUDA-event time."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydvs_b200.numba_nvs as nvs  # noqa: E402
import pydvs_b200.decode_spikes as dec  # noqa: E402

PEAK = 6552.3
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        PEAK = float(json.load(f).get("hbm_gbs", PEAK))
except (OSError, ValueError):
    pass


def timed(fn, reps=20, warm=3):
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


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                      "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3)}))


def main():
    shape = (64, 1080, 1920)  # cfg 4 geometry: 132.7 Mpx
    n = shape[0] * shape[1] * shape[2]
    g = torch.Generator(device="cuda").manual_seed(1)
    f = torch.rand(shape, device="cuda", generator=g) * 255
    r = torch.rand(shape, device="cuda", generator=g) * 255
    t = torch.rand(shape, device="cuda", generator=g) * 30 + 5
    s = torch.zeros(shape, device="cuda")
    m = (torch.rand(shape, device="cuda", generator=g) < 0.5).to(torch.uint8)
    a = (0.97, 10.0, 0.31, 0.93)
    # bytes per pixel: frame 4 R, ref 4 R + 4 W, spikes 4 W (+ thr 4 R, + 4 W when adaptive, + 1 R mask)
    report("nvs_0 (k_nvs<0>)", 20 * n, timed(lambda: nvs.nvs_0(f, r, t, s)))
    report("nvs_leaky (k_nvs<1>)", 20 * n, timed(lambda: nvs.nvs_leaky(f, r, t, s, a[0])))
    report("nvs_leaky_adaptive (k_nvs<2>)", 24 * n, timed(lambda: nvs.nvs_leaky_adaptive(f, r, t, s, *a)))
    report("nvs_noisy_leaky_adaptive (k_nvs<3>, mask given)", 25 * n,
           timed(lambda: nvs.nvs_noisy_leaky_adaptive(f, r, t, s, *a, 0.5, leak_mask=m)))
    ro, to = r.clone(), t.clone()
    report("nvs_bi_noisy_leaky_adaptive (k_nvs<4>, masks given)", 42 * n,
           timed(lambda: nvs.nvs_bi_noisy_leaky_adaptive(f, r, t, s, *a, 0.5, ro, to, 120.0, leak_mask=m, leak_mask_off=m)))
    del f, r, t, s, m, ro, to
    # receiver: 256 streams of 128x128, 3 % of the pixels carry a key
    S, H, W = 4096, 128, 128
    ref = torch.full((S, H, W), 128, dtype=torch.int16, device="cuda")
    per = int(0.03 * H * W)
    q = torch.randint(0, H * W, (S, per), device="cuda", generator=g)
    keys = (((q // W) << 7) | (q % W) | (torch.randint(0, 2, (S, per), device="cuda", generator=g) << 14)).to(torch.int16).reshape(-1)
    offs = (torch.arange(S + 1, device="cuda", dtype=torch.int64) * per)
    nk = keys.numel()
    ms = timed(lambda: dec.decode_rate_batch(ref, keys, offs, 12, 0.99, 14, 7, 127))
    # decay 2 R + 2 W per pixel; per key 2 B key + 2 B read-modify-write of the pixel
    report("decode_rate_batch (k_decay_i16 + k_decode_rate, incl. the error-word readback)", 4 * S * H * W + 6 * nk, ms)


if __name__ == "__main__":
    main()
