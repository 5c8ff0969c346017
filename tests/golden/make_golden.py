#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference.

Runs in the build container only: it imports the compiled pyDVS Cython modules from ``oracle/_ref``
(``oracle/build_ref.py`` builds them from /root/reference) and reads the reference's MNIST / fading-mask
fixtures.  Everything the tests need -- input frames included -- is stored in the .npz, so neither the
CPU tests nor the GPU tests touch /root/reference at run time.

    python tests/golden/make_golden.py

Cases (SURVEY.md section 8(d) shapes, reduced to fixture size):
  cfg1   128x128 moving bar, rate coded, thr 12, max_time_ms 8, hw 1.0, MERGED        (stand_alone.py loop)
  cfg2   4 MNIST digits padded 28->32, injected micro-saccades x fade mask, adaptive thresholds,
         time_bin_thr (6 bins), hw 0.99                                                  (mnist_to_spikes.py loop)
  cfg3   48x64 random frames, rate + local inhibition k=2
  cfg4   40x64 bar + noise, time coding (4 bins) + adaptive threshold
  unc    uncoded twin: 24x32 random frames, rate / time triples + time_bin keys
  misc   log2 tables, inh coords, key wrap-around probes, noise variants (seeded global RNG)
Lists of lists are stored flattened as (values, slot_offsets).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PYDVS_REFERENCE", "/root/reference")

warnings.simplefilter("ignore")


def flat(lists):
    off = np.cumsum([0] + [len(x) for x in lists]).astype(np.int64)
    vals = [v for sl in lists for v in sl]
    if vals and isinstance(vals[0], (list, tuple)):
        arr = np.array([[int(c) for c in v] for v in vals], dtype=np.int32).reshape(-1, 3)
    else:
        arr = np.array([int(v) for v in vals], dtype=np.int32)
    return arr, off


def bar_frames(n, H, W, rng, noise):
    frames = np.full((n, H, W), 64, dtype=np.int16)
    bw, sp = max(1, W // 16), max(1, W // 64)
    for t in range(n):
        x0 = (t * sp) % W
        cols = (np.arange(bw) + x0) % W
        frames[t][:, cols] = 200
        if noise:
            frames[t] += rng.integers(-noise, noise + 1, (H, W)).astype(np.int16)
    return np.clip(frames, 0, 255).astype(np.int16)


def main():
    from oracle import ref_loader
    mods = ref_loader.load()
    assert mods is not None, "build the reference first: python oracle/build_ref.py"
    gs, gsu, _ = mods
    import cv2
    out = {}
    rng = np.random.default_rng(20261019)

    def run_pipeline(name, frames, *, mod=gs, adaptive=False, inh=0, update, encode, ref0=128, thr0=12, pol=2):
        """frames [n, H, W]; stores per-frame spikes / ref / thr and the flattened lists."""
        n, H, W = frames.shape
        ref = np.full((H, W), ref0, dtype=np.int16)
        thr = np.full((H, W), thr0, dtype=np.int16)
        coords = mod.generate_inh_coords(W, H, inh) if inh > 1 else None
        sp_all, ref_all, thr_all, vals_all, off_all = [], [], [], [], []
        for t in range(n):
            if adaptive:
                d, a, s = mod.thresholded_difference_adpt(frames[t], ref, thr, 6, 168)
            else:
                d, a, s = mod.thresholded_difference(frames[t], ref, 12)
            if inh > 1:
                s = mod.local_inhibition(s, a, coords, W, H, inh)
            res = update(a, s, ref, thr)
            if adaptive:
                ref[:], thr[:] = res
            else:
                ref[:] = res
            neg, pos, g = mod.split_spikes(s, a, pol)
            v, o = flat(encode(pos, neg, g))
            sp_all.append(s.copy()); ref_all.append(ref.copy()); thr_all.append(thr.copy())
            vals_all.append(v); off_all.append(o)
        out[name + "_frames"] = frames
        out[name + "_spikes"] = np.stack(sp_all)
        out[name + "_ref"] = np.stack(ref_all)
        if adaptive:
            out[name + "_thr"] = np.stack(thr_all)
        out[name + "_list_vals"] = np.concatenate(vals_all) if vals_all else np.zeros(0, np.int32)
        out[name + "_list_vals_len"] = np.array([len(v) for v in vals_all], dtype=np.int64)
        out[name + "_list_off"] = np.stack(off_all)

    # ---- cfg1 ---------------------------------------------------------------------------------
    f = bar_frames(8, 128, 128, rng, 0)
    run_pipeline("cfg1", f,
                 update=lambda a, s, r, th: gs.update_reference_rate(a, s, r, 12, 8, 1.0),
                 encode=lambda p, n, g: gs.make_spike_lists_rate(p, n, g, 12, 14, 7, 127, 8))

    # ---- cfg2: MNIST saccades (offsets injected), fade mask, adaptive, time_bin_thr ---------------
    import glob
    pngs = sorted(glob.glob(os.path.join(REF, "mnist", "t10k", "*.png")))[:4]
    mask = cv2.imread(os.path.join(REF, "pydvs", "fading_mask.png"), cv2.IMREAD_GRAYSCALE)
    mask = np.float64(cv2.resize(mask, (32, 32))) / 255.0
    lut6 = gs.generate_log2_table(2, 6)[1]
    for i, png in enumerate(pngs):
        digit = cv2.imread(png, cv2.IMREAD_GRAYSCALE)
        padded = np.zeros((32, 32), dtype=np.int16)
        padded[2:30, 2:30] = digit
        frames, cx, cy = [], 0, 0
        for t in range(10):
            if t:
                cx = int(np.clip(cx + rng.integers(-1, 2), -1, 1))
                cy = int(np.clip(cy + rng.integers(-1, 2), -1, 1))
            moved = np.zeros_like(padded)
            ys, xs = np.mgrid[0:32, 0:32]
            sx, sy = xs - cx, ys + cy
            ok = (sx >= 0) & (sx < 32) & (sy >= 0) & (sy < 32)
            moved[ok] = padded[sy[ok], sx[ok]]
            frames.append(gs.mask_image(moved, mask))
        run_pipeline("cfg2_%d" % i, np.stack(frames), adaptive=True, ref0=0, thr0=12,
                     update=lambda a, s, r, th: gs.update_reference_rate_adpt(a, s, r, th, 6, 168, 2, 12, 8, 0.99),
                     encode=lambda p, n, g: gs.make_spike_lists_time_bin_thr(p, n, g, 10, 5, 31, 8, 6, 168, 6, lut6))
    out["cfg2_lut"] = lut6

    # ---- cfg3: rate + inhibition ------------------------------------------------------------------
    f = rng.integers(0, 256, (5, 48, 64)).astype(np.int16)
    run_pipeline("cfg3", f, inh=2,
                 update=lambda a, s, r, th: gs.update_reference_rate(a, s, r, 12, 8, 1.0),
                 encode=lambda p, n, g: gs.make_spike_lists_rate(p, n, g, 12, 12, 6, 63, 8))

    # ---- cfg4: time coding + adaptive ---------------------------------------------------------------
    f = bar_frames(8, 40, 64, rng, 6)
    run_pipeline("cfg4", f, adaptive=True, thr0=12,
                 update=lambda a, s, r, th: gs.update_reference_rate_adpt(a, s, r, th, 6, 168, 2, 12, 4, 1.0),
                 encode=lambda p, n, g: gs.make_spike_lists_time(p, n, g, 12, 6, 63, 4, 4, 6, 168))

    # ---- uncoded twin -------------------------------------------------------------------------------
    f = rng.integers(0, 256, (4, 24, 32)).astype(np.int16)
    run_pipeline("unc_rate", f, mod=gsu,
                 update=lambda a, s, r, th: gsu.update_reference_rate(a, s, r, 12, 8, 0.9),
                 encode=lambda p, n, g: gsu.make_spike_lists_rate(p, n, g, 12, 10, 5, 31, 8))
    run_pipeline("unc_time", f, mod=gsu,
                 update=lambda a, s, r, th: gsu.update_reference_time_thresh(a, s, r, 12, 8, 1.0),
                 encode=lambda p, n, g: gsu.make_spike_lists_time(p, n, g, 10, 5, 31, 6, 8, 12, 168))
    lut8 = gsu.generate_log2_table(4, 8)[3]
    run_pipeline("unc_bin", f, mod=gsu,
                 update=lambda a, s, r, th: gsu.update_reference_time_binary_raw(a, s, r, 12, 8, 3, 1.0, lut8),
                 encode=lambda p, n, g: gsu.make_spike_lists_time_bin(p.astype(np.int16), n.astype(np.int16), g, 10, 5,
                                                                      31, 8, 12, 168, 8, lut8))
    run_pipeline("unc_adpt", f, mod=gsu, adaptive=True, thr0=7,
                 update=lambda a, s, r, th: gsu.update_reference_rate_adpt(a, s, r, th, 6, 30, 2, 12, 8, 1.0),
                 encode=lambda p, n, g: gsu.make_spike_lists_rate(p, n, g, 6, 10, 5, 31, 8))
    out["unc_lut8"] = lut8

    # ---- misc known answers -------------------------------------------------------------------------
    out["log2_4_8"] = gs.generate_log2_table(4, 8)
    out["log2_2_6"] = gs.generate_log2_table(2, 6)
    out["inh_coords_8_6_2"] = gs.generate_inh_coords(8, 6, 2)
    # key wrap-around beyond 128-class geometries (Appendix A.6): single-pixel lists
    probes = [(0, 0), (127, 127), (2159, 3839), (1079, 1919), (5, 300)]
    keys = []
    for (r, c) in probes:
        pos = np.array([[r], [c], [50]], dtype=np.int16)
        empty = np.zeros((3, 0), dtype=np.int16)
        k_pos = gs.make_spike_lists_rate(pos, empty, 50, 12, 24, 12, 255, 8)[0][0]
        k_neg = gs.make_spike_lists_rate(empty, pos, 50, 12, 24, 12, 255, 8)[0][0]
        keys.append((r, c, k_pos, k_neg))
    out["key_probes_ds12"] = np.array(keys, dtype=np.int64)
    # noise variants with the global NumPy RNG seeded (generate_spikes.pyx:428-473)
    c = rng.integers(0, 256, (16, 24)).astype(np.int16)
    r = rng.integers(0, 256, (16, 24)).astype(np.int16)
    d, a, s = gs.thresholded_difference(c, r, 12)
    out["noise_a"], out["noise_s"], out["noise_r"] = a, s, r
    np.random.seed(4242)
    out["noise_raw"] = gs.update_reference_time_binary_raw_noise(a, s, r, 24, 16, 12, 8, 3, 0.9, lut8, 10)
    np.random.seed(4242)
    out["noise_thr"] = gs.update_reference_time_thresh_noise(a, s, r, 24, 16, 12, 8, 0.9, 10)
    # every polarity of split_spikes on one plane
    for pol in range(4):
        neg, pos, g = gs.split_spikes(s, a, pol)
        out["split_p%d_neg" % pol], out["split_p%d_pos" % pol], out["split_p%d_g" % pol] = neg, pos, np.int64(g)

    # frame animators: usaccade_image on a frame where it only moves (frame % frames_per_usaccade != 0), then mask
    digit = cv2.imread(pngs[0], cv2.IMREAD_GRAYSCALE)
    padded = np.zeros((32, 32), dtype=np.int16)
    padded[2:30, 2:30] = digit
    moves = [(0, 0), (1, 0), (0, 1), (-1, -1), (2, -3), (-5, 4), (31, 0), (0, -32), (40, 40)]
    out["move_src"], out["move_mask"] = padded, mask
    out["move_offsets"] = np.array(moves, dtype=np.int32)
    out["move_plain"] = np.stack([gs.usaccade_image(padded, 3, 2, 1, cx, cy, 9)[0] for cx, cy in moves])
    out["move_masked"] = np.stack([gs.mask_image(gs.usaccade_image(padded, 3, 2, 1, cx, cy, 0)[0], mask) for cx, cy in moves])

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KiB, %d arrays)" % (path, os.path.getsize(path) / 1024, len(out)))


if __name__ == "__main__":
    main()
