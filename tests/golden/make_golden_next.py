#!/usr/bin/env python
"""Generate tests/golden/golden_next_v1.npz: golden vectors for the rows next to the hot path
 by RUNNING THE REFERENCE in this container:

  * ``pydvs/decode_spikes.pyx`` compiled unmodified into oracle/_ref (oracle/build_ref.py);
  * ``pydvs/numba_nvs.py`` imported from /root/reference under numba (only ``numpy.float`` shimmed).

The noisy numba variants use an unseeded private generator; they are run with leak probability 1.0
(every pixel leaks) and -1.0 (none does), which makes the draw deterministic and pins both branches.

    python tests/golden/make_golden_next.py      # needs /root/reference; the .npz is committed
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

NVS_ARGS = (0.97, 10.0, 0.31, 0.93)  # reference_leak, threshold_base, threshold_mult_incr, threshold_leak
TARGET_OFF = 120.0


def make_keys(rng, H, W, n, flag_shift, data_shift, unique):
    if unique:
        q = rng.choice(H * W, min(n, H * W), replace=False)
        r, c = q // W, q % W
    else:
        r, c = rng.integers(0, H, n), rng.integers(0, W, n)
    f = rng.integers(0, 2, len(r))
    return ((f << flag_shift) | (r << data_shift) | c).astype(np.int16)


def nvs_planes(seed, H, W):
    r = np.random.default_rng(seed)
    frame = r.uniform(0, 255, (H, W)).astype(np.float32)
    ref = r.uniform(0, 255, (H, W)).astype(np.float32)
    thr = r.uniform(2, 40, (H, W)).astype(np.float32)
    frame[0, :8] = ref[0, :8] + 3 * thr[0, :8]      # exact multiples of the threshold
    frame[1, :8] = ref[1, :8] - 3 * thr[1, :8]
    thr[2, :4] = 0.0                                # zero / negative thresholds (inf / nan propagate)
    thr[3, :4] = -5.0
    frame[4, :8] = ref[4, :8]                       # zero difference
    return frame, ref, thr


def main():
    mods = ref_loader.load()
    if mods is None:
        raise SystemExit("oracle/_ref is not built: python oracle/build_ref.py")
    dec = mods[2]
    if not hasattr(np, "float"):
        np.float = float
    spec = importlib.util.spec_from_file_location("_numba_nvs_ref", "/root/reference/pydvs/numba_nvs.py")
    nvs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nvs)
    warnings.simplefilter("ignore")
    out = {}
    rng = np.random.default_rng(2024)

    # ---- decoder: rate -------------------------------------------------------------------------
    n_rate = 6
    for i in range(n_rate):
        H, W, ds, mask, fs = [(16, 16, 4, 15, 8), (32, 32, 5, 31, 10), (128, 128, 7, 127, 14)][i % 3]
        ref = rng.integers(-40, 300, (H, W)).astype(np.int16)
        keys = make_keys(rng, H, W, int(rng.integers(1, 200)), fs, ds, unique=bool(i % 2))
        thr, hw = int(rng.integers(1, 40)), [1.0, 0.99, 0.29][i % 3]
        out["dec_rate%d_ref" % i] = ref
        out["dec_rate%d_keys" % i] = keys
        out["dec_rate%d_par" % i] = np.array([thr, fs, ds, mask], np.int64)
        out["dec_rate%d_hw" % i] = np.array([hw])
        out["dec_rate%d_out" % i] = dec.update_ref_spike_list_rate(ref.copy(), keys, thr, hw, fs, ds, mask)
    out["dec_rate_n"] = np.array([n_rate])

    # ---- decoder: time / time_bin, a few frames each, until the reference raises ----------------
    n_time = 8
    for i in range(n_time):
        binary = i % 2
        fn = dec.update_ref_spike_list_time_bin if binary else dec.update_ref_spike_list_time
        H, W, ds, mask, fs = 8, 8, 3, 7, 6
        ref = rng.integers(0, 256, (H, W)).astype(np.int16)
        tl, vl = np.zeros((H, W), np.int16), np.full((H, W), -1, np.int16)
        nb, bw, per = int(rng.integers(2, 9)), [1.0, 0.5, 2.0, 1.5][i % 4], int(rng.integers(2, 20))
        thr, hw = int(rng.integers(1, 20)), [1.0, 0.99][(i // 2) % 2]
        t = int(rng.integers(0, 70000))
        out["dec_time%d_ref0" % i] = ref.copy()
        out["dec_time%d_par" % i] = np.array([binary, nb, per, thr, fs, ds, mask], np.int64)
        out["dec_time%d_fpar" % i] = np.array([bw, hw])
        keys_all, times, refs, tls, vls, err = [], [], [], [], [], ""
        for step in range(5):
            keys = make_keys(rng, H, W, int(rng.integers(1, 30)), fs, ds, unique=(i % 4 < 2))
            if binary:  # keep the signs positive so that log2 stays defined for a few frames
                keys = (keys & ~(1 << fs)).astype(np.int16)
            t += int(rng.integers(1, 40))
            keys_all.append(keys)
            times.append(t)
            try:
                ref, tl, vl = fn(ref, tl, vl, keys, nb, bw, per, t, thr, hw, fs, ds, mask)
            except Exception as e:  # noqa: BLE001
                err = type(e).__name__
                break
            refs.append(ref.copy()); tls.append(tl.copy()); vls.append(vl.copy())
        out["dec_time%d_keys" % i] = np.concatenate(keys_all)
        out["dec_time%d_keys_len" % i] = np.array([len(k) for k in keys_all], np.int64)
        out["dec_time%d_times" % i] = np.array(times, np.int64)
        out["dec_time%d_refs" % i] = np.array(refs, np.int16).reshape(-1, H, W)
        out["dec_time%d_tls" % i] = np.array(tls, np.int16).reshape(-1, H, W)
        out["dec_time%d_vls" % i] = np.array(vls, np.int16).reshape(-1, H, W)
        out["dec_time%d_err" % i] = np.array([err])
    out["dec_time_n"] = np.array([n_time])

    # ---- numba float variants --------------------------------------------------------------------
    H, W = 24, 40
    frame, ref0, thr0 = nvs_planes(11, H, W)
    frames = np.stack([np.roll(frame, it, 1) for it in range(3)])
    out["nvs_frames"], out["nvs_ref0"], out["nvs_thr0"] = frames, ref0, thr0
    out["nvs_args"] = np.array(NVS_ARGS + (TARGET_OFF,))
    cases = [("nvs_0", None), ("nvs_leaky", None), ("nvs_leaky_adaptive", None),
             ("nvs_noisy_leaky_adaptive", 1.0), ("nvs_noisy_leaky_adaptive", -1.0),
             ("nvs_bi_noisy_leaky_adaptive", 1.0), ("nvs_bi_noisy_leaky_adaptive", -1.0)]
    for name, p in cases:
        r, t, s = ref0.copy(), thr0.copy(), np.zeros_like(ref0)
        ro, to = ref0[::-1].copy(), thr0[::-1].copy()
        for f in frames:
            f = f.copy()
            if name == "nvs_0":
                nvs.nvs_0(f, r, t, s)
            elif name == "nvs_leaky":
                nvs.nvs_leaky(f, r, t, s, NVS_ARGS[0])
            elif name == "nvs_leaky_adaptive":
                nvs.nvs_leaky_adaptive(f, r, t, s, *NVS_ARGS)
            elif name == "nvs_noisy_leaky_adaptive":
                nvs.nvs_noisy_leaky_adaptive(f, r, t, s, *NVS_ARGS, p)
            else:
                nvs.nvs_bi_noisy_leaky_adaptive(f, r, t, s, *NVS_ARGS, p, ro, to, TARGET_OFF)
        tag = name + ("" if p is None else ("_all" if p >= 1 else "_none"))
        out[tag + "_ref"], out[tag + "_thr"], out[tag + "_spikes"] = r, t, s
        if name.startswith("nvs_bi"):
            out[tag + "_ref_off"], out[tag + "_thr_off"] = ro, to
    a = nvs.thresholded_difference(frames[0].copy(), ref0.copy(), thr0.copy(), np.zeros_like(ref0))
    out["nvs_td_active"] = np.asarray(a, np.uint8)

    # ---- frame animators (generate_spikes.pyx:1128-1182) ------------------------------------------
    gs = mods[0]
    rng = np.random.default_rng(77)
    img = rng.integers(0, 256, (12, 20)).astype(np.int16)
    out["anim_img"] = img
    trav = [(fn, sp) for sp in (0.5, 1.0, 1.7, 3.3, -1.0) for fn in (0, 1, 5, 12, 13, 20, 24, 39, 41, 70)]
    out["anim_trav_par"] = np.array(trav, np.float64)
    out["anim_trav_out"] = np.stack([gs.traverse_image(img, int(fn), float(sp), 37) for fn, sp in trav])
    fade = [(fn, hf) for hf in (1, 7, 15) for fn in (0, 1, 5, 6, 7, 8, 9, 14, 16, 29, 31)]
    out["anim_fade_par"] = np.array(fade, np.int64)
    out["anim_fade_out"] = np.stack([gs.fade_image(img, int(fn), int(hf), 37) for fn, hf in fade])

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_next_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %d bytes" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
