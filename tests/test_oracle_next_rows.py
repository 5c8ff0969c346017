"""CPU: the oracle's restatement of the rows next to the hot path --
``decode_spikes.pyx`` and ``numba_nvs.py`` -- against golden vectors produced by running the reference
(tests/golden/make_golden_next.py) and, where the reference is present, against the reference live."""
import importlib.util
import os
import warnings

import numpy as np
import pytest

from oracle import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_next_v1.npz"))
ERR = {"": None, "ValueError": ValueError, "OverflowError": OverflowError, "IndexError": IndexError}


def feq(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


# ---- decoder ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("i", range(int(G["dec_rate_n"][0])))
def test_decode_rate_golden(i):
    thr, fs, ds, mask = (int(v) for v in G["dec_rate%d_par" % i])
    out = orc.update_ref_spike_list_rate(G["dec_rate%d_ref" % i].copy(), G["dec_rate%d_keys" % i], thr,
                                         float(G["dec_rate%d_hw" % i][0]), fs, ds, mask)
    assert np.array_equal(out, G["dec_rate%d_out" % i])


def replay_time_case(i, fn_time, fn_time_bin, to_plane=lambda a: a, to_host=lambda a: a):
    """Replays golden decoder sequence i through the given implementation; shared with the GPU tests."""
    binary, nb, per, thr, fs, ds, mask = (int(v) for v in G["dec_time%d_par" % i])
    bw, hw = (float(v) for v in G["dec_time%d_fpar" % i])
    fn = fn_time_bin if binary else fn_time
    ref = to_plane(G["dec_time%d_ref0" % i].copy())
    tl = to_plane(np.zeros(ref.shape, np.int16))
    vl = to_plane(np.full(ref.shape, -1, np.int16))
    lens, keys = G["dec_time%d_keys_len" % i], G["dec_time%d_keys" % i]
    want_err = ERR[str(G["dec_time%d_err" % i][0])]
    n_ok = len(G["dec_time%d_refs" % i])
    pos = 0
    for step, t in enumerate(G["dec_time%d_times" % i]):
        k = to_plane(keys[pos:pos + int(lens[step])].copy())
        pos += int(lens[step])
        if step == n_ok:
            assert want_err is not None
            with pytest.raises(want_err):
                fn(ref, tl, vl, k, nb, bw, per, int(t), thr, hw, fs, ds, mask)
            return
        ref, tl, vl = fn(ref, tl, vl, k, nb, bw, per, int(t), thr, hw, fs, ds, mask)
        assert np.array_equal(to_host(ref), G["dec_time%d_refs" % i][step]), (i, step, "ref")
        assert np.array_equal(to_host(tl), G["dec_time%d_tls" % i][step]), (i, step, "time_last")
        assert np.array_equal(to_host(vl), G["dec_time%d_vls" % i][step]), (i, step, "val_last")
    assert want_err is None


@pytest.mark.parametrize("i", range(int(G["dec_time_n"][0])))
def test_decode_time_golden(i):
    replay_time_case(i, orc.update_ref_spike_list_time, orc.update_ref_spike_list_time_bin)


def test_decode_negative_key_is_overflow_error():
    ref = np.zeros((8, 8), np.int16)
    with pytest.raises(OverflowError):
        orc.update_ref_spike_list_rate(ref, np.array([3, -5], np.int16), 12, 1.0, 6, 3, 7)


def test_decoder_live_against_reference(reference_modules):
    if reference_modules is None:
        pytest.skip("oracle/_ref not built")
    dec = reference_modules[2]
    rng = np.random.default_rng(5)

    def keys_for(n, H, W, fs, ds):
        return ((rng.integers(0, 2, n) << fs) | (rng.integers(0, H, n) << ds) | rng.integers(0, W, n)).astype(np.int16)

    for _ in range(100):
        ref = rng.integers(-50, 300, (16, 16)).astype(np.int16)
        k = keys_for(int(rng.integers(0, 80)), 16, 16, 8, 4)
        thr, hw = int(rng.integers(1, 40)), float(rng.choice([1.0, 0.99, 0.5, 0.29]))
        assert np.array_equal(dec.update_ref_spike_list_rate(ref.copy(), k, thr, hw, 8, 4, 15),
                              orc.update_ref_spike_list_rate(ref.copy(), k, thr, hw, 8, 4, 15))
    for binary, f_ref, f_orc in ((0, dec.update_ref_spike_list_time, orc.update_ref_spike_list_time),
                                 (1, dec.update_ref_spike_list_time_bin, orc.update_ref_spike_list_time_bin)):
        for _ in range(100):
            a = [rng.integers(0, 256, (8, 8)).astype(np.int16), np.zeros((8, 8), np.int16), np.full((8, 8), -1, np.int16)]
            b = [x.copy() for x in a]
            nb, bw, per = int(rng.integers(2, 9)), float(rng.choice([1.0, 0.5, 2.0, 1.5])), int(rng.integers(1, 20))
            thr, hw, t = int(rng.integers(1, 20)), float(rng.choice([1.0, 0.99])), int(rng.integers(0, 70000))
            for _step in range(5):
                k = keys_for(int(rng.integers(0, 30)), 8, 8, 6, 3)
                t += int(rng.integers(1, 40))
                ea = eb = None
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    try:
                        a = list(f_ref(a[0], a[1], a[2], k, nb, bw, per, t, thr, hw, 6, 3, 7))
                    except Exception as e:  # noqa: BLE001
                        ea = type(e)
                try:
                    b = list(f_orc(b[0], b[1], b[2], k, nb, bw, per, t, thr, hw, 6, 3, 7))
                except Exception as e:  # noqa: BLE001
                    eb = type(e)
                assert ea is eb
                if ea is not None:
                    break
                assert all(np.array_equal(x, y) for x, y in zip(a, b))


# ---- numba float variants ------------------------------------------------------------------------
NVS_CASES = [("nvs_0", orc.NVS_0, None), ("nvs_leaky", orc.NVS_LEAKY, None),
             ("nvs_leaky_adaptive", orc.NVS_LEAKY_ADAPTIVE, None),
             ("nvs_noisy_leaky_adaptive_all", orc.NVS_NOISY_LEAKY_ADAPTIVE, 1),
             ("nvs_noisy_leaky_adaptive_none", orc.NVS_NOISY_LEAKY_ADAPTIVE, 0)]


def nvs_kwargs():
    leak, base, incr, tleak, target = (float(v) for v in G["nvs_args"])
    return dict(reference_leak=leak, threshold_base=base, threshold_mult_incr=incr, threshold_leak=tleak), target


@pytest.mark.parametrize("tag,mode,mask", NVS_CASES)
def test_nvs_golden(tag, mode, mask):
    kw, _ = nvs_kwargs()
    r, t = G["nvs_ref0"].copy(), G["nvs_thr0"].copy()
    s = np.zeros_like(r)
    for f in G["nvs_frames"]:
        m = None if mask is None else np.full(r.shape, mask, np.uint8)
        active = orc.nvs_step(mode, f.copy(), r, t, s, leak_mask=m, **kw)
    assert feq(r, G[tag + "_ref"]) and feq(t, G[tag + "_thr"]) and feq(s, G[tag + "_spikes"])
    assert active.dtype == np.uint8


@pytest.mark.parametrize("tag,mask", [("nvs_bi_noisy_leaky_adaptive_all", 1), ("nvs_bi_noisy_leaky_adaptive_none", 0)])
def test_nvs_two_cell_golden(tag, mask):
    kw, target = nvs_kwargs()
    r, t = G["nvs_ref0"].copy(), G["nvs_thr0"].copy()
    ro, to = r[::-1].copy(), t[::-1].copy()
    s = np.zeros_like(r)
    m = np.full(r.shape, mask, np.uint8)
    for f in G["nvs_frames"]:
        orc.nvs_bi_step(f.copy(), r, t, s, kw["reference_leak"], kw["threshold_base"], kw["threshold_mult_incr"],
                        kw["threshold_leak"], ro, to, target, leak_mask=m, leak_mask_off=m)
    for got, name in ((r, "_ref"), (t, "_thr"), (s, "_spikes"), (ro, "_ref_off"), (to, "_thr_off")):
        assert feq(got, G[tag + name]), name


def test_nvs_thresholded_difference_golden():
    r, t = G["nvs_ref0"].copy(), G["nvs_thr0"].copy()
    active = orc.nvs_step(orc.NVS_0, G["nvs_frames"][0].copy(), r, t, np.zeros_like(r))
    assert np.array_equal(active, G["nvs_td_active"])


def test_nvs_live_against_numba():
    path = "/root/reference/pydvs/numba_nvs.py"
    if not os.path.exists(path):
        pytest.skip("reference sources not present")
    pytest.importorskip("numba")
    if not hasattr(np, "float"):
        np.float = float
    spec = importlib.util.spec_from_file_location("_numba_nvs_ref", path)
    nvs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nvs)
    rng = np.random.default_rng(3)
    kw = dict(reference_leak=0.9, threshold_base=7.5, threshold_mult_incr=0.77, threshold_leak=0.81)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(4):
            f = rng.uniform(0, 255, (32, 48)).astype(np.float32)
            a = [rng.uniform(0, 255, f.shape).astype(np.float32), rng.uniform(-3, 40, f.shape).astype(np.float32),
                 np.zeros_like(f)]
            b = [x.copy() for x in a]
            for _it in range(3):
                nvs.nvs_leaky_adaptive(f, a[0], a[1], a[2], 0.9, 7.5, 0.77, 0.81)
                orc.nvs_step(orc.NVS_LEAKY_ADAPTIVE, f, b[0], b[1], b[2], **kw)
            assert all(feq(x, y) for x, y in zip(a, b))


# ---- frame animators -------------------------------------------------------------------------------
def test_traverse_and_fade_golden():
    from oracle import frame_oracle as fo
    img = G["anim_img"]
    for (fn, sp), want in zip(G["anim_trav_par"], G["anim_trav_out"]):
        assert np.array_equal(fo.traverse_image(img, int(fn), float(sp), 37), want), (fn, sp)
    for (fn, hf), want in zip(G["anim_fade_par"], G["anim_fade_out"]):
        got = fo.fade_image(img, int(fn), int(hf), 37)
        assert got.dtype == np.int16 and np.array_equal(got, want), (fn, hf)
