"""CPU: the C oracle (oracle/dvs_oracle.c) against the golden vectors generated from the compiled,
unmodified reference (tests/golden/make_golden.py).  This is what pins the oracle's parity."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def lists_of(name, t, triples=False):
    lens = G[name + "_list_vals_len"]
    start = int(lens[:t].sum())
    vals = G[name + "_list_vals"][start:start + int(lens[t])]
    off = G[name + "_list_off"][t]
    return [vals[int(off[j]):int(off[j + 1])].tolist() for j in range(len(off) - 1)]


def check_pipeline(name, pipe, adaptive=False):
    frames = G[name + "_frames"]
    for t in range(frames.shape[0]):
        got = pipe.step(frames[t])
        assert np.array_equal(pipe.ref, G[name + "_ref"][t]), (name, t, "ref")
        if adaptive:
            assert np.array_equal(pipe.thr, G[name + "_thr"][t]), (name, t, "thr")
        assert got == lists_of(name, t), (name, t, "lists")


def test_cfg1_rate_bar():
    check_pipeline("cfg1", orc.Pipeline(128, 128, mode=orc.ENC_RATE, threshold=12, max_time_ms=8, flag_shift=14,
                                        data_shift=7, data_mask=127))


@pytest.mark.parametrize("i", range(4))
def test_cfg2_mnist_time_bin_thr_adaptive(i):
    pipe = orc.Pipeline(32, 32, mode=orc.ENC_TIME_BIN_THR, threshold=12, max_time_ms=8, num_bins=6, history_weight=0.99,
                        adaptive=True, min_thr=6, max_thr=168, down=2, up=12, lut=G["cfg2_lut"], flag_shift=10,
                        data_shift=5, data_mask=31, ref0=0, thr0=12)
    check_pipeline("cfg2_%d" % i, pipe, adaptive=True)


def test_cfg3_rate_inhibition():
    check_pipeline("cfg3", orc.Pipeline(64, 48, mode=orc.ENC_RATE, threshold=12, max_time_ms=8, inh_k=2, flag_shift=12,
                                        data_shift=6, data_mask=63))


def test_cfg4_time_adaptive():
    pipe = orc.Pipeline(64, 40, mode=orc.ENC_TIME, threshold=12, max_time_ms=4, num_bins=4, adaptive=True, min_thr=6,
                        max_thr=168, down=2, up=12, flag_shift=12, data_shift=6, data_mask=63, thr0=12)
    check_pipeline("cfg4", pipe, adaptive=True)


def test_uncoded_twin():
    check_pipeline("unc_rate", orc.Pipeline(32, 24, mode=orc.ENC_RATE, threshold=12, max_time_ms=8, history_weight=0.9,
                                            uncoded=True, flag_shift=10, data_shift=5, data_mask=31))
    check_pipeline("unc_time", orc.Pipeline(32, 24, mode=orc.ENC_TIME, threshold=12, max_time_ms=8, num_bins=6,
                                            uncoded=True, flag_shift=10, data_shift=5, data_mask=31))
    check_pipeline("unc_bin", orc.Pipeline(32, 24, mode=orc.ENC_TIME_BIN, threshold=12, max_time_ms=8, num_bins=8,
                                           uncoded=True, lut=G["unc_lut8"], flag_shift=10, data_shift=5, data_mask=31))
    pipe = orc.Pipeline(32, 24, mode=orc.ENC_RATE, threshold=12, max_time_ms=8, adaptive=True, min_thr=6, max_thr=30,
                        down=2, up=12, uncoded=True, flag_shift=10, data_shift=5, data_mask=31, thr0=7)
    check_pipeline("unc_adpt", pipe, adaptive=True)


def test_tables_and_coords():
    assert np.array_equal(orc.generate_log2_table(4, 8), G["log2_4_8"])
    assert np.array_equal(orc.generate_log2_table(2, 6), G["log2_2_6"])
    assert np.array_equal(orc.Oracle.generate_inh_coords(8, 6, 2), G["inh_coords_8_6_2"])
    # the two docstring examples of the reference (generate_spikes.pyx:368-372, :412-417)
    assert orc.generate_log2_table(4, 8)[2][38] == 36
    assert int(orc.generate_log2_table(4, 8)[2][125 // 12]) * 12 == 120


def test_key_wraparound_probes():
    O = orc.Oracle()
    for r, c, k_pos, k_neg in G["key_probes_ds12"]:
        assert O.spike_to_key(int(r), int(c), 24, 12, 255, 1) == k_pos
        assert O.spike_to_key(int(r), int(c), 24, 12, 255, 0) == k_neg


def test_noise_variants_seeded():
    O = orc.Oracle()
    a, s, r = G["noise_a"], G["noise_s"], G["noise_r"]
    np.random.seed(4242)
    sample = np.random.randint(0, 16 * 24, size=int((10 / 100.) * (24 * 16))).astype(np.int16)
    assert np.array_equal(O.update_reference_noise(0, a, s, r, 24, 16, 12, 8, 0.9, G["unc_lut8"], sample), G["noise_raw"])
    assert np.array_equal(O.update_reference_noise(1, a, s, r, 24, 16, 12, 8, 0.9, None, sample), G["noise_thr"])


def test_split_spikes_all_polarities():
    O = orc.Oracle()
    for pol in range(4):
        neg, pos, g = O.split_spikes(G["noise_s"], G["noise_a"], pol)
        assert np.array_equal(neg, G["split_p%d_neg" % pol]) and np.array_equal(pos, G["split_p%d_pos" % pol])
        assert g == int(G["split_p%d_g" % pol])


def test_float_dvs_operator_golden():
    """golden_float_v1.npz was produced by the reference's own DVSOperator (tests/golden/make_golden_float.py)."""
    import os
    from oracle import oracle as orc
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_float_v1.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 3
    for name in names:
        frames, (thr0, relax, up, down), snaps = g[name + "/frames"], g[name + "/params"], g[name + "/snaps"]
        n, H, W = frames.shape
        ref, thr, diff = np.zeros((H, W), np.float32), np.full((H, W), thr0, np.float32), np.zeros((H, W), np.float32)
        for t in range(n):
            orc.dvs_op_f32(np.ascontiguousarray(frames[t]), diff, ref, thr, float(relax), float(up), float(down))
            if (t + 1) % 10 == 0:
                exp = snaps[(t + 1) // 10 - 1]
                for got, e, what in zip((diff, ref, thr), exp, ("diff", "ref", "thr")):
                    assert np.array_equal(got.view(np.uint32), e.view(np.uint32)), (name, t, what)
