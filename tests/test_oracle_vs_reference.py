"""CPU: live differential test of the C oracle against the compiled, unmodified reference in
oracle/_ref (skipped where the reference has not been built, e.g. on the GPU box before build())."""
import warnings

import numpy as np
import pytest

from oracle.oracle import Oracle, generate_log2_table


@pytest.fixture(scope="module")
def refs(reference_modules):
    if reference_modules is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    warnings.simplefilter("ignore")
    return reference_modules


def _eq(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), what


def _norm(lists):
    return [[[int(c) for c in e] if isinstance(e, (list, tuple)) else int(e) for e in sl] for sl in lists]


@pytest.mark.parametrize("uncoded", [False, True])
def test_every_hot_path_function(refs, uncoded):
    ref = refs[1] if uncoded else refs[0]
    O = Oracle(uncoded)
    rng = np.random.default_rng(7 + uncoded)
    lut = ref.generate_log2_table(4, 8)
    _eq(lut, generate_log2_table(4, 8), "log2 table")
    lut6 = ref.generate_log2_table(2, 6)[1]
    for trial in range(24):
        H, W = [(8, 8), (16, 12), (32, 32), (6, 10)][trial % 4]
        curr = rng.integers(0, 256, (H, W)).astype(np.int16)
        r0 = rng.integers(0, 256, (H, W)).astype(np.int16)
        T = int(rng.integers(1, 40))
        d, a, s = ref.thresholded_difference(curr, r0, T)
        for x, y in zip((d, a, s), O.thresholded_difference(curr, r0, T)):
            _eq(x, y, "thresholded_difference")
        thr = rng.integers(0, 60, (H, W)).astype(np.int16)
        da, aa, sa = ref.thresholded_difference_adpt(curr, r0, thr, 6, 168)
        for x, y in zip((da, aa, sa), O.thresholded_difference_adpt(curr, r0, thr, 6, 168)):
            _eq(x, y, "thresholded_difference_adpt")
        if H % 2 == 0 and W % 2 == 0:
            coords = ref.generate_inh_coords(W, H, 2)
            _eq(coords, O.generate_inh_coords(W, H, 2), "inh coords")
            s1, s2 = s.copy(), s.copy()
            ref.local_inhibition(s1, a, coords, W, H, 2)
            O.local_inhibition(s2, a, coords, W, H, 2)
            _eq(s1, s2, "local_inhibition")
        hw = [1.0, 0.99, 0.5, 0.29][trial % 4]
        mt = int(rng.integers(1, 20))
        _eq(ref.update_reference_rate(a, s, r0, T, mt, hw), O.update_reference_rate(a, s, r0, T, mt, hw), "rate")
        _eq(ref.update_reference_time_thresh(a, s, r0, T, mt, hw), O.update_reference_time_thresh(a, s, r0, T, mt, hw),
            "time_thresh")
        t1, t2 = thr.copy(), thr.copy()
        ra, ta = ref.update_reference_rate_adpt(aa, sa, r0, t1, 6, 168, 2, 12, mt, hw)
        rb, tb = O.update_reference_rate_adpt(aa, sa, r0, t2, 6, 168, 2, 12, mt, hw)
        _eq(ra, rb, "rate_adpt ref")
        _eq(np.asarray(ta), np.asarray(tb), "rate_adpt thr")
        _eq(t1, t2, "rate_adpt in-place thr")
        _eq(ref.update_reference_time_binary_raw(a, s, r0, T, mt, 3, hw, lut[2]),
            O.update_reference_time_binary_raw(a, s, r0, T, mt, 3, hw, lut[2]), "time_binary_raw")
        _eq(ref.update_reference_time_binary_thresh(a, s, r0, T, mt, 3, hw, lut[2]),
            O.update_reference_time_binary_thresh(a, s, r0, T, mt, 3, hw, lut[2]), "time_binary_thresh")
        for pol in ((0, 1, 2) if uncoded else (0, 1, 2, 3)):
            n1, p1, g1 = ref.split_spikes(s, a, pol)
            n2, p2, g2 = O.split_spikes(s, a, pol)
            _eq(n1, n2, "split neg"), _eq(p1, p2, "split pos")
            assert g1 == g2
            ds = int(np.log2(max(W, H))) + 1
            fs, dm = 2 * ds, (1 << ds) - 1
            assert _norm(ref.make_spike_lists_rate(p1, n1, g1, T, fs, ds, dm, mt)) == \
                O.make_spike_lists_rate(p2, n2, g2, T, fs, ds, dm, mt)
            nb = int(rng.integers(1, 10))
            assert _norm(ref.make_spike_lists_time(p1, n1, g1, fs, ds, dm, nb, mt, T, 180)) == \
                O.make_spike_lists_time(p2, n2, g2, fs, ds, dm, nb, mt, T, 180)
            pi, ni = p1.astype(np.int16), n1.astype(np.int16)
            assert ref.make_spike_lists_time_bin(pi, ni, g1, fs, ds, dm, mt, T, 180, 8, lut[3]) == \
                O.make_spike_lists_time_bin(pi, ni, g1, fs, ds, dm, mt, T, 180, 8, lut[3])
            if T >= 5:
                def call(fn):
                    try:
                        return fn(pi, ni, g1, fs, ds, dm, mt, T, 180, 6, lut6)
                    except IndexError:
                        return "IndexError"
                assert call(ref.make_spike_lists_time_bin_thr) == call(O.make_spike_lists_time_bin_thr)


def test_numpy_floor_division_and_fp64_history_weight(refs):
    gs = refs[0]
    O = Oracle()
    a = np.array([[0, 1, 100, 101, 255, 37]], dtype=np.int16)
    s = np.ones_like(a)
    r = np.array([[0, 1, 100, 101, 255, -3]], dtype=np.int16)
    for hw in [x / 100.0 for x in range(1, 100, 7)] + [0.29]:
        _eq(gs.update_reference_rate(a, s, r, 12, 8, hw), O.update_reference_rate(a, s, r, 12, 8, hw), hw)
    # division by a zero / negative per-pixel threshold follows NumPy (0, floor)
    thr = np.array([[0, -7, 5, 12, 1, -1]], dtype=np.int16)
    t1, t2 = thr.copy(), thr.copy()
    ra, ta = gs.update_reference_rate_adpt(a, s, r, t1, 6, 168, 2, 12, 8, 1.0)
    rb, tb = O.update_reference_rate_adpt(a, s, r, t2, 6, 168, 2, 12, 8, 1.0)
    _eq(ra, rb, "ref"), _eq(np.asarray(ta), np.asarray(tb), "thr")


def test_dvs_operator_restatement_matches_the_compiled_reference():
    """orc_dvs_op_f32 (oracle/dvs_oracle.c) against the reference's own DVSOperator::operator()
    (pydvs/cpp/dvs_op.cpp:4-33 compiled unchanged against oracle/opencv_stub): bit-identical float32 planes over
    300 frames per setting, thresholds running up to 1e30 and down to denormals / zero."""
    from oracle import ref_loader
    from oracle import oracle as orc
    op = ref_loader.load_dvs_op()
    if op is None:
        pytest.skip("oracle/_ref/libdvs_op_ref.so not built (python oracle/build_ref.py)")
    rng = np.random.default_rng(11)
    for (H, W, thr0, relax, up, down, scale) in [(24, 40, 10.0, 0.999, 1.5, 0.99, 255.0), (7, 9, 25.0, 1.0, 2.0, 0.5, 255.0),
                                                 (16, 16, 1.0, 0.5, 1.01, 0.999, 1.0), (5, 33, 3.0, 0.999, 10.0, 0.1, 1e4)]:
        a = [np.zeros((H, W), np.float32), np.zeros((H, W), np.float32), np.full((H, W), thr0, np.float32)]
        b = [x.copy() for x in a]
        for t in range(300):
            src = (rng.random((H, W)) * scale).astype(np.float32)
            op(src, a[0], a[1], a[2], relax, up, down)
            orc.dvs_op_f32(src, b[0], b[1], b[2], relax, up, down)
            for x, y, what in zip(a, b, ("diff", "ref", "thr")):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (t, what)
