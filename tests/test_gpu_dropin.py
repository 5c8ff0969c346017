"""GPU parity of the drop-in module API (through the C-ABI) against the CPU oracle, function by
function, on seeded inputs.  Integer results must be bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(8, 8), (16, 24), (32, 32), (6, 10), (33, 17), (128, 128), (48, 640)]


def _mods(uncoded):
    import pydvs_b200.generate_spikes as gs
    import pydvs_b200.generate_spikes_uncoded as gsu
    from oracle.oracle import Oracle
    return (gsu if uncoded else gs), Oracle(uncoded)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _np(t):
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.uint16:
            return t.view(torch.int16).cpu().numpy().view(np.uint16)
        return t.cpu().numpy()
    return np.asarray(t)


def _frames(rng, shape, lo=0, hi=256):
    return rng.integers(lo, hi, shape).astype(np.int16)


@pytest.mark.parametrize("uncoded", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_thresholded_difference(uncoded, shape):
    m, O = _mods(uncoded)
    rng = np.random.default_rng(hash(shape) % 1000)
    for T in (0, 1, 12, 40, 300, -5):
        c, r = _frames(rng, shape), _frames(rng, shape)
        got = m.thresholded_difference(_cu(c), _cu(r), T)
        exp = O.thresholded_difference(c, r, T)
        for g, e in zip(got, exp):
            assert np.array_equal(_np(g), e)
    # full int16 range incl. wrap-around of the difference and of abs(-32768)
    c, r = _frames(rng, shape, -32768, 32768), _frames(rng, shape, -32768, 32768)
    c.flat[0], r.flat[0] = -32768, 0
    for g, e in zip(m.thresholded_difference(_cu(c), _cu(r), 100), O.thresholded_difference(c, r, 100)):
        assert np.array_equal(_np(g), e)
    thr = _frames(rng, shape, -20, 80)
    c, r = _frames(rng, shape), _frames(rng, shape)
    for g, e in zip(m.thresholded_difference_adpt(_cu(c), _cu(r), _cu(thr), 6, 168),
                    O.thresholded_difference_adpt(c, r, thr, 6, 168)):
        assert np.array_equal(_np(g), e)


def test_numpy_adapter_and_batched():
    m, O = _mods(False)
    rng = np.random.default_rng(5)
    c, r = _frames(rng, (16, 16)), _frames(rng, (16, 16))
    got = m.thresholded_difference(c, r, 12)  # NumPy in -> NumPy out
    assert all(isinstance(g, np.ndarray) for g in got)
    for g, e in zip(got, O.thresholded_difference(c, r, 12)):
        assert np.array_equal(g, e)
    cb, rb = _frames(rng, (5, 16, 24)), _frames(rng, (5, 16, 24))
    d, a, s = m.thresholded_difference(_cu(cb), _cu(rb), 12)
    for i in range(5):
        e = O.thresholded_difference(cb[i], rb[i], 12)
        assert np.array_equal(_np(d[i]), e[0]) and np.array_equal(_np(a[i]), e[1]) and np.array_equal(_np(s[i]), e[2])


@pytest.mark.parametrize("shape,k", [((8, 8), 2), ((16, 24), 4), ((12, 18), 3), ((128, 128), 2), ((48, 640), 2),
                                     ((40, 40), 5), ((16, 16), 8), ((6, 10), 1)])
def test_local_inhibition(shape, k):
    m, O = _mods(False)
    rng = np.random.default_rng(k)
    H, W = shape
    for trial in range(3):
        c, r = _frames(rng, shape, 0, 64 if trial == 0 else 256), _frames(rng, shape, 0, 64 if trial == 0 else 256)
        d, a, s = O.thresholded_difference(c, r, 12)
        s_ref = s.copy()
        O.local_inhibition(s_ref, a, None, W, H, k)
        s_gpu = _cu(s)
        out = m.local_inhibition(s_gpu, _cu(a), m.generate_inh_coords(W, H, k), W, H, k)
        assert out is s_gpu, "local_inhibition must return the object it mutated"
        assert np.array_equal(_np(s_gpu), s_ref)
        s_np = s.copy()
        out = m.local_inhibition(s_np, a, None, W, H, k)
        assert out is s_np and np.array_equal(s_np, s_ref)


def test_local_inhibition_rejects_non_tiling():
    m, _ = _mods(False)
    s = torch.zeros((6, 10), dtype=torch.int16, device="cuda")
    with pytest.raises(IndexError):
        m.local_inhibition(s, s.clone(), None, 10, 6, 4)


@pytest.mark.parametrize("uncoded", [False, True])
@pytest.mark.parametrize("shape", SHAPES[:5])
def test_update_reference_variants(uncoded, shape):
    m, O = _mods(uncoded)
    rng = np.random.default_rng(11)
    lut = O.generate_log2_table(4, 8)
    for hw in (1.0, 0.99, 0.5, 0.29, 0.07):
        for T, mt in ((12, 8), (1, 3), (7, 33), (40, 1)):
            c, r = _frames(rng, shape), _frames(rng, shape)
            d, a, s = O.thresholded_difference(c, r, T)
            assert np.array_equal(_np(m.update_reference_rate(_cu(a), _cu(s), _cu(r), T, mt, hw)),
                                  O.update_reference_rate(a, s, r, T, mt, hw))
            assert np.array_equal(_np(m.update_reference_time_thresh(_cu(a), _cu(s), _cu(r), T, mt, hw)),
                                  O.update_reference_time_thresh(a, s, r, T, mt, hw))
            for row in (1, 2, 3):
                assert np.array_equal(
                    _np(m.update_reference_time_binary_raw(_cu(a), _cu(s), _cu(r), T, mt, row, hw, lut[row])),
                    O.update_reference_time_binary_raw(a, s, r, T, mt, row, hw, lut[row]))
                assert np.array_equal(
                    _np(m.update_reference_time_binary_thresh(_cu(a), _cu(s), _cu(r), T, mt, row, hw, _cu(lut[row]))),
                    O.update_reference_time_binary_thresh(a, s, r, T, mt, row, hw, lut[row]))
            # adaptive: per-pixel thresholds incl. zero (division by zero -> 0) and negatives
            thr = _frames(rng, shape, -3, 60)
            da, aa, sa = O.thresholded_difference_adpt(c, r, thr, 6, 168)
            t_cpu, t_gpu = thr.copy(), _cu(thr)
            r_exp, t_exp = O.update_reference_rate_adpt(aa, sa, r, t_cpu, 6, 168, 2, 12, mt, hw)
            r_got, t_got = m.update_reference_rate_adpt(_cu(aa), _cu(sa), _cu(r), t_gpu, 6, 168, 2, 12, mt, hw)
            assert np.array_equal(_np(r_got), r_exp)
            assert np.array_equal(_np(t_got), np.asarray(t_exp))
            assert np.array_equal(_np(t_gpu), t_cpu), "threshold must be mutated in place like the reference"
            if uncoded:
                assert t_got is t_gpu


def test_update_reference_rate_on_arbitrary_int16_planes():
    """The unfused kernels take whatever int16 planes the caller hands them: negative abs_diff, spikes other than
    -1/0/1, references outside [0, 255], products that wrap (the specialised rate path keeps every int16 wrap)."""
    m, O = _mods(False)
    rng = np.random.default_rng(21)
    for shape in ((16, 24), (33, 17), (64, 128)):
        for T, mt in ((12, 8), (2, 30000), (300, 200), (7, -3), (32767, 5)):
            a = rng.integers(-32768, 32768, shape).astype(np.int16)
            s = rng.integers(-4, 5, shape).astype(np.int16)
            s[rng.integers(0, 2, shape) == 0] = 0
            r = rng.integers(-400, 700, shape).astype(np.int16)
            for hw in (1.0, 0.5):
                with np.errstate(all="ignore"):
                    exp = O.update_reference_rate(a, s, r, T, mt, hw)
                assert np.array_equal(_np(m.update_reference_rate(_cu(a), _cu(s), _cu(r), T, mt, hw)), exp), (shape, T, mt, hw)


def test_local_inhibition_on_arbitrary_int16_planes():
    m, O = _mods(False)
    rng = np.random.default_rng(22)
    for shape in ((16, 24), (64, 128), (6, 10), (8, 4096)):
        a = rng.integers(-50, 50, shape).astype(np.int16)      # ties and negative values: first maximum wins
        s = rng.integers(-2, 3, shape).astype(np.int16)
        s[rng.integers(0, 3, shape) == 0] = 0
        exp = s.copy()
        O.local_inhibition(exp, a, None, shape[1], shape[0], 2)
        got = _cu(s)
        out = m.local_inhibition(got, _cu(a), None, shape[1], shape[0], 2)
        assert out is got and np.array_equal(_np(got), exp), shape
        batch_s, batch_a = np.stack([s, s[::-1].copy()]), np.stack([a, a[::-1].copy()])   # [S, H, W]
        exp_b = batch_s.copy()
        for i in range(2):
            O.local_inhibition(exp_b[i], batch_a[i], None, shape[1], shape[0], 2)
        got_b = _cu(batch_s)
        m.local_inhibition(got_b, _cu(batch_a), None, shape[1], shape[0], 2)
        assert np.array_equal(_np(got_b), exp_b), shape


def test_update_reference_lut_overrun_raises():
    m, O = _mods(False)
    a = np.full((4, 8), 200, dtype=np.int16)
    s = np.ones((4, 8), dtype=np.int16)
    lut = O.generate_log2_table(2, 6)[1]  # 64 entries < 200
    with pytest.raises(IndexError):
        m.update_reference_time_binary_raw(_cu(a), _cu(s), _cu(a), 12, 8, 2, 1.0, lut)


def test_noise_variants_with_injected_samples():
    m, O = _mods(False)
    rng = np.random.default_rng(3)
    H, W = 16, 24
    lut = O.generate_log2_table(4, 8)[2]
    c, r = _frames(rng, (H, W)), _frames(rng, (H, W))
    d, a, s = O.thresholded_difference(c, r, 12)
    sample = rng.integers(0, H * W, size=40).astype(np.int16)
    sample[:3] = sample[3:6]  # duplicates: the XOR must still be applied once
    got = m.update_reference_time_binary_raw_noise(_cu(a), _cu(s), _cu(r), W, H, 12, 8, 3, 0.9, lut, 10, _sample=sample)
    assert np.array_equal(_np(got), O.update_reference_noise(0, a, s, r, W, H, 12, 8, 0.9, lut, sample))
    got = m.update_reference_time_thresh_noise(_cu(a), _cu(s), _cu(r), W, H, 12, 8, 0.9, 10, _sample=sample)
    assert np.array_equal(_np(got), O.update_reference_noise(1, a, s, r, W, H, 12, 8, 0.9, None, sample))
    # default path draws from the global NumPy RNG exactly like the reference does
    np.random.seed(1234)
    num = int((10 / 100.) * (W * H))
    expect_sample = np.random.randint(0, H * W, size=num).astype(np.int16)
    np.random.seed(1234)
    got = m.update_reference_time_thresh_noise(_cu(a), _cu(s), _cu(r), W, H, 12, 8, 1.0, 10)
    assert np.array_equal(_np(got), O.update_reference_noise(1, a, s, r, W, H, 12, 8, 1.0, None, expect_sample))


@pytest.mark.parametrize("uncoded", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_split_and_encoders(uncoded, shape):
    m, O = _mods(uncoded)
    rng = np.random.default_rng(17)
    H, W = shape
    ds = int(np.ceil(np.log2(max(W, H))))
    fs, dm = min(2 * ds, 15), (1 << ds) - 1 if ds < 8 else 255
    lut8 = O.generate_log2_table(4, 8)
    lut6 = O.generate_log2_table(2, 6)[1]
    for trial, T in enumerate((12, 5, 30)):
        c, r = _frames(rng, shape), _frames(rng, shape)
        if trial == 2:
            c[:] = r  # no spikes at all
        d, a, s = O.thresholded_difference(c, r, T)
        for pol in ((0, 1, 2) if uncoded else (0, 1, 2, 3)):
            n_e, p_e, g_e = O.split_spikes(s, a, pol)
            n_g, p_g, g_g = m.split_spikes(_cu(s), _cu(a), pol)
            assert _np(n_g).dtype == n_e.dtype and np.array_equal(_np(n_g), n_e)
            assert np.array_equal(_np(p_g), p_e) and g_g == g_e
            for mt in (8, 3, 16):
                assert m.make_spike_lists_rate(p_g, n_g, g_g, T, fs, ds, dm, mt) == \
                    O.make_spike_lists_rate(p_e, n_e, g_e, T, fs, ds, dm, mt)
            for nb in (8, 4, 1):
                assert m.make_spike_lists_time(p_g, n_g, g_g, fs, ds, dm, nb, 8, T, 180) == \
                    O.make_spike_lists_time(p_e, n_e, g_e, fs, ds, dm, nb, 8, T, 180)
            pi, ni = p_e.astype(np.int16), n_e.astype(np.int16)  # the uncoded _bin encoders need int16
            assert m.make_spike_lists_time_bin(_cu(pi), _cu(ni), g_e, fs, ds, dm, 8, T, 180, 8, lut8[3]) == \
                O.make_spike_lists_time_bin(pi, ni, g_e, fs, ds, dm, 8, T, 180, 8, lut8[3])
            if T >= 5:  # keeps val // T inside the 64-entry table
                try:
                    exp = O.make_spike_lists_time_bin_thr(pi, ni, g_e, fs, ds, dm, 8, T, 180, 6, lut6)
                except IndexError:
                    with pytest.raises(IndexError):
                        m.make_spike_lists_time_bin_thr(_cu(pi), _cu(ni), g_e, fs, ds, dm, 8, T, 180, 6, lut6)
                else:
                    assert m.make_spike_lists_time_bin_thr(_cu(pi), _cu(ni), g_e, fs, ds, dm, 8, T, 180, 6, lut6) == exp


def test_survey_known_answers():
    """2x2 known-answer vectors probed on the compiled reference."""
    m, _ = _mods(False)
    s = _cu(np.array([[1, -1], [-1, 1]], dtype=np.int16))
    a = _cu(np.array([[30, 40], [50, 60]], dtype=np.int16))
    neg, pos, g = m.split_spikes(s, a, 2)
    assert _np(neg).tolist() == [[0, 1], [1, 0], [40, 50]] and _np(pos).tolist() == [[0, 1], [0, 1], [30, 60]] and g == 60
    assert m.make_spike_lists_rate(pos, neg, g, 12, 2, 1, 1, 8) == \
        [[1, 7, 4, 2], [1, 7, 4, 2], [1, 4, 2], [1, 4], [1], [], [], []]
    assert m.make_spike_lists_time(pos, neg, g, 2, 1, 1, 8, 8, 12, 180) == [[], [], [], [7], [2], [4], [1], []]
    mu, _ = _mods(True)
    negu, posu, gu = mu.split_spikes(s, a, 2)
    lists = mu.make_spike_lists_rate(posu, negu, gu, 12, 2, 1, 1, 8)
    first = [[0, 0, 1], [1, 1, 1], [0, 1, -1], [1, 0, -1]]
    assert lists[0] == first and lists[1] == first and [len(x) for x in lists] == [4, 4, 3, 2, 1, 0, 0, 0]


def test_argument_validation_like_cython():
    m, _ = _mods(False)
    f = torch.zeros((4, 8), dtype=torch.float32, device="cuda")
    i = torch.zeros((4, 8), dtype=torch.int16, device="cuda")
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        m.thresholded_difference(f, i, 12)
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        m.thresholded_difference(i[0], i[0], 12)
    with pytest.raises(OverflowError):
        m.thresholded_difference(i, i, 70000)
    with pytest.raises(TypeError):
        m.thresholded_difference(i, i)
    neg, pos, g = m.split_spikes(i, i, 2)
    with pytest.raises(OverflowError):
        m.make_spike_lists_rate(pos, neg, g, 12, 300, 1, 1, 8)
    with pytest.raises(TypeError):  # KEY_XYP is dead in the reference (Appendix A.6)
        m.make_spike_lists_rate(pos, neg, g, 12, 2, 1, 1, 8, key_coding=1)


def test_float_mode_matches_dvs_operator():
    from oracle import oracle as orc
    from pydvs_b200 import DVSOperatorF32
    rng = np.random.default_rng(0)
    S, H, W = 3, 37, 53
    op = DVSOperatorF32(H, W, streams=S, thr_init=10.0, relax=0.999, up=1.5, down=0.99)
    ref = np.zeros((S, H, W), dtype=np.float32)
    thr = np.full((S, H, W), 10.0, dtype=np.float32)
    diff = np.zeros_like(ref)
    for t in range(25):
        frame = rng.integers(0, 256, (S, H, W)).astype(np.float32)
        orc.dvs_op_f32(frame, diff, ref, thr, 0.999, 1.5, 0.99)
        op.update(torch.from_numpy(frame).cuda())
        # north star: 1e-6 relative; the kernel forbids FMA contraction, so this is bit-exact
        assert np.array_equal(op.ref.cpu().numpy(), ref)
        assert np.array_equal(op.thr.cpu().numpy(), thr)
        assert np.array_equal(op.diff.cpu().numpy(), diff)
        np.testing.assert_allclose(op.thr.cpu().numpy(), thr, rtol=1e-6)


def test_float_mode_matches_the_reference_built_dvs_operator():
    """k_step_f32 against planes produced by the reference's own DVSOperator::operator() (pydvs/cpp/dvs_op.cpp compiled
    unchanged, tests/golden/make_golden_float.py) and, where oracle/_ref travelled, against that library live."""
    import os
    from oracle import ref_loader
    from pydvs_b200 import DVSOperatorF32
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_float_v1.npz"))
    for name in sorted({k.split("/")[0] for k in g.files}):
        frames, (thr0, relax, up, down), snaps = g[name + "/frames"], g[name + "/params"], g[name + "/snaps"]
        n, H, W = frames.shape
        op = DVSOperatorF32(H, W, streams=1, thr_init=float(thr0), relax=float(relax), up=float(up), down=float(down))
        dev = torch.from_numpy(frames).cuda()
        for t in range(n):
            op.update(dev[t:t + 1].contiguous())
            if (t + 1) % 10 == 0:
                exp = snaps[(t + 1) // 10 - 1]
                for got, e, what in zip((op.diff, op.ref, op.thr), exp, ("diff", "ref", "thr")):
                    assert np.array_equal(got[0].cpu().numpy().view(np.uint32), e.view(np.uint32)), (name, t, what)
    live = ref_loader.load_dvs_op()
    if live is None:
        return
    rng = np.random.default_rng(5)
    S, H, W = 2, 37, 53
    op = DVSOperatorF32(H, W, streams=S, thr_init=10.0, relax=0.999, up=1.5, down=0.99)
    ref, thr, diff = (np.zeros((S, H, W), np.float32), np.full((S, H, W), 10.0, np.float32), np.zeros((S, H, W), np.float32))
    for t in range(120):
        frame = rng.uniform(0, 255, (S, H, W)).astype(np.float32)
        for s in range(S):
            live(frame[s], diff[s], ref[s], thr[s], 0.999, 1.5, 0.99)
        op.update(torch.from_numpy(frame).cuda())
    for got, e in zip((op.diff, op.ref, op.thr), (diff, ref, thr)):
        assert np.array_equal(got.cpu().numpy().view(np.uint32), e.view(np.uint32))
