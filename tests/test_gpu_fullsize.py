"""GPU: BASELINE.json's full frame sizes (3840x2160 full pipeline, 1920x1080 time + adaptive) checked
through size-independent properties, plus one stream of each against the oracle frame loop."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _events_np(res):
    ev = res.events().cpu().numpy()
    return ev[:, 0].astype(np.uint16).astype(np.int64), ev[:, 1].astype(np.uint16).astype(np.int64), \
        ev[:, 2].astype(np.uint16).astype(np.int64), ev[:, 3].astype(np.int64)


def _check_event_properties(dvs, res, W, H):
    """CSR consistency + the reference's ordering (Appendix A.7) for every (stream, slot, list) segment."""
    off = res.offsets_host()
    x, y, t, p = _events_np(res)
    n_slots = dvs.n_slots
    assert off[0] == 0 and np.all(np.diff(off) >= 0) and off[-1] == len(x)
    assert x.max(initial=0) < W and y.max(initial=0) < H
    seg = np.repeat(np.arange(len(off) - 1), np.diff(off))
    slot, pol = (seg // 2) % n_slots, seg % 2
    assert np.array_equal(t, slot), "event time slot must equal its CSR slot"
    assert np.array_equal(p, np.where(pol == 0, 1, -1)), "pos list carries +1, neg list -1"
    lin = y * W + x
    same = seg[1:] == seg[:-1]
    assert np.all(lin[1:][same] > lin[:-1][same]), "events of a segment must be strictly row-major"
    totals = dvs.stream_totals.view(dvs.S, dvs.C).cpu().numpy()
    per_stream = np.diff(off[::2 * n_slots])
    assert np.array_equal(per_stream, totals[:, 2:].sum(axis=1))
    # rate coding: slot j holds exactly the records with k > j, so segment sizes are non-increasing in j
    if dvs.cfg.output_type == "RATE":
        sizes = np.diff(off).reshape(dvs.S, n_slots, 2)
        assert np.all(sizes[:, 1:, :] <= sizes[:, :-1, :])


@pytest.mark.parametrize("state_dtype", ["int16", "uint8"])
def test_4k_full_pipeline_properties_and_kernel_agreement(state_dtype):
    from oracle import oracle as orc
    from pydvs_b200 import DVSConfig, DVSStreams, synth
    S, H, W = 3, 2160, 3840
    cfg = DVSConfig(width=W, height=H, streams=S, output_type="RATE", threshold=12, max_time_ms=8, inh_area_width=2)
    fast = DVSStreams(cfg, event_capacity=S * H * W, state_dtype=state_dtype)
    generic = DVSStreams(cfg, event_capacity=S * H * W, force_generic=True)   # int16 state, generic body
    pipe = orc.Pipeline(W, H, mode=orc.ENC_RATE, threshold=12, max_time_ms=8, inh_k=2, flag_shift=15, data_shift=12,
                        data_mask=255)
    for t in range(4):
        f = synth.moving_bar(S, H, W, t, noise=8, seed=5) if t != 2 else synth.random_frames(S, H, W, t, seed=5)
        ra, rb = fast.step(f), generic.step(f)
        assert torch.equal(fast.ref, generic.ref), "fast and generic tile kernels must agree bit-for-bit"
        assert torch.equal(ra.seg_offsets, rb.seg_offsets) and torch.equal(ra.events(), rb.events())
        assert int(fast.ref.min()) >= 0 and int(fast.ref.max()) <= 255
        _check_event_properties(fast, ra, W, H)
        exp = pipe.step(f[1].cpu().numpy())            # stream 1 against the oracle, keys wrap like the reference's
        assert np.array_equal(fast.ref[1].cpu().numpy(), pipe.ref)
        assert ra.spike_lists(1, 15, 12, 255) == exp
    # idempotence: a frame equal to the reference plane produces no event and leaves the state alone
    before = fast.ref.clone()
    res = fast.step(before.clone())
    assert res.total_events() == 0 and torch.equal(fast.ref, before)


@pytest.mark.parametrize("state_dtype", ["int16", "uint8"])
def test_1080p_time_adaptive_properties(state_dtype):
    from oracle import oracle as orc
    from pydvs_b200 import DVSConfig, DVSStreams, synth
    S, H, W = 3, 1080, 1920
    cfg = DVSConfig(width=W, height=H, streams=S, output_type="TIME", threshold=12, adaptive_threshold=True, max_time_ms=4,
                    num_bins=4, threshold0=12)
    dvs = DVSStreams(cfg, event_capacity=S * H * W, state_dtype=state_dtype)
    pipe = orc.Pipeline(W, H, mode=orc.ENC_TIME, threshold=12, max_time_ms=4, num_bins=4, adaptive=True, min_thr=6,
                        max_thr=168, down=2, up=12, flag_shift=15, data_shift=11, data_mask=255, thr0=12)
    for t in range(4):
        f = synth.moving_bar(S, H, W, t, noise=8, seed=9)
        res = dvs.step(f)
        _check_event_properties(dvs, res, W, H)
        assert int(dvs.thr.min()) >= 6 and int(dvs.thr.max()) <= 168
        x, y, tt, p = _events_np(res)
        assert len(np.unique(np.stack([res.offsets_host()[:-1]]))) >= 1
        exp = pipe.step(f[2].cpu().numpy())
        assert np.array_equal(dvs.ref[2].cpu().numpy(), pipe.ref) and np.array_equal(dvs.thr[2].cpu().numpy(), pipe.thr)
        assert res.spike_lists(2, 15, 11, 255) == exp
        # time coding: one event per record
        tot = dvs.stream_totals.view(S, dvs.C).cpu().numpy()
        assert np.array_equal(tot[:, 2:].sum(axis=1), tot[:, 0] + tot[:, 1])


def test_float_mode_4k_linearity_of_quiet_pixels():
    """DVSOperator float mode at 4K: pixels that never cross the threshold relax exactly geometrically."""
    from pydvs_b200 import DVSOperatorF32
    H, W = 2160, 3840
    op = DVSOperatorF32(H, W, streams=1, thr_init=1e9, relax=0.5, up=2.0, down=1.0)
    op.ref.fill_(64.0)
    frame = torch.full((1, H, W), 70.0, device="cuda")
    for _ in range(3):
        op.update(frame)
    assert torch.all(op.ref == 8.0) and torch.all(op.diff == 0) and torch.all(op.thr == 1e9)


@pytest.mark.parametrize("state_dtype", ["int16", "uint8"])
@pytest.mark.parametrize("workload", ["4k_rate_inhibition", "1080p_time_adaptive", "4k_dense"])
def test_repeated_steps_are_byte_identical(workload, state_dtype):
    """Determinism stress (stands in for racecheck, which is closed on this pool): the same multi-stream step from the
    same state, 50 times, must produce byte-identical events, CSR row pointers, reference and threshold planes --
    shared-memory histograms, atomics into the redo list and the ordered compaction leave no room for run-to-run order."""
    from pydvs_b200 import DVSConfig, DVSStreams, synth
    if workload == "1080p_time_adaptive":
        S, H, W = 16, 1080, 1920
        cfg = DVSConfig(width=W, height=H, streams=S, output_type="TIME", max_time_ms=4, num_bins=4,
                        adaptive_threshold=True, threshold0=12)
    else:
        S, H, W = 16, 2160, 3840
        cfg = DVSConfig(width=W, height=H, streams=S, output_type="RATE", max_time_ms=8, inh_area_width=2)
    pattern = "dense" if workload == "4k_dense" else "bar"
    dvs = DVSStreams(cfg, event_capacity=S * H * W * (2 if pattern == "dense" else 1), graph=False, state_dtype=state_dtype)
    for t in range(6):     # leave the start-up transient (the generic body handles its overfull tiles) but keep spikes
        dvs.step(synth.hashed_frames(S, H, W, t, pattern=pattern))
    frame = synth.hashed_frames(S, H, W, 6, pattern=pattern)
    ref0 = dvs.ref.clone()
    thr0 = dvs.thr.clone() if dvs.thr is not None else None
    first = None
    for rep in range(50):
        dvs.set_state(ref=ref0, thr=thr0)
        res = dvs.step(frame)
        n = res.total_events()
        got = (res.seg_offsets.clone(), res.events_raw[:n].clone(), dvs.ref.clone(),
               dvs.thr.clone() if thr0 is not None else None)
        if first is None:
            first = got
            assert n > 1000
            continue
        for a, b, what in zip(first, got, ("seg_offsets", "events", "ref", "thr")):
            assert a is None or torch.equal(a, b), (workload, rep, what)
    dvs.check_status()
