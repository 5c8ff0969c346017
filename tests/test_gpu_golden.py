"""GPU: the CUDA path (through the C-ABI) against the golden vectors generated from the compiled,
unmodified reference -- fused DVSStreams pipeline and the unfused drop-in functions."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def lists_of(name, t):
    lens = G[name + "_list_vals_len"]
    start = int(lens[:t].sum())
    vals = G[name + "_list_vals"][start:start + int(lens[t])]
    off = G[name + "_list_off"][t]
    return [vals[int(off[j]):int(off[j + 1])].tolist() for j in range(len(off) - 1)]


def run_streams(names, cfg_kwargs, key, adaptive=False, **stream_kwargs):
    """All fixtures in `names` share a geometry: they become the streams of ONE DVSStreams object."""
    from pydvs_b200 import DVSConfig, DVSStreams
    frames = np.stack([G[n + "_frames"] for n in names], axis=1)  # [n_frames, S, H, W]
    n, S, H, W = frames.shape
    dvs = DVSStreams(DVSConfig(width=W, height=H, streams=S, **cfg_kwargs), event_capacity=S * H * W * 8, **stream_kwargs)
    for t in range(n):
        res = dvs.step(torch.from_numpy(frames[t]).cuda())
        for s, name in enumerate(names):
            assert np.array_equal(dvs.ref[s].cpu().numpy(), G[name + "_ref"][t]), (name, t, "ref")
            if adaptive:
                assert np.array_equal(dvs.thr[s].cpu().numpy(), G[name + "_thr"][t]), (name, t, "thr")
            assert res.spike_lists(s, *key) == lists_of(name, t), (name, t, "lists")


@pytest.mark.parametrize("force_generic", [False, True])
def test_cfg1(force_generic):
    run_streams(["cfg1"], dict(output_type="RATE", threshold=12, max_time_ms=8), (14, 7, 127), force_generic=force_generic)


def test_cfg2_four_mnist_streams_in_one_batch():
    run_streams(["cfg2_%d" % i for i in range(4)],
                dict(output_type="TIME_BIN_THR", threshold=12, adaptive_threshold=True, history_weight=0.99, max_time_ms=8,
                     num_bins=6, log2_table=G["cfg2_lut"], ref0=0, threshold0=12), (10, 5, 31), adaptive=True)


@pytest.mark.parametrize("force_generic", [False, True])
def test_cfg3(force_generic):
    run_streams(["cfg3"], dict(output_type="RATE", threshold=12, max_time_ms=8, inh_area_width=2), (12, 6, 63),
                force_generic=force_generic)


def test_cfg4():
    run_streams(["cfg4"], dict(output_type="TIME", threshold=12, adaptive_threshold=True, max_time_ms=4, num_bins=4,
                               threshold0=12), (12, 6, 63), adaptive=True)


def test_uncoded_twin():
    run_streams(["unc_rate"], dict(output_type="RATE", threshold=12, max_time_ms=8, history_weight=0.9, uncoded=True), (10, 5, 31))
    run_streams(["unc_time"], dict(output_type="TIME", threshold=12, max_time_ms=8, num_bins=6, uncoded=True), (10, 5, 31))
    run_streams(["unc_bin"], dict(output_type="TIME_BIN", threshold=12, max_time_ms=8, num_bins=8, uncoded=True,
                                  log2_table=G["unc_lut8"]), (10, 5, 31))
    run_streams(["unc_adpt"], dict(output_type="RATE", threshold=12, adaptive_threshold=True, min_threshold=6, max_threshold=30,
                                   max_time_ms=8, uncoded=True, threshold0=7, encoder_threshold=6), (10, 5, 31), adaptive=True)


def test_dropin_functions_frame_by_frame_cfg3():
    """The reference's own call sequence (stand_alone.py:160-185) on CUDA tensors, against golden."""
    import pydvs_b200.generate_spikes as gs
    frames = G["cfg3_frames"]
    n, H, W = frames.shape
    ref = torch.full((H, W), 128, dtype=torch.int16, device="cuda")
    coords = gs.generate_inh_coords(W, H, 2)
    for t in range(n):
        curr = torch.from_numpy(frames[t]).cuda()
        diff, abs_diff, spikes = gs.thresholded_difference(curr, ref, 12)
        spikes = gs.local_inhibition(spikes, abs_diff, coords, W, H, 2)
        ref[:] = gs.update_reference_rate(abs_diff, spikes, ref, 12, 8, 1.0)
        neg, pos, gmax = gs.split_spikes(spikes, abs_diff, 2)
        lists = gs.make_spike_lists_rate(pos, neg, gmax, 12, 12, 6, 63, 8)
        assert np.array_equal(spikes.cpu().numpy(), G["cfg3_spikes"][t])
        assert np.array_equal(ref.cpu().numpy(), G["cfg3_ref"][t])
        assert lists == lists_of("cfg3", t)


def test_dropin_with_numpy_arrays_like_stand_alone_py():
    """stand_alone.py allocates NumPy buffers itself: the module must accept them unchanged."""
    import pydvs_b200.generate_spikes as gs
    frames = G["cfg1_frames"]
    ref = 128 * np.ones((128, 128), dtype=np.int16)
    diff, abs_diff, spikes = (np.zeros((128, 128), dtype=np.int16) for _ in range(3))
    for t in range(frames.shape[0]):
        diff[:], abs_diff[:], spikes[:] = gs.thresholded_difference(frames[t], ref, 12)
        ref[:] = gs.update_reference_rate(abs_diff, spikes, ref, 12, 8, 1.0)
        neg, pos, gmax = gs.split_spikes(spikes, abs_diff, 2)
        lists = gs.make_spike_lists_rate(pos, neg, gmax, 12, 14, 7, 127, 8)
        assert isinstance(lists, list) and isinstance(neg, np.ndarray)
        assert np.array_equal(ref, G["cfg1_ref"][t]) and lists == lists_of("cfg1", t)


def test_misc_known_answers():
    import pydvs_b200.generate_spikes as gs
    assert np.array_equal(gs.generate_log2_table(4, 8), G["log2_4_8"])
    for r, c, k_pos, k_neg in G["key_probes_ds12"]:
        pos = torch.tensor([[r], [c], [50]], dtype=torch.int16, device="cuda")
        empty = torch.zeros((3, 0), dtype=torch.int16, device="cuda")
        assert gs.make_spike_lists_rate(pos, empty, 50, 12, 24, 12, 255, 8)[0][0] == k_pos   # int16 wrap-around
        assert gs.make_spike_lists_rate(empty, pos, 50, 12, 24, 12, 255, 8)[0][0] == k_neg
    a, s, r = (torch.from_numpy(G[k]).cuda() for k in ("noise_a", "noise_s", "noise_r"))
    np.random.seed(4242)
    got = gs.update_reference_time_binary_raw_noise(a, s, r, 24, 16, 12, 8, 3, 0.9, G["unc_lut8"], 10)
    assert np.array_equal(got.cpu().numpy(), G["noise_raw"])
    np.random.seed(4242)
    got = gs.update_reference_time_thresh_noise(a, s, r, 24, 16, 12, 8, 0.9, 10)
    assert np.array_equal(got.cpu().numpy(), G["noise_thr"])
    for pol in range(4):
        neg, pos, g = gs.split_spikes(s, a, pol)
        assert np.array_equal(neg.cpu().numpy(), G["split_p%d_neg" % pol])
        assert np.array_equal(pos.cpu().numpy(), G["split_p%d_pos" % pol]) and g == int(G["split_p%d_g" % pol])
