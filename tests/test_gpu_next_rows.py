"""GPU parity of the rows next to the hot path through the C-ABI:
``pydvs_b200.decode_spikes`` (decode_spikes.pyx) and ``pydvs_b200.numba_nvs`` (numba_nvs.py) against the
CPU oracle and the golden vectors generated from the reference.  Integer planes and float32 planes must
both be bit-exact (the float variants reproduce numba's evaluation precision exactly)."""
import numpy as np
import pytest
import torch

from test_oracle_next_rows import G, NVS_CASES, feq, nvs_kwargs, replay_time_case

pytestmark = pytest.mark.gpu


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- decoder ---------------------------------------------------------------------------------------
def _keys(rng, n, H, W, fs, ds):
    return ((rng.integers(0, 2, n) << fs) | (rng.integers(0, H, n) << ds) | rng.integers(0, W, n)).astype(np.int16)


@pytest.mark.parametrize("i", range(int(G["dec_rate_n"][0])))
def test_decode_rate_golden(i):
    import pydvs_b200.decode_spikes as dec
    thr, fs, ds, mask = (int(v) for v in G["dec_rate%d_par" % i])
    hw = float(G["dec_rate%d_hw" % i][0])
    out = dec.update_ref_spike_list_rate(G["dec_rate%d_ref" % i].copy(), G["dec_rate%d_keys" % i], thr, hw, fs, ds, mask)
    assert isinstance(out, np.ndarray) and np.array_equal(out, G["dec_rate%d_out" % i])
    out = dec.update_ref_spike_list_rate(_cu(G["dec_rate%d_ref" % i]), _cu(G["dec_rate%d_keys" % i]), thr, hw, fs, ds, mask)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), G["dec_rate%d_out" % i])


@pytest.mark.parametrize("i", range(int(G["dec_time_n"][0])))
@pytest.mark.parametrize("device_tensors", [False, True])
def test_decode_time_golden(i, device_tensors):
    import pydvs_b200.decode_spikes as dec
    if device_tensors:
        replay_time_case(i, dec.update_ref_spike_list_time, dec.update_ref_spike_list_time_bin, _cu,
                         lambda t: t.cpu().numpy())
    else:
        replay_time_case(i, dec.update_ref_spike_list_time, dec.update_ref_spike_list_time_bin)


def test_decoders_random_vs_oracle():
    import pydvs_b200.decode_spikes as dec
    from oracle import oracle as orc
    rng = np.random.default_rng(17)
    for trial in range(40):
        H, W, fs, ds, mask = [(16, 16, 8, 4, 15), (64, 128, 14, 7, 127), (8, 8, 6, 3, 7)][trial % 3]
        ref = rng.integers(-50, 300, (H, W)).astype(np.int16)
        k = _keys(rng, int(rng.integers(0, 400)), H, W, fs, ds)   # duplicates accumulate
        thr, hw = int(rng.integers(1, 40)), float(rng.choice([1.0, 0.99, 0.5, 0.29]))
        assert np.array_equal(dec.update_ref_spike_list_rate(ref.copy(), k, thr, hw, fs, ds, mask),
                              orc.update_ref_spike_list_rate(ref.copy(), k, thr, hw, fs, ds, mask))
    for binary in (0, 1):
        f_gpu = dec.update_ref_spike_list_time_bin if binary else dec.update_ref_spike_list_time
        f_orc = orc.update_ref_spike_list_time_bin if binary else orc.update_ref_spike_list_time
        for trial in range(30):
            a = [rng.integers(0, 256, (8, 8)).astype(np.int16), np.zeros((8, 8), np.int16), np.full((8, 8), -1, np.int16)]
            b = [x.copy() for x in a]
            nb, bw, per = int(rng.integers(2, 9)), float(rng.choice([1.0, 0.5, 2.0, 1.5])), int(rng.integers(1, 20))
            thr, hw, t = int(rng.integers(1, 20)), float(rng.choice([1.0, 0.99])), int(rng.integers(0, 70000))
            for _step in range(5):
                k = _keys(rng, int(rng.integers(0, 40)), 8, 8, 6, 3)   # pixels repeat: several rounds on the GPU
                t += int(rng.integers(1, 40))
                ea = eb = None
                try:
                    a = list(f_gpu(a[0], a[1], a[2], k, nb, bw, per, t, thr, hw, 6, 3, 7))
                except (ValueError, OverflowError) as e:
                    ea = type(e)
                try:
                    b = list(f_orc(b[0], b[1], b[2], k, nb, bw, per, t, thr, hw, 6, 3, 7))
                except (ValueError, OverflowError) as e:
                    eb = type(e)
                assert ea is eb, (binary, trial, ea, eb)
                if ea is not None:
                    break
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), (binary, trial)


def test_decoder_errors():
    import pydvs_b200.decode_spikes as dec
    ref = np.zeros((8, 8), np.int16)
    with pytest.raises(OverflowError):
        dec.update_ref_spike_list_rate(ref, np.array([3, -5], np.int16), 12, 1.0, 6, 3, 7)
    with pytest.raises(IndexError):
        dec.update_ref_spike_list_rate(ref, np.array([(9 << 4) | 1], np.int16), 12, 1.0, 8, 4, 15)
    with pytest.raises(ValueError):
        dec.update_ref_spike_list_rate(ref.astype(np.int32), np.array([3], np.int16), 12, 1.0, 6, 3, 7)
    with pytest.raises(OverflowError):
        dec.update_ref_spike_list_rate(ref, np.array([3], np.int16), 70000, 1.0, 6, 3, 7)
    assert dec.key_to_spike(0x123, 8, 4, 15) == (2, 3, -1)
    with pytest.raises(OverflowError):
        dec.key_to_spike(-1, 8, 4, 15)


def test_encode_decode_round_trip_many_streams():
    """Size-independent property: with the uncoded rate encoder (k = min(max_time_ms - 1, v) events of weight
    `threshold` per pixel) and history_weight 1, the receiver's plane moves by exactly the opposite of the
    sender's reference (the decoder maps flag 1 to -1) as long as v < max_time_ms and nothing
    clips: ref_sender + ref_receiver == 2 * ref0."""
    import pydvs_b200.decode_spikes as dec
    from pydvs_b200 import DVSConfig, DVSStreams
    S, H, W, T, mt = 96, 128, 128, 12, 16
    cfg = DVSConfig(width=W, height=H, streams=S, output_type="RATE", threshold=T, max_time_ms=mt, uncoded=True, ref0=128)
    dvs = DVSStreams(cfg, event_capacity=S * H * W * 10)
    receiver = torch.full((S, H, W), 128, dtype=torch.int16, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(4):
        # frames within 128 +- 40: references stay well inside [0, 255] and v = |d| // 12 <= 8 < max_time_ms - 1
        frames = (128 + torch.randint(-40, 41, (S, H, W), device="cuda", generator=g)).to(torch.int16)
        res = dvs.step(frames)
        keys = res.keys(14, 7, 127)                     # uncoded layout, all streams, stream-major
        assert keys.numel() == res.total_events() > 0
        # per-stream offsets: every (2 * n_slots)-th entry of the CSR over (stream, slot, polarity)
        offs = res.seg_offsets[::2 * dvs.n_slots].contiguous()
        assert offs.numel() == S + 1
        receiver = dec.decode_rate_batch(receiver, keys, offs, T, 1.0, 14, 7, 127)
    assert torch.equal(dvs.ref.to(torch.int32) + receiver.to(torch.int32),
                       torch.full((S, H, W), 256, dtype=torch.int32, device="cuda"))


# ---- numba float variants ------------------------------------------------------------------------
@pytest.mark.parametrize("tag,mode,mask", NVS_CASES)
@pytest.mark.parametrize("device_tensors", [False, True])
def test_nvs_golden(tag, mode, mask, device_tensors):
    import pydvs_b200.numba_nvs as nvs
    kw, _ = nvs_kwargs()
    conv = _cu if device_tensors else (lambda a: a.copy())
    r, t = conv(G["nvs_ref0"]), conv(G["nvs_thr0"])
    s = conv(np.zeros_like(G["nvs_ref0"]))
    m = None if mask is None else np.full(G["nvs_ref0"].shape, mask, np.uint8)
    a = (kw["reference_leak"], kw["threshold_base"], kw["threshold_mult_incr"], kw["threshold_leak"])
    for f in G["nvs_frames"]:
        f = conv(f)
        if tag == "nvs_0":
            nvs.nvs_0(f, r, t, s)
        elif tag == "nvs_leaky":
            nvs.nvs_leaky(f, r, t, s, a[0])
        elif tag == "nvs_leaky_adaptive":
            nvs.nvs_leaky_adaptive(f, r, t, s, *a)
        else:
            nvs.nvs_noisy_leaky_adaptive(f, r, t, s, *a, 0.5, leak_mask=m)
    host = (lambda x: x.cpu().numpy()) if device_tensors else (lambda x: x)
    assert feq(host(r), G[tag + "_ref"]) and feq(host(t), G[tag + "_thr"]) and feq(host(s), G[tag + "_spikes"])


@pytest.mark.parametrize("tag,mask", [("nvs_bi_noisy_leaky_adaptive_all", 1), ("nvs_bi_noisy_leaky_adaptive_none", 0)])
def test_nvs_two_cell_golden(tag, mask):
    import pydvs_b200.numba_nvs as nvs
    kw, target = nvs_kwargs()
    r, t = G["nvs_ref0"].copy(), G["nvs_thr0"].copy()
    ro, to = r[::-1].copy(), t[::-1].copy()
    s = np.zeros_like(r)
    m = np.full(r.shape, mask, np.uint8)
    for f in G["nvs_frames"]:
        nvs.nvs_bi_noisy_leaky_adaptive(f.copy(), r, t, s, kw["reference_leak"], kw["threshold_base"],
                                        kw["threshold_mult_incr"], kw["threshold_leak"], 0.5, ro, to, target,
                                        leak_mask=m, leak_mask_off=m)
    for got, name in ((r, "_ref"), (t, "_thr"), (s, "_spikes"), (ro, "_ref_off"), (to, "_thr_off")):
        assert feq(got, G[tag + name]), name


def test_nvs_thresholded_difference():
    import pydvs_b200.numba_nvs as nvs
    r, t = G["nvs_ref0"].copy(), G["nvs_thr0"].copy()
    s = np.zeros_like(r)
    active = nvs.thresholded_difference(G["nvs_frames"][0].copy(), r, t, s)
    assert active.dtype == np.uint8 and np.array_equal(active, G["nvs_td_active"])
    assert np.array_equal(r, G["nvs_ref0"])  # the reference plane is left alone


def test_nvs_random_batched_vs_oracle():
    """[S, H, W] batches with random per-pixel leak masks, odd sizes (scalar tail) and unaligned views."""
    import pydvs_b200.numba_nvs as nvs
    from oracle import oracle as orc
    rng = np.random.default_rng(23)
    kw = dict(reference_leak=0.9, threshold_base=7.5, threshold_mult_incr=0.77, threshold_leak=0.81)
    a = (0.9, 7.5, 0.77, 0.81)
    for shape in ((3, 33, 47), (2, 64, 128), (1, 5, 7)):
        f = rng.uniform(0, 255, shape).astype(np.float32)
        planes = [rng.uniform(0, 255, shape).astype(np.float32), rng.uniform(-3, 40, shape).astype(np.float32),
                  np.zeros(shape, np.float32), rng.uniform(0, 255, shape).astype(np.float32),
                  rng.uniform(1, 40, shape).astype(np.float32)]
        m1, m2 = rng.integers(0, 2, shape).astype(np.uint8), rng.integers(0, 2, shape).astype(np.uint8)
        exp = [p.copy() for p in planes]
        got = [_cu(p) for p in planes]
        fc = _cu(f)
        for _it in range(3):
            orc.nvs_bi_step(f, exp[0], exp[1], exp[2], *a, exp[3], exp[4], 99.5, leak_mask=m1, leak_mask_off=m2)
            nvs.nvs_bi_noisy_leaky_adaptive(fc, got[0], got[1], got[2], *a, 0.5, got[3], got[4], 99.5, leak_mask=m1,
                                            leak_mask_off=m2)
        assert all(feq(g.cpu().numpy(), e) for g, e in zip(got, exp)), shape
        exp = [p.copy() for p in planes[:3]]
        got = [_cu(p) for p in planes[:3]]
        for _it in range(3):
            orc.nvs_step(orc.NVS_NOISY_LEAKY_ADAPTIVE, f, exp[0], exp[1], exp[2], leak_mask=m1, **kw)
            nvs.nvs_noisy_leaky_adaptive(fc, got[0], got[1], got[2], *a, 0.5, leak_mask=_cu(m1))
        assert all(feq(g.cpu().numpy(), e) for g, e in zip(got, exp)), shape


def test_nvs_device_generated_mask_is_seeded_and_has_the_right_rate():
    import pydvs_b200.numba_nvs as nvs
    shape = (4, 128, 256)
    outs = []
    for _rep in range(2):
        g = torch.Generator(device="cuda").manual_seed(77)
        f = torch.full(shape, 50.0, device="cuda")
        r = torch.full(shape, 100.0, device="cuda")
        t = torch.full(shape, 1000.0, device="cuda")   # never spikes: the reference only leaks
        s = torch.zeros(shape, device="cuda")
        nvs.nvs_noisy_leaky_adaptive(f, r, t, s, 0.5, 1000.0, 0.0, 1.0, 0.25, generator=g)
        outs.append(r.clone())
    assert torch.equal(outs[0], outs[1])
    leaked = (outs[0] == 50.0).float().mean().item()
    assert abs(leaked - 0.25) < 0.01 and bool(((outs[0] == 50.0) | (outs[0] == 100.0)).all())


# ---- batched frame animators -------------------------------------------------------------------------
def test_traverse_and_fade_golden_batched():
    """Every golden case is one stream of a single batched launch."""
    from pydvs_b200 import frame_sources as fs
    img = _cu(G["anim_img"]).unsqueeze(0)
    for sp in np.unique(G["anim_trav_par"][:, 1]):
        sel = G["anim_trav_par"][:, 1] == sp
        fn = G["anim_trav_par"][sel, 0].astype(np.int32)
        out = fs.traverse_images(img, np.zeros(len(fn), np.int32), fn, float(sp), bg=37)
        assert np.array_equal(out.cpu().numpy(), G["anim_trav_out"][sel]), sp
    for hf in np.unique(G["anim_fade_par"][:, 1]):
        sel = G["anim_fade_par"][:, 1] == hf
        fn = G["anim_fade_par"][sel, 0].astype(np.int32)
        out = fs.fade_images(img, np.zeros(len(fn), np.int32), fn, int(hf), bg=37)
        assert np.array_equal(out.cpu().numpy(), G["anim_fade_out"][sel]), hf


def test_traverse_and_fade_random_vs_oracle():
    from oracle import frame_oracle as fo
    from pydvs_b200 import frame_sources as fs
    rng = np.random.default_rng(9)
    for H, W in ((28, 28), (33, 47), (64, 128)):
        imgs = rng.integers(0, 256, (5, H, W)).astype(np.int16)
        S = 24
        idx = rng.integers(0, 5, S).astype(np.int32)
        fn = rng.integers(0, 3 * W, S).astype(np.int32)
        for sp in (0.5, 1.0, 2.6, -0.7):
            out = fs.traverse_images(_cu(imgs), idx, fn, sp, bg=11).cpu().numpy()
            for s in range(S):
                assert np.array_equal(out[s], fo.traverse_image(imgs[idx[s]], int(fn[s]), sp, 11)), (H, W, sp, s)
        for hf in (3, 20, 45):
            out = fs.fade_images(_cu(imgs), idx, fn, hf, bg=11).cpu().numpy()
            for s in range(S):
                assert np.array_equal(out[s], fo.fade_image(imgs[idx[s]], int(fn[s]), hf, 11)), (H, W, hf, s)


# ---- deterministic batched VirtualCam --------------------------------------------------------------------
@pytest.mark.parametrize("behaviour", ["SACCADE", "TRAVERSE", "FADE", "NONE", "ATTENTION"])
@pytest.mark.parametrize("with_mask", [False, True])
def test_virtual_cam_schedule_matches_scalar_restatement(behaviour, with_mask):
    # the restatement is pinned against the reference's own VirtualCam class (tests/test_virtual_cam_reference.py)
    from oracle.virtual_cam_oracle import ReplayDraws, VirtualCamOracle
    from pydvs_b200.virtual_cam import VirtualCam
    rng = np.random.default_rng(4)
    N, H, W, S = 3, 32, 32, 5
    imgs = rng.integers(0, 256, (N, H, W)).astype(np.int16)
    mask = rng.uniform(0, 1, (H, W)) if with_mask else None
    start = np.array([0, 1, 2, 1, 0])
    kw = dict(fps=30, image_on_time_ms=400, inter_off_time_ms=100, max_saccade_distance=2,
              frames_per_microsaccade=2, max_cycles=2, background_gray=9, frames_per_saccade=5)
    cam = VirtualCam(_cu(imgs), behaviour=behaviour, fade_mask=mask, streams=S, start_index=start, seed=11, **kw)
    replay = ReplayDraws(11, S)      # the same vectorised draws, handed out one camera at a time
    refs = [VirtualCamOracle(imgs, behaviour, fade_mask=mask, start_index=int(start[s]), rng=replay.stream(s), **kw)
            for s in range(S)]
    n_valid = 0
    for step in range(140):
        dvs_ref = rng.integers(0, 256, (S, H, W)).astype(np.int16)    # stands in for the DVS reference planes (ATTENTION)
        ok, frames = cam.read(_cu(dvs_ref))
        host = frames.cpu().numpy()
        centers = cam.center
        for s in range(S):
            ok_ref, img = refs[s].read(dvs_ref[s])
            assert ok == ok_ref, (step, s)
            assert np.array_equal(host[s], img), (behaviour, step, s)
            if behaviour in ("SACCADE", "ATTENTION"):
                assert tuple(centers[s]) == (refs[s].cx, refs[s].cy), (step, s)
        n_valid += ok
        if not ok and behaviour != "NONE":
            break
    assert n_valid > 50
    if behaviour != "NONE":
        assert not ok   # max_cycles reached: the camera reports end of stream and serves the background


def test_virtual_cam_replays_the_reference_class_recordings():
    """tests/golden/golden_vcam_v1.npz holds what the reference's own VirtualCam class served on its mnist/t10k fixtures
    (tests/golden/make_golden_vcam.py) together with the random draws it made; the GPU camera fed the same draws must
    serve the same frames.  ATTENTION frames whose five largest changes are tied are unspecified in the reference."""
    import os
    from pydvs_b200.virtual_cam import VirtualCam
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_vcam_v1.npz"))
    fps, on_ms, off_ms, max_dist, fp_usacc, fp_sacc, cycles, bg = (int(v) for v in g["args"])
    images, mask = g["images"], g["mask"]

    class Recorded:
        def __init__(self, steps, picks):
            self.steps, self.pk, self.i, self.j = steps, picks, 0, 0

        def offsets(self, S, d):
            self.i += 1
            return self.steps[self.i - 1:self.i]

        def picks(self, S):
            self.j += 1
            return self.pk[self.j - 1:self.j]

    for name in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        behaviour, masked = name.split("_")[0], name.endswith("_mask")
        frames, valid, centers = g[name + "/frames"], g[name + "/valid"], g[name + "/centers"]
        tied = set(int(i) for i in g[name + "/tied_at"])
        cam = VirtualCam(_cu(images), behaviour=behaviour, fps=fps, image_on_time_ms=on_ms, inter_off_time_ms=off_ms,
                         max_saccade_distance=max_dist, frames_per_microsaccade=fp_usacc, frames_per_saccade=fp_sacc,
                         max_cycles=cycles, background_gray=bg, fade_mask=mask if masked else None, streams=1,
                         draws=Recorded(g[name + "/steps"], g[name + "/picks"]))
        compared = 0
        for i in range(len(frames)):
            ref = _cu(g[name + "/ref_planes"][i][None]) if behaviour == "ATTENTION" else None
            ok, fr = cam.read(ref)
            assert bool(ok) == bool(valid[i]), (name, i)
            if i in tied and tuple(cam.center[0]) != tuple(centers[i]):
                cam.cx[:], cam.cy[:] = int(centers[i][0]), int(centers[i][1])
                cam.current_image = _cu(frames[i][None])
                continue
            assert np.array_equal(fr[0].cpu().numpy(), frames[i]), (name, i)
            compared += 1
        assert compared >= len(frames) - 6, name


def test_virtual_cam_attention_runs_and_recentres():
    from pydvs_b200.virtual_cam import VirtualCam
    rng = np.random.default_rng(8)
    imgs = rng.integers(0, 256, (2, 32, 32)).astype(np.int16)
    cam = VirtualCam(_cu(imgs), behaviour="ATTENTION", fps=30, image_on_time_ms=2000, frames_per_saccade=5,
                     streams=3, seed=2)
    ref = torch.full((3, 32, 32), 128, dtype=torch.int16, device="cuda")
    for step in range(12):
        ok, frames = cam.read(ref)
        assert ok and frames.shape == (3, 32, 32) and frames.dtype == torch.int16
        assert np.abs(cam.center).max() <= 16 + 12


# ---- the whole offline flow: frame source -> fused step -> spike files -----------------------------------
def test_images_to_spike_files_matches_the_cpu_restatement():
    """examples/images_to_spike_files.py (mnist_to_spikes.py's flow, batched on the GPU) against the same flow on the
    CPU: the script's micro-saccade loop restated with the pinned move_image -> oracle pipeline -> the reference's
    writer loop."""
    import importlib.util
    import os
    from oracle import oracle as orc
    from oracle import frame_oracle as fo
    from pydvs_b200.frame_sources import saccade_offsets
    import pydvs_b200.generate_spikes as gs
    spec = importlib.util.spec_from_file_location(
        "_example", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "images_to_spike_files.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    n, frames, fps = 6, 8, 100
    imgs = ex.blob_images(n, seed=3)
    writers = ex.convert(imgs, frames=frames, fps=fps, seed=5)
    frame_ms = int(1000. / fps)
    lut = gs.generate_log2_table(2, 8)[1]
    total = 0
    offsets = saccade_offsets(n, frames, 1, 1, seed=5)      # mnist_to_spikes.py:246-258 with seeded draws
    for s in range(n):
        pipe = orc.Pipeline(32, 32, mode=orc.ENC_TIME_BIN_THR, threshold=12, max_time_ms=frame_ms, num_bins=6,
                            history_weight=0.99, adaptive=True, min_thr=6, max_thr=168, down=2, up=12, lut=lut,
                            flag_shift=10, data_shift=5, data_mask=31, ref0=0, thr0=12)
        lines, t = [], 0
        for f in range(frames):
            frame = fo.move_image(imgs[s], int(offsets[f, s, 0]), int(offsets[f, s, 1]), 0)
            t_idx = 0
            for slot in pipe.step(frame):                  # mnist_to_spikes.py:290-303
                for key in slot:
                    lines.append("%s %f" % (key, t + t_idx))
                t_idx += max(1, frame_ms // 6)
            t += frame_ms
        assert writers[s].getvalue() == "\n".join(lines), s
        total += len(lines)
    assert total > 100
