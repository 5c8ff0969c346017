"""One-byte state planes (``DVSStreams(state_dtype="uint8")``, dvs_step_params.state_dtype == DVS_STATE_U8).

Every update rule of the reference ends in ``clip(.., 0, 255)`` (generate_spikes.pyx:251, 297, 378, 425) and the coded
threshold rule in ``clip(.., min, max)`` (:297), so a reference / threshold plane that starts inside [0, 255] never
leaves it and can be kept in bytes.  These tests run the same frame loops as tests/test_gpu_streams.py against the CPU
oracle (which works on int16 planes like the reference): reference planes, threshold planes and the full AER
lists must be identical for every kernel family that reads the state."""
import numpy as np
import pytest
import torch

from test_gpu_streams import _run_case

pytestmark = pytest.mark.gpu
U8 = dict(state_dtype="uint8", debug=False)


def test_headline_shape_ranked_kernel():
    # BASELINE config 5 shape: 2x2 inhibition on two-row tiles, coded rate / time, int16 and uint8 frames
    from pydvs_b200 import _lib
    for kw in (dict(), dict(frame_dtype="uint8"), dict(output_type="TIME", num_bins=8), dict(split_lists=False)):
        assert _run_case(4, 64, 480, inh=2, source="bar", frames=6, **U8, **kw).kernel_path == _lib.PATH_FAST
        _run_case(2, 8, 3840, inh=2, source="random", frames=4, **U8, **kw)          # two column blocks per thread
    _run_case(3, 32, 3840, inh=2, source="random", frames=3, **U8)


def test_linear_tiles_static_threshold():
    for kw in (dict(), dict(output_type="TIME", num_bins=8), dict(uncoded=True), dict(max_time_ms=33, threshold=3),
               dict(frame_dtype="uint8"), dict(polarity=0), dict(polarity=1), dict(polarity=3)):
        _run_case(2, 16, 64, frames=4, **U8, **kw)
    _run_case(2, 8, 3840, source="bar", frames=5, **U8)                                # four groups per thread
    _run_case(1, 128, 128, source="bar", frames=10, **U8)                              # BASELINE config 1
    _run_case(3, 26, 346, frames=4, source="bar", **U8)                                # rows not a multiple of 8 pixels


def test_adaptive_thresholds():
    # BASELINE config 4 shape (reduced) and the other adaptive kernels; thresholds live in bytes as well
    _run_case(5, 54, 96, output_type="TIME", adaptive=True, max_time_ms=4, num_bins=4, source="bar", frames=8, thr0=12, **U8)
    _run_case(2, 8, 1920, output_type="TIME", adaptive=True, max_time_ms=4, num_bins=4, source="bar", frames=6, thr0=12, **U8)
    _run_case(2, 4, 3840, output_type="RATE", adaptive=True, source="random", frames=5, thr0=12, **U8)
    _run_case(2, 8, 1920, output_type="RATE", adaptive=True, inh=2, source="random", frames=6, thr0=12, **U8)
    _run_case(2, 16, 64, output_type="RATE", adaptive=True, frame_dtype="uint8", frames=5, thr0=12, **U8)
    _run_case(2, 16, 64, output_type="RATE", adaptive=True, frames=5, thr0=0, **U8)      # zero thresholds (B.2)
    _run_case(2, 16, 64, output_type="RATE", adaptive=True, frames=5, thr0=255, **U8)    # above max: clipped down
    _run_case(48, 32, 32, output_type="TIME_BIN_THR", adaptive=True, hw=0.99, frames=8, source="narrow", ref0=0,
              num_bins=6, threshold=12, thr0=12, **U8)                                  # BASELINE config 2 shape


def test_history_decay_time_binary_and_4x4():
    for hw in (0.0, 0.5, 0.99):
        _run_case(2, 16, 64, hw=hw, frames=5, source="narrow", **U8)
    _run_case(2, 8, 1920, hw=0.99, inh=2, frames=5, source="bar", **U8)
    _run_case(2, 24, 40, output_type="TIME_BIN", num_bins=8, frames=4, source="narrow", lut_bits=(3, 8), threshold=5, **U8)
    _run_case(1, 128, 128, output_type="TIME_BIN_THR", num_bins=6, inh=2, frames=6, source="bar", threshold=12, **U8)
    _run_case(2, 16, 256, inh=4, frames=5, source="bar", **U8)
    _run_case(1, 8, 4096, inh=4, frames=3, **U8)


def test_declined_tiles_and_the_generic_body_keep_byte_state():
    from pydvs_b200 import _lib
    _run_case(2, 16, 256, source="wide", frames=4, **U8)                # frames outside [0, 255]: redo list
    _run_case(2, 16, 256, source="wide", frames=4, inh=2, **U8)
    _run_case(2, 16, 256, output_type="RATE", adaptive=True, source="wide", frames=4, thr0=12, **U8)
    for kw in (dict(), dict(inh=2), dict(inh=3, W=18), dict(adaptive=True, thr0=12), dict(hw=1.5), dict(hw=-0.25)):
        W = kw.pop("W", 64)
        assert _run_case(2, 12, W, frames=4, force_generic=True, **U8, **kw).kernel_path == _lib.PATH_GENERIC
    _run_case(2, 33, 17, frames=4, **U8)                                # scalar loads / stores of the byte planes
    _run_case(1, 10, 14, inh=2, frames=4, **U8)


def test_state_dtype_selection_and_accessors():
    from pydvs_b200 import DVSConfig, DVSStreams, synth
    cfg = DVSConfig(width=64, height=16, streams=2, output_type="RATE", threshold=12, inh_area_width=2)
    auto = DVSStreams(cfg, state_dtype="auto")
    assert auto.state_dtype == "uint8" and auto.ref.dtype == torch.int16 and auto.thr is None
    assert DVSStreams(cfg).state_dtype == "uint8"                       # the default: bytes wherever that is exact
    assert DVSStreams(DVSConfig(width=64, height=16, streams=1, ref0=300), state_dtype="auto").state_dtype == "int16"
    assert DVSStreams(cfg, state_dtype="auto", debug_planes=True).state_dtype == "int16"
    unc = DVSConfig(width=64, height=16, streams=1, adaptive_threshold=True, uncoded=True)
    assert DVSStreams(unc, state_dtype="auto").state_dtype == "int16"   # the uncoded rule keeps unclipped thresholds
    with pytest.raises(ValueError):
        DVSStreams(unc, state_dtype="uint8")
    with pytest.raises(ValueError):
        DVSStreams(DVSConfig(width=64, height=16, streams=1, ref0=-1), state_dtype="uint8")
    # both layouts walk the same states; set_state() writes either of them
    wide = DVSStreams(cfg, state_dtype="int16")
    assert wide.ref.data_ptr() == wide._ref.data_ptr()                  # int16 state: `ref` IS the resident tensor
    start = torch.randint(0, 256, (2, 16, 64), dtype=torch.int16, device="cuda")
    auto.set_state(ref=start)
    wide.set_state(ref=start)
    for t in range(5):
        f = synth.random_frames(2, 16, 64, t, seed=3)
        a, b = auto.step(f), wide.step(f)
        assert torch.equal(a.events_raw[:a.total_events()], b.events_raw[:b.total_events()])
        assert torch.equal(auto.ref, wide.ref) and torch.equal(auto.ref_plane(1), wide.ref_plane(1))
    with pytest.raises(ValueError):
        auto.set_state(ref=start + 1000)
    assert auto.thr_plane(0) is None


def test_byte_state_under_capture_replay_pipelined_expand_and_padded_rows():
    from pydvs_b200 import DVSConfig, DVSStreams, synth
    # capture() / replay() on byte state: the warm-up snapshot restores the byte planes
    cfg = DVSConfig(width=128, height=128, streams=1, output_type="RATE", inh_area_width=2)
    eager, graphed = DVSStreams(cfg, graph=False), DVSStreams(cfg, graph=False, state_dtype="uint8")
    static = synth.moving_bar(1, 128, 128, 0, noise=6, seed=1)
    graphed.capture(static)
    assert graphed.frames_done == 0 and int((graphed.ref != 128).sum()) == 0
    for t in range(6):
        f = synth.moving_bar(1, 128, 128, t, noise=6, seed=1)
        re = eager.step(f)
        static.copy_(f)
        rg = graphed.replay()
        assert torch.equal(eager.ref, graphed.ref)
        assert torch.equal(re.seg_offsets, rg.seg_offsets) and torch.equal(re.events(), rg.events())
    # expand on the side stream (pipeline_depth 2), rows padded to a multiple of 8 pixels, uint8 frames handed to an
    # int16 configuration: all at once against the plain int16-state object
    cfg = DVSConfig(width=346, height=20, streams=3, output_type="TIME", num_bins=8, max_time_ms=8, inh_area_width=2)
    plain = DVSStreams(cfg, graph=False, event_capacity=3 * 20 * 346)
    piped = DVSStreams(cfg, state_dtype="uint8", pipeline_depth=2, event_capacity=3 * 20 * 346)
    assert piped.Wp == 352 and plain.Wp == 352
    for t in range(7):
        f = synth.random_frames(3, 20, 346, t, seed=5)
        a = plain.step(f)
        b = piped.step(f.to(torch.uint8) if t % 2 else f)
        piped.drain()
        torch.cuda.synchronize()
        n = a.total_events()
        assert n == b.total_events() and torch.equal(a.events_raw[:n], b.events_raw[:n])
        assert torch.equal(plain.ref, piped.ref)
