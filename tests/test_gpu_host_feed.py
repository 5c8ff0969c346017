"""HostFeed: pinned host frames -> device -> DVSStreams.step, two frame sets ahead, with and without the exact
narrowing of int16 planes to bytes on the host.  Every mode must produce the events and state of stepping the same
frames directly from device memory; a frame set with a value outside [0, 255] falls back to the int16 copy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    from pydvs_b200 import DVSConfig
    base = dict(width=256, height=64, streams=3, output_type="RATE", threshold=12, inh_area_width=2)
    base.update(kw)
    return DVSConfig(**base)


@pytest.mark.parametrize("narrow", [True, False, "auto"])
@pytest.mark.parametrize("state_dtype", ["int16", "uint8"])
def test_host_feed_matches_device_resident_steps(narrow, state_dtype):
    from pydvs_b200 import DVSStreams, HostFeed, synth
    cfg = _cfg()
    direct = DVSStreams(cfg, state_dtype=state_dtype, graph=False)
    fed = DVSStreams(cfg, state_dtype=state_dtype, graph=False)
    feed = HostFeed(fed, narrow=narrow, threads=3)
    frames = [synth.moving_bar(3, 64, 256, t, noise=6, seed=2) for t in range(7)]
    frames[4][1, 10, 20] = 300 if narrow is not False else 200      # one frame set that cannot be narrowed
    frames[5][2, 3, 7] = -5 if narrow is not False else 5
    host = [f.cpu().pin_memory() for f in frames]
    tickets = [feed.submit(host[0]), feed.submit(host[1])]
    with pytest.raises(RuntimeError):
        feed.submit(host[2])                                         # at most two sets in flight
    for t in range(7):
        a = direct.step(frames[t])
        b = feed.step(tickets[t])
        if t + 2 < 7:
            tickets.append(feed.submit(host[t + 2]))
        torch.cuda.synchronize()
        n = a.total_events()
        assert n == b.total_events() and torch.equal(a.seg_offsets, b.seg_offsets)
        assert torch.equal(a.events_raw[:n], b.events_raw[:n])
        assert torch.equal(direct.ref, fed.ref)
    feed.close()
    fed.check_status()
    st = feed.stats
    if narrow is True:
        assert st["narrowed"] == 5 and st["out_of_range"] == 2 and st["plain"] == 2
    elif narrow is False:
        assert st["narrowed"] == 0 and st["plain"] == 7
    else:
        assert st["calibration"] is not None and st["narrowed"] + st["plain"] == 7


def test_host_feed_takes_uint8_host_frames_and_adaptive_thresholds():
    from pydvs_b200 import DVSStreams, HostFeed, synth
    cfg = _cfg(output_type="TIME", adaptive_threshold=True, num_bins=4, max_time_ms=4, threshold0=12, inh_area_width=0)
    cap = 3 * 64 * 256 * 2
    direct, fed = DVSStreams(cfg, state_dtype="auto", event_capacity=cap), DVSStreams(cfg, state_dtype="auto", event_capacity=cap)
    feed = HostFeed(fed, narrow=True)
    for t in range(5):
        f = synth.moving_bar(3, 64, 256, t, noise=6, seed=4)
        host = (f.to(torch.uint8) if t % 2 else f).cpu().pin_memory()     # bytes from a camera, or the API's int16
        a, b = direct.step(f), feed.step(feed.submit(host))
        torch.cuda.synchronize()
        n = a.total_events()
        assert n == b.total_events() and torch.equal(a.events_raw[:n], b.events_raw[:n])
        assert torch.equal(direct.ref, fed.ref) and torch.equal(direct.thr, fed.thr)
    feed.close()


def test_async_key_sink_delivers_every_step_in_order():
    """AsyncKeySink: keys + CSR row pointers of each step reach pinned host memory one push behind, identical to the
    synchronous path (StepResult.keys / offsets_host) -- including a step that emits more than the sized copy."""
    from pydvs_b200 import DVSStreams, synth
    from pydvs_b200.sinks import AsyncKeySink
    cfg = _cfg(width=64, height=32, streams=5)
    dvs = DVSStreams(cfg, event_capacity=5 * 32 * 64 * 8)
    sink = AsyncKeySink(dvs, 12, 6, 63, depth=2, initial_keys=16)     # tiny first guess: the top-up copy runs
    want = []
    got = []
    for t in range(9):
        f = synth.random_frames(5, 32, 64, t, seed=t % 3) if t != 4 else synth.moving_bar(5, 32, 64, t, noise=0)
        res = dvs.step(f)
        want.append((res.offsets_host().copy(), res.keys(12, 6, 63).cpu().numpy().copy()))
        sink.push(res)
        if t >= 1:
            off, keys = sink.pop()
            got.append((off.copy(), keys.copy()))
    with pytest.raises(RuntimeError):
        sink.push(res); sink.push(res); sink.push(res)
    off, keys = sink.pop()
    got.append((off.copy(), keys.copy()))
    assert len(got) >= 9
    for (wo, wk), (go, gk) in zip(want, got):
        assert np.array_equal(wo, go) and np.array_equal(wk, gk)
