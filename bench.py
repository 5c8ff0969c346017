#!/usr/bin/env python
"""bench.py -- throughput of the spike-generation hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (CUDA)
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K ...     # one rank per GPU
    python bench.py --impl reference ...                                # the reference's CPU path

One "step" = one frame for every stream: fused tile pass (threshold -> inhibition k=2 -> reference
update -> record compaction -> slot counts), scans, AER expand.  Headline workload = BASELINE.json
configs[4]: 3840x2160 synthetic video, full pipeline, rate coded, 256 independent streams IN TOTAL, split
over the GPUs in contiguous blocks (`"scaling": "strong"`: 256 on one GPU, 32 per GPU on eight; nothing but
a final event-count all-gather crosses GPUs).  `--scaling weak` keeps 256 streams per GPU instead.

Prints ONE JSON line on rank 0:
  value          stream-frames/s with the frames resident in HBM (CUDA events, max over ranks)
  e2e            the same through the public API with int16 HOST frames (the reference API's dtype): pinned
                 H2D of every frame + D2H of the AER result inside the timed region; `e2e.uint8_host_frames`
                 is the camera-native variant, `e2e.h2d_peak_gbs` the plain pinned-copy ceiling of the same run
  roofline       the fused tile pass against the measured HBM peak
  configs        the other BASELINE configs (cfg 5 dense, cfg 4, cfg 2, cfg 3, cfg 1), uint8 frames, a 16-frame
                 ring and the float DVSOperator mode, each timed the same way (N = 1 only)
  parity_check   stream 0 and one other stream of the SAME frame sequence re-run through the CPU oracle after
                 the timed loop: reference plane (+ thresholds) and the full AER list of the last step must be
                 identical; the process exits 1 otherwise
  cpu_baseline   the reference's Cython path on one host core, on a bounded sample, from the same state
The reference arm times the compiled, unmodified Cython reference (oracle/_ref; the C port if it is absent),
one worker process per stream on all host cores, on the same frames, the same settle / warm-up / timed phases
and the same `config`; it never imports pydvs_b200.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # S = streams in total (BASELINE.json); encoder threshold = 12 (min_threshold 6 when adaptive)
    "4k_full_pipeline_rate": dict(W=3840, H=2160, S=256, output_type="RATE", inh=2, adaptive=False, max_time_ms=8,
                                  num_bins=None, hw=1.0, ref0=128),
    "4k_rate_no_inhibition": dict(W=3840, H=2160, S=256, output_type="RATE", inh=0, adaptive=False, max_time_ms=8,
                                  num_bins=None, hw=1.0, ref0=128),
    "4k_rate_inhibition_4x4": dict(W=3840, H=2160, S=256, output_type="RATE", inh=4, adaptive=False, max_time_ms=8,
                                   num_bins=None, hw=1.0, ref0=128),
    "davis346_rate_inhibition": dict(W=346, H=260, S=1024, output_type="RATE", inh=2, adaptive=False, max_time_ms=8,
                                     num_bins=None, hw=1.0, ref0=128),
    "davis346_rate": dict(W=346, H=260, S=1024, output_type="RATE", inh=0, adaptive=False, max_time_ms=8, num_bins=None,
                          hw=1.0, ref0=128),
    "1080p_time_adaptive": dict(W=1920, H=1080, S=64, output_type="TIME", inh=0, adaptive=True, max_time_ms=4,
                                num_bins=4, hw=1.0, ref0=128),
    "vga_rate_inhibition": dict(W=640, H=480, S=1, output_type="RATE", inh=2, adaptive=False, max_time_ms=8,
                                num_bins=None, hw=1.0, ref0=128),
    "mnist_time_bin_thr_adaptive": dict(W=32, H=32, S=4096, output_type="TIME_BIN_THR", inh=0, adaptive=True,
                                        max_time_ms=8, num_bins=6, hw=0.99, ref0=0),
    "128_rate": dict(W=128, H=128, S=1, output_type="RATE", inh=0, adaptive=False, max_time_ms=8, num_bins=None,
                     hw=1.0, ref0=128),
}

# the other BASELINE configs, reported under "configs" (N = 1): name -> overrides
SUITE = [
    ("cfg5_4k_dense", dict(workload="4k_full_pipeline_rate", pattern="dense", settle=2)),
    ("cfg5_4k_uint8_frames", dict(workload="4k_full_pipeline_rate", frame_dtype="uint8", reuse_headline_ring=True)),
    ("cfg5_4k_int16_state", dict(workload="4k_full_pipeline_rate", state_dtype="int16", reuse_headline_ring=True)),  # rounds 1-2 layout
    ("cfg5_4k_ring4_periodic_orbit", dict(workload="4k_full_pipeline_rate", ring=4)),   # round 1's headline workload
    ("cfg4_1080p_time_adaptive_64_streams", dict(workload="1080p_time_adaptive", settle=8)),
    ("cfg4_1080p_time_adaptive_64_streams_int16_state", dict(workload="1080p_time_adaptive", settle=8, state_dtype="int16")),
    ("cfg4_1080p_time_adaptive_8_streams", dict(workload="1080p_time_adaptive", streams=8, settle=8)),
    ("cfg2_mnist_shape_4096_streams_time_bin_thr_adaptive", dict(workload="mnist_time_bin_thr_adaptive", settle=4)),
    ("cfg3_vga_rate_inhibition", dict(workload="vga_rate_inhibition")),
    ("cfg1_128_rate", dict(workload="128_rate", settle=4)),
    # a real sensor shape whose rows are not a multiple of 8 pixels, with inhibition: planes kept with padded rows
    ("davis346_rate_inhibition_1024_streams", dict(workload="davis346_rate_inhibition")),
]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="4k_full_pipeline_rate", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=None, help="streams in total (default: the workload's)")
    ap.add_argument("--pattern", default="bar", choices=["bar", "dense"],
                    help="bar: moving bar + noise (sparse events); dense: uniform random pixels")
    ap.add_argument("--frame-dtype", default="int16", choices=["int16", "uint8"])
    ap.add_argument("--state-dtype", default="auto", choices=["auto", "int16", "uint8"],
                    help="layout of the resident reference / threshold planes: int16 as the reference's API holds them, or "
                         "one byte per pixel where that is exact (auto: uint8 whenever DVSStreams accepts it)")
    ap.add_argument("--ring", type=int, default=16,
                    help="distinct frames resident per stream.  The bar travels W/64 px per frame; a 4-frame ring settles on a "
                         "period-4 orbit of the state that emits ~60x fewer events than video does, so the default keeps 16")
    ap.add_argument("--e2e-steps", type=int, default=16)
    ap.add_argument("--host-narrow", default="auto", choices=["auto", "on", "off"],
                    help="e2e: narrow int16 host frames to bytes on the host cores before the copy (exact; HostFeed); auto = "
                         "time both ways once and keep the faster")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--configs", default="all",
                    help="all | none | comma-separated names of the secondary configs to time (N = 1)")
    ap.add_argument("--config-steps", type=int, default=20)
    ap.add_argument("--cpu-frames", type=int, default=3)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--no-weak-leg", action="store_true", help="N > 1: skip the secondary weak-scaling measurement")
    ap.add_argument("--pipeline-depth", type=int, default=1, choices=[1, 2])
    ap.add_argument("--k1-variant", default="auto", choices=["auto", "plain", "staged", "generic"],
                    help="generic: force the generic tile body (what shapes outside the fast paths run)")
    ap.add_argument("--tile-rows", type=int, default=None, help="rows per tile (default: dvs_choose_tile_rows)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="CUDA-graph replay of the step (auto: DVSStreams decides by problem size)")
    ap.add_argument("--settle", type=int, default=24,
                    help="untimed steps before the warm-up: with inhibition only one pixel of each 2x2 area moves "
                         "its reference per frame, so the dense start-up transient (ref0 = 128 vs background 64) "
                         "takes ~16 frames to die out")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# frames: counter-based, so the GPU arm (torch, pydvs_b200.synth.hashed_frames) and the reference arm
# (NumPy, below) see the same pixels without sharing any code path through the product
# ------------------------------------------------------------------------------------------------
def np_hashed_frame(stream, H, W, t, pattern="bar", noise=8, bg=64, bar=200):
    """NumPy twin of pydvs_b200.synth.hashed_frames for ONE stream (int16 [H, W]); equality is a CPU test."""
    import numpy as np
    P = H * W
    M = 0xFFFFFFFF
    q = np.arange(P, dtype=np.int64)
    h = (((np.int64(stream) * 1000003 + t) * 2654435761) + q * 40503) & M
    h ^= h >> 16
    h = (h * 0x7FEB352D) & M
    h ^= h >> 15
    h = (h * 0x2C1B3C6D) & M
    h ^= h >> 16
    if pattern == "dense":
        return (h & 0xFF).astype(np.int16).reshape(H, W)
    bw, sp = max(1, W // 16), max(1, W // 64)
    x = q % W
    x0 = ((stream * 37) % W + t * sp) % W
    base = np.where(((x - x0) % W) < bw, bar, bg).astype(np.int64)
    if noise:
        base += (h % (2 * noise + 1)) - noise
    return np.clip(base, 0, 255).astype(np.int16).reshape(H, W)


def key_params(W, H):
    import math
    ds = int(math.ceil(math.log2(max(W, H, 2))))
    return min(2 * ds, 15), ds, min((1 << ds) - 1, 255)


def bench_config(args, wl, world):
    """The `config` object: identical for both arms (the reference arm states its sample in cpu_baseline)."""
    S_total = args.streams if args.streams is not None else wl["S"]
    if args.scaling == "weak":
        S_total *= world
    per_gpu_gb = (S_total // max(world, 1)) * wl["W"] * wl["H"] * 6 / 1e9
    return {"workload": args.workload, "width": wl["W"], "height": wl["H"], "streams_total": S_total,
            "coding": wl["output_type"], "inhibition": wl["inh"], "adaptive_threshold": wl["adaptive"],
            "pattern": args.pattern, "frame_dtype": args.frame_dtype, "ring_frames": args.ring,
            "settle_steps": args.settle, "ref0": wl["ref0"], "state_dtype": args.state_dtype,
            "l2": ("inputs larger than L2: every step reads another frame of the ring plus the state planes "
                   "(%.2f GB per GPU per step)" % per_gpu_gb) if per_gpu_gb * args.ring > 0.126 else
                  "working set fits in L2 (launch-bound configuration, not an HBM measurement)"}


# ------------------------------------------------------------------------------------------------
# clocks: sampled with NVML in a thread DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period, [], 0, None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                try:
                    self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        names = [n for bit, n in self.REASONS.items() if self.reasons & bit and n != "gpu_idle"]
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(s)}


def visible_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# CPU arms (the oracle is executed here and only here: as the checker, or as the thing timed as a BASELINE)
# ------------------------------------------------------------------------------------------------
def oracle_pipeline(wl, W, H):
    from oracle import oracle as orc
    enc = {"RATE": 0, "TIME": 1, "TIME_BIN": 2, "TIME_BIN_THR": 3}[wl["output_type"]]
    lut = orc.generate_log2_table(4, 8)[3] if "BIN" in wl["output_type"] else None
    fs, ds, dm = key_params(W, H)
    return orc.Pipeline(W, H, mode=enc, threshold=12, max_time_ms=wl["max_time_ms"], num_bins=wl["num_bins"],
                        history_weight=wl["hw"], inh_k=wl["inh"], adaptive=wl["adaptive"], lut=lut, flag_shift=fs,
                        data_shift=ds, data_mask=dm, ref0=wl["ref0"], thr0=12)


def reference_frame(gs, wl, curr, ref, thr, inh_coords, lut, keyp):
    """One frame with the reference's own call sequence (stand_alone.py:160-185 /
    external_dvs_emulator_device.py:549-683); returns the list of lists."""
    W, H = curr.shape[1], curr.shape[0]
    fs, ds, dm = keyp
    if wl["adaptive"]:
        diff, abs_diff, spikes = gs.thresholded_difference_adpt(curr, ref, thr, 6, 168)
    else:
        diff, abs_diff, spikes = gs.thresholded_difference(curr, ref, 12)
    if wl["inh"] > 1:
        spikes = gs.local_inhibition(spikes, abs_diff, inh_coords, W, H, wl["inh"])
    if wl["adaptive"]:
        ref[:], thr[:] = gs.update_reference_rate_adpt(abs_diff, spikes, ref, thr, 6, 168, 2, 12, wl["max_time_ms"], wl["hw"])
    elif wl["output_type"] in ("RATE", "TIME"):
        ref[:] = gs.update_reference_rate(abs_diff, spikes, ref, 12, wl["max_time_ms"], wl["hw"])
    elif wl["output_type"] == "TIME_BIN":
        ref[:] = gs.update_reference_time_binary_raw(abs_diff, spikes, ref, 12, wl["max_time_ms"], 2, wl["hw"], lut)
    else:
        ref[:] = gs.update_reference_time_binary_thresh(abs_diff, spikes, ref, 12, wl["max_time_ms"], 2, wl["hw"], lut)
    neg, pos, gmax = gs.split_spikes(spikes, abs_diff, 2)
    enc_thr = 6 if wl["adaptive"] else 12
    if wl["output_type"] == "RATE":
        return gs.make_spike_lists_rate(pos, neg, gmax, enc_thr, fs, ds, dm, wl["max_time_ms"])
    if wl["output_type"] == "TIME":
        return gs.make_spike_lists_time(pos, neg, gmax, fs, ds, dm, wl["num_bins"], wl["max_time_ms"], enc_thr, 168)
    if wl["output_type"] == "TIME_BIN":
        return gs.make_spike_lists_time_bin(pos, neg, gmax, fs, ds, dm, wl["max_time_ms"], enc_thr, 168, 8, lut)
    return gs.make_spike_lists_time_bin_thr(pos, neg, gmax, fs, ds, dm, wl["max_time_ms"], enc_thr, 168, wl["num_bins"], lut)


def load_reference():
    """The compiled reference modules (oracle/_ref) or None."""
    try:
        from oracle import ref_loader
        return ref_loader.load()
    except Exception:
        return None


_SHARED = {}


def _ref_worker(idx):
    """One worker process = one stream (the reference's parallelism model, stand_alone_threaded.py:236-247):
    the stream's own frames, the same settle / warm-up / timed phases as the GPU arm."""
    import numpy as np
    import warnings
    warnings.simplefilter("ignore")
    sh = _SHARED
    wl, rows, pattern = sh["wl"], sh["rows"], sh["pattern"]
    W, H = wl["W"], rows
    ring = [np_hashed_frame(idx, wl["H"], W, t, pattern)[:rows].copy() for t in range(sh["ring"])]
    keyp = key_params(W, wl["H"])
    mods = load_reference() if sh["kind"] == "reference" else None
    t_setup = time.perf_counter()
    pipe = oracle_pipeline(wl, W, H)
    i = 0
    for _ in range(sh["settle"]):   # untimed settle: the C port (bit-identical to the reference, tests/) advances the state
        pipe.step_raw(ring[i % len(ring)])
        i += 1
    n_events = 0
    if mods is None:
        for _ in range(sh["warmup"]):
            pipe.step_raw(ring[i % len(ring)])
            i += 1
        t0 = time.perf_counter()
        for _ in range(sh["steps"]):
            counts, _ev = pipe.step_raw(ring[i % len(ring)])
            n_events += int(counts.sum())
            i += 1
        t1 = time.perf_counter()
        return t1 - t0, t0 - t_setup, n_events
    gs = mods[0]
    inh_coords = gs.generate_inh_coords(W, H, wl["inh"]) if wl["inh"] > 1 else None
    lut = gs.generate_log2_table(4, 8)[3]
    ref, thr = pipe.ref.copy(), pipe.thr.copy()
    for _ in range(sh["warmup"]):
        reference_frame(gs, wl, ring[i % len(ring)], ref, thr, inh_coords, lut, keyp)
        i += 1
    t0 = time.perf_counter()
    for _ in range(sh["steps"]):
        lists = reference_frame(gs, wl, ring[i % len(ring)], ref, thr, inh_coords, lut, keyp)
        n_events += sum(len(x) for x in lists)
        i += 1
    t1 = time.perf_counter()
    return t1 - t0, t0 - t_setup, n_events


def run_reference_arm(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation with all the host cores, one process per stream."""
    if rank != 0:
        return
    import multiprocessing as mp
    import psutil
    kind = "reference" if load_reference() is not None else "port"
    W, H = wl["W"], wl["H"]
    # every worker holds ~20 frame-sized temporaries (+ ~160 B/pixel of Python tuples while the reference builds
    # inh_coords): bound the worker count by the host memory as well as by the cores
    per_worker = W * H * (200 if (kind == "reference" and wl["inh"] > 1) else 60) + (64 << 20)
    cores = max(1, min(os.cpu_count() or 1, 32, int(0.5 * psutil.virtual_memory().available / per_worker)))
    S_total = args.streams if args.streams is not None else wl["S"]
    cores = min(cores, S_total)
    # bounded sample: all K + W frames are run; if that would take too long the workers take the top rows of their
    # frames (a multiple of the inhibition width) instead of fewer steps
    per_frame_s = (2.5 if kind == "reference" else 0.2) * (W * H) / (3840 * 2160)
    rows = H
    while rows > 16 and (args.steps + args.warmup) * per_frame_s * rows / H > 150.0:
        rows = max(16, (rows // 2) // 4 * 4)
    _SHARED.update(kind=kind, wl=wl, rows=rows, pattern=args.pattern, ring=args.ring, settle=args.settle,
                   warmup=args.warmup, steps=args.steps)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, range(cores))
    wall = time.perf_counter() - t0
    loop = max(r[0] for r in res)     # slowest worker's timed loop
    frac = rows / H
    value = cores * args.steps * frac / loop
    line = {"impl": "reference", "metric": "stream_frames_per_s", "value": value, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int16",
            "data": "synthetic", "gpixel_per_s": value * W * H / 1e9,
            "config": bench_config(args, wl, world),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "%d worker processes = streams 0..%d of the workload, each its own frames; per "
                                       "step every worker advances its stream by one frame (rows 0..%d of %d); %d settle "
                                       "frames advanced untimed by the C port, %d warm-up + %d timed frames through the "
                                       "%s; events in the timed frames: %d; wall %.1f s"
                                       % (cores, cores - 1, rows - 1, H, args.settle, args.warmup, args.steps,
                                          "compiled Cython reference" if kind == "reference" else "C port",
                                          sum(r[2] for r in res), wall)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_single_core(wl, ring_host, ref_now, thr_now, next_i, n_frames, pattern):
    """The reference's path on ONE host core (it is single-threaded: serial prange, setup.py:13-29), continuing
    stream 0 from the state the GPU run left (same phase as the timed region)."""
    import numpy as np
    import warnings
    warnings.simplefilter("ignore")
    mods = load_reference()
    W, H = wl["W"], wl["H"]
    keyp = key_params(W, H)
    ref = ref_now.copy()
    thr = thr_now.copy() if thr_now is not None else np.full((H, W), 12, np.int16)
    n_ev = 0
    t_s = time.perf_counter()
    if mods is not None:
        gs = mods[0]
        inh_coords = gs.generate_inh_coords(W, H, wl["inh"]) if wl["inh"] > 1 else None
        lut = gs.generate_log2_table(4, 8)[3]
        t0 = time.perf_counter()
        for i in range(n_frames):
            lists = reference_frame(gs, wl, ring_host[(next_i + i) % len(ring_host)], ref, thr, inh_coords, lut, keyp)
            n_ev += sum(len(x) for x in lists)
        secs = time.perf_counter() - t0
        kind = "reference"
    else:
        pipe = oracle_pipeline(wl, W, H)
        pipe.ref[:], pipe.thr[:] = ref, thr
        t0 = time.perf_counter()
        for i in range(n_frames):
            counts, _ = pipe.step_raw(ring_host[(next_i + i) % len(ring_host)])
            n_ev += int(counts.sum())
        secs = time.perf_counter() - t0
        kind = "port"
    return {"value": n_frames / secs, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d consecutive %dx%d frames of stream 0 (%s pattern) continuing from the state after the timed "
                      "region, one core of %d; one-off setup (generate_inh_coords etc.) %.1f s excluded"
                      % (n_frames, W, H, pattern, os.cpu_count(), t0 - t_s),
            "mpixel_per_s": n_frames * W * H / secs / 1e6, "events": n_ev}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_ring(wl, S, ring, pattern, frame_dtype, device, stream_base):
    import torch
    from pydvs_b200 import synth
    dt = torch.uint8 if frame_dtype == "uint8" else torch.int16
    frames = []
    chunk = max(1, min(S, (1 << 25) // (wl["W"] * wl["H"])))  # bound the int64 temporaries of the generator
    for t in range(ring):
        buf = torch.empty((S, wl["H"], wl["W"]), dtype=dt, device=device)
        for s0 in range(0, S, chunk):
            n = min(chunk, S - s0)
            buf[s0:s0 + n] = synth.hashed_frames(n, wl["H"], wl["W"], t, pattern=pattern, device=device,
                                                 stream_offset=stream_base + s0, dtype=dt)
        frames.append(buf)
    return frames


class Case:
    """One timed configuration on this rank: state object + frame ring + step counter."""

    def __init__(self, args, wl, S, stream_base, device, pattern, frame_dtype, ring_n, graph, ring=None,
                 pipeline_depth=1, k1_variant="auto", state_dtype=None):
        import torch
        from pydvs_b200 import DVSConfig, DVSStreams
        self.wl, self.S, self.stream_base, self.device = wl, S, stream_base, device
        self.pattern, self.frame_dtype = pattern, frame_dtype
        self.W, self.H = wl["W"], wl["H"]
        self.P = self.W * self.H
        cfg = DVSConfig(width=self.W, height=self.H, streams=S, output_type=wl["output_type"], threshold=12,
                        adaptive_threshold=wl["adaptive"], history_weight=wl["hw"], max_time_ms=wl["max_time_ms"],
                        num_bins=wl["num_bins"], inh_area_width=wl["inh"], frame_dtype=frame_dtype,
                        threshold0=12 if wl["adaptive"] else None, ref0=wl["ref0"], tile_rows=getattr(args, "tile_rows", None))
        g = {"auto": "auto", "on": True, "off": False}[graph]
        self.dvs = DVSStreams(cfg, device=device, event_capacity=1 << 16, pipeline_depth=pipeline_depth,
                              force_generic={"auto": 0, "plain": 0, "staged": 3, "generic": 1}[k1_variant],
                              split_lists=False, graph=g,   # AER events are the output; split_spikes lists are not requested
                              state_dtype=state_dtype or getattr(args, "state_dtype", "auto"))
        self.ring = ring if ring is not None else make_ring(wl, S, ring_n, pattern, frame_dtype, device, stream_base)
        self.i = 0
        torch.cuda.synchronize()

    def frame(self):
        f = self.ring[self.i % len(self.ring)]
        self.i += 1
        return f

    def settle(self, n):
        """Untimed steps; event overflow during the start-up transient only drops events (the state advances
        regardless), so the event buffer is sized afterwards from the steady state."""
        import torch
        for _ in range(n):
            self.dvs.step(self.frame(), expand=False)
        torch.cuda.synchronize()
        # capacity policy: the worst case (one record per inhibition area, every slot it can reach) when that is
        # small, else 2x the largest step of one measured pass over the ring (timed() doubles it on overflow)
        d = self.dvs
        per_rec = {"RATE": max(d.n_slots - 1, 1), "TIME": 1}.get(self.wl["output_type"], 8)
        bound = self.S * (self.P // (d.inh_k * d.inh_k)) * per_rec
        peak = 0
        for _ in range(len(self.ring)):   # the whole ring: the frame on which the bar wraps around is the heaviest
            d.step(self.frame(), expand=False)
            peak = max(peak, int(d.seg_offsets[-1].item()))
        self.peak_events = peak
        d.reserve_events(bound if bound * 8 <= (2 << 30) else min(bound, int(peak * 1.5) + (1 << 16)))
        d.status.zero_()

    def launches_per_step(self):
        d = self.dvs
        # K1 (+ dvs_pad_rows in front of it when the planes are kept with padded rows), redo, scan, expand.  The scan is one
        # launch (the last CTA builds the row pointers, dvs_scan_tiles_ticket) up to 8192 (stream, slot, polarity) segments;
        # beyond that: k_scan_small + k_seg_fill for streams of at most 8 tiles, else the tile scan + three segment launches
        small = d.S * 2 * d.n_slots <= 8192
        scan = 1 if small else (2 if d.tps <= 8 else 4)
        return (1 if d.Wp != d.W else 0) + 2 + scan + 1

    def timed(self, K, warmup, barrier, sampler=None, split=True):
        """K steps bracketed by CUDA events on the launching stream.  split: per-launch events around K1."""
        import numpy as np
        import torch
        dvs = self.dvs
        if dvs.use_graph:   # every ring buffer gets its graph captured before the timed region
            warmup = max(warmup, len(self.ring) + 1)
        for _ in range(warmup):
            dvs.step(self.frame())
        torch.cuda.synchronize()
        try:
            dvs.check_status()
        except OverflowError:
            if int(os.environ.get("WORLD_SIZE", "1")) > 1:
                raise   # a retry on one rank would leave the others at the barrier
            dvs.reserve_events(2 * dvs.event_capacity)
            return self.timed(K, warmup, barrier, sampler, split)
        split = split and not dvs.use_graph
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)] if split else None
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.start()
        start.record()
        if split:
            for i in range(K):
                f = self.frame()
                dvs.begin_step()
                ev[i][0].record()
                dvs.tile_pass(f)
                ev[i][1].record()
                dvs.scan()
                dvs.expand_async()
                ev[i][2].record()
            dvs.drain()   # every side-stream expand finishes inside the timed region
        else:
            for i in range(K):
                dvs.step(self.frame())
        stop.record()
        torch.cuda.synchronize()
        barrier()
        out = {"elapsed_ms": start.elapsed_time(stop), "k1_ms": None, "rest_ms": None}
        if split:
            out["k1_ms"] = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
            out["rest_ms"] = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
        try:
            dvs.check_status()
        except OverflowError:   # events were dropped: the step did less work than it should have; measure again
            if int(os.environ.get("WORLD_SIZE", "1")) > 1:
                raise
            dvs.reserve_events(2 * dvs.event_capacity)
            return self.timed(K, warmup, barrier, None, split)
        out["events_last_step"] = int(dvs.seg_offsets[-1].item())
        return out

    def k1_only(self, n=10):
        """The tile pass alone (graph-replayed configs have no per-launch events inside the step)."""
        import numpy as np
        import torch
        dvs = self.dvs
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n)]
        for i in range(n):
            f = self.frame()
            dvs.begin_step()
            ev[i][0].record()
            dvs.tile_pass(f)
            ev[i][1].record()
            dvs.scan()
            dvs.expand_async()
        torch.cuda.synchronize()
        return float(np.median([e[0].elapsed_time(e[1]) for e in ev]))

    # -- parity: snapshot now, check later (the oracle runs in worker threads at the end of the run) --------
    def parity_snapshot(self, streams):
        """State and AER output of the LAST step for the given local streams, plus their ring frames on the host."""
        import torch
        dvs = self.dvs
        torch.cuda.synchronize()
        fs, ds, dm = key_params(self.W, self.H)
        from pydvs_b200.streams import StepResult
        res = StepResult(dvs, dvs.events, dvs.seg_offsets, dvs.status)
        off = res.offsets_host()
        per = 2 * dvs.n_slots
        jobs = []
        for s in streams:
            keys = res.keys(fs, ds, dm, stream=s).cpu().numpy()
            slot_off = off[s * per:(s + 1) * per + 1:2].copy() - off[s * per]
            jobs.append({"stream": self.stream_base + s, "wl": self.wl, "n_frames": self.i,
                         "ring": [f[s].cpu().numpy().astype("int16") for f in self.ring],
                         "ref": dvs.ref_plane(s).cpu().numpy(),
                         "thr": dvs.thr_plane(s).cpu().numpy() if dvs.thr_plane(s) is not None else None,
                         "keys": keys, "slot_counts": (slot_off[1:] - slot_off[:-1])})
        return jobs


def run_parity_job(job):
    """Re-run one stream's whole frame sequence through the CPU oracle and compare with the GPU snapshot."""
    import numpy as np
    wl, ring = job["wl"], job["ring"]
    H, W = ring[0].shape
    pipe = oracle_pipeline(wl, W, H)
    counts = flat = None
    for i in range(job["n_frames"]):
        counts, flat = pipe.step_raw(ring[i % len(ring)])
    ref_equal = bool(np.array_equal(pipe.ref, job["ref"]))
    thr_equal = True if job["thr"] is None else bool(np.array_equal(pipe.thr, job["thr"]))
    exp_keys = flat[:, 0].astype(np.int16)
    events_equal = bool(np.array_equal(np.asarray(counts, dtype=np.int64), np.asarray(job["slot_counts"], dtype=np.int64))
                        and np.array_equal(exp_keys, job["keys"]))
    return {"stream": job["stream"], "frames": job["n_frames"], "events_last_step": int(len(exp_keys)),
            "ref_equal": ref_equal, "thr_equal": thr_equal, "events_equal": events_equal}


def load_peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    return peak, src


def load_traffic(key):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def bytes_per_px(wl, frame_dtype):
    """frame read + int16 reference read and write (+ int16 threshold read and write)."""
    return (1 if frame_dtype == "uint8" else 2) + 4 + (4 if wl["adaptive"] else 0)


def layout_bytes_per_px(wl, frame_dtype, state_dtype):
    """The same count for the layout actually resident (one-byte state planes halve the state traffic)."""
    st = 1 if state_dtype == "uint8" else 2
    return (1 if frame_dtype == "uint8" else 2) + 2 * st + (2 * st if wl["adaptive"] else 0)


def pick_streams(S, name):
    """Stream 0 and one other, chosen by a hash of the configuration name."""
    import zlib
    if S <= 1:
        return [0]
    return [0, 1 + zlib.crc32(name.encode()) % (S - 1)]


def run_suite_entry(args, name, ov, device, barrier, headline_case, peak):
    """Time one secondary configuration (N = 1) the same way as the headline; returns (entry, parity jobs)."""
    import torch
    wl = WORKLOADS[ov["workload"]]
    S = ov.get("streams", wl["S"])
    pattern, fdt = ov.get("pattern", "bar"), ov.get("frame_dtype", "int16")
    ring_n = ov.get("ring", 4)
    ring = None
    if ov.get("reuse_headline_ring") and headline_case is not None and headline_case.S == S:
        dt = torch.uint8 if fdt == "uint8" else torch.int16
        ring = [f.to(dt) for f in headline_case.ring]
    case = Case(args, wl, S, 0, device, pattern, fdt, ring_n, args.graph, ring=ring, state_dtype=ov.get("state_dtype"))
    case.settle(ov.get("settle", args.settle))
    K = max(1, args.config_steps)
    t = case.timed(K, 3, barrier)
    jobs = [] if args.no_parity else case.parity_snapshot(pick_streams(S, name))
    k1_ms = t["k1_ms"] if t["k1_ms"] is not None else case.k1_only()
    ms = t["elapsed_ms"] / K
    bpp = bytes_per_px(wl, fdt)
    k1_bytes = bpp * S * case.P
    l2_resident = (bpp + 2) * S * case.P * len(case.ring) < 100e6
    sdt = case.dvs.state_dtype
    tr = load_traffic("%s:%s:%s%s" % (ov["workload"], pattern, fdt, ":s8" if sdt == "uint8" else ""))
    entry = {"workload": ov["workload"], "width": wl["W"], "height": wl["H"], "streams": S, "coding": wl["output_type"],
             "inhibition": wl["inh"], "adaptive_threshold": wl["adaptive"], "pattern": pattern, "frame_dtype": fdt,
             "ring_frames": len(case.ring), "steps": K, "ms_per_step": ms, "stream_frames_per_s": S * K / (t["elapsed_ms"] / 1e3),
             "gpixel_per_s": S * case.P * K / (t["elapsed_ms"] / 1e3) / 1e9, "k1_ms": k1_ms, "bytes_per_px": bpp,
             "achieved_gbs": k1_bytes / (k1_ms / 1e3) / 1e9, "frac": k1_bytes / (k1_ms / 1e3) / 1e9 / peak,
             "step_frac": (k1_bytes + 8 * t["events_last_step"]) / (ms / 1e3) / 1e9 / peak,
             "traffic": (tr["bytes_per_px"] * S * case.P) if tr else None,
             "cuda_graph": bool(case.dvs.use_graph), "launches_per_step": case.launches_per_step(),
             "events_last_step": t["events_last_step"], "tile_rows": case.dvs.tile_rows, "state_dtype": sdt,
             "layout_bytes_per_px": layout_bytes_per_px(wl, fdt, sdt),
             "frac_of_layout_bytes": layout_bytes_per_px(wl, fdt, sdt) * S * case.P / (k1_ms / 1e3) / 1e9 / peak,
             # (bytes/px * pixels + 8 * events) / time -- K1 writes one 8-byte record per spike
             "frac_with_events": (k1_bytes + 8 * t["events_last_step"]) / (k1_ms / 1e3) / 1e9 / peak}
    if l2_resident:
        entry["l2_resident"] = "ring + state fit in the 126 MB L2: launch-bound, `frac` is not an HBM fraction"
    del case
    torch.cuda.empty_cache()
    return entry, jobs


def run_float_entry(args, device, peak):
    """The float32 DVSOperator mode (pydvs/cpp/dvs_op.cpp:4-33): src 4 R + ref 4 RW + thr 4 RW + diff 4 W = 24 B/px
    with the diff plane the C++ object keeps (dvs_emu.hpp:89-108); parity of plane 0 against the oracle."""
    import numpy as np
    import torch
    from pydvs_b200 import DVSOperatorF32, synth
    S, H, W = 64, 1080, 1920
    op = DVSOperatorF32(H, W, streams=S, device=device)
    ring = [synth.hashed_frames(S, H, W, t, device=device).to(torch.float32) for t in range(4)]
    K = max(1, args.config_steps)
    n = 0
    for _ in range(3):
        op.update(ring[n % 4]); n += 1
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for _ in range(K):
        op.update(ring[n % 4]); n += 1
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / K
    entry = {"workload": "float_dvs_operator", "width": W, "height": H, "streams": S, "dtype": "f32", "steps": K,
             "ms_per_step": ms, "k1_ms": ms, "bytes_per_px": 24, "stream_frames_per_s": S / (ms / 1e3),
             "gpixel_per_s": S * H * W / (ms / 1e3) / 1e9, "achieved_gbs": 24 * S * H * W / (ms / 1e3) / 1e9,
             "frac": 24 * S * H * W / (ms / 1e3) / 1e9 / peak, "launches_per_step": 1, "cuda_graph": False}
    if not args.no_parity:
        from oracle import oracle as orc
        ref = np.zeros((H, W), np.float32)
        thr = np.full((H, W), 10.0, np.float32)
        diff = np.zeros((H, W), np.float32)
        host = [f[0].cpu().numpy() for f in ring]
        for i in range(n):
            orc.dvs_op_f32(host[i % 4], diff, ref, thr, 0.999, 1.5, 0.99)
        entry["parity"] = {"frames": n, "ref_equal": bool(np.array_equal(op.ref[0].cpu().numpy(), ref)),
                           "thr_equal": bool(np.array_equal(op.thr[0].cpu().numpy(), thr)),
                           "events_equal": bool(np.array_equal(op.diff[0].cpu().numpy(), diff))}
    del op, ring
    torch.cuda.empty_cache()
    return entry


def run_pipeline_entry(args, device, peak):
    """BASELINE config 2 end to end on the device: the reference's MNIST fixtures (28x28 padded to 32x32, from
    tests/golden/golden_vcam_v1.npz) shown by 4096 batched VirtualCam cameras (micro-saccades + fade mask) -> fused
    time-binary step with adaptive thresholds -> sink (16-bit keys + CSR row pointers of every stream copied to the host).
    Parity: two cameras re-run on the CPU (pinned VirtualCam restatement -> oracle pipeline)."""
    import numpy as np
    import torch
    from pydvs_b200 import DVSConfig, DVSStreams
    from pydvs_b200.virtual_cam import VirtualCam
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_vcam_v1.npz"))
    imgs, mask = g["images"].astype(np.int16), g["mask"]
    wl = WORKLOADS["mnist_time_bin_thr_adaptive"]
    S, H, W = 4096, 32, 32
    start = np.arange(S) % len(imgs)
    kw = dict(fps=1000, image_on_time_ms=10, inter_off_time_ms=1, max_saccade_distance=1, frames_per_microsaccade=1,
              max_cycles=10 ** 6, background_gray=0)
    cam = VirtualCam(torch.from_numpy(imgs).to(device), behaviour="SACCADE", fade_mask=mask, streams=S, start_index=start,
                     seed=7, **kw)
    cfg = DVSConfig(width=W, height=H, streams=S, output_type="TIME_BIN_THR", threshold=12, adaptive_threshold=True,
                    history_weight=wl["hw"], max_time_ms=wl["max_time_ms"], num_bins=wl["num_bins"], threshold0=12, ref0=0)
    dvs = DVSStreams(cfg, device=device, event_capacity=S * H * W * 4, split_lists=False, state_dtype=args.state_dtype)
    fs, ds, dm = key_params(W, H)
    frames_host = {0: [], 1234: []}

    from pydvs_b200.sinks import AsyncKeySink
    asink = AsyncKeySink(dvs, fs, ds, dm, depth=2, initial_keys=1 << 20)

    def one(keep, sink):
        ok, batch = cam.read()
        if keep:   # (kept on the device: a download here would synchronise every frame; converted after the timing)
            for s in frames_host:
                frames_host[s].append(batch[s].clone())
        res = dvs.step(batch)
        if sink == "async":   # keys + CSR of frame n land in pinned host memory while frame n + 1 is computed
            asink.push(res)
            return asink.pop() if asink._pushed - asink._popped == 2 else None
        if sink:
            return res.offsets_host(), res.keys(fs, ds, dm).cpu().numpy()
        return None

    warm, K = 24, max(1, args.config_steps)
    for _ in range(warm):
        one(not args.no_parity, False)
    torch.cuda.synchronize()
    out = {}
    for sink in (False, "async", True):
        t0 = time.perf_counter()
        for _ in range(K):
            last = one(not args.no_parity, sink)
        if sink == "async":
            last_async = asink.pop()          # the frame still in flight
            last_async = (last_async[0].copy(), last_async[1].copy())
        torch.cuda.synchronize()
        out[sink] = (time.perf_counter() - t0) / K
    dvs.check_status()
    frames_host = {s: [f.cpu().numpy() for f in frs] for s, frs in frames_host.items()}
    entry = {"workload": "mnist_virtualcam_pipeline", "width": W, "height": H, "streams": S, "coding": "TIME_BIN_THR",
             "adaptive_threshold": True, "source": "VirtualCam SACCADE + fade mask on the reference's mnist/t10k fixtures",
             "steps": K, "ms_per_step": out[False] * 1e3, "stream_frames_per_s": S / out[False],
             "ms_per_step_with_sink": out["async"] * 1e3, "stream_frames_per_s_with_sink": S / out["async"],
             "ms_per_step_with_synchronous_sink": out[True] * 1e3,
             "sink": "int16 keys + CSR row pointers of all streams copied to pinned host memory every frame "
                     "(pydvs_b200.sinks.AsyncKeySink: one frame behind the device; the synchronous variant waits per frame)",
             "events_last_step": int(last[0][-1]), "timing": "wall clock around source + step (+ sink), synchronised at both ends",
             "cuda_graph": bool(dvs.use_graph)}
    if not args.no_parity:
        from oracle import oracle as orc
        from oracle.virtual_cam_oracle import ReplayDraws, VirtualCamOracle
        lut = orc.generate_log2_table(4, 8)[3]
        replay = ReplayDraws(7, S)
        res_ok = []
        off, keys = last
        per = 2 * dvs.n_slots
        for s, frs in frames_host.items():
            ocam = VirtualCamOracle(list(imgs), "SACCADE", fade_mask=mask, start_index=int(start[s]), rng=replay.stream(s),
                                    frames_per_saccade=29, **kw)
            pipe = orc.Pipeline(W, H, mode=orc.ENC_TIME_BIN_THR, threshold=12, max_time_ms=wl["max_time_ms"],
                                num_bins=wl["num_bins"], history_weight=wl["hw"], adaptive=True, lut=lut, flag_shift=fs,
                                data_shift=ds, data_mask=dm, ref0=0, thr0=12)
            frames_equal = True
            for fr in frs:
                _ok, img = ocam.read()
                frames_equal = frames_equal and np.array_equal(img, fr)
                counts, flat = pipe.step_raw(img)
            got = keys[int(off[s * per]):int(off[(s + 1) * per])]
            res_ok.append({"stream": s, "frames_equal": bool(frames_equal),
                           "ref_equal": bool(np.array_equal(dvs.ref_plane(s).cpu().numpy(), pipe.ref)),
                           "thr_equal": bool(np.array_equal(dvs.thr_plane(s).cpu().numpy(), pipe.thr)),
                           "events_equal": bool(np.array_equal(got, flat[:, 0].astype(np.int16)))})
        entry["parity"] = {"streams": [r["stream"] for r in res_ok], "frames": len(frames_host[0]),
                           "frames_equal": all(r["frames_equal"] for r in res_ok), "ref_equal": all(r["ref_equal"] for r in res_ok),
                           "thr_equal": all(r["thr_equal"] for r in res_ok), "events_equal": all(r["events_equal"] for r in res_ok)}
    del cam, dvs
    torch.cuda.empty_cache()
    return entry


def run_ours(args, wl, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    from pydvs_b200 import sharding

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cfgd = bench_config(args, wl, world)
    S_total = cfgd["streams_total"]
    stream_base, S = sharding.shard_streams(S_total, world, rank)
    W, H = wl["W"], wl["H"]
    P = W * H
    peak, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pool = ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1)))
    pending = []   # (label, future)

    # ---- headline ------------------------------------------------------------------------------------
    case = Case(args, wl, S, stream_base, device, args.pattern, args.frame_dtype, args.ring,
                "off" if args.pipeline_depth > 1 else args.graph, pipeline_depth=args.pipeline_depth,
                k1_variant=args.k1_variant)
    case.settle(args.settle)
    K, Wm = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(visible_gpu_index(local_rank))
    t = case.timed(K, Wm, barrier, sampler=sampler)
    clocks = sampler.stop()
    elapsed_ms = allmax(t["elapsed_ms"])
    k1_ms = t["k1_ms"] if t["k1_ms"] is not None else case.k1_only()
    k1_ms = allmax(k1_ms)
    rest_ms = t["rest_ms"]
    if rest_ms is None:   # graph replay: no events inside the step; the scan + expand share is what the tile pass leaves
        rest_ms = max(0.0, elapsed_ms / K - k1_ms)
    events_last_step = t["events_last_step"]
    dvs = case.dvs
    if rank == 0 and not args.no_parity:
        for job in case.parity_snapshot(pick_streams(S, "headline")):
            pending.append(("headline", pool.submit(run_parity_job, job)))
    frames_done_timed = case.i
    lps = case.launches_per_step()

    # final event-count all-gather: the only collective of the path (NCCL over NVLink)
    counts = dvs.stream_totals.view(S, dvs.C)[:, 2:].sum(dim=1).to(torch.int64)
    all_counts = sharding.gather_event_counts(counts, total_streams=S_total) if world > 1 else counts
    value = S_total * K / (elapsed_ms / 1e3)

    # state of stream 0 for the single-core CPU baseline (same phase as the timed region)
    cpu_state = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_state = ([f[0].cpu().numpy().astype(np.int16) for f in case.ring], dvs.ref_plane(0).cpu().numpy(),
                     dvs.thr_plane(0).cpu().numpy() if dvs.thr_plane(0) is not None else None, case.i)

    # ---- e2e: host frames in pinned memory, H2D + step + D2H of the AER result per step ----------------
    e2e = None
    if not args.no_e2e:
        pool_pin = {}   # one pinned allocation serves both legs (page-locking runs at ~2 GB/s)
        e2e = run_e2e(args, case, world, barrier, allmax, "int16" if args.frame_dtype == "int16" else "uint8", pool_pin,
                      narrow=args.host_narrow)
        if args.frame_dtype == "int16":
            if args.host_narrow is not False:   # the straight copy of the int16 planes, for comparison
                e2e["plain_int16_copy"] = run_e2e(args, case, world, barrier, allmax, "int16", pool_pin, narrow=False)
            from pydvs_b200 import DVSConfig, DVSStreams
            cfg2 = DVSConfig(**{**dvs.cfg.__dict__, "frame_dtype": "uint8"})
            case2 = Case.__new__(Case)
            case2.__dict__.update(case.__dict__)
            case2.frame_dtype = "uint8"
            case2.dvs = DVSStreams(cfg2, device=device, event_capacity=dvs.event_capacity, split_lists=False, graph=False,
                                   state_dtype=dvs.state_dtype)
            case2.dvs.set_state(ref=dvs._ref)   # continue from the same state and position in the ring
            case2.ring = [f.to(torch.uint8) for f in case.ring]
            e2e["uint8_host_frames"] = run_e2e(args, case2, world, barrier, allmax, "uint8", pool_pin)
            del case2, pool_pin
            torch.cuda.empty_cache()

    # ---- N > 1: the weak-scaling leg (256 streams per GPU) as a secondary measurement ------------------
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_leg:
        del case, dvs
        torch.cuda.empty_cache()
        S_w = args.streams if args.streams is not None else wl["S"]
        wcase = Case(args, wl, S_w, rank * S_w, device, args.pattern, args.frame_dtype, args.ring, args.graph)
        wcase.settle(args.settle)
        tw = wcase.timed(K, Wm, barrier)
        w_ms = allmax(tw["elapsed_ms"])
        weak = {"scaling": "weak", "streams_per_gpu": S_w, "streams_total": S_w * world, "ms_per_step": w_ms / K,
                "value": S_w * world * K / (w_ms / 1e3), "unit": "frames/s",
                "gpixel_per_s": S_w * world * K / (w_ms / 1e3) * P / 1e9}
        case = wcase
        dvs = wcase.dvs

    # ---- the other BASELINE configs (N = 1) -----------------------------------------------------------
    configs = {}
    if world == 1 and args.configs != "none":
        want = None if args.configs == "all" else set(args.configs.split(","))
        headline_for_reuse = case if (args.pattern == "bar" and args.frame_dtype == "int16") else None
        for name, ov in SUITE:
            if want is not None and name not in want:
                continue
            try:
                entry, jobs = run_suite_entry(args, name, ov, device, barrier, headline_for_reuse, peak)
                configs[name] = entry
                for job in jobs:
                    pending.append((name, pool.submit(run_parity_job, job)))
            except Exception as exc:   # never lose the headline to a secondary configuration
                configs[name] = {"error": repr(exc)}
        if want is None or "cfg2_pipeline_virtualcam_4096_streams" in want:
            try:
                configs["cfg2_pipeline_virtualcam_4096_streams"] = run_pipeline_entry(args, device, peak)
            except Exception as exc:
                configs["cfg2_pipeline_virtualcam_4096_streams"] = {"error": repr(exc)}
        if want is None or "float_dvs_operator_1080p_64_streams" in want:
            try:
                configs["float_dvs_operator_1080p_64_streams"] = run_float_entry(args, device, peak)
            except Exception as exc:
                configs["float_dvs_operator_1080p_64_streams"] = {"error": repr(exc)}

    if rank != 0:
        return 0

    # ---- collect the parity results ---------------------------------------------------------------------
    parity = None
    ok = True
    if not args.no_parity:
        per = {}
        for label, fut in pending:
            r = fut.result()
            per.setdefault(label, []).append(r)
        for label, rs in per.items():
            summary = {"streams": [r["stream"] for r in rs], "frames": rs[0]["frames"],
                       "events_last_step": [r["events_last_step"] for r in rs],
                       "ref_equal": all(r["ref_equal"] for r in rs), "thr_equal": all(r["thr_equal"] for r in rs),
                       "events_equal": all(r["events_equal"] for r in rs)}
            if label == "headline":
                parity = summary
            elif label in configs:
                configs[label]["parity"] = summary
            ok = ok and summary["ref_equal"] and summary["thr_equal"] and summary["events_equal"]
        for extra in ("float_dvs_operator_1080p_64_streams", "cfg2_pipeline_virtualcam_4096_streams"):
            f = configs.get(extra, {}).get("parity")
            if f:
                ok = ok and f["ref_equal"] and f["thr_equal"] and f["events_equal"] and f.get("frames_equal", True)
        if parity is not None:
            parity["all_configs_equal"] = ok
            parity["checker"] = "oracle/ (C restatement pinned against the compiled reference), whole frame sequence from ref0"

    bpp = bytes_per_px(wl, args.frame_dtype)
    k1_bytes = bpp * S * P
    achieved = k1_bytes / (k1_ms / 1e3) / 1e9
    tr = load_traffic("%s:%s:%s%s" % (args.workload, args.pattern, args.frame_dtype, ":s8" if dvs.state_dtype == "uint8" else ""))
    traffic = tr["bytes_per_px"] * S * P if tr else None
    step_bytes = k1_bytes + 8 * events_last_step
    line = {
        "metric": "stream_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
        "warmup": Wm, "ms_per_step": elapsed_ms / K, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "gpixel_per_s": value * P / 1e9, "gpixel_per_s_per_gpu": value * P / 1e9 / world,
        "config": cfgd,
        "details": {"streams_per_gpu": S, "events_last_step": events_last_step, "frames_per_stream_before_timing": frames_done_timed - K,
                    "pipeline_depth": args.pipeline_depth, "cuda_graph": bool(t["k1_ms"] is None),
                    "k1_timing": ("the tile pass alone, CUDA events around 10 eager launches on the same ring right after the "
                                  "timed loop (the timed steps are CUDA-graph replays)") if t["k1_ms"] is None else
                                 "CUDA events around the tile pass of every timed step"},
        "roofline": {"bound": "hbm", "kernel": "K1 fused tile pass (k_tile_fast2 when the fast path applies; timed with the normally empty k_tile_redo launch)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "frac_of_8tbs_spec": achieved / 8000.0, "bytes_per_px": bpp, "k1_ms": k1_ms,
                     "state_dtype": dvs.state_dtype,
                     "layout_bytes_per_px": layout_bytes_per_px(wl, args.frame_dtype, dvs.state_dtype),
                     "frac_with_events": (k1_bytes + 8 * events_last_step) / (k1_ms / 1e3) / 1e9 / peak,
                     # the same fraction on the compulsory bytes of the layout actually resident (byte state: 2 + 1 + 1 B/px)
                     "frac_of_layout_bytes": layout_bytes_per_px(wl, args.frame_dtype, dvs.state_dtype) * S * P / (k1_ms / 1e3) / 1e9 / peak,
                     "scan_expand_ms": rest_ms, "k1_share_of_step": (k1_ms / (k1_ms + rest_ms)) if rest_ms else None,
                     "step_achieved_gbs": step_bytes / (elapsed_ms / K / 1e3) / 1e9,
                     "traffic_gbs": (traffic / (k1_ms / 1e3) / 1e9) if traffic else None,
                     "note": "achieved counts SURVEY 8(d)'s ALGORITHMIC bytes (frame read + int16 reference read + reference write "
                             "per pixel); the kernel does not write back pixels whose reference is unchanged and, with "
                             "state_dtype uint8, keeps the reference plane in one byte per pixel (exact: every update clips "
                             "to [0, 255]), so its DRAM traffic (`traffic`, from the ncu capture under profiles/) is lower "
                             "and frac can exceed 1 while the DRAM itself runs at traffic_gbs; configs.cfg5_4k_int16_state "
                             "is the same workload on int16 state planes"},
        "clocks": clocks, "gpu_launches": lps * K,
        "events_allgather": {"streams": int(all_counts.numel()), "total_events_last_step": int(all_counts.sum().item())},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if weak is not None:
        line["weak"] = weak
    if configs:
        line["configs"] = configs
    if parity is not None:
        line["parity_check"] = parity
    if cpu_state is not None:
        try:
            line["cpu_baseline"] = cpu_baseline_single_core(wl, cpu_state[0], cpu_state[1], cpu_state[2], cpu_state[3],
                                                            args.cpu_frames, args.pattern)
        except Exception as exc:  # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    return 0 if ok else 1


def run_e2e(args, case, world, barrier, allmax, frame_dtype, pool_pin, narrow=False):
    """Same metric through the public API with HOST buffers (pydvs_b200.HostFeed + DVSStreams.step): every step moves
    that step's frames from pinned host memory to the device, runs the step, and reads the AER result (CSR row pointers
    + events) back to pinned host memory.  The host ring holds the same frames in the same order as the device-timed
    loop.  The feed works two frame sets ahead: while step i runs, set i+1 is on the wire and -- with ``narrow`` --
    set i+2 is being narrowed from int16 to bytes on the host cores (exact, every value checked; see HostFeed)."""
    import psutil
    import torch
    from pydvs_b200 import HostFeed
    dvs, S, P, device = case.dvs, case.S, case.P, case.device
    elem = 1 if frame_dtype == "uint8" else 2
    frame_bytes = S * P * elem
    n_host = min(len(case.ring), 4)
    while n_host > 2 and n_host * frame_bytes * max(world, 1) > 0.5 * psutil.virtual_memory().available:
        n_host //= 2
    if n_host * frame_bytes * max(world, 1) > 0.6 * psutil.virtual_memory().available:
        return {"value": None, "unit": "frames/s", "skipped": "not enough host memory to pin the frames"}
    t_pin = time.perf_counter()
    if pool_pin.get("bytes", 0) < n_host * frame_bytes:
        pool_pin["buf"] = None
        pool_pin["buf"] = torch.empty(n_host * frame_bytes, dtype=torch.uint8).pin_memory()
        pool_pin["bytes"] = n_host * frame_bytes
    t_pin = time.perf_counter() - t_pin
    host = [pool_pin["buf"][j * frame_bytes:(j + 1) * frame_bytes].view(case.ring[0].dtype).view(case.ring[0].shape)
            for j in range(n_host)]
    # host[k % n_host] holds ring frame k for the next n_host steps (n_host divides the ring length)
    for j in range(n_host):
        host[(case.i + j) % n_host].copy_(case.ring[(case.i + j) % len(case.ring)])
    off_host = torch.empty(dvs.seg_offsets.shape, dtype=torch.int64).pin_memory()
    # the short host ring replays a few frames over and over (each wrap is a jump of the bar): heavier steps than the
    # resident ring's steady state, so make room for 4x its largest step; an overflow is reported, not fatal
    dvs.reserve_events(min(S * P * 2, int(getattr(case, "peak_events", 0) * 4) + (1 << 20)))
    ev_cap_host = dvs.event_capacity
    ev_host = torch.empty(ev_cap_host, dtype=torch.int64).pin_memory()
    s_x = torch.cuda.Stream(device)
    use_graph, dvs.use_graph = dvs.use_graph, False   # the device buffers alternate; keep the plain launches
    feed = HostFeed(dvs, narrow=narrow, threads=max(1, (os.cpu_count() or 2) // max(world, 1)))
    # plain pinned H2D ceiling of this run (same buffers, same ranks at the same time)
    probe = torch.empty_like(case.ring[0])
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    reps = 3
    for r in range(reps):
        probe.copy_(host[r % n_host], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    h2d_peak = reps * frame_bytes / allmax(time.perf_counter() - t0) / 1e9
    del probe
    K = max(1, args.e2e_steps)
    W_e = 7 if narrow == "auto" else 2   # untimed steps first: the pipeline fills, and an "auto" feed settles on its way
    overflow = False
    first = case.i
    x_done = [torch.cuda.Event() for _ in range(W_e + K + 1)]
    d2h_total = 0
    t0 = None
    # two phases through the same feed: W_e untimed steps (an "auto" feed settles on its way there), the pipeline drains,
    # then K timed steps from an empty pipeline -- the timed region pays its own fill
    for lo, hi in ((0, W_e), (W_e, W_e + K)):
        torch.cuda.synchronize()
        barrier()
        if lo == W_e:
            t0 = time.perf_counter()
        tickets = {lo: feed.submit(host[(first + lo) % n_host])}
        if lo + 1 < hi:
            tickets[lo + 1] = feed.submit(host[(first + lo + 1) % n_host])
        for i in range(lo, hi):
            with torch.cuda.stream(s_x):
                feed.step(tickets.pop(i))
                off_host.copy_(dvs.seg_offsets, non_blocking=True)
                x_done[i].record(s_x)
            if i + 2 < hi:   # two sets ahead: narrowed / copied while this step and the next copy run
                tickets[i + 2] = feed.submit(host[(first + i + 2) % n_host])
            x_done[i].synchronize()
            n_ev = int(off_host[-1])
            if n_ev > ev_cap_host:
                overflow, n_ev = True, ev_cap_host
            with torch.cuda.stream(s_x):
                ev_host[:n_ev].copy_(dvs.events[:n_ev], non_blocking=True)
            if i >= W_e:
                d2h_total += off_host.numel() * 8 + n_ev * 8
    torch.cuda.synchronize()
    barrier()
    secs = allmax(time.perf_counter() - t0)
    feed.close()
    case.i += W_e + K
    dvs.use_graph = use_graph
    try:
        dvs.check_status()
    except OverflowError:
        overflow = True
    value = bench_config(args, case.wl, world)["streams_total"] * K / secs
    narrowing = feed.narrow is True and frame_dtype == "int16" and feed.stats["out_of_range"] == 0
    narrowed = K if narrowing else 0
    wire_bytes = frame_bytes // 2 if narrowing else frame_bytes
    out = {"value": value, "unit": "frames/s", "h2d_bytes_per_step": wire_bytes, "d2h_bytes_per_step": d2h_total // K,
           "steps": K, "gpixel_per_s": value * P / 1e9, "host_frame_dtype": frame_dtype, "host_ring_frames": n_host,
           "host_frame_bytes_per_step": frame_bytes, "pin_seconds": t_pin, "h2d_peak_gbs": h2d_peak,
           "h2d_achieved_gbs": wire_bytes * K / secs / 1e9, "h2d_frac_of_peak": wire_bytes * K / secs / 1e9 / h2d_peak,
           "event_overflow": overflow,
           "host_narrowing": {"mode": str(narrow), "timed_steps_narrowed": narrowed, "untimed_steps": W_e,
                              "out_of_range_steps": feed.stats["out_of_range"], "threads": feed.threads,
                              "narrow_seconds_per_step": (feed.stats["narrow_seconds"] / feed.stats["narrowed"]) if feed.stats["narrowed"] else None,
                              "calibration": feed.stats["calibration"]},
           "note": "per GPU: pinned host frames -> (exact narrowing of int16 planes to bytes on the host cores when that is "
                   "faster than copying them) -> H2D -> fused step -> D2H of CSR offsets + events, every step; "
                   "h2d_peak_gbs = plain pinned-memory copies of the same buffers with all ranks copying at once"}
    return out


def main():
    args = parse_args()
    args.host_narrow = {"auto": "auto", "on": True, "off": False}[args.host_narrow]
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return 0
    import torch
    import torch.distributed as dist
    lib_path = os.path.join(ROOT, "pydvs_b200", "libpydvs_b200.so")
    if not os.path.exists(lib_path):  # fresh checkout: build the real thing (there is no fallback to fall back to)
        if local_rank == 0:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_pydvs_build", os.path.join(ROOT, "pydvs_b200", "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        else:
            while not os.path.exists(lib_path + ".sha256"):
                time.sleep(1.0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    try:
        return run_ours(args, wl, rank, local_rank, world)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
