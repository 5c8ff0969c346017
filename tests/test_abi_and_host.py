"""CPU: the C-ABI library loads and exports every symbol include/pydvs_b200.h declares; host-side logic
(argument validation that precedes any launch, tiling choice, SpikeLists, helper functions, config)."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pydvs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dvs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pydvs_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 16
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libpydvs_b200.so does not export %s" % n
    assert set(_lib.EXPORTS) == set(names), "ctypes binding and header disagree"
    assert _lib.lib.dvs_abi_version() == _lib.ABI_VERSION == 4


def test_struct_layouts_match_the_header():
    from pydvs_b200 import _lib
    assert ctypes.sizeof(_lib.EncParams) == 32
    assert ctypes.sizeof(_lib.Tiling) == 16
    assert _lib.StepParams.history_weight.offset % 8 == 0
    assert ctypes.sizeof(_lib.StepParams) == 16 + 16 * 4 + 8 + 32 + 10 * 8


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """A C99 translation unit that includes the header, checks the struct sizes the ctypes binding assumes and calls
    two entry points that need no GPU links against libpydvs_b200.so and runs (what a cgo / JNI / FFI user does)."""
    import shutil
    import subprocess
    from pydvs_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text("""
#include <stdio.h>
#include "pydvs_b200.h"
int main(void) {
  if (sizeof(dvs_step_params) != %d || sizeof(dvs_enc_params) != %d || sizeof(dvs_tiling) != %d) return 2;
  if (sizeof(dvs_nvs_params) != %d || sizeof(dvs_key_list) != %d || sizeof(dvs_event) != 8 || sizeof(dvs_record) != 8) return 3;
  if (dvs_abi_version() != DVS_ABI_VERSION) return 4;
  if (dvs_step_fused(NULL, NULL) != DVS_ERR_INVALID_ARGUMENT) return 5;
  if (dvs_choose_tile_rows(256, 2160, 3840, 2) != 2) return 6;
  puts(dvs_last_error_string());
  return 0;
}
""" % (ctypes.sizeof(_lib.StepParams), ctypes.sizeof(_lib.EncParams), ctypes.sizeof(_lib.Tiling),
       ctypes.sizeof(_lib.NvsParams), ctypes.sizeof(_lib.KeyList)))
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-L", lib_dir, "-lpydvs_b200", "-Wl,-rpath," + lib_dir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)


def test_choose_tile_rows_and_argument_errors_without_a_gpu():
    from pydvs_b200 import _lib
    lib = _lib.lib
    for (S, H, W, k) in [(256, 2160, 3840, 2), (1, 128, 128, 1), (4096, 32, 32, 1), (1, 480, 640, 2), (64, 1080, 1920, 1),
                         (1, 8, 2064, 4), (3, 33, 17, 1)]:
        rows = lib.dvs_choose_tile_rows(S, H, W, k)
        assert rows >= max(k, 1) and rows % max(k, 1) == 0 and rows * W <= 32768
    assert lib.dvs_choose_tile_rows(1, 16, 40000, 1) == -1
    # invalid arguments are rejected before any CUDA call
    assert lib.dvs_step_fused(None, None) == _lib.ERR_INVALID_ARGUMENT
    assert b"null" in lib.dvs_last_error_string()
    assert lib.dvs_local_inhibition_i16(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 6, 10, 4, None) == _lib.ERR_INVALID_ARGUMENT
    assert lib.dvs_scan_tiles(0, 1, 1, None, None, None, None, None) == _lib.ERR_INVALID_ARGUMENT
    # dvs_pad_rows: W even, Wp = W rounded up to a multiple of 8, 1- or 2-byte pixels, aligned buffers
    P = ctypes.c_void_p
    for args in ((P(64), P(64), 4, 345, 352, 2), (P(64), P(64), 4, 346, 360, 2), (P(64), P(64), 4, 346, 350, 2),
                 (P(64), P(64), 4, 346, 352, 4), (P(66), P(64), 4, 346, 352, 2), (P(64), P(72), 4, 346, 352, 2),
                 (None, P(64), 4, 346, 352, 2)):
        assert lib.dvs_pad_rows(*args, None) == _lib.ERR_INVALID_ARGUMENT, args
    assert lib.dvs_pad_rows(P(64), P(64), 0, 346, 352, 2, None) == _lib.OK      # nothing to do: no launch
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_INVALID_ARGUMENT)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED)


def test_module_api_surface_matches_the_reference_names():
    import pydvs_b200.generate_spikes as gs
    import pydvs_b200.generate_spikes_uncoded as gsu
    coded = ["thresholded_difference", "thresholded_difference_adpt", "generate_inh_coords", "local_inhibition",
             "update_reference_rate", "update_reference_rate_adpt", "update_reference_time_thresh",
             "update_reference_time_binary_raw", "update_reference_time_binary_thresh",
             "update_reference_time_binary_raw_noise", "update_reference_time_thresh_noise", "split_spikes",
             "make_spike_lists_rate", "make_spike_lists_time", "make_spike_lists_time_bin",
             "make_spike_lists_time_bin_thr", "generate_log2_table", "render_frame", "render_comparison", "count_bits",
             "average_bits", "root_mean_square", "mask_image", "traverse_image", "fade_image", "usaccade_image",
             "attention_image"]
    for n in coded:
        assert callable(getattr(gs, n)), n
    for n in coded:
        if n in ("update_reference_time_thresh_noise", "mask_image"):
            assert not hasattr(gsu, n)
        else:
            assert callable(getattr(gsu, n)), n
    assert gs.DTYPE is np.int16 and gs.DTYPE_U8 is np.uint8 and gs.DTYPE_U16 is np.uint16 and gs.DTYPE_FLOAT is float
    assert gs.RECTIFIED_POLARITY == 3 and not hasattr(gsu, "RECTIFIED_POLARITY")


def test_host_tables_match_reference_when_available(reference_modules):
    import pydvs_b200.generate_spikes as gs
    from oracle import oracle as orc
    for a, b in ((4, 8), (2, 6), (8, 8), (3, 5), (1, 4)):
        assert np.array_equal(gs.generate_log2_table(a, b), orc.generate_log2_table(a, b))
    for w, h, k in ((8, 8, 2), (16, 12, 4), (6, 10, 2), (9, 6, 3)):
        assert np.array_equal(gs.generate_inh_coords(w, h, k), orc.Oracle.generate_inh_coords(w, h, k))
    if reference_modules is not None:
        assert np.array_equal(gs.generate_log2_table(4, 8), reference_modules[0].generate_log2_table(4, 8))


def test_scalar_conversion_errors_like_cython():
    from pydvs_b200 import _api
    assert _api._i16(12) == 12 and _api._i16(np.int64(-5)) == -5
    with pytest.raises(OverflowError):
        _api._i16(70000)
    with pytest.raises(OverflowError):
        _api._u8(300)
    with pytest.raises(OverflowError):
        _api._u8(-1)
    with pytest.raises(TypeError):
        _api._i16(1.5)
    with pytest.raises(TypeError):
        _api._i16("12")


def test_spike_lists_behaves_like_a_list_of_lists():
    from pydvs_b200 import SpikeLists
    off = np.array([0, 2, 2, 5], dtype=np.int64)
    sl = SpikeLists(off, torch.tensor([5, -3, 7, 8, 9], dtype=torch.int16))
    assert len(sl) == 3 and sl[0] == [5, -3] and sl[1] == [] and list(sl)[2] == [7, 8, 9]
    assert sl == [[5, -3], [], [7, 8, 9]] and sl != [[5], [], []]
    assert pickle.loads(pickle.dumps(sl)) == [[5, -3], [], [7, 8, 9]]
    assert isinstance(pickle.loads(pickle.dumps(sl)), list)
    assert sl.slot_counts().tolist() == [2, 0, 3]
    tr = SpikeLists(np.array([0, 1, 2]), torch.tensor([[1, 2, 1], [3, 4, -1]], dtype=torch.int32))
    assert tr.tolist() == [[[1, 2, 1]], [[3, 4, -1]]]


def test_config_resolution_mirrors_the_device_defaults():
    from pydvs_b200 import DVSConfig
    c = DVSConfig(width=128, height=128, fps=125).resolved()
    assert c.max_time_ms == 8 and c.num_bins == 8 and c.encoder_threshold == 12
    c = DVSConfig(width=32, height=32, output_type="TIME_BIN_THR", adaptive_threshold=True, fps=60).resolved()
    assert c.max_time_ms == 16 and c.num_bins == 6 and c.encoder_threshold == 6
    assert DVSConfig(width=8, height=8, output_type="TIME_BIN").resolved().num_bins == 8
    with pytest.raises(ValueError):
        DVSConfig(width=8, height=8, output_type="nope").resolved()


def test_helpers_against_reference(reference_modules):
    """Non-hot-path helpers (plain tensor ops) on CPU tensors vs the compiled reference."""
    if reference_modules is None:
        pytest.skip("oracle/_ref not built")
    import pydvs_b200.generate_spikes as gs
    ref = reference_modules[0]
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (16, 16)).astype(np.int16)
    spikes = rng.integers(-1, 2, (16, 16)).astype(np.int16)
    t_img, t_spk = torch.from_numpy(img), torch.from_numpy(spikes)
    for pol in range(4):
        assert np.array_equal(gs.render_frame(t_spk, t_img, 16, 16, pol).numpy(), ref.render_frame(spikes, img, 16, 16, pol))
    mask = rng.random((16, 16))
    assert np.array_equal(gs.mask_image(t_img, torch.from_numpy(mask)).numpy(), ref.mask_image(img, mask))
    for fn, speed in ((0, 1.0), (3, 2.0), (9, 1.5), (20, 1.0), (40, 1.0), (5, 0.1)):
        assert np.array_equal(gs.traverse_image(t_img, fn, speed, 7).numpy(), ref.traverse_image(img, fn, speed, 7)), (fn, speed)
    for fn in (0, 3, 9, 10, 11, 15, 19):
        assert np.array_equal(gs.fade_image(t_img, fn, 10, 0).numpy(), ref.fade_image(img, fn, 10, 0)), fn
    for cx, cy in ((0, 0), (1, -1), (-3, 2), (5, 5), (-16, 0), (0, 17)):
        got, gx, gy = gs.usaccade_image(t_img, 3, 2, 1, cx, cy, 9)   # frame 3 % 2 != 0: pure move_image
        exp, ex, ey = ref.usaccade_image(img, 3, 2, 1, cx, cy, 9)
        assert np.array_equal(got.numpy(), exp) and (gx, gy) == (ex, ey), (cx, cy)
    assert np.array_equal(gs.count_bits(t_img).numpy(), ref.count_bits(img))
    nb = ref.count_bits(img)
    assert abs(gs.average_bits(torch.from_numpy(nb)) - ref.average_bits(nb)) < 1e-12
    assert abs(gs.root_mean_square(t_img, t_spk) - ref.root_mean_square(img, spikes)) < 1e-12
    frame3 = rng.integers(0, 256, (16, 16, 3)).astype(np.uint8)
    assert np.array_equal(gs.render_comparison(t_img, t_spk, t_img, t_img, torch.from_numpy(frame3), 16, 16).numpy(),
                          ref.render_comparison(img, spikes, img, img, frame3, 16, 16))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pydvs_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pydvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "dvs_oracle" not in text, f


def test_bench_frame_generators_agree_and_the_reference_arm_never_loads_the_product():
    """bench.py's NumPy frame generator (reference arm) and pydvs_b200.synth.hashed_frames (GPU arm) produce the same
    pixels; the reference arm runs without importing pydvs_b200 (so it cannot map libpydvs_b200.so)."""
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from pydvs_b200 import synth
    for pat in ("bar", "dense"):
        for (S, H, W, t, off) in [(3, 16, 64, 5, 7), (2, 54, 96, 0, 250)]:
            a = synth.hashed_frames(S, H, W, t, pattern=pat, device="cpu", stream_offset=off).numpy()
            for s in range(S):
                assert np.array_equal(a[s], bench.np_hashed_frame(off + s, H, W, t, pat)), (pat, s)
    code = ("import sys, json, io, contextlib; sys.argv=['bench.py','--impl','reference','--workload','vga_rate_inhibition',"
            "'--steps','3','--warmup','1','--settle','2']; import bench; buf=io.StringIO();\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "line=json.loads(buf.getvalue()); assert line['impl']=='reference' and line['steps']==3 and line['warmup']==1;\n"
            "assert not any(m.startswith('pydvs_b200') for m in sys.modules), 'reference arm imported the product';\n"
            "print(json.dumps(line['config']))")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    import argparse
    import json
    ref_cfg = json.loads(out.stdout.strip().splitlines()[-1])
    args = argparse.Namespace(streams=None, state_dtype="auto", scaling="strong", workload="vga_rate_inhibition", pattern="bar",
                              frame_dtype="int16", ring=16, settle=2)
    assert ref_cfg == bench.bench_config(args, bench.WORKLOADS["vga_rate_inhibition"], 1)   # both arms print this dict


def test_bench_parity_job_detects_a_flipped_pixel():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["vga_rate_inhibition"]
    H, W = 32, 64
    ring = [bench.np_hashed_frame(3, H, W, t) for t in range(4)]
    pipe = bench.oracle_pipeline(wl, W, H)
    for i in range(6):
        counts, flat = pipe.step_raw(ring[i % 4])
    job = {"stream": 3, "wl": wl, "n_frames": 6, "ring": ring, "ref": pipe.ref.copy(), "thr": None,
           "keys": flat[:, 0].astype(np.int16).copy(), "slot_counts": counts.copy()}
    r = bench.run_parity_job(job)
    assert r["ref_equal"] and r["events_equal"] and r["thr_equal"]
    job["ref"][5, 7] ^= 1
    assert not bench.run_parity_job(job)["ref_equal"]


def test_host_narrowing_is_exact_and_reports_values_outside_a_byte():
    """dvs_host_narrow_i16_u8 (no GPU): every size / thread count gives the bytes NumPy's cast gives, and any value
    outside [0, 255] -- negative ones included -- is reported."""
    from pydvs_b200 import _lib
    rng = np.random.default_rng(11)
    for n, threads in ((0, 1), (1, 1), (63, 2), (64, 1), (4097, 3), (1 << 16, 4), ((1 << 18) + 37, 8), (3 * 4096 * 5, 5)):
        src = rng.integers(0, 256, n, dtype=np.int16)
        raw = np.zeros(n + 64, np.uint8)
        off = (-raw.ctypes.data) % 32                    # 32-byte aligned destination: the AVX2 / streaming-store path
        for shift in (0, 1):                             # ... and an unaligned one: the SSE2 path
            dst = raw[off + shift:off + shift + n]
            flag = ctypes.c_int32(7)
            rc = _lib.lib.dvs_host_narrow_i16_u8(src.ctypes.data, dst.ctypes.data, n, threads, ctypes.byref(flag))
            assert rc == _lib.OK and flag.value == 0
            assert np.array_equal(dst, src.astype(np.uint8))
    src = rng.integers(0, 256, 1 << 17, dtype=np.int16)
    dst = np.zeros(src.size, np.uint8)
    for pos, bad in ((0, 256), (77, -1), (src.size - 1, 300), (70000, -32768), (65536, 32767)):
        s2 = src.copy()
        s2[pos] = bad
        flag = ctypes.c_int32(0)
        assert _lib.lib.dvs_host_narrow_i16_u8(s2.ctypes.data, dst.ctypes.data, s2.size, 4, ctypes.byref(flag)) == _lib.OK
        assert flag.value == 1, (pos, bad)
    assert _lib.lib.dvs_host_narrow_i16_u8(None, dst.ctypes.data, 4, 1, ctypes.byref(flag)) == _lib.ERR_INVALID_ARGUMENT
