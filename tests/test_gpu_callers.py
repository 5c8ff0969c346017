"""GPU: the reference's CALLERS run unchanged on the drop-in (north star: "external_dvs_emulator_device, VirtualCam and
the stand_alone* scripts run unchanged").  ``oracle/build_ref.py`` byte-compiles ``stand_alone.py`` and
``pydvs/virtual_cam.py`` from the reference tree into ``oracle/_ref/callers/`` (build output that travels to the GPU
box); here they are executed as they are, importing ``pydvs.generate_spikes`` through ``pydvs_b200.install_as_pydvs()``,
and their results are compared with recordings of the same scripts running on the compiled reference module
(``tests/golden/make_golden_callers.py``, ``make_golden_vcam.py``)."""
import os
import sys
import types
import zlib

import numpy as np
import pytest
import torch

from oracle import build_ref
import _callers

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_callers_v1.npz"))


def _code(name):
    code = build_ref.load_caller_code(name)
    if code is None:
        pytest.skip("oracle/_ref/callers not built (python oracle/build_ref.py needs the reference tree)")
    return code


@pytest.mark.parametrize("res", [128, 64])
def test_stand_alone_py_runs_unchanged_on_the_drop_in(res):
    """60 frames of stand_alone.py's own loop (NumPy buffers, :151-200) on the CUDA kernels: every displayed frame
    (render_frame), every spike list, the final reference plane and the last list of lists equal the reference run."""
    r = _callers.run_stand_alone(_code("stand_alone"), None, 60, res=res)
    assert r["n"] == 60
    assert np.array_equal(r["shown"], G["res%d/shown" % res])
    assert np.array_equal(r["lists"], G["res%d/lists" % res])
    assert np.array_equal(r["ref"], G["res%d/ref" % res])
    assert [len(s) for s in r["spike_lists"]] == G["res%d/last_lens" % res].tolist()
    assert [k for s in r["spike_lists"] for k in s] == G["res%d/last_keys" % res].tolist()
    assert "pydvs" not in sys.modules      # the alias is removed again


def test_stand_alone_loop_body_on_cuda_tensors():
    """The same loop body (stand_alone.py:160-185, same arguments) handed CUDA tensors: nothing leaves the device except
    what the test checks."""
    import cv2
    import pydvs_b200.generate_spikes as gs
    res = 128
    cap = _callers.FakeCapture()
    data_shift = np.uint8(np.log2(res))
    up_down_shift, data_mask, polarity = np.uint8(2 * data_shift), np.uint8(res - 1), np.uint8(2)
    threshold, max_threshold, max_time_ms, num_bits, num_active_bits, history_weight = 12, 180, 8, 6, 2, 1.0
    log2_table = gs.generate_log2_table(num_active_bits, num_bits)[num_active_bits - 1]
    inh_coords = gs.generate_inh_coords(res, res, 2)
    ref = torch.full((res, res), 128, dtype=torch.int16, device="cuda")
    lut = torch.as_tensor(log2_table, device="cuda")
    for i in range(60):
        _, raw = cap.read()
        h, w, _ = raw.shape
        new_w = int(float(res * w) / float(h))
        c0 = (new_w - res) // 2
        img = cv2.resize(cv2.cvtColor(raw, cv2.COLOR_BGR2GRAY).astype(np.int16), (new_w, res))[:, c0:c0 + res]
        curr = torch.from_numpy(np.ascontiguousarray(img)).cuda()
        diff, abs_diff, spikes = gs.thresholded_difference(curr, ref, threshold)
        spikes = gs.local_inhibition(spikes, abs_diff, inh_coords, res, res, 2)
        ref = gs.update_reference_time_binary_thresh(abs_diff, spikes, ref, threshold, max_time_ms, num_active_bits,
                                                     history_weight, lut)
        neg_spks, pos_spks, max_diff = gs.split_spikes(spikes, abs_diff, polarity)
        spike_lists = gs.make_spike_lists_time_bin_thr(pos_spks, neg_spks, max_diff, up_down_shift, data_shift, data_mask,
                                                       max_time_ms, threshold, max_threshold, num_bits, lut)
        spk_img = gs.render_frame(spikes, curr, res, res, polarity)
        assert spk_img.is_cuda and ref.is_cuda
        assert zlib.crc32(spk_img.cpu().numpy().astype(np.uint8).tobytes()) == int(G["res128/shown"][i]), i
        flat = [int(k) for slot in spike_lists for k in slot] + [len(slot) for slot in spike_lists]
        assert zlib.crc32(np.asarray(flat, dtype=np.int64).tobytes()) == int(G["res128/lists"][i]), i
    assert np.array_equal(ref.cpu().numpy(), G["res128/ref"])


@pytest.mark.parametrize("name", ["SACCADE_mask", "SACCADE", "TRAVERSE_mask", "FADE", "ATTENTION_mask", "NONE"])
def test_reference_virtual_cam_class_runs_unchanged_on_the_drop_in(name, tmp_path):
    """The reference's own VirtualCam class (pydvs/virtual_cam.py, byte-compiled, unchanged) calling back into the drop-in
    module for its animators (gs.usaccade_image / attention_image / traverse_image / fade_image / mask_image) serves
    the frames it served on the compiled reference module."""
    import cv2
    import glob
    import pydvs_b200
    from pydvs_b200 import _helpers
    code = _code("virtual_cam")
    g = np.load(os.path.join(HERE, "golden", "golden_vcam_v1.npz"))
    fps, on_ms, off_ms, max_dist, fp_usacc, fp_sacc, cycles, bg = (int(v) for v in g["args"])
    behaviour, masked = name.split("_")[0], name.endswith("_mask")
    # fixtures the class reads from disk, written from the golden arrays: its images and (next to its own file) the mask
    for i, img in enumerate(g["images"]):
        cv2.imwrite(str(tmp_path / ("img_%02d.png" % i)), img[2:30, 2:30].astype(np.uint8))
    cv2.imwrite(os.path.join(build_ref.CALLERS_DIR, "fading_mask.png"), np.rint(g["mask"] * 255.0).astype(np.uint8))

    class Clock:
        now = 0.0

    fake_time = types.ModuleType("time")
    fake_time.time = lambda: Clock.now
    stub_glob2 = types.ModuleType("glob2")
    stub_glob2.glob = glob.glob
    saved_time, saved_glob2 = _helpers.time, sys.modules.get("glob2")
    pydvs_b200.install_as_pydvs(force=True)
    sys.modules["glob2"] = stub_glob2
    try:
        mod = types.ModuleType("_ref_virtual_cam")
        mod.__file__ = code.co_filename
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            exec(code, mod.__dict__)
            mod.get_time = fake_time.time
            _helpers.time = fake_time
            np.random.seed(12345)
            cam = mod.VirtualCam(str(tmp_path), behaviour=behaviour, fps=fps, resolution=32, image_on_time_ms=on_ms,
                                 inter_off_time_ms=off_ms, max_saccade_distance=max_dist, frames_per_microsaccade=fp_usacc,
                                 frames_per_saccade=fp_sacc, max_cycles=cycles, fade_with_mask=masked, background_gray=bg)
        frames, valid, centers = g[name + "/frames"], g[name + "/valid"], g[name + "/centers"]
        tied = set(int(i) for i in g[name + "/tied_at"])
        compared = 0
        for i in range(len(frames)):
            ref = g[name + "/ref_planes"][i] if behaviour == "ATTENTION" else np.zeros((32, 32), np.int16)
            ok, img = cam.read(ref)
            assert bool(ok) == bool(valid[i]), (name, i)
            if i in tied and (int(cam.center_x), int(cam.center_y)) != tuple(centers[i]):
                cam.center_x, cam.center_y = int(centers[i][0]), int(centers[i][1])   # unspecified order in the reference
                cam.current_image[:] = frames[i]
            else:
                assert np.array_equal(np.asarray(img), frames[i]), (name, i)
                compared += 1
            Clock.now += 1.0 / fps
        cam.stop()
        assert compared >= len(frames) - 6
    finally:
        _helpers.time = saved_time
        pydvs_b200.uninstall_pydvs_alias()
        if saved_glob2 is None:
            sys.modules.pop("glob2", None)
        else:
            sys.modules["glob2"] = saved_glob2
