"""Test helper: execute the reference's caller scripts UNCHANGED (byte-compiled by oracle/build_ref.py from the
sources where they lie) against a chosen ``pydvs.generate_spikes`` -- the compiled reference module (to record golden
results) or ``pydvs_b200`` under ``install_as_pydvs()`` (to show that callers need no edits).

``stand_alone.py`` opens a webcam and a window; those two things are replaced from outside for the duration of a run:
``cv2.VideoCapture`` serves seeded synthetic BGR frames, ``cv2.imshow`` records what the script draws, ``cv2.waitKey``
answers 'q' after ``n_frames`` frames.  Everything else in the script -- its frame grabbing, buffers, the five-call
hot loop (``stand_alone.py:160-185``) and ``render_frame`` -- runs as written.
"""
import sys
import types
import zlib

import numpy as np


class FakeCapture:
    """A 320x240 BGR 'webcam': a bright bar moving over a noisy background (seeded)."""

    def __init__(self, *_a, **_k):
        self.rs, self.t = np.random.RandomState(7), 0

    def isOpened(self):
        return True

    def set(self, *_a):
        return False

    def get(self, _prop):
        return 0.0            # stand_alone.py:137-138 then assumes 125 fps -> max_time_ms = 8

    def read(self):
        g = np.full((240, 320), 60, np.int32)
        x0 = (self.t * 9) % 320
        g[:, x0:x0 + 24] = 210
        g[(self.t * 5) % 240:(self.t * 5) % 240 + 12, :] += 40
        g += self.rs.randint(-6, 7, size=g.shape)
        g = np.clip(g, 0, 255).astype(np.uint8)
        self.t += 1
        return True, np.stack([g, g, g], axis=2)

    def release(self):
        pass


def run_stand_alone(code, gs_module, n_frames, res=128):
    """``gs_module`` None: the script imports ``pydvs.generate_spikes`` through ``pydvs_b200.install_as_pydvs()``.
    -> dict(shown=[crc32 of every frame the script displayed], lists=[crc32 of every spike_lists], ref=final reference
    plane, spike_lists=the last list of lists, n=frames)."""
    import cv2
    shown, lists_crc = [], []
    state = {"frames": 0}
    saved_cv2 = {k: getattr(cv2, k) for k in ("VideoCapture", "namedWindow", "startWindowThread", "imshow", "waitKey",
                                              "destroyAllWindows")}
    saved_mod = {k: sys.modules.get(k) for k in ("pydvs", "pydvs.generate_spikes")}
    saved_argv = sys.argv
    g = {"__name__": "__main__", "__file__": code.co_filename}

    def imshow(_name, img):
        shown.append(zlib.crc32(np.ascontiguousarray(img).tobytes()))
        sl = g.get("spike_lists")
        flat = [int(k) for slot in sl for k in slot] + [len(slot) for slot in sl]
        lists_crc.append(zlib.crc32(np.asarray(flat, dtype=np.int64).tobytes()))

    def wait_key(_ms):
        state["frames"] += 1
        return ord("q") if state["frames"] >= n_frames else 0

    try:
        cv2.VideoCapture, cv2.imshow, cv2.waitKey = FakeCapture, imshow, wait_key
        cv2.namedWindow = cv2.startWindowThread = cv2.destroyAllWindows = lambda *a, **k: None
        if gs_module is None:
            import pydvs_b200
            pydvs_b200.install_as_pydvs(force=True)
        else:
            pkg = types.ModuleType("pydvs")
            pkg.generate_spikes = gs_module
            sys.modules["pydvs"], sys.modules["pydvs.generate_spikes"] = pkg, gs_module
        sys.argv = ["stand_alone.py", "--res", str(res)]
        import contextlib
        import io
        import warnings
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exec(code, g)
    finally:
        for k, v in saved_cv2.items():
            setattr(cv2, k, v)
        if gs_module is None:
            pydvs_b200.uninstall_pydvs_alias()
        for k, v in saved_mod.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.argv = saved_argv
    return {"shown": np.array(shown, dtype=np.int64), "lists": np.array(lists_crc, dtype=np.int64),
            "ref": np.array(g["ref"], dtype=np.int16), "spike_lists": [[int(k) for k in s] for s in g["spike_lists"]],
            "n": len(shown), "globals": g}
