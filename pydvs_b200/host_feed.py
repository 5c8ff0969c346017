"""Host -> device feed for :class:`pydvs_b200.DVSStreams` when the frames live in (pinned) host memory.

The reference's callers hold every frame as an int16 NumPy plane of 8-bit gray values (``stand_alone.py:160-166``,
``external_dvs_emulator_device.py:347-360``: ``cv2.cvtColor(..).astype(int16)``).  Fed from the host, a step is bound
by the PCIe copy of those planes (bench.py's e2e leg moves them at 0.98 of the pinned-copy ceiling), so this feed

* keeps one frame set in flight ahead of the step (copy stream + events, two device buffers), and
* optionally narrows int16 frames to bytes on the host cores first (``dvs_host_narrow_i16_u8``: multi-threaded AVX2,
  every value checked), copies half the bytes and lets the fused step read them through its uint8 frame path -- the
  results are identical for values in [0, 255]; a frame set with any other value is copied as int16 instead.

``narrow="auto"`` times both ways on the first frame set (all ranks of a job do so at the same time, so a host that is
short of cores or memory bandwidth is seen) and keeps the faster one.
"""
import ctypes as C
import os
import time
from concurrent.futures import ThreadPoolExecutor

import torch

from ._lib import check, lib


class HostFeed:
    def __init__(self, dvs, narrow="auto", threads=None):
        self.dvs = dvs
        self.device = dvs.device
        self.shape = (dvs.S, dvs.H, dvs.W)
        self.narrow = narrow
        self.threads = int(threads) if threads else max(1, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        self._copy = torch.cuda.Stream(device=self.device)
        self._dev16 = [None, None]
        self._dev8 = [None, None]
        self._stage8 = [None, None]
        self._h2d_done = [None, None]     # event: the copy out of stage slot i has finished
        self._step_done = [None, None]    # event: the step that read device slot i has finished
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._n = 0
        self._outstanding = 0
        self.stats = {"narrowed": 0, "plain": 0, "out_of_range": 0, "narrow_seconds": 0.0, "calibration": None}
        if narrow is True or narrow == "auto":
            # page-locking runs at a few GB/s: the staging buffers are pinned here, not inside the first steps
            for slot in (0, 1):
                self._stage(slot)
                self._dev(slot, torch.uint8)

    # -- buffers are created on first use (an int16-only run never allocates the byte buffers) -----------------
    def _dev(self, slot, dtype):
        pool = self._dev8 if dtype == torch.uint8 else self._dev16
        if pool[slot] is None:
            pool[slot] = torch.empty(self.shape, dtype=dtype, device=self.device)
        return pool[slot]

    def _stage(self, slot):
        if self._stage8[slot] is None:
            self._stage8[slot] = torch.empty(self.shape, dtype=torch.uint8).pin_memory()
        return self._stage8[slot]

    def _narrow_into(self, frames_host, stage):
        flag = C.c_int32(0)
        t0 = time.perf_counter()
        check(lib.dvs_host_narrow_i16_u8(C.c_void_p(frames_host.data_ptr()), C.c_void_p(stage.data_ptr()),
                                         frames_host.numel(), self.threads, C.byref(flag)))
        self.stats["narrow_seconds"] += time.perf_counter() - t0
        return flag.value == 0

    def _calibrate(self, frames_host):
        """Seconds per frame set: host narrowing, and the plain int16 copy (measured on a slice, scaled)."""
        n = frames_host.numel()
        m = min(n, 1 << 28)                      # up to 256 Mi pixels of the frame set (small copies under-rate the link)
        src = frames_host.view(-1)[:m]
        stage = self._stage(0).view(-1)
        dev = self._dev(0, torch.int16).view(-1)
        flag = C.c_int32(0)
        best_n = best_c = 1e9
        with torch.cuda.stream(self._copy):      # (first touch of the device buffer, outside the timing)
            dev[:m].copy_(src, non_blocking=True)
        for _ in range(2):
            t0 = time.perf_counter()
            check(lib.dvs_host_narrow_i16_u8(C.c_void_p(src.data_ptr()), C.c_void_p(stage.data_ptr()), m, self.threads, C.byref(flag)))
            best_n = min(best_n, time.perf_counter() - t0)
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            with torch.cuda.stream(self._copy):
                dev[:m].copy_(src, non_blocking=True)
            self._copy.synchronize()
            best_c = min(best_c, time.perf_counter() - t0)
        t_narrow, t_copy16 = best_n * n / m, best_c * n / m
        # pipelined, a narrowed step costs max(narrowing, half the copy) and a plain one the whole copy -- but the narrowing
        # threads and the copy engine share the host's memory bandwidth (two ranks narrowing at once on one box: a
        # predicted 45 against 50 ms became 52 against 47 ms), so narrowing has to win by a clear margin
        use = max(1.25 * t_narrow, 0.5 * t_copy16) < t_copy16
        self.stats["calibration"] = {"narrow_s_per_set": t_narrow, "copy_int16_s_per_set": t_copy16, "narrow": bool(use),
                                     "threads": self.threads}
        if not use:   # give the pinned staging memory back
            self._stage8 = [None, None]
            self._dev8 = [None, None]
        return use

    def _produce(self, frames_host, slot):
        """Worker thread: (narrow ->) enqueue the H2D copy of one frame set; returns (device tensor, copy-done event)."""
        if self.device.index is not None:
            torch.cuda.set_device(self.device)
        if self.narrow == "auto" and frames_host.dtype == torch.int16:
            self.narrow = self._calibrate(frames_host)
        want_narrow = self.narrow is True and frames_host.dtype == torch.int16
        src, dtype = frames_host, frames_host.dtype
        if want_narrow:
            if self._h2d_done[slot] is not None:
                self._h2d_done[slot].synchronize()       # the copy that last read this staging buffer
            stage = self._stage(slot)
            if self._narrow_into(frames_host, stage):
                src, dtype = stage, torch.uint8
                self.stats["narrowed"] += 1
            else:
                self.stats["out_of_range"] += 1
        if src is frames_host:
            self.stats["plain"] += 1
        dev = self._dev(slot, dtype)
        with torch.cuda.stream(self._copy):
            if self._step_done[slot] is not None:
                self._copy.wait_event(self._step_done[slot])   # the step that last read this device buffer
            dev.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._copy)
        if src is not frames_host:
            self._h2d_done[slot] = done
        return dev, done

    def submit(self, frames_host):
        """Start moving one frame set ([S, H, W] int16 or uint8, pinned host memory) to the device; returns a ticket for
        :meth:`step`.  At most two tickets may be outstanding."""
        if frames_host.is_cuda or tuple(frames_host.shape[-2:]) != self.shape[1:] or frames_host.numel() != self.dvs.S * self.dvs.P:
            raise ValueError("frames_host must be a host tensor of shape [%d, %d, %d]" % self.shape)
        if frames_host.dtype not in (torch.int16, torch.uint8) or not frames_host.is_contiguous():
            raise ValueError("frames_host must be contiguous int16 or uint8")
        if self._outstanding >= 2:
            raise RuntimeError("two frame sets are already in flight: call step() before the next submit()")
        self._outstanding += 1
        slot = self._n % 2
        self._n += 1
        return slot, self._pool.submit(self._produce, frames_host, slot)

    def step(self, ticket, **kw):
        """Run the DVS step on the frame set of ``ticket`` (on the current stream) once its copy has landed."""
        slot, fut = ticket
        dev, done = fut.result()
        self._outstanding -= 1
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(done)
        res = self.dvs.step(dev, **kw)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._step_done[slot] = ev
        return res

    def close(self):
        self._pool.shutdown(wait=True)
