"""Non-hot-path members of the module API: visualisation, statistics and the image animators that
``VirtualCam`` calls back into.  They exist so that callers can keep doing
``gs.<name>(...)``; they are plain torch tensor ops (no hand-written kernels -- SURVEY.md section 8(f)
ranks them "next") and run on the device of their input (NumPy input: the CUDA device when there is
one, and returned as NumPy).

Citations: ``generate_spikes.pyx`` ("gs") line ranges.
"""
import time

import numpy as np
import torch

UP_POLARITY, DOWN_POLARITY, MERGED_POLARITY, RECTIFIED_POLARITY = 0, 1, 2, 3


def _to_t(x, dtype=None):
    """-> (tensor, was_numpy)"""
    if isinstance(x, np.ndarray):
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        return (t if dtype is None else t.to(dtype)), True
    if isinstance(x, torch.Tensor):
        return (x if dtype is None else x.to(dtype)), False
    raise TypeError("expected numpy.ndarray or torch.Tensor, got %s" % type(x).__name__)


def _req_i16(x, name):
    dt = x.dtype
    if dt not in (np.dtype(np.int16), torch.int16):
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % (dt,))
    if x.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % x.ndim)


def _ret(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _i16wrap(v):
    return ((int(v) + 32768) % 65536) - 32768


def _trunc_i16(t):
    """DTYPE(float64 array): C cast toward zero."""
    return t.trunc().to(torch.int64).to(torch.int16)


def build_helpers(uncoded):
    api = {}

    def export(fn):
        api[fn.__name__] = fn
        return fn

    # ---- gs:566-601 -----------------------------------------------------------------------------
    @export
    def render_frame(spikes, curr_frame, width, height, polarity):
        _req_i16(spikes, "spikes"), _req_i16(curr_frame, "curr_frame")
        s, was_np = _to_t(spikes)
        c, _ = _to_t(curr_frame)
        c = c.to(s.device)
        gray = (c.to(torch.int32) & 0xff).to(torch.uint8)  # int16 -> uint8 assignment wraps
        out = gray.unsqueeze(-1).repeat(1, 1, 3).contiguous()
        if out.shape[0] != height or out.shape[1] != width:
            raise ValueError("could not broadcast input array from shape %s into shape (%d,%d)"
                             % (tuple(c.shape), height, width))
        green = torch.tensor([0, 255, 0], dtype=torch.uint8, device=s.device)
        red = torch.tensor([0, 0, 255], dtype=torch.uint8, device=s.device)
        if not uncoded and polarity == RECTIFIED_POLARITY:
            out[s != 0] = green
        if polarity in (UP_POLARITY, MERGED_POLARITY):
            out[s > 0] = green
        if polarity in (DOWN_POLARITY, MERGED_POLARITY):
            out[s < 0] = red
        return _ret(out, was_np)

    # ---- gs:605-648 -----------------------------------------------------------------------------
    @export
    def render_comparison(curr_frame, ref_frame, lap_curr, lap_ref, spikes_frame, width, height):
        c, was_np = _to_t(curr_frame)
        r, _ = _to_t(ref_frame)
        sf, _ = _to_t(spikes_frame)
        r, sf = r.to(c.device), sf.to(c.device)
        out = torch.zeros((height, 4 * width, 3), dtype=torch.uint8, device=c.device)

        def u8(t):
            return (t.to(torch.int32) & 0xff).to(torch.uint8).unsqueeze(-1)

        out[:height, 0:width, :] = u8(c)
        out[:height, width:2 * width, :] = u8(r)
        out[:height, 2 * width:3 * width, :] = u8((c - r).abs())  # int16 arithmetic, wraps
        out[:height, 3 * width:, :] = sf
        return _ret(out, was_np)

    # ---- gs:530-558 -----------------------------------------------------------------------------
    @export
    def count_bits(in_data):
        _req_i16(in_data, "in_data")
        t, was_np = _to_t(in_data)
        v = t.to(torch.int32) & 0xff  # astype(uint8)
        n = torch.zeros_like(v)
        for i in range(8):
            n += (v >> i) & 1
        return _ret(n.reshape(-1, 2).to(torch.int16), was_np)

    @export
    def average_bits(num_bits):
        _req_i16(num_bits, "num_bits")
        t, _ = _to_t(num_bits)
        sel = t[t > 0]
        return float(sel.to(torch.float64).mean().item()) if sel.numel() else float("nan")

    @export
    def root_mean_square(original, estimate):
        _req_i16(original, "original"), _req_i16(estimate, "estimate")
        o, _ = _to_t(original, torch.float64)
        e, _ = _to_t(estimate, torch.float64)
        return float(torch.sqrt(((o - e.to(o.device)) ** 2).mean()).item())

    # ---- animators, gs:1124-1298 -------------------------------------------------------------------
    if not uncoded:
        @export
        def mask_image(original, mask):
            _req_i16(original, "original")
            o, was_np = _to_t(original, torch.float64)
            m, _ = _to_t(mask, torch.float64)
            return _ret(_trunc_i16(o * m.to(o.device)), was_np)

    @export
    def traverse_image(original, frame_number, speed, bg_gray):
        _req_i16(original, "original")
        o, was_np = _to_t(original)
        moved = torch.full_like(o, int(bg_gray))
        width = o.shape[1]
        n = _i16wrap(int(float(frame_number) * float(speed)))  # DTYPE(frame_number*speed)
        if n == 0:
            n = 1
        if n <= width:
            moved[:, :n] = o[:, -n:]
        elif n < 2 * width:
            n = 2 * width - n + 1
            moved[:, -n:] = o[:, :n]
        return _ret(moved, was_np)

    @export
    def fade_image(original, frame_number, half_frame, bg_gray):
        _req_i16(original, "original")
        o, was_np = _to_t(original, torch.float64)
        if frame_number < half_frame - 1:
            alpha = float(frame_number) / float(half_frame)
        elif frame_number > half_frame + 1:
            alpha = float(half_frame * 2 - frame_number) / float(half_frame)
        else:
            alpha = 1.0
        return _ret(_trunc_i16(o * alpha), was_np)

    def move_image(o, delta_x, delta_y, bg_gray):
        """gs:1184-1215 on a tensor: content moves right for delta_x > 0 and up for delta_y > 0."""
        moved = torch.full_like(o, int(bg_gray))
        height, width = o.shape
        if delta_x < 0:
            nx, ox = slice(0, delta_x), slice(abs(delta_x), width)
        elif delta_x > 0:
            nx, ox = slice(delta_x, width), slice(0, -delta_x)
        else:
            nx = ox = slice(0, width)
        if delta_y > 0:
            ny, oy = slice(0, -delta_y), slice(delta_y, height)
        elif delta_y < 0:
            ny, oy = slice(abs(delta_y), height), slice(0, delta_y)
        else:
            ny = oy = slice(0, height)
        moved[ny, nx] = o[oy, ox]
        return moved

    @export
    def usaccade_image(original, frame_number, frames_per_usaccade, max_delta, center_x, center_y, bg_gray):
        _req_i16(original, "original")
        o, was_np = _to_t(original)
        center_x, center_y = int(center_x), int(center_y)
        if frame_number % frames_per_usaccade != 0:
            return _ret(move_image(o, center_x, center_y, bg_gray), was_np), center_x, center_y
        # gs:1238-1240: reseeds the global NumPy RNG from the wall clock, then two draws
        np.random.seed(seed=np.uint32(int(time.time() * (1000000000 if uncoded else 10000000000)) & 0xffffffff))
        center_x = _i16wrap(center_x + np.random.randint(-max_delta, max_delta + 1))
        center_y = _i16wrap(center_y + np.random.randint(-max_delta, max_delta + 1))
        return _ret(move_image(o, center_x, center_y, bg_gray), was_np), center_x, center_y

    @export
    def attention_image(original, previous, reference, frame_number, frames_per_usaccade, frame_saccade,
                        max_delta, center_x, center_y, bg_gray):
        _req_i16(original, "original")
        if frame_number % frame_saccade != 0:
            return usaccade_image(original, frame_number, frames_per_usaccade, max_delta, center_x, center_y,
                                  bg_gray)
        o, was_np = _to_t(original)
        p, _ = _to_t(previous)
        r, _ = _to_t(reference)
        new_w = o.shape[1] // 2

        def area_half(t):
            # cv2.resize(..., INTER_AREA) to (new_w, new_w): for an exact 2x reduction of int16 data
            # OpenCV averages each 2x2 block with round-half-up; other ratios go through cv2 on the host.
            t = t.to(o.device)
            if t.shape[0] == 2 * new_w and t.shape[1] == 2 * new_w:
                q = t.to(torch.int32).reshape(new_w, 2, new_w, 2).sum(dim=(1, 3))
                return ((q + 2) >> 2).to(torch.int16)
            import cv2
            small = cv2.resize(t.cpu().numpy(), (new_w, new_w), interpolation=cv2.INTER_AREA).astype(np.int16)
            return torch.from_numpy(small).to(o.device)

        diff = (area_half(p) - area_half(r)).abs()
        top_n = torch.argsort(diff.reshape(-1), stable=True)[-5:]   # equal values: by index (NumPy leaves it unspecified)
        max_idx = int(top_n[np.random.randint(5)].item())
        row, col = (max_idx // new_w) * 2, (max_idx % new_w) * 2
        center_x, center_y = new_w - col, new_w - row
        if center_x > new_w or center_x < -new_w or center_y > new_w or center_y < -new_w:
            center_x = center_y = 0
        return _ret(move_image(o, center_x, center_y, bg_gray), was_np), center_x, center_y

    api["_move_image"] = move_image
    return api
