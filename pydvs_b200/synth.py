"""Seeded synthetic frame sources: generated on the device with torch so that
benchmarks never feed frames over PCIe, and downloadable so the CPU oracle sees identical frames.
Not part of the hot path; frame synthesis is never inside a timed region.
"""
import torch


def moving_bar(S, H, W, t, device="cuda", bg=64, bar=200, bar_width=None, speed=None, noise=0, seed=0,
               dtype=torch.int16, stream_offset=0):
    """Vertical bar (value `bar`, width W/16) on background `bg`, moving W/64 px per frame with
    wrap-around; stream s starts at phase ((s + stream_offset) * 37) % W; optional uniform noise in [-noise, noise]."""
    bw = bar_width if bar_width is not None else max(1, W // 16)
    sp = speed if speed is not None else max(1, W // 64)
    x = torch.arange(W, device=device).view(1, 1, W)
    phase = ((torch.arange(S, device=device) + stream_offset) * 37) % W
    x0 = ((phase + t * sp) % W).view(S, 1, 1)
    inside = ((x - x0) % W) < bw
    frame = torch.where(inside, torch.tensor(bar, device=device), torch.tensor(bg, device=device))
    frame = frame.expand(S, H, W).to(torch.int32)
    if noise:
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + t)
        frame = frame + torch.randint(-noise, noise + 1, (S, H, W), generator=g, device=device, dtype=torch.int32)
    return frame.clamp_(0, 255).to(dtype).contiguous()


def random_frames(S, H, W, t, device="cuda", seed=0, lo=0, hi=256, dtype=torch.int16):
    """Dense worst case: i.i.d. uniform pixels in [lo, hi) per frame."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 7919 + t)
    return torch.randint(lo, hi, (S, H, W), generator=g, device=device, dtype=torch.int32).to(dtype).contiguous()


def saccade_frames(images, t, offsets, device="cuda", bg=0, dtype=torch.int16):
    """MNIST-style micro-saccade stream (mnist_to_spikes.py:246-260 with injected offsets):
    images [S, H, W]; offsets [S, 2] integer (dx, dy) for this frame; content moves right / up."""
    S, H, W = images.shape
    out = torch.full((S, H, W), bg, dtype=images.dtype, device=device)
    ys = torch.arange(H, device=device).view(1, H, 1)
    xs = torch.arange(W, device=device).view(1, 1, W)
    dx = offsets[:, 0].view(S, 1, 1).to(device)
    dy = offsets[:, 1].view(S, 1, 1).to(device)
    src_x, src_y = xs - dx, ys + dy
    ok = (src_x >= 0) & (src_x < W) & (src_y >= 0) & (src_y < H)
    idx = (src_y.clamp(0, H - 1) * W + src_x.clamp(0, W - 1)).expand(S, H, W)
    gathered = torch.gather(images.reshape(S, H * W).to(device), 1, idx.reshape(S, H * W)).reshape(S, H, W)
    out = torch.where(ok, gathered, out)
    return out.to(dtype).contiguous()


# ---- counter-based frames: the same pixels from torch (any device) and from NumPy (bench.py's reference arm) ----
_M32 = 0xFFFFFFFF


def _mix32(h):
    """lowbias32-style integer hash on int64 tensors holding 32-bit values (constants < 2**31, so no
    product leaves the int64 range: NumPy and torch agree bit for bit on every device)."""
    h = h ^ (h >> 16)
    h = (h * 0x7FEB352D) & _M32
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & _M32
    return h ^ (h >> 16)


def hashed_frames(S, H, W, t, pattern="bar", device="cuda", stream_offset=0, noise=8, bg=64, bar=200,
                  dtype=torch.int16):
    """Frame ``t`` of streams ``stream_offset .. stream_offset + S - 1``.

    ``bar``: the moving bar of :func:`moving_bar` plus uniform noise in ``[-noise, noise]`` drawn from a
    hash of (global stream id, t, pixel index); ``dense``: uniform pixels in [0, 255] from the same hash.
    ``bench.np_hashed_frame`` is the NumPy twin (checked in tests/test_abi_and_host.py)."""
    P = H * W
    q = torch.arange(P, device=device, dtype=torch.int64).view(1, P)
    sid = (torch.arange(S, device=device, dtype=torch.int64) + stream_offset).view(S, 1)
    h = _mix32((((sid * 1000003 + t) * 2654435761) + q * 40503) & _M32)
    if pattern == "dense":
        return (h & 0xFF).view(S, H, W).to(dtype).contiguous()
    bw, sp = max(1, W // 16), max(1, W // 64)
    x = q % W
    x0 = ((sid * 37) % W + t * sp) % W
    base = torch.where(((x - x0) % W) < bw, bar, bg)
    if noise:
        base = base + (h % (2 * noise + 1)) - noise
    return base.clamp_(0, 255).view(S, H, W).to(dtype).contiguous()
