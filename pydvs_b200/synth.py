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
