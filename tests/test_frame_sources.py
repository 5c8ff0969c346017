"""Frame sources: NumPy oracle vs golden vectors from the compiled reference
(CPU), batched CUDA kernel vs the oracle (GPU)."""
import os

import numpy as np
import pytest

from oracle import frame_oracle as fo

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_oracle_move_and_mask_match_the_reference():
    src, mask = G["move_src"], G["move_mask"]
    for i, (cx, cy) in enumerate(G["move_offsets"]):
        assert np.array_equal(fo.move_image(src, int(cx), int(cy), 9), G["move_plain"][i]), (cx, cy)
        assert np.array_equal(fo.mask_image(fo.move_image(src, int(cx), int(cy), 0), mask), G["move_masked"][i]), (cx, cy)


def test_saccade_offsets_follow_the_reference_state_machine():
    from pydvs_b200.frame_sources import saccade_offsets
    off = saccade_offsets(64, 10, frames_per_microsaccade=1, max_dist=1, seed=3)
    assert off.shape == (10, 64, 2) and np.all(off[0] == 0)
    # scalar replay of mnist_to_spikes.py:246-258 for every stream with the same draws
    steps = np.random.RandomState(3).randint(-1, 2, size=(10, 64, 2))
    for s in range(64):
        cx = cy = 0
        for t in range(1, 10):
            cx, cy = cx + steps[t, s, 0], cy + steps[t, s, 1]      # usaccade_image redraws on every frame here
            assert (off[t, s, 0], off[t, s, 1]) == (cx, cy)
            if abs(cx) > 1:
                cx = 0
            elif abs(cy) > 1:
                cy = 0
    again = saccade_offsets(64, 10, frames_per_microsaccade=1, max_dist=1, seed=3)
    assert np.array_equal(off, again)       # deterministic, unlike the wall-clock reseeding of the reference
    held = saccade_offsets(4, 9, frames_per_microsaccade=3, max_dist=1, seed=1)
    assert np.array_equal(held[1], held[2]) and np.array_equal(held[4], held[5])


@pytest.mark.gpu
def test_batched_move_mask_kernel_matches_the_oracle():
    import torch
    from pydvs_b200.frame_sources import SaccadeSource, move_images
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (5, 24, 40)).astype(np.int16)
    mask = rng.random((24, 40))
    S = 37
    idx = rng.integers(0, 5, S).astype(np.int32)
    dx = rng.integers(-45, 46, S).astype(np.int32)
    dy = rng.integers(-30, 31, S).astype(np.int32)
    d_imgs = torch.from_numpy(imgs).cuda()
    for m, bg in ((None, 9), (mask, 0), (mask, 17)):
        got = move_images(d_imgs, idx, dx, dy, bg, m).cpu().numpy()
        for s in range(S):
            exp = fo.move_image(imgs[idx[s]], int(dx[s]), int(dy[s]), bg)
            if m is not None:
                exp = fo.mask_image(exp, m)
            assert np.array_equal(got[s], exp), (s, dx[s], dy[s])
    # golden vectors through the kernel as well
    src = torch.from_numpy(G["move_src"][None]).cuda()
    offs = G["move_offsets"]
    got = move_images(src, np.zeros(len(offs), np.int32), offs[:, 0], offs[:, 1], 0, G["move_mask"]).cpu().numpy()
    assert np.array_equal(got, G["move_masked"])
    # the batched source reproduces the per-stream oracle loop
    cam = SaccadeSource(d_imgs, idx, frames_per_image=6, fade_mask=mask, seed=5)
    for t in range(6):
        fr = cam.frame(t).cpu().numpy()
        for s in range(S):
            cx, cy = cam.offsets_host[t, s]
            assert np.array_equal(fr[s], fo.mask_image(fo.move_image(imgs[idx[s]], int(cx), int(cy), 0), mask))
