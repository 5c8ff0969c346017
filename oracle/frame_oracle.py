"""NumPy restatement of the reference's frame animators.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

``move_image`` follows ``generate_spikes.pyx:1184-1215`` slice for slice, ``mask_image`` ``:1124-1127``.
Pinned against the compiled reference through tests/golden/golden_v1.npz (``move_*`` arrays, produced with
``usaccade_image`` on frames where it does not redraw its offsets, and ``mask_image``).
"""
import numpy as np


def move_image(original, delta_x, delta_y, bg_gray):
    moved = bg_gray * np.ones_like(original, dtype=np.int16)
    height, width = original.shape
    if delta_x < 0:
        new_x, old_x = slice(0, delta_x), slice(abs(delta_x), width)
    elif delta_x > 0:
        new_x, old_x = slice(delta_x, width), slice(0, -delta_x)
    else:
        new_x = old_x = slice(0, width)
    if delta_y > 0:
        new_y, old_y = slice(0, -delta_y), slice(delta_y, height)
    elif delta_y < 0:
        new_y, old_y = slice(abs(delta_y), height), slice(0, delta_y)
    else:
        new_y = old_y = slice(0, height)
    moved[new_y, new_x] = original[old_y, old_x]
    return moved


def mask_image(original, mask):
    return np.int16(original * mask)


def traverse_image(original, frame_number, speed, bg_gray):
    """generate_spikes.pyx:1128-1152, slice for slice."""
    moved = bg_gray * np.ones_like(original, dtype=np.int16)
    width = original.shape[1]
    n = int(np.int16(int(np.int16(frame_number) * float(speed))))
    if n == 0:
        n = 1
    if n <= width:
        moved[:, :n] = original[:, -n:]
    elif n < 2 * width:
        n = 2 * width - n + 1
        moved[:, -n:] = original[:, :n]
    return moved


def fade_image(original, frame_number, half_frame, bg_gray):
    """generate_spikes.pyx:1156-1182 (bg_gray is overwritten by the final assignment)."""
    if frame_number < half_frame - 1:
        alpha = float(frame_number) / float(half_frame)
    elif frame_number > half_frame + 1:
        alpha = float(int(np.int16(half_frame * 2 - frame_number))) / float(half_frame)
    else:
        alpha = 1.0
    return np.int16(original * alpha)
