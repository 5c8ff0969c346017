#!/usr/bin/env python
"""The offline conversion flow of the reference's ``mnist_to_spikes.py`` (images -> micro-saccade frames -> DVS ->
time-binary AER -> one text spike file per image), with every stage on the GPU and all images converted at once:

    SaccadeSource (batched micro-saccades)  ->  DVSStreams (fused step, state in HBM)  ->  SpikeFileWriter

    python examples/images_to_spike_files.py --out /tmp/spikes          # synthetic 32x32 digit-like blobs

Each image is one stream; ``frames`` frames are shown per image (``mnist_to_spikes.py:134``), encoded with the
time-binary encoder and adaptive thresholds of BASELINE config 2 and written as ``"<key> <t_ms>"`` lines.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydvs_b200 import DVSConfig, DVSStreams  # noqa: E402
from pydvs_b200.sinks import SpikeFileWriter, spike_file_lines_all_streams  # noqa: E402
from pydvs_b200.frame_sources import SaccadeSource  # noqa: E402


def blob_images(n, size=32, seed=0):
    """Digit-like synthetic images: a few bright strokes on black, 28x28 centred in size x size (mnist_to_spikes.py:239)."""
    rng = np.random.RandomState(seed)
    imgs = np.zeros((n, size, size), dtype=np.int16)
    pad = (size - 28) // 2
    for i in range(n):
        canvas = np.zeros((28, 28), dtype=np.float64)
        for _ in range(3):
            r0, c0 = rng.randint(4, 24, size=2)
            dr, dc = rng.randint(-1, 2, size=2)
            for k in range(rng.randint(6, 14)):
                r, c = r0 + k * dr, c0 + k * dc
                if 0 <= r < 28 and 0 <= c < 28:
                    canvas[max(r - 1, 0):r + 2, max(c - 1, 0):c + 2] = 255
        imgs[i, pad:pad + 28, pad:pad + 28] = canvas.astype(np.int16)
    return imgs


def convert(images, frames=10, fps=100, seed=0, key_bits=(10, 5, 31)):
    """images: int16 [N, H, W] (host or CUDA).  Returns one SpikeFileWriter per image."""
    images = torch.as_tensor(images, dtype=torch.int16, device="cuda")
    n, h, w = images.shape
    frame_ms = int(1000. / fps)
    # the frame loop of mnist_to_spikes.py:246-258 (it calls the animator directly, not VirtualCam.read) for all images
    cam = SaccadeSource(images, np.arange(n, dtype=np.int32), frames_per_image=frames, seed=seed)
    dvs = DVSStreams(DVSConfig(width=w, height=h, streams=n, output_type="TIME_BIN_THR", adaptive_threshold=True,
                               threshold=12, min_threshold=6, max_threshold=168, threshold_delta_down=2,
                               threshold_delta_up=12, history_weight=0.99, max_time_ms=frame_ms, num_bins=6, ref0=0,
                               num_bits_per_spike=2, log2_table=None),
                     event_capacity=n * h * w * 8)
    writers = [SpikeFileWriter(frame_ms, max(1, frame_ms // 6)) for _ in range(n)]
    flag_shift, data_shift, data_mask = key_bits
    for f in range(frames):
        res = dvs.step(cam.frame(f))
        # one key launch + two device-to-host copies for all images of the frame
        lines = spike_file_lines_all_streams(res, flag_shift, data_shift, data_mask, writers[0].t, writers[0].t_bin_ms)
        for s in range(n):
            writers[s].add_lines(lines[s])
    return writers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="spike_files")
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--frames", type=int, default=10)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    writers = convert(blob_images(args.images), frames=args.frames)
    for i, wtr in enumerate(writers):
        wtr.write(os.path.join(args.out, "image_%05d.txt" % i))
    print("wrote %d spike files to %s (%d events)" % (len(writers), args.out, sum(len(w.lines) for w in writers)))


if __name__ == "__main__":
    main()
