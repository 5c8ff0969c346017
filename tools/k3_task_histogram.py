import sys, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydvs_b200 import DVSConfig, DVSStreams, synth
S, H, W = 32, 2160, 3840
cfg = DVSConfig(width=W, height=H, streams=S, output_type="RATE", threshold=12, max_time_ms=8, inh_area_width=2)
dvs = DVSStreams(cfg, event_capacity=S * H * W, split_lists=False)
ring = [synth.moving_bar(S, H, W, t, noise=8, seed=1) for t in range(4)]
for i in range(28):
    dvs.step(ring[i % 4])
for i in range(28, 32):
    res = dvs.step(ring[i % 4])
    torch.cuda.synchronize()
    tc = dvs.tile_counts.view(S, dvs.C, dvs.tps)[:, :2, :].reshape(-1).cpu().numpy()
    ev = res.total_events()
    bins = [0, 1, 5, 9, 17, 33, 65, 129, 100000]
    h, _ = np.histogram(tc, bins=bins)
    print("frame", i % 4, "events", ev, "tasks", tc.size, "records", int(tc.sum()), "hist(0,1-4,5-8,9-16,17-32,33-64,65-128,>128)", h.tolist())
