"""Per-kernel CUDA-event timing of the step over the frame ring (K1 / scans / expand), 64 streams of 4K."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydvs_b200 import DVSConfig, DVSStreams, synth
import bench
wl = bench.WORKLOADS["4k_full_pipeline_rate"]
S=64
cfg = DVSConfig(width=3840, height=2160, streams=S, output_type="RATE", threshold=12, max_time_ms=8, inh_area_width=2)
dvs = DVSStreams(cfg, event_capacity=S*3840*2160)
ring = bench.make_ring(wl, S, 4, "bar", "int16", torch.device("cuda"), 0)
for i in range(32): dvs.step(ring[i%4])
torch.cuda.synchronize()
ev=[[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(8)]
tot=[]
for i in range(8):
    ev[i][0].record(); dvs.tile_pass(ring[i%4]); ev[i][1].record(); dvs.scan(); ev[i][2].record(); dvs.expand(); ev[i][3].record()
    tot.append(dvs.seg_offsets[-1:].clone())
torch.cuda.synchronize()
for i in range(8):
    rec = dvs.stream_totals.view(S, dvs.C)[:, :2].sum().item() if i==7 else -1
    print(i%4, 'K1 %.3f ms  scan %.3f ms  expand %.3f ms  events %d'%(ev[i][0].elapsed_time(ev[i][1]), ev[i][1].elapsed_time(ev[i][2]), ev[i][2].elapsed_time(ev[i][3]), int(tot[i].item())), rec)
