#!/usr/bin/env python
"""Per-kernel device time of one step at a given batch (default 1, the reference's test mode): where a single
autoregressive clip spends its 0.3 ms.   python benchmarks/batch1_profile.py [batch] [H] [W]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = int(sys.argv[2]) if len(sys.argv) > 2 else 720
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
dev = torch.device("cuda", 0)
net = ofs.FlowNetSPyramid(device=dev, max_batch=B)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
feats = torch.rand((B, 384, 512, 27), device=dev)
frames = torch.rand((B, H, W, 3), device=dev)
rows = net.profile(feats, frames, iters=10)
tot = sum(ms for _, ms, _ in rows)
for name, ms, macs in rows:
    print(f"{name:28s} {ms * 1e3:8.1f} us" + (f"  {2 * macs / (ms * 1e-3) / 1e12:7.1f} TF/s" if macs else ""))
print(f"sum of kernels {tot * 1e3:.1f} us")
out = net.stabilize(feats, frames)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    out = net.stabilize(feats, frames)
torch.cuda.synchronize()
print(f"graph replay: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per step at batch {B}")
