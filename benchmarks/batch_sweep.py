#!/usr/bin/env python
"""Device-resident step time (forward + flow glue + warp at 720p) for batch 1..16, with the per-kernel breakdown at batch 1."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
net = ofs.FlowNetSPyramid(device=dev, max_batch=16)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
for B in (1, 2, 4, 8, 16):
    feats = F.make_feats(1, B).to(dev)
    frames = torch.rand((B, 720, 1280, 3), device=dev)
    for _ in range(5):
        net.stabilize(feats, frames)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        net.stabilize(feats, frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    dense_ms, macs, _ = net.time_kernels("dense", B, iters=20)
    print(f"B={B:2d}: {ms:7.3f} ms/step  {B / ms * 1e3:8.0f} pairs/s   dense GEMMs {dense_ms:6.3f} ms = {2 * macs / dense_ms / 1e9:6.0f} TFLOP/s", flush=True)
    if B == 1:
        for name, t, m in net.profile(feats, frames, iters=5):
            print(f"      {name:26s} {t * 1e3:7.1f} us")
