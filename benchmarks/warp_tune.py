#!/usr/bin/env python
"""A/B of the fused flow-resize + warp kernel's launch shapes (rows per warp x resident blocks per SM) on the network's
own flow: the kernel replayed from one CUDA graph (ofs_net_time_kernels), 8 x 720p, as bench.py's roofline_warp does."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = 8, 720, 1280
net = ofs.FlowNetSPyramid(device=dev, max_batch=B)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
feats, frames = F.make_feats(11, B).to(dev), torch.rand((B, H, W, 3), device=dev)
ref = net.stabilize(feats, frames).clone()
names = {0: "2 rows, 8 blocks (shipped)", 1: "4 rows, 6", 2: "2 rows, 5", 3: "2 rows, 7", 4: "4 rows, 5", 5: "4 rows, 7", 6: "2 rows, 6"}
for rep in range(2):
    for t in range(7):
        ofs.set_warp_variant(3 + 16 * t)
        ms, _, _ = net.time_kernels("warp", B, frames=frames, iters=40)
        same = torch.equal(net.stabilize(feats, frames), ref) if rep == 0 else True
        gbs = (B * H * W * 24 + B * 382 * 510 * 8) / ms / 1e6
        print(f"{names[t]:28s} {ms * 1e3:7.2f} us  {gbs:6.0f} GB/s  {gbs / 6542.4:.3f} of peak  identical={same}", flush=True)
ofs.set_warp_variant(3)
