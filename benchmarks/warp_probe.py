#!/usr/bin/env python
"""Runs each warp / sampler kernel a few times at 8 x 720p (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = 8, 720, 1280
g = torch.Generator().manual_seed(3)
img = torch.rand((B, H, W, 3), generator=g).to(dev)
low = torch.randn((B, H // 32 + 2, W // 32 + 2, 2), generator=g) * 4.0
flow = torch.nn.functional.interpolate(low.permute(0, 3, 1, 2), size=(H, W), mode="bilinear").permute(0, 2, 3, 1).contiguous().to(dev)
f2 = (torch.randn((B, 382, 510, 2), generator=g) * 2.0).to(dev)
for variant in (0, 3):
    ofs.set_warp_variant(variant)
    for _ in range(2):
        ofs.tf_warp(img, flow, H, W)
        ofs.flow_resize_warp(img, f2)
ofs.set_warp_variant(3)
th = torch.tensor([[0.98, 0.087, 0.01, -0.087, 0.98, 0.0]]).repeat(B, 1).to(dev)
for _ in range(2):
    ofs.AffineTransformer((H, W)).transform(img, th)
torch.cuda.synchronize()
print("done")
