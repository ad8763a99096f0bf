#!/usr/bin/env python
"""A few steps of one clip set (for `ncu --metrics gpu__time_duration.sum`): python benchmarks/clip_one.py [n_clips] [steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
net = ofs.FlowNetSPyramid(device=dev, max_batch=n)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
stab = ofs.ClipStabilizer(net, n_clips=n, height=720, width=1280)
fin, fout = stab.pinned_buffer(), stab.pinned_buffer()
fin[...] = np.random.default_rng(n).integers(0, 256, fin.shape, dtype=np.uint8)
for _ in range(steps):
    stab.step(fin, out=fout)
stab.close()
