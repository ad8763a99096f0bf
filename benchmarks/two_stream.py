#!/usr/bin/env python
"""Does running two independent 8-pair batches on two streams (two net objects) raise whole-GPU throughput?
The step is a chain of 26 kernels, several of which leave SMs idle (wave quantisation, small layers): a second
stream can fill those holes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = 8, 720, 1280
w = F.make_weights(0, "calibrated", head_scale=0.02)
lib = ofs.load_library()
for nstreams in (1, 2, 3):
    nets = [ofs.FlowNetSPyramid(device=dev, max_batch=B) for _ in range(nstreams)]
    for n in nets:
        n.assign_weights(w)
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    sets = [(F.make_feats(10 + i, B).to(dev), torch.rand((B, H, W, 3), device=dev)) for i in range(2 * nstreams)]
    outs = [torch.empty_like(sets[0][1]) for _ in range(2 * nstreams)]

    def step(i):
        k = i % nstreams
        j = i % (2 * nstreams)
        feats, frames = sets[j]
        ofs._lib.check(lib.ofs_net_stabilize(nets[k]._h, ofs._lib.ptr(feats), ofs._lib.ptr(frames), ofs._lib.ptr(outs[j]), None,
                                             B, H, W, streams[k].cuda_stream))

    for i in range(4 * nstreams):
        step(i)
    torch.cuda.synchronize()
    steps = 200
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(nstreams)]
    e0.record(torch.cuda.current_stream())
    for s in streams:
        s.wait_event(e0)
    for i in range(steps):
        step(i)
    for s, e in zip(streams, ends):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in ends)
    print(f"{nstreams} stream(s): {steps} steps in {ms:.2f} ms -> {B * steps / ms * 1e3:.0f} pairs/s ({ms / steps:.4f} ms/step)", flush=True)
    for n in nets:
        n.close()
