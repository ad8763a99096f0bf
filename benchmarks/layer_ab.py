#!/usr/bin/env python
"""Per-layer A/B of the network's GEMM tilings (OFS_TUNE / other OFS_* switches set by the caller): prints one compact
line -- dense-set time (graph replay), its burst-roofline fraction, pairs/s at 1 and 2 streams and every kernel's time."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = int(os.environ.get("AB_BATCH", "8")), 720, 1280
tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("OFS_TUNE", "default")
w = F.make_weights(0, "calibrated", head_scale=0.02)
lib = ofs.load_library()
nets = [ofs.FlowNetSPyramid(device=dev, max_batch=B) for _ in range(2)]
for n in nets:
    n.assign_weights(w)
sets = [(F.make_feats(10 + i, B).to(dev), torch.rand((B, H, W, 3), device=dev)) for i in range(4)]
outs = [torch.empty_like(sets[0][1]) for _ in range(4)]
res = {"tag": tag}
for nstreams in (1, 2):
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]

    def step(i):
        k, j = i % nstreams, i % 4
        feats, frames = sets[j]
        ofs._lib.check(lib.ofs_net_stabilize(nets[k]._h, ofs._lib.ptr(feats), ofs._lib.ptr(frames), ofs._lib.ptr(outs[j]), None,
                                             B, H, W, streams[k].cuda_stream))

    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    steps = 100
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
    res[f"pairs_s_{nstreams}"] = round(B * steps / ms * 1e3)
gemm_ms, macs, nl = nets[0].time_kernels("dense", B, iters=20)
res["dense_ms"] = round(gemm_ms, 4)
res["frac_burst"] = round(2 * macs / (gemm_ms * 1e-3) / 1e12 / 1634.7, 4)
prof = nets[0].profile(sets[0][0], sets[0][1], iters=10)
res["us"] = {n.replace("gemm:", ""): round(ms * 1e3, 1) for n, ms, m in prof}
print(json.dumps(res), flush=True)
