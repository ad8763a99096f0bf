#!/usr/bin/env python
"""Frames/s of the device-side clip driver (one iteration of main_dl.py:540-630 per clip per step), pinned uint8
host buffers, for several clip counts / frame sizes; n_clips = 1 is the faithful autoregressive single-clip loop."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
net = ofs.FlowNetSPyramid(device=dev, max_batch=16)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
rows = []
for (H, W) in ((720, 1280), (1080, 1920)):
    for n in (1, 2, 4, 8, 16):
        stab = ofs.ClipStabilizer(net, n_clips=n, height=H, width=W)
        fin, fout = stab.pinned_buffer(), stab.pinned_buffer()
        fin[...] = np.random.default_rng(n).integers(0, 256, fin.shape, dtype=np.uint8)
        for _ in range(5):
            stab.step(fin, out=fout)
        steps = 40
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            stab.step(fin, out=fout)
        dt = time.perf_counter() - t0
        depth = stab.depth
        ins = [fin] + [stab.pinned_buffer() for _ in range(depth - 1)]
        outs = [fout] + [stab.pinned_buffer() for _ in range(depth - 1)]
        for b in ins[1:]:
            b[...] = fin[:, ::-1]
        stab.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            if stab.in_flight == depth:
                stab.wait()
            stab.submit(ins[i % depth], out=outs[i % depth])
        while stab.in_flight:
            stab.wait()
        dtp = time.perf_counter() - t0
        rows.append({"frame": [H, W], "n_clips": n, "ms_per_step": 1e3 * dt / steps, "frames_per_s": n * steps / dt,
                     "pipelined_ms_per_step": 1e3 * dtp / steps, "pipelined_frames_per_s": n * steps / dtp})
        print(f"{H}x{W}  clips {n:2d}: step() {1e3 * dt / steps:7.3f} ms/step {n * steps / dt:8.1f} frames/s | "
              f"submit/wait {1e3 * dtp / steps:7.3f} ms/step {n * steps / dtp:8.1f} frames/s", flush=True)
        stab.close()
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "clip_bench.json"), "w"), indent=1)
