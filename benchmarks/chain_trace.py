#!/usr/bin/env python
"""Phase timeline of the conv5 ... conv6_1 chain launch (OFS_CHAIN_TRACE=1): per layer the GEMM phase, the grid barrier,
the all-CTA reduction and the second barrier, as medians / maxima over the CTAs (globaltimer, ns)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["OFS_CHAIN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402

dev = torch.device("cuda", 0)
B = int(os.environ.get("AB_BATCH", "8"))
lib = ofs.load_library()
lib.ofs_chain_trace_read.restype = C.c_int
net = ofs.FlowNetSPyramid(device=dev, max_batch=B)
net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
x = F.make_feats(3, B).to(dev)
for _ in range(5):
    net.forward(x)
torch.cuda.synchronize()
buf = np.zeros(148 * 32, dtype=np.int64)
n = lib.ofs_chain_trace_read(buf.ctypes.data_as(C.c_void_p), buf.size)
t = buf[:n].reshape(-1, 32)
t0 = t[:, 16].min()
print(f"{n // 32} CTAs; kernel entry skew {t[:, 16].max() - t0} ns")
names = ["gemm done", "barrier", "reduced", "barrier"]
prev = t[:, 16]
for l, lname in enumerate(["conv5", "conv5_1", "conv6", "conv6_1"]):
    for k in range(4):
        cur = t[:, 4 * l + k]
        d = cur - prev
        print(f"{lname:8s} {names[k]:10s} at {np.median(cur - t0) / 1e3:7.2f} us (max {(cur.max() - t0) / 1e3:7.2f});  phase median {np.median(d) / 1e3:6.2f} us, min {d.min() / 1e3:6.2f}, max {d.max() / 1e3:6.2f}")
        prev = cur
print(f"total {(t[:, 15].max() - t0) / 1e3:.2f} us")
