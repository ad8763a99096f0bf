#!/usr/bin/env python
"""BASELINE configs[4]: microbench sweep of the gather kernels, achieved HBM GB/s vs peak.

    python benchmarks/sampler_sweep.py [--quick] [--out gpurun_out/sampler_sweep.json]

ops: tf_warp (staged / direct variants), fused flow_resize_warp, AffineTransformer, ProjectiveTransformer,
transformImage (Lie homography).  Sizes 256x256 .. 2160x3840, batch 1..64 (capped at 8 GB of traffic),
flows: zero / const (3.3,-2.7) / smooth (32x-upsampled N(0,4^2) px) / adversarial U(-32,32).
Algorithmic bytes: tf_warp 32 B/px, fused warp 24 B/px + 1.56 MB/frame, grid samplers 24 B/px.
Timing: CUDA events around each launch, L2 flushed (256 MB write) before every timed launch, median of N.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402

SIZES = [(256, 256), (384, 512), (720, 1280), (1080, 1920), (2160, 3840)]
BATCHES = [1, 2, 4, 8, 16, 32, 64]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def make_flow(kind, B, H, W, dev, gen):
    if kind == "zero":
        return torch.zeros(B, H, W, 2, device=dev)
    if kind == "const":
        f = torch.zeros(B, H, W, 2, device=dev)
        f[..., 0], f[..., 1] = 3.3, -2.7
        return f
    if kind == "smooth":
        lo = torch.randn((B, 2, max(H // 32, 2), max(W // 32, 2)), generator=gen) * 4.0
        return torch.nn.functional.interpolate(lo.to(dev), size=(H, W), mode="bilinear", align_corners=True) \
            .permute(0, 2, 3, 1).contiguous()
    return ((torch.rand((B, H, W, 2), generator=gen) - 0.5) * 64.0).to(dev)


def timed(fn, flush, iters):
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)                      # evict L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


class Cfg:
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sampler_sweep.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(3)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    peak = peaks()
    sizes = [(256, 256), (720, 1280), (2160, 3840)] if args.quick else SIZES
    batches = [1, 8] if args.quick else BATCHES
    iters = 5 if args.quick else 9
    results = []
    for (H, W) in sizes:
        for B in batches:
            if B * H * W * 32 > 8e9 or (B > 8 and H * W > 1080 * 1920):
                continue
            img = torch.rand((B, H, W, 3), device=dev)
            px = B * H * W
            flows = ["zero", "smooth", "adversarial"] if (B in (1, 8)) else ["smooth"]
            for kind in flows:
                flow = make_flow(kind, B, H, W, dev, gen)
                for variant in (3, 2, 1, 0):
                    ofs.set_warp_variant(variant)
                    ofs.tf_warp(img, flow, H, W)
                    ms = timed(lambda: ofs.tf_warp(img, flow, H, W), flush, iters)
                    results.append(dict(op="tf_warp", variant={0: "direct", 1: "staged12", 2: "staged16", 3: "lean"}[variant], flow=kind, B=B, H=H, W=W,
                                        ms=ms, gbs=px * 32 / ms / 1e6))
                ofs.set_warp_variant(3)
                del flow
            # the stand-alone fused op on two flows: white noise per flow pixel (every gather lands in its own sector: the
            # adversarial case) and a smooth field like the network's (what the in-network roofline line measures)
            f2 = torch.randn((B, 382, 510, 2), generator=gen).to(dev) * 2.0
            low = torch.randn((B, 2, 14, 18), generator=gen) * 3.0
            f2s = torch.nn.functional.interpolate(low, size=(382, 510), mode="bicubic", align_corners=False).permute(0, 2, 3, 1).contiguous().to(dev)
            for fl, label in ((f2, "white noise N(0,2^2)"), (f2s, "smooth (network-like)")):
                ofs.flow_resize_warp(img, fl)
                ms = timed(lambda: ofs.flow_resize_warp(img, fl), flush, iters)
                results.append(dict(op="flow_resize_warp", variant="lean", flow=label, B=B, H=H, W=W, ms=ms,
                                    gbs=(px * 24 + B * 382 * 510 * 8) / ms / 1e6))
            c, s_ = np.cos(np.deg2rad(5)) * 1.02, np.sin(np.deg2rad(5)) * 1.02
            th6 = torch.tensor([[c, -s_, 0.01, s_, c, -0.02]] * B, dtype=torch.float32, device=dev)
            aff = ofs.AffineTransformer((H, W))
            aff.transform(img, th6)
            ms = timed(lambda: aff.transform(img, th6), flush, iters)
            results.append(dict(op="AffineTransformer", variant="-", flow="rot5+zoom2%", B=B, H=H, W=W, ms=ms, gbs=px * 24 / ms / 1e6))
            th8 = torch.tensor([[1, 0, 0, 0, 1, 0, 0.02, -0.01]] * B, dtype=torch.float32, device=dev)
            prj = ofs.ProjectiveTransformer((H, W))
            prj.transform(img, th8)
            ms = timed(lambda: prj.transform(img, th8), flush, iters)
            results.append(dict(op="ProjectiveTransformer", variant="-", flow="small homography", B=B, H=H, W=W, ms=ms,
                                gbs=px * 24 / ms / 1e6))
            cfg = Cfg()
            cfg.warpType, cfg.warpApprox, cfg.batch_size, cfg.height, cfg.width = "homography", 4, B, H, W
            cfg.refMtrx = np.array([[(W - 1) / 2.0, 0, (W - 1) / 2.0], [0, (H - 1) / 2.0, (H - 1) / 2.0], [0, 0, 1]], np.float32)
            pm = ofs.vec2mtrx(cfg, (torch.randn((B, 8), generator=gen) * 0.02).to(dev))
            ofs.transformImage(cfg, img, pm)
            ms = timed(lambda: ofs.transformImage(cfg, img, pm), flush, iters)
            results.append(dict(op="transformImage", variant="-", flow="random small p", B=B, H=H, W=W, ms=ms, gbs=px * 24 / ms / 1e6))
            del img, f2
            torch.cuda.empty_cache()
    for r in results:
        r["frac_of_measured_peak"] = r["gbs"] / peak
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"hbm_peak_gbs": peak, "results": results}, f, indent=1)
    print(f"| op | variant | flow | B | HxW | ms | GB/s | of {peak:.0f} |")
    print("|---|---|---|---:|---|---:|---:|---:|")
    for r in results:
        print(f"| {r['op']} | {r['variant']} | {r['flow']} | {r['B']} | {r['H']}x{r['W']} | {r['ms']:.4f} | {r['gbs']:.0f} | "
              f"{100 * r['frac_of_measured_peak']:.1f}% |")


if __name__ == "__main__":
    main()
