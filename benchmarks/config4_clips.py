#!/usr/bin/env python
"""BASELINE configs[3]: 8 x 1024-frame 720p synthetic clips stabilised on N GPUs of one box by frame-batch partition,
the stabilised uint8 frames gathered to rank 0 over NCCL (the only collective of the design).

    python benchmarks/config4_clips.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/config4_clips.py [--frames 1024] [--clips 8]

Rank r owns frames [lo, hi) of EVERY clip (sharding.shard_range) and advances its 8 sub-clips in lockstep as one
batch through the device-side clip driver (ClipStabilizer.submit_device: frames and outputs stay in HBM; each shard
cold-starts its history from its own first frame exactly as main_dl.py:555-556 does for i < offset -- the labelled
approximate mode of a split clip).  Then ONE gather_output(..., dst=0) moves every rank's [frames, clips, H, W, 3]
uint8 block to rank 0.  Prints one JSON line on rank 0: frames/s of the stabilisation (max over ranks, CUDA events),
the gather's GB/s into rank 0 and its share of the run.  Total work is fixed as N grows ("scaling": "strong").
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs  # noqa: E402
from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    H, W, C_, n_clips = args.height, args.width, 3, args.clips
    lo, hi = ofs.shard_range(args.frames, rank, world)
    n_local = hi - lo
    net = ofs.FlowNetSPyramid(device=dev, max_batch=n_clips)
    net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
    stab = ofs.ClipStabilizer(net, n_clips=n_clips, height=H, width=W)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    frames = torch.randint(0, 256, (n_local, n_clips, H, W, C_), dtype=torch.uint8, device=dev, generator=gen)
    out = torch.empty_like(frames)

    def run_range(n):
        for i in range(n):
            if stab.in_flight == stab.depth:
                stab.wait()
            stab.submit_device(frames[i], out[i])
        while stab.in_flight:
            stab.wait()

    run_range(min(8, n_local))      # warm-up: graphs captured, clocks up
    stab.reset()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    run_range(n_local)
    e1.record()
    torch.cuda.synchronize()
    t_stab_wall = time.perf_counter() - t0
    ms_stab = e0.elapsed_time(e1)
    # gather: [n_local, clips, H, W, 3] uint8 blocks, frame-major, to rank 0.  The receive buffer is allocated and the
    # NCCL point-to-point channels are connected (a one-frame gather) BEFORE the timed call: both are one-off costs of a
    # process, not of a clip.
    ms_gather, full = 0.0, out
    if dist is not None:
        nmax = max(h - l for l, h in (ofs.shard_range(args.frames, r, world) for r in range(world)))
        recv = torch.empty((world * nmax, n_clips, H, W, C_), dtype=torch.uint8, device=dev) if rank == 0 else None
        ofs.gather_output(out[:1], world, dst=0)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e1.record()
        full = ofs.gather_output(out, args.frames, dst=0, out=recv)
        e2.record()
        torch.cuda.synchronize()
        ms_gather = e1.elapsed_time(e2)
    t = torch.tensor([ms_stab, ms_gather, t_stab_wall * 1e3], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_stab, ms_gather, ms_wall = (float(v) for v in t.tolist())
    if rank == 0:
        total_frames = args.frames * n_clips
        ok = tuple(full.shape) == (args.frames, n_clips, H, W, C_) and full.dtype == torch.uint8
        if dist is not None:
            ok = ok and torch.equal(full[lo:hi], out)   # rank 0's own block sits where the pair order puts it
        gathered = (args.frames - n_local) * n_clips * H * W * C_          # bytes that crossed NVLink into rank 0
        rec = {"metric": "frames/s (8 x 1024-frame 720p clips, frame-batch partition, NCCL gather of the uint8 output to rank 0)",
               "n_gpus": world, "frames_total": total_frames, "clips": n_clips, "frames_per_clip": args.frames, "frame": [H, W],
               "stabilise_ms": ms_stab, "stabilise_wall_ms": ms_wall, "gather_ms": ms_gather,
               "frames_per_s_stabilise": total_frames / (ms_stab * 1e-3),
               "frames_per_s_with_gather": total_frames / ((ms_stab + ms_gather) * 1e-3),
               "gather_bytes_into_rank0": gathered, "gather_gb_s": (gathered / (ms_gather * 1e-3) / 1e9) if ms_gather > 0 else None,
               "gather_share_of_run": ms_gather / (ms_stab + ms_gather), "payload": "uint8 (2.76 MB per frame)",
               "scaling": "strong", "data": "synthetic uint8 frames resident in HBM; device-side clip driver, 8 sub-clips in lockstep per rank",
               "gathered_ok": bool(ok)}
        print(json.dumps(rec), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(rec, f, indent=1)
    stab.close()
    net.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
