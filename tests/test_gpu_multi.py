"""Multi-GPU: clips sharded over the GPUs of one box (one process per GPU, weights replicated, no collective on the
inference path) and the stabilised uint8 frames collected over NCCL (BASELINE configs[3]).
Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

H, W, N_FRAMES = 96, 128, 12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _clip(seed):
    return torch.from_numpy(np.random.default_rng(seed).integers(0, 256, (N_FRAMES, 1, H, W, 3), dtype=np.uint8))


def _run_clip(ofs, dev, weights, clip):
    """One clip through the device-side clip driver, frames and outputs resident on the GPU."""
    net = ofs.FlowNetSPyramid(device=dev, max_batch=1)
    net.assign_weights(weights)
    stab = ofs.ClipStabilizer(net, n_clips=1, height=H, width=W)
    frames = clip.to(dev)
    out = torch.empty((N_FRAMES, 1, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(N_FRAMES):
        if stab.in_flight == stab.depth:
            stab.wait()
        stab.submit_device(frames[i], out[i])
    while stab.in_flight:
        stab.wait()
    torch.cuda.synchronize(dev)
    stab.close()
    net.close()
    return out[:, 0]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs
        from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic

        weights = synthetic.make_weights(0, "calibrated", head_scale=0.02)
        local = _run_clip(ofs, dev, weights, _clip(100 + rank))                    # this rank's clip
        full = ofs.gather_output(local, world * N_FRAMES, dst=0)                   # NCCL, uint8 payload, to rank 0
        every = ofs.gather_output(local, world * N_FRAMES)                         # and the all-gather form
        ok = True
        if rank == 0:
            ok = full.dtype == torch.uint8 and tuple(full.shape) == (world * N_FRAMES, H, W, 3)
            for r in range(world):                                                 # rank 0 replays every clip itself
                want = _run_clip(ofs, dev, weights, _clip(100 + r))
                ok = ok and torch.equal(full[r * N_FRAMES:(r + 1) * N_FRAMES], want)
            ok = ok and torch.equal(every, full)
        else:
            ok = full is None and tuple(every.shape) == (world * N_FRAMES, H, W, 3)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_clips_sharded_over_gpus_and_gathered_over_nccl():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]
