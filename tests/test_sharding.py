"""Multi-GPU host logic on CPU: pair-range sharding and the output gather with the gloo
backend, world size 2 -- the same code path bench.py / a split clip uses with NCCL on the GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from coupe.optical_flow_based_deep_video_stabilization_b200.sharding import gather_output, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 9, 1024, 1031):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (l0, h0), (l1, h1) in zip(spans, spans[1:]):
                assert h0 == l1 and l0 <= h0 and l1 <= h1
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1          # balanced: weak scaling, no straggler
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_gather_output_single_process_is_identity():
    x = torch.rand(3, 4, 5, 3)
    assert gather_output(x, 3) is x
    with pytest.raises(ValueError):
        gather_output(x, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_items, rank, world)
        # the "stabilised frames" of this rank's pair range: frame i is filled with the value i
        local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1, 1).expand(hi - lo, 6, 8, 3).contiguous()
        full = gather_output(local, n_items)
        ok = full.shape == (n_items, 6, 8, 3) and bool((full[:, 0, 0, 0] == torch.arange(n_items, dtype=torch.float32)).all())
        # the uint8 payload the video writer consumes (4x fewer bytes than float32), gathered to rank 0 only
        u8 = gather_output(local.to(torch.uint8), n_items, dst=0)
        if rank == 0:
            ok = ok and u8.dtype == torch.uint8 and u8.shape == (n_items, 6, 8, 3) and bool(
                (u8[:, 5, 7, 2] == torch.arange(n_items, dtype=torch.uint8)).all())
        else:
            ok = ok and u8 is None
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, ok and float(t) == float(world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7, 1])
def test_gather_output_world2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]
