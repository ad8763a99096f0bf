#!/usr/bin/env python
"""bench.py -- frame-pairs/sec of the stabiliser hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frame pairs: BASELINE.json configs[1]
= FlowNetS-pyramid forward + test-mode flow glue + dense warp (the sess.run of
main_flownetS_pyramid_noprevloss_dataloader.py:569) on 8 independent 720p frame pairs, bf16 operands,
random-init weights, synthetic frames.  Rank 0 prints ONE JSON line (see the keys below).

    value        whole-job pairs/s, inputs resident in HBM, CUDA-event timed, max over ranks
    e2e          same metric through the C-ABI host-buffer call (ofs_net_stabilize_host): pinned host
                 inputs, H2D + compute + D2H inside the timed region
    roofline     dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs of the 14 dense layers /
                 their launch time (replayed from one CUDA graph as in the step, CUDA events), vs MEASURED_PEAKS.json
    breakdown    per-launch times with an event after every launch (diagnostic: includes launch gaps)
    roofline_warp  the HBM-bound fused flow-resize + warp kernel
    cpu_baseline the CPU oracle (a port of the reference arithmetic; TensorFlow 1.10 is not installable)
                 timed on this box's host cores on a bounded sample
    --impl reference  times that CPU port alone, on all host threads, same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frame-pairs/sec (FlowNetS fwd+warp) @720p"
UNIT = "pairs/s"
BATCH, FRAME_H, FRAME_W = 8, 720, 1280
DENSE_GFLOP_PER_PAIR = 37.89      # 10 encoder convs + 4 transposed convs, literal MACs x 2
WARP_BYTES_PER_PX = 24.0          # fused flow-resize+warp: image 12 + out 12 (+1.56 MB of flow2 per frame)
FLOW2_BYTES = 382 * 510 * 2 * 4


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (set) of the dominant kernels, parsed from the newest
    committed profiles/rNN_ncu_traffic.json (written by benchmarks/ncu_summarize.py from an `ncu --set full` capture of
    this very command).  No file -> no claim (traffic: null)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        t = json.load(f)
    t["source"] = os.path.relpath(files[-1], ROOT)
    return t


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "reasons": reasons, "samples": len(self.samples)}


def synth_inputs(torch, seed, batch, h, w, device=None, pinned=False):
    from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F   # shared by both arms

    g = torch.Generator().manual_seed(seed)
    feats = F.make_feats(seed, batch)
    frames = torch.rand((batch, h, w, 3), generator=g)
    if device is not None:
        return feats.to(device), frames.to(device)
    if pinned:
        return feats.pin_memory(), frames.pin_memory()
    return feats, frames


def cpu_oracle_rate(torch, n_pairs, h, w, threads):
    """pairs/s of the CPU port (oracle) of the same step: forward_literal + flow glue + tf_warp."""
    from oracle import flownet as F
    from oracle import samplers as S

    torch.set_num_threads(threads)
    wts = F.make_weights(0, "calibrated", head_scale=0.02)
    feats, frames = synth_inputs(torch, 7, 1, h, w)
    out = F.forward_literal(feats, wts)                       # warm-up
    S.flow_resize_warp(frames, out["predict_flow2"], h, w)
    t0 = time.perf_counter()
    for _ in range(n_pairs):
        out = F.forward_literal(feats, wts)
        S.flow_resize_warp(frames, out["predict_flow2"], h, w)
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt


def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (TensorFlow 1.10 cannot be installed here),
    on all host threads, on the SAME workload as the GPU arm: every step is one full batch of 8 distinct 720p frame
    pairs (forward + flow glue + tf_warp each), two input sets alternating between steps."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import flownet as F
    from oracle import samplers as S

    wts = F.make_weights(0, "calibrated", head_scale=0.02)
    sets = [synth_inputs(torch, 100 + i, BATCH, FRAME_H, FRAME_W) for i in range(2)]

    def step(i):
        feats, frames = sets[i % 2]
        for b in range(BATCH):                                # the reference runs batch 1 per sess.run (main_dl.py:491)
            o = F.forward_literal(feats[b:b + 1], wts)
            S.flow_resize_warp(frames[b:b + 1], o["predict_flow2"], FRAME_H, FRAME_W)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = BATCH * args.steps / dt
    sample = (f"all {BATCH} pairs of every step ({args.steps} steps, 2 alternating input sets of {BATCH} distinct 720p pairs), "
              f"batch 1 per call as the reference does, fp32 torch-CPU oracle, {threads} torch threads, {dt:.1f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU port (oracle/) of the TF-1.10 graph; TensorFlow/TensorLayer cannot be installed offline",
    }), flush=True)


def workload_name():
    return (f"BASELINE configs[1]: FlowNetS-pyramid fwd (384x512x27) + flow glue + tf_warp at {FRAME_H}x{FRAME_W}, "
            f"batch {BATCH} independent frame pairs per GPU, random-init weights")


def workload_config():
    """The `config` object: identical for both arms (how each arm EXECUTES the workload is reported outside it)."""
    return {"workload": workload_name(), "pairs_per_step_per_gpu": BATCH, "frame": [FRAME_H, FRAME_W],
            "net_input": [384, 512, 27],
            "l2": "every step reads a fresh 258 MB input set (> 126 MB L2); consecutive steps alternate between input sets"}


def run_ours(args):
    import torch

    import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs
    from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as F   # random-init weights, synthetic inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = ofs.load_library()
    peaks = load_peaks()

    weights = F.make_weights(0, "calibrated", head_scale=0.02)
    # Frame-pair batches are independent, and the step is a chain of 26 kernels several of which leave SMs idle (192
    # tiles on 148 SMs, the M <= 1536 layers): `--streams` batches are kept in flight on as many CUDA streams, each
    # with its own net instance (weights + activation workspace), and fill those holes (+17 % at 2 streams).
    nstreams = max(1, args.streams)
    nets = [ofs.FlowNetSPyramid(device=dev, max_batch=BATCH, precision=args.precision) for _ in range(nstreams)]
    for n_ in nets:
        n_.assign_weights(weights)
    net = nets[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    # 2 input sets per stream alternate so no step re-reads what an earlier one left in the 126 MB L2
    nsets = 2 * nstreams
    sets = [synth_inputs(torch, 100 + rank * nsets + i, BATCH, FRAME_H, FRAME_W, device=dev) for i in range(nsets)]
    outs = [torch.empty_like(sets[0][1]) for _ in range(nsets)]

    def step(i):
        k, j = i % nstreams, i % nsets
        feats, frames = sets[j]
        ofs._lib.check(lib.ofs_net_stabilize(nets[k]._h, ofs._lib.ptr(feats), ofs._lib.ptr(frames), ofs._lib.ptr(outs[j]),
                                             None, BATCH, FRAME_H, FRAME_W, streams[k].cuda_stream))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(nsteps):
        """ms from one event on the current stream (every step stream waits for it) to the LAST step stream's end."""
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
        e0.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_event(e0)
        for i in range(nsteps):
            step(i)
        for s, e in zip(streams, ends):
            e.record(s)
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e) for e in ends)

    for i in range(max(args.warmup, 3) * nstreams):
        step(i)
    barrier()
    # ---- the two roofline kernels timed ALONE and FIRST, on a GPU that has only run the warm-up steps: the tensor
    # denominator is the BURST bf16 peak (MEASURED_PEAKS: a kernel timed alone from idle), so the numerator is taken in
    # the same state -- after the K-step region and the sustained block the SM clock sits at the 1 kW power cap
    # (~1.68 GHz against 1.965) and the same launch set reads ~10 % slower; that second reading is reported beside it,
    # against the SUSTAINED peak
    burst = None
    if rank == 0:
        time.sleep(0.5)
        g_ms, g_macs, g_launches = net.time_kernels("dense", BATCH, iters=20)
        w_ms, _, _ = net.time_kernels("warp", BATCH, frames=sets[0][1], iters=20)
        burst = (g_ms, g_macs, g_launches, w_ms)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = lib.ofs_launch_count()
    ms_local = timed_steps(args.steps)
    barrier()
    launches = int(lib.ofs_launch_count() - l0)
    ms_total = torch.tensor([ms_local], device=dev)
    if dist is not None:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    # the same K steps strictly one after the other on one stream (what a single synchronous caller gets)
    single_ms = None
    if nstreams > 1:
        nstreams_saved, nstreams = nstreams, 1
        streams_saved, streams = streams, streams[:1]
        barrier()
        single_ms = timed_steps(args.steps)
        nstreams, streams = nstreams_saved, streams_saved
        barrier()

    # ---- sustained: the same steps looped for >= args.sustained_seconds (the K-step region above is ~10 ms at boost
    # clocks; MEASURED_PEAKS shows a B200 settling near 1.3 GHz under seconds of dense GEMM load)
    sustained = None
    if args.sustained_seconds > 0:
        s_sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        if s_sampler:
            s_sampler.start()
        chunk = 40 * nstreams
        s_ms, s_steps, t_wall = 0.0, 0, time.perf_counter()
        while time.perf_counter() - t_wall < args.sustained_seconds:
            s_ms += timed_steps(chunk)
            s_steps += chunk
        barrier()
        s_clocks = s_sampler.stop() if s_sampler else None
        t_s = torch.tensor([s_ms], device=dev)
        if dist is not None:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        sustained = {"value": world * BATCH * s_steps / (float(t_s.item()) * 1e-3), "unit": UNIT, "steps": s_steps,
                     "seconds": float(t_s.item()) * 1e-3, "ms_per_step": float(t_s.item()) / s_steps, "clocks": s_clocks,
                     "how": f"chunks of {chunk} steps, {nstreams} in flight, CUDA-event timed back to back until "
                            f">= {args.sustained_seconds} s of wall time; nvidia-smi sampled every 0.2 s meanwhile"}

    # ---- per-kernel breakdown, measured live with CUDA events on the launching stream (rank 0)
    roofline = roofline_warp = breakdown = None
    if rank == 0:
        prof = net.profile(sets[0][0], sets[0][1], iters=max(3, min(args.steps, 10)))
        step_ms = (single_ms if single_ms else ms_total) / args.steps   # shares refer to one step executed alone
        # dominant kernel = the tcgen05 implicit-GEMM conv: its 14 launches per step (plus split-K reductions) are
        # replayed from one CUDA graph, exactly as the step runs them, between two CUDA events on the net's stream
        # ... and once more now, after the K-step region and the sustained block (hot, at the power cap)
        hot_ms, macs, gemm_launches = net.time_kernels("dense", BATCH, iters=20)
        hot_wms, _, _ = net.time_kernels("warp", BATCH, frames=sets[0][1], iters=20)
        gemm_ms, _, _, wms = burst
        flops = 2.0 * macs
        ach = flops / (gemm_ms * 1e-3) / 1e12
        hot_ach = flops / (hot_ms * 1e-3) / 1e12
        ncu = load_ncu_traffic()
        roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops"], "traffic": ncu.get("dense_set_bytes") if ncu else None,
                    "traffic_source": (ncu["source"] + " (ncu --set full: DRAM bytes of the launch set)") if ncu else None,
                    "kernel": "conv_gemm*_kernel (the 14 dense conv / transposed-conv layers of one step, incl. split-K reductions)",
                    "launches_per_step": gemm_launches, "ms_per_step_in_kernel": gemm_ms, "share_of_step": hot_ms / step_ms,
                    "algorithmic_gflop_per_launch_set": flops / 1e9,
                    "peak_source": peaks["source"] + " (burst bf16; the launch set timed alone right after the warm-up steps, before any long timed region)",
                    "after_sustained_load": {"ms_per_step_in_kernel": hot_ms, "achieved": hot_ach, "peak_sustained": peaks["bf16_tflops_sustained"],
                                             "frac_of_sustained_peak": hot_ach / peaks["bf16_tflops_sustained"],
                                             "frac_of_burst_peak": hot_ach / peaks["bf16_tflops"],
                                             "note": "the same launch set timed again after the K-step region and the sustained block (SM clock at the power cap)"},
                    "frac_of_sustained_peak": hot_ach / peaks["bf16_tflops_sustained"], "peak_sustained": peaks["bf16_tflops_sustained"],
                    "frac_of_nominal_2250": ach / 2250.0,
                    "timing": "20 repetitions of the launch set replayed from one CUDA graph, CUDA events on the launching stream"}
        wbytes = BATCH * (FRAME_H * FRAME_W * WARP_BYTES_PER_PX + FLOW2_BYTES)
        wach = wbytes / (wms * 1e-3) / 1e9
        roofline_warp = {"bound": "hbm", "achieved": wach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": wach / peaks["hbm_gbs"], "traffic": ncu.get("warp_bytes") if ncu else None,
                         "kernel": "warp5_kernel<true> (fused flow-resize + tf_warp)",
                         "ms_per_launch": wms, "algorithmic_bytes_per_launch": wbytes, "frac_of_nominal_7700": wach / 7700.0,
                         "share_of_step": hot_wms / step_ms, "peak_source": peaks["source"] + " (timed alone right after the warm-up steps)",
                         "after_sustained_load": {"ms_per_launch": hot_wms, "achieved": wbytes / (hot_wms * 1e-3) / 1e9,
                                                  "frac": wbytes / (hot_wms * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
        breakdown = [{"kernel": n, "ms": round(ms, 4), "tflops": (2 * m / (ms * 1e-3) / 1e12 if m else None)} for n, ms, m in prof]

    # ---- e2e: the C-ABI host-buffer call, pinned host inputs, H2D + compute + D2H every step
    # `--streams` caller threads, each with its own net and pinned buffers, call the synchronous C-ABI entry point
    # concurrently (ctypes drops the GIL): one call's D2H tail overlaps the next call's H2D
    host_sets = []
    for k in range(nstreams):
        hf, hfr = synth_inputs(torch, 300 + rank * nstreams + k, BATCH, FRAME_H, FRAME_W, pinned=True)
        host_sets.append((hf, hfr, torch.empty_like(hfr).pin_memory()))
    hf, hfr, hout = host_sets[0]
    e2e_steps = max(4, min(args.steps, 12))
    e2e_steps -= e2e_steps % nstreams

    def e2e_worker(k, nsteps):
        torch.cuda.set_device(local)
        f_, fr_, o_ = host_sets[k]
        for _ in range(nsteps):
            nets[k].stabilize_host(f_, fr_, o_)

    def e2e_run(nsteps_total):
        if nstreams == 1:
            e2e_worker(0, nsteps_total)
            return
        ts = [threading.Thread(target=e2e_worker, args=(k, nsteps_total // nstreams)) for k in range(nstreams)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    e2e_run(2 * nstreams)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(dt.item())
    # ---- the same loop through the device-side clip driver: uint8 frames in and out, history ring
    # on the device -- BATCH clips in lockstep, one reference loop iteration per step
    import numpy as np
    clip_steps = max(6, min(args.steps, 40))
    clip_steps -= clip_steps % nstreams
    stabs = [ofs.ClipStabilizer(nets[k], n_clips=BATCH, height=FRAME_H, width=FRAME_W) for k in range(nstreams)]
    clip_bufs = []
    for k, stab in enumerate(stabs):
        u8, u8_out = [stab.pinned_buffer() for _ in range(stab.depth)], [stab.pinned_buffer() for _ in range(stab.depth)]
        u8[0][...] = np.random.default_rng(7 + rank * nstreams + k).integers(0, 256, u8[0].shape, dtype=np.uint8)
        for b in u8[1:]:
            b[...] = u8[0][:, ::-1]
        clip_bufs.append((u8, u8_out))

    def clip_worker(k, nsteps):
        torch.cuda.set_device(local)
        u8_, out_ = clip_bufs[k]
        depth = stabs[k].depth
        for i in range(nsteps):          # steps in flight: upload i+1 / download i-1 overlap the kernels of i
            if stabs[k].in_flight == depth:
                stabs[k].wait()
            stabs[k].submit(u8_[i % depth], out=out_[i % depth])
        while stabs[k].in_flight:
            stabs[k].wait()

    def clip_run(nsteps_total):
        if nstreams == 1:
            clip_worker(0, nsteps_total)
            return
        ts = [threading.Thread(target=clip_worker, args=(k, nsteps_total // nstreams)) for k in range(nstreams)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    clip_run(3 * nstreams)
    barrier()
    t0 = time.perf_counter()
    clip_run(clip_steps)
    barrier()
    dtc = torch.tensor([time.perf_counter() - t0], device=dev)
    if dist is not None:
        dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
    clip_value = world * BATCH * clip_steps / float(dtc.item())
    for stab in stabs:
        stab.close()
    u8 = clip_bufs[0][0][0]
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_pairs = 128                                   # ~10-15 s of CPU work on a 16-24 core host
            rate, secs = cpu_oracle_rate(torch, n_pairs, FRAME_H, FRAME_W, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{n_pairs} 720p frame pairs (batch 1 each) of the same step, fp32 torch-CPU oracle, {secs:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(),
            "execution": {"steps_in_flight": nstreams, "input_sets": nsets,
                          "parallelism": f"replicas x{world}, no data-path collective; {nstreams} independent batches in flight per GPU "
                                         f"on {nstreams} CUDA streams"},
            "sustained": sustained,
            "value_one_step_at_a_time": (world * BATCH * args.steps / (single_ms * 1e-3)) if single_ms else None,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(lib.ofs_net_host_h2d_bytes(net._h, BATCH, FRAME_H, FRAME_W)),
                    "d2h_bytes_per_step": int(hout.numel() * 4), "steps": e2e_steps,
                    "host_buffer_bytes_per_step": int(hf.numel() * 4 + hfr.numel() * 4),
                    "api": f"ofs_net_stabilize_host (C ABI, pinned host float32 buffers in and out), {nstreams} concurrent caller thread(s)"},
            "e2e_clip_driver": {"value": clip_value, "unit": UNIT, "h2d_bytes_per_step": int(u8.nbytes), "d2h_bytes_per_step": int(u8.nbytes),
                                "steps": clip_steps, "api": "ofs_clips_submit_host / ofs_clips_wait (uint8 BGR frames in / out, device-side history ring, up to 3 steps in flight; "
                                f"one iteration of main_dl.py:540-630 per clip per step, pinned host buffers), {nstreams} clip set(s) of {BATCH} "
                                       "stepped by concurrent caller threads"},
            "gpu_launches": launches, "launches_per_step": launches // max(args.steps, 1),
            "clocks": clocks, "roofline": roofline, "roofline_warp": roofline_warp, "cpu_baseline": cpu,
            "breakdown": breakdown, "lib": os.path.relpath(ofs.lib_path(), ROOT),
        }
        print(json.dumps(line), flush=True)
    for n_ in nets:
        n_.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--precision", choices=["bf16", "fp16"], default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=3.0, help="0 skips the sustained-clock block")
    ap.add_argument("--streams", type=int, default=2, help="independent batches in flight per GPU (1 = strictly sequential steps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
