#!/usr/bin/env python
"""Per-layer microbenchmark of the tcgen05 implicit-GEMM conv kernel (needs a B200).

    python benchmarks/conv_bench.py [--batch 8] [--layers 3_1,4_1] [--variants "256:1:1,256:1:2"] [--trace]

Each layer of flownetS_pyramid (reference model.py:807-880) is run stand-alone exactly as the network runs it
(16-bit NHWC activations with the network's channel strides, packed weights, fused bias+lrelu epilogue) on
device-resident pseudo-random operands through the measurement entry ofs_conv2d_bench.  A variant is
block_n:ksplit:cta_group[:debug].  Prints us per launch, TFLOP/s on the LITERAL MACs of the layer, and with
--trace the per-CTA role timeline (SM cycles) of one extra launch.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (H, W, cin, in_cs, cout, out_cs, k, stride, transposed, default variant)
LAYERS = {
    "1": (384, 512, 27, 32, 64, 64, 7, 2, 0),
    "2": (192, 256, 64, 64, 128, 256, 5, 2, 0),
    "3": (96, 128, 128, 256, 256, 256, 5, 2, 0),
    "3_1": (48, 64, 256, 256, 256, 448, 3, 1, 0),
    "4": (48, 64, 256, 448, 512, 512, 3, 2, 0),
    "4_1": (24, 32, 512, 512, 512, 832, 3, 1, 0),
    "5": (24, 32, 512, 832, 512, 512, 3, 2, 0),
    "5_1": (12, 16, 512, 512, 512, 1088, 3, 1, 0),
    "6": (12, 16, 512, 1088, 1024, 1024, 3, 2, 0),
    "6_1": (6, 8, 1024, 1024, 1024, 1024, 3, 1, 0),
    "deconv5": (6, 8, 1024, 1024, 512, 1088, 4, 2, 1),
    "deconv4": (12, 16, 1026, 1088, 256, 832, 4, 2, 1),
    "deconv3": (24, 32, 770, 832, 128, 448, 4, 2, 1),
    "deconv2": (48, 64, 386, 448, 64, 256, 4, 2, 1),
    "predict2": (96, 128, 194, 256, 18, 18, 1, 1, 0),
    # the same layers with TIGHT concat strides (200 / 392 / 776 / 1032 channels: pixel rows not 128-byte aligned, a
    # 64-channel TMA box row straddles two lines) -- what the network used before round 2's alignment change
    "2t": (192, 256, 64, 64, 128, 200, 5, 2, 0),
    "3t": (96, 128, 128, 200, 256, 256, 5, 2, 0),
    "3_1t": (48, 64, 256, 256, 256, 392, 3, 1, 0),
    "4t": (48, 64, 256, 392, 512, 512, 3, 2, 0),
    "4_1t": (24, 32, 512, 512, 512, 776, 3, 1, 0),
    "5t": (24, 32, 512, 776, 512, 512, 3, 2, 0),
    "5_1t": (12, 16, 512, 512, 512, 1032, 3, 1, 0),
    "6t": (12, 16, 512, 1032, 1024, 1024, 3, 2, 0),
    "deconv5t": (6, 8, 1024, 1024, 512, 1032, 4, 2, 1),
    "deconv4t": (12, 16, 1026, 1032, 256, 776, 4, 2, 1),
    "deconv3t": (24, 32, 770, 776, 128, 392, 4, 2, 1),
    "deconv2t": (48, 64, 386, 392, 64, 200, 4, 2, 1),
    "predict2t": (96, 128, 194, 200, 18, 18, 1, 1, 0),
    # proxy for conv1 in space-to-depth form (25 K blocks of a 128-column 1-CTA tile on the 192 x 128 pair grid; the real thing has 35)
    "s2d_proxy": (192, 128, 64, 64, 128, 128, 5, 1, 0),
}
# the network's own tilings (flownet.cu, ofs_net_create): block_n:ksplit:cta_group code
DEFAULTS = {"1": "64:1:4", "2": "128:1:4", "3": "256:1:2", "3_1": "256:1:2", "4": "192:1:1", "4_1": "192:1:1",
            "5": "256:6:1", "5_1": "256:6:1", "6": "128:4:1", "6_1": "128:4:1", "deconv5": "64:1:34",
            "deconv4": "128:1:34", "deconv3": "128:1:36", "deconv2": "64:1:34", "predict2": "32:1:32", "s2d_proxy": "128:1:1"}
DEFAULTS.update({k + "t": v for k, v in DEFAULTS.items() if k + "t" in LAYERS})


def macs(name, B):
    H, W, cin, _, cout, _, k, s, tr = LAYERS[name]
    if tr:
        return B * H * W * 16 * cin * cout
    p = k // 2
    return B * ((H + 2 * p - k) // s + 1) * ((W + 2 * p - k) // s + 1) * k * k * cin * cout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--layers", default=",".join(LAYERS))
    ap.add_argument("--variants", default="", help="comma list of block_n:ksplit:cta_group[:debug]; empty = network default")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--flush-mb", type=int, default=0)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--json", default="")
    args = ap.parse_args()

    import torch

    import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs
    from coupe.optical_flow_based_deep_video_stabilization_b200 import _lib

    lib = ofs.load_library()
    fn = lib.ofs_conv2d_bench
    fn.restype = C.c_int
    fn.argtypes = [C.c_int] * 16 + [C.POINTER(C.c_float), C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    results = []
    for name in args.layers.split(","):
        H, W, cin, in_cs, cout, out_cs, k, s, tr = LAYERS[name]
        variants = args.variants.split(",") if args.variants else [DEFAULTS[name]]
        for v in variants:
            parts = [int(x) for x in v.split(":")]
            bn, ks, cg = parts[0], parts[1], parts[2]
            dbg = parts[3] if len(parts) > 3 else 0
            if cout % bn != 0 and not (bn == 192 and cout % 64 == 0 and not tr) and name != "predict2" and cg != 5:
                continue
            ms = C.c_float(0)
            grid = C.c_int(0)
            cap = 1024 * 32
            trace = (C.c_longlong * cap)()
            rc = fn(args.batch, H, W, cin, in_cs, cout, out_cs, k, s, tr, bn, ks, cg, dbg, args.iters, args.flush_mb,
                    C.byref(ms), C.cast(trace, C.c_void_p) if args.trace else None, cap, C.byref(grid),
                    _lib.current_stream_ptr(dev))
            _lib.check(rc)
            us = ms.value * 1e3
            tf = 2 * macs(name, args.batch) / (ms.value * 1e-3) / 1e12
            line = f"{name:8s} bn={bn:3d} ks={ks} cg={cg} dbg={dbg} grid={grid.value:3d}  {us:8.1f} us  {tf:7.1f} TF/s"
            rec = {"layer": name, "block_n": bn, "ksplit": ks, "cta_group": cg, "debug": dbg, "grid": grid.value,
                   "us": us, "tflops": tf}
            if args.trace:
                g = grid.value
                rows = [[trace[i * 32 + j] for j in range(32)] for i in range(g)]
                prev = [[trace[(g + i) * 32 + j] for j in range(32)] for i in range(g)]
                gap2_ns = min(r[11] for r in rows) - max(r[13] for r in prev)
                t0 = min(r[0] for r in rows)
                span_ns = max(r[13] for r in rows) - t0
                tk0 = min(r[11] for r in rows)
                entry_skew = max(r[11] for r in rows) - tk0
                full_span = max(r[13] for r in rows) - tk0
                import statistics as st

                def med(f):
                    vals = [f(r) for r in rows if f(r) is not None]
                    return st.median(vals) if vals else float("nan")

                def mx(f):
                    vals = [f(r) for r in rows if f(r) is not None]
                    return max(vals) if vals else float("nan")
                start_skew = mx(lambda r: r[0] - t0)
                total = med(lambda r: r[9] - r[1])
                tmax = mx(lambda r: r[9] - r[1])
                prod = med(lambda r: r[2] - r[1] if r[2] else None)
                first_full = med(lambda r: r[3] - r[1] if r[3] else None)
                mma_done = med(lambda r: r[4] - r[1] if r[4] else None)
                epi_first = med(lambda r: r[5] - r[1] if r[5] else None)
                epi_done = med(lambda r: r[6] - r[1] if r[6] else None)
                prologue = med(lambda r: r[1] - r[12])
                epi_loop = med(lambda r: r[10] - r[1] if r[10] else None)
                mhz = med(lambda r: (r[9] - r[12]) / max(r[7] - r[11], 1) * 1e3)
                line += (f"\n           previous launch end -> this launch first CTA entry: {gap2_ns / 1e3:.1f} us"
                         f"\n           SM clock during the traced launch {mhz:.0f} MHz; entry->end {full_span / 1e3:.1f} us, entry skew {entry_skew / 1e3:.1f} us, prologue {prologue:.0f} cyc, "
                         f"epi loop end {epi_loop:.0f}")
                nch = med(lambda r: r[21])
                if nch and nch == nch and nch > 0:
                    line += ("\n           epilogue per 64-col chunk (cycles): ld+math+sts %.0f | tmem release+fence %.0f | wait_read %.0f | bar %.0f | issue %.0f  (%d chunks)"
                             % tuple([med(lambda r, k=k: r[k] / max(r[21], 1)) for k in (16, 17, 18, 19, 20)] + [nch]))
                line += ("\n           per stage item (median CTA, %d items): MMA warp blocked on operands %.0f cyc, on a free accumulator %.0f cyc per item; "
                         "producer blocked on a free stage %.0f cyc per item"
                         % (med(lambda r: r[23]), med(lambda r: r[22] / max(r[23], 1)), med(lambda r: r[25] / max(r[23], 1)),
                            med(lambda r: r[24] / max(r[23], 1))))
                line += (f"\n           trace: span {span_ns / 1e3:.1f} us, start skew {start_skew / 1e3:.1f} us; cycles (median CTA): "
                         f"total {total:.0f} (max {tmax:.0f}) | producer done {prod:.0f} | first full {first_full:.0f} | "
                         f"mma issued {mma_done:.0f} | epi first {epi_first:.0f} | epi done {epi_done:.0f}")
                if dbg & 256:
                    fine = [trace[g * 64 + j] for j in range(4096)]
                    n_it = 0
                    while n_it < 1024 and fine[n_it * 4 + 3]:
                        n_it += 1
                    pw, pi, mw, mc = ([fine[i * 4 + k] for i in range(n_it)] for k in range(4))
                    lo, hi = min(8, n_it // 4), n_it
                    def md(vals):
                        vals = sorted(vals)
                        return vals[len(vals) // 2] if vals else float("nan")
                    line += ("\n           fine (CTA 0, %d items; medians over items %d..): producer wait-done -> loads issued %.0f | loads issued -> next wait-done %.0f"
                             " || MMA wait-done -> committed %.0f | committed -> next wait-done %.0f"
                             " || loads issued -> MMA sees full %.0f | period (MMA wait-done to wait-done) %.0f"
                             % (n_it, lo, md([pi[i] - pw[i] for i in range(lo, hi)]), md([pw[i + 1] - pi[i] for i in range(lo, hi - 1)]),
                                md([mc[i] - mw[i] for i in range(lo, hi)]), md([mw[i + 1] - mc[i] for i in range(lo, hi - 1)]),
                                md([mw[i] - pi[i] for i in range(lo, hi)]), md([mw[i + 1] - mw[i] for i in range(lo, hi - 1)])))
                    line += "\n           fine first 12 items (producer wait-done, issued, MMA wait-done, committed; cycles from the first stamp): " + " ".join(
                        "[%d %d %d %d]" % (pw[i] - pw[0], pi[i] - pw[0], mw[i] - pw[0], mc[i] - pw[0]) for i in range(min(12, n_it)))
                rec["trace"] = {"span_us": span_ns / 1e3, "start_skew_us": start_skew / 1e3, "total": total, "total_max": tmax,
                                "producer_done": prod, "first_full": first_full, "mma_issued": mma_done,
                                "epi_first": epi_first, "epi_done": epi_done}
            print(line, flush=True)
            results.append(rec)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
