#!/usr/bin/env python
"""Turns the scratch ncu outputs of tools/gpu_profile_r2.sh (gpurun_out/launches.csv, prof_conv.ncu-rep, prof_misc.ncu-rep)
into the tracked summaries under profiles/ -- launch shares, the --set full table, and <tag>_ncu_traffic.json (DRAM bytes
per launch set of the dominant kernels, which bench.py reports as roofline.traffic).   python benchmarks/ncu_summarize.py r02"""
import json
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def short(name):
    return (name.split("(")[0].replace("void ", "").replace("ofs::<unnamed>::", "").replace("<unnamed>::", "")
            .replace("unnamed>::", ""))


def launch_shares():
    rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg, tot = collections.OrderedDict(), 0.0
    for d in data:
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"]) / 1e3
        tot += float(d["Metric Value"]) / 1e3
    out = ["# ncu launch list, `OFS_GRAPH=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline` "
           "(gpu__time_duration.sum, --clock-control none)", "",
           "Per-launch times are cold-cache and serialised; compare SHARES with bench.py's live numbers "
           "(roofline.share_of_step).  Raw list: the .csv beside this file.", "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")
    out += ["", "The list above covers everything the command launches, including bench.py's own measurement replays (the",
            "dense-layer set replayed 2 x 20 times by ofs_net_time_kernels).  ONE step of the hot path (from the input pack",
            "to the fused warp, first complete step in the list):", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    names = [short(d["Kernel Name"]) for d in data]
    start = next(i for i, k in enumerate(names) if k.startswith(("pack_act", "pack27")))
    end = next(i for i in range(start, len(names)) if names[i].startswith("warp5"))
    step, stot = collections.OrderedDict(), 0.0
    for d in data[start:end + 1]:
        a = step.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(d["Metric Value"]) / 1e3
        stot += float(d["Metric Value"]) / 1e3
    for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / stot:.1f}% |")
    dense = sum(t for k, (n, t) in step.items() if ("conv_gemm" in k and "<32," not in k) or "deconv_stack" in k or "splitk" in k)
    warp = sum(t for k, (n, t) in step.items() if k.startswith("warp5"))
    out += ["", f"One step under ncu: {stot:.0f} us in {end + 1 - start} launches; dense conv / deconv GEMMs incl. split-K reductions "
            f"{100 * dense / stot:.1f} % (bench.py live: roofline.share_of_step), fused warp {100 * warp / stot:.1f} % "
            "(roofline_warp.share_of_step)."]
    open(os.path.join(P, f"{tag}_ncu_launch_shares.md"), "w").write("\n".join(out) + "\n")
    os.replace(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_ncu_launches_bench_steps2.csv")) if False else None
    import shutil
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_ncu_launches_bench_steps2.csv"))


def raw(rep):
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def full_summary():
    labels = ["conv1 (slab, pairs)", "conv2 (slab, pairs)", "conv3 (pairs)", "conv3_1 (pairs)", "conv4 (192)", "conv4_1 (192)",
              "conv5 (split-K 6)", "conv5 reduce", "conv5_1 (split-K 6)", "conv5_1 reduce", "conv6 (128 cols, split-K 4)", "conv6 reduce",
              "conv6_1 (128 cols, split-K 4)", "conv6_1 reduce", "deconv5 + predict6 (1 CTA)", "deconv4 + predict5 (1 CTA)",
              "deconv3 + predict4 (pairs)", "deconv2 + predict3 (per phase, 1 CTA)", "predict2 1x1 product"]
    traffic = {"how": "ncu --set full --clock-control none, one step of `OFS_GRAPH=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline "
                      "--sustained-seconds 0` in launch order; dram__bytes_read.sum + dram__bytes_write.sum per launch", "per_launch": []}
    out = [f"# ncu --set full summaries ({tag}, B200, `OFS_GRAPH=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline`)", "",
           "`--clock-control none --import-source on`; per-launch values are cold-cache and serialised (ncu flushes caches between",
           "replays): compare utilisations and shares, not absolute times.  Scratch reports: gpurun_out/prof_conv.ncu-rep, prof_misc.ncu-rep.", ""]
    for rep, title, labs in (("prof_conv.ncu-rep", "tcgen05 implicit-GEMM conv kernels, one step in launch order (batch 8)", labels),
                             ("prof_misc.ncu-rep", "other kernels", None)):
        hdr, units, rows = raw(rep)
        idx = {h: i for i, h in enumerate(hdr)}

        def g(r, k):
            return r[idx[k]] if k in idx else "n/a"

        def f(r, k):
            try:
                return "%.1f" % float(g(r, k))
            except ValueError:
                return g(r, k)

        def mb(r, k):
            v, u = float(g(r, k)), units[idx[k]]
            return v / 1e6 if u == "byte" else v * 1e-3 if u == "Kbyte" else v if u == "Mbyte" else v * 1e3

        out += [f"## {title}", "",
                "| layer / kernel | kernel | grid | time us | tensor pipe active % of elapsed | of active (avg / busiest SM) | DRAM read MB | "
                "DRAM write MB | L2 hit % | issue active % | regs |", "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
        for i, r in enumerate(rows):
            if labs and i >= len(labs):
                break
            t = float(g(r, "gpu__time_duration.sum"))
            t_us = t / 1e3 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else t
            name = short(g(r, "Kernel Name"))[:40]
            lab = labs[i] if labs and i < len(labs) else name
            out.append("| %s | `%s` | %s | %.1f | %s | %s / %s | %.1f | %.1f | %s | %s | %s |" % (
                lab, name, g(r, "Grid Size").replace(", 1, 1", ""), t_us,
                f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                f(r, "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active"),
                mb(r, "dram__bytes_read.sum"), mb(r, "dram__bytes_write.sum"), f(r, "lts__t_sector_hit_rate.pct"),
                f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread")))
            try:
                traffic["per_launch"].append({"label": lab, "kernel": name, "us": t_us,
                                              "dram_bytes": (mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum")) * 1e6})
            except ValueError:
                pass
        out.append("")
    dense = [e for e in traffic["per_launch"] if ("conv_gemm" in e["kernel"] or "deconv_stack" in e["kernel"] or "splitk" in e["kernel"])
             and "predict2" not in e["label"]]
    warp = [e for e in traffic["per_launch"] if e["kernel"].startswith("warp5")]
    traffic["dense_set_bytes"] = sum(e["dram_bytes"] for e in dense)
    traffic["dense_set_launches"] = len(dense)
    traffic["warp_bytes"] = warp[0]["dram_bytes"] if warp else None
    json.dump(traffic, open(os.path.join(P, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    extra = os.path.join(P, f"{tag}_ncu_reading.md")
    if os.path.exists(extra):
        out += open(extra).read().splitlines()
    open(os.path.join(P, f"{tag}_ncu_full_summary.md"), "w").write("\n".join(out) + "\n")


launch_shares()
full_summary()
print("wrote", [f for f in sorted(os.listdir(P)) if f.startswith(tag + "_ncu")])
