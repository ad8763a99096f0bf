import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs
from oracle import flownet as F
dev = torch.device("cuda", 0)
w = F.make_weights(0, "calibrated", head_scale=0.02)
x = F.make_feats(2, 2)
ref = F.forward_literal(x, w)
res = {}
for flag in ("0", "1"):
    os.environ["OFS_CONV1X2"] = flag
    net = ofs.FlowNetSPyramid(device=dev, max_batch=2, precision="bf16")
    net.assign_weights(w)
    out = net.forward(x.to(dev))
    c1 = net.activation("conv1", 2).cpu()
    c2 = net.activation("conv2", 2).cpu()
    f2 = out["predict_flow2"].cpu().clone()
    res[flag] = (c1, c2, f2)
    print("OFS_CONV1X2=" + flag, "EPE vs fp32 oracle", F.epe(f2, ref["predict_flow2"]))
    net.close()
a, b = res["0"][0], res["1"][0]
d = (a - b).abs()
print("conv1: differing elements", int((d > 0).sum()), "of", d.numel(), "max abs diff", float(d.max()), "max |a|", float(a.abs().max()))
idx = (d > 0).nonzero()
if len(idx):
    xs = idx[:, 2]
    print("x positions of diffs (hist of x % 128):", torch.bincount(xs % 128, minlength=128).tolist()[:16], "...")
    print("distinct x:", sorted(set(xs.tolist()))[:40])
    big = (d > 0.05 * a.abs().max()).nonzero()
    print("large diffs:", len(big), big[:10].tolist())
d2 = (res["0"][1] - res["1"][1]).abs()
print("conv2: differing", int((d2 > 0).sum()), "max", float(d2.max()))
print("flow2 EPE between the two forms", F.epe(res["0"][2], res["1"][2]))
