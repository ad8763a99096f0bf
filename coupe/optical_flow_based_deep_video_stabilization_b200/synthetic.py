"""Random-init checkpoints and synthetic inputs of the FlowNetS-pyramid shapes, for benchmarks and demos when no trained
checkpoint is at hand (the reference's Drive checkpoint is unreachable offline).

The variable names and layouts are those of the TensorLayer npz the reference loads
(main_flownetS_pyramid_noprevloss_dataloader.py:520): <layer>/W_conv2d [k,k,cin,cout], <layer>/b_conv2d,
<layer>/W_deconv2d [4,4,cout,cin], <layer>/b_deconv2d, <bn>/beta, <bn>/moving_mean, <bn>/moving_variance.
Initialisers follow the reference layers (model.py:807-887): Conv2d W ~ variance_scaling_initializer() (truncated
normal, sigma = sqrt(1.3 * 2 / fan_in)), DeConv2dLayer W ~ truncated normal sigma 0.02 (TensorLayer default).

The CPU oracle carries its own copy of these generators (it must not depend on the product); tests/test_host_logic.py
checks that the two produce identical arrays.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

# (name, k, stride, cin, cout)  -- model.py:807-844
ENCODER = [("1", 7, 2, 27, 64), ("2", 5, 2, 64, 128), ("3", 5, 2, 128, 256), ("3_1", 3, 1, 256, 256),
           ("4", 3, 2, 256, 512), ("4_1", 3, 1, 512, 512), ("5", 3, 2, 512, 512), ("5_1", 3, 1, 512, 512),
           ("6", 3, 2, 512, 1024), ("6_1", 3, 1, 1024, 1024)]
# (name, bn name, cin, cout) -- model.py:850,859,868,877
DECONVS = [("deconv5", "deconv5_bn", 1024, 512), ("deconv4", "deconv4_bn", 1026, 256),
           ("deconv3", "deconv3_bn", 770, 128), ("deconv2", "deconv2_bn", 386, 64)]
# (name, cin) -- model.py:848,856,865,874,885
HEADS = [("predict6", 1024), ("predict5", 1026), ("predict4", 770), ("predict3", 386), ("predict2", 194)]
FLOW_UPS = ["upsample6_5", "upsample5_4", "upsample4_3", "upsample3_2"]  # model.py:852,861,870,879
NET_H, NET_W, NET_C = 384, 512, 27  # main_dl.py:491 placeholder [B,384,512,27]


def _trunc_normal(gen, shape, std):
    """tf.truncated_normal: resample outside 2 sigma."""
    x = torch.empty(shape, dtype=torch.float64)
    x.normal_(0.0, 1.0, generator=gen)
    bad = x.abs() > 2.0
    while bad.any():
        x[bad] = torch.empty(int(bad.sum()), dtype=torch.float64).normal_(0.0, 1.0, generator=gen)
        bad = x.abs() > 2.0
    return (x * std).to(torch.float32)


def make_weights(seed=0, kind="he", head_scale=None):
    """Deterministic synthetic checkpoint: OrderedDict name -> float32 numpy array (TF layouts).

    kind="he"          the reference initialisers, biases 0, BatchNorm mu = 0, var = 1, beta = 0.
    kind="calibrated"  same W, plus non-trivial BN statistics (mu ~ N(0,.1), var ~ U(.5,1.5), beta ~ N(0,.1)), biases
                       ~ N(0,.05), and the five flow heads scaled by `head_scale` (default 0.08) so that |flow2| is in
                       the few-pixel regime of a trained stabiliser instead of ~30 px.
    """
    gen = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    cal = kind == "calibrated"
    if head_scale is None:
        head_scale = 0.08 if cal else 1.0

    def bn(name, c):
        if cal:
            w[f"{name}/beta"] = (torch.randn(c, generator=gen) * 0.1).numpy()
            w[f"{name}/moving_mean"] = (torch.randn(c, generator=gen) * 0.1).numpy()
            w[f"{name}/moving_variance"] = (torch.rand(c, generator=gen) + 0.5).numpy()
        else:
            w[f"{name}/beta"] = np.zeros(c, np.float32)
            w[f"{name}/moving_mean"] = np.zeros(c, np.float32)
            w[f"{name}/moving_variance"] = np.ones(c, np.float32)

    def bias(c):
        return (torch.randn(c, generator=gen) * 0.05).numpy() if cal else np.zeros(c, np.float32)

    for name, k, s, cin, cout in ENCODER:
        std = math.sqrt(1.3 * 2.0 / (k * k * cin))
        w[f"{name}/W_conv2d"] = _trunc_normal(gen, (k, k, cin, cout), std).numpy()
        w[f"{name}/b_conv2d"] = bias(cout)
        bn(name, cout)
    for name, bnname, cin, cout in DECONVS:
        w[f"{name}/W_deconv2d"] = _trunc_normal(gen, (4, 4, cout, cin), 0.02).numpy()
        w[f"{name}/b_deconv2d"] = bias(cout)
        bn(bnname, cout)
    for name, cin in HEADS:
        std = math.sqrt(1.3 * 2.0 / (9 * cin))
        w[f"{name}/W_conv2d"] = (_trunc_normal(gen, (3, 3, cin, 2), std) * head_scale).numpy()
        w[f"{name}/b_conv2d"] = bias(2) * (head_scale if cal else 1.0)
    for name in FLOW_UPS:
        w[f"{name}/W_deconv2d"] = _trunc_normal(gen, (4, 4, 2, 2), 0.02).numpy()
        w[f"{name}/b_deconv2d"] = bias(2)
    return w


def make_feats(seed, batch, kind="uniform"):
    """Synthetic [B,384,512,27] network input in [0,1]."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, NET_H, NET_W, NET_C), generator=gen, dtype=torch.float32)
    if kind == "smooth":
        k = torch.ones(NET_C, 1, 9, 9) / 81.0
        x = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), k, padding=4, groups=NET_C).permute(0, 2, 3, 1)
        x = ((x - x.amin()) / (x.amax() - x.amin())).contiguous()
    return x
