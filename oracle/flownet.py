"""CPU restatement of ``flownetS_pyramid`` (reference ``model.py:786-893``).

Test infrastructure only -- see ``oracle/__init__.py`` (PARITY UNPINNED: the
reference cannot run here and ships no golden vectors).

Two forwards are provided:

``forward_literal``   follows model.py line by line: PadLayer -> Conv2d(VALID)+bias ->
                      BatchNormLayer(moving stats, no gamma) -> lrelu, DeConv2dLayer,
                      ConcatLayer, ElementwiseLayer left-fold adds, NN-upsample +
                      predict2.  This is THE oracle.
``forward_folded``    the algebraically equal form the CUDA path computes (BN folded
                      into W'/b', concat-by-slice, predict2 as a 1x1 GEMM on the 96x128
                      grid + 9-tap gather).  With ``emulate_bf16=True`` it rounds
                      weights and stored activations to bf16 at exactly the points the
                      CUDA path does, which separates "indexing bug" (O(1) error) from
                      "bf16 operand rounding" (O(1e-2) relative) in the GPU tests.

Variable names follow the TensorLayer npz checkpoint the reference loads
(``main_flownetS_pyramid_noprevloss_dataloader.py:520``): ``<layer>/W_conv2d``,
``<layer>/b_conv2d``, ``<layer>/W_deconv2d``, ``<layer>/b_deconv2d``, ``<bn>/beta``,
``<bn>/moving_mean``, ``<bn>/moving_variance`` under ``main_net/flownetS/``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from . import tf1_ops as T

# (name, k, stride, cin, cout)  -- model.py:807-844
ENCODER = [
    ("1", 7, 2, 27, 64),
    ("2", 5, 2, 64, 128),
    ("3", 5, 2, 128, 256),
    ("3_1", 3, 1, 256, 256),
    ("4", 3, 2, 256, 512),
    ("4_1", 3, 1, 512, 512),
    ("5", 3, 2, 512, 512),
    ("5_1", 3, 1, 512, 512),
    ("6", 3, 2, 512, 1024),
    ("6_1", 3, 1, 1024, 1024),
]
# (name, bn name, cin, cout) -- model.py:850,859,868,877
DECONVS = [
    ("deconv5", "deconv5_bn", 1024, 512),
    ("deconv4", "deconv4_bn", 1026, 256),
    ("deconv3", "deconv3_bn", 770, 128),
    ("deconv2", "deconv2_bn", 386, 64),
]
# (name, cin) -- model.py:848,856,865,874,885
HEADS = [("predict6", 1024), ("predict5", 1026), ("predict4", 770), ("predict3", 386), ("predict2", 194)]
FLOW_UPS = ["upsample6_5", "upsample5_4", "upsample4_3", "upsample3_2"]  # model.py:852,861,870,879

NET_H, NET_W, NET_C = 384, 512, 27  # main:491 placeholder [B,384,512,27]


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
    """Deterministic synthetic checkpoint (the Drive checkpoint is unreachable offline).

    kind="he"          reference initialisers: Conv2d W ~ variance_scaling_initializer()
                       (truncated normal, sigma = sqrt(1.3*2/fan_in)), DeConv2dLayer W ~
                       truncated normal sigma 0.02 (TL default), biases 0, BN mu=0 var=1 beta=0.
    kind="calibrated"  same W, plus non-trivial BN stats (mu~N(0,.1), var~U(.5,1.5),
                       beta~N(0,.1)), biases ~N(0,.05), and the five flow heads scaled by
                       ``head_scale`` (default 0.08) so that |flow2| is in the few-pixel
                       regime of a trained stabiliser instead of ~30 px.
    Returns an OrderedDict name -> float32 numpy array (TF layouts).
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
    """Synthetic [B,384,512,27] input in [0,1]."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, NET_H, NET_W, NET_C), generator=gen, dtype=torch.float32)
    if kind == "smooth":
        k = torch.ones(NET_C, 1, 9, 9) / 81.0
        x = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), k, padding=4, groups=NET_C).permute(0, 2, 3, 1)
        x = ((x - x.amin()) / (x.amax() - x.amin())).contiguous()
    return x


def _t(w, name, dtype):
    return torch.from_numpy(np.ascontiguousarray(w[name])).to(dtype)


def forward_literal(feats, w, dtype=torch.float32, keep=False):
    """model.py:805-893, line by line.  feats: torch [B,384,512,27].  Returns the 6-key dict
    (plus every intermediate under its reference variable name when ``keep``)."""
    x = feats.to(dtype)
    acts = OrderedDict()

    def conv_bn(x, name, k, s):
        n = T.pad_constant(x, k // 2)                                                # PadLayer
        n = T.conv2d_valid(n, _t(w, f"{name}/W_conv2d", dtype), _t(w, f"{name}/b_conv2d", dtype), s)
        n = T.batchnorm_infer(n, _t(w, f"{name}/moving_mean", dtype), _t(w, f"{name}/moving_variance", dtype),
                              _t(w, f"{name}/beta", dtype))
        return T.lrelu(n, 0.1)

    enc = {}
    n = x
    for name, k, s, cin, cout in ENCODER:                                            # :807-844
        n = conv_bn(n, name, k, s)
        enc[name] = n
        acts["conv" + name] = n

    def head(x, name):
        return T.conv2d_valid(T.pad_constant(x, 1), _t(w, f"{name}/W_conv2d", dtype), _t(w, f"{name}/b_conv2d", dtype), 1)

    def deconv_bn(x, name, bnname):
        n = T.conv2d_transpose_k4s2_same(x, _t(w, f"{name}/W_deconv2d", dtype), _t(w, f"{name}/b_deconv2d", dtype))
        n = T.batchnorm_infer(n, _t(w, f"{bnname}/moving_mean", dtype), _t(w, f"{bnname}/moving_variance", dtype),
                              _t(w, f"{bnname}/beta", dtype))
        return T.lrelu(n, 0.1)

    def flow_up(f, name):
        return T.conv2d_transpose_k4s2_same(f, _t(w, f"{name}/W_deconv2d", dtype), _t(w, f"{name}/b_deconv2d", dtype))

    f6 = head(enc["6_1"], "predict6")                                                # :847-848
    deconv5 = deconv_bn(enc["6_1"], "deconv5", "deconv5_bn")                         # :850-851
    up65 = flow_up(f6, "upsample6_5")                                                # :852
    concat5 = torch.cat([enc["5_1"], deconv5, up65], dim=3)                          # :853
    u = T.resize_bilinear_tf1(f6, 12, 16)
    f5 = (head(concat5, "predict5") + u) + u                                         # :855-857
    deconv4 = deconv_bn(concat5, "deconv4", "deconv4_bn")                            # :859-860
    up54 = flow_up(f5, "upsample5_4")                                                # :861
    concat4 = torch.cat([enc["4_1"], deconv4, up54], dim=3)                          # :862
    u = T.resize_bilinear_tf1(f5, 24, 32)
    f4 = (head(concat4, "predict4") + u) + u                                         # :864-866
    deconv3 = deconv_bn(concat4, "deconv3", "deconv3_bn")                            # :868-869
    up43 = flow_up(f4, "upsample4_3")                                                # :870
    concat3 = torch.cat([enc["3_1"], deconv3, up43], dim=3)                          # :871
    u = T.resize_bilinear_tf1(f4, 48, 64)
    f3 = (head(concat3, "predict3") + u) + u                                         # :873-875
    deconv2 = deconv_bn(concat3, "deconv2", "deconv2_bn")                            # :877-878
    up32 = flow_up(f3, "upsample3_2")                                                # :879
    concat2 = torch.cat([enc["2"], deconv2, up32], dim=3)                            # :880
    n = T.pad_constant(concat2, 1)                                                   # :882  [B,98,130,194]
    n = T.resize_nearest_tf1_align(n, feats.shape[1], feats.shape[2])                # :883  [B,384,512,194]
    p2 = T.conv2d_valid(n, _t(w, "predict2/W_conv2d", dtype), _t(w, "predict2/b_conv2d", dtype), 1)  # :885
    u = T.resize_bilinear_tf1(f3, 382, 510)                                          # :886
    f2 = p2
    for _ in range(8):                                                               # :887 left fold of tf.add
        f2 = f2 + u
    out = {"predict_flow6": f6, "predict_flow5": f5, "predict_flow4": f4, "predict_flow3": f3,
           "predict_flow2": f2, "flow": f2}                                          # :893
    if keep:
        acts.update(deconv5=deconv5, concat5=concat5, deconv4=deconv4, concat4=concat4, deconv3=deconv3,
                    concat3=concat3, deconv2=deconv2, concat2=concat2, head2=p2)
        out["_acts"] = acts
    return out


# ---------------------------------------------------------------------------------------
# folded form (what the CUDA path computes) + bf16 emulation
# ---------------------------------------------------------------------------------------

def fold_bn(w):
    """W' = W * r, b' = (b - mu) * r + beta, r = 1/sqrt(var+eps) per output channel (fp64 math,
    fp32 result).  Conv W [kh,kw,cin,cout] scales the last axis; deconv W [4,4,cout,cin] axis 2."""
    f = OrderedDict()
    for name, k, s, cin, cout in ENCODER:
        r = 1.0 / np.sqrt(w[f"{name}/moving_variance"].astype(np.float64) + T.BN_EPS)
        f[f"{name}/W"] = (w[f"{name}/W_conv2d"].astype(np.float64) * r).astype(np.float32)
        f[f"{name}/b"] = ((w[f"{name}/b_conv2d"].astype(np.float64) - w[f"{name}/moving_mean"]) * r
                          + w[f"{name}/beta"]).astype(np.float32)
    for name, bnname, cin, cout in DECONVS:
        r = 1.0 / np.sqrt(w[f"{bnname}/moving_variance"].astype(np.float64) + T.BN_EPS)
        f[f"{name}/W"] = (w[f"{name}/W_deconv2d"].astype(np.float64) * r[None, None, :, None]).astype(np.float32)
        f[f"{name}/b"] = ((w[f"{name}/b_deconv2d"].astype(np.float64) - w[f"{bnname}/moving_mean"]) * r
                          + w[f"{bnname}/beta"]).astype(np.float32)
    for name, cin in HEADS:
        f[f"{name}/W"] = w[f"{name}/W_conv2d"].astype(np.float32)
        f[f"{name}/b"] = w[f"{name}/b_conv2d"].astype(np.float32)
    for name in FLOW_UPS:
        f[f"{name}/W"] = w[f"{name}/W_deconv2d"].astype(np.float32)
        f[f"{name}/b"] = w[f"{name}/b_deconv2d"].astype(np.float32)
    return f


def round_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def round_fp16(x):
    return x.to(torch.float16).to(x.dtype)


def forward_folded(feats, folded, emulate_bf16=False, acc_dtype=torch.float64, keep=False):
    """Folded network.  With emulate_bf16 (True = bf16, "fp16" = IEEE half) the GEMM operands (stored activations, folded conv/deconv/
    head weights) are rounded to bf16 exactly where the CUDA path stores bf16; accumulation is done
    in ``acc_dtype`` (fp64 = the ideal the fp32 tensor-core accumulator approximates).  Flow maps,
    biases, flow up-sampler weights and all pyramid arithmetic stay fp32/fp64 like the CUDA path."""
    if emulate_bf16 == "fp16":
        q = round_fp16
    else:
        q = round_bf16 if emulate_bf16 else (lambda t: t)
    dt = acc_dtype

    def W(name):
        return torch.from_numpy(folded[name]).to(dt)

    def Wq(name):
        return q(torch.from_numpy(folded[name]).to(torch.float32)).to(dt)

    acts = OrderedDict()
    n = q(feats.to(torch.float32)).to(dt)
    enc = {}
    for name, k, s, cin, cout in ENCODER:
        y = T.conv2d_valid(T.pad_constant(n, k // 2), Wq(f"{name}/W"), W(f"{name}/b"), s)
        n = q(T.lrelu(y, 0.1).to(torch.float32)).to(dt)
        enc[name] = n
        acts["conv" + name] = n

    def head(x, name):
        return T.conv2d_valid(T.pad_constant(x, 1), Wq(f"{name}/W"), W(f"{name}/b"), 1)

    def deconv(x, name):
        y = T.conv2d_transpose_k4s2_same(x, Wq(f"{name}/W"), W(f"{name}/b"))
        return q(T.lrelu(y, 0.1).to(torch.float32)).to(dt)

    def flow_up(f, name):  # fp32 math on the fp32 flow, stored into the bf16 concat slice
        y = T.conv2d_transpose_k4s2_same(f, W(f"{name}/W"), W(f"{name}/b"))
        return q(y.to(torch.float32)).to(dt)

    raw6 = head(enc["6_1"], "predict6")
    f6 = raw6
    d5 = deconv(enc["6_1"], "deconv5")
    concat5 = torch.cat([enc["5_1"], d5, flow_up(f6, "upsample6_5")], 3)
    f5 = head(concat5, "predict5") + 2.0 * T.resize_bilinear_tf1(f6, 12, 16)
    d4 = deconv(concat5, "deconv4")
    concat4 = torch.cat([enc["4_1"], d4, flow_up(f5, "upsample5_4")], 3)
    f4 = head(concat4, "predict4") + 2.0 * T.resize_bilinear_tf1(f5, 24, 32)
    d3 = deconv(concat4, "deconv3")
    concat3 = torch.cat([enc["3_1"], d3, flow_up(f4, "upsample4_3")], 3)
    f3 = head(concat3, "predict3") + 2.0 * T.resize_bilinear_tf1(f4, 48, 64)
    d2 = deconv(concat3, "deconv2")
    concat2 = torch.cat([enc["2"], d2, flow_up(f3, "upsample3_2")], 3)

    # predict2 restructured: P[t] = concat2 (96x128 grid) x W2[t] (1x1 GEMM, 18 columns), then the
    # 3x3 VALID conv over the NN-upsampled zero-padded map is a 9-tap gather-sum of P.
    B = feats.shape[0]
    w2 = Wq("predict2/W")                                   # [3,3,194,2]
    P = torch.einsum("bhwc,yxco->bhwyxo", concat2, w2)      # [B,96,128,3,3,2]
    Ppad = torch.nn.functional.pad(P, (0, 0, 0, 0, 0, 0, 1, 1, 1, 1))   # zero border = padded concat2
    yi = torch.from_numpy(T.nearest_align_table(98, NET_H))
    xi = torch.from_numpy(T.nearest_align_table(130, NET_W))
    head2 = torch.zeros((B, 382, 510, 2), dtype=dt) + W("predict2/b")
    for ky in range(3):
        for kx in range(3):
            head2 = head2 + Ppad[:, yi[ky:ky + 382]][:, :, xi[kx:kx + 510]][:, :, :, ky, kx, :]
    f2 = head2 + 8.0 * T.resize_bilinear_tf1(f3, 382, 510)
    out = {"predict_flow6": f6, "predict_flow5": f5, "predict_flow4": f4, "predict_flow3": f3,
           "predict_flow2": f2, "flow": f2}
    if keep:
        acts.update(deconv5=d5, concat5=concat5, deconv4=d4, concat4=concat4, deconv3=d3, concat3=concat3,
                    deconv2=d2, concat2=concat2, head2=head2)
        out["_acts"] = acts
    return out


def epe(a, b):
    """mean endpoint error between two [B,h,w,2] flow maps (pixels)."""
    d = (a.to(torch.float64) - b.to(torch.float64))
    return float(torch.sqrt((d * d).sum(-1)).mean())
