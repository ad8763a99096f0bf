"""Generates tests/golden/kat_v1.npz: known-answer vectors for the hot path.

The reference ships NO tests or golden vectors and cannot be executed offline (TensorFlow 1.10 /
TensorLayer are not installable here), so these vectors are derived ANALYTICALLY from the cited
reference lines -- closed forms written out below, with no call into oracle/ or the CUDA code.
Both the oracle (CPU tests) and the CUDA path (GPU tests) are checked against them.

    python tests/golden/make_golden.py      # rewrites kat_v1.npz deterministically
"""
import os

import numpy as np

f32 = np.float32
out = {}

# ---- K1 tf_warp, zero flow (main_dl.py:83-129): x0=x, x1=min(x+1,W-1); for x<W-1 and y<H-1 the weights are
# (1,0,0,0) -> identity; on the last column x1==x0 so (x1-x)=0 -> wa=wb=0 and (x-x0)=0 -> wc=wd=0 -> 0.
rng = np.random.RandomState(11)
img = rng.rand(2, 5, 7, 3).astype(f32)
exp = img.copy()
exp[:, -1, :, :] = 0
exp[:, :, -1, :] = 0
out["k1_img"], out["k1_flow"], out["k1_out"] = img, np.zeros((2, 5, 7, 2), f32), exp

# ---- K2 flow = (+0.5, 0): x = c+0.5, x0=c, x1=c+1 (c<=W-2) -> 0.5*img[c] + 0.5*img[c+1] in rows < H-1;
# column W-1: x=W-0.5 -> x0 = W-1, x1 clipped to W-1 -> wa=(W-1-x)=-0.5 * (y1-y), wc = (x-x0)=0.5*(y1-y) -> both taps
# read img[W-1] -> (-0.5+0.5)*img = 0.  Last row: y1==y0 -> (y1-y)=0 and (y-y0)=0 -> 0.
flow = np.zeros((2, 5, 7, 2), f32)
flow[..., 0] = 0.5
exp = np.zeros_like(img)
exp[:, :-1, :-1] = 0.5 * img[:, :-1, :-1] + 0.5 * img[:, :-1, 1:]
out["k2_flow"], out["k2_out"] = flow, exp.astype(f32)

# ---- K3 x in (-1,0): flow_x = -0.25 at column 0 -> x=-0.25, trunc -> 0, x0=0, x1=1: weights (1-x... ) = (x1-x, x-x0) =
# (1.25, -0.25): linear EXTRAPOLATION 1.25*img[0] - 0.25*img[1]   (rows < H-1, zero y flow)
flow = np.zeros((2, 5, 7, 2), f32)
flow[:, :, 0, 0] = -0.25
exp = img.copy()
exp[:, :, 0] = f32(1.25) * img[:, :, 0] + f32(-0.25) * img[:, :, 1]
exp[:, -1] = 0
exp[:, :, -1] = 0
out["k3_flow"], out["k3_out"] = flow, exp.astype(f32)

# ---- K4 far outside: x >= W-1 or x <= -1 -> both corners clip to the same column -> 0 everywhere
flow = np.zeros((2, 5, 7, 2), f32)
flow[0, ..., 0] = 100.0
flow[1, ..., 0] = -100.0
out["k4_flow"], out["k4_out"] = flow, np.zeros_like(img)

# ---- K5 TF1 legacy bilinear 2x up-sampling of a ramp (model.py:857): src = dst*0.5 ->
# out[2k]=in[k], out[2k+1]=(in[k]+in[k+1])/2, last = in[last] (hi index clamps)
src = (np.arange(6 * 8, dtype=f32).reshape(1, 6, 8, 1) * f32(0.5)) ** 1
src = np.concatenate([src, -src], 3)
def up2_1d(a, axis):
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    o = np.zeros((2 * n,) + a.shape[1:], f32)
    o[0::2] = a
    o[1:-1:2] = (a[:-1] + (a[1:] - a[:-1]) * f32(0.5))
    o[-1] = a[-1]
    return np.moveaxis(o, 0, axis)
out["k5_in"], out["k5_out"] = src, up2_1d(up2_1d(src, 2), 1)

# ---- K6 nearest-neighbour align_corners index tables (model.py:883): idx = round(i*(in-1)/(out-1)); exact rational
# arithmetic (no tie can occur: 194*i = 383*(2n+1) has no solution by parity; same for 258*i = 511*(2n+1))
out["k6_rows"] = np.array([(2 * i * 97 + 383) // (2 * 383) for i in range(384)], np.int64)
out["k6_cols"] = np.array([(2 * i * 129 + 511) // (2 * 511) for i in range(512)], np.int64)

# ---- K7 transposed conv k4 s2 SAME impulse (model.py:850): x = delta at (i,j), ci=0 ->
# y[2i+ky-1, 2j+kx-1, co] = W[ky,kx,co,0]
W = rng.randn(4, 4, 3, 2).astype(f32)
x = np.zeros((1, 4, 5, 2), f32)
x[0, 1, 2, 0] = 1.0
x[0, 0, 0, 1] = 2.0            # corner impulse on ci=1: rows/cols -1 fall off the output
y = np.zeros((1, 8, 10, 3), f32)
for ky in range(4):
    for kx in range(4):
        oy, ox = 2 * 1 + ky - 1, 2 * 2 + kx - 1
        y[0, oy, ox] += W[ky, kx, :, 0]
        oy, ox = ky - 1, kx - 1
        if oy >= 0 and ox >= 0:
            y[0, oy, ox] += f32(2.0) * W[ky, kx, :, 1]
out["k7_w"], out["k7_x"], out["k7_out"] = W, x, y

# ---- K8 bilinear_interp (spatial_transformer.py:902-964): identity theta reproduces the image when out_size==in size;
# a shift of exactly one pixel is theta[2] = 2/(W-1): out[:, :, c] = img[:, :, c+1], last column reads the zero border
img8 = rng.rand(2, 6, 9, 3).astype(f32)
out["k8_img"] = img8
out["k8_theta_id"] = np.tile(np.array([1, 0, 0, 0, 1, 0], f32), (2, 1))
out["k8_out_id"] = img8
th = np.tile(np.array([1, 0, 2.0 / 8.0, 0, 1, 0], f32), (2, 1))
exp = np.zeros_like(img8)
exp[:, :, :-1] = img8[:, :, 1:]
out["k8_theta_shift"], out["k8_out_shift"] = th, exp
# far outside: theta translation 10 -> x clipped to W (the zero border column) -> 0
out["k8_theta_far"] = np.tile(np.array([1, 0, 10.0, 0, 1, 0], f32), (2, 1))
out["k8_out_far"] = np.zeros_like(img8)

# ---- K9 transformImage (warp.py:46-86): pMtrx = I, refMtrx maps [-1,1] -> pixel coords -> identity
H9, W9 = 5, 9
ref = np.array([[(W9 - 1) / 2.0, 0, (W9 - 1) / 2.0], [0, (H9 - 1) / 2.0, (H9 - 1) / 2.0], [0, 0, 1]], f32)
img9 = rng.rand(2, H9, W9, 3).astype(f32)
out["k9_img"], out["k9_ref"], out["k9_p"] = img9, ref, np.tile(np.eye(3, dtype=f32), (2, 1, 1))
out["k9_out"] = img9            # W-1 = 8 and H-1 = 4 are powers of two: linspace*4+4 is exact, floor==ceil everywhere

# ---- K10 vec2mtrx (warp.py:25-43): p = 0 -> I;  homography p1 only (x translation generator, nilpotent): I + p1*E02
out["k10_p_zero"] = np.zeros((2, 8), f32)
out["k10_m_zero"] = np.tile(np.eye(3, dtype=f32), (2, 1, 1))
p = np.zeros((2, 8), f32)
p[:, 0] = [0.25, -1.5]
m = np.tile(np.eye(3, dtype=f32), (2, 1, 1))
m[:, 0, 2] = p[:, 0]
out["k10_p_tx"], out["k10_m_tx"] = p, m
# affine: A = [[a,0,0],[0,0,0],[0,0,0]] -> exp series truncated at warpApprox=4: 1 + a + a^2/2 + a^3/6 at [0,0]
pa = np.zeros((1, 6), f32)
pa[0, 0] = 0.5
ma = np.eye(3, dtype=f32)[None].copy()
ma[0, 0, 0] = f32(f32(f32(1.0 + 0.5) + f32(0.25 / 2.0)) + f32(0.125 / 6.0))
out["k10_p_aff"], out["k10_m_aff"] = pa, ma

# ---- K11 flow glue (main_dl.py:497-498): constant flow2 = (a, b) on 382x510 -> resize of a constant is the
# constant; out = (a*384/382 * W/512, b*384/382 * H/384)
a, b = f32(1.5), f32(-2.25)
H11, W11 = 16, 24
fx = f32(f32(f32(a * f32(384.0)) / f32(382.0)) * f32(W11)) / f32(512.0)
fy = f32(f32(f32(b * f32(384.0)) / f32(382.0)) * f32(H11)) / f32(384.0)
out["k11_ab"] = np.array([a, b], f32)
out["k11_hw"] = np.array([H11, W11], np.int64)
out["k11_out"] = np.array([fx, fy], f32)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, sorted(out))
