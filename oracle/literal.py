"""Second, independent restatement of the same semantics as ``tf1_ops`` / ``samplers``,
written as scalar loops in numpy float32 -- deliberately slow and literal, for SMALL inputs
only.  Test infrastructure only (see ``oracle/__init__.py``).  The CPU tests require the
vectorised oracle and these loops to agree, so that a slip in either restatement shows up.

Each function cites the reference line (or the TF-1.10 kernel behaviour) it follows.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def resize_bilinear(x, oh, ow):
    """TF-1.10 ResizeBilinear CPU kernel, align_corners=False: in = out_idx * (in_size/out_size)."""
    B, H, W, C = x.shape
    out = np.zeros((B, oh, ow, C), f32)
    hs = f32(H) / f32(oh)
    ws = f32(W) / f32(ow)
    for oy in range(oh):
        iy = f32(oy) * hs
        y0 = int(math.floor(iy)); y1 = min(y0 + 1, H - 1); yl = f32(iy - f32(y0))
        for ox in range(ow):
            ix = f32(ox) * ws
            x0 = int(math.floor(ix)); x1 = min(x0 + 1, W - 1); xl = f32(ix - f32(x0))
            for b in range(B):
                tl, tr = x[b, y0, x0], x[b, y0, x1]
                bl, br = x[b, y1, x0], x[b, y1, x1]
                top = tl + (tr - tl) * xl
                bot = bl + (br - bl) * xl
                out[b, oy, ox] = top + (bot - top) * yl
    return out


def resize_nearest_align(x, oh, ow):
    """TF-1.10 ResizeNearestNeighbor, align_corners=True: min(roundf(i*(in-1)/(out-1)), in-1)."""
    B, H, W, C = x.shape
    out = np.zeros((B, oh, ow, C), x.dtype)
    hs = f32(H - 1) / f32(oh - 1)
    ws = f32(W - 1) / f32(ow - 1)
    for oy in range(oh):
        v = float(f32(oy) * hs)
        iy = min(int(math.floor(v + 0.5)), H - 1)
        for ox in range(ow):
            v = float(f32(ox) * ws)
            ix = min(int(math.floor(v + 0.5)), W - 1)
            out[:, oy, ox] = x[:, iy, ix]
    return out


def conv2d_pad_valid(x, w, b, k, s):
    """PadLayer(p=k//2, zeros) + tf.nn.conv2d VALID stride s + bias (model.py:807-808)."""
    p = k // 2
    B, H, W, C = x.shape
    xp = np.zeros((B, H + 2 * p, W + 2 * p, C), np.float64)
    xp[:, p:p + H, p:p + W] = x
    oh = (H + 2 * p - k) // s + 1
    ow = (W + 2 * p - k) // s + 1
    out = np.zeros((B, oh, ow, w.shape[3]), np.float64)
    for oy in range(oh):
        for ox in range(ow):
            patch = xp[:, oy * s:oy * s + k, ox * s:ox * s + k, :]          # [B,k,k,C]
            out[:, oy, ox] = np.tensordot(patch, w.astype(np.float64), axes=([1, 2, 3], [0, 1, 2]))
    return (out + b).astype(f32)


def conv2d_transpose_k4s2(x, w, b):
    """tf.nn.conv2d_transpose k4 s2 SAME: y[2i+ky-1, 2j+kx-1, co] += x[i,j,ci] W[ky,kx,co,ci]."""
    B, H, W, Ci = x.shape
    Co = w.shape[2]
    out = np.zeros((B, 2 * H, 2 * W, Co), np.float64)
    for i in range(H):
        for j in range(W):
            for ky in range(4):
                oy = 2 * i + ky - 1
                if oy < 0 or oy >= 2 * H:
                    continue
                for kx in range(4):
                    ox = 2 * j + kx - 1
                    if ox < 0 or ox >= 2 * W:
                        continue
                    out[:, oy, ox] += x[:, i, j].astype(np.float64) @ w[ky, kx].astype(np.float64).T
    return (out + b).astype(f32)


def tf_warp(img, flow):
    """main_flownetS_pyramid_noprevloss_dataloader.py:70-130, one pixel at a time."""
    B, H, W, C = img.shape
    out = np.zeros_like(img, dtype=f32)
    for b in range(B):
        for py in range(H):
            for px in range(W):
                x = f32(f32(px) + flow[b, py, px, 0])
                y = f32(f32(py) + flow[b, py, px, 1])
                x0 = int(x); y0 = int(y)                       # C truncation == tf.cast(int32)
                x1 = x0 + 1; y1 = y0 + 1
                x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1)
                y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1)
                wa = f32(f32(x1) - x) * f32(f32(y1) - y)
                wb = f32(f32(x1) - x) * f32(y - f32(y0))
                wc = f32(x - f32(x0)) * f32(f32(y1) - y)
                wd = f32(x - f32(x0)) * f32(y - f32(y0))
                out[b, py, px] = wa * img[b, y0, x0] + wb * img[b, y1, x0] + wc * img[b, y0, x1] + wd * img[b, y1, x1]
    return out


def _linspace_f32(n):
    if n == 1:
        return np.array([-1.0], f32)
    step = f32(2.0) / f32(n - 1)
    return np.array([f32(-1.0) + step * f32(i) for i in range(n)], f32)


def grid_sample(im, theta, out_size, projective):
    """AffineTransformer / ProjectiveTransformer .transform (spatial_transformer.py:400-452,539-608)
    with bilinear_interp (:902-964), one output pixel at a time, on an explicitly padded image."""
    B, H, W, C = im.shape
    oh, ow = out_size
    pad = np.zeros((B, H + 2, W + 2, C), f32)
    pad[:, 1:H + 1, 1:W + 1] = im
    xs = _linspace_f32(ow)
    ys = _linspace_f32(oh)
    out = np.zeros((B, oh, ow, C), f32)
    for b in range(B):
        t = theta[b].astype(f32)
        for oy in range(oh):
            for ox in range(ow):
                xt, yt = xs[ox], ys[oy]
                if projective:
                    x = f32(f32(t[0] * xt + t[1] * yt) + t[2])
                    y = f32(f32(t[3] * xt + t[4] * yt) + t[5])
                    z = f32(f32(t[6] * xt + t[7] * yt) + f32(1.0))
                    if z == 0:
                        z = f32(z + f32(1e-8))
                    x = f32(x / z); y = f32(y / z)
                else:
                    x = f32(f32(t[0] * xt + t[1] * yt) + t[2])
                    y = f32(f32(t[3] * xt + t[4] * yt) + t[5])
                x = f32(f32(f32(x + f32(1.0)) / f32(2.0)) * f32(W - 1))
                y = f32(f32(f32(y + f32(1.0)) / f32(2.0)) * f32(H - 1))
                x = min(max(x, f32(-1)), f32(W)); y = min(max(y, f32(-1)), f32(H))
                x = f32(x + f32(1)); y = f32(y + f32(1))
                x0f = f32(math.floor(x)); y0f = f32(math.floor(y))
                x1f = f32(x0f + 1); y1f = f32(y0f + 1)
                x0 = int(x0f); y0 = int(y0f)
                x1 = int(min(x1f, f32(W + 1))); y1 = int(min(y1f, f32(H + 1)))
                out[b, oy, ox] = (f32(x1f - x) * f32(y1f - y)) * pad[b, y0, x0] \
                    + (f32(x - x0f) * f32(y1f - y)) * pad[b, y0, x1] \
                    + (f32(x1f - x) * f32(y - y0f)) * pad[b, y1, x0] \
                    + (f32(x - x0f) * f32(y - y0f)) * pad[b, y1, x1]
    return out


def transform_image(image, pMtrx, refMtrx, out_h, out_w):
    """warp.py:46-86 / :89-129, one output pixel at a time (source size = image.shape[1:3])."""
    B, sh, sw, _ = image.shape
    out = np.zeros((B, out_h, out_w, 3), f32)
    Xs = np.linspace(-1, 1, out_w).astype(f32)
    Ys = np.linspace(-1, 1, out_h).astype(f32)
    for b in range(B):
        M = (refMtrx.astype(f32) @ pMtrx[b].astype(f32)).astype(f32)
        for oy in range(out_h):
            for ox in range(out_w):
                v = np.array([Xs[ox], Ys[oy], 1.0], f32)
                h = np.array([f32(f32(M[r, 0] * v[0] + M[r, 1] * v[1]) + M[r, 2] * v[2]) for r in range(3)], f32)
                X = f32(h[0] / f32(h[2] + f32(1e-8)))
                Y = f32(h[1] / f32(h[2] + f32(1e-8)))
                xf, xc = math.floor(X), math.ceil(X)
                yf, yc = math.floor(Y), math.ceil(Y)
                xr = f32(X - f32(xf)); yr = f32(Y - f32(yf))

                def px(xi, yi):
                    if 0 <= xi < sw and 0 <= yi < sh:
                        return image[b, yi, xi]
                    return np.zeros(3, f32)
                out[b, oy, ox] = px(xf, yf) * (1 - xr) * (1 - yr) + px(xc, yf) * xr * (1 - yr) \
                    + px(xf, yc) * (1 - xr) * yr + px(xc, yc) * xr * yr
    return out
