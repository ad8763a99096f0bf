"""CPU restatement of the reference's samplers.  Test infrastructure only -- see
``oracle/__init__.py`` (PARITY UNPINNED).

  tf_warp / get_pixel_value   main_flownetS_pyramid_noprevloss_dataloader.py:44-130
  flow_resize (test-mode glue) main_flownetS_pyramid_noprevloss_dataloader.py:497-498
  _meshgrid, bilinear_interp, AffineTransformer, ProjectiveTransformer
                               spatial_transformer.py:755-779, 902-964, 373-452, 519-608
  vec2mtrx, transformImage, transformCropImage   warp.py:25-129

All functions take/return NHWC torch CPU tensors; arithmetic is fp32 in the same operation
order as the reference graph (pass float64 tensors for the noise-floor cross-check).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tf1_ops as T


# ------------------------------------------------------------------------------ tf_warp
def tf_warp(img, flow, H, W):
    """main:70-130.  img [B,H,W,C], flow [B,H,W,2] (ch0 = x, ch1 = y, pixel units)."""
    dt = img.dtype
    gx = torch.arange(W, dtype=dt).view(1, 1, W)
    gy = torch.arange(H, dtype=dt).view(1, H, 1)
    x = gx + flow[..., 0]                                   # :83  grid + flow
    y = gy + flow[..., 1]
    x0 = x.to(torch.int32)                                  # :92  tf.cast -> truncation toward zero
    y0 = y.to(torch.int32)
    x1 = x0 + 1
    y1 = y0 + 1
    x0 = x0.clamp(0, W - 1); x1 = x1.clamp(0, W - 1)        # :98-101
    y0 = y0.clamp(0, H - 1); y1 = y1.clamp(0, H - 1)
    B = img.shape[0]
    b = torch.arange(B).view(B, 1, 1).expand(B, H, W)
    Ia = img[b, y0.long(), x0.long()]                       # :104-107 gather_nd
    Ib = img[b, y1.long(), x0.long()]
    Ic = img[b, y0.long(), x1.long()]
    Id = img[b, y1.long(), x1.long()]
    x0f, x1f, y0f, y1f = x0.to(dt), x1.to(dt), y0.to(dt), y1.to(dt)   # :110-113 recast CLIPPED corners
    wa = ((x1f - x) * (y1f - y)).unsqueeze(3)               # :117-120
    wb = ((x1f - x) * (y - y0f)).unsqueeze(3)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(3)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(3)
    return wa * Ia + wb * Ib + wc * Ic + wd * Id            # :129 add_n


def flow_resize(flow2, out_h, out_w):
    """main:497-498.  flow2 [B,382,510,2] -> [B,out_h,out_w,2] in video-pixel units."""
    f = flow2 * 384.0 / flow2.shape[1]                      # (flow*384.0)/382
    f = T.resize_bilinear_tf1(f, out_h, out_w)
    return torch.cat([f[..., 0:1] * out_w / 512, f[..., 1:2] * out_h / 384], 3)


def flow_resize_warp(img, flow2, out_h, out_w):
    """main:497-514: the whole post-network part of the test-mode graph."""
    return tf_warp(img, flow_resize(flow2, out_h, out_w), out_h, out_w)


# ------------------------------------------------------------------ spatial_transformer
def tf_linspace(start, stop, num):
    """tf.linspace (LinSpace op, T=float32): start + step*i with step=(stop-start)/(num-1), all fp32."""
    if num == 1:
        return np.array([start], np.float32)
    step = (np.float32(stop) - np.float32(start)) / np.float32(num - 1)
    return (np.float32(start) + step * np.arange(num, dtype=np.float32)).astype(np.float32)


def meshgrid(out_size, dtype=torch.float32):
    """spatial_transformer.py:755-779: rows [x; y; 1] of linspace(-1,1) coords, [3, H*W]."""
    oh, ow = out_size
    xs = torch.from_numpy(tf_linspace(-1.0, 1.0, ow)).to(dtype)
    ys = torch.from_numpy(tf_linspace(-1.0, 1.0, oh)).to(dtype)
    xt = xs.view(1, ow).expand(oh, ow).reshape(-1)
    yt = ys.view(oh, 1).expand(oh, ow).reshape(-1)
    return torch.stack([xt, yt, torch.ones_like(xt)], 0)


def bilinear_interp(im, x, y, out_size):
    """spatial_transformer.py:902-964.  im [B,H,W,C]; x,y flat [B*oh*ow] in [-1,1] coords."""
    dt = im.dtype
    B, H, W, C = im.shape
    e = 1
    imp = torch.nn.functional.pad(im, (0, 0, e, e, e, e))                     # :906
    x = (x + 1.0) / 2.0 * (W - 1.0)                                           # :916
    y = (y + 1.0) / 2.0 * (H - 1.0)
    x = x.clamp(-e, W - 1 + e)                                                # :918
    y = y.clamp(-e, H - 1 + e)
    x = x + e                                                                 # :921
    y = y + e
    x0f = torch.floor(x); y0f = torch.floor(y)                                # :925
    x1f = x0f + 1; y1f = y0f + 1
    x0 = x0f.long(); y0 = y0f.long()
    x1 = torch.minimum(x1f, torch.tensor(W - 1 + 2 * e, dtype=dt)).long()     # :932
    y1 = torch.minimum(y1f, torch.tensor(H - 1 + 2 * e, dtype=dt)).long()
    dim2 = W + 2 * e
    dim1 = dim2 * (H + 2 * e)
    n_out = out_size[0] * out_size[1]
    base = (torch.arange(B) * dim1).repeat_interleave(n_out)                  # :938
    flat = imp.reshape(-1, C)
    I00 = flat[base + y0 * dim2 + x0]
    I01 = flat[base + y0 * dim2 + x1]
    I10 = flat[base + y1 * dim2 + x0]
    I11 = flat[base + y1 * dim2 + x1]
    w00 = ((x1f - x) * (y1f - y)).unsqueeze(1)                                # :958 (UNCLIPPED x1f)
    w01 = ((x - x0f) * (y1f - y)).unsqueeze(1)
    w10 = ((x1f - x) * (y - y0f)).unsqueeze(1)
    w11 = ((x - x0f) * (y - y0f)).unsqueeze(1)
    return w00 * I00 + w01 * I01 + w10 * I10 + w11 * I11                      # :963


def affine_transform(inp, theta, out_size):
    """AffineTransformer(out_size).transform(inp, theta)  spatial_transformer.py:400-452."""
    B, H, W, C = inp.shape
    grid = meshgrid(out_size, inp.dtype)                                       # :397
    th = theta.reshape(-1, 2, 3).to(inp.dtype)                                 # :442
    tg = torch.matmul(th, grid.unsqueeze(0).expand(B, 3, -1))                  # :447
    xs = tg[:, 0, :].reshape(-1)
    ys = tg[:, 1, :].reshape(-1)
    out = bilinear_interp(inp, xs, ys, out_size)
    return out.reshape(B, out_size[0], out_size[1], C)                         # :432


def projective_transform(inp, theta, out_size):
    """ProjectiveTransformer(out_size).transform(inp, theta)  spatial_transformer.py:539-608."""
    B, H, W, C = inp.shape
    grid = meshgrid(out_size, inp.dtype)
    th = torch.cat([theta.reshape(B, 8).to(inp.dtype), torch.ones(B, 1, dtype=inp.dtype)], 1).reshape(B, 3, 3)
    tg = torch.matmul(th, grid.unsqueeze(0).expand(B, 3, -1))
    xs, ys, zs = tg[:, 0, :], tg[:, 1, :], tg[:, 2, :]
    safe = torch.where(zs == 0, zs + 1e-8, zs)                                 # :598
    xs = (xs / safe).reshape(-1)
    ys = (ys / safe).reshape(-1)
    out = bilinear_interp(inp, xs, ys, out_size)
    return out.reshape(-1, out_size[0], out_size[1], C)


# --------------------------------------------------------------------------------- warp.py
def vec2mtrx(p, warp_type, warp_approx):
    """warp.py:25-43: Lie-algebra parameters -> 3x3 via truncated matrix-exponential series."""
    dt = p.dtype
    B = p.shape[0]
    O = torch.zeros(B, dtype=dt)
    if warp_type == "homography":
        p1, p2, p3, p4, p5, p6, p7, p8 = [p[:, i] for i in range(8)]
        A = torch.stack([torch.stack([p3, p2, p1], 1), torch.stack([p6, -p3 - p7, p5], 1),
                         torch.stack([p4, p8, p7], 1)], 1)                      # :29
    elif warp_type == "affine":
        p1, p2, p3, p4, p5, p6 = [p[:, i] for i in range(6)]
        A = torch.stack([torch.stack([p1, p2, p3], 1), torch.stack([p4, p5, p6], 1),
                         torch.stack([O, O, O], 1)], 1)                         # :33
    else:
        raise AssertionError("unknown warpType")                               # :34
    eye = torch.eye(3, dtype=dt).unsqueeze(0).repeat(B, 1, 1)
    pm = eye.clone()
    numer = eye.clone()
    denom = 1.0
    for i in range(1, warp_approx):                                            # :39-42
        numer = torch.matmul(numer, A)
        denom *= i
        pm = pm + numer / denom
    return pm


def transform_image(image, pMtrx, refMtrx, out_h, out_w, src_h=None, src_w=None):
    """warp.py:46-86 (transformImage; src==out) and :89-129 (transformCropImage; source
    dataH x dataW, output height x W).  image [B,src_h,src_w,3], pMtrx [B,3,3], refMtrx [3,3]."""
    dt = image.dtype
    B = image.shape[0]
    src_h = image.shape[1] if src_h is None else src_h
    src_w = image.shape[2] if src_w is None else src_w
    trans = torch.matmul(refMtrx.to(dt).unsqueeze(0).expand(B, 3, 3), pMtrx.to(dt))   # :49
    X, Y = np.meshgrid(np.linspace(-1, 1, out_w), np.linspace(-1, 1, out_h))        # :51 (float64 -> f32)
    X, Y = X.flatten(), Y.flatten()
    XYhom = torch.from_numpy(np.stack([X, Y, np.ones_like(X)], 1).T.astype(np.float32)).to(dt)
    wh = torch.matmul(trans, XYhom.unsqueeze(0).expand(B, 3, -1))                   # :55
    Xw = (wh[:, 0] / (wh[:, 2] + 1e-8)).reshape(B, out_h, out_w)                    # :57
    Yw = (wh[:, 1] / (wh[:, 2] + 1e-8)).reshape(B, out_h, out_w)
    Xf, Xc = torch.floor(Xw), torch.ceil(Xw)                                        # :60
    Yf, Yc = torch.floor(Yw), torch.ceil(Yw)
    Xfi, Xci, Yfi, Yci = Xf.long(), Xc.long(), Yf.long(), Yc.long()
    vec = torch.cat([image.reshape(-1, 3), torch.zeros(1, 3, dtype=dt)], 0)         # :65-66
    bi = torch.arange(B).view(B, 1, 1)
    outside = B * src_h * src_w

    def idx(xi, yi):
        inside = (xi >= 0) & (xi < src_w) & (yi >= 0) & (yi < src_h)                # :72-73
        return torch.where(inside, (bi * src_h + yi) * src_w + xi, torch.tensor(outside))

    xr = (Xw - Xf).unsqueeze(3)                                                     # :79
    yr = (Yw - Yf).unsqueeze(3)
    ul = vec[idx(Xfi, Yfi)] * (1 - xr) * (1 - yr)                                   # :81-84
    ur = vec[idx(Xci, Yfi)] * xr * (1 - yr)
    bl = vec[idx(Xfi, Yci)] * (1 - xr) * yr
    br = vec[idx(Xci, Yci)] * xr * yr
    return ul + ur + bl + br                                                        # :85
