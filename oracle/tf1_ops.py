"""Vectorised (torch, CPU) restatement of the TF-1.10 / TensorLayer-1.x ops that the
reference hot path calls.  Test infrastructure only -- see ``oracle/__init__.py``.

Third-party semantics restated here (not vendored under /root/reference; the
only pin is README.md:5 "Tensorflow 1.10"; TensorLayer 1.8-1.11 API):

  tf.nn.conv2d(NHWC, VALID) on an explicit tf.pad     <- model.py:807-885
  tf.nn.conv2d_transpose(k4, s2, SAME)                <- model.py:850,852,...
  tf.nn.batch_normalization(gamma=None, eps=1e-5)     <- model.py:809 (BatchNormLayer, is_train=False)
  tl.act.lrelu(x, 0.1) = max(x, 0.1 x)                <- model.py:788
  tf.image.resize_images(BILINEAR, align_corners=False) legacy (no half-pixel)
                                                      <- model.py:857,866,875,886, main:497
  tf.image.resize_images(NEAREST, align_corners=True) <- model.py:795-802,883

All tensors are NHWC like the reference.  ``dtype`` follows the input (fp32 for the
oracle proper, fp64 for the noise-floor cross-check).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # TensorLayer BatchNormLayer default epsilon


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def pad_constant(x, p):
    """PadLayer(n, [[0,0],[p,p],[p,p],[0,0]], "constant")  (model.py:807)."""
    return F.pad(x, (0, 0, p, p, p, p))


def conv2d_valid(x, w_tf, b, stride):
    """tl.layers.Conv2d(..., padding='VALID', act=None): tf.nn.conv2d + bias.

    w_tf: [kh, kw, cin, cout] (TF filter layout), b: [cout] or None.
    """
    w = w_tf.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(_nchw(x), w, b, stride=stride)
    return _nhwc(y)


def conv2d_transpose_k4s2_same(x, w_tf, b):
    """DeConv2dLayer(shape=(4,4,cout,cin), strides=(1,2,2,1)) (padding default SAME).

    y[2i+ky-1, 2j+kx-1, co] += x[i,j,ci] * W[ky,kx,co,ci];  w_tf: [4,4,cout,cin].
    """
    w = w_tf.permute(3, 2, 0, 1).contiguous()  # [cin, cout, kh, kw]
    y = F.conv_transpose2d(_nchw(x), w, b, stride=2, padding=1)
    return _nhwc(y)


def batchnorm_infer(x, mean, var, beta):
    """BatchNormLayer(is_train=False, gamma_init=None): no gamma."""
    return (x - mean) * torch.rsqrt(var + BN_EPS) + beta


def lrelu(x, alpha=0.1):
    return torch.maximum(x, alpha * x)


def bilinear_tables(in_size, out_size):
    """Index / lerp tables of TF-1.10 ResizeBilinear (align_corners=False, legacy).

    scale = in/out in float32; src = i*scale (float32); lo = floor; hi = min(lo+1,in-1).
    """
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.arange(out_size, dtype=np.float32) * scale
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(lo + 1, in_size - 1)
    lerp = (src - lo.astype(np.float32)).astype(np.float32)
    return lo, hi, lerp


def resize_bilinear_tf1(x, out_h, out_w):
    """tf.image.resize_images(x, [out_h,out_w]) method=BILINEAR, align_corners=False (TF1 legacy)."""
    B, H, W, C = x.shape
    if (H, W) == (out_h, out_w):
        return x.clone()
    ylo, yhi, yl = bilinear_tables(H, out_h)
    xlo, xhi, xl = bilinear_tables(W, out_w)
    yl_t = torch.from_numpy(yl).to(x.dtype).view(1, out_h, 1, 1)
    xl_t = torch.from_numpy(xl).to(x.dtype).view(1, 1, out_w, 1)
    ylo_t, yhi_t = torch.from_numpy(ylo), torch.from_numpy(yhi)
    xlo_t, xhi_t = torch.from_numpy(xlo), torch.from_numpy(xhi)
    top = x[:, ylo_t]
    bot = x[:, yhi_t]
    tl_, tr_ = top[:, :, xlo_t], top[:, :, xhi_t]
    bl_, br_ = bot[:, :, xlo_t], bot[:, :, xhi_t]
    t = tl_ + (tr_ - tl_) * xl_t
    b = bl_ + (br_ - bl_) * xl_t
    return t + (b - t) * yl_t


def nearest_align_table(in_size, out_size):
    """TF-1.10 ResizeNearestNeighbor(align_corners=True): min(roundf(i*(in-1)/(out-1)), in-1)."""
    scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
    src = np.arange(out_size, dtype=np.float32) * scale
    # roundf = half away from zero; src >= 0 so floor(src+0.5) in float32
    idx = np.floor(src + np.float32(0.5)).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def resize_nearest_tf1_align(x, out_h, out_w):
    B, H, W, C = x.shape
    yi = torch.from_numpy(nearest_align_table(H, out_h))
    xi = torch.from_numpy(nearest_align_table(W, out_w))
    return x[:, yi][:, :, xi]
