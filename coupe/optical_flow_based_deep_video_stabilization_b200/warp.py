"""Host mirror of the reference's warp.py (Lie-algebra homography / affine warps).

    vec2mtrx(config, p)                     warp.py:25-43
    transformImage(config, image, pMtrx)    warp.py:46-86
    transformCropImage(config, image, pMtrx) warp.py:89-129
    fit / compose / inverse                 warp.py:6-23 (host-side numpy helpers)

``config`` is any object with the attributes the reference reads: warpType, warpApprox, batch_size,
refMtrx (3x3) / refMtrx_b, height, width, W, dataH, dataW.  Device work runs through libofstab.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .ops import _cuda_f32


def fit(Xsrc, Xdst):
    """Least-squares affine fit between two point sets (warp.py:6-15); host numpy."""
    import scipy.linalg

    ptsN = len(Xsrc)
    X, Y, U, V = Xsrc[:, 0], Xsrc[:, 1], Xdst[:, 0], Xdst[:, 1]
    O, I = np.zeros([ptsN]), np.ones([ptsN])
    A = np.concatenate((np.stack([X, Y, I, O, O, O], axis=1), np.stack([O, O, O, X, Y, I], axis=1)), axis=0)
    b = np.concatenate((U, V), axis=0)
    p1, p2, p3, p4, p5, p6 = scipy.linalg.lstsq(A, b)[0].squeeze()
    return np.array([[p1, p2, p3], [p4, p5, p6], [0, 0, 1]], dtype=np.float32)


def compose(config, p, dp):
    return p + dp


def inverse(config, p):
    return -p


def _ref_tensor(ref, device):
    t = ref if isinstance(ref, torch.Tensor) else torch.as_tensor(np.asarray(ref, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def vec2mtrx(config, p):
    p = _cuda_f32(p, "p", ndim=2)
    B = p.shape[0]
    if config.warpType == "homography":
        wt, dim = 0, 8
    elif config.warpType == "affine":
        wt, dim = 1, 6
    else:
        raise AssertionError("unknown warpType")            # warp.py:34 assert(False)
    if p.shape[1] != dim:
        raise ValueError(f"vec2mtrx: p must be [B,{dim}] for warpType {config.warpType}")
    out = torch.empty((B, 3, 3), device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().ofs_vec2mtrx(_lib.ptr(p), _lib.ptr(out), B, wt, int(config.warpApprox),
                                            _lib.current_stream_ptr(p.device)))
    return out


def _lie(image, pMtrx, ref, src_h, src_w, out_h, out_w):
    image = _cuda_f32(image, "image")
    B = image.shape[0]
    if tuple(image.shape[1:]) != (src_h, src_w, 3):
        raise ValueError(f"image must be [B,{src_h},{src_w},3], got {tuple(image.shape)}")   # warp.py:65 hard-codes 3
    pMtrx = _cuda_f32(pMtrx.reshape(B, 3, 3), "pMtrx", ndim=3)
    ref = _ref_tensor(ref, image.device)
    out = torch.empty((B, out_h, out_w, 3), device=image.device, dtype=torch.float32)
    with torch.cuda.device(image.device):
        _lib.check(_lib.load().ofs_lie_warp(_lib.ptr(image), _lib.ptr(pMtrx), _lib.ptr(ref), _lib.ptr(out), B,
                                            src_h, src_w, out_h, out_w, _lib.current_stream_ptr(image.device)))
    return out


def transformImage(config, image, pMtrx):
    return _lie(image, pMtrx, config.refMtrx, int(config.height), int(config.width), int(config.height),
                int(config.width))


def transformCropImage(config, image, pMtrx):
    if int(config.W) != int(config.width):
        raise ValueError("transformCropImage: the reference reshapes width*height points to [height, W]; W must equal width")
    return _lie(image, pMtrx, config.refMtrx_b, int(config.dataH), int(config.dataW), int(config.height),
                int(config.W))
