"""Host mirror of the reference's warp.py (Lie-algebra homography / affine warps).

    vec2mtrx(config, p)                     warp.py:25-43
    transformImage(config, image, pMtrx)    warp.py:46-86
    transformCropImage(config, image, pMtrx) warp.py:89-129
    fit / compose / inverse                 warp.py:6-23 (host-side helpers, written independently: QR least squares)

``config`` is any object with the attributes the reference reads: warpType, warpApprox, batch_size,
refMtrx (3x3) / refMtrx_b, height, width, W, dataH, dataW.  Device work runs through libofstab.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .ops import _cuda_f32


def fit(Xsrc, Xdst):
    """Affine map (3x3, last row 0 0 1) taking the points Xsrc [N,2] onto Xdst [N,2] in the least-squares sense
    (same contract as the reference helper warp.py:6-15).  The two output coordinates are independent problems over
    the same design matrix [x y 1], so one QR-based solve with a two-column right-hand side covers both."""
    src = np.asarray(Xsrc, dtype=np.float64).reshape(-1, 2)
    dst = np.asarray(Xdst, dtype=np.float64).reshape(-1, 2)
    if src.shape != dst.shape or src.shape[0] < 3:
        raise ValueError("fit: need two [N,2] point sets with N >= 3")
    design = np.column_stack([src, np.ones(len(src))])          # [N,3]
    rows, *_ = np.linalg.lstsq(design, dst, rcond=None)         # [3,2]: column j holds the row of output coordinate j
    M = np.eye(3, dtype=np.float32)
    M[:2, :] = rows.T
    return M


def compose(config, p, dp):
    """Additive composition of warp parameters (first-order in the Lie algebra), as the reference does (warp.py:17-19)."""
    return torch.add(p, dp) if isinstance(p, torch.Tensor) else np.add(p, dp)


def inverse(config, p):
    """Parameter vector of the inverse warp under the same first-order approximation (warp.py:21-23)."""
    return torch.neg(p) if isinstance(p, torch.Tensor) else np.negative(p)


_REF_CACHE = {}


def _ref_tensor(ref, device):
    """config.refMtrx on the device.  The reference keeps it as a numpy constant of the config; uploading it on every
    call is a synchronous pageable copy (~30 us, the whole cost of a small transformImage), so host matrices are cached
    by value per device."""
    if isinstance(ref, torch.Tensor):
        return ref.to(device=device, dtype=torch.float32).contiguous()
    arr = np.ascontiguousarray(ref, dtype=np.float32)
    key = (arr.tobytes(), arr.shape, str(device))
    t = _REF_CACHE.get(key)
    if t is None:
        if len(_REF_CACHE) >= 64:
            _REF_CACHE.clear()
        t = _REF_CACHE[key] = torch.as_tensor(arr).to(device=device)
    return t


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
