"""Host mirror of the reference's graph-level samplers defined inside the main scripts.

    tf_warp(img, flow, H, W)        main_flownetS_pyramid_noprevloss_dataloader.py:70-130
    get_pixel_value(img, x, y)      ...:44-68
    flow_resize / flow_resize_warp  ...:497-498 / :497-514 (test-mode glue, fused)

Same names, argument order and semantics as the reference; eager instead of graph-building:
tensors are torch CUDA float32 NHWC, the work runs on the current CUDA stream through
libofstab.so.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _lib


def _cuda_f32(t, name, ndim=4):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {t.device}; this package has no CPU fallback (CUDA sm_100a only)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    return t.contiguous()


def tf_warp(img, flow, H, W):
    """Dense backward bilinear warp with the reference's truncate-and-clip corner rule."""
    img = _cuda_f32(img, "img")
    flow = _cuda_f32(flow, "flow")
    B, h, w, C = img.shape
    if (h, w) != (int(H), int(W)):
        raise ValueError(f"tf_warp: img is {h}x{w} but H,W = {H},{W}")
    if tuple(flow.shape) != (B, h, w, 2):
        raise ValueError(f"tf_warp: flow shape {tuple(flow.shape)} != {(B, h, w, 2)}")
    out = torch.empty_like(img)
    lib = _lib.load()
    with torch.cuda.device(img.device):
        _lib.check(lib.ofs_tf_warp(_lib.ptr(img), _lib.ptr(flow), _lib.ptr(out), B, h, w, C,
                                   _lib.current_stream_ptr(img.device)))
    return out


def get_pixel_value(img, x, y):
    """img [B,H,W,C]; x, y integer [B,H,W] -> img[b, y, x]  (tf.gather_nd of the reference)."""
    img = _cuda_f32(img, "img")
    B = img.shape[0]
    b = torch.arange(B, device=img.device).view(B, 1, 1).expand_as(x)
    return img[b, y.long(), x.long()]


def flow_resize(flow2, out_h, out_w):
    """predict_flow2 [B,382,510,2] -> video-resolution flow [B,out_h,out_w,2] in video pixels."""
    flow2 = _cuda_f32(flow2, "flow2")
    B, fh, fw, two = flow2.shape
    if two != 2:
        raise ValueError("flow_resize: last dim must be 2")
    out = torch.empty((B, int(out_h), int(out_w), 2), device=flow2.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(flow2.device):
        _lib.check(lib.ofs_flow_resize(_lib.ptr(flow2), _lib.ptr(out), B, fh, fw, int(out_h), int(out_w),
                                       _lib.current_stream_ptr(flow2.device)))
    return out


def flow_resize_warp(img, flow2, out_h=None, out_w=None):
    """Fused flow_resize + tf_warp: img [B,H,W,3], flow2 [B,382,510,2] -> [B,H,W,3]."""
    img = _cuda_f32(img, "img")
    flow2 = _cuda_f32(flow2, "flow2")
    B, H, W, C = img.shape
    if C != 3:
        raise ValueError("flow_resize_warp: img must have 3 channels")
    if out_h is not None and (int(out_h), int(out_w)) != (H, W):
        raise ValueError("flow_resize_warp: out size must equal the frame size")
    if flow2.shape[0] != B or flow2.shape[3] != 2:
        raise ValueError(f"flow_resize_warp: flow2 shape {tuple(flow2.shape)} does not match batch {B}")
    out = torch.empty_like(img)
    lib = _lib.load()
    with torch.cuda.device(img.device):
        _lib.check(lib.ofs_flow_resize_warp(_lib.ptr(img), _lib.ptr(flow2), _lib.ptr(out), B, H, W,
                                            flow2.shape[1], flow2.shape[2], _lib.current_stream_ptr(img.device)))
    return out


def set_warp_variant(variant):
    """A/B switch of the tf_warp kernels (identical results): 3 = lean direct gathers (default), 0 = first direct kernel,
    1 / 2 = shared-memory staged (12-byte / padded 16-byte pixels).  3 + 16 t selects launch shape t of the fused warp
    (benchmarks/warp_tune.py)."""
    _lib.check(_lib.load().ofs_set_warp_variant(int(variant)))


def conv2d_nhwc(x, w, b=None, stride=1, transposed=False, lrelu=False, precision="bf16", block_n=0, ksplit=1, cta_group=1,
                out16=False):
    """Stand-alone run of the network's tcgen05 implicit-GEMM conv kernel (tests / microbench).

    x [B,H,W,Cin] CUDA f32; w CPU f32 TF layout ([k,k,Cin,Cout], or [4,4,Cout,Cin] when transposed);
    zero padding k//2 (conv) / TF SAME (transposed k4 s2).  Returns [B,Ho,Wo,Cout] CUDA f32.
    block_n: N tile (0 = automatic); ksplit > 1: split-K (the result then carries one 16-bit rounding, as in the
    network); cta_group: 1 single CTAs, 2 CTA pairs (tcgen05 cta_group::2), 4 pairs + slab groups (conv1 / conv2 form),
    8 / 32 two / four K chunks per pipeline stage, 16 split-K inside a thread-block cluster (DSMEM reduction, same bits
    as the workspace split-K), 5 pairs + slab groups with two output pixels per GEMM row (conv1 form, block_n 128), 64 / 66
    the phase-stacked transposed conv (cout 64) on single CTAs / CTA pairs; out16: the network's 16-bit TMA-store
    epilogue, widened to f32 afterwards.
    """
    x = _cuda_f32(x, "x")
    w = w.detach().to("cpu", torch.float32).contiguous()
    B, H, W, Cin = x.shape
    if transposed:
        k, Cout = 4, w.shape[2]
        Ho, Wo = 2 * H, 2 * W
        stride = 2
    else:
        k, Cout = w.shape[0], w.shape[3]
        p = k // 2
        Ho, Wo = (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1
    bb = None if b is None else b.detach().to("cpu", torch.float32).contiguous()
    y = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    prec = {"bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}[precision]
    with torch.cuda.device(x.device):
        _lib.check(lib.ofs_conv2d_nhwc_ex(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bb), _lib.ptr(y), B, H, W, Cin, Cout, k,
                                          int(stride), int(bool(transposed)), int(bool(lrelu)), prec, int(block_n),
                                          int(ksplit), int(cta_group), int(bool(out16)), _lib.current_stream_ptr(x.device)))
    return y
