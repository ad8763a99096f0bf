"""Host mirror of the reference's spatial_transformer.py classes that are on the hot path.

    AffineTransformer(out_size, name=..., interp_method='bilinear').transform(inp, theta)   :373-452
    ProjectiveTransformer(out_size, ...).transform(inp, theta)                              :519-608
    transformer(inp, theta, out_size)                                                       :34-38

Same constructor/`transform` signatures and semantics (linspace(-1,1) target grid, 1-pixel zero
border, coordinates clipped to [-1, W]); eager on torch CUDA float32 NHWC tensors.  The legacy
``transformer`` wrapper of the reference passes three arguments to a two-argument method and would
raise TypeError; the name is kept with the evidently intended two-argument behaviour.  The other
transformer classes of the file (Elastic/TPS, 3-D, bicubic, symmetry variants) have no call sites on
the inference path and are out of scope.
"""
from __future__ import annotations

import torch

from . import _lib
from .ops import _cuda_f32


class _GridSampler(object):
    param_dim = 0
    _projective = False

    def __init__(self, out_size, name="SpatialTransformer", interp_method="bilinear", **kwargs):
        if interp_method != "bilinear":
            raise NotImplementedError("only interp_method='bilinear' is on the hot path")
        self.name = name
        self.out_size = (int(out_size[0]), int(out_size[1]))
        self.interp_method = interp_method

    def transform(self, inp, theta):
        inp = _cuda_f32(inp, "inp")
        B, H, W, C = inp.shape
        theta = _cuda_f32(theta.reshape(B, self.param_dim), "theta", ndim=2)
        oh, ow = self.out_size
        out = torch.empty((B, oh, ow, C), device=inp.device, dtype=torch.float32)
        lib = _lib.load()
        fn = lib.ofs_grid_sample_projective if self._projective else lib.ofs_grid_sample_affine
        with torch.cuda.device(inp.device):
            _lib.check(fn(_lib.ptr(inp), _lib.ptr(theta), _lib.ptr(out), B, H, W, C, oh, ow,
                          _lib.current_stream_ptr(inp.device)))
        return out


class AffineTransformer(_GridSampler):
    """theta [B,6] row-major 2x3; identity = [1,0,0, 0,1,0]."""
    param_dim = 6
    _projective = False

    def __init__(self, out_size, name="SpatialAffineTransformer", interp_method="bilinear", **kwargs):
        super().__init__(out_size, name, interp_method, **kwargs)


class ProjectiveTransformer(_GridSampler):
    """theta [B,8] = first 8 entries of a row-major 3x3 whose last entry is 1."""
    param_dim = 8
    _projective = True

    def __init__(self, out_size, name="SpatialProjectiveTransformer", interp_method="bilinear", **kwargs):
        super().__init__(out_size, name, interp_method, **kwargs)


def transformer(inp, theta, out_size, name="SpatialTransformer", **kwargs):
    return AffineTransformer(out_size, name=name).transform(inp, theta)
