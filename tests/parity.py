"""Parity criteria of the sampler tests (north star: max abs error <= 1e-3 on [0,1] pixels).

strict_max_abs     for CONTINUOUS samplers (bilinear_interp of the Affine / Projective transformers,
                   spatial_transformer.py:902-964, and the Lie warp, warp.py:46-86): a rounding difference in a source
                   coordinate moves the result by O(ulp x gradient), so the bound is max-abs, full stop.
warp_max_abs       for tf_warp behind a RESIZED flow (main_dl.py:497-514): tf_warp is discontinuous where a source
                   coordinate crosses -1 or size-1 (main_dl.py:88-101: truncation + clipping make both corner weights
                   cancel, the output jumps to 0), so a 1-ulp difference between two evaluations of the flow resize
                   may legitimately produce an O(1) difference there -- and ONLY there: every pixel above the tolerance
                   must sit within `eps` pixels of such a discontinuity; everything else obeys the max-abs bound.
Both report the worst pixel when they fail."""
import torch


def _worst(diff):
    flat = int(diff.argmax())
    idx = []
    for s in reversed(diff.shape):
        idx.append(flat % s)
        flat //= s
    return tuple(reversed(idx))


def strict_max_abs(got, ref, tol, what=""):
    diff = (got.double() - ref.double()).abs()
    m = float(diff.max())
    n_bad = int((diff > tol).sum())
    assert m <= tol, f"{what}: max|err| {m:.3e} > {tol:g} at {_worst(diff)}; {n_bad} of {diff.numel()} elements above the tolerance"
    return m


def warp_max_abs(got, ref, tol, src_x, src_y, H, W, eps=2e-3, what=""):
    """src_x / src_y [B,H,W]: the oracle's source coordinates (pixel grid + flow) of every output pixel."""
    diff = (got.double() - ref.double()).abs().amax(dim=-1)                       # per pixel, worst channel
    near = ((src_x - (W - 1)).abs() < eps) | ((src_x + 1).abs() < eps) | ((src_y - (H - 1)).abs() < eps) | ((src_y + 1).abs() < eps)
    bad = diff > tol
    stray = bad & ~near
    if bool(stray.any()):
        d2 = diff * stray
        i = _worst(d2)
        raise AssertionError(f"{what}: {int(stray.sum())} pixel(s) above {tol:g} away from any tf_warp discontinuity; worst "
                             f"{float(d2.max()):.3e} at {i}, source ({float(src_x[i]):.5f}, {float(src_y[i]):.5f}) of a {H}x{W} frame")
    m = float((diff * (~near)).max()) if bool((~near).any()) else 0.0
    return m, int(bad.sum())
