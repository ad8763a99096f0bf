"""CPU restatement of the OpenCV / SciPy / TF ops the reference's OTHER test modes call around the network
(SURVEY.md 8(f) row 4).  Test infrastructure only -- see ``oracle/__init__.py``.

  warp_perspective_u8     cv2.warpPerspective(frame, h, (w, h)), INTER_LINEAR, BORDER_CONSTANT 0
                          (evaluate_originalSize_homo, main_flownetS_pyramid_noprevloss_dataloader.py:743)
  invert3x3               cv::invert of the 3x3 double matrix that warpPerspective applies first
  flow3_glue              predict_flow3 * out_h / 48 -> TF1 resize -> per-axis scale      (main_dl.py:681-682)
  cv_resize_f32           cv2.resize(float32 image, (512, 384)), INTER_LINEAR               (main_dl.py:862)
  box_blur_same           tf.nn.conv2d(x, const 1/(75*75) [75,75,1,1], SAME)                (main_flownetS_pyramid.py:634-637)
  medfilt                 scipy.signal.medfilt(volume, k): k x k x k window, zero padded    (main_flownetS_pyramid.py:809)

UNLIKE the TensorFlow ops, OpenCV and SciPy ARE installed here, so these restatements are PINNED: tests/test_oracle_cv.py
checks every one of them against the real library (byte for byte for the uint8 warp) on this machine.

OpenCV's uint8 linear warp is fixed point: source coordinates are rounded to 1/32 pixel (cvRound of a double), the four
bilinear weights come from a 32 x 32 table of 15-bit integers that sum to 32768, and the result is
(sum + 2^14) >> 15.  Coordinates are evaluated in 64-wide blocks as X0 + M0 * x1 with X0 = M0 * x_block + M1 * y + M2
(imgwarp.cpp, WarpPerspectiveInvoker) -- restated in that operation order, in float64.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
INTER_REMAP_COEF_BITS = 15
INTER_REMAP_COEF_SCALE = 1 << INTER_REMAP_COEF_BITS
WARP_BLOCK_W = 64          # bw0 of WarpPerspectiveInvoker for images at least 64 wide and 16 high


def _sat_short(v):
    return int(min(max(int(np.rint(v)), -32768), 32767))


def bilinear_tab():
    """initInterTab2D(INTER_LINEAR, fixpt=true): int16 weights [32 (fy)][32 (fx)][4] in the order (y0x0, y0x1, y1x0, y1x1)."""
    one = np.float32(1.0)
    scale = np.float32(1.0 / INTER_TAB_SIZE)
    c1 = [(one - np.float32(i) * scale, np.float32(i) * scale) for i in range(INTER_TAB_SIZE)]
    tab = np.zeros((INTER_TAB_SIZE, INTER_TAB_SIZE, 4), np.int32)
    for i in range(INTER_TAB_SIZE):
        for j in range(INTER_TAB_SIZE):
            w = [np.float32(c1[i][k1] * c1[j][k2]) for k1 in range(2) for k2 in range(2)]
            iw = [_sat_short(np.float32(v) * np.float32(INTER_REMAP_COEF_SCALE)) for v in w]
            diff = sum(iw) - INTER_REMAP_COEF_SCALE
            # (32768 * (a/32) * (b/32) = 32 a b is an exact integer for every fraction, so the four weights always sum to
            # 32768 and only the fraction (0, 0) ever meets the int16 saturation: 32768 -> 32767, deficit 1.  Whether the
            # deficit is put back or not, (w * p + 2^14) >> 15 == p for every byte p when the other weights are 0.)
            if diff != 0:
                k = int(np.argmax(iw)) if diff < 0 else int(np.argmin(iw))
                iw[k] -= diff
            tab[i, j] = iw
    return tab


_TAB = None


def _tab():
    global _TAB
    if _TAB is None:
        _TAB = bilinear_tab()
    return _TAB


def invert3x3(m):
    """cv::invert of a 3x3 double matrix (closed form: adjugate times 1/det, each entry (a*b - c*d) * d)."""
    m = np.asarray(m, np.float64).reshape(3, 3)
    det = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0]) +
           m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if det == 0.0:
        return np.zeros((3, 3))
    d = 1.0 / det
    t = np.empty(9)
    t[0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[3] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[4] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[5] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[6] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[7] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[8] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t.reshape(3, 3)


def warp_perspective_coords(M, out_h, out_w, block_w=None):
    """Fixed-point source coordinates (X, Y in 1/32 px, int64) of every destination pixel for the INVERSE map M."""
    M = np.asarray(M, np.float64).reshape(9)
    bw = min(WARP_BLOCK_W, out_w) if block_w is None else block_w
    xs = np.arange(out_w)
    xb = (xs // bw) * bw
    x1 = (xs - xb).astype(np.float64)[None, :]
    xb = xb.astype(np.float64)[None, :]
    y = np.arange(out_h).astype(np.float64)[:, None]
    X0 = M[0] * xb + M[1] * y + M[2]
    Y0 = M[3] * xb + M[4] * y + M[5]
    W0 = M[6] * xb + M[7] * y + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide="ignore", invalid="ignore"):
        W = np.where(W != 0, INTER_TAB_SIZE / np.where(W != 0, W, 1.0), 0.0)
    lo, hi = float(-2 ** 31), float(2 ** 31 - 1)
    fX = np.maximum(lo, np.minimum(hi, (X0 + M[0] * x1) * W))
    fY = np.maximum(lo, np.minimum(hi, (Y0 + M[3] * x1) * W))
    return np.rint(fX).astype(np.int64), np.rint(fY).astype(np.int64)     # cvRound: nearest, ties to even


def remap_fixed_u8(img, X, Y):
    """remapBilinear<FixedPtCast<int, uchar, 15>> with BORDER_CONSTANT 0 on fixed-point coordinates X, Y (1/32 px)."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape[:2]
    sx = np.clip(X >> INTER_BITS, -32768, 32767)      # saturate_cast<short>
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    wts = _tab()[(Y & (INTER_TAB_SIZE - 1)), (X & (INTER_TAB_SIZE - 1))]          # [H,W,4]
    acc = np.zeros(X.shape + (img.shape[2],), np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        px = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64)
        acc += np.where(ok[..., None], px, 0) * wts[..., k][..., None]
    return np.clip((acc + (1 << (INTER_REMAP_COEF_BITS - 1))) >> INTER_REMAP_COEF_BITS, 0, 255).astype(np.uint8)


def warp_perspective_u8(img, H, dsize):
    """cv2.warpPerspective(img, H, dsize) for a uint8 HxWxC image: default flags (INTER_LINEAR, forward matrix H that is
    inverted first), BORDER_CONSTANT with value 0.  dsize = (width, height)."""
    out_w, out_h = dsize
    X, Y = warp_perspective_coords(invert3x3(H), out_h, out_w)
    return remap_fixed_u8(img, X, Y)


# ------------------------------------------------------------------------------ homography-mode flow glue
def flow3_glue(flow3, out_h, out_w):
    """main_dl.py:681-682: resize_images(predict_flow3 * out_h / 48, [out_h, out_w]) then x * out_w / 512, y * out_h / 384.
    flow3: torch [B,48,64,2] float32."""
    import torch

    from . import tf1_ops as T

    f = flow3 * float(out_h) / flow3.shape[1]
    f = T.resize_bilinear_tf1(f, out_h, out_w)
    return torch.cat([f[..., 0:1] * out_w / 512, f[..., 1:2] * out_h / 384], 3)


# ------------------------------------------------------------------------------ cv2.resize, float32, INTER_LINEAR
def cv_resize_f32(img, dsize):
    """cv2.resize(img_f32 [H,W,C], (dw, dh)) with INTER_LINEAR (resizeGeneric_ HResizeLinear / VResizeLinear, float):
    half-pixel source coordinates, coefficients in float32, horizontal pass then vertical pass, each s0*c0 + s1*c1."""
    img = np.asarray(img, np.float32)
    sh, sw = img.shape[:2]
    dw, dh = dsize

    def coeffs(dn, sn):
        scale = 1.0 / (float(dn) / float(sn))                     # double, as resize() computes it
        idx = np.empty(dn, np.int64)
        a = np.empty((dn, 2), np.float32)
        for d in range(dn):
            fd = (d + 0.5) * scale - 0.5                           # double: cv2 4.13's results need the FRACTION taken in
            s = int(np.floor(fd))                                  # double and only then rounded to float32 (checked
            f = np.float32(fd - s)                                 # against cv2.resize in tests/test_oracle_cv.py)
            if s < 0:
                f, s = np.float32(0), 0
            if s + 1 >= sn:                                        # the second tap would be outside: clamp, weight 0
                f, s = np.float32(0), sn - 1
            idx[d] = s
            a[d] = (np.float32(1.0) - f, f)
        return idx, a

    xi, xa = coeffs(dw, sw)
    yi, ya = coeffs(dh, sh)
    x1 = np.minimum(xi + 1, sw - 1)
    y1 = np.minimum(yi + 1, sh - 1)
    rows = img[:, xi] * xa[None, :, 0, None] + img[:, x1] * xa[None, :, 1, None]            # horizontal pass, float32
    rows = rows.astype(np.float32)
    out = rows[yi] * ya[:, None, 0, None] + rows[y1] * ya[:, None, 1, None]
    return out.astype(np.float32)


# ------------------------------------------------------------------------------ flow post-filters of the sibling driver
def box_blur_same(x, k=75):
    """tf.nn.conv2d(x[..., None], constant(1/(k*k)) [k,k,1,1], SAME) of a [H,W] float32 plane: every tap has the weight
    float32(1/(k*k)); zero padding.  float64 accumulation of the float32 products (the TF kernel's summation order is
    not specified; the test tolerance covers it)."""
    x = np.asarray(x, np.float32)
    h, w = x.shape
    r = k // 2
    wgt = np.float32(1.0 / (k * k))
    prod = (x * wgt).astype(np.float32).astype(np.float64)
    pad = np.zeros((h + 2 * r + 1, w + 2 * r + 1))
    pad[r + 1: r + 1 + h, r + 1: r + 1 + w] = prod
    ii = pad.cumsum(0).cumsum(1)
    out = ii[k:, k:] - ii[:-k, k:] - ii[k:, :-k] + ii[:-k, :-k]
    return out[:h, :w].astype(np.float32)


def medfilt(vol, k=5):
    """scipy.signal.medfilt(vol, k) of an N-d array with a scalar kernel size: a k x ... x k window in EVERY dimension
    (including a trailing channel axis), zero padded, median = element (k**nd) // 2 of the sorted window."""
    vol = np.asarray(vol)
    r = k // 2
    pad = np.pad(vol, r)
    wins = []
    import itertools

    for off in itertools.product(range(k), repeat=vol.ndim):
        sl = tuple(slice(o, o + s) for o, s in zip(off, vol.shape))
        wins.append(pad[sl])
    st = np.sort(np.stack(wins, 0), axis=0)
    return st[len(wins) // 2].astype(vol.dtype)
