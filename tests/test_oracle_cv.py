"""The oracle's restatements of the OpenCV / SciPy ops of the reference's other test modes against
the REAL libraries, which -- unlike TensorFlow -- are installed here: these restatements are pinned, byte for byte where
the arithmetic is integer.  (main_dl.py = main_flownetS_pyramid_noprevloss_dataloader.py)"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cvops


def _homographies(rng, w, h, n):
    out = []
    for _ in range(n):
        m = np.eye(3)
        m[:2, :2] += rng.normal(0, 0.03, (2, 2))
        m[:2, 2] = rng.normal(0, 6.0, 2)
        m[2, :2] = rng.normal(0, 2e-5, 2)
        out.append(m)
    out.append(np.eye(3))
    out.append(np.array([[1.3, 0.2, -40.0], [-0.25, 0.8, 30.0], [4e-4, -3e-4, 1.0]]))     # strong perspective: parts outside
    out.append(np.array([[1.0, 0.0, 0.5], [0.0, 1.0, 0.25], [0.0, 0.0, 1.0]]))            # coordinates on the rounding ties
    return out


def test_invert3x3_matches_cv2_invert_bit_for_bit():
    rng = np.random.default_rng(1)
    for m in _homographies(rng, 640, 480, 20):
        ok, inv = cv2.invert(m)
        assert ok != 0
        np.testing.assert_array_equal(cvops.invert3x3(m), inv)


def test_bilinear_table_via_cv2_remap_every_fraction():
    """Every one of the 32 x 32 fixed-point weight sets: cv2.remap with CV_16SC2 + CV_16UC1 maps runs the same
    remapBilinear on the same table; random pixel quadruples expose a weight that is off by one."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    n = 32 * 32 * 48
    frac = np.tile(np.arange(1024), 48)
    sx = rng.integers(-2, 65, n)
    sy = rng.integers(-2, 65, n)
    X = ((sx << 5) + (frac & 31)).reshape(192, -1)
    Y = ((sy << 5) + (frac >> 5)).reshape(192, -1)
    map1 = np.stack([X >> 5, Y >> 5], -1).astype(np.int16)
    map2 = (((Y & 31) << 5) + (X & 31)).astype(np.uint16)
    want = cv2.remap(img, map1, map2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    got = cvops.remap_fixed_u8(img, X.astype(np.int64), Y.astype(np.int64))
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("h,w", [(97, 131), (240, 320), (64, 64), (40, 50)])
def test_warp_perspective_u8_matches_cv2_byte_for_byte(h, w):
    rng = np.random.default_rng(h * w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for m in _homographies(rng, w, h, 6):
        want = cv2.warpPerspective(img, m, (w, h))
        got = cvops.warp_perspective_u8(img, m, (w, h))
        np.testing.assert_array_equal(got, want)


def test_warp_perspective_u8_720p_and_other_output_size():
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    m = _homographies(rng, 1280, 720, 1)[0]
    np.testing.assert_array_equal(cvops.warp_perspective_u8(img, m, (1280, 720)), cv2.warpPerspective(img, m, (1280, 720)))
    np.testing.assert_array_equal(cvops.warp_perspective_u8(img, m, (500, 300)), cv2.warpPerspective(img, m, (500, 300)))


def test_cv_resize_f32_matches_cv2():
    """main_dl.py:862: cv2.resize(warped float32 382x510, (512, 384)).  float arithmetic: cv2's vector code may fuse the
    multiply-adds, so the bound is a few ulp of a [0,1] value, not bit equality."""
    rng = np.random.default_rng(3)
    img = rng.random((382, 510, 3), dtype=np.float32)
    want = cv2.resize(img, (512, 384))
    got = cvops.cv_resize_f32(img, (512, 384))
    assert got.shape == want.shape and float(np.abs(got - want).max()) <= 2.5e-7
    img2 = rng.random((33, 47, 3), dtype=np.float32)
    assert float(np.abs(cvops.cv_resize_f32(img2, (90, 70)) - cv2.resize(img2, (90, 70))).max()) <= 2.5e-7


def test_box_blur_same_is_a_zero_padded_mean():
    """main_flownetS_pyramid.py:634-637: conv2d with a constant 1/(75*75) kernel, SAME = zero padding."""
    import torch

    rng = np.random.default_rng(4)
    x = rng.normal(0, 2, (90, 130)).astype(np.float32)
    k = 75
    want = torch.nn.functional.conv2d(torch.from_numpy(x)[None, None].double(),
                                      torch.full((1, 1, k, k), float(np.float32(1.0 / (k * k))), dtype=torch.float64),
                                      padding=k // 2)[0, 0].numpy()
    got = cvops.box_blur_same(x, k)
    assert float(np.abs(got - want).max()) <= 1e-6


def test_medfilt_matches_scipy_including_the_channel_axis():
    """main_flownetS_pyramid.py:809 calls scipy.signal.medfilt on the [382,510,2] flow with the scalar kernel size 5: the
    window is 5 x 5 x 5 -- it also spans the 2-channel axis with 3 zero pads, 75 of its 125 entries are padding zeros,
    and the median of the flow is therefore 0 everywhere.  The restatement follows scipy, quirk included."""
    import scipy.signal

    rng = np.random.default_rng(5)
    flow = rng.normal(0, 3, (40, 52, 2)).astype(np.float32)
    want = scipy.signal.medfilt(flow, 5)
    got = cvops.medfilt(flow, 5)
    np.testing.assert_array_equal(got, want)
    assert not want.any()                                   # the reference's median flow is identically zero
    plane = rng.normal(0, 3, (31, 29)).astype(np.float32)   # a genuine 2-d median for comparison
    np.testing.assert_array_equal(cvops.medfilt(plane, 5), scipy.signal.medfilt(plane, 5))
