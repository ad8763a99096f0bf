"""GPU parity of the reference's other test modes: the homography mode
(main_flownetS_pyramid_noprevloss_dataloader.py:634-751), the fixed-size mode (:758-866) and the flow post-filters of the
sibling driver (main_flownetS_pyramid.py:582-820).  OpenCV and SciPy are the references themselves here (they are
installed): the uint8 warp must match cv2 byte for byte."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from oracle import cvops
from oracle import flownet as F
from oracle import samplers as S
from oracle import tf1_ops as T
from parity import strict_max_abs, warp_max_abs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ofs(cuda_dev):
    import coupe.optical_flow_based_deep_video_stabilization_b200 as m

    m.load_library()
    return m


def _homographies(rng, n):
    out = []
    for _ in range(n):
        m = np.eye(3)
        m[:2, :2] += rng.normal(0, 0.03, (2, 2))
        m[:2, 2] = rng.normal(0, 6.0, 2)
        m[2, :2] = rng.normal(0, 2e-5, 2)
        out.append(m)
    out += [np.eye(3), np.array([[1.3, 0.2, -40.0], [-0.25, 0.8, 30.0], [4e-4, -3e-4, 1.0]]),
            np.array([[1.0, 0.0, 0.5], [0.0, 1.0, 0.25], [0.0, 0.0, 1.0]])]
    return out


@pytest.mark.parametrize("h,w", [(97, 131), (720, 1280), (64, 64), (40, 50), (1080, 1920)])
def test_warp_perspective_u8_is_cv2_byte_for_byte(ofs, cuda_dev, h, w):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dev_img = torch.from_numpy(img).to(cuda_dev)
    for m in _homographies(rng, 4 if h < 700 else 1):
        got = ofs.warp_perspective_u8(dev_img, m, (w, h)).cpu().numpy()
        np.testing.assert_array_equal(got, cv2.warpPerspective(img, m, (w, h)))
    got = ofs.warp_perspective_u8(dev_img, m, (w // 2 + 3, h // 2 + 1)).cpu().numpy()          # another output size
    np.testing.assert_array_equal(got, cv2.warpPerspective(img, m, (w // 2 + 3, h // 2 + 1)))


def test_warp_perspective_u8_batched(ofs, cuda_dev):
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (11, 90, 120, 3), dtype=np.uint8)          # > 8: two launches
    hs = np.stack(_homographies(rng, 8))
    got = ofs.warp_perspective_u8(torch.from_numpy(imgs).to(cuda_dev), hs).cpu().numpy()
    for b in range(11):
        np.testing.assert_array_equal(got[b], cv2.warpPerspective(imgs[b], hs[b], (120, 90)))
    with pytest.raises(TypeError):
        ofs.warp_perspective_u8(torch.from_numpy(imgs), hs)                # CPU tensor: no fallback


@pytest.mark.parametrize("H,W", [(90, 160), (256, 256), (720, 1280)])
def test_flow3_glue_vs_oracle(ofs, cuda_dev, H, W):
    gen = torch.Generator().manual_seed(H)
    f3 = torch.randn((2, 48, 64, 2), generator=gen) * 0.4
    ref = cvops.flow3_glue(f3, H, W)
    got = ofs.flow_resize_ex(f3.to(cuda_dev), H, W, pre_mul=float(H)).cpu()
    assert float((got - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    # pre_mul = 384 on a predict_flow2-shaped field is the regular test-mode glue
    f2 = torch.randn((1, 382, 510, 2), generator=gen) * 3.0
    a = ofs.flow_resize_ex(f2.to(cuda_dev), H, W, pre_mul=384.0)
    assert torch.equal(a, ofs.flow_resize(f2.to(cuda_dev), H, W))


def test_tf1_resize_and_cv_resize(ofs, cuda_dev):
    gen = torch.Generator().manual_seed(3)
    feats = torch.rand((2, 384, 512, 27), generator=gen)
    ref = T.resize_bilinear_tf1(feats[..., 24:27], 382, 510)                 # main_dl.py:806
    got = ofs.tf1_resize_images(feats.to(cuda_dev), (382, 510), c0=24, channels=3).cpu()
    strict_max_abs(got, ref, 1e-6, "tf1_resize_images 384x512 -> 382x510")
    img = torch.rand((2, 382, 510, 3), generator=gen)
    got = ofs.cv_resize_linear(img.to(cuda_dev), (512, 384)).cpu().numpy()
    for b in range(2):
        want = cv2.resize(img[b].numpy(), (512, 384))                        # main_dl.py:862
        assert float(np.abs(got[b] - want).max()) <= 2.5e-7
    got255 = ofs.cv_resize_linear(img.to(cuda_dev), (512, 384), post_mul=255.0).cpu().numpy()
    assert float(np.abs(got255[0] - cv2.resize(img[0].numpy(), (512, 384)) * 255).max()) <= 1e-4


def test_flow_box_blur_ema_vs_oracle(ofs, cuda_dev):
    gen = torch.Generator().manual_seed(4)
    flow = torch.randn((2, 382, 510, 2), generator=gen) * 2.0
    prev = torch.randn((2, 382, 510, 2), generator=gen)
    got = ofs.flow_box_blur_ema(flow.to(cuda_dev), prev.to(cuda_dev), k=75, a=0.9, b=0.1).cpu().numpy()
    blur = ofs.flow_box_blur_ema(flow.to(cuda_dev), None, k=75).cpu().numpy()
    for b in range(2):
        for c in range(2):
            sm = cvops.box_blur_same(flow[b, :, :, c].numpy(), 75)
            assert float(np.abs(blur[b, :, :, c] - sm).max()) <= 2e-6
            want = np.float32(0.9) * sm + np.float32(0.1) * prev[b, :, :, c].numpy()
            assert float(np.abs(got[b, :, :, c] - want).max()) <= 2e-6
    small = torch.randn((1, 20, 33, 2), generator=gen)                        # window larger than the field
    got = ofs.flow_box_blur_ema(small.to(cuda_dev), None, k=75).cpu().numpy()
    assert float(np.abs(got[0, :, :, 0] - cvops.box_blur_same(small[0, :, :, 0].numpy(), 75)).max()) <= 1e-6


def test_medfilt_vs_scipy(ofs, cuda_dev):
    import scipy.signal

    rng = np.random.default_rng(6)
    flow = rng.normal(0, 3, (60, 70, 2)).astype(np.float32)
    got = ofs.medfilt(torch.from_numpy(flow).to(cuda_dev), 5).cpu().numpy()
    np.testing.assert_array_equal(got, scipy.signal.medfilt(flow, 5))        # == 0: the window spans the channel axis
    vol = rng.normal(0, 3, (17, 19, 7)).astype(np.float32)
    for k in (3, 5):
        got = ofs.medfilt(torch.from_numpy(vol).to(cuda_dev), k).cpu().numpy()
        np.testing.assert_array_equal(got, scipy.signal.medfilt(vol, k))


@pytest.fixture(scope="module")
def small_net(ofs, cuda_dev):
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    net.assign_weights(F.make_weights(0, "calibrated", head_scale=0.02))
    yield net
    net.close()


def test_homography_mode_step_vs_host_replay(ofs, cuda_dev, small_net):
    """evaluate_originalSize_homo, frame by frame: the written frame is cv2.warpPerspective of the input with the fitted
    homography -- byte for byte; the dense-warp feedback and the homography itself are replayed on the host from the GPU's
    predict_flow3 with the oracle glue / tf_warp and cv2.findHomography."""
    H, W = 96, 128
    rng = np.random.default_rng(8)
    base = cv2.GaussianBlur(rng.integers(0, 256, (H + 8, W + 8, 3), dtype=np.uint8), (0, 0), 2.0)
    stab = ofs.HomographyStabilizer(small_net, H, W)
    for i in range(3):
        frame = np.ascontiguousarray(base[i:i + H, 2 * i:2 * i + W])
        out, h = stab.step(frame)
        np.testing.assert_array_equal(out, cv2.warpPerspective(frame, h, (W, H)))                     # main_dl.py:743,751
        f3 = stab.last["predict_flow3"].cpu()
        ref_flow = cvops.flow3_glue(f3, H, W)                                                         # :681-682
        assert float((stab.last["outflow"].cpu() - ref_flow).abs().max()) < 1e-4
        xv, yv = np.meshgrid(np.linspace(0, W - 1, W), np.linspace(0, H - 1, H))
        grid = np.stack([xv, yv], 2)
        h_ref, _ = cv2.findHomography(grid.reshape(-1, 2), (grid - ref_flow[0].numpy()).reshape(-1, 2), cv2.RANSAC)
        np.testing.assert_allclose(h / h[2, 2], h_ref / h_ref[2, 2], atol=2e-3, rtol=0)                # same fit from near-equal flows
        resized_input = torch.from_numpy((cv2.cvtColor(frame, cv2.COLOR_RGB2BGR) / 255.0).astype(np.float32))[None]
        ref_warp = S.tf_warp(resized_input, ref_flow, H, W)
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        warp_max_abs(stab.last["warped"].cpu(), ref_warp, 1e-3, xs + ref_flow[..., 0], ys + ref_flow[..., 1], H, W,
                     what=f"homography mode, frame {i}")
        assert len(stab.history) == i + 1 and stab.history[i].shape == (H, W, 3)


@pytest.mark.parametrize("flow_filter", [None, "blurNma", "medianNma"])
def test_fixed_size_modes_step_vs_host_replay(ofs, cuda_dev, small_net, flow_filter):
    """evaluate() / evaluate_blurNma() / evaluate_medianNma(): everything after the network replayed on the host from the
    GPU's predict_flow2 (oracle TF1 resize + filters + tf_warp, cv2.resize): max-abs <= 1e-3 on [0,1] pixels."""
    import scipy.signal

    rng = np.random.default_rng(9)
    base = cv2.GaussianBlur(rng.integers(0, 256, (200, 260, 3), dtype=np.uint8), (0, 0), 1.5)
    stab = ofs.FixedSizeStabilizer(small_net, flow_filter=flow_filter)
    prevof = np.zeros((382, 510, 2), np.float32)
    for i in range(2):
        frame = np.ascontiguousarray(base[i:i + 180, i:i + 240])
        side = stab.step(frame)
        assert side.shape == (384, 1024, 3) and side.dtype == np.uint8
        small = cv2.resize(frame, (512, 384))
        np.testing.assert_array_equal(side[:, :512], small)
        of = stab.last["predict_flow2"].cpu()
        cur = torch.from_numpy((cv2.cvtColor(small, cv2.COLOR_RGB2BGR) / 255.0).astype(np.float32))[None]
        unstab = T.resize_bilinear_tf1(cur, 382, 510)
        strict_max_abs(stab.last["unstabimg"].cpu(), unstab, 1e-6, "unstabimg")
        if flow_filter is None:
            flow = of
        elif flow_filter == "blurNma":
            sm = np.stack([cvops.box_blur_same(of[0, :, :, c].numpy(), 75) for c in range(2)], -1)
            flow = torch.from_numpy(np.float32(0.9) * sm + np.float32(0.1) * prevof)[None]
            prevof = (0.9 * prevof + 0.1 * of[0].numpy()).astype(np.float32)
        else:
            med = scipy.signal.medfilt(of[0].numpy(), 5)
            flow = torch.from_numpy(np.float32(0.9) * med + np.float32(0.1) * prevof)[None]
            prevof = (0.9 * prevof + 0.1 * med).astype(np.float32)
        assert float((stab.last["flow"].cpu() - flow).abs().max()) <= 1e-5
        ref_warp = S.tf_warp(unstab, flow, 382, 510)
        ys, xs = torch.meshgrid(torch.arange(382, dtype=torch.float32), torch.arange(510, dtype=torch.float32), indexing="ij")
        warp_max_abs(stab.last["warped"].cpu(), ref_warp, 1e-3, xs + flow[..., 0], ys + flow[..., 1], 382, 510,
                     what=f"fixed-size mode {flow_filter}, frame {i}")
        total = cv2.cvtColor(cv2.resize(stab.last["warped"][0].cpu().numpy(), (512, 384)) * 255, cv2.COLOR_RGB2BGR)
        assert float(np.abs(stab.history[i] - total).max()) <= 1e-3                                    # of 255
