"""GPU parity of the gather kernels (through the C ABI) against the CPU oracle and the golden
vectors.  Tolerance (BASELINE.json north star): max abs error <= 1e-3 on [0,1] pixels."""
import numpy as np
import pytest
import torch

from oracle import samplers as S
from parity import strict_max_abs, warp_max_abs

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def ofs(cuda_dev):
    import coupe.optical_flow_based_deep_video_stabilization_b200 as m

    m.load_library()
    return m


def g(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("case", ["k1", "k2", "k3", "k4"])
def test_tf_warp_golden(ofs, cuda_dev, golden, case, variant):
    ofs.set_warp_variant(variant)
    img = g(golden["k1_img"], cuda_dev)
    out = ofs.tf_warp(img, g(golden[f"{case}_flow"], cuda_dev), img.shape[1], img.shape[2])
    np.testing.assert_allclose(out.cpu().numpy(), golden[f"{case}_out"], rtol=0, atol=1e-6)
    ofs.set_warp_variant(3)


def _flows(kind, B, H, W, gen):
    if kind == "zero":
        return torch.zeros(B, H, W, 2)
    if kind == "const":
        f = torch.zeros(B, H, W, 2)
        f[..., 0], f[..., 1] = 3.3, -2.7
        return f
    if kind == "smooth":
        lo = torch.randn((B, 2, max(H // 32, 2), max(W // 32, 2)), generator=gen) * 4.0
        return torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1).contiguous()
    return (torch.rand((B, H, W, 2), generator=gen) - 0.5) * 64.0      # adversarial U(-32,32)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("kind", ["zero", "const", "smooth", "adversarial"])
@pytest.mark.parametrize("B,H,W,C", [(2, 64, 96, 3), (1, 256, 256, 3), (2, 37, 52, 3), (1, 33, 47, 3), (2, 40, 64, 5)])
def test_tf_warp_vs_oracle(ofs, cuda_dev, variant, kind, B, H, W, C):
    gen = torch.Generator().manual_seed(3)
    img = torch.rand((B, H, W, C), generator=gen)
    flow = _flows(kind, B, H, W, gen)
    ref = S.tf_warp(img, flow, H, W)
    ofs.set_warp_variant(variant)
    out = ofs.tf_warp(img.to(cuda_dev), flow.to(cuda_dev), H, W).cpu()
    ofs.set_warp_variant(3)
    assert float((out - ref).abs().max()) <= TOL


def test_tf_warp_720p_properties(ofs, cuda_dev):
    """Full BASELINE size: size-independent properties instead of a CPU comparison of every pixel.
    (a) zero flow = identity with black last row / column; (b) linearity in the image;
    (c) a 64-row crop matches the oracle."""
    gen = torch.Generator().manual_seed(5)
    B, H, W = 2, 720, 1280
    img = torch.rand((B, H, W, 3), generator=gen).to(cuda_dev)
    img2 = torch.rand((B, H, W, 3), generator=gen).to(cuda_dev)
    z = torch.zeros(B, H, W, 2, device=cuda_dev)
    out = ofs.tf_warp(img, z, H, W)
    assert torch.equal(out[:, :-1, :-1], img[:, :-1, :-1])
    assert float(out[:, -1].abs().max()) == 0.0 and float(out[:, :, -1].abs().max()) == 0.0
    flow = _flows("smooth", B, H, W, gen).to(cuda_dev)
    a = ofs.tf_warp(img, flow, H, W)
    b = ofs.tf_warp(img2, flow, H, W)
    ab = ofs.tf_warp(0.25 * img + 0.75 * img2, flow, H, W)
    assert float((ab - (0.25 * a + 0.75 * b)).abs().max()) < 1e-5
    # variant equivalence at full size
    ofs.set_warp_variant(0)
    a0 = ofs.tf_warp(img, flow, H, W)
    ofs.set_warp_variant(3)
    assert float((a0 - a).abs().max()) < 1e-6
    # exact oracle check on a window (warp of the cropped inputs differs only where taps leave the crop)
    fl = flow[:1, :64].clone()
    fl[..., 1].clamp_(-4, 4)
    ref = S.tf_warp(img[:1, :72].cpu(), torch.cat([fl.cpu(), torch.zeros(1, 8, W, 2)], 1), 72, W)
    got = ofs.tf_warp(img[:1, :72].contiguous(), torch.cat([fl, torch.zeros(1, 8, W, 2, device=cuda_dev)], 1).contiguous(), 72, W)
    assert float((got.cpu() - ref).abs().max()) <= TOL


def test_tf_warp_empty_and_errors(ofs, cuda_dev):
    out = ofs.tf_warp(torch.zeros(0, 8, 8, 3, device=cuda_dev), torch.zeros(0, 8, 8, 2, device=cuda_dev), 8, 8)
    assert out.shape == (0, 8, 8, 3)
    with pytest.raises(ValueError):
        ofs.tf_warp(torch.zeros(1, 8, 8, 3, device=cuda_dev), torch.zeros(1, 8, 9, 2, device=cuda_dev), 8, 8)
    with pytest.raises(RuntimeError):
        ofs.tf_warp(torch.zeros(1, 8, 8, 3), torch.zeros(1, 8, 8, 2), 8, 8)          # CPU tensors: no fallback
    with pytest.raises(TypeError):
        ofs.tf_warp(torch.zeros(1, 8, 8, 3, device=cuda_dev).double(), torch.zeros(1, 8, 8, 2, device=cuda_dev), 8, 8)


@pytest.mark.parametrize("H,W", [(16, 24), (256, 256), (90, 160), (45, 77)])
def test_flow_resize_and_fused_warp(ofs, cuda_dev, golden, H, W):
    gen = torch.Generator().manual_seed(H)
    f2 = torch.randn((2, 382, 510, 2), generator=gen) * 3.0
    img = torch.rand((2, H, W, 3), generator=gen)
    ref_flow = S.flow_resize(f2, H, W)
    got_flow = ofs.flow_resize(f2.to(cuda_dev), H, W).cpu()
    assert float((got_flow - ref_flow).abs().max()) < 1e-4
    ref = S.flow_resize_warp(img, f2, H, W)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    src_x, src_y = xs + ref_flow[..., 0], ys + ref_flow[..., 1]                   # the oracle's source coordinates
    for variant in (0, 1, 2, 3):
        ofs.set_warp_variant(variant)
        got = ofs.flow_resize_warp(img.to(cuda_dev), f2.to(cuda_dev)).cpu()
        # strict max-abs everywhere except within 2e-3 px of a tf_warp discontinuity (source x = -1 / W-1, y = -1 / H-1),
        # where a 1-ulp difference of the resized flow flips the output between a pixel value and 0
        m, n_disc = warp_max_abs(got, ref, TOL, src_x, src_y, H, W, what=f"fused warp variant {variant} {H}x{W}")
        assert n_disc <= 4, n_disc
    ofs.set_warp_variant(3)
    # two-step (resize then warp) == fused, same criterion
    two = ofs.tf_warp(img.to(cuda_dev), ofs.flow_resize(f2.to(cuda_dev), H, W), H, W).cpu()
    warp_max_abs(two, got, 1e-5, src_x, src_y, H, W, what="two-step vs fused")


def test_flow_glue_constant_golden(ofs, cuda_dev, golden):
    a, b = golden["k11_ab"]
    H, W = [int(v) for v in golden["k11_hw"]]
    f2 = torch.zeros(1, 382, 510, 2)
    f2[..., 0], f2[..., 1] = float(a), float(b)
    out = ofs.flow_resize(f2.to(cuda_dev), H, W).cpu()
    np.testing.assert_allclose(out[..., 0].numpy(), golden["k11_out"][0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out[..., 1].numpy(), golden["k11_out"][1], rtol=0, atol=1e-6)


@pytest.mark.parametrize("which", ["id", "shift", "far"])
def test_affine_golden(ofs, cuda_dev, golden, which):
    img = g(golden["k8_img"], cuda_dev)
    out = ofs.AffineTransformer((img.shape[1], img.shape[2])).transform(img, g(golden[f"k8_theta_{which}"], cuda_dev))
    np.testing.assert_allclose(out.cpu().numpy(), golden[f"k8_out_{which}"], rtol=0, atol=2e-5)


THETAS6 = {
    "identity": [1, 0, 0, 0, 1, 0],
    "rot5_zoom": [np.cos(np.deg2rad(5)) * 1.02, -np.sin(np.deg2rad(5)) * 1.02, 0.01, np.sin(np.deg2rad(5)) * 1.02,
                  np.cos(np.deg2rad(5)) * 1.02, -0.02],
    "wild": [0.5, -0.4, 0.9, 0.3, 1.4, -1.1],
}


@pytest.mark.parametrize("name", list(THETAS6))
@pytest.mark.parametrize("B,H,W,C,oh,ow", [(2, 48, 64, 3, 48, 64), (1, 37, 53, 3, 40, 52), (2, 30, 40, 5, 17, 23),
                                           (1, 64, 64, 1, 32, 128)])
def test_affine_vs_oracle(ofs, cuda_dev, name, B, H, W, C, oh, ow):
    gen = torch.Generator().manual_seed(17)
    img = torch.rand((B, H, W, C), generator=gen)
    th = torch.tensor([THETAS6[name]] * B, dtype=torch.float32) + torch.randn((B, 6), generator=gen) * 0.01
    ref = S.affine_transform(img, th, (oh, ow))
    got = ofs.AffineTransformer((oh, ow)).transform(img.to(cuda_dev), th.to(cuda_dev)).cpu()
    strict_max_abs(got, ref, TOL, f"AffineTransformer {name} {B}x{H}x{W}x{C} -> {oh}x{ow}")   # bilinear_interp is continuous
    got2 = ofs.transformer(img.to(cuda_dev), th.to(cuda_dev), (oh, ow)).cpu()
    assert torch.equal(got, got2)


@pytest.mark.parametrize("B,H,W,C,oh,ow", [(2, 48, 64, 3, 48, 64), (1, 37, 53, 2, 40, 52)])
def test_projective_vs_oracle(ofs, cuda_dev, B, H, W, C, oh, ow):
    gen = torch.Generator().manual_seed(19)
    img = torch.rand((B, H, W, C), generator=gen)
    th = torch.tensor([[1, 0, 0, 0, 1, 0, 0, 0]] * B, dtype=torch.float32) + torch.randn((B, 8), generator=gen) * 0.05
    ref = S.projective_transform(img, th, (oh, ow))
    got = ofs.ProjectiveTransformer((oh, ow)).transform(img.to(cuda_dev), th.to(cuda_dev)).cpu()
    strict_max_abs(got, ref, TOL, f"ProjectiveTransformer {B}x{H}x{W}x{C} -> {oh}x{ow}")


class Cfg:
    pass


def test_lie_warp_golden_and_oracle(ofs, cuda_dev, golden):
    np.testing.assert_allclose(
        ofs.vec2mtrx(_cfg("homography", 5, 2), g(golden["k10_p_tx"], cuda_dev)).cpu().numpy(), golden["k10_m_tx"], atol=1e-7)
    np.testing.assert_allclose(
        ofs.vec2mtrx(_cfg("affine", 4, 1), g(golden["k10_p_aff"], cuda_dev)).cpu().numpy(), golden["k10_m_aff"], atol=1e-7)
    with pytest.raises(AssertionError):
        ofs.vec2mtrx(_cfg("similarity", 4, 1), g(golden["k10_p_zero"], cuda_dev))
    img = g(golden["k9_img"], cuda_dev)
    cfg = _cfg("homography", 4, 2)
    cfg.height, cfg.width, cfg.refMtrx = img.shape[1], img.shape[2], golden["k9_ref"]
    out = ofs.transformImage(cfg, img, g(golden["k9_p"], cuda_dev))
    np.testing.assert_allclose(out.cpu().numpy(), golden["k9_out"], rtol=0, atol=1e-6)
    # random homographies vs oracle, plus the crop variant
    gen = torch.Generator().manual_seed(23)
    B, H, W = 2, 40, 56
    im = torch.rand((B, H, W, 3), generator=gen)
    p = torch.randn((B, 8), generator=gen) * 0.05
    ref_m = torch.tensor([[(W - 1) / 2.0, 0, (W - 1) / 2.0], [0, (H - 1) / 2.0, (H - 1) / 2.0], [0, 0, 1]])
    pm_ref = S.vec2mtrx(p, "homography", 4)
    cfg = _cfg("homography", 4, B)
    cfg.height, cfg.width, cfg.refMtrx = H, W, ref_m
    pm = ofs.vec2mtrx(cfg, p.to(cuda_dev))
    assert float((pm.cpu() - pm_ref).abs().max()) < 1e-6
    got = ofs.transformImage(cfg, im.to(cuda_dev), pm).cpu()
    ref = S.transform_image(im, pm_ref, ref_m, H, W)
    strict_max_abs(got, ref, TOL, "transformImage, random homographies")
    cfg.height, cfg.width, cfg.W, cfg.dataH, cfg.dataW, cfg.refMtrx_b = 32, 48, 48, H, W, ref_m
    got = ofs.transformCropImage(cfg, im.to(cuda_dev), pm).cpu()
    ref = S.transform_image(im, pm_ref, ref_m, 32, 48, H, W)
    strict_max_abs(got, ref, TOL, "transformCropImage, random homographies")


def _cfg(warp_type, approx, batch):
    c = Cfg()
    c.warpType, c.warpApprox, c.batch_size = warp_type, approx, batch
    return c
