"""GPU parity of the tcgen05 implicit-GEMM convolution and of the whole FlowNetS-pyramid forward
(through the C ABI) against the CPU oracle.

Tolerances
  * single conv layer, operands pre-rounded to the 16-bit format: the only differences are the fp32
    accumulation order of the tensor core -> |err| <= 2e-3 * (1 + |ref|).
  * whole network vs the bf16-emulating oracle (same rounding points, fp64 accumulate): per-layer mean
    relative error <= 1e-2 (a wrong tap / channel / swizzle gives O(1)).
  * whole network vs the fp32 oracle (BASELINE.json north star): mean EPE(predict_flow2) <= 2e-2 px on the
    calibrated weight set (head_scale 0.02, mean |flow2| ~ 1.7 px; EPE/|flow| is reported too).
"""
import numpy as np
import pytest
import torch

from oracle import flownet as F
from oracle import tf1_ops as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ofs(cuda_dev):
    import coupe.optical_flow_based_deep_video_stabilization_b200 as m

    m.load_library()
    return m


def _round(x, prec):
    return x.to(torch.bfloat16 if prec == "bf16" else torch.float16).to(torch.float32)


CONV_CASES = [
    # B, H, W, cin, cout, k, stride, transposed
    pytest.param(2, 6, 8, 64, 16, 3, 1, False, id="tiny_k3s1"),
    pytest.param(2, 6, 8, 70, 32, 3, 1, False, id="ragged_channels"),
    pytest.param(3, 6, 8, 128, 128, 3, 1, False, id="ragged_batch_n128"),
    pytest.param(1, 48, 64, 256, 256, 3, 1, False, id="conv3_1_shape"),
    pytest.param(1, 24, 32, 128, 64, 3, 2, False, id="k3s2"),
    pytest.param(2, 32, 64, 64, 128, 5, 2, False, id="k5s2"),
    pytest.param(1, 64, 128, 27, 64, 7, 2, False, id="conv1_form_paired"),
    pytest.param(2, 6, 8, 128, 64, 4, 2, True, id="deconv"),
    pytest.param(1, 12, 16, 130, 128, 4, 2, True, id="deconv_ragged_channels"),
    pytest.param(1, 16, 32, 194, 18, 1, 1, False, id="predict2_product"),
    pytest.param(1, 12, 16, 1026, 2, 3, 1, False, id="flow_head"),
    pytest.param(1, 8, 256, 64, 64, 3, 1, False, id="two_tiles_per_row"),
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,transposed", CONV_CASES)
def test_conv_gemm_vs_oracle(ofs, cuda_dev, prec, B, H, W, cin, cout, k, stride, transposed):
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + k)
    x = _round(torch.rand((B, H, W, cin), generator=gen), prec)
    if transposed:
        w = _round(torch.randn((4, 4, cout, cin), generator=gen) * (1.0 / np.sqrt(4 * cin)), prec)
    else:
        w = _round(torch.randn((k, k, cin, cout), generator=gen) * (1.0 / np.sqrt(k * k * cin)), prec)
    b = torch.randn(cout, generator=gen) * 0.1
    for lrelu in (False, True):
        got = ofs.conv2d_nhwc(x.to(cuda_dev), w, b, stride=stride, transposed=transposed, lrelu=lrelu, precision=prec).cpu()
        if transposed:
            ref = T.conv2d_transpose_k4s2_same(x.double(), w.double(), b.double())
        else:
            ref = T.conv2d_valid(T.pad_constant(x.double(), k // 2), w.double(), b.double(), stride)
        if lrelu:
            ref = T.lrelu(ref, 0.1)
        ref = ref.float()
        assert got.shape == ref.shape
        err = (got - ref).abs()
        assert float((err / (1 + ref.abs())).max()) <= 2e-3, float(err.max())


TILING_CASES = [
    # B, H, W, cin, cout, k, stride, transposed, block_n, ksplit, cta_group
    pytest.param(4, 6, 8, 512, 256, 3, 1, False, 128, 4, 1, id="splitk4_whole_image_tiles"),
    pytest.param(3, 6, 8, 256, 128, 3, 1, False, 128, 1, 1, id="whole_image_tiles_ragged_batch"),
    pytest.param(2, 12, 16, 256, 128, 3, 2, False, 128, 3, 1, id="splitk3_k3s2"),
    pytest.param(2, 6, 8, 256, 128, 4, 2, True, 128, 5, 1, id="splitk5_deconv_uneven"),
    pytest.param(1, 48, 64, 256, 256, 3, 1, False, 256, 1, 1, id="block_n256"),
    pytest.param(2, 24, 32, 128, 512, 3, 1, False, 256, 2, 1, id="block_n256_splitk2"),
    pytest.param(1, 24, 32, 128, 64, 3, 1, False, 32, 1, 1, id="block_n32_two_n_tiles"),
    pytest.param(1, 48, 64, 128, 128, 3, 1, False, 128, 1, 2, id="pair_n128"),
    pytest.param(1, 48, 64, 256, 256, 3, 1, False, 256, 1, 2, id="pair_n256_conv3_1_shape"),
    pytest.param(2, 32, 64, 64, 128, 5, 2, False, 128, 1, 2, id="pair_k5s2"),
    pytest.param(1, 64, 128, 27, 64, 7, 2, False, 64, 1, 2, id="pair_conv1_form"),
    pytest.param(3, 6, 8, 256, 256, 3, 1, False, 256, 1, 2, id="pair_odd_tile_count_whole_images"),
    pytest.param(1, 12, 16, 130, 128, 4, 2, True, 64, 1, 2, id="pair_deconv_two_n_tiles"),
    pytest.param(2, 24, 32, 256, 512, 3, 1, False, 256, 3, 2, id="pair_n256_splitk3"),
    pytest.param(1, 16, 32, 194, 18, 1, 1, False, 32, 1, 2, id="pair_fp32_out_mode"),
    # 16-bit epilogue: swizzled shared-memory staging + TMA stores (ksplit = -1 selects out16)
    pytest.param(1, 48, 64, 128, 256, 3, 1, False, 256, -1, 1, id="tma_store_n256_two_row_tiles"),
    pytest.param(2, 24, 32, 64, 128, 3, 1, False, 64, -1, 1, id="tma_store_n64_two_n_tiles_4_row_pieces"),
    pytest.param(3, 6, 8, 128, 128, 3, 1, False, 128, -1, 1, id="tma_store_whole_image_tiles_ragged_batch"),
    pytest.param(3, 12, 16, 64, 64, 3, 1, False, 64, -1, 1, id="tma_store_ragged_tail_tile"),
    pytest.param(2, 64, 128, 27, 64, 7, 2, False, 64, -1, 1, id="tma_store_conv1_form"),
    pytest.param(2, 12, 16, 130, 128, 4, 2, True, 64, -1, 1, id="tma_store_deconv_phases"),
    pytest.param(1, 24, 32, 96, 64, 4, 2, True, 64, -1, 1, id="tma_store_deconv_pieces"),
    pytest.param(1, 48, 64, 256, 256, 3, 1, False, 256, -1, 2, id="tma_store_pair_n256"),
    pytest.param(3, 6, 8, 256, 256, 3, 1, False, 128, -1, 2, id="tma_store_pair_odd_tiles"),
    pytest.param(2, 24, 32, 128, 512, 3, 1, False, 192, -1, 1, id="tma_store_n192_padded_last_tile"),
    pytest.param(1, 24, 32, 64, 320, 3, 1, False, 192, -1, 1, id="tma_store_n192_cout320"),
    pytest.param(2, 24, 32, 128, 512, 3, 1, False, 192, -1, 2, id="tma_store_pair_n192_padded_last_tile"),
    pytest.param(3, 24, 32, 64, 320, 3, 1, False, 192, -1, 2, id="tma_store_pair_n192_cout320_odd_tiles"),
    # chunk groups (cta_group = 8: two 64-channel K blocks of a tap per pipeline stage)
    pytest.param(1, 24, 32, 192, 64, 3, 1, False, 64, -1, 8, id="kgroup_odd_chunk_count"),
    pytest.param(2, 12, 16, 130, 128, 4, 2, True, 128, -1, 8, id="kgroup_deconv"),
    pytest.param(1, 24, 32, 386, 64, 4, 2, True, 64, -1, 8, id="kgroup_deconv2_form"),
    pytest.param(3, 6, 8, 256, 128, 3, 1, False, 128, 1, 8, id="kgroup_whole_image_tiles"),
    # four K chunks per stage for 32-column fp32 tiles (cta_group = 32: the predict2 product form)
    pytest.param(2, 96, 128, 194, 18, 1, 1, False, 32, 1, 32, id="kgroup4_predict2_form"),
    pytest.param(1, 24, 32, 300, 18, 3, 1, False, 32, 1, 32, id="kgroup4_k3_five_chunks"),
    pytest.param(3, 6, 8, 64, 32, 3, 1, False, 32, 1, 32, id="kgroup4_single_chunk_whole_image_tiles"),
    # slab groups (cta_group = 4: CTA pairs, x-shifted taps share one shared-memory slab per pipeline stage)
    pytest.param(1, 32, 256, 27, 64, 7, 2, False, 64, 1, 4, id="slab_conv1_form"),
    pytest.param(2, 16, 512, 27, 64, 7, 2, False, 64, -1, 4, id="slab_conv1_form_two_x_tiles_out16"),
    pytest.param(1, 6, 256, 27, 64, 7, 2, False, 64, 1, 4, id="slab_conv1_form_odd_tile_count"),
    pytest.param(1, 16, 256, 64, 128, 5, 2, False, 128, 1, 4, id="slab_conv2_form"),
    pytest.param(3, 10, 256, 64, 128, 5, 2, False, 128, -1, 4, id="slab_conv2_form_odd_tiles_out16"),
    pytest.param(1, 8, 256, 16, 64, 3, 2, False, 64, 1, 4, id="slab_k3_paired"),
    pytest.param(1, 8, 256, 64, 64, 3, 2, False, 64, 1, 4, id="slab_k3_cin64"),
    # phase-stacked transposed conv (cta_group = 64: 1 CTA, 66: CTA pairs): the deconv2 form, cout 64
    pytest.param(1, 24, 32, 386, 64, 4, 2, True, 64, -1, 64, id="stack_deconv2_form_1cta"),
    pytest.param(1, 24, 32, 386, 64, 4, 2, True, 64, -1, 66, id="stack_deconv2_form_pairs"),
    pytest.param(3, 12, 16, 130, 64, 4, 2, True, 64, -1, 64, id="stack_1cta_two_pieces_ragged_tail"),
    pytest.param(3, 12, 16, 130, 64, 4, 2, True, 64, -1, 66, id="stack_pairs_odd_tile_count"),
    pytest.param(3, 6, 8, 64, 64, 4, 2, True, 64, -1, 66, id="stack_pairs_whole_image_tiles"),
    pytest.param(2, 48, 64, 200, 64, 4, 2, True, 64, -1, 66, id="stack_pairs_two_row_tiles_many"),
    # conv1 form with two output pixels per GEMM row (cta_group = 5: CTA pairs, quad view, 128 accumulator columns)
    pytest.param(1, 32, 512, 27, 64, 7, 2, False, 128, -1, 5, id="slab2_conv1_two_pixel_form"),
    pytest.param(2, 16, 1024, 27, 64, 7, 2, False, 128, -1, 5, id="slab2_two_x_tiles"),
    pytest.param(1, 6, 512, 27, 64, 7, 2, False, 128, -1, 5, id="slab2_odd_tile_count"),
    # tail split: > 148 M tiles of 256 columns, the last wave runs as half tiles on twice as many CTAs
    pytest.param(7, 48, 64, 64, 256, 3, 1, False, 256, -1, 1, id="tail_half_168_tiles"),
    pytest.param(7, 48, 64, 64, 256, 3, 1, False, 256, -1, 2, id="tail_half_pairs_84_pair_tiles"),
    pytest.param(7, 46, 64, 64, 256, 3, 1, False, 256, -1, 2, id="tail_half_pairs_odd_tile_count_161"),
    pytest.param(5, 96, 64, 64, 256, 3, 2, False, 256, -1, 2, id="pairs_n256_k3s2_out16"),
    # transposed conv with the fused head on single CTAs (cta_group = 34) and split-K: the deconv4 form (the head's
    # shares leave per split; the network tests check their sum)
    pytest.param(2, 12, 16, 1026, 256, 4, 2, True, 128, 3, 34, id="deconv4_form_head_splitk3"),
    pytest.param(3, 6, 8, 200, 128, 4, 2, True, 128, 2, 34, id="deconv_head_splitk2_ragged_batch"),
    pytest.param(1, 12, 16, 130, 128, 4, 2, True, 64, 4, 34, id="deconv_head_splitk4_two_n_tiles"),
    # cluster split-K (cta_group = 16: the K splits of a tile are one thread-block cluster, reduced through DSMEM)
    pytest.param(3, 6, 8, 512, 512, 3, 1, False, 256, 8, 16, id="kcluster8_whole_image_tiles_ragged_batch"),
    pytest.param(2, 12, 16, 256, 256, 3, 2, False, 256, 4, 16, id="kcluster4_k3s2"),
    pytest.param(1, 24, 32, 128, 512, 3, 1, False, 256, 6, 16, id="kcluster6_two_n_tiles"),
    pytest.param(2, 6, 8, 320, 256, 3, 1, False, 256, 7, 16, id="kcluster7_uneven_splits"),
    pytest.param(8, 6, 8, 1024, 1024, 3, 1, False, 256, 8, 16, id="kcluster8_conv6_1_shape"),
]


@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,transposed,block_n,ksplit,cta_group", TILING_CASES)
def test_conv_gemm_tilings(ofs, cuda_dev, B, H, W, cin, cout, k, stride, transposed, block_n, ksplit, cta_group):
    """Explicit block_n / split-K / CTA-pair variants (split-K output carries one bf16 rounding)."""
    gen = torch.Generator().manual_seed(77 + block_n + ksplit)
    x = _round(torch.rand((B, H, W, cin), generator=gen), "bf16")
    shape = (4, 4, cout, cin) if transposed else (k, k, cin, cout)
    w = _round(torch.randn(shape, generator=gen) * (1.0 / np.sqrt(k * k * cin)), "bf16")
    b = torch.randn(cout, generator=gen) * 0.1
    out16 = ksplit < 0
    ksplit = max(ksplit, 1)
    got = ofs.conv2d_nhwc(x.to(cuda_dev), w, b, stride=stride, transposed=transposed, lrelu=True, precision="bf16",
                          block_n=block_n, ksplit=ksplit, cta_group=cta_group, out16=out16).cpu()
    if transposed:
        ref = T.conv2d_transpose_k4s2_same(x.double(), w.double(), b.double())
    else:
        ref = T.conv2d_valid(T.pad_constant(x.double(), k // 2), w.double(), b.double(), stride)
    ref = T.lrelu(ref, 0.1).float()
    tol = 6e-3 if (ksplit > 1 or out16) else 2e-3
    assert float(((got - ref).abs() / (1 + ref.abs())).max()) <= tol


@pytest.mark.parametrize("cta_group", [1, 2])
def test_tail_half_tiles_bit_identical(ofs, cuda_dev, cta_group, monkeypatch):
    """Splitting the last wave's tiles into two 128-column halves does not touch the K order: same bits."""
    gen = torch.Generator().manual_seed(11)
    x = _round(torch.rand((8, 48, 64, 128), generator=gen), "bf16").to(cuda_dev)   # 192 M tiles (conv3_1's grid)
    w = _round(torch.randn((3, 3, 128, 256), generator=gen) * (1.0 / np.sqrt(9 * 128)), "bf16")
    b = torch.randn(256, generator=gen) * 0.1
    monkeypatch.setenv("OFS_TAIL_HALF", "0")
    a = ofs.conv2d_nhwc(x, w, b, lrelu=True, precision="bf16", block_n=256, cta_group=cta_group, out16=True).cpu()
    monkeypatch.setenv("OFS_TAIL_HALF", "1")
    c = ofs.conv2d_nhwc(x, w, b, lrelu=True, precision="bf16", block_n=256, cta_group=cta_group, out16=True).cpu()
    assert torch.equal(a, c)


@pytest.mark.parametrize("B,H,W,cin,cout,ks", [(8, 12, 16, 512, 512, 4), (8, 6, 8, 512, 1024, 8), (3, 6, 8, 256, 256, 5)])
def test_cluster_splitk_bit_identical_to_workspace_splitk(ofs, cuda_dev, B, H, W, cin, cout, ks):
    """The DSMEM reduction adds bias + the partial tiles in split order, as splitk_reduce_kernel does: same bits."""
    gen = torch.Generator().manual_seed(5 + ks)
    x = _round(torch.rand((B, H, W, cin), generator=gen), "bf16").to(cuda_dev)
    w = _round(torch.randn((3, 3, cin, cout), generator=gen) * (1.0 / np.sqrt(9 * cin)), "bf16")
    b = torch.randn(cout, generator=gen) * 0.1
    a = ofs.conv2d_nhwc(x, w, b, lrelu=True, precision="bf16", block_n=256, ksplit=ks, cta_group=1).cpu()
    c = ofs.conv2d_nhwc(x, w, b, lrelu=True, precision="bf16", block_n=256, ksplit=ks, cta_group=16).cpu()
    assert torch.equal(a, c)


@pytest.mark.parametrize("B,H,W,cin,cout,ks", [(8, 12, 16, 512, 512, 6), (8, 6, 8, 1024, 1024, 8), (8, 24, 32, 512, 512, 6),
                                               (3, 6, 8, 256, 256, 5), (1, 6, 8, 512, 1024, 8)])
def test_fused_splitk_reduce_bit_identical_to_reduce_kernel(ofs, cuda_dev, monkeypatch, B, H, W, cin, cout, ks):
    """Split-K reduced inside the GEMM launch (arrival counters + dynamically claimed rows) sums bias + the partials in
    split order exactly as splitk_reduce_kernel does: same bits, launch after launch (the counters clean themselves)."""
    gen = torch.Generator().manual_seed(9 + ks)
    x = _round(torch.rand((B, H, W, cin), generator=gen), "bf16").to(cuda_dev)
    w = _round(torch.randn((3, 3, cin, cout), generator=gen) * (1.0 / np.sqrt(9 * cin)), "bf16")
    b = torch.randn(cout, generator=gen) * 0.1
    stride = 2 if H == 24 else 1
    monkeypatch.setenv("OFS_FUSED_REDUCE", "0")
    a = ofs.conv2d_nhwc(x, w, b, stride=stride, lrelu=True, precision="bf16", block_n=256, ksplit=ks, cta_group=1).cpu()
    monkeypatch.setenv("OFS_FUSED_REDUCE", "1")
    for _ in range(3):
        c = ofs.conv2d_nhwc(x, w, b, stride=stride, lrelu=True, precision="bf16", block_n=256, ksplit=ks, cta_group=1).cpu()
        assert torch.equal(a, c)


@pytest.fixture(scope="module")
def net_case(ofs, cuda_dev):
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    x = F.make_feats(2, 2)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=2, precision="bf16")
    net.assign_weights(w)
    out = net.forward(x.to(cuda_dev))
    torch.cuda.synchronize()
    return w, x, net, out


ACTS = ["conv1", "conv2", "conv3", "conv3_1", "conv4", "conv4_1", "conv5", "conv5_1", "conv6", "conv6_1",
        "concat5", "concat4", "concat3", "concat2"]


def test_network_per_layer_vs_bf16_emulating_oracle(net_case):
    w, x, net, out = net_case
    ref = F.forward_folded(x, F.fold_bn(w), emulate_bf16=True, keep=True)
    report = []
    for name in ["input"] + ACTS:
        got = net.activation(name, 2).cpu()
        exp = (F.round_bf16(x) if name == "input" else ref["_acts"][name]).float()
        assert got.shape == exp.shape, name
        rel = float((got - exp).abs().mean() / (exp.abs().mean() + 1e-12))
        report.append((name, rel))
        assert rel <= 1e-2, (name, rel, report)
    for lvl in (6, 5, 4, 3, 2):
        k = f"predict_flow{lvl}"
        e = F.epe(out[k].cpu(), ref[k])
        mag = float(torch.sqrt((ref[k] ** 2).sum(-1)).mean())
        assert e <= 1e-2 * max(mag, 0.1) + 1e-3, (k, e, mag)


def test_network_epe_vs_fp32_oracle_bf16(net_case):
    w, x, net, out = net_case
    ref = F.forward_literal(x, w)
    mag = float(torch.sqrt((ref["predict_flow2"] ** 2).sum(-1)).mean())
    e = F.epe(out["predict_flow2"].cpu(), ref["predict_flow2"])
    print(f"bf16: mean EPE(predict_flow2) = {e:.5f} px, mean |flow2| = {mag:.3f} px, EPE/|flow| = {e / mag:.4%}")
    assert e <= 2e-2, (e, mag)
    assert out["flow"].data_ptr() == out["predict_flow2"].data_ptr()
    for lvl, hw in {6: (6, 8), 5: (12, 16), 4: (24, 32), 3: (48, 64), 2: (382, 510)}.items():
        assert tuple(out[f"predict_flow{lvl}"].shape) == (2,) + hw + (2,)


def test_network_epe_fp16_operands_larger_flows(ofs, cuda_dev):
    """fp16 operands (same tensor rate) hold the 2e-2 px bound at 4x larger flows (|flow2| ~ 7 px)."""
    w = F.make_weights(0, "calibrated", head_scale=0.08)
    x = F.make_feats(4, 1)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1, precision="fp16")
    net.assign_weights(w)
    out = net.forward(x.to(cuda_dev))
    ref = F.forward_literal(x, w)
    mag = float(torch.sqrt((ref["predict_flow2"] ** 2).sum(-1)).mean())
    e = F.epe(out["predict_flow2"].cpu(), ref["predict_flow2"])
    print(f"fp16: mean EPE(predict_flow2) = {e:.5f} px, mean |flow2| = {mag:.3f} px, EPE/|flow| = {e / mag:.4%}")
    assert e <= 2e-2, (e, mag)
    net.close()


def test_network_he_init_relative_epe(ofs, cuda_dev):
    """Reference initialisers (stress case, |flow2| tens of pixels): report and bound EPE/|flow|."""
    w = F.make_weights(0, "he")
    x = F.make_feats(6, 1)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1, precision="bf16")
    net.assign_weights(w)
    out = net.forward(x.to(cuda_dev))
    ref = F.forward_literal(x, w)
    mag = float(torch.sqrt((ref["predict_flow2"] ** 2).sum(-1)).mean())
    e = F.epe(out["predict_flow2"].cpu(), ref["predict_flow2"])
    print(f"he-init bf16: EPE = {e:.4f} px, |flow2| = {mag:.2f} px, EPE/|flow| = {e / mag:.4%}")
    assert e / mag <= 1.5e-2
    net.close()


def test_precision_options_report(ofs, cuda_dev):
    """the operand precision is a per-net option.  Reports mean EPE(predict_flow2) against the fp32 oracle for
    bf16 and fp16 operands on the calibrated set (|flow2| ~ 1.8 px, the regime of a trained stabiliser) and on the
    reference's own initialisers (|flow2| ~ 90 px), and pins what each buys: fp16 operands (same tensor rate, 3 more
    mantissa bits) hold the calibrated set with a 4x margin under 2e-2 px and cut the He-init error several-fold."""
    res = {}
    for wname, w, seed in (("calibrated", F.make_weights(0, "calibrated", head_scale=0.02), 2), ("he", F.make_weights(0, "he"), 6)):
        x = F.make_feats(seed, 1)
        ref = F.forward_literal(x, w)["predict_flow2"]
        mag = float(torch.sqrt((ref ** 2).sum(-1)).mean())
        for prec in ("bf16", "fp16"):
            net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1, precision=prec)
            net.assign_weights(w)
            e = F.epe(net.forward(x.to(cuda_dev))["predict_flow2"].cpu(), ref)
            net.close()
            res[(wname, prec)] = (e, mag)
            print(f"{wname:10s} {prec}: EPE = {e:.5f} px, |flow2| = {mag:.2f} px, EPE/|flow| = {e / mag:.4%}")
    assert res[("calibrated", "bf16")][0] <= 2e-2
    assert res[("calibrated", "fp16")][0] <= 5e-3
    assert res[("he", "fp16")][0] <= 0.5 * res[("he", "bf16")][0]
    assert res[("he", "fp16")][0] / res[("he", "fp16")][1] <= 3e-3


def test_batch_independence_and_determinism(net_case, cuda_dev):
    """Frame pairs are independent units: a pair's result must not depend on its batch neighbours."""
    w, x, net, out = net_case
    o1 = net.forward(x[1:2].to(cuda_dev))
    assert torch.equal(o1["predict_flow2"], net.forward(x[1:2].to(cuda_dev))["predict_flow2"])
    torch.testing.assert_close(o1["predict_flow2"][0], out["predict_flow2"][1], rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        net.forward(torch.zeros(3, 384, 512, 27, device=cuda_dev))                   # > max_batch
    with pytest.raises(ValueError):
        net.forward(torch.zeros(1, 256, 256, 27, device=cuda_dev))                   # model.py:850 hard-wires 384x512


def test_reference_signature_and_stabilize(ofs, cuda_dev, net_case):
    """flownetS_pyramid(feats, batch_size, is_train=False) + tf_warp == fused stabilize == host-buffer call."""
    from oracle import samplers as S

    w, x, net, out = net_case
    with pytest.raises(RuntimeError):
        ofs.flownetS_pyramid(x.to(cuda_dev), 2, scope="never_loaded")
    ofs.assign_weights(w, scope="flownetS", device=cuda_dev, max_batch=2)
    o = ofs.flownetS_pyramid(x.to(cuda_dev), 2, is_train=False)
    assert set(o) == {"predict_flow6", "predict_flow5", "predict_flow4", "predict_flow3", "predict_flow2", "flow"}
    torch.testing.assert_close(o["predict_flow2"], out["predict_flow2"], rtol=0, atol=0)
    with pytest.raises(NotImplementedError):
        ofs.flownetS_pyramid(x.to(cuda_dev), 2, is_train=True)
    H, W = 256, 256                                                                  # BASELINE configs[0] frame size
    gen = torch.Generator().manual_seed(1)
    frames = torch.rand((2, H, W, 3), generator=gen)
    outflow = ofs.flow_resize(o["predict_flow2"], H, W)
    two_step = ofs.tf_warp(frames.to(cuda_dev), outflow, H, W)
    fused, f2 = net.stabilize(x.to(cuda_dev), frames.to(cuda_dev), return_flow=True)
    torch.testing.assert_close(f2, out["predict_flow2"], rtol=0, atol=0)
    from parity import warp_max_abs

    ref_flow = S.flow_resize(out["predict_flow2"].cpu(), H, W)                       # the oracle's resized flow of the GPU's flow2
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    src_x, src_y = xs + ref_flow[..., 0], ys + ref_flow[..., 1]
    warp_max_abs(fused.cpu(), two_step.cpu(), 1e-5, src_x, src_y, H, W, what="fused vs two-step")
    host = net.stabilize_host(x.pin_memory(), frames.pin_memory())
    torch.testing.assert_close(host, fused.cpu(), rtol=0, atol=0)
    ref = S.flow_resize_warp(frames, out["predict_flow2"].cpu(), H, W)               # oracle warp on the GPU's flow
    # max-abs <= 1e-3 everywhere except within 2e-3 px of a tf_warp discontinuity (parity.py)
    _, n_disc = warp_max_abs(fused.cpu(), ref, 1e-3, src_x, src_y, H, W, what="stabilize vs oracle warp")
    assert n_disc <= 4, n_disc


def test_missing_weight_is_an_error(ofs, cuda_dev):
    w = F.make_weights(0, "he")
    del w["4_1/W_conv2d"]
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    with pytest.raises(ofs.OfstabError):
        net.assign_weights(w)
    with pytest.raises(ofs.OfstabError):
        net.forward(torch.zeros(1, 384, 512, 27, device=cuda_dev))                   # forward before weights
    # names with the checkpoint's scope prefix and ':0' suffix are accepted
    w = {f"main_net/flownetS/{k}:0": v for k, v in F.make_weights(0, "he").items()}
    net.assign_weights(w)
    net.close()


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W", [(8, 720, 1280), (16, 1080, 1920)], ids=["configs1_b8_720p", "configs2_b16_1080p"])
def test_full_size_step_properties(ofs, cuda_dev, B, H, W):
    """BASELINE configs[1] / configs[2] at full size, through size-independent properties (the oracle needs ~0.1 s per
    pair, so only pair 0 is compared with it): batch invariance (pair i alone == pair i inside the batch, bit for bit,
    although the batch rides a different CUDA graph and different split-K / tile schedules per launch), replay
    determinism, host-buffer call == device call, warp linearity in the image, and the oracle on one pair."""
    from oracle import samplers as S

    w = F.make_weights(0, "calibrated", head_scale=0.02)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=B)
    net.assign_weights(w)
    feats = F.make_feats(11, B)
    gen = torch.Generator().manual_seed(12)
    frames = torch.rand((B, H, W, 3), generator=gen)
    fd, gd = feats.to(cuda_dev), frames.to(cuda_dev)
    out, f2 = net.stabilize(fd, gd, return_flow=True)
    out2, _ = net.stabilize(fd, gd, return_flow=True)                              # graph replay
    assert torch.equal(out, out2)
    for i in (0, B // 2, B - 1):
        oi, fi = net.stabilize(fd[i:i + 1].contiguous(), gd[i:i + 1].contiguous(), return_flow=True)
        assert torch.equal(fi[0], f2[i]) and torch.equal(oi[0], out[i])
    host = net.stabilize_host(feats.pin_memory(), frames.pin_memory())
    assert torch.equal(host, out.cpu())
    # linearity of the warp in the image for a fixed flow
    g2 = torch.rand((B, H, W, 3), generator=gen).to(cuda_dev)
    a = ofs.flow_resize_warp(gd[:2], f2[:2], H, W)
    b = ofs.flow_resize_warp(g2[:2], f2[:2], H, W)
    ab = ofs.flow_resize_warp(0.25 * gd[:2] + 0.75 * g2[:2], f2[:2], H, W)
    assert float((ab - (0.25 * a + 0.75 * b)).abs().max()) < 1e-5
    assert float((a - out[:2]).abs().max()) < 1e-5                                   # stand-alone fused op == network path
    # oracle on pair 0
    ref = F.forward_literal(feats[:1], w)
    e = F.epe(f2[:1].cpu(), ref["predict_flow2"])
    assert e <= 2e-2, e                                                              # north-star tolerance, px at bf16
    ref_img = S.flow_resize_warp(frames[:1], f2[:1].cpu(), H, W)
    from parity import warp_max_abs

    ref_flow = S.flow_resize(f2[:1].cpu(), H, W)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    _, n_disc = warp_max_abs(out[:1].cpu(), ref_img, 1e-3, xs + ref_flow[..., 0], ys + ref_flow[..., 1], H, W,
                             what="720p stabilize vs oracle warp")                  # 1e-3 on [0,1] pixels (parity.py)
    assert n_disc <= 8, n_disc
    net.close()


def test_host_call_wire_format_does_not_change_results(ofs, cuda_dev, monkeypatch):
    """ofs_net_stabilize_host with OFS_HOST_PACK=1 sends the network input as bf16 rounded on the host (half the PCIe
    bytes); by default it sends float32 and pack_act_kernel rounds it on the device: same bits either way, odd batch."""
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    gen = torch.Generator().manual_seed(21)
    B, H, W = 3, 240, 320
    feats, frames = F.make_feats(9, B), torch.rand((B, H, W, 3), generator=gen)
    outs, wire = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("OFS_HOST_PACK", flag)
        net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=4, precision="bf16")
        net.assign_weights(w)
        outs.append(net.stabilize_host(feats.pin_memory(), frames.pin_memory()))
        outs.append(net.stabilize_host(feats, frames))                                # pageable buffers, second call: graph replay
        wire.append(int(net._lib.ofs_net_host_h2d_bytes(net._h, B, H, W)))
        net.close()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert wire[0] == B * (384 * 512 * 27 * 2 + H * W * 3 * 4) and wire[1] == B * (384 * 512 * 27 * 4 + H * W * 3 * 4)


@pytest.mark.parametrize("B", [1, 3, 8])
def test_layer_chain_bit_identical_to_separate_launches(ofs, cuda_dev, monkeypatch, B):
    """conv5 ... conv6_1 (model.py:829-845) as ONE cooperative persistent launch (split-K GEMM phases and all-CTA
    reductions separated by grid barriers, conv_chain_kernel; opt-in with OFS_CHAIN=1, measured slower) gives the bits of
    the eight separate launches: same K order per split, same split order in the sums.  Repeated calls reuse the self-maintained
    barrier words; the step-graph path (stabilize) replays the chain from a CUDA graph; two nets on two streams run their
    chains concurrently (the launch is cooperative: no partial residency)."""
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    x = F.make_feats(5, B).to(cuda_dev)
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("OFS_CHAIN", flag)
        net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=B, precision="bf16")
        net.assign_weights(w)
        outs = []
        for _ in range(3):
            o = net.forward(x)
            outs.append({k: o[k].clone() for k in ("predict_flow6", "predict_flow2")})
            acts = {k: net.activation(k, B).clone() for k in ("conv5", "conv5_1", "conv6", "conv6_1")}
        for o in outs[1:]:
            assert torch.equal(o["predict_flow2"], outs[0]["predict_flow2"])
        res[flag] = (outs[0], acts, net.launches_per_forward)
        if flag == "1":
            frames = torch.rand((B, 96, 128, 3), device=cuda_dev)
            s0 = net.stabilize(x, frames)
            s1 = net.stabilize(x, frames)                      # graph replay
            assert torch.equal(s0, s1)
            if B == 8:
                net2 = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=B, precision="bf16")
                net2.assign_weights(w)
                st1, st2 = torch.cuda.Stream(device=cuda_dev), torch.cuda.Stream(device=cuda_dev)
                torch.cuda.synchronize()
                got = []
                for _ in range(6):
                    with torch.cuda.stream(st1):
                        a = net.forward(x)["predict_flow2"].clone()
                    with torch.cuda.stream(st2):
                        b = net2.forward(x)["predict_flow2"].clone()
                    got += [a, b]
                torch.cuda.synchronize()
                for g in got:
                    assert torch.equal(g, outs[0]["predict_flow2"])
                net2.close()
        net.close()
    assert res["0"][2] - res["1"][2] == 7, (res["0"][2], res["1"][2])   # 4 GEMMs + 4 reductions -> 1 launch
    for k in ("conv5", "conv5_1", "conv6", "conv6_1"):
        assert torch.equal(res["0"][1][k], res["1"][1][k]), k
    for k in ("predict_flow6", "predict_flow2"):
        assert torch.equal(res["0"][0][k], res["1"][0][k]), k


def test_deconv_splitk_with_fused_head_matches_default(ofs, cuda_dev, monkeypatch, net_case):
    """A transposed conv with the fused flow head can run split-K (OFS_TUNE=deconv4:128:3:1: single CTAs, three K splits):
    every split leaves its share of the head in its own plane and pyr_kernel sums the planes in split order.  The flows
    then differ from the default tiling only by fp32 summation order (and one bf16 rounding of deconv4's output)."""
    w, x, net, _ = net_case
    out = {k: v.clone() for k, v in net.forward(x.to(cuda_dev)).items()}      # (the fixture's own result may be stale:
    c4a = net.activation("concat4", 2).cpu()                                  #  other tests run this net on other inputs)
    monkeypatch.setenv("OFS_TUNE", "deconv4:128:3:1,deconv5:128:2:1")
    net2 = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=2, precision="bf16")
    net2.assign_weights(w)
    o2 = net2.forward(x.to(cuda_dev))
    assert net2.launches_per_forward == net.launches_per_forward + 2          # two more split-K reductions
    for lvl in (6, 5, 4, 3, 2):
        k = f"predict_flow{lvl}"
        a, b = out[k].cpu(), o2[k].cpu()
        mag = float(torch.sqrt((a ** 2).sum(-1)).mean())
        assert F.epe(b, a) <= 2e-3 * max(mag, 0.1) + 1e-4, (k, F.epe(b, a), mag)
    c4b = net2.activation("concat4", 2).cpu()
    assert float((c4a - c4b).abs().max()) <= 2e-2 * float(c4a.abs().max())
    net2.close()


def test_stacked_deconv2_matches_per_phase_form(ofs, cuda_dev, monkeypatch, net_case):
    """deconv2 + predict3 in the phase-stacked form (OFS_STACK=1: one 304-column accumulator tile, nine taps fetched once)
    against the default per-phase form at network level: same products, another fp32 summation order."""
    w, x, net, _ = net_case
    out = {k: v.clone() for k, v in net.forward(x.to(cuda_dev)).items()}
    c2a = net.activation("concat2", 2).cpu()
    monkeypatch.setenv("OFS_STACK", "1")
    net2 = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=2, precision="bf16")
    net2.assign_weights(w)
    o2 = net2.forward(x.to(cuda_dev))
    for lvl in (3, 2):
        k = f"predict_flow{lvl}"
        a, b = out[k].cpu(), o2[k].cpu()
        mag = float(torch.sqrt((a ** 2).sum(-1)).mean())
        assert F.epe(b, a) <= 2e-3 * max(mag, 0.1) + 1e-4, (k, F.epe(b, a), mag)
    for lvl in (6, 5, 4):
        assert torch.equal(out[f"predict_flow{lvl}"], o2[f"predict_flow{lvl}"])
    c2b = net2.activation("concat2", 2).cpu()
    assert float((c2a - c2b).abs().max()) <= 2e-2 * float(c2a.abs().max())
    net2.close()


def test_input_pack_kernels_agree(ofs, cuda_dev, monkeypatch):
    """The network input [B,384,512,27] float32 is rounded to 16-bit channels-32 pixels by pack27_kernel (a warp streams 32
    pixels through shared memory with 128-bit loads) when the caller's array is 16-byte aligned, and by the generic
    pack_act_kernel otherwise: same bits, and the step graph keeps one capture per alignment class."""
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    B = 2
    x = F.make_feats(12, B)
    buf = torch.empty(x.numel() + 4, device=cuda_dev)
    aligned = buf[4:].view(x.shape) if buf.data_ptr() % 16 == 0 else buf[:x.numel()].view(x.shape)
    shifted = buf[1:1 + x.numel()].view(x.shape)
    assert aligned.data_ptr() % 16 == 0 and shifted.data_ptr() % 16 != 0
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=B, precision="bf16")
    net.assign_weights(w)
    frames = torch.rand((B, 96, 128, 3), device=cuda_dev)
    res = []
    for view in (shifted, aligned):          # views of one buffer: fill, run, read back before the next fill
        view.copy_(x.to(cuda_dev))
        out = net.forward(view)["predict_flow2"].clone()
        act = net.activation("input", B).clone()
        stab = net.stabilize(view, frames).clone()
        res.append((out, act, stab))
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][1], F.round_bf16(x).to(cuda_dev))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][2], res[1][2])
    assert net.graph_stats()[0] == 2         # one capture per alignment class of feats
    net.close()


def test_npz_checkpoint_ingest(ofs, cuda_dev, tmp_path):
    """tl.files.load_and_assign_npz_dict (main_dl.py:520): an npz keyed by TF variable names
    ('main_net/flownetS/<layer>/<var>:0', as tl.files.save_npz_dict writes them, main_dl.py:424-426), with
    non-trivial BN statistics, gives the same flows as handing the arrays over directly."""
    w = F.make_weights(3, "calibrated", head_scale=0.02)
    path = tmp_path / "flownetS_pyramid.npz"
    np.savez(path, **{f"main_net/flownetS/{k}:0": np.asarray(v, dtype=np.float32) for k, v in w.items()})
    x = F.make_feats(5, 1).to(cuda_dev)
    direct = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    direct.assign_weights(w)
    want = direct.forward(x)
    ofs.load_and_assign_npz_dict(name=str(path), sess=None, scope="npz_test", device=cuda_dev, max_batch=1)
    got = ofs.flownetS_pyramid(x, 1, is_train=False, scope="npz_test")
    for k in want:
        assert torch.equal(got[k], want[k]), k
    direct.close()


def test_checkpoint_ingest_is_strict(ofs, cuda_dev):
    """A checkpoint that lacks a BatchNorm statistic / a bias, or carries an array no variable of flownetS_pyramid consumes
    (model.py:786-893 has gamma_init=None), is refused with the offending key named -- never loaded with defaults."""
    from coupe.optical_flow_based_deep_video_stabilization_b200._lib import OfstabError

    w = F.make_weights(0, "calibrated", head_scale=0.02)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=1)
    for missing in ("3_1/moving_variance", "deconv4_bn/beta", "4/b_conv2d", "predict5/b_conv2d", "upsample4_3/b_deconv2d"):
        bad = {k: v for k, v in w.items() if k != missing}
        with pytest.raises(OfstabError, match=missing):
            net.assign_weights(bad)
        with pytest.raises(OfstabError):                 # a failed ingest leaves the net unloaded
            net.forward(F.make_feats(5, 1).to(cuda_dev))
    extra = dict(w)
    extra["main_net/flownetS/2/gamma:0"] = np.ones(128, np.float32)
    with pytest.raises(OfstabError, match="2/gamma"):
        net.assign_weights(extra)
    net.assign_weights(w)                                 # and a complete one still loads afterwards
    assert net.forward(F.make_feats(5, 1).to(cuda_dev))["flow"].shape == (1, 382, 510, 2)
    net.close()


def test_stabilize_replays_one_graph_for_fresh_tensors(ofs, cuda_dev):
    """The reference feeds a NEW array every frame (main_dl.py:568-569) and the drop-in allocates a fresh output per
    call: 100 calls with freshly allocated tensors must be served by ONE captured graph (re-pointed, not
    re-captured), and give exactly what the plain stream-launched path gives."""
    w = F.make_weights(0, "calibrated", head_scale=0.02)
    net = ofs.FlowNetSPyramid(device=cuda_dev, max_batch=2)
    net.assign_weights(w)
    g = torch.Generator().manual_seed(11)
    H, W = 96, 128
    base_feats = F.make_feats(9, 2)
    base_frames = torch.rand((2, H, W, 3), generator=g)
    keep, outs = [], []
    for i in range(100):
        feats = (base_feats if i % 2 == 0 else base_feats.flip(0)).to(cuda_dev).clone()
        frames = (base_frames if i % 2 == 0 else base_frames.flip(0)).to(cuda_dev).clone()
        keep.append((feats, frames))                                   # keep them alive: every call sees new addresses
        if i % 3 == 0:
            keep.append(torch.empty(1 + 7 * i, device=cuda_dev))       # shift the allocator
        outs.append(net.stabilize(feats, frames))
    torch.cuda.synchronize()
    captures, updates, cached = net.graph_stats()
    assert captures == 1 and cached == 1, (captures, updates, cached)
    assert updates >= 90, updates
    assert len({o.data_ptr() for o in outs}) > 50
    for i in (2, 3, 98, 99):
        assert torch.equal(outs[i], outs[i % 2]), i                    # same inputs -> same bits, whatever the addresses
    assert torch.equal(outs[1], outs[0].flip(0))                       # batch order follows the inputs
    # with the flow output as well (another cached graph), and against the un-graphed path
    out_a, flow_a = net.stabilize(keep[0][0], keep[0][1], return_flow=True)
    out_b, flow_b = net.stabilize(keep[0][0].clone(), keep[0][1].clone(), return_flow=True)
    assert torch.equal(out_a, outs[0]) and torch.equal(out_b, outs[0]) and torch.equal(flow_a, flow_b)
    assert net.graph_stats()[0] == 2
    ref = net.forward(keep[0][0])["predict_flow2"]
    assert torch.equal(flow_a, ref)
    net.close()
