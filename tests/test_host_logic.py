"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
16-bit conversions, and the conv plan (tap tables, TMA view, tiling, weight packing) emulated in
numpy against the oracle convolution.  No compute call reaches a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import emu
from oracle import tf1_ops as T

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    from coupe.optical_flow_based_deep_video_stabilization_b200 import _lib

    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from coupe.optical_flow_based_deep_video_stabilization_b200 import _lib

    header = open(os.path.join(ROOT, "include", "ofstab.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ofs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ofstab.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototype table and header disagree"
    assert lib.ofs_version() == 100


def test_host_wire_packing_is_the_device_rounding(lib):
    """ofs_net_stabilize_host rounds the float32 network input to bf16 on the host (worker pool, AVX2 or scalar) before it
    crosses PCIe: the bits must be those of cvt.rn.bf16.f32 -- torch's float32 -> bfloat16 cast -- for ordinary values,
    ties, denormals, infinities and NaN, whatever the worker count and the array length."""
    import ctypes as C

    rng = np.random.RandomState(3)
    base = np.concatenate([rng.rand(100003).astype(np.float32), (rng.randn(5000) * 1e3).astype(np.float32),
                           np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 3.3895314e38, 65504.0], np.float32)])
    ties = (np.arange(4096, dtype=np.uint32) << 16 | 0x8000).view(np.float32)            # exactly half way: ties to even
    near = ((np.arange(4096, dtype=np.uint32) << 16) | 0x7fff).view(np.float32)
    x = np.ascontiguousarray(np.concatenate([base, ties, near]))
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    for threads, n in ((1, len(x)), (5, len(x)), (3, 1000), (4, 63), (2, 0)):
        got = np.full(len(x), 0xABCD, np.uint16)
        assert lib.ofs_debug_host_pack_bf16(x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), n, threads) == 0
        nan = np.isnan(x[:n])
        np.testing.assert_array_equal(got[:n][~nan], want[:n][~nan])
        assert ((got[:n][nan] & 0x7fff) > 0x7f80).all()                                   # NaN stays NaN
        assert (got[n:] == 0xABCD).all()                                                  # nothing past the end is touched


def test_no_cpu_fallback_without_gpu(lib):
    import coupe.optical_flow_based_deep_video_stabilization_b200 as ofs

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.ofs_device_check(0) != 0
    assert b"no CPU fallback" in lib.ofs_last_error() or b"sm_" in lib.ofs_last_error()
    with pytest.raises(RuntimeError):
        ofs.tf_warp(torch.zeros(1, 4, 4, 3), torch.zeros(1, 4, 4, 2), 4, 4)
    with pytest.raises(RuntimeError):
        ofs.FlowNetSPyramid()
    with pytest.raises(RuntimeError):
        ofs.AffineTransformer((4, 4)).transform(torch.zeros(1, 4, 4, 3), torch.zeros(1, 6))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "coupe")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_cvt16_matches_torch(lib):
    rng = np.random.RandomState(0)
    vals = np.concatenate([rng.randn(2000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1.0, 300.0, 7e4)])
    vals = np.concatenate([vals, np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 6.1e-5, 5.96e-8, 2.98e-8,
                                           2.9802322e-8, 8.9e-8, np.inf, -np.inf], np.float32)])
    tv = torch.from_numpy(vals)
    exp_bf = tv.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
    exp_fp = tv.to(torch.float16).view(torch.int16).numpy().astype(np.uint16)
    got_bf = np.array([lib.ofs_debug_cvt16(float(v), 1) for v in vals], np.uint16)
    got_fp = np.array([lib.ofs_debug_cvt16(float(v), 0) for v in vals], np.uint16)
    np.testing.assert_array_equal(got_bf, exp_bf)
    np.testing.assert_array_equal(got_fp, exp_fp)


# kind, B, H, W, cin, in_cs, cout, k, stride, block_n
CONV_CASES = [
    pytest.param(0, 2, 6, 8, 70, 72, 16, 3, 1, 16, id="k3s1_ragged_channels_tile_spans_images"),
    pytest.param(0, 3, 6, 8, 64, 64, 32, 3, 1, 32, id="k3s1_ragged_batch"),
    pytest.param(0, 1, 12, 16, 64, 72, 32, 3, 2, 32, id="k3s2_slice_of_wider_buffer"),
    pytest.param(0, 2, 16, 32, 64, 64, 16, 5, 2, 16, id="k5s2"),
    pytest.param(0, 1, 16, 32, 27, 32, 64, 7, 2, 64, id="k7s2_paired_conv1_form"),
    pytest.param(0, 1, 12, 16, 20, 32, 16, 3, 2, 16, id="k3s2_paired"),
    pytest.param(1, 2, 6, 8, 70, 72, 32, 4, 2, 32, id="deconv_k4s2"),
    pytest.param(1, 1, 12, 16, 130, 136, 64, 4, 2, 64, id="deconv_k4s2_3chunks"),
    pytest.param(0, 1, 8, 16, 194, 200, 18, 1, 1, 32, id="k1s1_predict2_product"),
    pytest.param(0, 1, 4, 256, 64, 64, 16, 3, 1, 16, id="wide_rows_two_tiles_per_row"),
]


SPLITK_CASES = [
    pytest.param(0, 3, 6, 8, 256, 256, 32, 3, 1, 32, 4, id="splitk4_whole_image_tiles"),
    pytest.param(0, 2, 12, 16, 128, 136, 16, 3, 2, 16, 3, id="splitk3_k3s2"),
    pytest.param(1, 2, 6, 8, 130, 136, 32, 4, 2, 32, 5, id="splitk5_deconv_uneven"),
]


@pytest.mark.parametrize("kind,B,H,W,cin,in_cs,cout,k,stride,bn,ks", SPLITK_CASES)
def test_conv_plan_splitk_emulation(lib, kind, B, H, W, cin, in_cs, cout, k, stride, bn, ks):
    test_conv_plan_emulation_matches_oracle(lib, kind, B, H, W, cin, in_cs, cout, k, stride, bn, ks)


@pytest.mark.parametrize("kind,B,H,W,cin,in_cs,cout,k,stride,bn", CONV_CASES)
def test_conv_plan_emulation_matches_oracle(lib, kind, B, H, W, cin, in_cs, cout, k, stride, bn, ks=1):
    rng = np.random.RandomState(1234 + k * 7 + stride)
    x = emu.bf16_round(rng.rand(B, H, W, cin).astype(np.float32))
    if kind == 0:
        w = emu.bf16_round(rng.randn(k, k, cin, cout).astype(np.float32) * 0.1)
    else:
        w = emu.bf16_round(rng.randn(4, 4, cout, cin).astype(np.float32) * 0.1)
    b = rng.randn(cout).astype(np.float32)
    plan = emu.get_plan(lib, kind, B, H, W, cin, in_cs, cout, k, stride, bn, w, b, ksplit=ks)
    if ks > 1:
        assert 1 < plan["ksplit"] <= ks
    act = np.zeros((B, H, W, in_cs), np.float32)
    act[..., :cin] = x
    act[..., cin:] = 7.0  # other layers' channels in the same buffer must never leak into this GEMM
    got = emu.emulate(plan, act)[..., :cout]
    xt, wt, bt = torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double()
    if kind == 0:
        ref = T.conv2d_valid(T.pad_constant(xt, k // 2), wt, bt, stride)
    else:
        ref = T.conv2d_transpose_k4s2_same(xt, wt, bt)
    assert got.shape == tuple(ref.shape)
    assert not np.isnan(got).any(), "some output pixel was never written"
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=1e-5)


SLAB_CASES = [
    pytest.param(1, 8, 256, 27, 32, 64, 7, 64, id="conv1_form_k7_paired"),
    pytest.param(2, 6, 512, 27, 32, 64, 7, 64, id="conv1_form_two_x_tiles_two_images"),
    pytest.param(1, 8, 256, 64, 64, 128, 5, 128, id="conv2_form_k5"),
    pytest.param(1, 6, 256, 16, 32, 64, 3, 64, id="k3_paired"),
    pytest.param(1, 6, 256, 64, 64, 64, 3, 64, id="k3_cin64"),
]


@pytest.mark.parametrize("B,H,W,cin,in_cs,cout,k,bn", SLAB_CASES)
def test_slab_group_plan_emulation_matches_oracle(lib, B, H, W, cin, in_cs, cout, k, bn):
    """Slab groups (conv1 / conv2): one TMA slab per group of x-shifted taps, tap t = the slab rows [off_t, off_t + 128),
    weights packed in group-major K order.  The numpy emulation of exactly that must equal the oracle stride-2 conv."""
    rng = np.random.RandomState(99 + k)
    x = emu.bf16_round(rng.rand(B, H, W, cin).astype(np.float32))
    w = emu.bf16_round(rng.randn(k, k, cin, cout).astype(np.float32) * 0.1)
    b = rng.randn(cout).astype(np.float32)
    plan = emu.get_plan_ex(lib, 0, B, H, W, cin, in_cs, cout, k, 2, bn, w, b, slab=True)
    assert plan["group_max"] <= 4 and plan["slab_extra"] <= 7 and plan["tiles_mp"] == (plan["tiles_m"] + 1) // 2
    assert plan["k_total"] == 64 * int(sum(plan["grp_n"][:plan["ntaps"]]))
    act = np.zeros((B, H, W, in_cs), np.float32)
    act[..., :cin] = x
    got, _ = emu.emulate_ex(plan, act)
    ref = T.conv2d_valid(T.pad_constant(torch.from_numpy(x).double(), k // 2), torch.from_numpy(w).double(),
                         torch.from_numpy(b).double(), 2)
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got[..., :cout], ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,H,W,cin,in_cs,cout,bn", [(2, 6, 8, 70, 72, 128, 128), (1, 12, 16, 130, 136, 64, 64),
                                                      (1, 12, 16, 96, 96, 256, 128)],
                         ids=["whole_image_tiles", "three_chunks", "two_n_tiles"])
def test_fused_head_plan_emulation_matches_oracle(lib, B, H, W, cin, in_cs, cout, bn):
    """Transposed conv with the level's 3x3 flow head fused as 16 extra accumulator columns: the deconv output must equal
    the oracle conv2d_transpose and the 4 phase shares of every input pixel must add up to the oracle 3x3 head."""
    rng = np.random.RandomState(5 + cout)
    x = emu.bf16_round(rng.rand(B, H, W, cin).astype(np.float32))
    w = emu.bf16_round(rng.randn(4, 4, cout, cin).astype(np.float32) * 0.1)
    hw = emu.bf16_round(rng.randn(3, 3, cin, 2).astype(np.float32) * 0.1)
    b = rng.randn(cout).astype(np.float32)
    plan = emu.get_plan_ex(lib, 1, B, H, W, cin, in_cs, cout, 4, 2, bn, w, b, head_w=hw)
    assert plan["w_rows_phase"] == plan["n_pad"] + 16 and plan["w_rows"] == 4 * plan["w_rows_phase"]
    act = np.zeros((B, H, W, in_cs), np.float32)
    act[..., :cin] = x
    act[..., cin:] = 7.0
    got, shares = emu.emulate_ex(plan, act)
    xt = torch.from_numpy(x).double()
    ref = T.conv2d_transpose_k4s2_same(xt, torch.from_numpy(w).double(), torch.from_numpy(b).double())
    np.testing.assert_allclose(got[..., :cout], ref.numpy(), rtol=1e-5, atol=1e-5)
    head = shares[:, 0::2, 0::2] + shares[:, 0::2, 1::2] + shares[:, 1::2, 0::2] + shares[:, 1::2, 1::2]
    href = T.conv2d_valid(T.pad_constant(xt, 1), torch.from_numpy(hw).double(), torch.zeros(2).double(), 1)
    np.testing.assert_allclose(head, href.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,H,W", [(1, 8, 512), (2, 6, 1024)], ids=["one_x_tile", "two_x_tiles_two_images"])
def test_two_pixel_slab_plan_emulation_matches_oracle(lib, B, H, W):
    """conv1 form with two output pixels per GEMM row (quad view of the 32-channel input, 5 quad taps per kernel row,
    [pixel 2m | pixel 2m+1] accumulator columns): the emulated tiles, read back as [B, H/2, W/2, 64], equal the oracle."""
    rng = np.random.RandomState(23)
    cin, in_cs, cout, k = 27, 32, 64, 7
    x = emu.bf16_round(rng.rand(B, H, W, cin).astype(np.float32))
    w = emu.bf16_round(rng.randn(k, k, cin, cout).astype(np.float32) * 0.1)
    b = rng.randn(cout).astype(np.float32)
    plan = emu.get_plan_ex(lib, 0, B, H, W, cin, in_cs, cout, k, 2, 128, w, b, slab2=True)
    assert plan["n_pad"] == 128 and plan["Wg"] == W // 4 and plan["ntaps"] == 14 and plan["k_total"] == 35 * 64
    act = np.zeros((B, H, W, in_cs), np.float32)
    act[..., :cin] = x
    got, _ = emu.emulate_ex(plan, act)                       # [B, H/2, W/4, 128]
    got = got.reshape(B, H // 2, W // 2, cout)               # the same memory as NHWC [B, H/2, W/2, 64]
    ref = T.conv2d_valid(T.pad_constant(torch.from_numpy(x).double(), k // 2), torch.from_numpy(w).double(),
                         torch.from_numpy(b).double(), 2)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,H,W,cin,in_cs", [(2, 6, 8, 70, 72), (1, 12, 16, 130, 136), (1, 24, 32, 64, 64)],
                         ids=["whole_image_tiles", "three_chunks_two_pieces", "four_row_tiles"])
def test_stacked_deconv_plan_emulation_matches_oracle(lib, B, H, W, cin, in_cs):
    """Phase-stacked transposed conv (deconv2 form, cout 64): 9 taps fetched once each, 11 weight-row entries into a
    304-column accumulator.  The four phase column blocks must equal the oracle conv2d_transpose, the three head copies
    must add up to the oracle 3x3 head, and the first two MMAs must cover every accumulator column exactly once."""
    rng = np.random.RandomState(17 + cin)
    cout = 64
    x = emu.bf16_round(rng.rand(B, H, W, cin).astype(np.float32))
    w = emu.bf16_round(rng.randn(4, 4, cout, cin).astype(np.float32) * 0.1)
    hw = emu.bf16_round(rng.randn(3, 3, cin, 2).astype(np.float32) * 0.1)
    b = rng.randn(cout).astype(np.float32)
    plan = emu.get_plan_ex(lib, 1, B, H, W, cin, in_cs, cout, 4, 2, 64, w, b, head_w=hw, stack=True)
    act = np.zeros((B, H, W, in_cs), np.float32)
    act[..., :cin] = x
    act[..., cin:] = 7.0      # pad channels of a concat buffer may hold anything finite: their weights are zero
    got, shares = emu.emulate_stack(plan, act)
    xt = torch.from_numpy(x).double()
    ref = T.conv2d_transpose_k4s2_same(xt, torch.from_numpy(w).double(), torch.from_numpy(b).double())
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=1e-5)
    head = shares[:, 0::2, 0::2] + shares[:, 0::2, 1::2] + shares[:, 1::2, 0::2] + shares[:, 1::2, 1::2]
    href = T.conv2d_valid(T.pad_constant(xt, 1), torch.from_numpy(hw).double(), torch.zeros(2).double(), 1)
    np.testing.assert_allclose(head, href.numpy(), rtol=1e-5, atol=1e-5)


def test_network_layer_plans_are_valid(lib):
    """Geometry of the 15 GEMM layers (default 1-CTA tilings; the flow heads ride in the deconvs): tiles cover the grid,
    K is whole 64-blocks, smem fits."""
    layers = [  # kind,H,W,cin,in_cs,cout,k,stride,bn
        (0, 384, 512, 27, 32, 64, 7, 2, 64), (0, 192, 256, 64, 64, 128, 5, 2, 128), (0, 96, 128, 128, 200, 256, 5, 2, 128),
        (0, 48, 64, 256, 256, 256, 3, 1, 128), (0, 48, 64, 256, 392, 512, 3, 2, 128), (0, 24, 32, 512, 512, 512, 3, 1, 128),
        (0, 24, 32, 512, 776, 512, 3, 2, 128), (0, 12, 16, 512, 512, 512, 3, 1, 128), (0, 12, 16, 512, 1032, 1024, 3, 2, 128),
        (0, 6, 8, 1024, 1024, 1024, 3, 1, 128), (1, 6, 8, 1024, 1024, 512, 4, 2, 128),
        (1, 12, 16, 1026, 1032, 256, 4, 2, 128),
        (1, 24, 32, 770, 776, 128, 4, 2, 128), (1, 48, 64, 386, 392, 64, 4, 2, 64),
        (0, 96, 128, 194, 200, 18, 1, 1, 32)]
    for B in (1, 8):
        for (kind, H, W, cin, in_cs, cout, k, s, bn) in layers:
            p = emu.get_plan(lib, kind, B, H, W, cin, in_cs, cout, k, s, bn)
            tileW = 1 << p["tileW_log2"]
            assert tileW * p["tile_rows"] <= 128 and p["npieces"] * p["piece_rows"] == p["tile_rows"]
            assert p["piece_rows"] == p["box_y"] * p["box_b"]
            assert p["box_b"] > 1 or p["Hg"] % p["box_y"] == 0, "a piece must not straddle two images"
            assert (p["piece_rows"] * tileW) % 8 == 0, "TMA piece must be whole 1024-byte swizzle atoms"
            assert p["tiles_m"] * p["tile_rows"] * tileW >= B * p["Hg"] * p["Wg"]
            assert p["k_total"] == p["ntaps"] * p["nchunks"] * 64 and p["nchunks"] * 64 >= (cin if not p["paired"] else 64)
            assert p["smem"] <= 227 * 1024 and 1 <= p["grid"] <= 148 or p["grid"] >= 1
            assert p["w_rows"] == p["phases"] * p["n_pad"] and p["n_pad"] % bn == 0


def test_plan_rejects_unsupported_shapes(lib):
    info, taps = (C.c_int * 44)(), (C.c_short * 256)()

    def rc(*a):
        return lib.ofs_debug_conv_plan(*a, 1, None, None, C.cast(info, C.c_void_p), C.cast(taps, C.c_void_p), None, 0, None)

    assert rc(0, 1, 7, 9, 64, 64, 16, 3, 2, 16) != 0     # odd H/W with stride 2
    assert rc(0, 1, 8, 12, 64, 64, 16, 3, 1, 16) != 0    # grid width 12 is not a multiple of a power of two >= 8
    assert rc(0, 1, 8, 16, 64, 60, 16, 3, 1, 16) != 0    # channel stride not a multiple of 8
    assert rc(0, 1, 8, 16, 64, 64, 16, 3, 1, 48) != 0    # block_n
    assert rc(1, 1, 8, 16, 64, 64, 16, 3, 2, 16) != 0    # transposed conv must be k4 s2
    assert b"k=4" in lib.ofs_last_error()


def _schedule(lib, kind, B, H, W, cin, in_cs, cout, k, stride, bn, cta_group=1, ksplit=1):
    out = (C.c_int * 10)()
    rc = lib.ofs_debug_conv_schedule(kind, B, H, W, cin, in_cs, cout, k, stride, bn, cta_group, ksplit, C.cast(out, C.c_void_p))
    assert rc == 0, lib.ofs_last_error()
    keys = ("grid", "tail_t0", "tiles_mp", "tiles_n", "phases", "ksplit", "kcluster", "tma_store", "ws_kib", "smem")
    return dict(zip(keys, list(out)))


OFF = 0x7FFFFFFF


def test_tail_split_schedule(lib, monkeypatch):
    """A last wave that fills at most half of the 148 SMs (74 CTA pairs) is issued as half tiles: every (M tile, column
    half) of the tail is visited exactly once by the unit -> tile decode the kernel uses."""
    monkeypatch.delenv("OFS_TAIL_HALF", raising=False)
    conv3_1 = (0, 8, 48, 64, 256, 256, 256, 3, 1, 256)
    s = _schedule(lib, *conv3_1)                                   # 192 tiles of 256 columns on 148 CTAs
    assert (s["tiles_mp"], s["tiles_n"], s["grid"], s["tail_t0"]) == (192, 1, 148, 148)
    units = s["tail_t0"] + 2 * (s["tiles_mp"] - s["tail_t0"])
    seen = set()
    for u in range(units):                                          # conv_gemm_body::decode_tile
        if u >= s["tail_t0"]:
            t = u - s["tail_t0"]
            seen.add((s["tail_t0"] + (t >> 1), t & 1))
        else:
            seen.add((u, -1))
    assert seen == {(m, -1) for m in range(148)} | {(m, h) for m in range(148, 192) for h in (0, 1)}
    assert max(sum(1 for u in range(c, units, s["grid"])) for c in range(s["grid"])) == 2      # one full + one half at most
    pairs = _schedule(lib, 0, 8, 96, 128, 128, 200, 256, 5, 2, 256, cta_group=2)              # conv3 on CTA pairs
    assert (pairs["tiles_mp"], pairs["grid"], pairs["tail_t0"]) == (96, 148, 74)
    assert _schedule(lib, 0, 16, 48, 64, 256, 256, 256, 3, 1, 256)["tail_t0"] == OFF           # 384 tiles: the rest (88) fills > half
    assert _schedule(lib, 0, 1, 48, 64, 256, 256, 256, 3, 1, 256)["tail_t0"] == OFF            # 24 tiles: a single partial wave
    assert _schedule(lib, 0, 8, 48, 64, 256, 392, 512, 3, 2, 256)["tail_t0"] == OFF            # two N tiles
    assert _schedule(lib, 0, 8, 12, 16, 512, 512, 512, 3, 1, 256, ksplit=6)["tail_t0"] == OFF  # split-K
    monkeypatch.setenv("OFS_TAIL_HALF", "0")
    assert _schedule(lib, *conv3_1)["tail_t0"] == OFF


def test_splitk_schedules(lib):
    """Workspace split-K: one persistent wave, fp32 partials in a workspace; cluster split-K: one CTA per (tile, split),
    no workspace, reduced on chip; shapes the cluster form cannot take are rejected."""
    conv6_1 = (0, 8, 6, 8, 1024, 1024, 1024, 3, 1, 256)
    ws = _schedule(lib, *conv6_1, ksplit=8)
    assert (ws["ksplit"], ws["kcluster"], ws["grid"], ws["tma_store"]) == (8, 0, 128, 2)
    assert ws["ws_kib"] == 8 * 8 * 6 * 8 * 1024 * 4 // 1024
    kc = _schedule(lib, *conv6_1, cta_group=16, ksplit=8)
    assert (kc["ksplit"], kc["kcluster"], kc["grid"], kc["tma_store"], kc["ws_kib"]) == (8, 1, 16 * 8, 0, 0)
    uneven = _schedule(lib, 0, 2, 6, 8, 320, 320, 256, 3, 1, 256, cta_group=16, ksplit=7)      # 45 K blocks: 7 + ... + 3
    assert uneven["ksplit"] == 7 and uneven["grid"] % 7 == 0
    out = (C.c_int * 10)()
    bad = lib.ofs_debug_conv_schedule(0, 8, 6, 8, 1024, 1024, 1024, 3, 1, 128, 16, 8, C.cast(out, C.c_void_p))   # 128-column tiles
    assert bad != 0 and b"cluster split-K" in lib.ofs_last_error()
    bad = lib.ofs_debug_conv_schedule(0, 8, 6, 8, 1024, 1024, 1024, 3, 1, 256, 16, 12, C.cast(out, C.c_void_p))  # 12 > 8 CTAs
    assert bad != 0


def test_build_is_a_no_op_when_fresh(monkeypatch):
    """A fresh tree never shells out to nvcc (the GPU box loads the shipped library); stale trees rebuild under a lock."""
    from coupe.optical_flow_based_deep_video_stabilization_b200 import build

    assert build.is_fresh()
    monkeypatch.setattr(build, "_nvcc", lambda: (_ for _ in ()).throw(AssertionError("nvcc invoked on a fresh tree")))
    assert build.build_library() == build.LIB_PATH


def test_byte_over_255_formula_matches_both_reference_quotients():
    """clip.cu / samplers.cu compute byte / 255 as q' = fma(fma(-q, 255, v), c, q), q = v * c, c = RN(1/255).  For all 256
    byte values that equals np.float32(v) / 255.0 (float32 division: the history taps, main_dl.py:556-558) and
    float32(v / 255.0) (float64 division cast at the feed: the current frame, main_dl.py:550, :568).  The two FMAs are
    emulated in float64, which holds their exact products and sums for these magnitudes before the single rounding."""
    v = np.arange(256, dtype=np.float64)
    c = np.float64(np.float32(1.0) / np.float32(255.0))
    assert np.float32(c) == np.float32(0.00392156885936856270)
    q = (v * c).astype(np.float32).astype(np.float64)                 # fmul.rn
    r = v - q * 255.0                                                 # fma(-q, 255, v): exact in float64, exact in float32
    assert np.array_equal(r.astype(np.float32).astype(np.float64), r)
    q2 = (r * c + q).astype(np.float32)                               # fma(r, c, q): one rounding
    hist = np.arange(256, dtype=np.float32) / np.float32(255.0)
    cur = (np.arange(256, dtype=np.float64) / 255.0).astype(np.float32)
    assert np.array_equal(q2, hist) and np.array_equal(q2, cur)
    assert np.count_nonzero(q.astype(np.float32) != cur) > 100        # the plain product is NOT enough


def test_times_384_over_382_formula_is_the_float32_division():
    """predict2_gather_kernel computes (flow2 * 384.0) / 382 (main_dl.py:497: float32 multiply, then true division) as
    t = x * 384, q = t * c, q' = fma(fma(-382, q, t), c, q) with c = RN(1/382).  Emulated with exact rational arithmetic and
    one rounding per operation, q' equals np.float32(t) / np.float32(382) on flow-sized values, tiny values and a sweep
    of exponents; the plain product t * c alone does not."""
    from fractions import Fraction

    c = np.float32(1.0) / np.float32(382.0)
    assert c == np.float32(0.00261780107393860817)
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.normal(0, 5, 6000), rng.random(3000) * 1e-3, rng.normal(0, 300, 3000),
                         2.0 ** rng.integers(-60, 60, 3000) * (rng.random(3000) + 0.5)]).astype(np.float32)

    def rn(fr):                       # exact rational -> float32, round to nearest even
        d = np.float64(float(fr))     # (float64 first: 53 bits hold every value here far beyond the float32 rounding point)
        return np.float32(d)

    plain_differs = 0
    for x in xs:
        t = np.float32(x * np.float32(384.0))
        q = np.float32(t * c)
        r = rn(Fraction(float(t)) - 382 * Fraction(float(q)))
        assert Fraction(float(r)) == Fraction(float(t)) - 382 * Fraction(float(q))      # the remainder is exact
        q2 = rn(Fraction(float(r)) * Fraction(float(c)) + Fraction(float(q)))
        want = t / np.float32(382.0)
        assert q2 == want, (x, q2, want)
        plain_differs += int(q != want)
    assert plain_differs > 100


def test_package_and_oracle_generators_agree():
    """bench.py / benchmarks draw random-init checkpoints and inputs from the package (they may not touch oracle/);
    the oracle keeps its own copy.  Same seeds, same arrays."""
    from coupe.optical_flow_based_deep_video_stabilization_b200 import synthetic as P
    from oracle import flownet as O

    for kind, hs in (("he", None), ("calibrated", 0.02)):
        a, b = P.make_weights(3, kind, head_scale=hs), O.make_weights(3, kind, head_scale=hs)
        assert list(a) == list(b)
        for k in a:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    assert torch.equal(P.make_feats(5, 1), O.make_feats(5, 1))
    assert torch.equal(P.make_feats(5, 1, "smooth"), O.make_feats(5, 1, "smooth"))


def test_warp_fit_compose_inverse_helpers():
    """warp.fit / compose / inverse (reference warp.py:6-23): least-squares affine fit recovers an exact affine map and
    averages noise; compose / inverse are the first-order parameter algebra."""
    import numpy as np

    from coupe.optical_flow_based_deep_video_stabilization_b200 import warp

    rng = np.random.default_rng(0)
    A = np.array([[1.1, -0.2, 3.0], [0.3, 0.9, -1.5]])
    src = rng.uniform(-5, 5, (40, 2))
    dst = src @ A[:, :2].T + A[:, 2]
    M = warp.fit(src, dst)
    assert M.shape == (3, 3) and M.dtype == np.float32
    np.testing.assert_allclose(M[:2], A, atol=1e-5)
    np.testing.assert_allclose(M[2], [0, 0, 1], atol=0)
    noisy = warp.fit(src, dst + rng.normal(0, 0.01, dst.shape))
    np.testing.assert_allclose(noisy[:2], A, atol=2e-2)
    with pytest.raises(ValueError):
        warp.fit(src[:2], dst[:2])
    p, dp = np.arange(8.0), np.ones(8)
    np.testing.assert_array_equal(warp.compose(None, p, dp), p + dp)
    np.testing.assert_array_equal(warp.inverse(None, p), -p)
