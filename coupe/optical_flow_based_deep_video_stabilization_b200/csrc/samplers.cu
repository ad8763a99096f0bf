// Bandwidth-bound gather kernels of the stabiliser hot path (sm_100a):
//   ofs_tf_warp              <- tf_warp / get_pixel_value        main_dl.py:44-130
//   ofs_flow_resize(_warp)   <- test-mode flow glue              main_dl.py:497-514
//   ofs_grid_sample_*        <- Affine/ProjectiveTransformer     spatial_transformer.py:373-452,519-608,755-779,902-964
//   ofs_vec2mtrx/ofs_lie_warp<- vec2mtrx / transformImage / transformCropImage   warp.py:25-129
//
// Design (HBM-bound, 32 B/px for tf_warp at C=3):
//   * a 256-thread block owns a 64x16 output tile; warp w owns rows 2w, 2w+1 and lane l the pixels
//     l and l+32 of a row, so neighbouring lanes always touch neighbouring pixels: flow loads are
//     coalesced, gathers of 12-byte pixels are bank-conflict free in shared memory (stride 3 words)
//     and sector-efficient in L1;
//   * results leave through a per-warp shared-memory transpose as 128-bit stores of whole rows
//     (12-byte pixels are not 16-byte aligned one by one, a row segment is);
//   * the "staged" tf_warp variant loads the bounding box of the source pixels the tile touches into
//     shared memory with fully coalesced loads (bbox found with warp-shuffle min/max reductions) and
//     gathers the 4 corners from there; tiles whose bbox does not fit fall back to read-only-path gathers;
//   * grids are a multiple of the SM count and grid-stride over the tiles.
// Arithmetic follows the reference operation by operation where that decides an index
// (truncation vs floor, clip order, fp32 coordinate products); see oracle/samplers.py.
#include <algorithm>
#include <climits>
#include <type_traits>

#include "ofs_common.cuh"

namespace ofs {
namespace {

struct Taps {
  int y[4];
  int x[4];  // source pixel; x < 0 marks "contributes zero"
  float w[4];
};

__device__ __forceinline__ int clip_i(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---- main_dl.py:83-120 ------------------------------------------------------------------------
__device__ __forceinline__ Taps taps_tfwarp(float fx, float fy, int ox, int oy, int H, int W) {
  Taps t;
  const float x = (float)ox + fx;
  const float y = (float)oy + fy;
  const int xi = __float2int_rz(x);  // tf.cast(float->int32): truncation toward zero
  const int yi = __float2int_rz(y);
  const int x0 = clip_i(xi, 0, W - 1);
  const int y0 = clip_i(yi, 0, H - 1);
  const int x1 = (xi >= W - 1) ? (W - 1) : max(xi + 1, 0);  // clip(xi+1) without int overflow
  const int y1 = (yi >= H - 1) ? (H - 1) : max(yi + 1, 0);
  const float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
  t.y[0] = y0; t.x[0] = x0; t.w[0] = (x1f - x) * (y1f - y);  // Ia
  t.y[1] = y1; t.x[1] = x0; t.w[1] = (x1f - x) * (y - y0f);  // Ib
  t.y[2] = y0; t.x[2] = x1; t.w[2] = (x - x0f) * (y1f - y);  // Ic
  t.y[3] = y1; t.x[3] = x1; t.w[3] = (x - x0f) * (y - y0f);  // Id
  return t;
}

// ---- main_dl.py:497-498 (TF1 legacy bilinear of the pre-scaled flow, then per-axis rescale) ----
struct FlowResize {
  const float* flow2;  // [B,fh,fw,2]
  int fh, fw, H, W;
  float hs, ws;  // fh/H, fw/W in fp32 like CalculateResizeScale
  int prescaled; // flow2 already holds (flow2 * 384.0) / fh (the network writes that copy itself)
  float pre_mul = 384.0f;   // the factor in front of "/ fh": 384.0 for predict_flow2 (main_dl.py:497), out_h for the
                            // predict_flow3-based flow of the homography mode (main_dl.py:681: flow3 * out_h / 48)
  __device__ __forceinline__ float2 at(int b, int oy, int ox) const {
    const float iy = (float)oy * hs;
    const float ix = (float)ox * ws;
    const int y0 = (int)floorf(iy), x0 = (int)floorf(ix);
    const int y1 = min(y0 + 1, fh - 1), x1 = min(x0 + 1, fw - 1);
    const float yl = iy - (float)y0, xl = ix - (float)x0;
    const float2* base = reinterpret_cast<const float2*>(flow2) + (size_t)b * fh * fw;
    float2 tl = __ldg(base + (size_t)y0 * fw + x0), tr = __ldg(base + (size_t)y0 * fw + x1);
    float2 bl = __ldg(base + (size_t)y1 * fw + x0), br = __ldg(base + (size_t)y1 * fw + x1);
    if (!prescaled) {
      const float fhf = (float)fh;
      // outputs['predict_flow2'] * 384.0 / 382, outputs['predict_flow3'] * out_h / 48  (multiply, then true division)
      tl.x = __fdiv_rn(tl.x * pre_mul, fhf); tl.y = __fdiv_rn(tl.y * pre_mul, fhf);
      tr.x = __fdiv_rn(tr.x * pre_mul, fhf); tr.y = __fdiv_rn(tr.y * pre_mul, fhf);
      bl.x = __fdiv_rn(bl.x * pre_mul, fhf); bl.y = __fdiv_rn(bl.y * pre_mul, fhf);
      br.x = __fdiv_rn(br.x * pre_mul, fhf); br.y = __fdiv_rn(br.y * pre_mul, fhf);
    }
    float2 top, bot, v;
    top.x = tl.x + (tr.x - tl.x) * xl; top.y = tl.y + (tr.y - tl.y) * xl;
    bot.x = bl.x + (br.x - bl.x) * xl; bot.y = bl.y + (br.y - bl.y) * xl;
    v.x = top.x + (bot.x - top.x) * yl;
    v.y = top.y + (bot.y - top.y) * yl;
    v.x = (v.x * (float)W) * 0.001953125f;    // outflow[...,0:1]*out_w/512 (power of two: exact)
    v.y = __fdiv_rn(v.y * (float)H, 384.0f);  // outflow[...,1:2]*out_h/384
    return v;
  }
};

// The four taps of a bilinear sample as two source rows x two source columns -- what every coordinate model here
// produces -- for the tiled kernel: r0 / r1 are ELEMENT offsets of the rows (y * srcW), c0 / c1 the columns, all clamped
// into the image so that every tap can be fetched unconditionally; w[] are the weights in the reference's summation
// order (r0c0, r0c1, r1c0, r1c1), already ZERO for a tap that falls outside the image: the reference multiplies such a
// tap by a zero pixel, w * 0 = +-0, and a rounded product of +-0 leaves the running sum unchanged, exactly as 0 * pixel does.
struct RowColTaps {
  int r0, r1, c0, c1;
  float w[4];
};
// yp in [lo, lo + n) as ONE unsigned compare
__device__ __forceinline__ bool in_range(int v, int lo, int n) { return (unsigned)(v - lo) < (unsigned)n; }

// ---- spatial_transformer.py:755-779 + 442-451 / 578-602 + 916-961 -------------------------------
struct GridSampleCoord {
  const float* theta;  // [B,6] or [B,8]
  int projective;
  int H, W, oH, oW;
  float step_x, step_y;  // tf.linspace fp32 step
  struct Ctx { float t[8]; };   // theta of the block's batch element, loaded once per thread
  __device__ __forceinline__ Ctx begin(int b) const {
    Ctx c;
    const int n = projective ? 8 : 6;
    const float* t = theta + (size_t)b * n;
#pragma unroll
    for (int i = 0; i < 8; ++i) c.t[i] = i < n ? __ldg(t + i) : 0.0f;
    return c;
  }
  __device__ __forceinline__ Taps taps(int b, int oy, int ox) const { return taps(begin(b), oy, ox); }
  __device__ __forceinline__ Taps taps(const Ctx& c, int oy, int ox) const {
    const float xt = (oW == 1) ? -1.0f : __fadd_rn(-1.0f, __fmul_rn(step_x, (float)ox));
    const float yt = (oH == 1) ? -1.0f : __fadd_rn(-1.0f, __fmul_rn(step_y, (float)oy));
    float xs = c.t[0] * xt + c.t[1] * yt + c.t[2];
    float ys = c.t[3] * xt + c.t[4] * yt + c.t[5];
    if (projective) {
      float zs = c.t[6] * xt + c.t[7] * yt + 1.0f;
      if (zs == 0.0f) zs = zs + 1e-8f;
      xs = __fdiv_rn(xs, zs);
      ys = __fdiv_rn(ys, zs);
    }
    const float Wf = (float)W, Hf = (float)H;
    float x = __fmul_rn(__fmul_rn(__fadd_rn(xs, 1.0f), 0.5f), Wf - 1.0f);   // / 2.0: an exact scaling
    float y = __fmul_rn(__fmul_rn(__fadd_rn(ys, 1.0f), 0.5f), Hf - 1.0f);
    x = fminf(fmaxf(x, -1.0f), Wf) + 1.0f;  // clip to [-edge, W-1+edge], then += edge
    y = fminf(fmaxf(y, -1.0f), Hf) + 1.0f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x1 = (int)fminf(x1f, Wf + 1.0f), y1 = (int)fminf(y1f, Hf + 1.0f);
    Taps tp;
    // coordinates are on the 1-px zero-padded image; map back, outside -> zero tap
    auto put = [&](int i, int yp, int xp, float w) {
      const bool in = (yp >= 1) && (yp <= H) && (xp >= 1) && (xp <= W);
      tp.y[i] = in ? yp - 1 : 0;
      tp.x[i] = in ? xp - 1 : -1;
      tp.w[i] = w;
    };
    put(0, y0, x0, (x1f - x) * (y1f - y));  // w00 I00
    put(1, y0, x1, (x - x0f) * (y1f - y));  // w01 I01
    put(2, y1, x0, (x1f - x) * (y - y0f));  // w10 I10
    put(3, y1, x1, (x - x0f) * (y - y0f));  // w11 I11
    return tp;
  }
  __host__ __device__ __forceinline__ bool flag() const { return projective != 0; }
  // the same coordinates and weights as taps() (operation for operation), in row / column form; kProj compiles the
  // projective division in or out
  template <bool kProj>
  __device__ __forceinline__ RowColTaps rc(const Ctx& c, int oy, int ox) const {
    const float xt = (oW == 1) ? -1.0f : __fadd_rn(-1.0f, __fmul_rn(step_x, (float)ox));
    const float yt = (oH == 1) ? -1.0f : __fadd_rn(-1.0f, __fmul_rn(step_y, (float)oy));
    float xs = c.t[0] * xt + c.t[1] * yt + c.t[2];
    float ys = c.t[3] * xt + c.t[4] * yt + c.t[5];
    if (kProj) {
      float zs = c.t[6] * xt + c.t[7] * yt + 1.0f;
      if (zs == 0.0f) zs = zs + 1e-8f;
      xs = __fdiv_rn(xs, zs);
      ys = __fdiv_rn(ys, zs);
    }
    const float Wf = (float)W, Hf = (float)H;
    float x = __fmul_rn(__fmul_rn(__fadd_rn(xs, 1.0f), 0.5f), Wf - 1.0f);
    float y = __fmul_rn(__fmul_rn(__fadd_rn(ys, 1.0f), 0.5f), Hf - 1.0f);
    x = fminf(fmaxf(x, -1.0f), Wf) + 1.0f;
    y = fminf(fmaxf(y, -1.0f), Hf) + 1.0f;
    const float x0f = floorf(x), y0f = floorf(y);
    const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const int x1 = (int)fminf(x1f, Wf + 1.0f), y1 = (int)fminf(y1f, Hf + 1.0f);
    // coordinates are on the 1-px zero-padded image: valid rows / columns are 1..H / 1..W
    const bool vy0 = in_range(y0, 1, H), vy1 = in_range(y1, 1, H), vx0 = in_range(x0, 1, W), vx1 = in_range(x1, 1, W);
    RowColTaps t;
    t.r0 = (vy0 ? y0 - 1 : 0) * W; t.r1 = (vy1 ? y1 - 1 : 0) * W;
    t.c0 = vx0 ? x0 - 1 : 0; t.c1 = vx1 ? x1 - 1 : 0;
    const float ax = x1f - x, bx = x - x0f, ay = y1f - y, by = y - y0f;
    t.w[0] = (vy0 && vx0) ? ax * ay : 0.0f;
    t.w[1] = (vy0 && vx1) ? bx * ay : 0.0f;
    t.w[2] = (vy1 && vx0) ? ax * by : 0.0f;
    t.w[3] = (vy1 && vx1) ? bx * by : 0.0f;
    return t;
  }
};

// ---- warp.py:46-86 / 89-129 ----------------------------------------------------------------------
struct LieCoord {
  const float* pMtrx;    // [B,3,3]
  const float* refMtrx;  // [3,3]
  int srcH, srcW, oH, oW;
  struct Ctx { float M[9]; double sx, sy; };   // refMtrx @ pMtrx[b] and the linspace steps, once per thread
  __device__ __forceinline__ Ctx begin(int b) const {
    Ctx c;
    const float* P = pMtrx + (size_t)b * 9;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        c.M[r * 3 + k] = __ldg(refMtrx + r * 3 + 0) * __ldg(P + 0 + k) + __ldg(refMtrx + r * 3 + 1) * __ldg(P + 3 + k) +
                         __ldg(refMtrx + r * 3 + 2) * __ldg(P + 6 + k);
    c.sx = oW > 1 ? 2.0 / (double)(oW - 1) : 0.0;
    c.sy = oH > 1 ? 2.0 / (double)(oH - 1) : 0.0;
    return c;
  }
  __device__ __forceinline__ Taps taps(int b, int oy, int ox) const { return taps(begin(b), oy, ox); }
  __device__ __forceinline__ Taps taps(const Ctx& c, int oy, int ox) const {
    const float* M = c.M;
    // np.linspace(-1,1,n) evaluated in float64, then astype(float32)
    const float X = (oW == 1) ? -1.0f : (ox == oW - 1 ? 1.0f : (float)(-1.0 + (double)ox * c.sx));
    const float Y = (oH == 1) ? -1.0f : (oy == oH - 1 ? 1.0f : (float)(-1.0 + (double)oy * c.sy));
    const float h0 = M[0] * X + M[1] * Y + M[2];
    const float h1 = M[3] * X + M[4] * Y + M[5];
    const float h2 = M[6] * X + M[7] * Y + M[8];
    const float Xw = __fdiv_rn(h0, h2 + 1e-8f);
    const float Yw = __fdiv_rn(h1, h2 + 1e-8f);
    const float xf = floorf(Xw), xc = ceilf(Xw), yf = floorf(Yw), yc = ceilf(Yw);
    const float xr = Xw - xf, yr = Yw - yf;
    // clamp before the int conversion (far-outside stays outside; avoids UB on huge values)
    const int xfi = (int)fminf(fmaxf(xf, -2.0f), (float)srcW + 1.0f);
    const int xci = (int)fminf(fmaxf(xc, -2.0f), (float)srcW + 1.0f);
    const int yfi = (int)fminf(fmaxf(yf, -2.0f), (float)srcH + 1.0f);
    const int yci = (int)fminf(fmaxf(yc, -2.0f), (float)srcH + 1.0f);
    Taps tp;
    auto put = [&](int i, int yi, int xi, float w) {
      const bool in = (xi >= 0) && (xi < srcW) && (yi >= 0) && (yi < srcH);
      tp.y[i] = in ? yi : 0;
      tp.x[i] = in ? xi : -1;
      tp.w[i] = w;
    };
    put(0, yfi, xfi, (1.0f - xr) * (1.0f - yr));  // UL
    put(1, yfi, xci, xr * (1.0f - yr));           // UR
    put(2, yci, xfi, (1.0f - xr) * yr);           // BL
    put(3, yci, xci, xr * yr);                    // BR
    return tp;
  }
  __host__ __device__ __forceinline__ bool flag() const { return false; }
  template <bool>
  __device__ __forceinline__ RowColTaps rc(const Ctx& c, int oy, int ox) const {
    const float* M = c.M;
    const float X = (oW == 1) ? -1.0f : (ox == oW - 1 ? 1.0f : (float)(-1.0 + (double)ox * c.sx));
    const float Y = (oH == 1) ? -1.0f : (oy == oH - 1 ? 1.0f : (float)(-1.0 + (double)oy * c.sy));
    const float h0 = M[0] * X + M[1] * Y + M[2];
    const float h1 = M[3] * X + M[4] * Y + M[5];
    const float h2 = M[6] * X + M[7] * Y + M[8];
    const float Xw = __fdiv_rn(h0, h2 + 1e-8f);
    const float Yw = __fdiv_rn(h1, h2 + 1e-8f);
    const float xf = floorf(Xw), xc = ceilf(Xw), yf = floorf(Yw), yc = ceilf(Yw);
    const float xr = Xw - xf, yr = Yw - yf;
    const int xfi = (int)fminf(fmaxf(xf, -2.0f), (float)srcW + 1.0f);
    const int xci = (int)fminf(fmaxf(xc, -2.0f), (float)srcW + 1.0f);
    const int yfi = (int)fminf(fmaxf(yf, -2.0f), (float)srcH + 1.0f);
    const int yci = (int)fminf(fmaxf(yc, -2.0f), (float)srcH + 1.0f);
    const bool vy0 = in_range(yfi, 0, srcH), vy1 = in_range(yci, 0, srcH), vx0 = in_range(xfi, 0, srcW), vx1 = in_range(xci, 0, srcW);
    RowColTaps t;
    t.r0 = (vy0 ? yfi : 0) * srcW; t.r1 = (vy1 ? yci : 0) * srcW;
    t.c0 = vx0 ? xfi : 0; t.c1 = vx1 ? xci : 0;
    const float ax = 1.0f - xr, ay = 1.0f - yr;
    t.w[0] = (vy0 && vx0) ? ax * ay : 0.0f;   // UL
    t.w[1] = (vy0 && vx1) ? xr * ay : 0.0f;   // UR
    t.w[2] = (vy1 && vx0) ? ax * yr : 0.0f;   // BL
    t.w[3] = (vy1 && vx1) ? xr * yr : 0.0f;   // BR
    return t;
  }
};

// -------------------------------------------------------------------------------------------------
// tap providers for the generic kernels
struct TfWarpProvider {
  const float* flow;  // [B,H,W,2]
  int H, W;
  __device__ __forceinline__ Taps taps(int b, int oy, int ox) const {
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)b * H + oy) * W + ox);
    return taps_tfwarp(f.x, f.y, ox, oy, H, W);
  }
  __device__ __forceinline__ float2 flow_at(int b, int oy, int ox) const {
    return __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)b * H + oy) * W + ox);
  }
};
struct ResizeWarpProvider {
  FlowResize fr;
  __device__ __forceinline__ Taps taps(int b, int oy, int ox) const {
    const float2 f = fr.at(b, oy, ox);
    return taps_tfwarp(f.x, f.y, ox, oy, fr.H, fr.W);
  }
  __device__ __forceinline__ float2 flow_at(int b, int oy, int ox) const { return fr.at(b, oy, ox); }
};
template <class Coord>
struct CoordProvider {
  Coord c;
  typedef typename Coord::Ctx Ctx;
  __device__ __forceinline__ Ctx begin(int b) const { return c.begin(b); }
  __device__ __forceinline__ Taps taps(const Ctx& ctx, int oy, int ox) const { return c.taps(ctx, oy, ox); }
  __device__ __forceinline__ Taps taps(int b, int oy, int ox) const { return c.taps(b, oy, ox); }
  template <bool kFlag>
  __device__ __forceinline__ RowColTaps rc(const Ctx& ctx, int oy, int ox) const { return c.template rc<kFlag>(ctx, oy, ox); }
  bool flag() const { return c.flag(); }
};

// acc + w*v as two separately rounded fp32 operations (no FMA contraction): the reference sums
// already-rounded products (tf.add_n([wa*Ia, ...])), which makes clipped corner pairs such as
// (-0.5*I) + (0.5*I) cancel to exactly zero.
__device__ __forceinline__ float mul_add_rn(float acc, float w, float v) { return __fadd_rn(acc, __fmul_rn(w, v)); }

// 12-byte pixel gather through the read-only path
__device__ __forceinline__ void gather3(const float* __restrict__ imgb, int srcW, int y, int x, float w, float* acc) {
  if (x >= 0) {
    const float* p = imgb + ((size_t)y * srcW + x) * 3;
    acc[0] = mul_add_rn(acc[0], w, __ldg(p + 0));
    acc[1] = mul_add_rn(acc[1], w, __ldg(p + 1));
    acc[2] = mul_add_rn(acc[2], w, __ldg(p + 2));
  }
}

// generic: any C, any width; one thread = one output pixel
template <class Provider>
__global__ void __launch_bounds__(256) sample_px_kernel(Provider prov, const float* __restrict__ img,
                                                        float* __restrict__ out, int B, int srcH, int srcW, int oH,
                                                        int oW, int C) {
  const size_t total = (size_t)B * oH * oW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % oW);
    const size_t r = i / oW;
    const int oy = (int)(r % oH);
    const int b = (int)(r / oH);
    const Taps t = prov.taps(b, oy, ox);
    const float* imgb = img + (size_t)b * srcH * srcW * C;
    float* o = out + i * C;
    for (int c = 0; c < C; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t.x[k] >= 0) acc = mul_add_rn(acc, t.w[k], __ldg(imgb + ((size_t)t.y[k] * srcW + t.x[k]) * C + c));
      o[c] = acc;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// C == 3 tile kernels: 64x16 output tile per 256-thread block (see the design note at the top)
constexpr int kTileW = 64, kTileH = 16;
constexpr int kStageMaxPx = 2560;  // 30 KB of fp32 RGB source pixels per block

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One output row segment (<= 64 px, a multiple of 4) of a warp: lane holds pixels lane and lane+32 as
// 2 x 3 floats; transposed through `obuf` (192 floats, this warp's) into whole-row 128-bit stores.
__device__ __forceinline__ void store_row3(float* obuf, float* __restrict__ grow, int lane, int valid_px,
                                           const float* v0, const float* v1) {
  obuf[lane * 3 + 0] = v0[0]; obuf[lane * 3 + 1] = v0[1]; obuf[lane * 3 + 2] = v0[2];
  obuf[(lane + 32) * 3 + 0] = v1[0]; obuf[(lane + 32) * 3 + 1] = v1[1]; obuf[(lane + 32) * 3 + 2] = v1[2];
  __syncwarp();
  const int nvec = (valid_px * 3) >> 2;
  const float4* o4 = reinterpret_cast<const float4*>(obuf);
  float4* g4 = reinterpret_cast<float4*>(grow);
  if (lane < nvec) __stcs(g4 + lane, o4[lane]);
  if (lane + 32 < nvec) __stcs(g4 + lane + 32, o4[lane + 32]);
  __syncwarp();
}

struct TileId { int b, ox0, oy0; };
__device__ __forceinline__ TileId decode_tile(size_t tile_id, int tiles_x, int tiles_y) {
  TileId t;
  t.ox0 = (int)(tile_id % tiles_x) * kTileW;
  const size_t r = tile_id / tiles_x;
  t.oy0 = (int)(r % tiles_y) * kTileH;
  t.b = (int)(r / tiles_y);
  return t;
}

// generic 4-tap sampler (grid_sample / Lie warp): direct read-only-path gathers
template <class Provider>
__global__ void __launch_bounds__(256) sample3_kernel(Provider prov, const float* __restrict__ img,
                                                      float* __restrict__ out, int B, int srcH, int srcW, int oH,
                                                      int oW) {
  __shared__ __align__(16) float obuf[8][192];
  const int tiles_x = (oW + kTileW - 1) / kTileW, tiles_y = (oH + kTileH - 1) / kTileH;
  const size_t ntiles = (size_t)B * tiles_y * tiles_x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    const TileId t = decode_tile(tile_id, tiles_x, tiles_y);
    const float* imgb = img + (size_t)t.b * srcH * srcW * 3;
    const int valid_px = min(kTileW, oW - t.ox0);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int oy = t.oy0 + wid * 2 + rr;
      if (oy >= oH) break;  // warp-uniform
      float v[2][3];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ox = t.ox0 + lane + 32 * h;
        v[h][0] = v[h][1] = v[h][2] = 0.0f;
        if (ox < oW) {
          const Taps tp = prov.taps(t.b, oy, ox);
#pragma unroll
          for (int k = 0; k < 4; ++k) gather3(imgb, srcW, tp.y[k], tp.x[k], tp.w[k], v[h]);
        }
      }
      store_row3(obuf[wid], out + (((size_t)t.b * oH + oy) * oW + t.ox0) * 3, lane, valid_px, v[0], v[1]);
    }
  }
}

// tf_warp / fused flow-resize + tf_warp; kStaged: source bounding box of the tile staged in shared memory
template <class Provider, bool kStaged>
__global__ void __launch_bounds__(256) warp3_kernel(Provider prov, const float* __restrict__ img,
                                                    float* __restrict__ out, int B, int H, int W) {
  __shared__ __align__(16) float obuf[8][192];
  __shared__ float tile[kStaged ? kStageMaxPx * 3 : 1];
  __shared__ int red[4][8];
  __shared__ int bbox[4];
  pdl_wait();
  pdl_launch_dependents();
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const size_t ntiles = (size_t)B * tiles_y * tiles_x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    const TileId t = decode_tile(tile_id, tiles_x, tiles_y);
    const float* imgb = img + (size_t)t.b * H * W * 3;
    const int valid_px = min(kTileW, W - t.ox0);
    // this thread's 4 pixels: rows 2w, 2w+1 x columns lane, lane+32; keep only their flow
    float2 fl[2][2];
    bool on[2][2];
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int oy = t.oy0 + wid * 2 + rr, ox = t.ox0 + lane + 32 * h;
        on[rr][h] = (oy < H) && (ox < W);
        fl[rr][h] = make_float2(0.f, 0.f);
        if (on[rr][h]) {
          fl[rr][h] = prov.flow_at(t.b, oy, ox);
          if (kStaged) {
            const Taps tp = taps_tfwarp(fl[rr][h].x, fl[rr][h].y, ox, oy, H, W);
            xmin = min(xmin, tp.x[0]); xmax = max(xmax, tp.x[3]);  // x0 <= x1, y0 <= y1 after clipping
            ymin = min(ymin, tp.y[0]); ymax = max(ymax, tp.y[3]);
          }
        }
      }
    bool staged = false;
    int bx0 = 0, by0 = 0, bw = 0;
    if (kStaged) {
      xmin = warp_min(xmin); ymin = warp_min(ymin); xmax = warp_max(xmax); ymax = warp_max(ymax);
      if (lane == 0) { red[0][wid] = xmin; red[1][wid] = xmax; red[2][wid] = ymin; red[3][wid] = ymax; }
      __syncthreads();
      if (threadIdx.x < 32) {
        int a = (lane < 8) ? red[0][lane] : INT_MAX, c = (lane < 8) ? red[1][lane] : INT_MIN;
        int d = (lane < 8) ? red[2][lane] : INT_MAX, e = (lane < 8) ? red[3][lane] : INT_MIN;
        a = warp_min(a); c = warp_max(c); d = warp_min(d); e = warp_max(e);
        if (lane == 0) { bbox[0] = a; bbox[1] = c; bbox[2] = d; bbox[3] = e; }
      }
      __syncthreads();
      bx0 = bbox[0]; by0 = bbox[2];
      bw = bbox[1] - bx0 + 1;
      const int bh = bbox[3] - by0 + 1;
      staged = (bw > 0) && (bh > 0) && ((long long)bw * bh <= kStageMaxPx);
      if (staged) {
        const int row_f = bw * 3;
        for (int r = wid; r < bh; r += 8) {  // one warp per source row: coalesced 128-byte lines
          const float* src = imgb + ((size_t)(by0 + r) * W + bx0) * 3;
          float* dst = tile + r * row_f;
          for (int i = lane; i < row_f; i += 32) dst[i] = __ldg(src + i);
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int oy = t.oy0 + wid * 2 + rr;
      if (oy < H) {  // warp-uniform
        float v[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          v[h][0] = v[h][1] = v[h][2] = 0.0f;
          if (on[rr][h]) {
            const Taps tp = taps_tfwarp(fl[rr][h].x, fl[rr][h].y, t.ox0 + lane + 32 * h, oy, H, W);
            if (kStaged && staged) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float* p = tile + ((tp.y[k] - by0) * bw + (tp.x[k] - bx0)) * 3;
                v[h][0] = mul_add_rn(v[h][0], tp.w[k], p[0]);
                v[h][1] = mul_add_rn(v[h][1], tp.w[k], p[1]);
                v[h][2] = mul_add_rn(v[h][2], tp.w[k], p[2]);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) gather3(imgb, W, tp.y[k], tp.x[k], tp.w[k], v[h]);
            }
          }
        }
        store_row3(obuf[wid], out + (((size_t)t.b * H + oy) * W + t.ox0) * 3, lane, valid_px, v[0], v[1]);
      }
    }
    if (kStaged) __syncthreads();  // tile / bbox reuse
  }
}

// tf_warp / fused flow-resize + tf_warp, variant 2: the tile's source bounding box is staged in shared memory
// as PADDED 16-byte pixels (x aligned to 4 px so the global side is whole 128-bit loads), every bilinear corner
// is then ONE 128-bit shared load instead of three 32-bit ones, and the taps are computed once (the sample
// position is kept in registers between the bounding-box pass and the gather pass).
constexpr int kStage4MaxPx = 2304;  // 36 KB of padded fp32 RGB source pixels per block
template <class Provider>
__global__ void __launch_bounds__(256) warp4_kernel(Provider prov, const float* __restrict__ img,
                                                    float* __restrict__ out, int B, int H, int W) {
  __shared__ __align__(16) float obuf[8][192];
  __shared__ __align__(16) float4 tile[kStage4MaxPx];
  __shared__ int red[4][8];
  pdl_wait();
  pdl_launch_dependents();
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const size_t ntiles = (size_t)B * tiles_y * tiles_x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    const TileId t = decode_tile(tile_id, tiles_x, tiles_y);
    const float* imgb = img + (size_t)t.b * H * W * 3;
    const int valid_px = min(kTileW, W - t.ox0);
    // this thread's 4 pixels: rows 2w, 2w+1 x columns lane, lane+32: sample position (main_dl.py:83)
    float sx[2][2], sy[2][2];
    bool on[2][2];
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int oy = t.oy0 + wid * 2 + rr, ox = t.ox0 + lane + 32 * h;
        on[rr][h] = (oy < H) && (ox < W);
        sx[rr][h] = sy[rr][h] = 0.f;
        if (on[rr][h]) {
          const float2 f = prov.flow_at(t.b, oy, ox);
          const float x = (float)ox + f.x, y = (float)oy + f.y;
          sx[rr][h] = x; sy[rr][h] = y;
          const int xi = __float2int_rz(x), yi = __float2int_rz(y);
          xmin = min(xmin, clip_i(xi, 0, W - 1)); ymin = min(ymin, clip_i(yi, 0, H - 1));
          xmax = max(xmax, (xi >= W - 1) ? (W - 1) : max(xi + 1, 0));
          ymax = max(ymax, (yi >= H - 1) ? (H - 1) : max(yi + 1, 0));
        }
      }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0) { red[0][wid] = xmin; red[1][wid] = xmax; red[2][wid] = ymin; red[3][wid] = ymax; }
    __syncthreads();
    int bx0 = red[0][0], bx1 = red[1][0], by0 = red[2][0], by1 = red[3][0];
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      bx0 = min(bx0, red[0][i]); bx1 = max(bx1, red[1][i]); by0 = min(by0, red[2][i]); by1 = max(by1, red[3][i]);
    }
    bx0 &= ~3;                                    // 16-byte aligned row segments (W % 4 == 0)
    const int bw = ((bx1 + 4) & ~3) - bx0;        // px, multiple of 4, <= W - bx0
    const int bh = by1 - by0 + 1;
    const bool staged = (bw > 0) && (bh > 0) && (bw * bh <= kStage4MaxPx);
    if (staged) {
      const int nvec = (bw * 3) >> 2;             // float4 per source row segment
      float* tf = reinterpret_cast<float*>(tile);
      for (int r = wid; r < bh; r += 8) {
        const float4* src = reinterpret_cast<const float4*>(imgb + ((size_t)(by0 + r) * W + bx0) * 3);
        float* dst = tf + (size_t)r * bw * 4;
        for (int k = lane; k < nvec; k += 32) {
          const float4 v = __ldg(src + k);
          const int i = 4 * k;                    // float index in the row segment: pixel i / 3, channel i % 3
          const int p0 = i / 3, c0 = i - 3 * p0;
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int cc = c0 + q;
            const int pp = p0 + (cc >= 3 ? 1 : 0) + (cc >= 6 ? 1 : 0);   // c0 <= 2, q <= 3: at most two carries
            const int ch = cc - 3 * (pp - p0);
            dst[pp * 4 + ch] = e[q];
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int oy = t.oy0 + wid * 2 + rr;
      if (oy < H) {  // warp-uniform
        float v[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          v[h][0] = v[h][1] = v[h][2] = 0.0f;
          if (on[rr][h]) {
            // main_dl.py:88-120 on the stored sample position
            const float x = sx[rr][h], y = sy[rr][h];
            const int xi = __float2int_rz(x), yi = __float2int_rz(y);
            const int x0 = clip_i(xi, 0, W - 1), y0 = clip_i(yi, 0, H - 1);
            const int x1 = (xi >= W - 1) ? (W - 1) : max(xi + 1, 0);
            const int y1 = (yi >= H - 1) ? (H - 1) : max(yi + 1, 0);
            const float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
            const float wa = (x1f - x) * (y1f - y), wb = (x1f - x) * (y - y0f);
            const float wc = (x - x0f) * (y1f - y), wd = (x - x0f) * (y - y0f);
            if (staged) {
              const float4* r0 = tile + (y0 - by0) * bw - bx0, *r1 = tile + (y1 - by0) * bw - bx0;
              const float4 Ia = r0[x0], Ib = r1[x0], Ic = r0[x1], Id = r1[x1];
              v[h][0] = mul_add_rn(mul_add_rn(mul_add_rn(mul_add_rn(0.f, wa, Ia.x), wb, Ib.x), wc, Ic.x), wd, Id.x);
              v[h][1] = mul_add_rn(mul_add_rn(mul_add_rn(mul_add_rn(0.f, wa, Ia.y), wb, Ib.y), wc, Ic.y), wd, Id.y);
              v[h][2] = mul_add_rn(mul_add_rn(mul_add_rn(mul_add_rn(0.f, wa, Ia.z), wb, Ib.z), wc, Ic.z), wd, Id.z);
            } else {
              gather3(imgb, W, y0, x0, wa, v[h]);
              gather3(imgb, W, y1, x0, wb, v[h]);
              gather3(imgb, W, y0, x1, wc, v[h]);
              gather3(imgb, W, y1, x1, wd, v[h]);
            }
          }
        }
        store_row3(obuf[wid], out + (((size_t)t.b * H + oy) * W + t.ox0) * 3, lane, valid_px, v[0], v[1]);
      }
    }
    __syncthreads();  // tile / red reuse
  }
}

// tf_warp / fused flow-resize + tf_warp, variant 3 (default): an instruction-lean rewrite of the direct-gather
// kernel.  ncu on variants 0-2 (profiles/r01_tuning.md): 174 (tf_warp) and 326 (fused) warp-instructions
// per 32 pixels-per-lane... i.e. per pixel, issue-bound at ~60 % issue-active with DRAM at 25-40 %.  Here:
//   * all image / flow indexing is 32-bit (one 64-bit base per image), rows and columns of a thread's 2x2
//     pixels share their resize set-up (y terms per row, x terms per column);
//   * the test-mode rescale flow_y * H / 384 is an exactly rounded division by a constant: q = t * RN(1/3),
//     r = fma(-3, q, t), q' = fma(r, RN(1/3), q) (Markstein), then an exact scale by 2^-7;
//   * 6 resident blocks per SM (launch bounds) instead of 4-5.
struct Resize2 {             // fused provider: pre-scaled flow2 [B,fh,fw,2] -> flow at the output pixel
  const float2* f2;
  int fh, fw;
  float hs, ws, Wf, Hf;
};
// Tap addresses of the tiled warps: a non-negative 32-bit element index times the element size, added to a per-image base
// that lives in ONE 64-bit register pair (pin_base) -- a single IMAD.WIDE.U32 per address.  Left to itself the compiler
// keeps the kernel-parameter part of the base apart and spends an LEA / LEA.HI.X (or IMAD.WIDE + IADD3 + IADD3.X) per tap;
// these kernels are issue-bound (ncu: 77 % issue-active), 8 tap addresses per pixel.
template <class T>
__device__ __forceinline__ const T* pin_base(const T* p) {
  asm("" : "+l"(p));
  return p;
}
template <int kBytes, class T>
__device__ __forceinline__ const T* tap_ptr(const T* base, int idx) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (size_t)(unsigned)idx * (unsigned)kBytes);
}

__device__ __forceinline__ float div384(float t) {
  const float c = 0.333333343267440796f;   // RN(1/3)
  float q = t * c;
  const float r = __fmaf_rn(-3.0f, q, t);
  q = __fmaf_rn(r, c, q);
  return q * 0.0078125f;
}

// flow at a thread's 2 x 2 pixels (rows oy0, oy0+1; columns ox0+lane, ox0+lane+32) of the fused step --
// main_dl.py:497-498: TF1 legacy bilinear of the pre-scaled flow, then x * W / 512, y * H / 384
__device__ __forceinline__ void fused_flow_2x2(const Resize2& rz, int b, int ox0, int oy0, int lane, float2 (&f)[2][2]) {
  const float2* __restrict__ fb = pin_base(rz.f2 + (size_t)b * rz.fh * rz.fw);
  int xa[2], xb[2], ya[2], yb[2];
  float xl[2], yl[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float ix = (float)(ox0 + lane + 32 * h) * rz.ws;
    const float fl = floorf(ix);
    xa[h] = min((int)fl, rz.fw - 1);          // columns past W are never stored; keep the loads in bounds
    xb[h] = min(xa[h] + 1, rz.fw - 1);
    xl[h] = ix - fl;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float iy = (float)(oy0 + r) * rz.hs;
    const float fl = floorf(iy);
    const int y0 = min((int)fl, rz.fh - 1);
    ya[r] = y0 * rz.fw;
    yb[r] = min(y0 + 1, rz.fh - 1) * rz.fw;
    yl[r] = iy - fl;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 tl = __ldg(tap_ptr<8>(fb, ya[r] + xa[h])), tr = __ldg(tap_ptr<8>(fb, ya[r] + xb[h]));
      const float2 bl = __ldg(tap_ptr<8>(fb, yb[r] + xa[h])), br = __ldg(tap_ptr<8>(fb, yb[r] + xb[h]));
      const float topx = tl.x + (tr.x - tl.x) * xl[h], topy = tl.y + (tr.y - tl.y) * xl[h];
      const float botx = bl.x + (br.x - bl.x) * xl[h], boty = bl.y + (br.y - bl.y) * xl[h];
      const float vx = topx + (botx - topx) * yl[r], vy = topy + (boty - topy) * yl[r];
      f[r][h].x = (vx * rz.Wf) * 0.001953125f;
      f[r][h].y = div384(vy * rz.Hf);
    }
}

// kRows output rows per warp (2, or 4: two row pairs, the per-thread column set-up shared), kMinBlocks resident blocks per SM
template <bool kFused, bool kDirectStore = false, int kRows = 2, int kMinBlocks = 6>
__global__ void __launch_bounds__(256, kMinBlocks) warp5_kernel(const float* __restrict__ img, const float2* __restrict__ flow,
                                                                 Resize2 rz, float* __restrict__ out, int B, int H, int W) {
  __shared__ __align__(16) float obuf[kDirectStore ? 1 : 8][192];
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Wm1 = W - 1, Hm1 = H - 1;
#pragma unroll
  for (int rp = 0; rp < kRows / 2; ++rp) {   // one block per 64 x (8 kRows) output tile: grid (tiles_x, tiles_y, B), no tile decode arithmetic
    const int b = blockIdx.z;
    const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * (8 * kRows) + wid * kRows + 2 * rp;
    const float* __restrict__ imgb = pin_base(img + (size_t)b * H * W * 3);
    const int valid_px = min(kTileW, W - ox0);
    // ---- flow at this thread's 2 x 2 pixels (rows oy0, oy0+1; columns ox0+lane, ox0+lane+32)
    float2 f[2][2];
    if (kFused) {
      fused_flow_2x2(rz, b, ox0, oy0, lane, f);
    } else {
      const float2* __restrict__ fb = flow + (size_t)b * H * W;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int oy = min(oy0 + r, Hm1), ox = min(ox0 + lane + 32 * h, Wm1);
          f[r][h] = __ldg(fb + oy * W + ox);
        }
    }
    // ---- main_dl.py:83-129 per pixel
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int oy = oy0 + r;
      if (oy < H) {   // warp-uniform
        float v[2][3];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ox = ox0 + lane + 32 * h;
          const float x = (float)ox + f[r][h].x, y = (float)oy + f[r][h].y;
          const int xi = __float2int_rz(x), yi = __float2int_rz(y);
          const int x0 = clip_i(xi, 0, Wm1), y0 = clip_i(yi, 0, Hm1);
          const int x1 = (xi >= Wm1) ? Wm1 : max(xi + 1, 0);
          const int y1 = (yi >= Hm1) ? Hm1 : max(yi + 1, 0);
          const float dx1 = (float)x1 - x, dx0 = x - (float)x0, dy1 = (float)y1 - y, dy0 = y - (float)y0;
          const float wa = dx1 * dy1, wb = dx1 * dy0, wc = dx0 * dy1, wd = dx0 * dy0;
          const int r0 = y0 * W, r1 = y1 * W;
          const float* pa = tap_ptr<12>(imgb, r0 + x0);
          const float* pb = tap_ptr<12>(imgb, r1 + x0);
          const float* pc = tap_ptr<12>(imgb, r0 + x1);
          const float* pd = tap_ptr<12>(imgb, r1 + x1);
          if (ox < W) {
            const float a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
            const float b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
            const float c0 = __ldg(pc), c1 = __ldg(pc + 1), c2 = __ldg(pc + 2);
            const float d0 = __ldg(pd), d1 = __ldg(pd + 1), d2 = __ldg(pd + 2);
            // tf.add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id]): rounded products, summed left to right
            v[h][0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, a0), __fmul_rn(wb, b0)), __fmul_rn(wc, c0)), __fmul_rn(wd, d0));
            v[h][1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, a1), __fmul_rn(wb, b1)), __fmul_rn(wc, c1)), __fmul_rn(wd, d1));
            v[h][2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, a2), __fmul_rn(wb, b2)), __fmul_rn(wc, c2)), __fmul_rn(wd, d2));
          } else {
            v[h][0] = v[h][1] = v[h][2] = 0.0f;
          }
        }
        if (kDirectStore) {
          float* o = out + (((size_t)b * H + oy) * W + ox0 + lane) * 3;
          if (ox0 + lane < W) { __stcs(o, v[0][0]); __stcs(o + 1, v[0][1]); __stcs(o + 2, v[0][2]); }
          if (ox0 + lane + 32 < W) { __stcs(o + 96, v[1][0]); __stcs(o + 97, v[1][1]); __stcs(o + 98, v[1][2]); }
        } else {
          store_row3(obuf[kDirectStore ? 0 : wid], out + (((size_t)b * H + oy) * W + ox0) * 3, lane, valid_px, v[0], v[1]);
        }
      }
    }
  }
}

// The clip driver's warp (main_dl.py:568-569, :625, :630) on the uint8 frame itself:
//   resizedInput = cvtColor(frame, RGB2BGR) / 255.0 (float64 quotient, float32 at the feed), tf_warp, then
//   totaloutputFrame = cvtColor(warped * 255, RGB2BGR) in float32 and np.uint8() of it.
// The two channel swaps cancel, so output byte k is computed from source byte k.  byte / 255 in the reference's
// rounding is q = v * RN(1/255), r = fma(-q, 255, v), q' = fma(r, RN(1/255), q): equal to
// (float)((double)v / 255.0) for all 256 byte values (tests/test_gpu_clip.py checks every one of them).
// 3 bytes in and 3 bytes out per pixel instead of three kernels moving 12 + 24 + 15.
__device__ __forceinline__ float byte_over_255(uint32_t v) {
  const float c = 0.00392156885936856270f;   // RN(1/255)
  const float x = (float)v;
  const float q = x * c;
  const float r = __fmaf_rn(-q, 255.0f, x);
  return __fmaf_rn(r, c, q);
}

template <bool kWantF32>
__global__ void __launch_bounds__(256, 6) warp5_u8_kernel(const uint8_t* __restrict__ img, Resize2 rz,
                                                           uint8_t* __restrict__ out_u8, float* __restrict__ out_f32,
                                                           int B, int H, int W) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int Wm1 = W - 1, Hm1 = H - 1;
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH + wid * 2;
  const uint8_t* __restrict__ imgb = pin_base(img + (size_t)b * H * W * 3);
  float2 f[2][2];
  fused_flow_2x2(rz, b, ox0, oy0, lane, f);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int oy = oy0 + r;
    if (oy < H) {   // warp-uniform
      uint32_t packed[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ox = ox0 + lane + 32 * h;
        const float x = (float)ox + f[r][h].x, y = (float)oy + f[r][h].y;
        const int xi = __float2int_rz(x), yi = __float2int_rz(y);
        const int x0 = clip_i(xi, 0, Wm1), y0 = clip_i(yi, 0, Hm1);
        const int x1 = (xi >= Wm1) ? Wm1 : max(xi + 1, 0);
        const int y1 = (yi >= Hm1) ? Hm1 : max(yi + 1, 0);
        const float dx1 = (float)x1 - x, dx0 = x - (float)x0, dy1 = (float)y1 - y, dy0 = y - (float)y0;
        const float wa = dx1 * dy1, wb = dx1 * dy0, wc = dx0 * dy1, wd = dx0 * dy0;
        const int r0 = y0 * W, r1 = y1 * W;
        const uint8_t* pa = tap_ptr<3>(imgb, r0 + x0);
        const uint8_t* pb = tap_ptr<3>(imgb, r1 + x0);
        const uint8_t* pc = tap_ptr<3>(imgb, r0 + x1);
        const uint8_t* pd = tap_ptr<3>(imgb, r1 + x1);
        packed[h] = 0;
        if (ox < W) {
          float v[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float a = byte_over_255(__ldg(pa + k)), bb = byte_over_255(__ldg(pb + k));
            const float cc = byte_over_255(__ldg(pc + k)), d = byte_over_255(__ldg(pd + k));
            // tf.add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id]): rounded products, summed left to right; then * 255 (float32)
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(wa, a), __fmul_rn(wb, bb)), __fmul_rn(wc, cc)), __fmul_rn(wd, d));
            v[k] = __fmul_rn(s, 255.0f);
            packed[h] |= (uint32_t)(__float2int_rz(v[k]) & 0xff) << (8 * k);   // np.uint8(): truncate, keep the low 8 bits
          }
          if (kWantF32) {
            float* o = out_f32 + (((size_t)b * H + oy) * W + ox) * 3;
            __stcs(o, v[0]); __stcs(o + 1, v[1]); __stcs(o + 2, v[2]);
          }
        }
      }
      // 32 lanes x 3 bytes -> 24 aligned 32-bit words per half row: word w takes its bytes from lanes 4w/3 and 4w/3 + 1
      const int src = (4 * lane) / 3, sh = 8 * (lane % 3);
      uint8_t* orow = out_u8 + (((size_t)b * H + oy) * W + ox0) * 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t lo = __shfl_sync(0xffffffffu, packed[h], src & 31);
        const uint32_t hi = __shfl_sync(0xffffffffu, packed[h], (src + 1) & 31);
        const uint32_t word = (lo >> sh) | (hi << (24 - sh));
        // W % 4 == 0: the valid part of a half row is a whole number of words
        if (lane < 24 && ox0 + 32 * h + (4 * lane) / 3 < W) __stcs(reinterpret_cast<uint32_t*>(orow + 96 * h) + lane, word);
      }
    }
  }
}

// any width (W % 4 != 0): one thread per pixel, the taps and summation order of sample_px_kernel<ResizeWarpProvider>
__global__ void __launch_bounds__(256) warp_u8_px_kernel(ResizeWarpProvider prov, const uint8_t* __restrict__ img,
                                                         uint8_t* __restrict__ out_u8, float* __restrict__ out_f32, int B,
                                                         int H, int W) {
  const size_t total = (size_t)B * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    const size_t r = i / W;
    const int oy = (int)(r % H);
    const int b = (int)(r / H);
    const Taps t = prov.taps(b, oy, ox);
    const uint8_t* imgb = img + (size_t)b * H * W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float acc = 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t.x[k] >= 0) acc = mul_add_rn(acc, t.w[k], byte_over_255(__ldg(imgb + ((size_t)t.y[k] * W + t.x[k]) * 3 + c)));
      const float v = __fmul_rn(acc, 255.0f);
      out_u8[i * 3 + c] = (uint8_t)(__float2int_rz(v) & 0xff);
      if (out_f32) out_f32[i * 3 + c] = v;
    }
  }
}

// 4-tap sampler of the coordinate models (grid_sample / Lie warp): one block per 64 x 16 output tile (3-D grid), a thread
// owns columns lane, lane + 32 of two rows.  The taps arrive as two rows x two columns with the weights of outside taps
// already zero (RowColTaps), so the 12 loads of a pixel are unconditional, every address is one IMAD.WIDE.U32 off a
// pinned base, and the sum is the reference's: rounded products added in tap order.  kFlag: the model's compile-time
// switch (GridSampleCoord: projective division).  [round 1's form built per-tap validity, selects on all 12 loaded values
// and both the affine and the projective path into one kernel: 250 instructions per pixel, 0.45-0.48 of HBM peak]
template <class Provider, bool kFlag>
__global__ void __launch_bounds__(256, 6) sample5_kernel(Provider prov, const float* __restrict__ img,
                                                         float* __restrict__ out, int B, int srcH, int srcW, int oH,
                                                         int oW) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH + wid * 2;
  const float* __restrict__ imgb = pin_base(img + (size_t)b * srcH * srcW * 3);
  const typename Provider::Ctx ctx = prov.begin(b);
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int oy = oy0 + rr;
    if (oy >= oH) break;  // warp-uniform
    float v[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ox = ox0 + lane + 32 * h;
      v[h][0] = v[h][1] = v[h][2] = 0.0f;
      if (ox < oW) {
        const RowColTaps t = prov.template rc<kFlag>(ctx, oy, ox);
        const float* pa = tap_ptr<12>(imgb, t.r0 + t.c0);
        const float* pb = tap_ptr<12>(imgb, t.r0 + t.c1);
        const float* pc = tap_ptr<12>(imgb, t.r1 + t.c0);
        const float* pd = tap_ptr<12>(imgb, t.r1 + t.c1);
        const float a0 = __ldg(pa), a1 = __ldg(pa + 1), a2 = __ldg(pa + 2);
        const float b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
        const float c0 = __ldg(pc), c1 = __ldg(pc + 1), c2 = __ldg(pc + 2);
        const float d0 = __ldg(pd), d1 = __ldg(pd + 1), d2 = __ldg(pd + 2);
        v[h][0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w[0], a0), __fmul_rn(t.w[1], b0)), __fmul_rn(t.w[2], c0)), __fmul_rn(t.w[3], d0));
        v[h][1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w[0], a1), __fmul_rn(t.w[1], b1)), __fmul_rn(t.w[2], c1)), __fmul_rn(t.w[3], d1));
        v[h][2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w[0], a2), __fmul_rn(t.w[1], b2)), __fmul_rn(t.w[2], c2)), __fmul_rn(t.w[3], d2));
      }
    }
    float* o = out + (((size_t)b * oH + oy) * oW + ox0 + lane) * 3;
    if (ox0 + lane < oW) { __stcs(o, v[0][0]); __stcs(o + 1, v[0][1]); __stcs(o + 2, v[0][2]); }
    if (ox0 + lane + 32 < oW) { __stcs(o + 96, v[1][0]); __stcs(o + 97, v[1][1]); __stcs(o + 98, v[1][2]); }
  }
}

__global__ void flow_resize_kernel(FlowResize fr, float* __restrict__ out, int B) {
  const size_t total = (size_t)B * fr.H * fr.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % fr.W);
    const size_t r = i / fr.W;
    const int oy = (int)(r % fr.H);
    const int b = (int)(r / fr.H);
    reinterpret_cast<float2*>(out)[i] = fr.at(b, oy, ox);
  }
}

// outputs['predict_flow2'] * 384.0 / 382 (main_dl.py:497): multiply, then true division
__global__ void prescale_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float fhf) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(__ldg(in + i) * 384.0f, fhf);
}

// warp.py:25-43, one thread per batch element
__global__ void vec2mtrx_kernel(const float* __restrict__ p, float* __restrict__ out, int B, int warp_type,
                                int warp_approx) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float A[9];
  if (warp_type == 0) {
    const float* q = p + (size_t)b * 8;
    const float p1 = q[0], p2 = q[1], p3 = q[2], p4 = q[3], p5 = q[4], p6 = q[5], p7 = q[6], p8 = q[7];
    A[0] = p3; A[1] = p2; A[2] = p1;
    A[3] = p6; A[4] = -p3 - p7; A[5] = p5;
    A[6] = p4; A[7] = p8; A[8] = p7;
  } else {
    const float* q = p + (size_t)b * 6;
    A[0] = q[0]; A[1] = q[1]; A[2] = q[2];
    A[3] = q[3]; A[4] = q[4]; A[5] = q[5];
    A[6] = 0.f; A[7] = 0.f; A[8] = 0.f;
  }
  float M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, N[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float denom = 1.0f;
  for (int i = 1; i < warp_approx; ++i) {
    float T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) T[r * 3 + c] = N[r * 3 + 0] * A[c] + N[r * 3 + 1] * A[3 + c] + N[r * 3 + 2] * A[6 + c];
    denom *= (float)i;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      N[k] = T[k];
      M[k] += __fdiv_rn(T[k], denom);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) out[(size_t)b * 9 + k] = M[k];
}

int grid_for(size_t work_items, int threads) {
  const int sms = sm_count();
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)sms * 8;  // 8 resident 256-thread blocks per SM
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (int)blocks;
}

int g_warp_tuning = 0;   // fused warp only: rows per warp / resident blocks per SM under test (see launch_warp3)
int g_warp_variant = 3;  // 0 = direct gathers, 1 = staged (12-byte pixels), 2 = staged (padded 16-byte pixels), 3 = lean direct (default)

size_t tile_count(int B, int oH, int oW) {
  return (size_t)B * ((oH + kTileH - 1) / kTileH) * ((oW + kTileW - 1) / kTileW);
}
dim3 tile_grid3(int B, int oH, int oW) {
  return dim3((unsigned)((oW + kTileW - 1) / kTileW), (unsigned)((oH + kTileH - 1) / kTileH), (unsigned)B);
}
int tile_grid(size_t ntiles, int blocks_per_sm) {
  const size_t cap = (size_t)sm_count() * blocks_per_sm;
  return (int)std::max<size_t>(1, std::min(ntiles, cap));
}

// generic 4-tap samplers (grid_sample, Lie warp)
template <class Provider>
int launch_sampler(Provider prov, const float* img, float* out, int B, int srcH, int srcW, int oH, int oW, int C,
                   cudaStream_t st) {
  if (B == 0 || oH == 0 || oW == 0) return OFS_OK;
  if (C == 3 && (oW % 4) == 0 && (((uintptr_t)out) % 16 == 0)) {
    if ((size_t)srcH * srcW * 3 < (1u << 31) && B <= 65535 && (oH + kTileH - 1) / kTileH <= 65535) {
      if (prov.flag()) sample5_kernel<Provider, true><<<tile_grid3(B, oH, oW), 256, 0, st>>>(prov, img, out, B, srcH, srcW, oH, oW);
      else sample5_kernel<Provider, false><<<tile_grid3(B, oH, oW), 256, 0, st>>>(prov, img, out, B, srcH, srcW, oH, oW);
    } else {
      const size_t nt = tile_count(B, oH, oW);
      sample3_kernel<Provider><<<tile_grid(nt, 8), 256, 0, st>>>(prov, img, out, B, srcH, srcW, oH, oW);
    }
  } else {
    const size_t px = (size_t)B * oH * oW;
    sample_px_kernel<Provider><<<grid_for(px, 256), 256, 0, st>>>(prov, img, out, B, srcH, srcW, oH, oW, C);
  }
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

// tf_warp family: flow providers, C == 3 fast path
template <class Provider>
int launch_warp3(Provider prov, const float* img, float* out, int B, int H, int W, cudaStream_t st) {
  const size_t nt = tile_count(B, H, W);
  if (g_warp_variant == 3 && (size_t)H * W * 3 < (1u << 31) && B <= 65535 && (H + kTileH - 1) / kTileH <= 65535) {
    if constexpr (std::is_same<Provider, ResizeWarpProvider>::value) {
      const FlowResize& fr = prov.fr;
      if (fr.prescaled) {
        Resize2 rz{reinterpret_cast<const float2*>(fr.flow2), fr.fh, fr.fw, fr.hs, fr.ws, (float)W, (float)H};
        // streaming 32-bit stores straight from registers: the three stores of a warp fill whole sectors between them,
        // and dropping the shared-memory transpose saves ~12 instructions per pixel (47.6 -> 45.0 us at 8 x 720p)
        const float2* nf = nullptr;
        const dim3 g4((unsigned)((W + kTileW - 1) / kTileW), (unsigned)((H + 31) / 32), (unsigned)B);   // 4 rows per warp: 64 x 32 tiles
        switch (g_warp_tuning) {   // measurement only (ofs_set_warp_variant(3 + 16 t)); 0 = the shipped form
          case 1: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 4, 6>, g4, dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          case 2: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 2, 5>, tile_grid3(B, H, W), dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          case 3: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 2, 7>, tile_grid3(B, H, W), dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          case 4: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 4, 5>, g4, dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          case 5: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 4, 7>, g4, dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          case 6: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 2, 6>, tile_grid3(B, H, W), dim3(256), 0, st, img, nf, rz, out, B, H, W)); break;
          // shipped: 8 resident blocks per SM (32 registers, no spills): 42.5 us against 42.9 with 6 at 8 x 720p -- the kernel
          // is issue-bound and every extra warp helps; 4 rows per warp measured 51 us (benchmarks/warp_tune.py)
          default: OFS_CUDA(launch_pdl(warp5_kernel<true, true, 2, 8>, tile_grid3(B, H, W), dim3(256), 0, st, img, nf, rz, out, B, H, W));
        }
        OFS_LAUNCH_CHECK();
        return OFS_OK;
      }
    } else {
      Resize2 rz{};
      OFS_CUDA(launch_pdl(warp5_kernel<false, true, 2, 8>, tile_grid3(B, H, W), dim3(256), 0, st, img,
                          reinterpret_cast<const float2*>(prov.flow), rz, out, B, H, W));
      OFS_LAUNCH_CHECK();
      return OFS_OK;
    }
  }
  if (g_warp_variant == 2 && (((uintptr_t)img) % 16 == 0))
    OFS_CUDA(launch_pdl(warp4_kernel<Provider>, dim3(tile_grid(nt, 4)), dim3(256), 0, st, prov, img, out, B, H, W));
  else if (g_warp_variant == 1)
    OFS_CUDA(launch_pdl(warp3_kernel<Provider, true>, dim3(tile_grid(nt, 5)), dim3(256), 0, st, prov, img, out, B, H, W));
  else
    OFS_CUDA(launch_pdl(warp3_kernel<Provider, false>, dim3(tile_grid(nt, 8)), dim3(256), 0, st, prov, img, out, B, H, W));
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

}  // namespace

int tf_warp_impl(const float* img, const float* flow, float* out, int B, int H, int W, int C, cudaStream_t st) {
  OFS_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, "ofs_tf_warp: bad shape B=%d H=%d W=%d C=%d", B, H, W, C);
  if (B == 0) return OFS_OK;  // empty batch: nothing to do (pointers may be null)
  OFS_REQUIRE(img && flow && out, "ofs_tf_warp: null pointer");
  OFS_REQUIRE(((uintptr_t)flow) % 8 == 0, "ofs_tf_warp: flow must be 8-byte aligned");
  TfWarpProvider prov{flow, H, W};
  if (C == 3 && (W % 4) == 0 && (((uintptr_t)out) % 16 == 0)) return launch_warp3(prov, img, out, B, H, W, st);
  const size_t px = (size_t)B * H * W;
  sample_px_kernel<TfWarpProvider><<<grid_for(px, 256), 256, 0, st>>>(prov, img, out, B, H, W, H, W, C);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int flow_resize_impl(const float* flow2, float* out, int B, int fh, int fw, int H, int W, cudaStream_t st, float pre_mul = 384.0f) {
  OFS_REQUIRE(B >= 0 && fh > 0 && fw > 0 && H > 0 && W > 0, "ofs_flow_resize: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(flow2 && out, "ofs_flow_resize: null pointer");
  FlowResize fr{flow2, fh, fw, H, W, (float)fh / (float)H, (float)fw / (float)W, 0, pre_mul};
  flow_resize_kernel<<<grid_for((size_t)B * H * W, 256), 256, 0, st>>>(fr, out, B);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int flow_resize_warp_impl(const float* img, const float* flow2, float* out, int B, int H, int W, int fh, int fw,
                          cudaStream_t st, int prescaled) {
  OFS_REQUIRE(B >= 0 && fh > 0 && fw > 0 && H > 0 && W > 0, "ofs_flow_resize_warp: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(img && flow2 && out, "ofs_flow_resize_warp: null pointer");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (!prescaled && g_warp_variant == 3 && cap == cudaStreamCaptureStatusNone && (W % 4) == 0 && (((uintptr_t)out) % 16 == 0)) {
    // stand-alone op: (flow2 * 384) / fh once per flow texel into a stream-ordered scratch buffer (the network writes
    // that copy itself), then the lean fused kernel
    float* scratch = nullptr;
    const size_t nflt = (size_t)B * fh * fw * 2;
    {   // keep freed blocks in the device's default pool (the default threshold of 0 returns them to the OS at every
        // synchronisation, which made this call 7x slower than the kernel it feeds)
      static bool pool_set[64] = {false};
      int dev = 0;
      OFS_CUDA(cudaGetDevice(&dev));
      if (dev >= 0 && dev < 64 && !pool_set[dev]) {
        cudaMemPool_t pool;
        OFS_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = ~0ull;
        OFS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        pool_set[dev] = true;
      }
    }
    OFS_CUDA(cudaMallocAsync((void**)&scratch, nflt * 4, st));
    prescale_kernel<<<grid_for(nflt, 256), 256, 0, st>>>(flow2, scratch, nflt, (float)fh);
    OFS_LAUNCH_CHECK();
    ResizeWarpProvider pv{FlowResize{scratch, fh, fw, H, W, (float)fh / (float)H, (float)fw / (float)W, 1}};
    const int rc = launch_warp3(pv, img, out, B, H, W, st);
    OFS_CUDA(cudaFreeAsync(scratch, st));
    return rc;
  }
  ResizeWarpProvider prov{FlowResize{flow2, fh, fw, H, W, (float)fh / (float)H, (float)fw / (float)W, prescaled}};
  if ((W % 4) == 0 && (((uintptr_t)out) % 16 == 0)) return launch_warp3(prov, img, out, B, H, W, st);
  const size_t px = (size_t)B * H * W;
  sample_px_kernel<ResizeWarpProvider><<<grid_for(px, 256), 256, 0, st>>>(prov, img, out, B, H, W, H, W, 3);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

// clip driver: uint8 BGR frame in, np.uint8(totaloutputFrame) out (+ the float32 totaloutputFrame when out_f32 != null)
int flow_resize_warp_u8_impl(const uint8_t* img, const float* flow2_prescaled, uint8_t* out_u8, float* out_f32, int B, int H,
                             int W, int fh, int fw, cudaStream_t st) {
  OFS_REQUIRE(B > 0 && fh > 0 && fw > 0 && H > 0 && W > 0, "flow_resize_warp_u8: bad shape");
  OFS_REQUIRE(img && flow2_prescaled && out_u8, "flow_resize_warp_u8: null pointer");
  if ((W % 4) != 0 || ((uintptr_t)out_u8) % 4 != 0 || (size_t)H * W * 3 >= (1u << 31) || B > 65535 ||
      (H + kTileH - 1) / kTileH > 65535) {
    ResizeWarpProvider prov{FlowResize{flow2_prescaled, fh, fw, H, W, (float)fh / (float)H, (float)fw / (float)W, 1}};
    const size_t px = (size_t)B * H * W;
    warp_u8_px_kernel<<<grid_for(px, 256), 256, 0, st>>>(prov, img, out_u8, out_f32, B, H, W);
    OFS_LAUNCH_CHECK();
    return OFS_OK;
  }
  Resize2 rz{reinterpret_cast<const float2*>(flow2_prescaled), fh, fw, (float)fh / (float)H, (float)fw / (float)W, (float)W, (float)H};
  if (out_f32)
    OFS_CUDA(launch_pdl(warp5_u8_kernel<true>, tile_grid3(B, H, W), dim3(256), 0, st, img, rz, out_u8, out_f32, B, H, W));
  else
    OFS_CUDA(launch_pdl(warp5_u8_kernel<false>, tile_grid3(B, H, W), dim3(256), 0, st, img, rz, out_u8, out_f32, B, H, W));
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

}  // namespace ofs

extern "C" {

int ofs_set_warp_variant(int v) {
  ofs::g_warp_tuning = (v >= 16 && (v & 15) == 3) ? (v >> 4) : 0;
  if (v >= 16) v &= 15;
  ofs::g_warp_variant = (v >= 0 && v <= 3) ? v : 3;
  return OFS_OK;
}

int ofs_tf_warp(const float* img, const float* flow, float* out, int B, int H, int W, int C, ofs_stream stream) {
  return ofs::tf_warp_impl(img, flow, out, B, H, W, C, (cudaStream_t)stream);
}

int ofs_flow_resize(const float* flow2, float* out, int B, int fh, int fw, int H, int W, ofs_stream stream) {
  return ofs::flow_resize_impl(flow2, out, B, fh, fw, H, W, (cudaStream_t)stream);
}

int ofs_flow_resize_ex(const float* flow, float* out, int B, int fh, int fw, int H, int W, float pre_mul, ofs_stream stream) {
  return ofs::flow_resize_impl(flow, out, B, fh, fw, H, W, (cudaStream_t)stream, pre_mul);
}

int ofs_flow_resize_warp(const float* img, const float* flow2, float* out, int B, int H, int W, int fh, int fw,
                         ofs_stream stream) {
  return ofs::flow_resize_warp_impl(img, flow2, out, B, H, W, fh, fw, (cudaStream_t)stream, 0);
}

static int grid_sample(const float* im, const float* theta, float* out, int B, int H, int W, int C, int oH, int oW,
                       int projective, ofs_stream stream) {
  OFS_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && oH > 0 && oW > 0, "ofs_grid_sample: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(im && theta && out, "ofs_grid_sample: null pointer");
  ofs::GridSampleCoord gc;
  gc.theta = theta; gc.projective = projective; gc.H = H; gc.W = W; gc.oH = oH; gc.oW = oW;
  gc.step_x = oW > 1 ? 2.0f / (float)(oW - 1) : 0.0f;
  gc.step_y = oH > 1 ? 2.0f / (float)(oH - 1) : 0.0f;
  ofs::CoordProvider<ofs::GridSampleCoord> prov{gc};
  return ofs::launch_sampler(prov, im, out, B, H, W, oH, oW, C, (cudaStream_t)stream);
}

int ofs_grid_sample_affine(const float* im, const float* theta, float* out, int B, int H, int W, int C, int oH,
                           int oW, ofs_stream stream) {
  return grid_sample(im, theta, out, B, H, W, C, oH, oW, 0, stream);
}
int ofs_grid_sample_projective(const float* im, const float* theta, float* out, int B, int H, int W, int C, int oH,
                               int oW, ofs_stream stream) {
  return grid_sample(im, theta, out, B, H, W, C, oH, oW, 1, stream);
}

int ofs_vec2mtrx(const float* p, float* pMtrx, int B, int warp_type, int warp_approx, ofs_stream stream) {
  OFS_REQUIRE(warp_type == 0 || warp_type == 1, "ofs_vec2mtrx: warp_type must be 0 (homography) or 1 (affine)");
  OFS_REQUIRE(B >= 0 && warp_approx >= 1, "ofs_vec2mtrx: bad B / warp_approx");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(p && pMtrx, "ofs_vec2mtrx: null pointer");
  ofs::vec2mtrx_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p, pMtrx, B, warp_type, warp_approx);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int ofs_lie_warp(const float* image, const float* pMtrx, const float* refMtrx, float* out, int B, int srcH, int srcW,
                 int outH, int outW, ofs_stream stream) {
  OFS_REQUIRE(B >= 0 && srcH > 0 && srcW > 0 && outH > 0 && outW > 0, "ofs_lie_warp: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(image && pMtrx && refMtrx && out, "ofs_lie_warp: null pointer");
  ofs::LieCoord lc{pMtrx, refMtrx, srcH, srcW, outH, outW};
  ofs::CoordProvider<ofs::LieCoord> prov{lc};
  return ofs::launch_sampler(prov, image, out, B, srcH, srcW, outH, outW, 3, (cudaStream_t)stream);
}

}  // extern "C"
