// Kernels of the reference's OTHER test modes around the same network:
//
//   evaluate_originalSize_homo   main_flownetS_pyramid_noprevloss_dataloader.py:634-751
//       cv2.warpPerspective(frame_unstab, h, (out_w, out_h))  (:743)             -> ofs_warp_perspective_u8
//   evaluate                     main_flownetS_pyramid_noprevloss_dataloader.py:758-866
//       tf.image.resize_images(inputs[..., 24:27], [382, 510])  (:806)           -> ofs_tf1_resize_bilinear
//       cv2.resize(warped float32, (512, 384))  (:862)                           -> ofs_cv_resize_linear_f32
//   evaluate_blurNma             main_flownetS_pyramid.py:582-700
//       conv2d(flow plane, const 1/(75*75) [75,75,1,1], SAME) (:634-637), 0.9 * smooth + 0.1 * prev (:641)
//                                                                                -> ofs_flow_box_blur_ema
//   evaluate_medianNma           main_flownetS_pyramid.py:703-820
//       scipy.signal.medfilt(flow [382,510,2], 5)  (:809)                        -> ofs_medfilt_nd3
//
// All HBM-bound byte / float work: one thread per output element group, coalesced stores, no tensor cores.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "ofs_common.cuh"

namespace ofs {
namespace {

int grid_for_n(size_t n, int threads) {
  const size_t blocks = (n + threads - 1) / threads;
  return (int)std::min<size_t>(std::max<size_t>(blocks, 1), (size_t)sm_count() * 32);
}

// ------------------------------------------------------------------------------------------------
// cv2.warpPerspective, uint8, INTER_LINEAR, BORDER_CONSTANT 0 (imgwarp.cpp: WarpPerspectiveInvoker + remapBilinear).
// OpenCV's arithmetic, restated:
//   M = invert(H)  (closed form, double)                      -- on the host, see invert3x3()
//   per 64-pixel block at x_b, pixel x1 inside it, row y (all double, each product / sum rounded separately):
//       X0 = M0*x_b + M1*y + M2,  W = W0 + M6*x1,  W = W ? 32/W : 0,  fX = clamp((X0 + M0*x1) * W, INT_MIN, INT_MAX)
//       X = cvRound(fX)  (nearest, ties to even);  sx = sat_short(X >> 5), ax = X & 31;  same for Y
//   weights = 32x32 table of four 15-bit integers summing to 32768 (float32 products of (1 - a/32, a/32), rounded, the
//             largest / smallest entry absorbing a rounding deficit / surplus)
//   value   = (sum_k w_k * pixel_k + 2^14) >> 15, a tap outside the source counting as 0
constexpr int kWarpBlockW = 64;

struct Mat9 { double m[9]; };
struct Mat9x8 { Mat9 v[8]; };   // the inverse maps of up to 8 batch elements travel as a kernel argument

unsigned short g_tab_host[32 * 32 * 4];   // 0 .. 32768: the entry for the fraction (0, 0) is 32768 and needs all 16 bits
std::once_flag g_tab_once;

void build_tab() {
  float c1[32][2];
  const float scale = 1.0f / 32.0f;
  for (int i = 0; i < 32; ++i) { c1[i][0] = 1.0f - (float)i * scale; c1[i][1] = (float)i * scale; }
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      int iw[4], sum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const float v = c1[i][k1] * c1[j][k2];
          // (32 a b: an exact integer for every fraction.  OpenCV saturates the one value 32768 to 32767 in its int16 table;
          // with the other three weights 0 both give (w * p + 2^14) >> 15 == p for every byte p.)
          long r = lrintf(v * 32768.0f);
          iw[k1 * 2 + k2] = (int)r;
          sum += (int)r;
        }
      const int diff = sum - 32768;
      if (diff != 0) {
        int k = 0;
        for (int q = 1; q < 4; ++q)
          if (diff < 0 ? iw[q] > iw[k] : iw[q] < iw[k]) k = q;
        iw[k] -= diff;
      }
      for (int q = 0; q < 4; ++q) g_tab_host[(i * 32 + j) * 4 + q] = (unsigned short)iw[q];
    }
}

void invert3x3(const double* s, double* t) {
  const double det = s[0] * (s[4] * s[8] - s[5] * s[7]) - s[1] * (s[3] * s[8] - s[5] * s[6]) + s[2] * (s[3] * s[7] - s[4] * s[6]);
  if (det == 0.0) { for (int i = 0; i < 9; ++i) t[i] = 0.0; return; }
  const double d = 1.0 / det;
  t[0] = (s[4] * s[8] - s[5] * s[7]) * d;
  t[1] = (s[2] * s[7] - s[1] * s[8]) * d;
  t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
  t[3] = (s[5] * s[6] - s[3] * s[8]) * d;
  t[4] = (s[0] * s[8] - s[2] * s[6]) * d;
  t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
  t[6] = (s[3] * s[7] - s[4] * s[6]) * d;
  t[7] = (s[1] * s[6] - s[0] * s[7]) * d;
  t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
}

__global__ void __launch_bounds__(256) warp_perspective_u8_kernel(const uint8_t* __restrict__ src, const __grid_constant__ Mat9x8 mats,
                                                                    const unsigned short* __restrict__ tab_g, uint8_t* __restrict__ dst,
                                                                    int sh, int sw, int oh, int ow, int block_w) {
  __shared__ unsigned short tab[32 * 32 * 4];
  for (int i = threadIdx.x; i < 32 * 32 * 4 / 2; i += blockDim.x)
    reinterpret_cast<int*>(tab)[i] = __ldg(reinterpret_cast<const int*>(tab_g) + i);
  __syncthreads();
  const int b = blockIdx.z;
  const int y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= ow) return;
  const double* M = mats.v[b].m;
  const int xb = (x / block_w) * block_w;
  const double x1 = (double)(x - xb), xbd = (double)xb, yd = (double)y;
  // every operation rounded on its own, in OpenCV's order (no fused multiply-add)
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], xbd), __dmul_rn(M[1], yd)), M[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], xbd), __dmul_rn(M[4], yd)), M[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], xbd), __dmul_rn(M[7], yd)), M[8]);
  double Wd = __dadd_rn(W0, __dmul_rn(M[6], x1));
  Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
  double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], x1)), Wd);
  double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], x1)), Wd);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
  const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
  const unsigned short* w = tab + (((Y & 31) << 5) + (X & 31)) * 4;
  const uint8_t* sb = src + (size_t)b * sh * sw * 3;
  int acc[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = sy + (k >> 1), xx = sx + (k & 1);
    if (yy >= 0 && yy < sh && xx >= 0 && xx < sw) {
      const uint8_t* p = sb + ((size_t)yy * sw + xx) * 3;
      const int wk = (int)w[k];
      acc[0] += wk * (int)__ldg(p);
      acc[1] += wk * (int)__ldg(p + 1);
      acc[2] += wk * (int)__ldg(p + 2);
    }
  }
  uint8_t* o = dst + (((size_t)b * oh + y) * ow + x) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (uint8_t)min(255, max(0, (acc[c] + (1 << 14)) >> 15));
}

// ------------------------------------------------------------------------------------------------
// tf.image.resize_images(x, [oh, ow]) -- TF-1.10 legacy bilinear (align_corners=False: src = dst * in/out, no half-pixel
// offset), any channel count, a channel window [c0, c0 + C) of a wider NHWC tensor as source (inputs[..., 24:27]).
__global__ void tf1_resize_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int cs, int c0, int C,
                                  int oh, int ow, float hs, float ws) {
  const size_t total = (size_t)B * oh * ow * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % ow);
    r /= ow;
    const int oy = (int)(r % oh);
    const int b = (int)(r / oh);
    const float iy = (float)oy * hs, ix = (float)ox * ws;
    const int y0 = (int)floorf(iy), x0 = (int)floorf(ix);
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float yl = iy - (float)y0, xl = ix - (float)x0;
    const float* base = in + (size_t)b * H * W * cs + c0 + c;
    const float tl = __ldg(base + ((size_t)y0 * W + x0) * cs), tr = __ldg(base + ((size_t)y0 * W + x1) * cs);
    const float bl = __ldg(base + ((size_t)y1 * W + x0) * cs), br = __ldg(base + ((size_t)y1 * W + x1) * cs);
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    out[i] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
  }
}

// ------------------------------------------------------------------------------------------------
// cv2.resize(float32 image, (ow, oh)), INTER_LINEAR: half-pixel source coordinates, the fraction taken in double and
// rounded to float32, horizontal pass (s0 * (1 - fx) + s1 * fx) rounded to float32, then the vertical pass.
__global__ void cv_resize_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C, int oh, int ow,
                                     double scale_y, double scale_x, float post_mul) {
  const size_t total = (size_t)B * oh * ow * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int ox = (int)(r % ow);
    r /= ow;
    const int oy = (int)(r % oh);
    const int b = (int)(r / oh);
    double fd = ((double)ox + 0.5) * scale_x - 0.5;
    int sx = (int)floor(fd);
    float fx = (float)(fd - (double)sx);
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx + 1 >= W) { fx = 0.0f; sx = W - 1; }
    fd = ((double)oy + 0.5) * scale_y - 0.5;
    int sy = (int)floor(fd);
    float fy = (float)(fd - (double)sy);
    if (sy < 0) { fy = 0.0f; sy = 0; }
    if (sy + 1 >= H) { fy = 0.0f; sy = H - 1; }
    const int sx1 = min(sx + 1, W - 1), sy1 = min(sy + 1, H - 1);
    const float* base = in + (size_t)b * H * W * C + c;
    const float a0 = 1.0f - fx, b0 = 1.0f - fy;
    const float r0 = __fadd_rn(__fmul_rn(__ldg(base + ((size_t)sy * W + sx) * C), a0), __fmul_rn(__ldg(base + ((size_t)sy * W + sx1) * C), fx));
    const float r1 = __fadd_rn(__fmul_rn(__ldg(base + ((size_t)sy1 * W + sx) * C), a0), __fmul_rn(__ldg(base + ((size_t)sy1 * W + sx1) * C), fx));
    const float v = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, fy));
    out[i] = post_mul == 1.0f ? v : __fmul_rn(v, post_mul);   // main_dl.py:862 multiplies the resized image by 255
  }
}

// ------------------------------------------------------------------------------------------------
// conv2d(plane, const w [k,k,1,1], SAME) of both flow planes + the 0.9 / 0.1 mix with the previous flow.
// Separable running sums: pass 1 sums k rows per column into a float scratch plane (zero padded), pass 2 sums k columns
// of that and applies  a * smooth + b * prev.  Each tap is the float32 product x * w, as the convolution forms it; the
// sums are float32 in window order.  2 x (read + write) of the flow field: HBM-bound.
__global__ void box_rows_kernel(const float2* __restrict__ flow, float2* __restrict__ tmp, int B, int H, int W, int r, float wgt) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * 32;
  if (x >= W) return;
  const float2* f = flow + (size_t)b * H * W;
  float2* t = tmp + (size_t)b * H * W;
  // the first window of this thread's 32-row strip, then slide
  float sx = 0.0f, sy = 0.0f;
  for (int yy = y0 - r; yy <= y0 + r; ++yy)
    if (yy >= 0 && yy < H) { const float2 v = __ldg(f + (size_t)yy * W + x); sx += v.x * wgt; sy += v.y * wgt; }
  for (int y = y0; y < min(y0 + 32, H); ++y) {
    t[(size_t)y * W + x] = make_float2(sx, sy);
    const int out_y = y - r, in_y = y + r + 1;
    if (out_y >= 0) { const float2 v = __ldg(f + (size_t)out_y * W + x); sx -= v.x * wgt; sy -= v.y * wgt; }
    if (in_y < H) { const float2 v = __ldg(f + (size_t)in_y * W + x); sx += v.x * wgt; sy += v.y * wgt; }
  }
}

__global__ void box_cols_mix_kernel(const float2* __restrict__ tmp, const float2* __restrict__ prev, float2* __restrict__ out, int B, int H,
                                    int W, int r, float a, float bmix) {
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (x0 >= W) return;
  const float2* t = tmp + ((size_t)b * H + y) * W;
  float sx = 0.0f, sy = 0.0f;
  for (int xx = x0 - r; xx <= x0 + r; ++xx)
    if (xx >= 0 && xx < W) { const float2 v = __ldg(t + xx); sx += v.x; sy += v.y; }
  for (int x = x0; x < min(x0 + 16, W); ++x) {
    const size_t o = ((size_t)b * H + y) * W + x;
    float2 res = make_float2(sx, sy);
    if (prev) { const float2 p = __ldg(prev + o); res.x = __fadd_rn(__fmul_rn(a, sx), __fmul_rn(bmix, p.x)); res.y = __fadd_rn(__fmul_rn(a, sy), __fmul_rn(bmix, p.y)); }
    out[o] = res;
    const int out_x = x - r, in_x = x + r + 1;
    if (out_x >= 0) { const float2 v = __ldg(t + out_x); sx -= v.x; sy -= v.y; }
    if (in_x < W) { const float2 v = __ldg(t + in_x); sx += v.x; sy += v.y; }
  }
}

// ------------------------------------------------------------------------------------------------
// scipy.signal.medfilt(vol [H,W,C], k): a k x k x k window in EVERY axis -- the channel axis included -- zero padded,
// median = element (k^3) / 2 of the sorted window.  (For the reference's [382,510,2] flow and k = 5, 75 of the 125 window
// entries are padding zeros and the result is 0 everywhere; the kernel computes the general rule, not that constant.)
// One thread per output element: the in-range values are gathered, the out-of-range ones counted as zeros, and the
// rank is found by counting (no sort): value v is the median iff  #(< v) <= rank < #(<= v).
constexpr int kMedMaxWin = 7 * 7 * 7;

__global__ void medfilt3_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int C, int k) {
  const size_t total = (size_t)H * W * C;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W);
  const int y = (int)(i / ((size_t)C * W));
  const int r = k / 2;
  float vals[kMedMaxWin];
  int n = 0;
  for (int dy = -r; dy <= r; ++dy)
    for (int dx = -r; dx <= r; ++dx)
      for (int dc = -r; dc <= r; ++dc) {
        const int yy = y + dy, xx = x + dx, cc = c + dc;
        vals[n++] = (yy >= 0 && yy < H && xx >= 0 && xx < W && cc >= 0 && cc < C) ? __ldg(in + ((size_t)yy * W + xx) * C + cc) : 0.0f;
      }
  const int rank = n / 2;
  float med = 0.0f;
  for (int a = 0; a < n; ++a) {
    const float v = vals[a];
    int lt = 0, le = 0;
    for (int q = 0; q < n; ++q) { lt += vals[q] < v; le += vals[q] <= v; }
    if (lt <= rank && rank < le) { med = v; break; }
  }
  out[i] = med;
}

}  // namespace
}  // namespace ofs

extern "C" {

int ofs_warp_perspective_u8(const uint8_t* src, const double* h_host, uint8_t* dst, int B, int src_h, int src_w, int out_h, int out_w,
                            ofs_stream stream) {
  using namespace ofs;
  cudaStream_t st = (cudaStream_t)stream;
  OFS_REQUIRE(B >= 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0, "ofs_warp_perspective_u8: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(src && h_host && dst, "ofs_warp_perspective_u8: null pointer");
  OFS_REQUIRE(src_h <= 32767 && src_w <= 32767 && B <= 65535 && out_h <= 65535, "ofs_warp_perspective_u8: image too large");
  int dev = 0;
  OFS_CUDA(cudaGetDevice(&dev));
  int rc = require_sm100(dev);
  if (rc != OFS_OK) return rc;
  std::call_once(g_tab_once, build_tab);
  // the weight table lives on the device once per device (uploaded synchronously on first use)
  static std::mutex mu;
  static unsigned short* tab_dev[64] = {nullptr};
  OFS_REQUIRE(dev >= 0 && dev < 64, "ofs_warp_perspective_u8: device index %d", dev);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!tab_dev[dev]) {
      unsigned short* t = nullptr;
      OFS_CUDA(cudaMalloc((void**)&t, sizeof(g_tab_host)));
      OFS_CUDA(cudaMemcpy(t, g_tab_host, sizeof(g_tab_host), cudaMemcpyHostToDevice));
      tab_dev[dev] = t;
    }
  }
  const int block_w = std::min(kWarpBlockW, out_w);   // bw0 of WarpPerspectiveInvoker (images at least 16 rows high)
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = std::min(8, B - b0);
    Mat9x8 mats = {};
    for (int b = 0; b < nb; ++b) invert3x3(h_host + (size_t)(b0 + b) * 9, mats.v[b].m);
    dim3 grid((out_w + 255) / 256, out_h, nb);
    warp_perspective_u8_kernel<<<grid, 256, 0, st>>>(src + (size_t)b0 * src_h * src_w * 3, mats, tab_dev[dev],
                                                     dst + (size_t)b0 * out_h * out_w * 3, src_h, src_w, out_h, out_w, block_w);
    OFS_LAUNCH_CHECK();
  }
  return OFS_OK;
}

int ofs_tf1_resize_bilinear(const float* in, float* out, int B, int H, int W, int in_channels, int c0, int C, int out_h, int out_w,
                            ofs_stream stream) {
  using namespace ofs;
  OFS_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && c0 >= 0 && c0 + C <= in_channels && out_h > 0 && out_w > 0,
              "ofs_tf1_resize_bilinear: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(in && out, "ofs_tf1_resize_bilinear: null pointer");
  const size_t total = (size_t)B * out_h * out_w * C;
  tf1_resize_kernel<<<grid_for_n(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, in_channels, c0, C, out_h, out_w,
                                                                               (float)H / (float)out_h, (float)W / (float)out_w);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int ofs_cv_resize_linear_f32(const float* in, float* out, int B, int H, int W, int C, int out_h, int out_w, float post_mul,
                             ofs_stream stream) {
  using namespace ofs;
  OFS_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && out_h > 0 && out_w > 0, "ofs_cv_resize_linear_f32: bad shape");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(in && out, "ofs_cv_resize_linear_f32: null pointer");
  const size_t total = (size_t)B * out_h * out_w * C;
  const double scale_x = 1.0 / ((double)out_w / (double)W), scale_y = 1.0 / ((double)out_h / (double)H);
  cv_resize_f32_kernel<<<grid_for_n(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, out_h, out_w, scale_y, scale_x,
                                                                                  post_mul);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int ofs_flow_box_blur_ema(const float* flow, const float* prev, float* out, float* scratch, int B, int H, int W, int k, float a, float b,
                          ofs_stream stream) {
  using namespace ofs;
  cudaStream_t st = (cudaStream_t)stream;
  OFS_REQUIRE(B >= 0 && H > 0 && W > 0 && k >= 1 && (k & 1) == 1, "ofs_flow_box_blur_ema: bad shape (k must be odd)");
  if (B == 0) return OFS_OK;
  OFS_REQUIRE(flow && out && scratch && out != flow && scratch != flow && scratch != out, "ofs_flow_box_blur_ema: null / aliased pointer");
  OFS_REQUIRE(B <= 65535 && H <= 65535 * 32, "ofs_flow_box_blur_ema: field too large");
  const float wgt = (float)(1.0 / ((double)k * (double)k));   // tf.constant(1/(75*75.0)) as float32
  box_rows_kernel<<<dim3((W + 127) / 128, (H + 31) / 32, B), 128, 0, st>>>(reinterpret_cast<const float2*>(flow),
                                                                          reinterpret_cast<float2*>(scratch), B, H, W, k / 2, wgt);
  OFS_LAUNCH_CHECK();
  box_cols_mix_kernel<<<dim3(((W + 15) / 16 + 31) / 32, H, B), 32, 0, st>>>(reinterpret_cast<const float2*>(scratch),
                                                                           reinterpret_cast<const float2*>(prev),
                                                                           reinterpret_cast<float2*>(out), B, H, W, k / 2, a, b);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

int ofs_medfilt_nd3(const float* in, float* out, int H, int W, int C, int k, ofs_stream stream) {
  using namespace ofs;
  OFS_REQUIRE(H > 0 && W > 0 && C > 0 && k >= 1 && k <= 7 && (k & 1) == 1, "ofs_medfilt_nd3: bad shape (odd k <= 7)");
  OFS_REQUIRE(in && out && in != out, "ofs_medfilt_nd3: null / aliased pointer");
  const size_t total = (size_t)H * W * C;
  medfilt3_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(in, out, H, W, C, k);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

}  // extern "C"
