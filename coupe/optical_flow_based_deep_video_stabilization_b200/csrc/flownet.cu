// ofs_net: the FlowNetS-pyramid forward of the stabiliser (reference model.py:786-893) on sm_100a.
//
//   feats fp32 [B,384,512,27] --pack--> 16-bit [B,384,512,32]
//   10 encoder convs + 4 transposed convs (+ the 4 flow heads predict6..3 fused as extra columns) -> conv_gemm.cu
//   BatchNorm (moving stats, no gamma, eps 1e-5) is folded into W'/b' at load time
//   ConcatLayer            -> never materialised: producers write channel slices of concat buffers
//   flow pyramid           -> pyr_kernel: f_n = head_n + up(f_{n+1}) + up(f_{n+1})  (TF1 legacy bilinear)
//                             fused with the 2->2 k4s2 flow up-sampler written into the next concat slice
//   pad + NN-upsample + predict2 (model.py:882-887) -> 1x1 GEMM on the 96x128 grid (18 columns = 9 taps x 2)
//                             + predict2_gather_kernel (9-tap gather through the NN index maps, + 8 x up(f3))
//
// HBM layout (per batch element, 16-bit): x0 384x512x32 | conv1 192x256x64 | concat2 96x128x200
// [conv2 0:128 | deconv2 128:192 | up3_2 192:194 | pad] | conv3 48x64x256 | concat3 48x64x392 |
// conv4 24x32x512 | concat4 24x32x776 | conv5 12x16x512 | concat5 12x16x1032 | conv6, conv6_1 6x8x1024.
// Pad channels of the concat buffers are zeroed once and never written.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include <thread>

#include "conv_gemm.cuh"
#include "host_pack.h"

namespace ofs {
int flow_resize_warp_u8_impl(const uint8_t* img, const float* flow2_prescaled, uint8_t* out_u8, float* out_f32, int B, int H,
                             int W, int fh, int fw, cudaStream_t st);
int flow_resize_warp_impl(const float* img, const float* flow2, float* out, int B, int H, int W, int fh, int fw,
                          cudaStream_t st, int prescaled);
}

namespace {

using namespace ofs;

constexpr int kNetH = 384, kNetW = 512, kNetC = 27;
// Channel strides of the concat buffers (logical channels 194 / 386 / 770 / 1026): padded to a multiple of 64 channels, so
// that every pixel row starts on a 128-byte boundary and a 64-channel TMA box row is ONE aligned 128-byte line.  With the
// tight strides (200 / 392 / 776 / 1032) a box row straddled two lines: conv3 34.3 -> 31.5 us, conv4 17.7 -> 16.0,
// conv6 11.7 -> 10.0, deconv3 32.2 -> 24.3, deconv2 40.9 -> 38.7 alone at batch 8 (profiles/r02_tuning.md section 11).
// The pad channels are zero from allocation on and never written; their packed weights are zero.
constexpr int kCat2 = 256, kCat3 = 448, kCat4 = 832, kCat5 = 1088;
constexpr float kBnEps = 1e-5f;

__device__ __forceinline__ uint16_t cvt16(float v, int is_bf16) {
  if (is_bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// TF-1.10 ResizeBilinear (align_corners=False, legacy): src = dst * (in/out), no half-pixel offset
__device__ __forceinline__ float2 tf1_bilinear2(const float2* __restrict__ src, int h, int w, float hs, float ws,
                                                 int oy, int ox) {
  const float iy = (float)oy * hs, ix = (float)ox * ws;
  const int y0 = (int)floorf(iy), x0 = (int)floorf(ix);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float yl = iy - (float)y0, xl = ix - (float)x0;
  const float2 tl = src[y0 * w + x0], tr = src[y0 * w + x1], bl = src[y1 * w + x0], br = src[y1 * w + x1];
  float2 top, bot, v;
  top.x = tl.x + (tr.x - tl.x) * xl; top.y = tl.y + (tr.y - tl.y) * xl;
  bot.x = bl.x + (br.x - bl.x) * xl; bot.y = bl.y + (br.y - bl.y) * xl;
  v.x = top.x + (bot.x - top.x) * yl; v.y = top.y + (bot.y - top.y) * yl;
  return v;
}

struct PyrParams {
  const float2* hpart;   // fused head: per-phase shares at this level [nsplit][B,2h,2w] (written by the level's deconv GEMM)
  int nsplit;            // K splits of that GEMM (1 unless it runs split-K): each split leaves its own plane of shares
  long long split_stride;  // B * 2h * 2w
  float hb0, hb1;        // head bias
  const float2* f_prev;  // summed flow of the coarser level [B,h/2,w/2] or null (level 6)
  float2* f_out;         // summed flow at this level [B,h,w]
  const float* up_w;     // [4,4,2,2] (ky,kx,co,ci) then [2] bias
  uint16_t* concat;      // destination concat buffer [B,2h,2w,cstride]
  int B, h, w, cstride, coff, is_bf16;
};

// model.py:857/866/875 (ElementwiseLayer left fold) + model.py:852/861/870/879 (flow up-sampler)
__device__ __forceinline__ float2 pyr_flow_at(const PyrParams& p, int b, int i, int j) {
  // predictN = sum of the 4 phase shares (fixed order) + bias (model.py:847-848,855-856,864-865,873-874)
  const float2* hp = p.hpart + ((size_t)b * 2 * p.h + 2 * i) * (2 * p.w) + 2 * j;
  float2 s00 = hp[0], s01 = hp[1], s10 = hp[2 * p.w], s11 = hp[2 * p.w + 1];
  for (int k = 1; k < p.nsplit; ++k) {   // split-K: a phase share is the sum of its K splits, in split order
    const float2* hk = hp + (size_t)k * p.split_stride;
    const float2 t00 = hk[0], t01 = hk[1], t10 = hk[2 * p.w], t11 = hk[2 * p.w + 1];
    s00.x += t00.x; s00.y += t00.y; s01.x += t01.x; s01.y += t01.y;
    s10.x += t10.x; s10.y += t10.y; s11.x += t11.x; s11.y += t11.y;
  }
  float2 v;
  v.x = ((s00.x + s01.x) + (s10.x + s11.x)) + p.hb0;
  v.y = ((s00.y + s01.y) + (s10.y + s11.y)) + p.hb1;
  if (p.f_prev) {
    const int hp = p.h >> 1, wp = p.w >> 1;
    const float2 u = tf1_bilinear2(p.f_prev + (size_t)b * hp * wp, hp, wp, (float)hp / (float)p.h,
                                   (float)wp / (float)p.w, i, j);
    v.x = (v.x + u.x) + u.x;
    v.y = (v.y + u.y) + u.y;
  }
  return v;
}

__global__ void pyr_kernel(PyrParams p) {
  pdl_wait();
  pdl_launch_dependents();
  const int H2 = 2 * p.h, W2 = 2 * p.w;
  const size_t total = (size_t)p.B * H2 * W2;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % W2);
    const size_t r = idx / W2;
    const int oy = (int)(r % H2);
    const int b = (int)(r / H2);
    if (((oy | ox) & 1) == 0) p.f_out[((size_t)b * p.h + (oy >> 1)) * p.w + (ox >> 1)] = pyr_flow_at(p, b, oy >> 1, ox >> 1);
    float a0 = __ldg(p.up_w + 64), a1 = __ldg(p.up_w + 65);
    // the 2 x 2 taps of the k4 s2 up-sampler: flows are fetched unconditionally at clamped positions (all ~32 loads of a
    // pixel in flight together -- these kernels are a handful of blocks deep and purely latency-bound) and a tap outside
    // the grid is skipped when it is accumulated, so the sum has the same terms in the same order as before
    float2 f[2][2];
    int kyv[2], kxv[2];
    bool vy[2], vx[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      kyv[s] = ((oy + 1) & 1) + 2 * s;
      const int i = (oy + 1 - kyv[s]) >> 1;  // 2i + ky - 1 == oy
      vy[s] = i >= 0 && i < p.h;
      kxv[s] = ((ox + 1) & 1) + 2 * s;
      const int j = (ox + 1 - kxv[s]) >> 1;
      vx[s] = j >= 0 && j < p.w;
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = min(max((oy + 1 - kyv[s]) >> 1, 0), p.h - 1), j = min(max((ox + 1 - kxv[t]) >> 1, 0), p.w - 1);
        f[s][t] = pyr_flow_at(p, b, i, j);
      }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float* wk = p.up_w + (kyv[s] * 4 + kxv[t]) * 4;  // [co][ci]
        const float t0 = f[s][t].x * __ldg(wk + 0) + f[s][t].y * __ldg(wk + 1);
        const float t1 = f[s][t].x * __ldg(wk + 2) + f[s][t].y * __ldg(wk + 3);
        if (vy[s] && vx[t]) { a0 += t0; a1 += t1; }
      }
    uint16_t* o = p.concat + idx * p.cstride + p.coff;
    const uint32_t packed = (uint32_t)cvt16(a0, p.is_bf16) | ((uint32_t)cvt16(a1, p.is_bf16) << 16);
    *reinterpret_cast<uint32_t*>(o) = packed;
  }
}

// 16-bit network input as it crosses PCIe from ofs_net_stabilize_host ([npix, 27], rounded on the host) -> the packed
// [npix, 32] layout conv1 reads (channels 27..31 zero)
__global__ void repack27_kernel(const uint16_t* __restrict__ in, uint4* __restrict__ out, size_t npix) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t total = npix * 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i >> 2;
    const int c0 = (int)(i & 3) << 3;
    const uint16_t* src = in + pix * 27 + c0;
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < 27) ? (uint32_t)__ldg(src + j) : 0u;
    out[i] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
  }
}

struct Predict2Params {
  const float* P;     // [B,96,128,18]: column (ky*3+kx)*2 + o
  const float2* f3;   // [B,48,64]
  float2* f2;         // [B,382,510]
  float2* f2s;        // [B,382,510]: (f2 * 384.0) / 382, the operand of the test-mode flow glue (main_dl.py:497)
  float bias0, bias1;
  float sy, sx;       // NN align_corners scales (98-1)/(384-1), (130-1)/(512-1) in fp32
  float hs, ws;       // bilinear scales 48/382, 64/510 in fp32
  const short* iy_tab; // [384] NN source row of padded-grid row r, minus 1 (row of the unpadded 96-row grid, -1 / 96 = padding)
  const short* ix_tab; // [512] same for columns (128-column grid)
  int B;
};

// model.py:882-887: zero-pad(1) -> nearest(align_corners) to 384x512 -> 3x3 VALID -> + 8 x up(f3).
// One block per 64 x 8 output tile (3-D grid), a thread owns columns lane, lane+32 of one row: the NN source
// indices come from two small tables (built once on the host with the same fp32 round(dst * (in-1)/(out-1))),
// the row terms are shared by both pixels, all indexing is 32-bit.
// x * 384.0f / 382.0f with the rounding of the two float32 operations the reference performs (multiply, then true
// division): the quotient by the constant is q = t * RN(1/382) corrected once, q' = fma(fma(-382, q, t), RN(1/382), q) --
// the correctly rounded t / 382 whenever t is a normal number well inside the exponent range (Markstein: a faithful q,
// the exact remainder by FMA, the correctly rounded reciprocal); anything else takes the IEEE division.
__device__ __forceinline__ float mul384_div382(float x) {
  const float t = x * 384.0f;
  const float c = 0.00261780107393860817f;   // RN(1/382)
  const float at = fabsf(t);
  if (at > 1e-30f && at < 1e30f) {
    const float q = t * c;
    return __fmaf_rn(__fmaf_rn(-382.0f, q, t), c, q);
  }
  return __fdiv_rn(t, 382.0f);
}

__global__ void __launch_bounds__(256) predict2_gather_kernel(Predict2Params p) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int oy = blockIdx.y * 8 + wid;
  if (oy >= 382) return;
  const int ox0 = blockIdx.x * 64 + lane;
  // one 64-bit base per array, 32-bit element offsets: a single IMAD.WIDE.U32 per address
  const float* __restrict__ Pb = p.P + (size_t)b * 96 * 128 * 18;
  const float2* __restrict__ f3b = p.f3 + (size_t)b * 48 * 64;
  asm("" : "+l"(Pb));
  asm("" : "+l"(f3b));
  int rowoff[3];    // (iy * 128) * 18 + ky * 6 of the tap row, clamped into the array when the tap reads the zero padding
  bool rowok[3];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = (int)p.iy_tab[oy + ky];
    rowok[ky] = iy >= 0 && iy < 96;
    rowoff[ky] = (rowok[ky] ? iy : 0) * (128 * 18) + ky * 6;
  }
  // TF1 bilinear of f3 (48x64 -> 382x510): row terms
  const float fy = (float)oy * p.hs;
  const int y0 = (int)floorf(fy), y1 = min(y0 + 1, 47);
  const float yl = fy - (float)y0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int ox = ox0 + 32 * h;
    if (ox >= 510) continue;
    int ix[3];
    bool colok[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int v = (int)p.ix_tab[ox + kx];
      colok[kx] = v >= 0 && v < 128;
      ix[kx] = (colok[kx] ? v : 0) * 18 + kx * 2;
    }
    // all nine taps and the four f3 taps are fetched unconditionally (in-bounds addresses) and in flight together; a tap
    // on the zero padding contributes -0.0f, which leaves every partial sum unchanged (x + -0 == x for all x)
    float2 t[3][3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        t[ky][kx] = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const char*>(Pb) + (size_t)(unsigned)(rowoff[ky] + ix[kx]) * 4u));
    const float fx = (float)ox * p.ws;
    const int x0 = (int)floorf(fx), x1 = min(x0 + 1, 63);
    const float xl = fx - (float)x0;
    const float2 tl = __ldg(f3b + (unsigned)(y0 * 64 + x0)), tr = __ldg(f3b + (unsigned)(y0 * 64 + x1));
    const float2 bl = __ldg(f3b + (unsigned)(y1 * 64 + x0)), br = __ldg(f3b + (unsigned)(y1 * 64 + x1));
    float a0 = p.bias0, a1 = p.bias1;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const bool ok = rowok[ky] && colok[kx];
        a0 += ok ? t[ky][kx].x : -0.0f;
        a1 += ok ? t[ky][kx].y : -0.0f;
      }
    const float topx = tl.x + (tr.x - tl.x) * xl, topy = tl.y + (tr.y - tl.y) * xl;
    const float botx = bl.x + (br.x - bl.x) * xl, boty = bl.y + (br.y - bl.y) * xl;
    const float ux = topx + (botx - topx) * yl, uy = topy + (boty - topy) * yl;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0 += ux; a1 += uy; }  // ElementwiseLayer left fold, model.py:887
    const int idx = (b * 382 + oy) * 510 + ox;
    p.f2[idx] = make_float2(a0, a1);
    p.f2s[idx] = make_float2(mul384_div382(a0), mul384_div382(a1));
  }
}


// model.py:882-887 in ONE kernel: the 1x1 product (18 dot products of 194 channels per source pixel of concat2) on
// mma.sync tensor cores straight into shared memory, and the 9-tap gather through the NN index maps + 8 x up(f3) out of
// it -- no 7 MB product tensor, no second launch.  A block owns a 4 x 32 tile of the 96 x 128 source grid plus the row /
// column after it (every output pixel's three taps per axis fall on its first tap's source row / column or the next one:
// the NN scale is ~1/3.95), i.e. 5 x 33 = 165 source pixels = 11 m16 tiles, one per warp; it writes the output pixels
// whose FIRST tap lands in its 4 x 32 interior (row / column ranges from two small host tables).
// The sum per output pixel has the terms and the order of predict2_gather_kernel (bias, then ky-major taps; a tap on the
// zero padding adds +0.0f instead of being skipped: the same value).
struct P2FusedParams {
  const uint16_t* concat2;  // [B,96,128,kCat2] 16-bit
  const uint16_t* w;        // packed predict2 product weights [32 rows][256] K-major; row (ky*3+kx)*2 + o
  const float2* f3;         // [B,48,64]
  float2* f2;               // [B,382,510]
  float2* f2s;              // [B,382,510]: (f2 * 384.0) / 382
  float bias0, bias1;
  float hs, ws;             // bilinear scales 48/382, 64/510 in fp32
  const short* iy_tab;      // [384] NN source row - 1 of padded row r (-1 / 96: padding)
  const short* ix_tab;      // [512]
  const short* oy_start;    // [25] first output row whose first tap falls in source rows >= 4 i  ([24] = 382)
  const short* ox_start;    // [5]  same for 32-column source tiles ([4] = 510)
  int is_bf16;
};
constexpr int kP2RegionW = 33, kP2Region = 5 * kP2RegionW, kP2Threads = 352, kP2WStride = 216, kP2PStride = 24;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int is_bf16) {
  if (is_bf16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kP2Threads, 3) predict2_fused_kernel(P2FusedParams p) {
  // B fragments pre-arranged per (32-element K block, k16 half, n8 tile): 32 lanes x 8 bytes each, conflict-free LDS.64
  __shared__ __align__(16) uint2 Bf[7 * 2 * 3 * 32];
  __shared__ __align__(16) float Ps[176 * kP2PStride];
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 4, c0 = blockIdx.x * 32;
  // The K order inside a 32-element block is permuted so that a thread's A operands of both k16 halves are ONE 16-byte
  // load per pixel row: lane (g, t) holds elements 32 kb + 8 t + [0, 8) of rows g and g + 8; half s multiplies elements
  // 4 s .. 4 s + 3 of them (the fragment's "k = 2t, 2t+1" and "k = 2t+8, 2t+9" slots).  The B fragments are gathered
  // from the packed [32][256] K-major weights with the same permutation: a sum over k does not care about its order.
  {
    const uint32_t* __restrict__ wsrc = reinterpret_cast<const uint32_t*>(p.w);
    for (int i = threadIdx.x; i < 7 * 2 * 3 * 32; i += kP2Threads) {
      const int l = i & 31, nt = (i >> 5) % 3, s2 = (i / 96) & 1, kb = i / 192;
      const int g = l >> 2, t = l & 3;
      const uint32_t* src = wsrc + (nt * 8 + g) * 128 + 16 * kb + 4 * t + 2 * s2;   // words: element 32 kb + 8 t + 4 s2
      Bf[i] = make_uint2(__ldg(src), __ldg(src + 1));
    }
  }
  __syncthreads();
  // ---- product: warp w computes region pixels [16 w, 16 w + 16) x 24 columns over K = 7 x 32 (194 real)
  {
    const int g = lane >> 2, t = lane & 3;
    const uint4* rowp[2];
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pp = warp * 16 + g + 8 * h;
      const int rr = pp / kP2RegionW, cc = pp - rr * kP2RegionW;
      const int sr = r0 + rr, sc = c0 + cc;
      valid[h] = pp < kP2Region && sr < 96 && sc < 128;
      rowp[h] = reinterpret_cast<const uint4*>(p.concat2 + ((size_t)(b * 96 + min(sr, 95)) * 128 + min(sc, 127)) * kCat2);
    }
    float acc[3][4] = {};
    // 16-byte unit 4 kb + t of the pixel row; the last block's units past element 200 (zero weights) re-read unit 24.  The loads of block kb + 2 are in flight while block kb is multiplied.
    uint4 va[3], vb[3];
    va[0] = __ldg(rowp[0] + t); vb[0] = __ldg(rowp[1] + t);
    va[1] = __ldg(rowp[0] + 4 + t); vb[1] = __ldg(rowp[1] + 4 + t);
#pragma unroll
    for (int kb = 0; kb < 7; ++kb) {
      if (kb + 2 < 7) {
        const int u = min(4 * (kb + 2) + t, 24);
        va[(kb + 2) % 3] = __ldg(rowp[0] + u);
        vb[(kb + 2) % 3] = __ldg(rowp[1] + u);
      }
      const uint4 xa = va[kb % 3], xb = vb[kb % 3];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        uint32_t a[4];
        a[0] = s2 ? xa.z : xa.x; a[1] = s2 ? xb.z : xb.x;
        a[2] = s2 ? xa.w : xa.y; a[3] = s2 ? xb.w : xb.y;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const uint2 bw = Bf[((kb * 2 + s2) * 3 + nt) * 32 + lane];
          mma16816(acc[nt], a, bw.x, bw.y, p.is_bf16);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pp = warp * 16 + g + 8 * h;
      float* dst = Ps + pp * kP2PStride + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const float v0 = valid[h] ? acc[nt][2 * h] : 0.0f, v1 = valid[h] ? acc[nt][2 * h + 1] : 0.0f;
        *reinterpret_cast<float2*>(dst + nt * 8) = make_float2(v0, v1);
      }
    }
  }
  __syncthreads();
  // ---- gather: output pixels whose first tap lands in source rows [r0, r0 + 4) x columns [c0, c0 + 32)
  const int oy_lo = p.oy_start[blockIdx.y], oy_hi = p.oy_start[blockIdx.y + 1];
  const int ox_lo = p.ox_start[blockIdx.x], ox_hi = p.ox_start[blockIdx.x + 1];
  const int ncols = ox_hi - ox_lo, total = (oy_hi - oy_lo) * ncols;
  const float2* __restrict__ f3b = p.f3 + (size_t)b * 48 * 64;
  for (int idx = threadIdx.x; idx < total; idx += kP2Threads) {
    const int dy = idx / ncols;
    const int oy = oy_lo + dy, ox = ox_lo + (idx - dy * ncols);
    float a0 = p.bias0, a1 = p.bias1;
    int lc[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int v = (int)p.ix_tab[ox + kx];
      lc[kx] = v < 0 ? -1 : (v - c0) * kP2PStride + kx * 2;    // v - c0 in [0, 32]: column 128 is a zeroed region column
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int v = (int)p.iy_tab[oy + ky];
      if (v < 0) continue;
      const int rbase = (v - r0) * kP2RegionW * kP2PStride + ky * 6;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if (lc[kx] < 0) continue;
        const float2 pv = *reinterpret_cast<const float2*>(Ps + rbase + lc[kx]);
        a0 += pv.x;
        a1 += pv.y;
      }
    }
    const float fy = (float)oy * p.hs;
    const int y0 = (int)floorf(fy), y1 = min(y0 + 1, 47);
    const float yl = fy - (float)y0;
    const float fx = (float)ox * p.ws;
    const int x0 = (int)floorf(fx), x1 = min(x0 + 1, 63);
    const float xl = fx - (float)x0;
    const float2 tl = f3b[y0 * 64 + x0], tr = f3b[y0 * 64 + x1], bl = f3b[y1 * 64 + x0], br = f3b[y1 * 64 + x1];
    const float topx = tl.x + (tr.x - tl.x) * xl, topy = tl.y + (tr.y - tl.y) * xl;
    const float botx = bl.x + (br.x - bl.x) * xl, boty = bl.y + (br.y - bl.y) * xl;
    const float ux = topx + (botx - topx) * yl, uy = topy + (boty - topy) * yl;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a0 += ux; a1 += uy; }  // ElementwiseLayer left fold, model.py:887
    const int o = (b * 382 + oy) * 510 + ox;
    p.f2[o] = make_float2(a0, a1);
    p.f2s[o] = make_float2(__fdiv_rn(a0 * 384.0f, 382.0f), __fdiv_rn(a1 * 384.0f, 382.0f));
  }
}

// Flow heads predict6..predict3 (model.py:847-848,855-856,864-865,873-874: zero-pad 1 + 3x3 conv to 2 channels +
// bias) read the same input as the level's transposed conv, and the 4 sub-pixel phases of that conv together visit
// exactly the head's 9 taps.  Each head is therefore FUSED into its level's deconv GEMM as 16 extra accumulator
// columns (2 real) on the last N tile: phase (py,px) accumulates the head taps (dy,dx) with (dy==1, dx==1) ==
// (py,px), the epilogue writes the phase share per output pixel, and pyr_kernel adds the 4 shares and the bias.
// (The stand-alone mma.sync head kernel this replaces cost 90 us per step at batch 8 for 36 MMAC.)
struct Head {
  std::string name;   // "predict6" .. "predict3"
  int level, cin;
  float bias[2] = {0, 0};
};

struct Layer {
  std::string name;       // checkpoint layer name ("3_1", "deconv5", "predict4", ...)
  std::string bn;         // BN scope or ""
  ConvDesc d;             // B filled in at prepare()
  const void* in = nullptr;
  void* out = nullptr;
  int ksplit = 1;         // fixed split-K factor: independent of the batch so results are batch-invariant
  int block_n_run = 0;    // BLOCK_N used at run time (0 = d.block_n)
  int cta_group = 1;      // 2 = CTA pairs (conv_gemm2_kernel)
  ConvPlan plan;          // plan of the currently prepared batch size
  std::map<int, ConvPlan> plans;  // bound plans per batch size (TMA descriptors are per batch)
  void* w_dev = nullptr;
  float* b_dev = nullptr;
  unsigned* counters = nullptr;   // arrival / departure counters of the in-kernel split-K reduction (kCounters, zeroed once)
  size_t w_elems = 0;
  int n_pad = 0;
};
constexpr int kCounters = 2048;
constexpr int kMaxHeadSplits = 4;   // planes of fused-head shares per level (a split-K deconv writes one per split)

struct ActInfo { void* ptr; int H, W, cs, coff, C; };

}  // namespace

struct ofs_net {
  int device = 0, max_batch = 0, is_bf16 = 1;
  bool loaded = false;
  int prepared_B = 0;
  cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[64] = {nullptr}, ev_comp[64] = {nullptr};
  // 16-bit activations
  void *x0 = nullptr, *conv1 = nullptr, *concat2 = nullptr, *conv3 = nullptr, *concat3 = nullptr, *conv4 = nullptr,
       *concat4 = nullptr, *conv5 = nullptr, *concat5 = nullptr, *conv6 = nullptr, *conv6_1 = nullptr;
  // fp32
  float *hpart[7] = {nullptr}, *f[7] = {nullptr};  // index = pyramid level 3..6 (hpart: fused-head phase shares); f[2] = flow2
  float* P2 = nullptr;
  float* f2s = nullptr;  // pre-scaled flow2 for the fused flow-resize + warp
  float* upw = nullptr;  // 4 x (64 + 2) floats: upsample6_5, 5_4, 4_3, 3_2
  short* nn_tab = nullptr;  // predict2: NN align_corners source index tables, 384 rows then 512 columns (value - 1)
  short* p2_tab = nullptr;  // predict2_fused_kernel: first output row / column per source tile row (25) / column (5)
  int p2_fused = 0;         // 1 (OFS_P2_FUSED=1): predict2 product + gather in one kernel instead of the 1x1 GEMM + predict2_gather_kernel
  float p2_bias[2] = {0, 0};
  std::vector<Layer> layers;
  std::vector<Head> heads;   // predict6, predict5, predict4, predict3
  float* ws = nullptr;       // split-K workspace (grow-only)
  size_t ws_bytes = 0;
  // conv5 ... conv6_1 as one cooperative persistent launch (conv_chain_launch); OFS_CHAIN=1 turns it on (A/B)
  int chain = 0;
  unsigned* chain_sync = nullptr;   // the chain's grid-barrier words (zeroed once, self-maintained)
  std::map<std::string, ActInfo> acts;
  std::vector<void*> allocs;
  // host-API staging
  float *st_feats = nullptr, *st_frames = nullptr, *st_out = nullptr;
  size_t st_frames_cap = 0;
  // ofs_net_stabilize_host, bf16 nets: the float32 input is rounded to bf16 on the host (worker pool, pinned staging) and
  // crosses PCIe as [B,384,512,27] 16-bit; `feats_packed16` tells forward_impl that its `feats` argument is that array
  HostPacker* packer = nullptr;
  uint16_t* h_feats16 = nullptr;   // pinned, max_batch inputs
  int feats_packed16 = 0;
  int host_pack = 0;               // OFS_HOST_PACK=1: bf16 on the wire (rounded on the host); default float32
  // CUDA graphs of whole stabilize() steps, one per (B, H, W, flow2 wanted, alignment class of out).  The caller's four
  // pointers are baked into a handful of kernel nodes (pack_act: feats; the warp: frames, out; predict2_gather:
  // flow2_out); `patches` records where, so a call with other addresses updates those nodes of the instantiated
  // graph (cudaGraphExecKernelNodeSetParams) instead of re-capturing -- see ofs_net_stabilize.
  struct StepGraph {
    const void* baked[4];          // feats, frames, out, flow2 as currently set in `exec`
    int B, H, W;
    bool out_aligned16;
    bool feats_aligned16;          // selects the input-pack kernel at capture time (launch_pack_act)
    bool packed16;                 // feats is the 16-bit wire format (repack27_kernel instead of pack_act_kernel)
    cudaGraph_t graph;             // kept alive: its node handles address the nodes of `exec`
    cudaGraphExec_t exec;
    int launches;
    uint64_t last_use;
    struct Loc { int arg, off, which; };
    struct Patch {
      cudaGraphNode_t node;
      cudaKernelNodeParams params;                  // kernelParams -> argptr
      std::vector<std::vector<uint8_t>> argbuf;     // private copy of every argument
      std::vector<void*> argptr;
      std::vector<Loc> locs;                        // 8-byte words holding one of the four pointers
    };
    std::vector<Patch> patches;
  };
  std::vector<StepGraph> graphs;
  uint64_t graph_clock = 0;
  uint64_t graph_captures = 0, graph_updates = 0;   // how often a step was captured / re-pointed (tests, diagnostics)
  int use_graphs = 1;
  unsigned weights_generation = 0;   // bumped by ofs_net_load_weights AND whenever prepare() reallocates the split-K workspace:
                                     // graphs captured by dependants (clip driver) bake both in and are stale afterwards
};

namespace {

void drop_graphs(ofs_net* n) {
  for (auto& g : n->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  n->graphs.clear();
}

// Finds every kernel-node argument word of `graph` that equals one of the call's four pointers.
int collect_patches(cudaGraph_t graph, const void* const ptrs[4], std::vector<ofs_net::StepGraph::Patch>& out) {
  size_t nn = 0;
  OFS_CUDA(cudaGraphGetNodes(graph, nullptr, &nn));
  std::vector<cudaGraphNode_t> nodes(nn);
  if (nn) OFS_CUDA(cudaGraphGetNodes(graph, nodes.data(), &nn));
  for (cudaGraphNode_t node : nodes) {
    cudaGraphNodeType type;
    OFS_CUDA(cudaGraphNodeGetType(node, &type));
    if (type != cudaGraphNodeTypeKernel) continue;
    ofs_net::StepGraph::Patch pt;
    pt.node = node;
    OFS_CUDA(cudaGraphKernelNodeGetParams(node, &pt.params));
    if (!pt.params.kernelParams) continue;   // (never the case for launches made with an argument array)
    for (size_t i = 0;; ++i) {
      size_t off = 0, size = 0;
      if (cudaFuncGetParamInfo(pt.params.func, i, &off, &size) != cudaSuccess) { cudaGetLastError(); break; }
      const uint8_t* src = static_cast<const uint8_t*>(pt.params.kernelParams[i]);
      pt.argbuf.emplace_back(src, src + size);
      for (size_t o = 0; o + 8 <= size; o += 8) {
        const void* v;
        memcpy(&v, src + o, 8);
        if (!v) continue;
        for (int w = 0; w < 4; ++w)
          if (ptrs[w] && v == ptrs[w]) pt.locs.push_back({(int)i, (int)o, w});
      }
    }
    if (pt.locs.empty()) continue;
    out.push_back(std::move(pt));
  }
  for (auto& pt : out) {   // argptr must point into the Patch's final resting place
    pt.argptr.clear();
    for (auto& b : pt.argbuf) pt.argptr.push_back(b.data());
    pt.params.kernelParams = pt.argptr.data();
    pt.params.extra = nullptr;
  }
  return OFS_OK;
}

int dev_alloc(ofs_net* n, void** p, size_t bytes, bool zero) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return OFS_ENOMEM;
  }
  n->allocs.push_back(*p);
  if (zero) OFS_CUDA(cudaMemset(*p, 0, bytes));
  return OFS_OK;
}

Layer make_layer(const char* name, const char* bn, ConvKind kind, int H, int W, int cin, int in_cs, int cout, int k,
                 int stride, int block_n, int out_mode, int lrelu, int out_cstride, int out_coff, const void* in,
                 void* out) {
  Layer L;
  L.name = name; L.bn = bn;
  L.d.kind = kind; L.d.B = 1; L.d.H = H; L.d.W = W; L.d.cin = cin; L.d.in_cs = in_cs; L.d.cout = cout; L.d.k = k;
  L.d.stride = stride; L.d.block_n = block_n; L.d.out_mode = out_mode; L.d.lrelu = lrelu; L.d.is_bf16 = 1;
  L.d.out_cstride = out_cstride; L.d.out_coff = out_coff;
  L.in = in; L.out = out;
  return L;
}

// OFS_TUNE="layer:block_n:ksplit[:cta_group],..." overrides the tiling of individual layers (experiments only;
// block_n must keep the packed weight layout valid, i.e. divide the layer's padded N).
bool tune_lookup(const std::string& name, int* block_n, int* ksplit, int* cta_group, int* debug) {
  const char* env = getenv("OFS_TUNE");
  if (!env) return false;
  std::string s(env);
  size_t pos = 0;
  while (pos < s.size()) {
    size_t end = s.find(',', pos);
    if (end == std::string::npos) end = s.size();
    const std::string item = s.substr(pos, end - pos);
    const size_t c1 = item.find(':'), c2 = item.find(':', c1 == std::string::npos ? 0 : c1 + 1);
    if (c1 != std::string::npos && c2 != std::string::npos && item.substr(0, c1) == name) {
      *block_n = atoi(item.substr(c1 + 1, c2 - c1 - 1).c_str());
      *ksplit = atoi(item.substr(c2 + 1).c_str());
      const size_t c3 = item.find(':', c2 + 1);
      if (c3 != std::string::npos) {
        *cta_group = atoi(item.substr(c3 + 1).c_str());
        const size_t c4 = item.find(':', c3 + 1);
        if (c4 != std::string::npos) *debug = atoi(item.substr(c4 + 1).c_str());
      }
      return true;
    }
    pos = end + 1;
  }
  return false;
}

int prepare(ofs_net* n, int B) {
  if (n->prepared_B == B) return OFS_OK;
  if (!n->layers.empty() && n->layers[0].plans.count(B)) {  // cached
    for (Layer& L : n->layers) L.plan = L.plans[B];
    n->prepared_B = B;
    return OFS_OK;
  }
  size_t ws_need = 0;
  for (Layer& L : n->layers) {
    L.d.B = B;
    L.d.is_bf16 = n->is_bf16;
    L.d.ksplit = 1;
    int rc = conv_plan_geometry(L.plan, L.d);
    if (rc != OFS_OK) return rc;
    int bn = L.block_n_run ? L.block_n_run : L.d.block_n, ks = L.ksplit, cg = L.cta_group, dbg = 0;
    const bool tuned = tune_lookup(L.name, &bn, &ks, &cg, &dbg);
    if (tuned && (bn != L.d.block_n) && (L.n_pad % bn != 0)) {
      set_error("OFS_TUNE: block_n %d does not divide the padded N %d of layer %s", bn, L.n_pad, L.name.c_str());
      return OFS_EINVAL;
    }
    if (L.d.slab) { cg = 2; bn = L.d.block_n; ks = 1; }   // the packed K order is the slab order: tiling is fixed
    if (L.d.head && ks > kMaxHeadSplits) {
      set_error("layer %s: split-K %d exceeds the %d planes of fused-head shares", L.name.c_str(), ks, kMaxHeadSplits);
      return OFS_EINVAL;
    }
    if (L.d.stack) {   // stacked deconv: only 1 CTA / CTA pair is a choice
      bn = L.d.block_n; ks = 1; if (cg != 1) cg = 2;
      L.d.cta_group = cg;
      rc = conv_plan_geometry(L.plan, L.d);
      if (rc != OFS_OK) return rc;
    }
    if (!tuned && !L.d.slab && !L.d.stack && ks == 1 && L.d.out_mode == 0) {
      // small batches leave the wide-N tilings with a handful of tiles (conv4_1 at batch 1: 18): narrow the N tile
      // until ~100 tiles exist.  The N tiling does not touch the K summation order, so results stay bit-identical
      // across batch sizes (tests: batch independence).
      auto tiles_for = [&](int b) {
        ConvDesc d = L.d;
        d.block_n = b;
        d.cta_group = cg == 2 ? 2 : 1;
        ConvPlan t;
        if (conv_plan_geometry(t, d) != OFS_OK) return -1;
        return t.p.tiles_mp * t.p.tiles_n * t.p.phases;
      };
      int t = tiles_for(bn);
      for (int cand : {128, 64}) {
        if (t >= (cg == 2 ? 48 : 96) || cand >= bn) continue;
        const int tc = tiles_for(cand);
        if (tc > t) { bn = cand; t = tc; }
      }
    }
    if (cg == 8 && !L.d.slab) { L.d.kgroup = 2; cg = 1; }   // OFS_TUNE cta_group 8 = chunk groups
    if (cg == 32 && !L.d.slab && bn == 32) { L.d.kgroup = 4; cg = 1; }   // 32 = four chunks per stage (32-column tiles)
    const bool kc = cg == 16 && ks > 1 && ks <= 8 && bn == 256 && !L.d.slab && L.d.out_mode == 0;   // 16 = cluster split-K
    if (L.d.kgroup >= 2) { rc = conv_plan_geometry(L.plan, L.d); if (rc != OFS_OK) return rc; }
    if (ks > 1 || bn != L.d.block_n || cg == 2 || dbg) {
      ConvDesc d = L.d;
      d.block_n = bn;
      d.cta_group = cg == 2 ? 2 : 1;
      d.debug = dbg;
      d.ksplit = (L.d.out_mode == 0 && ks > 1) ? ks : 1;
      d.kcluster = kc ? 1 : 0;
      rc = conv_plan_geometry(L.plan, d);
      if (rc != OFS_OK) return rc;
    }
    ws_need = std::max(ws_need, L.plan.ws_bytes);
  }
  if (ws_need > n->ws_bytes) {
    // grow-only; sized for max_batch at once so that cached plans of other batch sizes stay valid
    OFS_CUDA(cudaDeviceSynchronize());
    if (n->ws) cudaFree(n->ws);
    n->ws = nullptr;
    n->ws_bytes = 0;
    for (Layer& L : n->layers) L.plans.clear();
    drop_graphs(n);   // they hold the old workspace pointer
    ++n->weights_generation;   // ... and so do the graphs of dependants (clip driver slots): they compare this counter
    const size_t want = std::max(ws_need, (ws_need / (size_t)B) * (size_t)n->max_batch);
    OFS_CUDA(cudaMalloc((void**)&n->ws, want));
    n->ws_bytes = want;
  }
  for (Layer& L : n->layers) {
    float* head_out = nullptr;
    if (L.d.head) head_out = n->hpart[L.name == "deconv5" ? 6 : L.name == "deconv4" ? 5 : L.name == "deconv3" ? 4 : 3];
    int rc = conv_plan_bind(L.plan, L.in, L.w_dev, L.b_dev, L.out, n->ws, head_out,
                            L.plan.n_counters <= kCounters ? L.counters : nullptr);
    if (rc != OFS_OK) return rc;
    L.plans[B] = L.plan;
  }
  n->prepared_B = B;
  return OFS_OK;
}

std::string norm_key(const char* raw) {
  std::string s(raw);
  const size_t colon = s.rfind(':');
  if (colon != std::string::npos && colon + 2 >= s.size()) s = s.substr(0, colon);  // ":0"
  // keep the last two path components: "<layer>/<var>"
  const size_t last = s.rfind('/');
  if (last == std::string::npos) return s;
  const size_t prev = s.rfind('/', last - 1);
  return prev == std::string::npos ? s : s.substr(prev + 1);
}

int launch_pyr(ofs_net* n, int level, int B, cudaStream_t st) {
  // level in {6,5,4,3}; grid h x w of that level; writes into the concat buffer of level-1
  static const int hs[7] = {0, 0, 0, 48, 24, 12, 6}, ws[7] = {0, 0, 0, 64, 32, 16, 8};
  PyrParams p;
  p.hpart = reinterpret_cast<const float2*>(n->hpart[level]);
  p.nsplit = 1;
  for (const Layer& L : n->layers)
    if (L.d.head && L.name == (level == 6 ? "deconv5" : level == 5 ? "deconv4" : level == 4 ? "deconv3" : "deconv2"))
      p.nsplit = std::max(1, L.plan.p.ksplit);
  p.split_stride = (long long)B * 4 * hs[level] * ws[level];
  p.hb0 = n->heads[6 - level].bias[0]; p.hb1 = n->heads[6 - level].bias[1];
  p.f_prev = level == 6 ? nullptr : reinterpret_cast<const float2*>(n->f[level + 1]);
  p.f_out = reinterpret_cast<float2*>(n->f[level]);
  p.up_w = n->upw + (6 - level) * 66;
  p.B = B; p.h = hs[level]; p.w = ws[level]; p.is_bf16 = n->is_bf16;
  switch (level) {
    case 6: p.concat = (uint16_t*)n->concat5; p.cstride = kCat5; p.coff = 1024; break;
    case 5: p.concat = (uint16_t*)n->concat4; p.cstride = kCat4; p.coff = 768; break;
    case 4: p.concat = (uint16_t*)n->concat3; p.cstride = kCat3; p.coff = 384; break;
    default: p.concat = (uint16_t*)n->concat2; p.cstride = kCat2; p.coff = 192; break;
  }
  const size_t total = (size_t)B * 4 * p.h * p.w;
  const int blocks = (int)std::min<size_t>((total + 127) / 128, (size_t)sm_count() * 8);
  pdl_set_kind(2);
  OFS_CUDA(launch_pdl(pyr_kernel, dim3(blocks), dim3(128), 0, st, p));
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

// `marks` (optional): one event is recorded on `st` after every kernel launch of the forward
struct Marks {
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  std::vector<double> macs;
  size_t used = 0;
  int mark(cudaStream_t st, const std::string& name, double m) {
    if (used == ev.size()) {
      cudaEvent_t e;
      OFS_CUDA(cudaEventCreate(&e));
      ev.push_back(e);
    }
    OFS_CUDA(cudaEventRecord(ev[used++], st));
    names.push_back(name);
    macs.push_back(m);
    return OFS_OK;
  }
};
#define OFS_MARK(name, m)                                \
  do {                                                   \
    if (marks) {                                         \
      int _rc = marks->mark(st, (name), (m));            \
      if (_rc != OFS_OK) return _rc;                     \
    }                                                    \
  } while (0)

// Launches layer i of n->layers -- or, at conv5 with the chain on, the whole conv5 ... conv6_1 chain.  *covered = number of
// layers the launch stands for (the caller skips the rest).
int launch_layer(ofs_net* n, size_t i, cudaStream_t st, int* covered) {
  *covered = 1;
  std::vector<Layer>& Ls = n->layers;
  if (n->chain && n->chain_sync && Ls[i].name == "5" && i + 3 < Ls.size() && Ls[i + 1].name == "5_1" && Ls[i + 2].name == "6" &&
      Ls[i + 3].name == "6_1") {
    const ConvPlan* plans[4] = {&Ls[i].plan, &Ls[i + 1].plan, &Ls[i + 2].plan, &Ls[i + 3].plan};
    if (conv_chain_supported(plans)) {
      *covered = 4;
      return conv_chain_launch(plans, n->chain_sync, st);
    }
  }
  return conv_launch(Ls[i].plan, st);
}

// feats == kFromX0: the packed 16-bit input buffer x0 has already been filled on the device (clip driver)
const float* const kFromX0 = reinterpret_cast<const float*>(uintptr_t(1));
int forward_impl(ofs_net* n, const float* feats, int B, float* f2_target, cudaStream_t st, Marks* marks = nullptr) {
  OFS_REQUIRE(n && feats, "ofs_net_forward: null pointer");
  OFS_REQUIRE(B >= 1 && B <= n->max_batch, "ofs_net_forward: batch %d outside [1, %d]", B, n->max_batch);
  if (!n->loaded) {
    set_error("ofs_net_forward: weights not loaded (call ofs_net_load_weights first)");
    return OFS_ESTATE;
  }
  OFS_CUDA(cudaSetDevice(n->device));
  int rc = prepare(n, B);
  if (rc != OFS_OK) return rc;
  OFS_MARK("start", 0.0);
  if (feats != kFromX0 && n->feats_packed16) {
    const size_t npix = (size_t)B * kNetH * kNetW;
    const int blocks = (int)std::min<size_t>((npix * 4 + 255) / 256, (size_t)sm_count() * 8);
    OFS_CUDA(launch_pdl(repack27_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const uint16_t*>(feats),
                        reinterpret_cast<uint4*>(n->x0), npix));
    OFS_LAUNCH_CHECK();
  } else if (feats != kFromX0) {
    rc = launch_pack_act(feats, n->x0, (size_t)B * kNetH * kNetW, kNetC, 32, n->is_bf16, st);
    if (rc != OFS_OK) return rc;
  }
  OFS_MARK("pack_input", 0.0);
  for (size_t li = 0; li < n->layers.size(); ++li) {
    Layer& L = n->layers[li];
    if (n->p2_fused && L.name == "predict2") continue;   // product + gather run as one kernel below
    int lvl = 0;
    if (L.name == "deconv5") lvl = 6;
    else if (L.name == "deconv4") lvl = 5;
    else if (L.name == "deconv3") lvl = 4;
    else if (L.name == "deconv2") lvl = 3;
    int covered = 1;
    rc = launch_layer(n, li, st, &covered);
    if (rc != OFS_OK) return rc;
    if (covered > 1) {
      double macs = 0.0;
      for (int k = 0; k < covered; ++k) macs += n->layers[li + k].plan.macs;
      OFS_MARK("gemm:" + L.name + ".." + n->layers[li + covered - 1].name, macs);
      li += (size_t)covered - 1;
      continue;
    }
    OFS_MARK(L.name == "predict2" ? std::string("gemm:predict2_product") : "gemm:" + L.name,
             L.name == "predict2" ? (double)B * 96 * 128 * 194 * 18 : L.plan.macs);
    if (lvl) {
      rc = launch_pyr(n, lvl, B, st);
      if (rc != OFS_OK) return rc;
      OFS_MARK("pyramid:level" + std::to_string(lvl), 0.0);
    }
  }
  if (n->p2_fused) {
    P2FusedParams fp;
    fp.concat2 = reinterpret_cast<const uint16_t*>(n->concat2);
    fp.w = nullptr;
    for (const Layer& L : n->layers) if (L.name == "predict2") fp.w = reinterpret_cast<const uint16_t*>(L.w_dev);
    fp.f3 = reinterpret_cast<const float2*>(n->f[3]);
    fp.f2 = reinterpret_cast<float2*>(f2_target ? f2_target : n->f[2]);
    fp.f2s = reinterpret_cast<float2*>(n->f2s);
    fp.bias0 = n->p2_bias[0]; fp.bias1 = n->p2_bias[1];
    fp.hs = 48.0f / 382.0f; fp.ws = 64.0f / 510.0f;
    fp.iy_tab = n->nn_tab; fp.ix_tab = n->nn_tab + 384;
    fp.oy_start = n->p2_tab; fp.ox_start = n->p2_tab + 25;
    fp.is_bf16 = n->is_bf16;
    OFS_REQUIRE(fp.w, "internal: predict2 weights missing");
    OFS_CUDA(launch_pdl(predict2_fused_kernel, dim3(4, 24, (unsigned)B), dim3(kP2Threads), 0, st, fp));
    OFS_LAUNCH_CHECK();
    OFS_MARK("predict2_fused", (double)B * 96 * 128 * 194 * 18);
    return OFS_OK;
  }
  Predict2Params pp;
  pp.P = n->P2;
  pp.f3 = reinterpret_cast<const float2*>(n->f[3]);
  pp.f2 = reinterpret_cast<float2*>(f2_target ? f2_target : n->f[2]);
  pp.f2s = reinterpret_cast<float2*>(n->f2s);
  pp.bias0 = n->p2_bias[0]; pp.bias1 = n->p2_bias[1];
  pp.sy = 97.0f / 383.0f; pp.sx = 129.0f / 511.0f;
  pp.hs = 48.0f / 382.0f; pp.ws = 64.0f / 510.0f;
  pp.B = B;
  pp.iy_tab = n->nn_tab; pp.ix_tab = n->nn_tab + 384;
  OFS_CUDA(launch_pdl(predict2_gather_kernel, dim3((510 + 63) / 64, (382 + 7) / 8, (unsigned)B), dim3(256), 0, st, pp));
  OFS_LAUNCH_CHECK();
  OFS_MARK("predict2_gather", 0.0);
  return OFS_OK;
}

}  // namespace

// ---- internal interface for the clip driver (clip.cu) -------------------------------------------------------------
namespace ofs {
void* net_x0(ofs_net* n) { return n->x0; }
int net_max_batch(const ofs_net* n) { return n->max_batch; }
int net_is_bf16(const ofs_net* n) { return n->is_bf16; }
int net_device(const ofs_net* n) { return n->device; }
bool net_loaded(const ofs_net* n) { return n->loaded; }
unsigned net_weights_generation(const ofs_net* n) { return n->weights_generation; }
int net_prepare(ofs_net* n, int B) { return prepare(n, B); }
// forward from the pre-filled x0 + fused flow glue / warp of frames -> out (all device pointers), on `st`
// clip driver: x0 already assembled on the device, uint8 frames in, np.uint8 frames (+ optional float32) out
int net_stabilize_u8_from_x0(ofs_net* n, const uint8_t* frames, uint8_t* out_u8, float* out_f32, int B, int H, int W,
                             cudaStream_t st) {
  int rc = forward_impl(n, kFromX0, B, nullptr, st);
  if (rc != OFS_OK) return rc;
  return flow_resize_warp_u8_impl(frames, n->f2s, out_u8, out_f32, B, H, W, 382, 510, st);
}

int net_stabilize_from_x0(ofs_net* n, const float* frames, float* out, int B, int H, int W, cudaStream_t st) {
  int rc = forward_impl(n, kFromX0, B, nullptr, st);
  if (rc != OFS_OK) return rc;
  return flow_resize_warp_impl(frames, n->f2s, out, B, H, W, 382, 510, st, 1);
}
}  // namespace ofs

extern "C" {

int ofs_net_create(ofs_net** out, int device, int max_batch, int precision) {
  OFS_REQUIRE(out, "ofs_net_create: null out pointer");
  *out = nullptr;
  OFS_REQUIRE(max_batch >= 1 && max_batch <= 64, "ofs_net_create: max_batch %d outside [1,64]", max_batch);
  OFS_REQUIRE(precision == OFS_PREC_BF16 || precision == OFS_PREC_FP16, "ofs_net_create: bad precision %d", precision);
  int rc = require_sm100(device);
  if (rc != OFS_OK) return rc;
  OFS_CUDA(cudaSetDevice(device));
  ofs_net* n = new ofs_net();
  n->device = device; n->max_batch = max_batch; n->is_bf16 = precision == OFS_PREC_BF16;
  { const char* e = getenv("OFS_GRAPH"); n->use_graphs = (e && e[0] == '0') ? 0 : 1; }
  // measured on the 16-vCPU host of the B200 boxes: 1584-1609 pairs/s packed (5-8 workers) against 1622 with float32 on
  // the wire -- there the host's memory system, which now also carries the conversion's read + write, is the limit, not
  // the bus.  Bit-identical either way (tested); opt-in.
  { const char* e = getenv("OFS_HOST_PACK"); n->host_pack = (e && e[0] == '1') ? 1 : 0; }
  // predict2 product + gather as ONE mma.sync kernel: built, byte-for-byte the same flow pipeline, and measured at parity with
  // the two launches it replaces (38.4-39.0 vs 38.8 us; 17.6 k vs 17.9 k pairs/s): opt-in (profiles/r02_tuning.md)
  { const char* e = getenv("OFS_P2_FUSED"); n->p2_fused = (e && e[0] == '1') ? 1 : 0; }
  // conv5 ... conv6_1 as one cooperative launch: built, bit-identical, and measured SLOWER than the eight launches it replaces
  // (83.7 us against ~72 us in the step: a grid barrier costs 2.3-3.5 us, as much as a kernel boundary, and the chain
  // needs two per layer; profiles/r02_tuning.md section 7): opt-in
  { const char* e = getenv("OFS_CHAIN"); n->chain = (e && e[0] == '1') ? 1 : 0; }
  rc = dev_alloc(n, (void**)&n->chain_sync, 16, true);
  if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
  const size_t B = (size_t)max_batch;
  struct { void** p; size_t elems; } bufs[] = {
      {&n->x0, B * 384 * 512 * 32},    {&n->conv1, B * 192 * 256 * 64}, {&n->concat2, B * 96 * 128 * kCat2},
      {&n->conv3, B * 48 * 64 * 256},  {&n->concat3, B * 48 * 64 * kCat3}, {&n->conv4, B * 24 * 32 * 512},
      {&n->concat4, B * 24 * 32 * kCat4}, {&n->conv5, B * 12 * 16 * 512},  {&n->concat5, B * 12 * 16 * kCat5},
      {&n->conv6, B * 6 * 8 * 1024},   {&n->conv6_1, B * 6 * 8 * 1024}};
  for (auto& b : bufs) {
    rc = dev_alloc(n, b.p, b.elems * 2, true);
    if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
  }
  static const int hs[7] = {0, 0, 382, 48, 24, 12, 6}, ws[7] = {0, 0, 510, 64, 32, 16, 8};
  for (int l = 2; l <= 6; ++l) {
    rc = dev_alloc(n, (void**)&n->f[l], B * hs[l] * ws[l] * 2 * 4, true);
    if (rc == OFS_OK && l >= 3) rc = dev_alloc(n, (void**)&n->hpart[l], kMaxHeadSplits * B * 4 * hs[l] * ws[l] * 2 * 4, true);
    if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
  }
  rc = dev_alloc(n, (void**)&n->P2, B * 96 * 128 * 18 * 4, true);
  if (rc == OFS_OK) rc = dev_alloc(n, (void**)&n->f2s, B * 382 * 510 * 2 * 4, true);
  if (rc == OFS_OK) rc = dev_alloc(n, (void**)&n->upw, 4 * 66 * 4, true);
  if (rc == OFS_OK) rc = dev_alloc(n, (void**)&n->nn_tab, (384 + 512) * 2, true);
  if (rc == OFS_OK) {
    // model.py:795-802,883: resize_nearest_neighbor(align_corners=True): src = min(round(dst * (in-1)/(out-1)), in-1),
    // evaluated in fp32 exactly as the kernel used to do per tap
    std::vector<short> tab(384 + 512);
    const float sy = 97.0f / 383.0f, sx = 129.0f / 511.0f;
    for (int r = 0; r < 384; ++r) { const float v = (float)r * sy; tab[r] = (short)(std::min((int)roundf(v), 97) - 1); }
    for (int c = 0; c < 512; ++c) { const float v = (float)c * sx; tab[384 + c] = (short)(std::min((int)roundf(v), 129) - 1); }
    rc = check_cuda(cudaMemcpy(n->nn_tab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice), "nn_tab upload", __FILE__, __LINE__);
    // predict2_fused_kernel: tile (i, j) of the source grid (4 rows x 32 columns) writes the output rows / columns whose
    // FIRST tap lands in it (the tables are non-decreasing; rows / columns on the top / left padding go to tile 0)
    std::vector<short> st(32, 0);
    for (int i = 1; i < 24; ++i) { int oy = 0; while (oy < 382 && tab[oy] < 4 * i) ++oy; st[i] = (short)oy; }
    st[24] = 382;
    for (int j = 1; j < 4; ++j) { int ox = 0; while (ox < 510 && tab[384 + ox] < 32 * j) ++ox; st[25 + j] = (short)ox; }
    st[25] = 0; st[29] = 510;
    if (rc == OFS_OK) rc = dev_alloc(n, (void**)&n->p2_tab, 32 * 2, true);
    if (rc == OFS_OK) rc = check_cuda(cudaMemcpy(n->p2_tab, st.data(), 32 * 2, cudaMemcpyHostToDevice), "p2_tab upload", __FILE__, __LINE__);
  }
  if (rc == OFS_OK) rc = dev_alloc(n, (void**)&n->st_feats, B * kNetH * kNetW * kNetC * 4, false);
  if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
  bool ok = cudaStreamCreateWithFlags(&n->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&n->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&n->s_d2h, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 64 && ok; ++i)
    ok = cudaEventCreateWithFlags(&n->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&n->ev_comp[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("cudaStreamCreate / cudaEventCreate failed");
    ofs_net_destroy(n);
    return OFS_ECUDA;
  }
  // execution order (model.py:807-885).  in_cs / out_cstride are the physical channel strides.
  auto& Ls = n->layers;
  Ls.push_back(make_layer("1", "1", kConv, 384, 512, 27, 32, 64, 7, 2, 64, 0, 1, 64, 0, n->x0, n->conv1));
  Ls.push_back(make_layer("2", "2", kConv, 192, 256, 64, 64, 128, 5, 2, 128, 0, 1, kCat2, 0, n->conv1, n->concat2));
  Ls.push_back(make_layer("3", "3", kConv, 96, 128, 128, kCat2, 256, 5, 2, 128, 0, 1, 256, 0, n->concat2, n->conv3));
  Ls.push_back(make_layer("3_1", "3_1", kConv, 48, 64, 256, 256, 256, 3, 1, 128, 0, 1, kCat3, 0, n->conv3, n->concat3));
  Ls.push_back(make_layer("4", "4", kConv, 48, 64, 256, kCat3, 512, 3, 2, 128, 0, 1, 512, 0, n->concat3, n->conv4));
  Ls.push_back(make_layer("4_1", "4_1", kConv, 24, 32, 512, 512, 512, 3, 1, 128, 0, 1, kCat4, 0, n->conv4, n->concat4));
  Ls.push_back(make_layer("5", "5", kConv, 24, 32, 512, kCat4, 512, 3, 2, 128, 0, 1, 512, 0, n->concat4, n->conv5));
  Ls.push_back(make_layer("5_1", "5_1", kConv, 12, 16, 512, 512, 512, 3, 1, 128, 0, 1, kCat5, 0, n->conv5, n->concat5));
  Ls.push_back(make_layer("6", "6", kConv, 12, 16, 512, kCat5, 1024, 3, 2, 128, 0, 1, 1024, 0, n->concat5, n->conv6));
  Ls.push_back(make_layer("6_1", "6_1", kConv, 6, 8, 1024, 1024, 1024, 3, 1, 128, 0, 1, 1024, 0, n->conv6, n->conv6_1));
  Ls.push_back(make_layer("deconv5", "deconv5_bn", kDeconvK4S2, 6, 8, 1024, 1024, 512, 4, 2, 128, 0, 1, kCat5, 512, n->conv6_1, n->concat5));
  Ls.push_back(make_layer("deconv4", "deconv4_bn", kDeconvK4S2, 12, 16, 1026, kCat5, 256, 4, 2, 128, 0, 1, kCat4, 512, n->concat5, n->concat4));
  Ls.push_back(make_layer("deconv3", "deconv3_bn", kDeconvK4S2, 24, 32, 770, kCat4, 128, 4, 2, 128, 0, 1, kCat3, 256, n->concat4, n->concat3));
  Ls.push_back(make_layer("deconv2", "deconv2_bn", kDeconvK4S2, 48, 64, 386, kCat3, 64, 4, 2, 64, 0, 1, kCat2, 128, n->concat3, n->concat2));
  // predict2 as a 1x1 GEMM with 18 columns on the 96x128 grid
  Ls.push_back(make_layer("predict2", "", kConv, 96, 128, 194, kCat2, 18, 1, 1, 32, 1, 0, 18, 0, n->concat2, n->P2));
  n->heads = {Head{"predict6", 6, 1024}, Head{"predict5", 5, 1026}, Head{"predict4", 4, 770}, Head{"predict3", 3, 386}};
  // Deep layers have few output tiles (conv6_1 at B=8: 32) but long K loops: their K loop is split over
  // several CTAs (fp32 partials in a workspace, summed in a fixed order).  The factors are constants of the
  // layer -- not of the batch -- so a frame pair's result is bit-identical whatever batch it rides in.
  // BLOCK_N = 256 where cout allows: the kernel is bound by L2 -> shared-memory delivery, and a wider N
  // tile halves the A-operand re-fetch per FLOP (measured: profiles/r01_tuning.md).
  for (Layer& L : Ls) {
    if ((L.name == "1" || L.name == "2") && !(getenv("OFS_NOSLAB") && getenv("OFS_NOSLAB")[0] == '1')) {
      L.d.slab = 1; L.d.cta_group = 2;   // x-shifted taps share one A slab per stage (CTA pairs); fixes the packed K order
      // conv1 with two output pixels per GEMM row (quad view; per kernel row 3 128-column and 2 64-column MMAs over quad
      // taps instead of 2 x 4 64-column MMAs over pair taps): built and tested, 74.6 vs 76.3 us alone, +0.4 % pairs/s one
      // step at a time and +2 % with two in flight.  Its conv1 output differs from the one-pixel form's in 0.016 % of the
      // elements by one bf16 ulp (another fp32 summation order), which the network amplifies to 0.009 px of flow: the
      // EPE against the fp32 oracle lands at 0.0165 / 0.0179 px instead of 0.0153 / 0.0151 on the two test inputs -- inside
      // the 0.02 px bound, but with less margin, so the one-pixel form stays the default (profiles/r02_tuning.md
      // sections 4 and 9).  OFS_CONV1X2=1 selects it.
      if (L.name == "1" && getenv("OFS_CONV1X2") && getenv("OFS_CONV1X2")[0] == '1') { L.d.slab = 2; L.d.block_n = 128; }
    } else if (L.name == "3") { L.block_n_run = 256; L.cta_group = 2; }   // CTA pairs: 36.9 vs 40.3 us (conv_bench)
    // (tilings below re-measured on the 64-channel-aligned concat strides: profiles/r02_sweep_aligned_strides.txt)
    else if (L.name == "3_1") { L.block_n_run = 256; L.cta_group = 2; }   // pairs 24.4 vs 25.3 us
    // cout 512 on 48 M tiles: 3 N tiles of 192 (the last one a third empty, clipped by the TMA store) = 144 tiles, ONE
    // wave on 148 SMs, instead of 2 x 256 = 96 tiles on 65 % of the SMs (-25 % time, conv_bench A/B)
    // single CTAs for both: conv4 15.0 vs 16.1 us on pairs, conv4_1 23.8 vs 24.8 (with the tight strides conv4 was faster
    // on pairs, 17.8 vs 18.5)
    else if (L.name == "4" || L.name == "4_1") { L.d.block_n = 192; }
    else if (L.name == "5" || L.name == "5_1") { L.block_n_run = 256; L.ksplit = 6; }
    // conv6 / conv6_1 (M = 384 rows): 128-column tiles x split-K 4 = 128 units move half the fp32 partials of 256 x 8
    // (alone: 11.8 vs 12.2 and 13.1 vs 14.8 us, profiles/r02_sweep_small_layers.txt)
    else if (L.name == "6" || L.name == "6_1") { L.block_n_run = 128; L.ksplit = 4; }
    // the level's flow head rides in the deconv GEMM (128 / 64-column tiles); CTA pairs halve the B fetch per CTA
    else if (L.d.kind == kDeconvK4S2) {
      // single CTAs: deconv5 13.8 vs 14.5 us on pairs, deconv4 22.7 vs 24.4; deconv3 on pairs 24.7 vs 25.4
      L.d.head = 1; L.cta_group = L.name == "deconv3" ? 2 : 1;
      // deconv4 (M = 1536 rows: 12 M tiles x 4 phases = 48 units of 68 K blocks) keeps a third of the machine busy, but
      // single CTAs x split-K 3 = 144 units (the head's shares leave per split and are summed by pyr_kernel: built,
      // tested, OFS_TUNE=deconv4:128:3:1) measured 27.1 us against 24.5 us for the pairs: the fp32 partials and the
      // reduce launch cost more than the idle SMs (profiles/r02_tuning.md section 7)
      // deconv2 (cout 64) has a second form: all four sub-pixel phases stacked in one accumulator tile, each input tap
      // fetched once (deconv_stack_kernel; fixes the packed weight layout).  It beat the per-phase form while the A boxes
      // straddled cache lines (40.4 vs 46.4 us); on the aligned strides the per-phase form is the faster one -- 33.4 us on
      // single CTAs, 33.7 on pairs, against 38.5 / 39.2 stacked -- and is the default.  OFS_STACK=1: the stacked form (A/B).
      if (L.name == "deconv2" && getenv("OFS_STACK") && getenv("OFS_STACK")[0] == '1') { L.d.stack = 1; L.d.cta_group = 2; }
    }
  }
  for (Layer& L : Ls) {
    L.d.is_bf16 = n->is_bf16;
    rc = conv_plan_geometry(L.plan, L.d);
    if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
    L.w_elems = (size_t)L.plan.w_rows * L.plan.k_total;
    L.n_pad = L.plan.p.n_pad;
    rc = dev_alloc(n, &L.w_dev, L.w_elems * 2, true);
    if (rc == OFS_OK) rc = dev_alloc(n, (void**)&L.b_dev, (size_t)L.n_pad * 4, true);
    if (rc == OFS_OK && L.ksplit > 1) rc = dev_alloc(n, (void**)&L.counters, (size_t)kCounters * 4, true);
    if (rc != OFS_OK) { ofs_net_destroy(n); return rc; }
  }
  n->acts = {
      {"conv1", {n->conv1, 192, 256, 64, 0, 64}},       {"conv2", {n->concat2, 96, 128, kCat2, 0, 128}},
      {"conv3", {n->conv3, 48, 64, 256, 0, 256}},       {"conv3_1", {n->concat3, 48, 64, kCat3, 0, 256}},
      {"conv4", {n->conv4, 24, 32, 512, 0, 512}},       {"conv4_1", {n->concat4, 24, 32, kCat4, 0, 512}},
      {"conv5", {n->conv5, 12, 16, 512, 0, 512}},       {"conv5_1", {n->concat5, 12, 16, kCat5, 0, 512}},
      {"conv6", {n->conv6, 6, 8, 1024, 0, 1024}},       {"conv6_1", {n->conv6_1, 6, 8, 1024, 0, 1024}},
      {"concat5", {n->concat5, 12, 16, kCat5, 0, 1026}}, {"concat4", {n->concat4, 24, 32, kCat4, 0, 770}},
      {"concat3", {n->concat3, 48, 64, kCat3, 0, 386}},   {"concat2", {n->concat2, 96, 128, kCat2, 0, 194}},
      {"input", {n->x0, 384, 512, 32, 0, 27}}};
  *out = n;
  return OFS_OK;
}

int ofs_net_destroy(ofs_net* n) {
  if (!n) return OFS_OK;
  cudaSetDevice(n->device);
  cudaDeviceSynchronize();
  if (n->stream) cudaStreamDestroy(n->stream);
  if (n->s_h2d) cudaStreamDestroy(n->s_h2d);
  if (n->s_d2h) cudaStreamDestroy(n->s_d2h);
  for (int i = 0; i < 64; ++i) {
    if (n->ev_h2d[i]) cudaEventDestroy(n->ev_h2d[i]);
    if (n->ev_comp[i]) cudaEventDestroy(n->ev_comp[i]);
  }
  drop_graphs(n);
  if (n->packer) host_packer_destroy(n->packer);
  if (n->h_feats16) cudaFreeHost(n->h_feats16);
  for (void* p : n->allocs) cudaFree(p);
  if (n->ws) cudaFree(n->ws);
  if (n->st_frames) cudaFree(n->st_frames);
  if (n->st_out) cudaFree(n->st_out);
  delete n;
  return OFS_OK;
}

int ofs_net_load_weights(ofs_net* n, const ofs_named_array* arrays, int count) {
  OFS_REQUIRE(n && arrays && count > 0, "ofs_net_load_weights: null / empty input");
  OFS_CUDA(cudaSetDevice(n->device));
  n->loaded = false;   // a failed ingest leaves the net unusable rather than half-updated
  std::map<std::string, const ofs_named_array*> by_name;
  for (int i = 0; i < count; ++i) {
    OFS_REQUIRE(arrays[i].name && arrays[i].data, "ofs_net_load_weights: entry %d has a null name / data", i);
    by_name[norm_key(arrays[i].name)] = &arrays[i];
  }
  // Every variable the reference checkpoint carries for this network (main_dl.py:330 saves ALL 'main_net' variables,
  // moving statistics included) is REQUIRED: a missing or differently named bias / beta / moving_mean /
  // moving_variance must not silently load as 0 / 1, and an array nothing consumed (a gamma, a layer of another
  // model) is reported instead of dropped.
  std::map<std::string, bool> consumed;
  for (auto& kv : by_name) consumed[kv.first] = false;
  auto find = [&](const std::string& key, int64_t numel, const float** ptr, bool required) -> int {
    auto it = by_name.find(key);
    if (it == by_name.end()) {
      *ptr = nullptr;
      if (required) { set_error("ofs_net_load_weights: missing array '%s'", key.c_str()); return OFS_EINVAL; }
      return OFS_OK;
    }
    consumed[key] = true;
    if (it->second->numel != numel) {
      set_error("ofs_net_load_weights: '%s' has %lld elements, expected %lld", key.c_str(), (long long)it->second->numel,
                (long long)numel);
      return OFS_EINVAL;
    }
    *ptr = it->second->data;
    return OFS_OK;
  };
  for (Layer& L : n->layers) {
    const bool deconv = L.d.kind == kDeconvK4S2;
    const bool is_p2 = L.name == "predict2";
    const int cin = L.d.cin, cout = is_p2 ? 2 : L.d.cout, k = is_p2 ? 3 : L.d.k;
    const int64_t wn = (int64_t)k * k * cin * cout;
    const float *w = nullptr, *b = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr;
    int rc = find(L.name + (deconv ? "/W_deconv2d" : "/W_conv2d"), wn, &w, true);
    if (rc == OFS_OK) rc = find(L.name + (deconv ? "/b_deconv2d" : "/b_conv2d"), cout, &b, true);
    if (rc == OFS_OK && !L.bn.empty()) {
      rc = find(L.bn + "/beta", cout, &beta, true);
      if (rc == OFS_OK) rc = find(L.bn + "/moving_mean", cout, &mean, true);
      if (rc == OFS_OK) rc = find(L.bn + "/moving_variance", cout, &var, true);
    }
    if (rc != OFS_OK) return rc;
    // fold BN: W' = W r, b' = (b - mu) r + beta, r = 1/sqrt(var + eps)   (double math, fp32 result)
    std::vector<float> wf((size_t)wn), bf((size_t)cout, 0.0f);
    std::vector<double> r((size_t)cout, 1.0);
    for (int c = 0; c < cout; ++c) {
      if (!L.bn.empty()) r[c] = 1.0 / std::sqrt((double)(var ? var[c] : 1.0f) + (double)kBnEps);
      const double bb = b ? b[c] : 0.0, mu = mean ? mean[c] : 0.0, be = beta ? beta[c] : 0.0;
      bf[c] = (float)((bb - mu) * r[c] + be);
    }
    if (deconv) {  // [4,4,cout,cin]
      for (int t = 0; t < 16; ++t)
        for (int c = 0; c < cout; ++c) {
          const size_t o = ((size_t)t * cout + c) * cin;
          for (int ci = 0; ci < cin; ++ci) wf[o + ci] = (float)((double)w[o + ci] * r[c]);
        }
    } else {  // [k,k,cin,cout]
      for (size_t i = 0; i < (size_t)wn; ++i) wf[i] = (float)((double)w[i] * r[i % cout]);
    }
    std::vector<uint16_t> wp;
    std::vector<float> bp;
    if (is_p2) {
      // [3,3,194,2] -> 1x1 conv with 18 output columns: column (ky*3+kx)*2 + o; bias applied in the gather
      std::vector<float> w1((size_t)cin * 18);
      for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < cin; ++ci)
          for (int o = 0; o < 2; ++o) w1[(size_t)ci * 18 + t * 2 + o] = wf[((size_t)t * cin + ci) * 2 + o];
      conv_pack_weights(L.plan, w1.data(), nullptr, wp, bp);
      n->p2_bias[0] = bf[0]; n->p2_bias[1] = bf[1];
    } else if (L.d.head) {
      // the flow head of this level (predict6 with deconv5, ... predict3 with deconv2) shares the GEMM
      Head& h = n->heads[L.name == "deconv5" ? 0 : L.name == "deconv4" ? 1 : L.name == "deconv3" ? 2 : 3];
      const float *hw = nullptr, *hb = nullptr;
      rc = find(h.name + "/W_conv2d", (int64_t)9 * h.cin * 2, &hw, true);
      if (rc == OFS_OK) rc = find(h.name + "/b_conv2d", 2, &hb, true);
      if (rc != OFS_OK) return rc;
      OFS_REQUIRE(h.cin == cin, "internal: head / deconv input mismatch for %s", L.name.c_str());
      h.bias[0] = hb ? hb[0] : 0.f;
      h.bias[1] = hb ? hb[1] : 0.f;
      conv_pack_weights(L.plan, wf.data(), bf.data(), wp, bp, hw);
    } else {
      conv_pack_weights(L.plan, wf.data(), bf.data(), wp, bp);
    }
    OFS_REQUIRE(wp.size() == L.w_elems && (int)bp.size() == L.n_pad, "internal: packed size mismatch for %s", L.name.c_str());
    OFS_CUDA(cudaMemcpy(L.w_dev, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
    OFS_CUDA(cudaMemcpy(L.b_dev, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
  }
  static const char* ups[4] = {"upsample6_5", "upsample5_4", "upsample4_3", "upsample3_2"};
  std::vector<float> upw(4 * 66, 0.0f);
  for (int i = 0; i < 4; ++i) {
    const float *w = nullptr, *b = nullptr;
    int rc = find(std::string(ups[i]) + "/W_deconv2d", 64, &w, true);
    if (rc == OFS_OK) rc = find(std::string(ups[i]) + "/b_deconv2d", 2, &b, true);
    if (rc != OFS_OK) return rc;
    memcpy(&upw[i * 66], w, 64 * 4);
    if (b) { upw[i * 66 + 64] = b[0]; upw[i * 66 + 65] = b[1]; }
  }
  {
    std::string extra;
    int n_extra = 0;
    for (auto& kv : consumed)
      if (!kv.second) { if (n_extra < 6) extra += (n_extra ? ", " : "") + kv.first; ++n_extra; }
    if (n_extra) {
      set_error("ofs_net_load_weights: %d array(s) match no variable of flownetS_pyramid (model.py:786-893 has no gamma, "
                "no other layers): %s%s", n_extra, extra.c_str(), n_extra > 6 ? ", ..." : "");
      return OFS_EINVAL;
    }
  }
  OFS_CUDA(cudaMemcpy(n->upw, upw.data(), upw.size() * 4, cudaMemcpyHostToDevice));
  n->loaded = true;
  ++n->weights_generation;
  n->prepared_B = 0;
  for (Layer& L : n->layers) L.plans.clear();
  OFS_CUDA(cudaDeviceSynchronize());
  drop_graphs(n);   // head biases etc. are kernel arguments
  return OFS_OK;
}

int ofs_net_forward(ofs_net* n, const float* feats, int B, float* f6, float* f5, float* f4, float* f3, float* f2,
                    ofs_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward_impl(n, feats, B, f2, st);
  if (rc != OFS_OK) return rc;
  static const int hs[7] = {0, 0, 0, 48, 24, 12, 6}, ws[7] = {0, 0, 0, 64, 32, 16, 8};
  float* outs[7] = {nullptr, nullptr, nullptr, f3, f4, f5, f6};
  for (int l = 3; l <= 6; ++l)
    if (outs[l])
      OFS_CUDA(cudaMemcpyAsync(outs[l], n->f[l], (size_t)B * hs[l] * ws[l] * 2 * 4, cudaMemcpyDeviceToDevice, st));
  return OFS_OK;
}

// One step = 26 dependent kernels of 3-100 us each: replayed as ONE CUDA graph, so neither host launch latency nor
// inter-kernel drain sits between them.  A graph bakes in its pointers; the reference's call pattern feeds a NEW numpy
// array every frame (main_dl.py:568-569) and the Python drop-in allocates a fresh output per call, so the graph is
// cached per SHAPE and the few kernel-node arguments that hold the caller's pointers are re-pointed in the
// instantiated graph when the addresses differ from the previous launch (microseconds on the host, nothing on the
// device) -- one capture per (B, H, W), whatever the addresses.  OFS_GRAPH=0 (read at ofs_net_create) falls back to
// plain stream launches.
int ofs_net_stabilize(ofs_net* n, const float* feats, const float* frames, float* out, float* flow2_out, int B, int H,
                      int W, ofs_stream stream) {
  OFS_REQUIRE(n && feats && frames && out, "ofs_net_stabilize: null pointer");
  OFS_REQUIRE(H > 0 && W > 0, "ofs_net_stabilize: bad frame size %dx%d", H, W);
  OFS_REQUIRE(B >= 1 && B <= n->max_batch, "ofs_net_stabilize: batch %d outside [1, %d]", B, n->max_batch);
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (n->use_graphs) cudaStreamIsCapturing(st, &cap);
  const void* ptrs[4] = {feats, frames, out, flow2_out};
  bool distinct = true;   // the patch table tells the four pointers apart by value
  for (int i = 0; i < 4; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (ptrs[i] && ptrs[i] == ptrs[j]) distinct = false;
  if (!n->use_graphs || !n->loaded || cap != cudaStreamCaptureStatusNone || !distinct) {   // plain path (also when the caller captures)
    int rc = forward_impl(n, feats, B, flow2_out, st);
    if (rc != OFS_OK) return rc;
    return flow_resize_warp_impl(frames, n->f2s, out, B, H, W, 382, 510, st, 1);
  }
  OFS_CUDA(cudaSetDevice(n->device));
  ++n->graph_clock;
  const bool aligned = (((uintptr_t)out) % 16) == 0;   // selects the warp kernel variant at capture time
  const bool faligned = (((uintptr_t)feats) % 16) == 0;
  for (auto& g : n->graphs) {
    if (g.B != B || g.H != H || g.W != W || g.out_aligned16 != aligned || g.feats_aligned16 != faligned ||
        (g.baked[3] == nullptr) != (flow2_out == nullptr) ||
        g.packed16 != (n->feats_packed16 != 0)) continue;
    g.last_use = n->graph_clock;
    bool same = true;
    for (int w = 0; w < 4; ++w) same = same && g.baked[w] == ptrs[w];
    if (!same) {
      for (auto& pt : g.patches) {
        for (const auto& loc : pt.locs) memcpy(pt.argbuf[loc.arg].data() + loc.off, &ptrs[loc.which], 8);
        OFS_CUDA(cudaGraphExecKernelNodeSetParams(g.exec, pt.node, &pt.params));
      }
      for (int w = 0; w < 4; ++w) g.baked[w] = ptrs[w];
      ++n->graph_updates;
    }
    OFS_CUDA(cudaGraphLaunch(g.exec, st));
    count_launch(g.launches);
    return OFS_OK;
  }
  int rc = prepare(n, B);   // plans, TMA descriptors and workspace exist before the capture starts
  if (rc != OFS_OK) return rc;
  const uint64_t l0 = launch_count();
  // captured on the net's own stream (the caller's may be the legacy default stream, which cannot capture);
  // the instantiated graph is launched on the caller's stream
  cudaStream_t cs = n->stream;
  OFS_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  rc = forward_impl(n, feats, B, flow2_out, cs);
  if (rc == OFS_OK) rc = flow_resize_warp_impl(frames, n->f2s, out, B, H, W, 382, 510, cs, 1);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  const int launches = (int)(launch_count() - l0);
  count_launch(-launches);   // captured, not executed
  if (rc != OFS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  OFS_CUDA(ce);
  ofs_net::StepGraph sg{};
  for (int w = 0; w < 4; ++w) sg.baked[w] = ptrs[w];
  sg.packed16 = n->feats_packed16 != 0;
  sg.B = B; sg.H = H; sg.W = W; sg.out_aligned16 = aligned; sg.feats_aligned16 = faligned; sg.graph = graph; sg.launches = launches; sg.last_use = n->graph_clock;
  rc = collect_patches(graph, ptrs, sg.patches);
  if (rc != OFS_OK) { cudaGraphDestroy(graph); return rc; }
  const cudaError_t ie = cudaGraphInstantiate(&sg.exec, graph, 0);
  if (ie != cudaSuccess) cudaGraphDestroy(graph);
  OFS_CUDA(ie);
  ++n->graph_captures;
  if (n->graphs.size() >= 16) {   // evict the least recently used shape
    size_t lru = 0;
    for (size_t i = 1; i < n->graphs.size(); ++i) if (n->graphs[i].last_use < n->graphs[lru].last_use) lru = i;
    cudaGraphExecDestroy(n->graphs[lru].exec);
    cudaGraphDestroy(n->graphs[lru].graph);
    n->graphs.erase(n->graphs.begin() + lru);
  }
  n->graphs.push_back(std::move(sg));
  // (the moved Patch vectors keep their heap buffers, but argptr / kernelParams were taken before the move of the
  // enclosing vector only -- element addresses are unchanged by moving a std::vector, so they stay valid)
  OFS_CUDA(cudaGraphLaunch(n->graphs.back().exec, st));
  count_launch(launches);
  return OFS_OK;
}

long long ofs_net_graph_stats(const ofs_net* n, int what) {
  if (!n) return -1;
  return what == 0 ? (long long)n->graph_captures : what == 1 ? (long long)n->graph_updates : (long long)n->graphs.size();
}

int ofs_net_stabilize_host(ofs_net* n, const float* feats_host, const float* frames_host, float* out_host, int B,
                           int H, int W) {
  OFS_REQUIRE(n && feats_host && frames_host && out_host, "ofs_net_stabilize_host: null pointer");
  OFS_REQUIRE(B >= 1 && B <= n->max_batch && H > 0 && W > 0, "ofs_net_stabilize_host: bad shape");
  OFS_CUDA(cudaSetDevice(n->device));
  const size_t frame_elems = (size_t)H * W * 3, feat_elems = (size_t)kNetH * kNetW * kNetC;
  if ((size_t)B * frame_elems * 4 > n->st_frames_cap) {  // grow-only staging, sized on first use of a frame size
    OFS_CUDA(cudaDeviceSynchronize());
    if (n->st_frames) cudaFree(n->st_frames);
    if (n->st_out) cudaFree(n->st_out);
    n->st_frames = n->st_out = nullptr;
    n->st_frames_cap = 0;
    const size_t cap = (size_t)n->max_batch * frame_elems * 4;
    OFS_CUDA(cudaMalloc((void**)&n->st_frames, cap));
    OFS_CUDA(cudaMalloc((void**)&n->st_out, cap));
    n->st_frames_cap = cap;
  }
  // The call is PCIe-bound (21 MB of feats + 11 MB of frame in, 11 MB out per 720p pair), so it is
  // software-pipelined over sub-batches: H2D of chunk k+1, compute of chunk k and D2H of chunk k-1 run on
  // three streams.  Sub-batching does not change results (fixed split-K factors: batch-invariant).
  int chunk = B >= 4 ? 2 : 1;
  if (const char* e = getenv("OFS_HOST_CHUNK")) chunk = std::max(1, std::min(B, atoi(e)));   // experiments only
  const int nchunks = (B + chunk - 1) / chunk;
  // bf16 nets: the network input crosses the bus as bf16, rounded here on the host exactly as pack_act_kernel would round
  // it on the device (same bits, half the bytes); sub-batch c+1 is converted while sub-batch c is in flight
  const bool packed = n->is_bf16 && n->host_pack;
  if (packed && !n->packer) {
    const unsigned hw = std::thread::hardware_concurrency();
    int threads = (int)std::max(1u, std::min(6u, hw / 3));
    if (const char* e = getenv("OFS_HOST_PACK_THREADS")) threads = std::max(1, std::min(64, atoi(e)));
    n->packer = host_packer_create(threads);
    OFS_CUDA(cudaHostAlloc((void**)&n->h_feats16, (size_t)n->max_batch * feat_elems * 2, cudaHostAllocDefault));
  }
  if (packed) host_packer_start(n->packer, feats_host, n->h_feats16, (size_t)std::min(chunk, B) * feat_elems);
  for (int c = 0; c < nchunks; ++c) {
    const int b0 = c * chunk, nb = std::min(chunk, B - b0);
    float* d_feats = n->st_feats + (size_t)b0 * feat_elems;
    float* d_frames = n->st_frames + (size_t)b0 * frame_elems;
    float* d_out = n->st_out + (size_t)b0 * frame_elems;
    if (packed) {
      host_packer_wait(n->packer);                        // sub-batch c is converted
      if (c + 1 < nchunks) {
        const int b1 = (c + 1) * chunk, nb1 = std::min(chunk, B - b1);
        host_packer_start(n->packer, feats_host + (size_t)b1 * feat_elems, n->h_feats16 + (size_t)b1 * feat_elems, (size_t)nb1 * feat_elems);
      }
      uint16_t* d16 = reinterpret_cast<uint16_t*>(n->st_feats) + (size_t)b0 * feat_elems;   // the float32 staging, reused
      d_feats = reinterpret_cast<float*>(d16);
      OFS_CUDA(cudaMemcpyAsync(d16, n->h_feats16 + (size_t)b0 * feat_elems, (size_t)nb * feat_elems * 2, cudaMemcpyHostToDevice, n->s_h2d));
    } else
    OFS_CUDA(cudaMemcpyAsync(d_feats, feats_host + (size_t)b0 * feat_elems, (size_t)nb * feat_elems * 4,
                             cudaMemcpyHostToDevice, n->s_h2d));
    OFS_CUDA(cudaMemcpyAsync(d_frames, frames_host + (size_t)b0 * frame_elems, (size_t)nb * frame_elems * 4,
                             cudaMemcpyHostToDevice, n->s_h2d));
    OFS_CUDA(cudaEventRecord(n->ev_h2d[c], n->s_h2d));
    OFS_CUDA(cudaStreamWaitEvent(n->stream, n->ev_h2d[c], 0));
    n->feats_packed16 = packed ? 1 : 0;
    int rc = ofs_net_stabilize(n, d_feats, d_frames, d_out, nullptr, nb, H, W, (ofs_stream)n->stream);
    n->feats_packed16 = 0;
    if (rc != OFS_OK) { if (packed) host_packer_wait(n->packer); return rc; }
    OFS_CUDA(cudaEventRecord(n->ev_comp[c], n->stream));
    OFS_CUDA(cudaStreamWaitEvent(n->s_d2h, n->ev_comp[c], 0));
    OFS_CUDA(cudaMemcpyAsync(out_host + (size_t)b0 * frame_elems, d_out, (size_t)nb * frame_elems * 4,
                             cudaMemcpyDeviceToHost, n->s_d2h));
  }
  OFS_CUDA(cudaStreamSynchronize(n->s_d2h));
  return OFS_OK;
}

// bytes the host call above moves host -> device for a batch (bench.py reports them as e2e.h2d_bytes_per_step)
long long ofs_net_host_h2d_bytes(const ofs_net* n, int B, int H, int W) {
  if (!n) return -1;
  const long long feats = (long long)B * kNetH * kNetW * kNetC * ((n->is_bf16 && n->host_pack) ? 2 : 4);
  return feats + (long long)B * H * W * 3 * 4;
}

int ofs_net_get_activation(ofs_net* n, const char* name, int B, float* out, int64_t capacity, int* shape4,
                           ofs_stream stream) {
  OFS_REQUIRE(n && name && out, "ofs_net_get_activation: null pointer");
  auto it = n->acts.find(name);
  OFS_REQUIRE(it != n->acts.end(), "ofs_net_get_activation: unknown activation '%s'", name);
  const ActInfo& a = it->second;
  OFS_REQUIRE(B >= 1 && B <= n->max_batch, "ofs_net_get_activation: bad batch");
  const int64_t need = (int64_t)B * a.H * a.W * a.C;
  OFS_REQUIRE(capacity >= need, "ofs_net_get_activation: capacity %lld < %lld", (long long)capacity, (long long)need);
  if (shape4) { shape4[0] = B; shape4[1] = a.H; shape4[2] = a.W; shape4[3] = a.C; }
  return launch_unpack_act(a.ptr, out, (size_t)B * a.H * a.W, a.cs, a.coff, a.C, n->is_bf16, (cudaStream_t)stream);
}

int ofs_net_profile(ofs_net* n, const float* feats, const float* frames, float* out, int B, int H, int W, int iters,
                    float* ms, double* macs, char* names, int cap, int* count, ofs_stream stream) {
  OFS_REQUIRE(n && feats && ms && macs && names && count && iters >= 1, "ofs_net_profile: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  Marks marks;
  std::vector<double> acc;
  int rc = OFS_OK;
  for (int it = 0; it < iters && rc == OFS_OK; ++it) {
    marks.used = 0;
    marks.names.clear();
    marks.macs.clear();
    rc = forward_impl(n, feats, B, nullptr, st, &marks);
    if (rc == OFS_OK && frames && out) {
      rc = flow_resize_warp_impl(frames, n->f2s, out, B, H, W, 382, 510, st, 1);
      if (rc == OFS_OK) rc = marks.mark(st, "flow_resize_warp", 0.0);
    }
    if (rc != OFS_OK) break;
    rc = check_cuda(cudaStreamSynchronize(st), "profile sync", __FILE__, __LINE__);
    if (rc != OFS_OK) break;
    if (acc.empty()) acc.assign(marks.used, 0.0);
    for (size_t i = 1; i < marks.used; ++i) {
      float t = 0;
      cudaEventElapsedTime(&t, marks.ev[i - 1], marks.ev[i]);
      acc[i] += t;
    }
  }
  const int nmarks = (int)marks.used - 1;
  if (rc == OFS_OK && nmarks > cap) { set_error("ofs_net_profile: capacity %d < %d", cap, nmarks); rc = OFS_EINVAL; }
  if (rc == OFS_OK) {
    for (int i = 0; i < nmarks; ++i) {
      ms[i] = (float)(acc[i + 1] / iters);
      macs[i] = marks.macs[i + 1];
      snprintf(names + (size_t)i * 32, 32, "%s", marks.names[i + 1].c_str());
    }
    *count = nmarks;
  }
  for (cudaEvent_t e : marks.ev) cudaEventDestroy(e);
  return rc;
}

int ofs_net_time_kernels(ofs_net* n, int which, const float* frames, float* out, int B, int H, int W, int iters,
                         float* ms_per_set, double* macs_per_set, int* launches_per_set) {
  OFS_REQUIRE(n && ms_per_set && iters >= 1 && (which == 0 || which == 1), "ofs_net_time_kernels: bad arguments");
  OFS_REQUIRE(B >= 1 && B <= n->max_batch, "ofs_net_time_kernels: batch %d outside [1, %d]", B, n->max_batch);
  OFS_REQUIRE(which == 0 || (frames && out && H > 0 && W > 0), "ofs_net_time_kernels: the warp needs frames / out");
  if (!n->loaded) { set_error("ofs_net_time_kernels: weights not loaded"); return OFS_ESTATE; }
  OFS_CUDA(cudaSetDevice(n->device));
  int rc = prepare(n, B);
  if (rc != OFS_OK) return rc;
  cudaStream_t cs = n->stream;
  OFS_CUDA(cudaStreamSynchronize(cs));
  double macs = 0.0;
  const uint64_t l0 = launch_count();
  OFS_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  for (int it = 0; it < iters && rc == OFS_OK; ++it) {
    if (which == 0) {
      for (size_t li = 0; li < n->layers.size(); ++li) {
        if (n->layers[li].name == "predict2") continue;
        int covered = 1;
        rc = launch_layer(n, li, cs, &covered);
        if (rc != OFS_OK) break;
        for (int k = 0; k < covered; ++k) if (it == 0) macs += n->layers[li + k].plan.macs;
        li += (size_t)covered - 1;
      }
    } else {
      rc = flow_resize_warp_impl(frames, n->f2s, out, B, H, W, 382, 510, cs, 1);
    }
  }
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  const int launches = (int)(launch_count() - l0);
  count_launch(-launches);
  if (rc != OFS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  OFS_CUDA(ce);
  cudaGraphExec_t exec = nullptr;
  cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  OFS_CUDA(ie);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(exec, cs);           // warm-up replay
  cudaEventRecord(e0, cs);
  cudaGraphLaunch(exec, cs);
  cudaEventRecord(e1, cs);
  rc = check_cuda(cudaStreamSynchronize(cs), "time_kernels sync", __FILE__, __LINE__);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaGraphExecDestroy(exec);
  count_launch(2 * launches);
  if (rc != OFS_OK) return rc;
  *ms_per_set = ms / (float)iters;
  if (macs_per_set) *macs_per_set = macs;
  if (launches_per_set) *launches_per_set = launches / iters;
  return OFS_OK;
}

int ofs_chain_trace_read(long long* host_words, int max_words) {
  if (!host_words || max_words <= 0) return 0;
  return conv_chain_trace_read(host_words, max_words);
}

int ofs_net_launches_per_forward(const ofs_net* n) {
  if (!n) return 0;
  int k = 1 + (int)n->layers.size() + 4 + 1;  // pack + GEMMs (heads ride in the deconvs) + pyramid steps + gather
  if (n->p2_fused) --k;                       // predict2: product and gather are one kernel
  for (const Layer& L : n->layers) k += (L.plan.p.ksplit > 1 && !L.plan.p.fused_reduce) ? 1 : 0;  // separate split-K reductions (after prepare())
  for (size_t i = 0; i + 3 < n->layers.size(); ++i)
    if (n->chain && n->chain_sync && n->layers[i].name == "5") {
      const ConvPlan* plans[4] = {&n->layers[i].plan, &n->layers[i + 1].plan, &n->layers[i + 2].plan, &n->layers[i + 3].plan};
      if (conv_chain_supported(plans)) k -= 7;   // four GEMMs + four reductions are one launch
    }
  return k;
}

}  // extern "C"
