// Implicit-GEMM convolution / transposed convolution on tcgen05 (sm_100a): plan + launch API
// shared by the network (flownet.cu) and the stand-alone op (ofs_conv2d_nhwc).
#pragma once

#include <vector>

#include "ofs_common.cuh"

namespace ofs {

constexpr int kMaxTapEntries = 64;

// Everything the kernel needs, passed as one __grid_constant__ parameter.
struct ConvGemmParams {
  CUtensorMap tmap_a;  // 5-D view of the NHWC 16-bit activation (see conv_gemm.cu)
  CUtensorMap tmap_w;  // 2-D packed weights [phases * n_pad rows][K_total], K-major (box rows: BLOCK_N / cta_group)
  CUtensorMap tmap_o[4];  // per phase: 4-D view [c, x, y, b] of this layer's 16-bit output slice (TMA-store epilogue)
  CUtensorMap tmap_w_half;  // tail split: the weights with a box of 128 / cta_group rows (half a 256-column tile)
  int tail_t0;         // tail split (256-column layers with one N tile): scheduling units >= tail_t0 are HALF tiles --
                       // unit tail_t0 + 2j + h computes columns [128h, 128h+128) of M tile tail_t0 + j.  A last wave
                       // that would leave most SMs idle is spread over twice as many CTAs at ~0.64 of the tile time;
                       // N splits do not touch the K summation order, results stay bit-identical.  INT_MAX: off.
  int tma_store;       // 1: 16-bit epilogue goes through swizzled shared-memory staging + cp.async.bulk.tensor stores
  // M grid (output pixels; for the transposed conv: input pixels, one GEMM per sub-pixel phase)
  int Hg, Wg;          // grid height / width per image
  int rows_total;      // B * Hg
  int tileW_log2;      // tile = tile_rows "global rows" (b,y) x tileW pixels <= 128 GEMM rows
  int tile_rows;       // valid global rows per tile
  int npieces;         // TMA loads per A stage
  int piece_rows;      // global rows per piece = box_y * box_b (a piece never straddles an image boundary
                       // unless it is made of whole images)
  int box_y, box_b;    // TMA box extent along y and batch: {rpl, 1}, or {Hg, images per tile} for small grids
  int a_bytes;         // bytes one A stage receives = npieces * piece_rows * tileW * 128
  int tiles_x;         // Wg / tileW
  int tiles_m;         // tiles_x * ceil(rows_total / tile_rows)
  int tiles_mp;        // M tiles per scheduling unit: tiles_m (1 CTA) or ceil(tiles_m / 2) (CTA pairs)
  int tiles_n;         // n_pad / BLOCK_N
  int phases;          // 1 (conv) or 4 (transposed conv sub-pixel phases)
  int ksplit;          // split-K factor (partials go to an fp32 workspace, reduced by splitk_reduce_kernel)
  int kb_per_split;    // K blocks per split
  int kcluster;        // 1: the ksplit CTAs of a tile form one cluster and reduce their partials through DSMEM
  long long ws_split_stride;  // workspace elements between two splits = out pixels * n_pad
  // split-K reduced INSIDE the GEMM launch (no splitk_reduce_kernel): every (tile, split) unit is one resident CTA; after
  // its fp32 partial has landed in the workspace a CTA bumps the tile's arrival counter, waits until all ksplit partials
  // of the tile are there, and sums ITS share of the tile's rows over the splits in split order (bias first: the very
  // order, and therefore the very bits, of splitk_reduce_kernel) into the final 16-bit slice.
  int fused_reduce;           // 1: on (needs total units <= grid so that all of them are co-resident)
  unsigned* sk_counters;      // [3 * tiles]: arrivals, row claims, departures (the last CTA to leave zeroes all three: self-cleaning)
  void* final_out;            // the 16-bit destination of the fused reduction
  int ntaps, nchunks;  // K_total = ntaps * nchunks * 64
  int n_pad;           // padded output channels per phase
  int w_rows_phase;    // packed weight rows per phase = n_pad (+ 16 with a fused head)
  float2* head_out;    // fused head: this phase's share per OUTPUT pixel [B, out_H, out_W] (null: no head)
  long long head_split_stride;  // split-K with the fused head: K split s writes its share of the head into plane s,
                                // head_out + s * head_split_stride (= B * out_H * out_W); the consumer sums the planes in order
  // epilogue
  void* out;           // 16-bit activations (mode 0) or fp32 (mode 1)
  const float* bias;   // [n_pad]
  int out_mode;        // 0: act(v + b) -> 16-bit;  1: v + b -> fp32 (first n_valid columns);
                       // 2: raw fp32 partial sums -> split-K workspace [ks][pixel][n_pad]
  int out_H, out_W;    // full output image size
  int out_cstride;     // channels per output pixel in the destination buffer
  int out_coff;        // channel offset of this layer's slice in the destination (concat-by-slice)
  int out_scale;       // 1, or 2 for the transposed conv
  int n_valid;         // real output channels
  int lrelu;           // apply max(v, 0.1 v)
  int is_bf16;         // operand / storage format: 1 bf16, 0 fp16
  int debug;           // measurement only: bit0 skip MMAs, bit1 skip A loads, bit2 skip B loads (results garbage)
  long long* trace;    // measurement only: per-CTA timestamps [grid][16] (null in production)
  int out_oy[4], out_ox[4];  // per-phase sub-pixel offset
  // per (phase, tap) TMA coordinate offsets: channel base, x offset, parity plane, y offset
  short tap_c[kMaxTapEntries], tap_x[kMaxTapEntries], tap_p[kMaxTapEntries], tap_y[kMaxTapEntries];
  // slab mode (slab = 1): a table entry is a GROUP of up to 4 taps that differ only by an x shift.  One TMA
  // box of tileW + slab_extra pixels (the "slab") is loaded per group and tap t reads it through a UMMA
  // descriptor whose start address is advanced by grp_off[t] 128-byte rows: one pipeline stage, one
  // handshake and one A fetch per group instead of per tap.
  int slab, slab_extra;
  short grp_n[kMaxTapEntries];          // 16-bit like the tap tables: byte-wide constant loads are not uniform loads
  short grp_off[kMaxTapEntries][4];
};

enum ConvKind { kConv = 0, kDeconvK4S2 = 1 };

struct ConvDesc {
  ConvKind kind;
  int B, H, W;         // input image grid
  int cin;             // logical input channels (K covers ceil(cin/64)*64, extra weights are zero)
  int in_cs;           // channel stride of the input buffer (elements), multiple of 8
  int cout;            // logical output channels
  int k, stride;       // conv: k odd, pad k/2, stride 1|2;  deconv: k=4, stride=2
  int block_n;         // 16, 32, 64, 128 or 256
  int out_mode, lrelu, is_bf16;
  int out_cstride, out_coff;
  int ksplit = 1;      // > 1: split the K loop over this many CTAs per tile (16-bit output mode only)
  int tail_half = 1;   // 0: never split the tail wave into half tiles (see ConvGemmParams::tail_t0)
  int kcluster = 0;    // 1: split-K inside a thread-block cluster (ksplit <= 8, block_n 256, 1-CTA tiles): no workspace, no reduce kernel
  int cta_group = 1;   // 2: CTA pairs (tcgen05 cta_group::2): tile = 256 GEMM rows x BLOCK_N, B split over the pair
  int kgroup = 1;      // 2: two consecutive 64-channel K blocks of a tap per pipeline stage (narrow-N layers)
  int head = 0;        // 1: transposed conv with the level's 3x3 flow head fused as 16 extra accumulator columns
  int slab = 0;        // 1: x-shifted taps share one shared-memory slab (stride-2 convs whose tiles are one 128-px row)
  int stack = 0;       // 1: transposed conv (cout 64, fused head) with all 4 sub-pixel phases stacked in ONE accumulator tile: each
                       //    of the 9 distinct input taps is fetched once per chunk (deconv_stack_kernel); cta_group 1 or 2
  int debug = 0;       // see ConvGemmParams::debug
  long long* trace = nullptr;  // see ConvGemmParams::trace
};

struct ConvPlan {
  ConvGemmParams p;
  ConvDesc d;
  int block_n = 0;
  int grid = 0;
  size_t smem = 0;
  int k_total = 0;
  int w_rows = 0;            // phases * n_pad
  bool paired = false;       // conv1-style stride-2 layer with in_cs == 32: two x-taps per K chunk
  int group_max = 1;         // slab mode: most taps in one group (sizes the pipeline stage)
  // K order of the packed weights: K block i holds weight tap (wt_ky[i], wt_kx[i]) (paired form: kx and kx + 1
  // of the x-parity pair starting at wt_kx[i], which may be -1)
  std::vector<int> wt_ky, wt_kx;
  double macs = 0;           // literal MACs of the layer (roofline numerator)
  size_t ws_bytes = 0;       // split-K workspace this plan needs (0 when ksplit == 1)
  bool fused_reduce_ok = false;  // geometry allows the in-kernel reduction (one resident CTA per (tile, split))
  int n_counters = 0;        // counters the fused reduction needs (3 per tile)
  // split-K reduction (filled by bind)
  void* final_out = nullptr;
  const float* bias_dev = nullptr;
  float* ws = nullptr;
};

// Geometry only (no device pointers): tap table, tiling, grid.  OFS_EINVAL on unsupported shapes.
int conv_plan_geometry(ConvPlan& plan, const ConvDesc& d);
// Packs float32 weights (TF layout: conv [k,k,cin,cout]; deconv [4,4,cout,cin]) into the 16-bit
// K-major GEMM layout of `plan` (w_rows x k_total) and the padded bias [n_pad].
void conv_pack_weights(const ConvPlan& plan, const float* w_tf, const float* bias, std::vector<uint16_t>& w_packed,
                       std::vector<float>& b_padded, const float* head_w = nullptr /* [3,3,cin,2] when plan.d.head */);
// Binds device pointers and encodes the TMA descriptors.
// counters: zero-initialised device array of plan.n_counters unsigned (null: split-K falls back to splitk_reduce_kernel)
int conv_plan_bind(ConvPlan& plan, const void* act_in, const void* w_packed_dev, const float* bias_dev, void* out,
                   float* workspace = nullptr, float* head_out = nullptr, unsigned* counters = nullptr);
int conv_launch(const ConvPlan& plan, cudaStream_t st);

// conv5 -> conv5_1 -> conv6 -> conv6_1 as ONE cooperative persistent launch (split-K GEMM phases and all-CTA reductions
// separated by grid barriers; bit-identical to four conv_launch calls).  `plans`: the four bound split-K plans in layer
// order; `sync`: two zero-initialised device words owned by the caller (self-maintained afterwards).
bool conv_chain_supported(const ConvPlan* const plans[4]);
int conv_chain_launch(const ConvPlan* const plans[4], unsigned* sync, cudaStream_t st);
// measurement only (OFS_CHAIN_TRACE=1): the last chain launch's per-CTA globaltimer stamps, 32 words per CTA:
// [4l + {0 GEMM done, 1 barrier, 2 reduced, 3 barrier}] for layer l, [16] kernel entry; returns the words copied
int conv_chain_trace_read(long long* host, int max_words);

// 5-D TMA view of the input activation (dims in elements, strides in bytes; dim 0 is contiguous)
void conv_act_view(const ConvDesc& d, unsigned long long dims[5], unsigned long long strides_bytes[4]);

uint16_t f32_to_bf16_rn(float f);
uint16_t f32_to_fp16_rn(float f);

// fp32 NHWC [npix, cin] -> 16-bit [npix, cs] (channels >= cin zero-filled)
int launch_pack_act(const float* in, void* out, size_t npix, int cin, int cs, int is_bf16, cudaStream_t st);
// 16-bit [npix, cs] channels [coff, coff+c) -> fp32 [npix, c]
int launch_unpack_act(const void* in, float* out, size_t npix, int cs, int coff, int c, int is_bf16, cudaStream_t st);

}  // namespace ofs
