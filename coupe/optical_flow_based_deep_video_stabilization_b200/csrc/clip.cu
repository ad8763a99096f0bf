// ofs_clips: the per-frame loop of evaluate_originalSize() (reference main_dl.py:535-630) with its state on the device.
//
// The reference keeps a float64 history of ALL output frames on the host and, per frame, runs nine cv2.resize +
// cvtColor calls on full-resolution frames, feeds 21 MB + 11 MB of float32 through sess.run and fetches 11 MB back
// (main_dl.py:550-569).  Only np.uint8() versions of the history are ever consumed (main_dl.py:556-558, :630), and a
// frame is re-used by up to 8 later steps (offsets 31,23,15,7,4,3,2,1).  So the device keeps, per clip, a 32-slot
// ring of the RESIZED (512x384), channel-swapped uint8 outputs; a step uploads one uint8 frame (2.8 MB at 720p
// instead of 32 MB of float32), resizes it once, assembles the 27-channel network input from the ring, runs forward +
// flow glue, warps the uint8 frame itself into the np.uint8 output frame (ofs::flow_resize_warp_u8_impl), resizes that
// into the ring and downloads it.  n clips advance in lockstep as one batch; up to 3 steps are in flight.
//
// Arithmetic restated exactly (the GPU test replays main_dl.py:540-630 with cv2 on the host and compares bytes):
//   * cv2.resize(u8, (512,384)) INTER_LINEAR: OpenCV's fixed-point path -- coefficients cvRound(f * 2048) from
//     f = (float)((d + 0.5) * scale - 0.5), horizontal int pass, vertical ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2;
//   * history taps: np.float32(u8) / 255.0 (float32 division); current frame: u8 / 255.0 (float64 division, cast to
//     float32 by the feed); resizedInput: the same per pixel at full resolution;
//   * totaloutputFrame[i] = cvtColor(warped * 255) in float32; np.uint8(): truncation toward zero, low 8 bits.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <initializer_list>
#include <vector>

#include "conv_gemm.cuh"

struct ofs_net;
namespace ofs {
void* net_x0(ofs_net* n);
int net_max_batch(const ofs_net* n);
int net_is_bf16(const ofs_net* n);
int net_device(const ofs_net* n);
bool net_loaded(const ofs_net* n);
unsigned net_weights_generation(const ofs_net* n);
int net_prepare(ofs_net* n, int B);
int net_stabilize_u8_from_x0(ofs_net* n, const uint8_t* frames, uint8_t* out_u8, float* out_f32, int B, int H, int W,
                             cudaStream_t st);
}  // namespace ofs

namespace {

using namespace ofs;

constexpr int kNetH = 384, kNetW = 512, kRing = 32;
constexpr int kSlice = kNetH * kNetW * 3;   // bytes of one resized uint8 slice
constexpr int kOffsets[8] = {31, 23, 15, 7, 4, 3, 2, 1};   // main_dl.py:553

struct StepState {        // read by the kernels from device memory so the captured graph never changes
  int hist_slot[8];       // ring slot of history tap j
  int write_slot;         // ring slot of this step's output
  int first;              // step 0: the resized input frame also seeds ring slot 0 (main_dl.py:548-549)
  int pad[6];
};

// cv2.resize(src, (512, 384)), INTER_LINEAR, uint8, 3 channels; dst channels reversed (cvtColor RGB2BGR == swap)
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, int H, int W,
                                                        uint8_t* __restrict__ dst, size_t dst_clip_stride,
                                                        const StepState* __restrict__ state, int dst_is_ring,
                                                        uint8_t* __restrict__ seed_ring, size_t ring_clip_stride,
                                                        const int* __restrict__ sx_tab, const short* __restrict__ ax_tab,
                                                        const int* __restrict__ sy_tab, const short* __restrict__ ay_tab) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  const int clip = blockIdx.z;
  if (dx >= kNetW) return;
  const int sx = sx_tab[dx], sx1 = min(sx + 1, W - 1);
  const int a0 = ax_tab[2 * dx], a1 = ax_tab[2 * dx + 1];
  const int sy = sy_tab[dy];
  const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
  const int b0 = ay_tab[2 * dy], b1 = ay_tab[2 * dy + 1];
  const uint8_t* s = src + (size_t)clip * H * W * 3;
  const uint8_t* r0 = s + (size_t)y0 * W * 3;
  const uint8_t* r1 = s + (size_t)y1 * W * 3;
  uint8_t o[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
    const int h1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
    o[2 - c] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
  }
  const size_t px = ((size_t)dy * kNetW + dx) * 3;
  uint8_t* d = dst + (size_t)clip * dst_clip_stride + (dst_is_ring ? (size_t)state->write_slot * kSlice : 0) + px;
  d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
  if (seed_ring && state->first) {
    uint8_t* r = seed_ring + (size_t)clip * ring_clip_stride + px;   // slot 0
    r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
  }
}

// byte / 255 in float32.  The reference divides the history taps in float32 (np.float32(u8) / 255.0, main_dl.py:556-558)
// and the current frame in float64 before the feed casts it to float32 (:550, :568); the two agree for every byte
// value, and both equal q' = fma(fma(-q, 255, v), RN(1/255), q) with q = v * RN(1/255) (checked for all 256 values by
// ofs_clips_create against the host quotients, so a compiler that contracts differently cannot go unnoticed).
__device__ __forceinline__ float byte_over_255(uint32_t v) {
  const float c = 0.00392156885936856270f;   // RN(1/255)
  const float x = (float)v;
  const float q = x * c;
  const float r = __fmaf_rn(-q, 255.0f, x);
  return __fmaf_rn(r, c, q);
}

__global__ void byte_over_255_table_kernel(float* out) { out[threadIdx.x] = byte_over_255(threadIdx.x); }

template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (kBf16) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// curinput (main_dl.py:550-558) straight into the network's packed 16-bit input: 8 history taps + current frame.
// One pixel per thread: 27 byte loads (a warp reads 96 contiguous bytes per load), the quotients in registers (a
// 256-entry shared-memory table cost more in bank conflicts than the 3 FMAs it saved), four 16-byte stores.
template <bool kBf16>
__global__ void __launch_bounds__(256) assemble_x0_kernel(const uint8_t* __restrict__ ring, uint32_t ring_clip_stride,
                                                          const uint8_t* __restrict__ cur, const StepState* __restrict__ state,
                                                          uint4* __restrict__ x0) {
  const int clip = blockIdx.y;
  const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= (uint32_t)(kNetH * kNetW)) return;
  const uint8_t* rb = ring + (size_t)clip * ring_clip_stride + px * 3u;
  float v[28];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint8_t* s = rb + (uint32_t)__ldg(&state->hist_slot[j]) * (uint32_t)kSlice;
    v[3 * j] = byte_over_255(__ldg(s)); v[3 * j + 1] = byte_over_255(__ldg(s + 1)); v[3 * j + 2] = byte_over_255(__ldg(s + 2));
  }
  const uint8_t* c = cur + (size_t)clip * kSlice + px * 3u;
  v[24] = byte_over_255(__ldg(c)); v[25] = byte_over_255(__ldg(c + 1)); v[26] = byte_over_255(__ldg(c + 2));
  v[27] = 0.0f;
  uint4* o = x0 + ((size_t)clip * kNetH * kNetW + px) * 4;
#pragma unroll
  for (int q = 0; q < 3; ++q)
    o[q] = make_uint4(pack2<kBf16>(v[8 * q], v[8 * q + 1]), pack2<kBf16>(v[8 * q + 2], v[8 * q + 3]),
                      pack2<kBf16>(v[8 * q + 4], v[8 * q + 5]), pack2<kBf16>(v[8 * q + 6], v[8 * q + 7]));
  o[3] = make_uint4(pack2<kBf16>(v[24], v[25]), pack2<kBf16>(v[26], v[27]), 0u, 0u);
}

// OpenCV resize.cpp: fx = (float)((dx+0.5)*scale_x - 0.5); sx = cvFloor(fx); fx -= sx; edge clamps; cvRound(f * 2048)
void linear_tables(int src, int dst, std::vector<int>& s_tab, std::vector<short>& a_tab, bool clamp_x) {
  s_tab.resize(dst);
  a_tab.resize(2 * dst);
  const double scale = (double)src / (double)dst;
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_x) {   // horizontal pass only; the vertical pass clips the ROW indices instead
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    s_tab[d] = s;
    auto sat = [](float v) { long r = lrintf(v); return (short)std::max<long>(-32768, std::min<long>(32767, r)); };
    a_tab[2 * d] = sat((1.f - f) * 2048.f);
    a_tab[2 * d + 1] = sat(f * 2048.f);
  }
}

}  // namespace

// One step in flight owns one Slot: its own input / output staging and its own captured graph (a graph's kernel
// arguments are baked in), so the upload of step i+1 and the download of step i-1 overlap the kernels of step i.
struct ClipSlot {
  uint8_t *d_frame = nullptr, *d_out_u8 = nullptr;
  float* d_out_f32 = nullptr;
  StepState* d_state = nullptr;
  StepState* h_state = nullptr;       // pinned
  cudaGraphExec_t graph = nullptr;
  bool graph_f32 = false;
  int graph_launches = 0;
  unsigned graph_generation = 0;      // weights generation the graph was captured with
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
  bool busy = false;
};

// 3 slots: with 2, upload(i) could not start before download(i-2) had landed (the host only submits step i after
// waiting for step i-2), a chain of download + upload + kernels per two steps that left the GPU idle a third of the time
constexpr int kSlots = 3;

struct ofs_clips {
  ofs_net* net = nullptr;
  int n = 0, H = 0, W = 0, device = 0;
  long long frame = 0;                // steps submitted
  long long collected = 0;            // steps waited for
  ClipSlot slot[kSlots];
  uint8_t *d_ring = nullptr, *d_cur = nullptr;
  float* d_quot = nullptr;
  int *d_sx = nullptr, *d_sy = nullptr;
  short *d_ax = nullptr, *d_ay = nullptr;
  cudaStream_t st = nullptr, st_in = nullptr, st_out = nullptr;
  std::vector<void*> allocs;
};

namespace {

int clips_alloc(ofs_clips* c, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) { set_error("ofs_clips: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return OFS_ENOMEM; }
  c->allocs.push_back(*p);
  return check_cuda(cudaMemset(*p, 0, bytes), "memset", __FILE__, __LINE__);
}

int enqueue_step(ofs_clips* c, const ClipSlot& s, cudaStream_t st, bool want_f32) {
  const int n = c->n, H = c->H, W = c->W;
  const size_t ring_stride = (size_t)kRing * kSlice;
  const dim3 rgrid((kNetW + 255) / 256, kNetH, n);
  resize_u8_kernel<<<rgrid, 256, 0, st>>>(s.d_frame, H, W, c->d_cur, (size_t)kSlice, s.d_state, 0, c->d_ring, ring_stride,
                                          c->d_sx, c->d_ax, c->d_sy, c->d_ay);
  OFS_LAUNCH_CHECK();
  {
    const dim3 agrid((kNetH * kNetW + 255) / 256, n);
    uint4* x0 = reinterpret_cast<uint4*>(net_x0(c->net));
    if (net_is_bf16(c->net))
      assemble_x0_kernel<true><<<agrid, 256, 0, st>>>(c->d_ring, (uint32_t)ring_stride, c->d_cur, s.d_state, x0);
    else
      assemble_x0_kernel<false><<<agrid, 256, 0, st>>>(c->d_ring, (uint32_t)ring_stride, c->d_cur, s.d_state, x0);
  }
  OFS_LAUNCH_CHECK();
  // forward + flow glue + the warp on the uint8 frame: resizedInput (:568), sess.run (:569), * 255 and np.uint8 (:625, :630)
  int rc = net_stabilize_u8_from_x0(c->net, s.d_frame, s.d_out_u8, want_f32 ? s.d_out_f32 : nullptr, n, H, W, st);
  if (rc != OFS_OK) return rc;
  resize_u8_kernel<<<rgrid, 256, 0, st>>>(s.d_out_u8, H, W, c->d_ring, ring_stride, s.d_state, 1, nullptr, 0, c->d_sx,
                                          c->d_ax, c->d_sy, c->d_ay);
  OFS_LAUNCH_CHECK();
  return OFS_OK;
}

}  // namespace

extern "C" {

int ofs_clips_create(ofs_clips** out, ofs_net* net, int n_clips, int H, int W) {
  OFS_REQUIRE(out && net, "ofs_clips_create: null pointer");
  *out = nullptr;
  OFS_REQUIRE(n_clips >= 1 && n_clips <= net_max_batch(net), "ofs_clips_create: %d clips outside [1, max_batch %d]", n_clips,
              net_max_batch(net));
  OFS_REQUIRE(H >= 2 && W >= 2, "ofs_clips_create: bad frame size %dx%d", H, W);
  OFS_CUDA(cudaSetDevice(net_device(net)));
  ofs_clips* c = new ofs_clips();
  c->net = net; c->n = n_clips; c->H = H; c->W = W; c->device = net_device(net);
  const size_t fpx = (size_t)n_clips * H * W * 3;
  int rc = OFS_OK;
  for (ClipSlot& s : c->slot) {
    if (rc == OFS_OK) rc = clips_alloc(c, (void**)&s.d_frame, fpx);
    if (rc == OFS_OK) rc = clips_alloc(c, (void**)&s.d_out_u8, fpx);
    if (rc == OFS_OK) rc = clips_alloc(c, (void**)&s.d_out_f32, fpx * 4);
    if (rc == OFS_OK) rc = clips_alloc(c, (void**)&s.d_state, sizeof(StepState));
    if (rc == OFS_OK && cudaMallocHost((void**)&s.h_state, sizeof(StepState)) != cudaSuccess) {
      set_error("ofs_clips_create: cudaMallocHost failed");
      rc = OFS_ENOMEM;
    }
    for (cudaEvent_t* e : {&s.ev_in, &s.ev_done, &s.ev_out})
      if (rc == OFS_OK && cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) {
        set_error("ofs_clips_create: cudaEventCreate failed");
        rc = OFS_ECUDA;
      }
  }
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_ring, (size_t)n_clips * kRing * kSlice);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_cur, (size_t)n_clips * kSlice);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_quot, 1024);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_sx, kNetW * 4);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_sy, kNetH * 4);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_ax, kNetW * 4);
  if (rc == OFS_OK) rc = clips_alloc(c, (void**)&c->d_ay, kNetH * 4);
  for (cudaStream_t* s : {&c->st, &c->st_in, &c->st_out})
    if (rc == OFS_OK && cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("ofs_clips_create: cudaStreamCreate failed");
      rc = OFS_ECUDA;
    }
  if (rc == OFS_OK) {
    // the device's byte / 255 against the reference's two host quotients, all 256 values
    float got[256];
    byte_over_255_table_kernel<<<1, 256>>>(c->d_quot);
    rc = check_cuda(cudaMemcpy(got, c->d_quot, sizeof(got), cudaMemcpyDeviceToHost), "quotient check", __FILE__, __LINE__);
    for (int v = 0; v < 256 && rc == OFS_OK; ++v) {
      const float hist = (float)v / 255.0f;                 // np.float32(u8) / 255.0        (main_dl.py:556,558)
      const float cur = (float)((double)v / 255.0);         // u8 / 255.0 -> float32 at the feed (main_dl.py:550,568)
      if (got[v] != hist || got[v] != cur) {
        set_error("ofs_clips_create: device %d / 255 = %.9g, reference %.9g / %.9g", v, got[v], hist, cur);
        rc = OFS_ECUDA;
      }
    }
    std::vector<int> sx, sy;
    std::vector<short> ax, ay;
    linear_tables(W, kNetW, sx, ax, true);
    linear_tables(H, kNetH, sy, ay, false);
    const struct { void* dst; const void* src; size_t bytes; } ups[] = {
        {c->d_sx, sx.data(), (size_t)kNetW * 4}, {c->d_sy, sy.data(), (size_t)kNetH * 4},
        {c->d_ax, ax.data(), (size_t)kNetW * 4}, {c->d_ay, ay.data(), (size_t)kNetH * 4}};
    for (const auto& u : ups)
      if (rc == OFS_OK) rc = check_cuda(cudaMemcpy(u.dst, u.src, u.bytes, cudaMemcpyHostToDevice), "table upload", __FILE__, __LINE__);
  }
  if (rc != OFS_OK) { ofs_clips_destroy(c); return rc; }
  *out = c;
  return OFS_OK;
}

int ofs_clips_destroy(ofs_clips* c) {
  if (!c) return OFS_OK;
  cudaSetDevice(c->device);
  for (cudaStream_t s : {c->st_in, c->st, c->st_out})
    if (s) cudaStreamSynchronize(s);
  for (ClipSlot& s : c->slot) {
    if (s.graph) cudaGraphExecDestroy(s.graph);
    for (cudaEvent_t e : {s.ev_in, s.ev_done, s.ev_out})
      if (e) cudaEventDestroy(e);
    if (s.h_state) cudaFreeHost(s.h_state);
  }
  for (cudaStream_t s : {c->st_in, c->st, c->st_out})
    if (s) cudaStreamDestroy(s);
  for (void* p : c->allocs) cudaFree(p);
  delete c;
  return OFS_OK;
}

int ofs_clips_wait(ofs_clips* c) {
  OFS_REQUIRE(c, "ofs_clips_wait: null handle");
  if (c->collected == c->frame) { set_error("ofs_clips_wait: no step in flight"); return OFS_ESTATE; }
  ClipSlot& s = c->slot[c->collected % kSlots];
  OFS_CUDA(cudaEventSynchronize(s.ev_out));
  s.busy = false;
  ++c->collected;
  return OFS_OK;
}

int ofs_clips_in_flight(const ofs_clips* c) { return c ? (int)(c->frame - c->collected) : -1; }

int ofs_clips_depth(void) { return kSlots; }

int ofs_clips_reset(ofs_clips* c) {
  OFS_REQUIRE(c, "ofs_clips_reset: null handle");
  while (c->collected < c->frame) {
    const int rc = ofs_clips_wait(c);
    if (rc != OFS_OK) return rc;
  }
  c->frame = c->collected = 0;
  return OFS_OK;
}

long long ofs_clips_frame_index(const ofs_clips* c) { return c ? c->frame : -1; }

// Three streams per clip set: uploads, kernels, downloads.  Step i uses slot i % 3; the slot's events order
//   upload(i) -> graph(i) -> download(i),   graph(i-3) -> upload(i)  (input staging free),
//   download(i-3) -> graph(i)               (output staging free),
// and graphs run in submission order on one stream, which is the recurrence through the history ring.
int ofs_clips_submit_host(ofs_clips* c, const uint8_t* frames_bgr, uint8_t* out_bgr_u8, float* out_bgr_f32) {
  OFS_REQUIRE(c && frames_bgr && out_bgr_u8, "ofs_clips_submit_host: null pointer");
  if (!net_loaded(c->net)) { set_error("ofs_clips_submit_host: weights not loaded"); return OFS_ESTATE; }
  if (c->frame - c->collected >= kSlots) {
    set_error("ofs_clips_submit_host: %d steps already in flight; call ofs_clips_wait first", kSlots);
    return OFS_ESTATE;
  }
  OFS_CUDA(cudaSetDevice(c->device));
  const size_t fpx = (size_t)c->n * c->H * c->W * 3;
  const bool want_f32 = out_bgr_f32 != nullptr;
  const long long i = c->frame;
  ClipSlot& s = c->slot[i % kSlots];
  if (!s.graph || s.graph_f32 != want_f32 || s.graph_generation != net_weights_generation(c->net)) {
    if (s.graph) { cudaGraphExecDestroy(s.graph); s.graph = nullptr; }
    int rc = net_prepare(c->net, c->n);
    if (rc != OFS_OK) return rc;
    const uint64_t l0 = launch_count();
    OFS_CUDA(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_step(c, s, c->st, want_f32);
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->st, &g);
    s.graph_launches = (int)(launch_count() - l0);
    count_launch(-s.graph_launches);
    if (rc != OFS_OK) { if (g) cudaGraphDestroy(g); return rc; }
    OFS_CUDA(ce);
    const cudaError_t ie = cudaGraphInstantiate(&s.graph, g, 0);
    cudaGraphDestroy(g);
    OFS_CUDA(ie);
    s.graph_f32 = want_f32;
    s.graph_generation = net_weights_generation(c->net);
  }
  StepState& h = *s.h_state;   // the slot's previous upload completed before its step was waited for
  h = StepState{};
  for (int j = 0; j < 8; ++j) h.hist_slot[j] = (int)(std::max<long long>(i - kOffsets[j], 0) % kRing);
  h.write_slot = (int)(i % kRing);
  h.first = i == 0;
  if (i >= kSlots) OFS_CUDA(cudaStreamWaitEvent(c->st_in, s.ev_done, 0));
  OFS_CUDA(cudaMemcpyAsync(s.d_state, &h, sizeof(h), cudaMemcpyHostToDevice, c->st_in));
  OFS_CUDA(cudaMemcpyAsync(s.d_frame, frames_bgr, fpx, cudaMemcpyDefault, c->st_in));   // host (pinned) or device source
  OFS_CUDA(cudaEventRecord(s.ev_in, c->st_in));
  OFS_CUDA(cudaStreamWaitEvent(c->st, s.ev_in, 0));
  if (i >= kSlots) OFS_CUDA(cudaStreamWaitEvent(c->st, s.ev_out, 0));
  OFS_CUDA(cudaGraphLaunch(s.graph, c->st));
  count_launch(s.graph_launches);
  OFS_CUDA(cudaEventRecord(s.ev_done, c->st));
  OFS_CUDA(cudaStreamWaitEvent(c->st_out, s.ev_done, 0));
  OFS_CUDA(cudaMemcpyAsync(out_bgr_u8, s.d_out_u8, fpx, cudaMemcpyDefault, c->st_out));   // host or device destination
  if (want_f32) OFS_CUDA(cudaMemcpyAsync(out_bgr_f32, s.d_out_f32, fpx * 4, cudaMemcpyDefault, c->st_out));
  OFS_CUDA(cudaEventRecord(s.ev_out, c->st_out));
  s.busy = true;
  ++c->frame;
  return OFS_OK;
}

int ofs_clips_step_host(ofs_clips* c, const uint8_t* frames_bgr, uint8_t* out_bgr_u8, float* out_bgr_f32) {
  OFS_REQUIRE(c, "ofs_clips_step_host: null handle");
  if (c->frame != c->collected) {
    set_error("ofs_clips_step_host: %d submitted step(s) not yet waited for", (int)(c->frame - c->collected));
    return OFS_ESTATE;
  }
  const int rc = ofs_clips_submit_host(c, frames_bgr, out_bgr_u8, out_bgr_f32);
  return rc != OFS_OK ? rc : ofs_clips_wait(c);
}

}  // extern "C"
