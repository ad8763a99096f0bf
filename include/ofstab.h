/*
 * ofstab.h -- C ABI of libofstab.so: the B200 (sm_100a) inference hot path of
 * posgraph/coupe.optical_flow_based_deep_video_stabilization.
 *
 * The reference has no FFI / plugin layer: its "operator API" is a set of Python call
 * signatures evaluated inside one tf.Session.run per frame.  Each entry point below
 * replaces the device work behind one of those signatures (reference file:line cited
 * per function; paths relative to the reference repo, "main_dl.py" =
 * main_flownetS_pyramid_noprevloss_dataloader.py).  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every tensor is a raw pointer + sizes, NHWC, float32 unless
 *     stated.  "dev" pointers are CUDA device pointers on the net's / current device,
 *     "host" pointers are CPU memory (pinned memory makes the copies asynchronous).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device-
 *     pointer entry points are stream-ordered and return without synchronising.
 *   - every function returns an ofs_status (0 = OK).  On failure a thread-local message
 *     is available from ofs_last_error().  There is no CPU fallback: a device that is
 *     not compute capability 10.x yields OFS_ENOTSM100.
 */
#ifndef OFSTAB_H_
#define OFSTAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFS_VERSION 100 /* 0.1.0 */

typedef enum ofs_status {
  OFS_OK = 0,
  OFS_EINVAL = 1,    /* bad shape / null pointer / misaligned pointer                */
  OFS_ECUDA = 2,     /* a CUDA runtime / driver call failed (message has the detail) */
  OFS_ENOTSM100 = 3, /* device is not a Blackwell sm_100 part                         */
  OFS_ENOMEM = 4,
  OFS_ESTATE = 5     /* call order (e.g. forward before load_weights)                */
} ofs_status;

typedef void* ofs_stream; /* cudaStream_t */

/* operand format of the tensor-core GEMMs (accumulation is always fp32) */
typedef enum ofs_precision {
  OFS_PREC_BF16 = 0, /* bf16 operands (default; the north-star configuration)              */
  OFS_PREC_FP16 = 1  /* IEEE half operands: same tensor rate, 8x finer mantissa, less range */
} ofs_precision;

int ofs_version(void);
/* thread-local message of the last failure on this thread ("" if none) */
const char* ofs_last_error(void);
/* OFS_OK iff `device` exists and is compute capability 10.x */
int ofs_device_check(int device);
/* number of kernels this library has launched since load (all threads); bench.py's gpu_launches */
uint64_t ofs_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense backward warp.  Replaces tf_warp(img, flow, H, W) + get_pixel_value
 * (main_dl.py:44-130): x = col + flow[...,0], y = row + flow[...,1]; corners by truncation
 * toward zero, clipped to the image; weights from the clipped corners.
 *   img  [B,H,W,C] dev, flow [B,H,W,2] dev, out [B,H,W,C] dev.
 */
int ofs_tf_warp(const float* img, const float* flow, float* out, int B, int H, int W, int C, ofs_stream stream);

/* Tuning knob for ofs_tf_warp / ofs_flow_resize_warp at C=3: 3 (default) = instruction-lean direct
 * gathers, 0 = first direct-gather kernel, 1 / 2 = shared-memory staged source tiles (12-byte / padded
 * 16-byte pixels).  Results are identical. */
int ofs_set_warp_variant(int variant);

/* Test-mode flow glue (main_dl.py:497-498): flow2*(384/fh) -> TF1 legacy bilinear resize to
 * HxW -> x*W/512, y*H/384.   flow2 [B,fh,fw,2] dev (fh=382, fw=510 for the network),
 * out [B,H,W,2] dev. */
int ofs_flow_resize(const float* flow2, float* out, int B, int fh, int fw, int H, int W, ofs_stream stream);

/* The same glue with the factor in front of "/ fh" spelled out: pre_mul = 384 gives ofs_flow_resize; the homography mode
 * (evaluate_originalSize_homo, main_dl.py:681-682) resizes predict_flow3 * out_h / 48, i.e. flow [B,48,64,2], pre_mul = H. */
int ofs_flow_resize_ex(const float* flow, float* out, int B, int fh, int fw, int H, int W, float pre_mul, ofs_stream stream);

/* Fused main_dl.py:497-514: ofs_flow_resize + ofs_tf_warp without materialising the HxW flow.
 *   img [B,H,W,3] dev, flow2 [B,fh,fw,2] dev, out [B,H,W,3] dev. */
int ofs_flow_resize_warp(const float* img, const float* flow2, float* out, int B, int H, int W, int fh, int fw,
                         ofs_stream stream);

/* ------------------------------------------------------------------------------------------
 * Spatial transformers.  Replace AffineTransformer(out_size).transform(inp, theta)
 * (spatial_transformer.py:373-452) and ProjectiveTransformer (:519-608), both through
 * _meshgrid (:755-779) and bilinear_interp (:902-964: 1-px zero border, clip to [-1, W]).
 *   im [B,H,W,C] dev, theta [B,6] (affine) or [B,8] (projective) dev, out [B,oH,oW,C] dev.
 */
int ofs_grid_sample_affine(const float* im, const float* theta, float* out, int B, int H, int W, int C, int oH,
                           int oW, ofs_stream stream);
int ofs_grid_sample_projective(const float* im, const float* theta, float* out, int B, int H, int W, int C, int oH,
                               int oW, ofs_stream stream);

/* ------------------------------------------------------------------------------------------
 * Lie-algebra homography / affine warp (warp.py).
 * ofs_vec2mtrx replaces vec2mtrx(config, p) (warp.py:25-43): warp_type 0 = "homography"
 * (p [B,8]), 1 = "affine" (p [B,6]); pMtrx [B,3,3] = sum_{i<warp_approx} A^i / i!.
 * ofs_lie_warp replaces transformImage (warp.py:46-86; srcH==outH, srcW==outW) and
 * transformCropImage (warp.py:89-129; source dataH x dataW, output height x W):
 *   image [B,srcH,srcW,3] dev, pMtrx [B,3,3] dev, refMtrx [3,3] dev, out [B,outH,outW,3] dev.
 */
int ofs_vec2mtrx(const float* p, float* pMtrx, int B, int warp_type, int warp_approx, ofs_stream stream);
int ofs_lie_warp(const float* image, const float* pMtrx, const float* refMtrx, float* out, int B, int srcH,
                 int srcW, int outH, int outW, ofs_stream stream);

/* ------------------------------------------------------------------------------------------
 * FlowNetS-pyramid.  Replaces flownetS_pyramid(feats, batch_size, is_train=False)
 * (model.py:786-893) evaluated under sess.run (main_dl.py:569).
 */
typedef struct ofs_net ofs_net;

/* one named float32 host array of a TensorLayer npz checkpoint (main_dl.py:520) */
typedef struct ofs_named_array {
  const char* name;  /* e.g. "main_net/flownetS/3_1/W_conv2d:0"; scope prefix and ":0" optional */
  const float* data; /* host, C-contiguous, TF layout ([kh,kw,cin,cout] conv, [4,4,cout,cin] deconv) */
  int64_t numel;
} ofs_named_array;

/* Allocates every device buffer for batches up to max_batch (no allocation afterwards). */
int ofs_net_create(ofs_net** net, int device, int max_batch, int precision /* ofs_precision */);
int ofs_net_destroy(ofs_net* net);
/* BN fold (W' = W/sqrt(var+1e-5), b' = (b-mu)/sqrt(var+1e-5)+beta), 16-bit conversion and
 * GEMM-K-major packing; copies to the device.  Missing BN statistics default to the
 * TensorLayer initial values (mu 0, var 1, beta 0); a missing weight is OFS_EINVAL. */
int ofs_net_load_weights(ofs_net* net, const ofs_named_array* arrays, int n);
/* feats [B,384,512,27] dev -> the 5 flow maps of the returned dict (model.py:893), dev:
 * f6 [B,6,8,2] f5 [B,12,16,2] f4 [B,24,32,2] f3 [B,48,64,2] f2 [B,382,510,2]; any may be NULL. */
int ofs_net_forward(ofs_net* net, const float* feats, int B, float* f6, float* f5, float* f4, float* f3, float* f2,
                    ofs_stream stream);
/* The whole sess.run(outputs_warpedimg) of main_dl.py:569: forward + flow glue + warp.
 * feats [B,384,512,27] dev, frames [B,H,W,3] dev -> out [B,H,W,3] dev.  flow2_out
 * ([B,382,510,2] dev) may be NULL. */
int ofs_net_stabilize(ofs_net* net, const float* feats, const float* frames, float* out, float* flow2_out, int B,
                      int H, int W, ofs_stream stream);
/* Same call on HOST buffers: H2D of feats+frames, compute, D2H of out, synchronous on return.
 * (This is the boundary the reference's feed_dict / fetch crosses, main_dl.py:568-569.)
 * The call is PCIe-bound and pipelined over sub-batches.  With OFS_HOST_PACK=1 in the environment of ofs_net_create, a
 * bf16 net rounds the float32 network input to bf16 on the HOST (a small worker pool, one sub-batch ahead of the copies)
 * and sends it as 16-bit: the device's first kernel performs the very same round-to-nearest-even, so the result is
 * bit-identical and 85 MB instead of 170 MB of network input cross the bus per 8-pair step.  Off by default: on a host
 * whose memory system is the bottleneck (measured) the conversion's own traffic cancels the gain.
 * ofs_net_host_h2d_bytes: the bytes such a call copies host -> device. */
int ofs_net_stabilize_host(ofs_net* net, const float* feats_host, const float* frames_host, float* out_host, int B,
                           int H, int W);
long long ofs_net_host_h2d_bytes(const ofs_net* net, int B, int H, int W);
/* Debug / parity: copy a named activation, widened to float32, into out (dev, capacity in
 * elements).  shape4 receives [B,H,W,C].  Names: conv1 conv2 conv3 conv3_1 conv4 conv4_1 conv5
 * conv5_1 conv6 conv6_1 concat5 concat4 concat3 concat2 (logical channels only). */
int ofs_net_get_activation(ofs_net* net, const char* name, int B, float* out, int64_t capacity, int* shape4,
                           ofs_stream stream);
/* Measurement aid (bench.py): runs `iters` forwards (+ the fused flow-resize/warp when frames and out
 * are non-NULL) with a CUDA event recorded on `stream` after every kernel launch, and returns the
 * mean duration of each launch in launch order.  ms[cap], macs[cap] (literal multiply-accumulates of
 * the layer, 0 for non-GEMM kernels), names[cap*32] (NUL-terminated, 32 bytes each).  Synchronous. */
int ofs_net_profile(ofs_net* net, const float* feats, const float* frames, float* out, int B, int H, int W, int iters,
                    float* ms, double* macs, char* names, int cap, int* count, ofs_stream stream);
/* Measurement aid (bench.py roofline): `iters` back-to-back repetitions of one kernel set are captured into
 * one CUDA graph and replayed between two CUDA events on the net's stream -- the kernels' device time as the
 * step executes them, without per-launch event or host overhead.  which = 0: the 14 dense conv / transposed
 * conv GEMM launches of a forward (plus their split-K reductions) on the activations left by the last forward,
 * macs_per_set = their literal multiply-accumulates; which = 1: the fused flow-resize + warp of
 * frames [B,H,W,3] dev -> out dev on the last forward's flow.  Synchronous. */
int ofs_net_time_kernels(ofs_net* net, int which, const float* frames, float* out, int B, int H, int W, int iters,
                         float* ms_per_set, double* macs_per_set, int* launches_per_set);
/* Measurement aid (OFS_CHAIN_TRACE=1 in the environment): per-CTA globaltimer stamps (ns) of the last conv5 ... conv6_1
 * chain launch, 32 words per CTA: [4l + {0: layer l's GEMM phase done, 1: past the barrier, 2: reduction done, 3: past
 * the barrier}], [16] kernel entry.  Returns the number of words copied (0 when tracing is off). */
int ofs_chain_trace_read(long long* host_words, int max_words);
/* per-forward kernel launches (constant for a given B) */
int ofs_net_launches_per_forward(const ofs_net* net);
/* Diagnostics of the step-graph cache of ofs_net_stabilize: what = 0 -> graphs captured so far, 1 -> launches that
 * re-pointed a cached graph at new feats / frames / out / flow2 addresses, 2 -> graphs currently cached.  The
 * reference feeds a new array every frame (main_dl.py:568-569); one capture per (B, H, W) must serve them all. */
long long ofs_net_graph_stats(const ofs_net* net, int what);

/* ------------------------------------------------------------------------------------------
 * The reference's other test modes around the same network.  All pointers are device pointers
 * unless named *_host; every op runs on `stream`.
 *
 * cv2.warpPerspective(frame, h, (out_w, out_h)) of evaluate_originalSize_homo (main_dl.py:743): uint8 [B,src_h,src_w,3]
 * -> uint8 [B,out_h,out_w,3], default flags (INTER_LINEAR, h is the FORWARD map and is inverted first), BORDER_CONSTANT 0.
 * OpenCV's fixed-point arithmetic restated exactly: byte-identical to cv2 (tests/test_gpu_modes.py).
 * h_host: B row-major 3x3 double matrices on the HOST (what cv2.findHomography returns). */
int ofs_warp_perspective_u8(const uint8_t* src, const double* h_host, uint8_t* dst, int B, int src_h, int src_w, int out_h,
                            int out_w, ofs_stream stream);
/* tf.image.resize_images(x[..., c0:c0+C], [out_h, out_w]) of evaluate() (main_dl.py:806: the current frame, channels 24:27
 * of the network input, to 382x510): TF-1.10 legacy bilinear.  in [B,H,W,in_channels] float32 -> out [B,out_h,out_w,C]. */
int ofs_tf1_resize_bilinear(const float* in, float* out, int B, int H, int W, int in_channels, int c0, int C, int out_h,
                            int out_w, ofs_stream stream);
/* cv2.resize(float32 image, (out_w, out_h)), INTER_LINEAR, times post_mul (main_dl.py:862: cv2.resize(warped, (512, 384)) * 255). */
int ofs_cv_resize_linear_f32(const float* in, float* out, int B, int H, int W, int C, int out_h, int out_w, float post_mul,
                             ofs_stream stream);
/* evaluate_blurNma (main_flownetS_pyramid.py:634-641): k x k mean filter with zero padding of both flow planes (conv2d with
 * the constant 1/(k*k), SAME) and the mix a * smooth + b * prev (0.9 / 0.1).  prev may be NULL (out = smooth).
 * flow, prev, out, scratch: [B,H,W,2] float32, all distinct. */
int ofs_flow_box_blur_ema(const float* flow, const float* prev, float* out, float* scratch, int B, int H, int W, int k, float a,
                          float b, ofs_stream stream);
/* scipy.signal.medfilt(vol, k) of evaluate_medianNma (main_flownetS_pyramid.py:809) for a 3-d array [H,W,C] and a scalar
 * kernel size k: the window is k x k x k -- it spans the channel axis too -- zero padded; odd k <= 7. */
int ofs_medfilt_nd3(const float* in, float* out, int H, int W, int C, int k, ofs_stream stream);

/* ------------------------------------------------------------------------------------------
 * Clip driver (SURVEY 8(f) "next" row 1): the per-frame loop of evaluate_originalSize(), main_dl.py:535-630, with its
 * state on the device.  n_clips clips of H x W frames advance in lockstep as one batch (n_clips <= the net's
 * max_batch; the faithful single-clip loop is n_clips = 1).  Per clip the handle keeps a 32-slot ring of the
 * resized (512x384, cv2.resize INTER_LINEAR fixed-point arithmetic), channel-swapped np.uint8 outputs; one step =
 * one iteration of the reference loop body: curinput assembly (main_dl.py:544-558: 8 history taps at offsets
 * 31,23,15,7,4,3,2,1 with the frame-0 clamp, + the current frame), sess.run (:569), totaloutputFrame[i] (:625).
 *   frames_bgr   host uint8 [n_clips,H,W,3]: cap.read() of every clip (BGR)
 *   out_bgr_u8   host uint8 [n_clips,H,W,3]: np.uint8(totaloutputFrame[i]), what out.write() receives (:630)
 *   out_bgr_f32  host float32 [n_clips,H,W,3] or NULL: totaloutputFrame[i] itself (BGR, 0..255, not clipped)
 * 3 bytes per pixel cross PCIe each way instead of 32 MB + 11 MB of float32 per 720p frame.
 * ofs_clips_step_host is synchronous.  ofs_clips_submit_host / ofs_clips_wait are the same step split in two so
 * that the upload of frame i+1 and the download of frame i-1 overlap the kernels of frame i (the input frames do
 * not depend on earlier outputs; the recurrence through the history ring stays ordered on the device): at most
 * ofs_clips_depth() (= 3) steps may be in flight, ofs_clips_wait blocks until the OLDEST one has landed in its host buffers, which must
 * stay valid (and should be page-locked) until then.  Results are identical to stepping synchronously. */
typedef struct ofs_clips ofs_clips;
int ofs_clips_create(ofs_clips** clips, ofs_net* net, int n_clips, int H, int W);
int ofs_clips_destroy(ofs_clips* clips);
int ofs_clips_reset(ofs_clips* clips); /* next step is frame 0 again */
long long ofs_clips_frame_index(const ofs_clips* clips);
int ofs_clips_step_host(ofs_clips* clips, const uint8_t* frames_bgr, uint8_t* out_bgr_u8, float* out_bgr_f32);
int ofs_clips_submit_host(ofs_clips* clips, const uint8_t* frames_bgr, uint8_t* out_bgr_u8, float* out_bgr_f32);
int ofs_clips_wait(ofs_clips* clips);
int ofs_clips_in_flight(const ofs_clips* clips); /* submitted, not yet waited for */
int ofs_clips_depth(void);                        /* how many steps may be in flight */

/* ------------------------------------------------------------------------------------------
 * Stand-alone implicit-GEMM convolution on the same tcgen05 kernel the network uses (unit
 * tests, micro-benchmarks):  y = act(conv2d(zero_pad(x, k/2), W, stride) + b), NHWC.
 *   x [B,H,W,Cin] dev f32, w host [k,k,Cin,Cout] f32, b host [Cout] f32 (may be NULL),
 *   y [B,Ho,Wo,Cout] dev f32, Ho = (H + 2(k/2) - k)/stride + 1.  stride in {1,2}; stride 2
 *   needs even H, W.  transposed != 0: k must be 4, stride 2, w host [4,4,Cout,Cin] (TF
 *   conv2d_transpose SAME), y [B,2H,2W,Cout].  lrelu != 0 applies max(v, 0.1 v).
 * Operands are rounded to `precision` exactly as in the network.  Synchronous.
 */
int ofs_conv2d_nhwc(const float* x, const float* w_host, const float* b_host, float* y, int B, int H, int W, int Cin,
                    int Cout, int k, int stride, int transposed, int lrelu, int precision, ofs_stream stream);

/* Same with explicit tiling: block_n in {16,32,64,128,256} (0 = automatic), a split-K factor
 * (ksplit > 1 reduces fp32 partials from a workspace into the 16-bit activation format, so the result
 * carries one 16-bit rounding, exactly as the network's split-K layers do) and cta_group (1, or 2 = CTA
 * pairs: tcgen05 cta_group::2, 256-row tiles, weight tile split over the pair; needs block_n >= 32;
 * 4 = pairs + slab groups, 8 / 32 = two / four K chunks per stage, 16 = split-K inside a thread-block cluster of
 * ksplit <= 8 CTAs reduced through distributed shared memory, block_n 256 -- same bits as the workspace path;
 * 5 = pairs + slab groups with two output pixels per GEMM row (the k7 s2 conv1 form, block_n 128); 64 / 66 = the
 * phase-stacked transposed conv (Cout 64, fused head columns with zero head weights) on single CTAs / CTA pairs).
 * out16 = 1 runs the network's 16-bit activation epilogue (shared-memory staging + TMA stores when
 * block_n >= 64; needs Cout % block_n == 0) and widens the result to fp32 afterwards. */
int ofs_conv2d_nhwc_ex(const float* x, const float* w_host, const float* b_host, float* y, int B, int H, int W,
                       int Cin, int Cout, int k, int stride, int transposed, int lrelu, int precision, int block_n,
                       int ksplit, int cta_group, int out16, ofs_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* OFSTAB_H_ */
