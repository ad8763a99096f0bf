#include "ofs_common.cuh"

namespace ofs {

static thread_local char g_err[1024] = {0};
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)(int64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

static int pdl_mode() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("OFS_PDL");
    cached = e ? atoi(e) : 0;   // default off: with the step graph, PDL on every kernel measured 1.5 % slower
    if (cached < 0 || cached > 4) cached = 0;
  }
  return cached;
}
bool pdl_enabled() { return pdl_mode() != 0; }
static thread_local int g_pdl_kind = 0, g_pdl_prev = 0;
void pdl_set_kind(int kind) { g_pdl_kind = kind; }
bool pdl_allow() {
  const int kind = g_pdl_kind, prev = g_pdl_prev;
  g_pdl_prev = kind;
  g_pdl_kind = 0;
  const int mode = pdl_mode();
  if (mode == 1) return true;
  // selective: only GEMM / helper kernels following GEMM / helper kernels start early (a 225 KB-smem GEMM CTA that
  // lands beside a streaming kernel's blocks shrinks their L1)
  if (mode == 2) return kind != 0 && prev != 0;
  // 3: only the small helpers behind a GEMM (their blocks fit beside the draining GEMM CTAs: the launch latency of the
  // split-K reductions and pyramid steps overlaps the GEMM's tail); 4: also the GEMM behind a small helper
  if (mode == 3) return kind == 2 && prev == 1;
  if (mode == 4) return (kind == 2 && prev == 1) || (kind == 1 && prev == 2);
  return false;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMsB200;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMsB200;
    cached[dev] = n;
  }
  return cached[dev];
}

int check_cuda(cudaError_t e, const char* what, const char* file, int line) {
  if (e == cudaSuccess) return OFS_OK;
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return OFS_ECUDA;
}

int require_sm100(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libofstab has no CPU fallback", cudaGetErrorString(e));
    return OFS_ECUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (have %d)", device, n);
    return OFS_EINVAL;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    set_error("device %d is sm_%d%d; libofstab is built for sm_100a only (no fallback path)", device, major, minor);
    return OFS_ENOTSM100;
  }
  return OFS_OK;
}

}  // namespace ofs

extern "C" {
int ofs_version(void) { return OFS_VERSION; }
const char* ofs_last_error(void) { return ofs::get_error(); }
int ofs_device_check(int device) { return ofs::require_sm100(device); }
uint64_t ofs_launch_count(void) { return ofs::launch_count(); }
}
