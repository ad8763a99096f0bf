// Shared host-side plumbing of libofstab.so: status codes, thread-local error text,
// launch accounting, device checks.  No torch types anywhere below the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../../include/ofstab.h"

namespace ofs {

constexpr int kNumSMsB200 = 148;

void set_error(const char* fmt, ...);
const char* get_error();
void count_launch(int n = 1);
uint64_t launch_count();
int sm_count();  // SMs of the current device (cached)

// returns OFS_OK or sets the error text
int check_cuda(cudaError_t e, const char* what, const char* file, int line);
int require_sm100(int device);

// Programmatic dependent launch (PDL): every kernel of the hot path is launched with the
// programmatic-stream-serialization attribute and executes pdl_wait() before it touches global memory,
// so its CTAs are scheduled, and its prologue (barrier init, TMEM allocation, descriptor prefetch) runs,
// while the previous kernel of the stream drains.  The attribute is OFF unless OFS_PDL=1 is set: measured
// (B200, batch 8): stream launches 846 -> 820 us/step with PDL, one CUDA graph per step 786 us, graph + PDL 803 us.
bool pdl_enabled();
// kernel classes for the selective mode (OFS_PDL=2): 0 = streaming / large-grid kernel, 1 = persistent tcgen05 GEMM,
// 2 = small helper (split-K reduce, pyramid step).  pdl_allow() is called once per launch, in launch order.
void pdl_set_kind(int kind);
bool pdl_allow();

#ifdef __CUDACC__
// blocks until every prerequisite grid has completed and its memory operations are visible (no-op when the
// kernel was not launched as a programmatic dependent)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next kernel of the stream start launching (it still waits in its own pdl_wait())
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_allow() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace ofs

#define OFS_CUDA(expr)                                                          \
  do {                                                                          \
    int _st = ::ofs::check_cuda((expr), #expr, __FILE__, __LINE__);             \
    if (_st != OFS_OK) return _st;                                              \
  } while (0)

#define OFS_REQUIRE(cond, ...)                                                  \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ::ofs::set_error(__VA_ARGS__);                                            \
      return OFS_EINVAL;                                                        \
    }                                                                           \
  } while (0)

#define OFS_LAUNCH_CHECK()                                                      \
  do {                                                                          \
    ::ofs::count_launch();                                                      \
    OFS_CUDA(cudaGetLastError());                                               \
  } while (0)
