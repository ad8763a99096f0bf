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
