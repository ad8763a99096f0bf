// What does one dependent kernel launch cost on sm_100a, and which kernel attribute makes it expensive?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/launch_overhead benchmarks/launch_overhead.cu
//
// N back-to-back launches of a near-empty 148-CTA kernel between one event pair; variants add, one at a time,
// what the conv GEMM kernel carries: 225 KB of dynamic shared memory, a TMEM allocation, a 1.6 KB
// __grid_constant__ parameter block, a prefetch.tensormap, 192-thread CTAs, a body of a few microseconds.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

struct Big { unsigned char bytes[1664]; };

template <bool kTmem, bool kSpin>
__global__ void __launch_bounds__(192, 1) k_small(int spin_cycles, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (kTmem) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (kSpin) {
    const long long t0 = clock64();
    while (clock64() - t0 < spin_cycles) {}
  }
  if (spin_cycles < 0) sink[threadIdx.x] = smem[threadIdx.x];
  if (kTmem) {
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
  }
}

__global__ void __launch_bounds__(192, 1) k_bigparam(const __grid_constant__ Big p, int spin_cycles, int* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (spin_cycles < 0) sink[threadIdx.x] = smem[threadIdx.x] + p.bytes[threadIdx.x];
  const long long t0 = clock64();
  while (clock64() - t0 < spin_cycles) {}
}

template <class F>
float time_launches(F launch, int n) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 10; ++i) launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < n; ++i) launch();
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / n;
}

int main() {
  int* sink;
  cudaMalloc(&sink, 4096);
  const int n = 200;
  const int big = 225 * 1024;
  cudaFuncSetAttribute(k_small<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_small<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_small<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_small<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_bigparam, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  Big bp = {};
  for (int spin : {0, 10000}) {
    printf("--- body spin %d cycles (%.1f us at 1.965 GHz)\n", spin, spin / 1965.0);
    printf("empty, 0 smem, 148x192           : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, 0>>>(spin, sink); }, n));
    printf("empty, 0 smem, 148x1024          : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 1024, 0>>>(spin, sink); }, n));
    printf("48 KB smem                       : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, 48 * 1024>>>(spin, sink); }, n));
    printf("100 KB smem                      : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, 100 * 1024>>>(spin, sink); }, n));
    printf("225 KB smem                      : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, big>>>(spin, sink); }, n));
    printf("225 KB smem + TMEM 512           : %6.2f us/launch\n", time_launches([&] { k_small<true, true><<<148, 192, big>>>(spin, sink); }, n));
    printf("0 smem + TMEM 512                : %6.2f us/launch\n", time_launches([&] { k_small<true, true><<<148, 192, 0>>>(spin, sink); }, n));
    printf("225 KB smem + 1.6 KB params      : %6.2f us/launch\n", time_launches([&] { k_bigparam<<<148, 192, big>>>(bp, spin, sink); }, n));
    printf("alternating 225 KB / 0 smem      : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, big>>>(spin, sink); k_small<false, true><<<148, 192, 0>>>(spin, sink); }, n) / 2);
    printf("alternating 225 KB / 40 KB smem  : %6.2f us/launch\n", time_launches([&] { k_small<false, true><<<148, 192, big>>>(spin, sink); k_small<false, true><<<592, 256, 40 * 1024>>>(spin, sink); }, n) / 2);
  }
  return 0;
}
