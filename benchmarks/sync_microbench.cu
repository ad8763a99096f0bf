// Microbenchmark of the per-K-block handshake primitives of a warp-specialised tcgen05 pipeline on sm_100a:
// how many SM cycles does one producer/consumer iteration cost when nothing but the synchronisation is done?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/sync_microbench benchmarks/sync_microbench.cu -lcuda
//
// Variants (producer thread = warp 0 lane 0, consumer thread = warp 1 lane 0, S-stage ring of mbarriers):
//   0  producer: wait(empty) + arrive(full)              consumer: wait(full) + arrive(empty)
//   1  producer: wait(empty) + arrive.expect_tx(full, 0) consumer: wait(full) + tcgen05.commit(empty)
//   2  as 1, + the consumer issues 4 tcgen05.mma (M128 N=BN K16) on garbage smem per iteration
//   3  as 1, + the producer issues one 16 KB cp.async.bulk (global->smem) per iteration (expect_tx 16 KB)
//   4  as 3 with two 16 KB bulk copies per iteration
//   5  2 + 4 combined
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int S>
__global__ void __launch_bounds__(64, 1) k(int variant, int iters, int bn, const uint8_t* src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2 * S + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = s32(&bars[0]), empty0 = s32(&bars[S]), done = s32(&bars[2 * S]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t sbase = (s32(smem) + 1023u) & ~1023u;
  const int stage_bytes = bn > 128 ? 49152 : 32768;   // A 16 KB + B 16 / 32 KB
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0) {
    t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait(empty0 + 8 * st, ph ^ 1);
      if (variant == 0) arrive(full0 + 8 * st);
      else if (variant == 1 || variant == 2) expect_tx(full0 + 8 * st, 0);
      else {
        const int n = (variant == 3) ? 1 : 2;
        expect_tx(full0 + 8 * st, 16384 * n);
        for (int j = 0; j < n; ++j)
          bulk(sbase + st * stage_bytes + j * 16384, src + ((size_t)(i * 2 + j) % 512) * 16384 + (size_t)blockIdx.x * (8u << 20),
               16384, full0 + 8 * st);
      }
      if (++st == S) { st = 0; ph ^= 1; }
    }
    t1 = clock64();
    out[blockIdx.x * 4 + 0] = t1 - t0;
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    t0 = clock64();
    int st = 0; uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
      wait(full0 + 8 * st, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (variant == 0) arrive(empty0 + 8 * st);
      else {
        if (variant == 2 || variant == 5) {
          const uint64_t da = desc(sbase + st * stage_bytes), db = desc(sbase + st * stage_bytes + 16384);
          mma(tmem, da, db, idesc, i > 0); mma(tmem, da + 2, db + 2, idesc, 1);
          mma(tmem, da + 4, db + 4, idesc, 1); mma(tmem, da + 6, db + 6, idesc, 1);
        }
        commit(empty0 + 8 * st);
      }
      if (++st == S) { st = 0; ph ^= 1; }
    }
    if (variant != 0) { commit(done); wait(done, 0); }
    t1 = clock64();
    out[blockIdx.x * 4 + 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int S>
void run(int variant, int iters, int bn, const uint8_t* src, long long* out, int grid) {
  const size_t smem = (size_t)S * (bn > 128 ? 49152 : 32768) + 2048;
  cudaFuncSetAttribute(k<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<S><<<grid, 64, smem>>>(variant, iters, bn, src, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d S=%d: %s\n", variant, S, cudaGetErrorString(e)); exit(1); }
  long long h[4 * 148];
  cudaMemcpy(h, out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
  double p = 0, c = 0;
  for (int i = 0; i < grid; ++i) { p += h[i * 4]; c += h[i * 4 + 1]; }
  printf("variant %d  S=%d  BN=%3d  grid=%3d : producer %.0f cyc/iter, consumer %.0f cyc/iter\n", variant, S, bn, grid,
         p / grid / iters, c / grid / iters);
}

int main() {
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, (size_t)148 * (8u << 20) + (1u << 20));
  cudaMemset(src, 0, (size_t)148 * (8u << 20) + (1u << 20));
  cudaMalloc(&out, sizeof(long long) * 4 * 148);
  const int iters = 4000;
  for (int grid : {1, 148}) {
    for (int v = 0; v <= 5; ++v) {
      const int bn = 128;
      run<2>(v, iters, bn, src, out, grid);
      run<4>(v, iters, bn, src, out, grid);
      run<6>(v, iters, bn, src, out, grid);
    }
    run<6>(2, iters, 64, src, out, grid);
    run<4>(2, iters, 256, src, out, grid);
    run<6>(5, iters, 64, src, out, grid);
    run<4>(5, iters, 256, src, out, grid);
  }
  return 0;
}
