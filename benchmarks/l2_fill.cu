// L2 -> shared-memory fill microbenchmark (TMA, one 200 KB CTA per SM): what the implicit-GEMM producers can be fed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_fill benchmarks/l2_fill.cu -lcuda && ./l2_fill
// Every CTA runs a ring of STAGES shared-memory buffers; per iteration it fetches
//   `a_rows` rows of 128 B that only this CTA reads  (the A operand of a GEMM tile: distinct per CTA), and
//   `b_rows` rows of 128 B that EVERY CTA reads       (the B operand: the same weights for all M tiles),
// the latter either unicast (each CTA fetches all of them), or split over a cluster of C CTAs with TMA multicast (each
// CTA fetches b_rows / C rows and delivers them to all C).  Reported: bytes landed in shared memory per SM per cycle,
// bytes requested from L2 per cycle chip-wide, and the time per iteration.  The working set (64 MB) stays L2-resident.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int MAX_STAGES = 12;
constexpr int SMEM_ROWS = 1536;  // rows of 128 B in the ring (192 KB): stages x rows per stage <= SMEM_ROWS

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void arrive_cluster(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Params {
  CUtensorMap map_a;     // box {64, a_box}
  CUtensorMap map_b;     // box {64, b_box}   (b_box = b_rows / C in multicast mode)
  int a_rows, a_box;     // rows per iteration private to the CTA, fetched as a_rows / a_box boxes
  int b_rows, b_box;     // rows per iteration shared by all CTAs
  int iters, csize, multicast;
  int stages, stage_rows;           // ring geometry
  int a_rows_total, b_rows_total;   // tensor heights (wrap-around)
  long long* cycles;     // per CTA
};

__global__ void __launch_bounds__(64, 1) fill_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int STAGES = p.stages;
  const uint32_t bars = base + SMEM_ROWS * 128;
  const uint32_t full0 = bars, empty0 = bars + 8 * MAX_STAGES;
  const uint32_t rank = p.csize > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, p.multicast ? p.csize : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (p.csize > 1) cluster_sync();
  const uint32_t stage_bytes = (uint32_t)(p.a_rows + p.b_rows) * 128u;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer
    int a_row = (int)((blockIdx.x * 7919u * (unsigned)p.a_rows) % (unsigned)(p.a_rows_total - p.a_rows));
    a_row -= a_row % 8;
    int b_row = 0;
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      wait(empty0 + 8 * s, ph ^ 1);
      const uint32_t dst = base + (uint32_t)s * (uint32_t)p.stage_rows * 128u;
      expect_tx(full0 + 8 * s, stage_bytes);
      for (int r = 0; r < p.a_rows; r += p.a_box) tma2d(dst + r * 128, &p.map_a, full0 + 8 * s, 0, a_row + r);
      if (p.multicast) {
        // this CTA fetches slice `rank` of the shared rows and delivers it to every CTA of the cluster
        const int slice = p.b_rows / p.csize;
        for (int r = 0; r < slice; r += p.b_box)
          tma2d_mc(dst + (p.a_rows + rank * slice + r) * 128, &p.map_b, full0 + 8 * s, 0, b_row + rank * slice + r, (uint16_t)((1u << p.csize) - 1));
      } else {
        for (int r = 0; r < p.b_rows; r += p.b_box) tma2d(dst + (p.a_rows + r) * 128, &p.map_b, full0 + 8 * s, 0, b_row + r);
      }
      a_row += p.a_rows; if (a_row + p.a_rows > p.a_rows_total) a_row = 0;
      b_row += p.b_rows; if (b_row + p.b_rows > p.b_rows_total) b_row = 0;
    }
  } else if (threadIdx.x == 32) {
    // consumer: the stage is "used" as soon as it is full; release it (in every CTA that wrote into it)
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      wait(full0 + 8 * s, ph);
      if (p.multicast) { for (int c = 0; c < p.csize; ++c) arrive_cluster(mapa(empty0 + 8 * s, c)); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
  if (p.csize > 1) cluster_sync();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)sym;
  const int a_total = 384 * 1024, b_total = 128 * 1024;   // rows of 128 B: 48 MB + 16 MB, L2-resident
  uint8_t *da, *db;
  long long* dc;
  CK(cudaMalloc(&da, (size_t)a_total * 128));
  CK(cudaMalloc(&db, (size_t)b_total * 128));
  CK(cudaMalloc(&dc, sizeof(long long) * 1024));
  CK(cudaMemset(da, 1, (size_t)a_total * 128));
  CK(cudaMemset(db, 2, (size_t)b_total * 128));
  const size_t smem = SMEM_ROWS * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  auto make_map = [&](CUtensorMap* m, void* base, int rows_total, int box_rows) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  };
  struct Case { const char* name; int a_rows, b_rows, csize, multicast, ctas; int stages = 4; int a_box = 0; int b_box = 0; };
  std::vector<Case> cases = {
      {"A only 128 rows (16 KB distinct per CTA)", 128, 0, 1, 0, 0},
      {"A only 256 rows", 256, 0, 1, 0, 0},
      {"B only 256 rows, unicast (same rows for every CTA)", 0, 256, 1, 0, 0},
      {"A 128 + B 256 unicast (1-CTA N=256 tile)", 128, 256, 1, 0, 0},
      {"A 128 + B 128 unicast (pair N=256: half of B per CTA)", 128, 128, 1, 0, 0},
      {"A 128 + B 256 multicast over 2", 128, 256, 2, 1, 0},
      {"A 128 + B 256 multicast over 4", 128, 256, 4, 1, 0},
      {"A 128 + B 64 unicast (what multicast-4 requests from L2)", 128, 64, 1, 0, 0},
      {"A 128 + B 80 unicast (deconv2 1-CTA tile)", 128, 80, 1, 0, 0},
      {"A 128 + B 256 unicast, 74 CTAs", 128, 256, 1, 0, 74},
      {"A 128 + B 256 unicast, 37 CTAs", 128, 256, 1, 0, 37},
      {"A 128 + B 256 unicast, 8 CTAs", 128, 256, 1, 0, 8},
      // is a stage's cost its bytes or its number of TMA boxes?  same bytes, more stages / fewer, larger boxes
      {"A 128, 1 box, 4 stages", 128, 0, 1, 0, 0, 4},
      {"A 128, 1 box, 8 stages", 128, 0, 1, 0, 0, 8},
      {"A 128, 1 box, 12 stages", 128, 0, 1, 0, 0, 12},
      {"A 128 as 2 boxes of 64, 8 stages", 128, 0, 1, 0, 0, 8, 64},
      {"A 128 as 4 boxes of 32, 8 stages", 128, 0, 1, 0, 0, 8, 32},
      {"A 256, 1 box, 6 stages", 256, 0, 1, 0, 0, 6},
      {"A 256 as 2 boxes of 128, 6 stages", 256, 0, 1, 0, 0, 6, 128},
      {"A 128 + B 144 (deconv3 stage), 2 boxes, 5 stages", 128, 144, 1, 0, 0, 5},
      {"A 128 + B 144, 2 boxes, 3 stages", 128, 144, 1, 0, 0, 3},
      {"A 256 + B 256 (two K chunks, merged boxes), 3 stages", 256, 256, 1, 0, 0, 3},
      {"A 256 + B 256 as 4 boxes of 128, 3 stages", 256, 256, 1, 0, 0, 3, 128, 128},
      {"A 128 + B 80 (deconv2 stage), 7 stages", 128, 80, 1, 0, 0, 7},
      {"A 256 + B 160 (two deconv2 K chunks merged), 3 stages", 256, 160, 1, 0, 0, 3},
      {"A 64 + B 64, 12 stages", 64, 64, 1, 0, 0, 12},
  };
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
  printf("device %s, %d SMs, max clock %d MHz\n", prop.name, sms, clock_khz / 1000);
  printf("%-58s %6s %9s %9s %12s %12s %10s\n", "case", "CTAs", "us", "clk/iter", "B/clk/SM in", "L2 B/clk all", "TB/s in");
  for (const Case& c : cases) {
    Params p = {};
    p.a_rows = c.a_rows; p.b_rows = c.b_rows; p.csize = c.csize; p.multicast = c.multicast;
    p.a_box = c.a_box ? c.a_box : c.a_rows ? (c.a_rows > 256 ? 128 : c.a_rows) : 8;
    const int slice = c.multicast ? c.b_rows / c.csize : c.b_rows;
    p.b_box = c.b_box ? c.b_box : slice ? slice : 8;
    p.stages = c.stages; p.stage_rows = c.a_rows + c.b_rows;
    if (p.stages > MAX_STAGES || p.stages * p.stage_rows > SMEM_ROWS) { printf("%-58s skipped (ring too large)\n", c.name); continue; }
    p.iters = 2000;
    p.a_rows_total = a_total; p.b_rows_total = b_total;
    p.cycles = dc;
    make_map(&p.map_a, da, a_total, p.a_box);
    make_map(&p.map_b, db, b_total, p.b_box);
    int ctas = c.ctas ? c.ctas : sms;
    ctas -= ctas % c.csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    long long cyc_max = 0;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0));
      CK(cudaLaunchKernelEx(&cfg, fill_kernel, p));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) {
        best = ms;
        std::vector<long long> h(ctas);
        CK(cudaMemcpy(h.data(), dc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
        cyc_max = 0; for (long long v : h) cyc_max = v > cyc_max ? v : cyc_max;
      }
    }
    const double in_bytes = (double)(c.a_rows + c.b_rows) * 128.0 * p.iters;                 // landed per CTA
    const double l2_bytes = (double)(c.a_rows + slice) * 128.0 * p.iters * ctas;            // requested chip-wide
    printf("%-58s %6d %9.1f %9.0f %12.1f %12.0f %10.2f\n", c.name, ctas, best * 1e3, (double)cyc_max / p.iters, in_bytes / (double)cyc_max, l2_bytes / (double)cyc_max,
           in_bytes * ctas / (best * 1e-3) / 1e12);
  }
  return 0;
}
