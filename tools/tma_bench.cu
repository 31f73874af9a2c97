// Microbenchmark: sustained TMA (cp.async.bulk.tensor.2d) load rate per SM as a function of box shape,
// swizzle mode, L2 promotion and ring depth.  One producer thread issues boxes into an S-stage smem ring,
// one consumer thread waits for each and releases it (no math).  Build: nvcc -arch=sm_100a tools/tma_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph), "r"(0x989680u) : "memory");
  }
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* b, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"(m), "r"(s32(b)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(1024, 1) bench(const __grid_constant__ CUtensorMap tm, int stages, int box_bytes, int box_rows, int boxes_per_cta, int boxes_k,
                                               int inner_elems, int hint) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (box_bytes + 1023) & ~1023;
  uint64_t* full = (uint64_t*)(sm + (size_t)stages * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* done = empty + stages;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < boxes_per_cta; ++i) {
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect(&full[s], box_bytes);
      const int box = blockIdx.x * boxes_per_cta + i;
      tma2d(sm + (size_t)s * stage_bytes, &tm, &full[s], (box % boxes_k) * inner_elems, (box / boxes_k) * box_rows);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < boxes_per_cta; ++i) {
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    mbar_arrive(done);
  } else if (warp >= 2) {
    // extra warps that only wait on a barrier completing at the very end (like idle epilogue warps)
    uint32_t ok = 0;
    while (!ok) {
      if (hint) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(done)), "r"(0), "r"(0x989680u) : "memory");
      else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(done)), "r"(0) : "memory");
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  const size_t total_bytes = 512ull << 20;  // 512 MiB tensor (>> L2)
  __half* d; CK(cudaMalloc(&d, total_bytes)); CK(cudaMemset(d, 0, total_bytes));
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  struct Cfg { int row_elems, inner, rows, swz, promo, stages, threads, hint; };
  std::vector<Cfg> cfgs;
  // row_elems = K (tensor inner dim / pitch), inner = box inner elems, rows = box rows
  for (int threads : {64, 192, 576})
    for (int hint : {0, 1})
      for (auto c : std::vector<Cfg>{{48, 64, 128, 3, 0, 0}, {96, 32, 128, 2, 0, 0}, {24, 32, 128, 2, 0, 0}, {192, 64, 128, 3, 0, 0}})
        cfgs.push_back({c.row_elems, c.inner, c.rows, c.swz, 3, 8, threads, hint});
  printf("K(pitch) inner rows swz promo stages | boxKB  us   GB/s(valid)  cycles/box(@1.9GHz) rows/cycle/SM\n");
  for (auto c : cfgs) {
    const long long R = (long long)(total_bytes / 2) / c.row_elems;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)c.row_elems, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)c.row_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.inner, (cuuint32_t)c.rows};
    cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle sw = c.swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : (c.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (c.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, (CUtensorMapL2promotion)c.promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int box_bytes = c.inner * 2 * c.rows;
    const int boxes_k = (c.row_elems + c.inner - 1) / c.inner;
    const long long boxes_total = (R / c.rows) * boxes_k;
    const int per_cta = (int)(boxes_total / 148);
    const size_t smem = 1024 + (size_t)c.stages * ((box_bytes + 1023) & ~1023) + c.stages * 16 + 128;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<<<148, c.threads, smem>>>(tm, c.stages, box_bytes, c.rows, per_cta, boxes_k, c.inner, c.hint);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    bench<<<148, c.threads, smem>>>(tm, c.stages, box_bytes, c.rows, per_cta, boxes_k, c.inner, c.hint);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double valid = (double)per_cta * 148 * c.rows * (c.inner < c.row_elems ? c.inner : c.row_elems) * 2.0;
    const double cyc = ms * 1e-3 * 1.9e9 / per_cta;
    printf("thr=%4d hint=%d %5d %5d %4d %3d %5d %6d | %5.1f %7.1f %8.1f %10.0f %8.3f\n", c.threads, c.hint, c.row_elems, c.inner, c.rows, c.swz, c.promo, c.stages, box_bytes / 1024.0, ms * 1e3,
           valid / ms / 1e6, cyc, c.rows / cyc);
  }
  return 0;
}
