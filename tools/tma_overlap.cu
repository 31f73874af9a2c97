// Feasibility probe: does cuTensorMapEncodeTiled accept OVERLAPPING rows (dim-1 stride smaller than the dim-0 extent)
// and does TMA deliver them?  Used to merge the kx taps of a conv into one wider K block (rows = 2 or 3 adjacent pixels).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, __half* out, int n_elems, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(n_elems * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(s32(sm)), "l"(&tm),
                 "r"(s32(&bar)), "r"(0), "r"(c1), "r"(c2), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_elems; i += blockDim.x) out[i] = reinterpret_cast<__half*>(sm)[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  const int C = 16, W = 8, WP = W + 1, H = 4, N = 1;    // NHWC with one pad pixel per row
  std::vector<__half> h((size_t)N * H * WP * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)(i % 2048));
  __half *d, *o; CK(cudaMalloc(&d, h.size() * 2 + 256)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  const int KW = 2, inner = KW * C;                       // a "row" = 2 adjacent pixels = 32 channels
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)WP * C * 2, (cuuint64_t)H * WP * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)inner, 8, 2, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode overlapping-rows map: %s (code %d)\n", r == CUDA_SUCCESS ? "OK" : "REJECTED", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  const int n_elems = inner * 8 * 2;
  CK(cudaMalloc(&o, n_elems * 2));
  k<<<1, 128, 4096>>>(tm, o, n_elems, 0, 1);
  CK(cudaDeviceSynchronize());
  std::vector<__half> got(n_elems); CK(cudaMemcpy(got.data(), o, n_elems * 2, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int y = 0; y < 2; ++y) for (int x = 0; x < 8; ++x) for (int c = 0; c < inner; ++c) {
    float want = (float)((((size_t)(y + 1) * WP + x) * C + c) % 2048);
    float g = __half2float(got[(y * 8 + x) * inner + c]);
    if (g != want) { if (bad < 5) printf("mismatch y=%d x=%d c=%d got %g want %g\n", y, x, c, g, want); ++bad; }
  }
  printf("overlapping rows delivered %s (%d mismatches of %d)\n", bad ? "WRONG" : "correctly", bad, n_elems);
  return 0;
}
