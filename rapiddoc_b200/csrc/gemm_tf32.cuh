// tcgen05 tensor-core GEMM on fp32 STORAGE (kind::tf32):  out[M,N] = act(A[M,K] * W[N,K]^T + bias),  A, W, out: fp32.
//
// Why: the convolutions of the ONNX CNNs (SLANet backbone, orientation classifier, seal detector) run as fp32 GEMMs with
// K, N <= 256.  At ~19 FLOP per byte they are compute-bound on the CUDA cores (the SIMT kernel reaches 1.06 TB/s = 0.16 of the
// HBM roofline); on the tensor cores the same fp32 buffers are read once by TMA and multiplied as TF32 (10-bit mantissa inputs,
// fp32 accumulation in TMEM), which puts the op back on the HBM roofline without changing any storage format.  This is the
// opt-in RDB_PREC_TF32 mode of rdb_op_gemm; RDB_PREC_FP32 (SIMT, fp32 math) stays the exact mode.
//
// Structure (one output tile of 128 rows x BN <= 256 columns per CTA, CTAs are small enough that several are resident per SM and
// overlap each other's load / MMA / store phases): warp 4 = TMA producer (A and W k-blocks of 32 floats = one 128-byte swizzle
// span, 4-stage mbarrier ring), warp 5 = TMEM allocation + single-thread tcgen05.mma issue (UMMA M=128, N=BN, K=8 per
// instruction), warps 0-3 = epilogue (tcgen05.ld of their 32 TMEM lanes -> bias / activation -> a padded shared tile ->
// whole 128-byte fp32 row segments).
// Ragged M / N / K edges are zero-filled by TMA out-of-bounds handling.
#pragma once
#include "gemm_tc.cuh"

namespace rdb {
namespace tf32 {

using namespace rdb::tc;

constexpr int kThreads = 192, kMaxStages = 4, kBK = 32;  // 32 floats = 128 bytes = one SWIZZLE_128B span
constexpr int kStagePitch = 36;                          // epilogue staging: 32 rows x 32 columns per warp, row pitch 36 floats (16-byte aligned rows)

struct Args {
  int M, N, K, BN, k_blocks, tmem_cols, stages;
  int tiles_m, tiles_n, acc_stride;      // persistent kernel: tile grid, TMEM columns between the two accumulator stages
  uint32_t idesc;
  const float* bias;
  float* out;
  int ldc, c_off, act;
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kThreads) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t span = 128u, a_bytes = 128u * span;
  const uint32_t b_bytes = ((uint32_t)g.BN * span + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = RDB_ALIGNED_SMEM(smem_raw);
  const int kStages = g.stages;
  float* stage_out = reinterpret_cast<float*>(smem + (size_t)kStages * stage_bytes);        // [4 warps][32][kStagePitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 4 * 32 * kStagePitch);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_row = blockIdx.x * 128, n_row = blockIdx.y * g.BN;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], a_bytes + (uint32_t)g.BN * span);
        tma_load_2d(sa, &tmA, &full_bar[s], kb * kBK, m_row);
        tma_load_2d(sa + a_bytes, &tmB, &full_bar[s], kb * kBK, n_row);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < g.k_blocks; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t da = make_smem_desc(sa, span), db = make_smem_desc(sa + a_bytes, span);
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk)           // K = 8 tf32 = 32 bytes per instruction: +2 in the (addr >> 4) field
          umma_tf32(tmem_base, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), g.idesc, (kb | kk) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (kb == g.k_blocks - 1) umma_commit(tfull_bar);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = output rows m_row + 32 w + lane.  tcgen05.ld hands every thread 16
    // columns of ITS row; written out directly that is 32 row-strided 64-byte pieces per instruction group (measured: the
    // output-heavy small-K GEMMs ran 2x slower than the SIMT kernel).  So each 32 x 32 block goes through a padded shared
    // tile and leaves as whole 128-byte row segments: 8 lanes x float4 per row, 4 rows per store instruction.
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const long long row0 = (long long)m_row + warp * 32;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* tile = stage_out + warp * 32 * kStagePitch;
    const int rr = lane >> 3, cc = (lane & 7) * 4;
    for (int c0 = 0; c0 < g.BN; c0 += 32) {
      if (n_row + c0 >= g.N) break;                    // warp-uniform: the rest of the tile is padding
      uint32_t r0[16], r1[16];
      tmem_ld16(taddr + (uint32_t)c0, r0);
      tmem_ld16(taddr + (uint32_t)(c0 + 16), r1);      // BN is a multiple of 16: the second half may lie beyond BN (allocated, never stored)
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int ca = n_row + c0 + j, cb = ca + 16;
        v[j] = apply_act_rt(__uint_as_float(r0[j]) + ((g.bias != nullptr && ca < g.N) ? __ldg(g.bias + ca) : 0.f), g.act);
        v[16 + j] = apply_act_rt(__uint_as_float(r1[j]) + ((g.bias != nullptr && cb < g.N) ? __ldg(g.bias + cb) : 0.f), g.act);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)      // 128-bit stores, row pitch 144 bytes: the 8 lanes of a quarter-warp land on distinct bank groups
        *reinterpret_cast<float4*>(tile + lane * kStagePitch + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int cbase = n_row + c0 + cc;               // this lane's 4 columns of the block
#pragma unroll
      for (int r = 0; r < 32; r += 4) {
        const long long row = row0 + r + rr;
        if (row < g.M && cbase < g.N) {
          const float4 v = *reinterpret_cast<const float4*>(tile + (r + rr) * kStagePitch + cc);
          float* o = g.out + row * g.ldc + g.c_off + cbase;
          if (cbase + 4 <= g.N) *reinterpret_cast<float4*>(o) = v;
          else { o[0] = v.x; if (cbase + 1 < g.N) o[1] = v.y; if (cbase + 2 < g.N) o[2] = v.z; }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols); }
}

// Persistent form (RDB_TF32=persistent; correct, currently the slower of the two — it needs the 16-warp epilogue of gemm_tc to
// pay off): a CTA walks tiles t = blockIdx.x, + gridDim.x, ...; the TMA ring runs on across tile boundaries and the
// accumulator has TWO TMEM stages, so the epilogue of tile i (TMEM -> shared tile -> 128-byte row segments) overlaps the loads
// and MMAs of tile i + 1, and barrier setup / TMEM allocation are paid once per CTA instead of once per tile.
__global__ void __launch_bounds__(kThreads) gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                        const Args g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t span = 128u, a_bytes = 128u * span;
  const uint32_t b_bytes = ((uint32_t)g.BN * span + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* smem = RDB_ALIGNED_SMEM(smem_raw);
  const int kStages = g.stages;
  float* stage_out = reinterpret_cast<float*>(smem + (size_t)kStages * stage_bytes);        // [4 warps][32][kStagePitch]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + 4 * 32 * kStagePitch);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = g.tiles_m * g.tiles_n;

  if (warp == 4) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m_row = (t / g.tiles_n) * 128, n_row = (t % g.tiles_n) * g.BN;
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], a_bytes + (uint32_t)g.BN * span);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * kBK, m_row);
          tma_load_2d(sa + a_bytes, &tmB, &full_bar[s], kb * kBK, n_row);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * g.acc_stride);
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da = make_smem_desc(sa, span), db = make_smem_desc(sa + a_bytes, span);
#pragma unroll
          for (int kk = 0; kk < kBK / 8; ++kk)
            umma_tf32(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), g.idesc, (kb | kk) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == g.k_blocks - 1) umma_commit(&tfull_bar[as]);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    float* tile = stage_out + warp * 32 * kStagePitch;
    const int rr = lane >> 3, cc = (lane & 7) * 4;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const int m_row = (t / g.tiles_n) * 128, n_row = (t % g.tiles_n) * g.BN;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const long long row0 = (long long)m_row + warp * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * g.acc_stride);
      for (int c0 = 0; c0 < g.BN; c0 += 32) {
        if (n_row + c0 >= g.N) break;
        uint32_t r0[16], r1[16];
        tmem_ld16(taddr + (uint32_t)c0, r0);
        tmem_ld16(taddr + (uint32_t)(c0 + 16), r1);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ca = n_row + c0 + j, cb = ca + 16;
          v[j] = apply_act_rt(__uint_as_float(r0[j]) + ((g.bias != nullptr && ca < g.N) ? __ldg(g.bias + ca) : 0.f), g.act);
          v[16 + j] = apply_act_rt(__uint_as_float(r1[j]) + ((g.bias != nullptr && cb < g.N) ? __ldg(g.bias + cb) : 0.f), g.act);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(tile + lane * kStagePitch + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int cbase = n_row + c0 + cc;
#pragma unroll
        for (int r = 0; r < 32; r += 4) {
          const long long row = row0 + r + rr;
          if (row < g.M && cbase < g.N) {
            const float4 q = *reinterpret_cast<const float4*>(tile + (r + rr) * kStagePitch + cc);
            float* o = g.out + row * g.ldc + g.c_off + cbase;
            if (cbase + 4 <= g.N) *reinterpret_cast<float4*>(o) = q;
            else { o[0] = q.x; if (cbase + 1 < g.N) o[1] = q.y; if (cbase + 2 < g.N) o[2] = q.z; }
          }
        }
        __syncwarp();
      }
      // this warp has drained its lanes of accumulator stage `as`: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols); }
}

// 2-D fp32 row-major [rows, cols], row pitch ld (elements); box = [box_rows, 32] k-major, 128-byte swizzle
inline CUtensorMap make_map_f32(const void* base, long long rows, int cols, int ld, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  RDB_CHECK(((uintptr_t)base & 15) == 0 && (ld * 4) % 16 == 0, "tma: base/pitch must be 16-byte aligned");
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuda: cuTensorMapEncodeTiled (fp32) failed, code " + std::to_string((int)r));
  return m;
}

inline void launch_gemm_tf32(int device, const float* A, int lda, long long M, int K, const float* W, int N, const float* bias, int act, float* out, int ldc,
                             int c_off, cudaStream_t st) {
  RDB_CHECK(M > 0 && M < (1ll << 31) && K > 0 && N > 0, "gemm tf32: bad shape");
  RDB_CHECK(ldc % 4 == 0 && c_off % 4 == 0 && ((uintptr_t)out & 15) == 0, "gemm tf32: output pitch / offset must be multiples of 4 floats");
  Args a{};
  a.M = (int)M; a.N = N; a.K = K;
  const int tiles_n = (N + 255) / 256;
  a.BN = (((N + tiles_n - 1) / tiles_n) + 15) / 16 * 16;
  a.k_blocks = (K + kBK - 1) / kBK;
  a.tmem_cols = 32;
  while (a.tmem_cols < ((a.BN + 31) / 32) * 32) a.tmem_cols *= 2;     // the epilogue reads whole 32-column blocks
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp): c_format F32 (1) at [4,6), a/b_format TF32 (2) at [7,10) / [10,13),
  // both operands K-major, N >> 3 at [17,23), M >> 4 at [24,29)
  a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  a.bias = bias; a.out = out; a.ldc = ldc; a.c_off = c_off; a.act = act;
  const CUtensorMap mA = make_map_f32(A, M, K, lda, 128);
  const CUtensorMap mB = make_map_f32(W, N, K, K, a.BN);
  const uint32_t b_bytes = ((uint32_t)a.BN * 128u + 1023u) & ~1023u;
  const char* mode = sw_get("RDB_TF32");
  if (!(mode && std::string(mode) == "persistent")) {
    // default: one short-lived CTA per tile.  Measured on the SLANet backbone (32 x 488^2): 1075 GB/s over its GEMMs against 986
    // for the persistent kernel below — with 4 epilogue warps per CTA the store side is the bottleneck, and up to five small
    // CTAs per SM give it more warps than two persistent ones.  RDB_TF32=persistent selects the other kernel (A/B).
    a.stages = a.k_blocks < kMaxStages ? a.k_blocks : kMaxStages;          // short K: fewer stages = more CTAs resident per SM
    const size_t smem = (size_t)a.stages * (128 * 128 + b_bytes) + 4 * 32 * kStagePitch * sizeof(float) + 1024 + 256;
    static bool attr[kMaxDevices] = {};
    if (!attr[device]) {
      RDB_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr[device] = true;
    }
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)tiles_n);
    gemm_tf32_kernel<<<grid, kThreads, smem, st>>>(mA, mB, a);
    RDB_LAUNCH_CHECK();
    return;
  }
  // persistent: two accumulator stages of acc_stride TMEM columns each; as many CTAs per SM as TMEM (512 columns) and shared
  // memory allow, at most 2
  a.tiles_m = (int)((M + 127) / 128);
  a.tiles_n = tiles_n;
  a.acc_stride = ((a.BN + 31) / 32) * 32;
  a.tmem_cols = 32;
  while (a.tmem_cols < 2 * a.acc_stride) a.tmem_cols *= 2;
  const int per_sm = a.tmem_cols <= 256 ? 2 : 1;
  const size_t budget = (per_sm == 2 ? 110 : 220) * 1024 - (4 * 32 * kStagePitch * sizeof(float) + 1024 + 256);
  int stages = (int)(budget / (128 * 128 + b_bytes));
  a.stages = stages > kMaxStages ? kMaxStages : (stages < 1 ? 1 : stages);
  const size_t smem = (size_t)a.stages * (128 * 128 + b_bytes) + 4 * 32 * kStagePitch * sizeof(float) + 1024 + 256;
  static bool attr_p[kMaxDevices] = {};
  static int sms[kMaxDevices] = {};
  if (!attr_p[device]) {
    RDB_CUDA(cudaFuncSetAttribute(gemm_tf32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    RDB_CUDA(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
    attr_p[device] = true;
  }
  const long long tiles = (long long)a.tiles_m * a.tiles_n;
  const long long want = (long long)sms[device] * per_sm;
  gemm_tf32_persistent_kernel<<<(unsigned)(tiles < want ? tiles : want), kThreads, smem, st>>>(mA, mB, a);
  RDB_LAUNCH_CHECK();
}

}  // namespace tf32
}  // namespace rdb
