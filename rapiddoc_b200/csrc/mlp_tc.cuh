// PPLCNetV4 channel mixer fused into ONE kernel (fp16 / tcgen05 mode):
//     out = W2 * gelu(W1 * x + b1) + b2 (+ x)            channel_conv1 -> GELU -> channel_conv2 (+ residual),
//                                                         rec_lcnetv4.py:210-224,229-235 (BatchNorm folded)
// The 2C-wide intermediate — the largest tensor of every block — never leaves the SM: per 128-row tile
//   TMA (128B swizzle) -> A tile -> tcgen05.mma (N = 2C) -> TMEM acc1 -> 16 epilogue warps: tcgen05.ld, +b1, GELU, fp16 ->
//   shared memory as CHANNEL PLANES (no-swizzle K-major operand: 16-byte chunk c of row r at c*2048 + r*16, conflict-free
//   stores) -> tcgen05.mma (K = 2C, N = Cout) -> TMEM acc2 -> epilogue: +b2, +residual, fp16 store.
// acc1 and the plane buffer are double-buffered and the epilogue warps run E1(i+1) before E2(i), so the tensor pipe works on
// GEMM1(i+1) / GEMM2(i) while the SIMT side is busy with the (dominant) GELU epilogue.
// HBM traffic per block: x read once (+ once more from L2 for the residual) and out written once: 3*C*2 B per pixel instead of 9*C*2.
#pragma once
#include "gemm_tc.cuh"
#include "stem_planar.cuh"   // planar_fill_w, umma_f16_lh

namespace rdb {

struct MlpArgs {
  long long M; int tiles;
  const __half *w1, *w2;          // [2C][C], [COUT][2C]
  const float *b1, *b2;
  const __half* res;              // residual rows [M, COUT] (the block input) or null
  __half* out;                    // [M, COUT]
  int act;                        // ACT_GELU (exact erf) or ACT_GELUF
  const __half *w1p, *w2p;        // mlp_big: weight slices pre-arranged as swizzled shared-memory images ([slice][k-block][row][64])
  int dbg;                        // RDB_MLPBIG_DBG=1 (timing experiment, wrong results): fetch the weight slices for the first tile only
  long long* marks;               // RDB_MLPBIG_MARKS=1: clock64 trace of CTA 0 ([0,128): epilogue thread 0, [128,256): MMA thread)
};

template <int C, int COUT>
struct MlpCfg {
  static constexpr int N1 = 2 * C, K2 = 2 * C;
  static constexpr int KB1 = (C + 63) / 64, KB2 = (K2 + 63) / 64;
  static constexpr int SA = C <= 48 ? 3 : 1;                       // A stages (one stage = KB1 swizzled k-blocks of 128 x 64 halves)
  static constexpr int A_STAGE = KB1 * 16384;
  static constexpr int P2 = K2 / 8, A2BUF = P2 * 2048;             // planes of the intermediate, one 128-row buffer
  static constexpr int oA = 0;
  static constexpr int oA2 = oA + SA * A_STAGE;
  static constexpr int oW1 = oA2 + 2 * A2BUF;
  static constexpr int oW2 = oW1 + KB1 * N1 * 128;
  static constexpr int oB = oW2 + KB2 * COUT * 128;                // b1[N1] b2[COUT]
  static constexpr int oBAR = (oB + (N1 + COUT) * 4 + 15) / 16 * 16;
  static constexpr int kBars = 2 * SA + 2 + 2 + 2 + 2 + 1 + 1;
  static constexpr int kSmem = oBAR + kBars * 8 + 16 + 1024;
  static constexpr int T1 = 0, T2 = 2 * N1;                        // TMEM: acc1[2] x N1, acc2 x COUT
  static_assert(T2 + COUT <= 512, "mlp_tc: TMEM");
  static_assert(kSmem <= 227 * 1024, "mlp_tc: shared memory");
  static_assert(C % 16 == 0 && COUT % 16 == 0, "mlp_tc: channel counts must be multiples of 16");
};

constexpr int kMlpEpi = 512;                      // 16 epilogue warps; warp 16 = TMA producer, warp 17 = MMA issuer
constexpr int kMlpThreads = kMlpEpi + 64;

template <int C, int COUT, int ACT>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_tc_kernel(const __grid_constant__ CUtensorMap tmA, const MlpArgs g) {
  using S = MlpCfg<C, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = RDB_ALIGNED_SMEM(smem_raw);
  float* sb1 = reinterpret_cast<float*>(sm + S::oB);
  float* sb2 = sb1 + S::N1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::oBAR);
  uint64_t* a_full = bars;                         // [SA]  TMA bytes landed
  uint64_t* a_empty = a_full + S::SA;              // [SA]  GEMM1 retired (tcgen05.commit)
  uint64_t* acc1_full = a_empty + S::SA;           // [2]   GEMM1 retired
  uint64_t* acc1_empty = acc1_full + 2;            // [2]   512 epilogue threads drained acc1
  uint64_t* a2_full = acc1_empty + 2;              // [2]   512 epilogue threads wrote the planes
  uint64_t* a2_empty = a2_full + 2;                // [2]   GEMM2 retired
  uint64_t* acc2_full = a2_empty + 2;              // [1]
  uint64_t* acc2_empty = acc2_full + 1;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tmA);
    for (int s = 0; s < S::SA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc1_full[s], 1); tc::mbar_init(&acc1_empty[s], kMlpEpi); tc::mbar_init(&a2_full[s], kMlpEpi); tc::mbar_init(&a2_empty[s], 1); }
    tc::mbar_init(acc2_full, 1); tc::mbar_init(acc2_empty, kMlpEpi);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid < kStemThreads) {
    planar_fill_w(sm + S::oW1, g.w1, S::N1, S::N1, 1, C, C);
    planar_fill_w(sm + S::oW2, g.w2, COUT, COUT, 1, S::K2, S::K2);
    for (int i = tid; i < S::N1; i += kStemThreads) sb1[i] = g.b1[i];
    for (int i = tid; i < COUT; i += kStemThreads) sb2[i] = g.b2[i];
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_local = g.tiles > (int)blockIdx.x ? (g.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kMlpEpi / 32) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int s = i % S::SA;
        if (i >= S::SA) tc::mbar_wait(&a_empty[s], (uint32_t)((i / S::SA) & 1) ^ 1u);
        tc::mbar_expect_tx(&a_full[s], (uint32_t)S::A_STAGE);
        const int row0 = (blockIdx.x + i * gridDim.x) * 128;
#pragma unroll
        for (int kb = 0; kb < S::KB1; ++kb) tc::tma_load_2d(sm + S::oA + s * S::A_STAGE + kb * 16384, &tmA, &a_full[s], kb * 64, row0);
      }
    }
  } else if (warp == kMlpEpi / 32 + 1) {
    // ================= MMA issuer =================
    if (lane == 0 && n_local > 0) {
      constexpr uint32_t hiSw = (1024u >> 4) | (1u << 14) | (2u << 29);      // 128-byte swizzle operand
      constexpr uint32_t hiRow = (128u >> 4) | (1u << 14);                    // planar operand
      const uint32_t loA = (tc::smem_u32(sm + S::oA) >> 4) | (1u << 16);
      const uint32_t loW1 = (tc::smem_u32(sm + S::oW1) >> 4) | (1u << 16), loW2 = (tc::smem_u32(sm + S::oW2) >> 4) | (1u << 16);
      const uint32_t loA2 = (tc::smem_u32(sm + S::oA2) >> 4) | ((2048u >> 4) << 16);
      constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(S::N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto g1 = [&](int i) {
        const int s = i % S::SA, b = i & 1;
        if (i >= 2) tc::mbar_wait(&acc1_empty[b], (uint32_t)((i >> 1) & 1) ^ 1u);
        tc::mbar_wait(&a_full[s], (uint32_t)((i / S::SA) & 1));
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks)
          umma_f16_lh(tmem_base + (uint32_t)(S::T1 + b * S::N1), loA + (uint32_t)((s * S::A_STAGE + (ks >> 2) * 16384) >> 4) + 2u * (ks & 3), hiSw,
                      loW1 + (uint32_t)((ks >> 2) * S::N1 * 8 + 2 * (ks & 3)), hiSw, idesc1, ks != 0 ? 1u : 0u);
        tc::umma_commit(&a_empty[s]);
        tc::umma_commit(&acc1_full[b]);
      };
      g1(0);
      for (int i = 0; i < n_local; ++i) {
        if (i + 1 < n_local) g1(i + 1);
        const int b = i & 1;
        tc::mbar_wait(&a2_full[b], (uint32_t)((i >> 1) & 1));
        if (i >= 1) tc::mbar_wait(acc2_empty, (uint32_t)(i & 1) ^ 1u);
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < S::K2 / 16; ++ks)
          umma_f16_lh(tmem_base + (uint32_t)S::T2, loA2 + (uint32_t)((b * S::A2BUF + 2 * ks * 2048) >> 4), hiRow,
                      loW2 + (uint32_t)((ks >> 2) * COUT * 8 + 2 * (ks & 3)), hiSw, idesc2, ks != 0 ? 1u : 0u);
        tc::umma_commit(&a2_empty[b]);
        tc::umma_commit(acc2_full);
      }
    }
  } else {
    // ================= 16 epilogue warps =================
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const int r = q * 32 + lane;
    auto wait = [&](uint64_t* b, uint32_t ph) { tc::mbar_wait(b, ph); __syncwarp(); tc::tc_fence_after(); };
    auto e1 = [&](int i) {       // acc1 -> +b1 -> GELU -> planes
      const int b = i & 1;
      wait(&acc1_full[b], (uint32_t)((i >> 1) & 1));
      if (i >= 2) wait(&a2_empty[b], (uint32_t)((i >> 1) & 1) ^ 1u);
      uint8_t* dst = sm + S::oA2 + b * S::A2BUF + r * 16;
#pragma unroll
      for (int c0 = sub * 16; c0 < S::N1; c0 += 64) {
        uint32_t rr[16];
        tc::tmem_ld16(tq + (uint32_t)(S::T1 + b * S::N1 + c0), rr);
        tc::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = apply_act<ACT>(__uint_as_float(rr[k]) + sb1[c0 + k]);
        uint4 u0, u1;
        __half2* h0 = reinterpret_cast<__half2*>(&u0); __half2* h1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
        for (int k = 0; k < 4; ++k) { h0[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]); h1[k] = __floats2half2_rn(v[8 + 2 * k], v[9 + 2 * k]); }
        *reinterpret_cast<uint4*>(dst + (c0 >> 3) * 2048) = u0;
        *reinterpret_cast<uint4*>(dst + ((c0 >> 3) + 1) * 2048) = u1;
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc1_empty[b]);
      tc::fence_proxy_async();
      tc::mbar_arrive(&a2_full[b]);
    };
    auto e2 = [&](int i) {       // acc2 -> +b2 (+res) -> global
      wait(acc2_full, (uint32_t)(i & 1));
      const long long row = (long long)(blockIdx.x + i * gridDim.x) * 128 + r;
      const bool ok = row < g.M;
#pragma unroll
      for (int c0 = sub * 16; c0 < COUT; c0 += 64) {
        uint32_t rr[16];
        tc::tmem_ld16(tq + (uint32_t)(S::T2 + c0), rr);
        tc::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(rr[k]) + sb2[c0 + k];
        if (ok) {
          if (g.res != nullptr) {
            float a[8], b2[8];
            Vec8<__half>::load(g.res + row * COUT + c0, a);
            Vec8<__half>::load(g.res + row * COUT + c0 + 8, b2);
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] += a[k]; v[8 + k] += b2[k]; }
          }
          float a[8], b2[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) { a[k] = v[k]; b2[k] = v[8 + k]; }
          Vec8<__half>::store(g.out + row * COUT + c0, a);
          Vec8<__half>::store(g.out + row * COUT + c0 + 8, b2);
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc2_empty);
    };
    if (n_local > 0) e1(0);
    for (int i = 0; i < n_local; ++i) {
      if (i + 1 < n_local) e1(i + 1);
      e2(i);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

template <int C, int COUT>
inline void launch_mlp_tc(Ctx& cx, const __half* x, long long M, const Tensor& w1, const Tensor& b1, const Tensor& w2, const Tensor& b2, const __half* res,
                          __half* out, int act) {
  using S = MlpCfg<C, COUT>;
  MlpArgs a{};
  a.M = M; a.tiles = (int)((M + 127) / 128);
  a.w1 = w1.h; a.w2 = w2.h; a.b1 = b1.d; a.b2 = b2.d; a.res = res; a.out = out; a.act = act;
  CUtensorMap mA = tc::make_map(x, M, C, C, 64, 128);
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  cx.begin("mlp_tc[M=" + std::to_string(M) + ",C=" + std::to_string(C) + ",N=" + std::to_string(COUT) + ",res=" + (res ? "1" : "0") + "]");
  if (act == ACT_GELU) {
    auto k = mlp_tc_kernel<C, COUT, ACT_GELU>;
    static bool done[rdb::kMaxDevices] = {};
    if (rdb::first_on_device(done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
    k<<<grid, kMlpThreads, S::kSmem, cx.st>>>(mA, a);
  } else {
    auto k = mlp_tc_kernel<C, COUT, ACT_GELUF>;
    static bool done[rdb::kMaxDevices] = {};
    if (rdb::first_on_device(done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
    k<<<grid, kMlpThreads, S::kSmem, cx.st>>>(mA, a);
  }
  cx.end();
}


// ============================================================================================================================
// Channel mixer for C = 192 (stage 3 of the det backbone, the dominant blocks of the recogniser): W1 (2C x C) and W2 (C x 2C)
// are 147 KB each — they cannot stay resident, so the hidden dimension is processed in NCH chunks of 128 columns and the weight
// slices stream through shared memory by TMA (from L2, shared by all CTAs):
//   per chunk h:  acc1 = x * W1[h]^T        (K = C,   N = 128)   -> E1: +b1, GELU -> 16 planes of the intermediate
//                 acc2 += planes * W2[:,h]^T (K = 128, N = Cout)
//   after the last chunk: E2: +b2 +residual -> store.
// acc1 and the planes are double-buffered; the weight slices have WS (1 or 2) stages.  The slices are pre-arranged in global
// memory as swizzled smem images and move with one bulk copy each.  Measured (rec, M = 122880 per launch, 6.5 tiles per CTA):
// three 128-column slices single-buffered 94 us; six 64-column slices double-buffered 106 us; strided TMA boxes instead of
// images 95 us; weight fetches switched off after the first tile (RDB_MLPBIG_DBG) 92 us — so weight streaming is NOT the bound.
// The clock64 trace (RDB_MLPBIG_MARKS) shows per tile: 3 x 3.5 k cycles of GELU epilogue, ~4 k of pipeline waits and 8 k cycles
// in the final epilogue, whose 32-byte-per-lane loads/stores at a 384-byte row pitch cost one LSU wavefront per lane.  Next:
// residual from the x tile already in smem and a TMA store of a swizzled staging tile (needs the 64-column configuration to
// free 48 KB).  Against the two unfused GEMMs (134 us) the kernel is 1.4x faster as it stands.
// one contiguous global -> shared bulk copy (no tensor map: the source already is the shared-memory image)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(tc::smem_u32(bar))
               : "memory");
}

// rows [row0, row0+nrows) x K columns [k0, k0+kcols) of a row-major [rows][ld] fp16 matrix -> swizzled K-major image
// [kcols/64 k-blocks][nrows][64 halves] (16-byte chunk c of row n at n*128 + ((c ^ (n&7)) << 4)): what TMA SWIZZLE_128B would
// have written, so a plain bulk copy of the image feeds tcgen05.mma.  A strided TMA box of the same slice costs one 128-byte
// segment per row (profiles/r01_tma_microbench.txt); the image moves as one contiguous transfer.
static __global__ void pack_swizzled_kernel(const __half* __restrict__ w, int ld, int row0, int nrows, int k0, int kcols, __half* __restrict__ dst) {
  const int chunks = (kcols / 64) * nrows * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += gridDim.x * blockDim.x) {
    const int c = i & 7, n = (i >> 3) % nrows, kb = (i >> 3) / nrows;
    __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = w[(size_t)(row0 + n) * ld + k0 + kb * 64 + c * 8 + j];
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(dst) + ((size_t)kb * nrows + n) * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(v);
  }
}

template <int C, int COUT, int NCH, int WS>
struct MlpBigCfg {
  static constexpr int NC = 2 * C / NCH;                           // hidden columns per chunk
  static_assert(NC % 64 == 0 && NC * NCH == 2 * C, "mlp_big: chunk width must be a multiple of 64");
  static constexpr int KB1 = C / 64, KB2 = NC / 64;
  static constexpr int A_BYTES = KB1 * 16384;                      // x tile: KB1 swizzled k-blocks of 128 x 64 halves
  static constexpr int W1_BYTES = KB1 * NC * 128;                  // W1 slice [128 rows][C]
  static constexpr int W2_BYTES = KB2 * COUT * 128;                // W2 slice [COUT rows][128]
  static constexpr int P2 = NC / 8, A2BUF = P2 * 2048;
  static constexpr int oA = 0, oW1 = oA + A_BYTES, oW2 = oW1 + WS * W1_BYTES, oA2 = oW2 + WS * W2_BYTES;
  static constexpr int oB = oA2 + 2 * A2BUF;
  static constexpr int oBAR = (oB + (2 * C + COUT) * 4 + 15) / 16 * 16;
  static constexpr int kBars = 20;
  static constexpr int kSmem = oBAR + kBars * 8 + 16 + 1024;
  static constexpr int T1 = 0, T2 = 2 * NC;
  static_assert(T2 + COUT <= 512, "mlp_big: TMEM");
  static_assert(kSmem <= 227 * 1024, "mlp_big: shared memory");
  static_assert(C % 64 == 0 && COUT % 16 == 0 && COUT <= 256, "mlp_big: shapes");
};

template <int C, int COUT, int NCH, int WS, int ACT>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_big_kernel(const __grid_constant__ CUtensorMap tmA, const MlpArgs g) {
  using S = MlpBigCfg<C, COUT, NCH, WS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = RDB_ALIGNED_SMEM(smem_raw);
  float* sb1 = reinterpret_cast<float*>(sm + S::oB);
  float* sb2 = sb1 + 2 * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::oBAR);
  uint64_t* a_full = bars + 0;  uint64_t* a_empty = bars + 1;
  uint64_t* w1_full = bars + 2;     // [2]
  uint64_t* w1_empty = bars + 4;    // [2]
  uint64_t* w2_full = bars + 6;     // [2]
  uint64_t* w2_empty = bars + 8;    // [2]
  uint64_t* acc1_full = bars + 10;  // [2]
  uint64_t* acc1_empty = bars + 12; // [2] 512 arrivals
  uint64_t* a2_full = bars + 14;    // [2] 512 arrivals
  uint64_t* a2_empty = bars + 16;   // [2]
  uint64_t* acc2_full = bars + 18;
  uint64_t* acc2_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tmA);
    for (int s = 0; s < 10; ++s) tc::mbar_init(&bars[s], 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc1_full[s], 1); tc::mbar_init(&acc1_empty[s], kMlpEpi); tc::mbar_init(&a2_full[s], kMlpEpi); tc::mbar_init(&a2_empty[s], 1); }
    tc::mbar_init(acc2_full, 1); tc::mbar_init(acc2_empty, kMlpEpi);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * C; i += kMlpThreads) sb1[i] = g.b1[i];
  for (int i = tid; i < COUT; i += kMlpThreads) sb2[i] = g.b2[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_local = g.tiles > (int)blockIdx.x ? (g.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total = n_local * NCH;      // chunk counter c = i*NCH + h
  int n_mark = 0;
  auto mark = [&](int who) { if (g.marks != nullptr && blockIdx.x == 0 && n_mark < 128) g.marks[who * 128 + n_mark++] = clock64(); };

  if (warp == kMlpEpi / 32) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int c = 0; c < total; ++c) {
        const int i = c / NCH, h = c % NCH;
        if (h == 0) {
          if (i >= 1) tc::mbar_wait(a_empty, (uint32_t)(i & 1) ^ 1u);
          tc::mbar_expect_tx(a_full, (uint32_t)S::A_BYTES);
          const int row0 = (blockIdx.x + i * gridDim.x) * 128;
#pragma unroll
          for (int kb = 0; kb < S::KB1; ++kb) tc::tma_load_2d(sm + S::oA + kb * 16384, &tmA, a_full, kb * 64, row0);
        }
        const int ws = WS == 2 ? (c & 1) : 0;
        const uint32_t wpar = (uint32_t)((WS == 2 ? (c >> 1) : c) & 1);
        if (c >= WS) tc::mbar_wait(&w1_empty[ws], wpar ^ 1u);
        if (g.dbg && c >= NCH) tc::mbar_arrive(&w1_full[ws]);
        else {
          tc::mbar_expect_tx(&w1_full[ws], (uint32_t)S::W1_BYTES);
          bulk_load(sm + S::oW1 + ws * S::W1_BYTES, reinterpret_cast<const uint8_t*>(g.w1p) + (size_t)h * S::W1_BYTES, (uint32_t)S::W1_BYTES, &w1_full[ws]);
        }
        if (c >= WS) tc::mbar_wait(&w2_empty[ws], wpar ^ 1u);
        if (g.dbg && c >= NCH) tc::mbar_arrive(&w2_full[ws]);
        else {
          tc::mbar_expect_tx(&w2_full[ws], (uint32_t)S::W2_BYTES);
          bulk_load(sm + S::oW2 + ws * S::W2_BYTES, reinterpret_cast<const uint8_t*>(g.w2p) + (size_t)h * S::W2_BYTES, (uint32_t)S::W2_BYTES, &w2_full[ws]);
        }
      }
    }
  } else if (warp == kMlpEpi / 32 + 1) {
    // ================= MMA issuer =================
    if (lane == 0 && total > 0) {
      constexpr uint32_t hiSw = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t hiRow = (128u >> 4) | (1u << 14);
      const uint32_t loA = (tc::smem_u32(sm + S::oA) >> 4) | (1u << 16);
      const uint32_t loW1 = (tc::smem_u32(sm + S::oW1) >> 4) | (1u << 16), loW2 = (tc::smem_u32(sm + S::oW2) >> 4) | (1u << 16);
      const uint32_t loA2 = (tc::smem_u32(sm + S::oA2) >> 4) | ((2048u >> 4) << 16);
      constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(S::NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto g1 = [&](int c) {
        const int i = c / NCH, h = c % NCH, b = c & 1;
        mark(1);
        if (c >= 2) tc::mbar_wait(&acc1_empty[b], (uint32_t)((c >> 1) & 1) ^ 1u);
        if (h == 0) tc::mbar_wait(a_full, (uint32_t)(i & 1));
        const int ws = WS == 2 ? b : 0;
        tc::mbar_wait(&w1_full[ws], (uint32_t)((WS == 2 ? (c >> 1) : c) & 1));
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < C / 16; ++ks)
          umma_f16_lh(tmem_base + (uint32_t)(S::T1 + b * S::NC), loA + (uint32_t)(((ks >> 2) * 16384) >> 4) + 2u * (ks & 3), hiSw,
                      loW1 + (uint32_t)((ws * S::W1_BYTES) >> 4) + (uint32_t)((ks >> 2) * S::NC * 8 + 2 * (ks & 3)), hiSw, idesc1, ks != 0 ? 1u : 0u);
        tc::umma_commit(&w1_empty[ws]);
        if (h == NCH - 1) tc::umma_commit(a_empty);
        tc::umma_commit(&acc1_full[b]);
        mark(1);
      };
      g1(0);
      for (int c = 0; c < total; ++c) {
        if (c + 1 < total) g1(c + 1);
        const int i = c / NCH, h = c % NCH, b = c & 1;
        tc::mbar_wait(&a2_full[b], (uint32_t)((c >> 1) & 1));
        const int ws = WS == 2 ? b : 0;
        mark(1);
        tc::mbar_wait(&w2_full[ws], (uint32_t)((WS == 2 ? (c >> 1) : c) & 1));
        if (h == 0 && i >= 1) tc::mbar_wait(acc2_empty, (uint32_t)(i & 1) ^ 1u);
        tc::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < S::NC / 16; ++ks)
          umma_f16_lh(tmem_base + (uint32_t)S::T2, loA2 + (uint32_t)((b * S::A2BUF + 2 * ks * 2048) >> 4), hiRow,
                      loW2 + (uint32_t)((ws * S::W2_BYTES) >> 4) + (uint32_t)((ks >> 2) * COUT * 8 + 2 * (ks & 3)), hiSw, idesc2, (h != 0 || ks != 0) ? 1u : 0u);
        tc::umma_commit(&w2_empty[ws]);
        tc::umma_commit(&a2_empty[b]);
        if (h == NCH - 1) tc::umma_commit(acc2_full);
        mark(1);
      }
    }
  } else {
    // ================= 16 epilogue warps =================
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const int r = q * 32 + lane;
    auto wait = [&](uint64_t* b, uint32_t ph) { tc::mbar_wait(b, ph); __syncwarp(); tc::tc_fence_after(); };
    auto e1 = [&](int c) {
      const int h = c % NCH, b = c & 1;
      if (tid == 0) mark(0);
      wait(&acc1_full[b], (uint32_t)((c >> 1) & 1));
      if (tid == 0) mark(0);
      if (c >= 2) wait(&a2_empty[b], (uint32_t)((c >> 1) & 1) ^ 1u);
      uint8_t* dst = sm + S::oA2 + b * S::A2BUF + r * 16;
      const float* bias = sb1 + h * S::NC;
#pragma unroll
      for (int c0 = sub * 16; c0 < S::NC; c0 += 64) {
        uint32_t rr[16];
        tc::tmem_ld16(tq + (uint32_t)(S::T1 + b * S::NC + c0), rr);
        tc::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + c0 + 4 * k4);
          v[4 * k4 + 0] = apply_act<ACT>(__uint_as_float(rr[4 * k4 + 0]) + bb.x); v[4 * k4 + 1] = apply_act<ACT>(__uint_as_float(rr[4 * k4 + 1]) + bb.y);
          v[4 * k4 + 2] = apply_act<ACT>(__uint_as_float(rr[4 * k4 + 2]) + bb.z); v[4 * k4 + 3] = apply_act<ACT>(__uint_as_float(rr[4 * k4 + 3]) + bb.w);
        }
        uint4 u0, u1;
        __half2* h0 = reinterpret_cast<__half2*>(&u0); __half2* h1 = reinterpret_cast<__half2*>(&u1);
#pragma unroll
        for (int k = 0; k < 4; ++k) { h0[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]); h1[k] = __floats2half2_rn(v[8 + 2 * k], v[9 + 2 * k]); }
        *reinterpret_cast<uint4*>(dst + (c0 >> 3) * 2048) = u0;
        *reinterpret_cast<uint4*>(dst + ((c0 >> 3) + 1) * 2048) = u1;
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc1_empty[b]);
      tc::fence_proxy_async();
      tc::mbar_arrive(&a2_full[b]);
      if (tid == 0) mark(0);
    };
    auto e2 = [&](int i) {
      if (tid == 0) mark(0);
      constexpr int NE2 = (COUT / 16 + 3) / 4;           // 16-column chunks per warp
      const long long row = (long long)(blockIdx.x + i * gridDim.x) * 128 + r;
      const bool ok = row < g.M;
      // the residual rows do not depend on the accumulators: fetch them before waiting for GEMM2
      uint4 rs[NE2][2];
#pragma unroll
      for (int k = 0; k < NE2; ++k) {
        const int c0 = sub * 16 + 64 * k;
        rs[k][0] = rs[k][1] = make_uint4(0, 0, 0, 0);
        if (ok && g.res != nullptr && c0 < COUT) {
          rs[k][0] = __ldg(reinterpret_cast<const uint4*>(g.res + row * COUT + c0));
          rs[k][1] = __ldg(reinterpret_cast<const uint4*>(g.res + row * COUT + c0 + 8));
        }
      }
      wait(acc2_full, (uint32_t)(i & 1));
      if (tid == 0) mark(0);
      uint32_t rr[NE2][16];
#pragma unroll
      for (int k = 0; k < NE2; ++k)
        if (sub * 16 + 64 * k < COUT) tc::tmem_ld16(tq + (uint32_t)(S::T2 + sub * 16 + 64 * k), rr[k]);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(acc2_empty);                        // the accumulator is in registers: GEMM2 of the next tile may start
#pragma unroll
      for (int k = 0; k < NE2; ++k) {
        const int c0 = sub * 16 + 64 * k;
        if (c0 >= COUT || !ok) continue;
        const __half2* h0 = reinterpret_cast<const __half2*>(&rs[k][0]);
        const __half2* h1 = reinterpret_cast<const __half2*>(&rs[k][1]);
        float a[8], b2[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
          const float4 bb0 = *reinterpret_cast<const float4*>(sb2 + c0 + 0), bb1 = *reinterpret_cast<const float4*>(sb2 + c0 + 4);
          (void)bb0; (void)bb1;
          a[2 * j] = __uint_as_float(rr[k][2 * j]) + sb2[c0 + 2 * j] + f0.x;
          a[2 * j + 1] = __uint_as_float(rr[k][2 * j + 1]) + sb2[c0 + 2 * j + 1] + f0.y;
          b2[2 * j] = __uint_as_float(rr[k][8 + 2 * j]) + sb2[c0 + 8 + 2 * j] + f1.x;
          b2[2 * j + 1] = __uint_as_float(rr[k][8 + 2 * j + 1]) + sb2[c0 + 8 + 2 * j + 1] + f1.y;
        }
        Vec8<__half>::store(g.out + row * COUT + c0, a);
        Vec8<__half>::store(g.out + row * COUT + c0 + 8, b2);
      }
      if (tid == 0) mark(0);
    };
    if (total > 0) e1(0);
    for (int c = 0; c < total; ++c) {
      if (c + 1 < total) e1(c + 1);
      if (c % NCH == NCH - 1) e2(c / NCH);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

template <int C, int COUT, int NCH, int WS>
inline void launch_mlp_big(Ctx& cx, const Weights& wts, const std::string& name, const __half* x, long long M, const Tensor& w1, const Tensor& b1, const Tensor& w2, const Tensor& b2, const __half* res,
                           __half* out, int act) {
  using S = MlpBigCfg<C, COUT, NCH, WS>;
  MlpArgs a{};
  a.M = M; a.tiles = (int)((M + 127) / 128);
  a.w1 = w1.h; a.w2 = w2.h; a.b1 = b1.d; a.b2 = b2.d; a.res = res; a.out = out; a.act = act;
  CUtensorMap mA = tc::make_map(x, M, C, C, 64, 128);
  // weight slices as shared-memory images, packed once per model
  bool fresh1, fresh2;
  __half* w1p = wts.derived(name + "pw1.img", (size_t)NCH * S::W1_BYTES / 2, &fresh1);
  __half* w2p = wts.derived(name + "pw2.img", (size_t)NCH * S::W2_BYTES / 2, &fresh2);
  if (fresh1 || fresh2) {
    for (int h = 0; h < NCH; ++h) {
      pack_swizzled_kernel<<<32, 256, 0, cx.st>>>(w1.h, C, h * S::NC, S::NC, 0, C, w1p + (size_t)h * S::W1_BYTES / 2);
      pack_swizzled_kernel<<<32, 256, 0, cx.st>>>(w2.h, 2 * C, 0, COUT, h * S::NC, S::NC, w2p + (size_t)h * S::W2_BYTES / 2);
    }
    RDB_LAUNCH_CHECK();
    RDB_CUDA(cudaStreamSynchronize(cx.st));   // first use only: the other compute lane (another stream) reads the same images
  }
  a.w1p = w1p; a.w2p = w2p;
  a.dbg = sw_debug("RDB_MLPBIG_DBG") != nullptr;
  static const bool want_marks = sw_debug("RDB_MLPBIG_MARKS") != nullptr;
  if (want_marks) { RDB_CUDA(cudaMalloc(&a.marks, 256 * sizeof(long long))); RDB_CUDA(cudaMemset(a.marks, 0, 256 * sizeof(long long))); }
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  cx.begin("mlp_tc[M=" + std::to_string(M) + ",C=" + std::to_string(C) + ",N=" + std::to_string(COUT) + ",res=" + (res ? "1" : "0") + "]");
  if (act == ACT_GELU) {
    auto k = mlp_big_kernel<C, COUT, NCH, WS, ACT_GELU>;
    static bool done[rdb::kMaxDevices] = {};
    if (rdb::first_on_device(done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
    k<<<grid, kMlpThreads, S::kSmem, cx.st>>>(mA, a);
  } else {
    auto k = mlp_big_kernel<C, COUT, NCH, WS, ACT_GELUF>;
    static bool done[rdb::kMaxDevices] = {};
    if (rdb::first_on_device(done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
    k<<<grid, kMlpThreads, S::kSmem, cx.st>>>(mA, a);
  }
  cx.end();
  if (want_marks) {
    long long h[256];
    RDB_CUDA(cudaDeviceSynchronize());
    RDB_CUDA(cudaMemcpy(h, a.marks, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(a.marks);
    for (int wv = 0; wv < 2; ++wv) {
      fprintf(stderr, "mlp_big_marks who=%d:", wv);
      for (int i = 0; i < 128 && h[wv * 128 + i] != 0; ++i) fprintf(stderr, " %lld", h[wv * 128 + i] - h[0]);
      fprintf(stderr, "\n");
    }
  }
}

}  // namespace rdb
