// PPLCNetV4 "large stem" after stem1, fused into ONE persistent kernel (fp16 / tcgen05 mode):
//   e1 --F.pad--> stem2a 2x2 (C1 -> C1/2) --F.pad--> stem2b 2x2 (-> C1)  \
//   e1 --F.pad--> maxpool 2x2 s1 ceil                                     +--> concat (2*C1) --> stem3 3x3 s2 (-> C1)
//   --> stem4 1x1 (-> 2*C1)                                (rec_lcnetv4.py:143-169; every conv + folded BN + ReLU)
// Unfused, these five ops move ~1.2 GB per 16 pages at HALF input resolution (the largest tensors of the whole
// network) through six kernels; fused, a CTA loads one e1 halo tile (19x35 px for an 8x16 output tile) and every
// intermediate lives in shared memory: HBM traffic = e1 read once + the quarter-resolution output written once.
//
// Every conv is an implicit GEMM on the tensor cores.  The A operand is gathered by the CTA's own threads from the
// smem-resident source tile into the K-major 128-byte-swizzle layout tcgen05.mma reads (16-byte chunk c of row r at
// r*128 + ((c ^ (r&7)) << 4), k-block = 64 halves), one elected thread issues the MMAs into a TMEM accumulator, and the
// epilogue (tcgen05.ld -> bias -> ReLU -> border mask -> fp16) writes the NEXT stage's source tile, again in smem.
// Inside a phase M-tiles are double-buffered (A slots + TMEM accumulators): the gather of tile j+1 overlaps MMA j and
// the epilogue of tile j-1; the max-pool runs in the shadow of stem2a's last MMA and the next tile's e1 halo is
// prefetched with cp.async (zero-fill outside the image = the F.pad / conv padding zeros) while stem2b..stem4 run.
#pragma once
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace rdb {

struct StemArgs {
  const __half* e1; int N, H1, W1, e1_pitch;          // stem1 output [N, H1, e1_pitch(px), C1]
  const __half *w2a, *w2b, *w3, *w4;                  // fp16 [CA][4*C1], [C1][4*CA], [C1][9*2*C1], [2*C1][C1]
  const float *b2a, *b2b, *b3, *b4;
  __half* out; int H2, W2;                            // [N, H2, W2, 2*C1]
  int tiles_x, tiles_y, tiles;
  long long* dbg;                                     // RDB_STEM_DBG=1: clock64 marks of CTA 0's second tile [3][128]
};

template <int C1>
struct StemCfg {
  static constexpr int TY = 8, TX = 16;               // output tile (stem3/stem4 resolution), TY*TX = 128 = UMMA M
  static constexpr int CA = (C1 / 2 + 7) / 8 * 8;      // stem2a channels padded to 8 (zero weights/bias in the pad)
  static constexpr int C2 = 2 * C1;
  static constexpr int ER_H = 2 * TY + 3, ER_W = 2 * TX + 3;   // e1 halo tile
  static constexpr int AR_H = 2 * TY + 2, AR_W = 2 * TX + 2;   // stem2a output region
  static constexpr int CR_H = 2 * TY + 1, CR_W = 2 * TX + 1;   // concat region
  static constexpr int K2A = 4 * C1, K2B = 4 * CA, K3 = 9 * C2, K4 = C1;
  static constexpr int KB2A = (K2A + 63) / 64, KB2B = (K2B + 63) / 64, KB3 = (K3 + 63) / 64, KB4 = (K4 + 63) / 64;
  static constexpr int N2A = (CA + 15) / 16 * 16, N2B = (C1 + 15) / 16 * 16, N3 = N2B, N4 = (C2 + 15) / 16 * 16;
  static constexpr int kSlot = 128 * 128;             // one A k-block: 128 rows x 128 B
  static constexpr int kSlots = 4;
  static constexpr int oA = 0;
  static constexpr int oW2A = oA + kSlots * kSlot;
  static constexpr int oW2B = oW2A + KB2A * N2A * 128;
  static constexpr int oW3 = oW2B + KB2B * N2B * 128;
  static constexpr int oW4 = oW3 + KB3 * N3 * 128;
  static constexpr int oE1 = oW4 + KB4 * N4 * 128;
  static constexpr int oAT = oE1 + ER_H * ER_W * C1 * 2;
  static constexpr int oCAT = oAT + AR_H * AR_W * CA * 2;
  static constexpr int oBIAS = oCAT + CR_H * CR_W * C2 * 2;
  static constexpr int oBAR = (oBIAS + (CA + C1 + C1 + C2) * 4 + 15) / 16 * 16;
  static constexpr int kSmem = oBAR + 128 + 1024;      // + alignment slack
  static_assert(KB2A <= 2 && KB2B <= 2 && KB4 == 1, "stem_fused: A double buffers hold two k-blocks each");
  static_assert((N2A * 128) % 1024 == 0 && (N2B * 128) % 1024 == 0 && (N4 * 128) % 1024 == 0, "weight k-blocks must stay 1024-byte aligned");
};

constexpr int kStemThreads = 512;

// weight matrix [rows_real][K_real] fp16 -> smem B tiles [kblocks][rows_pad][64 halves], 128-byte swizzle, zero padded
__device__ __forceinline__ void stem_fill_w(uint8_t* dst, const __half* __restrict__ w, int rows_real, int rows_pad, int K_real, int kblocks) {
  const int chunks = kblocks * rows_pad * 8;
  for (int i = threadIdx.x; i < chunks; i += kStemThreads) {
    const int c = i & 7, n = (i >> 3) % rows_pad, kb = (i >> 3) / rows_pad;
    uint4 u = make_uint4(0, 0, 0, 0);
    __half* h = reinterpret_cast<__half*>(&u);
    if (n < rows_real) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kb * 64 + c * 8 + j;
        if (k < K_real) h[j] = w[n * K_real + k];
      }
    }
    *reinterpret_cast<uint4*>(dst + (size_t)kb * rows_pad * 128 + n * 128 + ((c ^ (n & 7)) << 4)) = u;
  }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t sz = valid ? 16u : 0u;   // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void stem_worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kStemThreads) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int C1>
__global__ void __launch_bounds__(kStemThreads + 32, 1) stem_fused_kernel(const StemArgs g) {
  using S = StemCfg<C1>;
  constexpr int TY = S::TY, TX = S::TX, CA = S::CA, C2 = S::C2;
  constexpr int EPX = C1 * 2, APX = CA * 2, CPX = C2 * 2;    // bytes per pixel of the three smem tiles
  constexpr int ECH = C1 / 8, ACH = CA / 8, CCH = C2 / 8;    // 16-byte chunks per pixel
  constexpr int SUBS = kStemThreads / 128;                   // epilogue warps per TMEM lane quarter
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = RDB_ALIGNED_SMEM(smem_raw);
  uint8_t* sA = sm + S::oA;
  uint8_t* sE1 = sm + S::oE1;
  uint8_t* sAT = sm + S::oAT;
  uint8_t* sCAT = sm + S::oCAT;
  float* sb2a = reinterpret_cast<float*>(sm + S::oBIAS);
  float* sb2b = sb2a + CA;
  float* sb3 = sb2b + C1;
  float* sb4 = sb3 + C1;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + S::oBAR);     // [0,4): MMA-done (tcgen05.commit), one per A slot / accumulator
  uint64_t* ready = bar + 4;                                      // [0,4): A slot filled (one arrival per worker thread)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, sub = warp >> 2;        // TMEM lane quarter of this warp / which 8-column groups it drains
  // gather roles (thread constants): 2x2 convs — row r4 of the M-tile, tap (tky,tkx); stem3 — chunk cc of rows r8, r8+64
  const int r4 = tid >> 2, tky = (tid >> 1) & 1, tkx = tid & 1, tap = tid & 3;
  const int cc8 = tid & 7, r8 = tid >> 3;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { tc::mbar_init(&bar[s], 1); tc::mbar_init(&ready[s], kStemThreads); }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 128);
  if (tid < kStemThreads) {
    stem_fill_w(sm + S::oW2A, g.w2a, CA, S::N2A, S::K2A, S::KB2A);
    stem_fill_w(sm + S::oW2B, g.w2b, C1, S::N2B, S::K2B, S::KB2B);
    stem_fill_w(sm + S::oW3, g.w3, C1, S::N3, S::K3, S::KB3);
    stem_fill_w(sm + S::oW4, g.w4, C2, S::N4, S::K4, S::KB4);
    for (int i = tid; i < CA; i += kStemThreads) sb2a[i] = g.b2a[i];
    for (int i = tid; i < C1; i += kStemThreads) { sb2b[i] = g.b2b[i]; sb3[i] = g.b3[i]; }
    for (int i = tid; i < C2; i += kStemThreads) sb4[i] = g.b4[i];
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
  auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  uint32_t phase = 0;   // bit s = parity of the next completion of bar[s]; every commit is waited exactly once by all threads
  auto wait_bar = [&](int s) { tc::mbar_wait(&bar[s], (phase >> s) & 1u); phase ^= 1u << s; __syncwarp(); tc::tc_fence_after(); };

  int n_mark = 0;
  const int who = tid == 0 ? 0 : (tid == kStemThreads ? 1 : (tid == kStemThreads - 1 ? 2 : -1));
  auto mark = [&](int t) {
    if (g.dbg != nullptr && who >= 0 && blockIdx.x == 0 && t == (int)gridDim.x && n_mark < 128) g.dbg[who * 128 + n_mark++] = clock64();
  };

  // e1 halo tile of tile t -> smem (cp.async; pixels outside the image are zero-filled)
  auto load_e1 = [&](int t) {
    const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
    const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;
    for (int i = tid; i < S::ER_H * S::ER_W * ECH; i += kStemThreads) {
      const int c = i % ECH, p = i / ECH;
      const int py = p / S::ER_W, px = p % S::ER_W;
      const int gy = gy0 + py, gx = gx0 + px;
      const bool ok = gy >= 0 && gy < g.H1 && gx >= 0 && gx < g.W1;
      const __half* src = ok ? g.e1 + (((long long)n * g.H1 + gy) * g.e1_pitch + gx) * C1 + c * 8 : g.e1;
      cp_async16(sE1 + p * EPX + c * 16, src, ok);
    }
  };

  if (warp == kStemThreads / 32) {
    // ================= MMA issuer: one thread follows the workers' "A slot ready" barriers =================
    if (lane == 0) {
      uint32_t rph = 0;
      auto wait_ready = [&](int s) { tc::mbar_wait(&ready[s], (rph >> s) & 1u); rph ^= 1u << s; tc::tc_fence_after(); };
      auto issue = [&](int slot0, const uint8_t* wtile, int n_pad, int k0, int k1, uint32_t acc_col, bool fresh) {
        // K steps [k0, k1) (16 halves each) of an operand pair whose k-blocks are kSlot / n_pad*128 bytes apart
        for (int ks = k0; ks < k1; ++ks) {
          const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA + (slot0 + ks / 4) * S::kSlot), 128) + (uint64_t)(2 * (ks & 3));
          const uint64_t db = tc::make_smem_desc(tc::smem_u32(wtile + (ks / 4) * n_pad * 128), 128) + (uint64_t)(2 * (ks & 3));
          tc::umma_f16(tmem_base + acc_col, da, db, idesc(n_pad), (ks != k0 || !fresh) ? 1u : 0u);
        }
      };
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
        for (int j = 0; j < (S::AR_H * S::AR_W + 127) / 128; ++j) {      // stem2a
          const int b = j & 1;
          wait_ready(b);
          mark(t);
          issue(2 * b, sm + S::oW2A, S::N2A, 0, S::K2A / 16, (uint32_t)(b * 64), true);
          tc::umma_commit(&bar[b]);
          mark(t);
        }
        for (int j = 0; j < (S::CR_H * S::CR_W + 127) / 128; ++j) {      // stem2b
          const int b = j & 1;
          wait_ready(b);
          mark(t);
          issue(2 * b, sm + S::oW2B, S::N2B, 0, S::K2B / 16, (uint32_t)(b * 64), true);
          tc::umma_commit(&bar[b]);
          mark(t);
        }
        for (int kb = 0; kb < S::KB3; ++kb) {                            // stem3: one k-block per slot
          const int sl = kb & 3;
          wait_ready(sl);
          mark(t);
          const int ksteps = (S::K3 - kb * 64) >= 64 ? 4 : (S::K3 - kb * 64) / 16;
          issue(sl, sm + S::oW3 + kb * S::N3 * 128, S::N3, 0, ksteps, 0u, kb == 0);
          tc::umma_commit(&bar[sl]);
          mark(t);
        }
        wait_ready(0);                                                   // stem4
        issue(0, sm + S::oW4, S::N4, 0, (S::K4 + 15) / 16, 64u, true);
        tc::umma_commit(&bar[0]);
      }
    }
  } else {
  if ((int)blockIdx.x < g.tiles) load_e1(blockIdx.x);
  for (int t = blockIdx.x; t < g.tiles; t += gridDim.x) {
    const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
    const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;   // image coords (stem1 resolution) of region pixel (0,0)
    mark(t);
    cp_async_wait_all();
    stem_worker_sync();
    mark(t);

    // ---------------- stem2a: AR_H x AR_W pixels, K = (ky,kx,ci) = 4*C1, source = e1 tile ----------------
    {
      constexpr int M = S::AR_H * S::AR_W, MT = (M + 127) / 128;
      auto epi = [&](int j) {
        const int b = j & 1;
        wait_bar(b);
        const int m = j * 128 + q * 32 + lane;
        const int py = m / S::AR_W, px = m % S::AR_W;
        const bool live = m < M;
        const bool inimg = live && (gy0 + py) >= 0 && (gy0 + py) < g.H1 && (gx0 + px) >= 0 && (gx0 + px) < g.W1;
        for (int c0 = sub * 8; c0 < CA; c0 += SUBS * 8) {
          uint32_t r[8];
          tc::tmem_ld8(tq + (uint32_t)(b * 64 + c0), r);
          tc::tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = inimg ? fmaxf(__uint_as_float(r[k]) + sb2a[c0 + k], 0.f) : 0.f;
          if (live) Vec8<__half>::store(reinterpret_cast<__half*>(sAT + m * APX) + c0, v);
        }
        tc::tc_fence_before();
      };
      int py = r4 / S::AR_W, px = r4 % S::AR_W;
      for (int j = 0; j < MT; ++j) {
        const int b = j & 1;
        {
          uint4 u[ECH];
#pragma unroll
          for (int k = 0; k < ECH; ++k) u[k] = make_uint4(0, 0, 0, 0);
          if (j * 128 + r4 < M) {
            const uint8_t* src = sE1 + ((py + tky) * S::ER_W + px + tkx) * EPX;
#pragma unroll
            for (int k = 0; k < ECH; ++k) u[k] = *reinterpret_cast<const uint4*>(src + k * 16);
          }
#pragma unroll
          for (int k = 0; k < ECH; ++k) {
            const int c = tap * ECH + k;
            *reinterpret_cast<uint4*>(sA + (2 * b + (c >> 3)) * S::kSlot + r4 * 128 + (((c & 7) ^ (r4 & 7)) << 4)) = u[k];
          }
          px += 128 % S::AR_W; py += 128 / S::AR_W;
          if (px >= S::AR_W) { px -= S::AR_W; ++py; }
        }
        tc::fence_proxy_async();     // this thread's generic-proxy smem writes -> visible to the tensor core
        tc::tc_fence_before();       // its TMEM loads of the accumulator about to be overwritten are complete
        mark(t);
        tc::mbar_arrive(&ready[b]);
        mark(t);
        if (j >= 1) epi(j - 1);
        mark(t);
      }
      // max-pool 2x2 s1 (ceil_mode, on the zero-padded e1) -> concat channels [0, C1): runs while the last stem2a MMA drains
      for (int i = tid; i < S::CR_H * S::CR_W * ECH; i += kStemThreads) {
        const int c = i % ECH, p = i / ECH;
        const int py2 = p / S::CR_W, px2 = p % S::CR_W;
        const bool inimg = (gy0 + py2) >= 0 && (gy0 + py2) < g.H1 && (gx0 + px2) >= 0 && (gx0 + px2) < g.W1;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (inimg) {
          const uint8_t* e = sE1 + (py2 * S::ER_W + px2) * EPX + c * 16;
          const uint4 a0 = *reinterpret_cast<const uint4*>(e), a1 = *reinterpret_cast<const uint4*>(e + EPX);
          const uint4 a2 = *reinterpret_cast<const uint4*>(e + S::ER_W * EPX), a3 = *reinterpret_cast<const uint4*>(e + S::ER_W * EPX + EPX);
          const __half2* h0 = reinterpret_cast<const __half2*>(&a0); const __half2* h1 = reinterpret_cast<const __half2*>(&a1);
          const __half2* h2 = reinterpret_cast<const __half2*>(&a2); const __half2* h3 = reinterpret_cast<const __half2*>(&a3);
          __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) ho[k] = __hmax2(__hmax2(h0[k], h1[k]), __hmax2(h2[k], h3[k]));
        }
        *reinterpret_cast<uint4*>(sCAT + p * CPX + c * 16) = o;
      }
      mark(t);
      epi(MT - 1);
      mark(t);
    }
    stem_worker_sync();   // stem2a tile complete; the e1 tile is dead
    mark(t);
    if (t + (int)gridDim.x < g.tiles) load_e1(t + gridDim.x);

    // ---------------- stem2b: CR_H x CR_W pixels, K = 4*CA, source = stem2a tile -> concat channels [C1, 2*C1) ----------------
    {
      constexpr int M = S::CR_H * S::CR_W, MT = (M + 127) / 128;
      auto epi = [&](int j) {
        const int b = j & 1;
        wait_bar(b);
        const int m = j * 128 + q * 32 + lane;
        const int py = m / S::CR_W, px = m % S::CR_W;
        const bool live = m < M;
        const bool inimg = live && (gy0 + py) >= 0 && (gy0 + py) < g.H1 && (gx0 + px) >= 0 && (gx0 + px) < g.W1;
        for (int c0 = sub * 8; c0 < C1; c0 += SUBS * 8) {
          uint32_t r[8];
          tc::tmem_ld8(tq + (uint32_t)(b * 64 + c0), r);
          tc::tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = inimg ? fmaxf(__uint_as_float(r[k]) + sb2b[c0 + k], 0.f) : 0.f;
          if (live) Vec8<__half>::store(reinterpret_cast<__half*>(sCAT + m * CPX) + C1 + c0, v);
        }
        tc::tc_fence_before();
      };
      int py = r4 / S::CR_W, px = r4 % S::CR_W;
      for (int j = 0; j < MT; ++j) {
        const int b = j & 1;
        {
          uint4 u[ACH];
#pragma unroll
          for (int k = 0; k < ACH; ++k) u[k] = make_uint4(0, 0, 0, 0);
          if (j * 128 + r4 < M) {
            const uint8_t* src = sAT + ((py + tky) * S::AR_W + px + tkx) * APX;
#pragma unroll
            for (int k = 0; k < ACH; ++k) u[k] = *reinterpret_cast<const uint4*>(src + k * 16);
          }
#pragma unroll
          for (int k = 0; k < ACH; ++k) {
            const int c = tap * ACH + k;
            *reinterpret_cast<uint4*>(sA + (2 * b + (c >> 3)) * S::kSlot + r4 * 128 + (((c & 7) ^ (r4 & 7)) << 4)) = u[k];
          }
          px += 128 % S::CR_W; py += 128 / S::CR_W;
          if (px >= S::CR_W) { px -= S::CR_W; ++py; }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        mark(t);
        tc::mbar_arrive(&ready[b]);
        mark(t);
        if (j >= 1) epi(j - 1);
        mark(t);
      }
      epi(MT - 1);
      mark(t);
    }
    stem_worker_sync();   // concat tile complete
    mark(t);

    // ---------------- stem3: 3x3 s2 on the concat tile, K = 9*C2 streamed through the 4 A slots ----------------
    {
      constexpr int KC = S::K3 / 8, ROWC = 3 * CCH;
      static_assert(ROWC >= 8, "stem3 gather: at most one kernel-row carry per k-block");
      int ky = 0, rem = cc8;   // K chunk c = kb*8 + cc8 = ky*ROWC + rem
      for (int kb = 0; kb < S::KB3; ++kb) {
        const int s = kb & 3;
        if (kb >= 4) wait_bar(s);
        const bool kvalid = kb * 8 + cc8 < KC;
#pragma unroll
        for (int u2 = 0; u2 < 128 * 8 / kStemThreads; ++u2) {
          const int r = r8 + u2 * (kStemThreads / 8);
          const int oy = r / TX, ox = r % TX;
          uint4 u = make_uint4(0, 0, 0, 0);
          if (kvalid) u = *reinterpret_cast<const uint4*>(sCAT + ((2 * oy + ky) * S::CR_W + 2 * ox) * CPX + rem * 16);
          *reinterpret_cast<uint4*>(sA + s * S::kSlot + r * 128 + ((cc8 ^ (r & 7)) << 4)) = u;
        }
        rem += 8;
        if (rem >= ROWC) { rem -= ROWC; ++ky; }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        mark(t);
        tc::mbar_arrive(&ready[s]);
        mark(t);
      }
      // drain: every outstanding commit is waited once, in issue order (the last one publishes the accumulator)
      for (int kb = (S::KB3 > 4 ? S::KB3 - 4 : 0); kb < S::KB3; ++kb) wait_bar(kb & 3);
      mark(t);
      // epilogue -> A operand of stem4 (K = C1, zero-padded to a multiple of 16) in slot 0
      {
        const int r = q * 32 + lane;
        for (int c0 = sub * 8; c0 < (S::K4 + 15) / 16 * 16; c0 += SUBS * 8) {
          float v[8];
          if (c0 < C1) {
            uint32_t rr[8];
            tc::tmem_ld8(tq + (uint32_t)c0, rr);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(__uint_as_float(rr[k]) + sb3[c0 + k], 0.f);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = 0.f;
          }
          uint4 u;
          __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
          *reinterpret_cast<uint4*>(sA + r * 128 + (((c0 >> 3) ^ (r & 7)) << 4)) = u;
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      mark(t);
      tc::mbar_arrive(&ready[0]);
      // ---------------- stem4: 1x1, K = C1 (issued by the MMA warp) ----------------
      wait_bar(0);
      mark(t);
      {
        const int r = q * 32 + lane;
        const int oy = ty * TY + r / TX, ox = tx * TX + r % TX;
        const bool ok = oy < g.H2 && ox < g.W2;
        __half* op = g.out + (((long long)n * g.H2 + oy) * g.W2 + ox) * C2;
        for (int c0 = sub * 8; c0 < C2; c0 += SUBS * 8) {
          uint32_t rr[8];
          tc::tmem_ld8(tq + (uint32_t)(64 + c0), rr);
          tc::tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(__uint_as_float(rr[k]) + sb4[c0 + k], 0.f);
          if (ok) Vec8<__half>::store(op + c0, v);
        }
        tc::tc_fence_before();
      }
      mark(t);
    }
  }
  cp_async_wait_all();
  }   // workers
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 128); }
}

template <int C1>
inline void launch_stem_fused(Ctx& cx, const Weights& w, const __half* e1, int n, int H1, int W1, int e1_pitch, __half* out, int H2, int W2) {
  using S = StemCfg<C1>;
  StemArgs a{};
  a.e1 = e1; a.N = n; a.H1 = H1; a.W1 = W1; a.e1_pitch = e1_pitch;
  a.w2a = w.get("stem2a.wp").h; a.w2b = w.get("stem2b.wp").h; a.w3 = w.get("stem3.w").h; a.w4 = w.get("stem4.w").h;
  a.b2a = w.get("stem2a.bp").d; a.b2b = w.get("stem2b.b").d; a.b3 = w.get("stem3.b").d; a.b4 = w.get("stem4.b").d;
  a.out = out; a.H2 = H2; a.W2 = W2;
  a.tiles_x = (W2 + S::TX - 1) / S::TX; a.tiles_y = (H2 + S::TY - 1) / S::TY; a.tiles = n * a.tiles_x * a.tiles_y;
  auto k = stem_fused_kernel<C1>;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  static const bool dbg = sw_debug("RDB_STEM_DBG") != nullptr;
  if (dbg) { RDB_CUDA(cudaMalloc(&a.dbg, 3 * 128 * sizeof(long long))); RDB_CUDA(cudaMemset(a.dbg, 0, 3 * 128 * sizeof(long long))); }
  cx.begin("stem_fused[P=" + std::to_string((long long)n * H2 * W2) + "]");
  k<<<grid, kStemThreads + 32, S::kSmem, cx.st>>>(a);
  cx.end();
  if (dbg) {
    long long h[3 * 128];
    RDB_CUDA(cudaDeviceSynchronize());
    RDB_CUDA(cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(a.dbg);
    const long long t0 = h[0];
    for (int wv = 0; wv < 3; ++wv) {
      fprintf(stderr, "stem_dbg who=%d:", wv);
      for (int i = 0; i < 128 && h[wv * 128 + i] != 0; ++i) fprintf(stderr, " %lld", h[wv * 128 + i] - t0);
      fprintf(stderr, "\n");
    }
  }
}

}  // namespace rdb
