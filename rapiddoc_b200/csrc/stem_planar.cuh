// PPLCNetV4 "large stem" after stem1 (stem2a, stem2b, max-pool, concat, stem3, stem4; rec_lcnetv4.py:143-169) fused into ONE
// persistent kernel whose convolutions are TAP-DECOMPOSED implicit GEMMs read straight out of shared memory:
//
//   * every intermediate of a 16x8 output tile lives in smem as CHANNEL PLANES: plane c holds the 16-byte chunk (8 fp16
//     channels) c of every pixel, pixels contiguous at a 16-byte stride.  In the tcgen05 "no-swizzle, K-major" operand
//     layout an 8-row core matrix is 8 consecutive 16-byte rows (128 B), 8-row groups are SBO bytes apart and the two
//     16-byte K chunks of one MMA are LBO bytes apart — so with LBO = plane size, ANY 16-byte-aligned start address is a
//     valid A operand.  A conv tap (ky,kx) is therefore just the same planes read at start + (ky*pitch + kx) pixels:
//     no im2col, no copies; D += A(tap) * W(tap) accumulates the taps in TMEM.
//   * stem3 is stride 2: the concat tile is stored with even / odd columns de-interleaved (row = [9 even | 9 odd] pixels),
//     an output row of the 16x8 tile is one 8-row core-matrix group and SBO = two concat rows.
//   * the only SIMT work left is the epilogues (tcgen05.ld -> bias -> ReLU -> border mask -> fp16 -> next stage's planes),
//     the 2x2 max-pool and the cp.async halo load (zero-fill outside the image = the F.pad / conv padding zeros).
//
// HBM traffic: e1 read once (+30% halo) and the quarter-resolution output written once, instead of ~1.2 GB per 16 pages.
#pragma once
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "stem_fused.cuh"   // StemArgs, stem worker helpers

namespace rdb {

template <int C1>
struct PlanarCfg {
  static constexpr int TY = 16, TX = 8;                 // output tile (stem3/stem4 resolution); an output row = one 8-row group
  static constexpr int CA = (C1 / 2 + 7) / 8 * 8, C2 = 2 * C1;
  static constexpr int ECH = C1 / 8, ACH = CA / 8, CCH = C2 / 8;
  static constexpr int ECHP = (ECH + 1) / 2 * 2, ACHP = (ACH + 1) / 2 * 2, CCHP = (CCH + 1) / 2 * 2;   // planes incl. a zero plane (K multiple of 16)
  static constexpr int ER_H = 2 * TY + 3, ER_W = 2 * TX + 3;      // e1 halo tile; ER_W is the row pitch of the e1 / stem2a planes
  static constexpr int AR_H = 2 * TY + 2, AR_W = 2 * TX + 2;      // stem2a region actually needed
  static constexpr int CR_H = 2 * TY + 1, CR_W = 2 * TX + 1;      // concat region
  static constexpr int MT2 = ((AR_H - 1) * ER_W + AR_W + 127) / 128;     // M-tiles (128 consecutive pixel indices, pitch ER_W)
  static constexpr int MT3 = ((CR_H - 1) * ER_W + CR_W + 127) / 128;
  static constexpr int EROWS = (MT2 * 128 + ER_W + 1 + 7) / 8 * 8;       // rows a shifted stem2a M-tile may touch
  static constexpr int AROWS = (((MT3 * 128 + ER_W + 1) > MT2 * 128 ? (MT3 * 128 + ER_W + 1) : MT2 * 128) + 7) / 8 * 8;
  static constexpr int EPLANE = EROWS * 16, APLANE = AROWS * 16;
  static constexpr int CPAR = (TX + 1) * 16, CROW = 2 * CPAR, CPLANE = CR_H * CROW;    // concat: [row][parity][TX+1] x 16 B
  static constexpr int A2PLANE = 128 * 16;
  static constexpr int C1P = ECHP * 8, CAP = ACHP * 8, C2P = CCHP * 8;                   // channel counts padded to 16
  static constexpr int K2A = 4 * C1P, K2B = 4 * CAP, K3 = 9 * C2P, K4 = C1P;
  static constexpr int KB2A = (K2A + 63) / 64, KB2B = (K2B + 63) / 64, KB3 = (K3 + 63) / 64, KB4 = (K4 + 63) / 64;
  static constexpr int N2A = (CA + 15) / 16 * 16, N2B = (C1 + 15) / 16 * 16, N3 = N2B, N4 = (C2 + 15) / 16 * 16;
  // TMEM columns
  static constexpr int T2 = 0, T3 = T2 + MT2 * N2A, T4 = T3 + MT3 * N2B, T5 = T4 + N3, TEND = T5 + N4;
  static_assert(TEND <= 512, "stem_planar: TMEM accumulators exceed 512 columns");
  // smem
  static constexpr int oE1 = 0;
#ifndef RDB_STEM_PAD
#define RDB_STEM_PAD 0
#endif
  static constexpr int oAT = oE1 + ECHP * EPLANE + RDB_STEM_PAD;
  static constexpr int oCAT = oAT + ACHP * APLANE;
  static constexpr int oA2 = oCAT + CCHP * CPLANE;
  static constexpr int oTILES_END = oA2 + ECHP * A2PLANE;
  static constexpr int oW2A = (oTILES_END + 1023) / 1024 * 1024;
  static constexpr int oW2B = oW2A + KB2A * N2A * 128;
  static constexpr int oW3 = oW2B + KB2B * N2B * 128;
  static constexpr int oW4 = oW3 + KB3 * N3 * 128;
  static constexpr int oBIAS = oW4 + KB4 * N4 * 128;
  static constexpr int oBAR = (oBIAS + (CA + C1 + C1 + C2) * 4 + 15) / 16 * 16;
  static constexpr int kBars = 4 + MT2 + MT3 + 2;
  static constexpr int kSmem = oBAR + kBars * 8 + 16 + 1024;
  static_assert(kSmem <= 227 * 1024, "stem_planar: shared memory");
};

// weights [rows_real][taps][c_real] fp16 -> smem B tiles, K index = tap*c_pad + ci (zero for ci >= c_real), 128-byte swizzle k-blocks
__device__ __forceinline__ void planar_fill_w(uint8_t* dst, const __half* __restrict__ w, int rows_real, int rows_pad, int taps, int c_real, int c_pad) {
  const int K = taps * c_pad, kblocks = (K + 63) / 64;
  const int chunks = kblocks * rows_pad * 8;
  for (int i = threadIdx.x; i < chunks; i += kStemThreads) {
    const int c = i & 7, n = (i >> 3) % rows_pad, kb = (i >> 3) / rows_pad;
    uint4 u = make_uint4(0, 0, 0, 0);
    __half* h = reinterpret_cast<__half*>(&u);
    if (n < rows_real) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = kb * 64 + c * 8 + j;
        const int tap = k / c_pad, ci = k % c_pad;
        if (k < K && ci < c_real) h[j] = w[(n * taps + tap) * c_real + ci];
      }
    }
    *reinterpret_cast<uint4*>(dst + (size_t)kb * rows_pad * 128 + n * 128 + ((c ^ (n & 7)) << 4)) = u;
  }
}

// K-major, no-swizzle operand: 8-row core matrices of 8 x 16 B; lbo = bytes between the two K chunks, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t planar_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// tcgen05.mma with the two smem descriptors passed as 32-bit halves: the issuing thread only ever adds a constant to the
// low word (start address in 16-byte units; LBO sits above it) — the single-thread issue rate is what bounds this kernel
__device__ __forceinline__ void umma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int C1>
__global__ void __launch_bounds__(kStemThreads + 32, 1) stem_planar_kernel(const StemArgs g) {
  using S = PlanarCfg<C1>;
  constexpr int TY = S::TY, TX = S::TX, CA = S::CA, C2 = S::C2, P = S::ER_W;
  constexpr int SUBS = kStemThreads / 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = RDB_ALIGNED_SMEM(smem_raw);
  uint8_t* sE1 = sm + S::oE1;
  uint8_t* sAT = sm + S::oAT;
  uint8_t* sCAT = sm + S::oCAT;
  uint8_t* sA2 = sm + S::oA2;
  float* sb2a = reinterpret_cast<float*>(sm + S::oBIAS);
  float* sb2b = sb2a + CA;
  float* sb3 = sb2b + C1;
  float* sb4 = sb3 + C1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::oBAR);
  uint64_t* e1_ready = bars + 0; uint64_t* a_ready = bars + 1; uint64_t* cat_ready = bars + 2; uint64_t* a2_ready = bars + 3;
  uint64_t* done2 = bars + 4; uint64_t* done3 = done2 + S::MT2; uint64_t* done4 = done3 + S::MT3; uint64_t* done5 = done4 + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kBars);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, sub = warp >> 2;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) tc::mbar_init(&bars[s], kStemThreads);
    for (int s = 4; s < S::kBars; ++s) tc::mbar_init(&bars[s], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid < kStemThreads) {
    for (int i = tid; i < S::oTILES_END / 16; i += kStemThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);   // incl. the zero planes
    planar_fill_w(sm + S::oW2A, g.w2a, CA, S::N2A, 4, C1, S::C1P);
    planar_fill_w(sm + S::oW2B, g.w2b, C1, S::N2B, 4, CA, S::CAP);
    planar_fill_w(sm + S::oW3, g.w3, C1, S::N3, 9, C2, S::C2P);
    planar_fill_w(sm + S::oW4, g.w4, C2, S::N4, 1, C1, S::C1P);
    for (int i = tid; i < CA; i += kStemThreads) sb2a[i] = g.b2a[i];
    for (int i = tid; i < C1; i += kStemThreads) { sb2b[i] = g.b2b[i]; sb3[i] = g.b3[i]; }
    for (int i = tid; i < C2; i += kStemThreads) sb4[i] = g.b4[i];
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
  int n_mark = 0;
  const int who = tid == 0 ? 0 : (tid == kStemThreads ? 1 : (tid == kStemThreads - 1 ? 2 : -1));
  auto mark = [&](int t) {
    if (g.dbg != nullptr && who >= 0 && blockIdx.x == 0 && t == (int)gridDim.x && n_mark < 128) g.dbg[who * 128 + n_mark++] = clock64();
  };

  if (warp == kStemThreads / 32) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version(1)<<14 | layout<<29
      const uint32_t loE1 = (tc::smem_u32(sE1) >> 4) | ((uint32_t)(S::EPLANE >> 4) << 16);
      const uint32_t loAT = (tc::smem_u32(sAT) >> 4) | ((uint32_t)(S::APLANE >> 4) << 16);
      const uint32_t loCAT = (tc::smem_u32(sCAT) >> 4) | ((uint32_t)(S::CPLANE >> 4) << 16);
      const uint32_t loA2 = (tc::smem_u32(sA2) >> 4) | ((uint32_t)(S::A2PLANE >> 4) << 16);
      constexpr uint32_t hiRow = (128u >> 4) | (1u << 14);                       // planar operand, 8-row groups contiguous
      constexpr uint32_t hiCat = ((uint32_t)(2 * S::CROW) >> 4) | (1u << 14);   // stem3: an 8-row group = one output row = two concat rows apart
      constexpr uint32_t hiW = (1024u >> 4) | (1u << 14) | (2u << 29);           // weights: 128-byte swizzle k-blocks
      const uint32_t loW2A = (tc::smem_u32(sm + S::oW2A) >> 4) | (1u << 16), loW2B = (tc::smem_u32(sm + S::oW2B) >> 4) | (1u << 16);
      const uint32_t loW3 = (tc::smem_u32(sm + S::oW3) >> 4) | (1u << 16), loW4 = (tc::smem_u32(sm + S::oW4) >> 4) | (1u << 16);
      // K step ks of a weight tile with n_pad rows per 64-wide k-block, in 16-byte units
      auto wofs = [](int n_pad, int ks) { return (uint32_t)((ks >> 2) * n_pad * 8 + 2 * (ks & 3)); };
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x, ph ^= 1u) {
        mark(t);
        tc::mbar_wait(e1_ready, ph);
        tc::tc_fence_after();
        mark(t);
        for (int j = 0; j < S::MT2; ++j) {                               // stem2a: 4 taps x ECHP/2 K steps per M-tile
          const uint32_t aj = loE1 + (uint32_t)(j * 128), dj = tmem_base + (uint32_t)(S::T2 + j * S::N2A);
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int h = 0; h < S::ECHP / 2; ++h)
              umma_f16_lh(dj, aj + (uint32_t)(2 * h * (S::EPLANE >> 4) + (tap >> 1) * P + (tap & 1)), hiRow, loW2A + wofs(S::N2A, tap * (S::ECHP / 2) + h), hiW,
                          idesc(S::N2A), (tap | h) != 0 ? 1u : 0u);
          tc::umma_commit(&done2[j]);
        }
        mark(t);
        tc::mbar_wait(a_ready, ph);
        tc::tc_fence_after();
        mark(t);
        for (int j = 0; j < S::MT3; ++j) {                               // stem2b
          const uint32_t aj = loAT + (uint32_t)(j * 128), dj = tmem_base + (uint32_t)(S::T3 + j * S::N2B);
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int h = 0; h < S::ACHP / 2; ++h)
              umma_f16_lh(dj, aj + (uint32_t)(2 * h * (S::APLANE >> 4) + (tap >> 1) * P + (tap & 1)), hiRow, loW2B + wofs(S::N2B, tap * (S::ACHP / 2) + h), hiW,
                          idesc(S::N2B), (tap | h) != 0 ? 1u : 0u);
          tc::umma_commit(&done3[j]);
        }
        mark(t);
        tc::mbar_wait(cat_ready, ph);
        tc::tc_fence_after();
        mark(t);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {                              // stem3: 3x3 stride 2 on the parity-split concat planes
          const int ky = tap / 3, kx = tap % 3;
#pragma unroll
          for (int h = 0; h < S::CCHP / 2; ++h)
            umma_f16_lh(tmem_base + (uint32_t)S::T4, loCAT + (uint32_t)((2 * h * S::CPLANE + ky * S::CROW + (kx & 1) * S::CPAR + (kx >> 1) * 16) >> 4), hiCat,
                        loW3 + wofs(S::N3, tap * (S::CCHP / 2) + h), hiW, idesc(S::N3), (tap | h) != 0 ? 1u : 0u);
        }
        tc::umma_commit(done4);
        mark(t);
        tc::mbar_wait(a2_ready, ph);
        tc::tc_fence_after();
        mark(t);
#pragma unroll
        for (int h = 0; h < S::ECHP / 2; ++h)                            // stem4: 1x1
          umma_f16_lh(tmem_base + (uint32_t)S::T5, loA2 + (uint32_t)(2 * h * (S::A2PLANE >> 4)), hiRow, loW4 + wofs(S::N4, h), hiW, idesc(S::N4), h != 0 ? 1u : 0u);
        tc::umma_commit(done5);
        mark(t);
      }
    }
  } else {
    // ================= workers: halo load, pool, epilogues =================
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    auto wait_done = [&](uint64_t* b, uint32_t ph) { tc::mbar_wait(b, ph); __syncwarp(); tc::tc_fence_after(); };
    // tile-invariant gather tables (registers): which halo pixels / pool outputs this thread owns
    constexpr int NPE = S::ER_H * S::ER_W, NPC = S::CR_H * S::CR_W;
    constexpr int LDI = (NPE * S::ECH + kStemThreads - 1) / kStemThreads, PLI = (NPC * S::ECH + kStemThreads - 1) / kStemThreads;
    int ld_yx[LDI], pl_yx[PLI];
    uint32_t ld_dst[LDI], pl_src[PLI], pl_dst[PLI];
#pragma unroll
    for (int k = 0; k < LDI; ++k) {
      const int i = tid + k * kStemThreads;
      const int p = i % NPE, c = i / NPE, py = p / S::ER_W, px = p % S::ER_W;   // consecutive threads -> consecutive pixels of one plane
      ld_yx[k] = i < NPE * S::ECH ? (py | (px << 8) | (c << 16)) : -1;
      ld_dst[k] = (uint32_t)(c * S::EPLANE + p * 16);
    }
#pragma unroll
    for (int k = 0; k < PLI; ++k) {
      const int i = tid + k * kStemThreads;
      const int p = i % NPC, c = i / NPC, py = p / S::CR_W, px = p % S::CR_W;
      pl_yx[k] = i < NPC * S::ECH ? (py | (px << 8)) : -1;
      pl_src[k] = (uint32_t)(c * S::EPLANE + (py * P + px) * 16);
      pl_dst[k] = (uint32_t)(c * S::CPLANE + py * S::CROW + (px & 1) * S::CPAR + (px >> 1) * 16);
    }
    auto load_e1 = [&](int t) {
      const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
      const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;
      const __half* img = g.e1 + (long long)n * g.H1 * g.e1_pitch * C1;
#pragma unroll
      for (int k = 0; k < LDI; ++k) {
        if (ld_yx[k] < 0) continue;
        const int gy = gy0 + (ld_yx[k] & 255), gx = gx0 + ((ld_yx[k] >> 8) & 255), c = ld_yx[k] >> 16;
        const bool ok = (unsigned)gy < (unsigned)g.H1 && (unsigned)gx < (unsigned)g.W1;
        const __half* src = ok ? img + ((gy * g.e1_pitch + gx) * C1 + c * 8) : g.e1;
        cp_async16(sE1 + ld_dst[k], src, ok);
      }
    };
    uint32_t ph = 0;
    if ((int)blockIdx.x < g.tiles) load_e1(blockIdx.x);
    for (int t = blockIdx.x; t < g.tiles; t += gridDim.x, ph ^= 1u) {
      const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
      const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;   // image coords (stem1 resolution) of region pixel (0,0)
      mark(t);
      cp_async_wait_all();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(e1_ready);
      mark(t);
      stem_worker_sync();       // the pool below reads halo pixels loaded by other threads
      mark(t);
      // ---- max-pool 2x2 s1 (ceil_mode, on the zero-padded e1) -> concat planes [0, ECH), overlapping the stem2a MMAs
#pragma unroll
      for (int k = 0; k < PLI; ++k) {
        if (pl_yx[k] < 0) continue;
        const int gy = gy0 + (pl_yx[k] & 255), gx = gx0 + (pl_yx[k] >> 8);
        uint4 o = make_uint4(0, 0, 0, 0);
        if ((unsigned)gy < (unsigned)g.H1 && (unsigned)gx < (unsigned)g.W1) {
          const uint8_t* e = sE1 + pl_src[k];
          const uint4 a0 = *reinterpret_cast<const uint4*>(e), a1 = *reinterpret_cast<const uint4*>(e + 16);
          const uint4 a2 = *reinterpret_cast<const uint4*>(e + P * 16), a3 = *reinterpret_cast<const uint4*>(e + P * 16 + 16);
          const __half2* h0 = reinterpret_cast<const __half2*>(&a0); const __half2* h1 = reinterpret_cast<const __half2*>(&a1);
          const __half2* h2 = reinterpret_cast<const __half2*>(&a2); const __half2* h3 = reinterpret_cast<const __half2*>(&a3);
          __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) ho[kk] = __hmax2(__hmax2(h0[kk], h1[kk]), __hmax2(h2[kk], h3[kk]));
        }
        *reinterpret_cast<uint4*>(sCAT + pl_dst[k]) = o;
      }
      mark(t);
      // ---- stem2a epilogue: accumulators -> stem2a planes (pixel index m, pitch P)
      for (int i = sub; i < S::MT2 * S::ACH; i += SUBS) {
        const int j = i / S::ACH, c = i % S::ACH;
        wait_done(&done2[j], ph);
        uint32_t r[8];
        tc::tmem_ld8(tq + (uint32_t)(S::T2 + j * S::N2A + c * 8), r);
        tc::tmem_ld_wait();
        const int m = j * 128 + q * 32 + lane;
        const int py = m / P, px = m % P;
        const bool inimg = (gy0 + py) >= 0 && (gy0 + py) < g.H1 && (gx0 + px) >= 0 && (gx0 + px) < g.W1;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = inimg ? fmaxf(__uint_as_float(r[k]) + sb2a[c * 8 + k], 0.f) : 0.f;
        Vec8<__half>::store(reinterpret_cast<__half*>(sAT + c * S::APLANE + m * 16), v);
      }
      mark(t);
      wait_done(&done2[S::MT2 - 1], ph);   // every stem2a MMA has retired: the e1 planes are dead after the pool
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(a_ready);
      mark(t);
      stem_worker_sync();       // pool + stem2a of this tile are done: the halo buffer is free
      mark(t);
      if (t + (int)gridDim.x < g.tiles) load_e1(t + gridDim.x);
      mark(t);
      // ---- stem2b epilogue -> concat planes [ECH, 2*ECH)
      for (int i = sub; i < S::MT3 * S::ECH; i += SUBS) {
        const int j = i / S::ECH, c = i % S::ECH;
        wait_done(&done3[j], ph);
        uint32_t r[8];
        tc::tmem_ld8(tq + (uint32_t)(S::T3 + j * S::N2B + c * 8), r);
        tc::tmem_ld_wait();
        const int m = j * 128 + q * 32 + lane;
        const int py = m / P, px = m % P;
        const bool live = py < S::CR_H && px < S::CR_W;
        const bool inimg = live && (gy0 + py) >= 0 && (gy0 + py) < g.H1 && (gx0 + px) >= 0 && (gx0 + px) < g.W1;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = inimg ? fmaxf(__uint_as_float(r[k]) + sb2b[c * 8 + k], 0.f) : 0.f;
        if (live) Vec8<__half>::store(reinterpret_cast<__half*>(sCAT + (S::ECH + c) * S::CPLANE + py * S::CROW + (px & 1) * S::CPAR + (px >> 1) * 16), v);
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      mark(t);
      tc::mbar_arrive(cat_ready);
      // ---- stem3 epilogue -> A planes of stem4
      wait_done(done4, ph);
      mark(t);
      for (int c = sub; c < S::ECH; c += SUBS) {
        uint32_t r[8];
        tc::tmem_ld8(tq + (uint32_t)(S::T4 + c * 8), r);
        tc::tmem_ld_wait();
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(__uint_as_float(r[k]) + sb3[c * 8 + k], 0.f);
        Vec8<__half>::store(reinterpret_cast<__half*>(sA2 + c * S::A2PLANE + (q * 32 + lane) * 16), v);
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(a2_ready);
      mark(t);
      // ---- stem4 epilogue -> global
      wait_done(done5, ph);
      mark(t);
      {
        const int r0 = q * 32 + lane;
        const int oy = ty * TY + r0 / TX, ox = tx * TX + r0 % TX;
        const bool ok = oy < g.H2 && ox < g.W2;
        __half* op = g.out + (((long long)n * g.H2 + oy) * g.W2 + ox) * C2;
        for (int c = sub; c < S::CCH; c += SUBS) {
          uint32_t r[8];
          tc::tmem_ld8(tq + (uint32_t)(S::T5 + c * 8), r);
          tc::tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaxf(__uint_as_float(r[k]) + sb4[c * 8 + k], 0.f);
          if (ok) Vec8<__half>::store(op + c * 8, v);
        }
      }
      tc::tc_fence_before();
      mark(t);
    }
    cp_async_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

template <int C1>
inline void launch_stem_planar(Ctx& cx, const Weights& w, const __half* e1, int n, int H1, int W1, int e1_pitch, __half* out, int H2, int W2) {
  using S = PlanarCfg<C1>;
  StemArgs a{};
  a.e1 = e1; a.N = n; a.H1 = H1; a.W1 = W1; a.e1_pitch = e1_pitch;
  a.w2a = w.get("stem2a.wp").h; a.w2b = w.get("stem2b.wp").h; a.w3 = w.get("stem3.w").h; a.w4 = w.get("stem4.w").h;
  a.b2a = w.get("stem2a.bp").d; a.b2b = w.get("stem2b.b").d; a.b3 = w.get("stem3.b").d; a.b4 = w.get("stem4.b").d;
  a.out = out; a.H2 = H2; a.W2 = W2;
  a.tiles_x = (W2 + S::TX - 1) / S::TX; a.tiles_y = (H2 + S::TY - 1) / S::TY; a.tiles = n * a.tiles_x * a.tiles_y;
  auto k = stem_planar_kernel<C1>;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  static const bool dbg = sw_debug("RDB_STEM_DBG") != nullptr;
  if (dbg) { RDB_CUDA(cudaMalloc(&a.dbg, 3 * 128 * sizeof(long long))); RDB_CUDA(cudaMemset(a.dbg, 0, 3 * 128 * sizeof(long long))); }
  cx.begin("stem_planar[P=" + std::to_string((long long)n * H2 * W2) + "]");
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
