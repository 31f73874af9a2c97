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
  static constexpr int oAT = oE1 + ECHP * EPLANE;
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

template <int C1>
__global__ void __launch_bounds__(kStemThreads + 32, 1) stem_planar_kernel(const StemArgs g) {
  using S = PlanarCfg<C1>;
  constexpr int TY = S::TY, TX = S::TX, CA = S::CA, C2 = S::C2, P = S::ER_W;
  constexpr int SUBS = kStemThreads / 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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

  if (warp == kStemThreads / 32) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t aE1 = tc::smem_u32(sE1), aAT = tc::smem_u32(sAT), aCAT = tc::smem_u32(sCAT), aA2 = tc::smem_u32(sA2);
      const uint32_t aW2A = tc::smem_u32(sm + S::oW2A), aW2B = tc::smem_u32(sm + S::oW2B), aW3 = tc::smem_u32(sm + S::oW3), aW4 = tc::smem_u32(sm + S::oW4);
      // B operand K step ks of a 128B-swizzled weight tile with n_pad rows per k-block
      auto bdesc = [](uint32_t base, int n_pad, int ks) { return tc::make_smem_desc(base + (uint32_t)((ks >> 2) * n_pad * 128), 128) + (uint64_t)(2 * (ks & 3)); };
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < g.tiles; t += gridDim.x, ph ^= 1u) {
        tc::mbar_wait(e1_ready, ph);
        tc::tc_fence_after();
        for (int j = 0; j < S::MT2; ++j) {                               // stem2a: 4 taps x ECHP/2 K steps per M-tile
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int h = 0; h < S::ECHP / 2; ++h) {
              const uint32_t a = aE1 + (uint32_t)(2 * h * S::EPLANE + (j * 128 + (tap >> 1) * P + (tap & 1)) * 16);
              tc::umma_f16(tmem_base + (uint32_t)(S::T2 + j * S::N2A), planar_desc(a, S::EPLANE, 128), bdesc(aW2A, S::N2A, tap * (S::ECHP / 2) + h),
                           idesc(S::N2A), (tap | h) != 0 ? 1u : 0u);
            }
          tc::umma_commit(&done2[j]);
        }
        tc::mbar_wait(a_ready, ph);
        tc::tc_fence_after();
        for (int j = 0; j < S::MT3; ++j) {                               // stem2b
#pragma unroll
          for (int tap = 0; tap < 4; ++tap)
#pragma unroll
            for (int h = 0; h < S::ACHP / 2; ++h) {
              const uint32_t a = aAT + (uint32_t)(2 * h * S::APLANE + (j * 128 + (tap >> 1) * P + (tap & 1)) * 16);
              tc::umma_f16(tmem_base + (uint32_t)(S::T3 + j * S::N2B), planar_desc(a, S::APLANE, 128), bdesc(aW2B, S::N2B, tap * (S::ACHP / 2) + h),
                           idesc(S::N2B), (tap | h) != 0 ? 1u : 0u);
            }
          tc::umma_commit(&done3[j]);
        }
        tc::mbar_wait(cat_ready, ph);
        tc::tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {                              // stem3: 3x3 stride 2 on the parity-split concat planes
          const int ky = tap / 3, kx = tap % 3;
#pragma unroll
          for (int h = 0; h < S::CCHP / 2; ++h) {
            const uint32_t a = aCAT + (uint32_t)(2 * h * S::CPLANE + ky * S::CROW + (kx & 1) * S::CPAR + (kx >> 1) * 16);
            tc::umma_f16(tmem_base + (uint32_t)S::T4, planar_desc(a, S::CPLANE, 2 * S::CROW), bdesc(aW3, S::N3, tap * (S::CCHP / 2) + h), idesc(S::N3),
                         (tap | h) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(done4);
        tc::mbar_wait(a2_ready, ph);
        tc::tc_fence_after();
#pragma unroll
        for (int h = 0; h < S::ECHP / 2; ++h)                            // stem4: 1x1
          tc::umma_f16(tmem_base + (uint32_t)S::T5, planar_desc(aA2 + (uint32_t)(2 * h * S::A2PLANE), S::A2PLANE, 128), bdesc(aW4, S::N4, h), idesc(S::N4),
                       h != 0 ? 1u : 0u);
        tc::umma_commit(done5);
      }
    }
  } else {
    // ================= workers: halo load, pool, epilogues =================
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    auto wait_done = [&](uint64_t* b, uint32_t ph) { tc::mbar_wait(b, ph); __syncwarp(); tc::tc_fence_after(); };
    auto load_e1 = [&](int t) {
      const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
      const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;
      const __half* img = g.e1 + (long long)n * g.H1 * g.e1_pitch * C1;
      for (int i = tid; i < S::ER_H * S::ER_W * S::ECH; i += kStemThreads) {
        const int p = i % (S::ER_H * S::ER_W), c = i / (S::ER_H * S::ER_W);     // consecutive threads -> consecutive pixels of one plane
        const int py = p / S::ER_W, px = p % S::ER_W;
        const int gy = gy0 + py, gx = gx0 + px;
        const bool ok = gy >= 0 && gy < g.H1 && gx >= 0 && gx < g.W1;
        const __half* src = ok ? img + ((gy * g.e1_pitch + gx) * C1 + c * 8) : g.e1;
        cp_async16(sE1 + c * S::EPLANE + p * 16, src, ok);
      }
    };
    uint32_t ph = 0;
    if ((int)blockIdx.x < g.tiles) load_e1(blockIdx.x);
    for (int t = blockIdx.x; t < g.tiles; t += gridDim.x, ph ^= 1u) {
      const int tx = t % g.tiles_x, ty = (t / g.tiles_x) % g.tiles_y, n = t / (g.tiles_x * g.tiles_y);
      const int gy0 = 2 * ty * TY - 1, gx0 = 2 * tx * TX - 1;   // image coords (stem1 resolution) of region pixel (0,0)
      cp_async_wait_all();
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(e1_ready);
      stem_worker_sync();       // the pool below reads halo pixels loaded by other threads
      // ---- max-pool 2x2 s1 (ceil_mode, on the zero-padded e1) -> concat planes [0, ECH), overlapping the stem2a MMAs
      for (int i = tid; i < S::CR_H * S::CR_W * S::ECH; i += kStemThreads) {
        const int p = i % (S::CR_H * S::CR_W), c = i / (S::CR_H * S::CR_W);
        const int py = p / S::CR_W, px = p % S::CR_W;
        const bool inimg = (gy0 + py) >= 0 && (gy0 + py) < g.H1 && (gx0 + px) >= 0 && (gx0 + px) < g.W1;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (inimg) {
          const uint8_t* e = sE1 + c * S::EPLANE + (py * P + px) * 16;
          const uint4 a0 = *reinterpret_cast<const uint4*>(e), a1 = *reinterpret_cast<const uint4*>(e + 16);
          const uint4 a2 = *reinterpret_cast<const uint4*>(e + P * 16), a3 = *reinterpret_cast<const uint4*>(e + P * 16 + 16);
          const __half2* h0 = reinterpret_cast<const __half2*>(&a0); const __half2* h1 = reinterpret_cast<const __half2*>(&a1);
          const __half2* h2 = reinterpret_cast<const __half2*>(&a2); const __half2* h3 = reinterpret_cast<const __half2*>(&a3);
          __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) ho[k] = __hmax2(__hmax2(h0[k], h1[k]), __hmax2(h2[k], h3[k]));
        }
        *reinterpret_cast<uint4*>(sCAT + c * S::CPLANE + py * S::CROW + (px & 1) * S::CPAR + (px >> 1) * 16) = o;
      }
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
      wait_done(&done2[S::MT2 - 1], ph);   // every stem2a MMA has retired: the e1 planes are dead after the pool
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(a_ready);
      stem_worker_sync();
      if (t + (int)gridDim.x < g.tiles) load_e1(t + gridDim.x);
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
      tc::mbar_arrive(cat_ready);
      // ---- stem3 epilogue -> A planes of stem4
      wait_done(done4, ph);
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
      // ---- stem4 epilogue -> global
      wait_done(done5, ph);
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
  static bool attr_done = false;
  if (!attr_done) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); attr_done = true; }
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  cx.begin("stem_planar[P=" + std::to_string((long long)n * H2 * W2) + "]");
  k<<<grid, kStemThreads + 32, S::kSmem, cx.st>>>(a);
  cx.end();
}

}  // namespace rdb
