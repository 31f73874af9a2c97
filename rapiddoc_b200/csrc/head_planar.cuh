// RepLKFPN output concat + DBHead, fused into ONE persistent kernel (fp16 / tcgen05 mode):
//   concat_{L=3..0}( nearest_up_{2^L}(p_L) * (1 + se_gate_L) )                    (db_fpn.py:401-415)
//   -> conv3x3(96->24)+BN+ReLU -> ConvT2x2 s2(24->24)+BN+ReLU -> ConvT2x2 s2(24->1) -> sigmoid -> nan_to_num
//   -> prob fp32 [4H,4W] + (prob > thresh) byte                                   (det_db_head.py:103-147, ocr_patch.py:229)
// Unfused this was neck_concat (writes 96-ch map) + head_conv3x3 (TMA patch conv) + head_tail: 550 us per 16 pages, the 96-channel
// concat written and read back through HBM.  Here a CTA builds the concat halo tile (10 x 32 px) of an 8 x 30 output tile
// directly in shared memory as CHANNEL PLANES (see stem_planar.cuh): the 3x3 conv is 9 tap-shifted no-swizzle UMMA operands
// per K step, its epilogue writes the 24-channel A planes of the first transposed conv (one [128,24]x[24,96] MMA), and the
// second epilogue — sub-warp = ConvT tap — finishes ReLU / final ConvT / sigmoid / threshold in registers.
// Two halo buffers + two accumulator sets: the MMA warp always has the next tile's conv queued while the 16 worker warps
// build tile t+1 and drain tiles t / t-1, so the tensor pipe (A-operand read bound at these tiny N) never idles.
#pragma once
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "stem_planar.cuh"

namespace rdb {

struct HeadArgs {
  const __half* f[4]; const float* gate[4];     // p_L [N, H>>L, W>>L, 24], gate_L [N][24] (already 1 + g)
  int N, H, W;                                  // level-0 size
  const __half* w_down; const float* b_down;    // [24][3][3][96], [24]
  const __half* w_up; const float* b_up;        // [4 taps * 24][24], [24]
  const float* w_fin; const float* b_fin;       // [4][24], [1]
  float thresh; float* prob; uint8_t* seg;      // [N, 4H, 4W]
  int tiles_x, tiles_y, tiles;
  int dbg;                                      // RDB_HEAD_DBG bit mask (timing experiments only): 1 skip the halo build, 2 skip the tail math, 4 skip the conv MMAs
};

struct HeadCfg {
  static constexpr int TH = 8, TW = 30, PITCH = 32;
  static constexpr int RH = TH + 2;                              // halo rows
  static constexpr int MT = TH * PITCH / 128;                    // M-tiles per tile (2)
  static constexpr int PROWS = (MT * 128 + 2 * PITCH + 2 + 7) / 8 * 8;   // plane rows a shifted M-tile may touch
  static constexpr int PLANE = PROWS * 16;
  static constexpr int CIN = 96, CCH = 12, CMID = 24, MCH = 3;
  static constexpr int K1 = 9 * CIN, KB1 = (K1 + 63) / 64, N1 = 32;       // conv3x3, N padded 24 -> 32
  static constexpr int N2 = 96, A2PLANE = 128 * 16;
  static constexpr int oPL = 0;                                  // [2 buffers][12 planes]
  static constexpr int oA2 = oPL + 2 * CCH * PLANE;              // [2 buffers][MT][3 planes] + one shared zero plane
  static constexpr int oZERO = oA2 + 2 * MT * MCH * A2PLANE;
  static constexpr int oW1 = (oZERO + A2PLANE + 1023) / 1024 * 1024;
  static constexpr int oW2 = oW1 + KB1 * N1 * 128;
  static constexpr int oF = oW2 + N2 * 128;                      // floats: b_down[24] b_up[24] w_fin[24][4] b_fin[4] gates[2][96]
  static constexpr int oBAR = (oF + (24 + 24 + 96 + 4 + 192) * 4 + 15) / 16 * 16;
  static constexpr int kBars = 2 + 2 + 2 * MT + 2;
  static constexpr int kSmem = oBAR + kBars * 8 + 16 + 1024;
  static constexpr int T1 = 0, T2 = 2 * MT * N1;                 // TMEM: conv acc [2][MT] x 32, tail acc [2][MT] x 96
  static_assert(T2 + 2 * MT * N2 <= 512, "head_planar: TMEM");
  static_assert(kSmem <= 227 * 1024, "head_planar: shared memory");
};

static __global__ void __launch_bounds__(kStemThreads + 32, 1) head_planar_kernel(const HeadArgs g) {
  using S = HeadCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = RDB_ALIGNED_SMEM(smem_raw);
  float* sF = reinterpret_cast<float*>(sm + S::oF);
  float* sbd = sF; float* sbu = sF + 24; float* swf = sF + 48; float* sbf = sF + 144; float* sgate = sF + 148;   // swf: [c][4 final taps]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::oBAR);
  uint64_t* planes_ready = bars;            // [2], one arrival per worker
  uint64_t* a2_ready = bars + 2;            // [2]
  uint64_t* conv_done = bars + 4;           // [2][MT]
  uint64_t* tail_done = bars + 4 + 2 * S::MT;   // [2]
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
    for (int i = tid; i < (S::oZERO + S::A2PLANE) / 16; i += kStemThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    planar_fill_w(sm + S::oW1, g.w_down, S::CMID, S::N1, 9, S::CIN, S::CIN);
    planar_fill_w(sm + S::oW2, g.w_up, S::N2, S::N2, 1, S::CMID, 32);
    for (int i = tid; i < 24; i += kStemThreads) { sbd[i] = g.b_down[i]; sbu[i] = g.b_up[i]; }
    for (int i = tid; i < 96; i += kStemThreads) swf[(i % 24) * 4 + i / 24] = g.w_fin[i];
    if (tid == 0) sbf[0] = g.b_fin[0];
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_local = g.tiles > (int)blockIdx.x ? (g.tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles of this CTA

  if (warp == kStemThreads / 32) {
    // ================= MMA issuer =================
    if (lane == 0 && n_local > 0) {
      const uint32_t loPL = (tc::smem_u32(sm + S::oPL) >> 4) | ((uint32_t)(S::PLANE >> 4) << 16);
      constexpr uint32_t hiRow = (128u >> 4) | (1u << 14);
      constexpr uint32_t hiW = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t loW1 = (tc::smem_u32(sm + S::oW1) >> 4) | (1u << 16), loW2 = (tc::smem_u32(sm + S::oW2) >> 4) | (1u << 16);
      const uint32_t aA2 = tc::smem_u32(sm + S::oA2), aZ = tc::smem_u32(sm + S::oZERO);
      constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(S::N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(S::N2 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto wofs = [](int n_pad, int ks) { return (uint32_t)((ks >> 2) * n_pad * 8 + 2 * (ks & 3)); };
      auto conv = [&](int it) {
        const int b = it & 1;
        tc::mbar_wait(&planes_ready[b], (uint32_t)(it >> 1) & 1u);
        tc::tc_fence_after();
        for (int j = 0; j < S::MT; ++j) {
          const uint32_t aj = loPL + (uint32_t)(b * S::CCH * (S::PLANE >> 4) + j * 128), dj = tmem_base + (uint32_t)(S::T1 + (b * S::MT + j) * S::N1);
#pragma unroll
          for (int tap = 0; tap < ((g.dbg & 4) ? 1 : 9); ++tap)
#pragma unroll
            for (int h = 0; h < S::CCH / 2; ++h)
              umma_f16_lh(dj, aj + (uint32_t)(2 * h * (S::PLANE >> 4) + (tap / 3) * S::PITCH + (tap % 3)), hiRow, loW1 + wofs(S::N1, tap * (S::CCH / 2) + h), hiW, idesc1,
                          (tap | h) != 0 ? 1u : 0u);
          tc::umma_commit(&conv_done[b * S::MT + j]);
        }
      };
      conv(0);
      for (int it = 0; it < n_local; ++it) {
        if (it + 1 < n_local) conv(it + 1);
        const int b = it & 1;
        tc::mbar_wait(&a2_ready[b], (uint32_t)(it >> 1) & 1u);
        tc::tc_fence_after();
        for (int j = 0; j < S::MT; ++j) {
          const uint32_t a0 = aA2 + (uint32_t)((b * S::MT + j) * S::MCH * S::A2PLANE);
          const uint32_t dj = tmem_base + (uint32_t)(S::T2 + (b * S::MT + j) * S::N2);
          // K step 0: planes 0,1; K step 1: plane 2 + the shared zero plane (its LBO depends on the buffer)
          umma_f16_lh(dj, (a0 >> 4) | ((uint32_t)(S::A2PLANE >> 4) << 16), hiRow, loW2 + wofs(S::N2, 0), hiW, idesc2, 0u);
          const uint32_t a2 = a0 + 2 * S::A2PLANE;
          umma_f16_lh(dj, (a2 >> 4) | (((aZ - a2) >> 4) << 16), hiRow, loW2 + wofs(S::N2, 1), hiW, idesc2, 1u);
        }
        tc::umma_commit(&tail_done[b]);
      }
    }
  } else {
    // ================= workers =================
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    auto wait_done = [&](uint64_t* b, uint32_t ph) { tc::mbar_wait(b, ph); __syncwarp(); tc::tc_fence_after(); };
    auto tile_xy = [&](int it, int& n, int& y0, int& x0) {
      const int t = blockIdx.x + it * gridDim.x;
      n = t / (g.tiles_x * g.tiles_y);
      const int r = t % (g.tiles_x * g.tiles_y);
      y0 = (r / g.tiles_x) * S::TH; x0 = (r % g.tiles_x) * S::TW;
    };
    // concat halo tile of local tile `it` -> planes[it & 1]
    auto build = [&](int it) {
      int n, y0, x0;
      tile_xy(it, n, y0, x0);
      uint8_t* pl = sm + S::oPL + (it & 1) * S::CCH * S::PLANE;
      float* gt = sgate + (it & 1) * 96;            // (1 + gate) of this tile's image, concat channel order
      if (tid < 96 && !(g.dbg & 1)) gt[tid] = __ldg(g.gate[3 - tid / 24] + n * 24 + tid % 24);
      stem_worker_sync();
      constexpr int NPX = S::RH * S::PITCH, ITEMS = (NPX * S::CCH + kStemThreads - 1) / kStemThreads;
      uint4 v[ITEMS];
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {             // all global loads of this thread in flight before the first use
        const int i = tid + k * kStemThreads;
        const int c = i / NPX, p = i % NPX;
        const int gy = y0 - 1 + (p >> 5), gx = x0 - 1 + (p & 31);
        v[k] = make_uint4(0, 0, 0, 0);
        if (i < NPX * S::CCH && !(g.dbg & 1) && (unsigned)gy < (unsigned)g.H && (unsigned)gx < (unsigned)g.W) {
          const int L = 3 - c / 3;
          v[k] = __ldg(reinterpret_cast<const uint4*>(g.f[L] + (((long long)n * (g.H >> L) + (gy >> L)) * (g.W >> L) + (gx >> L)) * 24 + (c % 3) * 8));
        }
      }
#pragma unroll
      for (int k = 0; k < ITEMS; ++k) {
        const int i = tid + k * kStemThreads;
        if (i >= NPX * S::CCH) break;
        const int c = i / NPX, p = i % NPX;
        const float4 g0 = *reinterpret_cast<const float4*>(gt + c * 8), g1 = *reinterpret_cast<const float4*>(gt + c * 8 + 4);
        const __half2* hv = reinterpret_cast<const __half2*>(&v[k]);
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
        float2 f;
        f = __half22float2(hv[0]); ho[0] = __floats2half2_rn(f.x * g0.x, f.y * g0.y);
        f = __half22float2(hv[1]); ho[1] = __floats2half2_rn(f.x * g0.z, f.y * g0.w);
        f = __half22float2(hv[2]); ho[2] = __floats2half2_rn(f.x * g1.x, f.y * g1.y);
        f = __half22float2(hv[3]); ho[3] = __floats2half2_rn(f.x * g1.z, f.y * g1.w);
        *reinterpret_cast<uint4*>(pl + c * S::PLANE + p * 16) = o;
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&planes_ready[it & 1]);
    };
    // conv3x3 epilogue of local tile `it`: +bias, ReLU -> A planes of the first transposed conv
    auto epi1 = [&](int it) {
      const int b = it & 1;
      const uint32_t ph = (uint32_t)(it >> 1) & 1u;
      for (int i = sub; i < S::MT * S::MCH; i += kStemThreads / 128) {
        const int j = i / S::MCH, c = i % S::MCH;
        wait_done(&conv_done[b * S::MT + j], ph);
        uint32_t r[8];
        tc::tmem_ld8(tq + (uint32_t)(S::T1 + (b * S::MT + j) * S::N1 + c * 8), r);
        tc::tmem_ld_wait();
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(__uint_as_float(r[k]) + sbd[c * 8 + k], 0.f);
        Vec8<__half>::store(reinterpret_cast<__half*>(sm + S::oA2 + ((b * S::MT + j) * S::MCH + c) * S::A2PLANE + (q * 32 + lane) * 16), v);
      }
      wait_done(&conv_done[b * S::MT + S::MT - 1], ph);   // every conv MMA of this tile has retired: its halo buffer is free
      tc::fence_proxy_async();
      tc::tc_fence_before();
      tc::mbar_arrive(&a2_ready[b]);
    };
    // tail epilogue of local tile `it`: sub-warp = tap (dy,dx) of the first ConvT; ReLU, final ConvT(24->1), sigmoid, threshold
    auto epi2 = [&](int it) {
      const int b = it & 1;
      int n, y0, x0;
      tile_xy(it, n, y0, x0);
      wait_done(&tail_done[b], (uint32_t)(it >> 1) & 1u);
      const int W4 = 4 * g.W;
      for (int j = 0; j < ((g.dbg & 2) ? 0 : S::MT); ++j) {
        uint32_t r16[16], r8[8];
        const uint32_t ta = tq + (uint32_t)(S::T2 + (b * S::MT + j) * S::N2 + 24 * sub);
        tc::tmem_ld16(ta, r16);
        tc::tmem_ld8(ta + 16u, r8);
        tc::tmem_ld_wait();
        float o[4];
        o[0] = o[1] = o[2] = o[3] = sbf[0];
#pragma unroll
        for (int c = 0; c < 24; ++c) {
          const float acc = __uint_as_float(c < 16 ? r16[c < 16 ? c : 0] : r8[c >= 16 ? c - 16 : 0]);
          const float hv = fmaxf(acc + sbu[c], 0.f);
          const float4 wf = *reinterpret_cast<const float4*>(swf + c * 4);
          o[0] = fmaf(hv, wf.x, o[0]); o[1] = fmaf(hv, wf.y, o[1]); o[2] = fmaf(hv, wf.z, o[2]); o[3] = fmaf(hv, wf.w, o[3]);
        }
        const int m = j * 128 + q * 32 + lane;
        const int y = y0 + (m >> 5), x = x0 + (m & 31);
        if ((m & 31) < S::TW && y < g.H && x < g.W) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float pz = 1.f / (1.f + expf(-o[k]));
            if (pz != pz) pz = 0.f;
            o[k] = pz;
          }
          const long long base = ((long long)n * (4 * g.H) + 4 * y + 2 * (sub >> 1)) * W4 + 4 * x + 2 * (sub & 1);
          *reinterpret_cast<float2*>(g.prob + base) = make_float2(o[0], o[1]);
          *reinterpret_cast<float2*>(g.prob + base + W4) = make_float2(o[2], o[3]);
          if (g.seg != nullptr) {
            *reinterpret_cast<uchar2*>(g.seg + base) = make_uchar2(o[0] > g.thresh, o[1] > g.thresh);
            *reinterpret_cast<uchar2*>(g.seg + base + W4) = make_uchar2(o[2] > g.thresh, o[3] > g.thresh);
          }
        }
      }
      tc::tc_fence_before();
    };
    if (n_local > 0) build(0);
    for (int it = 0; it < n_local; ++it) {
      if (it + 1 < n_local) build(it + 1);
      epi1(it);
      if (it >= 1) epi2(it - 1);
    }
    if (n_local > 0) epi2(n_local - 1);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

inline void launch_head_planar(Ctx& cx, const Weights& w, const __half* const* f, const float* const* gate, int n, int H, int W, float thresh, float* prob,
                               uint8_t* seg) {
  using S = HeadCfg;
  HeadArgs a{};
  for (int i = 0; i < 4; ++i) { a.f[i] = f[i]; a.gate[i] = gate[i]; }
  a.N = n; a.H = H; a.W = W;
  a.w_down = w.get("head.down.w").h; a.b_down = w.get("head.down.b").d;
  a.w_up = w.get("head.up.w").h; a.b_up = w.get("head.up.b").d;
  a.w_fin = w.get("head.final.w").d; a.b_fin = w.get("head.final.b").d;
  a.thresh = thresh; a.prob = prob; a.seg = seg;
  if (const char* e = sw_debug("RDB_HEAD_DBG")) a.dbg = std::atoi(e);
  a.tiles_x = (W + S::TW - 1) / S::TW; a.tiles_y = (H + S::TH - 1) / S::TH; a.tiles = n * a.tiles_x * a.tiles_y;
  auto k = head_planar_kernel;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kSmem)); }
  const int grid = a.tiles < cx.num_sms ? a.tiles : cx.num_sms;
  cx.begin("head_planar[P=" + std::to_string((long long)n * H * W) + "]");
  k<<<grid, kStemThreads + 32, S::kSmem, cx.st>>>(a);
  cx.end();
}

}  // namespace rdb
