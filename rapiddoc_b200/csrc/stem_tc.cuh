// stem1 (conv3x3 s2 p1, 3 -> C1, +BN +ReLU; rec_lcnetv4.py:151,160) as an implicit GEMM on the tensor cores.
// The input is uint8 HWC (or fp32 NCHW at the InferSession seam), which TMA cannot normalise or convert, so the A operand
// is built by the CTA's own threads: thread t of a 128-thread CTA gathers the 27 taps of output pixel t of the tile,
// normalises them through the 768-entry table (the reference's float32 op order) and writes one 64-byte K-major row
// (K = 27 padded to 32 halves) straight into shared memory in the 64-byte-swizzle layout tcgen05.mma expects
// (16-byte chunk c of row r lives at r*64 + ((c ^ ((r>>1)&3)) << 4) — the layout TMA SWIZZLE_64B would have produced).
// One elected thread then issues two tcgen05.mma (K = 2 x 16) against the [C1 x 32] weight tile; the fp32 accumulators
// come back from TMEM one row per thread (lane = pixel) for bias + ReLU + the fp16 NHWC store.
// CTAs are small (128 threads, ~12 KB smem, 32/64 TMEM columns) so 8+ of them share an SM and hide each other's
// gather / MMA / store latencies; no intra-CTA pipeline is needed.
// The SIMT stem1_kernel (kernels.cuh) did 648 FFMA + 162 LDS per pixel and ran at 12% of the HBM roofline.
#pragma once
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace rdb {

template <typename IN, int C1>
__global__ void __launch_bounds__(128)
stem1_tc_kernel(IN in, int N, const __half* __restrict__ wh /*[C1][27] = [co][ky][kx][ci]*/, const float* __restrict__ b,
                __half* __restrict__ out, int OH, int OW, int out_wp, int tiles) {
  constexpr int NB = (C1 + 15) / 16 * 16;          // UMMA N
  constexpr uint32_t TCOLS = NB <= 32 ? 32 : 64;   // TMEM columns (power of two >= 32)
  __shared__ __align__(1024) uint8_t sAB[128 * 64 + NB * 64];
  __shared__ float lut[IN::kU8 ? 768 : 1];
  __shared__ float sb[C1];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = sAB;
  uint8_t* sB = sAB + 128 * 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TCOLS);
  for (int i = tid; i < NB * 4; i += 128) {        // weight tile, zero-padded to [NB][32]
    const int n = i >> 2, c = i & 3;
    uint4 u = make_uint4(0, 0, 0, 0);
    __half* h = reinterpret_cast<__half*>(&u);
    if (n < C1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c * 8 + j < 27) h[j] = wh[n * 27 + c * 8 + j];
    }
    *reinterpret_cast<uint4*>(sB + n * 64 + ((c ^ ((n >> 1) & 3)) << 4)) = u;
  }
  for (int i = tid; i < C1; i += 128) sb[i] = b[i];
  if (IN::kU8)
    for (int i = tid; i < 768; i += 128) lut[i] = in.norm(i >> 8, i & 255);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const long long total = (long long)N * OH * out_wp;
  uint32_t ph = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long idx = (long long)tile * 128 + tid;
    const bool ok = idx < total;
    int ox = 0, oy = 0, n = 0;
    if (ok) { ox = (int)(idx % out_wp); const long long r = idx / out_wp; oy = (int)(r % OH); n = (int)(r / OH); }
    const bool real = ok && ox < OW;               // pad pixels (ox >= OW) are written as zeros
    uint4 row[4];
    {
      __half2* hp = reinterpret_cast<__half2*>(row);
      float v[28];
      const bool xin = real && in.row_inside(n, ox * 2 - 1);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (real && iy >= 0 && iy < in.H) {
          float r9[9];
          if (xin) in.row9_fast(lut, n, iy, ox * 2 - 1, r9);
          else in.row9(lut, n, iy, ox * 2 - 1, r9);
#pragma unroll
          for (int j = 0; j < 9; ++j) v[ky * 9 + j] = r9[j];
        } else {
#pragma unroll
          for (int j = 0; j < 9; ++j) v[ky * 9 + j] = 0.f;
        }
      }
      v[27] = 0.f;
#pragma unroll
      for (int j = 0; j < 14; ++j) hp[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      hp[14] = __floats2half2_rn(0.f, 0.f);
      hp[15] = hp[14];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sA + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = row[c];
    tc::fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc::tc_fence_before();       // orders this thread's TMEM loads of the previous tile before the next MMA
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint64_t da = tc::make_smem_desc(tc::smem_u32(sA), 64), db = tc::make_smem_desc(tc::smem_u32(sB), 64);
      tc::umma_f16(tmem_base, da, db, idesc, 0u);
      tc::umma_f16(tmem_base, da + 2, db + 2, idesc, 1u);
      tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, ph);
    ph ^= 1;
    __syncwarp();
    tc::tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    __half* op = out + idx * C1;
#pragma unroll
    for (int c0 = 0; c0 < C1; c0 += 16) {
      float v[16];
      if (c0 + 16 <= C1) {
        uint32_t r[16];
        tc::tmem_ld16(taddr + (uint32_t)c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      } else {
        uint32_t r[8];
        tc::tmem_ld8(taddr + (uint32_t)c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = __uint_as_float(r[j]); v[8 + j] = 0.f; }
      }
      if (ok) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = real ? fmaxf(v[j] + sb[c0 + j], 0.f) : 0.f;
        Vec8<__half>::store(op + c0, a);
        if (c0 + 16 <= C1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = real ? fmaxf(v[8 + j] + sb[c0 + 8 + j], 0.f) : 0.f;
          Vec8<__half>::store(op + c0 + 8, a);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, TCOLS); }
}

template <typename IN, int C1>
inline void launch_stem1_tc(Ctx& cx, const IN& src, int n, const Tensor& w, const Tensor& b, __half* out, int OH, int OW, int out_wp) {
  const long long total = (long long)n * OH * out_wp;
  const int tiles = cdiv(total, 128);
  int grid = cx.num_sms * 8;
  if (grid > tiles) grid = tiles;
  cx.begin("stem1_tc[P=" + std::to_string((long long)n * OH * OW) + ",N=" + std::to_string(C1) + "]");
  stem1_tc_kernel<IN, C1><<<grid, 128, 0, cx.st>>>(src, n, w.h, b.d, out, OH, OW, out_wp, tiles);
  cx.end();
}

}  // namespace rdb
