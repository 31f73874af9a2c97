// tcgen05 tensor-core GEMM for the pointwise convs / linears (precision mode fp16):
//   out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ res)          A, W, out: fp16; accumulate fp32
// A is the NHWC activation tensor viewed as a K-major [M,K] matrix, W the K-major weight.
//
// Blackwell structure (sm_100a only; nothing here compiles for another arch):
//   * TMA (cp.async.bulk.tensor.2d, hardware 32/64/128-byte swizzle) stages A/B k-blocks into a
//     multi-stage shared-memory ring guarded by full/empty mbarriers,
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA M=128, N<=256, K=16)
//     with the fp32 accumulator in TMEM; tcgen05.commit releases smem stages / publishes the
//     accumulator,
//   * two TMEM accumulator stages: the epilogue warps drain tile i (tcgen05.ld -> bias/act/
//     residual -> fp16 global stores, or the fused CTC argmax/sum-exp) while the tensor core
//     already works on tile i+1,
//   * persistent CTAs (grid = #SMs) walk the (m-tile, n-tile) list.
// K is tiny on this path (24..768): the k-block is ONE swizzle atom wide (16/32/64 halves, the
// largest that divides K) so any channel count that is a multiple of 16 maps without padding
// HBM traffic; ragged K/N/M edges are zero-filled by TMA out-of-bounds handling.
// Reference ops covered: channel_conv1/2 (backbones/rec_lcnetv4.py:210-224), RepLKFPN 1x1s
// (necks/db_fpn.py:326-331,350-356), LightSVTR convs/linears (necks/rnn.py:238-290), CTC head
// (heads/rec_multi_head.py:45,70).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace rdb {
namespace tc {

// ------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  // suspend-time hint: the warp sleeps in hardware until the phase flips (or the hint expires) instead of
  // hammering the mbarrier unit — 16 spinning epilogue warps otherwise starve the TMA / tcgen05.commit arrivals
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { printf("rdb gemm_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand in a hardware-swizzled tile whose rows are one
// swizzle span (32/64/128 B) wide: 8-row groups are SBO = 8*span bytes apart.
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor — start>>4 [0,14), LBO>>4 [16,30),
//  SBO>>4 [32,46), version=1 [46,48), layout_type [61,64): 2=128B, 4=64B, 6=32B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t span_bytes) {
  uint64_t lt = span_bytes == 128 ? 2ull : (span_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * span_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= lt << 61;
  return d;
}

// 1024-byte aligned base of the dynamic shared memory.  The default re-derives the pointer through an integer cast, which makes
// nvcc treat everything behind it as GENERIC memory (LD.E/ST.E instead of LDS/STS — see DESIGN.md section 6).  Building with
// RDB_SMEM_BASE=shared (rapiddoc_b200/build.py -> -DRDB_SMEM_SHARED_BASE) keeps the address space; it has not run on a GPU yet,
// so it is a build-time experiment, not the default.
#ifdef RDB_SMEM_SHARED_BASE
#define RDB_ALIGNED_SMEM(raw) ((raw) + ((1024u - (::rdb::tc::smem_u32(raw) & 1023u)) & 1023u))
#else
#define RDB_ALIGNED_SMEM(raw) reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023)
#endif

// ------------------------------------------------------------------------ kernel
struct Args {
  int M, N, K;
  int BN;          // n-tile (UMMA N), multiple of 16, <= 256
  int tiles_m, tiles_n;
  int k_blocks;    // ceil(K / AW)
  int AW;          // k-block width in halves: 16, 32 or 64 (= one swizzle span)
  int stages;
  int resident;    // 1: the CTA keeps its n-tile of B (all k-blocks) in smem for its whole life; stages hold A only
  int ctas_per_n;  // resident: CTAs per n-tile, each walks m-tiles tm = blockIdx.x / tiles_n + i * ctas_per_n
  int walk_dm, walk_dn;          // per-iteration tile step of one CTA: tm += walk_dm, tn += walk_dn (with carry)
  int walk_sx, walk_sy, walk_sn; // conv: walk_dm decomposed into (x, y, image) tile steps
  int epi_subs;                  // epilogue sub-warps per lane quarter that own at least one 16-column chunk
  int tmem_cols;   // power of two >= 2*BN
  uint32_t idesc;
  const float* bias;
  const __half* res; int ldr;
  __half* out; int ldc; int c_off;
  int act;
  float* pmax; int* pidx; float* psum;  // EPI_CTC partials [M, tiles_n * kEpiSubs]
  // fused RepLKFPN input stage (db_fpn.py:342-363,394-399): out = acc * colscale[image][col] + nearest_up2(up_res)
  const float* colscale; int rows_per_img;   // per-(image, column) SE factor 1+gate, rows_per_img = H*W of this level
  const __half* up_res; int up_H, up_W;      // coarser level [n, up_H/2, up_W/2, N] added at (y/2, x/2)
  // EPI_HEAD (fused DBHead tail): rows are pixels of an [n, hH, hW] map, N = 4*24 ConvT outputs; the
  // epilogue applies ReLU, the final ConvT(24->1, 2x2 s2), sigmoid and the DB threshold.
  int hH, hW;
  const float* w_fin; const float* b_fin; float thresh;
  float* prob; uint8_t* seg;
  // implicit-GEMM conv mode (conv != 0): A is the NHWC input behind a 4-D tensor map
  // {C, W, H, n}; an m-tile is a TH x TW patch of output pixels (TH*TW = 128) of one image and
  // k-block kb = (tap, channel chunk) is the same patch shifted by the tap — TMA out-of-bounds
  // zero fill provides the conv padding.
  int conv;
  int OH, OW, TH, TW, tiles_y, tiles_x;
  int KW, sh, sw, pt, pl, cchunks, C, tw_shift;
  // patch mode (stride-1 convs, B resident): one TMA box per (kx, channel chunk) holds TH+KH-1 input rows x 8 columns; the
  // KH vertical taps are the SAME smem box read at +8-row (= one swizzle atom) descriptor offsets -> KH x fewer loads/bytes
  int patch, KH, b_blocks;
  int out_wp;                // conv epilogue: output row pitch in pixels (>= OW; 0 = OW)   // b_blocks: resident B k-blocks (= KH*KW*cchunks in conv modes, k_blocks otherwise)
};

constexpr int kEpiSubs = 4;                        // epilogue warps per TMEM lane quarter
constexpr int kThreadsTc = 64 + 128 * kEpiSubs;    // warps 0..15: epilogue, warp 16: TMA producer, warp 17: MMA + TMEM alloc
constexpr int kWarpTma = 4 * kEpiSubs, kWarpMma = 4 * kEpiSubs + 1;
// (the SM's warp arbiter favours HIGH warp ids: the two single-thread critical roles sit above the 16 epilogue warps so
//  their instruction streams are never starved by epilogue warps waking up to poll their barriers)
enum { EPI_STORE = 0, EPI_CTC = 1, EPI_HEAD = 2 };

template <int EPI, int ACT>
__global__ void __launch_bounds__(kThreadsTc, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t span = (uint32_t)g.AW * 2u;
  const uint32_t a_bytes = g.patch ? (((uint32_t)(g.TH + g.KH - 1) * 8u * span + 1023u) & ~1023u) : 128u * span;
  const uint32_t a_tx = g.patch ? (uint32_t)(g.TH + g.KH - 1) * 8u * span : a_bytes;
  const uint32_t b_bytes = ((uint32_t)g.BN * span + 1023u) & ~1023u;
  const uint32_t stage_bytes = g.resident ? a_bytes : a_bytes + b_bytes;
  uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem_al;                                                         // resident B: [k_blocks][b_bytes]
  uint8_t* smem = smem_al + (g.resident ? (size_t)g.b_blocks * b_bytes : 0);          // stage ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)g.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + g.stages;
  uint64_t* tfull_bar = bars + 2 * g.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bfull_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4 * g.epi_subs); }
    mbar_init(bfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = g.tiles_m * g.tiles_n;
  const uint32_t tx_bytes = g.resident ? a_tx : a_bytes + (uint32_t)g.BN * span;
  // Tile walk shared by the three roles.  Divisions happen ONCE per CTA; every further tile is reached by
  // adds + carries (16 epilogue warps x a few runtime divisions per tile used to cost more issue slots than
  // the tile's actual work).
  struct Walk {
    int tm, tn, tx, ty, n;
    __device__ __forceinline__ void init(const Args& g) {
      if (g.resident) { tn = blockIdx.x % g.tiles_n; tm = blockIdx.x / g.tiles_n; }
      else { tm = blockIdx.x / g.tiles_n; tn = blockIdx.x % g.tiles_n; }
      tx = ty = n = 0;
      if (g.conv != 0) { tx = tm % g.tiles_x; ty = (tm / g.tiles_x) % g.tiles_y; n = tm / (g.tiles_x * g.tiles_y); }
    }
    __device__ __forceinline__ bool valid(const Args& g) const { return tm < g.tiles_m; }
    __device__ __forceinline__ void next(const Args& g) {
      tm += g.walk_dm; tn += g.walk_dn;
      if (tn >= g.tiles_n) { tn -= g.tiles_n; ++tm; }
      if (g.conv != 0) {
        tx += g.walk_sx; ty += g.walk_sy; n += g.walk_sn;
        if (tx >= g.tiles_x) { tx -= g.tiles_x; ++ty; }
        if (ty >= g.tiles_y) { ty -= g.tiles_y; ++n; }
      }
    }
  };
  if (warp == kWarpTma) {
    // ================= TMA producer =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      Walk wk; wk.init(g);
      if (g.resident && wk.valid(g)) {
        const int tn = wk.tn;
        // one-time load of this CTA's B n-tile: every k-block, laid out exactly like a stage's B buffer
        mbar_expect_tx(bfull_bar, (uint32_t)g.b_blocks * (uint32_t)g.BN * span);
        int kx = 0, cc = 0, bk = 0;
        for (int kb = 0; kb < g.b_blocks; ++kb) {
          const int kcoord = g.conv == 0 ? kb * g.AW : bk + cc * g.AW;
          tma_load_2d(smem_b + (size_t)kb * b_bytes, &tmB, bfull_bar, kcoord, tn * g.BN);
          if (g.conv != 0 && ++cc == g.cchunks) { cc = 0; bk += g.C; ++kx; }
        }
      }
      for (; wk.valid(g); wk.next(g)) {
        const int tm = wk.tm, tn = wk.tn;
        // per-tile coordinates hoisted out of the k loop: this single thread is the latency-critical
        // instruction stream of the CTA, so the k loop must stay free of integer divisions
        const int cn = wk.n;
        const int cx0 = wk.tx * g.TW * g.sw - g.pl;
        const int cy0 = wk.ty * g.TH * g.sh - g.pt;
        const int m_row = tm * 128, n_row = tn * g.BN;
        int kx = 0, ky = 0, cc = 0, bk = 0;   // conv: tap (ky,kx), channel chunk cc, B k-offset of the tap
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx_bytes);
          if (g.conv == 0) {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * g.AW, m_row);
            if (!g.resident) tma_load_2d(sa + a_bytes, &tmB, &full_bar[s], kb * g.AW, n_row);
          } else {
            tma_load_4d(sa, &tmA, &full_bar[s], cc * g.AW, cx0 + kx, cy0 + ky, cn);
            if (!g.resident) tma_load_2d(sa + a_bytes, &tmB, &full_bar[s], bk + cc * g.AW, n_row);
            if (++cc == g.cchunks) {
              cc = 0; bk += g.C;
              if (g.patch) ++kx;                       // patch mode walks (kx, cc); ky lives in the MMA loop
              else if (++kx == g.KW) { kx = 0; ++ky; }
            }
          }
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ================= MMA issuer (single thread) =================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      Walk wk; wk.init(g);
      if (g.resident && wk.valid(g)) mbar_wait(bfull_bar, 0);
      for (int it = 0; wk.valid(g); wk.next(g), ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * g.BN);
        int pkx = 0, pcc = 0;   // patch mode: (kx, channel chunk) of the stage being consumed
        for (int kb = 0; kb < g.k_blocks; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const int ksteps = g.AW / 16;
          if (!g.patch) {
            const uint64_t da = make_smem_desc(sa, span);
            const uint64_t db = make_smem_desc(g.resident ? smem_u32(smem_b + (size_t)kb * b_bytes) : sa + a_bytes, span);
            for (int kk = 0; kk < ksteps; ++kk) {
              // advancing K by 16 halves = 32 bytes inside the swizzle span: +2 in the (addr>>4) field
              umma_f16(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), g.idesc, (kb | kk) != 0 ? 1u : 0u);
            }
          } else {
            for (int ky = 0; ky < g.KH; ++ky) {
              // vertical tap ky = the same box, 8 rows (one 8-row swizzle atom, 8*span bytes) further down
              const uint64_t da = make_smem_desc(sa + (uint32_t)ky * 8u * span, span);
              const int bidx = (ky * g.KW + pkx) * g.cchunks + pcc;
              const uint64_t db = make_smem_desc(smem_u32(smem_b + (size_t)bidx * b_bytes), span);
              for (int kk = 0; kk < ksteps; ++kk)
                umma_f16(d_tmem, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), g.idesc, (kb | ky | kk) != 0 ? 1u : 0u);
            }
            if (++pcc == g.cchunks) { pcc = 0; ++pkx; }
          }
          umma_commit(&empty_bar[s]);
          if (kb == g.k_blocks - 1) umma_commit(&tfull_bar[as]);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue warps (TMEM -> registers -> global) =================
    // 16 warps: warp w may only touch TMEM lanes [32*(w%4), +32) (hardware rule), so the four warps
    // sharing a lane quarter split the tile's 16-column chunks between them (sub = 0..3).
    const int q = warp & 3;
    const int sub = warp >> 2;
    Walk wk; wk.init(g);
    for (int it = 0; sub < g.epi_subs && wk.valid(g); wk.next(g), ++it) {
      const int tm = wk.tm, tn = wk.tn;
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      long long row = (long long)tm * 128 + q * 32 + lane;   // GEMM: output row; conv: output pixel index
      bool row_ok = row < g.M;
      if (g.conv != 0) {
        const int r = q * 32 + lane;
        const int oy = wk.ty * g.TH + (r >> g.tw_shift), ox = wk.tx * g.TW + (r & (g.TW - 1));
        row_ok = (oy < g.OH) && (ox < g.OW);
        row = ((long long)wk.n * g.OH + oy) * (g.out_wp ? g.out_wp : g.OW) + ox;
      }
      const int n0 = tn * g.BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * g.BN);
      if (EPI == EPI_HEAD) {
        // sub = ConvT tap (dy,dx) of the first up-conv; this thread owns its 24 hidden channels of one pixel
        uint32_t r16[16], r8[8];
        tmem_ld16(taddr + (uint32_t)(24 * sub), r16);
        tmem_ld8(taddr + (uint32_t)(24 * sub + 16), r8);
        float o[4];
        {
          const float bf = __ldg(g.b_fin);
          o[0] = bf; o[1] = bf; o[2] = bf; o[3] = bf;
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 24; ++c) {
          const float acc = __uint_as_float(c < 16 ? r16[c < 16 ? c : 0] : r8[c >= 16 ? c - 16 : 0]);
          const float hv = fmaxf(acc + __ldg(g.bias + c), 0.f);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = fmaf(hv, __ldg(g.w_fin + k * 24 + c), o[k]);
        }
        if (row_ok) {
          const int hw = g.hH * g.hW;
          const int n = (int)(row / hw), rem = (int)(row % hw);
          const int y = rem / g.hW, x = rem % g.hW;
          const int W4 = 4 * g.hW;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float pz = 1.f / (1.f + expf(-o[k]));
            if (pz != pz) pz = 0.f;
            o[k] = pz;
          }
          const long long base = ((long long)n * (4 * g.hH) + 4 * y + 2 * (sub >> 1)) * W4 + 4 * x + 2 * (sub & 1);
          *reinterpret_cast<float2*>(g.prob + base) = make_float2(o[0], o[1]);
          *reinterpret_cast<float2*>(g.prob + base + W4) = make_float2(o[2], o[3]);
          if (g.seg != nullptr) {
            *reinterpret_cast<uchar2*>(g.seg + base) = make_uchar2(o[0] > g.thresh, o[1] > g.thresh);
            *reinterpret_cast<uchar2*>(g.seg + base + W4) = make_uchar2(o[2] > g.thresh, o[3] > g.thresh);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        continue;
      }
      float cmax = -INFINITY, csum = 0.f; int cidx = 0x7fffffff;
      for (int c0 = sub * 16; c0 < g.BN; c0 += 16 * kEpiSubs) {
        if (n0 + c0 >= g.N) break;                 // warp-uniform: the rest of this n-tile is padding
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        const bool full = (n0 + c0 + 16 <= g.N);
        float v[16];
        if (g.bias != nullptr) {
          if (full) {
            const float4* bp = reinterpret_cast<const float4*>(g.bias + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) { float4 b4 = __ldg(bp + j); v[4 * j] = b4.x; v[4 * j + 1] = b4.y; v[4 * j + 2] = b4.z; v[4 * j + 3] = b4.w; }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (n0 + c0 + j < g.N) ? __ldg(g.bias + n0 + c0 + j) : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r[j]);
        if (EPI == EPI_STORE) {
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = apply_act<ACT>(v[j]);
            if (g.colscale != nullptr) {
              const int img = (int)(row / g.rows_per_img);
              const float* sp = g.colscale + (long long)img * g.N + n0 + c0;
#pragma unroll
              for (int j = 0; j < 16; ++j) if (full || n0 + c0 + j < g.N) v[j] *= __ldg(sp + j);
              if (g.up_res != nullptr) {
                const int rem = (int)(row - (long long)img * g.rows_per_img);
                const int y = rem / g.up_W, x = rem - y * g.up_W;
                const __half* up = g.up_res + (((long long)img * (g.up_H >> 1) + (y >> 1)) * (g.up_W >> 1) + (x >> 1)) * g.N + n0 + c0;
                if (full) {
                  float a[8], b2[8];
                  Vec8<__half>::load(up, a);
                  Vec8<__half>::load(up + 8, b2);
#pragma unroll
                  for (int j = 0; j < 8; ++j) { v[j] += a[j]; v[8 + j] += b2[j]; }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) if (n0 + c0 + j < g.N) v[j] += __half2float(up[j]);
                }
              }
            }
            if (g.res != nullptr) {
              const __half* rp = g.res + row * g.ldr + n0 + c0;
              if (full) {
                float a[8], b2[8];
                Vec8<__half>::load(rp, a);
                Vec8<__half>::load(rp + 8, b2);
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] += a[j]; v[8 + j] += b2[j]; }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (n0 + c0 + j < g.N) v[j] += __half2float(rp[j]);
              }
            }
            __half* op = g.out + row * g.ldc + g.c_off + n0 + c0;
            if (full) {
              float a[8], b2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { a[j] = v[j]; b2[j] = v[8 + j]; }
              Vec8<__half>::store(op, a);
              Vec8<__half>::store(op + 8, b2);
            } else {
              // ragged last chunk (N = 24 -> 8 valid columns): still one 16-byte store for the first 8
              // (every loop over v[] is fully unrolled with compile-time indices: one dynamic index would move the
              // whole accumulator row to local memory)
              const int valid = g.N - (n0 + c0);
              if (valid >= 8) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = v[j];
                Vec8<__half>::store(op, a);
#pragma unroll
                for (int j = 8; j < 16; ++j) if (j < valid) op[j] = __float2half_rn(v[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (j < valid) op[j] = __float2half_rn(v[j]);
              }
            }
          }
        } else {
          // fused greedy-decode partials: (max, argmax, sum exp(x - max)) of this 16-column chunk computed with
          // independent operations (tree max, 16 parallel exps), then ONE merge into the running row state —
          // a per-element online update would serialise 16 dependent exp/compare steps per chunk
          float m16 = -INFINITY; int i16 = 0x7fffffff;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (n0 + c0 + j >= g.N) v[j] = -INFINITY;
            if (v[j] > m16) { m16 = v[j]; i16 = n0 + c0 + j; }      // ascending j: ties keep the lowest index
          }
          float s16 = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) s16 += __expf(v[j] - m16);   // exp(-inf) = 0 for masked columns
          if (m16 > cmax) { csum = csum * __expf(cmax - m16) + s16; cmax = m16; cidx = i16; }
          else csum += s16 * __expf(m16 - cmax);
        }
      }
      if (EPI == EPI_CTC && row_ok) {
        const long long pi = row * ((long long)g.tiles_n * kEpiSubs) + tn * kEpiSubs + sub;
        g.pmax[pi] = cmax;
        g.pidx[pi] = cidx;
        g.psum[pi] = csum;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

// ------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    RDB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    RDB_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuda: cuTensorMapEncodeTiled unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp16 row-major [rows, cols] with row pitch ld (elements); box = [box_rows, aw] k-major
inline CUtensorMap make_map(const void* base, long long rows, int cols, int ld, int aw, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)aw, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapSwizzle sw = aw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (aw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  RDB_CHECK(((uintptr_t)base & 15) == 0 && (ld * 2) % 16 == 0, "tma: base/pitch must be 16-byte aligned");
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuda: cuTensorMapEncodeTiled failed, code " + std::to_string((int)r));
  return m;
}

// 4-D fp16 NHWC tensor {C, W, H, n}; box {aw, TW*sw, TH*sh, 1} traversed with element strides {1, sw, sh, 1}
// ld: pixel pitch in elements (0 = C): a channel SLICE of a wider NHWC buffer (the formula encoder's dense-concat buffers) is a
// valid conv input — dims[0] = C bounds the slice, channels beyond it are zero-filled by TMA
inline CUtensorMap make_map_nhwc(const void* base, int n, int H, int W, int C, int aw, int TH, int TW, int sh, int sw, int ld = 0) {
  CUtensorMap m;
  const cuuint64_t P = ld > 0 ? ld : C;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {P * 2, (cuuint64_t)W * P * 2, (cuuint64_t)H * W * P * 2};
  cuuint32_t box[4] = {(cuuint32_t)aw, (cuuint32_t)(TW * sw), (cuuint32_t)(TH * sh), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1};
  CUtensorMapSwizzle swz = aw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (aw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  RDB_CHECK(((uintptr_t)base & 15) == 0 && (P * 2) % 16 == 0, "tma: NHWC base/channel pitch must be 16-byte aligned");
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuda: cuTensorMapEncodeTiled(4d) failed, code " + std::to_string((int)r));
  return m;
}

// "Wide-row" view of an NHWC tensor whose rows are padded to WP pixels: row (x, y, n) = KW adjacent pixels = KW*C contiguous
// channels starting at padded pixel x (dim-1 stride = ONE pixel, i.e. rows overlap — accepted and delivered by TMA, see
// tools/tma_overlap.cu).  The KW horizontal taps of a conv become one K range; zero pad pixels in memory replace x-OOB fill.
inline CUtensorMap make_map_wide(const void* base, int n, int H, int W, int WP, int C, int KW, int aw, int box_rows_y) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)KW * C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)WP * C * 2, (cuuint64_t)H * WP * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)aw, 8, (cuuint32_t)box_rows_y, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle swz = aw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (aw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  RDB_CHECK(((uintptr_t)base & 15) == 0 && (C * 2) % 16 == 0, "tma: wide-row base/pixel pitch must be 16-byte aligned");
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuda: cuTensorMapEncodeTiled(wide rows) failed, code " + std::to_string((int)r));
  return m;
}

struct Plan {
  Args a;
  size_t smem;
  int grid;
};

// k-block width (one swizzle span): prefer few wide TMA boxes; a ragged last block is zero-filled by
// TMA (no HBM traffic for the out-of-bounds part), e.g. K=48 -> one 64-wide block with 16 zero columns.
inline int pick_aw(int K) {
  // TMA cost is per box row ("segment"), not per byte (profiles/r01_tma_microbench.txt): use the widest span (64 halves =
  // 128 B) as soon as K > 32 even when the last block is only partly valid (K = 48, 96, 120, 240 ...)
  const char* e = sw_get("RDB_TC_AW");
  if (e != nullptr && e[0] == 'n') {            // RDB_TC_AW=narrow: previous rule (largest span dividing K), for A/B runs
    if (K % 64 == 0) return 64;
    if (K % 32 == 0) return 32;
    if (K <= 16) return 16;
    if (K <= 32) return 32;
    return 64;
  }
  if (K <= 16) return 16;
  if (K <= 32) return 32;
  return 64;
}

inline int pick_bn(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  // fewest tiles, then least padding
  int best = 256, best_tiles = (N + 255) / 256, best_pad = best_tiles * 256 - N;
  for (int bn = 240; bn >= 128; bn -= 16) {
    int tiles = (N + bn - 1) / bn, pad = tiles * bn - N;
    if (tiles < best_tiles || (tiles == best_tiles && pad < best_pad)) { best = bn; best_tiles = tiles; best_pad = pad; }
  }
  return best;
}

// stage count / residency / grid, common to GEMM and conv plans
inline void finish_plan(Plan& p, int num_sms) {
  Args& a = p.a;
  const size_t span = (size_t)a.AW * 2;
  const size_t a_bytes = a.patch ? (((size_t)(a.TH + a.KH - 1) * 8 * span + 1023) & ~(size_t)1023) : 128 * span;
  const size_t b_bytes = ((size_t)a.BN * span + 1023) & ~(size_t)1023;
  const size_t budget = 200 * 1024;
  if (a.b_blocks == 0) a.b_blocks = a.k_blocks;
  const size_t b_res = (size_t)a.b_blocks * b_bytes;
  const char* e = sw_get("RDB_TC_RESIDENT");
  const bool allow = !(e != nullptr && e[0] == '0');
  int min_stages = a.k_blocks < 3 ? a.k_blocks + 2 : 4;
  if (allow && a.tiles_n <= num_sms && b_res + (size_t)min_stages * a_bytes <= budget) {
    a.resident = 1;
    int stages = (int)((budget - b_res) / a_bytes);
    if (stages > 12) stages = 12;
    int want = a.k_blocks * 3;
    if (stages > want) stages = want;
    if (stages < 2) stages = 2;
    a.stages = stages;
    a.ctas_per_n = num_sms / a.tiles_n;
    if (a.ctas_per_n > a.tiles_m) a.ctas_per_n = a.tiles_m;
    if (a.ctas_per_n < 1) a.ctas_per_n = 1;
    p.grid = a.ctas_per_n * a.tiles_n;
    p.smem = 1024 + b_res + stages * a_bytes + (2 * stages + 8) * 8 + 16;
    a.walk_dm = a.ctas_per_n; a.walk_dn = 0;
  } else {
    a.resident = 0;
    const size_t stage = a_bytes + b_bytes;
    int stages = (int)(budget / stage);
    if (stages > 8) stages = 8;
    int want = a.k_blocks * 3;
    if (stages > want) stages = want;
    if (stages < 2) stages = 2;
    a.stages = stages;
    p.smem = 1024 + stages * stage + (2 * stages + 8) * 8 + 16;
    long long tiles = (long long)a.tiles_m * a.tiles_n;
    p.grid = (int)(tiles < num_sms ? tiles : num_sms);
    a.walk_dm = p.grid / a.tiles_n; a.walk_dn = p.grid % a.tiles_n;
  }
  if (a.conv) {
    a.walk_sx = a.walk_dm % a.tiles_x;
    a.walk_sy = (a.walk_dm / a.tiles_x) % a.tiles_y;
    a.walk_sn = a.walk_dm / (a.tiles_x * a.tiles_y);
  }
  if (a.epi_subs == 0) {   // sub-warps that own >= 1 chunk (EPI_HEAD overrides to 4)
    const int chunks = a.BN / 16;
    a.epi_subs = chunks < kEpiSubs ? chunks : kEpiSubs;
  }
}

inline Plan make_plan(long long M, int N, int K, int num_sms) {
  Plan p{};
  Args& a = p.a;
  a.M = (int)M; a.N = N; a.K = K;
  a.AW = pick_aw(K);
  a.k_blocks = (K + a.AW - 1) / a.AW;
  a.BN = pick_bn(N);
  a.tiles_m = (int)((M + 127) / 128);
  a.tiles_n = (N + a.BN - 1) / a.BN;
  int cols = 32;
  while (cols < 2 * a.BN) cols *= 2;
  a.tmem_cols = cols;
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c=F32 [4,6)=1, a/b = F16 (0),
  // a_major/b_major = K (0), n>>3 at [17,23), m>>4 at [24,29)
  a.idesc = (1u << 4) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  finish_plan(p, num_sms);
  return p;
}

// conv plan: out[n,OH,OW,N] = act(conv(in[n,H,W,C], w[N][KH][KW][C]) + bias)
inline Plan make_conv_plan_mode(int n, int H, int W, int C, int N, int KH, int KW, int sh, int sw, int pt, int pl, int OH, int OW, int num_sms,
                                bool patch) {
  Plan p{};
  Args& a = p.a;
  a.conv = 1;
  a.N = N; a.K = KH * KW * C; a.C = C;
  a.AW = pick_aw(C);
  a.cchunks = (C + a.AW - 1) / a.AW;
  a.KH = KH;
  a.b_blocks = KH * KW * a.cchunks;
  a.patch = patch ? 1 : 0;
  a.k_blocks = patch ? KW * a.cchunks : KH * KW * a.cchunks;   // smem stages consumed per tile
  a.BN = pick_bn(N);
  a.OH = OH; a.OW = OW; a.KW = KW; a.sh = sh; a.sw = sw; a.pt = pt; a.pl = pl;
  a.TW = patch ? 8 : (OW >= 16 ? 16 : (OW >= 8 ? 8 : 4));
  a.TH = 128 / a.TW;
  a.tw_shift = a.TW == 16 ? 4 : (a.TW == 8 ? 3 : 2);
  a.tiles_x = (OW + a.TW - 1) / a.TW;
  a.tiles_y = (OH + a.TH - 1) / a.TH;
  a.tiles_m = n * a.tiles_x * a.tiles_y;
  a.tiles_n = (N + a.BN - 1) / a.BN;
  a.M = n * OH * OW;
  int cols = 32;
  while (cols < 2 * a.BN) cols *= 2;
  a.tmem_cols = cols;
  a.idesc = (1u << 4) | ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  finish_plan(p, num_sms);
  return p;
}

// wide-row patch plan (stride-1 conv on a row-padded input): K of one vertical tap = KW*C contiguous channels
inline Plan make_conv_plan_wide(int n, int H, int W, int C, int N, int KH, int KW, int pt, int OH, int OW, int num_sms) {
  // identical to a patch-mode conv with KW_eff = 1 and C_eff = KW*C read through make_map_wide (pl is provided by the pad pixels)
  Plan p = make_conv_plan_mode(n, H, W, KW * C, N, KH, 1, 1, 1, pt, 0, OH, OW, num_sms, true);
  return p;
}

inline Plan make_conv_plan(int n, int H, int W, int C, int N, int KH, int KW, int sh, int sw, int pt, int pl, int OH, int OW, int num_sms) {
  const char* e = sw_get("RDB_TC_PATCH");
  const bool allow = !(e != nullptr && e[0] == '0');
  if (allow && sh == 1 && sw == 1 && OW >= 8) {
    Plan p = make_conv_plan_mode(n, H, W, C, N, KH, KW, sh, sw, pt, pl, OH, OW, num_sms, true);
    if (p.a.resident) return p;   // patch mode needs the weights resident (B is indexed by tap inside the MMA loop)
  }
  return make_conv_plan_mode(n, H, W, C, N, KH, KW, sh, sw, pt, pl, OH, OW, num_sms, false);
}

}  // namespace tc
}  // namespace rdb
