// SLANet structure decoder (SLAHead: GRU attention loop) on the device — SURVEY row T4.
//
// What it replaces: the `Loop` node of slanet-1m.onnx / SLANet_plus as onnxruntime executes it inside
// `OrtInferSession.__call__` (rapid_doc/model/table/rapid_table_self/inference_engine/onnxruntime/main.py:70-76, called by
// PPTableStructurer.__call__, table_structure/pp_structure/main.py:41-51): up to 501 sequential steps of
//   onehot(prev) ; e = w_s . tanh(H Wi + (h Wh + bh)) ; alpha = softmax(e) ; ctx = alpha H ;
//   h' = GRU([ctx, onehot], h) ; logits = W4 (W3 h' + b3) + b4 ; loc = sigmoid(W6 (W5 h' + b5) + b6) ; prev = argmax(logits)
// stopping after the first step at which EVERY row of the batch has emitted the end token.
//
// B200 mapping: the whole loop is ONE launch.  One persistent CTA per table image keeps h / attention scores in shared memory
// and walks its own sequence; the 2.2 MB of weights are shared by all CTAs and stay L2-resident.  The only cross-CTA coupling
// of the reference loop — the common stop step — is two global words (count of finished rows, max first-eos step) that each
// CTA publishes once and polls once per step; no grid barrier.  H Wi is hoisted out of the loop (it does not depend on h).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdint>
#include <string>

#include "common.cuh"
#include "engine.cuh"
#include "../../include/rapiddoc_b200.h"

namespace rdb {
namespace sla {

constexpr int HID = 256, THREADS = 256, MAX_CLASSES = 64, MAX_LOC = 8;

struct Weights {
  const float* Wh;   // [HID][HID]   h2h, [in][out]
  const float* bh;   // [HID]
  const float* ws;   // [HID]        score
  const float* WihT; // [C + classes][3*HID]
  const float* WhhT; // [HID][3*HID]
  const float* bih;  // [3*HID]
  const float* bhh;  // [3*HID]
  const float* W3; const float* b3;   // [HID][HID]
  const float* W4; const float* b4;   // [HID][classes]
  const float* W5; const float* b5;   // [HID][HID]
  const float* W6; const float* b6;   // [HID][loc]
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// y[j] = b[j] + sum_k x[k] * W[k][ld*0 + j]  for the thread's column j; x in shared memory; 4 independent accumulators
__device__ __forceinline__ float matvec_col(const float* __restrict__ W, int ld, int K, const float* x, int j) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    a0 = fmaf(x[k], __ldg(W + (long long)k * ld + j), a0);
    a1 = fmaf(x[k + 1], __ldg(W + (long long)(k + 1) * ld + j), a1);
    a2 = fmaf(x[k + 2], __ldg(W + (long long)(k + 2) * ld + j), a2);
    a3 = fmaf(x[k + 3], __ldg(W + (long long)(k + 3) * ld + j), a3);
  }
  for (; k < K; ++k) a0 = fmaf(x[k], __ldg(W + (long long)k * ld + j), a0);
  return (a0 + a1) + (a2 + a3);
}

// sync words: [0] rows that have emitted eos, [1] max over rows of the first eos step
__global__ void __launch_bounds__(THREADS) sla_decode_kernel(const float* __restrict__ H, const float* __restrict__ Hp, int HW, int C, Weights w, int classes,
                                                            int loc_dim, int max_steps, int eos, float* __restrict__ logits_out, float* __restrict__ loc_out,
                                                            int* __restrict__ ids_out, int* __restrict__ sync_words, int* __restrict__ steps_run) {
  extern __shared__ float sm[];
  float* h = sm;                    // [HID]
  float* hp = h + HID;              // [HID]
  float* hn = hp + HID;             // [HID]
  float* t3 = hn + HID;             // [HID]
  float* t5 = t3 + HID;             // [HID]
  float* ctx = t5 + HID;            // [C] (+ 2 x C partials)
  float* part = ctx + C;            // [2][C]
  float* red = part + 2 * C;        // [32]
  float* lg = red + 32;             // [MAX_CLASSES + MAX_LOC]
  float* e = lg + MAX_CLASSES + MAX_LOC;   // [HW]
  __shared__ int s_prev, s_stop;

  const int B = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Hb = H + (long long)b * HW * C;
  const float* Hpb = Hp + (long long)b * HW * HID;
  volatile int* sw = sync_words;
  h[tid] = 0.f;
  if (tid == 0) { s_prev = 0; s_stop = 0; }
  bool had_eos = false;           // thread 0 only
  __syncthreads();

  int i = 0;
  for (; i < max_steps; ++i) {
    // 1. hp = h Wh + bh
    hp[tid] = matvec_col(w.Wh, HID, HID, h, tid) + __ldg(w.bh + tid);
    __syncthreads();
    // 2. attention scores: one warp per position, lanes over the hidden axis
    {
      float wsv[HID / 32], hpv[HID / 32];
#pragma unroll
      for (int t = 0; t < HID / 32; ++t) { wsv[t] = __ldg(w.ws + lane + 32 * t); hpv[t] = hp[lane + 32 * t]; }
      for (int p = warp; p < HW; p += THREADS / 32) {
        const float* row = Hpb + (long long)p * HID;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < HID / 32; ++t) s = fmaf(wsv[t], tanhf(__ldg(row + lane + 32 * t) + hpv[t]), s);
        s = warp_sum(s);
        if (lane == 0) e[p] = s;
      }
    }
    __syncthreads();
    // 3. softmax over the positions
    {
      float m = -INFINITY;
      for (int p = tid; p < HW; p += THREADS) m = fmaxf(m, e[p]);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) red[warp] = m;
      __syncthreads();
      m = red[0];
#pragma unroll
      for (int q = 1; q < THREADS / 32; ++q) m = fmaxf(m, red[q]);
      __syncthreads();
      float s = 0.f;
      for (int p = tid; p < HW; p += THREADS) { const float v = expf(e[p] - m); e[p] = v; s += v; }
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      s = 0.f;
#pragma unroll
      for (int q = 0; q < THREADS / 32; ++q) s += red[q];
      const float inv = 1.f / s;
      __syncthreads();
      for (int p = tid; p < HW; p += THREADS) e[p] *= inv;
    }
    __syncthreads();
    // 4. ctx = alpha H  (two halves of the positions on 2 x C threads)
    if (tid < 2 * C) {
      const int c = tid % C, half = tid / C;
      float a0 = 0.f, a1 = 0.f;
      int p = half;
      for (; p + 2 < HW; p += 4) {
        a0 = fmaf(e[p], __ldg(Hb + (long long)p * C + c), a0);
        a1 = fmaf(e[p + 2], __ldg(Hb + (long long)(p + 2) * C + c), a1);
      }
      for (; p < HW; p += 2) a0 = fmaf(e[p], __ldg(Hb + (long long)p * C + c), a0);
      part[half * C + c] = a0 + a1;
    }
    __syncthreads();
    if (tid < C) ctx[tid] = part[tid] + part[C + tid];
    __syncthreads();
    // 5. GRU cell on [ctx, onehot(prev)] and h: thread j owns hidden unit j (its r, z, c columns)
    {
      const int j = tid, prev = s_prev;
      float x[3], g[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        x[q] = matvec_col(w.WihT, 3 * HID, C, ctx, q * HID + j) + __ldg(w.WihT + (long long)(C + prev) * 3 * HID + q * HID + j) + __ldg(w.bih + q * HID + j);
        g[q] = matvec_col(w.WhhT, 3 * HID, HID, h, q * HID + j) + __ldg(w.bhh + q * HID + j);
      }
      const float r = sigmoidf_(x[0] + g[0]);
      const float z = sigmoidf_(x[1] + g[1]);
      const float cand = tanhf(x[2] + r * g[2]);
      hn[j] = (h[j] - cand) * z + cand;
    }
    __syncthreads();
    // 6. first layers of the two generators
    t3[tid] = matvec_col(w.W3, HID, HID, hn, tid) + __ldg(w.b3 + tid);
    t5[tid] = matvec_col(w.W5, HID, HID, hn, tid) + __ldg(w.b5 + tid);
    h[tid] = hn[tid];
    __syncthreads();
    // 7. logits (classes) and box (loc_dim): one warp per output, lanes over k
    for (int n = warp; n < classes + loc_dim; n += THREADS / 32) {
      const bool is_loc = n >= classes;
      const int col = is_loc ? n - classes : n;
      const float* W = is_loc ? w.W6 : w.W4;
      const int ld = is_loc ? loc_dim : classes;
      const float* x = is_loc ? t5 : t3;
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < HID / 32; ++t) s = fmaf(x[lane + 32 * t], __ldg(W + (long long)(lane + 32 * t) * ld + col), s);
      s = warp_sum(s);
      if (lane == 0) lg[n] = is_loc ? sigmoidf_(s + __ldg(w.b6 + col)) : s + __ldg(w.b4 + col);
    }
    __syncthreads();
    // 8. outputs of the step, greedy token, stop bookkeeping
    if (tid < classes) logits_out[((long long)b * max_steps + i) * classes + tid] = lg[tid];
    else if (tid >= 64 && tid < 64 + loc_dim) loc_out[((long long)b * max_steps + i) * loc_dim + (tid - 64)] = lg[classes + tid - 64];
    if (tid == 0) {
      int best = 0;
      float bv = lg[0];
      for (int n = 1; n < classes; ++n) if (lg[n] > bv) { bv = lg[n]; best = n; }     // first maximum, as ArgMax(select_last_index=0)
      s_prev = best;
      ids_out[(long long)b * max_steps + i] = best;
      if (best == eos && !had_eos) {
        had_eos = true;
        atomicMax(sync_words + 1, i);
        __threadfence();
        atomicAdd(sync_words + 0, 1);
      }
      // the reference loop ends after the first step at which every row has an eos: that step is max_b(first eos step)
      s_stop = (sw[0] == B && i >= sw[1]) ? 1 : 0;
    }
    __syncthreads();
    if (s_stop) { ++i; break; }
  }
  if (tid == 0) steps_run[b] = i;
}


// ---- the same loop spread over a thread-block cluster ------------------------------------------------------------------------
// One CTA per image is bound by what one SM can pull from L2 (2.2 MB of weights + the image's projected features per step).
// Here a cluster of CL CTAs owns one image: CTA r holds hidden units [r*HID/CL, (r+1)*HID/CL) of every matrix-vector product
// and 1/CL of the attention positions, so each SM streams 1/CL of the bytes; the slices are all-gathered through distributed
// shared memory (remote stores + three cluster barriers per step).  Attention is a two-level softmax: every CTA reduces its own
// positions relative to its local maximum and ships (partial context, max, sum); the combine is exact up to fp32 rounding.
namespace cg = cooperative_groups;

template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS) sla_decode_cluster_kernel(
    const float* __restrict__ H, const float* __restrict__ Hp, int HW, int C, Weights w, int classes, int loc_dim, int max_steps, int eos,
    float* __restrict__ logits_out, float* __restrict__ loc_out, int* __restrict__ ids_out, int* __restrict__ sync_words, int* __restrict__ steps_run) {
  constexpr int SL = HID / CL;        // hidden units owned by this CTA
  constexpr int Q = THREADS / SL;     // K-split of every matrix-vector product
  constexpr int XW = 132;             // exchange record: partial context (<= 128), local max, local sum
  __shared__ float h_buf[2][HID];
  __shared__ float hp[HID], t3[HID], t5[HID];
  __shared__ float ctx[128];
  __shared__ float xch[CL][XW];
  __shared__ float part[2][128];
  __shared__ float e[1024 / CL];
  __shared__ float4 red[Q][SL];
  __shared__ float lg[MAX_CLASSES + MAX_LOC];
  __shared__ float redw[32];
  __shared__ int s_prev, s_stop;

  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int B = gridDim.x / CL, b = blockIdx.x / CL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = tid % SL, q = tid / SL;
  const int PP = (HW + CL - 1) / CL, p0 = r * PP, p1 = min(HW, p0 + PP);
  const float* Hb = H + (long long)b * HW * C;
  const float* Hpb = Hp + (long long)b * HW * HID;
  volatile int* sw = sync_words;

  h_buf[0][tid] = 0.f;
  hp[tid] = __ldg(w.bh + tid);                       // h = 0: h Wh + bh = bh
  if (tid == 0) { s_prev = 0; s_stop = 0; }
  bool had_eos = false;                              // rank 0, thread 0
  int cur = 0, i = 0;
  cluster.sync();

  for (; i < max_steps; ++i) {
    // ---- A: attention over this CTA's positions, relative to the local maximum
    {
      float wsv[HID / 32], hpv[HID / 32];
#pragma unroll
      for (int t = 0; t < HID / 32; ++t) { wsv[t] = __ldg(w.ws + lane + 32 * t); hpv[t] = hp[lane + 32 * t]; }
      for (int p = p0 + warp; p < p1; p += THREADS / 32) {
        const float* row = Hpb + (long long)p * HID;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < HID / 32; ++t) s = fmaf(wsv[t], tanhf(__ldg(row + lane + 32 * t) + hpv[t]), s);
        s = warp_sum(s);
        if (lane == 0) e[p - p0] = s;
      }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int p = tid; p < p1 - p0; p += THREADS) m = fmaxf(m, e[p]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) redw[warp] = m;
    __syncthreads();
    m = redw[0];
#pragma unroll
    for (int t = 1; t < THREADS / 32; ++t) m = fmaxf(m, redw[t]);
    __syncthreads();
    float ssum = 0.f;
    for (int p = tid; p < p1 - p0; p += THREADS) { const float v = expf(e[p] - m); e[p] = v; ssum += v; }
    ssum = warp_sum(ssum);
    if (lane == 0) redw[warp] = ssum;
    __syncthreads();
    ssum = 0.f;
#pragma unroll
    for (int t = 0; t < THREADS / 32; ++t) ssum += redw[t];
    if (tid < 2 * C) {
      const int c = tid % C, half = tid / C;
      float a0 = 0.f, a1 = 0.f;
      int p = p0 + half;
      for (; p + 2 < p1; p += 4) {
        a0 = fmaf(e[p - p0], __ldg(Hb + (long long)p * C + c), a0);
        a1 = fmaf(e[p + 2 - p0], __ldg(Hb + (long long)(p + 2) * C + c), a1);
      }
      for (; p < p1; p += 2) a0 = fmaf(e[p - p0], __ldg(Hb + (long long)p * C + c), a0);
      part[half][c] = a0 + a1;
    }
    __syncthreads();
    if (tid < C + 2) {
      const float v = tid < C ? part[0][tid] + part[1][tid] : (tid == C ? m : ssum);
      const int slot = tid < C ? tid : 128 + (tid - C);
#pragma unroll
      for (int peer = 0; peer < CL; ++peer) cluster.map_shared_rank(&xch[0][0], peer)[r * XW + slot] = v;
    }
    cluster.sync();                                                               // (1) partial contexts exchanged
    if (s_stop) break;                       // set by rank 0 during the previous step, visible to every CTA after this barrier
    if (tid < C) {
      float M = xch[0][128];
#pragma unroll
      for (int t = 1; t < CL; ++t) M = fmaxf(M, xch[t][128]);
      float S = 0.f, a = 0.f;
#pragma unroll
      for (int t = 0; t < CL; ++t) {
        const float f = (xch[t][128] == -INFINITY) ? 0.f : expf(xch[t][128] - M);
        S = fmaf(xch[t][129], f, S);
        a = fmaf(xch[t][tid], f, a);
      }
      ctx[tid] = a / S;
    }
    __syncthreads();
    // ---- B: GRU cell for this CTA's hidden units; K split over Q thread groups
    {
      const float* hc = h_buf[cur];
      const int col = SL * r + j, prev = s_prev;
      float ar = 0.f, az = 0.f, axc = 0.f, ahc = 0.f;
      for (int k = q; k < C; k += Q) {
        const float x = ctx[k];
        const float* row = w.WihT + (long long)k * 3 * HID + col;
        ar = fmaf(x, __ldg(row), ar); az = fmaf(x, __ldg(row + HID), az); axc = fmaf(x, __ldg(row + 2 * HID), axc);
      }
#pragma unroll 4
      for (int k = q; k < HID; k += Q) {
        const float x = hc[k];
        const float* row = w.WhhT + (long long)k * 3 * HID + col;
        ar = fmaf(x, __ldg(row), ar); az = fmaf(x, __ldg(row + HID), az); ahc = fmaf(x, __ldg(row + 2 * HID), ahc);
      }
      if (q == 0) {
        const float* row = w.WihT + (long long)(C + prev) * 3 * HID + col;
        ar += __ldg(row) + __ldg(w.bih + col) + __ldg(w.bhh + col);
        az += __ldg(row + HID) + __ldg(w.bih + HID + col) + __ldg(w.bhh + HID + col);
        axc += __ldg(row + 2 * HID) + __ldg(w.bih + 2 * HID + col);
        ahc += __ldg(w.bhh + 2 * HID + col);
      }
      red[q][j] = make_float4(ar, az, axc, ahc);
      __syncthreads();
      if (tid < SL) {
        float4 t = red[0][tid];
#pragma unroll
        for (int u = 1; u < Q; ++u) { const float4 v = red[u][tid]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        const float rg = sigmoidf_(t.x), zg = sigmoidf_(t.y);
        const float cand = tanhf(t.z + rg * t.w);
        const float hn = (hc[SL * r + tid] - cand) * zg + cand;
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) cluster.map_shared_rank(&h_buf[0][0], peer)[(cur ^ 1) * HID + SL * r + tid] = hn;
      }
    }
    cluster.sync();                                                               // (2) new hidden state gathered
    // ---- C: next step's h Wh + bh and the first generator layers, for this CTA's columns
    {
      const float* hn = h_buf[cur ^ 1];
      const int col = SL * r + j;
      float a1 = 0.f, a3 = 0.f, a5 = 0.f;
#pragma unroll 4
      for (int k = q; k < HID; k += Q) {
        const float x = hn[k];
        a1 = fmaf(x, __ldg(w.Wh + (long long)k * HID + col), a1);
        a3 = fmaf(x, __ldg(w.W3 + (long long)k * HID + col), a3);
        a5 = fmaf(x, __ldg(w.W5 + (long long)k * HID + col), a5);
      }
      red[q][j] = make_float4(a1, a3, a5, 0.f);
      __syncthreads();
      if (tid < SL) {
        float4 t = red[0][tid];
#pragma unroll
        for (int u = 1; u < Q; ++u) { const float4 v = red[u][tid]; t.x += v.x; t.y += v.y; t.z += v.z; }
        const int c2 = SL * r + tid;
        const float v1 = t.x + __ldg(w.bh + c2), v3 = t.y + __ldg(w.b3 + c2), v5 = t.z + __ldg(w.b5 + c2);
#pragma unroll
        for (int peer = 0; peer < CL; ++peer) {
          cluster.map_shared_rank(&hp[0], peer)[c2] = v1;
          cluster.map_shared_rank(&t3[0], peer)[c2] = v3;
          cluster.map_shared_rank(&t5[0], peer)[c2] = v5;
        }
      }
    }
    cluster.sync();                                                               // (3) hp / t3 / t5 gathered
    // ---- D: logits and box on every CTA (identical arithmetic, so every CTA derives the same token); rank 0 publishes
    for (int n = warp; n < classes + loc_dim; n += THREADS / 32) {
      const bool is_loc = n >= classes;
      const int col = is_loc ? n - classes : n;
      const float* W = is_loc ? w.W6 : w.W4;
      const int ld = is_loc ? loc_dim : classes;
      const float* x = is_loc ? t5 : t3;
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < HID / 32; ++t) s = fmaf(x[lane + 32 * t], __ldg(W + (long long)(lane + 32 * t) * ld + col), s);
      s = warp_sum(s);
      if (lane == 0) lg[n] = is_loc ? sigmoidf_(s + __ldg(w.b6 + col)) : s + __ldg(w.b4 + col);
    }
    __syncthreads();
    if (r == 0) {
      if (tid < classes) logits_out[((long long)b * max_steps + i) * classes + tid] = lg[tid];
      else if (tid >= 64 && tid < 64 + loc_dim) loc_out[((long long)b * max_steps + i) * loc_dim + (tid - 64)] = lg[classes + tid - 64];
    }
    if (tid == 0) {
      int best = 0;
      float bv = lg[0];
      for (int n = 1; n < classes; ++n) if (lg[n] > bv) { bv = lg[n]; best = n; }
      s_prev = best;
      if (r == 0) {
        ids_out[(long long)b * max_steps + i] = best;
        if (best == eos && !had_eos) {
          had_eos = true;
          atomicMax(sync_words + 1, i);
          __threadfence();
          atomicAdd(sync_words + 0, 1);
        }
        if (sw[0] == B && i >= sw[1]) {
#pragma unroll
          for (int peer = 0; peer < CL; ++peer) *cluster.map_shared_rank(&s_stop, peer) = 1;
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  cluster.sync();                            // nobody leaves while a peer may still store into its shared memory
  if (r == 0 && tid == 0) steps_run[b] = i;
}

// probabilities of the kept steps; everything at or beyond the common stop step reads as the reference's untouched zero rows
// (softmax of zeros = 1/classes, box = 0)
__global__ void sla_finish_kernel(const float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ loc, int B, int max_steps, int classes, int loc_dim,
                                  const int* __restrict__ sync_words, int* __restrict__ total_steps) {
  const int T = (sync_words[0] == B) ? sync_words[1] + 1 : max_steps;
  if (blockIdx.x == 0 && threadIdx.x == 0) *total_steps = T;
  const long long rows = (long long)B * max_steps;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(r % max_steps);
    float* p = probs + r * classes;
    if (i >= T) {
      for (int n = 0; n < classes; ++n) p[n] = 1.f / (float)classes;
      for (int n = 0; n < loc_dim; ++n) loc[r * loc_dim + n] = 0.f;
      continue;
    }
    const float* l = logits + r * classes;
    float m = l[0];
    for (int n = 1; n < classes; ++n) m = fmaxf(m, l[n]);
    float s = 0.f;
    for (int n = 0; n < classes; ++n) s += expf(l[n] - m);
    for (int n = 0; n < classes; ++n) p[n] = expf(l[n] - m) / s;
  }
}

}  // namespace sla
}  // namespace rdb

namespace {
thread_local std::string g_sla_err;
}

extern "C" {

const char* rdb_sla_last_error(void) { return g_sla_err.c_str(); }

int rdb_sla_decode(int device, const float* feat, const float* feat_proj, int batch, int hw, int c, const rdb_sla_weights_t* w, int classes, int loc_dim,
                   int max_steps, int eos, float* logits, float* probs, float* loc, int32_t* ids, int32_t* sync_words, int32_t* steps_run, int32_t* total_steps,
                   void* stream) {
  try {
    RDB_CHECK(feat && feat_proj && w && logits && probs && loc && ids && sync_words && steps_run && total_steps, "sla_decode: null argument");
    RDB_CHECK(batch > 0 && hw > 0 && c > 0 && 2 * c <= rdb::sla::THREADS && classes > 0 && classes <= rdb::sla::MAX_CLASSES && loc_dim > 0 &&
                  loc_dim <= rdb::sla::MAX_LOC && max_steps > 0 && eos >= 0 && eos < classes,
              "sla_decode: bad shape (hidden 256, C <= 128, classes <= 64, loc <= 8)");
    RDB_CHECK(w->hidden == rdb::sla::HID, "sla_decode: the kernel is built for hidden size 256");
    rdb::DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    rdb::sla::Weights k{w->Wh, w->bh, w->ws, w->WihT, w->WhhT, w->bih, w->bhh, w->W3, w->b3, w->W4, w->b4, w->W5, w->b5, w->W6, w->b6};
    const size_t smem = sizeof(float) * (5 * rdb::sla::HID + 3 * c + 32 + rdb::sla::MAX_CLASSES + rdb::sla::MAX_LOC + hw);
    RDB_CHECK(smem <= 200 * 1024, "sla_decode: feature map too large for the shared-memory score buffer");
    static bool attr_set[rdb::kMaxDevices] = {};
    if (smem > 48 * 1024 && !attr_set[device]) {
      RDB_CUDA(cudaFuncSetAttribute(rdb::sla::sla_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[device] = true;
    }
    RDB_CUDA(cudaMemsetAsync(sync_words, 0, 2 * sizeof(int32_t), st));
    // cluster of 4 CTAs per image (distributed shared memory) unless switched off; RDB_SLA=cta keeps one CTA per image
    const char* mode = rdb::sw_get("RDB_SLA");
    const bool use_cluster = hw <= 1024 && c <= 126 && !(mode && std::string(mode) == "cta");
    rdb::Ctx cx;
    cx.st = st;
    cx.begin(std::string(use_cluster ? "sla_decode_cluster4" : "sla_decode_cta") + "[B=" + std::to_string(batch) + ",HW=" + std::to_string(hw) + "]");
    if (use_cluster) {
      rdb::sla::sla_decode_cluster_kernel<4><<<batch * 4, rdb::sla::THREADS, 0, st>>>(feat, feat_proj, hw, c, k, classes, loc_dim, max_steps, eos, logits, loc, ids,
                                                                                      sync_words, steps_run);
    } else {
      rdb::sla::sla_decode_kernel<<<batch, rdb::sla::THREADS, smem, st>>>(feat, feat_proj, hw, c, k, classes, loc_dim, max_steps, eos, logits, loc, ids, sync_words,
                                                                          steps_run);
    }
    cx.end();
    const long long rows = (long long)batch * max_steps;
    rdb::sla::sla_finish_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(logits, probs, loc, batch, max_steps, classes, loc_dim, sync_words, total_steps);
    RDB_LAUNCH_CHECK();
    return RDB_OK;
  } catch (const std::exception& e) {
    g_sla_err = e.what();
    return g_sla_err.find("cuda") != std::string::npos ? RDB_ERR_CUDA : RDB_ERR_INVALID;
  }
}

}  // extern "C"
