// Tensor ops of the formula path (PP-FormulaNet_plus: PPHGNetV2-B6 encoder + MBart decoder) behind the
// C-ABI (include/rapiddoc_b200.h, "formula engine ops").  The host side (rapiddoc_b200/formula.py) walks the network and
// calls these on DEVICE buffers; every op is enqueued on the caller's stream and returns without synchronising.
//
// Layout: NHWC activations, a feature map is a [pixels, C] matrix with a row pitch `ld` — so the dense concatenation of an
// HGV2 block (rec_pphgnetv2.py:1124-1136: input + 6 layer outputs, 672..5632 channels) is never copied: every layer writes
// its channel slice of the block's wide buffer (ldc / c_off) and the next layer reads its input slice in place (lda).
//   conv kxk dense   im2col (one pass, K order = ky,kx,c = the packed weight order) + GEMM
//   conv 1x1 / linear  GEMM straight on the activations
//   GEMM             prec 0: fp32 SIMT (exact-parity mode)        prec 1: fp16 tcgen05 / TMEM / TMA (gemm_tc.cuh)
//   depthwise k x k, 2x2/s1 max-pool, LayerNorm, token embedding, single-query attention over a KV cache, row argmax
#include "../../include/rapiddoc_b200.h"

#include "engine.cuh"
#include "gemm_tf32.cuh"

namespace rdb {
namespace ops {

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld, int KH, int KW, int sh, int sw, int pt, int pl, int OH, int OW,
                              T* __restrict__ out) {
  // one thread per (output pixel, tap, 8-channel group) when C % 8 == 0, else per element
  const long long K = (long long)KH * KW * C;
  const long long total = (long long)n * OH * OW * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K;
    const int k = (int)(i % K);
    const int c = k % C, tap = k / C, kx = tap % KW, ky = tap / KW;
    const int ox = (int)(row % OW), oy = (int)((row / OW) % OH), b = (int)(row / ((long long)OW * OH));
    const int iy = oy * sh - pt + ky, ix = ox * sw - pl + kx;
    T v = from_f32<T>(0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((long long)(b * H + iy) * W + ix) * ld + c];
    out[i] = v;
  }
}

// the same gather with 8 channels per thread (C % 8 == 0, ld % 8 == 0, 16-byte aligned base): 128-bit loads / stores
template <typename T>
__global__ void im2col_vec8_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld, int KH, int KW, int sh, int sw, int pt, int pl, int OH, int OW,
                                   T* __restrict__ out) {
  const int C8 = C / 8;
  const long long K8 = (long long)KH * KW * C8;
  const long long total = (long long)n * OH * OW * K8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K8;
    const int k = (int)(i % K8);
    const int c8 = k % C8, tap = k / C8, kx = tap % KW, ky = tap / KW;
    const int ox = (int)(row % OW), oy = (int)((row / OW) % OH), b = (int)(row / ((long long)OW * OH));
    const int iy = oy * sh - pt + ky, ix = ox * sw - pl + kx;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) Vec8<T>::load(x + ((long long)(b * H + iy) * W + ix) * ld + c8 * 8, v);
    Vec8<T>::store(out + i * 8, v);
  }
}

// depthwise k x k, 8 channels per thread (C % 8 == 0, pitches % 8 == 0)
template <typename T>
__global__ void dwconv_vec8_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld_in, int K, int s, const float* __restrict__ w,
                                   const float* __restrict__ bias, int relu, T* __restrict__ out, int OH, int OW, int ld_out, int c_off) {
  const int C8 = C / 8, p = (K - 1) / 2;
  const long long total = (long long)n * OH * OW * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const long long px = i / C8;
    const int ox = (int)(px % OW), oy = (int)((px / OW) % OH), b = (int)(px / ((long long)OW * OH));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * s - p + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ox * s - p + kx;
        if (ix < 0 || ix >= W) continue;
        float v[8], wv[8];
        Vec8<T>::load(x + ((long long)(b * H + iy) * W + ix) * ld_in + c, v);
        Vec8<float>::load(w + (ky * K + kx) * C + c, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], wv[j], acc[j]);
      }
    }
    float bv[8];
    Vec8<float>::load(bias + c, bv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = rdb::apply_act_rt(acc[j] + bv[j], relu);
    Vec8<T>::store(out + px * ld_out + c_off + c, acc);
  }
}

// stride-1 depthwise k x k with a 4-pixel strip per thread: the k + 3 input vectors of a kernel row are loaded once and
// reused by the four outputs (k = 5: 40 vector loads per strip instead of 100)
template <typename T, int K>
__global__ void dwconv_strip4_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld_in, const float* __restrict__ w,
                                     const float* __restrict__ bias, int relu, T* __restrict__ out, int ld_out, int c_off) {
  constexpr int P = (K - 1) / 2;
  const int C8 = C / 8, W4 = (W + 3) / 4;
  const long long total = (long long)n * H * W4 * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    const long long t = i / C8;
    const int ox0 = (int)(t % W4) * 4, oy = (int)((t / W4) % H), b = (int)(t / ((long long)W4 * H));
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy - P + ky;
      if (iy < 0 || iy >= H) continue;
      const T* row = x + ((long long)(b * H + iy) * W) * ld_in + c;
      float v[K + 3][8];
#pragma unroll
      for (int q = 0; q < K + 3; ++q) {
        const int ix = ox0 - P + q;
        if (ix >= 0 && ix < W) Vec8<T>::load(row + (long long)ix * ld_in, v[q]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        float wv[8];
        Vec8<float>::load(w + (ky * K + kx) * C + c, wv);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(v[kx + p][j], wv[j], acc[p][j]);
      }
    }
    float bv[8];
    Vec8<float>::load(bias + c, bv);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (ox0 + p >= W) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = rdb::apply_act_rt(acc[p][j] + bv[j], relu);
      Vec8<T>::store(out + ((long long)(b * H + oy) * W + ox0 + p) * ld_out + c_off + c, acc[p]);
    }
  }
}

// depthwise k x k, pad (k-1)/2, stride s, weights [k][k][C] fp32, bias [C]; in pitch ld_in, out pitch ld_out (+ c_off)
template <typename T>
__global__ void dwconv_generic_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld_in, int K, int s, const float* __restrict__ w,
                                      const float* __restrict__ bias, int relu, T* __restrict__ out, int OH, int OW, int ld_out, int c_off) {
  const long long total = (long long)n * OH * OW * C;
  const int p = (K - 1) / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long px = i / C;
    const int ox = (int)(px % OW), oy = (int)((px / OW) % OH), b = (int)(px / ((long long)OW * OH));
    float acc = 0.f;
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * s - p + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ox * s - p + kx;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(to_f32<T>(x[((long long)(b * H + iy) * W + ix) * ld_in + c]), w[(ky * K + kx) * C + c], acc);
      }
    }
    acc += bias[c];
    acc = rdb::apply_act_rt(acc, relu);
    out[px * ld_out + c_off + c] = from_f32<T>(acc);
  }
}

// PaddingSameAsPaddleMaxPool2d(kernel 2, stride 1) (rec_pphgnetv2.py:962-976): zero pad one row / column at the bottom / right
template <typename T>
__global__ void maxpool2x2s1_kernel(const T* __restrict__ x, int n, int H, int W, int C, int ld_in, T* __restrict__ out, int ld_out, int c_off) {
  const long long total = (long long)n * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long px = i / C;
    const int ox = (int)(px % W), oy = (int)((px / W) % H), b = (int)(px / ((long long)W * H));
    float m = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = oy + dy, ix = ox + dx;
        const float v = (iy < H && ix < W) ? to_f32<T>(x[((long long)(b * H + iy) * W + ix) * ld_in + c]) : 0.f;
        m = fmaxf(m, v);
      }
    out[px * ld_out + c_off + c] = from_f32<T>(m);
  }
}

// rows of C (pitch ld_in) -> channel slice of a wider buffer, optional dtype change
template <typename TI, typename TO>
__global__ void copy_cols_kernel(const TI* __restrict__ x, long long rows, int C, int ld_in, TO* __restrict__ out, int ld_out, int c_off) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    out[r * ld_out + c_off + c] = from_f32<TO>(to_f32<TI>(x[r * ld_in + c]));
  }
}

// MBart decoder input (rec_unimernet_head.py:440-456, rec_ppformulanet_head.py:449-486): embed_tokens[id] * embed_scale +
// embed_positions[pos + 2]
__global__ void embed_kernel(const long long* __restrict__ ids, int B, int D, const float* __restrict__ tok, float scale, const float* __restrict__ pos_tab,
                             int pos, float* __restrict__ out, const int* __restrict__ step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i % D;
  if (step) { pos = *step; ids += (long long)pos * B; }        // device-side step counter: ids = row `step` of the token table
  out[i] = tok[ids[b] * D + d] * scale + pos_tab[(long long)(pos + 2) * D + d];
}

// one query per (batch row, head) against T cached keys / values: q [B, H*HD] (already scaled), k / v [B, Tcap, H*HD]
// softmax(q k^T) v in fp32 (MBartAttention.forward, rec_unimernet_head.py:541-633, tgt_len = 1, no mask needed)
__global__ void attn_decode_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int T, int Tcap, int H, int HD,
                                   float* __restrict__ out, const int* __restrict__ step) {
  extern __shared__ float sc[];          // scores [T]
  if (step) T = *step + 1;
  const int b = blockIdx.x / H, h = blockIdx.x % H, D = H * HD;
  const float* qp = q + (long long)b * D + h * HD;
  const float* kp = k + (long long)b * Tcap * D + h * HD;
  const float* vp = v + (long long)b * Tcap * D + h * HD;
  float mx = -INFINITY;
  if (HD == 32 && (D & 3) == 0) {
    // head_dim 32 (MBart d_model 512 / 16 heads): the query sits in registers and each key row is 8 x 128-bit loads (same
    // summation order as the scalar loop)
    float4 qv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) qv[j] = __ldg(reinterpret_cast<const float4*>(qp) + j);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const float4* kr = reinterpret_cast<const float4*>(kp + (long long)t * D);
      float4 kv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) kv[j] = __ldg(kr + j);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s = fmaf(qv[j].x, kv[j].x, s); s = fmaf(qv[j].y, kv[j].y, s); s = fmaf(qv[j].z, kv[j].z, s); s = fmaf(qv[j].w, kv[j].w, s);
      }
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
  } else {
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      float s = 0.f;
      for (int d = 0; d < HD; ++d) s = fmaf(qp[d], kp[(long long)t * D + d], s);
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
  }
  __shared__ float red[32];
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (blockDim.x + 31) / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) { const float e = expf(sc[t] - mx); sc[t] = e; sum += e; }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) sum += red[i];
  const float inv = 1.f / sum;
  // value pass: thread (d, part) sums every (blockDim / HD)-th key, the parts are folded through shared memory
  __shared__ float part[128];
  const int parts = blockDim.x / HD;
  if (parts >= 2 && HD <= 64) {
    const int d = threadIdx.x % HD, pi = threadIdx.x / HD;
    float acc = 0.f;
    if (pi < parts)
      for (int t = pi; t < T; t += parts) acc = fmaf(sc[t], vp[(long long)t * D + d], acc);
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < HD) {
      float tot = 0.f;
      for (int q = 0; q < parts; ++q) tot += part[q * HD + threadIdx.x];
      out[(long long)b * D + h * HD + threadIdx.x] = tot * inv;
    }
  } else {
    for (int d = threadIdx.x; d < HD; d += blockDim.x) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc = fmaf(sc[t], vp[(long long)t * D + d], acc);
      out[(long long)b * D + h * HD + d] = acc * inv;
    }
  }
}

// out[m, n] = act(A[m,:] . W[n,:] + bias[n]) (+ res[m,n]) for M <= 32 rows (one decode step of the whole batch): a GEMM this
// thin is a weight-streaming problem (HBM-bound, 4 B/MAC/32 rows), so each warp owns 2 output columns, reads their two weight
// rows exactly once with 128-bit loads, and the 32 activation rows of the current 128-wide K chunk sit in shared memory.
// out_step: optional device counter; the output rows are then written at out + *out_step * out_step_stride (KV-cache append).
template <int ACT>
__global__ void __launch_bounds__(128) skinny_gemm_kernel(const float* __restrict__ A, int lda, int M, int K, const float* __restrict__ W, int N,
                                                          const float* __restrict__ bias, const float* __restrict__ res, int ldr, float* __restrict__ out,
                                                          int ldc, const int* __restrict__ out_step, long long out_step_stride) {
  __shared__ float4 As[32][32];            // [row][k4] of the current chunk (128 k)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 4 + warp) * 2;
  float acc0[32], acc1[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
  const bool v0 = n0 < N, v1 = n0 + 1 < N;
  for (int k0 = 0; k0 < K; k0 += 128) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 32; i += 128) {
      const int m = i >> 5, k4 = i & 31;
      As[m][k4] = (m < M && k0 + k4 * 4 < K) ? *reinterpret_cast<const float4*>(A + (long long)m * lda + k0 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const bool kin = k0 + lane * 4 < K;         // K % 4 == 0: a float4 is either fully inside or fully outside
    const float4 w0 = (v0 && kin) ? *reinterpret_cast<const float4*>(W + (long long)n0 * K + k0 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 w1 = (v1 && kin) ? *reinterpret_cast<const float4*>(W + (long long)(n0 + 1) * K + k0 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < 32; ++m) {
      const float4 a = As[m][lane];
      acc0[m] = fmaf(a.x, w0.x, fmaf(a.y, w0.y, fmaf(a.z, w0.z, fmaf(a.w, w0.w, acc0[m]))));
      acc1[m] = fmaf(a.x, w1.x, fmaf(a.y, w1.y, fmaf(a.z, w1.z, fmaf(a.w, w1.w, acc1[m]))));
    }
  }
  // lane m ends up with the totals of row m: butterfly over the 32 rows
#pragma unroll
  for (int m = 0; m < 32; ++m) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      acc0[m] += __shfl_xor_sync(0xffffffffu, acc0[m], o);
      acc1[m] += __shfl_xor_sync(0xffffffffu, acc1[m], o);
    }
  }
  float r0 = 0.f, r1 = 0.f;
#pragma unroll
  for (int m = 0; m < 32; ++m) if (lane == m) { r0 = acc0[m]; r1 = acc1[m]; }
  if (lane < M) {
    float* o = out + (out_step ? (long long)(*out_step) * out_step_stride : 0) + (long long)lane * ldc;
    if (v0) { float y = apply_act<ACT>(r0 + (bias ? bias[n0] : 0.f)); if (res) y += res[(long long)lane * ldr + n0]; o[n0] = y; }
    if (v1) { float y = apply_act<ACT>(r1 + (bias ? bias[n0 + 1] : 0.f)); if (res) y += res[(long long)lane * ldr + n0 + 1]; o[n0 + 1] = y; }
  }
}

__global__ void add_kernel(const float* a, const float* b, float* out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// greedy step bookkeeping of PPFormulaNet_Head.generate_export (rec_ppformulanet_head.py:1118-1160): next = argmax (or eos when
// the forced-EOS length is reached), finished rows emit pad; a row finishes when it emits eos.  all_done: every row has an eos.
__global__ void greedy_step_kernel(const int* __restrict__ arg, int B, int force_eos, int eos, int pad, long long* __restrict__ next, int* __restrict__ unfinished,
                                   int* __restrict__ has_eos, int* __restrict__ all_done, int* __restrict__ step, int forced_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int st = step ? *step : 0;
  if (step) { next += (long long)(st + 1) * B; all_done += st + 1; force_eos = (st + 1) == forced_len - 1; }
  if (b < B) {
    long long t = force_eos ? eos : arg[b];
    t = unfinished[b] ? t : pad;
    next[b] = t;
    if (t == eos) { unfinished[b] = 0; has_eos[b] = 1; }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) {   // single block (B <= 1024, checked by the launcher)
    int d = 1;
    for (int i = 0; i < B; ++i) d &= has_eos[i];
    *all_done = d;
    if (step) *step = st + 1;
  }
}

// ---- elementwise / reshaping ops of the small ONNX CNNs (orientation classifier, seal detector), NHWC fp32 -------------------
// one launch applies a short chain of per-element steps: BatchNorm / bias / learnable-affine (scalar or per channel), ReLU,
// HardSigmoid(alpha, beta), x * HardSigmoid (HardSwish with the graph's own alpha), Sigmoid
struct ChainStep { int kind; float a, b; const float* va; const float* vb; };
struct Chain { int n; ChainStep s[8]; };
enum { CH_AFFINE = 0, CH_AFFINE_VEC = 1, CH_RELU = 2, CH_HSIG = 3, CH_HSWISH = 4, CH_SIGMOID = 5 };

__global__ void chain_kernel(const float* __restrict__ x, long long rows, int C, int ld_in, float* __restrict__ out, int ld_out, int c_off, Chain ch) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    float v = x[r * ld_in + c];
    for (int k = 0; k < ch.n; ++k) {
      const ChainStep& s = ch.s[k];
      switch (s.kind) {
        case CH_AFFINE: v = __fadd_rn(__fmul_rn(v, s.a), s.b); break;
        case CH_AFFINE_VEC: {
          if (s.va) v = __fmul_rn(v, s.va[c]);
          if (s.vb) v = __fadd_rn(v, s.vb[c]);
        } break;
        case CH_RELU: v = fmaxf(v, 0.f); break;
        case CH_HSIG: v = fminf(fmaxf(__fadd_rn(__fmul_rn(s.a, v), s.b), 0.f), 1.f); break;
        case CH_HSWISH: v = __fmul_rn(v, fminf(fmaxf(__fadd_rn(__fmul_rn(s.a, v), s.b), 0.f), 1.f)); break;
        case CH_SIGMOID: v = 1.f / (1.f + expf(-v)); break;
      }
    }
    out[r * ld_out + c_off + c] = v;
  }
}

// GlobalAveragePool: x [n, hw, C] (pitch ld) -> out [n, C]; block (32 channels x 8 row lanes)
__global__ void global_avgpool_kernel(const float* __restrict__ x, int hw, int C, int ld, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int b = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < hw; r += 8) acc += x[((long long)b * hw + r) * ld + c];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x];
    out[(long long)b * C + c] = s / (float)hw;
  }
}

// squeeze-excite scaling: out[n, p, c] = x[n, p, c] * gate[n, c]
__global__ void mul_gate_kernel(const float* __restrict__ x, const float* __restrict__ gate, long long rows, int rows_per_image, int C, int ld_in,
                                float* __restrict__ out, int ld_out) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    out[r * ld_out + c] = x[r * ld_in + c] * gate[(r / rows_per_image) * C + c];
  }
}

// Resize(mode nearest, coordinate_transformation_mode asymmetric, nearest_mode floor) into a channel slice:
// src = min(floor(dst / scale), in - 1) with scale = out / in in float32, as the ONNX definition (and onnxruntime) evaluate it
__global__ void resize_nearest_kernel(const float* __restrict__ x, int n, int H, int W, int C, int ld_in, int OH, int OW, float sy, float sx,
                                      float* __restrict__ out, int ld_out, int c_off) {
  const long long total = (long long)n * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long px = i / C;
    const int ox = (int)(px % OW), oy = (int)((px / OW) % OH), b = (int)(px / ((long long)OW * OH));
    const int iy = min((int)floorf(__fdiv_rn((float)oy, sy)), H - 1), ix = min((int)floorf(__fdiv_rn((float)ox, sx)), W - 1);
    out[px * ld_out + c_off + c] = x[((long long)(b * H + iy) * W + ix) * ld_in + c];
  }
}

// the same gather with 4 channels per thread and 32-bit index arithmetic (C, pitches, offsets multiples of 4; < 2^31 elements)
__global__ void resize_nearest_vec4_kernel(const float* __restrict__ x, int n, int H, int W, int C4, int ld_in, int OH, int OW, float sy, float sx,
                                           float* __restrict__ out, int ld_out, int c_off) {
  const unsigned total = (unsigned)n * OH * OW * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4, px = i / C4;
    const unsigned ox = px % OW, t = px / OW, oy = t % OH, b = t / OH;
    const int iy = min((int)floorf(__fdiv_rn((float)oy, sy)), H - 1), ix = min((int)floorf(__fdiv_rn((float)ox, sx)), W - 1);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((long long)(b * H + iy) * W + ix) * ld_in) + c4);
    *(reinterpret_cast<float4*>(out + (long long)px * ld_out + c_off) + c4) = v;
  }
}

// ConvTranspose(k = stride = s) after its GEMM: g [n*H*W, s*s*C] with columns (dy, dx, c) -> out [n, s*H, s*W, C]
__global__ void depth_to_space_kernel(const float* __restrict__ g, int n, int H, int W, int C, int s, float* __restrict__ out) {
  const int OH = H * s, OW = W * s;
  const long long total = (long long)n * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long px = i / C;
    const int ox = (int)(px % OW), oy = (int)((px / OW) % OH), b = (int)(px / ((long long)OW * OH));
    out[i] = g[((long long)(b * H + oy / s) * W + ox / s) * (s * s * C) + ((oy % s) * s + ox % s) * C + c];
  }
}

// Softmax over the last axis, one warp per row
__global__ void softmax_rows_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (r >= rows) return;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, x[r * C + c]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(x[r * C + c] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int c = lane; c < C; c += 32) out[r * C + c] = expf(x[r * C + c] - m) / s;
}

// uint8 HWC images (one padded canvas per image, valid region rh x rw at the top-left) -> normalised fp32 NHWC with the channels
// padded to 4: out = lut[c][v] inside the valid region, 0 outside.  The 3 x 256 table is built by the caller with the very numpy
// expression of the reference preprocessing, so the result is bit-identical to it by construction.
__global__ void lut_u8_nhwc4_kernel(const uint8_t* __restrict__ img, const int* __restrict__ valid, const float* __restrict__ lut, int n, int H, int W,
                                    float4* __restrict__ out) {
  __shared__ float t[3 * 256];
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) t[i] = lut[i];
  __syncthreads();
  const long long total = (long long)n * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < valid[2 * b] && x < valid[2 * b + 1]) {
      const uint8_t* p = img + i * 3;
      v = make_float4(t[p[0]], t[256 + p[1]], t[512 + p[2]], 0.f);
    }
    out[i] = v;
  }
}

// First layer of the ONNX CNNs: dense 3x3 conv on the 4-channel (3 + zero pad) fp32 NHWC input, 16 output channels.  im2col +
// GEMM moves the 9x-expanded input through HBM twice; here one thread owns one output pixel, reads its 9 input pixels as
// float4 and keeps the 16 x 36 weights in shared memory.  The products are accumulated in the same (ky, kx, c) order, with
// fmaf, as the GEMM path does, and the bias is added last: results are bit-identical to im2col + gemm_simt.
template <int CO>
__global__ void conv3x3_c4_kernel(const float4* __restrict__ x, int n, int H, int W, const float* __restrict__ w /*[CO][36]*/, const float* __restrict__ bias,
                                  int act, int stride, int pad, float* __restrict__ out, int OH, int OW, int ldc, int c_off) {
  __shared__ float sw[36][CO];
  for (int i = threadIdx.x; i < 36 * CO; i += blockDim.x) sw[i % 36][i / 36] = w[i];
  __syncthreads();
  const long long total = (long long)n * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), b = (int)(i / ((long long)OW * OH));
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - pad + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - pad + kx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + ((long long)(b * H + iy) * W + ix));
        const int t = (ky * 3 + kx) * 4;
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          acc[c] = fmaf(v.x, sw[t][c], acc[c]);
          acc[c] = fmaf(v.y, sw[t + 1][c], acc[c]);
          acc[c] = fmaf(v.z, sw[t + 2][c], acc[c]);
          acc[c] = fmaf(v.w, sw[t + 3][c], acc[c]);
        }
      }
    }
    float* o = out + i * ldc + c_off;
#pragma unroll
    for (int c = 0; c < CO; c += 4) {
      float4 r;
      r.x = rdb::apply_act_rt(acc[c] + (bias ? bias[c] : 0.f), act);
      r.y = rdb::apply_act_rt(acc[c + 1] + (bias ? bias[c + 1] : 0.f), act);
      r.z = rdb::apply_act_rt(acc[c + 2] + (bias ? bias[c + 2] : 0.f), act);
      r.w = rdb::apply_act_rt(acc[c + 3] + (bias ? bias[c + 3] : 0.f), act);
      *reinterpret_cast<float4*>(o + c) = r;
    }
  }
}

inline int grid_for(long long total) { long long g = (total + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

}  // namespace ops
}  // namespace rdb

namespace {
thread_local std::string g_ops_err;
template <typename F>
int op_guard(F&& f) {
  try { f(); return RDB_OK; }
  catch (const std::exception& e) { g_ops_err = e.what(); return g_ops_err.find("cuda") != std::string::npos ? RDB_ERR_CUDA : RDB_ERR_INVALID; }
}
rdb::Pool& ops_pool(int device) {       // gemm_tc needs no workspace; the pool only satisfies the launch context
  static rdb::Pool pools[rdb::kMaxDevices];
  return pools[device];
}
// per-kernel timing of the simple ops through the same profiler the engines use (rdb_profile_*)
struct OpTimer {
  rdb::Ctx cx;
  OpTimer(const std::string& name, cudaStream_t st) { cx.st = st; cx.begin(name); }
  ~OpTimer() { rdb::Profiler& p = rdb::Profiler::global(); if (p.on && !p.recs.empty()) cudaEventRecord(p.recs.back().b, cx.st); }
};
int sm_count(int device) {
  static int n[rdb::kMaxDevices] = {};
  if (!n[device]) RDB_CUDA(cudaDeviceGetAttribute(&n[device], cudaDevAttrMultiProcessorCount, device));
  return n[device];
}
}  // namespace

extern "C" {

const char* rdb_ops_last_error(void) { return g_ops_err.c_str(); }

int rdb_op_gemm(int device, int prec, const void* A, int lda, long long M, int K, const void* W, int N, const float* bias, int act, const void* res,
                int ldr, void* out, int ldc, int c_off, void* stream, const int32_t* out_step, long long out_step_stride) {
  return op_guard([&] {
    RDB_CHECK(A && W && out && M > 0 && K > 0 && N > 0, "gemm: bad argument");
    rdb::DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (prec == RDB_PREC_TF32 && M > 32 && res == nullptr && out_step == nullptr &&
        (act == rdb::ACT_NONE || act == rdb::ACT_RELU || act == rdb::ACT_HSWISH)) {
      RDB_CHECK(K % 4 == 0 && lda % 4 == 0, "gemm tf32: K and lda must be multiples of 4 (16-byte TMA rows)");
      OpTimer tm("gemm_tf32_op[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + "]", st);
      rdb::tf32::launch_gemm_tf32(device, static_cast<const float*>(A), lda, M, K, static_cast<const float*>(W), N, bias, act, static_cast<float*>(out), ldc,
                                  c_off, st);
      return;
    }
    if (prec == RDB_PREC_TF32) prec = RDB_PREC_FP32;
    if (prec == RDB_PREC_FP32) {
      RDB_CHECK(K % 4 == 0 && lda % 4 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "gemm fp32: K and lda must be multiples of 4, A and W 16-byte aligned (vector loads)");
      if (M <= 32 && (act == rdb::ACT_NONE || act == rdb::ACT_GELU || act == rdb::ACT_RELU)) {   // decode step: weight-streaming kernel
        const float* Af = static_cast<const float*>(A);
        const float* Wf = static_cast<const float*>(W);
        const float* Rf = static_cast<const float*>(res);
        float* Of = static_cast<float*>(out) + c_off;
        const int grid = (N + 7) / 8;
        OpTimer tm("skinny_gemm[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + "]", st);
        if (act == rdb::ACT_GELU) rdb::ops::skinny_gemm_kernel<rdb::ACT_GELU><<<grid, 128, 0, st>>>(Af, lda, (int)M, K, Wf, N, bias, Rf, ldr, Of, ldc, out_step, out_step_stride);
        else if (act == rdb::ACT_RELU) rdb::ops::skinny_gemm_kernel<rdb::ACT_RELU><<<grid, 128, 0, st>>>(Af, lda, (int)M, K, Wf, N, bias, Rf, ldr, Of, ldc, out_step, out_step_stride);
        else rdb::ops::skinny_gemm_kernel<rdb::ACT_NONE><<<grid, 128, 0, st>>>(Af, lda, (int)M, K, Wf, N, bias, Rf, ldr, Of, ldc, out_step, out_step_stride);
        RDB_LAUNCH_CHECK();
        return;
      }
      RDB_CHECK(out_step == nullptr, "gemm: out_step is only supported on the decode (M <= 32 rows) path");
      rdb::GemmArgs a{};
      a.A = A; a.lda = lda; a.W = static_cast<const float*>(W); a.bias = bias; a.res = res; a.ldr = ldr; a.out = out; a.ldc = ldc; a.c_off = c_off;
      a.M = (int)M; a.N = N; a.K = K; a.act = act;
      RDB_CHECK(M < (1ll << 31), "gemm: M too large");
      OpTimer tm("gemm_simt_op[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + "]", st);
      rdb::launch_gemm_simt<float, float>(a, st);
    } else {
      RDB_CHECK(out_step == nullptr, "gemm fp16: out_step not supported");
      RDB_CHECK(K % 8 == 0 && lda % 8 == 0 && c_off % 8 == 0 && ldc % 8 == 0, "gemm fp16: K, lda, ldc, c_off must be multiples of 8 (16-byte TMA rows)");
      rdb::Ctx cx;
      cx.st = st; cx.pool = &ops_pool(device); cx.precision = 1; cx.use_tc = true; cx.num_sms = sm_count(device);
      rdb::launch_gemm_tc(cx, static_cast<const __half*>(A), lda, M, K, static_cast<const __half*>(W), N, bias, act == rdb::ACT_GELU ? rdb::ACT_GELU : act,
                          static_cast<const __half*>(res), ldr, static_cast<__half*>(out), ldc, c_off);
    }
  });
}

int rdb_op_conv_tc(int device, const void* x, int n, int h, int w, int c, int ld, const void* wt, int cout, const float* bias, int act, int kh, int kw,
                   int sh, int sw, int pt, int pl, void* out, int oh, int ow, int ldc, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && wt && out && n > 0, "conv_tc: bad argument");
    RDB_CHECK(c % 8 == 0 && ld % 8 == 0 && ldc % 8 == 0 && c_off % 8 == 0, "conv_tc: channel counts / pitches must be multiples of 8");
    rdb::DeviceGuard g(device);
    rdb::Ctx cx;
    cx.st = (cudaStream_t)stream; cx.pool = &ops_pool(device); cx.precision = 1; cx.use_tc = true; cx.num_sms = sm_count(device);
    rdb::launch_conv_tc(cx, "conv_op", static_cast<const __half*>(x), n, h, w, c, static_cast<const __half*>(wt), cout, bias, act, kh, kw, sh, sw, pt, pl,
                        static_cast<__half*>(out), oh, ow, ldc, c_off, 0, 0, ld == c ? 0 : ld);
  });
}

int rdb_op_im2col(int device, int prec, const void* x, int n, int h, int w, int c, int ld, int kh, int kw, int sh, int sw, int pt, int pl, int oh, int ow,
                  void* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && n > 0, "im2col: bad argument");
    rdb::DeviceGuard g(device);
    const long long total = (long long)n * oh * ow * kh * kw * c;
    OpTimer tm("im2col[P=" + std::to_string((long long)n * oh * ow) + ",K=" + std::to_string(kh * kw * c) + "]", (cudaStream_t)stream);
    if (c % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x % 16) == 0) {
      if (prec == RDB_PREC_FP32)
        rdb::ops::im2col_vec8_kernel<float><<<rdb::ops::grid_for(total / 8), 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x), n, h, w, c, ld, kh, kw, sh, sw, pt, pl, oh, ow, static_cast<float*>(out));
      else
        rdb::ops::im2col_vec8_kernel<__half><<<rdb::ops::grid_for(total / 8), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x), n, h, w, c, ld, kh, kw, sh, sw, pt, pl, oh, ow, static_cast<__half*>(out));
      RDB_LAUNCH_CHECK();
      return;
    }
    if (prec == RDB_PREC_FP32)
      rdb::ops::im2col_kernel<float><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x), n, h, w, c, ld, kh, kw, sh, sw, pt, pl, oh, ow, static_cast<float*>(out));
    else
      rdb::ops::im2col_kernel<__half><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x), n, h, w, c, ld, kh, kw, sh, sw, pt, pl, oh, ow, static_cast<__half*>(out));
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_dwconv(int device, int prec, const void* x, int n, int h, int w, int c, int ld_in, int k, int stride, const float* wt, const float* bias, int relu,
                  void* out, int oh, int ow, int ld_out, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && wt && bias && n > 0 && (k & 1), "dwconv: bad argument");
    rdb::DeviceGuard g(device);
    const long long total = (long long)n * oh * ow * c;
    OpTimer tm("dwconv_op[P=" + std::to_string((long long)n * oh * ow) + ",C=" + std::to_string(c) + ",k=" + std::to_string(k) + "]", (cudaStream_t)stream);
    const bool vec = c % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && c_off % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0;
    if (vec && stride == 1 && (k == 5 || k == 3) && oh == h && ow == w) {
      const long long strips = (long long)n * h * ((w + 3) / 4) * (c / 8);
      const int grid = rdb::ops::grid_for(strips);
      cudaStream_t st = (cudaStream_t)stream;
      if (prec == RDB_PREC_FP32) {
        if (k == 5) rdb::ops::dwconv_strip4_kernel<float, 5><<<grid, 256, 0, st>>>(static_cast<const float*>(x), n, h, w, c, ld_in, wt, bias, relu, static_cast<float*>(out), ld_out, c_off);
        else rdb::ops::dwconv_strip4_kernel<float, 3><<<grid, 256, 0, st>>>(static_cast<const float*>(x), n, h, w, c, ld_in, wt, bias, relu, static_cast<float*>(out), ld_out, c_off);
      } else {
        if (k == 5) rdb::ops::dwconv_strip4_kernel<__half, 5><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), n, h, w, c, ld_in, wt, bias, relu, static_cast<__half*>(out), ld_out, c_off);
        else rdb::ops::dwconv_strip4_kernel<__half, 3><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), n, h, w, c, ld_in, wt, bias, relu, static_cast<__half*>(out), ld_out, c_off);
      }
      RDB_LAUNCH_CHECK();
      return;
    }
    if (vec) {
      if (prec == RDB_PREC_FP32)
        rdb::ops::dwconv_vec8_kernel<float><<<rdb::ops::grid_for(total / 8), 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x), n, h, w, c, ld_in, k, stride, wt, bias, relu, static_cast<float*>(out), oh, ow, ld_out, c_off);
      else
        rdb::ops::dwconv_vec8_kernel<__half><<<rdb::ops::grid_for(total / 8), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x), n, h, w, c, ld_in, k, stride, wt, bias, relu, static_cast<__half*>(out), oh, ow, ld_out, c_off);
      RDB_LAUNCH_CHECK();
      return;
    }
    if (prec == RDB_PREC_FP32)
      rdb::ops::dwconv_generic_kernel<float><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x), n, h, w, c, ld_in, k, stride, wt, bias, relu, static_cast<float*>(out), oh, ow, ld_out, c_off);
    else
      rdb::ops::dwconv_generic_kernel<__half><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x), n, h, w, c, ld_in, k, stride, wt, bias, relu, static_cast<__half*>(out), oh, ow, ld_out, c_off);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_maxpool2x2s1(int device, int prec, const void* x, int n, int h, int w, int c, int ld_in, void* out, int ld_out, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && n > 0, "maxpool: bad argument");
    rdb::DeviceGuard g(device);
    const long long total = (long long)n * h * w * c;
    if (prec == RDB_PREC_FP32)
      rdb::ops::maxpool2x2s1_kernel<float><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x), n, h, w, c, ld_in, static_cast<float*>(out), ld_out, c_off);
    else
      rdb::ops::maxpool2x2s1_kernel<__half><<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x), n, h, w, c, ld_in, static_cast<__half*>(out), ld_out, c_off);
    RDB_LAUNCH_CHECK();
  });
}

/* src_prec / dst_prec: 0 fp32, 1 fp16 */
int rdb_op_copy_cols(int device, int src_prec, int dst_prec, const void* x, long long rows, int c, int ld_in, void* out, int ld_out, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && rows > 0 && c > 0, "copy_cols: bad argument");
    rdb::DeviceGuard g(device);
    const int grid = rdb::ops::grid_for(rows * c);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_prec == 0 && dst_prec == 0) rdb::ops::copy_cols_kernel<float, float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), rows, c, ld_in, static_cast<float*>(out), ld_out, c_off);
    else if (src_prec == 0) rdb::ops::copy_cols_kernel<float, __half><<<grid, 256, 0, st>>>(static_cast<const float*>(x), rows, c, ld_in, static_cast<__half*>(out), ld_out, c_off);
    else if (dst_prec == 0) rdb::ops::copy_cols_kernel<__half, float><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), rows, c, ld_in, static_cast<float*>(out), ld_out, c_off);
    else rdb::ops::copy_cols_kernel<__half, __half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), rows, c, ld_in, static_cast<__half*>(out), ld_out, c_off);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_layernorm(int device, const float* x, long long rows, int c, const float* gamma, const float* beta, float eps, float* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && gamma && beta && rows > 0, "layernorm: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("layernorm_op", (cudaStream_t)stream);
    rdb::layernorm_kernel<float><<<rdb::cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, c, gamma, beta, eps, nullptr, out);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_embed(int device, const int64_t* ids, int batch, int dim, const float* tok, float scale, const float* pos_tab, int pos, float* out, void* stream,
                 const int32_t* step) {
  return op_guard([&] {
    RDB_CHECK(ids && tok && pos_tab && out && batch > 0, "embed: bad argument");
    rdb::DeviceGuard g(device);
    rdb::ops::embed_kernel<<<rdb::cdiv((long long)batch * dim, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(ids), batch, dim, tok, scale, pos_tab, pos, out, step);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_attn_decode(int device, const float* q, const float* k, const float* v, int batch, int t, int t_cap, int heads, int head_dim, float* out,
                       void* stream, const int32_t* step) {
  return op_guard([&] {
    RDB_CHECK(q && k && v && out && batch > 0 && t > 0 && t <= t_cap, "attn_decode: bad argument");
    if (step) t = t_cap;                     // score buffer sized for the whole cache when the length lives on the device
    RDB_CHECK((size_t)t * 4 <= 200 * 1024, "attn_decode: sequence too long for the score buffer");
    rdb::DeviceGuard g(device);
    auto kern = rdb::ops::attn_decode_kernel;
    const size_t sm = (size_t)t * sizeof(float);
    if (sm > 48 * 1024) RDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    OpTimer tm(step ? "attn_decode_self" : "attn_decode_cross", (cudaStream_t)stream);
    kern<<<batch * heads, 128, sm, (cudaStream_t)stream>>>(q, k, v, t, t_cap, heads, head_dim, out, step);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_chain(int device, const float* x, long long rows, int c, int ld_in, const int32_t* kinds, const float* a, const float* b,
                 const float* const* va, const float* const* vb, int n_steps, float* out, int ld_out, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && rows > 0 && c > 0 && n_steps >= 0 && n_steps <= 8 && (n_steps == 0 || (kinds && a && b)), "chain: bad argument (at most 8 steps)");
    rdb::DeviceGuard g(device);
    rdb::ops::Chain ch{};
    ch.n = n_steps;
    for (int i = 0; i < n_steps; ++i) {
      RDB_CHECK(kinds[i] >= rdb::ops::CH_AFFINE && kinds[i] <= rdb::ops::CH_SIGMOID, "chain: unknown step kind");
      ch.s[i] = rdb::ops::ChainStep{kinds[i], a[i], b[i], va ? va[i] : nullptr, vb ? vb[i] : nullptr};
    }
    OpTimer tm("chain_op[n=" + std::to_string(n_steps) + "]", (cudaStream_t)stream);
    rdb::ops::chain_kernel<<<rdb::ops::grid_for(rows * c), 256, 0, (cudaStream_t)stream>>>(x, rows, c, ld_in, out, ld_out, c_off, ch);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_global_avgpool(int device, const float* x, int n, int hw, int c, int ld, float* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && n > 0 && hw > 0 && c > 0, "global_avgpool: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("global_avgpool_op", (cudaStream_t)stream);
    rdb::ops::global_avgpool_kernel<<<dim3(n, (c + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, hw, c, ld, out);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_mul_gate(int device, const float* x, const float* gate, int n, int hw, int c, int ld_in, float* out, int ld_out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && gate && out && n > 0 && hw > 0 && c > 0, "mul_gate: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("mul_gate_op", (cudaStream_t)stream);
    rdb::ops::mul_gate_kernel<<<rdb::ops::grid_for((long long)n * hw * c), 256, 0, (cudaStream_t)stream>>>(x, gate, (long long)n * hw, hw, c, ld_in, out, ld_out);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_resize_nearest(int device, const float* x, int n, int h, int w, int c, int ld_in, int oh, int ow, float* out, int ld_out, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "resize_nearest: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("resize_nearest_op", (cudaStream_t)stream);
    const long long total = (long long)n * oh * ow * c;
    if (c % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && c_off % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && total < (1ll << 31))
      rdb::ops::resize_nearest_vec4_kernel<<<rdb::ops::grid_for(total / 4), 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c / 4, ld_in, oh, ow, (float)oh / (float)h,
                                                                                                           (float)ow / (float)w, out, ld_out, c_off);
    else
      rdb::ops::resize_nearest_kernel<<<rdb::ops::grid_for(total), 256, 0, (cudaStream_t)stream>>>(x, n, h, w, c, ld_in, oh, ow, (float)oh / (float)h,
                                                                                                  (float)ow / (float)w, out, ld_out, c_off);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_depth_to_space(int device, const float* g_in, int n, int h, int w, int c, int scale, float* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(g_in && out && n > 0 && h > 0 && w > 0 && c > 0 && scale >= 1, "depth_to_space: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("depth_to_space_op", (cudaStream_t)stream);
    rdb::ops::depth_to_space_kernel<<<rdb::ops::grid_for((long long)n * h * scale * w * scale * c), 256, 0, (cudaStream_t)stream>>>(g_in, n, h, w, c, scale, out);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_softmax_rows(int device, const float* x, long long rows, int c, float* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && out && rows > 0 && c > 0, "softmax_rows: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("softmax_rows_op", (cudaStream_t)stream);
    rdb::ops::softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, rows, c, out);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_lut_u8_nhwc4(int device, const uint8_t* img, const int32_t* valid_hw, const float* lut, int n, int h, int w, float* out, void* stream) {
  return op_guard([&] {
    RDB_CHECK(img && valid_hw && lut && out && n > 0 && h > 0 && w > 0 && ((uintptr_t)out % 16) == 0, "lut_u8_nhwc4: bad argument");
    rdb::DeviceGuard g(device);
    OpTimer tm("lut_u8_nhwc4_op", (cudaStream_t)stream);
    rdb::ops::lut_u8_nhwc4_kernel<<<rdb::ops::grid_for((long long)n * h * w), 256, 0, (cudaStream_t)stream>>>(img, valid_hw, lut, n, h, w, reinterpret_cast<float4*>(out));
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_conv3x3_c4(int device, const float* x, int n, int h, int w, const float* wt, int cout, const float* bias, int act, int stride, int pad, float* out,
                      int oh, int ow, int ldc, int c_off, void* stream) {
  return op_guard([&] {
    RDB_CHECK(x && wt && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "conv3x3_c4: bad argument");
    RDB_CHECK(cout == 16, "conv3x3_c4: built for 16 output channels (the stem of the PP-LCNet family)");
    RDB_CHECK(ldc % 4 == 0 && c_off % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0, "conv3x3_c4: 16-byte alignment");
    rdb::DeviceGuard g(device);
    OpTimer tm("conv3x3_c4_op[P=" + std::to_string((long long)n * oh * ow) + "]", (cudaStream_t)stream);
    rdb::ops::conv3x3_c4_kernel<16><<<rdb::ops::grid_for((long long)n * oh * ow), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), n, h, w, wt, bias, act, stride, pad, out, oh, ow, ldc, c_off);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_add(int device, const float* a, const float* b, float* out, long long n, void* stream) {
  return op_guard([&] {
    RDB_CHECK(a && b && out && n > 0, "add: bad argument");
    rdb::DeviceGuard g(device);
    rdb::ops::add_kernel<<<rdb::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
    RDB_LAUNCH_CHECK();
  });
}

int rdb_op_greedy_step(int device, const int32_t* argmax, int batch, int force_eos, int eos, int pad, int64_t* next, int32_t* unfinished, int32_t* has_eos,
                       int32_t* all_done, void* stream, int32_t* step, int forced_len) {
  return op_guard([&] {
    RDB_CHECK(argmax && next && unfinished && has_eos && all_done && batch > 0 && batch <= 1024, "greedy_step: bad argument");
    rdb::DeviceGuard g(device);
    rdb::ops::greedy_step_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(argmax, batch, force_eos, eos, pad, reinterpret_cast<long long*>(next), unfinished, has_eos, all_done, step, forced_len);
    RDB_LAUNCH_CHECK();
  });
}

}  // extern "C"
