// HBM-bound SIMT kernels of the PP-OCRv6 det/rec forward passes (NHWC activations,
// fp32 math, storage type T in {float, __half}).  Each kernel cites the reference op it
// implements (paths relative to rapid_doc/model/ocr/ppocrv6_pytorch/modeling/).
#pragma once
#include "common.cuh"

namespace rdb {

// =====================================================================================
// stem1: conv3x3 s2 p1, Cin=3 -> C1, +bias(BN folded) + ReLU.   rec_lcnetv4.py:151,160
// Input variants: fp32 NCHW (InferSession seam) or uint8 HWC + fused normalisation
// (facade seam; DetPreProcess / resize_norm_img arithmetic).
// =====================================================================================
struct InF32NCHW {
  static constexpr bool kU8 = false;
  const float* x; int H, W;
  __device__ __forceinline__ float get(int n, int y, int xx, int c) const {
    return x[(((long long)n * 3 + c) * H + y) * W + xx];
  }
  __device__ __forceinline__ float norm(int, int) const { return 0.f; }
  __device__ __forceinline__ float look(const float*, int n, int y, int xx, int c) const { return get(n, y, xx, c); }
  __device__ __forceinline__ bool row_inside(int, int ix0) const { return ix0 >= 0 && ix0 + 2 < W; }
  __device__ __forceinline__ void row9_fast(const float*, int n, int iy, int ix0, float (&v)[9]) const {
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] = get(n, iy, ix0 + j / 3, j % 3);
  }
  // the 3 pixels x 3 channels [ix0, ix0+3) of row iy as the 9 consecutive K entries (kx*3+ci) of one conv row; 0 outside
  __device__ __forceinline__ void row9(const float*, int n, int iy, int ix0, float (&v)[9]) const {
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int ix = ix0 + j / 3;
      v[j] = (ix >= 0 && ix < W) ? get(n, iy, ix, j % 3) : 0.f;
    }
  }
};
// norm_mode 0: (v*(1/255) - mean)/std   (DetPreProcess)
// norm_mode 1: (v/255 - 0.5)/0.5        (resize_norm_img); pixels with xx >= valid_w[n] are 0 (right pad)
struct InU8HWC {
  static constexpr bool kU8 = true;
  const uint8_t* x; int H, W; int norm_mode; float mean[3], stdv[3]; const int* valid_w;
  // the reference's float32 op order, evaluated once per (channel, byte value) into a 768-entry table
  __device__ __forceinline__ float norm(int c, int v8) const {
    const float v = (float)v8;
    if (norm_mode == 0) return __fdiv_rn(__fsub_rn(__fmul_rn(v, 1.0f / 255.0f), mean[c]), stdv[c]);
    return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 0.5f);
  }
  __device__ __forceinline__ float look(const float* lut, int n, int y, int xx, int c) const {
    if (norm_mode != 0 && valid_w != nullptr && xx >= valid_w[n]) return 0.f;
    return lut[c * 256 + x[(((long long)n * H + y) * W + xx) * 3 + c]];
  }
  // interior pixels: all three taps of the row are inside the (valid part of the) image -> 9 unpredicated byte loads
  __device__ __forceinline__ bool row_inside(int n, int ix0) const {
    const int wlim = (norm_mode != 0 && valid_w != nullptr) ? min(W, valid_w[n]) : W;
    return ix0 >= 0 && ix0 + 2 < wlim;
  }
  __device__ __forceinline__ void row9_fast(const float* lut, int n, int iy, int ix0, float (&v)[9]) const {
    const uint8_t* p = x + (((long long)n * H + iy) * W + ix0) * 3;
#pragma unroll
    for (int j = 0; j < 9; ++j) v[j] = lut[(j % 3) * 256 + __ldg(p + j)];
  }
  __device__ __forceinline__ void row9(const float* lut, int n, int iy, int ix0, float (&v)[9]) const {
    const uint8_t* p = x + (((long long)n * H + iy) * W + ix0) * 3;    // 9 consecutive bytes (ix0 may be -1: guarded below)
    const int wlim = (norm_mode != 0 && valid_w != nullptr) ? min(W, valid_w[n]) : W;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int ix = ix0 + j / 3;
      v[j] = (ix >= 0 && ix < wlim) ? lut[(j % 3) * 256 + __ldg(p + j)] : 0.f;
    }
  }
};

template <typename T, typename IN, int C1>
__global__ void __launch_bounds__(128) stem1_kernel(IN in, int N, const float* __restrict__ w /*[C1][3][3][3]*/,
                                                    const float* __restrict__ b, T* __restrict__ out, int OH, int OW, int out_wp) {
  __shared__ float sw[27 * C1];
  __shared__ float sb[C1];
  __shared__ float lut[IN::kU8 ? 768 : 1];
  for (int i = threadIdx.x; i < 27 * C1; i += blockDim.x) {
    int co = i / 27, r = i % 27;  // r = (ky*3+kx)*3+ci
    sw[r * C1 + co] = w[i];
  }
  for (int i = threadIdx.x; i < C1; i += blockDim.x) sb[i] = b[i];
  if (IN::kU8)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = in.norm(i >> 8, i & 255);
  __syncthreads();
  // out_wp >= OW: output rows are padded to out_wp pixels (pad pixels are written as zeros: they are the F.pad column
  // the wide-row stem2a conv reads)
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * OH * out_wp;
  if (idx >= total) return;
  int ox = idx % out_wp; int oy = (idx / out_wp) % OH; int n = idx / ((long long)out_wp * OH);
  float acc[C1];
  if (ox >= OW) {
    T* zp = out + idx * C1;
    float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C1; c += 8) Vec8<T>::store(zp + c, z);
    return;
  }
#pragma unroll
  for (int c = 0; c < C1; ++c) acc[c] = sb[c];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= in.H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= in.W) continue;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        float v = in.look(lut, n, iy, ix, ci);
        const float* wp = &sw[((ky * 3 + kx) * 3 + ci) * C1];
#pragma unroll
        for (int c = 0; c < C1; c += 4) {
          float4 w4 = *reinterpret_cast<const float4*>(wp + c);
          acc[c] = fmaf(v, w4.x, acc[c]); acc[c + 1] = fmaf(v, w4.y, acc[c + 1]);
          acc[c + 2] = fmaf(v, w4.z, acc[c + 2]); acc[c + 3] = fmaf(v, w4.w, acc[c + 3]);
        }
      }
    }
  }
  T* op = out + idx * C1;
#pragma unroll
  for (int c = 0; c < C1; c += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[c + j], 0.f);
    Vec8<T>::store(op + c, v);
  }
}

// =====================================================================================
// Generic small dense conv, NHWC -> NHWC, zero outside [0,H)x[0,W) (covers both the
// symmetric conv padding and the F.pad(0,1,0,1) of the stem, rec_lcnetv4.py:161-164).
// thread = one output pixel x CO_T output channels; weights staged in smem as
// [tap][ci][COUT] so a warp reads them as broadcast float4.
// =====================================================================================
template <typename T, int KH, int KW, int SH, int SW, int CIN, int COUT, int CO_T, int ACT>
__global__ void __launch_bounds__(128) conv_direct_kernel(const T* __restrict__ in, int N, int H, int W, int pt, int pl,
                                                          const float* __restrict__ w /*[COUT][KH][KW][CIN]*/,
                                                          const float* __restrict__ b, T* __restrict__ out, int OH, int OW,
                                                          int ld_out, int co_off) {
  extern __shared__ float smem[];
  float* sw = smem;                       // [KH*KW*CIN][COUT]
  float* sb = smem + KH * KW * CIN * COUT;
  constexpr int TAPS = KH * KW * CIN;
  for (int i = threadIdx.x; i < TAPS * COUT; i += blockDim.x) {
    int co = i / TAPS, r = i % TAPS;
    sw[r * COUT + co] = w[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = b[i];
  __syncthreads();
  constexpr int GROUPS = COUT / CO_T;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * OH * OW;
  if (idx >= total) return;
  const int g = blockIdx.y;  // output-channel group (warp-uniform)
  int ox = idx % OW; int oy = (idx / OW) % OH; int n = idx / ((long long)OW * OH);
  float acc[CO_T];
#pragma unroll
  for (int c = 0; c < CO_T; ++c) acc[c] = sb[g * CO_T + c];
#pragma unroll
  for (int ky = 0; ky < KH; ++ky) {
    int iy = oy * SH - pt + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      int ix = ox * SW - pl + kx;
      if (ix < 0 || ix >= W) continue;
      const T* ip = in + (((long long)n * H + iy) * W + ix) * CIN;
      const float* wt = sw + (ky * KW + kx) * CIN * COUT + g * CO_T;
#pragma unroll 1
      for (int c8 = 0; c8 < CIN; c8 += 8) {
        float v[8];
        if (CIN % 8 == 0) {
          Vec8<T>::load(ip + c8, v);
        } else {  // CIN % 4 == 0 (12-channel stem2a output)
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (c8 + j < CIN) ? to_f32<T>(ip[c8 + j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c8 + j < CIN) {
            const float* wp = wt + (c8 + j) * COUT;
#pragma unroll
            for (int c = 0; c < CO_T; c += 4) {
              float4 w4 = *reinterpret_cast<const float4*>(wp + c);
              acc[c] = fmaf(v[j], w4.x, acc[c]); acc[c + 1] = fmaf(v[j], w4.y, acc[c + 1]);
              acc[c + 2] = fmaf(v[j], w4.z, acc[c + 2]); acc[c + 3] = fmaf(v[j], w4.w, acc[c + 3]);
            }
          }
        }
      }
    }
  }
  T* op = out + idx * ld_out + co_off + g * CO_T;
#pragma unroll
  for (int c = 0; c < CO_T; c += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) op[c + j] = from_f32<T>(apply_act<ACT>(acc[c + j]));
  }
}

// =====================================================================================
// stem max-pool 2x2 s1 over the zero-padded (bottom/right) stem1 output, written into the
// first C channels of the concat buffer.  rec_lcnetv4.py:156,165-166
// =====================================================================================
template <typename T>
__global__ void pool2x2_concat_kernel(const T* __restrict__ in, int N, int H, int W, int C, T* __restrict__ out, int ld_out, int in_wp) {
  int G = C / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W * G;
  if (idx >= total) return;
  int g = idx % G; long long p = idx / G;
  int x = p % W; int y = (p / W) % H; int n = p / ((long long)W * H);
  float m[8];
  Vec8<T>::load(in + (((long long)n * H + y) * in_wp + x) * C + g * 8, m);
#pragma unroll
  for (int d = 1; d < 4; ++d) {
    int yy = y + (d >> 1), xx = x + (d & 1);
    float v[8];
    if (yy < H && xx < W) {
      Vec8<T>::load(in + (((long long)n * H + yy) * in_wp + xx) * C + g * 8, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;  // F.pad zeros take part in the max
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
  }
  Vec8<T>::store(out + p * ld_out + g * 8, m);
}

// =====================================================================================
// Depthwise conv KHxKW, stride (sh,sw), pad (KH/2,KW/2), +bias, optional act, optional
// "+ centre input" (LightSVTR h + dw(h), necks/rnn.py:367).  thread = pixel x 8 channels.
// rec_lcnetv4.py:187-206 (token_conv), db_fpn.py:317-325 (7x7), rnn.py:352 (1x7)
// =====================================================================================
template <typename T, int KH, int KW, int ACT, bool ADD_IN>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, int N, int H, int W, int C, int sh, int sw,
                                                     const float* __restrict__ w /*[KH][KW][C]*/, const float* __restrict__ b,
                                                     T* __restrict__ out, int OH, int OW) {
  int G = C / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * OH * OW * G;
  if (idx >= total) return;
  int g = idx % G; long long p = idx / G;
  int ox = p % OW; int oy = (p / OW) % OH; int n = p / ((long long)OW * OH);
  float acc[8];
  {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(b + g * 8));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(b + g * 8 + 4));
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
#pragma unroll
  for (int ky = 0; ky < KH; ++ky) {
    int iy = oy * sh - KH / 2 + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      int ix = ox * sw - KW / 2 + kx;
      if (ix < 0 || ix >= W) continue;
      float v[8];
      Vec8<T>::load(in + (((long long)n * H + iy) * W + ix) * C + g * 8, v);
      const float* wp = w + (ky * KW + kx) * C + g * 8;
      float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]);
      acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
      acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]);
      acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = apply_act<ACT>(acc[j]);
  if (ADD_IN) {
    float v[8];
    Vec8<T>::load(in + (((long long)n * H + oy) * W + ox) * C + g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  Vec8<T>::store(out + p * C + g * 8, acc);
}

// Shared-memory tiled depthwise conv, stride 1, for the large kernels (RepLKFPN 7x7,
// db_fpn.py:317-325): a block owns TH x TW output pixels x 8*G channels, stages the halo tile
// once in smem (pixel pitch 16*G+16 bytes => the 4-pixel strips of a quarter-warp land on
// distinct banks) and every thread produces a strip of 4 consecutive outputs x 8 channels so
// each staged input vector is reused K times in registers.
template <typename T, int K, int G, int TH, int TW>
__global__ void __launch_bounds__(G * (TW / 4) * TH)
dwconv_tiled_kernel(const T* __restrict__ in, int N, int H, int W, int C, const float* __restrict__ w /*[K][K][C]*/,
                    const float* __restrict__ b, T* __restrict__ out, float* __restrict__ pool = nullptr) {
  constexpr int HH = TH + K - 1, HW = TW + K - 1;
  constexpr int PITCH = 16 * G + 16;                 // bytes per staged pixel
  constexpr int XS = TW / 4;
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* tile = dsm;                                // [HH][HW][PITCH]
  float* sw = reinterpret_cast<float*>(dsm + HH * HW * PITCH);  // [K*K][8*G]
  const int cblocks = C / (8 * G);
  const int cb = blockIdx.z % cblocks, n = blockIdx.z / cblocks;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int c0 = cb * 8 * G;
  const int tid = threadIdx.x;
  for (int i = tid; i < K * K * 8 * G; i += blockDim.x) sw[i] = w[(i / (8 * G)) * C + c0 + i % (8 * G)];
  if (sizeof(T) == 2) {
    // halo staging with every load of a thread in flight before the first shared-memory store (ncu: the per-iteration
    // load -> store dependency left the kernel waiting on long_scoreboard 4 cycles per issued instruction)
    constexpr int NT = G * (TW / 4) * TH, NLD = (HH * HW * G + NT - 1) / NT;
    uint4 v[NLD];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * NT;
      const int g = i % G, p = i / G;
      const int hx = p % HW, hy = p / HW;
      const int iy = y0 + hy - K / 2, ix = x0 + hx - K / 2;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < HH * HW * G && iy >= 0 && iy < H && ix >= 0 && ix < W)
        v[u] = __ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + iy) * W + ix) * C + c0 + g * 8));
    }
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * NT;
      if (i < HH * HW * G) *reinterpret_cast<uint4*>(tile + (i / G) * PITCH + (i % G) * 16) = v[u];
    }
  } else {
    for (int i = tid; i < HH * HW * G; i += blockDim.x) {
      const int g = i % G, p = i / G;
      const int hx = p % HW, hy = p / HW;
      const int iy = y0 + hy - K / 2, ix = x0 + hx - K / 2;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        float f[8];
        Vec8<T>::load(in + (((long long)n * H + iy) * W + ix) * C + c0 + g * 8, f);
        __half2* hp = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) hp[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
      }
      *reinterpret_cast<uint4*>(tile + (hy * HW + hx) * PITCH + g * 16) = v;
    }
  }
  __syncthreads();
  const int g = tid % G, xs = (tid / G) % XS, ty = tid / (G * XS);
  float acc[4][8];
  {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0 + g * 8));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + g * 8 + 4));
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    uint4 iv[K + 3];
#pragma unroll
    for (int i = 0; i < K + 3; ++i) iv[i] = *reinterpret_cast<const uint4*>(tile + ((ty + ky) * HW + xs * 4 + i) * PITCH + g * 16);
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(sw + (ky * K + kx) * 8 * G + g * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + (ky * K + kx) * 8 * G + g * 8 + 4);
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const __half2* hp = reinterpret_cast<const __half2*>(&iv[o + kx]);
        float2 f0 = __half22float2(hp[0]), f1 = __half22float2(hp[1]), f2 = __half22float2(hp[2]), f3 = __half22float2(hp[3]);
        acc[o][0] = fmaf(f0.x, w0.x, acc[o][0]); acc[o][1] = fmaf(f0.y, w0.y, acc[o][1]);
        acc[o][2] = fmaf(f1.x, w0.z, acc[o][2]); acc[o][3] = fmaf(f1.y, w0.w, acc[o][3]);
        acc[o][4] = fmaf(f2.x, w1.x, acc[o][4]); acc[o][5] = fmaf(f2.y, w1.y, acc[o][5]);
        acc[o][6] = fmaf(f3.x, w1.z, acc[o][6]); acc[o][7] = fmaf(f3.y, w1.w, acc[o][7]);
      }
    }
  }
  const int oy = y0 + ty;
  if (oy < H) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int ox = x0 + xs * 4 + o;
      if (ox < W) Vec8<T>::store(out + (((long long)n * H + oy) * W + ox) * C + c0 + g * 8, acc[o]);
    }
  }
  if (pool != nullptr) {
    // squeeze-excitation pool fused into the producer: this block's channel sums (valid pixels only, fp32, fixed order) go to
    // pool[n][tile][C]; se_fc_kernel adds the tiles up — the separate pooling pass over the whole tensor disappears
    __syncthreads();                                   // every thread is done reading the staged tile: reuse it
    float* red = reinterpret_cast<float*>(dsm);        // [XS*TH][8*G]
    float sacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sacc[j] = 0.f;
    if (oy < H) {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (x0 + xs * 4 + o < W) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sacc[j] += acc[o][j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(xs + XS * ty) * 8 * G + g * 8 + j] = sacc[j];
    __syncthreads();
    if (tid < 8 * G) {
      float t = 0.f;
      for (int k = 0; k < XS * TH; ++k) t += red[k * 8 * G + tid];
      const int tiles = gridDim.x * gridDim.y, tile_id = blockIdx.y * gridDim.x + blockIdx.x;
      pool[((long long)n * tiles + tile_id) * C + c0 + tid] = t;
    }
  }
}

// Packed-half variant of the tiled depthwise conv for the 7x7 RepLK kernels (fp16 storage mode only).  The fp32 kernel
// above is instruction-issue bound (49 FFMA + ~18 half->float converts per output element, ncu: 61% issue slots, 9% DRAM);
// here the K horizontal taps of one kernel row run as HFMA2 on channel PAIRS (inputs already sit in smem as half2, weights
// staged as half2) and every row's K-term partial sum is flushed into the fp32 accumulators, so the fp16 accumulation depth
// is K = 7, never K*K: per 4x8 output strip and row 112 HFMA2 + 64 convert/add instead of 224 FFMA + 80 converts.
template <int K, int G, int TH, int TW, int ROWS>
__global__ void __launch_bounds__(G * (TW / 4) * TH)
dwconv_tiled_h2_kernel(const __half* __restrict__ in, int N, int H, int W, int C, const __half* __restrict__ w /*[K][K][C]*/,
                       const float* __restrict__ b, __half* __restrict__ out) {
  constexpr int HH = TH + K - 1, HW = TW + K - 1;
  constexpr int PITCH = 16 * G + 16;                 // bytes per staged pixel
  constexpr int XS = TW / 4;
  extern __shared__ __align__(16) uint8_t dsm[];
  uint8_t* tile = dsm;                                // [HH][HW][PITCH]
  uint4* sw = reinterpret_cast<uint4*>(dsm + HH * HW * PITCH);  // [K*K][G] x 8 halves
  const int cblocks = C / (8 * G);
  const int cb = blockIdx.z % cblocks, n = blockIdx.z / cblocks;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int c0 = cb * 8 * G;
  const int tid = threadIdx.x;
  for (int i = tid; i < K * K * G; i += blockDim.x) sw[i] = *reinterpret_cast<const uint4*>(w + (i / G) * C + c0 + (i % G) * 8);
  for (int i = tid; i < HH * HW * G; i += blockDim.x) {
    const int g = i % G, p = i / G;
    const int hx = p % HW, hy = p / HW;
    const int iy = y0 + hy - K / 2, ix = x0 + hx - K / 2;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(in + (((long long)n * H + iy) * W + ix) * C + c0 + g * 8);
    *reinterpret_cast<uint4*>(tile + (hy * HW + hx) * PITCH + g * 16) = v;
  }
  __syncthreads();
  const int g = tid % G, xs = (tid / G) % XS, ty = tid / (G * XS);
  float acc[4][8];
  {
    float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0 + g * 8));
    float4 b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + g * 8 + 4));
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
  // ROWS kernel rows share one packed-half partial sum before it is flushed into the fp32 accumulators
  // (fp16 accumulation depth = ROWS*K taps)
#pragma unroll 1
  for (int ky0 = 0; ky0 < K; ky0 += ROWS) {
    __half2 r[4][4];
#pragma unroll
    for (int kr = 0; kr < ROWS; ++kr) {
      const int ky = ky0 + kr;
      if (ky < K) {
        uint4 iv[K + 3];
#pragma unroll
        for (int i = 0; i < K + 3; ++i) iv[i] = *reinterpret_cast<const uint4*>(tile + ((ty + ky) * HW + xs * 4 + i) * PITCH + g * 16);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const uint4 wv = sw[(ky * K + kx) * G + g];
          const __half2* wp = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const __half2* hp = reinterpret_cast<const __half2*>(&iv[o + kx]);
#pragma unroll
            for (int q = 0; q < 4; ++q) r[o][q] = (kr == 0 && kx == 0) ? __hmul2(hp[q], wp[q]) : __hfma2(hp[q], wp[q], r[o][q]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(r[o][q]);
        acc[o][2 * q] += f.x; acc[o][2 * q + 1] += f.y;
      }
  }
  const int oy = y0 + ty;
  if (oy >= H) return;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int ox = x0 + xs * 4 + o;
    if (ox < W) Vec8<__half>::store(out + (((long long)n * H + oy) * W + ox) * C + c0 + g * 8, acc[o]);
  }
}

// =====================================================================================
// Squeeze-excitation: deterministic two-stage global average pool, the two tiny FCs and
// the gate.  rec_lcnetv4.py:120-142 (gate = clip(x/6+.5,0,1)), db_fpn.py:288-308
// (gate = clip(.2x+.5,0,1), used as x + x*gate -> we emit 1+gate).
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pool_partial_kernel(const T* __restrict__ in, int HW, int C, float* __restrict__ partial,
                                                           int chunks) {
  extern __shared__ float red[];  // [P][C]
  int G = C / 8;
  int P = blockDim.x / G;
  int n = blockIdx.y, chunk = blockIdx.x;
  int g = threadIdx.x % G, p = threadIdx.x / G;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (p < P) {
    const T* base = in + (long long)n * HW * C + g * 8;
    for (long long px = (long long)chunk * P + p; px < HW; px += (long long)chunks * P) {
      float v[8];
      Vec8<T>::load(base + px * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[p * C + g * 8 + j] = s[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < P; ++q) t += red[q * C + c];
    partial[((long long)n * chunks + chunk) * C + c] = t;
  }
}

// mode 0: gate = hardsigmoid(z) = clip(z/6+.5,0,1);  mode 1: gate = 1 + clip(.2z+.5,0,1)
// w0 != null: the pooled tensor is the INPUT of a bias-free 1x1 conv w0 [C][C0] and the SE acts on its output;
// mean(conv(x)) = w0 * mean(x) (linearity), so the conv output never has to be pooled (RepLKFPN insert_conv).
// Latency-bound by construction (a few hundred MACs): 32 warps, every dot product / partial sum is one warp with
// the loads of all lanes in flight at once and a fixed-order shuffle tree (deterministic).
__device__ __forceinline__ float warp_sum(float t) {
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffff, t, o);
  return t;
}

// 512 threads; every stage spreads its loads over the whole block so that a stage costs about one memory latency:
//   pool   : thread (channel c, part j) adds the chunks j, j+P, ... ; the P parts are then added in a fixed order
//   matvec : 16 threads per output row, each with K/64 independent float4 loads, xor-shuffle tree inside the 16-lane group
// (the first version used one warp per row and C/32 dependent rounds: 14.5 us per launch, 3 % of a det step over 26 launches).
// All sums run in a fixed order: deterministic.
__device__ __forceinline__ void se_matvec16(const float* __restrict__ Wm, const float* vec, int rows, int K, float* out_sm) {
  // out_sm[r] = dot(Wm[r][0..K), vec), K % 4 == 0 handled by float4, any K by the scalar tail
  const int t16 = threadIdx.x & 15, grp = threadIdx.x >> 4, ngrp = blockDim.x >> 4;
  for (int r0 = 0; r0 < rows; r0 += ngrp) {
    const int r = r0 + grp;
    float t = 0.f;
    if (r < rows) {
      const float* wr = Wm + (long long)r * K;
      if ((K & 3) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(wr);
        for (int c = t16; c < K / 4; c += 16) { const float4 w = __ldg(w4 + c); t = fmaf(w.x, vec[4 * c], fmaf(w.y, vec[4 * c + 1], fmaf(w.z, vec[4 * c + 2], fmaf(w.w, vec[4 * c + 3], t)))); }
      } else {
        for (int c = t16; c < K; c += 16) t = fmaf(__ldg(wr + c), vec[c], t);
      }
    }
#pragma unroll
    for (int o = 8; o; o >>= 1) t += __shfl_xor_sync(0xffffffff, t, o);
    if (r < rows && t16 == 0) out_sm[r] = t;
  }
}

static __global__ void __launch_bounds__(512) se_fc_kernel(const float* __restrict__ partial, int chunks, int HW, int C, int Cr,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2, int mode,
                                                         float* __restrict__ gate, const float* __restrict__ w0, int C0) {
  extern __shared__ float sm[];  // mean[C] + hid[Cr] + mean0[C0] + red[512] + tmp[C]
  float* mean = sm;
  float* hid = sm + C;
  float* mean0 = hid + Cr;
  const int n = blockIdx.x;
  const int Cp = (w0 != nullptr) ? C0 : C;         // channels of the pooled tensor
  float* pooled = (w0 != nullptr) ? mean0 : mean;
  float* red = mean0 + ((w0 != nullptr) ? C0 : 0);
  float* tmp = red + 512;
  {
    const int P = blockDim.x / Cp > 0 ? blockDim.x / Cp : 1;   // parts per channel
    const int c = threadIdx.x % Cp, j = threadIdx.x / Cp;
    if (j < P) {
      const float* p = partial + (long long)n * chunks * Cp + c;
      float t = 0.f;
      for (int k = j; k < chunks; k += P) t += p[(long long)k * Cp];
      red[j * Cp + c] = t;
    }
    __syncthreads();
    for (int cc = threadIdx.x; cc < Cp; cc += blockDim.x) {
      float t = 0.f;
      for (int jj = 0; jj < P; ++jj) t += red[jj * Cp + cc];
      pooled[cc] = t / (float)HW;
    }
    __syncthreads();
  }
  if (w0 != nullptr) {                                           // mean of the 1x1 conv output = W0 * mean of its input
    se_matvec16(w0, mean0, C, C0, mean);
    __syncthreads();
  }
  se_matvec16(w1, mean, Cr, C, tmp);                             // hidden = relu(W1 mean + b1)
  __syncthreads();
  for (int r = threadIdx.x; r < Cr; r += blockDim.x) hid[r] = fmaxf(tmp[r] + b1[r], 0.f);
  __syncthreads();
  se_matvec16(w2, hid, C, Cr, tmp);                              // gate = f(W2 hidden + b2)
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float t = tmp[c] + b2[c];
    float gt;
    if (mode == 0) gt = fminf(fmaxf(t / 6.f + 0.5f, 0.f), 1.f);
    else gt = 1.f + fminf(fmaxf(0.2f * t + 0.5f, 0.f), 1.f);
    gate[(long long)n * C + c] = gt;
  }
}

template <typename T>
__global__ void scale_channels_kernel(T* __restrict__ x, long long HW, int C, const float* __restrict__ gate, long long total8) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total8) return;
  int G = C / 8;
  int g = idx % G; long long p = idx / G;
  int n = p / HW;
  float v[8];
  Vec8<T>::load(x + p * C + g * 8, v);
  const float* gp = gate + (long long)n * C + g * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= gp[j];
  Vec8<T>::store(x + p * C + g * 8, v);
}

// =====================================================================================
// RepLKFPN top-down: dst += nearest_up2(src).   db_fpn.py:394-399
// =====================================================================================
template <typename T>
__global__ void upsample2_add_kernel(T* __restrict__ dst, const T* __restrict__ src, int N, int H, int W, int C) {
  int G = C / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W * G;
  if (idx >= total) return;
  int g = idx % G; long long p = idx / G;
  int x = p % W; int y = (p / W) % H; int n = p / ((long long)W * H);
  float a[8], b[8];
  Vec8<T>::load(dst + p * C + g * 8, a);
  Vec8<T>::load(src + (((long long)n * (H / 2) + (y >> 1)) * (W / 2) + (x >> 1)) * C + g * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  Vec8<T>::store(dst + p * C + g * 8, a);
}

// RepLKFPN output: nearest upsample x1/2/4/8 of the four 24-channel maps (each scaled by
// its SE gate 1+g), concatenated coarse->fine.   db_fpn.py:401-415
template <typename T>
struct NeckSrc { const T* f[4]; const float* gate[4]; };

template <typename T>
__global__ void neck_concat_kernel(NeckSrc<T> s, int N, int H, int W, T* __restrict__ out, int out_wp, int x_off) {
  // out [N,H,W,96]; channel group j (24 ch) comes from level 3-j (level L has size H>>L)
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W * 12;
  if (idx >= total) return;
  int g = idx % 12; long long p = idx / 12;
  int x = p % W; int y = (p / W) % H; int n = p / ((long long)W * H);
  int j = g / 3, c8 = (g % 3) * 8;
  int L = 3 - j;
  int h = H >> L, w = W >> L;
  float v[8];
  Vec8<T>::load(s.f[L] + (((long long)n * h + (y >> L)) * w + (x >> L)) * 24 + c8, v);
  const float* gp = s.gate[L] + n * 24 + c8;
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] *= gp[k];
  Vec8<T>::store(out + (((long long)n * H + y) * out_wp + x + x_off) * 96 + g * 8, v);
}

// zero the pad pixel columns [c0, c0+nc) of a row-padded NHWC buffer ([rows = n*H][wp][C])
template <typename T>
__global__ void zero_cols_kernel(T* __restrict__ buf, long long rows, int wp, int C, int c0, int nc) {
  int G = C / 8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = rows * nc * G;
  if (idx >= total) return;
  int g = idx % G; long long q = idx / G;
  int k = q % nc; long long r = q / nc;
  float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  Vec8<T>::store(buf + (r * wp + c0 + k) * C + g * 8, z);
}

// =====================================================================================
// DBHead tail: ConvT2x2s2(24->24)+BN+ReLU -> ConvT2x2s2(24->1)+bias -> sigmoid ->
// nan_to_num, fused; also emits the pre-dilation segmentation byte (prob > thresh).
// det_db_head.py:117-147; threshold: ocr_patch.py:229.
// thread = one pixel of the intermediate 2H x 2W grid -> 2x2 output probabilities.
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(128) head_tail_kernel(const T* __restrict__ in /*[N,H,W,24]*/, int N, int H, int W,
                                                        const float* __restrict__ w_up /*[2][2][24][24]*/,
                                                        const float* __restrict__ b_up, const float* __restrict__ w_fin /*[2][2][24]*/,
                                                        const float* __restrict__ b_fin, float thresh, float* __restrict__ prob,
                                                        uint8_t* __restrict__ seg) {
  __shared__ float s_up[4 * 24 * 24];
  __shared__ float s_bu[24];
  __shared__ float s_fin[4 * 24];
  for (int i = threadIdx.x; i < 4 * 24 * 24; i += blockDim.x) {
    // store as [q][ci][co] for broadcast float4 reads over co
    int q = i / 576, r = i % 576, co = r / 24, ci = r % 24;
    s_up[q * 576 + ci * 24 + co] = w_up[i];
  }
  for (int i = threadIdx.x; i < 24; i += blockDim.x) s_bu[i] = b_up[i];
  for (int i = threadIdx.x; i < 96; i += blockDim.x) s_fin[i] = w_fin[i];
  __syncthreads();
  const int H2 = 2 * H, W2 = 2 * W;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H2 * W2;
  if (idx >= total) return;
  int x2 = idx % W2; int y2 = (idx / W2) % H2; int n = idx / ((long long)W2 * H2);
  int q = (y2 & 1) * 2 + (x2 & 1);
  const T* ip = in + (((long long)n * H + (y2 >> 1)) * W + (x2 >> 1)) * 24;
  float hid[24];
#pragma unroll
  for (int c = 0; c < 24; ++c) hid[c] = s_bu[c];
#pragma unroll
  for (int c8 = 0; c8 < 24; c8 += 8) {
    float v[8];
    Vec8<T>::load(ip + c8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wp = &s_up[q * 576 + (c8 + j) * 24];
#pragma unroll
      for (int c = 0; c < 24; c += 4) {
        float4 w4 = *reinterpret_cast<const float4*>(wp + c);
        hid[c] = fmaf(v[j], w4.x, hid[c]); hid[c + 1] = fmaf(v[j], w4.y, hid[c + 1]);
        hid[c + 2] = fmaf(v[j], w4.z, hid[c + 2]); hid[c + 3] = fmaf(v[j], w4.w, hid[c + 3]);
      }
    }
  }
  const float bf = b_fin[0];
  float o[4] = {bf, bf, bf, bf};
#pragma unroll
  for (int c = 0; c < 24; ++c) {
    float hv = fmaxf(hid[c], 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = fmaf(hv, s_fin[k * 24 + c], o[k]);
  }
  const int W4 = 2 * W2;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float pz = 1.f / (1.f + expf(-o[k]));
    if (pz != pz) pz = 0.f;  // nan_to_num
    o[k] = pz;
  }
  long long base = ((long long)n * (2 * H2) + 2 * y2) * W4 + 2 * x2;
  *reinterpret_cast<float2*>(prob + base) = make_float2(o[0], o[1]);
  *reinterpret_cast<float2*>(prob + base + W4) = make_float2(o[2], o[3]);
  if (seg != nullptr) {
    *reinterpret_cast<uchar2*>(seg + base) = make_uchar2(o[0] > thresh, o[1] > thresh);
    *reinterpret_cast<uchar2*>(seg + base + W4) = make_uchar2(o[2] > thresh, o[3] > thresh);
  }
}

// cv2.resize(INTER_LINEAR) for uint8 HWC, bit-exact with OpenCV's generic fixed-point path (DetPreProcess resize,
// SURVEY App. B): 11-bit coefficients (xa/ya, built on the host in float32 exactly as OpenCV does), horizontal pass in
// int, vertical pass  ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2.  One thread per output pixel (3 channels).
static __global__ void resize_linear_u8_kernel(const uint8_t* __restrict__ src, int N, int SH, int SW, uint8_t* __restrict__ dst, int DH,
                                               int DW, const int* __restrict__ xi, const short* __restrict__ xa,
                                               const int* __restrict__ yi, const short* __restrict__ ya) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * DH * DW;
  if (idx >= total) return;
  int x = idx % DW; int y = (idx / DW) % DH; int n = idx / ((long long)DW * DH);
  const int x0 = xi[x], x1 = min(x0 + 1, SW - 1);
  const int ys = yi[y];
  const int y0 = min(max(ys, 0), SH - 1), y1 = min(max(ys + 1, 0), SH - 1);
  const int a0 = xa[2 * x], a1 = xa[2 * x + 1], b0 = ya[2 * y], b1 = ya[2 * y + 1];
  const uint8_t* r0 = src + ((long long)n * SH + y0) * SW * 3;
  const uint8_t* r1 = src + ((long long)n * SH + y1) * SW * 3;
  uint8_t* o = dst + idx * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int S0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
    const int S1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
    int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

// prob -> seg byte (used when the prob map came from elsewhere, e.g. the DB C-ABI entry)
static __global__ void threshold_kernel(const float* __restrict__ prob, float thresh, uint8_t* __restrict__ seg, long long total) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < total) {
    float4 p = *reinterpret_cast<const float4*>(prob + i);
    *reinterpret_cast<uchar4*>(seg + i) = make_uchar4(p.x > thresh, p.y > thresh, p.z > thresh, p.w > thresh);
  } else {
    for (; i < total; ++i) seg[i] = prob[i] > thresh;
  }
}

// cv2.dilate(seg, ones(2,2)) : anchor (1,1) -> out[y,x] = OR seg[y-1..y, x-1..x]
// (ocr_patch.py:232-235; border pixels outside the image are ignored).
static __global__ void dilate2x2_kernel(const uint8_t* __restrict__ seg, int N, int H, int W, uint8_t* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // 4 pixels per thread along x
  int W4 = W / 4;
  long long total = (long long)N * H * W4;
  if (idx >= total) return;
  int xq = idx % W4; int y = (idx / W4) % H; int n = idx / ((long long)W4 * H);
  const uint8_t* row = seg + ((long long)n * H + y) * W + xq * 4;
  uchar4 c = *reinterpret_cast<const uchar4*>(row);
  uint8_t cl = (xq > 0) ? row[-1] : 0;
  uchar4 u = make_uchar4(0, 0, 0, 0);
  uint8_t ul = 0;
  if (y > 0) {
    u = *reinterpret_cast<const uchar4*>(row - W);
    ul = (xq > 0) ? row[-W - 1] : 0;
  }
  uchar4 o;
  o.x = c.x | cl | u.x | ul;
  o.y = c.y | c.x | u.y | u.x;
  o.z = c.z | c.y | u.z | u.y;
  o.w = c.w | c.z | u.w | u.z;
  *reinterpret_cast<uchar4*>(out + ((long long)n * H + y) * W + xq * 4) = o;
}

// =====================================================================================
// rec: avg_pool2d([3,2]) on [N,3,W,C] -> [N,1,W/2,C].   rec_lcnetv4.py:311
// =====================================================================================
template <typename T>
__global__ void avgpool3x2_kernel(const T* __restrict__ in, int N, int W, int C, T* __restrict__ out) {
  int G = C / 8, OW = W / 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * OW * G;
  if (idx >= total) return;
  int g = idx % G; long long p = idx / G;
  int ox = p % OW; int n = p / OW;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (int y = 0; y < 3; ++y)
    for (int dx = 0; dx < 2; ++dx) {
      float v[8];
      Vec8<T>::load(in + (((long long)n * 3 + y) * W + ox * 2 + dx) * C + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] *= (1.f / 6.f);
  Vec8<T>::store(out + p * C + g * 8, s);
}

// LayerNorm over the last dim (C <= 1024), one warp per row; optional residual add
// (out = ln(x)*g+b [+ res]).   necks/rnn.py:301-318,376-379
template <typename T>
__global__ void layernorm_kernel(const T* __restrict__ x, long long rows, int C, const float* __restrict__ g,
                                 const float* __restrict__ b, float eps, const T* __restrict__ res, T* __restrict__ out) {
  long long row = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xp = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f32<T>(xp[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
  float mean = s / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = to_f32<T>(xp[c]) - mean; v += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffff, v, o);
  float rstd = rsqrtf(v / C + eps);
  for (int c = lane; c < C; c += 32) {
    float y = (to_f32<T>(xp[c]) - mean) * rstd * g[c] + b[c];
    if (res != nullptr) y += to_f32<T>(res[row * C + c]);
    out[row * C + c] = from_f32<T>(y);
  }
}

// LightSVTR self-attention: qkv [N,T,3,heads,hd] -> out [N,T,heads*hd].  One block per
// (n, head); K,V of the head staged in smem; one thread per query with online softmax.
// necks/rnn.py:253-265 (scale = hd^-0.5)
template <typename T, int HD>
__global__ void attention_kernel(const T* __restrict__ qkv, int Tn, int heads, float scale, T* __restrict__ out) {
  extern __shared__ float kv[];  // K[T][HD], V[T][HD]
  float* K = kv;
  float* V = kv + Tn * HD;
  int n = blockIdx.x / heads, h = blockIdx.x % heads;
  int C = heads * HD;
  const T* base = qkv + (long long)n * Tn * 3 * C;
  for (int i = threadIdx.x; i < Tn * HD; i += blockDim.x) {
    int t = i / HD, d = i % HD;
    K[i] = to_f32<T>(base[(long long)t * 3 * C + C + h * HD + d]);
    V[i] = to_f32<T>(base[(long long)t * 3 * C + 2 * C + h * HD + d]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = to_f32<T>(base[(long long)t * 3 * C + h * HD + d]) * scale;
    float m = -INFINITY, l = 0.f, acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    for (int j = 0; j < Tn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], K[j * HD + d], s);
      float mn = fmaxf(m, s);
      float corr = __expf(m - mn), pj = __expf(s - mn);
      l = l * corr + pj;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = acc[d] * corr + pj * V[j * HD + d];
      m = mn;
    }
    float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) out[((long long)n * Tn + t) * C + h * HD + d] = from_f32<T>(acc[d] * inv);
  }
}

// =====================================================================================
// CTC greedy decode.  Stage 1 (fused in the head GEMM epilogue) leaves per-row, per-N-tile
// partials (max, argmax, sum exp(x-max)); this kernel merges them into (id, prob) where
// prob = softmax max = 1/sum exp(x - max)   (torch.py:186-187 + CTCLabelDecode argmax/max).
// Ties resolve to the lowest index, as numpy argmax does.
// =====================================================================================
static __global__ void ctc_merge_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx, const float* __restrict__ psum,
                                 int rows, int tiles, int* __restrict__ ids, float* __restrict__ probs) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float m = -INFINITY; int id = 0;
  for (int t = 0; t < tiles; ++t) {
    float v = pmax[(long long)r * tiles + t];
    if (v > m) { m = v; id = pidx[(long long)r * tiles + t]; }
  }
  float s = 0.f;
  for (int t = 0; t < tiles; ++t) s += psum[(long long)r * tiles + t] * __expf(pmax[(long long)r * tiles + t] - m);
  ids[r] = id;
  probs[r] = 1.f / s;
}

// CTC collapse: keep t where id != blank(0) and id != id[t-1]; one warp per text line.
// out_ids [N,T] (compacted, -1 padded), out_len [N], conf [N] = mean kept prob (0 if none).
static __global__ void ctc_collapse_kernel(const int* __restrict__ ids, const float* __restrict__ probs, int N, int T,
                                    int* __restrict__ out_ids, int* __restrict__ out_len, float* __restrict__ conf) {
  int n = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (n >= N) return;
  int count = 0; float sum = 0.f;
  for (int t0 = 0; t0 < T; t0 += 32) {
    int t = t0 + lane;
    int id = (t < T) ? ids[(long long)n * T + t] : 0;
    int prev = (t > 0 && t < T) ? ids[(long long)n * T + t - 1] : -1;
    bool keep = (t < T) && id != 0 && id != prev;
    unsigned mask = __ballot_sync(0xffffffff, keep);
    int pos = count + __popc(mask & ((1u << lane) - 1));
    if (keep) out_ids[(long long)n * T + pos] = id;
    float pv = keep ? probs[(long long)n * T + t] : 0.f;
    // sequential-order sum (lane 0..31) for run-to-run determinism
    for (int l = 0; l < 32; ++l) { float x = __shfl_sync(0xffffffff, pv, l); sum += x; }
    count += __popc(mask);
  }
  for (int t = count + lane; t < T; t += 32) out_ids[(long long)n * T + t] = -1;
  if (lane == 0) { out_len[n] = count; conf[n] = count ? sum / count : 0.f; }
}

// row-wise softmax over V (compat path for the InferSession seam that must return the
// full [N,T,V] probability tensor, torch.py:186-192); one block per row.
static __global__ void softmax_rows_kernel(const float* __restrict__ logits, int V, float* __restrict__ out) {
  __shared__ float red[32];
  const float* x = logits + (long long)blockIdx.x * V;
  float* o = out + (long long)blockIdx.x * V;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += blockDim.x) m = fmaxf(m, x[i]);
#pragma unroll
  for (int k = 16; k; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffff, m, k));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += blockDim.x) s += expf(x[i] - m);
#pragma unroll
  for (int k = 16; k; k >>= 1) s += __shfl_xor_sync(0xffffffff, s, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
  float inv = 1.f / s;
  for (int i = threadIdx.x; i < V; i += blockDim.x) o[i] = expf(x[i] - m) * inv;
}

}  // namespace rdb
