// fp32 SIMT GEMM  out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ res), fused epilogues.
// This is the EXACT-precision pointwise-conv / linear engine (precision mode "fp32");
// the tensor-core engine lives in gemm_tc.cuh.  A is NHWC activations viewed as [M,K].
// Covers: channel_conv1/2 (rec_lcnetv4.py:210-224), insert/pointwise convs
// (db_fpn.py:326-331,350-356), LightSVTR 1x1 convs + linears (rnn.py:238-290) and the
// CTC head Linear(120->18710) (rec_multi_head.py:45,70) with a fused greedy-decode epilogue.
#pragma once
#include "common.cuh"

namespace rdb {

struct GemmArgs {
  const void* A; int lda;        // [M,K], row stride lda (elements)
  const float* W;                // [N,K] fp32
  const float* bias;             // [N] or null
  const void* res; int ldr;      // residual [M,N] (same storage type as out) or null
  void* out; int ldc; int c_off; // out[r*ldc + c_off + c]
  int M, N, K;
  int act;
  // EPI_CTC outputs
  float* pmax; int* pidx; float* psum; int tiles;
};

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

template <typename TA, typename TO, int ACT, bool CTC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const TA* A = reinterpret_cast<const TA*>(g.A);
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_row = tid / 2, a_k = (tid % 2) * 8;
  const int b_n = tid / 4, b_k = (tid % 4) * 4;
  for (int k0 = 0; k0 < g.K; k0 += SG_BK) {
    {  // A tile
      float v[8];
      long long r = m0 + a_row;
      if (r < g.M && k0 + a_k + 8 <= g.K) {
        Vec8<TA>::load(A + r * g.lda + k0 + a_k, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < g.M && k0 + a_k + j < g.K) ? to_f32<TA>(A[r * g.lda + k0 + a_k + j]) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) As[a_k + j][a_row] = v[j];
    }
    {  // W tile
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      int n = n0 + b_n;
      if (n < g.N) {
        if (k0 + b_k + 4 <= g.K) {
          w4 = __ldg(reinterpret_cast<const float4*>(g.W + (long long)n * g.K + k0 + b_k));
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j) if (k0 + b_k + j < g.K) t[j] = g.W[(long long)n * g.K + k0 + b_k + j];
          w4 = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      Bs[b_k + 0][b_n] = w4.x; Bs[b_k + 1][b_n] = w4.y; Bs[b_k + 2][b_n] = w4.z; Bs[b_k + 3][b_n] = w4.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int c0 = n0 + tx * 4;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (g.bias != nullptr && c0 + j < g.N) ? g.bias[c0 + j] : 0.f;

  if (!CTC) {
    TO* out = reinterpret_cast<TO*>(g.out);
    const TO* res = reinterpret_cast<const TO*>(g.res);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      long long r = m0 + ty * 8 + i;
      if (r >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = c0 + j;
        if (c >= g.N) continue;
        float v = apply_act<ACT>(acc[i][j] + bv[j]);
        if (res != nullptr) v += to_f32<TO>(res[r * g.ldr + c]);
        out[r * g.ldc + g.c_off + c] = from_f32<TO>(v);
      }
    }
  } else {
    // fused greedy-decode partials over this tile's 64 columns (16 lanes share a row group)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float m = -INFINITY; int id = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = c0 + j;
        float v = (c < g.N) ? acc[i][j] + bv[j] : -INFINITY;
        acc[i][j] = v;
        if (v > m) { m = v; id = c; }
      }
#pragma unroll
      for (int o = 8; o; o >>= 1) {
        float m2 = __shfl_xor_sync(0xffffffff, m, o);
        int id2 = __shfl_xor_sync(0xffffffff, id, o);
        if (m2 > m || (m2 == m && id2 < id)) { m = m2; id = id2; }
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (acc[i][j] == -INFINITY) ? 0.f : __expf(acc[i][j] - m);
#pragma unroll
      for (int o = 8; o; o >>= 1) s += __shfl_xor_sync(0xffffffff, s, o);
      long long r = m0 + ty * 8 + i;
      if (tx == 0 && r < g.M) {
        g.pmax[r * g.tiles + blockIdx.y] = m;
        g.pidx[r * g.tiles + blockIdx.y] = id;
        g.psum[r * g.tiles + blockIdx.y] = s;
      }
    }
  }
}

template <typename TA, typename TO>
void launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  dim3 grid(cdiv(g.M, SG_BM), cdiv(g.N, SG_BN));
  switch (g.act) {
    case ACT_NONE: gemm_simt_kernel<TA, TO, ACT_NONE, false><<<grid, 256, 0, st>>>(g); break;
    case ACT_RELU: gemm_simt_kernel<TA, TO, ACT_RELU, false><<<grid, 256, 0, st>>>(g); break;
    case ACT_GELU: gemm_simt_kernel<TA, TO, ACT_GELU, false><<<grid, 256, 0, st>>>(g); break;
    case ACT_SILU: gemm_simt_kernel<TA, TO, ACT_SILU, false><<<grid, 256, 0, st>>>(g); break;
    case ACT_HSWISH: gemm_simt_kernel<TA, TO, ACT_HSWISH, false><<<grid, 256, 0, st>>>(g); break;
    default: throw Error("gemm_simt: bad act");
  }
  RDB_LAUNCH_CHECK();
}

template <typename TA>
void launch_gemm_simt_ctc(GemmArgs g, cudaStream_t st) {
  dim3 grid(cdiv(g.M, SG_BM), cdiv(g.N, SG_BN));
  g.tiles = grid.y;
  gemm_simt_kernel<TA, float, ACT_NONE, true><<<grid, 256, 0, st>>>(g);
  RDB_LAUNCH_CHECK();
}

}  // namespace rdb
