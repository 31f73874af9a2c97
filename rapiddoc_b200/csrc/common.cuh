// Shared helpers for the rapiddoc_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace rdb {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define RDB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw rdb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + \
                       ":" + std::to_string(__LINE__));                                      \
  } while (0)

#define RDB_CHECK(cond, msg)                                                           \
  do {                                                                                 \
    if (!(cond)) throw rdb::Error(std::string("check failed: ") + #cond + " — " + msg); \
  } while (0)

#define RDB_LAUNCH_CHECK() RDB_CUDA(cudaGetLastError())

// cudaFuncSetAttribute is per DEVICE: one-time flags are kept per device index, so a second engine on another GPU of the
// same process sets its kernels' dynamic shared memory limit too.
constexpr int kMaxDevices = 64;
inline bool first_on_device(bool (&flags)[kMaxDevices]) {
  int d = 0;
  cudaGetDevice(&d);
  if (d < 0 || d >= kMaxDevices) return true;
  if (flags[d]) return false;
  flags[d] = true;
  return true;
}

// A/B switches (RDB_*): read from the environment ONCE per process and cached — the hot path never calls getenv.  Tests that
// flip a switch between calls use rdb_switches_reload().  Switches that exist only to measure what a stage costs (and give
// wrong results by design, e.g. RDB_SE_SKIP) or that print debug marks compile in only with -DRDB_DEBUG_SWITCHES.
struct Switches {
  std::unordered_map<std::string, std::string> vals;   // name -> value for the names that are set
  std::unordered_map<std::string, bool> seen;
  std::mutex mu;
  static Switches& get() { static Switches s; return s; }
  const char* lookup(const char* name) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = seen.find(name);
    if (it == seen.end()) {
      const char* e = std::getenv(name);
      seen[name] = e != nullptr;
      if (e) vals[name] = e;
      return e ? vals[name].c_str() : nullptr;
    }
    return it->second ? vals[name].c_str() : nullptr;
  }
  void reload() { std::lock_guard<std::mutex> lk(mu); vals.clear(); seen.clear(); }
};
inline const char* sw_get(const char* name) { return Switches::get().lookup(name); }
#ifdef RDB_DEBUG_SWITCHES
inline const char* sw_debug(const char* name) { return sw_get(name); }
#else
inline const char* sw_debug(const char*) { return nullptr; }
#endif

// Every C-ABI entry point selects its engine's device; the caller's current device (torch's, for instance) is put back on exit.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (device != prev) cudaSetDevice(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SILU = 3, ACT_SIGMOID = 4, ACT_GELUF = 5, ACT_HSWISH = 6 };

// erf-GELU for the fp16 path, 11 issue slots instead of erff's ~24 (the GELU GEMM epilogues are issue-bound):
// gelu(x) = x * Phi(x) with Phi(x) = sigmoid(x * h(x^2)), h = degree-4 least-max fit of logit(Phi(x)) / x on |x| <= 8
// (x^2 clamped to 64 beyond).  Max |error| against the exact erf form, evaluated in float32 over [-60, 60]: 3.4e-6 —
// 1/140 of the fp16 rounding step of the stored result at |x| = 1.  Coefficients carry the -log2(e) of exp -> ex2.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(x * x, 64.f);
  float h = -3.2291283e-06f;
  h = fmaf(h, u, 8.8241026e-05f);
  h = fmaf(h, u, 3.6025618e-04f);
  h = fmaf(h, u, -1.05226646e-01f);
  h = fmaf(h, u, -2.30204543e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * h));
  return __fdividef(x, 1.f + e);
}


template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));  // exact erf GELU
  if (ACT == ACT_GELUF) return gelu_fast(x);
  if (ACT == ACT_SILU) return x / (1.f + __expf(-x));
  if (ACT == ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  if (ACT == ACT_HSWISH) return x * fminf(fmaxf(fmaf(x, 1.f / 6.f, 0.5f), 0.f), 1.f);   // ONNX HardSwish: x * max(0, min(1, x/6 + 0.5))
  return x;
}

__device__ __forceinline__ float apply_act_rt(float x, int act) {
  switch (act) {
    case ACT_RELU: return apply_act<ACT_RELU>(x);
    case ACT_GELU: return apply_act<ACT_GELU>(x);
    case ACT_GELUF: return apply_act<ACT_GELUF>(x);
    case ACT_SILU: return apply_act<ACT_SILU>(x);
    case ACT_SIGMOID: return apply_act<ACT_SIGMOID>(x);
    case ACT_HSWISH: return apply_act<ACT_HSWISH>(x);
    default: return x;
  }
}

// ---- 8-channel vector access, fp32 math, storage T in {float, __half} -----------------
template <typename T>
struct Vec8;

template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};

template <>
struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace rdb
