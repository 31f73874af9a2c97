// Host-side runtime shared by the det / rec engines: weight blob, device buffer pool,
// launch helpers.  One engine = one (GPU, model); all work of a call goes to one stream.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace rdb {

// ---------------------------------------------------------------- weights ("RDW1" blob)
struct Tensor {
  const float* d = nullptr;  // device fp32
  const __half* h = nullptr; // device fp16 copy (same layout), used by the tcgen05 GEMMs
  int ndim = 0;
  int shape[4] = {1, 1, 1, 1};
  size_t numel = 0;
};

static __global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

class Weights {
 public:
  Weights(const void* blob, size_t nbytes) {
    RDB_CHECK(nbytes >= 8 && std::memcmp(blob, "RDW1", 4) == 0, "not an RDW1 weight blob");
    const uint8_t* p = static_cast<const uint8_t*>(blob);
    uint32_t n;
    std::memcpy(&n, p + 4, 4);
    RDB_CHECK(8 + (size_t)n * 100 <= nbytes, "truncated RDW1 header");
    RDB_CUDA(cudaMalloc(&dev_, nbytes));
    RDB_CUDA(cudaMemcpy(dev_, blob, nbytes, cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < n; ++i) {
      const uint8_t* e = p + 8 + (size_t)i * 100;
      char name[65];
      std::memcpy(name, e, 64);
      name[64] = 0;
      Tensor t;
      uint32_t nd, shp[4];
      uint64_t off, numel;
      std::memcpy(&nd, e + 64, 4);
      std::memcpy(shp, e + 68, 16);
      std::memcpy(&off, e + 84, 8);
      std::memcpy(&numel, e + 92, 8);
      RDB_CHECK(off + numel * 4 <= nbytes, "RDW1 entry out of bounds");
      t.ndim = nd;
      for (int k = 0; k < 4; ++k) t.shape[k] = shp[k];
      t.numel = numel;
      t.d = reinterpret_cast<const float*>(static_cast<uint8_t*>(dev_) + off);
      map_[name] = t;
    }
    // fp16 mirror of the whole blob (element i of the fp32 payload <-> element i of the mirror)
    const size_t nfl = nbytes / 4;
    RDB_CUDA(cudaMalloc(&half_, nfl * 2));
    f32_to_f16_kernel<<<(unsigned)((nfl + 255) / 256), 256>>>(reinterpret_cast<const float*>(dev_), half_, nfl);
    RDB_CUDA(cudaGetLastError());
    RDB_CUDA(cudaDeviceSynchronize());
    for (auto& kv : map_) kv.second.h = half_ + (kv.second.d - reinterpret_cast<const float*>(dev_));
  }
  ~Weights() { if (dev_) cudaFree(dev_); if (half_) cudaFree(half_); for (auto& kv : derived_) cudaFree(kv.second); }
  const Tensor& get(const std::string& name) const {
    auto it = map_.find(name);
    if (it == map_.end()) throw Error("weight tensor missing: " + name);
    return it->second;
  }
  bool has(const std::string& name) const { return map_.count(name) != 0; }
  // derived device buffers (e.g. weight slices pre-arranged as shared-memory images), built once on first use and owned here
  __half* derived(const std::string& key, size_t halves, bool* fresh) const {
    auto it = derived_.find(key);
    *fresh = it == derived_.end();
    if (!*fresh) return it->second;
    __half* p = nullptr;
    RDB_CUDA(cudaMalloc(&p, halves * sizeof(__half)));
    derived_[key] = p;
    return p;
  }

 private:
  void* dev_ = nullptr;
  __half* half_ = nullptr;
  std::unordered_map<std::string, Tensor> map_;
  mutable std::unordered_map<std::string, __half*> derived_;
};

// ---------------------------------------------------------------- device buffer pool
// Caching allocator for the call-scoped activation / staging buffers of one compute lane.  Every alloc/free of one engine
// call happens in program order on ONE stream, so a block may be handed out again as soon as it is freed.
//  * best-fit reuse: a request takes the smallest cached block of [bytes, 2*bytes] (RapidDoc feeds many page / crop shapes:
//    det buckets at 64-px granularity, rec widths per batch — exact-size keys would keep one full set per shape);
//  * the cache is bounded: when the cached (idle) bytes exceed the cap (default 12 GiB per lane, RDB_POOL_CAP_MB or
//    rdb_set_pool_cap_bytes) the least recently used blocks go back to the driver, so the torch models that share the GPU
//    keep their memory;
//  * reclaim(): after an exception unwound an infer call, its live blocks return to the cache instead of leaking.
class Pool {
 public:
  Pool() {
    if (const char* e = std::getenv("RDB_POOL_CAP_MB")) cap_ = (size_t)std::atoll(e) << 20;
  }
  ~Pool() { release_all(); }
  void* alloc(size_t bytes) {
    bytes = (bytes + 511) / 512 * 512;
    if (bytes == 0) bytes = 512;
    auto it = free_.lower_bound(bytes);
    while (it != free_.end() && it->second.empty()) it = free_.erase(it);
    if (it != free_.end() && it->first <= 2 * bytes) {
      Block b = it->second.back();
      it->second.pop_back();
      if (it->second.empty()) free_.erase(it);
      cached_ -= b.bytes;
      live_[b.p] = b.bytes;
      return b.p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim();
      RDB_CUDA(cudaMalloc(&p, bytes));
    }
    total_ += bytes;
    live_[p] = bytes;
    return p;
  }
  template <typename T>
  T* alloc_t(size_t n) { return static_cast<T*>(alloc(n * sizeof(T))); }
  void free(void* p) {
    if (!p) return;
    auto it = live_.find(p);
    RDB_CHECK(it != live_.end(), "pool: free of unknown pointer");
    free_[it->second].push_back(Block{p, it->second, ++tick_});
    cached_ += it->second;
    live_.erase(it);
  }
  // called by the engines at the end of a call: evict least-recently-used cached blocks down to the cap
  void enforce_cap() {
    if (cached_ <= cap_) return;
    std::vector<Block> all;
    for (auto& kv : free_) for (auto& b : kv.second) all.push_back(b);
    std::sort(all.begin(), all.end(), [](const Block& a, const Block& b) { return a.tick < b.tick; });
    std::unordered_map<void*, bool> drop;
    size_t c = cached_;
    for (auto& b : all) { if (c <= cap_) break; drop[b.p] = true; c -= b.bytes; }
    cudaDeviceSynchronize();
    for (auto it = free_.begin(); it != free_.end();) {
      auto& v = it->second;
      for (size_t i = 0; i < v.size();) {
        if (drop.count(v[i].p)) { cudaFree(v[i].p); total_ -= v[i].bytes; cached_ -= v[i].bytes; v[i] = v.back(); v.pop_back(); }
        else ++i;
      }
      it = v.empty() ? free_.erase(it) : std::next(it);
    }
  }
  void trim() {  // give every cached block back to the driver
    cudaDeviceSynchronize();
    for (auto& kv : free_)
      for (auto& b : kv.second) { cudaFree(b.p); total_ -= b.bytes; }
    free_.clear();
    cached_ = 0;
  }
  void reclaim() {  // an exception unwound the call that owned the live blocks
    std::vector<void*> ps;
    for (auto& kv : live_) ps.push_back(kv.first);
    for (void* p : ps) free(p);
  }
  void release_all() {
    trim();
    for (auto& kv : live_) cudaFree(kv.first);
    live_.clear();
  }
  void set_cap(size_t bytes) { cap_ = bytes; }
  size_t total_bytes() const { return total_; }
  size_t cached_bytes() const { return cached_; }
  size_t live_blocks() const { return live_.size(); }

 private:
  struct Block { void* p; size_t bytes; unsigned long long tick; };
  std::map<size_t, std::vector<Block>> free_;
  std::unordered_map<void*, size_t> live_;
  size_t total_ = 0, cached_ = 0, cap_ = (size_t)12 << 30;
  unsigned long long tick_ = 0;
};

inline bool is_device_ptr(const void* p) {
  if (p == nullptr) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------- launch context
// Optional per-kernel timing (rdb_profile_*): CUDA events recorded on the launching stream
// around every launch, resolved after the call's final synchronise.  Off by default.
struct Profiler {
  bool on = false;
  struct Rec { std::string name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> spare;
  std::map<std::string, std::pair<double, long long>> acc;  // name -> (ms, launches)
  cudaEvent_t get() {
    if (!spare.empty()) { cudaEvent_t e = spare.back(); spare.pop_back(); return e; }
    cudaEvent_t e; RDB_CUDA(cudaEventCreate(&e)); return e;
  }
  void resolve() {
    for (auto& r : recs) {
      RDB_CUDA(cudaEventSynchronize(r.b));
      float ms = 0.f;
      RDB_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      auto& p = acc[r.name]; p.first += ms; p.second += 1;
      spare.push_back(r.a); spare.push_back(r.b);
    }
    recs.clear();
  }
  static Profiler& global() { static Profiler p; return p; }
};

struct Ctx {
  cudaStream_t st = nullptr;
  Pool* pool = nullptr;
  long long launches = 0;
  int precision = 0;
  bool use_tc = false;   // fp16 mode: tcgen05 GEMMs (default) vs SIMT GEMMs on fp16 storage (RDB_GEMM=simt)
  int num_sms = 148;
  // bracket ONE kernel launch:  cx.begin("name"); kernel<<<...>>>(...); cx.end();
  void begin(const std::string& name) {
    Profiler& p = Profiler::global();
    if (!p.on) return;
    Profiler::Rec r{name, p.get(), p.get()};
    RDB_CUDA(cudaEventRecord(r.a, st));
    p.recs.push_back(r);
  }
  void end() {
    RDB_LAUNCH_CHECK();
    launches++;
    Profiler& p = Profiler::global();
    if (p.on) RDB_CUDA(cudaEventRecord(p.recs.back().b, st));
  }
  void finish() {  // after the work of a call is enqueued; resolves timings if profiling
    Profiler& p = Profiler::global();
    if (p.on) { RDB_CUDA(cudaStreamSynchronize(st)); p.resolve(); }
  }
};

constexpr int kThreads = 256;

template <typename K>
inline void set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) RDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

inline bool env_is(const char* name, const char* val) {
  // timing-only switches that change results are debug-build only
  const bool debug_only = std::strcmp(name, "RDB_SE_SKIP") == 0;
  const char* e = debug_only ? sw_debug(name) : sw_get(name);
  return e != nullptr && std::strcmp(e, val) == 0;
}

inline bool env_gemm_simt() {
  const char* e = sw_get("RDB_GEMM");
  return e != nullptr && std::strcmp(e, "simt") == 0;
}

template <int ACT>
inline void launch_tc_store_act(const tc::Plan& p, const CUtensorMap& mA, const CUtensorMap& mB, cudaStream_t st) {
  auto k = tc::gemm_tc_kernel<tc::EPI_STORE, ACT>;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
  k<<<p.grid, tc::kThreadsTc, p.smem, st>>>(mA, mB, p.a);
}
inline void launch_tc_store(const tc::Plan& p, const CUtensorMap& mA, const CUtensorMap& mB, int act, cudaStream_t st) {
  switch (act) {
    case ACT_NONE: launch_tc_store_act<ACT_NONE>(p, mA, mB, st); break;
    case ACT_RELU: launch_tc_store_act<ACT_RELU>(p, mA, mB, st); break;
    case ACT_GELU: launch_tc_store_act<ACT_GELU>(p, mA, mB, st); break;
    case ACT_GELUF: launch_tc_store_act<ACT_GELUF>(p, mA, mB, st); break;
    case ACT_SILU: launch_tc_store_act<ACT_SILU>(p, mA, mB, st); break;
    default: throw Error("gemm_tc: bad act");
  }
}

// tcgen05 GEMM launch: out[M,N] = act(A[M,K] W^T + b) (+res), everything fp16 in HBM.
struct TcFuse {   // optional fused RepLKFPN input stage: out = acc*colscale[img][col] + nearest_up2(up_res)
  const float* colscale = nullptr; int rows_per_img = 0;
  const __half* up_res = nullptr; int up_H = 0, up_W = 0;
};

inline void launch_gemm_tc(Ctx& cx, const __half* A, int lda, long long M, int K, const __half* Wh, int N, const float* bias, int act,
                           const __half* res, int ldr, __half* out, int ldc, int c_off, const TcFuse* fuse = nullptr) {
  tc::Plan p = tc::make_plan(M, N, K, cx.num_sms);
  tc::Args& a = p.a;
  a.bias = bias; a.res = res; a.ldr = ldr; a.out = out; a.ldc = ldc; a.c_off = c_off; a.act = act;
  if (fuse) { a.colscale = fuse->colscale; a.rows_per_img = fuse->rows_per_img; a.up_res = fuse->up_res; a.up_H = fuse->up_H; a.up_W = fuse->up_W; }
  CUtensorMap mA = tc::make_map(A, M, K, lda, a.AW, 128);
  CUtensorMap mB = tc::make_map(Wh, N, K, K, a.AW, a.BN);
  cx.begin("gemm_tc[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + ",res=" + ((res || (fuse && fuse->up_res)) ? "1" : "0") + "]");
  launch_tc_store(p, mA, mB, act, cx.st);
  cx.end();
}

// tcgen05 implicit-GEMM convolution (dense KHxKW conv, NHWC fp16): the stem 2x2 / 3x3-s2 convs and the
// DBHead 3x3 conv (rec_lcnetv4.py:152-154, det_db_head.py:104-109) on the tensor cores.
// in_wp > 0: the input rows are padded to in_wp pixels with zero pad pixels in memory and the conv runs in wide-row patch
// mode (the KW horizontal taps are one K range).  out_wp > 0: the output is written with that row pitch.
inline void launch_conv_tc(Ctx& cx, const char* name, const __half* in, int n, int H, int W, int C, const __half* Wh, int N, const float* bias,
                           int act, int KH, int KW, int sh, int sw, int pt, int pl, __half* out, int OH, int OW, int ldc, int c_off,
                           int in_wp = 0, int out_wp = 0, int in_ld = 0) {
  const bool wide = in_wp > 0 && sh == 1 && sw == 1;
  RDB_CHECK(!(wide && in_ld > 0), "conv_tc: a pixel pitch and the wide-row mode exclude each other");
  tc::Plan p = wide ? tc::make_conv_plan_wide(n, H, W, C, N, KH, KW, pt, OH, OW, cx.num_sms)
                    : tc::make_conv_plan(n, H, W, C, N, KH, KW, sh, sw, pt, pl, OH, OW, cx.num_sms);
  tc::Args& a = p.a;
  RDB_CHECK(!wide || (a.patch && a.resident), "conv_tc: wide-row mode needs the weights resident");
  a.bias = bias; a.res = nullptr; a.ldr = 0; a.out = out; a.ldc = ldc; a.c_off = c_off; a.act = act; a.out_wp = out_wp;
  CUtensorMap mA = wide ? tc::make_map_wide(in, n, H, W, in_wp, C, KW, a.AW, a.TH + a.KH - 1)
                        : tc::make_map_nhwc(in, n, H, in_wp > 0 ? in_wp : W, C, a.AW, a.patch ? a.TH + a.KH - 1 : a.TH, a.TW, sh, sw, in_ld);
  CUtensorMap mB = tc::make_map(Wh, N, KH * KW * C, KH * KW * C, a.AW, a.BN);
  cx.begin(std::string(name) + (wide ? "_tcw[P=" : (a.patch ? "_tcp[P=" : "_tc[P=")) + std::to_string((long long)n * OH * OW) + ",C=" + std::to_string(C) + ",N=" + std::to_string(N) + "]");
  launch_tc_store(p, mA, mB, act, cx.st);
  cx.end();
}

// DBHead tail on the tensor cores: ConvT2x2(24->24)+BN+ReLU as one [P,24]x[24,96] tcgen05 GEMM whose epilogue
// applies the final ConvT2x2(24->1), sigmoid, nan_to_num and the DB threshold (det_db_head.py:117-147).
inline void launch_head_tail_tc(Ctx& cx, const __half* hd, int n, int H, int W, const __half* w_up_h /*[96][24]*/, const float* b_up,
                                const float* w_fin, const float* b_fin, float thresh, float* prob, uint8_t* seg) {
  const long long M = (long long)n * H * W;
  tc::Plan p = tc::make_plan(M, 96, 24, cx.num_sms);
  tc::Args& a = p.a;
  a.epi_subs = 4;  // EPI_HEAD: sub-warp = ConvT tap, all four always active
  a.bias = b_up; a.hH = H; a.hW = W; a.w_fin = w_fin; a.b_fin = b_fin; a.thresh = thresh; a.prob = prob; a.seg = seg;
  CUtensorMap mA = tc::make_map(hd, M, 24, 24, a.AW, 128);
  CUtensorMap mB = tc::make_map(w_up_h, 96, 24, 24, a.AW, a.BN);
  auto k = tc::gemm_tc_kernel<tc::EPI_HEAD, ACT_NONE>;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
  cx.begin("head_tail_tc[P=" + std::to_string(M) + "]");
  k<<<p.grid, tc::kThreadsTc, p.smem, cx.st>>>(mA, mB, a);
  cx.end();
}

inline void launch_gemm_tc_ctc(Ctx& cx, const __half* A, int lda, long long M, int K, const __half* Wh, int N, const float* bias,
                               float* pmax, int* pidx, float* psum, int* tiles_out) {
  tc::Plan p = tc::make_plan(M, N, K, cx.num_sms);
  tc::Args& a = p.a;
  a.bias = bias; a.pmax = pmax; a.pidx = pidx; a.psum = psum;
  *tiles_out = a.tiles_n * tc::kEpiSubs;
  CUtensorMap mA = tc::make_map(A, M, K, lda, a.AW, 128);
  CUtensorMap mB = tc::make_map(Wh, N, K, K, a.AW, a.BN);
  auto k = tc::gemm_tc_kernel<tc::EPI_CTC, ACT_NONE>;
  static bool attr_done[rdb::kMaxDevices] = {};
  if (rdb::first_on_device(attr_done)) { RDB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
  cx.begin("gemm_tc_ctc[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + "]");
  k<<<p.grid, tc::kThreadsTc, p.smem, cx.st>>>(mA, mB, a);
  cx.end();
}

// ---------------------------------------------------------------- typed launch helpers
template <typename T>
struct Ops {
  // NHWC activation handle
  struct Act {
    T* p = nullptr; int n = 0, h = 0, w = 0, c = 0;
    int wp = 0;   // row pitch in pixels when the rows carry zero pad pixels (wide-row conv inputs); 0 = w
    int pitch() const { return wp ? wp : w; }
    long long pixels() const { return (long long)n * h * w; }
    long long numel() const { return pixels() * c; }
  };
  static Act make(Ctx& cx, int n, int h, int w, int c) {
    Act a; a.n = n; a.h = h; a.w = w; a.c = c;
    a.p = cx.pool->alloc_t<T>((size_t)a.numel());
    return a;
  }
  static void release(Ctx& cx, Act& a) { cx.pool->free(a.p); a.p = nullptr; }

  // pointwise conv / linear: out[M,N] = act(A W^T + b) (+res)
  static void gemm(Ctx& cx, const T* A, int lda, long long M, int K, const Tensor& W, const Tensor* bias, int act,
                   const T* res, int ldr, T* out, int ldc, int c_off) {
    int N = W.shape[0];
    RDB_CHECK(W.shape[1] == K, "gemm: weight K mismatch");
    GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W.d; g.bias = bias ? bias->d : nullptr; g.res = res; g.ldr = ldr;
    g.out = out; g.ldc = ldc; g.c_off = c_off; g.M = (int)M; g.N = N; g.K = K; g.act = act;
    RDB_CHECK(M < (1ll << 31), "gemm: M too large");
    if constexpr (std::is_same<T, __half>::value) {
      if (cx.use_tc) {
        launch_gemm_tc(cx, A, lda, M, K, W.h, N, bias ? bias->d : nullptr, act, res, ldr, out, ldc, c_off);
        return;
      }
    }
    cx.begin("gemm_simt[M=" + std::to_string(M) + ",K=" + std::to_string(K) + ",N=" + std::to_string(N) + ",res=" + (res ? "1" : "0") + "]");
    launch_gemm_simt<T, T>(g, cx.st);
    cx.end();
  }
  static void pw(Ctx& cx, const Act& in, const Tensor& W, const Tensor* bias, int act, const T* res, Act& out) {
    gemm(cx, in.p, in.c, in.pixels(), in.c, W, bias, act, res, out.c, out.p, out.c, 0);
  }

  // tile shape of the smem-tiled 3x3 depthwise kernel for a map of h x w
  static void dw3_tile(int h, int& TH, int& TW) {
    TH = h >= 8 ? 8 : 4; TW = 32;
    if (h == 12 || h == 6 || h == 3) { TH = h; TW = 16; }
  }
  static int dw3_tiles(int h, int w) { int TH, TW; dw3_tile(h, TH, TW); return cdiv(w, TW) * cdiv(h, TH); }

  // pool_partial != null (3x3 stride-1 only): the kernel also writes per-tile channel sums [n][tiles][C] for the SE pool
  // and *pooled is set when it did
  template <int KH, int KW, int ACT, bool ADD_IN>
  static void dwconv(Ctx& cx, const Act& in, int sh, int sw, const Tensor& w, const Tensor& b, Act& out, float* pool_partial = nullptr,
                     bool* pooled = nullptr) {
    if (pooled) *pooled = false;
    if constexpr (std::is_same<T, __half>::value && KH == 7 && KW == 7 && ACT == ACT_NONE && !ADD_IN) {
      if (sh == 1 && sw == 1 && in.c % 32 == 0 && !env_is("RDB_DW", "simple")) {
        constexpr int TH = 8, TW = 32, G = 4;
        if (cx.use_tc && !env_is("RDB_DW7", "f32")) {   // fp16/tcgen05 mode: packed-half row taps, fp32 accumulation across rows
          const size_t sm = (size_t)(TH + 6) * (TW + 6) * (16 * G + 16) + 49 * G * 16;
          dim3 grid(cdiv(in.w, TW), cdiv(in.h, TH), in.n * (in.c / (8 * G)));
          cx.begin("dwconv7x7_h2[P=" + std::to_string(out.pixels()) + ",C=" + std::to_string(in.c) + ",s=1]");
          auto k = dwconv_tiled_h2_kernel<7, G, TH, TW, 1>;
          set_smem(k, sm);
          k<<<grid, G * (TW / 4) * TH, sm, cx.st>>>(in.p, in.n, in.h, in.w, in.c, w.h, b.d, out.p);
          cx.end();
          return;
        }
        auto k = dwconv_tiled_kernel<T, 7, G, TH, TW>;
        const size_t sm = (size_t)(TH + 6) * (TW + 6) * (16 * G + 16) + 49 * 8 * G * sizeof(float);
        set_smem(k, sm);
        dim3 grid(cdiv(in.w, TW), cdiv(in.h, TH), in.n * (in.c / (8 * G)));
        cx.begin("dwconv7x7_tiled[P=" + std::to_string(out.pixels()) + ",C=" + std::to_string(in.c) + ",s=1]");
        k<<<grid, G * (TW / 4) * TH, sm, cx.st>>>(in.p, in.n, in.h, in.w, in.c, w.d, b.d, out.p, nullptr);
        cx.end();
        return;
      }
    }
    if constexpr (std::is_same<T, __half>::value && KH == 3 && KW == 3 && ACT == ACT_NONE && !ADD_IN) {
      if (sh == 1 && sw == 1 && in.c % 16 == 0 && !env_is("RDB_DW", "simple")) {
        // tile shape: the recogniser's maps are 12 / 6 / 3 rows x 80 columns -> tiles that cover the height exactly and
        // 16-column tiles (80 = 5 x 16); page-sized maps use 8 x 32
        const bool g4 = (in.c % 32 == 0);
        const int G = g4 ? 4 : 2;
        int TH, TW;
        dw3_tile(in.h, TH, TW);
        const size_t sm = (size_t)(TH + 2) * (TW + 2) * (16 * G + 16) + 9 * 8 * G * sizeof(float);
        dim3 grid(cdiv(in.w, TW), cdiv(in.h, TH), in.n * (in.c / (8 * G)));
        const int threads = G * (TW / 4) * TH;
        if (cx.use_tc && env_is("RDB_DW3", "h2")) {
          // EXPERIMENT, off by default: packed-half 3x3 taps with fp16-rounded weights.  +2.8 % det / +4.9 % rec, but the 14
          // backbone blocks compound the extra rounding: max |prob - oracle| reaches 4.3e-2 on a noise page (> the 3e-2 this
          // mode promises, tools/err_stats.py), so the fp32-FMA kernel below stays the product path
          const size_t smh = (size_t)(TH + 2) * (TW + 2) * (16 * G + 16) + 9 * G * 16;
          cx.begin("dwconv3x3_h2[P=" + std::to_string(out.pixels()) + ",C=" + std::to_string(in.c) + ",s=1]");
#define RDB_DW3H(GG, TTH, TTW) dwconv_tiled_h2_kernel<3, GG, TTH, TTW, 1><<<grid, threads, smh, cx.st>>>(in.p, in.n, in.h, in.w, in.c, w.h, b.d, out.p)
          if (g4) {
            if (TH == 8) RDB_DW3H(4, 8, 32); else if (TH == 4) RDB_DW3H(4, 4, 32); else if (TH == 12) RDB_DW3H(4, 12, 16);
            else if (TH == 6) RDB_DW3H(4, 6, 16); else RDB_DW3H(4, 3, 16);
          } else {
            if (TH == 8) RDB_DW3H(2, 8, 32); else if (TH == 4) RDB_DW3H(2, 4, 32); else if (TH == 12) RDB_DW3H(2, 12, 16);
            else if (TH == 6) RDB_DW3H(2, 6, 16); else RDB_DW3H(2, 3, 16);
          }
#undef RDB_DW3H
          cx.end();
          return;
        }
        cx.begin("dwconv3x3_tiled[P=" + std::to_string(out.pixels()) + ",C=" + std::to_string(in.c) + ",s=1]");
        if (pooled) *pooled = pool_partial != nullptr;
#define RDB_DW3(GG, TTH, TTW) dwconv_tiled_kernel<T, 3, GG, TTH, TTW><<<grid, threads, sm, cx.st>>>(in.p, in.n, in.h, in.w, in.c, w.d, b.d, out.p, pool_partial)
        if (g4) {
          if (TH == 8) RDB_DW3(4, 8, 32); else if (TH == 4) RDB_DW3(4, 4, 32); else if (TH == 12) RDB_DW3(4, 12, 16);
          else if (TH == 6) RDB_DW3(4, 6, 16); else RDB_DW3(4, 3, 16);
        } else {
          if (TH == 8) RDB_DW3(2, 8, 32); else if (TH == 4) RDB_DW3(2, 4, 32); else if (TH == 12) RDB_DW3(2, 12, 16);
          else if (TH == 6) RDB_DW3(2, 6, 16); else RDB_DW3(2, 3, 16);
        }
#undef RDB_DW3
        cx.end();
        return;
      }
    }
    long long total = out.pixels() * (in.c / 8);
    cx.begin("dwconv" + std::to_string(KH) + "x" + std::to_string(KW) + "[P=" + std::to_string(out.pixels()) + ",C=" + std::to_string(in.c) + ",s=" + std::to_string(sh * sw) + "]");
    dwconv_kernel<T, KH, KW, ACT, ADD_IN><<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(
        in.p, in.n, in.h, in.w, in.c, sh, sw, w.d, b.d, out.p, out.h, out.w);
    cx.end();
  }

  // SE gate (mode 0: hardsigmoid; mode 1: 1+clip(.2z+.5)) -> float gate[n][C] from pool.
  // pre != null: x is the input of the bias-free 1x1 conv `pre` [C][x.c] and the SE acts on the conv OUTPUT (C channels).
  // partial != null: per-tile channel sums [n][chunks_in][x.c] already written by the producer of x (fused pool)
  static float* se_gate(Ctx& cx, const Act& x, const Tensor& w1, const Tensor& b1, const Tensor& w2, const Tensor& b2, int mode,
                        const Tensor* pre = nullptr, float* partial_in = nullptr, int chunks_in = 0) {
    int HW = x.h * x.w, Cp = x.c, Cr = w1.shape[0];
    const int C = pre ? pre->shape[0] : Cp;
    RDB_CHECK(Cp <= 512, "se: more than 512 pooled channels");
    if (partial_in != nullptr && !env_is("RDB_SE_SKIP", "1")) {
      float* gate = cx.pool->alloc_t<float>((size_t)x.n * C);
      cx.begin("se_fc");
      se_fc_kernel<<<x.n, 512, (size_t)(2 * C + Cr + Cp + 512) * sizeof(float), cx.st>>>(partial_in, chunks_in, HW, C, Cr, w1.d, b1.d, w2.d, b2.d, mode, gate,
                                                                          pre ? pre->d : nullptr, Cp);
      cx.end();
      return gate;
    }
    int G = Cp / 8;
    int threads = 256;
    int P = threads / G;
    RDB_CHECK(P >= 1, "se: too many channels");
    int chunks = HW / (P * 16);
    if (chunks < 1) chunks = 1;
    if (chunks > 32) chunks = 32;
    if (env_is("RDB_SE_SKIP", "1")) {   // timing experiment only (wrong results): what do all SE pool/FC launches cost?
      float* g0 = cx.pool->alloc_t<float>((size_t)x.n * C);
      RDB_CUDA(cudaMemsetAsync(g0, 0x3f, (size_t)x.n * C * sizeof(float), cx.st));
      return g0;
    }
    float* partial = cx.pool->alloc_t<float>((size_t)x.n * chunks * Cp);
    float* gate = cx.pool->alloc_t<float>((size_t)x.n * C);
    cx.begin("se_pool[P=" + std::to_string(x.pixels()) + ",C=" + std::to_string(Cp) + "]");
    pool_partial_kernel<T><<<dim3(chunks, x.n), threads, (size_t)P * Cp * sizeof(float), cx.st>>>(x.p, HW, Cp, partial, chunks);
    cx.end();
    cx.begin("se_fc");
    se_fc_kernel<<<x.n, 512, (size_t)(2 * C + Cr + Cp + 512) * sizeof(float), cx.st>>>(partial, chunks, HW, C, Cr, w1.d, b1.d, w2.d, b2.d, mode, gate,
                                                                         pre ? pre->d : nullptr, Cp);
    cx.end();
    cx.pool->free(partial);
    return gate;
  }
  static void scale(Ctx& cx, Act& x, const float* gate) {
    if (env_is("RDB_SE_SKIP", "1")) return;
    long long total8 = x.numel() / 8;
    cx.begin("se_scale[P=" + std::to_string(x.pixels()) + ",C=" + std::to_string(x.c) + "]");
    scale_channels_kernel<T><<<cdiv(total8, kThreads), kThreads, 0, cx.st>>>(x.p, (long long)x.h * x.w, x.c, gate, total8);
    cx.end();
  }
};


}  // namespace rdb
#include "stem_fused.cuh"
#include "stem_planar.cuh"
#include "head_planar.cuh"
#include "mlp_tc.cuh"
namespace rdb {

// PPLCNetV4 block table (cin, cout, stride_h, stride_w, se) — rec_lcnetv4.py:7-43
struct BlockCfg { int cin, cout, sh, sw, se; };

// stem + 4 stages, shared by det and rec.  x: stem1 output already computed by caller.
template <typename T>
struct Backbone {
  using O = Ops<T>;
  using Act = typename O::Act;

  // stem2a..stem4 (rec_lcnetv4.py:160-169).  C1 = stem1 channels (24 det / 48 rec).
  template <int C1>
  static Act stem_rest(Ctx& cx, const Weights& w, Act& e1) {
    constexpr int CH = C1 / 2, C2 = 2 * C1;
    constexpr int CHP = (CH + 7) / 8 * 8;   // stem2a output channels padded to a 16-byte pixel pitch (TMA)
    const int n = e1.n, H1 = e1.h, W1 = e1.w;
    const int H2 = (H1 - 1) / 2 + 1, W2 = (W1 - 1) / 2 + 1;
    bool tc_path = false;
    if constexpr (std::is_same<T, __half>::value) tc_path = cx.use_tc && !env_is("RDB_CONV", "simt");
    RDB_CHECK(tc_path || e1.wp == 0, "stem: row-padded input needs the tcgen05 conv path");
    if constexpr (std::is_same<T, __half>::value && C1 == 24) {
      if (tc_path && !env_is("RDB_STEM", "unfused")) {   // stem2a .. stem4 in one kernel, intermediates in shared memory
        Act x = O::make(cx, n, H2, W2, C2);
        if (env_is("RDB_STEM", "copy")) launch_stem_fused<C1>(cx, w, e1.p, n, H1, W1, e1.pitch(), x.p, H2, W2);   // im2col-copy variant
        else launch_stem_planar<C1>(cx, w, e1.p, n, H1, W1, e1.pitch(), x.p, H2, W2);
        O::release(cx, e1);
        return x;
      }
    }
    Act cat = O::make(cx, n, H1, W1, C2);
    Act s3 = O::make(cx, n, H2, W2, C1);
    if (tc_path) {
      if constexpr (std::is_same<T, __half>::value) {
        // F.pad(0,1,0,1) + conv2x2 == conv with pt=pl=0 and zero fill past the bottom/right edge.  With row-padded
        // buffers (e1.wp = W1+1) the right pad pixel is real memory and both kx taps are ONE wide TMA row.
        const int a_wp = e1.wp ? W1 + 1 : 0;
        Act a = O::make(cx, n, H1, a_wp ? a_wp : W1, CHP);
        a.w = W1; a.wp = a_wp;
        launch_conv_tc(cx, "stem2a", e1.p, n, H1, W1, C1, w.get("stem2a.wp").h, CHP, w.get("stem2a.bp").d, ACT_RELU, 2, 2, 1, 1, 0, 0,
                       a.p, H1, W1, CHP, 0, e1.wp, a_wp);
        if (a_wp) {
          const long long rows = (long long)n * H1;
          zero_cols_kernel<T><<<cdiv(rows * (CHP / 8), kThreads), kThreads, 0, cx.st>>>(a.p, rows, a_wp, CHP, W1, 1);
          RDB_LAUNCH_CHECK();
          cx.launches++;
        }
        launch_conv_tc(cx, "stem2b", a.p, n, H1, W1, CHP, w.get("stem2b.wp").h, C1, w.get("stem2b.b").d, ACT_RELU, 2, 2, 1, 1, 0, 0,
                       cat.p, H1, W1, C2, C1, a_wp, 0);
        O::release(cx, a);
      }
    } else {
      Act a = O::make(cx, n, H1, W1, CH);
      {
        auto k = conv_direct_kernel<T, 2, 2, 1, 1, C1, CH, CH, ACT_RELU>;
        size_t sm = (size_t)(4 * C1 * CH + CH) * sizeof(float);
        set_smem(k, sm);
        cx.begin("stem2a");
        k<<<dim3(cdiv(a.pixels(), 128), 1), 128, sm, cx.st>>>(e1.p, n, H1, W1, 0, 0, w.get("stem2a.w").d, w.get("stem2a.b").d,
                                                              a.p, H1, W1, CH, 0);
        cx.end();
      }
      {
        auto k = conv_direct_kernel<T, 2, 2, 1, 1, CH, C1, C1 / 2, ACT_RELU>;
        size_t sm = (size_t)(4 * CH * C1 + C1) * sizeof(float);
        set_smem(k, sm);
        cx.begin("stem2b");
        k<<<dim3(cdiv(a.pixels(), 128), 2), 128, sm, cx.st>>>(a.p, n, H1, W1, 0, 0, w.get("stem2b.w").d, w.get("stem2b.b").d,
                                                              cat.p, H1, W1, C2, C1);
        cx.end();
      }
      O::release(cx, a);
    }
    {
      long long total = e1.pixels() * (C1 / 8);
      cx.begin("stem_pool");
      pool2x2_concat_kernel<T><<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(e1.p, n, H1, W1, C1, cat.p, C2, e1.pitch());
      cx.end();
    }
    O::release(cx, e1);
    if (tc_path) {
      if constexpr (std::is_same<T, __half>::value)
        launch_conv_tc(cx, "stem3", cat.p, n, H1, W1, C2, w.get("stem3.w").h, C1, w.get("stem3.b").d, ACT_RELU, 3, 3, 2, 2, 1, 1, s3.p, H2, W2,
                       C1, 0);
    } else {
      auto k = conv_direct_kernel<T, 3, 3, 2, 2, C2, C1, C1 / 2, ACT_RELU>;
      size_t sm = (size_t)(9 * C2 * C1 + C1) * sizeof(float);
      set_smem(k, sm);
      cx.begin("stem3");
      k<<<dim3(cdiv(s3.pixels(), 128), 2), 128, sm, cx.st>>>(cat.p, n, H1, W1, 1, 1, w.get("stem3.w").d, w.get("stem3.b").d,
                                                             s3.p, H2, W2, C1, 0);
      cx.end();
    }
    O::release(cx, cat);
    Act x = O::make(cx, n, H2, W2, C2);
    // stem4 weight is stored [C2][1][1][C1] == [C2][C1]
    Tensor w4 = w.get("stem4.w");
    w4.shape[1] = C1;
    O::pw(cx, s3, w4, &w.get("stem4.b"), ACT_RELU, nullptr, x);
    O::release(cx, s3);
    return x;
  }

  // one PPLCNetV4 block (rec_lcnetv4.py:226-236).  Consumes x unless keep_in.
  static Act block(Ctx& cx, const Weights& w, const std::string& name, const BlockCfg& c, Act& x, bool keep_in) {
    const int OH = (x.h + 2 - 3) / c.sh + 1, OW = (x.w + 2 - 3) / c.sw + 1;
    Act t = O::make(cx, x.n, OH, OW, c.cin);
    float* pool_partial = nullptr;
    bool pooled = false;
    const int tiles = O::dw3_tiles(x.h, x.w);
    if (c.se && c.sh == 1 && c.sw == 1 && !env_is("RDB_SE_POOL", "separate")) pool_partial = cx.pool->template alloc_t<float>((size_t)x.n * tiles * c.cin);
    O::template dwconv<3, 3, ACT_NONE, false>(cx, x, c.sh, c.sw, w.get(name + "dw.w"), w.get(name + "dw.b"), t, pool_partial, &pooled);
    if (!keep_in) O::release(cx, x);
    if (c.se) {
      float* gate = O::se_gate(cx, t, w.get(name + "se.w1"), w.get(name + "se.b1"), w.get(name + "se.w2"), w.get(name + "se.b2"), 0, nullptr,
                               pooled ? pool_partial : nullptr, tiles);
      O::scale(cx, t, gate);
      cx.pool->free(gate);
    }
    if (pool_partial) cx.pool->free(pool_partial);
    const bool rep = (c.sh == 1 && c.sw == 1 && c.cin == c.cout);
    if constexpr (std::is_same<T, __half>::value) {
      // channel mixer (pw1 -> GELU -> pw2 + residual) as one kernel, the 2C-wide intermediate stays in shared memory
      const bool big = c.cin == 192 && c.cout == 192 && !env_is("RDB_MLP", "small");   // weights streamed in three 128-column slices
      if (cx.use_tc && !env_is("RDB_MLP", "unfused") && ((c.cin == 48 && (c.cout == 48 || c.cout == 96)) || (c.cin == 96 && c.cout == 96) || big)) {
        Act y = O::make(cx, x.n, OH, OW, c.cout);
        const int act = env_is("RDB_GELU", "exact") ? ACT_GELU : ACT_GELUF;
        const long long M = t.pixels();
        const Tensor &w1 = w.get(name + "pw1.w"), &b1 = w.get(name + "pw1.b"), &w2 = w.get(name + "pw2.w"), &b2 = w.get(name + "pw2.b");
        if (c.cin == 48 && c.cout == 48) launch_mlp_tc<48, 48>(cx, t.p, M, w1, b1, w2, b2, rep ? t.p : nullptr, y.p, act);
        else if (c.cin == 48) launch_mlp_tc<48, 96>(cx, t.p, M, w1, b1, w2, b2, nullptr, y.p, act);
        else if (c.cin == 96) launch_mlp_tc<96, 96>(cx, t.p, M, w1, b1, w2, b2, rep ? t.p : nullptr, y.p, act);
        else launch_mlp_big<192, 192, 3, 1>(cx, w, name, t.p, M, w1, b1, w2, b2, rep ? t.p : nullptr, y.p, act);
        O::release(cx, t);
        return y;
      }
    }
    Act u = O::make(cx, x.n, OH, OW, 2 * c.cin);
    // fp16/tcgen05 mode: erf-GELU through the 11-instruction sigmoid form (common.cuh gelu_fast, |err| <= 3.4e-6)
    int gelu = ACT_GELU;
    if constexpr (std::is_same<T, __half>::value) { if (cx.use_tc && !env_is("RDB_GELU", "exact")) gelu = ACT_GELUF; }
    O::pw(cx, t, w.get(name + "pw1.w"), &w.get(name + "pw1.b"), gelu, nullptr, u);
    Act y = O::make(cx, x.n, OH, OW, c.cout);
    O::pw(cx, u, w.get(name + "pw2.w"), &w.get(name + "pw2.b"), ACT_NONE, rep ? t.p : nullptr, y);
    O::release(cx, u);
    O::release(cx, t);
    return y;
  }
};

}  // namespace rdb

#include "stem_tc.cuh"
