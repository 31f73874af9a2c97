// PP-OCRv6-small text recogniser: PPLCNetV4-rec backbone + LightSVTR neck + CTC head with
// fused greedy decode.  Reference network: rapid_doc/model/ocr/ppocrv6_pytorch/modeling/
// backbones/rec_lcnetv4.py:26-43,306-311, necks/rnn.py:321-379, heads/rec_multi_head.py:66-77;
// engine seam rapid_doc/model/ocr/torch.py:171-192; decode rapid_ocr.py:443-449 (CTCLabelDecode).
#include "rec.h"

namespace rdb {

static const BlockCfg kRecBlocks[4][7] = {
    {{96, 96, 1, 1, 1}},
    {{96, 96, 1, 1, 0}, {96, 96, 1, 1, 0}},
    {{96, 192, 2, 1, 0}, {192, 192, 1, 1, 1}, {192, 192, 1, 1, 0}, {192, 192, 1, 1, 1}, {192, 192, 1, 1, 0}, {192, 192, 1, 1, 1}, {192, 192, 1, 1, 0}},
    {{192, 384, 2, 1, 0}, {384, 384, 1, 1, 1}, {384, 384, 1, 1, 0}},
};
static const int kRecBlockCount[4] = {1, 2, 7, 3};

RecEngine::RecEngine(const void* blob, size_t nbytes, int device, int precision) : device_(device), precision_(precision) {
  RDB_CUDA(cudaSetDevice(device));
  weights_.reset(new Weights(blob, nbytes));
  RDB_CHECK(weights_->has("ctc.w") && weights_->has("svtr.norm.g"), "blob is not a rec model");
  vocab_ = weights_->get("ctc.w").shape[0];
  RDB_CUDA(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, device));
}

RecEngine::~RecEngine() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  pools_[0].release_all();
  pools_[1].release_all();
  if (ev_fork_) {
    cudaEventDestroy(ev_fork_);
    for (int s = 0; s < 2; ++s) { cudaStreamDestroy(lane_[s]); cudaEventDestroy(ev_join_[s]); }
  }
}

template <typename T>
void RecEngine::forward_chunk(Ctx& cx, const RecInput& in, int n, int W, int32_t* ids, float* probs, int32_t* text_ids,
                              int32_t* text_len, float* conf, float* softmax) {
  using O = Ops<T>;
  using Act = typename O::Act;
  const Weights& w = *weights_;
  const int H = 48, H1 = 24, W1 = (W - 1) / 2 + 1;
  bool wide = false;
  if constexpr (std::is_same<T, __half>::value) wide = cx.use_tc && !env_is("RDB_CONV", "simt") && !env_is("RDB_TC_WIDE", "0");
  const int e1_wp = wide ? W1 + 1 : W1;
  Act e1 = O::make(cx, n, H1, e1_wp, 48);
  e1.w = W1; e1.wp = wide ? e1_wp : 0;
  bool stem1_tc = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_STEM1", "simt")) {
      stem1_tc = true;
      if (in.f32 != nullptr) {
        InF32NCHW src{in.f32, H, W};
        launch_stem1_tc<InF32NCHW, 48>(cx, src, n, w.get("stem1.w"), w.get("stem1.b"), e1.p, H1, W1, e1_wp);
      } else {
        InU8HWC src{in.u8, H, W, 1, {0, 0, 0}, {1, 1, 1}, in.valid_w};
        launch_stem1_tc<InU8HWC, 48>(cx, src, n, w.get("stem1.w"), w.get("stem1.b"), e1.p, H1, W1, e1_wp);
      }
    }
  }
  if (!stem1_tc) {
    long long total = (long long)n * H1 * e1_wp;
    cx.begin("stem1");
    if (in.f32 != nullptr) {
      InF32NCHW src{in.f32, H, W};
      stem1_kernel<T, InF32NCHW, 48><<<cdiv(total, 128), 128, 0, cx.st>>>(src, n, w.get("stem1.w").d, w.get("stem1.b").d, e1.p, H1, W1, e1_wp);
    } else {
      InU8HWC src{in.u8, H, W, 1, {0, 0, 0}, {1, 1, 1}, in.valid_w};
      stem1_kernel<T, InU8HWC, 48><<<cdiv(total, 128), 128, 0, cx.st>>>(src, n, w.get("stem1.w").d, w.get("stem1.b").d, e1.p, H1, W1, e1_wp);
    }
    cx.end();
  }
  Act x = Backbone<T>::template stem_rest<48>(cx, w, e1);
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < kRecBlockCount[s]; ++b) {
      std::string name = "s" + std::to_string(s) + ".b" + std::to_string(b) + ".";
      Act y = Backbone<T>::block(cx, w, name, kRecBlocks[s][b], x, false);
      x = y;
    }
  RDB_CHECK(x.h == 3 && x.c == 384, "rec: unexpected backbone output shape");
  const int Tn = x.w / 2;
  Act tok = O::make(cx, n, 1, Tn, 384);
  {
    long long total = tok.pixels() * 48;
    cx.begin("avgpool3x2");
    avgpool3x2_kernel<T><<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(x.p, n, x.w, 384, tok.p);
    cx.end();
  }
  O::release(cx, x);
  // ---- LightSVTR (rnn.py:361-379)
  const long long M = (long long)n * Tn;
  Act res0 = O::make(cx, n, 1, Tn, 120);
  O::pw(cx, tok, w.get("svtr.c0.w"), &w.get("svtr.c0.b"), ACT_SILU, nullptr, res0);
  Act h = O::make(cx, n, 1, Tn, 120);
  O::pw(cx, tok, w.get("svtr.c1.w"), &w.get("svtr.c1.b"), ACT_SILU, nullptr, h);
  O::release(cx, tok);
  Act t = O::make(cx, n, 1, Tn, 120);
  O::template dwconv<1, 7, ACT_SILU, true>(cx, h, 1, 1, w.get("svtr.dw.w"), w.get("svtr.dw.b"), t);
  O::release(cx, h);
  auto ln = [&](const Act& a, const Tensor& g, const Tensor& b, const T* res, Act& o) {
    cx.begin("layernorm");
    layernorm_kernel<T><<<cdiv(M, 8), 256, 0, cx.st>>>(a.p, M, 120, g.d, b.d, 1e-6f, res, o.p);
    cx.end();
  };
  for (int i = 0; i < 2; ++i) {
    std::string p = "svtr.blk" + std::to_string(i) + ".";
    Act y = O::make(cx, n, 1, Tn, 120);
    ln(t, w.get(p + "ln1.g"), w.get(p + "ln1.b"), nullptr, y);
    Act qkv = O::make(cx, n, 1, Tn, 360);
    O::pw(cx, y, w.get(p + "qkv.w"), &w.get(p + "qkv.b"), ACT_NONE, nullptr, qkv);
    Act att = O::make(cx, n, 1, Tn, 120);
    {
      size_t sm = (size_t)2 * Tn * 15 * sizeof(float);
      auto k = attention_kernel<T, 15>;
      set_smem(k, sm);
      int threads = Tn < 128 ? ((Tn + 31) / 32) * 32 : 128;
      cx.begin("attention");
      k<<<n * 8, threads, sm, cx.st>>>(qkv.p, Tn, 8, 0.2581988897471611f /* 15^-0.5 */, att.p);
      cx.end();
    }
    O::release(cx, qkv);
    Act t2 = O::make(cx, n, 1, Tn, 120);
    O::pw(cx, att, w.get(p + "proj.w"), &w.get(p + "proj.b"), ACT_NONE, t.p, t2);
    O::release(cx, att);
    O::release(cx, t);
    ln(t2, w.get(p + "ln2.g"), w.get(p + "ln2.b"), nullptr, y);
    Act m1 = O::make(cx, n, 1, Tn, 240);
    O::pw(cx, y, w.get(p + "fc1.w"), &w.get(p + "fc1.b"), ACT_SILU, nullptr, m1);
    O::release(cx, y);
    Act t3 = O::make(cx, n, 1, Tn, 120);
    O::pw(cx, m1, w.get(p + "fc2.w"), &w.get(p + "fc2.b"), ACT_NONE, t2.p, t3);
    O::release(cx, m1);
    O::release(cx, t2);
    t = t3;
  }
  Act seq = O::make(cx, n, 1, Tn, 120);
  ln(t, w.get("svtr.norm.g"), w.get("svtr.norm.b"), res0.p, seq);
  O::release(cx, t);
  O::release(cx, res0);
  // ---- CTC head + fused greedy decode
  const Tensor& cw = w.get("ctc.w");
  const Tensor& cb = w.get("ctc.b");
  const int V = cw.shape[0];
  if (softmax != nullptr) {  // compat: materialise logits, softmax rows (torch.py:186-187)
    float* logits = cx.pool->alloc_t<float>((size_t)M * V);
    GemmArgs g{};
    g.A = seq.p; g.lda = 120; g.W = cw.d; g.bias = cb.d; g.out = logits; g.ldc = V; g.M = (int)M; g.N = V; g.K = 120; g.act = ACT_NONE;
    cx.begin("ctc_logits_gemm");
    launch_gemm_simt<T, float>(g, cx.st);
    cx.end();
    cx.begin("softmax_rows");
    softmax_rows_kernel<<<(unsigned)M, 256, 0, cx.st>>>(logits, V, softmax);
    cx.end();
    cx.pool->free(logits);
  }
  {
    int tiles = cdiv(V, SG_BN);
    const int cap = tiles < 320 ? 320 : tiles;  // room for the tcgen05 kernel's (n-tile x sub-warp) partials
    float* pmax = cx.pool->alloc_t<float>((size_t)M * cap);
    float* psum = cx.pool->alloc_t<float>((size_t)M * cap);
    int* pidx = cx.pool->alloc_t<int>((size_t)M * cap);
    bool done = false;
    if constexpr (std::is_same<T, __half>::value) {
      if (cx.use_tc) {
        launch_gemm_tc_ctc(cx, seq.p, 120, M, 120, cw.h, V, cb.d, pmax, pidx, psum, &tiles);
        done = true;
      }
    }
    if (!done) {
      GemmArgs g{};
      g.A = seq.p; g.lda = 120; g.W = cw.d; g.bias = cb.d; g.M = (int)M; g.N = V; g.K = 120;
      g.pmax = pmax; g.pidx = pidx; g.psum = psum;
      cx.begin("ctc_head_gemm_argmax");
      launch_gemm_simt_ctc<T>(g, cx.st);
      cx.end();
    }
    cx.begin("ctc_merge");
    ctc_merge_kernel<<<cdiv(M, 128), 128, 0, cx.st>>>(pmax, pidx, psum, (int)M, tiles, ids, probs);
    cx.end();
    cx.begin("ctc_collapse");
    ctc_collapse_kernel<<<cdiv(n, 4), 128, 0, cx.st>>>(ids, probs, n, Tn, text_ids, text_len, conf);
    cx.end();
    cx.pool->free(pmax); cx.pool->free(psum); cx.pool->free(pidx);
  }
  O::release(cx, seq);
}

void RecEngine::infer_impl(const RecInput& in0, int n, int W, const RecOutput& out, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device_));
  RDB_CHECK(n > 0 && W >= 16, "rec: width must be >= 16");
  RDB_CHECK((in0.f32 != nullptr) != (in0.u8 != nullptr), "rec: exactly one input");
  Ctx cx;
  cx.st = st; cx.pool = &pools_[0]; cx.precision = precision_;
  cx.use_tc = (precision_ == 1) && !env_gemm_simt();
  cx.num_sms = num_sms_;
  const int Tn = tokens_for_width(W);
  const void* src = in0.f32 ? (const void*)in0.f32 : (const void*)in0.u8;
  const size_t crop_in = (size_t)3 * 48 * W * (in0.f32 ? sizeof(float) : 1);
  const bool in_dev = is_device_ptr(src);
  int chunk = chunk_crops_;
  if (chunk > n) chunk = n;
  // a batch that fits one chunk is still cut in two when it is large enough: the two compute lanes overlap each other's
  // small kernels (SE FCs, LightSVTR) — the recogniser of a window is called with one <= rec_batch_num batch at a time
  if (chunk == n && (long long)n * W >= 64ll * 320 && n >= 8 && !env_is("RDB_REC_SPLIT", "0")) chunk = (n + 1) / 2;
  // two compute lanes (stream + pool each), chunks alternate (see DetEngine::infer)
  const int n_chunks = (n + chunk - 1) / chunk;
  const int lanes = (n_chunks >= 2 && !env_is("RDB_LANES", "1")) ? 2 : 1;
  ensure_streams();
  int32_t* d_vw = nullptr;
  if (in0.valid_w) {
    if (is_device_ptr(in0.valid_w)) d_vw = const_cast<int32_t*>(in0.valid_w);
    else {
      d_vw = pools_[0].alloc_t<int32_t>(n);
      RDB_CUDA(cudaMemcpyAsync(d_vw, in0.valid_w, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    }
  }
  RDB_CUDA(cudaEventRecord(ev_fork_, st));
  Ctx cxs[2];
  void* d_in[2] = {nullptr, nullptr};
  int32_t *d_ids[2], *d_tids[2], *d_tlen[2];
  float *d_probs[2], *d_conf[2], *d_sm[2] = {nullptr, nullptr};
  const bool sm_dev = out.softmax ? is_device_ptr(out.softmax) : true;
  for (int l = 0; l < lanes; ++l) {
    cxs[l] = cx; cxs[l].st = lane_[l]; cxs[l].pool = &pools_[l]; cxs[l].launches = 0;
    RDB_CUDA(cudaStreamWaitEvent(lane_[l], ev_fork_, 0));
    if (!in_dev) d_in[l] = pools_[l].alloc(crop_in * chunk);
    d_ids[l] = pools_[l].alloc_t<int32_t>((size_t)chunk * Tn);
    d_probs[l] = pools_[l].alloc_t<float>((size_t)chunk * Tn);
    d_tids[l] = pools_[l].alloc_t<int32_t>((size_t)chunk * Tn);
    d_tlen[l] = pools_[l].alloc_t<int32_t>(chunk);
    d_conf[l] = pools_[l].alloc_t<float>(chunk);
    if (out.softmax && !sm_dev) d_sm[l] = pools_[l].alloc_t<float>((size_t)chunk * Tn * vocab_);
  }
  bool any_host = !in_dev;
  int it = 0;
  for (int i0 = 0; i0 < n; i0 += chunk, ++it) {
    const int l = it % lanes;
    cudaStream_t ls = lane_[l];
    int m = (n - i0 < chunk) ? (n - i0) : chunk;
    const uint8_t* src_i = static_cast<const uint8_t*>(src) + (size_t)i0 * crop_in;
    const void* dsrc = src_i;
    if (!in_dev) {
      RDB_CUDA(cudaMemcpyAsync(d_in[l], src_i, crop_in * m, cudaMemcpyHostToDevice, ls));
      dsrc = d_in[l];
    }
    RecInput in = in0;
    if (in.f32) in.f32 = static_cast<const float*>(dsrc); else in.u8 = static_cast<const uint8_t*>(dsrc);
    in.valid_w = d_vw ? d_vw + i0 : nullptr;
    float* smx = out.softmax ? (sm_dev ? out.softmax + (size_t)i0 * Tn * vocab_ : d_sm[l]) : nullptr;
    if (precision_ == 0) forward_chunk<float>(cxs[l], in, m, W, d_ids[l], d_probs[l], d_tids[l], d_tlen[l], d_conf[l], smx);
    else forward_chunk<__half>(cxs[l], in, m, W, d_ids[l], d_probs[l], d_tids[l], d_tlen[l], d_conf[l], smx);
    auto emit = [&](void* dst, const void* dsrc2, size_t bytes) {
      if (!dst) return;
      if (is_device_ptr(dst)) RDB_CUDA(cudaMemcpyAsync(dst, dsrc2, bytes, cudaMemcpyDeviceToDevice, ls));
      else { RDB_CUDA(cudaMemcpyAsync(dst, dsrc2, bytes, cudaMemcpyDeviceToHost, ls)); any_host = true; }
    };
    emit(out.ids ? out.ids + (size_t)i0 * Tn : nullptr, d_ids[l], (size_t)m * Tn * 4);
    emit(out.probs ? out.probs + (size_t)i0 * Tn : nullptr, d_probs[l], (size_t)m * Tn * 4);
    emit(out.text_ids ? out.text_ids + (size_t)i0 * Tn : nullptr, d_tids[l], (size_t)m * Tn * 4);
    emit(out.text_len ? out.text_len + i0 : nullptr, d_tlen[l], (size_t)m * 4);
    emit(out.conf ? out.conf + i0 : nullptr, d_conf[l], (size_t)m * 4);
    if (out.softmax && !sm_dev) emit(out.softmax + (size_t)i0 * Tn * vocab_, d_sm[l], (size_t)m * Tn * vocab_ * 4);
  }
  long long launches = 0;
  for (int l = 0; l < lanes; ++l) {
    RDB_CUDA(cudaEventRecord(ev_join_[l], lane_[l]));
    RDB_CUDA(cudaStreamWaitEvent(st, ev_join_[l], 0));
    launches += cxs[l].launches;
  }
  if (any_host) RDB_CUDA(cudaStreamSynchronize(st));
  for (int l = 0; l < lanes; ++l) {
    if (d_in[l]) pools_[l].free(d_in[l]);
    pools_[l].free(d_ids[l]); pools_[l].free(d_probs[l]); pools_[l].free(d_tids[l]); pools_[l].free(d_tlen[l]); pools_[l].free(d_conf[l]);
    if (d_sm[l]) pools_[l].free(d_sm[l]);
  }
  if (d_vw && d_vw != in0.valid_w) pools_[0].free(d_vw);
  if (Profiler::global().on) { RDB_CUDA(cudaDeviceSynchronize()); Profiler::global().resolve(); }
  last_launches_ = launches;
}

void RecEngine::infer(const RecInput& in, int n, int W, const RecOutput& out, cudaStream_t st) {
  try {
    infer_impl(in, n, W, out, st);
  } catch (...) {
    cudaDeviceSynchronize();
    for (auto& p : pools_) p.reclaim();
    throw;
  }
  for (auto& p : pools_) p.enforce_cap();
}

void RecEngine::ensure_streams() {
  if (ev_fork_) return;
  RDB_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  for (int s = 0; s < 2; ++s) {
    RDB_CUDA(cudaStreamCreateWithFlags(&lane_[s], cudaStreamNonBlocking));
    RDB_CUDA(cudaEventCreateWithFlags(&ev_join_[s], cudaEventDisableTiming));
  }
}

}  // namespace rdb
