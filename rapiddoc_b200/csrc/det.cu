// PP-OCRv6-small DBNet text detector forward + DB binarise/dilate on one B200.
// Reference network: rapid_doc/model/ocr/ppocrv6_pytorch/modeling/backbones/rec_lcnetv4.py:7-23,
// necks/db_fpn.py:366-415, heads/det_db_head.py:103-147; engine seam rapid_doc/model/ocr/torch.py:171-184.
#include "det.h"

namespace rdb {

static const BlockCfg kDetBlocks[4][5] = {
    {{48, 48, 1, 1, 1}, {48, 48, 1, 1, 0}},
    {{48, 96, 2, 2, 0}, {96, 96, 1, 1, 1}, {96, 96, 1, 1, 0}},
    {{96, 192, 2, 2, 0}, {192, 192, 1, 1, 1}, {192, 192, 1, 1, 0}, {192, 192, 1, 1, 1}, {192, 192, 1, 1, 0}},
    {{192, 384, 2, 2, 0}, {384, 384, 1, 1, 1}, {384, 384, 1, 1, 0}},
};
static const int kDetBlockCount[4] = {2, 3, 5, 3};

DetEngine::DetEngine(const void* blob, size_t nbytes, int device, int precision) : device_(device), precision_(precision) {
  RDB_CUDA(cudaSetDevice(device));
  weights_.reset(new Weights(blob, nbytes));
  RDB_CHECK(weights_->has("head.final.w") && weights_->has("neck.lk3.pw.w"), "blob is not a det model");
  RDB_CUDA(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, device));
}

DetEngine::~DetEngine() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (int s = 0; s < kMaxLanes; ++s) pools_[s].release_all();
  if (copy_in_) {
    cudaStreamDestroy(copy_in_); cudaStreamDestroy(copy_out_);
    cudaEventDestroy(ev_fork_);
    for (int s = 0; s < kMaxLanes; ++s) { cudaStreamDestroy(lane_[s]); cudaEventDestroy(ev_join_[s]); }
    for (int s = 0; s < kSlots; ++s) { cudaEventDestroy(ev_in_[s]); cudaEventDestroy(ev_out_[s]); cudaEventDestroy(ev_compute_[s]); }
  }
}

void DetEngine::ensure_streams() {
  if (copy_in_) return;
  RDB_CUDA(cudaStreamCreateWithFlags(&copy_in_, cudaStreamNonBlocking));
  RDB_CUDA(cudaStreamCreateWithFlags(&copy_out_, cudaStreamNonBlocking));
  RDB_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  for (int s = 0; s < kMaxLanes; ++s) {
    RDB_CUDA(cudaStreamCreateWithFlags(&lane_[s], cudaStreamNonBlocking));
    RDB_CUDA(cudaEventCreateWithFlags(&ev_join_[s], cudaEventDisableTiming));
  }
  for (int s = 0; s < kSlots; ++s) {
    RDB_CUDA(cudaEventCreateWithFlags(&ev_in_[s], cudaEventDisableTiming));
    RDB_CUDA(cudaEventCreateWithFlags(&ev_out_[s], cudaEventDisableTiming));
    RDB_CUDA(cudaEventCreateWithFlags(&ev_compute_[s], cudaEventDisableTiming));
  }
}

static void linear_coeffs(int dsize, int ssize, bool vertical, std::vector<int>& idx, std::vector<short>& ab);

template <typename T>
void DetEngine::forward_chunk(Ctx& cx, const DetInput& in, int n, int H, int W, float thresh, bool dilate, float* prob,
                              uint8_t* bitmap) {
  using O = Ops<T>;
  using Act = typename O::Act;
  const Weights& w = *weights_;
  const int H1 = H / 2, W1 = W / 2;
  // ---- stem1
  bool wide = false;   // fp16/tcgen05 path: conv inputs carry one (2x2) or two (3x3) zero pad pixels per row (wide-row TMA)
  if constexpr (std::is_same<T, __half>::value) wide = cx.use_tc && !env_is("RDB_CONV", "simt") && !env_is("RDB_TC_WIDE", "0");
  const int e1_wp = wide ? W1 + 1 : W1;
  Act e1 = O::make(cx, n, H1, e1_wp, 24);
  e1.w = W1; e1.wp = wide ? e1_wp : 0;
  bool stem1_tc = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_STEM1", "simt")) {
      stem1_tc = true;
      if (in.f32 != nullptr) {
        InF32NCHW src{in.f32, H, W};
        launch_stem1_tc<InF32NCHW, 24>(cx, src, n, w.get("stem1.w"), w.get("stem1.b"), e1.p, H1, W1, e1_wp);
      } else {
        InU8HWC src{in.u8, H, W, 0, {in.mean[0], in.mean[1], in.mean[2]}, {in.stdv[0], in.stdv[1], in.stdv[2]}, nullptr};
        launch_stem1_tc<InU8HWC, 24>(cx, src, n, w.get("stem1.w"), w.get("stem1.b"), e1.p, H1, W1, e1_wp);
      }
    }
  }
  if (!stem1_tc) {
    long long total = (long long)n * H1 * e1_wp;
    cx.begin("stem1");
    if (in.f32 != nullptr) {
      InF32NCHW src{in.f32, H, W};
      stem1_kernel<T, InF32NCHW, 24><<<cdiv(total, 128), 128, 0, cx.st>>>(src, n, w.get("stem1.w").d, w.get("stem1.b").d, e1.p, H1, W1, e1_wp);
    } else {
      InU8HWC src{in.u8, H, W, 0, {in.mean[0], in.mean[1], in.mean[2]}, {in.stdv[0], in.stdv[1], in.stdv[2]}, nullptr};
      stem1_kernel<T, InU8HWC, 24><<<cdiv(total, 128), 128, 0, cx.st>>>(src, n, w.get("stem1.w").d, w.get("stem1.b").d, e1.p, H1, W1, e1_wp);
    }
    cx.end();
  }
  Act x = Backbone<T>::template stem_rest<24>(cx, w, e1);
  // ---- stages
  Act feats[4];
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < kDetBlockCount[s]; ++b) {
      bool keep = (b == 0 && s > 0);  // the previous stage output also feeds the neck
      std::string name = "s" + std::to_string(s) + ".b" + std::to_string(b) + ".";
      Act y = Backbone<T>::block(cx, w, name, kDetBlocks[s][b], x, keep);
      x = y;
    }
    feats[s] = x;
  }
  // ---- RepLKFPN
  Act f[4];
  bool neck_fused = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_NECK", "unfused")) {
      // fused input stage: the SE gate comes from the pooled backbone feature (mean(conv(x)) = W mean(x)), and ONE
      // tcgen05 GEMM per level applies the 1x1 conv, the (1+gate) column scale and the top-down nearest-up add.
      neck_fused = true;
      for (int i = 3; i >= 0; --i) {
        std::string p = "neck.in" + std::to_string(i) + ".";
        float* gate = O::se_gate(cx, feats[i], w.get(p + "se.w1"), w.get(p + "se.b1"), w.get(p + "se.w2"), w.get(p + "se.b2"), 1, &w.get(p + "w"));
        f[i] = O::make(cx, n, feats[i].h, feats[i].w, 96);
        TcFuse fu;
        fu.colscale = gate; fu.rows_per_img = feats[i].h * feats[i].w;
        if (i < 3) { fu.up_res = f[i + 1].p; fu.up_H = feats[i].h; fu.up_W = feats[i].w; }
        launch_gemm_tc(cx, feats[i].p, feats[i].c, feats[i].pixels(), feats[i].c, w.get(p + "w").h, 96, nullptr, ACT_NONE, nullptr, 96, f[i].p, 96, 0, &fu);
        O::release(cx, feats[i]);
        cx.pool->free(gate);
      }
    }
  }
  if (!neck_fused) {
    for (int i = 0; i < 4; ++i) {
      std::string p = "neck.in" + std::to_string(i) + ".";
      f[i] = O::make(cx, n, feats[i].h, feats[i].w, 96);
      O::pw(cx, feats[i], w.get(p + "w"), nullptr, ACT_NONE, nullptr, f[i]);
      O::release(cx, feats[i]);
      float* gate = O::se_gate(cx, f[i], w.get(p + "se.w1"), w.get(p + "se.b1"), w.get(p + "se.w2"), w.get(p + "se.b2"), 1);
      O::scale(cx, f[i], gate);  // x + x*g = x*(1+g)
      cx.pool->free(gate);
    }
    for (int i = 2; i >= 0; --i) {
      long long total = f[i].pixels() * 12;
      cx.begin("neck_upadd");
      upsample2_add_kernel<T><<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(f[i].p, f[i + 1].p, n, f[i].h, f[i].w, 96);
      cx.end();
    }
  }
  Act pq[4];
  NeckSrc<T> ns;
  float* gates[4];
  for (int i = 0; i < 4; ++i) {
    std::string p = "neck.lk" + std::to_string(i) + ".";
    Act d = O::make(cx, n, f[i].h, f[i].w, 96);
    O::template dwconv<7, 7, ACT_NONE, false>(cx, f[i], 1, 1, w.get(p + "dw.w"), w.get(p + "dw.b"), d);
    O::release(cx, f[i]);
    pq[i] = O::make(cx, n, d.h, d.w, 24);
    O::pw(cx, d, w.get(p + "pw.w"), nullptr, ACT_NONE, nullptr, pq[i]);
    O::release(cx, d);
    gates[i] = O::se_gate(cx, pq[i], w.get(p + "se.w1"), w.get(p + "se.b1"), w.get(p + "se.w2"), w.get(p + "se.b2"), 1);
    ns.f[i] = pq[i].p;
    ns.gate[i] = gates[i];
  }
  uint8_t* seg = nullptr;
  if (bitmap != nullptr) seg = dilate ? cx.pool->alloc_t<uint8_t>((size_t)n * H * W) : bitmap;
  bool head_fused = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_HEAD", "unfused") && !env_is("RDB_HEAD", "simt") && !env_is("RDB_CONV", "simt")) {
      // concat + conv3x3 + both transposed convs + sigmoid + threshold in one kernel (head_planar.cuh)
      const __half* fp[4]; const float* gp[4];
      for (int i = 0; i < 4; ++i) { fp[i] = pq[i].p; gp[i] = gates[i]; }
      launch_head_planar(cx, w, fp, gp, n, pq[0].h, pq[0].w, thresh, prob, seg);
      for (int i = 0; i < 4; ++i) { O::release(cx, pq[i]); cx.pool->free(gates[i]); }
      head_fused = true;
    }
  }
  if (!head_fused) {
  const int neck_wp = wide ? pq[0].w + 2 : pq[0].w;
  Act neck = O::make(cx, n, pq[0].h, neck_wp, 96);
  neck.w = pq[0].w; neck.wp = wide ? neck_wp : 0;
  {
    long long total = neck.pixels() * 12;
    cx.begin("neck_concat");
    neck_concat_kernel<T><<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(ns, n, neck.h, neck.w, neck.p, neck_wp, wide ? 1 : 0);
    cx.end();
    if (wide) {
      const long long rows = (long long)n * neck.h;
      zero_cols_kernel<T><<<cdiv(rows * 12, kThreads), kThreads, 0, cx.st>>>(neck.p, rows, neck_wp, 96, 0, 1);
      zero_cols_kernel<T><<<cdiv(rows * 12, kThreads), kThreads, 0, cx.st>>>(neck.p, rows, neck_wp, 96, neck_wp - 1, 1);
      RDB_LAUNCH_CHECK();
      cx.launches += 2;
    }
  }
  for (int i = 0; i < 4; ++i) { O::release(cx, pq[i]); cx.pool->free(gates[i]); }
  // ---- DBHead
  RDB_CHECK(!wide || cx.use_tc, "wide rows need the tcgen05 conv");
  Act hd = O::make(cx, n, neck.h, neck.w, 24);
  bool head_tc = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_CONV", "simt")) {
      launch_conv_tc(cx, "head_conv3x3", neck.p, n, neck.h, neck.w, 96, w.get("head.down.w").h, 24, w.get("head.down.b").d, ACT_RELU, 3, 3, 1, 1,
                     1, 1, hd.p, hd.h, hd.w, 24, 0, wide ? neck_wp : 0, 0);
      head_tc = true;
    }
  }
  if (!head_tc) {
    auto k = conv_direct_kernel<T, 3, 3, 1, 1, 96, 24, 24, ACT_RELU>;
    size_t sm = (size_t)(9 * 96 * 24 + 24) * sizeof(float);
    set_smem(k, sm);
    cx.begin("head_conv3x3");
    k<<<dim3(cdiv(hd.pixels(), 128), 1), 128, sm, cx.st>>>(neck.p, n, neck.h, neck.w, 1, 1, w.get("head.down.w").d,
                                                          w.get("head.down.b").d, hd.p, hd.h, hd.w, 24, 0);
    cx.end();
  }
  O::release(cx, neck);
  bool tail_tc = false;
  if constexpr (std::is_same<T, __half>::value) {
    if (cx.use_tc && !env_is("RDB_HEAD", "simt")) {
      launch_head_tail_tc(cx, hd.p, n, hd.h, hd.w, w.get("head.up.w").h, w.get("head.up.b").d, w.get("head.final.w").d,
                          w.get("head.final.b").d, thresh, prob, seg);
      tail_tc = true;
    }
  }
  if (!tail_tc) {
    long long total = (long long)n * (2 * hd.h) * (2 * hd.w);
    cx.begin("head_tail");
    head_tail_kernel<T><<<cdiv(total, 128), 128, 0, cx.st>>>(hd.p, n, hd.h, hd.w, w.get("head.up.w").d, w.get("head.up.b").d,
                                                            w.get("head.final.w").d, w.get("head.final.b").d, thresh, prob, seg);
    cx.end();
  }
  O::release(cx, hd);
  }   // !head_fused
  if (bitmap != nullptr && dilate) {
    long long total = (long long)n * H * (W / 4);
    cx.begin("db_dilate");
    dilate2x2_kernel<<<cdiv(total, kThreads), kThreads, 0, cx.st>>>(seg, n, H, W, bitmap);
    cx.end();
    cx.pool->free(seg);
  }
}

void DetEngine::infer_impl(const DetInput& in_host_or_dev, int n, int H, int W, float thresh, bool dilate, float* prob, uint8_t* bitmap,
                      cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device_));
  RDB_CHECK(n > 0 && H > 0 && W > 0 && H % 32 == 0 && W % 32 == 0, "det: h and w must be positive multiples of 32");
  RDB_CHECK((in_host_or_dev.f32 != nullptr) != (in_host_or_dev.u8 != nullptr), "det: exactly one input");
  Ctx cx;
  cx.st = st; cx.pool = &pools_[0]; cx.precision = precision_;
  cx.use_tc = (precision_ == 1) && !env_gemm_simt();
  cx.num_sms = num_sms_;
  const void* src = in_host_or_dev.f32 ? (const void*)in_host_or_dev.f32 : (const void*)in_host_or_dev.u8;
  const size_t in_elem = in_host_or_dev.f32 ? sizeof(float) : 1;
  const bool do_resize = in_host_or_dev.u8 != nullptr && in_host_or_dev.src_h > 0 && (in_host_or_dev.src_h != H || in_host_or_dev.src_w != W);
  const int SH = do_resize ? in_host_or_dev.src_h : H, SW = do_resize ? in_host_or_dev.src_w : W;
  const size_t page_in = (size_t)3 * SH * SW * in_elem;
  const size_t page_px = (size_t)H * W;
  const bool in_dev = is_device_ptr(src);
  const bool prob_dev = prob ? is_device_ptr(prob) : true;
  const bool bm_dev = bitmap ? is_device_ptr(bitmap) : true;
  // chunk so that one chunk's activations stay a bounded working set
  // device-resident buffers: no copies to overlap, so larger chunks (fewer launches, longer persistent kernels) win;
  // host buffers keep the smaller chunk so H2D / D2H of neighbouring chunks hide behind compute
  const bool all_dev = is_device_ptr(src) && (prob == nullptr || is_device_ptr(prob)) && (bitmap == nullptr || is_device_ptr(bitmap));
  long long px_budget = (all_dev && !chunk_pixels_set_) ? 2 * chunk_pixels_ : chunk_pixels_;
  int chunk = (int)(px_budget / (long long)(H * (long long)W));
  if (chunk < 1) chunk = 1;
  if (chunk > n) chunk = n;
  // Two compute LANES (stream + buffer pool each): consecutive chunks alternate lanes so the many tiny
  // kernels of one chunk (SE FCs, stage-4 layers) overlap the other chunk's work, and — with host buffers —
  // the H2D of chunk i+1 and the D2H of chunk i-1 (dedicated copy streams) overlap the compute of chunk i.
  // With device buffers the call stays asynchronous: lanes fork from `st` and join back into it by events.
  const bool any_host = !in_dev || (prob && !prob_dev) || (bitmap && !bm_dev);
  const int n_chunks = (n + chunk - 1) / chunk;
  int lanes = 2;
  if (const char* e = sw_get("RDB_LANES")) lanes = std::atoi(e);
  if (lanes > kMaxLanes) lanes = kMaxLanes;
  if (lanes > n_chunks) lanes = n_chunks;
  if (lanes < 1) lanes = 1;
  ensure_streams();
  Ctx cxs[kMaxLanes];
  // host staging: kSlots buffers (two per lane), so a chunk's H2D / D2H never stalls the lane that owns the slot
  const int slots = lanes * 2 <= n_chunks ? lanes * 2 : lanes;
  void* d_in[kSlots] = {};
  float* d_prob[kSlots] = {};
  uint8_t* d_bm[kSlots] = {};
  // GPU resize (DetPreProcess): coefficient tables once per call
  int *d_xi = nullptr, *d_yi = nullptr; short *d_xa = nullptr, *d_ya = nullptr;
  uint8_t* d_rs[kMaxLanes] = {};
  if (do_resize) {
    std::vector<int> xi, yi; std::vector<short> xa, ya;
    linear_coeffs(W, SW, false, xi, xa);
    linear_coeffs(H, SH, true, yi, ya);
    d_xi = pools_[0].alloc_t<int>(W); d_yi = pools_[0].alloc_t<int>(H);
    d_xa = pools_[0].alloc_t<short>(2 * (size_t)W); d_ya = pools_[0].alloc_t<short>(2 * (size_t)H);
    RDB_CUDA(cudaMemcpyAsync(d_xi, xi.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
    RDB_CUDA(cudaMemcpyAsync(d_yi, yi.data(), (size_t)H * 4, cudaMemcpyHostToDevice, st));
    RDB_CUDA(cudaMemcpyAsync(d_xa, xa.data(), (size_t)W * 4, cudaMemcpyHostToDevice, st));
    RDB_CUDA(cudaMemcpyAsync(d_ya, ya.data(), (size_t)H * 4, cudaMemcpyHostToDevice, st));
    RDB_CUDA(cudaStreamSynchronize(st));   // the host vectors go out of scope
  }
  RDB_CUDA(cudaEventRecord(ev_fork_, st));
  for (int l = 0; l < lanes; ++l) {
    if (do_resize) d_rs[l] = pools_[l].alloc_t<uint8_t>((size_t)chunk * H * W * 3);
    cxs[l] = cx;
    cxs[l].st = lane_[l];
    cxs[l].pool = &pools_[l];
    cxs[l].launches = 0;
    RDB_CUDA(cudaStreamWaitEvent(lane_[l], ev_fork_, 0));
  }
  for (int s = 0; s < slots; ++s) {
    Pool& pl = pools_[s % lanes];
    if (!in_dev) d_in[s] = pl.alloc(page_in * chunk);
    if (!(prob_dev && prob)) d_prob[s] = pl.alloc_t<float>(page_px * chunk);
    if (bitmap && !bm_dev) d_bm[s] = pl.alloc_t<uint8_t>(page_px * chunk);
  }
  // chunk schedule: uniform, except that with host buffers the first and last chunks are ramped (c/4, c/2, c ... c/2, c/4)
  // so the un-overlappable first H2D and last D2H are short
  std::vector<int> sched;
  {
    int rest = n;
    int ramp = (any_host && chunk >= 4 && n >= 2 * chunk) ? 1 : 0;    // 1: one half-size chunk at each end; 2: quarter + half
    if (const char* e = sw_get("RDB_RAMP")) ramp = (any_host && chunk >= 4 && n >= 2 * chunk) ? std::atoi(e) : 0;
    std::vector<int> head;
    if (ramp >= 2) head.push_back(chunk / 4);
    if (ramp >= 1) head.push_back(chunk / 2);
    for (int h : head) { sched.push_back(h); rest -= 2 * h; }
    while (rest > 0) { const int m = rest < chunk ? rest : chunk; sched.push_back(m); rest -= m; }
    for (int i = (int)head.size() - 1; i >= 0; --i) sched.push_back(head[i]);
  }
  int i0 = 0;
  for (int it = 0; it < (int)sched.size(); i0 += sched[it], ++it) {
    const int l = it % lanes, sl = it % slots;
    cudaStream_t ls = lane_[l];
    const int m = sched[it];
    const uint8_t* src_i = static_cast<const uint8_t*>(src) + (size_t)i0 * page_in;
    DetInput in = in_host_or_dev;
    const void* dsrc = src_i;
    if (!in_dev) {
      if (it >= slots) RDB_CUDA(cudaStreamWaitEvent(copy_in_, ev_compute_[sl], 0));   // slot's previous reader done
      RDB_CUDA(cudaMemcpyAsync(d_in[sl], src_i, page_in * m, cudaMemcpyHostToDevice, copy_in_));
      RDB_CUDA(cudaEventRecord(ev_in_[sl], copy_in_));
      RDB_CUDA(cudaStreamWaitEvent(ls, ev_in_[sl], 0));
      dsrc = d_in[sl];
    }
    if (any_host && it >= slots) RDB_CUDA(cudaStreamWaitEvent(ls, ev_out_[sl], 0));     // slot's previous D2H done
    if (do_resize) {
      cxs[l].begin("resize_linear_u8");
      resize_linear_u8_kernel<<<cdiv((long long)m * H * W, 256), 256, 0, ls>>>(static_cast<const uint8_t*>(dsrc), m, SH, SW, d_rs[l], H, W, d_xi, d_xa,
                                                                              d_yi, d_ya);
      cxs[l].end();
      dsrc = d_rs[l];
    }
    if (in.f32) in.f32 = static_cast<const float*>(dsrc); else in.u8 = static_cast<const uint8_t*>(dsrc);
    float* p_out = (prob_dev && prob) ? prob + (size_t)i0 * page_px : d_prob[sl];
    uint8_t* b_out = bitmap ? (bm_dev ? bitmap + (size_t)i0 * page_px : d_bm[sl]) : nullptr;
    if (precision_ == 0) forward_chunk<float>(cxs[l], in, m, H, W, thresh, dilate, p_out, b_out);
    else forward_chunk<__half>(cxs[l], in, m, H, W, thresh, dilate, p_out, b_out);
    if (any_host) {
      RDB_CUDA(cudaEventRecord(ev_compute_[sl], ls));
      RDB_CUDA(cudaStreamWaitEvent(copy_out_, ev_compute_[sl], 0));
      if (prob && !prob_dev) RDB_CUDA(cudaMemcpyAsync(prob + (size_t)i0 * page_px, d_prob[sl], page_px * m * sizeof(float), cudaMemcpyDeviceToHost, copy_out_));
      if (bitmap && !bm_dev) RDB_CUDA(cudaMemcpyAsync(bitmap + (size_t)i0 * page_px, d_bm[sl], page_px * m, cudaMemcpyDeviceToHost, copy_out_));
      RDB_CUDA(cudaEventRecord(ev_out_[sl], copy_out_));
    }
  }
  long long launches = 0;
  for (int l = 0; l < lanes; ++l) {
    RDB_CUDA(cudaEventRecord(ev_join_[l], lane_[l]));
    RDB_CUDA(cudaStreamWaitEvent(st, ev_join_[l], 0));
    launches += cxs[l].launches;
  }
  if (any_host) {
    RDB_CUDA(cudaStreamSynchronize(copy_out_));
    RDB_CUDA(cudaStreamSynchronize(st));
  }
  for (int s = 0; s < slots; ++s) {
    Pool& pl = pools_[s % lanes];
    if (d_in[s]) pl.free(d_in[s]);
    if (d_prob[s]) pl.free(d_prob[s]);
    if (d_bm[s]) pl.free(d_bm[s]);
  }
  for (int l = 0; l < lanes; ++l) if (d_rs[l]) pools_[l].free(d_rs[l]);
  if (do_resize) { pools_[0].free(d_xi); pools_[0].free(d_yi); pools_[0].free(d_xa); pools_[0].free(d_ya); }
  if (Profiler::global().on) { RDB_CUDA(cudaDeviceSynchronize()); Profiler::global().resolve(); }
  last_launches_ = launches;
}

void DetEngine::infer(const DetInput& in, int n, int H, int W, float thresh, bool dilate, float* prob, uint8_t* bitmap, cudaStream_t st) {
  try {
    infer_impl(in, n, H, W, thresh, dilate, prob, bitmap, st);
  } catch (...) {   // the call's pool blocks go back to the cache instead of staying "live" for ever
    cudaDeviceSynchronize();
    for (auto& p : pools_) p.reclaim();
    throw;
  }
  for (auto& p : pools_) p.enforce_cap();
}

// OpenCV resize() coefficient tables for INTER_LINEAR / 8U (float32 maths as in cv::resize -> saturate_cast<short>)
// Horizontal taps are clamped with the weight moved onto the edge pixel; vertical taps keep their weights and clamp
// only the row index (cv::resize builds yofs/beta without the edge adjustment and clips rows in the row loop).
static void linear_coeffs(int dsize, int ssize, bool vertical, std::vector<int>& idx, std::vector<short>& ab) {
  idx.resize(dsize); ab.resize(2 * (size_t)dsize);
  const double inv = (double)dsize / ssize, scale = 1.0 / inv;
  for (int d = 0; d < dsize; ++d) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (!vertical) {
      if (s < 0) { fx = 0.f; s = 0; }
      if (s >= ssize - 1) { fx = 0.f; s = ssize - 1; }
    }
    idx[d] = s;
    ab[2 * d] = (short)lrintf((1.f - fx) * 2048.f);
    ab[2 * d + 1] = (short)lrintf(fx * 2048.f);
  }
}

void linear_coeffs_cv(int dsize, int ssize, bool vertical, std::vector<int>& idx, std::vector<short>& ab) { linear_coeffs(dsize, ssize, vertical, idx, ab); }

void resize_linear_u8(int device, const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  RDB_CHECK(n > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, "resize: bad shape");
  std::vector<int> xi, yi; std::vector<short> xa, ya;
  linear_coeffs(dw, sw, false, xi, xa);
  linear_coeffs(dh, sh, true, yi, ya);
  const size_t in_b = (size_t)n * sh * sw * 3, out_b = (size_t)n * dh * dw * 3;
  const bool s_dev = is_device_ptr(src), d_dev = is_device_ptr(dst);
  uint8_t* ds = const_cast<uint8_t*>(src); uint8_t* dd = dst;
  int *dxi, *dyi; short *dxa, *dya;
  RDB_CUDA(cudaMalloc(&dxi, dw * 4)); RDB_CUDA(cudaMalloc(&dyi, dh * 4)); RDB_CUDA(cudaMalloc(&dxa, dw * 4)); RDB_CUDA(cudaMalloc(&dya, dh * 4));
  RDB_CUDA(cudaMemcpyAsync(dxi, xi.data(), dw * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dyi, yi.data(), dh * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dxa, xa.data(), dw * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dya, ya.data(), dh * 4, cudaMemcpyHostToDevice, st));
  if (!s_dev) { RDB_CUDA(cudaMalloc(&ds, in_b)); RDB_CUDA(cudaMemcpyAsync(ds, src, in_b, cudaMemcpyHostToDevice, st)); }
  if (!d_dev) RDB_CUDA(cudaMalloc(&dd, out_b));
  resize_linear_u8_kernel<<<cdiv((long long)n * dh * dw, 256), 256, 0, st>>>(ds, n, sh, sw, dd, dh, dw, dxi, dxa, dyi, dya);
  RDB_LAUNCH_CHECK();
  if (!d_dev) RDB_CUDA(cudaMemcpyAsync(dst, dd, out_b, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));   // the coefficient tables are freed below
  cudaFree(dxi); cudaFree(dyi); cudaFree(dxa); cudaFree(dya);
  if (!s_dev) cudaFree(ds);
  if (!d_dev) cudaFree(dd);
}

void db_bitmap(int device, const float* prob, int n, int H, int W, float thresh, bool dilate, uint8_t* bitmap, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  RDB_CHECK(W % 4 == 0, "db_bitmap: width must be a multiple of 4");
  const size_t px = (size_t)n * H * W;
  const bool p_dev = is_device_ptr(prob), b_dev = is_device_ptr(bitmap);
  float* dp = const_cast<float*>(prob);
  uint8_t* db = bitmap;
  uint8_t* seg = nullptr;
  if (!p_dev) { RDB_CUDA(cudaMalloc(&dp, px * 4)); RDB_CUDA(cudaMemcpyAsync(dp, prob, px * 4, cudaMemcpyHostToDevice, st)); }
  if (!b_dev) RDB_CUDA(cudaMalloc(&db, px));
  if (dilate) RDB_CUDA(cudaMalloc(&seg, px));
  threshold_kernel<<<cdiv((long long)(px + 3) / 4, 256), 256, 0, st>>>(dp, thresh, dilate ? seg : db, (long long)px);
  RDB_LAUNCH_CHECK();
  if (dilate) {
    dilate2x2_kernel<<<cdiv((long long)n * H * (W / 4), 256), 256, 0, st>>>(seg, n, H, W, db);
    RDB_LAUNCH_CHECK();
  }
  if (!b_dev) RDB_CUDA(cudaMemcpyAsync(bitmap, db, px, cudaMemcpyDeviceToHost, st));
  if (!p_dev || !b_dev || dilate) {
    RDB_CUDA(cudaStreamSynchronize(st));
    if (!p_dev) cudaFree(dp);
    if (!b_dev) cudaFree(db);
    if (seg) cudaFree(seg);
  }
}

}  // namespace rdb
