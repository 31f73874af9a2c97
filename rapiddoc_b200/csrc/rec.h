#pragma once
#include <memory>

#include "engine.cuh"

namespace rdb {

struct RecInput {
  const float* f32 = nullptr;      // [n,3,48,W] fp32 NCHW
  const uint8_t* u8 = nullptr;     // [n,48,W,3] uint8 BGR
  const int32_t* valid_w = nullptr;  // [n] (u8 only): columns >= valid_w are the zero right-pad
};

struct RecOutput {  // host or device pointers, any may be null
  int32_t* ids = nullptr; float* probs = nullptr;
  int32_t* text_ids = nullptr; int32_t* text_len = nullptr; float* conf = nullptr;
  float* softmax = nullptr;
};

class RecEngine {
 public:
  RecEngine(const void* blob, size_t nbytes, int device, int precision);
  ~RecEngine();
  void infer(const RecInput& in, int n, int W, const RecOutput& out, cudaStream_t st);
  void set_pool_cap(size_t bytes) { for (auto& p : pools_) p.set_cap(bytes); }
  size_t pool_bytes() const { size_t t = 0; for (auto& p : pools_) t += p.total_bytes(); return t; }
  int vocab() const { return vocab_; }
  // T for an input of width W: stem1 s2, stem3 s2, avg_pool [3,2]  (rec_lcnetv4.py:151,154,311)
  static int tokens_for_width(int W) { int w1 = (W - 1) / 2 + 1; int w2 = (w1 - 1) / 2 + 1; return w2 / 2; }
  long long last_launches() const { return last_launches_; }
  int device() const { return device_; }
  void set_chunk_crops(int c) { chunk_crops_ = c; }

 private:
  void infer_impl(const RecInput& in, int n, int W, const RecOutput& out, cudaStream_t st);
  template <typename T>
  void forward_chunk(Ctx& cx, const RecInput& in, int n, int W, int32_t* ids, float* probs, int32_t* text_ids, int32_t* text_len,
                     float* conf, float* softmax);
  int device_, precision_, vocab_ = 0;
  std::unique_ptr<Weights> weights_;
  void ensure_streams();
  Pool pools_[2];
  cudaStream_t lane_[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork_ = nullptr, ev_join_[2] = {nullptr, nullptr};
  long long last_launches_ = 0;
  int num_sms_ = 148;
  int chunk_crops_ = 256;
};

}  // namespace rdb
