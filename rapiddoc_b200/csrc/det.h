#pragma once
#include <memory>

#include "engine.cuh"

namespace rdb {

struct DetInput {
  const float* f32 = nullptr;    // [n,3,H,W] fp32 NCHW (InferSession seam)
  const uint8_t* u8 = nullptr;   // [n,H,W,3] uint8 BGR (facade seam)
  float mean[3] = {0, 0, 0}, stdv[3] = {1, 1, 1};
  int src_h = 0, src_w = 0;      // u8 only: pages are [n,src_h,src_w,3] and are resized on the GPU (cv2 INTER_LINEAR, bit-exact) to H x W
};

class DetEngine {
 public:
  DetEngine(const void* blob, size_t nbytes, int device, int precision);
  ~DetEngine();
  // prob [n,H,W] f32 / bitmap [n,H,W] u8: host or device pointers, may be null.
  void infer(const DetInput& in, int n, int H, int W, float thresh, bool dilate, float* prob, uint8_t* bitmap, cudaStream_t st);
  void set_pool_cap(size_t bytes) { for (auto& p : pools_) p.set_cap(bytes); }
  size_t pool_bytes() const { size_t t = 0; for (auto& p : pools_) t += p.total_bytes(); return t; }
  long long last_launches() const { return last_launches_; }
  int device() const { return device_; }
  void set_chunk_pixels(long long px) { chunk_pixels_ = px; chunk_pixels_set_ = true; }

 private:
  void infer_impl(const DetInput& in, int n, int H, int W, float thresh, bool dilate, float* prob, uint8_t* bitmap, cudaStream_t st);
  template <typename T>
  void forward_chunk(Ctx& cx, const DetInput& in, int n, int H, int W, float thresh, bool dilate, float* prob, uint8_t* bitmap);
  void ensure_streams();
  int device_, precision_;
  static constexpr int kMaxLanes = 4;
  cudaStream_t copy_in_ = nullptr, copy_out_ = nullptr, lane_[kMaxLanes] = {};
  static constexpr int kSlots = 2 * kMaxLanes;   // host-staging buffers: two per lane, so a lane never waits on its own copies
  cudaEvent_t ev_in_[kSlots] = {}, ev_out_[kSlots] = {}, ev_compute_[kSlots] = {};
  cudaEvent_t ev_join_[kMaxLanes] = {}, ev_fork_ = nullptr;
  std::unique_ptr<Weights> weights_;
  Pool pools_[kMaxLanes];   // one per compute lane
  long long last_launches_ = 0;
  int num_sms_ = 148;
  bool chunk_pixels_set_ = false;
  long long chunk_pixels_ = 16ll * 1024 * 1024;  // pages per internal chunk = chunk_pixels / (H*W)
};

// cv2.resize(INTER_LINEAR) on uint8 HWC images [n,sh,sw,3] -> [n,dh,dw,3], bit-exact (host or device pointers)
void resize_linear_u8(int device, const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, cudaStream_t st);

void db_bitmap(int device, const float* prob, int n, int H, int W, float thresh, bool dilate, uint8_t* bitmap, cudaStream_t st);

}  // namespace rdb
