// C-ABI of librapiddoc_b200.so (see include/rapiddoc_b200.h for the contract and the
// reference interfaces each entry point replaces).
#include "../../include/rapiddoc_b200.h"

#include <algorithm>
#include <cmath>
#include <vector>

#include "det.h"
#include "rec.h"
#include "warp.cuh"
#include "dbpost.cuh"
#include "layout.cuh"

struct rdb_det { rdb::DetEngine* e; };
struct rdb_rec { rdb::RecEngine* e; };

namespace {
thread_local std::string g_err;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return RDB_OK;
  } catch (const rdb::Error& e) {
    g_err = e.what();
    return g_err.find("cuda") != std::string::npos ? RDB_ERR_CUDA : RDB_ERR_INVALID;
  } catch (const std::exception& e) {
    g_err = e.what();
    return RDB_ERR_INVALID;
  }
}

// Clipper 6.4.2 ClipperOffset for ONE closed polygon, JT_ROUND, ET_CLOSEDPOLYGON, delta>0
// (published algorithm: AddPath / FixOrientations / DoOffset / OffsetPoint / DoRound with
// arc tolerance 0.25; Round() = half away from zero).  Replaces the pyclipper call at
// rapid_doc/model/ocr/ocr_patch.py:166-171.  The trailing clean-up union is omitted: it
// does not change the vertex set of the convex quads DB feeds it, and the consumer
// (cv2.minAreaRect) only sees the convex hull.
inline long long cround(double v) { return v < 0 ? (long long)(v - 0.5) : (long long)(v + 0.5); }

int clipper_offset(const double* xy, int n_in, double delta, int64_t* out, int max_pts) {
  struct P { long long x, y; };
  std::vector<P> pts;
  for (int i = 0; i < n_in; ++i) pts.push_back({(long long)xy[2 * i], (long long)xy[2 * i + 1]});  // pyclipper truncates
  int hi = n_in - 1;
  if (hi < 0) return 0;
  while (hi > 0 && pts[0].x == pts[hi].x && pts[0].y == pts[hi].y) --hi;
  std::vector<P> src{pts[0]};
  for (int i = 1; i <= hi; ++i)
    if (src.back().x != pts[i].x || src.back().y != pts[i].y) src.push_back(pts[i]);
  const int n = (int)src.size();
  if (n < 3) return 0;
  double a = 0;
  for (int i = 0, j = n - 1; i < n; j = i++) a += ((double)src[j].x + src[i].x) * ((double)src[j].y - src[i].y);
  if (-a * 0.5 < 0) std::reverse(src.begin(), src.end());
  std::vector<P> dst;
  if (std::fabs(delta) < 1e-20) dst = src;
  else {
    const double pi = 3.141592653589793238;
    double y = 0.25;
    if (y > std::fabs(delta) * 0.25) y = std::fabs(delta) * 0.25;
    double steps = pi / std::acos(1 - y / std::fabs(delta));
    if (steps > std::fabs(delta) * pi) steps = std::fabs(delta) * pi;
    double m_sin = std::sin(2 * pi / steps), m_cos = std::cos(2 * pi / steps);
    const double steps_per_rad = steps / (2 * pi);
    if (delta < 0) m_sin = -m_sin;
    std::vector<double> nx(n), ny(n);
    for (int i = 0; i < n; ++i) {
      const P& p1 = src[i];
      const P& p2 = src[(i + 1) % n];
      double dx = (double)(p2.x - p1.x), dy = (double)(p2.y - p1.y);
      double f = 1.0 / std::sqrt(dx * dx + dy * dy);
      nx[i] = dy * f;
      ny[i] = -dx * f;
    }
    int k = n - 1;
    for (int j = 0; j < n; ++j) {
      double sin_a = nx[k] * ny[j] - nx[j] * ny[k];
      if (std::fabs(sin_a * delta) < 1.0) {
        double cos_a = nx[k] * nx[j] + ny[j] * ny[k];
        if (cos_a > 0) {   // angle ~ 0: OffsetPoint returns here, BEFORE its trailing `k = j` (Clipper 6.4.2 clipper.cpp)
          dst.push_back({cround(src[j].x + nx[k] * delta), cround(src[j].y + ny[k] * delta)});
          continue;
        }
      } else if (sin_a > 1.0) sin_a = 1.0;
      else if (sin_a < -1.0) sin_a = -1.0;
      {
        if (sin_a * delta < 0) {
          dst.push_back({cround(src[j].x + nx[k] * delta), cround(src[j].y + ny[k] * delta)});
          dst.push_back(src[j]);
          dst.push_back({cround(src[j].x + nx[j] * delta), cround(src[j].y + ny[j] * delta)});
        } else {
          double ang = std::atan2(sin_a, nx[k] * nx[j] + ny[k] * ny[j]);
          long long st = cround(steps_per_rad * std::fabs(ang));
          if (st < 1) st = 1;
          double X = nx[k], Y = ny[k];
          for (long long i = 0; i < st; ++i) {
            dst.push_back({cround(src[j].x + X * delta), cround(src[j].y + Y * delta)});
            double X2 = X;
            X = X * m_cos - m_sin * Y;
            Y = X2 * m_sin + Y * m_cos;
          }
          dst.push_back({cround(src[j].x + nx[j] * delta), cround(src[j].y + ny[j] * delta)});
        }
      }
      k = j;
    }
  }
  int cnt = (int)dst.size();
  if (cnt > max_pts) throw rdb::Error("clipper_offset: output buffer too small");
  for (int i = 0; i < cnt; ++i) { out[2 * i] = dst[i].x; out[2 * i + 1] = dst[i].y; }
  return cnt;
}
}  // namespace

extern "C" {

int rdb_version(void) { return 100; }
const char* rdb_last_error(void) { return g_err.c_str(); }

int rdb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int rdb_pinned_alloc(size_t nbytes, void** out) {
  return guarded([&] { RDB_CHECK(out != nullptr, "null out"); RDB_CUDA(cudaHostAlloc(out, nbytes, cudaHostAllocDefault)); });
}
int rdb_pinned_free(void* p) {
  return guarded([&] { RDB_CUDA(cudaFreeHost(p)); });
}

static void require_device(int device) {
  // cudaGetDeviceProperties costs milliseconds per call: the verdict per device index is cached (it cannot change within a process)
  static int ok[rdb::kMaxDevices] = {};
  if (device >= 0 && device < rdb::kMaxDevices && ok[device]) return;
  int n = rdb_device_count();
  if (n <= 0) throw rdb::Error("cuda: no CUDA device visible — rapiddoc_b200 has no CPU fallback");
  if (device < 0 || device >= n) throw rdb::Error("invalid device index");
  int major = 0, minor = 0;
  RDB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  RDB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) throw rdb::Error(std::string("cuda: device is sm_") + std::to_string(major * 10 + minor) + ", this library is built for sm_100a only");
  if (device < rdb::kMaxDevices) ok[device] = 1;
}

int rdb_det_create(const void* weights, size_t nbytes, int device, int precision, rdb_det_t** out) {
  int rc = guarded([&] {
    RDB_CHECK(weights && out, "null argument");
    RDB_CHECK(precision == RDB_PREC_FP32 || precision == RDB_PREC_FP16, "bad precision");
    require_device(device);
    rdb::DeviceGuard g(device);
    *out = new rdb_det{new rdb::DetEngine(weights, nbytes, device, precision)};
  });
  if (rc == RDB_ERR_CUDA && g_err.find("no CUDA device") != std::string::npos) rc = RDB_ERR_NO_DEVICE;
  return rc;
}
void rdb_det_destroy(rdb_det_t* h) {
  if (h) { rdb::DeviceGuard g(h->e->device()); delete h->e; delete h; }
}

int rdb_det_infer_f32(rdb_det_t* h, const float* x, int n, int hgt, int wid, float* prob, void* stream) {
  return guarded([&] {
    RDB_CHECK(h && x && prob, "null argument");
    rdb::DeviceGuard g(h->e->device());
    rdb::DetInput in;
    in.f32 = x;
    h->e->infer(in, n, hgt, wid, 0.3f, false, prob, nullptr, (cudaStream_t)stream);
  });
}

int rdb_det_infer_u8(rdb_det_t* h, const uint8_t* pages, int n, int hgt, int wid, const float mean[3], const float stdv[3],
                     float thresh, int use_dilation, float* prob, uint8_t* bitmap, void* stream) {
  return guarded([&] {
    RDB_CHECK(h && pages && mean && stdv && (prob || bitmap), "null argument");
    rdb::DeviceGuard g(h->e->device());
    rdb::DetInput in;
    in.u8 = pages;
    for (int i = 0; i < 3; ++i) { in.mean[i] = mean[i]; in.stdv[i] = stdv[i]; }
    h->e->infer(in, n, hgt, wid, thresh, use_dilation != 0, prob, bitmap, (cudaStream_t)stream);
  });
}

int rdb_det_infer_u8_resize(rdb_det_t* h, const uint8_t* pages, int n, int src_h, int src_w, int hgt, int wid, const float mean[3],
                            const float stdv[3], float thresh, int use_dilation, float* prob, uint8_t* bitmap, void* stream) {
  return guarded([&] {
    RDB_CHECK(h && pages && mean && stdv && (prob || bitmap) && src_h > 0 && src_w > 0, "null argument");
    rdb::DeviceGuard g(h->e->device());
    rdb::DetInput in;
    in.u8 = pages; in.src_h = src_h; in.src_w = src_w;
    for (int i = 0; i < 3; ++i) { in.mean[i] = mean[i]; in.stdv[i] = stdv[i]; }
    h->e->infer(in, n, hgt, wid, thresh, use_dilation != 0, prob, bitmap, (cudaStream_t)stream);
  });
}

int rdb_db_bitmap(int device, const float* prob, int n, int hgt, int wid, float thresh, int use_dilation, uint8_t* bitmap,
                  void* stream) {
  return guarded([&] {
    RDB_CHECK(prob && bitmap && n > 0, "null argument");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::db_bitmap(device, prob, n, hgt, wid, thresh, use_dilation != 0, bitmap, (cudaStream_t)stream);
  });
}

int rdb_resize_linear_u8(int device, const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, void* stream) {
  return guarded([&] {
    RDB_CHECK(src && dst, "null argument");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::resize_linear_u8(device, src, n, sh, sw, dst, dh, dw, (cudaStream_t)stream);
  });
}

int rdb_warp_crops(int device, const uint8_t* page, int hgt, int wid, int n, const double* minv, const int32_t* sizes, const int32_t* rotate,
                   uint8_t* out, const int64_t* offsets, int64_t out_bytes, void* stream) {
  return guarded([&] {
    RDB_CHECK(page && out && (n == 0 || (minv && sizes && offsets)), "null argument");
    RDB_CHECK(hgt > 0 && wid > 0 && n >= 0 && out_bytes >= 0, "warp: bad shape");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::warp_crops(device, page, hgt, wid, n, minv, sizes, rotate, out, reinterpret_cast<const long long*>(offsets), (long long)out_bytes, (cudaStream_t)stream);
  });
}

int rdb_resize_pack_u8(int device, const uint8_t* src, int64_t src_bytes, int n, const int64_t* src_offsets, const int32_t* sizes, const int32_t* dst_w,
                       uint8_t* dst, int hgt, int wid_max, void* stream) {
  return guarded([&] {
    RDB_CHECK(src && dst && (n == 0 || (src_offsets && sizes && dst_w)), "null argument");
    RDB_CHECK(n >= 0 && hgt > 0 && wid_max > 0 && src_bytes >= 0, "resize_pack: bad shape");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::resize_pack_u8(device, src, (long long)src_bytes, n, reinterpret_cast<const long long*>(src_offsets), sizes, dst_w, dst, hgt, wid_max, (cudaStream_t)stream);
  });
}

int rdb_warp_crops_batch(int device, const uint8_t* pages, int n_pages, int hgt, int wid, int n, const int32_t* page_idx, const double* minv,
                         const int32_t* sizes, const int32_t* rotate, uint8_t* out, const int64_t* offsets, int64_t out_bytes, void* stream) {
  return guarded([&] {
    RDB_CHECK(pages && out && (n == 0 || (minv && sizes && offsets)), "null argument");
    RDB_CHECK(hgt > 0 && wid > 0 && n >= 0 && n_pages > 0 && out_bytes >= 0, "warp: bad shape");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::warp_crops(device, pages, hgt, wid, n, minv, sizes, rotate, out, reinterpret_cast<const long long*>(offsets), (long long)out_bytes,
                    (cudaStream_t)stream, n_pages, page_idx);
  });
}

int rdb_resize_pack_slots(int device, const uint8_t* src, int64_t src_bytes, int n, const int64_t* src_offsets, const int32_t* sizes,
                          const int32_t* dst_w, const int64_t* dst_offsets, const int32_t* dst_pitch, uint8_t* dst, int64_t dst_bytes, int hgt,
                          void* stream) {
  return guarded([&] {
    RDB_CHECK(src && dst && (n == 0 || (src_offsets && sizes && dst_w && dst_offsets && dst_pitch)), "null argument");
    RDB_CHECK(n >= 0 && hgt > 0 && src_bytes >= 0 && dst_bytes >= 0, "resize_pack_slots: bad shape");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::resize_pack_slots(device, src, (long long)src_bytes, n, reinterpret_cast<const long long*>(src_offsets), sizes, dst_w,
                           reinterpret_cast<const long long*>(dst_offsets), dst_pitch, dst, (long long)dst_bytes, hgt, (cudaStream_t)stream);
  });
}

int rdb_db_box_scores(int device, const float* prob, int n, int hgt, int wid, int m, const float* quads, const int32_t* page_idx, double* scores,
                      int32_t* flags, void* stream) {
  return guarded([&] {
    RDB_CHECK(prob && (m == 0 || (quads && scores && flags)), "null argument");
    RDB_CHECK(n > 0 && hgt > 0 && wid > 0 && m >= 0, "box_scores: bad shape");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::db_box_scores(device, prob, n, hgt, wid, m, quads, page_idx, scores, flags, (cudaStream_t)stream);
  });
}

int rdb_debug_fill_quad(const int32_t* pts_xy, int mw, int mh, uint8_t* mask) {
  return guarded([&] {
    RDB_CHECK(pts_xy && mask && mw > 0 && mh > 0, "null argument");
    for (int i = 0; i < 4; ++i)
      RDB_CHECK(pts_xy[2 * i] >= 0 && pts_xy[2 * i] < mw && pts_xy[2 * i + 1] >= 0 && pts_xy[2 * i + 1] < mh, "fill_quad: vertex outside the mask");
    rdb::debug_fill_quad(pts_xy, mw, mh, mask);
  });
}

int rdb_det_set_pool_cap_bytes(rdb_det_t* h, size_t bytes) { if (!h) return RDB_ERR_INVALID; h->e->set_pool_cap(bytes); return RDB_OK; }
int rdb_rec_set_pool_cap_bytes(rdb_rec_t* h, size_t bytes) { if (!h) return RDB_ERR_INVALID; h->e->set_pool_cap(bytes); return RDB_OK; }
long long rdb_det_pool_bytes(rdb_det_t* h) { return h ? (long long)h->e->pool_bytes() : -1; }
long long rdb_rec_pool_bytes(rdb_rec_t* h) { return h ? (long long)h->e->pool_bytes() : -1; }

int rdb_layout_nms(int device, const float* boxes, int stride, const int32_t* order, const int32_t* offsets, int pages, float iou_same, float iou_diff,
                   int32_t* keep, int32_t* keep_n, void* stream) {
  return guarded([&] {
    RDB_CHECK(boxes && order && offsets && keep && keep_n && stride >= 6 && pages >= 0, "null argument");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::layout_nms(device, boxes, stride, order, offsets, pages, iou_same, iou_diff, keep, keep_n, (cudaStream_t)stream);
  });
}

int rdb_layout_containment(int device, const float* boxes, int stride, const int32_t* offsets, int pages, int formula_index, int category_index,
                           int mode, int32_t* contains_other, int32_t* contained_by_other, void* stream) {
  return guarded([&] {
    RDB_CHECK(boxes && offsets && contains_other && contained_by_other && stride >= 6 && pages >= 0, "null argument");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::layout_containment(device, boxes, stride, offsets, pages, formula_index, category_index, mode, contains_other, contained_by_other,
                            (cudaStream_t)stream);
  });
}

int rdb_argmax_rows(int device, const float* x, long long rows, int vocab, int32_t* idx, float* val, void* stream) {
  return guarded([&] {
    RDB_CHECK(x && idx && val && rows >= 0 && vocab > 0, "null argument");
    require_device(device);
    rdb::DeviceGuard g(device);
    rdb::argmax_rows(device, x, rows, vocab, idx, val, (cudaStream_t)stream);
  });
}

int rdb_debug_cubic_tab(int16_t* out) {
  return guarded([&] {
    RDB_CHECK(out != nullptr, "null argument");
    std::vector<short> t;
    rdb::build_cubic_tab(t);
    for (size_t i = 0; i < t.size(); ++i) out[i] = t[i];
  });
}

int rdb_clipper_offset(const double* box_xy, int n_pts, double distance, int64_t* out_xy, int max_pts) {
  int cnt = 0;
  int rc = guarded([&] {
    RDB_CHECK(box_xy && out_xy && n_pts >= 0, "null argument");
    cnt = clipper_offset(box_xy, n_pts, distance, out_xy, max_pts);
  });
  return rc == RDB_OK ? cnt : rc;
}

int rdb_clipper_offset_batch(const double* boxes_xy, int m, const double* distances, int64_t* out_xy, int max_pts_total, int32_t* counts) {
  int total = 0;
  int rc = guarded([&] {
    RDB_CHECK(m >= 0 && (m == 0 || (boxes_xy && distances && out_xy && counts)), "null argument");
    for (int i = 0; i < m; ++i) {
      const int n = clipper_offset(boxes_xy + (size_t)i * 8, 4, distances[i], out_xy + (size_t)total * 2, max_pts_total - total);
      counts[i] = n;
      total += n;
    }
  });
  return rc == RDB_OK ? total : rc;
}

int rdb_rec_create(const void* weights, size_t nbytes, int device, int precision, rdb_rec_t** out) {
  int rc = guarded([&] {
    RDB_CHECK(weights && out, "null argument");
    RDB_CHECK(precision == RDB_PREC_FP32 || precision == RDB_PREC_FP16, "bad precision");
    require_device(device);
    rdb::DeviceGuard g(device);
    *out = new rdb_rec{new rdb::RecEngine(weights, nbytes, device, precision)};
  });
  if (rc == RDB_ERR_CUDA && g_err.find("no CUDA device") != std::string::npos) rc = RDB_ERR_NO_DEVICE;
  return rc;
}
void rdb_rec_destroy(rdb_rec_t* h) {
  if (h) { rdb::DeviceGuard g(h->e->device()); delete h->e; delete h; }
}
int rdb_rec_vocab(rdb_rec_t* h) { return h ? h->e->vocab() : RDB_ERR_INVALID; }
int rdb_rec_tokens(int wid) { return rdb::RecEngine::tokens_for_width(wid); }

int rdb_rec_infer_f32(rdb_rec_t* h, const float* x, int n, int wid, int32_t* ids, float* probs, int32_t* text_ids,
                      int32_t* text_len, float* conf, float* softmax, void* stream) {
  return guarded([&] {
    RDB_CHECK(h && x, "null argument");
    rdb::DeviceGuard g(h->e->device());
    rdb::RecInput in;
    in.f32 = x;
    rdb::RecOutput o;
    o.ids = ids; o.probs = probs; o.text_ids = text_ids; o.text_len = text_len; o.conf = conf; o.softmax = softmax;
    h->e->infer(in, n, wid, o, (cudaStream_t)stream);
  });
}

int rdb_rec_infer_u8(rdb_rec_t* h, const uint8_t* crops, const int32_t* valid_w, int n, int wid, int32_t* ids, float* probs,
                     int32_t* text_ids, int32_t* text_len, float* conf, void* stream) {
  return guarded([&] {
    RDB_CHECK(h && crops, "null argument");
    rdb::DeviceGuard g(h->e->device());
    rdb::RecInput in;
    in.u8 = crops;
    in.valid_w = valid_w;
    rdb::RecOutput o;
    o.ids = ids; o.probs = probs; o.text_ids = text_ids; o.text_len = text_len; o.conf = conf;
    h->e->infer(in, n, wid, o, (cudaStream_t)stream);
  });
}

long long rdb_det_last_launches(rdb_det_t* h) { return h ? h->e->last_launches() : -1; }
long long rdb_rec_last_launches(rdb_rec_t* h) { return h ? h->e->last_launches() : -1; }

// Diagnostic entry: one GEMM through the fp16 engines (use_tc=1: tcgen05 kernel, 0: SIMT kernel on
// fp16 storage).  Host fp32 in/out; inputs are rounded to fp16 on the device exactly as the
// engines store them.  out[M,N] = act(A W^T + bias) (+res).  mode 1 = CTC partial epilogue:
// out receives [M,2] = (argmax id, softmax max prob).
static __global__ void f2h_kernel(const float* in, __half* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
static __global__ void h2f_kernel(const __half* in, float* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half2float(in[i]);
}
int rdb_debug_gemm(int device, int use_tc, int mode, const float* A, const float* W, const float* bias, const float* res, int M,
                   int N, int K, int act, float* out) {
  return guarded([&] {
    require_device(device);
    rdb::DeviceGuard g(device);
    RDB_CUDA(cudaSetDevice(device));
    rdb::Pool pool;
    rdb::Ctx cx;
    cx.pool = &pool; cx.st = nullptr; cx.use_tc = use_tc != 0;
    RDB_CUDA(cudaDeviceGetAttribute(&cx.num_sms, cudaDevAttrMultiProcessorCount, device));
    auto up = [&](const float* h, size_t n, float** df, __half** dh) {
      *df = pool.alloc_t<float>(n);
      *dh = pool.alloc_t<__half>(n);
      RDB_CUDA(cudaMemcpy(*df, h, n * 4, cudaMemcpyHostToDevice));
      f2h_kernel<<<(unsigned)((n + 255) / 256), 256>>>(*df, *dh, n);
    };
    float *dAf, *dWf, *dRf = nullptr, *dB = nullptr; __half *dAh, *dWh, *dRh = nullptr;
    up(A, (size_t)M * K, &dAf, &dAh);
    up(W, (size_t)N * K, &dWf, &dWh);
    if (res) up(res, (size_t)M * N, &dRf, &dRh);
    if (bias) { dB = pool.alloc_t<float>(N); RDB_CUDA(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice)); }
    if (mode == 0) {
      __half* dO = pool.alloc_t<__half>((size_t)M * N);
      float* dOf = pool.alloc_t<float>((size_t)M * N);
      if (use_tc) {
        const char* reps_env = rdb::sw_get("RDB_DEBUG_GEMM_REPS");   // timing aid: repeat the launch (see rdb_profile_*)
        const int reps = reps_env ? atoi(reps_env) : 1;
        for (int r = 0; r < (reps > 1 ? reps : 1); ++r) rdb::launch_gemm_tc(cx, dAh, K, M, K, dWh, N, dB, act, dRh, N, dO, N, 0);
      } else {
        rdb::GemmArgs g{};
        g.A = dAh; g.lda = K; g.W = dWf; g.bias = dB; g.res = dRh; g.ldr = N; g.out = dO; g.ldc = N; g.M = M; g.N = N; g.K = K; g.act = act;
        rdb::launch_gemm_simt<__half, __half>(g, nullptr);
      }
      h2f_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(dO, dOf, (size_t)M * N);
      RDB_CUDA(cudaMemcpy(out, dOf, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    } else {
      int tiles = rdb::cdiv(N, rdb::SG_BN);
      if (tiles < 320) tiles = 320;
      float* pmax = pool.alloc_t<float>((size_t)M * tiles);
      float* psum = pool.alloc_t<float>((size_t)M * tiles);
      int* pidx = pool.alloc_t<int>((size_t)M * tiles);
      int* ids = pool.alloc_t<int>(M);
      float* pr = pool.alloc_t<float>(M);
      if (use_tc) {
        rdb::launch_gemm_tc_ctc(cx, dAh, K, M, K, dWh, N, dB, pmax, pidx, psum, &tiles);
      } else {
        rdb::GemmArgs g{};
        g.A = dAh; g.lda = K; g.W = dWf; g.bias = dB; g.M = M; g.N = N; g.K = K; g.pmax = pmax; g.pidx = pidx; g.psum = psum;
        tiles = rdb::cdiv(N, rdb::SG_BN);
        rdb::launch_gemm_simt_ctc<__half>(g, nullptr);
      }
      rdb::ctc_merge_kernel<<<rdb::cdiv(M, 128), 128>>>(pmax, pidx, psum, M, tiles, ids, pr);
      std::vector<int> hi(M); std::vector<float> hp(M);
      RDB_CUDA(cudaMemcpy(hi.data(), ids, (size_t)M * 4, cudaMemcpyDeviceToHost));
      RDB_CUDA(cudaMemcpy(hp.data(), pr, (size_t)M * 4, cudaMemcpyDeviceToHost));
      for (int i = 0; i < M; ++i) { out[2 * i] = (float)hi[i]; out[2 * i + 1] = hp[i]; }
    }
    RDB_CUDA(cudaDeviceSynchronize());
    if (rdb::Profiler::global().on) rdb::Profiler::global().resolve();
  });
}

int rdb_switches_reload(void) {
  rdb::Switches::get().reload();
  return RDB_OK;
}

int rdb_profile_enable(int on) {
  rdb::Profiler::global().on = (on != 0);
  return RDB_OK;
}
int rdb_profile_reset(void) {
  rdb::Profiler::global().acc.clear();
  return RDB_OK;
}
// JSON object {"kernel name": [total_ms, launches], ...}; returns bytes needed (incl. NUL)
int rdb_profile_dump(char* buf, size_t cap) {
  if (!rdb::Profiler::global().recs.empty()) {      // op-level records (rdb_op_*) are resolved here
    cudaDeviceSynchronize();
    try { rdb::Profiler::global().resolve(); } catch (...) {}
  }
  std::string s = "{";
  bool first = true;
  for (auto& kv : rdb::Profiler::global().acc) {
    if (!first) s += ", ";
    first = false;
    s += "\"" + kv.first + "\": [" + std::to_string(kv.second.first) + ", " + std::to_string(kv.second.second) + "]";
  }
  s += "}";
  if (buf && cap > 0) {
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return (int)s.size() + 1;
}

int rdb_det_set_chunk_pixels(rdb_det_t* h, long long px) {
  if (!h || px <= 0) return RDB_ERR_INVALID;
  h->e->set_chunk_pixels(px);
  return RDB_OK;
}
int rdb_rec_set_chunk_crops(rdb_rec_t* h, int crops) {
  if (!h || crops <= 0) return RDB_ERR_INVALID;
  h->e->set_chunk_crops(crops);
  return RDB_OK;
}

}  // extern "C"
