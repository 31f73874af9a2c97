// get_rotate_crop_image (rapid_doc/utils/ocr_utils.py:494-537) on the GPU: cv2.warpPerspective(INTER_CUBIC, BORDER_REPLICATE)
// for a batch of text-line quads of one page, bit-exact with OpenCV's imgwarp.cpp:
//   * inverse map in double, evaluated per OpenCV block (block origin x first, then + M*x1, in that order, plain IEEE
//     mul/add — __dmul_rn/__dadd_rn keep nvcc from contracting into FMA), scaled by INTER_TAB_SIZE/W, clamped to int range and
//     rounded half-to-even into a 5-bit fixed-point source coordinate;
//   * remapBicubic: 4x4 int16 weights from the 1024-entry table (a = -0.75 cubic, float 1-D tables multiplied in float,
//     x 2^15, rounded, corrected so each 16 weights sum to 32768), replicate border = clamped indices, (sum + 2^14) >> 15.
// One thread per destination pixel (3 channels); the optional np.rot90 of tall crops is folded into the store index.
// The CPU restatement that pins this against cv2 itself is oracle/warp.py (tests/test_oracle.py, tests/test_gpu_warp.py).
#pragma once
#include <vector>

#include "engine.cuh"

namespace rdb {

// Grow-only per-device scratch for the (synchronous) crop entry points: cudaMalloc / cudaFree per call cost ~40 ms per page at the
// facade level (dozens of them, each a device-wide synchronisation) — measured with tools/pipeline_probe.py.  Not thread-safe,
// like the engine handles; every call ends with a stream synchronise, so the buffer is free again when the next call starts.
struct ScratchCarver {
  uint8_t* base; size_t off = 0;
  template <typename T>
  T* take(size_t n) { T* p = reinterpret_cast<T*>(base + off); off += (n * sizeof(T) + 255) / 256 * 256; return p; }
};
inline uint8_t* device_scratch(int device, size_t bytes) {
  // per host thread: the detection stage of one window and the recognition stage of the previous one may run on two
  // threads at once (B200OcrModel.ocr_pages_stream); each keeps its own grow-only scratch
  static thread_local uint8_t* buf[64] = {};
  static thread_local size_t cap[64] = {};
  RDB_CHECK(device >= 0 && device < 64, "scratch: device index");
  if (cap[device] < bytes) {
    if (buf[device]) { RDB_CUDA(cudaDeviceSynchronize()); cudaFree(buf[device]); buf[device] = nullptr; cap[device] = 0; }
    size_t want = bytes + bytes / 2;
    if (want < (size_t)8 << 20) want = (size_t)8 << 20;
    RDB_CUDA(cudaMalloc(&buf[device], want));
    cap[device] = want;
  }
  return buf[device];
}
inline size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

// Grow-only PINNED host staging per host thread.  Small descriptor tables / results go through it instead of pageable memory:
// a pageable cudaMemcpyAsync is a blocking call inside the driver, and two of them from different threads serialise — with the
// two-stage window pipeline the box scorer of window k+1 was observed waiting 13 ms per call behind the recogniser's
// (pageable) result copy of window k.  Pinned copies are truly asynchronous; the caller synchronises its own stream.
struct HostCarver {
  uint8_t* base; size_t off = 0;
  template <typename T>
  T* take(size_t n) { T* p = reinterpret_cast<T*>(base + off); off += (n * sizeof(T) + 255) / 256 * 256; return p; }
};
inline uint8_t* pinned_scratch(size_t bytes) {
  static thread_local uint8_t* buf = nullptr;
  static thread_local size_t cap = 0;
  if (cap < bytes) {
    if (buf) cudaFreeHost(buf);
    size_t want = bytes + bytes / 2;
    if (want < (size_t)1 << 20) want = (size_t)1 << 20;
    RDB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&buf), want, cudaHostAllocDefault));
    cap = want;
  }
  return buf;
}

struct WarpCrop {
  double m[9];          // dst -> src homography (cv::invert of getPerspectiveTransform)
  int w, h;             // warp output size (before rotation)
  int rotate;           // 1: store np.rot90(dst) ([w][h][3]) instead of dst ([h][w][3])
  int bw0;              // OpenCV's block width for this destination size
  long long offset;     // byte offset of this crop in the output buffer
  long long src_off;    // byte offset of this crop's page in the page buffer (pages are same-size [n_pages,H,W,3])
};

static __global__ void __launch_bounds__(256) warp_cubic_kernel(const uint8_t* __restrict__ src, int H, int W, const WarpCrop* __restrict__ crops,
                                                                const short* __restrict__ tab /*[1024][16]*/, uint8_t* __restrict__ out) {
  const WarpCrop& c = crops[blockIdx.y];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)c.w * c.h) return;
  const int x = (int)(idx % c.w), y = (int)(idx / c.w);
  const int xb = (x / c.bw0) * c.bw0, x1 = x - xb;
  const double dx = (double)xb, dy = (double)y, d1 = (double)x1;
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(c.m[0], dx), __dmul_rn(c.m[1], dy)), c.m[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(c.m[3], dx), __dmul_rn(c.m[4], dy)), c.m[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(c.m[6], dx), __dmul_rn(c.m[7], dy)), c.m[8]);
  double Wd = __dadd_rn(W0, __dmul_rn(c.m[6], d1));
  Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(c.m[0], d1)), Wd)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(c.m[3], d1)), Wd)));
  const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
  const int sx = max(-32768, min(32767, X >> 5)) - 1, sy = max(-32768, min(32767, Y >> 5)) - 1;
  const short* w = tab + (((Y & 31) << 5) + (X & 31)) * 16;
  int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    const int iy = max(0, min(H - 1, sy + ky));
    const uint8_t* row = src + c.src_off + (long long)iy * W * 3;
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int ix = max(0, min(W - 1, sx + kx));
      const int wt = w[ky * 4 + kx];
      const uint8_t* p = row + ix * 3;
      s0 += p[0] * wt; s1 += p[1] * wt; s2 += p[2] * wt;
    }
  }
  const long long o = c.rotate ? ((long long)(c.w - 1 - x) * c.h + y) : idx;
  uint8_t* d = out + c.offset + o * 3;
  d[0] = (uint8_t)max(0, min(255, (s0 + (1 << 14)) >> 15));
  d[1] = (uint8_t)max(0, min(255, (s1 + (1 << 14)) >> 15));
  d[2] = (uint8_t)max(0, min(255, (s2 + (1 << 14)) >> 15));
}

// initInterTab2D(INTER_CUBIC, fixpt): float arithmetic exactly as OpenCV (no contraction: host code, no FMA ISA enabled)
inline void build_cubic_tab(std::vector<short>& tab) {
  float t1[32][4];
  const float A = -0.75f, scale = 1.f / 32.f;
  for (int i = 0; i < 32; ++i) {
    volatile float x = (float)i * scale;
    volatile float x1 = x + 1.f;
    volatile float c0 = ((A * x1 - 5.f * A) * x1 + 8.f * A) * x1 - 4.f * A;
    volatile float c1 = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    volatile float xm = 1.f - x;
    volatile float c2 = ((A + 2.f) * xm - (A + 3.f)) * xm * xm + 1.f;
    volatile float c3 = 1.f - c0 - c1 - c2;
    t1[i][0] = c0; t1[i][1] = c1; t1[i][2] = c2; t1[i][3] = c3;
  }
  tab.assign(1024 * 16, 0);
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      int it[4][4], isum = 0;
      for (int k1 = 0; k1 < 4; ++k1)
        for (int k2 = 0; k2 < 4; ++k2) {
          volatile float v = t1[i][k1] * t1[j][k2];
          volatile float sv = v * 32768.f;
          long r = lrintf(sv);                      // cvRound: round half to even (default rounding mode)
          if (r < -32768) r = -32768;
          if (r > 32767) r = 32767;
          it[k1][k2] = (int)r;
          isum += (int)r;
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
        for (int k1 = 2; k1 < 4; ++k1)
          for (int k2 = 2; k2 < 4; ++k2) {
            if (it[k1][k2] < it[mk1][mk2]) { mk1 = k1; mk2 = k2; }
            else if (it[k1][k2] > it[Mk1][Mk2]) { Mk1 = k1; Mk2 = k2; }
          }
        if (diff < 0) it[Mk1][Mk2] -= diff; else it[mk1][mk2] -= diff;
      }
      for (int k = 0; k < 16; ++k) tab[(i * 32 + j) * 16 + k] = (short)it[k / 4][k % 4];
    }
}

// page [H,W,3] uint8 (host or device); minv [n][9]; sizes [n][2] = (w,h); rotate [n]; offsets [n] bytes into out (host or device)
inline void warp_crops(int device, const uint8_t* page, int H, int W, int n, const double* minv, const int32_t* sizes, const int32_t* rotate,
                       uint8_t* out, const long long* offsets, long long out_bytes, cudaStream_t st, int n_pages = 1, const int32_t* page_idx = nullptr) {
  RDB_CUDA(cudaSetDevice(device));
  if (n <= 0) return;
  static const std::vector<short> host_tab = [] { std::vector<short> t; build_cubic_tab(t); return t; }();   // thread-safe one-time init
  static short* dev_tab[64] = {};
  RDB_CHECK(device >= 0 && device < 64, "warp: device index");
  if (!dev_tab[device]) {
    RDB_CUDA(cudaMalloc(&dev_tab[device], host_tab.size() * sizeof(short)));
    RDB_CUDA(cudaMemcpy(dev_tab[device], host_tab.data(), host_tab.size() * sizeof(short), cudaMemcpyHostToDevice));
  }
  std::vector<WarpCrop> hc(n);
  long long max_px = 0;
  for (int i = 0; i < n; ++i) {
    WarpCrop& c = hc[i];
    for (int k = 0; k < 9; ++k) c.m[k] = minv[i * 9 + k];
    c.w = sizes[2 * i]; c.h = sizes[2 * i + 1]; c.rotate = rotate ? rotate[i] : 0; c.offset = offsets[i];
    const int pg = page_idx ? page_idx[i] : 0;
    RDB_CHECK(pg >= 0 && pg < n_pages, "warp: page index out of range");
    c.src_off = (long long)pg * H * W * 3;
    RDB_CHECK(c.w > 0 && c.h > 0, "warp: empty crop");
    RDB_CHECK(c.offset >= 0 && c.offset + (long long)c.w * c.h * 3 <= out_bytes, "warp: crop exceeds the output buffer");
    int bh0 = c.h < 16 ? c.h : 16;                    // WarpPerspectiveInvoker: BLOCK_SZ = 32
    int bw0 = 1024 / bh0 < c.w ? 1024 / bh0 : c.w;
    c.bw0 = bw0;
    const long long px = (long long)c.w * c.h;
    if (px > max_px) max_px = px;
  }
  const bool p_dev = is_device_ptr(page), o_dev = is_device_ptr(out);
  uint8_t* dp = const_cast<uint8_t*>(page);
  uint8_t* dout = out;
  const size_t page_b = (size_t)n_pages * H * W * 3;
  ScratchCarver sc{device_scratch(device, pad256(sizeof(WarpCrop) * n) + (p_dev ? 0 : pad256(page_b)) + (o_dev ? 0 : pad256((size_t)out_bytes)))};
  WarpCrop* dc = sc.take<WarpCrop>(n);
  if (!p_dev) { dp = sc.take<uint8_t>(page_b); RDB_CUDA(cudaMemcpyAsync(dp, page, page_b, cudaMemcpyHostToDevice, st)); }
  if (!o_dev) dout = sc.take<uint8_t>((size_t)out_bytes);
  WarpCrop* hp = reinterpret_cast<WarpCrop*>(pinned_scratch(sizeof(WarpCrop) * n));
  std::memcpy(hp, hc.data(), sizeof(WarpCrop) * n);
  RDB_CUDA(cudaMemcpyAsync(dc, hp, sizeof(WarpCrop) * n, cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((max_px + 255) / 256), (unsigned)n);
  warp_cubic_kernel<<<grid, 256, 0, st>>>(dp, H, W, dc, dev_tab[device], dout);
  RDB_LAUNCH_CHECK();
  if (!o_dev) RDB_CUDA(cudaMemcpyAsync(out, dout, (size_t)out_bytes, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));      // hc goes out of scope; the scratch is free for the next call
}

// ---------------------------------------------------------------------------------------------------------------------------
// resize_norm_img geometry on the GPU (rapidocr TextRecognizer.resize_norm_img as called from
// rapid_doc/model/ocr/rapid_ocr.py:423-440): every crop of a recognition batch is cv2.resize'd (INTER_LINEAR, uint8, bit-exact
// 11-bit fixed-point path, same arithmetic as resize_linear_u8_kernel) to height dh and its own width dst_w[i], written
// left-aligned into a [n][dh][dw_max][3] batch; the columns right of dst_w[i] are zero (the rec stem applies the zero pad after
// normalisation).  Crops are ragged: per-crop coefficient tables are built on the host with OpenCV's float32 rule.
struct ResizeCrop {
  long long src_off;    // byte offset of the crop in the packed source buffer
  int sw, sh, dw;       // source size, destination width
  int tab_off;          // offset (in ints) of this crop's tables: xi[dw] yi[dh] then shorts xa[2*dw] ya[2*dh] at tab_off_s
  int tab_off_s;
};

static __global__ void __launch_bounds__(256) resize_pack_kernel(const uint8_t* __restrict__ src, const ResizeCrop* __restrict__ crops, const int* __restrict__ itab,
                                                                 const short* __restrict__ stab, uint8_t* __restrict__ dst, int dh, int dw_max) {
  const ResizeCrop& c = crops[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dh * dw_max) return;
  const int x = idx % dw_max, y = idx / dw_max;
  uint8_t* o = dst + ((long long)blockIdx.y * dh * dw_max + idx) * 3;
  if (x >= c.dw) { o[0] = 0; o[1] = 0; o[2] = 0; return; }
  const int* xi = itab + c.tab_off; const int* yi = xi + c.dw;
  const short* xa = stab + c.tab_off_s; const short* ya = xa + 2 * c.dw;
  const int x0 = xi[x], x1 = min(x0 + 1, c.sw - 1);
  const int ys = yi[y];
  const int y0 = min(max(ys, 0), c.sh - 1), y1 = min(max(ys + 1, 0), c.sh - 1);
  const int a0 = xa[2 * x], a1 = xa[2 * x + 1], b0 = ya[2 * y], b1 = ya[2 * y + 1];
  const uint8_t* r0 = src + c.src_off + (long long)y0 * c.sw * 3;
  const uint8_t* r1 = src + c.src_off + (long long)y1 * c.sw * 3;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int S0 = r0[x0 * 3 + ch] * a0 + r0[x1 * 3 + ch] * a1;
    const int S1 = r1[x0 * 3 + ch] * a0 + r1[x1 * 3 + ch] * a1;
    const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
    o[ch] = (uint8_t)min(max(v, 0), 255);
  }
}

void linear_coeffs_cv(int dsize, int ssize, bool vertical, std::vector<int>& idx, std::vector<short>& ab);   // det.cu

// src: packed crops (host or device), crop i = [sizes[i][1]][sizes[i][0]][3] at src_offsets[i]; dst [n][dh][dw_max][3] (host or device)
inline void resize_pack_u8(int device, const uint8_t* src, long long src_bytes, int n, const long long* src_offsets, const int32_t* sizes, const int32_t* dst_w,
                           uint8_t* dst, int dh, int dw_max, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (n <= 0) return;
  std::vector<ResizeCrop> hc(n);
  std::vector<int> itab;
  std::vector<short> stab;
  std::vector<int> xi, yi;
  std::vector<short> xa, ya;
  for (int i = 0; i < n; ++i) {
    ResizeCrop& c = hc[i];
    c.src_off = src_offsets[i]; c.sw = sizes[2 * i]; c.sh = sizes[2 * i + 1]; c.dw = dst_w[i];
    RDB_CHECK(c.sw > 0 && c.sh > 0 && c.dw > 0 && c.dw <= dw_max, "resize_pack: bad crop geometry");
    RDB_CHECK(c.src_off >= 0 && c.src_off + (long long)c.sw * c.sh * 3 <= src_bytes, "resize_pack: crop exceeds the source buffer");
    linear_coeffs_cv(c.dw, c.sw, false, xi, xa);
    linear_coeffs_cv(dh, c.sh, true, yi, ya);
    c.tab_off = (int)itab.size(); c.tab_off_s = (int)stab.size();
    itab.insert(itab.end(), xi.begin(), xi.end()); itab.insert(itab.end(), yi.begin(), yi.end());
    stab.insert(stab.end(), xa.begin(), xa.end()); stab.insert(stab.end(), ya.begin(), ya.end());
  }
  const bool s_dev = is_device_ptr(src), d_dev = is_device_ptr(dst);
  const size_t dst_b = (size_t)n * dh * dw_max * 3;
  uint8_t* ds = const_cast<uint8_t*>(src); uint8_t* dd = dst;
  ScratchCarver sc{device_scratch(device, pad256(sizeof(ResizeCrop) * n) + pad256(sizeof(int) * itab.size()) + pad256(sizeof(short) * stab.size()) +
                                              (s_dev ? 0 : pad256((size_t)src_bytes)) + (d_dev ? 0 : pad256(dst_b)))};
  ResizeCrop* dc = sc.take<ResizeCrop>(n);
  int* dit = sc.take<int>(itab.size());
  short* dst_tab = sc.take<short>(stab.size());
  if (!s_dev) { ds = sc.take<uint8_t>((size_t)src_bytes); RDB_CUDA(cudaMemcpyAsync(ds, src, (size_t)src_bytes, cudaMemcpyHostToDevice, st)); }
  if (!d_dev) dd = sc.take<uint8_t>(dst_b);
  RDB_CUDA(cudaMemcpyAsync(dc, hc.data(), sizeof(ResizeCrop) * n, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dit, itab.data(), sizeof(int) * itab.size(), cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dst_tab, stab.data(), sizeof(short) * stab.size(), cudaMemcpyHostToDevice, st));
  dim3 grid((unsigned)((dh * dw_max + 255) / 256), (unsigned)n);
  resize_pack_kernel<<<grid, 256, 0, st>>>(ds, dc, dit, dst_tab, dd, dh, dw_max);
  RDB_LAUNCH_CHECK();
  if (!d_dev) RDB_CUDA(cudaMemcpyAsync(dst, dd, dst_b, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same resize for EVERY recognition batch of a window in one launch (cross-page batching): crop i goes to its
// own slot [dh][pitch_i][3] at byte dst_off[i] of one packed buffer (a batch = consecutive slots of equal pitch), columns
// >= dw_i of the slot are zeroed.  OpenCV's coefficient rule (float32, 11-bit fixed point) is evaluated in the kernel with
// explicitly rounded IEEE operations, so no per-crop tables travel and the call needs no synchronisation of its own.
struct ResizeSlot {
  long long src_off, dst_off;
  int sw, sh, dw, pitch;
};

__device__ __forceinline__ void cv_linear_coeff(int d, int dsize, int ssize, bool vertical, int* idx, int* a0, int* a1) {
  const double inv = __ddiv_rn((double)dsize, (double)ssize), scale = __ddiv_rn(1.0, inv);
  float fx = __double2float_rn(__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5));
  int s = (int)floorf(fx);
  fx = __fsub_rn(fx, (float)s);
  if (!vertical) {
    if (s < 0) { fx = 0.f; s = 0; }
    if (s >= ssize - 1) { fx = 0.f; s = ssize - 1; }
  }
  *idx = s;
  *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f));
  *a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
}

static __global__ void __launch_bounds__(256) resize_slots_kernel(const uint8_t* __restrict__ src, const ResizeSlot* __restrict__ slots, uint8_t* __restrict__ dst, int dh) {
  const ResizeSlot c = slots[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dh * c.pitch) return;
  const int x = idx % c.pitch, y = idx / c.pitch;
  uint8_t* o = dst + c.dst_off + (long long)idx * 3;
  if (x >= c.dw) { o[0] = 0; o[1] = 0; o[2] = 0; return; }
  int x0, a0, a1, ys, b0, b1;
  cv_linear_coeff(x, c.dw, c.sw, false, &x0, &a0, &a1);
  cv_linear_coeff(y, dh, c.sh, true, &ys, &b0, &b1);
  const int x1 = min(x0 + 1, c.sw - 1);
  const int y0 = min(max(ys, 0), c.sh - 1), y1 = min(max(ys + 1, 0), c.sh - 1);
  const uint8_t* r0 = src + c.src_off + (long long)y0 * c.sw * 3;
  const uint8_t* r1 = src + c.src_off + (long long)y1 * c.sw * 3;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const int S0 = r0[x0 * 3 + ch] * a0 + r0[x1 * 3 + ch] * a1;
    const int S1 = r1[x0 * 3 + ch] * a0 + r1[x1 * 3 + ch] * a1;
    const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
    o[ch] = (uint8_t)min(max(v, 0), 255);
  }
}

// src / dst: DEVICE buffers; the per-crop arrays are host arrays
inline void resize_pack_slots(int device, const uint8_t* src, long long src_bytes, int n, const long long* src_offsets, const int32_t* sizes,
                              const int32_t* dst_w, const long long* dst_offsets, const int32_t* dst_pitch, uint8_t* dst, long long dst_bytes, int dh,
                              cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (n <= 0) return;
  RDB_CHECK(is_device_ptr(src) && is_device_ptr(dst), "resize_pack_slots: src and dst must be device buffers");
  std::vector<ResizeSlot> hs(n);
  int max_pitch = 1;
  for (int i = 0; i < n; ++i) {
    ResizeSlot& c = hs[i];
    c.src_off = src_offsets[i]; c.dst_off = dst_offsets[i]; c.sw = sizes[2 * i]; c.sh = sizes[2 * i + 1]; c.dw = dst_w[i]; c.pitch = dst_pitch[i];
    RDB_CHECK(c.sw > 0 && c.sh > 0 && c.dw > 0 && c.dw <= c.pitch, "resize_pack_slots: bad crop geometry");
    RDB_CHECK(c.src_off >= 0 && c.src_off + (long long)c.sw * c.sh * 3 <= src_bytes, "resize_pack_slots: crop exceeds the source buffer");
    RDB_CHECK(c.dst_off >= 0 && c.dst_off + (long long)dh * c.pitch * 3 <= dst_bytes, "resize_pack_slots: slot exceeds the destination buffer");
    if (c.pitch > max_pitch) max_pitch = c.pitch;
  }
  ScratchCarver sc{device_scratch(device, pad256(sizeof(ResizeSlot) * n))};
  ResizeSlot* ds = sc.take<ResizeSlot>(n);
  ResizeSlot* hp = reinterpret_cast<ResizeSlot*>(pinned_scratch(sizeof(ResizeSlot) * n));
  std::memcpy(hp, hs.data(), sizeof(ResizeSlot) * n);
  RDB_CUDA(cudaMemcpyAsync(ds, hp, sizeof(ResizeSlot) * n, cudaMemcpyHostToDevice, st));
  for (int i0 = 0; i0 < n; i0 += 32768) {     // grid.y limit
    const int m = n - i0 < 32768 ? n - i0 : 32768;
    dim3 grid((unsigned)((dh * max_pitch + 255) / 256), (unsigned)m);
    resize_slots_kernel<<<grid, 256, 0, st>>>(src, ds + i0, dst, dh);
    RDB_LAUNCH_CHECK();
  }
  RDB_CUDA(cudaStreamSynchronize(st));      // the slot table lives in the shared scratch
}

}  // namespace rdb
