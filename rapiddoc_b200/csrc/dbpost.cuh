// DBPostProcess.box_score_fast on the GPU (rapidocr DBPostProcess as patched by rapid_doc/model/ocr/ocr_patch.py:223-241;
// upstream PaddleOCR db_postprocess.py box_score_fast): the mean of the prob map over cv2.fillPoly's raster of the
// (integer-truncated, bbox-shifted) mini-box quad.  With the score computed where the prob map lives, the fp32 map
// (4 MB per 1024x1024 page) never travels to the host — only the 1 MB bitmap does.
//
// cv2.fillPoly(mask, [quad], 1) (LINE_8, shift 0; OpenCV drawing.cpp CollectPolyEdges + FillEdgeCollection) is evaluated in
// CLOSED FORM per scanline instead of by walking an edge list, so rows are independent and a warp can take a row:
//   * outline: every edge is an 8-connected Bresenham line drawn left-to-right; with OpenCV's error term the minor-axis
//     offset after j major steps is k_j = floor((2*minor*j + major - 1) / (2*major)), so the pixels of an edge on row y are
//     one pixel (y-major edges) or one run [j_lo, j_hi] (x-major edges);
//   * interior: edges with y0 != y1 become 16.16 fixed-point scan edges x(y) = ((x_top << 16) + 0x8000) + (y - y_top) * dx,
//     dx = trunc((x1 - x0) * 65536 / (y1 - y0)); on row y the active edges (y_top <= y < y_bot) are sorted by x and paired;
//     a pair fills [x_lo >> 16, (x_hi - 0x8000) >> 16].
// The union of those <= 6 intervals is the mask row.  The function below is __host__ __device__: the same code rasterises on
// the CPU for the `-m "not gpu"` tests that pin it against cv2.fillPoly itself (rdb_debug_fill_quad).
// Valid when all four vertices lie inside the mask (always true for box_score_fast unless the mini box sticks out of the page;
// those quads are flagged and scored by the caller from a fetched ROI with cv2, so every score follows OpenCV's raster).
#pragma once
#include "engine.cuh"
#include "warp.cuh"

namespace rdb {

struct QuadRaster {
  int px[4], py[4];   // integer vertices in mask coordinates
  int mw, mh;         // mask size
};

__host__ __device__ inline long long floordiv_ll(long long a, long long b) {   // b > 0
  long long q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}
__host__ __device__ inline long long ceildiv_ll(long long a, long long b) { return -floordiv_ll(-a, b); }

// intervals [lo[i], hi[i]] (inclusive, clipped to the mask) of row y; returns their number (<= 6)
__host__ __device__ inline int quad_row_intervals(const QuadRaster& q, int y, int* lo, int* hi) {
  int n = 0;
  long long ex[4];   // active scan edges' x at row y
  int ne = 0;
  for (int e = 0; e < 4; ++e) {
    const int x0 = q.px[(e + 3) & 3], y0 = q.py[(e + 3) & 3], x1 = q.px[e], y1 = q.py[e];
    // ---- outline: Bresenham from the left end point
    int xl = x0, yl = y0, xr = x1, yr = y1;
    if (xr < xl) { xl = x1; yl = y1; xr = x0; yr = y0; }
    const int dx = xr - xl, dy = yr - yl, ady = dy < 0 ? -dy : dy, sy = dy < 0 ? -1 : 1;
    if (ady > dx) {                    // y-major: one pixel per row
      const int j = (y - yl) * sy;
      if (j >= 0 && j <= ady) {
        const int k = j == 0 ? 0 : (int)((2ll * dx * j + ady - 1) / (2ll * ady));
        lo[n] = hi[n] = xl + k; ++n;
      }
    } else {                           // x-major: a run per row
      const int k = (y - yl) * sy;
      if (k >= 0 && k <= ady) {
        long long jl = 0, jh = dx;
        if (ady > 0) {
          jl = ceildiv_ll(2ll * dx * k - dx + 1, 2ll * ady);
          jh = floordiv_ll(2ll * dx * k + dx, 2ll * ady);
          if (jl < 0) jl = 0;
          if (jh > dx) jh = dx;
        }
        if (jl <= jh) { lo[n] = xl + (int)jl; hi[n] = xl + (int)jh; ++n; }
      }
    }
    // ---- scan edge
    if (y0 != y1) {
      const long long c0 = ((long long)x0 << 16) + 32768, c1 = ((long long)x1 << 16) + 32768;
      const long long d = (c1 - c0) / (y1 - y0);       // C truncation toward zero
      if (y0 < y1) { if (y >= y0 && y < y1) ex[ne++] = c0 + (long long)(y - y0) * d; }
      else         { if (y >= y1 && y < y0) ex[ne++] = c1 + (long long)(y - y1) * d; }
    }
  }
  // sort the (<= 4) active edges by x, pair them up
  for (int a = 1; a < ne; ++a) {
    const long long v = ex[a];
    int b = a - 1;
    while (b >= 0 && ex[b] > v) { ex[b + 1] = ex[b]; --b; }
    ex[b + 1] = v;
  }
  for (int a = 0; a + 1 < ne; a += 2) {
    const long long x1 = ex[a] >> 16, x2 = (ex[a + 1] - 32768) >> 16;
    if (x1 < q.mw && x2 >= 0 && x2 >= x1) { lo[n] = (int)(x1 < 0 ? 0 : x1); hi[n] = (int)(x2 > q.mw - 1 ? q.mw - 1 : x2); ++n; }
  }
  return n;
}

struct ScoreQuad {
  float x[4], y[4];   // mini-box corners in prob-map coordinates (float32, as cv2.boxPoints returns them)
  int page;
};

// box_score_fast's geometry: clipped integer bounding box, float32 shift, truncation to int32 (np.floor / np.ceil / np.clip /
// astype(np.int32) in the reference's order).  Returns false if a vertex leaves the mask (caller's cv2 fallback).
__host__ __device__ inline bool score_quad_setup(const ScoreQuad& s, int W, int H, QuadRaster* q, int* xmin_o, int* ymin_o) {
  float fx0 = s.x[0], fx1 = s.x[0], fy0 = s.y[0], fy1 = s.y[0];
  for (int i = 1; i < 4; ++i) {
    fx0 = s.x[i] < fx0 ? s.x[i] : fx0; fx1 = s.x[i] > fx1 ? s.x[i] : fx1;
    fy0 = s.y[i] < fy0 ? s.y[i] : fy0; fy1 = s.y[i] > fy1 ? s.y[i] : fy1;
  }
  auto clampi = [](float v, int hi) { int i = (int)v; return i < 0 ? 0 : (i > hi ? hi : i); };
  const int xmin = clampi(floorf(fx0), W - 1), xmax = clampi(ceilf(fx1), W - 1);
  const int ymin = clampi(floorf(fy0), H - 1), ymax = clampi(ceilf(fy1), H - 1);
  q->mw = xmax - xmin + 1; q->mh = ymax - ymin + 1;
  bool inside = true;
  for (int i = 0; i < 4; ++i) {
#ifdef __CUDA_ARCH__
    const float sx = __fsub_rn(s.x[i], (float)xmin), sy = __fsub_rn(s.y[i], (float)ymin);
#else
    const float sx = s.x[i] - (float)xmin, sy = s.y[i] - (float)ymin;
#endif
    q->px[i] = (int)sx; q->py[i] = (int)sy;     // C cast = numpy astype(int32): truncation toward zero
    inside = inside && q->px[i] >= 0 && q->px[i] < q->mw && q->py[i] >= 0 && q->py[i] < q->mh;
  }
  *xmin_o = xmin; *ymin_o = ymin;
  return inside;
}

// one CTA per quad; warps take mask rows, lanes take columns; double accumulation (cv2.mean accumulates float32 in double)
static __global__ void __launch_bounds__(256) box_score_kernel(const float* __restrict__ prob, int H, int W, const ScoreQuad* __restrict__ quads,
                                                               double* __restrict__ scores, int* __restrict__ flags) {
  const ScoreQuad s = quads[blockIdx.x];
  QuadRaster q;
  int xmin, ymin;
  const bool inside = score_quad_setup(s, W, H, &q, &xmin, &ymin);
  if (!inside) {
    if (threadIdx.x == 0) { flags[blockIdx.x] = 1; scores[blockIdx.x] = 0.0; }
    return;
  }
  const float* base = prob + (size_t)s.page * H * W + (size_t)ymin * W + xmin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sum = 0.0;
  int cnt = 0;
  for (int y = warp; y < q.mh; y += 8) {
    int lo[6], hi[6];
    const int n = quad_row_intervals(q, y, lo, hi);
    if (n == 0) continue;
    int a = lo[0], b = hi[0];
    for (int i = 1; i < n; ++i) { a = lo[i] < a ? lo[i] : a; b = hi[i] > b ? hi[i] : b; }
    const float* row = base + (size_t)y * W;
    for (int x = a + lane; x <= b; x += 32) {
      bool in = false;
      for (int i = 0; i < n; ++i) in = in || (x >= lo[i] && x <= hi[i]);
      if (in) { sum += (double)row[x]; ++cnt; }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, o);
    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  }
  __shared__ double ssum[8];
  __shared__ int scnt[8];
  if (lane == 0) { ssum[warp] = sum; scnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int c = 0;
    for (int i = 0; i < 8; ++i) { t += ssum[i]; c += scnt[i]; }
    scores[blockIdx.x] = c ? t * (1.0 / (double)c) : 0.0;      // cv::mean: s * (1. / nz)
    flags[blockIdx.x] = 0;
  }
}

// prob [n,H,W] f32 (host or device); quads [m][4][2] f32, page_idx [m], scores [m] f64, flags [m] i32: host arrays
inline void db_box_scores(int device, const float* prob, int n, int H, int W, int m, const float* quads, const int32_t* page_idx, double* scores,
                          int32_t* flags, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (m <= 0) return;
  std::vector<ScoreQuad> hq(m);
  for (int i = 0; i < m; ++i) {
    for (int k = 0; k < 4; ++k) { hq[i].x[k] = quads[i * 8 + 2 * k]; hq[i].y[k] = quads[i * 8 + 2 * k + 1]; }
    hq[i].page = page_idx ? page_idx[i] : 0;
    RDB_CHECK(hq[i].page >= 0 && hq[i].page < n, "box_scores: page index out of range");
  }
  const bool p_dev = is_device_ptr(prob);
  const size_t prob_b = (size_t)n * H * W * sizeof(float);
  ScratchCarver sc{device_scratch(device, pad256(sizeof(ScoreQuad) * m) + pad256(sizeof(double) * m) + pad256(sizeof(int) * m) + (p_dev ? 0 : pad256(prob_b)))};
  ScoreQuad* dq = sc.take<ScoreQuad>(m);
  double* ds = sc.take<double>(m);
  int* df = sc.take<int>(m);
  const float* dp = prob;
  if (!p_dev) { float* t = sc.take<float>((size_t)n * H * W); RDB_CUDA(cudaMemcpyAsync(t, prob, prob_b, cudaMemcpyHostToDevice, st)); dp = t; }
  static const bool dbg = sw_debug("RDB_SCORE_TIMING") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, st); }
  HostCarver hcv{pinned_scratch(pad256(sizeof(ScoreQuad) * m) + pad256(sizeof(double) * m) + pad256(sizeof(int) * m))};
  ScoreQuad* hqp = hcv.take<ScoreQuad>(m);
  double* hsp = hcv.take<double>(m);
  int* hfp = hcv.take<int>(m);
  std::memcpy(hqp, hq.data(), sizeof(ScoreQuad) * m);
  RDB_CUDA(cudaMemcpyAsync(dq, hqp, sizeof(ScoreQuad) * m, cudaMemcpyHostToDevice, st));
  if (dbg) cudaEventRecord(e1, st);
  box_score_kernel<<<m, 256, 0, st>>>(dp, H, W, dq, ds, df);
  RDB_LAUNCH_CHECK();
  if (dbg) cudaEventRecord(e2, st);
  RDB_CUDA(cudaMemcpyAsync(hsp, ds, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaMemcpyAsync(hfp, df, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));
  std::memcpy(scores, hsp, sizeof(double) * m);
  std::memcpy(flags, hfp, sizeof(int) * m);
  if (dbg) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    fprintf(stderr, "[box_scores] m=%d h2d %.3f ms kernel %.3f ms\n", m, a, b);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  }
}

// host-only: the mask quad_row_intervals produces (tests pin it against cv2.fillPoly without a GPU)
inline void debug_fill_quad(const int32_t* pts_xy, int mw, int mh, uint8_t* mask) {
  QuadRaster q;
  for (int i = 0; i < 4; ++i) { q.px[i] = pts_xy[2 * i]; q.py[i] = pts_xy[2 * i + 1]; }
  q.mw = mw; q.mh = mh;
  for (int y = 0; y < mh; ++y) {
    int lo[6], hi[6];
    const int n = quad_row_intervals(q, y, lo, hi);
    for (int i = 0; i < n; ++i)
      for (int x = lo[i] < 0 ? 0 : lo[i]; x <= hi[i] && x < mw; ++x) mask[(size_t)y * mw + x] = 1;
  }
}

}  // namespace rdb
