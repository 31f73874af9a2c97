// PP-DocLayout post-processing on the GPU for a whole window of pages:
//   class-aware greedy NMS          rapid_layout_self/model_handler/pp_doclayout/post_process.py:948-979 (iou :925-946)
//   containment relations           post_process.py:981-1022 (is_contained, check_containment)
// Both are O(n^2) Python loops per page in the reference (n <= 300 boxes); here one CTA handles one page and the window's pages
// run side by side.  The arithmetic is the reference's float32 arithmetic under NumPy-2 promotion rules (np.float32 scalars
// with weak Python ints / floats): every operation is an explicitly rounded float32 op (no FMA contraction), thresholds are
// compared as float32, so keep-sets and flags are bit-identical to the Python loops (tests/test_layout.py).
#pragma once
#include "engine.cuh"
#include "warp.cuh"

namespace rdb {

// boxes: rows of `stride` floats [cls, score, x1, y1, x2, y2, ...]
__device__ __forceinline__ float nms_iou(const float* a, const float* b) {
  const float x1 = fmaxf(a[2], b[2]), y1 = fmaxf(a[3], b[3]);
  const float x2 = fminf(a[4], b[4]), y2 = fminf(a[5], b[5]);
  const float w = __fadd_rn(__fsub_rn(x2, x1), 1.f), h = __fadd_rn(__fsub_rn(y2, y1), 1.f);
  const float inter = __fmul_rn(w > 0.f ? w : 0.f, h > 0.f ? h : 0.f);                    // max(0, .) * max(0, .)
  const float aa = __fmul_rn(__fadd_rn(__fsub_rn(a[4], a[2]), 1.f), __fadd_rn(__fsub_rn(a[5], a[3]), 1.f));
  const float ab = __fmul_rn(__fadd_rn(__fsub_rn(b[4], b[2]), 1.f), __fadd_rn(__fsub_rn(b[5], b[3]), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}

// one CTA per page.  order: indices sorted by descending score (np.argsort(scores)[::-1], made on the host so that ties fall
// exactly as NumPy's sort leaves them).  keep: kept indices in selection order, keep_n: their number.
static __global__ void __launch_bounds__(256) layout_nms_kernel(const float* __restrict__ boxes, int stride, const int* __restrict__ order,
                                                                const int* __restrict__ offsets, float iou_same, float iou_diff,
                                                                int* __restrict__ keep, int* __restrict__ keep_n) {
  const int p = blockIdx.x, base = offsets[p], n = offsets[p + 1] - base;
  extern __shared__ unsigned char alive[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) alive[i] = 1;
  __shared__ int kept;
  if (threadIdx.x == 0) kept = 0;
  __syncthreads();
  for (int t = 0; t < n; ++t) {
    if (!alive[t]) continue;                         // uniform: every thread reads the same flag after the barrier below
    const int cur = order[base + t];
    const float* cb = boxes + (size_t)(base + cur) * stride;
    if (threadIdx.x == 0) keep[base + kept++] = cur;
    for (int j = t + 1 + threadIdx.x; j < n; j += blockDim.x) {
      if (!alive[j]) continue;
      const float* ob = boxes + (size_t)(base + order[base + j]) * stride;
      const float thr = cb[0] == ob[0] ? iou_same : iou_diff;
      if (!(nms_iou(cb + 0, ob + 0) < thr)) alive[j] = 0;       // kept only `if iou_value < threshold`
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) keep_n[p] = kept;
}

__device__ __forceinline__ bool box_contained(const float* a, const float* b) {   // is_contained(box1 = a, box2 = b)
  const float area = __fmul_rn(__fsub_rn(a[4], a[2]), __fsub_rn(a[5], a[3]));
  const float xi1 = fmaxf(a[2], b[2]), yi1 = fmaxf(a[3], b[3]), xi2 = fminf(a[4], b[4]), yi2 = fminf(a[5], b[5]);
  const float iw = __fsub_rn(xi2, xi1), ih = __fsub_rn(yi2, yi1);
  const float inter = __fmul_rn(iw > 0.f ? iw : 0.f, ih > 0.f ? ih : 0.f);
  const float r = area > 0.f ? __fdiv_rn(inter, area) : 0.f;
  return r >= 0.9f;
}

// check_containment(boxes, formula_index, category_index, mode): mode 0 = plain, 1 = "large", 2 = "small"; index < 0 = None
static __global__ void __launch_bounds__(256) layout_containment_kernel(const float* __restrict__ boxes, int stride, const int* __restrict__ offsets,
                                                                        int formula_index, int category_index, int mode,
                                                                        int* __restrict__ contains_other, int* __restrict__ contained_by_other) {
  const int p = blockIdx.x, base = offsets[p], n = offsets[p + 1] - base;
  for (int k = threadIdx.x; k < n * n; k += blockDim.x) {
    const int i = k / n, j = k % n;
    if (i == j) continue;
    const float* bi = boxes + (size_t)(base + i) * stride;
    const float* bj = boxes + (size_t)(base + j) * stride;
    if (formula_index >= 0 && bi[0] == (float)formula_index && bj[0] != (float)formula_index) continue;
    bool test;
    if (category_index >= 0 && mode != 0) test = (mode == 1 && bj[0] == (float)category_index) || (mode == 2 && bi[0] == (float)category_index);
    else test = true;
    if (test && box_contained(bi, bj)) {
      contained_by_other[base + i] = 1;
      contains_other[base + j] = 1;
    }
  }
}

inline void layout_nms(int device, const float* boxes, int stride, const int32_t* order, const int32_t* offsets, int pages, float iou_same, float iou_diff,
                       int32_t* keep, int32_t* keep_n, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (pages <= 0) return;
  const int total = offsets[pages];
  int max_n = 0;
  for (int p = 0; p < pages; ++p) max_n = std::max(max_n, offsets[p + 1] - offsets[p]);
  RDB_CHECK(max_n <= 40000, "layout_nms: too many boxes on one page");
  ScratchCarver sc{device_scratch(device, pad256((size_t)total * stride * 4) + 2 * pad256((size_t)total * 4) + 2 * pad256((size_t)(pages + 1) * 4))};
  float* db = sc.take<float>((size_t)total * stride);
  int* dord = sc.take<int>(total);
  int* dkeep = sc.take<int>(total);
  int* doff = sc.take<int>(pages + 1);
  int* dkn = sc.take<int>(pages + 1);
  RDB_CUDA(cudaMemcpyAsync(db, boxes, (size_t)total * stride * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(dord, order, (size_t)total * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(doff, offsets, (size_t)(pages + 1) * 4, cudaMemcpyHostToDevice, st));
  layout_nms_kernel<<<pages, 256, max_n + 16, st>>>(db, stride, dord, doff, iou_same, iou_diff, dkeep, dkn);
  RDB_LAUNCH_CHECK();
  RDB_CUDA(cudaMemcpyAsync(keep, dkeep, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaMemcpyAsync(keep_n, dkn, (size_t)pages * 4, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));
}

inline void layout_containment(int device, const float* boxes, int stride, const int32_t* offsets, int pages, int formula_index, int category_index,
                               int mode, int32_t* contains_other, int32_t* contained_by_other, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (pages <= 0) return;
  const int total = offsets[pages];
  if (total == 0) return;
  ScratchCarver sc{device_scratch(device, pad256((size_t)total * stride * 4) + 2 * pad256((size_t)total * 4) + pad256((size_t)(pages + 1) * 4))};
  float* db = sc.take<float>((size_t)total * stride);
  int* dc = sc.take<int>(total);
  int* dcb = sc.take<int>(total);
  int* doff = sc.take<int>(pages + 1);
  RDB_CUDA(cudaMemcpyAsync(db, boxes, (size_t)total * stride * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemcpyAsync(doff, offsets, (size_t)(pages + 1) * 4, cudaMemcpyHostToDevice, st));
  RDB_CUDA(cudaMemsetAsync(dc, 0, (size_t)total * 4, st));
  RDB_CUDA(cudaMemsetAsync(dcb, 0, (size_t)total * 4, st));
  layout_containment_kernel<<<pages, 256, 0, st>>>(db, stride, doff, formula_index, category_index, mode, dc, dcb);
  RDB_LAUNCH_CHECK();
  RDB_CUDA(cudaMemcpyAsync(contains_other, dc, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaMemcpyAsync(contained_by_other, dcb, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
  RDB_CUDA(cudaStreamSynchronize(st));
}

// TableLabelDecode's argmax / max over the structure vocabulary (rapid_table_self/table_structure/pp_structure/post_process.py:39-80):
// probs [B,T,V] f32 -> idx [B,T] i32 (first maximum, as np.argmax), val [B,T] f32.  One warp per (b, t) row.
static __global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ x, long long rows, int V, int* __restrict__ idx, float* __restrict__ val) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* row = x + r * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  bool nan = false;
  int nan_i = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float f = row[v];
    if (f != f) { if (!nan) { nan = true; nan_i = v; } continue; }       // np.argmax returns the first NaN if there is one
    if (f > best || (f == best && v < bi)) { best = f; bi = v; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_down_sync(0xffffffffu, best, o);
    const int oi = __shfl_down_sync(0xffffffffu, bi, o);
    const int on = __shfl_down_sync(0xffffffffu, nan_i, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    nan_i = on < nan_i ? on : nan_i;
  }
  if (lane == 0) {
    if (nan_i != 0x7fffffff) { idx[r] = nan_i; val[r] = row[nan_i]; }
    else { idx[r] = bi == 0x7fffffff ? 0 : bi; val[r] = best; }
  }
}

inline void argmax_rows(int device, const float* x, long long rows, int V, int32_t* idx, float* val, cudaStream_t st) {
  RDB_CUDA(cudaSetDevice(device));
  if (rows <= 0) return;
  const bool x_dev = is_device_ptr(x), o_dev = is_device_ptr(idx);
  RDB_CHECK(o_dev == is_device_ptr(val), "argmax_rows: idx and val must live on the same side");
  ScratchCarver sc{device_scratch(device, (x_dev ? 0 : pad256((size_t)rows * V * 4)) + (o_dev ? 0 : 2 * pad256((size_t)rows * 4)))};
  const float* dx = x;
  if (!x_dev) { float* t = sc.take<float>((size_t)rows * V); RDB_CUDA(cudaMemcpyAsync(t, x, (size_t)rows * V * 4, cudaMemcpyHostToDevice, st)); dx = t; }
  int* di = o_dev ? idx : sc.take<int>(rows);
  float* dv = o_dev ? val : sc.take<float>(rows);
  argmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(dx, rows, V, di, dv);
  RDB_LAUNCH_CHECK();
  if (!o_dev) {
    RDB_CUDA(cudaMemcpyAsync(idx, di, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
    RDB_CUDA(cudaMemcpyAsync(val, dv, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
  }
  if (!x_dev || !o_dev) RDB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace rdb
