// cv2.findContours(bitmap, RETR_LIST, CHAIN_APPROX_SIMPLE) for every page of a window, on host threads without the GIL
// ("serial single-thread OpenCV per page becomes the critical path" — measured here: 0.55 ms of Python-held time
// per page, and the cv2 call does not scale across Python threads).
//
// OpenCV is the third-party dependency of rapidocr's DBPostProcess.boxes_from_bitmap (called from
// rapid_doc/model/ocr/ocr_patch.py:236-239); the algorithm restated is Suzuki-Abe border following as implemented in
// modules/imgproc/src/contours.cpp (cvFindNextContour + icvFetchContour): the image is padded by one zero pixel, non-zero
// pixels become 1, a raster scan starts an OUTER border at a 0 -> 1 transition (pixel still unlabelled) and a HOLE border at
// a foreground -> 0 transition (foreground = value >= 1), the border is followed with the 8-neighbourhood search of
// icvFetchContour (clockwise start search from direction 4 / 0, counter-clockwise continuation), visited pixels are
// labelled 2 or -126 ("right edge" pixels), a point is emitted whenever the chain direction changes (CHAIN_APPROX_SIMPLE),
// and — RETR_LIST — contours are returned in REVERSE order of discovery.  tests/test_contours.py pins it against
// cv2.findContours itself (contour count, order, every point) on random, blob and text bitmaps.  Host-only code.
#include "../../include/rapiddoc_b200.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct rdb_contours {
  int n = 0;
  std::vector<std::vector<int32_t>> sizes;   // per page: points per contour, in cv2 order
  std::vector<std::vector<int32_t>> pts;     // per page: x, y, x, y, ... in cv2 order
};

namespace {

void trace_page(const uint8_t* bm, int h, int w, std::vector<int32_t>& sizes_out, std::vector<int32_t>& pts_out) {
  const int W = w + 2, H = h + 2;
  std::vector<signed char> buf((size_t)W * H, 0);
  for (int y = 0; y < h; ++y) {
    signed char* row = buf.data() + (size_t)(y + 1) * W + 1;
    const uint8_t* src = bm + (size_t)y * w;
    for (int x = 0; x < w; ++x) row[x] = src[x] ? 1 : 0;
  }
  const int step = W;
  int deltas[16] = {1, -step + 1, -step, -step - 1, -1, step - 1, step, step + 1, 1, -step + 1, -step, -step - 1, -1, step - 1, step, step + 1};
  static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  std::vector<int32_t> sz;                    // discovery order
  std::vector<int32_t> pts;
  for (int y = 1; y < H - 1; ++y) {
    signed char* img = buf.data() + (size_t)y * W;
    int prev = 0;
    for (int x = 1; x < W - 1; ++x) {
      if (img[x] == prev) {                   // run of equal pixels: skip 8 at a time (most of a page is background)
        const uint64_t pat = 0x0101010101010101ull * (uint8_t)prev;
        while (x + 8 < W - 1) {
          uint64_t v;
          std::memcpy(&v, img + x, 8);
          if (v != pat) break;
          x += 8;
        }
        while (x < W - 1 && img[x] == prev) ++x;
        if (x >= W - 1) break;
      }
      const int p = img[x];
      int is_hole = 0;
      bool start = true;
      if (!(prev == 0 && p == 1)) {
        if (p != 0 || prev < 1) start = false;
        else is_hole = 1;
      }
      if (start) {
        // ---- icvFetchContour from (x - is_hole, y), reported with the (-1, -1) offset of the padding
        signed char* i0 = img + x - is_hole;
        int px = x - is_hole - 1, py = y - 1;
        const size_t first = pts.size();
        int s_end = is_hole ? 0 : 4, s = s_end;
        signed char* i1;
        do {
          s = (s - 1) & 7;
          i1 = i0 + deltas[s];
        } while (*i1 == 0 && s != s_end);
        if (s == s_end) {                     // single pixel
          *i0 = (signed char)(2 | -128);
          pts.push_back(px); pts.push_back(py);
        } else {
          signed char* i3 = i0;
          signed char* i4 = nullptr;
          int prev_s = s ^ 4;
          for (;;) {
            s_end = s;
            if (s > 15) s = 15;
            while (s < 15) {
              i4 = i3 + deltas[++s];
              if (*i4 != 0) break;
            }
            s &= 7;
            if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = (signed char)(2 | -128);
            else if (*i3 == 1) *i3 = 2;
            if (s != prev_s) {
              pts.push_back(px); pts.push_back(py);
              prev_s = s;
            }
            px += dx[s]; py += dy[s];
            if (i4 == i0 && i3 == i1) break;
            i3 = i4;
            s = (s + 4) & 7;
          }
        }
        sz.push_back((int32_t)((pts.size() - first) / 2));
      }
      prev = img[x];                          // the (possibly re-labelled) value, as the scanner re-reads it on resume
    }
  }
  // RETR_LIST: every new contour is linked in FRONT of the list -> reverse discovery order
  sizes_out.resize(sz.size());
  pts_out.resize(pts.size());
  size_t off_out = 0;
  std::vector<size_t> offs(sz.size() + 1, 0);
  for (size_t i = 0; i < sz.size(); ++i) offs[i + 1] = offs[i] + (size_t)sz[i] * 2;
  for (size_t k = 0; k < sz.size(); ++k) {
    const size_t i = sz.size() - 1 - k;
    sizes_out[k] = sz[i];
    std::memcpy(pts_out.data() + off_out, pts.data() + offs[i], (size_t)sz[i] * 2 * sizeof(int32_t));
    off_out += (size_t)sz[i] * 2;
  }
}

}  // namespace

extern "C" {

int rdb_contours_trace(const uint8_t* bitmaps, int n, int hgt, int wid, int max_threads, rdb_contours_t** out) {
  if (!bitmaps || !out || n <= 0 || hgt <= 0 || wid <= 0) return RDB_ERR_INVALID;
  try {
    rdb_contours* c = new rdb_contours;
    c->n = n;
    c->sizes.resize(n);
    c->pts.resize(n);
    int nt = max_threads > 0 ? max_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > n) nt = n;
    std::atomic<int> next{0};
    auto work = [&] {
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        trace_page(bitmaps + (size_t)i * hgt * wid, hgt, wid, c->sizes[i], c->pts[i]);
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    *out = c;
    return RDB_OK;
  } catch (...) {
    return RDB_ERR_INVALID;
  }
}

int rdb_contours_counts(rdb_contours_t* c, int32_t* per_page, int64_t* total_contours, int64_t* total_points) {
  if (!c) return RDB_ERR_INVALID;
  int64_t tc = 0, tp = 0;
  for (int i = 0; i < c->n; ++i) {
    if (per_page) per_page[i] = (int32_t)c->sizes[i].size();
    tc += (int64_t)c->sizes[i].size();
    tp += (int64_t)c->pts[i].size() / 2;
  }
  if (total_contours) *total_contours = tc;
  if (total_points) *total_points = tp;
  return RDB_OK;
}

int rdb_contours_fetch(rdb_contours_t* c, int32_t* contour_sizes, int32_t* points_xy) {
  if (!c || !contour_sizes || !points_xy) return RDB_ERR_INVALID;
  size_t so = 0, po = 0;
  for (int i = 0; i < c->n; ++i) {
    std::memcpy(contour_sizes + so, c->sizes[i].data(), c->sizes[i].size() * sizeof(int32_t));
    std::memcpy(points_xy + po, c->pts[i].data(), c->pts[i].size() * sizeof(int32_t));
    so += c->sizes[i].size();
    po += c->pts[i].size();
  }
  return RDB_OK;
}

void rdb_contours_free(rdb_contours_t* c) { delete c; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------------
// sorted_boxes (twice: detector + caller) and merge_det_boxes for every page of a window, in float32 like the NumPy scalars the
// reference computes with (rapid_doc/utils/ocr_utils.py:105-127, 257-317 with merge_spans_to_line :16-38,
// _is_overlaps_y_exceeds_threshold :40-52, merge_overlapping_spans :219-254, calculate_is_angle :478-485).  The Python
// restatement (rapiddoc_b200/lines.py, pinned against the reference's functions) stays the specification; this is the same
// logic off the interpreter: 64 pages x 25 boxes cost 14 ms of GIL time in Python.
namespace {

struct Box { float p[8]; };   // x0,y0,x1,y1,x2,y2,x3,y3

void sort_boxes(std::vector<Box>& b) {
  std::stable_sort(b.begin(), b.end(), [](const Box& a, const Box& c) { return a.p[1] < c.p[1] || (a.p[1] == c.p[1] && a.p[0] < c.p[0]); });
  const int n = (int)b.size();
  for (int i = 0; i < n - 1; ++i)
    for (int j = i; j >= 0; --j) {
      const float d = b[j + 1].p[1] - b[j].p[1];
      if ((d < 0 ? -d : d) < 10.f && b[j + 1].p[0] < b[j].p[0]) std::swap(b[j], b[j + 1]);
      else break;
    }
}

bool is_angle(const Box& b) {
  const float height = ((b.p[7] - b.p[1]) + (b.p[5] - b.p[3])) / 2.f;
  const float dy = b.p[5] - b.p[1];
  return !(0.8f * height <= dy && dy <= 1.2f * height);
}

struct Span { float x0, y0, x1, y1; };

bool y_overlap_exceeds(const Span& a, const Span& b, float thr) {
  const float lo = a.y0 > b.y0 ? a.y0 : b.y0, hi = a.y1 < b.y1 ? a.y1 : b.y1;
  float overlap = hi - lo;
  if (!(overlap > 0.f)) overlap = 0.f;                       // max(0, .)
  const float h1 = a.y1 - a.y0, h2 = b.y1 - b.y0, mh = h1 < h2 ? h1 : h2;
  return mh > 0.f ? (overlap / mh) > thr : false;
}

void put_span(const Span& s, std::vector<Box>& out) {
  out.push_back(Box{{s.x0, s.y0, s.x1, s.y0, s.x1, s.y1, s.x0, s.y1}});
}

void merge_boxes(const std::vector<Box>& in, std::vector<Box>& out) {
  std::vector<Span> spans;
  std::vector<Box> angled;
  for (const Box& b : in) {
    if (is_angle(b)) angled.push_back(b);
    else spans.push_back(Span{b.p[0], b.p[1], b.p[2], b.p[5]});
  }
  std::stable_sort(spans.begin(), spans.end(), [](const Span& a, const Span& b) { return a.y0 < b.y0; });
  std::vector<std::vector<Span>> lines;
  for (const Span& s : spans) {
    if (lines.empty() || !y_overlap_exceeds(s, lines.back().back(), 0.6f)) lines.emplace_back();
    lines.back().push_back(s);
  }
  for (auto& line : lines) {
    float mnx = line[0].x0, mxx = line[0].x1, mny = line[0].y0, mxy = line[0].y1;
    for (const Span& s : line) {
      mnx = s.x0 < mnx ? s.x0 : mnx; mxx = s.x1 > mxx ? s.x1 : mxx;
      mny = s.y0 < mny ? s.y0 : mny; mxy = s.y1 > mxy ? s.y1 : mxy;
    }
    if ((mxx - mnx) > (mxy - mny) * 4.f) {
      std::stable_sort(line.begin(), line.end(), [](const Span& a, const Span& b) { return a.x0 < b.x0; });
      std::vector<Span> merged;
      for (const Span& s : line) {
        if (merged.empty() || merged.back().x1 < s.x0) merged.push_back(s);
        else {
          Span& m = merged.back();
          m.x0 = m.x0 < s.x0 ? m.x0 : s.x0; m.y0 = m.y0 < s.y0 ? m.y0 : s.y0;
          m.x1 = m.x1 > s.x1 ? m.x1 : s.x1; m.y1 = m.y1 > s.y1 ? m.y1 : s.y1;
        }
      }
      for (const Span& s : merged) put_span(s, out);
    } else {
      for (const Span& s : line) put_span(s, out);
    }
  }
  for (const Box& b : angled) out.push_back(b);
}

}  // namespace

extern "C" int rdb_lines_sort_merge(const float* boxes, const int32_t* page_offsets, int n_pages, int merge, float* out, int32_t* out_offsets) {
  if (!boxes || !page_offsets || !out || !out_offsets || n_pages < 0) return RDB_ERR_INVALID;
  try {
    int total = 0;
    out_offsets[0] = 0;
    for (int p = 0; p < n_pages; ++p) {
      std::vector<Box> b(page_offsets[p + 1] - page_offsets[p]);
      if (!b.empty()) std::memcpy(b.data(), boxes + (size_t)page_offsets[p] * 8, b.size() * sizeof(Box));
      sort_boxes(b);          // TextDetector.sorted_boxes
      sort_boxes(b);          // the caller's sorted_boxes on the sorted result (rapid_ocr.py:372 / analyze_utils.py:194)
      std::vector<Box> res;
      if (merge) merge_boxes(b, res); else res = b;
      if (!res.empty()) std::memcpy(out + (size_t)total * 8, res.data(), res.size() * sizeof(Box));
      total += (int)res.size();
      out_offsets[p + 1] = total;
    }
    return total;
  } catch (...) {
    return RDB_ERR_INVALID;
  }
}
