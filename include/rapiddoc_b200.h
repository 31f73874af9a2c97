/* rapiddoc_b200 — C-ABI of the B200-native OCR hot path that drops in under RapidDoc.
 *
 * RapidDoc (the reference) is pure Python and has no FFI of its own; the seam it already
 * swaps engines at is the `InferSession.__call__(np.ndarray) -> np.ndarray` protocol
 * (rapidocr InferSession, replaced by rapid_doc/model/ocr/ocr_patch.py:95-105 with
 * rapid_doc/model/ocr/torch.py:33-200).  Every entry point below names the reference
 * interface (file:line under /root/reference) it replaces.  INTEGRATION.md shows the
 * ctypes binding a RapidDoc maintainer would add.
 *
 * Conventions: plain pointers + sizes, no torch types.  Every data pointer may be a HOST
 * pointer (pageable or pinned) or a DEVICE pointer on the engine's GPU; the library
 * detects which (cudaPointerGetAttributes).  With host pointers the call copies in,
 * computes, copies out and returns after the results are in host memory.  With device
 * pointers the work is enqueued on `stream` (a cudaStream_t, may be NULL = legacy default
 * stream) and the call returns without synchronising.  Handles are not thread-safe;
 * distinct handles are independent.  All functions return RDB_OK (0) or a negative code;
 * rdb_last_error() gives the message of the calling thread's last failure.
 */
#ifndef RAPIDDOC_B200_H_
#define RAPIDDOC_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDB_OK 0
#define RDB_ERR_INVALID (-1)
#define RDB_ERR_CUDA (-2)
#define RDB_ERR_NO_DEVICE (-3)

/* arithmetic mode of the pointwise-conv / linear engine */
#define RDB_PREC_FP32 0 /* fp32 storage + fp32 SIMT math: the exact-parity mode            */
#define RDB_PREC_FP16 1 /* fp16 NHWC storage, tcgen05 (fp16 x fp16 -> fp32 TMEM) GEMMs      */
#define RDB_PREC_TF32 2 /* rdb_op_gemm only: fp32 storage, tcgen05 kind::tf32 math          */

typedef struct rdb_det rdb_det_t; /* PP-OCRv6-small DBNet detector on one GPU  */
typedef struct rdb_rec rdb_rec_t; /* PP-OCRv6-small LightSVTR/CTC recogniser   */

int rdb_version(void);
const char* rdb_last_error(void);
int rdb_device_count(void);
/* pinned host memory for the callers' page / crop buffers */
int rdb_pinned_alloc(size_t nbytes, void** out);
int rdb_pinned_free(void* p);

/* ---- text detection -------------------------------------------------------------------
 * weights: "RDW1" blob (rapiddoc_b200/weights.py), BN folded.  Replaces building
 * BaseModel + load_state_dict + .to(device): rapid_doc/model/ocr/torch.py:78-89,156-168. */
int rdb_det_create(const void* weights, size_t nbytes, int device, int precision, rdb_det_t** out);
void rdb_det_destroy(rdb_det_t* h);

/* InferSession seam for det.  x [n,3,h,w] f32 (h,w multiples of 32) -> prob [n,1,h,w] f32.
 * Replaces TorchInferSession.__call__ for "maps": rapid_doc/model/ocr/torch.py:171-184
 * (network: backbones/rec_lcnetv4.py:283-305, necks/db_fpn.py:366-415,
 * heads/det_db_head.py:103-147). */
int rdb_det_infer_f32(rdb_det_t* h, const float* x, int n, int hgt, int wid, float* prob, void* stream);

/* Facade seam for det: uint8 BGR HWC pages (already resized to h,w multiples of 32) ->
 * prob map + DB bitmap.  Fuses DetPreProcess' normalisation ((v/255-mean)/std, rapidocr
 * ch_ppocr_det/utils.py as pinned by ocr_patch.py:33-40, rapid_ocr.py:59-62), the forward
 * above, and DBPostProcess' binarise + 2x2 dilate (ocr_patch.py:223-235).
 * prob may be NULL (bitmap only) and bitmap may be NULL. */
int rdb_det_infer_u8(rdb_det_t* h, const uint8_t* pages, int n, int hgt, int wid, const float mean[3],
                     const float stdv[3], float thresh, int use_dilation, float* prob, uint8_t* bitmap,
                     void* stream);

/* rdb_det_infer_u8 with DetPreProcess' resize on the GPU too: pages [n,src_h,src_w,3] are resized (cv2 INTER_LINEAR,
 * bit-exact) to hgt x wid (multiples of 32, chosen by the caller with DetPreProcess' limit_side_len rule), then as above. */
int rdb_det_infer_u8_resize(rdb_det_t* h, const uint8_t* pages, int n, int src_h, int src_w, int hgt, int wid,
                            const float mean[3], const float stdv[3], float thresh, int use_dilation, float* prob,
                            uint8_t* bitmap, void* stream);

/* DetPreProcess' cv2.resize(img, (dw, dh)) (INTER_LINEAR, uint8 HWC), bit-exact with OpenCV's fixed-point path
 * (rapidocr ch_ppocr_det/utils.py DetPreProcess as pinned by rapid_doc/model/ocr/ocr_patch.py:33-40).
 * src [n,sh,sw,3] -> dst [n,dh,dw,3]; host or device pointers. */
int rdb_resize_linear_u8(int device, const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw,
                         void* stream);

/* get_rotate_crop_image for n text-line quads of ONE page (rapid_doc/utils/ocr_utils.py:494-537, called from
 * rapid_doc/model/ocr/rapid_ocr.py ocr()/__call__ and backend/pipeline/analyze_utils.py): cv2.warpPerspective(INTER_CUBIC,
 * BORDER_REPLICATE), bit-exact with OpenCV's fixed-point remap.  page [hgt,wid,3] uint8; minv [n][9] = the dst->src
 * homography (cv2.invert(cv2.getPerspectiveTransform(quad, rect))[1], row-major double); sizes [n][2] = (crop_w, crop_h) of the
 * warp output; rotate [n] (may be NULL): 1 = store np.rot90(crop) (the reference rotates crops with h/w >= 2); crop i is
 * written as a contiguous [h][w][3] (or [w][h][3]) block at byte offsets[i] of out (out_bytes long).  Host or device pointers. */
int rdb_warp_crops(int device, const uint8_t* page, int hgt, int wid, int n, const double* minv, const int32_t* sizes,
                   const int32_t* rotate, uint8_t* out, const int64_t* offsets, int64_t out_bytes, void* stream);

/* resize_norm_img geometry for one recognition batch (rapidocr TextRecognizer.resize_norm_img as driven by
 * rapid_doc/model/ocr/rapid_ocr.py:423-440): crop i ([sizes[i][1]][sizes[i][0]][3] uint8 at byte src_offsets[i] of the packed
 * buffer src, e.g. the output of rdb_warp_crops) is cv2.resize'd (INTER_LINEAR, bit-exact) to hgt x dst_w[i] and written
 * left-aligned into dst [n][hgt][wid_max][3]; columns >= dst_w[i] are zero.  The (x/255-0.5)/0.5 normalisation and the zero pad
 * after it are applied by rdb_rec_infer_u8 (valid_w = dst_w).  Host or device pointers. */
int rdb_resize_pack_u8(int device, const uint8_t* src, int64_t src_bytes, int n, const int64_t* src_offsets, const int32_t* sizes,
                       const int32_t* dst_w, uint8_t* dst, int hgt, int wid_max, void* stream);

/* get_rotate_crop_image for the quads of a whole WINDOW of same-size pages in one launch (cross-page batching of
 * rapid_doc/backend/pipeline/analyze_utils.py:193-212 -> utils/ocr_utils.py:494-537): as rdb_warp_crops, with
 * pages [n_pages,hgt,wid,3] and page_idx [n] selecting each quad's page. */
int rdb_warp_crops_batch(int device, const uint8_t* pages, int n_pages, int hgt, int wid, int n, const int32_t* page_idx,
                         const double* minv, const int32_t* sizes, const int32_t* rotate, uint8_t* out, const int64_t* offsets,
                         int64_t out_bytes, void* stream);

/* resize_norm_img geometry for EVERY recognition batch of a window in one launch (the batches of
 * rapid_doc/model/ocr/rapid_ocr.py:423-440 keep their own padded width): crop i is cv2.resize'd (INTER_LINEAR, bit-exact) to
 * hgt x dst_w[i] into its slot [hgt][dst_pitch[i]][3] at byte dst_offsets[i] of dst; slot columns >= dst_w[i] are zero.
 * src / dst are DEVICE buffers, the per-crop arrays host arrays. */
int rdb_resize_pack_slots(int device, const uint8_t* src, int64_t src_bytes, int n, const int64_t* src_offsets, const int32_t* sizes,
                          const int32_t* dst_w, const int64_t* dst_offsets, const int32_t* dst_pitch, uint8_t* dst, int64_t dst_bytes,
                          int hgt, void* stream);

/* DBPostProcess.box_score_fast for m mini-box quads over n same-size prob maps (rapidocr DBPostProcess as patched by
 * rapid_doc/model/ocr/ocr_patch.py:223-241; upstream boxes_from_bitmap -> box_score_fast): scores[i] = cv2.mean of
 * prob[page_idx[i]] over cv2.fillPoly's raster of the bbox-shifted, int32-truncated quad (same pixels as OpenCV; the
 * float64 sum differs from cv2.mean only in summation order, <= 1e-12).  prob [n,hgt,wid] f32 host or DEVICE (the point:
 * the map stays on the GPU); quads [m,4,2] f32, page_idx [m] (NULL = page 0), scores [m] f64, flags [m] i32 are host arrays.
 * flags[i] = 1: a vertex of quad i lies outside its clipped bounding box (mini box sticking out of the page) — not scored
 * here; the caller scores those from an ROI with cv2 so that every score follows OpenCV's raster rule. */
int rdb_db_box_scores(int device, const float* prob, int n, int hgt, int wid, int m, const float* quads, const int32_t* page_idx,
                      double* scores, int32_t* flags, void* stream);

/* host-only diagnostic: the cv2.fillPoly(mask, [quad], 1) raster rdb_db_box_scores uses, for a quad whose vertices lie
 * inside the mw x mh mask; pts_xy [4][2] int32, mask [mh][mw] uint8 (zero-initialised by the caller). */
int rdb_debug_fill_quad(const int32_t* pts_xy, int mw, int mh, uint8_t* mask);

/* device-memory hygiene: each engine caches its call-scoped buffers per compute lane (best-fit reuse); idle cached bytes
 * beyond the cap go back to the driver at the end of a call (default 12 GiB per lane). */
int rdb_det_set_pool_cap_bytes(rdb_det_t* h, size_t bytes);
int rdb_rec_set_pool_cap_bytes(rdb_rec_t* h, size_t bytes);
long long rdb_det_pool_bytes(rdb_det_t* h);
long long rdb_rec_pool_bytes(rdb_rec_t* h);

/* cv2.findContours(bitmap, RETR_LIST, CHAIN_APPROX_SIMPLE) for n same-size {0, non-zero} uint8 bitmaps on host threads (no GIL):
 * the contour source of DBPostProcess.boxes_from_bitmap (rapidocr, called from rapid_doc/model/ocr/ocr_patch.py:236-239).
 * Same contours, same order, same points as OpenCV (Suzuki-Abe border following, reverse discovery order).  Host-only.
 *   rdb_contours_trace   traces every page (max_threads <= 0: all hardware threads) and returns a handle
 *   rdb_contours_counts  per_page [n] contour counts, totals
 *   rdb_contours_fetch   contour_sizes [total_contours] (page after page, cv2 order), points_xy [total_points][2] int32 */
typedef struct rdb_contours rdb_contours_t;
int rdb_contours_trace(const uint8_t* bitmaps, int n, int hgt, int wid, int max_threads, rdb_contours_t** out);
int rdb_contours_counts(rdb_contours_t* c, int32_t* per_page, int64_t* total_contours, int64_t* total_points);
int rdb_contours_fetch(rdb_contours_t* c, int32_t* contour_sizes, int32_t* points_xy);
void rdb_contours_free(rdb_contours_t* c);

/* sorted_boxes (applied twice: TextDetector + caller) and, with merge != 0, merge_det_boxes for every page of a window
 * (rapid_doc/utils/ocr_utils.py:105-127, 257-317), float32 like the NumPy scalars of the reference.  boxes [total][4][2] f32
 * delimited by page_offsets [n_pages+1]; out (capacity = total boxes) / out_offsets [n_pages+1]; returns the output count.
 * Host-only. */
int rdb_lines_sort_merge(const float* boxes, const int32_t* page_offsets, int n_pages, int merge, float* out, int32_t* out_offsets);

/* DBPostProcess binarise (+ optional cv2.dilate 2x2) on an existing prob map [n,h,w]:
 * rapid_doc/model/ocr/ocr_patch.py:228-235. */
int rdb_db_bitmap(int device, const float* prob, int n, int hgt, int wid, float thresh, int use_dilation,
                  uint8_t* bitmap, void* stream);

/* DBPostProcess.unclip for one quad (host-only, integer Clipper offset JT_ROUND /
 * ET_CLOSEDPOLYGON): rapid_doc/model/ocr/ocr_patch.py:161-172 (pyclipper).  box: 4 (x,y)
 * pairs; distance = area*ratio/perimeter is computed by the caller.  Writes up to
 * max_pts points to out_xy, returns the count (or a negative error). */
int rdb_clipper_offset(const double* box_xy, int n_pts, double distance, int64_t* out_xy, int max_pts);

/* rdb_clipper_offset for m quads in one call (every box of a window): boxes_xy [m][4][2], distances [m]; the polygons are
 * written back to back into out_xy (max_pts_total points), counts[i] = points of polygon i; returns the total. */
int rdb_clipper_offset_batch(const double* boxes_xy, int m, const double* distances, int64_t* out_xy, int max_pts_total,
                             int32_t* counts);

/* ---- layout / table post-processing -------------------------------------------------------------------
 * PP-DocLayout class-aware greedy NMS for a window of pages in one launch (one CTA per page): replaces the Python loop
 * `nms` rapid_doc/model/layout/rapid_layout_self/model_handler/pp_doclayout/post_process.py:948-979 (+ `iou` :925-946, the +1
 * pixel convention).  boxes [total][stride] f32 rows = [cls, score, x1, y1, x2, y2, ...]; offsets [pages+1] delimit the pages;
 * order [total] = per page np.argsort(scores)[::-1] (made by the caller, so ties fall as NumPy leaves them); keep [total]
 * receives per page (at the page's offset) the kept row indices in selection order, keep_n [pages] their number.  All host arrays.
 * float32 arithmetic with the reference's operation order: keep-sets are bit-identical. */
int rdb_layout_nms(int device, const float* boxes, int stride, const int32_t* order, const int32_t* offsets, int pages, float iou_same,
                   float iou_diff, int32_t* keep, int32_t* keep_n, void* stream);

/* `check_containment` post_process.py:996-1022 (+ `is_contained` :981-994) for a window of pages: formula_index /
 * category_index < 0 = None; mode 0 = None, 1 = "large", 2 = "small".  contains_other / contained_by_other [total] i32. */
int rdb_layout_containment(int device, const float* boxes, int stride, const int32_t* offsets, int pages, int formula_index,
                           int category_index, int mode, int32_t* contains_other, int32_t* contained_by_other, void* stream);

/* argmax / max over the last axis of x [rows][V] f32 (first maximum wins, as np.argmax): the reduction TableLabelDecode.decode
 * starts with, rapid_doc/model/table/rapid_table_self/table_structure/pp_structure/post_process.py:49-50 (structure_probs
 * [B,T,50] -> idx, prob).  x host or device; idx / val host or device (both on the same side). */
int rdb_argmax_rows(int device, const float* x, long long rows, int vocab, int32_t* idx, float* val, void* stream);

/* ---- formula engine ops ------------------------------------------------------------------------------
 * PP-FormulaNet_plus = PPHGNetV2-B6 encoder (rapid_doc/model/formula/rapid_formula_self/networks/backbones/rec_pphgnetv2.py:
 * 858-1207, 1587-1642) + MBart decoder with greedy generation (networks/heads/rec_unimernet_head.py:502-748,931-976 and
 * networks/heads/rec_ppformulanet_head.py:400-632,1052-1171).  The reference runs it as a torch module behind
 * `InferSession.__call__` (rapid_formula_self/inference_engine/torch.py:25-131); here the host side
 * (rapiddoc_b200/formula.py) walks the network and enqueues these ops on DEVICE buffers (NHWC activations = [pixels, C]
 * matrices with a row pitch).  All pointers are device pointers; nothing synchronises; prec = RDB_PREC_FP32 (fp32 SIMT,
 * exact-parity mode) or RDB_PREC_FP16 (fp16 storage, tcgen05 GEMMs).  act: 0 none, 1 ReLU, 2 GELU(erf), 6 HardSwish (fp32 path).
 * rdb_op_gemm also takes RDB_PREC_TF32: fp32 buffers multiplied on the tensor cores as TF32 (10-bit mantissa inputs, fp32
 * accumulation; act 0 / 1 / 6, no residual) — the fast mode of the ONNX CNN executor; rows <= 32 fall back to the fp32 path. */
const char* rdb_ops_last_error(void);
/* out[M, c_off : c_off+N] (row pitch ldc) = act(A[M,K] (pitch lda) * W[N,K]^T + bias) (+ res [M,N] pitch ldr): every conv1x1 /
 * nn.Linear, and every dense k x k conv after rdb_op_im2col.  W fp32 (prec 0) or fp16 (prec 1); bias fp32.
 * fp32 with M <= 32 rows (one decode step of a batch) runs a weight-streaming kernel; there out_step (device counter, may be
 * NULL) shifts the output by *out_step * out_step_stride elements — the k / v projections append to their KV-cache row. */
int rdb_op_gemm(int device, int prec, const void* A, int lda, long long M, int K, const void* W, int N, const float* bias, int act,
                const void* res, int ldr, void* out, int ldc, int c_off, void* stream, const int32_t* out_step, long long out_step_stride);
/* dense kh x kw conv as a tcgen05 IMPLICIT GEMM (fp16): x [n,h,w,c] with pixel pitch ld (a channel slice of a wider buffer is
 * fine) read through a 4-D TMA map — padding = TMA out-of-bounds zero fill, stride = element strides, no im2col buffer;
 * wt [cout][kh][kw][c] fp16; out channel slice (ldc, c_off) = act(conv + bias).  ConvBNAct of PPHGNetV2 (rec_pphgnetv2.py:858-913). */
int rdb_op_conv_tc(int device, const void* x, int n, int h, int w, int c, int ld, const void* wt, int cout, const float* bias, int act,
                   int kh, int kw, int sh, int sw, int pt, int pl, void* out, int oh, int ow, int ldc, int c_off, void* stream);
/* x [n,h,w,c] (pixel pitch ld) -> out [n*oh*ow, kh*kw*c], K order (ky, kx, c) = the packed conv weight order */
int rdb_op_im2col(int device, int prec, const void* x, int n, int h, int w, int c, int ld, int kh, int kw, int sh, int sw, int pt, int pl,
                  int oh, int ow, void* out, void* stream);
/* depthwise k x k conv + folded BN + activation (`relu`: 0 none, 1 ReLU, 6 HardSwish — the act codes of rdb_op_gemm):
 * LightConvBNAct.conv2 and the stage downsample (rec_pphgnetv2.py:916-959,1177-1187), the depthwise layers of PP-LCNet */
int rdb_op_dwconv(int device, int prec, const void* x, int n, int h, int w, int c, int ld_in, int k, int stride, const float* wt,
                  const float* bias, int relu, void* out, int oh, int ow, int ld_out, int c_off, void* stream);
/* PaddingSameAsPaddleMaxPool2d(2, stride 1) (rec_pphgnetv2.py:962-976) into a channel slice */
int rdb_op_maxpool2x2s1(int device, int prec, const void* x, int n, int h, int w, int c, int ld_in, void* out, int ld_out, int c_off,
                        void* stream);
/* rows of c (pitch ld_in) -> channel slice of a wider buffer, with dtype change (0 fp32 / 1 fp16) */
int rdb_op_copy_cols(int device, int src_prec, int dst_prec, const void* x, long long rows, int c, int ld_in, void* out, int ld_out,
                     int c_off, void* stream);
int rdb_op_layernorm(int device, const float* x, long long rows, int c, const float* gamma, const float* beta, float eps, float* out,
                     void* stream);
/* embed_tokens[id] * scale + embed_positions[pos + 2]  (MBartLearnedPositionalEmbedding, offset 2).
 * step (device counter, may be NULL): pos = *step and the ids are row *step of a [steps+1, batch] token table — the form every
 * decode-step op takes so that ONE captured CUDA graph replays for all steps. */
int rdb_op_embed(int device, const int64_t* ids, int batch, int dim, const float* tok, float scale, const float* pos_tab, int pos,
                 float* out, void* stream, const int32_t* step);
/* softmax(q k^T) v for ONE new query per row against t cached positions: q [B, heads*head_dim] (already scaled),
 * k / v caches [B, t_cap, heads*head_dim] (MBartAttention.forward with tgt_len 1) */
int rdb_op_attn_decode(int device, const float* q, const float* k, const float* v, int batch, int t, int t_cap, int heads, int head_dim,
                       float* out, void* stream, const int32_t* step /* t = *step + 1 */);
int rdb_op_add(int device, const float* a, const float* b, float* out, long long n, void* stream);
/* ---- ops of the small ONNX CNNs RapidDoc ships (rapid_orientation.onnx: RapidOrientation.__call__,
 * rapid_doc/model/orientation/rapid_orientation/rapid_orientation.py:43-56; pp-ocrv4_mobile_seal_det.onnx), NHWC fp32.
 * rdb_op_chain applies up to 8 per-element steps in one launch; kinds: 0 x*a+b, 1 x*va[c]+vb[c] (either pointer may be NULL:
 * folded BatchNormalization / conv bias), 2 Relu, 3 HardSigmoid(alpha=a, beta=b), 4 x*HardSigmoid(x) (HardSwish: a=1/6, b=.5),
 * 5 Sigmoid.  va / vb are host arrays of n_steps device pointers. */
int rdb_op_chain(int device, const float* x, long long rows, int c, int ld_in, const int32_t* kinds, const float* a, const float* b,
                 const float* const* va, const float* const* vb, int n_steps, float* out, int ld_out, int c_off, void* stream);
int rdb_op_global_avgpool(int device, const float* x, int n, int hw, int c, int ld, float* out /* [n,c] */, void* stream);
/* squeeze-excite scaling out[n,p,c] = x[n,p,c] * gate[n,c] */
int rdb_op_mul_gate(int device, const float* x, const float* gate, int n, int hw, int c, int ld_in, float* out, int ld_out, void* stream);
/* Resize(nearest, asymmetric, floor) to oh x ow, into a channel slice: src = min(floor(dst / (out/in)), in - 1) in float32 */
int rdb_op_resize_nearest(int device, const float* x, int n, int h, int w, int c, int ld_in, int oh, int ow, float* out, int ld_out, int c_off,
                          void* stream);
/* ConvTranspose(kernel = stride = scale) = rdb_op_gemm to [n*h*w, scale*scale*c] (columns dy, dx, c) + this pixel shuffle */
int rdb_op_depth_to_space(int device, const float* g, int n, int h, int w, int c, int scale, float* out, void* stream);
int rdb_op_softmax_rows(int device, const float* x, long long rows, int c, float* out, void* stream);
/* Image normalisation of TablePreprocess (rapid_table_self/table_structure/pp_structure/pre_process.py) on the device:
 * img [n,h,w,3] uint8 canvases whose top-left valid_hw[i] = (rh, rw) region holds the resized image; out [n,h,w,4] fp32 NHWC
 * (4th channel and the padding = 0); lut [3][256] fp32 = the reference's numpy expression evaluated for every byte value. */
/* First layer of the ONNX CNNs: dense 3x3 conv (stride, symmetric pad) on the 4-channel fp32 NHWC input (3 channels + a zero
 * one, the layout rdb_op_lut_u8_nhwc4 writes), 16 output channels, wt [16][3][3][4]; direct form, bit-identical to
 * rdb_op_im2col + rdb_op_gemm(RDB_PREC_FP32) in a ninth of the HBM traffic. */
int rdb_op_conv3x3_c4(int device, const float* x, int n, int h, int w, const float* wt, int cout, const float* bias, int act, int stride, int pad,
                      float* out, int oh, int ow, int ldc, int c_off, void* stream);
int rdb_op_lut_u8_nhwc4(int device, const uint8_t* img, const int32_t* valid_hw, const float* lut, int n, int h, int w, float* out,
                        void* stream);

/* ---- table structure: SLANet head ---------------------------------------------------------
 * The GRU-attention decode loop of SLAHead (the `Loop` node of slanet-1m.onnx / SLANet_plus, executed by onnxruntime inside
 * OrtInferSession.__call__, rapid_doc/model/table/rapid_table_self/inference_engine/onnxruntime/main.py:70-76; caller
 * PPTableStructurer.__call__, table_structure/pp_structure/main.py:41-51) as ONE persistent launch, one CTA per table image.
 * All matrices fp32 on the device, stored [in][out]; the GRU matrices transposed to [in][3*hidden] (gate order r, z, c). */
typedef struct rdb_sla_weights {
  int hidden;                              /* 256 */
  const float *Wh, *bh;                    /* h2h   [hidden][hidden], [hidden]            (linear_1) */
  const float *ws;                         /* score [hidden]                              (linear_2) */
  const float *WihT, *WhhT, *bih, *bhh;    /* GRUCell: [C+classes][3h], [h][3h], [3h], [3h] */
  const float *W3, *b3, *W4, *b4;          /* structure generator: [h][h], [h][classes]   (linear_3, linear_4) */
  const float *W5, *b5, *W6, *b6;          /* box generator: [h][h], [h][loc] + sigmoid   (linear_5, linear_6) */
} rdb_sla_weights_t;
const char* rdb_sla_last_error(void);
/* feat [batch, hw, c] (the neck output, NHWC) and feat_proj = feat * Wi2h [batch, hw, hidden] (hoisted out of the loop);
 * logits / probs [batch, max_steps, classes], loc [batch, max_steps, loc_dim], ids [batch, max_steps]: rows >= *total_steps read
 * as the reference's untouched rows (probability 1/classes, box 0).  *total_steps = steps the reference loop would run: first
 * step after which every image has emitted `eos` (or max_steps); the caller keeps min(total_steps + 1, max_steps) rows, as the
 * graph's final Slice does.  sync_words: 2 int32 of device scratch; steps_run [batch] (diagnostic). */
int rdb_sla_decode(int device, const float* feat, const float* feat_proj, int batch, int hw, int c, const rdb_sla_weights_t* w, int classes,
                   int loc_dim, int max_steps, int eos, float* logits, float* probs, float* loc, int32_t* ids, int32_t* sync_words,
                   int32_t* steps_run, int32_t* total_steps, void* stream);
/* one greedy step of generate_export (rec_ppformulanet_head.py:1118-1160) on the device: next token = argmax (eos when force_eos),
 * finished rows emit pad, a row finishes on eos; *all_done = every row has produced an eos */
int rdb_op_greedy_step(int device, const int32_t* argmax, int batch, int force_eos, int eos, int pad, int64_t* next, int32_t* unfinished,
                       int32_t* has_eos, int32_t* all_done, void* stream, int32_t* step /* token-table row + counter, incremented */,
                       int forced_len);

/* ---- text recognition ----------------------------------------------------------------- */
int rdb_rec_create(const void* weights, size_t nbytes, int device, int precision, rdb_rec_t** out);
void rdb_rec_destroy(rdb_rec_t* h);
int rdb_rec_vocab(rdb_rec_t* h); /* 18710 */
/* T (CTC steps) the network produces for an input of width wid: stem1 s2, stem3 s2,
 * avg_pool [3,2] (backbones/rec_lcnetv4.py:151,154,311); wid = int(48*max_wh_ratio) is
 * arbitrary (rapid_ocr.py:425-438), T = wid/8 for multiples of 8. */
int rdb_rec_tokens(int wid);

/* InferSession seam for rec.  x [n,3,48,w] f32 (w multiple of 8), T = w/8.
 * Replaces TorchInferSession.__call__ + softmax (rapid_doc/model/ocr/torch.py:171-192;
 * network backbones/rec_lcnetv4.py:26-43,306-311, necks/rnn.py:321-379,
 * heads/rec_multi_head.py:66-77) AND CTCLabelDecode's argmax/max (rapidocr
 * ch_ppocr_rec/utils.py, called rapid_ocr.py:443-449) fused in the head GEMM epilogue:
 *   ids   [n,T] int32  argmax token per step          (may be NULL)
 *   probs [n,T] f32    softmax probability of that token (may be NULL)
 *   text_ids [n,T] int32 CTC-collapsed ids, -1 padded; text_len [n]; conf [n] mean kept prob
 *   softmax [n,T,V] f32 full probability tensor (may be NULL; the compat output the
 *   reference engine returns — 3 MB per 48x320 crop, avoid on the fast path). */
int rdb_rec_infer_f32(rdb_rec_t* h, const float* x, int n, int wid, int32_t* ids, float* probs,
                      int32_t* text_ids, int32_t* text_len, float* conf, float* softmax, void* stream);

/* Facade seam for rec: uint8 BGR crops already resized to height 48, packed [n,48,w,3] with
 * valid_w[i] real columns (right part zero-padded AFTER normalisation, as
 * TextRecognizer.resize_norm_img does; called rapid_ocr.py:438).  Fuses (v/255-0.5)/0.5. */
int rdb_rec_infer_u8(rdb_rec_t* h, const uint8_t* crops, const int32_t* valid_w, int n, int wid, int32_t* ids,
                     float* probs, int32_t* text_ids, int32_t* text_len, float* conf, void* stream);

/* stats of the last infer call on a handle: number of kernel launches it enqueued */
long long rdb_det_last_launches(rdb_det_t* h);
long long rdb_rec_last_launches(rdb_rec_t* h);

/* Diagnostic: one GEMM out[M,N] = act(A[M,K] W[N,K]^T + bias) (+res) through the fp16
 * pointwise-conv engines (use_tc=1 tcgen05 kernel, 0 SIMT kernel); host fp32 in/out, inputs
 * rounded to fp16 as the engines store them.  mode 1: fused CTC epilogue, out = [M,2]
 * (argmax id, softmax max prob).  Used by the kernel unit tests. */
int rdb_debug_gemm(int device, int use_tc, int mode, const float* A, const float* W, const float* bias,
                   const float* res, int M, int N, int K, int act, float* out);

/* per-kernel device timing (CUDA events on the launching stream around every launch of
 * subsequent infer calls; process-wide).  dump writes a JSON object
 * {"kernel": [total_ms, launches], ...} and returns the bytes needed. */
/* host-only: the 1024 x 16 int16 bicubic weight table rdb_warp_crops uses (OpenCV initInterTab2D(INTER_CUBIC, fixpt)) */
int rdb_debug_cubic_tab(int16_t* out);
/* the RDB_* A/B switches are read from the environment once per process; call this after changing one */
int rdb_switches_reload(void);
int rdb_profile_enable(int on);
int rdb_profile_reset(void);
int rdb_profile_dump(char* buf, size_t cap);

/* scheduling knobs: pixels (det) / crops (rec) processed per internal chunk */
int rdb_det_set_chunk_pixels(rdb_det_t* h, long long pixels);
int rdb_rec_set_chunk_crops(rdb_rec_t* h, int crops);

#ifdef __cplusplus
}
#endif
#endif /* RAPIDDOC_B200_H_ */
