/* witw_b200.h -- C ABI of libwitw_b200.so: B200 (sm_100a) kernels for the WITW
 * cross-view retrieval hot path.
 *
 * The reference (IQTLabs/WITW) has no FFI: its hot path is six Python functions in
 * model/cvig_fov.py (byte-identical copies in model/cvig_semantic.py) plus the rank
 * loop of test().  This header is the boundary a binding for that path uses; each
 * entry point names the reference code it replaces (paths relative to the reference
 * repository root).  The Python binding (ctypes) is witw_b200/_lib.py, and
 * witw_b200/ops.py mirrors the reference's function names and signatures.
 *
 * Conventions
 *   - every function returns 0 (WITW_OK) or a negative WITW_ERR_* code; the message
 *     is available from witw_last_error() (thread-local)
 *   - pointers named *_dev are device pointers on the current CUDA device, all
 *     others are host pointers; tensors are dense, row-major, in the reference's
 *     layouts (NCHW features, [gallery, query] matrices)
 *   - no function allocates device memory; sizes of operand/workspace buffers are
 *     returned by the *_bytes queries and the caller owns the buffers
 *   - work is enqueued on `stream` (a cudaStream_t) and not synchronised
 *   - there is no CPU fallback: without an sm_100 device every compute call fails
 */
#ifndef WITW_B200_H
#define WITW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* witw_stream_t; /* cudaStream_t */

#define WITW_OK 0
#define WITW_ERR_INVALID (-1)     /* bad argument (shape, null pointer, alignment) */
#define WITW_ERR_UNSUPPORTED (-2) /* valid request outside what the kernels cover */
#define WITW_ERR_CUDA (-3)        /* CUDA runtime / driver error */
#define WITW_ERR_DEVICE (-4)      /* no sm_100 device */

const char* witw_last_error(void);
int witw_version(void);
/* 0 when the current device is compute capability 10.x, WITW_ERR_DEVICE otherwise */
int witw_device_check(void);
/* Opt-in: keep [ptr, ptr + bytes) resident in L2 for the kernels launched on `stream` from now on (an access-policy window
 * backed by the device's persisting-L2 set-aside, which the call raises -- a device-wide setting -- to what is needed, at most
 * the device's maximum).  bytes == 0 removes the window.  Meant for the query operand of a sweep, which is re-read once per 8
 * gallery items while uploads and the gallery operand stream through the same cache. */
int witw_stream_l2_window(const void* ptr_dev, size_t bytes, witw_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1  polar transform            replaces model/cvig_fov.py:156-209
 *                                (bilinear_interpolate + PolarTransform.__call__)
 * ------------------------------------------------------------------------------------------ */

/* Host, float64.  Sample coordinates of PolarTransform (cvig_fov.py:197-201) for an
 * [h_s, w_s] panorama-aligned image resampled from an [s_o, s_o] tile.  x, y: [h_s*w_s]. */
int witw_polar_grid(int h_s, int w_s, int s_o, double* x, double* y);

/* Host.  Tap indices and fp32 weights of bilinear_interpolate (cvig_fov.py:163-181):
 * floor, +1, clip all four to the image, weights from the clipped integers in float64,
 * rounded to fp32.  idx4: [n][4] = x0,x1,y0,y1;  w4: [n][4] = wa,wb,wc,wd. */
int witw_bilinear_lut(const double* x, const double* y, int64_t n, int src_h, int src_w,
                      int32_t* idx4, float* w4);

/* Device.  dst[i, p] = wa*src[i,y0,x0] + wb*src[i,y1,x0] + wc*src[i,y0,x1] + wd*src[i,y1,x1]
 * (cvig_fov.py:173-183, same order, every product and sum rounded to fp32) for n_img
 * planes [src_h, src_w] and n_out sample points.  Bit-exact with the reference; generic
 * geometry.  This is what the bilinear_interpolate drop-in calls. */
int witw_bilinear_gather_f32(const float* src_dev, float* dst_dev, const int32_t* idx4_dev,
                             const float* w4_dev, int64_t n_img, int src_h, int src_w,
                             int64_t n_out, witw_stream_t stream);

/* Fast polar resample: shared-memory staged quadrants, register-resident sample table.
 * A plan holds the packed table for one geometry; build it once on the host, copy it to
 * the device, pass both copies.  Weights are (1-fx, fx) x (1-fy, fy) with fx, fy the
 * float64 fractions rounded to fp32, i.e. within 2^-24 absolute of the reference's
 * weights; pixels whose taps the reference clips are patched with the exact formula. */
size_t witw_polar_plan_bytes(int h_s, int w_s, int s_o);
int witw_polar_plan_build(int h_s, int w_s, int s_o, void* plan_host);
int witw_polar_resample_f32(const float* src_dev, float* dst_dev, int64_t n_img,
                            const void* plan_host, const void* plan_dev, witw_stream_t stream);

/* uint8 tiles in, normalised polar images out: ImageNormalization (cvig_fov.py:137-149, norm(data / 255.)) fused
 * into the polar transform (cvig_fov.py:186-209) for tiles that already have the model's size, where Resize
 * (cvig_fov.py:133) is the identity.  The source is read at one byte per pixel instead of four.
 *   witw_norm_lut (host): lut[c][v] = ((v / divisor[c]) - mean[c]) / std[c] for v = 0..255, each step rounded to
 *        fp32 as torch does (divisor 255 for image channels).
 *   witw_bilinear_gather_u8: the bit-exact path -- every tap is looked up in the table, then blended in the
 *        reference's order; plane p uses channel p % n_ch.  lut_dev: [n_ch][256] fp32 on the device.
 *   witw_polar_resample_u8: the staged throughput path on a plan built by witw_polar_plan_build_u8 (box rows of
 *        160 bytes starting on a 16-pixel boundary); the raw pixels are blended and the channel's affine map
 *        a*v + b applied once per output pixel (within 1e-6 of the table form), clipped pixels patched exactly. */
int witw_norm_lut(const float* divisor, const float* mean, const float* std, int n_ch, float* lut_host);
int witw_bilinear_gather_u8(const uint8_t* src_dev, float* dst_dev, const int32_t* idx4_dev,
                            const float* w4_dev, int64_t n_planes, int src_h, int src_w, int64_t n_out,
                            const float* lut_dev, int n_ch, witw_stream_t stream);
size_t witw_polar_plan_bytes_u8(int h_s, int w_s, int s_o);
int witw_polar_plan_build_u8(int h_s, int w_s, int s_o, void* plan_host);
int witw_polar_resample_u8(const uint8_t* src_dev, float* dst_dev, int64_t n_planes, int n_ch,
                           const float* lut_host, const float* lut_dev, const void* plan_host,
                           const void* plan_dev, witw_stream_t stream);

/* Raw images in, model-sized normalised images out: Resize (cvig_fov.py:117-134 -- torchvision's bilinear
 * resize of a float image, align_corners=False; the surface panorama's wrap-around column window of
 * cvig_fov.py:120-129) fused with ImageNormalization (cvig_fov.py:137-149; cvig_semantic.py:163-176 divides only the
 * first three channels by 255, hence a per-channel divisor).  The overhead output feeds witw_polar_resample_f32.
 *   plan (host-built, then copied to the device; pass both copies): separable tap tables with ATen's arithmetic,
 *        antialias = 0: upsample_bilinear2d (the reference's pinned torchvision 0.9.1 / torch 1.8.1),
 *        antialias = 1: _upsample_bilinear2d_aa (torchvision >= 0.17's default: the reference as it runs today).
 *   witw_resize_norm: n_planes planes [in_h, in_w] (uint8 when src_is_u8, else fp32) -> [out_h, col_count] fp32;
 *        output column x shows column (col_start + x) % out_w of the resized image; plane p is channel p % n_ch;
 *        mean == NULL: resize only; else dst = ((v / divisor[c]) - mean[c]) / std[c], each step rounded to fp32
 *        (n_ch <= 8).  Resampling: width first with an fp32 intermediate, then height (ATen's order). */
size_t witw_resize_plan_bytes(int in_h, int in_w, int out_h, int out_w, int antialias);
int witw_resize_plan_build(int in_h, int in_w, int out_h, int out_w, int antialias, void* plan_host);
int witw_resize_norm(const void* src_dev, int src_is_u8, float* dst_dev, int64_t n_planes, int n_ch,
                     const void* plan_host, const void* plan_dev, int col_start, int col_count,
                     const float* divisor, const float* mean, const float* std, witw_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2/K3  orientation-searched distance       replaces model/cvig_fov.py:297-363
 *        (correlation -> crop_overhead -> l2_distance)
 *
 *   corr[g,q,s] = sum_{c,h,k<sw} ov[g,c,h,(s+k)%W] * su[q,c,h,k]
 *   ori[g,q]    = argmax_s corr[g,q,s]                (first maximum)
 *   dist[g,q]   = 2 - 2*corr[g,q,ori] / (||crop(ov_g, ori)|| * ||su_q||)
 *
 * CH = C*H feature rows of W (gallery) / sw (query) columns.
 * ------------------------------------------------------------------------------------------ */

/* Exact fp32 path on CUDA cores; any CH, W <= 128, 1 <= sw <= W.  Outputs are optional
 * (NULL to skip): dist [G,Q] fp32, ori [G,Q] int64, corr [G,Q,W] fp32. */
int witw_match_f32(const float* ov_dev, const float* su_dev, int64_t G, int64_t Q, int CH, int W,
                   int sw, float* dist_dev, int64_t* ori_dev, float* corr_dev,
                   witw_stream_t stream);

/* Same, for explicit (gallery, query) index pairs: pair p compares ov[pair_g[p]] with
 * su[pair_q[p]].  dist/ori: [n_pairs].  Used for the true-match distances of the rank
 * evaluation and for fp32 re-checks of tensor-core results. */
int witw_match_pairs_f32(const float* ov_dev, const float* su_dev, const int64_t* pair_g_dev,
                         const int64_t* pair_q_dev, int64_t n_pairs, int CH, int W, int sw,
                         float* dist_dev, int64_t* ori_dev, witw_stream_t stream);

/* Standalone a4 (cvig_fov.py:318-343): out[g,q,ch,k] = ov[g,ch,(k+ori[g,q])%W], k<sw. */
int witw_crop_gather_f32(const float* ov_dev, const int64_t* ori_dev, float* out_dev, int64_t G,
                         int64_t Q, int CH, int W, int sw, witw_stream_t stream);

/* Standalone a5 (cvig_fov.py:346-363): crop [G,Q,K], su [Q,K] -> dist [G,Q]. */
int witw_l2_distance_f32(const float* crop_dev, const float* su_dev, float* dist_dev, int64_t G,
                         int64_t Q, int64_t K, witw_stream_t stream);

/* Backward of a4 / a5 for the reference's train() (cvig_fov.py:450-460; correlation is an argmax and has no
 * gradient).  crop_backward: grad_ov [G,CH,W] = scatter of grad_out [G,Q,CH,sw] back through the roll
 * (deterministic, no atomics).  l2_distance_backward: grad_crop [G,Q,K] and/or grad_su [Q,K] (either may be
 * NULL) from grad_dist [G,Q]; coef_scratch [G,Q,2] fp32 is required when grad_su is requested. */
int witw_crop_backward_f32(const float* grad_out_dev, const int64_t* ori_dev, float* grad_ov_dev, int64_t G,
                           int64_t Q, int CH, int W, int sw, witw_stream_t stream);
int witw_l2_distance_backward_f32(const float* crop_dev, const float* su_dev, const float* grad_dist_dev,
                                  float* grad_crop_dev, float* grad_su_dev, float* coef_scratch_dev,
                                  int64_t G, int64_t Q, int64_t K, witw_stream_t stream);

/* The training slice of train() (cvig_fov.py:450-460) without the [G,Q,C,H,sw] crop: the forward is witw_match_f32
 * (orientation + distance); this is the backward of distance = l2_distance(crop_overhead(ov, ori), su) with respect to
 * both feature sets, from the features, the orientation and grad_dist [G,Q].  grad_ov [G,CH,W] / grad_su [Q,CH,sw]:
 * either may be NULL; coef_scratch: [G,Q,3] fp32.  Deterministic (no atomics). */
int witw_match_backward_f32(const float* ov_dev, const float* su_dev, const int64_t* ori_dev,
                            const float* grad_dist_dev, float* grad_ov_dev, float* grad_su_dev,
                            float* coef_scratch_dev, int64_t G, int64_t Q, int CH, int W, int sw,
                            witw_stream_t stream);

/* triplet_loss (cvig_fov.py:366-382): the soft-margin triplet loss of a [N,N] distance matrix whose diagonal holds the
 * matching pairs, loss_dev[0], and (optionally) its gradient grad_dist_dev [N,N].  2 <= N <= 4096. */
int witw_triplet_loss_f32(const float* dist_dev, int N, float alpha, float* loss_dev,
                          float* grad_dist_dev, witw_stream_t stream);

/* ---- Whole-gallery sweeps on the tensor cores (tcgen05, fp16 operands, fp32 accumulation in TMEM); W must be 64 ----
 *
 * Two sweeps with one contract -- correlation -> crop_overhead -> l2_distance of cvig_fov.py:297-363 for every (gallery
 * item, query) pair, the rank rule of cvig_fov.py:552 and a per-query top-k fused behind it:
 *   witw_match_spec  (csrc/match_spec.cu, needs C*H == 64) evaluates the circular correlation through the correlation
 *        theorem, corr[g,q,:] = irfft( sum_r O[g,r,f] conj(S[q,r,f]) ): 33 bins x 64 feature rows of complex MACs on
 *        tcgen05 (16.9 kFLOP per pair instead of 524 kFLOP at 360 degrees), the 64-point inverse real FFT, the argmax
 *        over the shift and everything after it in the epilogue, register-local.
 *   witw_match_tc    (csrc/match_tc.cu) is the same search as one dense contraction over all 64 shifts; the gallery
 *        operand is a set of pre-shifted 16-byte rows that a no-swizzle UMMA descriptor reads as the Hankel matrix.
 *
 * Operands (csrc/sweep_common.cuh).  Features are scaled to unit norm before they are rounded to fp16, so an accumulator is
 * acc[s] = unit * corr[g,q,s] / (||ov_g|| ||su_q||) and
 *        dist = 2 - 2 * acc[s*] * gal_scale[g,s*] * qry_aux[q][0].
 * The prep functions write, per gallery item: the operand, gal_scale [G_pad,64], gal_aux [G_pad,4] = (max and spread of
 * gal_scale[g,:], the item's rounding scale, the operand's scale) and -- optional, for the fp32 finish -- crop_inv_norm
 * [G_pad,64] = 1/||crop(ov_g,s)||; per query: the operand, qry_aux [Q,2] = (1, or NaN for a zero-norm query; the query's
 * rounding scale) and q_inv_norm [Q].  G_pad = G rounded up to 4 (dense) / 8 (spectral); the tables of a call start at its
 * first item.  witw_spec_gallery_prep: g_first = index of ov[0] inside the operand (a multiple of 8 unless it continues a
 * partial group), so an encode loop (cvig_fov.py:519-532) can append batch by batch.  spec_out (optional): the fp32
 * spectra of the same rows in the layout of witw_spectral_rows_f32 -- the operands of the fp32 finish, from the same pass. */
size_t witw_gallery_operand_bytes(int64_t G, int CH, int sw);
size_t witw_query_operand_bytes(int64_t Q, int CH, int sw);
int witw_gallery_prep(const float* ov_dev, int64_t G, int CH, int W, int sw, void* gal_op_dev, float* gal_scale_dev,
                      float* gal_aux_dev, float* crop_inv_norm_dev /* optional */, witw_stream_t stream);
int witw_query_prep(const float* su_dev, int64_t Q, int CH, int sw, void* qry_op_dev, float* qry_aux_dev,
                    float* q_inv_norm_dev, witw_stream_t stream);
int witw_spec_supported(int CH, int W, int sw);
size_t witw_spec_gallery_operand_bytes(int64_t G, int CH);
size_t witw_spec_query_operand_bytes(int64_t Q, int CH);
int witw_spec_gallery_prep(const float* ov_dev, int64_t G, int64_t g_first, int CH, int W, int sw, void* gal_op_dev,
                           float* gal_scale_dev, float* gal_aux_dev, float* crop_inv_norm_dev /* optional */,
                           float* spec_out_dev /* optional */, witw_stream_t stream);
int witw_spec_query_prep(const float* su_dev, int64_t Q, int CH, int sw, void* qry_op_dev, float* qry_aux_dev,
                         float* q_inv_norm_dev, float* spec_out_dev /* optional */, witw_stream_t stream);

/* One sweep of Q queries over G gallery items.  All pointers are device pointers; every output is optional (NULL).
 *   dist [G,Q] fp32, ori [G,Q] uint8
 *   rank_count [Q] int32 += #{g : dist[g,q] <= d_true[q]}  (needs d_true [Q]; the rank rule of cvig_fov.py:552; the caller
 *        zeroes rank_count, shards add into it).  true_idx [Q] (optional, global gallery index of each query's match):
 *        that item is counted iff d_true[q] is not NaN -- d[idx] <= d[idx] in the reference.
 *   topk_key / topk_idx [n_slots,Q,topk]: per query the topk smallest (key, gallery index + g_index_offset) candidates
 *        of each of n_slots gallery slices, ascending; n_slots from witw_match_*_topk_slots(); merge with
 *        witw_topk_merge().  topk <= 16.
 * err_sigmas > 0 bounds what the fp16 operands can have done to a result (err_sigmas standard deviations of the
 * accumulated rounding error, per pair; 5 is the library's default) and changes three things: a top-k key is
 * dist - slack, a lower bound of the pair's fp32 distance; a rank decision within slack of d_true is not taken but
 * deferred; and for the matrix outputs, pairs whose argmax is not certain or whose slack exceeds fix_rel * dist are
 * deferred for an fp32 overwrite.  Deferred pairs go to per-query lists: list_n [Q] (zeroed by the caller) counts them,
 * list_g [Q,list_cap] holds the local gallery index with bit 31 set when the rank decision is pending.  A count above
 * list_cap means the query's list is incomplete: witw_finish_spec_f32 flags it and the caller re-does that query with
 * witw_match_columns_spec_f32.  err_sigmas == 0: keys are the fp16 distances and every decision is taken from them. */
typedef struct witw_sweep_args {
  const void* gal_op;
  const float* gal_scale;
  const float* gal_aux;
  const void* qry_op;
  const float* qry_aux;
  int64_t G, Q;
  int32_t CH, sw;
  int32_t g_index_offset;
  int32_t topk;
  float* dist;
  uint8_t* ori;
  const float* d_true;
  const int32_t* true_idx;
  int32_t* rank_count;
  float* topk_key;
  int32_t* topk_idx;
  int32_t* list_g;
  int32_t* list_n;
  int32_t list_cap;
  float err_sigmas;
  float fix_rel;
} witw_sweep_args;
int witw_match_tc_topk_slots(int64_t G, int64_t Q);
int witw_match_tc(const witw_sweep_args* args, witw_stream_t stream);
int witw_match_spec_topk_slots(int64_t G, int64_t Q);
/* How witw_match_spec tiles the sweep: 2 (default) = a CTA pair per 128-query x 16-item tile (tcgen05 cta_group::2: the pair
 * shares the gallery operand, each CTA stages 64 queries), 1 = one CTA per 128-query x 8-item tile.  Same results; a
 * process-wide setting meant for measurements (call before witw_match_spec_topk_slots). */
int witw_match_spec_variant(int cta_group);
int witw_match_spec(const witw_sweep_args* args, witw_stream_t stream);

/* ---- fp32 evaluation in the azimuth-frequency domain (csrc/spectral.cu, csrc/finish.cu) ----
 * The circular correlation of cvig_fov.py:297-312 through the correlation theorem on packed 64-point spectra of the
 * feature rows: ~13k MACs per pair instead of 262k, and in fp32 closer to the float64 value than a 4096-term fp32 dot
 * product.  W must be 64; narrower query rows are zero-padded.
 *   witw_spectral_rows_f32: x [n_rows,row_len] fp32 -> spec [n_rows,64] fp32 (32 float2 slots per row: slot f =
 *        (Re X_f, Im X_f) for f = 1..31, slot 0 = (X_0, X_32)).  Gallery: n_rows = G*CH, row_len = 64; queries:
 *        n_rows = Q*CH, row_len = sw.
 *   witw_match_pairs_spec_f32: (dist [n_pairs], ori [n_pairs] int64) of explicit pairs; crop_inv_norm [G,64] and
 *        q_inv_norm [Q] are the fp32 tables of the prep functions.  The true-match distances of the rank rule.
 *   witw_match_columns_spec_f32: the distances / orientations of F selected queries (q_sel [F], NULL = 0..F-1) against
 *        every gallery item: element (g, column) at [g*ld + column], column = the query index when col_is_q, else
 *        0..F-1; ori as int64 and / or uint8; count_out [F] += #{g : d <= d_true[q]} (the match itself, true_idx[q] -
 *        g_index_offset, counted by index).  The whole answer for a few queries (heat map, cvig_fov.py:545-552 one query
 *        at a time) and the fallback for queries witw_finish_spec_f32 flags.
 *   witw_finish_spec_f32: the fp32 finish of a sweep run with err_sigmas > 0: evaluates every query's deferred pairs
 *        (pending rank decisions are added to rank_count, matrix entries overwritten) and
 *        re-ranks the sweep's merged top-k candidates cand_key / cand_idx [Q,kc] (ascending keys) into out_dist /
 *        out_idx [Q,k_out] -- exact distances, ties by lower index.  qflag [Q] / n_flagged [1] (zeroed by the caller):
 *        bit 0 = the query's list overflowed, bit 1 = the keys do not prove that no item outside the candidate list
 *        belongs to the top k_out; flagged queries must be re-done with witw_match_columns_spec_f32. */
int witw_spectral_rows_f32(const float* x_dev, int64_t n_rows, int row_len, float* spec_dev,
                           witw_stream_t stream);
int witw_match_pairs_spec_f32(const float* gal_spec_dev, const float* crop_inv_norm_dev,
                              const float* qry_spec_dev, const float* q_inv_norm_dev,
                              const int64_t* pair_g_dev, const int64_t* pair_q_dev, int64_t n_pairs, int CH,
                              float* dist_dev, int64_t* ori_dev, witw_stream_t stream);
int witw_match_columns_spec_f32(const float* gal_spec_dev, const float* crop_inv_norm_dev, const float* qry_spec_dev,
                                const float* q_inv_norm_dev, int64_t G, int CH, const int32_t* q_sel_dev, int64_t F,
                                float* dist_dev, int64_t* ori64_dev, uint8_t* ori8_dev, int64_t ld, int col_is_q,
                                const float* d_true_dev, const int32_t* true_idx_dev, int32_t g_index_offset,
                                int32_t* count_out_dev, witw_stream_t stream);
typedef struct witw_finish_args {
  const float* gal_spec;
  const float* crop_inv_norm;
  const float* qry_spec;
  const float* q_inv_norm;
  int64_t G, Q;
  int32_t CH;
  int32_t g_index_offset;
  const int32_t* list_g;
  const int32_t* list_n;
  int32_t list_cap;
  int32_t kc;
  const float* d_true;
  int32_t* rank_count;
  float* dist;
  uint8_t* ori;
  const float* cand_key;
  const int32_t* cand_idx;
  float* out_dist;
  int32_t* out_idx;
  int32_t k_out;
  int32_t reserved;
  int32_t* qflag;
  int32_t* n_flagged;
  void* scratch; /* witw_finish_scratch_bytes(Q, kc) bytes, 16-byte aligned */
} witw_finish_args;
size_t witw_finish_scratch_bytes(int64_t Q, int kc);
int witw_finish_spec_f32(const witw_finish_args* args, witw_stream_t stream);
/* sizeof(witw_sweep_args) / sizeof(witw_finish_args) as this library was built: a binding checks its own declaration. */
size_t witw_sizeof_sweep_args(void);
size_t witw_sizeof_finish_args(void);

/* ---- peer-memory exchange of a gallery-sharded evaluation (csrc/peer.cu; witw_b200/sharded.py) ----
 * The reference evaluates on one GPU (cvig_fov.py:519-552); with the gallery sharded over the GPUs of a box the rank rule needs
 * the true-match distances everywhere before the sweep and the sum of the rank counts / the merge of the top-k lists after it.
 * Each rank owns one exchange buffer (witw_peer_alloc: cudaMalloc + a 64-byte CUDA IPC handle for the other ranks'
 * witw_peer_open); peers_dev is a DEVICE array of `world` buffer addresses as this rank sees them, own buffer at [rank].
 * Both calls store this rank's part into every rank's buffer over NVLink, signal, wait for every peer's signal of the same
 * sequence number (seq = 1, 2, ... identical on all ranks, one per call) and finish locally:
 *   witw_peer_thresholds: d_local [Q] holds the fp32 distance of query q to item true_idx[q] where this rank owns that item
 *        (g_offset <= true_idx[q] < g_offset + g_local); d_true_out [Q] receives the complete vector.
 *   witw_peer_results: counts [Q] int32, td / ti [Q,k] (k may be 0), flagged [1] or NULL of this shard -> total_out [Q] int64
 *        (the summed counts), n_flag_out [1] (the summed flags; bit 30: a wait timed out), merged_dist / merged_idx [Q,k]
 *        (witw_topk_merge over the ranks' lists in rank order).
 * world <= 16; k the same in both calls of an evaluation. */
size_t witw_peer_exchange_bytes(int64_t Q, int k, int world);
int witw_peer_alloc(size_t bytes, void** buf_dev, void* ipc_handle_64);
int witw_peer_open(const void* ipc_handle_64, void** buf_dev);
int witw_peer_close(void* buf_dev);
int witw_peer_free(void* buf_dev);
int witw_peer_thresholds(const float* d_local_dev, const int64_t* true_idx_dev, int64_t g_offset, int64_t g_local,
                         int64_t Q, int k, void* const* peers_dev, int world, int rank, uint32_t seq,
                         float* d_true_out_dev, witw_stream_t stream);
int witw_peer_results(const int32_t* counts_dev, const float* td_dev, const int32_t* ti_dev, const int32_t* flagged_dev,
                      int64_t Q, int k, void* const* peers_dev, const void* own_buf_dev, int world, int rank,
                      uint32_t seq, int64_t* total_out_dev, int32_t* n_flag_out_dev, float* merged_dist_dev,
                      int32_t* merged_idx_dev, witw_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  rank / top-k               replaces model/cvig_fov.py:550-552 and
 *                                model/cvig_baseline.py:456-460
 * ------------------------------------------------------------------------------------------ */

/* ranks[q] = #{g : dist[g,q] <= dist[true_idx[q], q]} over a materialised [G,Q] matrix
 * (true_idx NULL = identity).  NaN compares false, ties and the match itself count. */
int witw_rank_from_dist_f32(const float* dist_dev, int64_t G, int64_t Q,
                            const int64_t* true_idx_dev, int64_t* ranks_dev,
                            witw_stream_t stream);

/* Baseline variant (cvig_baseline.py:458-460): Euclidean distance between [N,D] gallery and
 * [Q,D] query embeddings, then the same rank rule.  dist_dev [N,Q] is required: it is both an
 * output and the workspace the ranks are counted from.  ranks_dev optional. */
int witw_l2_rank_f32(const float* ov_dev, const float* su_dev, int64_t N, int64_t Q, int64_t D,
                     const int64_t* true_idx_dev, float* dist_dev, int64_t* ranks_dev,
                     witw_stream_t stream);

/* Per query (column) the k smallest distances of dist [G,Q], ascending, ties by lower gallery index;
 * idx gets g + g_index_offset; NaN / +inf never enter (unfilled slots are (+inf, -1)).  k <= 128.
 * The gallery is cut into n_slices row slices (witw_topk_slices() suggests a count that fills the
 * machine); slice s writes its candidate list to out[s][q][k].  n_slices == 1 gives the final [Q,k];
 * otherwise merge with witw_topk_merge(). */
int witw_topk_slices(int64_t G, int64_t Q);
int witw_topk_from_dist_f32(const float* dist_dev, int64_t G, int64_t Q, int k, int n_slices,
                            float* topk_dist_dev, int32_t* topk_idx_dev, int32_t g_index_offset,
                            witw_stream_t stream);

/* The same result for large galleries (G >= 1024, k <= 32, Q a multiple of 4) without candidate lists: thresholds from
 * a strided row sample, one streaming pass that appends the elements at or below the threshold to per-column survivor
 * buffers, and a warp-per-column selection.  scratch: witw_topk_select_scratch_bytes(). */
size_t witw_topk_select_scratch_bytes(int64_t Q, int k);
int witw_topk_select_f32(const float* dist_dev, int64_t G, int64_t Q, int k, float* topk_dist_dev,
                         int32_t* topk_idx_dev, int32_t g_index_offset, void* scratch_dev,
                         witw_stream_t stream);

/* Merge n_lists sorted candidate lists per query ([n_lists,Q,k] each) into one [Q,k]. */
int witw_topk_merge(const float* cand_dist_dev, const int32_t* cand_idx_dev, int n_lists,
                    int64_t Q, int k, float* topk_dist_dev, int32_t* topk_idx_dev,
                    witw_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WITW_B200_H */
