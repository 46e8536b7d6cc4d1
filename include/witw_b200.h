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

/* Tensor-core path (tcgen05, bf16 operands, fp32 accumulation in TMEM); W must be 64.
 * gallery_prep writes the gallery operand (pre-shifted 16-byte rows that a no-swizzle UMMA
 * descriptor reads as the Hankel matrix of all 64 azimuth shifts) and the table
 * crop_inv_norm[g,s] = 1/||crop(ov_g,s)|| (fp32, from the fp32 inputs).  query_prep writes
 * the bf16 query operand [Q, CH*sw_pad] and q_inv_norm[q]. */
size_t witw_gallery_operand_bytes(int64_t G, int CH, int sw);
size_t witw_query_operand_bytes(int64_t Q, int CH, int sw);
int witw_gallery_prep(const float* ov_dev, int64_t G, int CH, int W, int sw, void* gal_op_dev,
                      float* crop_inv_norm_dev /* [G_pad4,64], G_pad4 = G rounded up to 4 */,
                      witw_stream_t stream);
int witw_query_prep(const float* su_dev, int64_t Q, int CH, int sw, void* qry_op_dev,
                    float* q_inv_norm_dev /* [Q] */, witw_stream_t stream);

/* One sweep of Q queries over G gallery items.  Optional outputs (NULL to skip):
 *   dist [G,Q] fp32, ori [G,Q] uint8
 *   rank_count [Q] int32 += #{g : dist[g,q] <= d_true[q]}   (needs d_true [Q]; the rank rule
 *        of cvig_fov.py:552; the caller zeroes rank_count, shards add into it).  true_idx [Q]
 *        (optional, global gallery index of each query's match): that item is counted iff
 *        d_true[q] is not NaN, whatever its bf16 distance -- d[idx] <= d[idx] in the reference
 *   topk_dist/topk_idx [n_slots,Q,topk]: per-query k smallest (distance, gallery index +
 *        g_index_offset) candidates of each of n_slots gallery slices, ascending; n_slots from
 *        witw_match_tc_topk_slots(); merge them with witw_topk_merge(). topk <= 16.
 */
int witw_match_tc_topk_slots(int64_t G, int64_t Q);
int witw_match_tc(const void* gal_op_dev, const float* crop_inv_norm_dev, const void* qry_op_dev,
                  const float* q_inv_norm_dev, int64_t G, int64_t Q, int CH, int sw,
                  float* dist_dev, uint8_t* ori_dev, const float* d_true_dev,
                  const int32_t* true_idx_dev, int32_t* rank_count_dev, int topk,
                  float* topk_dist_dev, int32_t* topk_idx_dev, int32_t g_index_offset,
                  float recheck_band, int64_t* recheck_g_dev, int64_t* recheck_q_dev,
                  int32_t* recheck_count_dev, int32_t recheck_capacity, witw_stream_t stream);

/* Exact finish of the tensor-core rank count.  With recheck_capacity > 0 witw_match_tc does not
 * count pairs whose bf16 distance is within recheck_band of d_true[q]; it appends them (local gallery
 * index, query) to the lists and bumps recheck_count_dev[0] (recheck_count_dev[1] counts pairs that did
 * not fit and were decided in bf16; the caller zeroes both).  witw_recheck_apply_f32 recomputes the
 * listed pairs in exact fp32 and adds d_exact <= d_true[q] to rank_count -- so the ranks are those of
 * the fp32 reference chain (cvig_fov.py:547-552) unless the list overflowed.  scratch: [capacity] fp32. */
int witw_recheck_apply_f32(const float* ov_dev, const float* su_dev, const int64_t* recheck_g_dev,
                           const int64_t* recheck_q_dev, const int32_t* recheck_count_dev,
                           int32_t capacity, int CH, int W, int sw, const float* d_true_dev,
                           int32_t* rank_count_dev, float* scratch_dev, witw_stream_t stream);

/* Exact re-ranking of top-k candidates: cand_idx [Q,kc] (global indices, -1 = empty) of a gallery
 * [G,...] whose first item has global index g_index_offset; distances are recomputed in fp32 and the
 * k_out best kept, ascending, ties by lower index.  kc <= 32.  scratch: witw_topk_refine_scratch_bytes(). */
size_t witw_topk_refine_scratch_bytes(int64_t Q, int kc);
int witw_topk_refine_f32(const float* ov_dev, const float* su_dev, int64_t G, int64_t Q, int CH, int W,
                         int sw, const int32_t* cand_idx_dev, int kc, int32_t g_index_offset, int k_out,
                         float* topk_dist_dev, int32_t* topk_idx_dev, void* scratch_dev,
                         witw_stream_t stream);

/* The same finishes in the azimuth-frequency domain (csrc/spectral.cu).  The circular correlation of
 * cvig_fov.py:297-312 is evaluated through the correlation theorem on packed 64-point spectra of the
 * feature rows: ~13k MACs per pair instead of 262k, and in fp32 closer to the float64 value than a
 * 4096-term fp32 dot product.  W must be 64; narrower query rows are zero-padded.
 *   witw_spectral_rows_f32: x [n_rows,row_len] fp32 -> spec [n_rows,64] fp32 (32 float2 slots per row:
 *        slot f = (Re X_f, Im X_f) for f = 1..31, slot 0 = (X_0, X_32)).  Gallery: n_rows = G*CH,
 *        row_len = 64; queries: n_rows = Q*CH, row_len = sw.
 *   witw_match_pairs_spec_f32: exact (dist [n_pairs], ori [n_pairs] int64) of explicit pairs;
 *        crop_inv_norm [G,64] and q_inv_norm [Q] are the fp32 tables of witw_gallery_prep / witw_query_prep.
 *   witw_recheck_apply_spec_f32 / witw_topk_refine_spec_f32: as the _f32 forms above, on spectra. */
int witw_spectral_rows_f32(const float* x_dev, int64_t n_rows, int row_len, float* spec_dev,
                           witw_stream_t stream);
int witw_match_pairs_spec_f32(const float* gal_spec_dev, const float* crop_inv_norm_dev,
                              const float* qry_spec_dev, const float* q_inv_norm_dev,
                              const int64_t* pair_g_dev, const int64_t* pair_q_dev, int64_t n_pairs, int CH,
                              float* dist_dev, int64_t* ori_dev, witw_stream_t stream);
int witw_recheck_apply_spec_f32(const float* gal_spec_dev, const float* crop_inv_norm_dev,
                                const float* qry_spec_dev, const float* q_inv_norm_dev,
                                const int64_t* recheck_g_dev, const int64_t* recheck_q_dev,
                                const int32_t* recheck_count_dev, int32_t capacity, int CH,
                                const float* d_true_dev, int32_t* rank_count_dev, float* scratch_dev,
                                witw_stream_t stream);
int witw_topk_refine_spec_f32(const float* gal_spec_dev, const float* crop_inv_norm_dev,
                              const float* qry_spec_dev, const float* q_inv_norm_dev, int64_t G, int64_t Q,
                              int CH, const int32_t* cand_idx_dev, int kc, int32_t g_index_offset, int k_out,
                              float* topk_dist_dev, int32_t* topk_idx_dev, void* scratch_dev,
                              witw_stream_t stream);

/* The whole-gallery sweep in the azimuth-frequency domain (csrc/match_spec.cu): the same contract as
 * witw_match_tc -- correlation -> crop_overhead -> l2_distance of cvig_fov.py:297-363, the rank rule of :552 and a
 * per-query top-k -- but the circular correlation is evaluated through the correlation theorem:
 *     corr[g,q,:] = irfft( sum_r O[g,r,f] conj(S[q,r,f]) )
 * The per-frequency products (33 bins x 64 feature rows of complex MACs, 16.9 kFLOP per pair instead of 524 kFLOP at
 * 360 degrees) run on tcgen05 with bf16 spectra and fp32 accumulation in TMEM; the 64-point inverse real FFT, the
 * argmax over the shift and everything after it run in the epilogue, register-local.  Needs C*H == 64 and W == 64.
 *   witw_spec_gallery_prep: ov [G,64,64] fp32 -> bf16 spectra in the operand layout (16 KB per item, groups of 8
 *        items) and crop_inv_norm [G rounded up to 8, 64].  g_first = index of ov[0] inside the operand (a multiple
 *        of 8 unless it continues a partial group), so an encode loop can append batch by batch.
 *   witw_spec_query_prep: su [Q,64,sw] fp32 -> bf16 spectra of the zero-padded rows, scaled by 1/64 (8 KB per
 *        query, laid out per tile of 128 queries), and q_inv_norm [Q].
 *   spec_out (both, optional): the fp32 spectra of the same rows in the layout of witw_spectral_rows_f32
 *        ([G*64,64] / [Q*64,64]) -- the operands of the exact finish, produced in the same pass.
 *   witw_match_spec: arguments as witw_match_tc; top-k candidate lists: witw_match_spec_topk_slots(). */
int witw_spec_supported(int CH, int W, int sw);
size_t witw_spec_gallery_operand_bytes(int64_t G, int CH);
size_t witw_spec_query_operand_bytes(int64_t Q, int CH);
int witw_spec_gallery_prep(const float* ov_dev, int64_t G, int64_t g_first, int CH, int W, int sw,
                           void* gal_op_dev, float* crop_inv_norm_dev, float* spec_out_dev /* optional */,
                           witw_stream_t stream);
int witw_spec_query_prep(const float* su_dev, int64_t Q, int CH, int sw, void* qry_op_dev,
                         float* q_inv_norm_dev, float* spec_out_dev /* optional */, witw_stream_t stream);
int witw_match_spec_topk_slots(int64_t G, int64_t Q);
int witw_match_spec(const void* gal_op_dev, const float* crop_inv_norm_dev, const void* qry_op_dev,
                    const float* q_inv_norm_dev, int64_t G, int64_t Q, int CH, int sw,
                    float* dist_dev, uint8_t* ori_dev, const float* d_true_dev,
                    const int32_t* true_idx_dev, int32_t* rank_count_dev, int topk,
                    float* topk_dist_dev, int32_t* topk_idx_dev, int32_t g_index_offset,
                    float recheck_band, int64_t* recheck_g_dev, int64_t* recheck_q_dev,
                    int32_t* recheck_count_dev, int32_t recheck_capacity, witw_stream_t stream);

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
