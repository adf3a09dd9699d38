/* libtmae_sm100.so -- C ABI of the B200-native (sm_100a) T-MAE sparse-window voxel-encoder hot path.
 *
 * Drop-in boundary.  The reference reaches this path through
 *   - its pybind module `sst_ops_cuda` (pcdet/ops/sst_ops/src/sst_ops_api.cpp:6-8), and
 *   - library calls (torch_scatter, spconv, pytorch3d, ATen/cuBLAS) made from the pcdet modules
 *     vfe (temporal_dyn_vfe.py / dyn_vfe.py) and backbone_3d (SiamWCA.py / SiamWCA_MAE.py).
 * Every entry point below names the reference interface it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes stub a pcdet maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory unless marked HOST;
 *   - the caller owns every buffer, including the workspace (query *_workspace_bytes first);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - returns 0 on success, a negative TMAE_ERR_* otherwise (never exits the process, unlike
 *     pcdet/ops/sst_ops/src/sst_ops.cpp:7-19); tmae_last_error_string() describes the failure;
 *   - the data path keeps no mutable global state: the error string is thread-local, per-(kernel, device) attribute
 *     flags are mutex-protected, counters are atomics.  Process-wide, opt-in and NOT thread-safe: the measurement
 *     switches of tmae_set_option, the event profiler (tmae_profile_begin/_end) and the GEMM timeline
 *     (tmae_debug_set_trace) -- diagnostics, never needed for results;
 *   - element counts are int64_t; index tensors use the reference's dtypes where they cross the
 *     module API (int64 coords / inverse indices) and int32 internally.
 *   - features are fp32 row-major (rows, channels) in the fp32 and tf32 modes; the bf16-storage mode (tmae_bf16_*
 *     entry points) keeps encoder activations and their gradients in bf16, statistics / accumulators / weights'
 *     master copies and weight gradients in fp32.
 */
#ifndef TMAE_SM100_H_
#define TMAE_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TMAE_API __attribute__((visibility("default")))
#else
#define TMAE_API
#endif

#define TMAE_ABI_VERSION 1

#define TMAE_ERR_INVALID_ARG (-1)
#define TMAE_ERR_CUDA (-2)
#define TMAE_ERR_UNSUPPORTED (-3)

/* GEMM arithmetic selector for the ops that take `precision`. */
#define TMAE_PREC_FP32 0 /* fp32 FFMA, parity mode (rtol 1e-5 vs the fp32 oracle) */
#define TMAE_PREC_TF32 1 /* tensor-core mode with fp32 STORAGE: tcgen05.mma.kind::tf32, fp32 accumulate in TMEM; operands stay fp32
                            in HBM and are read through TMA (dense) or cp.async gathers (sparse conv) as TF32 */
#define TMAE_PREC_BF16 2 /* tensor-core mode with bf16 STORAGE (the throughput mode): tcgen05.mma.kind::f16 on bf16 operands, fp32
                            accumulate in TMEM, bf16 activations out of the TMEM epilogue; only the tmae_bf16_* entry points take it */

#define TMAE_ACT_NONE 0
#define TMAE_ACT_GELU 1 /* exact erf GELU (torch default, sst_basic_block.py:121-122) */
#define TMAE_ACT_RELU 2
#define TMAE_ACT_GELU_DERIV 3 /* bf16 mode only: y = GELU(.), and `preact` receives GELU'(x w^T + bias) instead of the pre-activation */
/* flag bits of tmae_bf16_linear_bwd_data's `accumulate` argument */
#define TMAE_BWD_ACCUMULATE 1          /* dx += */
#define TMAE_BWD_PRE_IS_DERIVATIVE 2   /* gelu_pre holds GELU'(pre-activation) (a TMAE_ACT_GELU_DERIV forward wrote it): multiply, do not re-derive */

#define TMAE_MAX_LEVELS 8
#define TMAE_WIN_TOKENS 64 /* 8 x 8 x 1 windows (t_mae_ssl.yaml:61) */

/* ---- library ------------------------------------------------------------------------------- */
TMAE_API const char* tmae_last_error_string(void);
TMAE_API int tmae_version(void);
TMAE_API int tmae_device_check(void); /* 0 iff the current device is sm_100 */
/* Per-kernel timing with CUDA events on the launching stream (used by bench.py for the roofline object): begin, run some
 * steps, end -> text table "name calls total_ms algorithmic_flops algorithmic_bytes" per kernel family. */
TMAE_API int64_t tmae_launch_count(void); /* instrumented kernel launches so far (lower bound of all launches) */
/* Which kernel served the GEMM-shaped calls so far: out[0] TMA-fed tcgen05 kernel, out[1] fp32 FFMA kernel in parity mode,
 * out[2] fp32 FFMA kernel taken in a tensor-core mode because TMA cannot express the shape (must stay 0 on the hot path),
 * out[3] thin-k kernel (k <= 16: the first VFE layer).  n <= 4 entries are written. */
TMAE_API int tmae_dispatch_counts(int64_t* out, int32_t n);
TMAE_API void tmae_profile_begin(void);
TMAE_API int64_t tmae_profile_end(char* buf, int64_t cap);
TMAE_API int64_t tmae_scan_scratch_elems(int64_t n);
TMAE_API int tmae_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total, int32_t* scratch, void* stream);

/* ---- A1-A2 dynamic voxelisation -------------------------------------------------------------
 * Replaces common_utils.get_in_range_mask (pcdet/utils/common_utils.py:66-76), points[keep],
 * coords.unique(dim=0, return_inverse=True) and torch_scatter.scatter(reduce='mean')
 * (pcdet/models/backbones_3d/vfe/temporal_dyn_vfe.py:69-85, dyn_vfe.py:66-82).
 *   points        (n_points, point_stride) f32  [b, x, y, z, feat...]
 *   range_lo, voxel  HOST float[3];  grid HOST int32[3] = [gx, gy, gz]
 * Outputs are allocated by the caller at capacity n_points rows (mcap = min(n_points, cells) voxel
 * rows); the first counts[0] / counts[1] rows are valid:
 *   points_out    (n_points, point_stride) f32   kept points, original order
 *   point_coords  (n_points, 4) i64  [b, z, y, x]
 *   inverse       (n_points,)  i64   voxel row of every kept point
 *   voxel_coords  (mcap, 4) i64      lexicographic (b,z,y,x) = torch.unique(dim=0) order
 *   voxel_mean    (mcap, point_stride-1) f32
 *   voxel_npts    (mcap,) i32 ; voxel_offset (mcap+1,) i32 ; pt_order (n_points,) i32
 *                 CSR of kept-point rows per voxel, ascending inside a voxel (canonical slot order)
 *   counts        (2 + batch,) i64: [n_kept, n_voxels, first voxel row of sample 0..batch-1]
 */
TMAE_API size_t tmae_voxelize_workspace_bytes(int64_t n_points, int32_t batch, const int32_t* grid);
TMAE_API int tmae_voxelize(const float* points, int64_t n_points, int32_t point_stride, const float* range_lo, const float* voxel,
                  const int32_t* grid, int32_t batch, float* points_out, int64_t* point_coords, int64_t* inverse,
                  int64_t* voxel_coords, float* voxel_mean, int32_t* voxel_npts, int32_t* voxel_offset, int32_t* pt_order,
                  int64_t* counts, void* workspace, size_t workspace_bytes, void* stream);

/* A3 (first half): per-point VFE input  [f_center(3), x,y,z,feat..., f_cluster(3)]
 * (temporal_dyn_vfe.py:89-110); features is (n_kept, point_stride - 1 + 6) f32. */
TMAE_API int tmae_vfe_point_features(const float* points_kept, int64_t n_kept, int32_t point_stride, const int64_t* point_coords,
                            const int64_t* inverse, const float* voxel_mean, const float* range_lo, const float* voxel,
                            float* features, void* stream);

/* ---- A4-A6 window partition -------------------------------------------------------------------
 * Replaces get_window_coors (pcdet/models/model_utils/sst_utils.py:6-58), drop_single_shift / drop_voxel
 * (pcdet/models/backbones_3d/spt_backbone.py:47-135), the temporal form drop_single_shift_ref_to_prv
 * (pcdet/models/backbones_3d/SiamWCA.py:65-199), get_inner_win_inds (pcdet/ops/sst_ops/sst_ops_utils.py:5-12
 * -> src/sst_ops_gpu.cu:14-20) and make_continuous_inds / get_flat2win_inds (sst_utils.py:61-115), for BOTH
 * shifts in one call.  coords_* are (m,3) i32 [b,y,x] in ascending lexicographic order (what the voxeliser
 * and the strided-conv table emit).  coords_b == NULL: single frame (SSTInputLayer); otherwise temporal
 * (frame a = current, b = previous): level from max(count_a, count_b), windows empty in either frame are
 * dropped.  wcap = tmae_partition_window_capacity().  Outputs, leading dimension 2 = shift:
 *   win_*   [2][m]  compact window id (level-major order) or -1 if the voxel is dropped
 *   slot_*  [2][m]  canonical inner-window slot (stable rank by element index)
 *   posidx_*[2][m]  u8  ly*8+lx  (row of the 64-entry position-embedding table)
 *   tok_*   [2][wcap*64] voxel row per (window, slot);  cnt_* [2][wcap] tokens per window
 *   win_level [2][wcap] ; n_win [2] ; level_base [2][n_levels+1] first compact id of each level
 *   status  [1]  bit0 coords not ascending, bit1 out-of-grid coordinate, bit2 count in no level,
 *                bit3 a window exceeds its level's max_tokens (voxel dropping is unsupported)
 *   ref_*   optional (NULL to skip) reference-format tables: batch_win_inds, voxel_drop_level,
 *           flat2win index (= compact id inside the level * max_tokens + slot), all [2][m] i64
 */
TMAE_API int64_t tmae_partition_window_capacity(int32_t batch, int32_t grid_x, int32_t grid_y);
TMAE_API size_t tmae_window_partition_workspace_bytes(int32_t batch, int32_t grid_x, int32_t grid_y, int32_t n_levels);
TMAE_API int tmae_window_partition(const int32_t* coords_a, int64_t m_a, const int32_t* coords_b, int64_t m_b, int32_t batch,
                          int32_t grid_x, int32_t grid_y, int32_t n_levels, const int32_t* lvl_lo, const int32_t* lvl_hi,
                          const int32_t* lvl_tokens,
                          int32_t* win_a, int32_t* slot_a, uint8_t* posidx_a, int32_t* tok_a, int32_t* cnt_a,
                          int32_t* win_b, int32_t* slot_b, uint8_t* posidx_b, int32_t* tok_b, int32_t* cnt_b,
                          int32_t* win_level, int32_t* n_win, int32_t* level_base, int32_t* status,
                          int64_t* ref_bwi_a, int64_t* ref_lvl_a, int64_t* ref_f2w_a,
                          int64_t* ref_bwi_b, int64_t* ref_lvl_b, int64_t* ref_f2w_b,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense contractions (fp32 parity mode) ------------------------------------------------------
 * Replace torch.nn.functional.linear at cosine_msa.py:57-62,431, sst_basic_block.py:81, wca_block.py:99,
 * network_utils.py:30, SiamWCA_MAE.py:117-119 and their autograd backward.  w is (n, k) row-major
 * (torch Linear layout).  y = act(x w^T + bias) + residual ; preact (nullable) receives x w^T + bias. */
/* options = measurement switches (A/B runs; defaults are the measured best): "wide_st" (1: 256-bit epilogue stores in the TMA GEMM),
 * "attn_occ" / "attn_occ_fwd" (1: mma attention kernels compiled for more resident CTAs per SM), "bn_colsum_cap" (0 = automatic
 * blocks per SM of the BatchNorm column sums), "ln_bwd_cap" (6 blocks per SM of the LayerNorm backward).
 * The Python binding applies TMAE_OPT_<NAME>=<int> environment variables through this call when the library is loaded. */
TMAE_API int tmae_set_option(const char* name, int32_t value);
/* Diagnostics: when device_u64 is non-null, CTA 0 of every following TMA GEMM launch writes a timeline of its TMA
 * producer, MMA issuer and epilogue (four equal slices of `capacity` 64-bit words: event << 56 | index << 40 | SM clock).
 * Pass NULL to switch it off.  Not part of the data path (no reference counterpart). */
TMAE_API int tmae_debug_set_trace(void* device_u64, int64_t capacity);
TMAE_API int tmae_linear_fwd(const float* x, const float* w, const float* bias, const float* residual, float* y, float* preact,
                    int64_t m, int64_t n, int64_t k, int32_t act, int32_t precision, void* stream);
/* y[m, :] = x[m, :] @ w^T + lut[rowidx[m], :]: the packed q/k/v projection of an encoder layer with the window position
 * embedding folded in.  q = (x + pos) Wq^T + bq = x Wq^T + (pos_lut Wq^T + bq)[posidx] because the embedding depends only on
 * the voxel's cell inside its 8x8 window (64 rows, spt_backbone.py:186-231); lut = tmae_pos_table output, (64, n). */
TMAE_API int tmae_linear_fwd_lut(const float* x, const float* w, const float* lut, const uint8_t* rowidx, float* y, int64_t m, int64_t n,
                        int64_t k, int32_t precision, void* stream);
/* table[p, j] = (j < n_pos ? sum_c pos_lut[p, c] * w[j, c] : 0) + bias[j]   for p < 64, j < n   (always fp32) */
TMAE_API int tmae_pos_table(const float* pos_lut, const float* w, const float* bias, float* table, float* table_t, int32_t n, int32_t n_pos,
                   int32_t c, void* stream);
/* The tensor-core form of the same projection: y = x w^T + x2 w2^T in ONE GEMM whose reduction runs over k then k2
 * (x2 = tmae_onehot64(posidx) (m, 64), w2 = table_t (n, 64) from tmae_pos_table, so x2 w2^T = table[posidx]). */
TMAE_API int tmae_linear_fwd_dual(const float* x, const float* w, const float* x2, const float* w2, float* y, int64_t m, int64_t n, int64_t k,
                         int64_t k2, int32_t precision, void* stream);
TMAE_API int tmae_onehot64(const uint8_t* idx, float* out, int64_t m, void* stream);
/* backward of the table bias: dtable[p, :] = sum over rows with rowidx == p of dy[row, :]  (64 x n, overwritten);
 * dbias[j] = sum_p dtable[p, j];  dw[j, :] += sum_p dtable[p, j] * pos_lut[p, :] for j < n_pos (adds to dw).
 * transposed = 1: dtable is (n, 64) = dy^T onehot, i.e. tmae_linear_bwd_weight(dy, onehot) -- the tensor-core form. */
TMAE_API int tmae_binned_colsum(const float* dy, const uint8_t* rowidx, float* dtable, int64_t rows, int32_t n, void* stream);
TMAE_API int tmae_pos_table_bwd(const float* dtable, int32_t transposed, const float* pos_lut, float* dw, float* dbias, int32_t n, int32_t n_pos,
                       int32_t c, void* stream);
TMAE_API int tmae_linear_bwd_data(const float* dy, const float* w, float* dx, int64_t m, int64_t n, int64_t k, int32_t accumulate,
                         int32_t precision, void* stream);
TMAE_API int tmae_linear_bwd_data_gelu(const float* dy, const float* w, const float* preact, float* dx, int64_t m, int64_t n, int64_t k,
                              int32_t precision, void* stream); /* dx = (dy w) * gelu'(preact) */
TMAE_API int tmae_linear_bwd_weight(const float* dy, const float* x, float* dw, float* dbias, int64_t m, int64_t n, int64_t k,
                           int32_t precision, void* stream);
TMAE_API int tmae_colsum(const float* x, float* out, int64_t rows, int32_t cols, void* stream);
TMAE_API int tmae_gelu_bwd(const float* dy, const float* preact, float* dx, int64_t n, void* stream);

/* ---- A9 sparse 2-D convolution ------------------------------------------------------------------
 * Replaces spconv SubMConv2d / SparseConv2d inside post_act_block (pcdet/utils/spconv_utils.py:37-56).
 * Tables: (rows, 9) i32 input row per tap (ky*3+kx) or -1.  Weights (cout, 3, 3, cin) = (cout, taps, cin). */
TMAE_API size_t tmae_subm_table_workspace_bytes(int32_t batch, int32_t y, int32_t x);
TMAE_API int tmae_subm_table(const int32_t* indices, int64_t rows_cap, const int32_t* rows_dev, int32_t batch, int32_t y, int32_t x,
                    int32_t* table, void* workspace, size_t workspace_bytes, void* stream);
TMAE_API size_t tmae_strided_table_workspace_bytes(int32_t batch, int32_t y_in, int32_t x_in);
TMAE_API int tmae_strided_table(const int32_t* indices, int64_t rows_cap, const int32_t* rows_dev, int32_t batch, int32_t y_in, int32_t x_in,
                       int32_t* indices_out, int64_t out_cap, int32_t* n_out, int32_t* table, int32_t* table_t, void* workspace,
                       size_t workspace_bytes, void* stream);
TMAE_API int tmae_sparse_conv_fwd(const float* x, const int32_t* table, const float* w, float* y, int64_t rows_out, int32_t taps,
                         int32_t cin, int32_t cout, int32_t accumulate, int32_t precision, void* stream);
TMAE_API int tmae_sparse_conv_bwd_weight(const float* dy, const float* x, const int32_t* table, float* dw, int64_t rows_out, int32_t taps,
                                int32_t cin, int32_t cout, int32_t precision, void* stream);
TMAE_API int tmae_transpose_taps(const float* w, float* wt, int32_t cout, int32_t taps, int32_t cin, int32_t flip, void* stream);

/* ---- row-wise kernels ---------------------------------------------------------------------------- */
/* y = x + lut[posidx]  : position embedding (spt_backbone.py:186-231) as a 64-row table lookup */
TMAE_API int tmae_add_pos(const float* x, const uint8_t* posidx, const float* lut, float* y, int64_t rows, int32_t c, void* stream);
/* y = LayerNorm(x + res) ; rows with rowmask == 0 ignore res (wca_block.py:96-98).  (sst_basic_block.py:78-83)
 * backward: dv = grad wrt (x + res) ; dres (nullable) = dv on rows whose res was used, else 0 */
TMAE_API int tmae_add_layernorm_fwd(const float* x, const float* res, const uint8_t* rowmask, const float* gamma, const float* beta, float* y,
                           float* mean, float* rstd, int64_t rows, int32_t c, float eps, void* stream);
TMAE_API int tmae_add_layernorm_bwd(const float* dy, const float* x, const float* res, const uint8_t* rowmask, const float* gamma,
                           const float* mean, const float* rstd, float* dv, float* dres, float* dgamma, float* dbeta, int64_t rows,
                           int32_t c, void* stream);
/* same, and dcolsum[c] (nullable; overwritten) = column sums of dres when dres is written, else of dv: the bias gradient of the
 * linear layer that produced res (out_proj / linear2, sst_basic_block.py:78-83), folded into the pass that computes it */
TMAE_API int tmae_add_layernorm_bwd_colsum(const float* dy, const float* x, const float* res, const uint8_t* rowmask, const float* gamma,
                                  const float* mean, const float* rstd, float* dv, float* dres, float* dgamma, float* dbeta, float* dcolsum,
                                  int64_t rows, int32_t c, void* stream);
/* BatchNorm1d (+ReLU) over rows: network_utils.py:31 (VFE), spconv_utils.py:50-54 (after sparse convs) */
TMAE_API size_t tmae_bn_workspace_bytes(int32_t c);
TMAE_API int tmae_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                      float eps, float* y, float* save_mean, float* save_rstd, int64_t rows, int32_t c, int32_t relu,
                      void* workspace, size_t workspace_bytes, void* stream);
TMAE_API int tmae_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y, int64_t rows,
                  int32_t c, int32_t relu, void* stream);
/* y may be NULL: the ReLU mask is recomputed from x with the forward's own expression (beta is needed for that) */
TMAE_API int tmae_bn_bwd(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                float* dx, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t relu, int32_t training, void* workspace,
                size_t workspace_bytes, void* stream);
/* BatchNorm2d (+ReLU) of the dense decoder (SiamWCA_MAE.py:79-115, nn.BatchNorm2d eps 1e-3 momentum 0.01) on a
 * channels-last bf16 map viewed as (B*Y*X, C) rows with explicit row pitches (elements): the output may be a column
 * slice of the concatenated (rows, 384) decoder buffer, which replaces torch.cat (SiamWCA_MAE.py:246-249).
 * Statistics and arithmetic in fp32 / double; training = 0 applies the given mean / rstd. */
TMAE_API int tmae_bn_bf16_fwd(const void* x, int64_t ldx, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float momentum, float eps, void* y, int64_t ldy, float* mean, float* rstd, int64_t rows, int32_t c, int32_t relu,
                     int32_t training, void* workspace, size_t workspace_bytes, void* stream);
TMAE_API int tmae_bn_bf16_bwd(const void* dy, int64_t ldd, const void* x, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                     const float* beta, void* dx, int64_t ldo, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t relu,
                     int32_t training, void* workspace, size_t workspace_bytes, void* stream);
/* per-voxel max over its points: torch_scatter.scatter_max (temporal_dyn_vfe.py:113) through the CSR */
TMAE_API int tmae_segment_max_fwd(const float* x, const int32_t* voxel_offset, const int32_t* pt_order, int64_t n_voxels, int32_t c, float* out,
                         int32_t* argmax, void* stream);
TMAE_API int tmae_segment_max_bwd(const float* dout, const int32_t* argmax, int64_t n_voxels, int32_t c, float* dx, int64_t n_points,
                         void* stream);
/* SparseConvTensor.dense() as a channels-last (B,Y,X,C) map (SiamWCA_MAE.py:235) and its inverse gather
 * (also spatial_features[b,y,x] at SiamWCA_MAE.py:311-312) */
TMAE_API int tmae_densify_nhwc(const float* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x,
                      float* dense, int32_t zero_fill, void* stream);
TMAE_API int tmae_gather_nhwc(const float* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, float* rows, void* stream);
/* same with a bf16 dense map (the cuDNN decoder runs bf16 channels-last in throughput mode) */
TMAE_API int tmae_densify_nhwc_bf16(const float* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x,
                           void* dense, int32_t zero_fill, void* stream);
TMAE_API int tmae_gather_nhwc_bf16(const void* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, float* rows, void* stream);
TMAE_API int tmae_gather_rows(const float* src, const int32_t* sel, int64_t m, int32_t c, float* out, void* stream);
TMAE_API int tmae_scatter_rows(const float* rows, const int32_t* sel, int64_t m, int32_t c, float* dst, void* stream);

/* ---- A7-A8 windowed cosine attention core -------------------------------------------------------
 * Replaces flat2window + _scaled_cosine_attention + window2flat (cosine_msa.py:114-176, sst_basic_block.py:22-54,
 * wca_block.py:26-67).  q/k/v are the PROJECTED rows (rows, C) in flat voxel order; tok/cnt tables come from
 * tmae_window_partition (self: q and k tables are the same; cross: q = current frame, k = previous frame).
 * max_windows bounds the grid; the live count is read from n_win on the device.  small_end (device i32) = number of
 * leading windows that hold <= 16 tokens on both sides (= level_base[first level with max_tokens > 16]; windows are
 * level-sorted): those run one warp per window, the rest one CTA per window with shared memory sized for 32 tokens
 * (windows [small_end, mid_end), mid_end = level_base[first level with max_tokens > 32]) or 64.  Backward: dsum (q rows,
 * heads) scratch.  ld_q / ld_k / ld_v: row pitches (elements, >= channels, multiples of 4) of q / k / v and of
 * dq / dk / dv, so the three may be column blocks of one packed projection output (rows, 3*channels); o, dout: channels.
 * precision: TMAE_PREC_FP32 = IEEE math, fp32 FFMA kernels for > 16-token windows; TMAE_PREC_TF32 = MUFU math, mma.sync TF32.
 * rows_q / rows_kv: row counts of q and k/v (only used for the profiler's algorithmic byte count; 0 = unknown). */
TMAE_API int tmae_window_attention_fwd(const float* q, const float* k, const float* v, float* o, float* lse, const int32_t* qtok,
                              const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                              const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                              int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                              int32_t precision, void* stream);
TMAE_API int tmae_window_attention_bwd(const float* dout, const float* q, const float* k, const float* v, const float* o, const float* lse,
                              float* dsum, float* dq, float* dk, float* dv, float* dtau, const int32_t* qtok, const int32_t* qcnt,
                              const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win, const int32_t* small_end,
                              const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min, int32_t channels,
                              int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv, int32_t precision,
                              void* stream);

/* ---- N2 / N3 (the step in front of the VFE): batch assembly of one frame set on the GPU ---------------
 * raw (n_points, feats) fp32 [x, y, z, feat...] = the samples' point arrays back to back, sample_offsets (batch + 1) i64.
 * Per point: remove_ego_points (once_utils.py:43-45; |x| < r and |y| < r on the RAW coordinates, r = 2:
 * once_temporal_dataset.py:167-168), convert_prv_frame_to_cur (once_utils.py:4-29) as two float64 affine maps per sample
 * (xform (batch, 2, 3, 4) row-major: prev -> global, global -> current = the reference's np.linalg.inv result, computed
 * on the host; xform_flags (batch, 2): 0 skips a map, which is what the reference does for an all-zero pose; both NULL
 * for the current frame), mask_points_by_range (common_utils.py:124-127; closed interval, compared in float64 like the
 * reference), collate_batch (dataset.py:203-208; sample index in column 0) and the final .float().  Stable compaction:
 * out (capacity n_points, 1 + feats) holds the kept points in input order, *count (device i64) how many; rows
 * [count, n_points) are set to the out-of-range sentinel [0, 1e6, 1e6, 1e6, 0...] which tmae_voxelize's range test drops,
 * so a caller may pass all n_points rows on without reading *count.  Two launches, no atomics; workspace = one i32 per
 * 512 points. */
TMAE_API size_t tmae_assemble_frames_workspace_bytes(int64_t n_points);
TMAE_API int tmae_assemble_frames(const float* raw, const int64_t* sample_offsets, int32_t batch, int32_t feats, const double* xform,
                         const uint8_t* xform_flags, float ego_radius, const float* crop_xyxy, float* out, int64_t* count,
                         int64_t n_points, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the reference's native op, 1:1 (pcdet/ops/sst_ops/src/sst_ops_api.cpp:6-8) ------------------------------------------------------
 * Same tensors as sst_ops_cuda.ingroup_inds_wrapper(group_inds i64 (N,), out_inds i64 (N,)) and
 * sst_ops_cuda.group_inner_inds_wrapper(inverse_inds i64 (N,), group_inds i64 (M,K)) (sst_ops.cpp:21-48), so sst_ops_utils.py:5-27 binds
 * them unchanged.  Results are the CANONICAL ones (what a serial run of sst_ops_gpu.cu:14-39 yields): out_inds[i] = number of earlier
 * elements with the same group id; group_inds[g] = the first K element indices of group g in ascending order, cyclically padded
 * (group_inds[g][i] = group_inds[g][i mod cnt]); rows of groups without elements are left untouched (the reference pre-fills -1).
 * No host sync, no allocation: workspace from tmae_sst_ops_workspace_bytes(N).  N < 2^31. */
TMAE_API size_t tmae_sst_ops_workspace_bytes(int64_t n);
TMAE_API int tmae_ingroup_inds(const int64_t* group_inds, int64_t* out_inds, int64_t n, void* workspace, size_t workspace_bytes, void* stream);
TMAE_API int tmae_group_inner_inds(const int64_t* inverse_inds, int64_t n, int64_t* group_inds, int64_t m, int32_t k, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ---- A10-A11 reconstruction target + Chamfer loss ------------------------------------------------
 * Replaces sst_ops_cuda.group_inner_inds_wrapper (pcdet/ops/sst_ops/src/sst_ops_api.cpp:8, sst_ops_gpu.cu:22-39),
 * the GT gather / centre subtraction (SiamWCA_MAE.py:132-139) and pytorch3d chamfer_distance (SiamWCA_MAE.py:163). */
TMAE_API int tmae_gt_group(const float* points_kept, int32_t point_stride, const int32_t* voxel_offset, const int32_t* pt_order,
                  const int64_t* voxel_coords, const float* range_lo, const float* voxel, int64_t n_voxels, int32_t k, float* gt,
                  int64_t* group_inds, void* stream);
TMAE_API size_t tmae_chamfer_workspace_bytes(void);
TMAE_API int tmae_chamfer_fwd(const float* pred, const float* gt, const float* w, int64_t n_voxels, int32_t p1, int32_t p2,
                     const float* points_kept, int32_t point_stride, const int32_t* voxel_offset, const int32_t* pt_order,
                     const int64_t* voxel_coords, const float* range_lo, const float* voxel, float* loss, void* state, void* stream);
TMAE_API int tmae_chamfer_bwd(const float* grad_loss, const float* pred, const float* gt, const float* w, int64_t n_voxels, int32_t p1,
                     int32_t p2, const float* points_kept, int32_t point_stride, const int32_t* voxel_offset,
                     const int32_t* pt_order, const int64_t* voxel_coords, const float* range_lo, const float* voxel,
                     const void* state, float* dpred, void* stream);

/* ---- A7/A8 whole encoder layer ---------------------------------------------------------------------
 * One call = one EncoderLayer.forward of pcdet/models/model_utils/sst_basic_block.py:58-84 (x_kv == NULL: windowed
 * self-attention) or wca_block.py:70-103 (x_kv = previous-frame rows: temporal window cross-attention):
 *   q = (x + pos) Wq, k = (x_kv + pos_kv) Wk, v = x_kv Wv -> cosine attention per window -> out_proj
 *   x1 = LN1(x + attn [rows with rowmask == 0 skip the attention term]) ; y = LN2(x1 + W2 gelu(W1 x1)).
 * `saved` (tmae_encoder_layer_saved_bytes) receives the intermediates backward needs; backward writes the input
 * gradient(s) and the 13 parameter gradients through a second tmae_layer_params whose pointers are gradient buffers. */
typedef struct tmae_layer_params {
  const float *in_w, *in_b, *out_w, *out_b, *tau, *ln1_g, *ln1_b, *w1, *b1, *w2, *b2, *ln2_g, *ln2_b;
} tmae_layer_params;
typedef struct tmae_layer_tables {
  const uint8_t* posidx_q;  /* (m_q)  row of the position table per query row  */
  const uint8_t* posidx_kv; /* (m_kv) cross only */
  const int32_t *qtok, *qcnt, *ktok, *kcnt, *n_win, *small_end, *mid_end; /* from tmae_window_partition (one shift) */
  const uint8_t* rowmask;   /* (m_q) cross only: 1 = row belongs to a paired window */
  int64_t max_windows;
  const float* onehot_q;    /* (m_q, 64) tmae_onehot64(posidx_q); NULL = use the table-bias epilogue (fp32 SIMT path) */
  const float* onehot_kv;   /* (m_kv, 64) cross only */
} tmae_layer_tables;
TMAE_API size_t tmae_encoder_layer_saved_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross);
TMAE_API size_t tmae_encoder_layer_scratch_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross);
TMAE_API int tmae_encoder_layer_fwd(const float* x, const float* x_kv, const tmae_layer_params* P, const tmae_layer_tables* T, const float* pos_lut,
                           float tau_min, float eps, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t precision,
                           int32_t need_backward, float* y, void* saved, size_t saved_size, void* stream);
TMAE_API int tmae_encoder_layer_bwd(const float* dy, const float* x, const float* x_kv, const tmae_layer_params* P, const tmae_layer_tables* T,
                           const float* pos_lut, float tau_min, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t precision,
                           const void* saved, size_t saved_size, float* dx, float* dx_kv, const tmae_layer_params* G, void* scratch,
                           size_t scratch_size, void* stream);

/* ---- N4: CenterHead box decoding + rotated BEV NMS ------------------------------------------------------------------------
 * Replaces centernet_utils.decode_bbox_from_heatmap (pcdet/models/model_utils/centernet_utils.py:154-220) after its top-K selection,
 * iou3d_nms_cuda.boxes_iou_bev_gpu / nms_gpu (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp, iou3d_nms_kernel.cu:236-311, iou3d_nms.cpp:90-135)
 * and the [:NMS_POST_MAXSIZE] cut of model_nms_utils.class_agnostic_nms (:6-25).  Boxes are (.., 7) f32 [x, y, z, dx, dy, dz, heading].
 * Unlike the reference, the suppression sweep runs on the device (no mask copy to the host, no cudaMalloc per call), batched over samples. */
TMAE_API int tmae_boxes_iou_bev(const float* boxes_a, int64_t na, const float* boxes_b, int64_t nb, float* ans_iou, void* stream);
TMAE_API size_t tmae_nms_bev_workspace_bytes(int32_t samples, int64_t cap);
/* boxes (samples, cap, 7): each sample's first counts[s] rows in descending score order -> keep (samples, cap) i64 = kept row indices (the
 * first num_keep[s] <= post_max entries are valid), num_keep (samples) i32 */
TMAE_API int tmae_nms_bev(const float* boxes, const int32_t* counts, int32_t samples, int64_t cap, float thresh, int32_t post_max, int64_t* keep,
                 int32_t* num_keep, void* workspace, size_t workspace_bytes, void* stream);
/* one head: scores / inds (batch, k) = top-k of sigmoid(hm).flatten(1) (descending; index = class * h * w + y * w + x); head maps NCHW f32
 * (dim = log sizes, rot = [cos, sin], iou nullable); class_map (ncls) i64 nullable; voxel_size / range_lo HOST float[2+], limit_range HOST
 * float[6].  Outputs at capacity k per sample, the first counts[b] rows valid, score order preserved. */
TMAE_API int tmae_centerhead_decode(const float* scores, const int64_t* inds, const float* center, const float* center_z, const float* dim,
                           const float* rot, const float* iou, const int64_t* class_map, int32_t batch, int32_t k, int32_t h, int32_t w,
                           int32_t ncls, float feature_map_stride, const float* voxel_size, const float* range_lo, const float* limit_range,
                           float score_thresh, float* boxes, float* out_scores, int64_t* labels, float* ious, int32_t* counts, void* stream);

/* ==== bf16-STORAGE mode (TMAE_PREC_BF16) ==========================================================================
 * Same reference functions as above (file:line cited there), for callers that keep encoder activations and their gradients
 * as bf16 rows (rows, channels) in HBM: what the reference itself does under its fp16 autocast
 * (tools/train_utils/train_utils.py:73-77; normalisation, softmax and LayerNorm statistics stay fp32: cosine_msa.py:151-176).
 * `void*` activation pointers are bf16 and must be 32-byte aligned; weights are bf16 copies of the fp32 master parameters
 * (tmae_cast_f32_bf16 / _multi); biases, LayerNorm / BatchNorm parameters, statistics and EVERY parameter gradient are fp32.
 * All GEMMs run on the TMA-fed tcgen05.mma.kind::f16 kernel (gemm_bf16.cu); there is no other implementation to fall back to. */
TMAE_API int tmae_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
TMAE_API int tmae_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
/* segs: DEVICE array of n_seg records {const float* src; void* dst; int64_t n} (24 bytes each): every cast in ONE launch */
TMAE_API int tmae_cast_f32_bf16_multi(const void* segs, int32_t n_seg, void* stream);
/* F.linear (cosine_msa.py:57-62,431; sst_basic_block.py:81): y = act(x w^T + bias) [+= y]; preact (nullable) = x w^T + bias */
TMAE_API int tmae_bf16_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* preact, int64_t m, int64_t n, int64_t k,
                         int32_t act, int32_t accumulate, void* stream);
/* packed q/k/v projection (cosine_msa.py:57-62 with q = k = x + pos, sst_basic_block.py:44): y = x w^T + table[posidx], the first
 * norm_cols columns (q, k) L2-normalised per head of width hd (F.normalize eps 1e-12, cosine_msa.py:151-152);
 * inv (m, norm_cols / hd) fp32 = 1 / max(|.|, 1e-12).  table (64, n) fp32 from tmae_pos_table. */
TMAE_API int tmae_bf16_qkv_fwd(const void* x, const void* w, const float* table, const uint8_t* posidx, void* y, float* inv, int64_t m, int64_t n,
                      int64_t k, int32_t norm_cols, int32_t hd, void* stream);
/* the same projection with the position term inside the MMA: y = [x | onehot] wcat^T; onehot (m, 64) bf16 from tmae_onehot64_bf16,
 * wcat (n, k + 64) bf16 = [w | table^T] from tmae_bf16_qkv_wcat (k a multiple of 64).  What the fused layer uses. */
TMAE_API int tmae_bf16_qkv_fwd_onehot(const void* x, const void* onehot, const void* wcat, void* y, float* inv, int64_t m, int64_t n, int64_t k,
                             int32_t norm_cols, int32_t hd, void* stream);
TMAE_API int tmae_bf16_qkv_wcat(const float* pos_lut, const float* w, const float* bias, void* wcat, int32_t n, int32_t n_pos, int32_t c, void* stream);
/* the same for every layer of a model in ONE launch (the operand depends on the weights only: once per optimizer step).  segs: DEVICE array of
 * n_seg records {const float* pos_lut; const float* w; const float* bias; void* wcat; int64_t n, n_pos, c} (56 bytes); max_n = the largest n */
TMAE_API int tmae_bf16_qkv_wcat_multi(const void* segs, int32_t n_seg, int32_t max_n, void* stream);
/* out_proj / linear2 + residual + LayerNorm in one pass (sst_basic_block.py:78,83; wca_block.py:96-102):
 * v = res + (rowmask == NULL || rowmask[row] ? a w^T + bias : 0); y = LayerNorm(v) * gamma + beta; n in {128, 256}; v nullable */
TMAE_API int tmae_bf16_linear_ln_fwd(const void* a, const void* w, const float* bias, const void* res, const uint8_t* rowmask, const float* gamma,
                            const float* beta, float eps, void* v, void* y, float* mean, float* rstd, int64_t m, int64_t n, int64_t k,
                            void* stream);
/* dx = dy w [* gelu'(gelu_pre)] [+= dx] ;  dw (n, k) fp32 = dy^T x (overwrites); `accumulate` = TMAE_BWD_* flag bits */
TMAE_API int tmae_bf16_linear_bwd_data(const void* dy, const void* w, const void* gelu_pre, void* dx, int64_t m, int64_t n, int64_t k,
                              int32_t accumulate, void* stream);
/* onehot (m, 64) bf16 = tmae_onehot64_bf16(posidx), nullable: also dtab_t (n, 64) fp32 = dy^T onehot from the same pass over dy = the
 * binned column sums behind the position-table gradient (tmae_pos_table_bwd with transposed = 1); needs k % 64 == 0 */
TMAE_API int tmae_bf16_linear_bwd_weight(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n,
                                int64_t k, void* stream);
TMAE_API int tmae_onehot64_bf16(const uint8_t* idx, void* out, int64_t m, void* stream);
/* LayerNorm backward from the saved pre-norm sum v: dv = grad wrt v; dres (nullable) = dv on rows with rowmask != 0 else 0;
 * dcolsum (nullable, fp32) = column sums of dres when written, else of dv (the bias gradient of the producing linear layer) */
TMAE_API int tmae_bf16_layernorm_bwd(const void* dy, const void* v, const uint8_t* rowmask, const float* gamma, const float* mean, const float* rstd,
                            void* dv, void* dres, float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int32_t c, void* stream);
TMAE_API int tmae_bf16_colsum(const void* x, float* out, int64_t rows, int32_t cols, void* stream);
TMAE_API int tmae_bf16_binned_colsum(const void* dy, const uint8_t* rowidx, float* dtable, int64_t rows, int32_t n, void* stream);
/* sparse convolution (spconv_utils.py:37-56) on bf16 rows: gathered tcgen05 GEMMs; w (cout, taps, cin) bf16; dw fp32 */
TMAE_API int tmae_bf16_sparse_conv_fwd(const void* x, const int32_t* table, const void* w, void* y, int64_t rows_out, int32_t taps, int32_t cin,
                              int32_t cout, void* stream);
TMAE_API int tmae_bf16_sparse_conv_bwd_weight(const void* dy, const void* x, const int32_t* table, float* dw, int64_t rows_out, int32_t taps,
                                     int32_t cin, int32_t cout, void* stream);
TMAE_API int tmae_transpose_taps_bf16(const float* w, void* wt, int32_t cout, int32_t taps, int32_t cin, int32_t flip, void* stream);
/* SparseConvTensor.dense() / the BEV gather with 16-bit rows AND a 16-bit map */
TMAE_API int tmae_densify_nhwc_b16(const void* rows, const int32_t* indices, int64_t m, int32_t c, int32_t batch, int32_t y, int32_t x, void* dense,
                          int32_t zero_fill, void* stream);
TMAE_API int tmae_gather_nhwc_b16(const void* dense, const int32_t* indices, int64_t m, int32_t c, int32_t y, int32_t x, void* rows, void* stream);
/* dst = bf16(src * (col < norm_cols ? scale[row, col / hd] : 1)) */
TMAE_API int tmae_scale_cast_bf16(const float* src, const float* scale, void* dst, int64_t rows, int32_t n, int32_t norm_cols, int32_t hd, void* stream);
/* window attention core on the tensor cores (attention_tc.cu; see tmae_window_attention_fwd for the tables): q, k are UNIT vectors per head
 * (the output of tmae_bf16_qkv_fwd), q / k / v / o bf16 with row pitches ld_* (elements, multiples of 8), lse (rows_q, heads) fp32 */
TMAE_API int tmae_bf16_window_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* qtok,
                                   const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                                   const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                                   int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                                   void* stream);
/* backward of the same: dq / dk / dv (bf16, pitches ld_q / ld_k / ld_v) are gradients wrt the UN-normalised projections: the kernel takes
 * dq_hat, dk_hat back through the per-head L2 normalisation with inv_q / inv_k = 1 / |.| per (row, head) (row pitches ld_inv_*) as written
 * by tmae_bf16_qkv_fwd; P is recomputed from lse; dtau (device, fp32) accumulates the temperature gradient.  o = the forward output
 * (bf16, pitch channels): the <= 16-token windows (warp kernels) take D = dO . O from it; NULL routes every window to the tcgen05 tiles,
 * which form D from P and dP in registers. */
TMAE_API int tmae_bf16_window_attention_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse, const float* inv_q,
                                   int32_t ld_inv_q, const float* inv_k, int32_t ld_inv_k, void* dq, void* dk, void* dv, float* dtau,
                                   const int32_t* qtok, const int32_t* qcnt, const int32_t* ktok, const int32_t* kcnt, const int32_t* n_win,
                                   const int32_t* small_end, const int32_t* mid_end, int64_t max_windows, const float* tau, float tau_min,
                                   int32_t channels, int32_t heads, int32_t ld_q, int32_t ld_k, int32_t ld_v, int64_t rows_q, int64_t rows_kv,
                                   void* stream);
/* whole encoder layer (see tmae_encoder_layer_fwd): x, x_kv, y, dy, dx, dx_kv are bf16; P = fp32 master parameters (biases, LayerNorm,
 * tau and the weights behind the position table), W = bf16 copies of the four weight matrices, G = fp32 gradient buffers. */
typedef struct tmae_bf16_weights {
  const void *in_w, *out_w, *w1, *w2;
  const void* in_wcat; /* nullable: (3C, C + 64) bf16 [in_w | table^T] with n_pos = 2C (tmae_bf16_qkv_wcat), kept by the caller across calls while
                          the weights do not change; NULL: the layer rebuilds it per call inside `saved` */
} tmae_bf16_weights;
TMAE_API size_t tmae_bf16_encoder_layer_saved_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross);
TMAE_API size_t tmae_bf16_encoder_layer_scratch_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross);
TMAE_API int tmae_bf16_encoder_layer_fwd(const void* x, const void* x_kv, const tmae_layer_params* P, const tmae_bf16_weights* W,
                                const tmae_layer_tables* T, const float* pos_lut, float tau_min, float eps, int64_t m_q, int64_t m_kv, int32_t c,
                                int32_t ff, int32_t heads, int32_t need_backward, void* y, void* saved, size_t saved_size, void* stream);
TMAE_API int tmae_bf16_encoder_layer_bwd(const void* dy, const void* x, const void* x_kv, const tmae_layer_params* P, const tmae_bf16_weights* W,
                                const tmae_layer_tables* T, const float* pos_lut, float tau_min, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff,
                                int32_t heads, const void* saved, size_t saved_size, void* dx, void* dx_kv, const tmae_layer_params* G,
                                void* g_base, size_t g_bytes, void* scratch, size_t scratch_size, void* stream);
/* g_base / g_bytes (nullable): when the 13 gradient buffers of G are carved from ONE allocation, its extent -- the call then clears it with a
 * single memset; NULL: every kernel clears its own target.  With g_base the weight-gradient GEMMs run on a per-device auxiliary stream owned
 * by the library, forked from and joined to `stream` by events inside the call (tmae_set_option "wgrad_stream"). */
/* which attention core the bf16 layers use: 1 = tcgen05 window kernel (attention_tc.cu), 0 = cast bridge to the fp32-I/O mma.sync
 * kernels (kept as the checker of the former; measurement switch, process-wide) */
TMAE_API int tmae_bf16_set_attention_impl(int32_t impl);
TMAE_API int tmae_bf16_attention_tc_available(void); /* 1 iff the tcgen05 window-attention kernel is built into this library */

#ifdef __cplusplus
}
#endif
#endif /* TMAE_SM100_H_ */
