/*
 * gasfm_b200 -- C ABI of the B200-native GASFM graph-attention path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The Python
 * host layer (gasfm_b200/_lib.py) binds exactly these symbols with ctypes; any other host
 * (C++, the reference's own Python through ctypes/cffi) can bind them the same way, see
 * INTEGRATION.md.  Each entry point names the reference interface it replaces; paths are
 * relative to the reference's ``code/`` directory.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in ``_host``;
 *   - ``stream`` is a ``cudaStream_t`` passed as ``void*`` (NULL = legacy default stream);
 *   - matrices are row-major fp32; ``ld*`` arguments are row strides in ELEMENTS;
 *   - index arrays produced here are int32; the reference-facing ``indices[2,E]`` stays int64;
 *   - every function returns 0 on success or a non-zero code; ``gasfm_last_error()`` then
 *     returns a human-readable message (thread-local).  Nothing here falls back to the CPU.
 *
 * Segment plans
 *   A "plan" describes one aggregation graph (AxialAggregationGraphWrapper,
 *   utils/dataset_utils.py:464-597) as segments over the E observation rows:
 *     seg_ptr[T+1]   exclusive prefix of segment lengths (CSR over views, CSC over tracks)
 *     perm[E]|NULL   edge id of the k-th entry in segment order (NULL = storage order,
 *                    i.e. the row-major CSR the reference's indices already have)
 *   Long segments are cut into chunks of ``chunk`` edges so that the work fills 148 SMs:
 *     chunk_ptr[T+1] exclusive prefix of ceil(len/chunk)
 *     chunk_seg[max_chunks] segment id of every chunk; max_chunks >= E/chunk + T
 *   chunk == 0 selects the short-segment schedule (one lane group per segment, no chunks).
 */
#ifndef GASFM_B200_H
#define GASFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GASFM_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define GASFM_API __attribute__((visibility("default")))
#else
#define GASFM_API
#endif

GASFM_API int gasfm_abi_version(void);
GASFM_API const char* gasfm_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Observation index build  (bit-exact integer work)
 * ------------------------------------------------------------------------------------- */

/* Replaces get_M_valid_points + the counting half of M2sparse
 * (utils/dataset_utils.py:86-113, 116-127).
 * M[2m,n] dense measurements -> valid[m*n] (0/1), cam_per_pts[n], pts_per_cam[m] (int64, the
 * reference's dtype) and n_obs[1] (int64) = number of valid observations E. */
GASFM_API int gasfm_m2sparse_count(const float* M, int m, int n, int min_views_per_point,
                         uint8_t* valid, int64_t* cam_per_pts, int64_t* pts_per_cam,
                         int64_t* n_obs, void* stream);

/* Replaces the index / value half of M2sparse + geo_utils.normalize_M
 * (utils/dataset_utils.py:128-156, utils/geo_utils.py:689-703).
 * Writes indices[2,E] (int64, row-major order == np.nonzero order) and values[E,2] =
 * (Ns[i] @ [x;y;1])[:2] (Ns may be NULL: raw coordinates).  ``scan_ws`` needs
 * gasfm_m2sparse_ws_bytes(m,n) bytes. */
GASFM_API size_t gasfm_m2sparse_ws_bytes(int m, int n);
GASFM_API int gasfm_m2sparse_fill(const float* M, const float* Ns, const uint8_t* valid, int m, int n,
                        int64_t n_obs, int64_t* indices, float* values, void* scan_ws,
                        void* stream);

/* CSR / CSC of the observation graph from the reference's indices[2,E] (int64, row-major
 * sorted).  Replaces the per-call ``sparse_coo_tensor(...).coalesce()`` sorts
 * (utils/sparse_utils.py:436-449; models/layers.py:545-549,561-565,823,922) and the edge lists
 * of AxialAggregationGraphWrapper.create_sparse_axial_aggregation_edges
 * (utils/dataset_utils.py:511-537): built once per scene, never re-sorted.
 *   row_idx[E], col_idx[E]  int32 copies of indices
 *   row_ptr[m+1]            CSR over views (storage order is already CSR order)
 *   col_ptr[n+1], csc_perm[E]  CSC over tracks; csc_perm is the STABLE sort by column
 * ``status`` (device int32) is set to 1 if indices are out of range, 2 if they are not sorted row-major.
 * ``ws`` needs gasfm_csr_build_ws_bytes(n_obs, n) bytes of device scratch (no allocation inside). */
GASFM_API size_t gasfm_csr_build_ws_bytes(int64_t n_obs, int n);
GASFM_API int gasfm_csr_build(const int64_t* indices, int64_t n_obs, int m, int n,
                    int32_t* row_idx, int32_t* col_idx, int32_t* row_ptr,
                    int32_t* col_ptr, int32_t* csc_perm, int32_t* status, void* ws, void* stream);

/* Chunk tables of a plan (see header comment).  chunk_ptr[T+1], chunk_seg[max_chunks]. */
GASFM_API int gasfm_plan_chunks(const int32_t* seg_ptr, int n_seg, int chunk, int32_t* chunk_ptr,
                      int32_t* chunk_seg, int max_chunks, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused GATv2 edge attention  (replaces torch_geometric.nn.GATv2Conv's message passing as
 * called at models/layers.py:329-335, 426-432, 550-556, 566-572, minus lin_l / lin_r)
 *
 *   z = XL[e] + XR[t(e)];  s[e,h] = sum_c att[h,c] * leaky_relu(z, slope)
 *   alpha = softmax over the segment of s;  out[t] = concat_h sum_e alpha * XL[e] + bias
 *   empty segment -> out[t] = bias, seg_max = -inf, seg_sum = 0
 *
 * XL[E,H*C] (row stride ldxl), XR[T,H*C] (row stride ldxr; ldxr == 0 broadcasts one row,
 * used for the stateless first block where the query is lin_r.bias), att[H*C], bias[H*C]|NULL.
 * out[T,H*C], seg_max[T,H], seg_sum[T,H] (softmax statistics, needed by backward and by the
 * multi-GPU merge).  ``ws`` needs gasfm_gat_ws_bytes(max_chunks, H, C) bytes when chunk > 0.
 * If ``normalize`` == 0 the un-normalised partial sum (sum_e exp(s-max) * XL[e]) is written
 * to ``out`` without bias: that is the per-GPU partial for a track-sharded graph.
 * ------------------------------------------------------------------------------------- */
GASFM_API size_t gasfm_gat_ws_bytes(int max_chunks, int heads, int head_dim);

GASFM_API int gasfm_gat_edge_fwd(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                       const float* att, const float* bias,
                       const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                       int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                       int heads, int head_dim, float slope, int normalize,
                       float* out, float* seg_max, float* seg_sum, void* ws, void* stream);

/* Backward.  Given dOut[T,H*C] and the forward's out / statistics:
 *   dXL[E,H*C] (row stride lddxl), dXR[T,H*C], datt[H*C] (accumulated over all edges).
 * ``out_nobias`` = out - bias (the aggregated values), row stride H*C.
 * ``datt_ws`` needs gasfm_gat_bwd_ws_bytes(...) bytes. */
GASFM_API size_t gasfm_gat_bwd_ws_bytes(int64_t n_obs, int n_seg, int max_chunks, int heads, int head_dim);

GASFM_API int gasfm_gat_edge_bwd(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                       const float* att, const float* out_nobias, const float* seg_max,
                       const float* seg_sum, const float* dOut,
                       const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                       int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                       int heads, int head_dim, float slope,
                       float* dXL, int64_t lddxl, float* dXR, float* datt, void* ws, void* stream);

/* gasfm_gat_edge_bwd that also writes dxl_rowmax[E] = max |dXL[e, :]| of every edge row it produces (head shapes 4 x 32 and
 * 4 x 64, gasfm_gat_edge_bwd_rowmax_supported): the row scale of the fp16 input-gradient GEMM that consumes dXL next. */
GASFM_API int gasfm_gat_edge_bwd_rowmax_supported(int heads, int head_dim);
GASFM_API int gasfm_gat_edge_bwd_rowmax(const float* XL, int64_t ldxl, const float* XR, int64_t ldxr,
                       const float* att, const float* out_nobias, const float* seg_max,
                       const float* seg_sum, const float* dOut,
                       const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                       int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                       int heads, int head_dim, float slope,
                       float* dXL, int64_t lddxl, float* dXR, float* datt, float* dxl_rowmax, void* ws, void* stream);

/* The same two kernels with the projected sources STORED as bf16 (XL / dXL: 2-byte elements, row strides in elements,
 * 8-byte aligned; everything else fp32, arithmetic fp32): half the bytes per edge of the bandwidth-bound pass
 * (BASELINE.json configs[4], "fp32 vs bf16").  Head shapes 4 x 32 and 4 x 64.  The result is exact for the bf16-rounded
 * inputs; against fp32 inputs the storage rounding (2^-9 relative per element) is the error. */
GASFM_API int gasfm_gat_edge_fwd_bf16(const void* XL_bf16, int64_t ldxl, const float* XR, int64_t ldxr,
                       const float* att, const float* bias,
                       const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                       int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                       int heads, int head_dim, float slope, int normalize,
                       float* out, float* seg_max, float* seg_sum, void* ws, void* stream);
GASFM_API int gasfm_gat_edge_bwd_bf16(const void* XL_bf16, int64_t ldxl, const float* XR, int64_t ldxr,
                       const float* att, const float* out_nobias, const float* seg_max,
                       const float* seg_sum, const float* dOut,
                       const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                       int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                       int heads, int head_dim, float slope,
                       void* dXL_bf16, int64_t lddxl, float* dXR, float* datt, void* ws, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row / column pooling  (SparseMat.sum / .mean, utils/sparse_utils.py:406-419; sparse_mean,
 * utils/sparse_utils.py:91-131; and the segment sums of the backward passes)
 *   out[t] = scale * sum_{e in segment t} X[e]   (mean_mode: divide by the segment length;
 *   empty segments give 0)
 * ------------------------------------------------------------------------------------- */
GASFM_API int gasfm_seg_sum(const float* X, int64_t ldx, int width,
                  const int32_t* seg_ptr, const int32_t* perm, int n_seg,
                  int chunk, const int32_t* chunk_ptr, const int32_t* chunk_seg, int max_chunks,
                  float scale, int mean_mode, float* out, void* ws, void* stream);
GASFM_API size_t gasfm_seg_sum_ws_bytes(int max_chunks, int width);

/* Backward of pooling: dX[e] = scale * dOut[seg(e)] (/ len if mean_mode). seg_of_edge[E]. */
GASFM_API int gasfm_seg_bcast(const float* dOut, int width, const int32_t* seg_of_edge,
                    const int32_t* seg_ptr, int64_t n_obs, float scale, int mean_mode,
                    float* dX, void* stream);

/* ---------------------------------------------------------------------------------------
 * Per-observation feature ops
 * ------------------------------------------------------------------------------------- */

/* out[width] = column sums of the row-major x[rows, ld>=width] (bias gradients: the reference gets them from
 * autograd's sum over the target rows, e.g. GATv2Conv.bias at models/layers.py:329-335).  Deterministic
 * two-stage reduction; ws: gasfm_col_sum_ws_bytes(rows, width) bytes (may be 0 -> NULL allowed). */
GASFM_API size_t gasfm_col_sum_ws_bytes(int64_t rows, int width);
GASFM_API int gasfm_col_sum(const float* x, int64_t ld, int64_t rows, int width, float* out, void* ws, void* stream);

/* y = relu(layer_norm(x) * gamma + beta)  (normalize_projection_features +
 * relu_on_projection_features, models/layers.py:232-234, 972-984).  gamma == NULL skips the
 * normalisation (use_norm_proj_update = false): y = relu(x).  mean/rstd[E] saved for backward. */
GASFM_API int gasfm_ln_relu_fwd(const float* x, int64_t n_rows, int width, const float* gamma,
                      const float* beta, float eps, float* y, float* mean, float* rstd,
                      void* stream);
GASFM_API size_t gasfm_ln_relu_bwd_ws_bytes(int64_t n_rows, int width);
/* Backward.  The ReLU mask is recomputed from x, mean, rstd, gamma, beta (y is not read).  add: optional
 * [n_rows,width] tensor added to dx -- the gradient of a residual branch that reads the same x. */
GASFM_API int gasfm_ln_relu_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                      const float* gamma, const float* beta, const float* add,
                      int64_t n_rows, int width, float* dx, float* dgamma, float* dbeta, void* ws, void* stream);

/* out[e] = pscale * P[e] + scale * (sum_k x0[e,k] * W0[:,k] + S[col[e]] + V[row[e]] + g) + skip[e]
 * (GraphAttnSfMProjectionFeatureUpdate.forward + the residual of GraphAttnSfMLayer.forward,
 * models/layers.py:941-945, 254-261; also SetOfSetProjectionFeatureUpdate, :141-143).
 * x0 / W0 (width x d0, row-major, d0 <= 4), g and skip may be NULL; P may be NULL when d0 > 0 -- with
 * g = bias and scale = 1 that is the write-bound linear layer of the first block, whose input is only
 * d0 = 2 wide (lin_l / lin_proj / skip_projection of block 0).  ``pscale`` lets the caller fold
 * the 1/4 into lin_proj's weights so that dP == dOut in backward (no extra pass over [E,width]). */
GASFM_API int gasfm_edge_update_fwd(const float* P, int64_t ldp, const float* x0, int d0, const float* W0,
                          const float* S, const float* V, const float* g, const float* skip,
                          int64_t ldskip, const int32_t* row_idx, const int32_t* col_idx,
                          int64_t n_obs, int width, float pscale, float scale, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense per-observation projection on the tensor cores (tcgen05, 3xTF32 split = fp32-level accuracy)
 *   C[M,N] = A[M,K] * B[N,K]^T + bias      fp32 in, fp32 out, N <= 256 (multiple of 16), K multiple of 4
 * Replaces cuBLAS SGEMM for GATv2Conv.lin_l (models/layers.py:329,426 via PyG) and lin_proj
 * (models/layers.py:941) and their input gradients.  B_hi / B_lo are the tf32 split of the weight
 * matrix, produced by gasfm_split_tf32 (hi = tf32(w), lo = w - hi).
 * ------------------------------------------------------------------------------------- */
GASFM_API int gasfm_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream);
GASFM_API int gasfm_linear_tf32x3_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc);
GASFM_API int gasfm_linear_tf32x3(const float* A, int64_t lda, const float* B_hi, const float* B_lo,
                                  const float* bias, float* C, int64_t ldc, int64_t M, int N, int K,
                                  int accumulate /* 1: C += A B^T + bias */, void* stream);

/* The same with A given as the column-wise concatenation [A_0 | .. | A_{n_seg-1}] of n_seg <= 4 matrices of seg_k
 * columns each (host arrays of device pointers / row strides); B is [N, n_seg * seg_k].  Input gradient of
 * several projections of one input in ONE pass: dX = sum_i dY_i W_i (GraphAttnSfMLayer: lin_l x 2 + lin_proj,
 * models/layers.py:329,426,941) instead of one GEMM plus two read-modify-write accumulations. */
GASFM_API int gasfm_linear_tf32x3_cat(const float* const* A, const int64_t* lda, int n_seg, int seg_k,
                            const float* B_hi, const float* B_lo, const float* bias, float* C, int64_t ldc,
                            int64_t M, int N, int accumulate,
                            float* a_amax /* optional [n_seg]: receives max |A_i| (for gasfm_wgrad_f16x2) */,
                            void* stream);

/* Same product on the fp16 tensor-core path (twice the tf32 MMA rate) with a SCALED 2 x FP16 split:
 * every row of A and of B is multiplied by a power of two that puts its largest magnitude in [2^14, 2^15)
 * (exact), split into fp16 hi + lo, multiplied as A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation and
 * descaled in the epilogue.  Accuracy ~2^-22 relative to |A_row| |B_row| like the 3xTF32 form; the kernel is
 * HBM-bound (reads A once, writes C once) instead of tensor-bound.  K <= 256, K % 8 == 0, N as above.
 * gasfm_split_f16 prepares B: hi/lo are [n_rows, k] fp16 (2-byte) matrices, descale[n_rows] fp32. */
GASFM_API int gasfm_split_f16(const float* w, int n_rows, int k, void* hi, void* lo, float* descale, void* stream);
/* Profiling hook: dev_buffer = 3*16*16 int64 on the device; the following gasfm_linear_f16x2 launches record
 * SM-clock timestamps of CTA 0's producer / MMA / epilogue milestones there.  NULL switches it off. */
GASFM_API int gasfm_debug_set_gemm_trace(void* dev_buffer);
GASFM_API int gasfm_linear_f16x2_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc);
/* groups (1..3): B_hi/B_lo/b_descale/bias hold ``groups`` stacked [N, K] weights; group g is written to columns
 * [g*N, (g+1)*N) of C (ldc >= groups*N).  Several projections of the same input read A from HBM once. */
GASFM_API int gasfm_linear_f16x2(const float* A, int64_t lda, const void* B_hi, const void* B_lo,
                       const float* b_descale, const float* bias, float* C, int64_t ldc,
                       int64_t M, int N, int K, int groups, int accumulate,
                       float* a_amax /* optional [1]: receives max |A| (for gasfm_wgrad_f16x2) */, void* stream);

/* The fp16 path for a column-wise concatenated A = [A_0 | .. | A_{n_seg-1}] (separate buffers, seg_k in {128, 256} columns each,
 * B = [N, n_seg * seg_k] split by gasfm_split_f16): the input gradient of a block's projections, dX = sum_i dY_i W_i
 * (models/layers.py:329,426,941), at twice the tensor rate of gasfm_linear_tf32x3_cat.  One power-of-two scale per ROW across all
 * segments: the operand producer passes over the tile twice (row maxima, then scale + split), the second pass and -- through an
 * L2 prefetch issued one tile ahead -- most of the first are served by L2, so HBM still sees every A_i once.
 * a_amax[n_seg] (optional) receives max |A_i|. */
GASFM_API int gasfm_linear_f16x2_cat_supported(int64_t M, int N, int n_seg, int seg_k, int64_t ldc);
GASFM_API int gasfm_linear_f16x2_cat(const float* const* A, const int64_t* lda, int n_seg, int seg_k,
                           const void* B_hi, const void* B_lo, const float* b_descale, const float* bias,
                           float* C, int64_t ldc, int64_t M, int N, float* a_amax, void* stream);

/* ... and in ONE pass when the kernels that produced the A_i left their row maxima behind (rowmax[i][M] = max |A_i[m, :]|; host
 * array of device pointers): gasfm_gat_edge_bwd_rowmax for the two attention gradients, gasfm_x0_bwd_rowmax for the gradient that
 * reaches lin_proj.  The operand producer then streams the K blocks with three blocks of loads in flight; n_seg * seg_k / 64 must
 * be a multiple of 4. */
GASFM_API int gasfm_linear_f16x2_cat_rowmax(const float* const* A, const int64_t* lda, const float* const* rowmax, int n_seg,
                           int seg_k, const void* B_hi, const void* B_lo, const float* b_descale, const float* bias,
                           float* C, int64_t ldc, int64_t M, int N, float* a_amax, void* stream);

/* The same kernel with LayerNorm + ReLU applied to A on the fly: the operand is relu(layer_norm(A) * gamma + beta)
 * (normalize_projection_features + relu_on_projection_features feeding lin_l / lin_proj, models/layers.py:232-234 ->
 * :329,426,941), evaluated on the register-resident A tile, so the normalised features are never written to or read
 * from memory.  ln_mean / ln_rstd [M] receive the row statistics (needed by gasfm_ln_relu_bwd); a_amax is max |operand|. */
GASFM_API int gasfm_linear_f16x2_ln(const float* A, int64_t lda, const float* ln_gamma, const float* ln_beta, float ln_eps,
                          float* ln_mean, float* ln_rstd, const void* B_hi, const void* B_lo,
                          const float* b_descale, const float* bias, float* C, int64_t ldc,
                          int64_t M, int N, int K, int groups, float* a_amax, void* stream);
/* ... which also writes the normalised operand y[M, K] (row stride ldy) from the registers that hold it: the weight gradients of
 * the same projections read it (dW = dY^T y), and a forward that keeps activations gets it without a separate LayerNorm pass.
 * Only the N = K = 256 kernel (CTA pairs, cta_group::2) does this: ask gasfm_linear_f16x2_ln_y_supported first. */
GASFM_API int gasfm_linear_f16x2_ln_y_supported(int64_t M, int N, int K, int64_t lda, int64_t ldc);
GASFM_API int gasfm_linear_f16x2_ln_y(const float* A, int64_t lda, const float* ln_gamma, const float* ln_beta, float ln_eps,
                          float* ln_mean, float* ln_rstd, float* y, int64_t ldy, const void* B_hi, const void* B_lo,
                          const float* b_descale, const float* bias, float* C, int64_t ldc,
                          int64_t M, int N, int K, int groups, float* a_amax, void* stream);

/* Weight gradient on the fp16 path: dW = dY^T X (+ db), each operand scaled by ONE power of two derived from its largest
 * magnitude (amax_dy[1], amax_x[1]: DEVICE scalars, as left behind by the GEMMs above that read the same matrices).
 * Nout in {128, 256}, Kout in {64, 128, 192, 256}; ws: gasfm_wgrad_f16x2_ws_bytes.  Error ~1e-6 of max|dW|. */
GASFM_API int gasfm_wgrad_f16x2_supported(int64_t E, int Nout, int Kout, int64_t lddy, int64_t ldx);
GASFM_API size_t gasfm_wgrad_f16x2_ws_bytes(int Nout, int Kout);
GASFM_API int gasfm_wgrad_f16x2(const float* dY, int64_t lddy, const float* X, int64_t ldx,
                      const float* amax_dy, const float* amax_x, int64_t E, int Nout, int Kout,
                      float* dW, float* dbias, void* ws, void* stream);

/* n_groups <= 3 weight gradients that share X in ONE launch: dW[g] = dY[g]^T X, db[g] = column sums of dY[g]
 * (lin_l x2 + lin_proj of one block all multiply the same x, models/layers.py:329,426,941).  dY / lddy: host arrays of device
 * pointers / row strides; amax_dy[n_groups]: device array; dW: stacked [n_groups * Nout, Kout], dbias: [n_groups * Nout].
 * The CTAs working on the same rows run side by side, so X is fetched from HBM once and from L2 by the other groups. */
GASFM_API int gasfm_wgrad_f16x2_multi(const float* const* dY, const int64_t* lddy, int n_groups, const float* X, int64_t ldx,
                            const float* amax_dy, const float* amax_x, int64_t E, int Nout, int Kout,
                            float* dW, float* dbias, void* ws, void* stream);

/* Weight gradient of the same projections: dW[Nout,Kout] = dY[E,Nout]^T * X[E,Kout], 3xTF32 on tcgen05,
 * deterministic split-K over the SMs.  ``ws`` needs gasfm_wgrad_tf32x3_ws_bytes(Nout,Kout) bytes. */
GASFM_API int gasfm_wgrad_tf32x3_supported(int64_t E, int Nout, int Kout, int64_t lddy, int64_t ldx);
GASFM_API size_t gasfm_wgrad_tf32x3_ws_bytes(int Nout, int Kout);
GASFM_API int gasfm_wgrad_tf32x3(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t E,
                                 int Nout, int Kout, float* dW, float* dbias /* [Nout] column sums of dY, or NULL */,
                                 void* ws, void* stream);

/* The same weight gradient for the narrow shipped widths (Nout, Kout in {32, 64}) on the SIMT pipes, where a
 * 128-wide tensor-core tile would be mostly padding; fp32 round-to-nearest accumulation, deterministic. */
GASFM_API int gasfm_wgrad_small_supported(int Nout, int Kout, int64_t lddy, int64_t ldx);
GASFM_API size_t gasfm_wgrad_small_ws_bytes(int Nout, int Kout);
GASFM_API int gasfm_wgrad_small(const float* dY, int64_t lddy, const float* X, int64_t ldx, int64_t E, int Nout,
                                int Kout, float* dW, float* dbias /* or NULL */, void* ws, void* stream);

/* Backward of the rank-d0 term of gasfm_edge_update_fwd in one pass over dOut[E,width]:
 *   dx0[E,d0] = scale * dOut W0,   dW0[width,d0] = scale * dOut^T x0        (d0 <= 4)
 * (the cat(x, x0) half of lin_proj, models/layers.py:245-251, 941). */
GASFM_API size_t gasfm_x0_bwd_ws_bytes(int64_t E, int width);
GASFM_API int gasfm_x0_bwd(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0,
                           float scale, float* dx0, float* dW0, void* ws, void* stream);

/* gasfm_x0_bwd that also writes rowmax[E] = max |dOut[e, :]| (the same gradient is lin_proj's output gradient,
 * models/layers.py:941: its row scale for the fp16 input-gradient GEMM). */
GASFM_API int gasfm_x0_bwd_rowmax(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0,
                                  float scale, float* dx0, float* dW0, void* ws, float* rowmax, void* stream);

/* The whole backward of gasfm_edge_update_fwd that runs in storage order, in ONE pass over dOut[E,width] (the reference's autograd
 * makes one scatter-add per gathered term, models/layers.py:941-945): dV[n_seg,width] = scale * sum over the view's rows of dOut
 * (gradient of V[row]; seg_ptr / chunk tables of the view plan, chunk > 0), dx0 / dW0 as gasfm_x0_bwd, and rowmax[E] (optional).
 * ws: gasfm_update_bwd_views_ws_bytes(max_chunks, width). */
GASFM_API size_t gasfm_update_bwd_views_ws_bytes(int max_chunks, int width);
GASFM_API int gasfm_update_bwd_views(const float* dOut, int64_t E, int width, const float* x0, const float* W0, int d0,
                                     float scale, const int32_t* seg_ptr, int n_seg, int chunk, const int32_t* chunk_ptr,
                                     const int32_t* chunk_seg, int max_chunks, float* dV, float* dx0, float* dW0,
                                     float* rowmax, void* ws, void* stream);

/* ---------------------------------------------------------------------------------------
 * Sparse ESFM reprojection loss over the E observed (view, point) pairs (ESFMLoss.forward,
 * loss_functions.py:85-123, without the dense [m,3,n] tensors).  Ps[m,3,4], pts3D[4,n] (row-major), obs[E,2]
 * = SparseMat.values (normalised measurements).  out[0] = loss, out[1] = #observations with a valid depth.
 * Backward: G[E,16] = per-observation contributions (cols 0-11 -> dPs[row], 12-15 -> dpts3D[:,col]) after the
 * reference's gradient hook (grad_mode 0 none, 1 = normalise where depth valid / #valid, 2 = normalise all / E);
 * ``upstream`` and ``stats`` (= out of the forward) are DEVICE scalars, so nothing synchronises.
 * ------------------------------------------------------------------------------------- */
GASFM_API size_t gasfm_esfm_loss_ws_bytes(int64_t E);
GASFM_API int gasfm_esfm_loss_fwd(const float* Ps, const float* pts3D, int64_t n, const float* obs,
                                  const int32_t* row_idx, const int32_t* col_idx, int64_t E, float margin,
                                  int hinge, float hinge_weight, float* out, void* ws, void* stream);
GASFM_API int gasfm_esfm_loss_bwd(const float* Ps, const float* pts3D, int64_t n, const float* obs,
                                  const int32_t* row_idx, const int32_t* col_idx, int64_t E, float margin,
                                  int hinge, float hinge_weight, const float* upstream, const float* stats,
                                  int grad_mode, float* G, void* stream);

/* Per-step reprojection metric on the device (evaluation.compute_core_errors -> core_errors['our_repro'],
 * evaluation.py:8-32; geo_utils.reprojection_error_with_points, utils/geo_utils.py:371-391): the nan-mean over the E
 * observed pairs of || u_e - pflat(P_i X_j)_xy ||.  Ps[m,3,4] are the UNNORMALISED cameras (Ns^-1 Ps_norm), pts3D[4,n],
 * obs[E,2] the raw image points.  out[0] = mean error, out[1] = number of pairs counted; ws as for the loss.  Nothing
 * leaves the device (the reference moves predictions to numpy and builds dense [m,n] arrays every training step). */
GASFM_API int gasfm_reproj_error(const float* Ps, const float* pts3D, int64_t n, const float* obs,
                                 const int32_t* row_idx, const int32_t* col_idx, int64_t E, float* out, void* ws,
                                 void* stream);

/* ---------------------------------------------------------------------------------------
 * Track-sharded multi-GPU exchange over NVLink peer memory (the reference is single-GPU,
 * main.py:78, so there is no reference interface to cite -- these serve GATv2Conv's per-view softmax,
 * models/layers.py:329-335, when the observations of a view are spread over several GPUs).
 *
 * Setup, once per process: every rank allocates an exchange buffer (gasfm_peer_buffer_bytes) and a flag array
 * (gasfm_peer_flags_bytes) with gasfm_peer_alloc, exports both as 64-byte IPC handles, exchanges the handles through
 * any host channel (torch.distributed here), imports its peers' handles and hands the world pointers (its own ones at
 * index ``rank``) to gasfm_peer_comm_create.  A communicator may also be built from buffers of ONE process (several
 * "ranks" on streams of one device): the kernels only see pointers.
 *
 * Every rank must issue the same sequence of exchanges with the same sizes.  An exchange is ONE kernel: push the local
 * partial into every peer's buffer, release a flag per CTA, wait for the peers' matching CTAs, combine in rank order
 * (bit-identical results on every rank).  No host synchronisation, no NCCL: the launches can be captured in a CUDA graph.
 * A peer that does not show up within ``timeout_s`` sets the communicator's error flag (gasfm_peer_comm_error) instead
 * of hanging the GPU.
 * ------------------------------------------------------------------------------------- */
#define GASFM_PEER_HANDLE_BYTES 64
GASFM_API int gasfm_peer_alloc(size_t bytes, void** ptr);                 /* cudaMalloc + zero fill */
GASFM_API int gasfm_peer_free(void* ptr);
GASFM_API int gasfm_peer_export(void* ptr, void* handle64);               /* ptr from gasfm_peer_alloc */
GASFM_API int gasfm_peer_import(const void* handle64, void** ptr);
GASFM_API int gasfm_peer_close(void* ptr);                                /* ptr from gasfm_peer_import */
GASFM_API size_t gasfm_peer_buffer_bytes(int world, int64_t region_floats);
GASFM_API size_t gasfm_peer_flags_bytes(int world);
GASFM_API int gasfm_peer_comm_create(int rank, int world, void* const* bufs, void* const* flags, int64_t region_floats,
                                     double timeout_s, void** comm_out);
GASFM_API int gasfm_peer_comm_destroy(void* comm);
GASFM_API int gasfm_peer_comm_error(void* comm, int* error_out);          /* synchronising read of the error flag */

/* out[n] = scale * sum over ranks of in[n] (n % 4 == 0, n <= region_floats; n == 0: a pure barrier).  Gradient
 * exchanges of the sharded backward: dXR of the per-view queries, the per-view term of the observation update
 * (models/layers.py:941-945), and the flat bucket of the observation-/point-level parameter gradients. */
GASFM_API int gasfm_peer_allreduce_sum(void* comm, const float* in, float* out, int64_t n, float scale, void* stream);

/* Log-sum-exp merge of the per-rank un-normalised softmax partials that gasfm_gat_edge_fwd(normalize = 0) produces:
 *   M = max_r max_r;  L = sum_r e^(max_r - M) sum_r;  out = sum_r e^(max_r - M) acc_r / L (+ bias);  L == 0 -> out = bias.
 * acc[T,H*C], seg_max / seg_sum[T,H] are this rank's partial; out[T,H*C], M[T,H], L[T,H] are identical on all ranks. */
GASFM_API int gasfm_peer_lse_merge(void* comm, const float* acc, const float* seg_max, const float* seg_sum,
                                   const float* bias, int n_seg, int heads, int head_dim,
                                   float* out, float* M, float* L, void* stream);

/* The same two combines on a buffer that already holds every rank's contribution (gathered[world][region_floats],
 * region = acc | max | sum resp. the n summands), e.g. after an NCCL / gloo all_gather: the library-collective arm. */
GASFM_API int gasfm_lse_merge_gathered(const float* gathered, int world, int64_t region_floats, const float* bias,
                                       int n_seg, int heads, int head_dim, float* out, float* M, float* L, void* stream);
GASFM_API int gasfm_sum_gathered(const float* gathered, int world, int64_t region_floats, int64_t n, float scale,
                                 float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Host-buffer convenience entry points (inputs and outputs in HOST memory; allocation and the
 * host<->device copies happen inside the call).  These are what a non-torch host binds.
 * ------------------------------------------------------------------------------------- */
GASFM_API int gasfm_csr_build_host(const int64_t* indices_host, int64_t n_obs, int m, int n,
                         int32_t* row_ptr_host, int32_t* col_ptr_host, int32_t* csc_perm_host);

GASFM_API int gasfm_gat_edge_fwd_host(const float* XL_host, const float* XR_host, const float* att_host,
                            const float* bias_host, const int64_t* target_host,
                            int64_t n_obs, int n_seg, int heads, int head_dim, float slope,
                            float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* GASFM_B200_H */
