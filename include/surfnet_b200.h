/* surfnet_b200.h -- C ABI of libsurfnet_b200.so
 *
 * B200 (sm_100a) operator-application kernels for SurfaceNetworks' LapResNet2 / DirResNet2 blocks.
 * This is the drop-in boundary that replaces the reference's only native seam:
 *
 *   reference                                              replaced by
 *   -----------------------------------------------------  ------------------------------------------
 *   src/utils/cuda/batch_csr.cu:13-47  (+ batch_csr.py:28-59)   sn_coo_to_csr32, sn_csr32_to_bsr4_*
 *   src/utils/cuda/sparse_bmm.cu:16-61 (+ sparse_bmm.py:28-61)  sn_csr_spmm_f32, sn_bsr4_spmm_f32
 *   torch.mm(sparse_coo, dense) at src/utils/utils_pt.py:167,176   sn_csr_spmm_f32   (Laplacian)
 *   torch.mm(sparse_coo, dense) at src/utils/utils_pt.py:202,214   sn_bsr4_spmm_f32  (Dirac / adjoint)
 *   src/utils/cuda/sparse_bmm_func.py:53-72 (backward = A^T grad)  same SpMM entry points on a
 *                                                                  transposed structure built once
 *   GraphConv1x1 "pre" BN + Linear, utils_pt.py:91-104             sn_colstats_f32, sn_bn_fold_{fwd,bwd}_f32,
 *                                                                  sn_gemm_tf32_f32, sn_gemm_tn_tf32_f32
 *   autograd backward of elu + torch.mm + cat, utils_pt.py:161-216 sn_{csr,bsr4}_spmm_epilogue_f32
 *   AvgResNet2 / global_average, utils_pt.py:120-122,230-243       sn_segment_sum_f32, sn_elu_bwd_group_f32
 *   sparse_diag_cat + upload per step, utils_pt.py:41-53           sn_assemble_block_diag
 *   mesh.dirac / cotangent_weights / graph.laplacian (offline)     sn_mesh_dirac_bsr4, sn_mesh_laplacian_csr
 *
 * Conventions (same as the reference's cupy launch seam, sparse_bmm.py:49-59, made explicit):
 *   - every pointer is a DEVICE pointer unless its name starts with host_;
 *   - the library never allocates, never synchronises and keeps no mutable global state:
 *     the caller owns every buffer (outputs and workspaces included) and passes the stream;
 *   - work is enqueued on `stream` (a cudaStream_t; pass torch's current stream) and the call returns;
 *   - shapes are run-time arguments (the reference re-JITs a kernel per shape, sparse_bmm.py:29-47);
 *   - return value: SN_OK, a negative SN_ERR_* argument error, or a positive cudaError_t.
 *   - dense matrices are row-major fp32 with an explicit leading dimension (in floats).
 */
#ifndef SURFNET_B200_H
#define SURFNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sn_stream_t; /* cudaStream_t */

#define SN_VERSION 100 /* 0.1.0 */

#define SN_OK 0
#define SN_ERR_ARG (-1)          /* null pointer / negative size / inconsistent arguments        */
#define SN_ERR_UNSUPPORTED (-2)  /* shape or alignment not supported by this entry point          */
#define SN_ERR_WORKSPACE (-3)    /* workspace smaller than the matching *_ws_bytes() query        */
#define SN_ERR_OVERFLOW (-4)     /* nnz / rows / cols do not fit the 32-bit index format          */

/* flags for sn_coo_to_csr32 */
#define SN_COO_SORTED 1 /* input is coalesced: sorted by (batch,row,col), as torch .coalesce() returns */

/* flags for the SpMM entry points */
#define SN_SPMM_ELU_INPUT 1     /* apply ELU(alpha=1) to the gathered dense operand: Y = S * elu(X)  */
#define SN_SPMM_DIRECT_GATHER 2 /* force the first-generation direct-gather kernel (one row per lane group)       */
#define SN_SPMM_SMEM_STREAM 4   /* sn_bsr4_spmm_f32: force the cp.async shared-memory streaming kernel (C=128/256/512) */
#define SN_SPMM_VARIANT(v) (((v) & 15) << 8) /* tuning variant of the row-group kernel (benchmarks only; 0 = default;
                                                6 = force the small-operator kernel, 7 = force the persistent one;
                                                sn_*_spmm_epilogue_f32: 8 = operand loads at the row's end;
                                                5 / 9 = two gathers in flight per row group, in registers /
                                                through shared memory) */
#define SN_SPMM_ROW_ENTRIES(n) (((n) & 15) << 12) /* caller's hint: typical (mean, rounded up) entries per row, 0 = unknown.
                                                     Picks the pipeline shape only, never the result: <= 3 (D) -> the
                                                     small-operator kernel keeps three gathers in flight; >= 5 (D*) -> the
                                                     BSR4 row-group kernel keeps two, landing in shared memory */

int sn_version(void);
const char* sn_status_string(int status);

/* ------------------------------------------------------------------------------------------------
 * Format layer.  COO (int64 indices, as torch.sparse / sparse_cat / sparse_diag_cat produce,
 * src/utils/utils_pt.py:21-69) -> CSR32 (int32 row pointers / column indices, fp32 values).
 *
 * batch == NULL : 2-D operator, indices are (row, col)                      [sparse_diag_cat layout]
 * batch != NULL : 3-D operator [B, rows_per_batch, cols_per_batch]          [sparse_cat layout, the
 *                 input of the reference's batch_csr kernel]; it is flattened to the block-diagonal
 *                 2-D operator: row' = b*rows_per_batch + row, col' = b*cols_per_batch + col.
 * n_rows is the total (flattened) row count; rowptr has n_rows+1 entries.  Unlike the reference
 * kernel (batch_csr.cu:36-42, which leaves interior empty rows pointing at 0) empty rows anywhere are
 * handled.  Without SN_COO_SORTED entries may come in any order; duplicates are kept (not summed) in
 * (col, input position) order, so results are deterministic.  Passing (col,row) swapped without
 * SN_COO_SORTED builds the transpose.
 * ---------------------------------------------------------------------------------------------- */
size_t sn_coo_to_csr32_ws_bytes(int64_t nnz, int64_t n_rows);
int sn_coo_to_csr32(const int64_t* batch, const int64_t* row, const int64_t* col, const float* val,
                    int64_t nnz, int64_t rows_per_batch, int64_t cols_per_batch, int64_t n_rows,
                    int64_t n_cols, int flags, int32_t* rowptr, int32_t* colind, float* out_val,
                    void* ws, size_t ws_bytes, sn_stream_t stream);

/* CSR32 -> BSR4 (4x4 blocks; n_rows % 4 == 0).  Two steps because the block count is data dependent:
 *   count : browptr[0..n_rows/4] <- exclusive scan of blocks per block-row; browptr[n_rows/4] = #blocks
 *           (read it back, allocate bcolind[#blocks], bval[16*#blocks]), then
 *   fill  : bcolind ascending per block-row; bval holds each block column-major with every column rotated so
 *           that its diagonal entry comes first:
 *           bval[16*k + 4*q + s] = block_k[(q + s) mod 4][q]   (q = column in block, s = 0..3).
 *           (Lane q of the SpMM kernels then accumulates output component (q + s) mod 4 in slot s, and the
 *           cross-lane reduction needs no per-lane register selection.)
 * The Dirac view of utils_pt.py:201-203 makes q index the q-th quarter of the channel vector. */
size_t sn_csr32_to_bsr4_ws_bytes(int64_t n_rows);
int sn_csr32_to_bsr4_count(const int32_t* rowptr, const int32_t* colind, int64_t n_rows,
                           int32_t* browptr, void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_csr32_to_bsr4_fill(const int32_t* rowptr, const int32_t* colind, const float* val, int64_t n_rows,
                          const int32_t* browptr, int32_t* bcolind, float* bval, sn_stream_t stream);

/* GPU-resident batch assembly: block-diagonal concatenation of per-mesh CSR32 (vals_per_entry = 1) or BSR4
 * (vals_per_entry = 16) operators that already live on the device, every mesh padded to rows_pad x cols_pad (block)
 * rows / columns -- what sparse_diag_cat(...).coalesce() + upload does per step in the reference (utils_pt.py:41-53,
 * as_rigid_as_possible/main.py:172-183).  `parts` is a device table of 6 int64 per mesh: rowptr pointer, colind
 * pointer, value pointer, rows, entries (nnz / blocks), exclusive prefix sum of entries.  Outputs: rowptr_out
 * [n_parts*rows_pad + 1], colind_out / val_out [total_entries (* 16)]. */
int sn_assemble_block_diag(const int64_t* parts, int64_t n_parts, int64_t rows_pad, int64_t cols_pad,
                           int64_t total_entries, int vals_per_entry, int32_t* rowptr_out, int32_t* colind_out,
                           float* val_out, sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Operator construction on the GPU (the step in front of the path; reference src/utils/mesh.py:17-125,
 * src/utils/graph.py:40-49, recipe src/as_rigid_as_possible/add_laplacian.py:39-70 -- O(V^2) dense numpy, offline).
 *
 * A batch of n_meshes triangle meshes, every mesh padded to v_pad vertices and f_pad faces:
 *   V [n_meshes, v_pad, 3] fp64 positions;  F [n_meshes, f_pad, 3] int32 LOCAL vertex indices, a face with a negative
 *   (or >= v_pad) index is padding.  A vertex may belong to any number of faces (the reference is dense O(V^2) numpy and
 *   has no limit either); vertices in more than 64 faces are processed over global scratch instead of registers.
 * Outputs are the block-diagonal batch operators in the formats of the SpMM entry points, values computed in fp64 with
 * the reference's operation order and rounded to fp32 once (the reference's .astype('float32')):
 *   sn_mesh_dirac_bsr4   : D  [n*f_pad x n*v_pad] block rows/cols: d_browptr [n*f_pad + 1], d_bcolind [<= 3 n f_pad],
 *                          d_bval [<= 48 n f_pad];  D* [n*v_pad x n*f_pad]: da_browptr [n*v_pad + 1], da_bcolind /
 *                          da_bval with the same capacities.  Block counts = the last row pointers.  Optionally
 *                          (non-NULL pairs) the transposes used by backward: D^T = (da_browptr, dt_bcolind, dt_bval)
 *                          shares D*'s structure, (D*)^T = (d_browptr, dat_bcolind, dat_bval) shares D's.
 *   sn_mesh_laplacian_csr: L  [n*v_pad x n*v_pad]: rowptr [n*v_pad + 1], colind / val [<= n (v_pad + 6 f_pad)].
 * status: one device int32, informational: 0, or the largest per-vertex face count when some vertex exceeds 64 faces
 * (its rows are complete and exact like every other row).  Deterministic (no floating-point atomics).
 * ws: sn_mesh_ws_bytes(n_meshes, v_pad, f_pad).
 * ---------------------------------------------------------------------------------------------- */
size_t sn_mesh_ws_bytes(int64_t n_meshes, int64_t v_pad, int64_t f_pad);
int sn_mesh_dirac_bsr4(const double* V, const int32_t* F, int64_t n_meshes, int64_t v_pad, int64_t f_pad,
                       int32_t* d_browptr, int32_t* d_bcolind, float* d_bval, int32_t* da_browptr, int32_t* da_bcolind,
                       float* da_bval, int32_t* dt_bcolind, float* dt_bval, int32_t* dat_bcolind, float* dat_bval,
                       int32_t* status, void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_mesh_laplacian_csr(const double* V, const int32_t* F, int64_t n_meshes, int64_t v_pad, int64_t f_pad,
                          int32_t* rowptr, int32_t* colind, float* val, int32_t* status, void* ws, size_t ws_bytes,
                          sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Operator application.
 *
 * sn_csr_spmm_f32 :  Y[n_rows x C] = S * X            (scalar Laplacian, utils_pt.py:167,176)
 * sn_bsr4_spmm_f32:  Y[n_brows x C] = S * X with the quaternion view of utils_pt.py:201-203,213-215:
 *                    Y[r, p*C/4 + c] = sum_{blocks (r,j)} sum_q block[p][q] * X[j, q*C/4 + c],  C % 4 == 0
 * X rows are addressed through colind, so X must have at least n_cols rows.  Accumulation is fp32 FMA
 * in ascending storage order (bit-reproducible run to run).  X and Y must not alias.
 * Both entry points run the row-group kernel (spmm_rowgroup.cu: C/16 lanes own a sparse row and keep the whole 4x4
 * block product in registers, software-pipelined LDG.128 gathers, per-warp index rings in shared memory) for
 * C in {32, 64, 128, 256, 512} with 16-byte aligned operands, and the direct-gather kernels for other widths /
 * alignments.  SN_SPMM_SMEM_STREAM / SN_SPMM_DIRECT_GATHER select the earlier kernels (kept for A/B measurements).
 * ---------------------------------------------------------------------------------------------- */
int sn_csr_spmm_f32(const int32_t* rowptr, const int32_t* colind, const float* val,
                    const float* X, int64_t ldx, float* Y, int64_t ldy,
                    int64_t n_rows, int64_t C, int flags, sn_stream_t stream);
int sn_bsr4_spmm_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval,
                     const float* X, int64_t ldx, float* Y, int64_t ldy,
                     int64_t n_brows, int64_t C, int flags, sn_stream_t stream);

/* Backward-pass SpMM with the activation derivative in its store path:
 *     Y = (S * X + G) .* elu'(A) + G2   G, A, G2 optional (NULL: absent), each [n_rows x C] with its leading dimension
 * A holds ACTIVATED values a = elu(x), so elu' = 1 for a > 0 and a + 1 otherwise (as sn_elu_bwd_f32 with a_is_raw = 0).
 * With S = the transposed operator, X = the gradient of the gathered half and G = the gradient of the un-gathered half
 * this is the whole backward of "elu, then [x | S x]" (src/utils/utils_pt.py:161-168,195-216 through autograd in the
 * reference) in one launch instead of an SpMM plus an elementwise pass; G2 carries a gradient that bypasses the
 * activation (the block residual, utils_pt.py:180,220), so autograd's accumulation adds disappear too.  Y may alias G / G2.  Row-group kernel only:
 * returns SN_ERR_UNSUPPORTED for C not in {32, 64, 128, 256, 512} or unaligned operands (callers then run
 * sn_*_spmm_f32 followed by sn_elu_bwd_f32). */
int sn_csr_spmm_epilogue_f32(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X,
                             int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int64_t C, const float* G, int64_t ldg,
                             const float* A, int64_t lda, const float* G2, int64_t ldg2, int flags, sn_stream_t stream);
int sn_bsr4_spmm_epilogue_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                              int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, const float* G, int64_t ldg,
                              const float* A, int64_t lda, const float* G2, int64_t ldg2, int flags, sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Activation pass in front of each operator application (F.elu at utils_pt.py:161,172,195,208).
 * sn_elu_f32     : Y = elu(X), strided in/out so the result lands in the left half of the stage's
 *                  concat buffer [rows x 2C] (replaces elu + torch.cat, utils_pt.py:168,177,204,216).
 * sn_elu_bwd_f32 : Y = (G + G2) * elu'(x); G2 may be NULL.  a_is_raw = 0: A holds the activated value
 *                  elu(x) (elu' = 1 if A > 0 else A + 1); a_is_raw = 1: A holds x (elu' = 1 if x > 0
 *                  else exp(x)).  Y may alias G or G2.
 * ---------------------------------------------------------------------------------------------- */
int sn_elu_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t rows, int64_t C, sn_stream_t stream);
int sn_elu_bwd_f32(const float* A, int64_t lda, int a_is_raw, const float* G, int64_t ldg, const float* G2,
                   int64_t ldg2, float* Y, int64_t ldy, int64_t rows, int64_t C, sn_stream_t stream);

/* Per-mesh (segment) column sums: out[s, c] = sum over the rows_per_seg rows of segment s of w[r] * X[r, c]
 * (w NULL = 1).  With w = mask this is the numerator of global_average (utils_pt.py:120-122); unweighted it is the
 * per-mesh gradient sum of its backward.  Deterministic (8 row slices per segment, fixed-order reductions).
 * sn_elu_bwd_group_f32: Y = (G + w[r] * GB[r / rows_per_seg]) * elu'(x) + G3 -- the backward of elu followed by
 * global_average's broadcast term, A holding the activated values as in sn_elu_bwd_f32 (a_is_raw = 0); G3 (may be NULL)
 * is a gradient that bypasses the activation: the residual of AvgResNet2 (utils_pt.py:243). */
size_t sn_segment_sum_ws_bytes(int64_t n_seg, int64_t C);
int sn_segment_sum_f32(const float* X, int64_t ldx, const float* w, int64_t rows_per_seg, int64_t n_seg, int64_t C,
                       float* out, void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_elu_bwd_group_f32(const float* A, int64_t lda, const float* G, int64_t ldg, const float* GB, const float* w,
                         int64_t rows_per_seg, const float* G3, int64_t ldg3, float* Y, int64_t ldy, int64_t rows, int64_t C,
                         sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Dense half of a stage on the tensor cores (tcgen05, kind::tf32, accumulators in TMEM):
 *
 *   C[M x N] = A[M x K] * B[N x K]^T + bias[N] + group_bias[row / rows_per_group][N] + rscale[N] .* R[M x N]
 *
 * bias, group_bias, R, rscale may be NULL (rscale NULL with R given means R is added unscaled).  group_bias is the
 * per-mesh term of AvgResNet2 (utils_pt.py:233,239: the broadcast global average times its half of the weights).  A is the stage's concat
 * buffer Z, B the Linear weight of GraphConv1x1 (utils_pt.py:89,99) with the training-mode BatchNorm of
 * utils_pt.py:84,98 folded in by the caller; R carries the block residual (utils_pt.py:180,220).  The same
 * entry point serves the backward product dZ = dY * W_s + p .* Z + q.
 * Default precision is 3xTF32 (hi/lo operand split in shared memory, three MMAs per k-step): results agree
 * with an fp32 GEMM to ~1e-6 relative to |A||B|.  SN_GEMM_SINGLE_PASS issues one TF32 MMA (~1e-3).
 * Supported: N in {64, 128, 256}, K % 32 == 0, 16-byte aligned pointers, leading dimensions % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
#define SN_GEMM_SINGLE_PASS 1
#define SN_GEMM_ELU_BWD_LEFT 4  /* multiply output columns [0, N/2) by elu'(R) (R = activated values: 1 if R > 0 else R + 1); needs R */
#define SN_GEMM_NO_L2_PREFETCH 2 /* A/B switch: disable the L2 prefetch of upcoming operands (residual rows / split-K boxes) */
#define SN_GEMM_LEGACY_SS 8      /* A/B switch: the round-1 kernel (both operands from shared memory) instead of the TS-mode one */
size_t sn_gemm_tf32_ws_bytes(int64_t N, int64_t K); /* workspace for the pre-split weights (3xTF32 mode) */
int sn_gemm_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, const float* R,
                     int64_t ldr, const float* rscale, const float* group_bias, int64_t rows_per_group, float* C,
                     int64_t ldc, int64_t M, int64_t N, int64_t K, int flags, void* ws, size_t ws_bytes,
                     sn_stream_t stream);

/* Same product with the weights ALREADY split by the caller: B_hi = tf32(B), B_lo = B - B_hi, both [N x K] with leading
 * dimension ldb (sn_bn_fold_{fwd,bwd}_f32 emit them; sn_split_tf32_f32 splits any matrix).  No workspace, no per-launch
 * split kernel -- the form the fused stages use.  3xTF32 only. */
int sn_gemm_tf32_presplit_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                              const float* bias, const float* R, int64_t ldr, const float* rscale, const float* group_bias,
                              int64_t rows_per_group, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int flags,
                              sn_stream_t stream);
/* The dense stage with the NEXT stage's activation (utils_pt.py:161,172,195,208: F.elu in front of every operator
 * application) and that stage's left-half BatchNorm statistics (utils_pt.py:98) fused into the epilogue:
 *   raw   = A * (B_hi + B_lo)^T + bias + group_bias + rscale .* R          (C, may be NULL: raw value not stored)
 *   C_act = elu(raw)                                                        (leading dimension ldc_act: any row-strided view)
 *   act_mean / act_var = per-column mean / biased variance of C_act over the M rows (both NULL: no statistics)
 * Statistics are deterministic: per-warp shared-memory accumulators, per-CTA partials in `ws`
 * (sn_gemm_act_ws_bytes(N) bytes), fixed-order fp64 final reduction.  Flags must be 0. */
size_t sn_gemm_act_ws_bytes(int64_t N);
int sn_gemm_tf32_presplit_act_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                                  const float* bias, const float* R, int64_t ldr, const float* rscale,
                                  const float* group_bias, int64_t rows_per_group, float* C, int64_t ldc, float* C_act,
                                  int64_t ldc_act, float* act_mean, float* act_var, int64_t M, int64_t N, int64_t K,
                                  int flags, void* ws, size_t ws_bytes, sn_stream_t stream);
/* The dZ product of an AvgResNet2 stage backward with the ELU backward of utils_pt.py:231,237 in its epilogue (replaces
 * sn_gemm_tf32_presplit_f32 + sn_elu_bwd_group_f32: the intermediate dZ is never written):
 *   C = ((A * (B_hi + B_lo)^T + bias + row_scale[row] * group_bias[row / rows_per_group] + rscale .* R) .* elu'(R)) + R2
 * R = the activated values a = elu(x) (elu' = 1 if a > 0 else a + 1), required; row_scale [M] (the mask weights) and
 * R2 [M x N] (a gradient that bypasses the activation, e.g. the block residual's) may be NULL. */
int sn_gemm_tf32_presplit_elubwd_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                                     const float* bias, const float* R, int64_t ldr, const float* rscale,
                                     const float* group_bias, int64_t rows_per_group, const float* row_scale, const float* R2,
                                     int64_t ldr2, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, sn_stream_t stream);
/* All-pairs feature correlation of the dense_correspondence SiameseModel (dense_correspondence/models.py:199-203,
 * torch.bmm(FA, FB^T) per shape pair): C[M x N] = A[M x K] * (B_hi + B_lo)[N x K]^T for WIDE N (thousands of columns),
 * K <= 128, K % 4 == 0, N % 4 == 0 (K and N tails are zero-filled / clipped by the tensor maps: no padding copies).
 * Same tcgen05 3xTF32 kernel as above; the A row tile is split once into tensor memory and stays resident while its
 * column passes stream the B rows; (row tile, column group) work items fill all SMs; the result leaves through the
 * TMA-store epilogue (the product is bound by writing M x N floats). */
int sn_gemm_nt_wide_tf32_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb, float* C,
                             int64_t ldc, int64_t M, int64_t N, int64_t K, sn_stream_t stream);
/* hi = tf32(X) (round to nearest), lo = X - hi; X [rows x cols] with leading dimension ldx, outputs contiguous. */
int sn_split_tf32_f32(const float* X, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, sn_stream_t stream);

/* Weight-gradient product of a stage, reduction over the rows (split-K over the SMs, deterministic):
 *
 *   G[M x N] = A[R x M]^T * B[R x N]        M = 128, N % 32 == 0, 32 <= N <= 256
 *
 * With A = dY and B = Z (the stage's concat buffer) G gives dW of the Linear and, with colsum(dY), every reduction
 * the BatchNorm backward needs (utils_pt.py:91-104 through autograd in the reference).  Same 3xTF32 precision and
 * SN_GEMM_SINGLE_PASS flag as sn_gemm_tf32_f32; operands are read through MN-major tcgen05 descriptors. */
size_t sn_gemm_tn_tf32_ws_bytes(int64_t R, int64_t N);
int sn_gemm_tn_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* G, int64_t ldg, int64_t R,
                        int64_t M, int64_t N, int flags, void* ws, size_t ws_bytes, sn_stream_t stream);
/* The same product plus colsum_A[M] = column sums of A over the R rows, from the same pass (the A tile crosses the
 * registers of the warps that move it into tensor memory): with A = dY this is db of the Linear and the second reduction
 * the BatchNorm backward needs -- no separate pass over dY.  Deterministic (per-CTA partials added in a fixed order). */
int sn_gemm_tn_colsum_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* G, int64_t ldg,
                               float* colsum_A, int64_t R, int64_t M, int64_t N, int flags, void* ws, size_t ws_bytes,
                               sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Batch statistics for the training-mode BatchNorm in front of the Linear (utils_pt.py:84,98): per-column mean and
 * biased variance of X [rows x C] over all rows, one HBM pass, deterministic (fixed-order fp64 final reduction).
 * The normalisation itself is never run as a pass: the caller folds it into the weights given to sn_gemm_tf32_f32.
 * Supported: C % 4 == 0, C/4 divides 256, ldx % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
size_t sn_colstats_ws_bytes(int64_t C);
int sn_colstats_f32(const float* X, int64_t ldx, int64_t rows, int64_t C, float* mean, float* var_biased,
                    void* ws, size_t ws_bytes, sn_stream_t stream);
/* sn_elu_f32 and sn_colstats_f32 of its OUTPUT in one pass: Y = elu(X), mean / var_biased = statistics of Y's columns
 * (the left half of a stage's concat buffer needs no separate statistics pass).  Same workspace as sn_colstats_f32. */
int sn_elu_colstats_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t rows, int64_t C, float* mean,
                        float* var_biased, void* ws, size_t ws_bytes, sn_stream_t stream);

/* Y = S X together with the per-column mean / biased variance of Y over all rows, from ONE launch: the BatchNorm statistics
 * of the right half of a stage's concat buffer (utils_pt.py:98 applied to the torch.cat of :168,177,204,216) are taken in
 * the row-group kernel's store path (per-(warp, row group) shared-memory accumulators, per-CTA partials in `ws`,
 * fixed-order fp64 final reduction: bit-reproducible).  Y is bit-identical to sn_{bsr4,csr}_spmm_f32.
 * Row-group kernel only: C in {32, 64, 128, 256, 512}; SN_ERR_UNSUPPORTED otherwise (nothing launched). */
size_t sn_spmm_stats_ws_bytes(int64_t C);
int sn_bsr4_spmm_stats_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X, int64_t ldx,
                           float* Y, int64_t ldy, int64_t n_brows, int64_t C, float* mean, float* var_biased, int flags,
                           void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_csr_spmm_stats_f32(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                          float* Y, int64_t ldy, int64_t n_rows, int64_t C, float* mean, float* var_biased, int flags,
                          void* ws, size_t ws_bytes, sn_stream_t stream);

/* O(C^2) glue of the fused dense stage, one launch each way.
 * forward : s = gamma*rstd, t = beta - mean*s, Wf = W diag(s) [N x K], bf = b + W t, rstd = 1/sqrt(var+eps); when
 *           running_mean/var are given they are updated with `momentum` (unbiased variance, rows/(rows-1));
 *           num_batches_tracked (device int64, NULL = none) is incremented by one -- nn.BatchNorm's counter, kept by
 *           the kernel so that a training step launches no separate increment per layer.
 * backward: from G = dY^T Z [N x K] and sdY = colsum(dY) [N]:  dW, db, dgamma, dbeta, the coefficients p, q of
 *           dZ = dY (W diag(s)) + p .* Z + q (training-mode BatchNorm backward folded), and WsT = (W diag(s))^T [K x N]. */
int sn_bn_fold_fwd_f32(const float* mean, const float* var, const float* gamma, const float* beta, const float* W,
                       const float* b, int64_t N, int64_t K, float eps, float* Wf, float* bf, float* s, float* t,
                       float* rstd, float* running_mean, float* running_var, float momentum, int64_t rows,
                       float* Wf_hi, float* Wf_lo, int64_t* num_batches_tracked, sn_stream_t stream);
int sn_bn_fold_bwd_f32(const float* G, const float* sdY, const float* W, const float* s, const float* t,
                       const float* rstd, const float* mean, int64_t N, int64_t K, int64_t rows, int training,
                       float* dW, float* db, float* dgamma, float* dbeta, float* p, float* q, float* WsT,
                       float* WsT_hi, float* WsT_lo, sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * AvgResNet2 stage (utils_pt.py:222-243): x -> elu -> [a | global_average(a, mask) broadcast] -> BatchNorm -> Linear.
 * The broadcast half is never built (it is constant per mesh): the stage is a K = C GEMM plus a per-mesh bias.
 *   sn_avg_stage_pre_f32  A = elu(X) [n_seg * rows_per_seg x C]; mean / var_biased [2C] = training-mode statistics of
 *                         [A | avg broadcast]; avg [n_seg x C] = (sum_r w[r] A[r]) * inv_cnt[seg] (w NULL = 1).  One pass
 *                         over X plus a tiny reduction; workspace sn_avg_stage_ws_bytes(n_seg, C).
 *   sn_avg_fold_fwd_f32   sn_bn_fold_fwd_f32 for K = 2C (outputs only the tf32 hi / lo split of W' = W diag(s), [N x 2C])
 *                         plus u [n_seg x N] = b' + W'[:, C:] avg_seg: the group_bias of sn_gemm_tf32_presplit_f32.
 *   sn_avg_fold_bwd_f32   backward glue in one launch: with GL = dY^T A [N x C] and SdY [n_seg x N] = per-mesh sums of dY
 *                         it forms G = [GL | SdY^T avg], runs the folded BatchNorm backward of sn_bn_fold_bwd_f32 on it
 *                         (dW [N x 2C], db, dgamma, dbeta, p, q [2C], WsT hi / lo [2C x N]) and returns
 *                         gb [n_seg x C] = inv_cnt (SdY (W_R diag(s_R)) + rows_per_seg (p_R avg + q_R)), the gradient that
 *                         sn_elu_bwd_group_f32 hands back to every row of the mesh.
 * ---------------------------------------------------------------------------------------------- */
size_t sn_avg_stage_ws_bytes(int64_t n_seg, int64_t C);
int sn_avg_stage_pre_f32(const float* X, int64_t ldx, const float* w, const float* inv_cnt, int64_t rows_per_seg,
                         int64_t n_seg, int64_t C, float* A, int64_t lda, float* mean, float* var_biased, float* avg,
                         void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_avg_fold_fwd_f32(const float* mean, const float* var, const float* gamma, const float* beta, const float* W,
                        const float* b, int64_t N, int64_t C, float eps, float* Wf_hi, float* Wf_lo, float* s, float* t,
                        float* rstd, float* running_mean, float* running_var, float momentum, int64_t rows,
                        const float* avg, int64_t n_seg, float* u, int64_t* num_batches_tracked, sn_stream_t stream);
int sn_avg_fold_bwd_f32(const float* GL, int64_t ldgl, const float* SdY, const float* avg, const float* W, const float* s,
                        const float* t, const float* rstd, const float* mean, const float* inv_cnt, int64_t N, int64_t C,
                        int64_t n_seg, int64_t rows_per_seg, int training, float* dW, float* db, float* dgamma,
                        float* dbeta, float* p, float* q, float* WsT_hi, float* WsT_lo, float* gb, sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The two ends of the model stacks.
 * Input layer conv1 = GraphConv1x1(K -> N, no BatchNorm), K = 3 or 6 channels (as_rigid_as_possible/models.py:112,
 * dense_correspondence/models.py:144, mesh_mnist/models_vae.py:26; nn.Linear at utils_pt.py:89,99):
 *   sn_linear_smallk_fwd_f32   Y[rows x N] = X[rows x K] W[N x K]^T + b          (K in {3, 6}, N % 4 == 0; one pass, write-bound)
 *   sn_linear_smallk_bwd_f32   dW[N x K] = dY^T X, db[N] = colsum(dY) (db may be NULL) from ONE pass over dY; K in {3, 6},
 *                              N <= 256; deterministic (per-CTA partials, fixed-order fp64 final sum)
 * as_rigid_as_possible output head and loss (models.py:152, main.py:225-226):
 *   sn_head_add_tiled_f32      Out[r, j] = Y[r, j] + In[r, c_in - 3 + j % 3]     (`+ inputs[:, :, -3:].repeat(1, 1, 40)`)
 *   sn_head_pad_grad_f32       dYp[r, j] = j < n_out ? G[r, j] : 0               (gradient of the first n_out columns of the
 *                              zero-padded 128-wide conv2 output)
 *   sn_masked_smooth_l1_*      loss[0] = scale * sum smooth_l1(Out .* M - T)   (M: one weight per row, beta = 1);
 *                              dOut = grad_loss[0] * scale * M .* smooth_l1'(Out .* M - T)   (grad_loss NULL = 1)
 * ---------------------------------------------------------------------------------------------- */
int sn_linear_smallk_fwd_f32(const float* X, int64_t ldx, const float* W, const float* b, float* Y, int64_t ldy, int64_t rows,
                             int64_t N, int64_t K, sn_stream_t stream);
size_t sn_linear_smallk_bwd_ws_bytes(int64_t N, int64_t K);
int sn_linear_smallk_bwd_f32(const float* dY, int64_t ldd, const float* X, int64_t ldx, int64_t rows, int64_t N, int64_t K,
                             float* dW, float* db, void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_head_add_tiled_f32(const float* Y, int64_t ldy, const float* In, int64_t ldi, int64_t c_in, float* Out, int64_t ldo,
                          int64_t rows, int64_t n_out, sn_stream_t stream);
int sn_head_pad_grad_f32(const float* G, int64_t ldg, float* dYp, int64_t ldd, int64_t rows, int64_t n_out, int64_t n_pad,
                         sn_stream_t stream);
size_t sn_masked_smooth_l1_ws_bytes(void);
int sn_masked_smooth_l1_fwd_f32(const float* Out, const float* T, const float* M, int64_t rows, int64_t C, float scale,
                                float* loss, void* ws, size_t ws_bytes, sn_stream_t stream);
int sn_masked_smooth_l1_bwd_f32(const float* Out, const float* T, const float* M, const float* grad_loss, int64_t rows,
                                int64_t C, float scale, float* dOut, sn_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * One ResNet stage as ONE call (orchestration of the entry points above).  What the reference's cupy seam
 * (src/utils/cuda/sparse_bmm_func.py:27-72) plus F.elu / torch.cat / GraphConv1x1 (utils_pt.py:161-169, 195-205, 208-218)
 * do per stage:
 *   forward   Y = Linear(BatchNorm_train([ elu(x_self) | S elu(x_gather) ])) (+ residual)
 *             Dirac:     S = D (x_self = f, x_gather = v) or S = D* (x_self = v, x_gather = f_out)
 *             Laplacian: S = L, x_gather = x_self (the SpMM gathers from the activated left half of Z)
 *             saved for backward (caller-allocated): Z [rows_out x 2C], act_gather [rows_in x C] (Dirac), stk [3 x 2C]
 *             (s | t | rstd), mean [2C]; var_biased [2C] is an output only.  running_mean / running_var may be NULL.
 *   backward  dZ [rows_out x 2C] (scratch and result: for the Dirac stage its LEFT half is the gradient of x_self),
 *             d_gather [rows_in x C] = gradient of x_gather (Laplacian: the full gradient of x) + g_extra (optional, e.g.
 *             a residual-path gradient), dgamma / dbeta [2C], dW [C x 2C], db [C].  t_* = the TRANSPOSED operator's arrays.
 * C = 128 (the width of every reference model); workspaces from sn_stage_{fwd,bwd}_ws_bytes.
 * ---------------------------------------------------------------------------------------------- */
size_t sn_stage_fwd_ws_bytes(int64_t C);
size_t sn_stage_bwd_ws_bytes(int64_t rows_out, int64_t C);
int sn_dir_stage_fwd_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, int64_t n_brows, int64_t n_bcols,
                         const float* x_self, int64_t ld_self, const float* x_gather, int64_t ld_gather, int64_t C,
                         const float* gamma, const float* beta, const float* W, const float* b, const float* residual,
                         int64_t ldr, float* running_mean, float* running_var, float momentum, float eps, float* Z,
                         float* act_gather, float* stk, float* mean, float* var_biased, float* Y, int64_t ldy, void* ws,
                         size_t ws_bytes, sn_stream_t stream);
int sn_lap_stage_fwd_f32(const int32_t* rowptr, const int32_t* colind, const float* val, int64_t n_rows, const float* x,
                         int64_t ldx, int64_t C, const float* gamma, const float* beta, const float* W, const float* b,
                         const float* residual, int64_t ldr, float* running_mean, float* running_var, float momentum, float eps,
                         float* Z, float* stk, float* mean, float* var_biased, float* Y, int64_t ldy, void* ws, size_t ws_bytes,
                         sn_stream_t stream);
int sn_dir_stage_bwd_f32(const int32_t* t_browptr, const int32_t* t_bcolind, const float* t_bval, int64_t rows_out,
                         int64_t rows_in, const float* dY, int64_t ldd, const float* Z, const float* act_gather, const float* W,
                         const float* stk, const float* mean, int64_t C, float* dZ, float* d_gather, int64_t ld_dg,
                         const float* g_extra, int64_t ld_ge, float* dgamma, float* dbeta, float* dW, float* db, void* ws,
                         size_t ws_bytes, sn_stream_t stream);
int sn_lap_stage_bwd_f32(const int32_t* t_rowptr, const int32_t* t_colind, const float* t_val, int64_t n_rows, const float* dY,
                         int64_t ldd, const float* Z, const float* W, const float* stk, const float* mean, int64_t C, float* dZ,
                         float* dx, int64_t ld_dx, const float* g_extra, int64_t ld_ge, float* dgamma, float* dbeta, float* dW,
                         float* db, void* ws, size_t ws_bytes, sn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SURFNET_B200_H */
