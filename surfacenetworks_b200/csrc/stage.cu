// stage.cu -- one ResNet stage of the reference as ONE call of the C ABI:
//
//   forward   Y = Linear(BatchNorm([ elu(x_self) | S elu(x_gather) ])) (+ residual)
//             reference src/utils/utils_pt.py:161-169 / 172-178 (LapResNet2: S = L, x_gather = x_self),
//             :195-205 (DirResNet2 faces <- vertices: S = D) and :208-218 (vertices <- faces: S = D*)
//   backward  everything autograd derives from that expression: gradients of x_self, x_gather, gamma, beta, W, b
//
// These are orchestration only -- every kernel they launch is also exported on its own (surfnet_b200.h) -- so that a
// non-Python host (the reference's cupy seam, src/utils/cuda/sparse_bmm_func.py:27-72, or a C++ trainer) gets the fused
// stage without re-implementing the sequencing, workspace carving and statistics plumbing of fused.py / ops.py:
//
//   fwd:  sn_elu_colstats_f32 (left half + its statistics)  [sn_elu_f32 (gather operand, Dirac only)]
//         sn_{bsr4,csr}_spmm_stats_f32 (right half + its statistics; falls back to spmm + sn_colstats_f32)
//         sn_bn_fold_fwd_f32 -> sn_gemm_tf32_presplit_f32 (residual in the epilogue)
//   bwd:  sn_gemm_tn_colsum_tf32_f32 (G = dY^T Z, colsum dY) -> sn_bn_fold_bwd_f32 -> sn_gemm_tf32_presplit_f32
//         (dZ = dY W_s + p Z + q, elu'(x_self) applied to the left half) -> sn_{bsr4,csr}_spmm_epilogue_f32 on S^T
//         ((S^T dZ_right [+ dZ_left]) .* elu'(x_gather))
//
// Supported: training-mode BatchNorm, C = 128 (the width of every reference model; both GEMM shapes and the split-K
// weight-gradient product on the tensor-core kernels, the SpMM on the row-group kernel), fp32, 16-byte aligned row-major operands.  Anything else: SN_ERR_UNSUPPORTED before the first launch.
#include "common.cuh"

namespace sn {
namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline char* carve(char*& p, size_t bytes) {
  char* r = p;
  p += align256(bytes);
  return r;
}
inline size_t max2(size_t a, size_t b) { return a > b ? a : b; }

struct Op {
  int blk;                    // 4: BSR4 (Dirac), 1: CSR (Laplacian)
  const int32_t* ptr;
  const int32_t* ind;
  const float* val;
};

int stage_shape_ok(int64_t rows_out, int64_t rows_in, int64_t C) {
  if (rows_out <= 0 || rows_in <= 0) return SN_ERR_ARG;
  if (C != 128) return SN_ERR_UNSUPPORTED;     // the split-K weight-gradient kernel takes 128 output channels
  return SN_OK;
}

int stage_fwd(const Op& S, int64_t rows_out, int64_t rows_in, const float* x_self, int64_t ld_self, const float* x_gather,
              int64_t ld_gather, int64_t C, const float* gamma, const float* beta, const float* W, const float* b,
              const float* residual, int64_t ldr, float* running_mean, float* running_var, float momentum, float eps, float* Z,
              float* act_gather, float* stk, float* mean, float* var, float* Y, int64_t ldy, void* ws, size_t ws_bytes,
              sn_stream_t stream) {
  const int64_t K = 2 * C;
  int rc = stage_shape_ok(rows_out, rows_in, C);
  if (rc != SN_OK) return rc;
  if (!S.ptr || !S.ind || !S.val || !x_self || !gamma || !beta || !W || !b || !Z || !stk || !mean || !var || !Y) return SN_ERR_ARG;
  const bool same = x_gather == nullptr;           // Laplacian stage: the SpMM gathers from the activated left half of Z
  if (same && rows_in != rows_out) return SN_ERR_ARG;
  if (!same && !act_gather) return SN_ERR_ARG;
  if (!ws || ws_bytes < sn_stage_fwd_ws_bytes(C)) return SN_ERR_WORKSPACE;
  char* p = reinterpret_cast<char*>(ws);
  float* Wf = reinterpret_cast<float*>(carve(p, (size_t)C * K * 4));
  float* Wf_hi = reinterpret_cast<float*>(carve(p, (size_t)C * K * 4));
  float* Wf_lo = reinterpret_cast<float*>(carve(p, (size_t)C * K * 4));
  float* bf = reinterpret_cast<float*>(carve(p, (size_t)C * 4));
  const size_t stat_bytes = max2(sn_colstats_ws_bytes(C), sn_spmm_stats_ws_bytes(C));
  void* stat_ws = carve(p, stat_bytes);

  rc = sn_elu_colstats_f32(x_self, ld_self, Z, K, rows_out, C, mean, var, stat_ws, stat_bytes, stream);
  if (rc != SN_OK) return rc;
  const float* gat = Z;                            // gather operand: [rows_in x C] with leading dimension ldg
  int64_t ldg = K;
  if (!same) {
    rc = sn_elu_f32(x_gather, ld_gather, act_gather, C, rows_in, C, stream);
    if (rc != SN_OK) return rc;
    gat = act_gather;
    ldg = C;
  }
  rc = S.blk == 4 ? sn_bsr4_spmm_stats_f32(S.ptr, S.ind, S.val, gat, ldg, Z + C, K, rows_out, C, mean + C, var + C, 0, stat_ws,
                                           stat_bytes, stream)
                  : sn_csr_spmm_stats_f32(S.ptr, S.ind, S.val, gat, ldg, Z + C, K, rows_out, C, mean + C, var + C, 0, stat_ws,
                                          stat_bytes, stream);
  if (rc == SN_ERR_UNSUPPORTED) {                  // two passes
    rc = S.blk == 4 ? sn_bsr4_spmm_f32(S.ptr, S.ind, S.val, gat, ldg, Z + C, K, rows_out, C, 0, stream)
                    : sn_csr_spmm_f32(S.ptr, S.ind, S.val, gat, ldg, Z + C, K, rows_out, C, 0, stream);
    if (rc != SN_OK) return rc;
    rc = sn_colstats_f32(Z + C, K, rows_out, C, mean + C, var + C, stat_ws, stat_bytes, stream);
  }
  if (rc != SN_OK) return rc;
  rc = sn_bn_fold_fwd_f32(mean, var, gamma, beta, W, b, C, K, eps, Wf, bf, stk, stk + K, stk + 2 * K, running_mean, running_var,
                          momentum, rows_out, Wf_hi, Wf_lo, nullptr, stream);
  if (rc != SN_OK) return rc;
  return sn_gemm_tf32_presplit_f32(Z, K, Wf_hi, Wf_lo, K, bf, residual, ldr, nullptr, nullptr, 0, Y, ldy, rows_out, C, K, 0, stream);
}

int stage_bwd(const Op& ST, int64_t rows_out, int64_t rows_in, const float* dY, int64_t ldd, const float* Z,
              const float* act_gather, const float* W, const float* stk, const float* mean, int64_t C, float* dZ,
              float* d_gather, int64_t ld_dg, const float* g_extra, int64_t ld_ge, float* dgamma, float* dbeta, float* dW,
              float* db, void* ws, size_t ws_bytes, sn_stream_t stream) {
  const int64_t K = 2 * C;
  int rc = stage_shape_ok(rows_out, rows_in, C);
  if (rc != SN_OK) return rc;
  if (!ST.ptr || !ST.ind || !ST.val || !dY || !Z || !W || !stk || !mean || !dZ || !d_gather || !dgamma || !dbeta || !dW || !db)
    return SN_ERR_ARG;
  const bool same = act_gather == nullptr;
  if (same && rows_in != rows_out) return SN_ERR_ARG;
  if (!ws || ws_bytes < sn_stage_bwd_ws_bytes(rows_out, C)) return SN_ERR_WORKSPACE;
  char* p = reinterpret_cast<char*>(ws);
  const size_t tn_bytes = sn_gemm_tn_tf32_ws_bytes(rows_out, K);
  void* tn_ws = carve(p, tn_bytes);
  float* G = reinterpret_cast<float*>(carve(p, (size_t)C * K * 4));
  float* sdY = reinterpret_cast<float*>(carve(p, (size_t)C * 4));
  float* WsT = reinterpret_cast<float*>(carve(p, (size_t)K * C * 4));
  float* WsT_hi = reinterpret_cast<float*>(carve(p, (size_t)K * C * 4));
  float* WsT_lo = reinterpret_cast<float*>(carve(p, (size_t)K * C * 4));
  float* pq = reinterpret_cast<float*>(carve(p, (size_t)2 * K * 4));

  rc = sn_gemm_tn_colsum_tf32_f32(dY, ldd, Z, K, G, K, sdY, rows_out, C, K, 0, tn_ws, tn_bytes, stream);
  if (rc != SN_OK) return rc;
  rc = sn_bn_fold_bwd_f32(G, sdY, W, stk, stk + K, stk + 2 * K, mean, C, K, rows_out, 1, dW, db, dgamma, dbeta, pq, pq + K, WsT,
                          WsT_hi, WsT_lo, stream);
  if (rc != SN_OK) return rc;
  // dZ = dY (W diag(s)) + p .* Z + q.  Dirac stage: the left half leaves already multiplied by elu'(x_self) (= the gradient
  // of x_self).  Laplacian stage: x_self is also the gather operand, the derivative is applied once, by the SpMM below.
  rc = sn_gemm_tf32_presplit_f32(dY, ldd, WsT_hi, WsT_lo, C, pq + K, Z, K, pq, nullptr, 0, dZ, K, rows_out, K, C,
                                 same ? 0 : SN_GEMM_ELU_BWD_LEFT, stream);
  if (rc != SN_OK) return rc;
  const float* A = same ? Z : act_gather;
  const int64_t lda = same ? K : C;
  const float* Gl = same ? dZ : nullptr;           // Laplacian: (S^T dZ_right + dZ_left) .* elu'(x)
  return ST.blk == 4 ? sn_bsr4_spmm_epilogue_f32(ST.ptr, ST.ind, ST.val, dZ + C, K, d_gather, ld_dg, rows_in, C, Gl, K, A, lda,
                                                 g_extra, ld_ge, 0, stream)
                     : sn_csr_spmm_epilogue_f32(ST.ptr, ST.ind, ST.val, dZ + C, K, d_gather, ld_dg, rows_in, C, Gl, K, A, lda,
                                                g_extra, ld_ge, 0, stream);
}

}  // namespace
}  // namespace sn

SN_API size_t sn_stage_fwd_ws_bytes(int64_t C) {
  using namespace sn;
  if (C <= 0) return 0;
  const size_t K = 2 * (size_t)C;
  return 3 * align256((size_t)C * K * 4) + align256((size_t)C * 4) +
         align256(max2(sn_colstats_ws_bytes(C), sn_spmm_stats_ws_bytes(C)));
}

SN_API size_t sn_stage_bwd_ws_bytes(int64_t rows_out, int64_t C) {
  using namespace sn;
  if (C <= 0 || rows_out <= 0) return 0;
  const size_t K = 2 * (size_t)C;
  return align256(sn_gemm_tn_tf32_ws_bytes(rows_out, (int64_t)K)) + 4 * align256((size_t)C * K * 4) + align256((size_t)C * 4) +
         align256(2 * K * 4);
}

SN_API int sn_dir_stage_fwd_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, int64_t n_brows,
                                int64_t n_bcols, const float* x_self, int64_t ld_self, const float* x_gather, int64_t ld_gather,
                                int64_t C, const float* gamma, const float* beta, const float* W, const float* b,
                                const float* residual, int64_t ldr, float* running_mean, float* running_var, float momentum,
                                float eps, float* Z, float* act_gather, float* stk, float* mean, float* var_biased, float* Y,
                                int64_t ldy, void* ws, size_t ws_bytes, sn_stream_t stream) {
  if (!x_gather) return SN_ERR_ARG;
  return sn::stage_fwd(sn::Op{4, browptr, bcolind, bval}, n_brows, n_bcols, x_self, ld_self, x_gather, ld_gather, C, gamma, beta,
                       W, b, residual, ldr, running_mean, running_var, momentum, eps, Z, act_gather, stk, mean, var_biased, Y, ldy,
                       ws, ws_bytes, stream);
}

SN_API int sn_lap_stage_fwd_f32(const int32_t* rowptr, const int32_t* colind, const float* val, int64_t n_rows, const float* x,
                                int64_t ldx, int64_t C, const float* gamma, const float* beta, const float* W, const float* b,
                                const float* residual, int64_t ldr, float* running_mean, float* running_var, float momentum,
                                float eps, float* Z, float* stk, float* mean, float* var_biased, float* Y, int64_t ldy, void* ws,
                                size_t ws_bytes, sn_stream_t stream) {
  return sn::stage_fwd(sn::Op{1, rowptr, colind, val}, n_rows, n_rows, x, ldx, nullptr, 0, C, gamma, beta, W, b, residual, ldr,
                       running_mean, running_var, momentum, eps, Z, nullptr, stk, mean, var_biased, Y, ldy, ws, ws_bytes, stream);
}

SN_API int sn_dir_stage_bwd_f32(const int32_t* t_browptr, const int32_t* t_bcolind, const float* t_bval, int64_t rows_out,
                                int64_t rows_in, const float* dY, int64_t ldd, const float* Z, const float* act_gather,
                                const float* W, const float* stk, const float* mean, int64_t C, float* dZ, float* d_gather,
                                int64_t ld_dg, const float* g_extra, int64_t ld_ge, float* dgamma, float* dbeta, float* dW,
                                float* db, void* ws, size_t ws_bytes, sn_stream_t stream) {
  if (!act_gather) return SN_ERR_ARG;
  return sn::stage_bwd(sn::Op{4, t_browptr, t_bcolind, t_bval}, rows_out, rows_in, dY, ldd, Z, act_gather, W, stk, mean, C, dZ,
                       d_gather, ld_dg, g_extra, ld_ge, dgamma, dbeta, dW, db, ws, ws_bytes, stream);
}

SN_API int sn_lap_stage_bwd_f32(const int32_t* t_rowptr, const int32_t* t_colind, const float* t_val, int64_t n_rows,
                                const float* dY, int64_t ldd, const float* Z, const float* W, const float* stk,
                                const float* mean, int64_t C, float* dZ, float* dx, int64_t ld_dx, const float* g_extra,
                                int64_t ld_ge, float* dgamma, float* dbeta, float* dW, float* db, void* ws, size_t ws_bytes,
                                sn_stream_t stream) {
  return sn::stage_bwd(sn::Op{1, t_rowptr, t_colind, t_val}, n_rows, n_rows, dY, ldd, Z, nullptr, W, stk, mean, C, dZ, dx, ld_dx,
                       g_extra, ld_ge, dgamma, dbeta, dW, db, ws, ws_bytes, stream);
}
