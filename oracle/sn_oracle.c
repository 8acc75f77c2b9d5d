/* sn_oracle.c -- CPU restatement of the reference's operator-application arithmetic.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (surfacenetworks_b200/) never links, imports or calls anything under oracle/.
 *
 * What is restated, and from where:
 *   oracle_coo_mm_*      torch.mm(sparse_coo, dense) as called at reference src/utils/utils_pt.py:167,176,202,214.
 *                        The arithmetic lives in PyTorch ATen (third-party; the reference pins torch==1.0.0,
 *                        README.md:45; this image has 2.11.0): for a coalesced COO matrix the CPU kernel walks the
 *                        non-zeros in (row, col) order and accumulates val * dense[col, :] into out[row, :]
 *                        (an axpy per non-zero), i.e. fp32 sums in ascending storage order.
 *   oracle_batch_csr     src/utils/cuda/batch_csr.cu:13-47 -- [3, nnz] sorted COO -> col_ind[nnz] +
 *                        col_ptr[B, R+1] with GLOBAL nnz offsets.  Restated with the intended semantics: an
 *                        empty row gets an empty range (the reference kernel leaves 0 there, batch_csr.py:48-49).
 *   oracle_sparse_bmm    src/utils/cuda/sparse_bmm.cu:16-61 -- C[b,i,j] = sum_k values[k] * dense[b, col_ind[k], j]
 *                        for k in [col_ptr[b,i], col_ptr[b,i+1]).
 *   oracle_dirac_view_mm the quaternion `view` of utils_pt.py:201-203,213-215: the [4R x 4Cn] operator applied to
 *                        x.view(Cn*4, C/4) and viewed back as [R, C].
 * The _f64 variants accumulate in double: ground truth for the componentwise bound
 *   |y - y_ref| <= 32 eps_32 (|S| |x|)   (SURVEY.md section 8(c)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* out[n_rows, C] = S @ X, S given as COO in storage order; fp32 accumulation like ATen's axpy loop */
API void oracle_coo_mm_f32(const int64_t* row, const int64_t* col, const float* val, int64_t nnz,
                           int64_t n_rows, const float* X, int64_t ldx, float* out, int64_t ldo, int64_t C) {
  for (int64_t r = 0; r < n_rows; ++r) memset(out + r * ldo, 0, sizeof(float) * (size_t)C);
  for (int64_t k = 0; k < nnz; ++k) {
    const float v = val[k];
    const float* x = X + col[k] * ldx;
    float* o = out + row[k] * ldo;
    for (int64_t c = 0; c < C; ++c) o[c] += v * x[c];
  }
}

/* same product in double; also returns bound[r,c] = sum_k |val| * |x| (the |S||x| of the error bound) */
API void oracle_coo_mm_f64(const int64_t* row, const int64_t* col, const float* val, int64_t nnz,
                           int64_t n_rows, const float* X, int64_t ldx, double* out, double* bound, int64_t C) {
  memset(out, 0, sizeof(double) * (size_t)(n_rows * C));
  if (bound) memset(bound, 0, sizeof(double) * (size_t)(n_rows * C));
  for (int64_t k = 0; k < nnz; ++k) {
    const double v = (double)val[k];
    const float* x = X + col[k] * ldx;
    double* o = out + row[k] * C;
    for (int64_t c = 0; c < C; ++c) o[c] += v * (double)x[c];
    if (bound) {
      double* b = bound + row[k] * C;
      for (int64_t c = 0; c < C; ++c) b[c] += fabs(v) * fabs((double)x[c]);
    }
  }
}

/* batch_csr: indices is [3, nnz] (batch, row, col), sorted.  col_ptr is [B, R+1], global offsets. */
API void oracle_batch_csr(const int64_t* indices, int64_t nnz, int64_t B, int64_t R, int64_t* col_ind,
                          int64_t* col_ptr) {
  int64_t k = 0;
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t r = 0; r <= R; ++r) {
      /* first entry at or after (b, r) */
      while (k < nnz && (indices[k] < b || (indices[k] == b && indices[k + nnz] < r))) ++k;
      col_ptr[b * (R + 1) + r] = k;
    }
  }
  for (int64_t i = 0; i < nnz; ++i) col_ind[i] = indices[i + 2 * nnz];
}

/* sparse_bmm: values/col_ind/col_ptr as produced above; dense [B, Rd, Cd]; C [B, R, Cd] */
API void oracle_sparse_bmm(const float* values, const int64_t* col_ind, const int64_t* col_ptr, int64_t B, int64_t R,
                           const float* dense, int64_t Rd, int64_t Cd, float* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t i = 0; i < R; ++i) {
      const int64_t s = col_ptr[b * (R + 1) + i], e = col_ptr[b * (R + 1) + i + 1];
      for (int64_t j = 0; j < Cd; ++j) {
        float acc = 0.0f;
        for (int64_t k = s; k < e; ++k) acc += values[k] * dense[b * Rd * Cd + col_ind[k] * Cd + j];
        out[b * R * Cd + i * Cd + j] = acc;
      }
    }
}

/* Dirac view: S is [4R x 4Cn] COO; X is [Cn, C] node features; out [R, C].
 * out[r, p*C4 + c] = sum over entries (4r+p, 4j+q): val * X[j, q*C4 + c]          (utils_pt.py:201-203) */
API void oracle_dirac_view_mm_f32(const int64_t* row, const int64_t* col, const float* val, int64_t nnz, int64_t R,
                                  const float* X, float* out, int64_t C) {
  const int64_t C4 = C / 4;
  memset(out, 0, sizeof(float) * (size_t)(R * C));
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t r = row[k] / 4, p = row[k] % 4, j = col[k] / 4, q = col[k] % 4;
    const float v = val[k];
    const float* x = X + j * C + q * C4;
    float* o = out + r * C + p * C4;
    for (int64_t c = 0; c < C4; ++c) o[c] += v * x[c];
  }
}

API void oracle_dirac_view_mm_f64(const int64_t* row, const int64_t* col, const float* val, int64_t nnz, int64_t R,
                                  const float* X, double* out, double* bound, int64_t C) {
  const int64_t C4 = C / 4;
  memset(out, 0, sizeof(double) * (size_t)(R * C));
  if (bound) memset(bound, 0, sizeof(double) * (size_t)(R * C));
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t r = row[k] / 4, p = row[k] % 4, j = col[k] / 4, q = col[k] % 4;
    const double v = (double)val[k];
    const float* x = X + j * C + q * C4;
    double* o = out + r * C + p * C4;
    for (int64_t c = 0; c < C4; ++c) o[c] += v * (double)x[c];
    if (bound) {
      double* b = bound + r * C + p * C4;
      for (int64_t c = 0; c < C4; ++c) b[c] += fabs(v) * fabs((double)x[c]);
    }
  }
}

/* ELU(alpha=1) as torch's CPU kernel evaluates it: expm1 on the negative branch (F.elu, utils_pt.py:161). */
API void oracle_elu_f32(const float* x, float* y, int64_t n) {
  for (int64_t i = 0; i < n; ++i) y[i] = x[i] > 0.0f ? x[i] : expm1f(x[i]);
}
