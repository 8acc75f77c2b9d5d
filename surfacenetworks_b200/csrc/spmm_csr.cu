// spmm_csr.cu -- scalar-CSR SpMM  Y = S * X  (cotangent Laplacian application).
//
// Replaces torch.mm(L, x.view(-1, feat)) at reference src/utils/utils_pt.py:167,176 and the reference's
// own batched kernel src/utils/cuda/sparse_bmm.cu:16-61 (one thread per output element, batch index on
// threadIdx.x => dense reads strided by R*C floats).  Here the feature dimension is the fast axis:
//
//   * LANES lanes (LANES*4 = padded feature width, <= 32) own one sparse row; a lane owns the float4
//     column slice c = 4*sl .. 4*sl+3 (+ 4*LANES per extra chunk), so every gathered dense row is read
//     with coalesced 128-bit loads (C = 128: one 512 B row per warp instruction);
//   * the row's (colind, val) segment is fetched once by the group's lanes (coalesced) and broadcast with
//     shuffles -- no shared memory, no atomics;
//   * U gathers are issued back to back before their FMAs (memory-level parallelism; rows have ~7 nnz);
//   * fp32 FMA accumulation in ascending storage order => bit-reproducible.
//
// Bound: HBM.  Algorithmic bytes per launch = 4(R+1) + 8 nnz + 4 R C (X once) + 4 R C (Y once)
// (BASELINE.md section 3).
#include "common.cuh"

namespace sn {

template <int LANES, bool ELU>
__global__ void __launch_bounds__(256)
csr_spmm_vec4_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                     const float* __restrict__ val, const float* __restrict__ X, int64_t ldx,
                     float* __restrict__ Y, int64_t ldy, int64_t n_rows, int C) {
  constexpr int RPW = kWarp / LANES;  // rows per warp
  constexpr int U = 4;                // gathers in flight per lane
  const int lane = threadIdx.x & 31;
  const int sl = lane % LANES;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp * RPW + lane / LANES;
  const bool row_ok = row < n_rows;

  int start = 0, end = 0;
  if (row_ok) {
    start = __ldg(rowptr + row);
    end = __ldg(rowptr + row + 1);
  }
  // Loop bounds must be warp-uniform (shuffles inside): iterate to the longest row of the warp.
  int len = end - start;
  int maxlen = len;
  if (RPW > 1) {
#pragma unroll
    for (int m = LANES; m < kWarp; m <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, m));
  }

  for (int c0 = 0; c0 < C; c0 += 4 * LANES) {  // feature chunks (one for C <= 128)
    const int c = c0 + 4 * sl;
    const bool col_ok = c < C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = 0; base < maxlen; base += LANES) {
      const int k = start + base + sl;
      int my_col = 0;
      float my_val = 0.f;
      if (k < end) {
        my_col = __ldg(colind + k);
        my_val = __ldg(val + k);
      }
      const int cnt = min(LANES, len - base);  // entries of this row in the batch (may be <= 0)
      int maxcnt = min(LANES, maxlen - base);
      for (int j0 = 0; j0 < maxcnt; j0 += U) {
        float4 xv[U];
        float w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u;
          const int src = j < LANES ? j : 0;
          const int col = __shfl_sync(0xffffffffu, my_col, src, LANES);
          w[u] = __shfl_sync(0xffffffffu, my_val, src, LANES);
          xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < cnt && col_ok) {
            xv[u] = ldg_f4(X + (int64_t)col * ldx + c);
            if (ELU) xv[u] = elu4(xv[u]);
          } else {
            w[u] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc = fma4(w[u], xv[u], acc);
      }
    }
    if (row_ok && col_ok) st_stream_f4(Y + row * ldy + c, acc);
  }
}

// Any C / any alignment: one warp per row, lanes stride over scalar columns.
template <bool ELU>
__global__ void __launch_bounds__(256)
csr_spmm_scalar_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                       const float* __restrict__ val, const float* __restrict__ X, int64_t ldx,
                       float* __restrict__ Y, int64_t ldy, int64_t n_rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int start = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int k = start; k < end; ++k) {
      float x = __ldg(X + (int64_t)__ldg(colind + k) * ldx + c);
      if (ELU) x = elu1(x);
      acc = fmaf(__ldg(val + k), x, acc);
    }
    Y[row * ldy + c] = acc;
  }
}

template <int LANES>
static int launch_vec4(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X,
                       int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int C, bool elu, cudaStream_t st) {
  constexpr int RPW = kWarp / LANES;
  constexpr int kThreads = 256;
  const int64_t rows_per_cta = (int64_t)RPW * (kThreads / 32);
  const int64_t grid = ceil_div(n_rows, rows_per_cta);
  if (grid > 0x7fffffffLL) return SN_ERR_OVERFLOW;
  if (elu)
    csr_spmm_vec4_kernel<LANES, true><<<(unsigned)grid, kThreads, 0, st>>>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C);
  else
    csr_spmm_vec4_kernel<LANES, false><<<(unsigned)grid, kThreads, 0, st>>>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C);
  return launch_status();
}

int launch_csr_rowgroup(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                        float* Y, int64_t ldy, int64_t n_rows, int64_t C, bool elu, int variant, const float* G,
                        int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2, cudaStream_t st,
                        float* stat_partial = nullptr, int* grid_out = nullptr);
int64_t rowgroup_max_grid();
// spmm_rowdirect.cu: small operators (one or two waves of threads)
bool rowdirect_applies(int64_t n_rows, int64_t C, bool three_in_flight);
int launch_csr_rowdirect(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                         float* Y, int64_t ldy, int64_t n_rows, int64_t nnz, int64_t C, cudaStream_t st);

}  // namespace sn

SN_API int sn_csr_spmm_f32(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X,
                           int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int64_t C, int flags,
                           sn_stream_t stream) {
  using namespace sn;
  if (n_rows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (n_rows == 0 || C == 0) return SN_OK;
  if (!rowptr || !X || !Y || ldx < C || ldy < C) return SN_ERR_ARG;
  if ((!colind || !val)) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool elu = (flags & SN_SPMM_ELU_INPUT) != 0;
  const bool vec_ok = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y);
  if (!vec_ok) {
    const int64_t grid = ceil_div(n_rows, 8);
    if (grid > 0x7fffffffLL) return SN_ERR_OVERFLOW;
    if (elu)
      csr_spmm_scalar_kernel<true><<<(unsigned)grid, 256, 0, st>>>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C);
    else
      csr_spmm_scalar_kernel<false><<<(unsigned)grid, 256, 0, st>>>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C);
    return launch_status();
  }
  if (!(flags & SN_SPMM_DIRECT_GATHER)) {
    const int variant = (flags >> 8) & 15, hint = (flags >> 12) & 15;
    // small operators (a mesh_mnist batch, a single mesh): the latency-oriented kernel, bit-identical results
    if (!elu && aligned16(val) && variant != 7 && (variant == 6 || (variant == 0 && rowdirect_applies(n_rows, C, hint >= 1 && hint <= 3)))) {
      const int rc = launch_csr_rowdirect(rowptr, colind, val, X, ldx, Y, ldy, n_rows, hint >= 1 && hint <= 3 ? 3 * n_rows : -1,
                                          C, st);
      if (rc != SN_ERR_UNSUPPORTED) return rc;
    }
    // row-group kernel: C = 32 ... 512
    const int rc = launch_csr_rowgroup(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C, elu, (variant == 6 || variant == 7) ? 0 : variant, nullptr, 0,
                                       nullptr, 0, nullptr, 0, st);
    if (rc != SN_ERR_UNSUPPORTED) return rc;
  }
  const int64_t v = C / 4;  // float4 columns
  if (v <= 1) return launch_vec4<1>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
  if (v <= 2) return launch_vec4<2>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
  if (v <= 4) return launch_vec4<4>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
  if (v <= 8) return launch_vec4<8>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
  if (v <= 16) return launch_vec4<16>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
  return launch_vec4<32>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, (int)C, elu, st);
}

// Y = (S X + G) .* elu'(A), see sn_bsr4_spmm_epilogue_f32.
SN_API int sn_csr_spmm_epilogue_f32(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X,
                                    int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int64_t C, const float* G,
                                    int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2, int flags,
                                     sn_stream_t stream) {
  using namespace sn;
  if (n_rows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (n_rows == 0 || C == 0) return SN_OK;
  if (!rowptr || !colind || !val || !X || !Y || ldx < C || ldy < C || (G && ldg < C) || (A && lda < C) ||
      (G2 && ldg2 < C))
    return SN_ERR_ARG;
  if (flags & (SN_SPMM_DIRECT_GATHER | SN_SPMM_ELU_INPUT)) return SN_ERR_UNSUPPORTED;
  if (C % 16 || ldx % 4 || ldy % 4 || ldg % 4 || lda % 4 || ldg2 % 4 || !aligned16(G2) || !aligned16(X) || !aligned16(Y) || !aligned16(G) || !aligned16(A))
    return SN_ERR_UNSUPPORTED;
  return launch_csr_rowgroup(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C, false, (flags >> 8) & 15, G, ldg, A, lda, G2, ldg2,
                             (cudaStream_t)stream);
}

// CSR twin of sn_bsr4_spmm_stats_f32 (Laplacian stages, utils_pt.py:98 after :167 / :176).
SN_API int sn_csr_spmm_stats_f32(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X,
                                 int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int64_t C, float* mean,
                                 float* var_biased, int flags, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (n_rows <= 0 || C <= 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (!rowptr || !colind || !val || !X || !Y || !mean || !var_biased || ldx < C || ldy < C) return SN_ERR_ARG;
  if (flags & (SN_SPMM_DIRECT_GATHER | SN_SPMM_ELU_INPUT)) return SN_ERR_UNSUPPORTED;
  if (C % 16 || ldx % 4 || ldy % 4 || !aligned16(X) || !aligned16(Y)) return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < (size_t)rowgroup_max_grid() * 2 * (size_t)C * sizeof(float) || !aligned16(ws)) return SN_ERR_WORKSPACE;
  int grid = 0;
  const int rc = launch_csr_rowgroup(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C, false, (flags >> 8) & 15, nullptr, 0, nullptr,
                                     0, nullptr, 0, (cudaStream_t)stream, reinterpret_cast<float*>(ws), &grid);
  if (rc != SN_OK) return rc;
  return launch_colstats_final(reinterpret_cast<const float*>(ws), grid, n_rows, (int)C, nullptr, mean, var_biased,
                               (cudaStream_t)stream);
}
