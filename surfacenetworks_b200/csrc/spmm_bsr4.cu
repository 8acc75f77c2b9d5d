// spmm_bsr4.cu -- 4x4-block CSR SpMM for the quaternion Dirac operator D (faces x vertices) and its
// adjoint D* (vertices x faces).
//
// Replaces torch.mm(Di, v.view(B*V*4, C/4)) / torch.mm(DiA, f.view(B*F*4, C/4)) at reference
// src/utils/utils_pt.py:201-203,213-215 (and the dead SparseBMMFunc branch :197-199,209-211).
// The reference's `view` makes quaternion component q of node n the q-th QUARTER of its channel vector:
//
//     Y[r, p*C4 + c] = sum_{blocks (r,j)} sum_q  B[p][q] * X[j, q*C4 + c],      C4 = C/4
//
// Mapping (C4 % 4 == 0 path):  G = 4*LPQ lanes own one block-row; lane (q, t) loads the float4
// X[j, q*C4 + 4t .. +3] -- so a warp instruction still reads whole contiguous 128 B lines of the dense
// row -- multiplies it by the four entries of block column q (one 128-bit load thanks to the rotated
// column-major block storage chosen by sn_csr32_to_bsr4_fill: slot s holds B[(q+s)%4][q]) and keeps four
// float4 partial sums, slot s feeding output component (q+s)%4.  After the row's blocks are consumed, lane q
// adds slot (4-d)%4 of the lane d quaternion-components further on (three rotating shuffles, no per-lane
// register selection) and holds output component p = q, which it stores as one coalesced float4.
// No shared memory, no atomics, deterministic summation order.
//
// Bound: HBM.  Algorithmic bytes per launch = 4(Rb+1) + 68 nb + 4 Cb C + 4 Rb C.
#include "common.cuh"

namespace sn {

template <int LPQ, bool ELU>
__global__ void __launch_bounds__(256)
bsr4_spmm_vec4_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ bcolind,
                      const float* __restrict__ bval, const float* __restrict__ X, int64_t ldx,
                      float* __restrict__ Y, int64_t ldy, int64_t n_brows, int C4) {
  constexpr int G = 4 * LPQ;        // lanes per block-row
  constexpr int RPW = kWarp / G;    // block-rows per warp
  constexpr int U = 3;              // blocks in flight (a face row of D has exactly 3 blocks)
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int q = gl / LPQ;           // quaternion component (block column) this lane reads
  const int t = gl % LPQ;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp * RPW + lane / G;
  const bool row_ok = row < n_brows;

  int start = 0, end = 0;
  if (row_ok) {
    start = __ldg(browptr + row);
    end = __ldg(browptr + row + 1);
  }
  int maxlen = end - start;
  if (RPW > 1) {
#pragma unroll
    for (int m = G; m < kWarp; m <<= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, m));
  }
  const int grp = lane - gl;  // first lane of this block-row's lane group

  for (int cb0 = 0; cb0 < C4; cb0 += 4 * LPQ) {  // chunks of the quarter-width (one for C <= 128)
    const int cb = cb0 + 4 * t;
    const bool col_ok = cb < C4;
    const int xoff = q * C4 + cb;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;
    for (int k0 = 0; k0 < maxlen; k0 += U) {
      float4 xv[U], wv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = start + k0 + u;
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        wv[u] = xv[u];
        if (k < end && col_ok) {
          const int j = __ldg(bcolind + k);
          wv[u] = ldg_f4(bval + (int64_t)k * 16 + 4 * q);  // B[(q+s)%4][q], s = 0..3
          xv[u] = ldg_f4(X + (int64_t)j * ldx + xoff);
          if (ELU) xv[u] = elu4(xv[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc0 = fma4(wv[u].x, xv[u], acc0);
        acc1 = fma4(wv[u].y, xv[u], acc1);
        acc2 = fma4(wv[u].z, xv[u], acc2);
        acc3 = fma4(wv[u].w, xv[u], acc3);
      }
    }
    // lane q: out_{p=q} = slot0(q) + slot3(q+1) + slot2(q+2) + slot1(q+3)   (lane indices mod 4 groups of LPQ)
    float4 out = acc0;
    out = add4(out, shfl_idx4(acc3, grp + (gl + LPQ) % G));
    out = add4(out, shfl_idx4(acc2, grp + (gl + 2 * LPQ) % G));
    out = add4(out, shfl_idx4(acc1, grp + (gl + 3 * LPQ) % G));
    if (row_ok && col_ok) st_stream_f4(Y + row * ldy + xoff, out);
  }
}

// Any C % 4 == 0 / any alignment: one thread per (block-row, quarter column).
template <bool ELU>
__global__ void __launch_bounds__(256)
bsr4_spmm_scalar_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ bcolind,
                        const float* __restrict__ bval, const float* __restrict__ X, int64_t ldx,
                        float* __restrict__ Y, int64_t ldy, int64_t n_brows, int C4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = idx / C4;
  const int c = (int)(idx % C4);
  if (row >= n_brows) return;
  const int start = __ldg(browptr + row), end = __ldg(browptr + row + 1);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = start; k < end; ++k) {
    const float* xr = X + (int64_t)__ldg(bcolind + k) * ldx + c;
    const float* b = bval + (int64_t)k * 16;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      float x = __ldg(xr + (int64_t)qq * C4);
      if (ELU) x = elu1(x);
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[p] = fmaf(__ldg(b + 4 * qq + ((p - qq) & 3)), x, acc[p]);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) Y[row * ldy + (int64_t)p * C4 + c] = acc[p];
}

template <int LPQ>
static int launch_vec4(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                       int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int C4, bool elu, cudaStream_t st) {
  constexpr int RPW = kWarp / (4 * LPQ);
  constexpr int kThreads = 256;
  const int64_t grid = ceil_div(n_brows, (int64_t)RPW * (kThreads / 32));
  if (grid > 0x7fffffffLL) return SN_ERR_OVERFLOW;
  if (elu)
    bsr4_spmm_vec4_kernel<LPQ, true><<<(unsigned)grid, kThreads, 0, st>>>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4);
  else
    bsr4_spmm_vec4_kernel<LPQ, false><<<(unsigned)grid, kThreads, 0, st>>>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4);
  return launch_status();
}

int launch_bsr4_stream(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X, int64_t ldx,
                       float* Y, int64_t ldy, int64_t n_brows, int64_t C, bool elu, cudaStream_t st);
int launch_bsr4_rowgroup(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                         int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, bool elu, int variant,
                         const float* G, int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2, cudaStream_t st,
                         float* stat_partial = nullptr, int* grid_out = nullptr);
int64_t rowgroup_max_grid();
// spmm_rowdirect.cu: small operators (one or two waves of threads)
bool rowdirect_applies(int64_t n_rows, int64_t C, bool three_in_flight);
int launch_bsr4_rowdirect(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X, int64_t ldx,
                          float* Y, int64_t ldy, int64_t n_brows, int64_t n_blocks, int64_t C, cudaStream_t st);

}  // namespace sn

SN_API int sn_bsr4_spmm_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval,
                            const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, int flags,
                            sn_stream_t stream) {
  using namespace sn;
  if (n_brows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (C % 4 != 0) return SN_ERR_UNSUPPORTED;  // the reference's view(.., C/4) needs it too
  if (n_brows == 0 || C == 0) return SN_OK;
  if (!browptr || !bcolind || !bval || !X || !Y || ldx < C || ldy < C) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool elu = (flags & SN_SPMM_ELU_INPUT) != 0;
  const int C4 = (int)(C / 4);
  const bool vec_ok = (C4 % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                      aligned16(bval);
  if (!vec_ok) {
    const int64_t grid = ceil_div(n_brows * C4, 256);
    if (grid > 0x7fffffffLL) return SN_ERR_OVERFLOW;
    if (elu)
      bsr4_spmm_scalar_kernel<true><<<(unsigned)grid, 256, 0, st>>>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4);
    else
      bsr4_spmm_scalar_kernel<false><<<(unsigned)grid, 256, 0, st>>>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4);
    return launch_status();
  }
  if (flags & SN_SPMM_SMEM_STREAM) {       // cp.async streaming kernel: C = 128 / 256 / 512
    const int rc = launch_bsr4_stream(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, elu, st);
    if (rc != SN_ERR_UNSUPPORTED) return rc;
  } else if (!(flags & SN_SPMM_DIRECT_GATHER)) {
    const int variant = (flags >> 8) & 15, hint = (flags >> 12) & 15;
    // small operators (a single mesh, a mesh_mnist batch): the latency-oriented kernel, bit-identical results
    if (!elu && variant != 7 && (variant == 6 || (variant == 0 && rowdirect_applies(n_brows, C, hint >= 1 && hint <= 3)))) {
      const int rc = launch_bsr4_rowdirect(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows,
                                           hint >= 1 && hint <= 3 ? 3 * n_brows : -1, C, st);
      if (rc != SN_ERR_UNSUPPORTED) return rc;
    }
    // row-group kernel: C = 32 ... 512 (powers of two).  Long rows (D*: one block per incident face, ~6) keep two gathers
    // in flight per row group, landing in shared memory (variant 9: D* 69.4 -> 63.5 us at the cfg3 size, C = 128;
    // 111.5 -> 100.7 us at C = 256); for rows of three blocks (D) the extra shared-memory traffic loses (58 -> 64 us)
    const int rg_variant = (variant == 6 || variant == 7) ? 0 : (variant == 0 && hint >= 5 ? 9 : variant);
    const int rc = launch_bsr4_rowgroup(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, elu, rg_variant, nullptr, 0,
                                        nullptr, 0, nullptr, 0, st);
    if (rc != SN_ERR_UNSUPPORTED) return rc;
  }
  const int v = C4 / 4;  // float4 columns per quaternion component
  if (v <= 1) return launch_vec4<1>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4, elu, st);
  if (v <= 2) return launch_vec4<2>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4, elu, st);
  if (v <= 4) return launch_vec4<4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4, elu, st);
  return launch_vec4<8>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C4, elu, st);
}

// Y = (S X + G) .* elu'(A) + G2: the SpMM of a backward pass with the activation derivative (and the gradient of the
// un-gathered half, G) applied in the store path -- replaces op.T.apply + sn_elu_bwd_f32.  Row-group kernel only.
SN_API int sn_bsr4_spmm_epilogue_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                                     int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, const float* G,
                                     int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2, int flags,
                                     sn_stream_t stream) {
  using namespace sn;
  if (n_brows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (n_brows == 0 || C == 0) return SN_OK;
  if (!browptr || !bcolind || !bval || !X || !Y || ldx < C || ldy < C || (G && ldg < C) || (A && lda < C) ||
      (G2 && ldg2 < C))
    return SN_ERR_ARG;
  if (flags & (SN_SPMM_DIRECT_GATHER | SN_SPMM_SMEM_STREAM | SN_SPMM_ELU_INPUT)) return SN_ERR_UNSUPPORTED;
  if (C % 16 || ldx % 4 || ldy % 4 || ldg % 4 || lda % 4 || ldg2 % 4 || !aligned16(G2) || !aligned16(X) || !aligned16(Y) || !aligned16(bval) ||
      !aligned16(G) || !aligned16(A))
    return SN_ERR_UNSUPPORTED;
  return launch_bsr4_rowgroup(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, false, (flags >> 8) & 15, G, ldg, A, lda, G2, ldg2,
                              (cudaStream_t)stream);
}

// Y = S X and the per-column mean / biased variance of Y over all n_brows rows from the same pass (the BatchNorm
// statistics of the right half of the stage's concat buffer, utils_pt.py:98 after :203): every row's values are added to
// per-(warp, row group) accumulators in shared memory where the row is stored; per-CTA partials, fixed-order fp64 final
// reduction.  Row-group kernel only (C in {32, ..., 512}); SN_ERR_UNSUPPORTED otherwise (nothing launched).
SN_API size_t sn_spmm_stats_ws_bytes(int64_t C) {
  return C <= 0 ? 0 : (size_t)sn::rowgroup_max_grid() * 2 * (size_t)C * sizeof(float);
}

SN_API int sn_bsr4_spmm_stats_f32(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                                  int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, float* mean,
                                  float* var_biased, int flags, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (n_brows <= 0 || C <= 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (!browptr || !bcolind || !bval || !X || !Y || !mean || !var_biased || ldx < C || ldy < C) return SN_ERR_ARG;
  if (flags & (SN_SPMM_DIRECT_GATHER | SN_SPMM_SMEM_STREAM | SN_SPMM_ELU_INPUT)) return SN_ERR_UNSUPPORTED;
  if (C % 16 || ldx % 4 || ldy % 4 || !aligned16(X) || !aligned16(Y) || !aligned16(bval)) return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_spmm_stats_ws_bytes(C) || !aligned16(ws)) return SN_ERR_WORKSPACE;
  int grid = 0;
  const int variant = (flags >> 8) & 15, hint = (flags >> 12) & 15;
  const int rc = launch_bsr4_rowgroup(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, false,
                                      variant == 0 && hint >= 5 ? 9 : variant, nullptr, 0,
                                      nullptr, 0, nullptr, 0, (cudaStream_t)stream, reinterpret_cast<float*>(ws), &grid);
  if (rc != SN_OK) return rc;
  return launch_colstats_final(reinterpret_cast<const float*>(ws), grid, n_brows, (int)C, nullptr, mean, var_biased,
                               (cudaStream_t)stream);
}
