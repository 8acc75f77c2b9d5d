// colstats.cu -- per-column batch statistics of a stage's concat buffer Z [rows x C]: the reduction that
// training-mode BatchNorm needs in front of the Linear (reference src/utils/utils_pt.py:84,98 -- nn.BatchNorm1d on
// x.transpose(1,2), i.e. statistics over ALL B*N rows, padded rows included).
//
// One HBM pass (bound: HBM, 4 * rows * C bytes).  Deterministic: every CTA writes its fp32 partial sums (sum, sum of
// squares) to a workspace row; a second tiny kernel adds the partials in a fixed order in fp64 and emits
// mean / biased variance.  BatchNorm itself is never applied as a pass: it is folded into the Linear weights.
#include "common.cuh"

namespace sn {

constexpr int kStatThreads = 256;

// thread (rg, cv): row group rg = tid / CV walks rows rg, rg + RG, ...; cv owns float4 column cv (cv < C/4)
__global__ void __launch_bounds__(kStatThreads)
colstats_partial_kernel(const float* __restrict__ X, int64_t ldx, int64_t rows, int C, float* __restrict__ partial) {
  extern __shared__ float red[];                      // [RG][2][C]
  const int CV = C / 4;
  const int RG = kStatThreads / CV;                   // row groups per CTA (C <= 1024, C % 4 == 0, CV divides 256)
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  // shifted sums: accumulate x - K with K = row 0 of the column, so that constant columns give variance exactly 0
  // and E[(x-K)^2] - E[x-K]^2 does not cancel catastrophically when |mean| >> std
  const float4 K = (cv < CV) ? __ldg(reinterpret_cast<const float4*>(X) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < RG) {
    const int64_t stride = (int64_t)gridDim.x * RG;
    int64_t r = (int64_t)blockIdx.x * RG + rg;
    // 4 rows in flight per thread
    for (; r + 3 * stride < rows; r += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(X + (r + u * stride) * ldx) + cv);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u].x -= K.x; v[u].y -= K.y; v[u].z -= K.z; v[u].w -= K.w;
        s = add4(s, v[u]);
        q.x = fmaf(v[u].x, v[u].x, q.x); q.y = fmaf(v[u].y, v[u].y, q.y);
        q.z = fmaf(v[u].z, v[u].z, q.z); q.w = fmaf(v[u].w, v[u].w, q.w);
      }
    }
    for (; r < rows; r += stride) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(X + r * ldx) + cv);
      v.x -= K.x; v.y -= K.y; v.z -= K.z; v.w -= K.w;
      s = add4(s, v);
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    float* base = red + (size_t)rg * 2 * C;
    *reinterpret_cast<float4*>(base + 4 * cv) = s;
    *reinterpret_cast<float4*>(base + C + 4 * cv) = q;
  }
  __syncthreads();
  // fixed-order reduction over the row groups; thread t < 2C handles one (stat, column)
  for (int i = threadIdx.x; i < 2 * C; i += kStatThreads) {
    float a = 0.f;
    for (int g = 0; g < RG; ++g) a += red[(size_t)g * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// fixed-order fp64 reduction of the per-CTA partials: CTA = 8 columns x 128 partial groups (1024 threads), C / 8 CTAs; group g
// adds partials g, g + 128, ... (at most 8 rows for the largest producer grid: all loads of a thread are independent and in
// flight together -- the launch is pure latency; the first version, 32 columns x 32 groups on C / 32 CTAs, walked 28 rows
// per thread in dependent batches and took 9-11 us) and the 128 group sums are added in order by the column's first thread.
constexpr int kFinCols = 8, kFinGroups = 128;
__global__ void __launch_bounds__(kFinCols * kFinGroups)
colstats_final_kernel(const float* __restrict__ partial, int n_partials, int64_t rows, int C, const float* __restrict__ X,
                      float* __restrict__ mean, float* __restrict__ var) {
  __shared__ double red[2][kFinGroups][kFinCols + 1];
  const int lc = threadIdx.x % kFinCols, g = threadIdx.x / kFinCols;
  const int c = blockIdx.x * kFinCols + lc;
  double s = 0.0, q = 0.0;
  if (c < C) {
    int i = g;
    for (; i + 3 * kFinGroups < n_partials; i += 4 * kFinGroups) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(partial + (size_t)(i + u * kFinGroups) * 2 * C + c);
        b[u] = __ldg(partial + (size_t)(i + u * kFinGroups) * 2 * C + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += (double)a[u];
        q += (double)b[u];
      }
    }
    for (; i < n_partials; i += kFinGroups) {
      s += (double)__ldg(partial + (size_t)i * 2 * C + c);
      q += (double)__ldg(partial + (size_t)i * 2 * C + C + c);
    }
  }
  red[0][g][lc] = s;
  red[1][g][lc] = q;
  __syncthreads();
  if (g == 0 && c < C) {
    s = 0.0;
    q = 0.0;
    for (int k = 0; k < kFinGroups; ++k) {
      s += red[0][k][lc];
      q += red[1][k][lc];
    }
    const double m = s / (double)rows;            // mean of the shifted values
    double v = q / (double)rows - m * m;
    if (v < 0.0) v = 0.0;
    mean[c] = (float)(m + (X ? (double)X[c] : 0.0));      // X: the shift the partial sums were taken around (null: none)
    var[c] = (float)v;
  }
}

// Activation pass with the statistics fused in: Y = elu(X) AND the per-CTA partial sums of Y's columns (same partial
// layout and final kernel as above), so the left half of a stage's concat buffer never needs a separate statistics
// pass.  gridDim.x * RG row slots are walked with a fixed stride, every thread keeps one float4 column.
__global__ void __launch_bounds__(kStatThreads)
elu_colstats_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int64_t rows, int C,
                    float* __restrict__ partial) {
  extern __shared__ float red[];                      // [RG][2][C]
  const int CV = C / 4;
  const int RG = kStatThreads / CV;
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  const float4 K = elu4(__ldg(reinterpret_cast<const float4*>(X) + cv));     // shift = activated row 0
  const int64_t stride = (int64_t)gridDim.x * RG;
  int64_t r = (int64_t)blockIdx.x * RG + rg;
  for (; r + stride < rows; r += 2 * stride) {          // two rows in flight per thread
    float4 v0 = __ldcs(reinterpret_cast<const float4*>(X + r * ldx) + cv);
    float4 v1 = __ldcs(reinterpret_cast<const float4*>(X + (r + stride) * ldx) + cv);
    v0 = elu4(v0);
    v1 = elu4(v1);
    *(reinterpret_cast<float4*>(Y + r * ldy) + cv) = v0;
    *(reinterpret_cast<float4*>(Y + (r + stride) * ldy) + cv) = v1;
    v0.x -= K.x; v0.y -= K.y; v0.z -= K.z; v0.w -= K.w;
    v1.x -= K.x; v1.y -= K.y; v1.z -= K.z; v1.w -= K.w;
    s = add4(s, add4(v0, v1));
    q.x = fmaf(v0.x, v0.x, fmaf(v1.x, v1.x, q.x)); q.y = fmaf(v0.y, v0.y, fmaf(v1.y, v1.y, q.y));
    q.z = fmaf(v0.z, v0.z, fmaf(v1.z, v1.z, q.z)); q.w = fmaf(v0.w, v0.w, fmaf(v1.w, v1.w, q.w));
  }
  for (; r < rows; r += stride) {
    float4 v = elu4(__ldcs(reinterpret_cast<const float4*>(X + r * ldx) + cv));
    *(reinterpret_cast<float4*>(Y + r * ldy) + cv) = v;
    v.x -= K.x; v.y -= K.y; v.z -= K.z; v.w -= K.w;
    s = add4(s, v);
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  float* base = red + (size_t)rg * 2 * C;
  *reinterpret_cast<float4*>(base + 4 * cv) = s;
  *reinterpret_cast<float4*>(base + C + 4 * cv) = q;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kStatThreads) {
    float a = 0.f;
    for (int g = 0; g < RG; ++g) a += red[(size_t)g * 2 * C + i];
    partial[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}

// Final reduction for producers that emit the partial sums themselves (the GEMM's ACT epilogue, the row-group SpMM's
// statistics store path): partial [n_partials][2][C] (sum | sum of squares of x - shift), shift [C] or null.
int launch_colstats_final(const float* partial, int n_partials, int64_t rows, int C, const float* shift, float* mean,
                          float* var, cudaStream_t st) {
  colstats_final_kernel<<<(unsigned)ceil_div(C, kFinCols), kFinCols * kFinGroups, 0, st>>>(partial, n_partials, rows, C, shift, mean, var);
  return launch_status();
}

static int stat_grid() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 6;   // 6 CTAs x 256 threads per SM, one partial row per CTA
}

}  // namespace sn

SN_API size_t sn_colstats_ws_bytes(int64_t C) {
  return C <= 0 ? 0 : (size_t)sn::stat_grid() * 2 * (size_t)C * sizeof(float);
}

SN_API int sn_colstats_f32(const float* X, int64_t ldx, int64_t rows, int64_t C, float* mean, float* var_biased,
                           void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows <= 0 || C <= 0 || !X || !mean || !var_biased || ldx < C) return SN_ERR_ARG;
  if (C % 4 || C > 1024 || (kStatThreads % (C / 4)) || ldx % 4 || !aligned16(X)) return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_colstats_ws_bytes(C)) return SN_ERR_WORKSPACE;
  int grid = stat_grid();
  const int RG = kStatThreads / (int)(C / 4);
  if ((int64_t)grid * RG > rows) grid = (int)ceil_div(rows, RG);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)RG * 2 * C * sizeof(float);
  if (smem > 48 * 1024) return SN_ERR_UNSUPPORTED;
  colstats_partial_kernel<<<grid, kStatThreads, smem, st>>>(X, ldx, rows, (int)C, (float*)ws);
  colstats_final_kernel<<<(unsigned)ceil_div(C, kFinCols), kFinCols * kFinGroups, 0, st>>>((const float*)ws, grid, rows, (int)C, X, mean, var_biased);
  return launch_status();
}

SN_API int sn_elu_colstats_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t rows, int64_t C, float* mean,
                               float* var_biased, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows <= 0 || C <= 0 || !X || !Y || !mean || !var_biased || ldx < C || ldy < C) return SN_ERR_ARG;
  if (C % 4 || C > 1024 || (kStatThreads % (C / 4)) || ldx % 4 || ldy % 4 || !aligned16(X) || !aligned16(Y))
    return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_colstats_ws_bytes(C)) return SN_ERR_WORKSPACE;
  int grid = stat_grid();
  const int RG = kStatThreads / (int)(C / 4);
  if ((int64_t)grid * RG > rows) grid = (int)ceil_div(rows, RG);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)RG * 2 * C * sizeof(float);
  if (smem > 48 * 1024) return SN_ERR_UNSUPPORTED;
  elu_colstats_kernel<<<grid, kStatThreads, smem, st>>>(X, ldx, Y, ldy, rows, (int)C, (float*)ws);
  // the shift used by the partial sums is the activated row 0, which the kernel above has just written to Y
  colstats_final_kernel<<<(unsigned)ceil_div(C, kFinCols), kFinCols * kFinGroups, 0, st>>>((const float*)ws, grid, rows, (int)C, Y, mean, var_biased);
  return launch_status();
}
