// bn_fold.cu -- the O(C^2) glue of the fused dense stage, one launch each way (replaces ~25 tiny elementwise launches).
//
// forward : BatchNorm (reference utils_pt.py:84,98) folded into the Linear (utils_pt.py:89,99):
//             s = gamma * rstd, t = beta - mean * s, W' = W diag(s), b' = b + W t, running statistics update
// backward: from G = dY^T Z and sdY = colsum(dY) (see fused.py):
//             dW = G diag(s) + sdY (x) t, db = sdY, dbeta = W^T sdY, dgamma = rstd (sum_c W.*G - mean dbeta),
//             p = -s rstd dgamma / n, q = -s dbeta / n - p mean, and the transposed, scaled weights (W diag(s))^T
#include "common.cuh"

namespace sn {

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

constexpr int kFoldCols = 32;      // columns per CTA
constexpr int kFoldRowGroups = 32; // row groups per CTA of the backward kernel (1024 threads: 4 rows per thread at N = 128)

// ONE launch for the forward fold.  The launch sits between two full-device kernels, so what matters is its critical
// path, not its throughput: one CTA per kFwdRows output rows, thread k owns column k (and k + 256, ...) of those rows, so
// every global read of a thread -- statistics, affine parameters, running statistics, its kFwdRows weights -- is issued
// before the first dependent instruction; b'[n] = b[n] + sum_k W[n,k] t[k] is one block reduction (fixed order).  Every
// CTA recomputes s_k / t_k (a few flops) instead of waiting for a first kernel to publish them; CTA 0 publishes s, t,
// rstd and updates the running statistics.  (History: two dependent launches, 22 us; one warp per row, 11 us -- the
// publishing warp walked its eight columns as eight dependent load -> store round trips; now ~5 us.)
constexpr int kFwdRows = 4;
constexpr int kFwdThreads = 256;
__global__ void __launch_bounds__(kFwdThreads)
bn_fold_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ W, const float* __restrict__ b, int N, int K,
                   float eps, float* __restrict__ Wf, float* __restrict__ bf, float* __restrict__ s_out,
                   float* __restrict__ t_out, float* __restrict__ rstd_out, float* __restrict__ running_mean,
                   float* __restrict__ running_var, float momentum, float unbias, float* __restrict__ Wf_hi,
                   float* __restrict__ Wf_lo, long long* __restrict__ batches_tracked) {
  __shared__ float red[kFwdThreads / 32][kFwdRows];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * kFwdRows;
  const bool publish = blockIdx.x == 0;
  const bool running = publish && running_mean != nullptr;
  if (publish && threadIdx.x == 0 && batches_tracked) *batches_tracked += 1;   // nn.BatchNorm's num_batches_tracked
  float acc[kFwdRows];
#pragma unroll
  for (int j = 0; j < kFwdRows; ++j) acc[j] = 0.f;
#pragma unroll 2
  for (int k = tid; k < K; k += kFwdThreads) {
    const float m = mean[k], v = var[k], ga = gamma[k], be = beta[k];
    float w[kFwdRows];
#pragma unroll
    for (int j = 0; j < kFwdRows; ++j) w[j] = n0 + j < N ? W[(size_t)(n0 + j) * K + k] : 0.f;
    const float rm = running ? running_mean[k] : 0.f, rv = running ? running_var[k] : 0.f;
    const float rstd = rsqrtf(v + eps);
    const float sk = ga * rstd;
    const float tk = be - m * sk;
#pragma unroll
    for (int j = 0; j < kFwdRows; ++j) {
      if (n0 + j < N) {
        const float ws = w[j] * sk;
        Wf[(size_t)(n0 + j) * K + k] = ws;
        if (Wf_hi) {                       // the 3xTF32 GEMM's pre-split B operand (sn_gemm_tf32_presplit_f32)
          const float h = tf32_round(ws);
          Wf_hi[(size_t)(n0 + j) * K + k] = h;
          Wf_lo[(size_t)(n0 + j) * K + k] = tf32_round(ws - h);
        }
      }
      acc[j] = fmaf(w[j], tk, acc[j]);
    }
    if (publish) {
      s_out[k] = sk;
      t_out[k] = tk;
      rstd_out[k] = rstd;
      if (running) {
        running_mean[k] = (1.f - momentum) * rm + momentum * m;
        running_var[k] = (1.f - momentum) * rv + momentum * v * unbias;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < kFwdRows; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kFwdRows; ++j) red[warp][j] = acc[j];
  }
  __syncthreads();
  if (tid < kFwdRows && n0 + tid < N) {
    float a = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kFwdThreads / 32; ++w8) a += red[w8][tid];       // fixed order
    bf[n0 + tid] = b[n0 + tid] + a;
  }
}

constexpr int kFoldMaxN = 256;     // rows of W a column slab stages for the transposed store
__global__ void __launch_bounds__(kFoldCols * kFoldRowGroups)
bn_fold_bwd_kernel(const float* __restrict__ G, const float* __restrict__ sdY, const float* __restrict__ W,
                   const float* __restrict__ s, const float* __restrict__ t, const float* __restrict__ rstd,
                   const float* __restrict__ mean, int N, int K, float inv_rows, int training, float* __restrict__ dW,
                   float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ p_out,
                   float* __restrict__ q_out, float* __restrict__ WsT, float* __restrict__ WsT_hi,
                   float* __restrict__ WsT_lo) {
  __shared__ float red[2][kFoldRowGroups][kFoldCols];
  __shared__ float ws_s[kFoldMaxN][kFoldCols + 1];      // W[n][k] s[k] of this slab: written row-wise, stored transposed
  const int c = threadIdx.x % kFoldCols, rg = threadIdx.x / kFoldCols;
  const int k = blockIdx.x * kFoldCols + c;
  float dbeta_p = 0.f, wg_p = 0.f;
  if (k < K) {
    const float sk = s[k], tk = t[k];
#pragma unroll 4
    for (int n = rg; n < N; n += kFoldRowGroups) {
      const float w = W[(size_t)n * K + k], g = G[(size_t)n * K + k], d = sdY[n];
      dW[(size_t)n * K + k] = fmaf(g, sk, d * tk);
      dbeta_p = fmaf(w, d, dbeta_p);
      wg_p = fmaf(w, g, wg_p);
      ws_s[n][c] = w * sk;
    }
  }
  red[0][rg][c] = dbeta_p;
  red[1][rg][c] = wg_p;
  __syncthreads();
  // (W diag(s))^T [K x N]: consecutive threads write consecutive n of one row k (the direct store had a stride of N floats
  // between the lanes of a warp: 32 sectors per instruction, three arrays)
  const int k0 = blockIdx.x * kFoldCols;
  for (int i = threadIdx.x; i < kFoldCols * N; i += kFoldCols * kFoldRowGroups) {
    const int kl = i / N, n = i - kl * N;
    if (k0 + kl < K) {
      const float ws = ws_s[n][kl];
      WsT[(size_t)(k0 + kl) * N + n] = ws;
      if (WsT_hi) {
        const float h = tf32_round(ws);
        WsT_hi[(size_t)(k0 + kl) * N + n] = h;
        WsT_lo[(size_t)(k0 + kl) * N + n] = tf32_round(ws - h);
      }
    }
  }
  if (rg == 0 && k < K) {
    float dbeta_k = 0.f, wg = 0.f;
#pragma unroll
    for (int g = 0; g < kFoldRowGroups; ++g) {   // fixed order
      dbeta_k += red[0][g][c];
      wg += red[1][g][c];
    }
    const float sk = s[k];
    const float dgamma_k = rstd[k] * (wg - mean[k] * dbeta_k);
    dgamma[k] = dgamma_k;
    dbeta[k] = dbeta_k;
    const float pk = training ? -sk * rstd[k] * dgamma_k * inv_rows : 0.f;
    p_out[k] = pk;
    q_out[k] = training ? -sk * dbeta_k * inv_rows - pk * mean[k] : 0.f;
  }
  if (blockIdx.x == 0)
    for (int n = threadIdx.x; n < N; n += blockDim.x) db[n] = sdY[n];
}

}  // namespace sn

SN_API int sn_bn_fold_fwd_f32(const float* mean, const float* var, const float* gamma, const float* beta, const float* W,
                              const float* b, int64_t N, int64_t K, float eps, float* Wf, float* bf, float* s, float* t,
                              float* rstd, float* running_mean, float* running_var, float momentum, int64_t rows,
                              float* Wf_hi, float* Wf_lo, int64_t* num_batches_tracked, sn_stream_t stream) {
  using namespace sn;
  if (N <= 0 || K <= 0 || !mean || !var || !gamma || !beta || !W || !b || !Wf || !bf || !s || !t || !rstd) return SN_ERR_ARG;
  if ((Wf_hi == nullptr) != (Wf_lo == nullptr)) return SN_ERR_ARG;
  const float unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  bn_fold_fwd_kernel<<<(unsigned)ceil_div(N, kFwdRows), kFwdThreads, 0, st>>>(mean, var, gamma, beta, W, b, (int)N, (int)K, eps,
                                                                              Wf, bf, s, t, rstd, running_mean, running_var,
                                                                              momentum, unbias, Wf_hi, Wf_lo,
                                                                              reinterpret_cast<long long*>(num_batches_tracked));
  return launch_status();
}

SN_API int sn_bn_fold_bwd_f32(const float* G, const float* sdY, const float* W, const float* s, const float* t,
                              const float* rstd, const float* mean, int64_t N, int64_t K, int64_t rows, int training,
                              float* dW, float* db, float* dgamma, float* dbeta, float* p, float* q, float* WsT,
                              float* WsT_hi, float* WsT_lo, sn_stream_t stream) {
  using namespace sn;
  if (N <= 0 || K <= 0 || rows <= 0 || !G || !sdY || !W || !s || !t || !rstd || !mean || !dW || !db || !dgamma || !dbeta ||
      !p || !q || !WsT || (WsT_hi == nullptr) != (WsT_lo == nullptr))
    return SN_ERR_ARG;
  if (N > kFoldMaxN) return SN_ERR_UNSUPPORTED;
  bn_fold_bwd_kernel<<<(unsigned)ceil_div(K, kFoldCols), kFoldCols * kFoldRowGroups, 0, (cudaStream_t)stream>>>(
      G, sdY, W, s, t, rstd, mean, (int)N, (int)K, (float)(1.0 / (double)rows), training, dW, db, dgamma, dbeta, p, q, WsT,
      WsT_hi, WsT_lo);
  return launch_status();
}
