// avg_stage.cu -- the O(rows * C) prologue and the O(C^2 + B C^2) glue of one AvgResNet2 stage
// (reference src/utils/utils_pt.py:222-243:  x -> elu -> cat[x, global_average(x, mask)] -> BatchNorm -> Linear).
//
// Round 1 ran this glue as ~17 launches per stage forward and ~15 backward (ATen reductions, cat, addmm, pow, ...), 137 us
// of launch-bound work around a 51 us GEMM (profiles/r2_launches_step_summary.json).  Here a stage is
//
//   forward   avg_pre_kernel        a = elu(x); per-(mesh, slice) partial sums: sum / sum of squares of a (shifted) and
//                                   the masked sum of a                                           ONE pass over x
//             avg_stats_kernel      partials -> left-half statistics, per-mesh averages avg [B x C], right-half statistics
//                                   (the broadcast half of the concat is constant per mesh: its BatchNorm statistics are
//                                   the equal-weight statistics of the B averages)
//             avg_fold_fwd_kernel   BatchNorm folded into the Linear (as bn_fold_fwd_kernel) + the per-mesh bias
//                                   u[b] = b' + W'_R avg_b that replaces the broadcast half of the GEMM
//             (GEMM: sn_gemm_tf32_presplit_f32 with K = C, group_bias = u)
//   backward  avg_fold_bwd_kernel   G_R = SdY^T avg, the folded BatchNorm backward on G = [G_L | G_R] (as bn_fold_bwd_kernel)
//                                   and the gradient of the per-mesh averages gb [B x C], one launch
//
// All reductions run in a fixed order (bit-reproducible).
#include "common.cuh"

namespace sn {

namespace {

constexpr int kSlices = 8;          // row slices per mesh in avg_pre_kernel
constexpr int kPreThreads = 256;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// grid (n_seg, kSlices).  Thread (rg, cv): row group rg walks the slice's rows rg, rg + RG, ...; cv owns float4 column cv.
// partial[(seg * kSlices + slice)][3][C] = { sum(a - K), sum((a - K)^2), sum(w a) },  K = elu(X[0, :]) (the shift that
// keeps E[(a-K)^2] - E[a-K]^2 from cancelling, see colstats.cu).
__global__ void __launch_bounds__(kPreThreads)
avg_pre_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, int rows_per_seg, int C,
               float* __restrict__ A, int64_t lda, float* __restrict__ partial) {
  extern __shared__ float red[];               // [RG][3][C]
  const int CV = C / 4, RG = kPreThreads / CV;
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  const int slice = (rows_per_seg + kSlices - 1) / kSlices;
  const int rb = blockIdx.y * slice, re = min(rb + slice, rows_per_seg);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_seg;
  const float4 K = elu4(__ldg(reinterpret_cast<const float4*>(X) + cv));
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s, m = s;
  int r = rb + rg;
  for (; r + RG < re; r += 2 * RG) {           // two rows in flight per thread (four: no faster, 29.8 vs 29.4 us)
    const float4 x0 = __ldcs(reinterpret_cast<const float4*>(X + (r0 + r) * ldx) + cv);
    const float4 x1 = __ldcs(reinterpret_cast<const float4*>(X + (r0 + r + RG) * ldx) + cv);
    const float w0 = w ? __ldg(w + r0 + r) : 1.f, w1 = w ? __ldg(w + r0 + r + RG) : 1.f;
    const float4 a0 = elu4(x0), a1 = elu4(x1);
    *(reinterpret_cast<float4*>(A + (r0 + r) * lda) + cv) = a0;
    *(reinterpret_cast<float4*>(A + (r0 + r + RG) * lda) + cv) = a1;
    m = fma4(w0, a0, m);
    m = fma4(w1, a1, m);
    const float4 d0 = make_float4(a0.x - K.x, a0.y - K.y, a0.z - K.z, a0.w - K.w);
    const float4 d1 = make_float4(a1.x - K.x, a1.y - K.y, a1.z - K.z, a1.w - K.w);
    s = add4(s, add4(d0, d1));
    q.x = fmaf(d0.x, d0.x, fmaf(d1.x, d1.x, q.x)); q.y = fmaf(d0.y, d0.y, fmaf(d1.y, d1.y, q.y));
    q.z = fmaf(d0.z, d0.z, fmaf(d1.z, d1.z, q.z)); q.w = fmaf(d0.w, d0.w, fmaf(d1.w, d1.w, q.w));
  }
  for (; r < re; r += RG) {
    const float4 a0 = elu4(__ldcs(reinterpret_cast<const float4*>(X + (r0 + r) * ldx) + cv));
    const float w0 = w ? __ldg(w + r0 + r) : 1.f;
    *(reinterpret_cast<float4*>(A + (r0 + r) * lda) + cv) = a0;
    m = fma4(w0, a0, m);
    const float4 d0 = make_float4(a0.x - K.x, a0.y - K.y, a0.z - K.z, a0.w - K.w);
    s = add4(s, d0);
    q.x = fmaf(d0.x, d0.x, q.x); q.y = fmaf(d0.y, d0.y, q.y); q.z = fmaf(d0.z, d0.z, q.z); q.w = fmaf(d0.w, d0.w, q.w);
  }
  float* base = red + (size_t)rg * 3 * C;
  *reinterpret_cast<float4*>(base + 4 * cv) = s;
  *reinterpret_cast<float4*>(base + C + 4 * cv) = q;
  *reinterpret_cast<float4*>(base + 2 * C + 4 * cv) = m;
  __syncthreads();
  float* out = partial + ((size_t)blockIdx.x * kSlices + blockIdx.y) * 3 * C;
  for (int i = threadIdx.x; i < 3 * C; i += kPreThreads) {
    float a = 0.f;
    for (int g = 0; g < RG; ++g) a += red[(size_t)g * 3 * C + i];      // fixed order
    out[i] = a;
  }
}

// grid = ceil(C / 32), 1024 threads = 32 columns x 32 groups.  Group g reduces the partial rows g, g + 32, ... (fp64) and the
// meshes g, g + 32, ...; the 32 group results are added in order.
__global__ void __launch_bounds__(1024)
avg_stats_kernel(const float* __restrict__ partial, int n_seg, int64_t rows, int C, const float* __restrict__ A0,
                 const float* __restrict__ inv_cnt, float* __restrict__ mean, float* __restrict__ var,
                 float* __restrict__ avg) {
  __shared__ double red[3][32][33];
  __shared__ double mean_r_s[32];
  const int lc = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lc;
  const bool live = c < C;
  double s = 0.0, q = 0.0, ma = 0.0;
  if (live) {
    const int n_part = n_seg * kSlices;
    int i = g;
    for (; i + 3 * 32 < n_part; i += 4 * 32) {   // eight independent loads in flight, same (fixed) order of additions
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = __ldg(partial + (size_t)(i + u * 32) * 3 * C + c);
        b[u] = __ldg(partial + (size_t)(i + u * 32) * 3 * C + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s += (double)a[u];
        q += (double)b[u];
      }
    }
    for (; i < n_part; i += 32) {
      s += (double)__ldg(partial + (size_t)i * 3 * C + c);
      q += (double)__ldg(partial + (size_t)i * 3 * C + C + c);
    }
    for (int b = g; b < n_seg; b += 32) {
      float m = 0.f;
#pragma unroll
      for (int j = 0; j < kSlices; ++j) m += __ldg(partial + ((size_t)b * kSlices + j) * 3 * C + 2 * C + c);
      const float a = m * __ldg(inv_cnt + b);
      avg[(size_t)b * C + c] = a;
      ma += (double)a;
    }
  }
  red[0][g][lc] = s;
  red[1][g][lc] = q;
  red[2][g][lc] = ma;
  __syncthreads();
  if (g == 0 && live) {
    s = q = ma = 0.0;
    for (int k = 0; k < 32; ++k) {
      s += red[0][k][lc];
      q += red[1][k][lc];
      ma += red[2][k][lc];
    }
    const double m = s / (double)rows;              // mean of the shifted values
    double v = q / (double)rows - m * m;
    if (v < 0.0) v = 0.0;
    mean[c] = (float)(m + (double)A0[c]);
    var[c] = (float)v;
    const double mr = ma / (double)n_seg;
    mean[C + c] = (float)mr;
    mean_r_s[lc] = mr;
  }
  __syncthreads();
  // right half, second pass: biased variance of the per-mesh averages around their mean (they are close together: two-pass)
  double dv = 0.0;
  if (live) {
    const double mr = mean_r_s[lc];
    for (int b = g; b < n_seg; b += 32) {
      const double d = (double)avg[(size_t)b * C + c] - mr;   // written by this very thread above
      dv += d * d;
    }
  }
  red[0][g][lc] = dv;
  __syncthreads();
  if (g == 0 && live) {
    dv = 0.0;
    for (int k = 0; k < 32; ++k) dv += red[0][k][lc];
    var[C + c] = (float)(dv / (double)n_seg);
  }
}

// bn_fold_fwd_kernel (bn_fold.cu) for K = 2C plus the per-mesh bias:
//   s = gamma rstd, t = beta - mean s, W' = W diag(s) (tf32 hi / lo split), b'[n] = b[n] + sum_k W[n,k] t[k],
//   u[b][n] = b'[n] + sum_c avg[b][c] W'[n][C + c]
// A launch of this kernel sits between two full-device kernels with nothing to overlap it: its cost is its critical path.
// (The first version -- one warp per output row, shuffle reductions, averages staged by a rolled load/store loop --
// took 21 us for 1 M multiply-adds: 8 dependent L2 round trips of staging and 16 rounds of five dependent shuffles.)
// Now: one CTA per kFoldRows output rows; EVERY global read of the CTA is issued up front (the per-mesh averages by
// cp.async, no registers held), thread k folds column k of the CTA's rows, one block reduction gives b', and thread
// (mesh, row) computes one u[mesh][row] as a shared-memory dot product with four independent accumulators.
constexpr int kMeshChunk = 64;
constexpr int kFoldRows = 4;
constexpr int kFoldThreads = 256;
__global__ void __launch_bounds__(kFoldThreads)
avg_fold_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ W, const float* __restrict__ b, int N, int C,
                    float eps, float* __restrict__ Wf_hi, float* __restrict__ Wf_lo, float* __restrict__ s_out,
                    float* __restrict__ t_out, float* __restrict__ rstd_out, float* __restrict__ running_mean,
                    float* __restrict__ running_var, float momentum, float unbias, const float* __restrict__ avg, int n_seg,
                    float* __restrict__ u, long long* __restrict__ batches_tracked) {
  extern __shared__ __align__(16) float fold_sm[];
  const int AS = C + 16, WS = C + 4;                      // row strides: conflict-free 16-byte reads (see the dot loop)
  float* avg_s = fold_sm;                                 // [kMeshChunk][AS]
  float* w_s = avg_s + kMeshChunk * AS;                   // [kFoldRows][WS]   W'[n][C + c] of this CTA's rows
  float* red = w_s + kFoldRows * WS;                      // [8 warps][kFoldRows]
  float* bf_s = red + 8 * kFoldRows;                      // [kFoldRows]
  const int K = 2 * C, CV = C / 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * kFoldRows;
  const bool publish = blockIdx.x == 0;
  const bool running = publish && running_mean != nullptr;
  if (publish && threadIdx.x == 0 && batches_tracked) *batches_tracked += 1;   // nn.BatchNorm's num_batches_tracked
  auto stage_avg = [&](int b0, int nb) {                  // meshes b0 .. b0 + nb - 1 -> avg_s, 16 bytes per cp.async
    for (int i = tid; i < nb * CV; i += kFoldThreads) {
      const int bl = i / CV, cv = i - bl * CV;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(avg_s + bl * AS + 4 * cv)),
                   "l"(avg + (size_t)(b0 + bl) * C + 4 * cv) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_avg(0, min(kMeshChunk, n_seg));
  float acc[kFoldRows];
#pragma unroll
  for (int j = 0; j < kFoldRows; ++j) acc[j] = 0.f;
#pragma unroll 2
  for (int k = tid; k < K; k += kFoldThreads) {           // K <= 512: at most two columns per thread, all loads in flight
    const float m = mean[k], v = var[k], ga = gamma[k], be = beta[k];
    float w[kFoldRows];
#pragma unroll
    for (int j = 0; j < kFoldRows; ++j) w[j] = n0 + j < N ? W[(size_t)(n0 + j) * K + k] : 0.f;
    const float rm = running ? running_mean[k] : 0.f, rv = running ? running_var[k] : 0.f;
    const float rstd = rsqrtf(v + eps);
    const float sk = ga * rstd;
    const float tk = be - m * sk;
#pragma unroll
    for (int j = 0; j < kFoldRows; ++j) {
      const float ws = w[j] * sk;
      if (n0 + j < N) {
        const float h = tf32_rna(ws);
        Wf_hi[(size_t)(n0 + j) * K + k] = h;
        Wf_lo[(size_t)(n0 + j) * K + k] = tf32_rna(ws - h);
      }
      acc[j] = fmaf(w[j], tk, acc[j]);
      if (k >= C) w_s[j * WS + (k - C)] = ws;
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
    for (int j = 0; j < kFoldRows; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kFoldRows; ++j) red[warp * kFoldRows + j] = acc[j];
  }
  __syncthreads();
  if (tid < kFoldRows) {
    float a = 0.f;
    for (int w8 = 0; w8 < kFoldThreads / 32; ++w8) a += red[w8 * kFoldRows + tid];      // fixed order
    bf_s[tid] = (n0 + tid < N ? b[n0 + tid] : 0.f) + a;
  }
  // thread (bl, j): u[b0 + bl][n0 + j].  A quarter-warp is 2 meshes x 4 rows: its 16-byte reads of avg_s touch two rows
  // AS floats apart (AS mod 32 = 16: disjoint banks), those of w_s four rows WS floats apart (WS mod 32 = 4: disjoint)
  const int bl = tid >> 2, j = tid & 3;
  for (int b0 = 0; b0 < n_seg; b0 += kMeshChunk) {
    const int nb = min(kMeshChunk, n_seg - b0);
    if (b0 > 0) {
      __syncthreads();                                    // previous chunk consumed
      stage_avg(b0, nb);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                      // averages, w_s and bf_s visible
    if (bl < nb && n0 + j < N) {
      const float4* ap = reinterpret_cast<const float4*>(avg_s + bl * AS);
      const float4* wp = reinterpret_cast<const float4*>(w_s + j * WS);
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int c4 = 0; c4 < CV; ++c4) {
        const float4 a = ap[c4], ww = wp[c4];
        d.x = fmaf(a.x, ww.x, d.x); d.y = fmaf(a.y, ww.y, d.y); d.z = fmaf(a.z, ww.z, d.z); d.w = fmaf(a.w, ww.w, d.w);
      }
      u[(size_t)(b0 + bl) * N + n0 + j] = bf_s[j] + ((d.x + d.y) + (d.z + d.w));
    }
  }
}

// Column-slab CTAs (32 columns k of the 2C, 32 row groups = 1024 threads), as bn_fold_bwd_kernel, with
//   G[n][k] = G_L[n][k] (k < C)   |   sum_b SdY[b][n] avg[b][k - C] (k >= C),     sdY[n] = sum_b SdY[b][n]
// and, for the right-half columns, the gradient reaching the per-mesh averages
//   gb[b][c] = inv_cnt[b] ( sum_n SdY[b][n] W[n][C+c] s[C+c] + rps (p[C+c] avg[b][c] + q[C+c]) ).
// Both are small dense products (B x N x C multiply-adds); they run out of shared memory with 4-wide register tiles:
// SdY (all of it, in chunks of kMeshChunk meshes), the slab's averages and the slab's scaled weights are staged once.
// N = 128 or 256 (thread (rg, c) owns the N / 32 rows NPT rg .. NPT rg + NPT - 1).
constexpr int kCols = 32, kGroups = 32;
template <int N>
__global__ void __launch_bounds__(kCols * kGroups)
avg_fold_bwd_kernel(const float* __restrict__ GL, int64_t ldgl, const float* __restrict__ SdY, const float* __restrict__ avg,
                    const float* __restrict__ W, const float* __restrict__ s, const float* __restrict__ t,
                    const float* __restrict__ rstd, const float* __restrict__ mean, const float* __restrict__ inv_cnt,
                    int C, int n_seg, float rps, float inv_rows, int training, float* __restrict__ dW,
                    float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ p_out,
                    float* __restrict__ q_out, float* __restrict__ WsT_hi, float* __restrict__ WsT_lo,
                    float* __restrict__ gb) {
  constexpr int NPT = N / kGroups;
  extern __shared__ __align__(16) float sm[];
  float* sdy_s = sm;                                    // [kMeshChunk][N]
  float* avg_s = sdy_s + kMeshChunk * N;                // [kMeshChunk][kCols]
  float* ws_s = avg_s + kMeshChunk * kCols;             // [N][kCols + 1]   W[n][k] s[k] of this slab (padded: transposed reads)
  float* red = ws_s + N * (kCols + 1);                        // [2][kGroups][kCols]
  float* pq = red + 2 * kGroups * kCols;                // [2][kCols]
  const int K = 2 * C;
  const int c = threadIdx.x % kCols, rg = threadIdx.x / kCols;
  const int k = blockIdx.x * kCols + c;                 // K % 32 == 0: always < K
  const bool right = blockIdx.x * kCols >= C;           // slab-uniform (C % 32 == 0)
  const float sk = s[k], tk = t[k];
  float d[NPT], g[NPT], wreg[NPT], glreg[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {      // this thread's weights / G_L entries: requested now, used after phase 1
    d[i] = g[i] = 0.f;
    wreg[i] = W[(size_t)(NPT * rg + i) * K + k];
    glreg[i] = right ? 0.f : GL[(size_t)(NPT * rg + i) * ldgl + k];
  }
  // ---- phase 1: sdY[n] (and G_R for right-half slabs), meshes in a fixed order
  for (int b0 = 0; b0 < n_seg; b0 += kMeshChunk) {
    const int nb = min(kMeshChunk, n_seg - b0);
    __syncthreads();
    for (int i = threadIdx.x * 4; i < nb * N; i += kCols * kGroups * 4)          // cp.async: every unit in flight at once
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdy_s + i)),
                   "l"(SdY + (size_t)b0 * N + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (right)
      for (int i = threadIdx.x; i < nb * kCols; i += kCols * kGroups)
        avg_s[i] = __ldg(avg + (size_t)(b0 + i / kCols) * C + (blockIdx.x * kCols - C) + (i % kCols));
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int b = 0; b < nb; ++b) {
      const float a = right ? avg_s[b * kCols + c] : 0.f;
#pragma unroll
      for (int i4 = 0; i4 < NPT; i4 += 4) {
        const float4 sd = *reinterpret_cast<const float4*>(sdy_s + b * N + NPT * rg + i4);     // warp-wide broadcast
        d[i4] += sd.x; d[i4 + 1] += sd.y; d[i4 + 2] += sd.z; d[i4 + 3] += sd.w;
        if (right) {
          g[i4] = fmaf(sd.x, a, g[i4]); g[i4 + 1] = fmaf(sd.y, a, g[i4 + 1]);
          g[i4 + 2] = fmaf(sd.z, a, g[i4 + 2]); g[i4 + 3] = fmaf(sd.w, a, g[i4 + 3]);
        }
      }
    }
  }
  float dbeta_p = 0.f, wg_p = 0.f;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int n = NPT * rg + i;
    if (!right) g[i] = glreg[i];
    const float w = wreg[i];
    dW[(size_t)n * K + k] = fmaf(g[i], sk, d[i] * tk);
    dbeta_p = fmaf(w, d[i], dbeta_p);
    wg_p = fmaf(w, g[i], wg_p);
    ws_s[n * (kCols + 1) + c] = w * sk;
    if (blockIdx.x == 0 && c == 0) db[n] = d[i];
  }
  red[(0 * kGroups + rg) * kCols + c] = dbeta_p;
  red[(1 * kGroups + rg) * kCols + c] = wg_p;
  __syncthreads();
  // (W diag(s))^T rows of this slab, coalesced over n (a direct store strides N floats between the lanes of a warp)
  for (int i = threadIdx.x; i < kCols * N; i += kCols * kGroups) {
    const int kl = i / N, n = i - kl * N;
    const float ws = ws_s[n * (kCols + 1) + kl];
    const float h = tf32_rna(ws);
    WsT_hi[(size_t)(blockIdx.x * kCols + kl) * N + n] = h;
    WsT_lo[(size_t)(blockIdx.x * kCols + kl) * N + n] = tf32_rna(ws - h);
  }
  if (rg == 0) {
    float dbeta_k = 0.f, wg = 0.f;
#pragma unroll
    for (int gq = 0; gq < kGroups; ++gq) {   // fixed order
      dbeta_k += red[(0 * kGroups + gq) * kCols + c];
      wg += red[(1 * kGroups + gq) * kCols + c];
    }
    const float dgamma_k = rstd[k] * (wg - mean[k] * dbeta_k);
    dgamma[k] = dgamma_k;
    dbeta[k] = dbeta_k;
    const float pk = training ? -sk * rstd[k] * dgamma_k * inv_rows : 0.f;
    const float qk = training ? -sk * dbeta_k * inv_rows - pk * mean[k] : 0.f;
    p_out[k] = pk;
    q_out[k] = qk;
    pq[c] = pk;
    pq[kCols + c] = qk;
  }
  if (!right) return;
  // ---- phase 2 (right-half slabs): gb[b][c]; thread (rg, c) owns meshes rg and rg + 32 of every chunk
  for (int b0 = 0; b0 < n_seg; b0 += kMeshChunk) {
    const int nb = min(kMeshChunk, n_seg - b0);
    __syncthreads();                                   // pq / ws_s written; previous chunk consumed
    if (n_seg > kMeshChunk) {                          // single chunk: SdY / avg are still staged from phase 1
      for (int i = threadIdx.x * 4; i < nb * N; i += kCols * kGroups * 4)
        *reinterpret_cast<float4*>(sdy_s + i) = __ldg(reinterpret_cast<const float4*>(SdY + (size_t)b0 * N + i));
      for (int i = threadIdx.x; i < nb * kCols; i += kCols * kGroups)
        avg_s[i] = __ldg(avg + (size_t)(b0 + i / kCols) * C + (blockIdx.x * kCols - C) + (i % kCols));
      __syncthreads();
    }
    const float pk = pq[c], qk = pq[kCols + c];
    const int bA = rg, bB = rg + 32;
    float hA = 0.f, hB = 0.f;
    for (int n = 0; n < N; n += 4) {
      const float4 sa = *reinterpret_cast<const float4*>(sdy_s + (bA < nb ? bA : 0) * N + n);
      const float4 sb = *reinterpret_cast<const float4*>(sdy_s + (bB < nb ? bB : 0) * N + n);
      const float w0 = ws_s[(n + 0) * (kCols + 1) + c], w1 = ws_s[(n + 1) * (kCols + 1) + c],
                  w2 = ws_s[(n + 2) * (kCols + 1) + c], w3 = ws_s[(n + 3) * (kCols + 1) + c];
      hA = fmaf(sa.x, w0, hA); hA = fmaf(sa.y, w1, hA); hA = fmaf(sa.z, w2, hA); hA = fmaf(sa.w, w3, hA);
      hB = fmaf(sb.x, w0, hB); hB = fmaf(sb.y, w1, hB); hB = fmaf(sb.z, w2, hB); hB = fmaf(sb.w, w3, hB);
    }
    if (bA < nb)
      gb[(size_t)(b0 + bA) * C + (k - C)] = (hA + rps * fmaf(pk, avg_s[bA * kCols + c], qk)) * __ldg(inv_cnt + b0 + bA);
    if (bB < nb)
      gb[(size_t)(b0 + bB) * C + (k - C)] = (hB + rps * fmaf(pk, avg_s[bB * kCols + c], qk)) * __ldg(inv_cnt + b0 + bB);
  }
}

}  // namespace

}  // namespace sn

SN_API size_t sn_avg_stage_ws_bytes(int64_t n_seg, int64_t C) {
  return (n_seg <= 0 || C <= 0) ? 0 : (size_t)n_seg * sn::kSlices * 3 * (size_t)C * sizeof(float);
}

SN_API int sn_avg_stage_pre_f32(const float* X, int64_t ldx, const float* w, const float* inv_cnt, int64_t rows_per_seg,
                                int64_t n_seg, int64_t C, float* A, int64_t lda, float* mean, float* var_biased, float* avg,
                                void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows_per_seg <= 0 || n_seg <= 0 || C <= 0 || !X || !A || !inv_cnt || !mean || !var_biased || !avg || ldx < C || lda < C)
    return SN_ERR_ARG;
  if (C % 4 || C > 1024 || (kPreThreads % (C / 4)) || ldx % 4 || lda % 4 || !aligned16(X) || !aligned16(A) ||
      n_seg > 65535 || rows_per_seg > 0x7fffffffLL)
    return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_avg_stage_ws_bytes(n_seg, C)) return SN_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int RG = kPreThreads / (int)(C / 4);
  const size_t smem = (size_t)RG * 3 * C * sizeof(float);
  if (smem > 48 * 1024) return SN_ERR_UNSUPPORTED;
  avg_pre_kernel<<<dim3((unsigned)n_seg, kSlices), kPreThreads, smem, st>>>(X, ldx, w, (int)rows_per_seg, (int)C, A, lda,
                                                                          (float*)ws);
  avg_stats_kernel<<<(unsigned)ceil_div(C, 32), 1024, 0, st>>>((const float*)ws, (int)n_seg, n_seg * rows_per_seg, (int)C, A,
                                                               inv_cnt, mean, var_biased, avg);
  return launch_status();
}

SN_API int sn_avg_fold_fwd_f32(const float* mean, const float* var, const float* gamma, const float* beta, const float* W,
                               const float* b, int64_t N, int64_t C, float eps, float* Wf_hi, float* Wf_lo, float* s,
                               float* t, float* rstd, float* running_mean, float* running_var, float momentum, int64_t rows,
                               const float* avg, int64_t n_seg, float* u, int64_t* num_batches_tracked, sn_stream_t stream) {
  using namespace sn;
  if (N <= 0 || C <= 0 || n_seg <= 0 || !mean || !var || !gamma || !beta || !W || !b || !Wf_hi || !Wf_lo || !s || !t ||
      !rstd || !avg || !u)
    return SN_ERR_ARG;
  if (C > 256) return SN_ERR_UNSUPPORTED;
  const float unbias = rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f;
  if (C % 4 || !aligned16(avg)) return SN_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(kMeshChunk * (C + 16) + kFoldRows * (C + 4) + 8 * kFoldRows + kFoldRows) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(avg_fold_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return SN_ERR_UNSUPPORTED;
  }
  avg_fold_fwd_kernel<<<(unsigned)ceil_div(N, kFoldRows), kFoldThreads, smem, (cudaStream_t)stream>>>(
      mean, var, gamma, beta, W, b, (int)N, (int)C, eps, Wf_hi, Wf_lo, s, t, rstd, running_mean, running_var, momentum, unbias,
      avg, (int)n_seg, u, reinterpret_cast<long long*>(num_batches_tracked));
  return launch_status();
}

SN_API int sn_avg_fold_bwd_f32(const float* GL, int64_t ldgl, const float* SdY, const float* avg, const float* W,
                               const float* s, const float* t, const float* rstd, const float* mean, const float* inv_cnt,
                               int64_t N, int64_t C, int64_t n_seg, int64_t rows_per_seg, int training, float* dW, float* db,
                               float* dgamma, float* dbeta, float* p, float* q, float* WsT_hi, float* WsT_lo, float* gb,
                               sn_stream_t stream) {
  using namespace sn;
  if (N <= 0 || C <= 0 || n_seg <= 0 || rows_per_seg <= 0 || !GL || !SdY || !avg || !W || !s || !t || !rstd || !mean ||
      !inv_cnt || !dW || !db || !dgamma || !dbeta || !p || !q || !WsT_hi || !WsT_lo || !gb || ldgl < C)
    return SN_ERR_ARG;
  if (C % 32 || (N != 128 && N != 256) || !aligned16(SdY)) return SN_ERR_UNSUPPORTED;
  const double rows = (double)n_seg * (double)rows_per_seg;
  const size_t smem = (size_t)(kMeshChunk * N + kMeshChunk * kCols + N * (kCols + 1) + 2 * kGroups * kCols + 2 * kCols) * sizeof(float);
  auto kern = N == 128 ? avg_fold_bwd_kernel<128> : avg_fold_bwd_kernel<256>;
  // > 48 KB of dynamic shared memory: opt in (cheap, idempotent)
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return SN_ERR_UNSUPPORTED;
  }
  kern<<<(unsigned)ceil_div(2 * C, kCols), kCols * kGroups, smem, (cudaStream_t)stream>>>(
      GL, ldgl, SdY, avg, W, s, t, rstd, mean, inv_cnt, (int)C, (int)n_seg, (float)rows_per_seg, (float)(1.0 / rows), training,
      dW, db, dgamma, dbeta, p, q, WsT_hi, WsT_lo, gb);
  return launch_status();
}
