// head_tail.cu -- the two ends of the model stacks, which the reference runs as generic ATen / cuBLAS launches:
//
//   * the input layer conv1 = GraphConv1x1(k -> N, no BatchNorm) with k = 3 or 6 input channels
//     (reference src/as_rigid_as_possible/models.py:112, dense_correspondence/models.py:144, mesh_mnist/models_vae.py:26):
//     cuBLAS picks a SIMT sgemm for K = 6 (30 us forward; backward 226 us for dW = dY^T X over 128 000 rows plus a 150 us
//     column-sum for db -- profiles/r2_launches_step_summary.json).  Both directions are one HBM pass here.
//   * the as_rigid_as_possible output head and loss: `conv2(...) + inputs[:, :, -3:].repeat(1, 1, 40)` (models.py:152) and
//     `smooth_l1_loss(outputs * mask, targets, size_average=False) / batch` (main.py:225-226): 10 elementwise / reduction
//     launches over 61 MB tensors in the reference, 4 here.
//
// Reductions are deterministic: per-CTA partials, fixed-order fp64 final sum.
#include "common.cuh"

namespace sn {

namespace {

constexpr int kThreads = 256;

// thread (rg, cv): cv owns output columns 4 cv .. 4 cv + 3 and keeps their K weights in registers; row group rg walks the
// rows rg, rg + RG * gridDim.x, ... (the K inputs of a row are a warp-wide broadcast load)
template <int K>
__global__ void __launch_bounds__(kThreads)
linear_smallk_fwd_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, const float* __restrict__ b,
                         float* __restrict__ Y, int64_t ldy, int64_t rows, int N) {
  const int CV = N / 4, RG = kThreads / CV;
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  if (rg >= RG) return;
  float w[4][K];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int k = 0; k < K; ++k) w[e][k] = __ldg(W + (size_t)(4 * cv + e) * K + k);
  const float4 bias = b ? __ldg(reinterpret_cast<const float4*>(b) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * RG;
  for (int64_t r = (int64_t)blockIdx.x * RG + rg; r < rows; r += stride) {
    float4 y = bias;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float x = __ldg(X + r * ldx + k);
      y.x = fmaf(x, w[0][k], y.x); y.y = fmaf(x, w[1][k], y.y); y.z = fmaf(x, w[2][k], y.z); y.w = fmaf(x, w[3][k], y.w);
    }
    *(reinterpret_cast<float4*>(Y + r * ldy) + cv) = y;
  }
}

// thread (rg, cv): row group rg walks rows rg, rg + RG * gridDim.x ...; cv owns float4 column cv of dY.
// partial[cta][(K + 1)][N]: row 0 = column sums of dY (db), row 1 + k = sum_r dY[r, :] X[r, k] (dW[:, k]).
template <int K>
__global__ void __launch_bounds__(kThreads)
linear_smallk_bwd_kernel(const float* __restrict__ dY, int64_t ldd, const float* __restrict__ X, int64_t ldx, int64_t rows,
                         int N, float* __restrict__ partial) {
  extern __shared__ float red[];                      // [RG][(K + 1)][N]
  const int CV = N / 4, RG = kThreads / CV;
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  float4 acc[K + 1];
#pragma unroll
  for (int k = 0; k <= K; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < RG) {
    const int64_t stride = (int64_t)gridDim.x * RG;
    int64_t r = (int64_t)blockIdx.x * RG + rg;
    for (; r + stride < rows; r += 2 * stride) {        // two rows in flight per thread
      const float4 d0 = __ldcs(reinterpret_cast<const float4*>(dY + r * ldd) + cv);
      const float4 d1 = __ldcs(reinterpret_cast<const float4*>(dY + (r + stride) * ldd) + cv);
      float x0[K], x1[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        x0[k] = __ldg(X + r * ldx + k);
        x1[k] = __ldg(X + (r + stride) * ldx + k);
      }
      acc[0] = add4(acc[0], add4(d0, d1));
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k + 1] = fma4(x1[k], d1, fma4(x0[k], d0, acc[k + 1]));
    }
    for (; r < rows; r += stride) {
      const float4 d = __ldcs(reinterpret_cast<const float4*>(dY + r * ldd) + cv);
      acc[0] = add4(acc[0], d);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k + 1] = fma4(__ldg(X + r * ldx + k), d, acc[k + 1]);
    }
#pragma unroll
    for (int k = 0; k <= K; ++k) *reinterpret_cast<float4*>(red + ((size_t)rg * (K + 1) + k) * N + 4 * cv) = acc[k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (K + 1) * N; i += kThreads) {
    float a = 0.f;
    for (int g = 0; g < RG; ++g) a += red[(size_t)g * (K + 1) * N + i];
    partial[(size_t)blockIdx.x * (K + 1) * N + i] = a;
  }
}

// out[i] = sum_p partial[p][i] (fp64, fixed order): CTA = 32 outputs x 32 groups; group g adds partials g, g + 32, ... and
// the 32 group sums are added in order.  dW gets the [N x K] layout of nn.Linear.
__global__ void __launch_bounds__(1024)
linear_smallk_final_kernel(const float* __restrict__ partial, int n_partials, int N, int K, float* __restrict__ dW,
                           float* __restrict__ db) {
  __shared__ double red[32][33];
  const int lc = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lc;                       // i = k1 * N + n, k1 = 0: db, k1 = k + 1: dW[n][k]
  const int total = (K + 1) * N;
  double a = 0.0;
  if (i < total)
    for (int p = g; p < n_partials; p += 32) a += (double)__ldg(partial + (size_t)p * total + i);
  red[g][lc] = a;
  __syncthreads();
  if (g == 0 && i < total) {
    a = 0.0;
    for (int q = 0; q < 32; ++q) a += red[q][lc];
    const int k1 = i / N, n = i - k1 * N;
    if (k1 == 0) {
      if (db) db[n] = (float)a;
    } else {
      dW[(size_t)n * K + (k1 - 1)] = (float)a;
    }
  }
}

__device__ __forceinline__ float smooth_l1(float d) {
  const float a = fabsf(d);
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float smooth_l1_grad(float d) { return d < -1.f ? -1.f : (d > 1.f ? 1.f : d); }

// out[r, j] = Y[r, j] + In[r, c_in - 3 + j % 3]   (j < n_out; Y may be the first n_out columns of a wider, padded buffer)
// float4 per thread, two in flight (n_out % 4 == 0, 16-byte aligned rows)
__global__ void __launch_bounds__(kThreads)
head_add_tiled_kernel(const float* __restrict__ Y, int64_t ldy, const float* __restrict__ In, int64_t ldi, int c_in,
                      float* __restrict__ Out, int64_t ldo, int64_t rows, int n4) {
  const int64_t total = rows * n4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    float4 y[2];
    int64_t r[2];
    int j[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t ii = i + u * step;
      r[u] = ii / n4;
      j[u] = (int)(ii - r[u] * n4) * 4;
      if (ii < total) y[u] = __ldcs(reinterpret_cast<const float4*>(Y + r[u] * ldy + j[u]));
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i + u * step < total) {
        const float* in = In + r[u] * ldi + (c_in - 3);
        const float t0 = __ldg(in), t1 = __ldg(in + 1), t2 = __ldg(in + 2);
        const int ph = j[u] % 3;                          // column j + e carries channel (j + e) % 3
        const float a = ph == 0 ? t0 : ph == 1 ? t1 : t2, b = ph == 0 ? t1 : ph == 1 ? t2 : t0, c = ph == 0 ? t2 : ph == 1 ? t0 : t1;
        *reinterpret_cast<float4*>(Out + r[u] * ldo + j[u]) = make_float4(y[u].x + a, y[u].y + b, y[u].z + c, y[u].w + a);
      }
    }
  }
}
// gradient of the slice Y[:, :n_out] of a padded [rows x n_pad] buffer: dYp[r, j] = j < n_out ? G[r, j] : 0
__global__ void __launch_bounds__(kThreads)
head_pad_grad_kernel(const float* __restrict__ G, int64_t ldg, float* __restrict__ dYp, int64_t ldd, int64_t rows, int n_out,
                     int np4) {
  const int64_t total = rows * np4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / np4;
    const int j = (int)(i - r * np4) * 4;
    const float4 g = j < n_out ? __ldcs(reinterpret_cast<const float4*>(G + r * ldg + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(dYp + r * ldd + j) = g;
  }
}

// loss partials: sum over elements of smooth_l1(out * m - target); contiguous [rows x C], C % 4 == 0
__global__ void __launch_bounds__(kThreads)
masked_sl1_fwd_kernel(const float* __restrict__ Out, const float* __restrict__ T, const float* __restrict__ M, int64_t rows,
                      int C4, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  const int64_t total = rows * C4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    float4 o[2], t[2];
    float m[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t ii = i + u * step;
      if (ii < total) {
        o[u] = __ldcs(reinterpret_cast<const float4*>(Out) + ii);
        t[u] = __ldcs(reinterpret_cast<const float4*>(T) + ii);
        m[u] = __ldg(M + ii / C4);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (i + u * step < total)
        acc += (smooth_l1(o[u].x * m[u] - t[u].x) + smooth_l1(o[u].y * m[u] - t[u].y)) +
               (smooth_l1(o[u].z * m[u] - t[u].z) + smooth_l1(o[u].w * m[u] - t[u].w));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) a += red[w];
    partial[blockIdx.x] = a;
  }
}
__global__ void __launch_bounds__(1024)
scalar_final_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ double red[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) a += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += red[w];
    out[0] = (float)(s * (double)scale);
  }
}
// dOut[r, c] = gscale[0] * scale * m[r] * smooth_l1'(out * m - target)
__global__ void __launch_bounds__(kThreads)
masked_sl1_bwd_kernel(const float* __restrict__ Out, const float* __restrict__ T, const float* __restrict__ M,
                      const float* __restrict__ gscale, float scale, int64_t rows, int C4, float* __restrict__ dOut) {
  const int64_t total = rows * C4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const float gs = (gscale ? __ldg(gscale) : 1.f) * scale;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * step) {
    float4 o[2], t[2];
    float m[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t ii = i + u * step;
      if (ii < total) {
        o[u] = __ldcs(reinterpret_cast<const float4*>(Out) + ii);
        t[u] = __ldcs(reinterpret_cast<const float4*>(T) + ii);
        m[u] = __ldg(M + ii / C4);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t ii = i + u * step;
      if (ii < total) {
        const float w = gs * m[u];
        reinterpret_cast<float4*>(dOut)[ii] =
            make_float4(w * smooth_l1_grad(o[u].x * m[u] - t[u].x), w * smooth_l1_grad(o[u].y * m[u] - t[u].y),
                        w * smooth_l1_grad(o[u].z * m[u] - t[u].z), w * smooth_l1_grad(o[u].w * m[u] - t[u].w));
      }
    }
  }
}

int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}
unsigned stream_grid(int64_t elems) {
  const int64_t want = ceil_div(elems, kThreads);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

}  // namespace sn

SN_API int sn_linear_smallk_fwd_f32(const float* X, int64_t ldx, const float* W, const float* b, float* Y, int64_t ldy,
                                    int64_t rows, int64_t N, int64_t K, sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || N <= 0 || K <= 0) return SN_ERR_ARG;
  if (rows == 0) return SN_OK;
  if (!X || !W || !Y || ldx < K || ldy < N) return SN_ERR_ARG;
  if ((K != 3 && K != 6) || N % 4 || N > 1024 || (kThreads % (N / 4)) || ldy % 4 || !aligned16(Y) || (b && !aligned16(b)))
    return SN_ERR_UNSUPPORTED;
  const int RG = kThreads / (int)(N / 4);
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid * RG > rows) grid = ceil_div(rows, RG);
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 3)
    linear_smallk_fwd_kernel<3><<<(unsigned)grid, kThreads, 0, st>>>(X, ldx, W, b, Y, ldy, rows, (int)N);
  else
    linear_smallk_fwd_kernel<6><<<(unsigned)grid, kThreads, 0, st>>>(X, ldx, W, b, Y, ldy, rows, (int)N);
  return launch_status();
}

SN_API size_t sn_linear_smallk_bwd_ws_bytes(int64_t N, int64_t K) {
  return (N <= 0 || K <= 0) ? 0 : (size_t)sn::sm_count() * 4 * (size_t)(K + 1) * (size_t)N * sizeof(float);
}

SN_API int sn_linear_smallk_bwd_f32(const float* dY, int64_t ldd, const float* X, int64_t ldx, int64_t rows, int64_t N,
                                    int64_t K, float* dW, float* db, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows <= 0 || N <= 0 || K <= 0 || !dY || !X || !dW || ldd < N || ldx < K) return SN_ERR_ARG;
  if ((K != 3 && K != 6) || N % 4 || N > 256 || (kThreads % (N / 4)) || ldd % 4 || !aligned16(dY)) return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_linear_smallk_bwd_ws_bytes(N, K)) return SN_ERR_WORKSPACE;
  const int RG = kThreads / (int)(N / 4);
  int grid = sm_count() * 4;
  if ((int64_t)grid * RG > rows) grid = (int)ceil_div(rows, RG);
  const size_t smem = (size_t)RG * (K + 1) * N * sizeof(float);
  if (smem > 48 * 1024) return SN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 3)
    linear_smallk_bwd_kernel<3><<<grid, kThreads, smem, st>>>(dY, ldd, X, ldx, rows, (int)N, (float*)ws);
  else
    linear_smallk_bwd_kernel<6><<<grid, kThreads, smem, st>>>(dY, ldd, X, ldx, rows, (int)N, (float*)ws);
  linear_smallk_final_kernel<<<(unsigned)ceil_div((K + 1) * N, 32), 1024, 0, st>>>((const float*)ws, grid, (int)N, (int)K, dW,
                                                                                db);
  return launch_status();
}

SN_API int sn_head_add_tiled_f32(const float* Y, int64_t ldy, const float* In, int64_t ldi, int64_t c_in, float* Out,
                                 int64_t ldo, int64_t rows, int64_t n_out, sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || n_out <= 0 || c_in < 3) return SN_ERR_ARG;
  if (rows == 0) return SN_OK;
  if (!Y || !In || !Out || ldy < n_out || ldo < n_out || ldi < c_in) return SN_ERR_ARG;
  if (n_out % 4 || ldy % 4 || ldo % 4 || !aligned16(Y) || !aligned16(Out)) return SN_ERR_UNSUPPORTED;
  head_add_tiled_kernel<<<stream_grid(rows * (n_out / 4)), kThreads, 0, (cudaStream_t)stream>>>(Y, ldy, In, ldi, (int)c_in, Out,
                                                                                           ldo, rows, (int)(n_out / 4));
  return launch_status();
}

SN_API int sn_head_pad_grad_f32(const float* G, int64_t ldg, float* dYp, int64_t ldd, int64_t rows, int64_t n_out,
                                int64_t n_pad, sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || n_out <= 0 || n_pad < n_out) return SN_ERR_ARG;
  if (rows == 0) return SN_OK;
  if (!G || !dYp || ldg < n_out || ldd < n_pad) return SN_ERR_ARG;
  if (n_out % 4 || n_pad % 4 || ldg % 4 || ldd % 4 || !aligned16(G) || !aligned16(dYp)) return SN_ERR_UNSUPPORTED;
  head_pad_grad_kernel<<<stream_grid(rows * (n_pad / 4)), kThreads, 0, (cudaStream_t)stream>>>(G, ldg, dYp, ldd, rows, (int)n_out,
                                                                                          (int)(n_pad / 4));
  return launch_status();
}

SN_API size_t sn_masked_smooth_l1_ws_bytes(void) { return (size_t)sn::sm_count() * 8 * sizeof(float); }

SN_API int sn_masked_smooth_l1_fwd_f32(const float* Out, const float* T, const float* M, int64_t rows, int64_t C, float scale,
                                       float* loss, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows <= 0 || C <= 0 || !Out || !T || !M || !loss) return SN_ERR_ARG;
  if (C % 4 || !aligned16(Out) || !aligned16(T)) return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_masked_smooth_l1_ws_bytes()) return SN_ERR_WORKSPACE;
  const unsigned grid = stream_grid(rows * (C / 4));
  cudaStream_t st = (cudaStream_t)stream;
  masked_sl1_fwd_kernel<<<grid, kThreads, 0, st>>>(Out, T, M, rows, (int)(C / 4), (float*)ws);
  scalar_final_kernel<<<1, 1024, 0, st>>>((const float*)ws, (int)grid, scale, loss);
  return launch_status();
}

SN_API int sn_masked_smooth_l1_bwd_f32(const float* Out, const float* T, const float* M, const float* grad_loss, int64_t rows,
                                       int64_t C, float scale, float* dOut, sn_stream_t stream) {
  using namespace sn;
  if (rows <= 0 || C <= 0 || !Out || !T || !M || !dOut) return SN_ERR_ARG;
  if (C % 4 || !aligned16(Out) || !aligned16(T) || !aligned16(dOut)) return SN_ERR_UNSUPPORTED;
  masked_sl1_bwd_kernel<<<stream_grid(rows * (C / 4)), kThreads, 0, (cudaStream_t)stream>>>(Out, T, M, grad_loss, scale, rows,
                                                                                        (int)(C / 4), dOut);
  return launch_status();
}
