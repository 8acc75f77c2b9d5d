// elementwise.cu -- the activation pass in front of each operator application.
//
// Every LapResNet2 / DirResNet2 stage starts with F.elu (reference src/utils/utils_pt.py:161,172,195,208)
// and concatenates the activated features with S * activated features (:168,177,204,216).  These kernels
// write the activated rows straight into the left half of the concat buffer [rows x 2C] (strided
// output), so torch.cat never runs and the SpMM gathers from / writes into the same buffer.
#include "common.cuh"

namespace sn {

// grid-stride over float4 elements of a [rows x C] matrix with arbitrary (16 B aligned) leading dims
__global__ void __launch_bounds__(256)
elu_vec4_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int64_t rows, int C4) {
  const int64_t total = rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(X + r * ldx + c));
    *reinterpret_cast<float4*>(Y + r * ldy + c) = elu4(v);
  }
}
__global__ void __launch_bounds__(256)
elu_scalar_kernel(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    Y[r * ldy + c] = elu1(X[r * ldx + c]);
  }
}
// Y[r,c] = (G[r,c] + G2[r,c]) * elu'(x).  RAW: A holds x itself (elu' = x > 0 ? 1 : exp(x));
// otherwise A holds the activated value a = elu(x) (elu' = a > 0 ? 1 : a + 1).  G2 may be null.
template <bool RAW>
__global__ void __launch_bounds__(256)
elu_bwd_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ G, int64_t ldg,
               const float* __restrict__ G2, int64_t ldg2, float* __restrict__ Y, int64_t ldy, int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float a = A[r * lda + c];
    float g = G[r * ldg + c];
    if (G2) g += G2[r * ldg2 + c];
    const float d = a > 0.f ? 1.f : (RAW ? expf(a) : a + 1.f);
    Y[r * ldy + c] = g * d;
  }
}

template <bool RAW>
__global__ void __launch_bounds__(256)
elu_bwd_vec4_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ G, int64_t ldg,
                    const float* __restrict__ G2, int64_t ldg2, float* __restrict__ Y, int64_t ldy, int64_t rows, int C4) {
  const int64_t total = rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(A + r * lda + c);
    float4 g = __ldcs(reinterpret_cast<const float4*>(G + r * ldg + c));
    if (G2) g = add4(g, __ldcs(reinterpret_cast<const float4*>(G2 + r * ldg2 + c)));
    float4 y;
    y.x = g.x * (a.x > 0.f ? 1.f : (RAW ? expf(a.x) : a.x + 1.f));
    y.y = g.y * (a.y > 0.f ? 1.f : (RAW ? expf(a.y) : a.y + 1.f));
    y.z = g.z * (a.z > 0.f ? 1.f : (RAW ? expf(a.z) : a.z + 1.f));
    y.w = g.w * (a.w > 0.f ? 1.f : (RAW ? expf(a.w) : a.w + 1.f));
    *reinterpret_cast<float4*>(Y + r * ldy + c) = y;
  }
}

// grid = (n_seg, kSegSplit): CTA (s, j) sums the j-th slice of segment s's rows; thread (rg, cv) walks rows of float4
// column cv; fixed-order smem reduction; partial[s][j][C].  A second tiny kernel adds the kSegSplit slices in order.
constexpr int kSegSplit = 8;
__global__ void __launch_bounds__(256)
segment_sum_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, int rows_per_seg, int C,
                   float* __restrict__ partial) {
  extern __shared__ float red[];               // [RG][C]
  const int CV = C / 4, RG = 256 / CV;
  const int cv = threadIdx.x % CV, rg = threadIdx.x / CV;
  const int slice = (rows_per_seg + kSegSplit - 1) / kSegSplit;
  const int rb = blockIdx.y * slice, re = min(rb + slice, rows_per_seg);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_seg;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rg < RG) {
    // (eight predicated rows in flight were measured slower: 20.9 vs 16.4 us per launch at 128000 x 128)
    int r = rb + rg;
    for (; r + 3 * RG < re; r += 4 * RG) {      // four rows in flight per thread, added in row order
      float4 v[4];
      float ww[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = __ldcs(reinterpret_cast<const float4*>(X + (r0 + r + u * RG) * ldx) + cv);
        ww[u] = w ? __ldg(w + r0 + r + u * RG) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s.x = fmaf(ww[u], v[u].x, s.x); s.y = fmaf(ww[u], v[u].y, s.y); s.z = fmaf(ww[u], v[u].z, s.z); s.w = fmaf(ww[u], v[u].w, s.w);
      }
    }
    for (; r < re; r += RG) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(X + (r0 + r) * ldx) + cv);
      const float ww = w ? __ldg(w + r0 + r) : 1.f;
      s.x = fmaf(ww, v.x, s.x); s.y = fmaf(ww, v.y, s.y); s.z = fmaf(ww, v.z, s.z); s.w = fmaf(ww, v.w, s.w);
    }
    *reinterpret_cast<float4*>(red + (size_t)rg * C + 4 * cv) = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f;
    for (int g = 0; g < RG; ++g) a += red[(size_t)g * C + c];
    partial[((size_t)blockIdx.x * kSegSplit + blockIdx.y) * C + c] = a;
  }
}
__global__ void segment_sum_final_kernel(const float* __restrict__ partial, int64_t n, int C, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // i = seg * C + c
  if (i >= n) return;
  const int64_t sgm = i / C;
  const int c = (int)(i - sgm * C);
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < kSegSplit; ++j) a += partial[(sgm * kSegSplit + j) * C + c];
  out[i] = a;
}

__global__ void __launch_bounds__(256)
elu_bwd_group_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ G, int64_t ldg,
                     const float* __restrict__ GB, const float* __restrict__ w, int rows_per_seg,
                     const float* __restrict__ G3, int64_t ldg3, float* __restrict__ Y, int64_t ldy, int64_t rows, int C4) {
  const int64_t total = rows * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C4;
    const int c = (int)(i - r * C4) * 4;
    const float4 a = *reinterpret_cast<const float4*>(A + r * lda + c);
    float4 g = __ldcs(reinterpret_cast<const float4*>(G + r * ldg + c));
    const float4 gb = __ldg(reinterpret_cast<const float4*>(GB + (r / rows_per_seg) * (int64_t)(C4 * 4) + c));
    const float ww = w ? __ldg(w + r) : 1.f;
    g.x = fmaf(ww, gb.x, g.x); g.y = fmaf(ww, gb.y, g.y); g.z = fmaf(ww, gb.z, g.z); g.w = fmaf(ww, gb.w, g.w);
    float4 y;
    y.x = g.x * (a.x > 0.f ? 1.f : a.x + 1.f);
    y.y = g.y * (a.y > 0.f ? 1.f : a.y + 1.f);
    y.z = g.z * (a.z > 0.f ? 1.f : a.z + 1.f);
    y.w = g.w * (a.w > 0.f ? 1.f : a.w + 1.f);
    if (G3) {                     // gradient that bypasses the activation (the block residual)
      const float4 g3 = __ldcs(reinterpret_cast<const float4*>(G3 + r * ldg3 + c));
      y.x += g3.x; y.y += g3.y; y.z += g3.z; y.w += g3.w;
    }
    *reinterpret_cast<float4*>(Y + r * ldy + c) = y;
  }
}

}  // namespace sn

SN_API size_t sn_segment_sum_ws_bytes(int64_t n_seg, int64_t C) {
  return (n_seg <= 0 || C <= 0) ? 0 : (size_t)n_seg * sn::kSegSplit * (size_t)C * sizeof(float);
}

SN_API int sn_segment_sum_f32(const float* X, int64_t ldx, const float* w, int64_t rows_per_seg, int64_t n_seg, int64_t C,
                              float* out, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (rows_per_seg <= 0 || n_seg < 0 || C <= 0 || !X || !out || ldx < C) return SN_ERR_ARG;
  if (n_seg == 0) return SN_OK;
  if (C % 4 || C > 1024 || (256 % (C / 4)) || ldx % 4 || !aligned16(X) || rows_per_seg > 0x7fffffffLL || n_seg > 0x7fffffffLL)
    return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_segment_sum_ws_bytes(n_seg, C)) return SN_ERR_WORKSPACE;
  if (n_seg > 65535LL * 1024) return SN_ERR_UNSUPPORTED;
  const int RG = 256 / (int)(C / 4);
  cudaStream_t st = (cudaStream_t)stream;
  segment_sum_kernel<<<dim3((unsigned)n_seg, kSegSplit), 256, (size_t)RG * C * sizeof(float), st>>>(
      X, ldx, w, (int)rows_per_seg, (int)C, (float*)ws);
  segment_sum_final_kernel<<<(unsigned)ceil_div(n_seg * C, 256), 256, 0, st>>>((const float*)ws, n_seg * C, (int)C, out);
  return launch_status();
}

SN_API int sn_elu_bwd_group_f32(const float* A, int64_t lda, const float* G, int64_t ldg, const float* GB, const float* w,
                                int64_t rows_per_seg, const float* G3, int64_t ldg3, float* Y, int64_t ldy, int64_t rows,
                                int64_t C, sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || C <= 0 || rows_per_seg <= 0 || !A || !G || !GB || !Y || lda < C || ldg < C || ldy < C || (G3 && ldg3 < C))
    return SN_ERR_ARG;
  if (rows == 0) return SN_OK;
  if (C % 4 || lda % 4 || ldg % 4 || ldy % 4 || !aligned16(A) || !aligned16(G) || !aligned16(GB) || !aligned16(Y) ||
      rows_per_seg > 0x7fffffffLL || (G3 && (ldg3 % 4 || !aligned16(G3))))
    return SN_ERR_UNSUPPORTED;
  const int64_t work = rows * (C / 4);
  const unsigned grid = (unsigned)(ceil_div(work, 256) < 148 * 16 ? ceil_div(work, 256) : 148 * 16);
  elu_bwd_group_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, G, ldg, GB, w, (int)rows_per_seg, G3, ldg3, Y, ldy, rows,
                                                               (int)(C / 4));
  return launch_status();
}

SN_API int sn_elu_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t rows, int64_t C,
                      sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (rows == 0 || C == 0) return SN_OK;
  if (!X || !Y || ldx < C || ldy < C) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y);
  const int64_t work = vec ? rows * (C / 4) : rows * C;
  const unsigned grid = (unsigned)(ceil_div(work, 256) < 148 * 16 ? ceil_div(work, 256) : 148 * 16);
  if (vec)
    elu_vec4_kernel<<<grid, 256, 0, st>>>(X, ldx, Y, ldy, rows, (int)(C / 4));
  else
    elu_scalar_kernel<<<grid, 256, 0, st>>>(X, ldx, Y, ldy, rows, (int)C);
  return launch_status();
}

SN_API int sn_elu_bwd_f32(const float* A, int64_t lda, int a_is_raw, const float* G, int64_t ldg, const float* G2,
                          int64_t ldg2, float* Y, int64_t ldy, int64_t rows, int64_t C, sn_stream_t stream) {
  using namespace sn;
  if (rows < 0 || C < 0 || C > 0x7fffffffLL) return SN_ERR_ARG;
  if (rows == 0 || C == 0) return SN_OK;
  if (!A || !G || !Y || lda < C || ldg < C || ldy < C || (G2 && ldg2 < C)) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && (lda % 4 == 0) && (ldg % 4 == 0) && (ldy % 4 == 0) && (!G2 || ldg2 % 4 == 0) &&
                   aligned16(A) && aligned16(G) && aligned16(Y) && (!G2 || aligned16(G2));
  if (vec) {
    const int64_t work = rows * (C / 4);
    const unsigned vgrid = (unsigned)(ceil_div(work, 256) < 148 * 16 ? ceil_div(work, 256) : 148 * 16);
    if (a_is_raw)
      elu_bwd_vec4_kernel<true><<<vgrid, 256, 0, st>>>(A, lda, G, ldg, G2, ldg2, Y, ldy, rows, (int)(C / 4));
    else
      elu_bwd_vec4_kernel<false><<<vgrid, 256, 0, st>>>(A, lda, G, ldg, G2, ldg2, Y, ldy, rows, (int)(C / 4));
    return launch_status();
  }
  const unsigned grid = (unsigned)(ceil_div(rows * C, 256) < 148 * 16 ? ceil_div(rows * C, 256) : 148 * 16);
  if (a_is_raw)
    elu_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, G, ldg, G2, ldg2, Y, ldy, rows, (int)C);
  else
    elu_bwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, G, ldg, G2, ldg2, Y, ldy, rows, (int)C);
  return launch_status();
}
