// common.cuh -- shared helpers for libsurfnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/surfnet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsurfnet_b200 targets sm_100a (B200) only"
#endif

#define SN_API extern "C" __attribute__((visibility("default")))

namespace sn {

constexpr int kWarp = 32;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky-less launch error so the next call starts clean
    return (int)e;
  }
  return SN_OK;
}

// Device-wide exclusive scan (convert.cu): out[0..n) = exclusive prefix sums of counts, out[n] = total; out may alias
// counts; tile_ws holds exclusive_scan_tiles(n) ints.
int64_t exclusive_scan_tiles(int64_t n);
int exclusive_scan(const int* counts, int64_t n, int* out, int* tile_ws, cudaStream_t st);

// colstats.cu: fixed-order fp64 reduction of per-CTA partial sums [n_partials][2][C] -> mean / biased variance
int launch_colstats_final(const float* partial, int n_partials, int64_t rows, int C, const float* shift, float* mean,
                          float* var, cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ELU(alpha = 1), the activation in front of every operator application (utils_pt.py:161,172,195,208).
// torch evaluates the negative branch with expm1.  expm1f from libdevice costs ~35 instructions with branches, which made
// the GEMM's activation epilogue issue-bound (profiles/r2_fusion_notes.md); this branch-free Cody-Waite form is 18:
//   n = rint(x log2 e),  r = x - n ln 2 (hi/lo),  e^r - 1 = r + r^2 (1/2 + r/6 + ... + r^5/7!)   (|r| <= 0.347),
//   e^x - 1 = 2^n (e^r - 1) + (2^n - 1)            -- one FMA; 2^n - 1 is exact
// Measured against float64 expm1 over [-100, 0] (tools/elu_accuracy.py): max error 0.87 ulp, i.e. as close to the
// reference's activation as expm1f (1 ulp) is.
__device__ __forceinline__ float expm1_nonpos(float x) {
  const float xn = fmaxf(x, -30.f);                            // e^-30 - 1 rounds to -1
  const float t = fmaf(xn, 1.4426950408889634f, 12582912.f);   // 1.5 * 2^23 + rint(x log2 e)
  const float n = t - 12582912.f;
  float r = fmaf(n, -0.693145751953125f, xn);
  r = fmaf(n, -1.428606765330187045e-06f, r);
  float p = 1.f / 5040.f;
  p = fmaf(p, r, 1.f / 720.f);
  p = fmaf(p, r, 1.f / 120.f);
  p = fmaf(p, r, 1.f / 24.f);
  p = fmaf(p, r, 1.f / 6.f);
  p = fmaf(p, r, 0.5f);
  const float pm1 = fmaf(p, r * r, r);
  const float s = __int_as_float((__float_as_int(t) << 23) + 0x3f800000);   // 2^n, n in [-44, 0]
  return fmaf(s, pm1, s - 1.f);
}
__device__ __forceinline__ float elu1(float x) { return x <= 0.f ? expm1_nonpos(x) : x; }   // NaN propagates
__device__ __forceinline__ float4 elu4(float4 v) {
  return make_float4(elu1(v.x), elu1(v.y), elu1(v.z), elu1(v.w));
}

// Read-only 128-bit gather of a dense feature row segment.  Rows are re-read by neighbouring sparse
// rows (mesh valence ~6), so they are allowed to allocate in L1.
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Streaming 128-bit store of an output row segment: written once, never re-read by this kernel.
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}

__device__ __forceinline__ float4 fma4(float a, float4 x, float4 acc) {
  acc.x = fmaf(a, x.x, acc.x);
  acc.y = fmaf(a, x.y, acc.y);
  acc.z = fmaf(a, x.z, acc.z);
  acc.w = fmaf(a, x.w, acc.w);
  return acc;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 sel4(bool c, float4 a, float4 b) { return c ? a : b; }
__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
  v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
  v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
  v.z = __shfl_xor_sync(0xffffffffu, v.z, m);
  v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
  return v;
}

__device__ __forceinline__ float4 shfl_idx4(float4 v, int src_lane) {
  v.x = __shfl_sync(0xffffffffu, v.x, src_lane);
  v.y = __shfl_sync(0xffffffffu, v.y, src_lane);
  v.z = __shfl_sync(0xffffffffu, v.z, src_lane);
  v.w = __shfl_sync(0xffffffffu, v.w, src_lane);
  return v;
}

}  // namespace sn
