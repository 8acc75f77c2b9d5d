// spmm_rowdirect.cu -- the latency-bound regime of the two SpMM families: operators so small that the whole launch is one or
// two waves of threads (ONE mesh of a few thousand vertices -- BASELINE cfg5, FAUST-sized Dirac at C = 16 ... 512 -- or a
// mesh_mnist batch of 32 x 500 vertices, cfg2; reference src/utils/utils_pt.py:167,176,201-203,213-215).
//
// Why a separate kernel (round-1 verdict, BENCH spmm table: 0.04 - 0.27 of the HBM roofline at these sizes): the persistent
// row-group kernel (spmm_rowgroup.cu) is built for throughput -- row pointers and column indices staged two warp-tiles
// ahead, ONE entry in flight per row group, 64 registers for occupancy.  With at most one warp-tile per resident warp
// none of that pipelining ever overlaps anything: a row of n entries costs 2 + n DEPENDENT memory round trips
// (row pointers -> column indices -> one gather per entry), ~1 us each out of a cold L2.  Here
//
//   * the grid is sized by the operator (one row group = C/16 lanes per sparse row, same lane mapping and the same
//     summation order as the row-group kernel, hence bit-identical results), no persistent loop, no index ring;
//   * a row group reads its two row pointers straight into registers, then keeps E entries in flight at once: their
//     column indices, their 4 x 16-byte X segments per lane (registers) and their 4x4 blocks (cp.async into a per-group
//     shared-memory slot, issued BEFORE the indices arrive: the address only needs the row pointer);
//   * the column indices of the NEXT E entries are requested while the current gathers fly.
//
// A row of n entries costs 2 + ceil(n / E) round trips: 3 for D (exactly three blocks per face), 4 for D* / L at valence
// <= 8.  Registers are spent freely (96 / 114 per thread), so the kernel is used only while ALL row groups of the launch
// are resident at once (rowdirect_applies below); beyond one wave the persistent kernel is as fast or faster.
//
// Measured on B200, one 7000-vertex mesh, L2 flushed (profiles/r2_small_operator_spmm.jsonl): D* 16.5 -> 12.2 us at
// C <= 128, D 10.6 -> 9.3 us at C <= 64.  The timing method itself reports 7.1 us for a 32-row operator, so these
// launches sit 2 - 5 us above the floor of "launch + first cold misses"; the rest is not the kernel's to win.
//
// Bound: latency; algorithmic bytes as in spmm_rowgroup.cu.
#include "common.cuh"

namespace sn {

namespace {

constexpr int kDirectWarps = 4;
constexpr int kDirectThreads = kDirectWarps * 32;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// LPR lanes per sparse row (C = 16 LPR), BLK = 4: BSR4 (rotated column-major 4x4 blocks, see sn_csr32_to_bsr4_fill),
// BLK = 1: CSR; E entries in flight per row group.
template <int LPR, int BLK, int E>
__global__ void __launch_bounds__(kDirectThreads)
rowdirect_spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                      const float* __restrict__ val, const float* __restrict__ X, uint32_t ldxb,
                      float* __restrict__ Y, uint32_t ldyb, int n_rows) {
  constexpr int G = 32 / LPR;                  // row groups per warp
  constexpr int C = 16 * LPR;
  constexpr int kQuarterBytes = C;             // (C/4 floats) * 4 bytes
  __shared__ __align__(16) float vring[BLK == 4 ? kDirectWarps * G * E * 16 : 4];   // [warp][group][E][16]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane / LPR, t = lane % LPR;
  const int row = (blockIdx.x * kDirectWarps + warp) * G + g;
  const bool live = row < n_rows;
  int k = 0, kend = 0;
  if (live) {
    k = __ldg(rowptr + row);
    kend = __ldg(rowptr + row + 1);
  }
  const int n_iter = __reduce_max_sync(0xffffffffu, (kend - k + E - 1) / E);   // warp-uniform trip count
  float* vslot = vring + (warp * G + g) * E * 16;
  const char* Xl = reinterpret_cast<const char*>(X) + t * 16;

  auto load_values = [&](int kk) {             // blocks kk .. kk + E - 1 -> this group's slot, 16-byte units over its lanes
    if (BLK == 4) {
      const uint32_t dst = smem_addr(vslot);
      const char* src = reinterpret_cast<const char*>(val) + (size_t)kk * 64u;
#pragma unroll
      for (int u0 = 0; u0 < 4 * E; u0 += LPR) {
        const int u = u0 + t;
        if (u < 4 * E && kk + (u >> 2) < kend)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * 16), "l"(src + u * 16) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  auto load_indices = [&](int (&j)[E], float (&w)[E], int kk) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      j[e] = kk + e < kend ? __ldg(colind + kk + e) : -1;
      if (BLK == 1) w[e] = kk + e < kend ? __ldg(val + kk + e) : 0.f;
    }
  };

  float4 acc[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
  int jn[E];
  float wn[E];
  load_values(k);
  load_indices(jn, wn, k);
  for (int it = 0; it < n_iter; ++it) {
    float4 xs[E][4];
    float w[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      w[e] = wn[e];
      if (jn[e] >= 0) {
        const char* xp = Xl + (size_t)(uint32_t)jn[e] * ldxb;
#pragma unroll
        for (int q = 0; q < 4; ++q) xs[e][q] = __ldg(reinterpret_cast<const float4*>(xp + q * kQuarterBytes));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) xs[e][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const int kcur = k;
    k += E;
    if (it + 1 < n_iter) load_indices(jn, wn, k);     // next chunk's indices travel with this chunk's gathers
    if (BLK == 4) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (kcur + e < kend) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 xv = xs[e][q];
          if (BLK == 4) {       // w = (B[q][q], B[q+1][q], B[q+2][q], B[q+3][q]), rows mod 4
            const float4 wq = *reinterpret_cast<const float4*>(vslot + 16 * e + 4 * q);
            acc[q] = fma4(wq.x, xv, acc[q]);
            acc[(q + 1) & 3] = fma4(wq.y, xv, acc[(q + 1) & 3]);
            acc[(q + 2) & 3] = fma4(wq.z, xv, acc[(q + 2) & 3]);
            acc[(q + 3) & 3] = fma4(wq.w, xv, acc[(q + 3) & 3]);
          } else {
            acc[q] = fma4(w[e], xv, acc[q]);
          }
        }
      }
    }
    if (BLK == 4 && it + 1 < n_iter) {
      __syncwarp();                                    // every lane has read the slot before it is refilled
      load_values(k);
    }
  }
  if (live) {
    char* yrow = reinterpret_cast<char*>(Y) + t * 16 + (size_t)(uint32_t)row * ldyb;
#pragma unroll
    for (int p = 0; p < 4; ++p) *reinterpret_cast<float4*>(yrow + p * kQuarterBytes) = acc[p];
  }
}

template <int LPR, int BLK, int E>
int launch_direct(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx, float* Y,
                  int64_t ldy, int64_t n_rows, cudaStream_t st) {
  constexpr int G = 32 / LPR;
  const int64_t grid = ceil_div(n_rows, (int64_t)kDirectWarps * G);
  if (grid > 0x7fffffffLL) return SN_ERR_UNSUPPORTED;
  rowdirect_spmm_kernel<LPR, BLK, E><<<(unsigned)grid, kDirectThreads, 0, st>>>(rowptr, colind, val, X, (uint32_t)(ldx * 4), Y,
                                                                               (uint32_t)(ldy * 4), (int)n_rows);
  return launch_status();
}

template <int BLK>
int launch_direct_family(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                         float* Y, int64_t ldy, int64_t n_rows, int64_t n_entries, int64_t C, cudaStream_t st) {
  if (n_rows >= 0x7fffff00LL || ldx >= (1LL << 30) || ldy >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  // three entries in flight for rows of up to three entries (D: exactly three blocks per face), four otherwise
  const bool e3 = n_entries >= 0 && n_entries <= 3 * n_rows;
#define SN_RD(LPR) \
  (e3 ? launch_direct<LPR, BLK, 3>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, st) \
      : launch_direct<LPR, BLK, 4>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, st))
  switch (C) {
    case 16: return SN_RD(1);
    case 32: return SN_RD(2);
    case 64: return SN_RD(4);
    case 128: return SN_RD(8);
    case 256: return SN_RD(16);
    case 512: return SN_RD(32);
    default: return SN_ERR_UNSUPPORTED;
  }
#undef SN_RD
}

}  // namespace

// The launch is "small" when all its row groups are resident at once (one wave: 96 / 114 registers -> 640 / 512 threads
// per SM); beyond that the persistent row-group kernel is as fast or faster.  Measured on B200 (tools/spmm_bench.py
// --variants rg6,rg7, one 7000-vertex mesh, L2 flushed): D* at C <= 128 16.5 -> 12.2 us, D at C <= 64 10.6 -> 9.3 us;
// at 1.2+ waves (D at C = 128: 11.6 vs 10.8 us, D at C = 512: 21 vs 18 us) the persistent kernel wins.
bool rowdirect_applies(int64_t n_rows, int64_t C, bool three_in_flight) {
  if (C != 16 && C != 32 && C != 64 && C != 128 && C != 256 && C != 512) return false;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return n_rows * (C / 16) <= (int64_t)sms * (three_in_flight ? 640 : 512);
}

// n_entries < 0: unknown (four entries in flight).  Both return SN_ERR_UNSUPPORTED for widths outside 16 ... 512.
int launch_bsr4_rowdirect(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X, int64_t ldx,
                          float* Y, int64_t ldy, int64_t n_brows, int64_t n_blocks, int64_t C, cudaStream_t st) {
  return launch_direct_family<4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, n_blocks, C, st);
}
int launch_csr_rowdirect(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                         float* Y, int64_t ldy, int64_t n_rows, int64_t nnz, int64_t C, cudaStream_t st) {
  return launch_direct_family<1>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, nnz, C, st);
}

}  // namespace sn
