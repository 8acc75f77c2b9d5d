// spmm_rowgroup.cu -- register-streaming SpMM for both operator families: the 4x4-block Dirac / adjoint
// (reference src/utils/utils_pt.py:201-203,213-215) and the scalar cotangent Laplacian (utils_pt.py:167,176).
//
// Why a third design (measured history in profiles/r1_bsr4_notes.md): the cp.async streaming kernel
// (spmm_bsr4_stream.cu) reached 0.50 of the HBM roofline and then ran out of ON-CHIP throughput, not DRAM: every
// gathered 512 B row crossed shared memory twice (LDGSTS write + LDS read), and the lane mapping (one lane per
// quaternion component q) needed 12 shuffles per output row -- ncu: L1TEX/shared wavefronts 80 %, issue slots 63 %,
// 61.6 SASS instructions per block.  This kernel removes both:
//
//   * LPR = C/16 lanes own one sparse row ("row group"); a warp works on 32/LPR rows at once.  Lane t loads the four
//     float4 X[j, q*C/4 + 4t .. +3], q = 0..3 -- for the Dirac view that is all four quaternion components of its
//     column slice, so the whole 4x4 block product (64 FMAs) stays inside the lane: NO shuffles, no shared-memory
//     round trip for X, 4 row groups per warp instruction at C = 128 (each still reads whole 128 B lines).
//   * gathers go straight to registers (LDG.128), software-pipelined PD blocks ahead per row group; with the
//     indices already in shared memory the three-deep dependent chain of the first direct-gather kernel is gone;
//   * warps are independent persistent workers: each walks warp-tiles of WR = (32/LPR)*RPG consecutive rows, keeps
//     that tile's row pointers / column indices in a private shared-memory ring filled by its own cp.async two
//     tiles ahead, and synchronises with nobody (no __syncthreads in the kernel);
//   * a row group streams through the contiguous run of blocks of its RPG rows; row ends are detected against the
//     staged row pointers, empty rows (padding faces / vertices of ragged batches) store zeros.
//
// Summation order per output element: ascending block (storage) order, q = 0..3 inside a block, fp32 FMA -- the same
// order as bsr4_spmm_scalar_kernel / csr_spmm_scalar_kernel, bit-reproducible run to run.
//
// Bound: HBM.  Algorithmic bytes per launch:
//   BSR4: 4(Rb+1) + 68 nb + 4 Cb C + 4 Rb C        CSR: 4(R+1) + 8 nnz + 8 R C
#include "common.cuh"

namespace sn {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// base + a * b with 32-bit a, b and a 64-bit base: one IMAD.WIDE.U32
__device__ __forceinline__ const char* ptr_mad(const char* base, uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const char*>(r);
}

template <int LPR, int RPG>
struct Geo {
  static constexpr int G = 32 / LPR;        // row groups per warp
  static constexpr int WR = G * RPG;        // rows per warp-tile
  static constexpr int CAP = WR * 10;       // staged column indices per warp-tile (excess: global loads)
  static constexpr int BPS = WR + 4;        // ints per row-pointer buffer (WR + 1 used, +1 read past the end)
  static constexpr int WARP_INTS = 3 * BPS + 2 * CAP;
  static constexpr size_t kSmem = (size_t)kWarps * WARP_INTS * sizeof(int);
};

}  // namespace

// LPR lanes per row (C = 16 LPR); RPG rows per row group per warp-tile; BLK = 4: BSR4 (16 values per entry, rotated
// column-major, see sn_csr32_to_bsr4_fill), BLK = 1: CSR (one value per entry); PD = prefetch depth (entries in flight
// per row group); MINB = CTAs per SM the register allocation is tuned for.
template <int LPR, int RPG, int BLK, bool ELU, int PD, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
rowgroup_spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                     const float* __restrict__ val, const float* __restrict__ X, uint32_t ldxb,
                     float* __restrict__ Y, uint32_t ldyb, int n_rows, int n_wtiles) {
  using Gm = Geo<LPR, RPG>;
  constexpr int C = 16 * LPR;
  constexpr int kQuarterBytes = C;             // (C/4 floats) * 4 bytes
  constexpr int WR = Gm::WR, CAP = Gm::CAP, BPS = Gm::BPS;
  constexpr int NW = BLK == 4 ? 4 : 1;         // float4 weight registers per entry
  extern __shared__ int smem_i[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* bp_buf = smem_i + warp * Gm::WARP_INTS;   // [3][BPS]
  int* bc_buf = bp_buf + 3 * BPS;                // [2][CAP]
  const int g = lane / LPR, t = lane % LPR;
  const char* Xl = reinterpret_cast<const char*>(X) + t * 16;   // this lane's float4 of quarter 0 of every row
  char* Yl = reinterpret_cast<char*>(Y) + t * 16;
  const char* vbase = reinterpret_cast<const char*>(val);
  const int wstride = gridDim.x * kWarps;
  int wt = blockIdx.x * kWarps + warp;
  if (wt >= n_wtiles) return;                    // warps never synchronise with each other

  auto prefetch_bp = [&](int tile, int buf) {    // row pointers of warp-tile `tile`
    if (tile < n_wtiles) {
      const int r0 = tile * WR;
      for (int i = lane; i <= WR; i += 32)
        cp_async4(bp_buf + buf * BPS + i, rowptr + min(r0 + i, n_rows));  // past the end: empty rows
    }
  };
  auto prefetch_bc = [&](int tile, int bpb, int buf) {   // its column indices (needs its row pointers in smem)
    if (tile < n_wtiles) {
      const int k0 = bp_buf[bpb * BPS], k1 = bp_buf[bpb * BPS + WR];
      const int n = min(k1 - k0, CAP);
      for (int i = lane; i < n; i += 32) cp_async4(bc_buf + buf * CAP + i, colind + k0 + i);
    }
  };

  prefetch_bp(wt, 0);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  prefetch_bc(wt, 0, 0);
  prefetch_bp(wt + wstride, 1);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();

  float4 xs[PD][4], ws[PD][NW];
#pragma unroll
  for (int s = 0; s < PD; ++s) {
#pragma unroll
    for (int q = 0; q < 4; ++q) xs[s][q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < NW; ++q) ws[s][q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  int b3 = 0, b2 = 0;    // ring positions: row-pointer buffer (mod 3) and column-index buffer (mod 2) of this tile
  for (; wt < n_wtiles; wt += wstride) {
    const int* bp = bp_buf + b3 * BPS;
    const int* bc = bc_buf + b2 * CAP;
    const int b3n = b3 == 2 ? 0 : b3 + 1, b3nn = b3n == 2 ? 0 : b3n + 1;
    // indices of the next two warp-tiles travel while this one is computed
    prefetch_bc(wt + wstride, b3n, b2 ^ 1);
    prefetch_bp(wt + 2 * wstride, b3nn);
    cp_async_commit();

    const int k0 = bp[0];
    const int rl0 = g * RPG;                                // the group's first row inside the warp-tile
    const int kend = bp[rl0 + RPG];
    int k = bp[rl0];                                        // next entry to accumulate
    const int n_iter = __reduce_max_sync(0xffffffffu, kend - k);
    int r = 0;                                              // row inside the group
    int next_end = bp[rl0 + 1];
    const uint32_t grow0 = (uint32_t)wt * WR + rl0;

    float4 acc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);

    // store every row that ends at entry k (the current one, then the empty rows behind it)
    auto flush = [&]() {
      while (r < RPG && k == next_end) {
        const uint32_t grow = grow0 + r;
        if (grow < (uint32_t)n_rows) {
          char* yrow = const_cast<char*>(ptr_mad(Yl, grow, ldyb));
#pragma unroll
          for (int p = 0; p < 4; ++p) st_stream_f4(reinterpret_cast<float*>(yrow + p * kQuarterBytes), acc[p]);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        ++r;
        next_end = bp[rl0 + r + 1];
      }
    };
    auto load = [&](float4 (&x)[4], float4 (&w)[NW], int kk) {
      if (kk < kend) {
        const int rel = kk - k0;
        const int j = rel < CAP ? bc[rel] : __ldg(colind + kk);
        const char* xp = ptr_mad(Xl, (uint32_t)j, ldxb);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(reinterpret_cast<const float4*>(xp + q * kQuarterBytes));
        if (BLK == 4) {
          const char* wp = ptr_mad(vbase, (uint32_t)kk, 64u);
#pragma unroll
          for (int q = 0; q < NW; ++q) w[q] = __ldg(reinterpret_cast<const float4*>(wp + q * 16));
        } else {
          w[0].x = __ldg(reinterpret_cast<const float*>(ptr_mad(vbase, (uint32_t)kk, 4u)));
        }
      }
    };
    auto compute = [&](float4 (&x)[4], float4 (&w)[NW]) {
      // finished lanes (k == kend) run the FMAs on stale registers; their accumulators are never stored
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 xv = ELU ? elu4(x[q]) : x[q];
        if (BLK == 4) {       // w[q] = (B[q][q], B[q+1][q], B[q+2][q], B[q+3][q]), rows mod 4
          acc[q] = fma4(w[q].x, xv, acc[q]);
          acc[(q + 1) & 3] = fma4(w[q].y, xv, acc[(q + 1) & 3]);
          acc[(q + 2) & 3] = fma4(w[q].z, xv, acc[(q + 2) & 3]);
          acc[(q + 3) & 3] = fma4(w[q].w, xv, acc[(q + 3) & 3]);
        } else {
          acc[q] = fma4(w[0].x, xv, acc[q]);
        }
      }
      if (k < kend) {
        ++k;
        flush();
      }
    };

    flush();                                               // leading empty rows
#pragma unroll
    for (int s = 0; s < PD; ++s) load(xs[s], ws[s], k + s);
    for (int i = 0; i < n_iter; i += PD) {
#pragma unroll
      for (int s = 0; s < PD; ++s) {
        compute(xs[s], ws[s]);
        load(xs[s], ws[s], k + PD - 1);
      }
    }

    cp_async_wait_all();          // this lane's share of the index prefetch has landed ...
    __syncwarp();                 // ... and the other lanes'; everyone is done with this tile's bp / bc
    b3 = b3n;
    b2 ^= 1;
  }
}

namespace {

template <int LPR, int RPG, int BLK, int PD, int MINB>
int launch_rg(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx, float* Y,
              int64_t ldy, int64_t n_rows, bool elu, cudaStream_t st) {
  using Gm = Geo<LPR, RPG>;
  auto kern = elu ? rowgroup_spmm_kernel<LPR, RPG, BLK, true, PD, MINB>
                  : rowgroup_spmm_kernel<LPR, RPG, BLK, false, PD, MINB>;
  cudaError_t e;
  if (Gm::kSmem > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm::kSmem);
    if (e != cudaSuccess) return (int)e;
  }
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, Gm::kSmem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return SN_ERR_UNSUPPORTED;
  const int64_t n_wtiles = ceil_div(n_rows, Gm::WR);
  const int64_t ctas = ceil_div(n_wtiles, kWarps);
  const int64_t grid = ctas < (int64_t)sms * per_sm ? ctas : (int64_t)sms * per_sm;
  kern<<<(unsigned)grid, kThreads, Gm::kSmem, st>>>(rowptr, colind, val, X, (uint32_t)(ldx * 4), Y,
                                                    (uint32_t)(ldy * 4), (int)n_rows, (int)n_wtiles);
  return launch_status();
}

// Small operators get short warp-tiles (more warps busy), large ones long tiles (index staging amortised).
template <int LPR, int BLK, int PD, int MINB>
int launch_lpr(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx, float* Y,
               int64_t ldy, int64_t n_rows, bool elu, int tile_mode, cudaStream_t st) {
  constexpr int G = 32 / LPR;
  constexpr int RS = G >= 4 ? 1 : 4 / G;      // short tile: >= 4 rows per warp
  constexpr int RL = 4 * RS;                  // long tile: >= 16 rows per warp
  const int64_t long_tiles = n_rows / (G * RL);
  const bool use_long = tile_mode == 2 || (tile_mode == 0 && long_tiles >= 2 * 148 * kWarps * 2);
  if (use_long)
    return launch_rg<LPR, RL, BLK, PD, MINB>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, elu, st);
  return launch_rg<LPR, RS, BLK, PD, MINB>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, elu, st);
}

template <int BLK>
int launch_family(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                  float* Y, int64_t ldy, int64_t n_rows, int64_t C, bool elu, int variant, cudaStream_t st) {
  // ldx / ldy in bytes and entry offsets (64 B per block) must fit 32 bits
  if (n_rows >= 0x7fffff00LL || ldx >= (1LL << 30) || ldy >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  const int tile_mode = variant == 4 ? 1 : variant == 5 ? 2 : 0;   // 1: force short warp-tiles, 2: force long
#define SN_RG(LPR, PD, MINB) launch_lpr<LPR, BLK, PD, MINB>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, elu, tile_mode, st)
  switch (C) {
    case 16: return SN_RG(1, 2, 2);
    case 32: return SN_RG(2, 2, 2);
    case 64: return SN_RG(4, 2, 2);
    case 128:
      switch (variant) {           // tuning variants (tools/spmm_bench.py --variants rg1,rg2; rg4 / rg5 force short / long tiles)
        case 1: return SN_RG(8, 1, 3);
        case 2: return SN_RG(8, 3, 2);
        default: return SN_RG(8, 2, 2);
      }
    case 256:
      switch (variant) {
        case 1: return SN_RG(16, 1, 3);
        case 2: return SN_RG(16, 3, 2);
        default: return SN_RG(16, 2, 2);
      }
    case 512: return SN_RG(32, 2, 2);
    default: return SN_ERR_UNSUPPORTED;
  }
#undef SN_RG
}

}  // namespace

// Both return SN_ERR_UNSUPPORTED when the kernel does not apply (C not in {16,...,512}); callers fall back.
int launch_bsr4_rowgroup(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                         int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, bool elu, int variant,
                         cudaStream_t st) {
  return launch_family<4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, elu, variant, st);
}
int launch_csr_rowgroup(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                        float* Y, int64_t ldy, int64_t n_rows, int64_t C, bool elu, int variant, cudaStream_t st) {
  return launch_family<1>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C, elu, variant, st);
}

}  // namespace sn
