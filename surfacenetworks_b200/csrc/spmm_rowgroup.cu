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
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// base + a * b with 32-bit a, b and a 64-bit base: one IMAD.WIDE.U32
__device__ __forceinline__ const char* ptr_mad(const char* base, uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const char*>(r);
}

template <int LPR, int RPG, int BLK, int PD, int EPS, bool XS = false>
struct Geo {
  static constexpr int G = 32 / LPR;        // row groups per warp
  static constexpr int WR = G * RPG;        // rows per warp-tile
  static constexpr int CAP = WR * 10;       // staged column indices (and CSR values) per warp-tile (excess: global loads)
  static constexpr int BPS = WR + 4;        // ints per row-pointer buffer (WR + 1 used, +1 read past the end)
  static constexpr int WRING = BLK == 4 ? PD * G * 16 * EPS : 0;   // floats: PD stages x G groups x EPS blocks
  // XS: the gathered rows of the PD stages in flight land in shared memory ([stage][entry][quarter][lane] x 16 bytes, every
  // lane reads back exactly the pieces it requested) instead of registers
  static constexpr int XZONE = XS ? PD * EPS * 4 * 32 * 4 : 0;
  static constexpr int WARP_WORDS = ((XZONE + WRING + 3 * BPS + 2 * CAP + (BLK == 1 ? 2 * CAP : 0)) + 3) / 4 * 4;
  static constexpr size_t kSmem = (size_t)kWarps * WARP_WORDS * 4;
  // MODE 3 (statistics in the store path): one [2][C] accumulator per (warp, row group) behind the per-warp regions
  static constexpr size_t kStatSmem = (size_t)kWarps * G * 2 * (16 * LPR) * 4;
  // MODE 4 (epilogue operands staged): per lane 3 operands x 4 quarters x 16 bytes, laid out [operand][quarter][lane]
  // (mode 4 + n: n = 1 ... 3 operands present -- the zone is sized by what the launch uses: 16 KB per operand and CTA)
  static constexpr size_t kOperandSmem = (size_t)kWarps * 4 * 32 * 16;
  static constexpr size_t smem_bytes(int mode) {
    return kSmem + (mode == 3 ? kStatSmem : 0) + (mode >= 5 ? (size_t)(mode - 4) * kOperandSmem : 0);
  }
};

// Optional output epilogue Y = (S X + G) .* elu'(A) + G2 (backward of "elu, then gather": the activation derivative and the
// gradient of the un-gathered half are applied where the row is stored instead of in a separate pass).
struct Epilogue {
  const float* G;     // [n_rows x C], leading dimension ldgb bytes, or null
  const float* A;     // activated values elu(x) [n_rows x C], leading dimension ldab bytes, or null
  const float* G2;    // [n_rows x C] added AFTER the derivative (a residual-path gradient), or null
  uint32_t ldgb, ldab, ldg2b;
  // MODE 3: per-CTA column sums / sums of squares of Y [gridDim.x][2][C] -- the BatchNorm statistics of the right half of
  // the stage's concat buffer, taken where the rows are stored instead of by a second pass over Y (sn_colstats_f32)
  float* stat_partial;
  int* host_grid_out; // HOST pointer (never dereferenced on the device): receives the grid size = number of partial rows
};

}  // namespace

// LPR lanes per row (C = 16 LPR); RPG rows per row group per warp-tile; BLK = 4: BSR4 (16 values per entry, rotated
// column-major, see sn_csr32_to_bsr4_fill), BLK = 1: CSR (one value per entry); PD = pipeline depth in stages of EPS
// consecutive entries per row group; MINB = CTAs per SM the register allocation is tuned for.
template <int LPR, int RPG, int BLK, int MODE, int PD, int EPS, int MINB, bool XS = false>
__global__ void __launch_bounds__(kThreads, MINB)
rowgroup_spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                     const float* __restrict__ val, const float* __restrict__ X, uint32_t ldxb,
                     float* __restrict__ Y, uint32_t ldyb, int n_rows, int n_wtiles, const Epilogue epi) {
  using Gm = Geo<LPR, RPG, BLK, PD, EPS, XS>;
  constexpr bool ELU = MODE == 1;      // ELU on the gathered operand
  constexpr bool EPI = MODE == 2 || MODE == 4;   // output epilogue (G, elu', G2)
  // MODE 4: the epilogue operands of a row are requested when the row STARTS -- cp.async into a per-lane landing zone in
  // shared memory (no registers held; every lane copies exactly the 16-byte pieces it will read, so no synchronisation) --
  // instead of by loads at the row's end, where their latency sat behind the whole gather chain of the row
  constexpr bool STAGED = MODE == 4;
  constexpr bool STATS = MODE == 3;    // column statistics of the output in the store path
  constexpr int C = 16 * LPR;
  constexpr int kQuarterBytes = C;             // (C/4 floats) * 4 bytes
  constexpr int G = Gm::G, WR = Gm::WR, CAP = Gm::CAP, BPS = Gm::BPS;
  extern __shared__ __align__(16) int smem_i[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* wbase = smem_i + warp * Gm::WARP_WORDS;
  float4* xzone = reinterpret_cast<float4*>(wbase) + lane;         // XS: [PD][EPS][4][32] float4, this lane's column
  float* wring = reinterpret_cast<float*>(wbase + Gm::XZONE);      // [PD][G][16 EPS]   (BSR4 only; 16-byte aligned)
  int* bp_buf = wbase + Gm::XZONE + Gm::WRING;           // [3][BPS]
  int* bc_buf = bp_buf + 3 * BPS;                        // [2][CAP]
  float* bv_buf = reinterpret_cast<float*>(bc_buf + 2 * CAP);   // [2][CAP]  (CSR only)
  const int g = lane / LPR, t = lane % LPR;
  const char* Xl = reinterpret_cast<const char*>(X) + t * 16;   // this lane's float4 of quarter 0 of every row
  char* Yl = reinterpret_cast<char*>(Y) + t * 16;
  const char* vbase = reinterpret_cast<const char*>(val);
  constexpr int kSlot = G * 16 * EPS;                    // floats per stage of the value ring
  const float* wslot0 = wring + g * 16 * EPS;            // this group's blocks in stage 0
  const int wstride = gridDim.x * kWarps;
  int wt = blockIdx.x * kWarps + warp;
  if (!STATS && wt >= n_wtiles) return;          // warps never synchronise with each other (STATS: once, at the very end)
  // STATS: this (warp, row group)'s accumulator; lane t owns columns q C/4 + 4t .. +3 (q = 0..3) of both statistics
  float* wstat = reinterpret_cast<float*>(smem_i + kWarps * Gm::WARP_WORDS) + (warp * G + g) * 2 * C + 4 * t;
  // STAGED: this lane's landing zone; piece (operand o, quarter q) at ops_s + ((4 o + q) * 32) float4s (conflict-free)
  const int n_ops = (epi.G != nullptr) + (epi.A != nullptr) + (epi.G2 != nullptr);
  float4* ops_s = reinterpret_cast<float4*>(smem_i + kWarps * Gm::WARP_WORDS) + warp * (n_ops * 4 * 32) + lane;
  const int slot_a = epi.G != nullptr ? 4 * 32 : 0, slot_g2 = slot_a + (epi.A != nullptr ? 4 * 32 : 0);   // G: slot 0
  auto stage_operands = [&](uint32_t grow) {
    if (grow < (uint32_t)n_rows) {
      const char* src[3] = {reinterpret_cast<const char*>(epi.G), reinterpret_cast<const char*>(epi.A),
                            reinterpret_cast<const char*>(epi.G2)};
      const uint32_t ld[3] = {epi.ldgb, epi.ldab, epi.ldg2b};
      const int slot[3] = {0, slot_a, slot_g2};
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        if (src[o] != nullptr) {
          const char* rp = ptr_mad(src[o] + t * 16, grow, ld[o]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ops_s + slot[o] + q * 32)),
                         "l"(rp + q * kQuarterBytes) : "memory");
        }
      }
    }
    cp_async_commit();
  };
  if (STATS) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      *reinterpret_cast<float4*>(wstat + p * (C / 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(wstat + C + p * (C / 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

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
      for (int i = lane; i < n; i += 32) {
        cp_async4(bc_buf + buf * CAP + i, colind + k0 + i);
        if (BLK == 1) cp_async4(bv_buf + buf * CAP + i, val + k0 + i);
      }
    }
  };

  prefetch_bp(wt, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();
  prefetch_bc(wt, 0, 0);
  prefetch_bp(wt + wstride, 1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncwarp();

  float4 xs[PD][4 * EPS];  // gathered rows of the EPS entries of each stage in flight
  float wv[PD][EPS];       // CSR: their values
#pragma unroll
  for (int s = 0; s < PD; ++s) {
#pragma unroll
    for (int q = 0; q < 4 * EPS; ++q) xs[s][q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < EPS; ++e) wv[s][e] = 0.f;
  }

  int b3 = 0, b2 = 0;    // ring positions: row-pointer buffer (mod 3) and column-index buffer (mod 2) of this tile
  for (; wt < n_wtiles; wt += wstride) {
    const int* bp = bp_buf + b3 * BPS;
    const int* bc = bc_buf + b2 * CAP;
    const float* bv = bv_buf + b2 * CAP;
    const int b3n = b3 == 2 ? 0 : b3 + 1, b3nn = b3n == 2 ? 0 : b3n + 1;
    // indices of the next two warp-tiles travel while this one is computed
    prefetch_bc(wt + wstride, b3n, b2 ^ 1);
    prefetch_bp(wt + 2 * wstride, b3nn);
    cp_async_commit();

    const int k0 = bp[0];
    const int rl0 = g * RPG;                                // the group's first row inside the warp-tile
    const int kend = bp[rl0 + RPG];
    int k = bp[rl0];                                        // next entry to accumulate
    const int n_iter = __reduce_max_sync(0xffffffffu, (kend - k + EPS - 1) / EPS);   // stages of EPS entries
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
          if (STAGED) cp_async_wait<0>();       // this lane's operand pieces (requested at the row's start) have landed
          if (EPI && epi.G != nullptr) {        // Y = S X + G
            const char* grow_p = ptr_mad(reinterpret_cast<const char*>(epi.G) + t * 16, grow, epi.ldgb);
#pragma unroll
            for (int p = 0; p < 4; ++p)
              acc[p] = add4(acc[p], STAGED ? ops_s[p * 32] : __ldg(reinterpret_cast<const float4*>(grow_p + p * kQuarterBytes)));
          }
          // (fetching this operand when the row STARTS, 16 more registers and 3 CTAs/SM, was measured slower: 104 / 92 us
          // against 96 / 89 us for D^T / (D*)^T at the cfg3 size; a register-free prefetch.global.L2 of the three operand
          // rows one row ahead changed nothing either -- 130.7 / 113.6 us against 130.5 / 107.0 us in the step, round 2)
          if (EPI && epi.A != nullptr) {        // ... times elu'(x) taken from the activated values a = elu(x): 1 or a + 1
            const char* arow_p = ptr_mad(reinterpret_cast<const char*>(epi.A) + t * 16, grow, epi.ldab);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const float4 a = STAGED ? ops_s[slot_a + p * 32] : __ldg(reinterpret_cast<const float4*>(arow_p + p * kQuarterBytes));
              acc[p].x *= a.x > 0.f ? 1.f : a.x + 1.f;
              acc[p].y *= a.y > 0.f ? 1.f : a.y + 1.f;
              acc[p].z *= a.z > 0.f ? 1.f : a.z + 1.f;
              acc[p].w *= a.w > 0.f ? 1.f : a.w + 1.f;
            }
          }
          if (EPI && epi.G2 != nullptr) {       // ... + G2
            const char* g2row_p = ptr_mad(reinterpret_cast<const char*>(epi.G2) + t * 16, grow, epi.ldg2b);
#pragma unroll
            for (int p = 0; p < 4; ++p)
              acc[p] = add4(acc[p], STAGED ? ops_s[slot_g2 + p * 32] : __ldg(reinterpret_cast<const float4*>(g2row_p + p * kQuarterBytes)));
          }
#pragma unroll
          for (int p = 0; p < 4; ++p) st_stream_f4(reinterpret_cast<float*>(yrow + p * kQuarterBytes), acc[p]);
          if (STATS) {                          // rows in the order this group stores them: fixed, run to run
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              float4* sp = reinterpret_cast<float4*>(wstat + p * (C / 4));
              float4* qp = reinterpret_cast<float4*>(wstat + C + p * (C / 4));
              const float4 a = acc[p];
              float4 sq = *qp;
              sq.x = fmaf(a.x, a.x, sq.x); sq.y = fmaf(a.y, a.y, sq.y); sq.z = fmaf(a.z, a.z, sq.z); sq.w = fmaf(a.w, a.w, sq.w);
              *sp = add4(*sp, a);
              *qp = sq;
            }
          }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        ++r;
        next_end = bp[rl0 + r + 1];
        if (STAGED && r < RPG) stage_operands(grow0 + r);   // the next row of this group: its operands travel with its gathers
      }
    };
    // stage s <- entries kk .. kk + EPS - 1 of this group's run: X rows to registers; BSR4 values (64 bytes per
    // entry) to the shared-memory ring with 16-byte cp.async, CSR values from the staged tile ring
    auto load = [&](const int s, float4 (&x)[4 * EPS], float (&w)[EPS], int kk) {
#pragma unroll
      for (int e = 0; e < EPS; ++e) {
        if (kk + e < kend) {
          const int rel = kk + e - k0;
          const int j = rel < CAP ? bc[rel] : __ldg(colind + kk + e);
          const char* xp = ptr_mad(Xl, (uint32_t)j, ldxb);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (XS)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(xzone + ((s * EPS + e) * 4 + q) * 32)),
                           "l"(xp + q * kQuarterBytes) : "memory");
            else
              x[4 * e + q] = __ldg(reinterpret_cast<const float4*>(xp + q * kQuarterBytes));
          }
          if (BLK == 1) w[e] = rel < CAP ? bv[rel] : __ldg(val + kk + e);
        }
      }
      if (BLK == 4) {
        const uint32_t dst = smem_u32(wslot0 + s * kSlot);
        const char* src = ptr_mad(vbase, (uint32_t)kk, 64u);
#pragma unroll
        for (int u0 = 0; u0 < 4 * EPS; u0 += LPR) {
          const int u = u0 + t;                             // 16-byte unit of the stage this lane copies
          if (u < 4 * EPS && kk + (u >> 2) < kend)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + u * 16), "l"(src + u * 16)
                         : "memory");
        }
        cp_async_commit();                                  // every lane, every call: uniform group counting
      } else if (XS) {
        cp_async_commit();
      }
    };
    auto fma_entry = [&](const float4* x, const float* wsm, float wscalar) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 xv = ELU ? elu4(x[q]) : x[q];
        if (BLK == 4) {       // w = (B[q][q], B[q+1][q], B[q+2][q], B[q+3][q]), rows mod 4
          const float4 w = *reinterpret_cast<const float4*>(wsm + 4 * q);
          acc[q] = fma4(w.x, xv, acc[q]);
          acc[(q + 1) & 3] = fma4(w.y, xv, acc[(q + 1) & 3]);
          acc[(q + 2) & 3] = fma4(w.z, xv, acc[(q + 2) & 3]);
          acc[(q + 3) & 3] = fma4(w.w, xv, acc[(q + 3) & 3]);
        } else {
          acc[q] = fma4(wscalar, xv, acc[q]);
        }
      }
    };
    auto compute = [&](const int s, float4 (&x)[4 * EPS], const float (&w)[EPS]) {
      const float* wsm = wslot0 + s * kSlot;
      if (BLK == 4 || XS) cp_async_wait<PD - 1>();      // this lane's share of stage s has landed ...
      if (BLK == 4) __syncwarp();                       // ... and the other lanes' shares (the blocks' values)
#pragma unroll
      for (int e = 0; e < EPS; ++e) {
        if (k < kend) {               // a run that ends inside the stage skips the remaining entries
          if (XS) {
#pragma unroll
            for (int q = 0; q < 4; ++q) x[4 * e + q] = xzone[((s * EPS + e) * 4 + q) * 32];
          }
          fma_entry(x + 4 * e, wsm + 16 * e, w[e]);
          ++k;
          flush();
        }
      }
      if (BLK == 4) __syncwarp();     // everyone has read stage s before its slot is refilled
    };

    if (STAGED) stage_operands(grow0);
    flush();                                               // leading empty rows
#pragma unroll
    for (int s = 0; s < PD; ++s) load(s, xs[s], wv[s], k + EPS * s);
    for (int i = 0; i < n_iter; i += PD) {
#pragma unroll
      for (int s = 0; s < PD; ++s) {
        compute(s, xs[s], wv[s]);
        load(s, xs[s], wv[s], k + EPS * (PD - 1));
      }
    }

    cp_async_wait<0>();           // this lane's share of the index prefetch has landed ...
    __syncwarp();                 // ... and the other lanes'; everyone is done with this tile's bp / bc
    b3 = b3n;
    b2 ^= 1;
  }
  if (STATS) {
    // the CTA's partial: its kWarps * G accumulators added in a fixed order (idle warps hold zeros)
    cp_async_wait<0>();
    __syncthreads();
    const float* all = reinterpret_cast<const float*>(smem_i + kWarps * Gm::WARP_WORDS);
    for (int i = threadIdx.x; i < 2 * C; i += kThreads) {
      float a = 0.f;
      for (int w = 0; w < kWarps * G; ++w) a += all[w * 2 * C + i];
      epi.stat_partial[(size_t)blockIdx.x * 2 * C + i] = a;
    }
  }
}

namespace {

struct DeviceInfo { int sms; };
inline DeviceInfo device_info() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return {sms};
}

template <int LPR, int RPG, int BLK, int PD, int EPS, int MINB, bool XS = false>
struct Launcher {
  using Gm = Geo<LPR, RPG, BLK, PD, EPS, XS>;
  using Kernel = void (*)(const int32_t*, const int32_t*, const float*, const float*, uint32_t, float*, uint32_t, int, int,
                          const Epilogue);
  // Which (pipeline shape, store-path mode) pairs exist.  The store-path modes (epilogue, statistics) are built for the
  // default shape -- one stage of one entry, four CTAs per SM -- only; the tuning shapes exist for the plain product,
  // the shared-memory gather variant also with the statistics.  (Everything else would be ~300 more kernels to compile.)
  template <int MODE>
  static constexpr bool exists() {
    return (PD == 1 && EPS == 1 && !XS) || MODE <= 1 || (XS && MODE == 3);
  }
  template <int MODE>
  static Kernel kernel_or_null() {
    if constexpr (exists<MODE>()) return rowgroup_spmm_kernel<LPR, RPG, BLK, MODE, PD, EPS, MINB, XS>;
    else return nullptr;
  }
  static Kernel pick(int mode) {      // runtime mode: 0 .. 3 as the template's, 5 .. 7 = staged epilogue with 1 .. 3 operands
    return mode == 1 ? kernel_or_null<1>() : mode == 2 ? kernel_or_null<2>() : mode == 3 ? kernel_or_null<3>()
           : mode >= 5 ? kernel_or_null<4>() : kernel_or_null<0>();
  }
  // persistent warps resident on the device for this instantiation (0: kernel cannot run)
  static int64_t resident_warps(int mode, int sms) {
    // occupancy is a property of (kernel, device model): queried once per process and device ordinal (a benign race:
    // concurrent first calls compute the same value)
    static int64_t cached[8][32] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const bool cacheable = dev >= 0 && dev < 32;
    if (cacheable && cached[mode][dev] != 0) return cached[mode][dev] < 0 ? 0 : cached[mode][dev];
    const int64_t w = query_resident_warps(mode, sms);
    if (cacheable) cached[mode][dev] = w > 0 ? w : -1;
    return w;
  }
  static int64_t query_resident_warps(int mode, int sms) {
    const Kernel kern = pick(mode);
    if (kern == nullptr) return 0;
    const size_t smem = Gm::smem_bytes(mode);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return (int64_t)sms * per_sm * kWarps;
  }
  // fraction of the resident warps' time spent on tiles when every warp walks ceil(tiles / warps) of them
  static double efficiency(int64_t n_rows, int64_t warps) {
    if (warps <= 0) return 0.0;
    const int64_t tiles = ceil_div(n_rows, Gm::WR);
    const int64_t w = tiles < warps ? tiles : warps;
    return (double)tiles / (double)(ceil_div(tiles, w) * warps);
  }
  static int launch(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                    float* Y, int64_t ldy, int64_t n_rows, int mode, int64_t warps, const Epilogue& epi,
                    cudaStream_t st) {
    if (warps <= 0) return SN_ERR_UNSUPPORTED;
    const Kernel kern = pick(mode);
    if (kern == nullptr) return SN_ERR_UNSUPPORTED;
    const int64_t n_wtiles = ceil_div(n_rows, Gm::WR);
    const int64_t ctas = ceil_div(n_wtiles, kWarps);
    const int64_t grid = ctas < warps / kWarps ? ctas : warps / kWarps;
    if (mode == 3 && epi.host_grid_out) *epi.host_grid_out = (int)grid;
    kern<<<(unsigned)grid, kThreads, Gm::smem_bytes(mode), st>>>(rowptr, colind, val, X, (uint32_t)(ldx * 4), Y,
                                                                 (uint32_t)(ldy * 4), (int)n_rows, (int)n_wtiles, epi);
    return launch_status();
  }
};

// Warp-tile length: long tiles amortise the index staging and the pipeline fill per tile, short tiles keep every
// persistent warp busy on small operators and shrink the last-wave quantisation (each warp walks an integer number of
// tiles).  Pick the longest of {4 RS, 2 RS, RS} rows per group whose quantisation efficiency is >= 0.93, else the
// most efficient one.  tile_mode 1 / 2 / 3 force short / medium / long (benchmarks).
template <int LPR, int BLK, int PD, int EPS, int MINB, bool XS = false>
int launch_lpr(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx, float* Y,
               int64_t ldy, int64_t n_rows, int mode, int tile_mode, const Epilogue& epi, cudaStream_t st) {
  constexpr int G = 32 / LPR;
  constexpr int RS = G >= 4 ? 1 : 4 / G;      // short tile: >= 4 rows per warp
  using LS = Launcher<LPR, RS, BLK, PD, EPS, MINB, XS>;
  using LM = Launcher<LPR, 2 * RS, BLK, PD, EPS, MINB, XS>;
  using LL = Launcher<LPR, 4 * RS, BLK, PD, EPS, MINB, XS>;
  const int sms = device_info().sms;
  const int64_t ws = LS::resident_warps(mode, sms), wm = LM::resident_warps(mode, sms), wl = LL::resident_warps(mode, sms);
  int pick = tile_mode;
  // small operators at C >= 256 (one mesh of a few thousand rows, BASELINE cfg5): a single row per row group, so the
  // serial chain of a warp is one row's entries instead of RS rows'
  if ((pick < 1 || pick > 3) && RS > 1 && ceil_div(n_rows, G * RS) < ws) {
    using LT = Launcher<LPR, 1, BLK, PD, EPS, MINB, XS>;
    return LT::launch(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, LT::resident_warps(mode, sms), epi, st);
  }
  if ((pick < 1 || pick > 3) && LPR == 32) pick = 1;   // C = 512: one row per warp pass; short tiles measured fastest
  if (pick < 1 || pick > 3) {
    const double es = LS::efficiency(n_rows, ws), em = LM::efficiency(n_rows, wm), el = LL::efficiency(n_rows, wl);
    // a tile length only qualifies when it gives every resident warp at least one tile
    const bool ql = wl > 0 && ceil_div(n_rows, G * 4 * RS) >= wl, qm = wm > 0 && ceil_div(n_rows, G * 2 * RS) >= wm;
    if (ql && el >= 0.93) pick = 3;
    else if (qm && em >= 0.93) pick = 2;
    else if (ql && el >= em && el >= es) pick = 3;
    else if (qm && em >= es) pick = 2;
    else pick = 1;
  }
  if (pick == 3) return LL::launch(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, wl, epi, st);
  if (pick == 2) return LM::launch(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, wm, epi, st);
  return LS::launch(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, ws, epi, st);
}

// Variant 9: two gathers in flight per row group, landing in shared memory -- built for the block operators (the scalar
// Laplacian was measured slower with it: 47.3 -> 49.9 us) and for the plain product and the statistics mode.
template <int LPR, int BLK>
int launch_xs(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx, float* Y,
              int64_t ldy, int64_t n_rows, int mode, int tile_mode, const Epilogue& epi, cudaStream_t st) {
  if constexpr (BLK == 4) {
    if (mode <= 1 || mode == 3)
      return launch_lpr<LPR, BLK, 2, 1, 4, true>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, tile_mode, epi, st);
  }
  return launch_lpr<LPR, BLK, 1, 1, 4>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, tile_mode, epi, st);
}

template <int BLK>
int launch_family(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                  float* Y, int64_t ldy, int64_t n_rows, int64_t C, bool elu, int variant, const Epilogue& epi,
                  cudaStream_t st) {
  // epilogue: operands staged through shared memory (4) unless variant 8 asks for the loads at the row's end (2, A/B runs)
  const bool has_epi = epi.G || epi.A || epi.G2;
  const int n_ops = (epi.G != nullptr) + (epi.A != nullptr) + (epi.G2 != nullptr);
  const int mode0 = has_epi ? (variant == 8 ? 2 : 4 + n_ops) : (epi.stat_partial ? 3 : (elu ? 1 : 0));   // 5 .. 7: staged
  if (has_epi && variant == 8) variant = 0;
  // the epilogues exist for the default pipeline shape only; 9 (rows through shared memory) also carries the statistics
  if (mode0 >= 2 && (elu || (variant >= 4 && !(mode0 == 3 && variant == 9)))) return SN_ERR_UNSUPPORTED;
  if (has_epi && epi.stat_partial) return SN_ERR_UNSUPPORTED;
  // ldx / ldy in bytes and entry offsets (64 B per block) must fit 32 bits
  if (n_rows >= 0x7fffff00LL || ldx >= (1LL << 30) || ldy >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  // tuning variants (tools/spmm_bench.py --variants rgN): 1 / 2 / 3 force short / medium / long warp-tiles;
  // 5 / 9 (C = 128 / 256 only) change the pipeline shape (two gathers in flight: in registers at 3 CTAs per SM / through
  // shared memory at 4).  (Two entries per stage and three stages through shared memory were measured too and removed.)
  // Measured on B200 at 64 x 2000 V (profiles/r1_spmm_rowgroup_notes.md): occupancy beats per-warp prefetch depth --
  // (1, 1, 4) = 32 warps per SM with one entry in flight per row group is the fastest shape at every width.
  const int tile_mode = variant >= 1 && variant <= 3 ? variant : 0;
#define SN_RG(LPR, PD, EPS, MINB) \
  launch_lpr<LPR, BLK, PD, EPS, MINB>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, tile_mode, epi, st)
#define SN_RG_TUNE(LPR)                          \
  switch (variant) {                             \
    case 5: return SN_RG(LPR, 2, 1, 3);          /* two gathers in flight per row group, registers */ \
    case 9: return launch_xs<LPR, BLK>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, mode, tile_mode, epi, st); \
    default: return SN_RG(LPR, 1, 1, 4);         \
  }
  auto run = [&](const int mode) -> int {
    switch (C) {
      case 32: return SN_RG(2, 1, 1, 4);
      case 64: return SN_RG(4, 1, 1, 4);
      case 128: SN_RG_TUNE(8)
      case 256: SN_RG_TUNE(16)
      case 512: return SN_RG(32, 1, 1, 4);
      default: return SN_ERR_UNSUPPORTED;
    }
  };
  const int rc = run(mode0);
  // (the staged epilogue needs 16 KB more shared memory per operand and CTA: where that does not fit, the loads at the row's end)
  return (rc == SN_ERR_UNSUPPORTED && mode0 >= 5) ? run(2) : rc;
#undef SN_RG_TUNE
#undef SN_RG
}

}  // namespace

// Both return SN_ERR_UNSUPPORTED when the kernel does not apply (C not in {32,...,512}); callers fall back to the
// direct-gather kernels (at C = 16 a row is one 64-byte segment and a lane per row has nothing left to share).
// G / A / G2 (optional, null = none): output epilogue Y = (S X + G) .* elu'(A) + G2, leading dimensions in floats.
int launch_bsr4_rowgroup(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                         int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, int64_t C, bool elu, int variant,
                         const float* G, int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2,
                         cudaStream_t st, float* stat_partial, int* grid_out) {
  if (ldg >= (1LL << 30) || lda >= (1LL << 30) || ldg2 >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  const Epilogue epi{G, A, G2, (uint32_t)(ldg * 4), (uint32_t)(lda * 4), (uint32_t)(ldg2 * 4), stat_partial, grid_out};
  return launch_family<4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, C, elu, variant, epi, st);
}
int launch_csr_rowgroup(const int32_t* rowptr, const int32_t* colind, const float* val, const float* X, int64_t ldx,
                        float* Y, int64_t ldy, int64_t n_rows, int64_t C, bool elu, int variant, const float* G,
                        int64_t ldg, const float* A, int64_t lda, const float* G2, int64_t ldg2, cudaStream_t st,
                        float* stat_partial, int* grid_out) {
  if (ldg >= (1LL << 30) || lda >= (1LL << 30) || ldg2 >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  const Epilogue epi{G, A, G2, (uint32_t)(ldg * 4), (uint32_t)(lda * 4), (uint32_t)(ldg2 * 4), stat_partial, grid_out};
  return launch_family<1>(rowptr, colind, val, X, ldx, Y, ldy, n_rows, C, elu, variant, epi, st);
}

// upper bound of the grid any instantiation launches (at most 2048 / 256 = 8 CTAs per SM): rows of the statistics workspace
int64_t rowgroup_max_grid() { return (int64_t)device_info().sms * 8; }

}  // namespace sn
