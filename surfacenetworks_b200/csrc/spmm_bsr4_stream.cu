// spmm_bsr4_stream.cu -- Dirac / adjoint BSR4 SpMM, streaming kernel (the hot kernel for C = 128 / 256 / 512).
//
// Same arithmetic and summation order as spmm_bsr4.cu (reference src/utils/utils_pt.py:201-203,213-215); what
// changes is how the memory system is driven.  Measured on B200 (profiles/r1_bsr4_notes.md): the direct-gather
// kernel sits at 23% of the HBM roofline with every unit < 40% busy -- each block-row pays a chain of three
// dependent global loads (browptr -> bcolind -> X rows) with one row per warp in flight.  A TMA variant (one
// cp.async.bulk per gathered 512 B row) was slower still: the per-SM TMA unit retires one small bulk copy
// every ~80 cycles.  This kernel removes the dependent chain and keeps ~3 KB of gathers in flight per warp:
//
//   * persistent CTAs walk tiles of 128 block-rows; the tile's row pointers and column indices are staged in
//     shared memory, and the NEXT tile's are prefetched with cp.async while the current tile is computed, so
//     index latency is paid once per CTA, not once per row;
//   * each warp owns 16 consecutive block-rows = one contiguous run of blocks, which it streams in chunks of
//     GB blocks through a private double buffer: 16-byte cp.async (LDGSTS) gathers of the dense rows
//     X[j, :] (whole 128 B lines) + the chunk's contiguous block values; chunk c+1 is in flight while
//     chunk c is consumed -- no registers are tied up by loads in flight;
//   * math: lane (q, t) reads its float4 of X and the (rotated) block column q from shared memory (conflict-
//     free LDS.128), 16 FMAs per block into four slots; when a row's run of blocks ends, three rotating
//     shuffles add the slots of the other q-lanes (no per-lane register selection) and the lane streams out
//     one coalesced STG.128.  Rows of any length (and empty rows) fall out of the streaming formulation; no
//     per-operator hints are needed.
//
// Bound: HBM.  Algorithmic bytes per launch = 4(Rb+1) + 68 nb + 4 Cb C + 4 Rb C.
#include "common.cuh"

namespace sn {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTileRows = 128;              // block-rows per tile
constexpr int kRowsPerWarp = kTileRows / kWarps;
constexpr int kBcCap = 1024;                // staged column indices per tile (rest falls back to global loads)
constexpr int kSlots = 2;                   // per-warp double buffer

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// base + a * b with 32-bit a, b and a 64-bit base: one IMAD.WIDE.U32
__device__ __forceinline__ const char* ptr_mad(const char* base, uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(reinterpret_cast<uint64_t>(base)));
  return reinterpret_cast<const char*>(r);
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace

// NCH = C / 128 feature passes per row; GB = blocks per chunk.
template <int NCH, int GB, bool ELU>
__global__ void __launch_bounds__(kThreads)
bsr4_spmm_stream_kernel(const int32_t* __restrict__ browptr, const int32_t* __restrict__ bcolind,
                        const float* __restrict__ bval, const float* __restrict__ X, int64_t ldx,
                        float* __restrict__ Y, int64_t ldy, int n_brows, int n_tiles) {
  constexpr int C = 128 * NCH;
  constexpr int C4 = C / 4;
  constexpr int UPR = C / 4;                       // 16-byte units per dense row
  constexpr int kRowBytes = C * 4;
  constexpr int kSlotBytes = GB * (kRowBytes + 64);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* slots = smem_raw;                                   // [kWarps][kSlots][kSlotBytes]
  int* bc_buf = reinterpret_cast<int*>(smem_raw + kWarps * kSlots * kSlotBytes);  // [2][kBcCap]
  int* bp_buf = bc_buf + 2 * kBcCap;                                 // [3][kTileRows + 4]
  constexpr int kBpStride = kTileRows + 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 3, tl = lane & 7;
  const int lane_xoff = q * C4 + 4 * tl;                  // float offset of this lane's float4 inside a row
  const int src1 = (lane + 8) & 31, src2 = (lane + 16) & 31, src3 = (lane + 24) & 31;
  // 32-bit strides (bytes) keep the per-block address arithmetic to one IMAD.WIDE
  const uint32_t ldxb = (uint32_t)ldx * 4u, ldyb = (uint32_t)ldy * 4u;
  const char* Xl = reinterpret_cast<const char*>(X) + lane * 16;       // this lane's 16-byte unit of every row
  char* Yl = reinterpret_cast<char*>(Y + lane_xoff);
  const unsigned char* my_slots = slots + (size_t)warp * kSlots * kSlotBytes;
  const uint32_t slot_u32 = smem_u32(my_slots) + lane * 16;            // shared-memory destination of that unit

  // ---- index staging helpers (all threads of the CTA)
  auto prefetch_bp = [&](int tile, int buf) {   // row pointers of `tile` -> bp_buf[buf]
    if (tile < n_tiles) {
      const int r0 = tile * kTileRows;
      for (int i = threadIdx.x; i <= kTileRows; i += kThreads) {
        const int r = min(r0 + i, n_brows);     // rows past the end repeat the last pointer => empty rows
        cp_async4(bp_buf + buf * kBpStride + i, browptr + r);
      }
    }
  };
  auto prefetch_bc = [&](int tile, int bpb, int buf) {  // column indices of `tile` (needs its bp in smem)
    if (tile < n_tiles) {
      const int k0 = bp_buf[bpb * kBpStride], k1 = bp_buf[bpb * kBpStride + kTileRows];
      const int n = min(k1 - k0, kBcCap);
      for (int i = threadIdx.x; i < n; i += kThreads) cp_async4(bc_buf + buf * kBcCap + i, bcolind + k0 + i);
    }
  };

  int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  // prologue: bp(tile) ; then bc(tile) + bp(next)
  prefetch_bp(tile, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  prefetch_bc(tile, 0, 0);
  prefetch_bp(tile + gridDim.x, 1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
    const int* bp = bp_buf + (it % 3) * kBpStride;
    const int* bc = bc_buf + (it & 1) * kBcCap;
    // prefetch the indices of the next two tiles while this one is computed
    prefetch_bc(tile + gridDim.x, (it + 1) % 3, (it + 1) & 1);
    prefetch_bp(tile + 2 * gridDim.x, (it + 2) % 3);
    cp_async_commit();

    const int k0 = bp[0];
    const int rw0 = warp * kRowsPerWarp;                       // this warp's rows inside the tile
    const int grow0 = tile * kTileRows + rw0;                  // ... their global index
    const int kb = bp[rw0], ke = bp[rw0 + kRowsPerWarp];       // ... and its contiguous run of blocks
    const int nchunks = (ke - kb + GB - 1) / GB;
    // lane l < 16 keeps the end pointer of the warp's row l: row boundaries inside a chunk become a bit mask,
    // and so do empty rows (end pointer equal to the previous row's)
    const int my_re = lane < kRowsPerWarp ? bp[rw0 + lane + 1] : -1;
    const int up_re = __shfl_up_sync(0xffffffffu, my_re, 1);   // every lane takes part in the shuffle
    const int prev_re = lane == 0 ? kb : up_re;
    const unsigned emptymask = __ballot_sync(0xffffffffu, lane < kRowsPerWarp && my_re == prev_re);
    const char* vsrc_lane = reinterpret_cast<const char*>(bval) + lane * 16;

    auto issue_n = [&](int c, const int nn) {  // gather chunk c (nn blocks from kb + c*GB) into slot c & 1
      const int kc = kb + c * GB;
      const uint32_t dst = slot_u32 + (c & 1) * kSlotBytes;
      const int kk = kc - k0;
      int j[GB];
      if (kk + GB <= kBcCap) {                 // staged indices (the normal case)
#pragma unroll
        for (int b = 0; b < GB; ++b) j[b] = bc[kk + b];
      } else {
#pragma unroll
        for (int b = 0; b < GB; ++b) j[b] = b < nn ? __ldg(bcolind + kc + b) : 0;
      }
#pragma unroll
      for (int b = 0; b < GB; ++b) {
        if (b < nn) {
          const char* src = ptr_mad(Xl, (uint32_t)j[b], ldxb);
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + b * kRowBytes + ch * 512),
                         "l"(src + ch * 512)
                         : "memory");
        }
      }
      if (lane < nn * 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + GB * kRowBytes),
                     "l"(ptr_mad(vsrc_lane, (uint32_t)kc, 64u))
                     : "memory");
    };
    auto issue = [&](int c) {
      const int n = ke - (kb + c * GB);
      if (n >= GB) issue_n(c, GB); else issue_n(c, n);   // full chunks compile without per-block checks
    };

    int row = 0;                   // warp-local index of the row being accumulated
    float4 acc[NCH][4];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[ch][p] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto finish_row = [&]() {      // lane q: out_q = slot0(q) + slot3(q+1) + slot2(q+2) + slot1(q+3); store; reset
      const int grow = grow0 + row;
      char* yrow = const_cast<char*>(ptr_mad(Yl, (uint32_t)grow, ldyb));
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float4 out = acc[ch][0];
        out = add4(out, shfl_idx4(acc[ch][3], src1));
        out = add4(out, shfl_idx4(acc[ch][2], src2));
        out = add4(out, shfl_idx4(acc[ch][1], src3));
        if (grow < n_brows) st_stream_f4(reinterpret_cast<float*>(yrow) + ch * 32, out);
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[ch][p] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      ++row;
    };
    // finish the current row, then any empty rows that follow it (they store zeros)
    auto finish_rows = [&]() {
      finish_row();
      while (row < kRowsPerWarp && ((emptymask >> row) & 1u)) finish_row();
    };

    if (emptymask & 1u) {                             // leading empty rows
      do finish_row(); while (row < kRowsPerWarp && ((emptymask >> row) & 1u));
    }
    auto compute_n = [&](int c, const int nn) {
      const int kc = kb + c * GB;
      const unsigned char* xs = my_slots + (c & 1) * kSlotBytes;
      const float* xp = reinterpret_cast<const float*>(xs) + lane_xoff;
      const float* wp = reinterpret_cast<const float*>(xs + GB * kRowBytes) + 4 * q;
      // bit b set <=> block kc + b is the last block of its row
      const unsigned rel = (unsigned)(my_re - kc - 1);
      const unsigned endmask = __reduce_or_sync(0xffffffffu, rel < (unsigned)nn ? 1u << rel : 0u);
#pragma unroll
      for (int b = 0; b < GB; ++b) {             // fully unrolled: shared-memory offsets become immediates
        if (b < nn) {
          const float4 w = *reinterpret_cast<const float4*>(wp + b * 16);  // B[(q+s)%4][q], s = 0..3
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) {
            float4 x = *reinterpret_cast<const float4*>(xp + b * C + ch * 32);
            if (ELU) x = elu4(x);
            acc[ch][0] = fma4(w.x, x, acc[ch][0]);
            acc[ch][1] = fma4(w.y, x, acc[ch][1]);
            acc[ch][2] = fma4(w.z, x, acc[ch][2]);
            acc[ch][3] = fma4(w.w, x, acc[ch][3]);
          }
          if (endmask & (1u << b)) finish_rows();
        }
      }
    };

    if (nchunks > 0) issue(0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) issue(c + 1);
      cp_async_commit();
      cp_async_wait<1>();          // chunk c (and everything older) has landed
      __syncwarp();
      const int n = ke - (kb + c * GB);
      if (n >= GB) compute_n(c, GB); else compute_n(c, n);
      __syncwarp();                // all lanes done with slot c & 1 before chunk c + 2 overwrites it
    }

    cp_async_wait<0>();            // this thread's share of the index prefetch has landed
    __syncthreads();               // ... and everyone else's; also: all warps are done with bp / bc of this tile
  }
}

template <int NCH, int GB>
static int launch_stream(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X,
                         int64_t ldx, float* Y, int64_t ldy, int64_t n_brows, bool elu, cudaStream_t st) {
  constexpr int kSlotBytes = GB * (128 * NCH * 4 + 64);
  constexpr size_t smem = (size_t)kWarps * kSlots * kSlotBytes + 2 * kBcCap * 4 + 3 * (kTileRows + 4) * 4;
  auto kern = elu ? bsr4_spmm_stream_kernel<NCH, GB, true> : bsr4_spmm_stream_kernel<NCH, GB, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return SN_ERR_UNSUPPORTED;
  const int64_t n_tiles = ceil_div(n_brows, kTileRows);
  const int64_t grid = n_tiles < (int64_t)sms * per_sm ? n_tiles : (int64_t)sms * per_sm;
  kern<<<(unsigned)grid, kThreads, smem, st>>>(browptr, bcolind, bval, X, ldx, Y, ldy, (int)n_brows, (int)n_tiles);
  return launch_status();
}

// Returns SN_ERR_UNSUPPORTED when the streaming kernel does not apply (caller falls back to direct gather).
int launch_bsr4_stream(const int32_t* browptr, const int32_t* bcolind, const float* bval, const float* X, int64_t ldx,
                       float* Y, int64_t ldy, int64_t n_brows, int64_t C, bool elu, cudaStream_t st) {
  if (n_brows >= 0x7fffffffLL - kTileRows || ldx >= (1LL << 30) || ldy >= (1LL << 30)) return SN_ERR_UNSUPPORTED;
  switch (C) {
    case 128: return launch_stream<1, 6>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    case 256: return launch_stream<2, 4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    case 512: return launch_stream<4, 3>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    default: return SN_ERR_UNSUPPORTED;
  }
}

}  // namespace sn
