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
  unsigned char* my_slots = slots + (size_t)warp * kSlots * kSlotBytes;
  const float* Xlane = X + lane * 4;                      // this lane's 16-byte unit of every gathered row
  const int lane_xoff = q * C4 + 4 * tl;                  // float offset of this lane's float4 inside a row
  const int src1 = (lane + 8) & 31, src2 = (lane + 16) & 31, src3 = (lane + 24) & 31;

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

    const int r0 = tile * kTileRows;
    const int k0 = bp[0];
    const int rw0 = warp * kRowsPerWarp;                       // this warp's rows inside the tile
    const int kb = bp[rw0], ke = bp[rw0 + kRowsPerWarp];       // ... and its contiguous run of blocks
    const int nchunks = (ke - kb + GB - 1) / GB;

    const int kbase = kb - k0;
    auto issue = [&](int c) {  // gather chunk c (blocks [kb + c*GB, ...)) into slot c & 1
      const int kc = kb + c * GB;
      const int n = min(GB, ke - kc);
      unsigned char* xs = my_slots + (c & 1) * kSlotBytes;
#pragma unroll
      for (int b = 0; b < GB; ++b) {
        if (b < n) {
          const int kk = kbase + c * GB + b;
          const int j = kk < kBcCap ? bc[kk] : __ldg(bcolind + kc + b);
          const float* src = Xlane + (int64_t)j * ldx;
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) cp_async16(xs + b * kRowBytes + ch * 512 + lane * 16, src + ch * 128);
        }
      }
      if (lane < n * 4) cp_async16(xs + GB * kRowBytes + lane * 16, bval + (int64_t)kc * 16 + lane * 4);
    };

    int row = rw0;                 // current row (tile-local) and the end of its run of blocks
    int row_end = bp[row + 1];
    float4 acc[NCH][4];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[ch][p] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto finish_row = [&]() {      // lane q: out_q = slot0(q) + slot3(q+1) + slot2(q+2) + slot1(q+3); store; reset
      const int grow = r0 + row;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float4 out = acc[ch][0];
        out = add4(out, shfl_idx4(acc[ch][3], src1));
        out = add4(out, shfl_idx4(acc[ch][2], src2));
        out = add4(out, shfl_idx4(acc[ch][1], src3));
        if (grow < n_brows) st_stream_f4(Y + (int64_t)grow * ldy + lane_xoff + ch * 32, out);
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[ch][p] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };

    if (nchunks > 0) issue(0);
    cp_async_commit();
    for (int c = 0; c < nchunks; ++c) {
      if (c + 1 < nchunks) issue(c + 1);
      cp_async_commit();
      cp_async_wait<1>();          // chunk c (and everything older) has landed
      __syncwarp();
      const int kc = kb + c * GB;
      const int n = min(GB, ke - kc);
      const unsigned char* xs = my_slots + (c & 1) * kSlotBytes;
      const float* xp = reinterpret_cast<const float*>(xs) + lane_xoff;
      const float* wp = reinterpret_cast<const float*>(xs + GB * kRowBytes) + 4 * q;
      int b = 0;
      while (b < n) {
        const int stop = min(n, row_end - kc);   // blocks of the current row that live in this chunk
#pragma unroll 3
        for (; b < stop; ++b) {
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
        }
        if (b < n) {                             // the row's run ended inside the chunk (or the row is empty)
          finish_row();
          ++row;
          row_end = bp[row + 1];
        }
      }
      __syncwarp();                // all lanes done with slot c & 1 before chunk c + 2 overwrites it
    }
    // remaining rows of this warp (the last non-empty one and any trailing empty rows)
    for (; row < rw0 + kRowsPerWarp; ++row) finish_row();

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
  if (n_brows >= 0x7fffffffLL - kTileRows) return SN_ERR_UNSUPPORTED;
  switch (C) {
    case 128: return launch_stream<1, 6>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    case 256: return launch_stream<2, 4>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    case 512: return launch_stream<4, 3>(browptr, bcolind, bval, X, ldx, Y, ldy, n_brows, elu, st);
    default: return SN_ERR_UNSUPPORTED;
  }
}

}  // namespace sn
