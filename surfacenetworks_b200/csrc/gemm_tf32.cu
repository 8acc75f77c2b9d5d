// gemm_tf32.cu -- the dense half of a ResNet stage on the 5th-generation tensor cores.
//
//   C[M x N] = A[M x K] * B[N x K]^T + bias[N] (+ rscale[N] .* R[M x N])        fp32 in, fp32 out
//
// Replaces nn.Linear inside GraphConv1x1 (reference src/utils/utils_pt.py:89,99) with the training-mode
// BatchNorm of :84,98 folded into B / bias by the caller (W' = W diag(gamma*rstd), b' = b + W (beta - gamma*mu*rstd)),
// and serves the backward product dZ = dY W_s + p .* Z + q through the same epilogue.  M is B*V or B*F
// (1e5..1e6 rows), N and K are feature widths (128 / 256).
//
// Precision: the reference Linear is true fp32.  tcgen05 `kind::tf32` keeps 10 mantissa bits, so every
// operand tile is split in shared memory into hi = tf32(x) and lo = x - hi and three MMAs are issued per
// k-step (hi*hi + hi*lo + lo*hi; the dropped lo*lo term is ~2^-22 relative): fp32-grade results at 3 MMAs.
//
// Structure (one CTA per SM, persistent over 128-row tiles; 16 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the A tile [128 x 32] and B tile [N x 32]
//               (128-byte swizzle) into a STAGES-deep ring, completion on `full` mbarriers
//   warps 4-7   split of the A tile: hi written back in place, lo into the twin tile, fence.proxy.async, `ready`
//               mbarrier (B is split once per launch by a tiny pre-kernel into the caller's workspace)
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N, K=8) x 12 per
//               k-block; tcgen05.commit frees the smem stage (`empty`) and publishes the accumulator (`tmem_full`)
//   warp 2      TMEM allocation (2 accumulator buffers of N columns)
//   warps 8-15  epilogue: tcgen05.ld 32 lanes x 32 columns, per-warp transpose buffer, bias / (scaled) residual with
//               the residual loads issued one chunk ahead, coalesced streaming stores; overlaps the next tile's
//               MMAs through the second TMEM buffer
// Bound: the fused stage is HBM-bound once on tensor cores (A read once, C written once; B stays in L2).
//
// Two generations live here.  `gemm_tf32_ss_kernel` (round 1, structure above) feeds BOTH operands from shared memory and was
// bound by shared-memory bandwidth: per 32-wide k-block the A tile crossed shared memory five times (TMA write, split
// read, hi + lo write-back, 12 operand reads by the MMAs).  `gemm_tf32_ts_kernel` (round 2, the default) takes A out of
// shared memory for the tensor core: the split warps read the TMA-landed tile ONCE, and write hi / lo straight into
// TENSOR MEMORY (tcgen05.st); the MMAs run in TS mode (A operand from TMEM, B from shared memory).  The A ring (16 KB
// stages, freed as soon as the split warps hold the tile in registers) and the B ring (pre-split weights streamed from
// L2) are decoupled, so TMA runs 6 A stages + 4 TMEM stages ahead of the tensor core instead of 3.  SN_GEMM_LEGACY_SS
// selects the old kernel (A/B runs, tools/gemm_bench.py).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace sn {
namespace gemm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;            // 32 fp32 = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;              // tf32: 32 bytes per MMA k-step
constexpr int kThreads = 512;            // 4 control/split warp slots x 2 + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct Params {
  const float* bias;     // [N] or null
  const float* R;        // [M x N] or null: out += rscale .* R
  const float* rscale;   // [N] or null (= 1)
  const float* gbias;    // [ceil(M / rows_per_group) x N] or null: per-row-group bias (AvgResNet2's per-mesh term)
  int rows_per_group;
  float* C;
  int64_t ldr, ldc;
  int M, N, K;
  int stages;
  int split;             // 1: 3xTF32 (hi/lo split), 0: single-pass TF32
  int l2_prefetch;       // 1: prefetch the next tile's residual rows into L2 while this tile is processed
  int elu_left;          // 1: multiply output columns [0, N/2) by elu'(R) (backward of the activated left half of Z)
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_ss_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_blo, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle (TMA and UMMA agree on address bits [7,10))
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N;
  const uint32_t a_bytes = kBlockM * kBlockK * 4;           // 16 KB
  const uint32_t b_bytes = (uint32_t)N * kBlockK * 4;       // 16 / 32 KB
  const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);     // A, A_lo, B, B_lo
  __shared__ uint64_t full_bar[kMaxStages], ready_bar[kMaxStages], empty_bar[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float epi_buf[kEpiWarps * 32 * 20];   // per-epilogue-warp transpose buffers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_kb = p.K / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(ready_bar + s, 4);      // one arrival per split warp
      mbar_init(empty_bar + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full + b, 1);
      mbar_init(tmem_empty + b, kEpiWarps);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {                       // TMEM: 2 accumulator buffers of N fp32 columns (power of two >= 32)
    const uint32_t cols = 2u * (uint32_t)N;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1u);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          unsigned char* sb = sa + 2 * a_bytes;
          mbar_arrive_expect_tx(full_bar + stage, a_bytes + (p.split ? 2 : 1) * b_bytes);
          tma_load_2d(sa, &map_a, kb * kBlockK, tile * kBlockM, full_bar + stage);
          tma_load_2d(sb, &map_b, kb * kBlockK, 0, full_bar + stage);
          if (p.split) tma_load_2d(sb + b_bytes, &map_blo, kb * kBlockK, 0, full_bar + stage);   // pre-split weights
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(tmem_empty + buf, ((it >> 1) & 1) ^ 1u);       // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * N);
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(ready_bar + stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sal = sa + a_bytes, sb = sa + 2 * a_bytes, sbl = sb + b_bytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint32_t off = k * kUmmaK * 4;               // 32 bytes per k-step inside the swizzle row
            const uint64_t da = umma_desc_sw128(sa + off), db = umma_desc_sw128(sb + off);
            umma_tf32(tmem_d, da, db, idesc, (kb | k) != 0);
            if (p.split) {
              umma_tf32(tmem_d, da, umma_desc_sw128(sbl + off), idesc, true);
              umma_tf32(tmem_d, umma_desc_sw128(sal + off), db, idesc, true);
            }
          }
          umma_commit(empty_bar + stage);                      // smem stage reusable once these MMAs retire
          if (kb == n_kb - 1) umma_commit(tmem_full + buf);    // accumulator complete
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ hi / lo split
    const int t = threadIdx.x - 128;                           // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar + stage, phase);
        if (p.split) {
          // A tile: 1024 float4, 8 per thread, all loads first (ILP); B arrives pre-split (hi | lo) by TMA
          unsigned char* sa = smem + (size_t)stage * stage_bytes + (size_t)t * 16;
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<float4*>(sa + i * 2048);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 hi;   // round-to-nearest tf32: |lo| <= 2^-12 |x|, and hi is exact for the tensor core
            hi.x = to_tf32(x[i].x); hi.y = to_tf32(x[i].y); hi.z = to_tf32(x[i].z); hi.w = to_tf32(x[i].w);
            *reinterpret_cast<float4*>(sa + i * 2048) = hi;
            *reinterpret_cast<float4*>(sa + a_bytes + i * 2048) =
                make_float4(x[i].x - hi.x, x[i].y - hi.y, x[i].z - hi.z, x[i].w - hi.w);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_bar + stage);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue
    // 8 epilogue warps: warp e reads TMEM lanes [32 (e % 4), +32) and owns the column half e / 4.
    // Each thread receives 32 consecutive columns of ONE row from TMEM; a per-warp 32 x 36-float transpose buffer
    // turns that into row-contiguous float4 accesses (8 lanes cover one 128-byte line of C / R): conflict-free
    // STS.128 / LDS.128, fully coalesced global traffic.  The residual operand R does not depend on the
    // accumulator, so its loads are issued one chunk ahead (and before the wait on the MMA): their latency is
    // hidden behind the tensor-core work instead of serialising the epilogue.
    const int e = warp - 8;
    const int quarter = e & 3, half = e >> 2;
    float* tbuf = epi_buf + e * (32 * 20);
    const int tr = lane >> 2, tc = (lane & 3) * 4;            // 4 lanes cover one row's 16 columns (64 bytes)
    const int ncol = N / 2;                                    // columns owned by this warp
    const int cbase = half * ncol;
    auto load_r = [&](float4 (&rr)[4], int row0, int c0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row0 + i * 8 + tr;
        rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.M) rr[i] = __ldcs(reinterpret_cast<const float4*>(p.R + (int64_t)row * p.ldr + c0 + tc));
      }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int row0 = tile * kBlockM + quarter * 32;
      float4 rr[4], rn[4], rn2[4];
      if (p.R) {                                               // prefetch before waiting for the accumulator
        load_r(rr, row0, cbase);
        load_r(rn, row0, cbase + 16);
        if (p.l2_prefetch) {                                   // ... and the next tile's residual rows into L2
          const int nrow0 = (tile + (int)gridDim.x) * kBlockM + quarter * 32;
          const int lines = ncol / 32;                         // 128-byte lines per row in this warp's column half
          for (int i = lane; i < 32 * lines; i += 32) {
            const int row = nrow0 + i / lines;
            if (row < p.M) prefetch_l2(p.R + (int64_t)row * p.ldr + cbase + (i % lines) * 32);
          }
        }
      }
      mbar_wait(tmem_full + buf, (it >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = cbase; c0 < cbase + ncol; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N + c0), v);
        if (p.R && c0 + 32 < cbase + ncol) load_r(rn2, row0, c0 + 32);   // residual two chunks ahead
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(tbuf + lane * 20 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        __syncwarp();
        float4 bia = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.bias) bia = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + tc));
        if (p.rscale) rs = __ldg(reinterpret_cast<const float4*>(p.rscale + c0 + tc));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + tr;
          const int row = row0 + r;
          if (row < p.M) {
            float4 o = add4(*reinterpret_cast<const float4*>(tbuf + r * 20 + tc), bia);
            if (p.gbias)
              o = add4(o, __ldg(reinterpret_cast<const float4*>(p.gbias + (int64_t)(row / p.rows_per_group) * N + c0 + tc)));
            if (p.R) {
              o.x = fmaf(rs.x, rr[i].x, o.x); o.y = fmaf(rs.y, rr[i].y, o.y);
              o.z = fmaf(rs.z, rr[i].z, o.z); o.w = fmaf(rs.w, rr[i].w, o.w);
              if (p.elu_left && c0 < (N >> 1)) {   // chunk-uniform: 16-column chunks never straddle N/2
                o.x *= rr[i].x > 0.f ? 1.f : rr[i].x + 1.f; o.y *= rr[i].y > 0.f ? 1.f : rr[i].y + 1.f;
                o.z *= rr[i].z > 0.f ? 1.f : rr[i].z + 1.f; o.w *= rr[i].w > 0.f ? 1.f : rr[i].w + 1.f;
              }
            }
            st_stream_f4(p.C + (int64_t)row * p.ldc + c0 + tc, o);
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) { rr[i] = rn[i]; rn[i] = rn2[i]; }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + buf);
    }
  }

  // ------------------------------------------------------------------ teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t cols = 2u * (uint32_t)N;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols));
  }
}

// ====================================================================================================================
// Round-2 kernel: A operand in TENSOR MEMORY (TS mode)
// ====================================================================================================================
constexpr int kATmemStages = 4;        // TMEM A stages: 64 columns each (32 hi | 32 lo) = one 32-wide k-block
constexpr int kATmemCol0 = 256;        // accumulators live in columns [0, 256), A stages in [256, 512)
constexpr int kMaxAStages = 8;         // shared-memory A ring (16 KB stages)
constexpr int kMaxBStages = 4;         // shared-memory B ring (hi | lo, nmma x 128 B each)

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// One elected lane of a converged warp.  Code under `if (elect_one())` is compiled as single-thread, warp-uniform code:
// UTCHMMA / UTMALDG / UTCBAR are emitted bare.  Under `if (lane == 0)` ptxas wraps every one of them in an
// ELECT ... BRA.U.ANY loop (it must assume divergence), which made the MMA-issuing thread the bottleneck of the first
// version of this kernel (ncu: ~200 dependent uniform-datapath instructions per k-block, 1750 clk against 768 clk of
// tensor-pipe time -- profiles/r2_gemm_notes.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// TS-mode MMA with the shared-memory descriptor given as (lo, hi) words: hi is a constant of the layout, lo advances by
// (byte offset >> 4), so the issuing thread spends one integer add per MMA on descriptors.
__device__ __forceinline__ void umma_tf32_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 d;\nsetp.ne.b32 p, %5, 0;\nmov.b64 d, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}

struct ParamsTS {
  const float* bias;     // [N] or null
  const float* R;        // [M x N] or null: out += rscale .* R
  const float* rscale;   // [N] or null (= 1)
  const float* gbias;    // per-row-group bias or null
  int rows_per_group;
  float* C;
  int64_t ldr, ldc;
  int M, N, K;
  int nmma;              // accumulator width = columns per MMA: N for N <= 128, else 128
  int n_halves;          // ceil(N / nmma) column passes per row tile ("jobs")
  int n_groups, hpg;     // work items per row tile and column passes per item (1, n_halves unless N is wide)
  int a_resident;        // n_halves > 1 and K / 32 <= kATmemStages: the tile's whole A extent stays in TMEM for both passes
  int a_stages, b_stages;
  int split;             // 1: 3xTF32, 0: single pass
  int l2_prefetch;
  int elu_left;
  // ACT epilogue (template parameter of the kernel): C2 = elu(result) with leading dimension ldc2 -- the activated copy the
  // NEXT stage gathers from / takes as the left half of its concat buffer (C may then be null: the raw value has no other
  // consumer) -- and, when stat_partial is given, the per-CTA column sums / sums of squares of C2 [gridDim.x][2][N] for that
  // stage's BatchNorm (finished by colstats_final_kernel in a fixed order: bit-reproducible).
  float* C2;
  int64_t ldc2;
  float* stat_partial;
  // "+ elu-backward" epilogue of the AvgResNet2 stage backward (non-ACT kernel only):
  //   C = ((acc + bias + row_scale[row] * gbias[row / rows_per_group] + rscale .* R) .* elu'(R)) + R2
  // row_scale [M] (the mask weights), elu_all: the derivative applies to every column, R2 (optional second residual, e.g.
  // the block-residual gradient) arrives through map_c2 into its own slots.
  const float* row_scale;
  int elu_all, has_r2;
  int r_slots;           // residual boxes in flight per epilogue warp (1 or 2)
  int epi_slots;         // 4 KB staging boxes per epilogue warp: output (+ residual slots / activated copy)
  int debug;             // timing experiments only (SN_GEMM_DEBUG_*, tools/gemm_stage_bench.py): results are garbage
};

// Roles (16 warps): 0 = A producer (TMA), 3 = B producer (TMA), 1 = MMA issuer, 2 = TMEM allocation, 4-7 = A split into
// TMEM (warp w owns TMEM lanes [32 (w - 4), +32) = tile rows), 8-15 = epilogue.
// Barriers: a_full / a_free  (TMA -> split warps -> TMA)           shared-memory A ring
//           a_ready / a_tfree (split warps -> MMA -> split warps)  TMEM A stages (tcgen05.st ... tcgen05.commit)
//           b_full / b_free  (TMA -> MMA -> TMA)                   shared-memory B ring
//           tmem_full / tmem_empty (MMA -> epilogue -> MMA)        two accumulators
template <bool ACT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_c,
                    const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_c2,
                    const ParamsTS p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nmma = p.nmma;
  constexpr uint32_t a_bytes = kBlockM * kBlockK * 4;       // 16 KB
  const uint32_t b_bytes = (uint32_t)nmma * kBlockK * 4;    // 8 / 16 KB per half (hi or lo)
  unsigned char* a_ring = smem;
  unsigned char* b_ring = smem + (size_t)p.a_stages * a_bytes;
  // epilogue staging, per epilogue warp: [output box 4 KB][p.epi_slots - 1 more 4 KB boxes: residual slots / activated copy]
  unsigned char* epi_ring = b_ring + (size_t)p.b_stages * 2 * b_bytes;
  __shared__ uint64_t a_full[kMaxAStages], a_free[kMaxAStages], a_ready[kATmemStages], a_tfree[kATmemStages];
  __shared__ uint64_t b_full[kMaxBStages], b_free[kMaxBStages], tmem_full[2], tmem_empty[2];
  __shared__ uint64_t r_full[kEpiWarps * 2];                  // [epilogue warp][residual slot]
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_bias[256], s_rscale[256];   // epilogue vectors: LDS instead of an L1-missing __ldg per chunk
  __shared__ __align__(16) float s_gb[kEpiWarps][64];          // per-warp copy of the job's group-bias row (when one mesh covers the slab)
  __shared__ __align__(16) float s_stat[ACT ? kEpiWarps : 1][2][128];   // ACT: per-epilogue-warp column sums / sums of squares

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items.  Standard case (N <= 256): one item per 128-row tile, p.n_halves column passes each.  Wide-N case (the
  // dense_correspondence correlation FA . FB^T, N ~ 7000): the column passes of a row tile are split into p.n_groups
  // groups so that every SM gets items; an item = (row tile, group) runs p.hpg passes (the last group what is left).
  const int n_groups = p.n_groups;
  const int n_tiles = ((p.M + kBlockM - 1) / kBlockM) * n_groups;
  const int n_kb = (p.K + kBlockK - 1) / kBlockK;              // a K tail is zero-filled by TMA
  auto item_row = [&](int item) { return (item / n_groups) * kBlockM; };
  auto item_half0 = [&](int item) { return (item % n_groups) * p.hpg; };
  auto item_nh = [&](int item) {
    const int left = p.n_halves - (item % n_groups) * p.hpg;
    return left < p.hpg ? left : p.hpg;
  };
  if (threadIdx.x < p.N && threadIdx.x < 256) {
    s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
    s_rscale[threadIdx.x] = p.rscale ? p.rscale[threadIdx.x] : 1.f;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(a_full + s, 1);
      mbar_init(a_free + s, 4);          // one arrival per split warp
    }
    for (int s = 0; s < kATmemStages; ++s) {
      mbar_init(a_ready + s, 4);
      mbar_init(a_tfree + s, 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_free + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full + b, 1);
      mbar_init(tmem_empty + b, kEpiWarps);
    }
    for (int w = 0; w < kEpiWarps; ++w) {
      mbar_init(r_full + 2 * w, 1);
      mbar_init(r_full + 2 * w + 1, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {                       // all 512 TMEM columns: 2 accumulators + 4 A stages (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n_a_pass = p.a_resident ? 1 : item_nh(tile);   // times the A tile is streamed per item
        for (int pass = 0; pass < n_a_pass; ++pass)
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(a_free + stage, phase ^ 1u);
            mbar_arrive_expect_tx(a_full + stage, a_bytes);
            tma_load_2d(a_ring + (size_t)stage * a_bytes, &map_a, kb * kBlockK, item_row(tile), a_full + stage);
            if (++stage == p.a_stages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ B producer (pre-split weights, L2-resident)
    if (elect_one() && !(p.debug & 16)) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int half = 0, nh = item_nh(tile), h0 = item_half0(tile); half < nh; ++half)
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(b_free + stage, phase ^ 1u);
            unsigned char* sb = b_ring + (size_t)stage * 2 * b_bytes;
            mbar_arrive_expect_tx(b_full + stage, (p.split ? 2u : 1u) * b_bytes);
            tma_load_2d(sb, &map_b, kb * kBlockK, (h0 + half) * nmma, b_full + stage);
            if (p.split) tma_load_2d(sb + b_bytes, &map_blo, kb * kBlockK, (h0 + half) * nmma, b_full + stage);
            if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread runs the whole role)
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nmma >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
      const uint64_t d0 = umma_desc_sw128(smem_u32(b_ring));
      const uint32_t d_lo0 = (uint32_t)d0, d_hi = (uint32_t)(d0 >> 32);
      const uint32_t stage_units = (2u * b_bytes) >> 4, lo_units = b_bytes >> 4;   // descriptor address units of 16 bytes
      const uint32_t split = (uint32_t)p.split;
      int bs = 0;
      uint32_t bph = 0;
      uint32_t a_cnt = 0;                // TMEM A stages consumed so far (stage = a_cnt % 4, parity = (a_cnt / 4) & 1)
      uint32_t job = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n_halves = item_nh(tile);
        for (int half = 0; half < n_halves; ++half, ++job) {
          const uint32_t buf = job & 1u;
          mbar_wait(tmem_empty + buf, ((job >> 1) & 1u) ^ 1u);    // epilogue drained this accumulator
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_d = tmem_base + buf * (uint32_t)nmma;
          for (int kb = 0; kb < n_kb; ++kb) {
            const uint32_t cnt = p.a_resident ? a_cnt + (uint32_t)kb : a_cnt;
            const uint32_t at = cnt % kATmemStages;
            if (half == 0 || !p.a_resident) mbar_wait(a_ready + at, (cnt / kATmemStages) & 1u);
            if (!(p.debug & 16)) mbar_wait(b_full + bs, bph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ta = tmem_base + (uint32_t)(kATmemCol0 + at * 64);
            const uint32_t dlo = d_lo0 + (uint32_t)bs * stage_units;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK && !(p.debug & 4); ++k) {
              // 32 bytes per k-step inside the swizzle row = 2 descriptor units; 8 TMEM columns per k-step of A
              umma_tf32_ts_lohi(tmem_d, ta + k * kUmmaK, dlo + 2 * k, d_hi, idesc, (uint32_t)((kb | k) != 0));
              if (split) {
                umma_tf32_ts_lohi(tmem_d, ta + k * kUmmaK, dlo + lo_units + 2 * k, d_hi, idesc, 1u);      // hi * lo
                umma_tf32_ts_lohi(tmem_d, ta + 32 + k * kUmmaK, dlo + 2 * k, d_hi, idesc, 1u);            // lo * hi
              }
            }
            if (!(p.debug & 16)) umma_commit(b_free + bs);
            if (!p.a_resident || half == n_halves - 1) umma_commit(a_tfree + at);
            if (kb == n_kb - 1) umma_commit(tmem_full + buf);
            if (++bs == p.b_stages) { bs = 0; bph ^= 1u; }
            if (!p.a_resident) ++a_cnt;
          }
        }
        if (p.a_resident) a_cnt += (uint32_t)n_kb;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ A: shared memory -> hi / lo -> tensor memory
    // Thread r (0..127) owns tile row r = TMEM lane r.  The TMA box is 128-byte swizzled: logical 16-byte chunk i of
    // row r sits at chunk i ^ (r & 7), so the eight LDS.128 of a quarter-warp hit eight distinct bank groups.
    const int r = threadIdx.x - 128;
    const uint32_t lane_base = (uint32_t)((warp - 4) * 32) << 16;
    const uint32_t a_ring_s = smem_u32(a_ring);
    int as = 0;
    uint32_t aph = 0, t_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int pass = 0, n_a_pass = p.a_resident ? 1 : item_nh(tile); pass < n_a_pass; ++pass)
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(a_full + as, aph);
          const uint32_t row = a_ring_s + (uint32_t)as * a_bytes + (uint32_t)r * 128u;
          uint32_t hi[32], lo[32];
          if (p.debug & 8) {
#pragma unroll
            for (int i = 0; i < 32; ++i) hi[i] = lo[i] = (uint32_t)i;
          } else
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = lds_f4(row + (uint32_t)((i ^ (r & 7)) << 4));
            const float h0 = to_tf32(x.x), h1 = to_tf32(x.y), h2 = to_tf32(x.z), h3 = to_tf32(x.w);
            hi[4 * i] = __float_as_uint(h0); hi[4 * i + 1] = __float_as_uint(h1);
            hi[4 * i + 2] = __float_as_uint(h2); hi[4 * i + 3] = __float_as_uint(h3);
            // lo is rounded to tf32 too: the tensor core TRUNCATES the low 13 mantissa bits of its operands (measured,
            // tests/test_gpu_gemm.py::test_tf32_operand_truncation_probe), rounding here halves that error
            lo[4 * i] = __float_as_uint(to_tf32(x.x - h0)); lo[4 * i + 1] = __float_as_uint(to_tf32(x.y - h1));
            lo[4 * i + 2] = __float_as_uint(to_tf32(x.z - h2)); lo[4 * i + 3] = __float_as_uint(to_tf32(x.w - h3));
          }
          __syncwarp();                                        // every lane has consumed its shared-memory row
          if (lane == 0) mbar_arrive(a_free + as);             // the TMA producer may refill this slot
          const uint32_t at = t_cnt % kATmemStages;
          mbar_wait(a_tfree + at, ((t_cnt / kATmemStages) & 1u) ^ 1u);   // the MMAs that read this TMEM stage retired
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ta = tmem_base + lane_base + (uint32_t)(kATmemCol0 + at * 64);
          tmem_st32(ta, hi);
          if (p.split) tmem_st32(ta + 32, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready + at);
          ++t_cnt;
          if (++as == p.a_stages) { as = 0; aph ^= 1u; }
        }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ epilogue (one job = one row tile x one column pass)
    // Round-2 rewrite.  The first TS epilogue moved the tile through per-thread global accesses (register prefetch of the
    // residual, transpose buffer, 64-byte store segments): ~375 instructions per 16-column chunk and warp, and it -- not
    // HBM, not the tensor pipe -- set the kernel's period (tools/gemm_stage_bench.py: [128000 x 128] . [128 x 128] + R
    // 54 us, 25 us with the epilogue switched off).  Now every global access of the epilogue is a TMA box:
    //   * warp e owns tile rows [32 (e % 4), +32) (its TMEM lanes) and the column half e / 4 of the job; it works in
    //     32-column chunks: tcgen05.ld gives lane r the 32 columns of row r = one 128-byte row of a [32 x 32] box;
    //   * the residual box of that chunk was TMA-loaded (128-byte swizzle) into the warp's slot one or two chunks ahead
    //     (flat chunk sequence across jobs and tiles: loads for the next job fly while the tensor core works on it);
    //   * results go to the warp's output box with conflict-free swizzled STS.128 and leave with ONE
    //     cp.async.bulk.tensor store (rows >= M are clipped by the tensor map): no address arithmetic, no predicates.
    const int e = warp - 8;
    const int quarter = e & 3, chalf = e >> 2;
    const int ncol = nmma / 2;                                 // columns owned by this warp inside the job
    const int nchunks = ncol / 32;                             // 2 (nmma = 128) or 1 (nmma = 64)
    const int N = p.N;
    unsigned char* my_g = epi_ring + (size_t)e * p.epi_slots * 4096;
    const uint32_t my_s = smem_u32(my_g);
    const uint32_t out_s = my_s;                               // output box
    const int r_slots = p.r_slots;                             // residual slots (ACT: the last box stages the activated copy)
    const uint32_t r_s = my_s + 4096;
    const uint32_t r2_s = r_s + (uint32_t)r_slots * 4096;      // second residual: same slot count, behind the first
    const uint32_t act_s = my_s + (uint32_t)(p.epi_slots - 1) * 4096;
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    struct ChunkIt { int tile, half, c; };
    auto advance = [&](ChunkIt& it) {
      if (++it.c == nchunks) {
        it.c = 0;
        if (++it.half == item_nh(it.tile)) { it.half = 0; it.tile += (int)gridDim.x; }
      }
    };
    ChunkIt pf = {(int)blockIdx.x, 0, 0};
    uint32_t r_issued = 0, r_used = 0;
    auto issue_r = [&]() {                                     // TMA load of the residual box of chunk `pf` into the next slot
      if (pf.tile < n_tiles) {
        if (lane == 0) {
          const uint32_t slot = r_issued % (uint32_t)r_slots;
          mbar_arrive_expect_tx(r_full + 2 * e + slot, (!ACT && p.has_r2) ? 8192u : 4096u);
          tma_load_2d(my_g + 4096 + slot * 4096u, &map_r, (item_half0(pf.tile) + pf.half) * nmma + chalf * ncol + pf.c * 32,
                      item_row(pf.tile) + quarter * 32, r_full + 2 * e + slot);
          if (!ACT && p.has_r2)
            tma_load_2d(my_g + 4096 + ((uint32_t)r_slots + slot) * 4096u, &map_c2,
                        (item_half0(pf.tile) + pf.half) * nmma + chalf * ncol + pf.c * 32, item_row(pf.tile) + quarter * 32,
                        r_full + 2 * e + slot);
        }
        ++r_issued;
        advance(pf);
      }
    };
    if (p.R)
      for (int i = 0; i < r_slots; ++i) issue_r();
    if (ACT) {
      for (int i = lane; i < 2 * 128; i += 32) (&s_stat[e][0][0])[i] = 0.f;
      __syncwarp();
    }
    uint32_t job = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int n_halves = item_nh(tile), h0 = item_half0(tile);
      for (int half = 0; half < n_halves; ++half, ++job) {
        const uint32_t buf = job & 1u;
        const int row0 = item_row(tile) + quarter * 32;
        const int cbase = (h0 + half) * nmma + chalf * ncol;   // first global column of this warp's share
        const int row_c = row0 + lane < p.M ? row0 + lane : p.M - 1;
        const float* gb_row = p.gbias ? p.gbias + (int64_t)(row_c / p.rows_per_group) * N : nullptr;
        const float gb_w = (!ACT && p.row_scale) ? __ldg(p.row_scale + row_c) : 1.f;
        // group bias: the 32 rows of this slab almost always belong to one group (a mesh is thousands of rows): stage that
        // row's 64 columns once per job and read them as shared-memory broadcasts; slabs that straddle a boundary load per lane
        bool gb_uniform = false;
        if (p.gbias) {
          const int last = row0 + 31 < p.M ? row0 + 31 : p.M - 1;
          gb_uniform = row0 < p.M && row0 / p.rows_per_group == last / p.rows_per_group;
          if (gb_uniform) {
            __syncwarp();
            if (lane * 4 < ncol)
              *reinterpret_cast<float4*>(&s_gb[e][lane * 4]) =
                  __ldg(reinterpret_cast<const float4*>(p.gbias + (int64_t)(row0 / p.rows_per_group) * N + cbase + lane * 4));
            __syncwarp();
          }
        }
        mbar_wait(tmem_full + buf, (job >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int ci = 0; ci < nchunks && !(p.debug & 1); ++ci) {
          const int c0 = cbase + ci * 32;
          uint32_t rslot_s = 0, r2slot_s = 0;
          if (p.R) {
            const uint32_t slot = r_used % (uint32_t)r_slots;
            mbar_wait(r_full + 2 * e + slot, (r_used / (uint32_t)r_slots) & 1u);
            rslot_s = r_s + slot * 4096u;
            r2slot_s = r2_s + slot * 4096u;
          }
          // the previous bulk store(s) of this warp must have read the staging boxes before they are overwritten
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {                        // 16 accumulator columns per tcgen05.ld
            uint32_t v[16];
            if (p.debug & 2) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = (uint32_t)(c0 + j);
            } else {
              tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * nmma + (c0 - (h0 + half) * nmma) + h * 16), v);
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int cidx = h * 4 + j4;                     // 16-byte chunk of the 128-byte box row
              const int col = c0 + cidx * 4;
              const uint32_t swz = row_off + (((uint32_t)cidx ^ sw) << 4);
              float4 o = make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                     __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
              if (n_groups == 1) o = add4(o, *reinterpret_cast<const float4*>(s_bias + col));   // wide-N: no epilogue vectors
              if (p.gbias)
                o = fma4(gb_w, gb_uniform ? *reinterpret_cast<const float4*>(&s_gb[e][col - cbase])
                                          : __ldg(reinterpret_cast<const float4*>(gb_row + col)), o);
              if (p.R) {
                const float4 rr = lds_f4(rslot_s + swz);
                const float4 rs = *reinterpret_cast<const float4*>(s_rscale + col);
                o.x = fmaf(rs.x, rr.x, o.x); o.y = fmaf(rs.y, rr.y, o.y);
                o.z = fmaf(rs.z, rr.z, o.z); o.w = fmaf(rs.w, rr.w, o.w);
                if (!ACT && ((p.elu_left && c0 < (N >> 1)) || p.elu_all)) {   // chunk-uniform: 32-column chunks never straddle N/2
                  o.x *= rr.x > 0.f ? 1.f : rr.x + 1.f; o.y *= rr.y > 0.f ? 1.f : rr.y + 1.f;
                  o.z *= rr.z > 0.f ? 1.f : rr.z + 1.f; o.w *= rr.w > 0.f ? 1.f : rr.w + 1.f;
                }
                if (!ACT && p.has_r2) o = add4(o, lds_f4(r2slot_s + swz));
              }
              if (!ACT || p.C)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_s + swz), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
              if (ACT) {
                const float4 a = elu4(o);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(act_s + swz), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
              }
            }
          }
          if (ci == nchunks - 1) {                             // accumulator drained: hand it back before the stores
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty + buf);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem accesses -> visible to TMA
          __syncwarp();
          if (p.R) {                                           // every lane has read this slot: refill it
            ++r_used;
            issue_r();
          }
          if (lane == 0) {
            if (!ACT || p.C)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_c), "r"(c0),
                           "r"(row0), "r"(out_s) : "memory");
            if (ACT)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_c2), "r"(c0),
                           "r"(row0), "r"(act_s) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (ACT && p.stat_partial) {
            // column sums of the activated box: lane L adds column L over the valid rows, in row order
            const int nvalid = p.M - row0 < 32 ? (p.M - row0 > 0 ? p.M - row0 : 0) : 32;
            float cs = 0.f, cq = 0.f;
            const uint32_t cpos = (uint32_t)(lane >> 2), cin = (uint32_t)(lane & 3) * 4u;
            for (int r = 0; r < nvalid; ++r) {
              float a;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(act_s + (uint32_t)r * 128u + ((cpos ^ (uint32_t)(r & 7)) << 4) + cin));
              cs += a;
              cq = fmaf(a, a, cq);
            }
            const int lc = half * ncol + ci * 32 + lane;       // column inside this warp's share of the N columns
            s_stat[e][0][lc] += cs;
            s_stat[e][1][lc] += cq;
          }
        }
        if (p.debug & 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + buf);
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
    __syncwarp();
    if (ACT && p.stat_partial) {
      // the CTA's partial: the four row-quarter warps that own a column, added in a fixed order
      asm volatile("bar.sync 1, 256;" ::: "memory");           // the eight epilogue warps only
      for (int i = threadIdx.x - 256; i < 2 * N; i += 256) {
        const int stat = i / N, col = i - stat * N;
        const int hf = col / nmma, within = col - hf * nmma;
        const int ch = within / ncol, lc = hf * ncol + (within - ch * ncol);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) a += s_stat[ch * 4 + q][stat][lc];
        p.stat_partial[(size_t)blockIdx.x * 2 * N + i] = a;
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// B -> (tf32(B), B - tf32(B)) into the workspace, row-major with leading dimension K
__global__ void split_weights_kernel(const float* __restrict__ B, int64_t ldb, float* __restrict__ hi, float* __restrict__ lo,
                                     int N, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const float x = B[(int64_t)n * ldb + k];
  const float h = to_tf32(x);
  hi[i] = h;
  lo[i] = to_tf32(x - h);     // rounded: the tensor core truncates its operands
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// row-major fp32 matrix [rows x cols] with leading dimension ld -> tensor map with a [box_rows x 32] box, 128B swizzle
static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gemm
}  // namespace sn

SN_API size_t sn_gemm_tf32_ws_bytes(int64_t N, int64_t K) {
  return (N <= 0 || K <= 0) ? 0 : (size_t)(2 * N * K) * sizeof(float) + 256;
}

namespace sn {
namespace gemm {

// Shared launcher: B_hi / B_lo are the pre-split weights (B_lo null = single pass on B_hi as given).
static int launch_gemm(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb, const float* bias,
                       const float* R, int64_t ldr, const float* rscale, const float* group_bias, int64_t rows_per_group,
                       float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int flags, cudaStream_t stream,
                       float* C_act = nullptr, int64_t ldc_act = 0, float* stat_partial = nullptr, bool wide = false,
                       const float* row_scale = nullptr, const float* R2 = nullptr, int64_t ldr2 = 0, bool elu_all = false) {
  const bool split = B_lo != nullptr;
  int dev = 0, sms = 148, smem_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int64_t tiles = ceil_div(M, kBlockM);
  unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  const int l2_prefetch = (!(flags & SN_GEMM_NO_L2_PREFETCH) && R && N >= 256) ? 1 : 0;
  CUtensorMap map_a, map_b, map_blo;
  if (!make_map(&map_a, A, M, K, lda, kBlockM)) return SN_ERR_UNSUPPORTED;

  if (flags & SN_GEMM_LEGACY_SS) {       // round-1 kernel: both operands from shared memory
    if (C_act) return SN_ERR_UNSUPPORTED;
    if (!make_map(&map_b, B_hi, N, K, ldb, (int)N)) return SN_ERR_UNSUPPORTED;
    if (split) {
      if (!make_map(&map_blo, B_lo, N, K, ldb, (int)N)) return SN_ERR_UNSUPPORTED;
    } else {
      map_blo = map_b;
    }
    Params p;
    p.bias = bias; p.R = R; p.rscale = rscale; p.C = C; p.ldr = ldr; p.ldc = ldc;
    p.gbias = group_bias; p.rows_per_group = group_bias ? (int)rows_per_group : 1;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.split = split ? 1 : 0;
    p.l2_prefetch = l2_prefetch;
    p.elu_left = (flags & SN_GEMM_ELU_BWD_LEFT) ? 1 : 0;
    const size_t stage_bytes = 2 * ((size_t)kBlockM * kBlockK * 4 + (size_t)N * kBlockK * 4);
    int stages = (int)(((size_t)smem_optin - 2048 - 22 * 1024) / stage_bytes);   // static smem: barriers + transpose buffers
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return SN_ERR_UNSUPPORTED;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gemm_tf32_ss_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_b, map_blo, p);
    return launch_status();
  }

  ParamsTS p;
  p.bias = bias; p.R = R; p.rscale = rscale; p.C = C; p.ldr = ldr; p.ldc = ldc;
  p.gbias = group_bias; p.rows_per_group = group_bias ? (int)rows_per_group : 1;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.nmma = N <= 128 ? (int)N : 128;
  p.n_halves = (int)ceil_div(N, p.nmma);
  p.n_groups = 1;
  p.hpg = p.n_halves;
  if (wide) {
    // enough (row tile, column group) items for two rounds over the SMs, every group non-empty
    int64_t g = ceil_div(2 * (int64_t)sms, tiles);
    if (g > p.n_halves) g = p.n_halves;
    if (g < 1) g = 1;
    p.hpg = (int)ceil_div(p.n_halves, g);
    p.n_groups = (int)ceil_div(p.n_halves, p.hpg);
    const int64_t items = tiles * p.n_groups;
    grid = (unsigned)(items < sms ? items : sms);
  }
  p.a_resident = (p.n_halves > 1 && ceil_div(K, kBlockK) <= kATmemStages) ? 1 : 0;
  p.split = split ? 1 : 0;
  p.l2_prefetch = l2_prefetch;
  p.elu_left = (flags & SN_GEMM_ELU_BWD_LEFT) ? 1 : 0;
  p.C2 = C_act; p.ldc2 = ldc_act; p.stat_partial = stat_partial;
  p.debug = (flags >> 8) & 0xff;
  if (!make_map(&map_b, B_hi, N, K, ldb, p.nmma)) return SN_ERR_UNSUPPORTED;
  if (split) {
    if (!make_map(&map_blo, B_lo, N, K, ldb, p.nmma)) return SN_ERR_UNSUPPORTED;
  } else {
    map_blo = map_b;
  }
  // epilogue boxes: [32 rows x 32 columns], 128-byte swizzle, one tensor map per matrix the epilogue touches
  CUtensorMap map_c, map_r, map_c2;
  if (C && !make_map(&map_c, C, M, N, ldc, 32)) return SN_ERR_UNSUPPORTED;
  if (C_act && !make_map(&map_c2, C_act, M, N, ldc_act, 32)) return SN_ERR_UNSUPPORTED;
  if (!C) map_c = map_c2;
  if (!C_act) map_c2 = map_c;
  if (R2) {                              // non-ACT kernel: map_c2 carries the second residual
    if (C_act || !R || !make_map(&map_c2, R2, M, N, ldr2, 32)) return SN_ERR_UNSUPPORTED;
  }
  p.row_scale = row_scale;
  p.elu_all = elu_all ? 1 : 0;
  p.has_r2 = R2 ? 1 : 0;
  if (R) {
    if (!make_map(&map_r, R, M, N, ldr, 32)) return SN_ERR_UNSUPPORTED;
  } else {
    map_r = map_c;
  }
  // shared memory: per epilogue warp one output box + the residual slots (+ the activated copy); a B ring of pre-split
  // weights streamed from L2; everything else goes to the A ring (HBM latency)
  static const int env_rslots = [] { const char* v = getenv("SN_GEMM_RSLOTS"); return v ? atoi(v) : 0; }();
  static const int env_bstages = [] { const char* v = getenv("SN_GEMM_BSTAGES"); return v ? atoi(v) : 0; }();
  p.r_slots = R ? ((env_rslots == 1 || env_rslots == 2) ? env_rslots : (C_act ? 1 : 2)) : 0;
  if ((C_act || R2) && p.r_slots > 1) p.r_slots = 1;
  p.epi_slots = 1 + p.r_slots * (R2 ? 2 : 1) + (C_act ? 1 : 0);
  const size_t a_bytes = (size_t)kBlockM * kBlockK * 4, b_stage = 2 * (size_t)p.nmma * kBlockK * 4;
  const size_t epi_bytes = (size_t)kEpiWarps * p.epi_slots * 4096;
  // static: epilogue vectors, barriers (+ 8 KB of statistics accumulators in the ACT kernel); 1 KB alignment
  const size_t budget = (size_t)smem_optin - (C_act ? 14 : 6) * 1024 - 1024;
  p.b_stages = (env_bstages >= 2 && env_bstages <= kMaxBStages) ? env_bstages : (epi_bytes <= 64 * 1024 ? 3 : 2);
  int a_stages = (int)((budget - p.b_stages * b_stage - epi_bytes) / a_bytes);
  if (a_stages > kMaxAStages) a_stages = kMaxAStages;
  if (a_stages < 2) return SN_ERR_UNSUPPORTED;
  p.a_stages = a_stages;
  const size_t smem = a_stages * a_bytes + p.b_stages * b_stage + epi_bytes + 1024;
  auto kern = C_act ? gemm_tf32_ts_kernel<true> : gemm_tf32_ts_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kThreads, smem, stream>>>(map_a, map_b, map_blo, map_c, map_r, map_c2, p);
  return launch_status();
}

// CTAs launch_gemm uses for M rows (= rows of the ACT kernel's statistics partials)
static int64_t gemm_grid(int64_t M) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = ceil_div(M, kBlockM);
  return tiles < sms ? tiles : sms;
}

static int check_gemm_args(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, const float* R,
                           int64_t ldr, const float* rscale, const float* group_bias, int64_t rows_per_group, const float* C,
                           int64_t ldc, int64_t M, int64_t N, int64_t K, int flags) {
  if (M < 0 || N <= 0 || K <= 0) return SN_ERR_ARG;
  if (M == 0) return SN_OK;
  if (!A || !B || !C || lda < K || ldb < K || ldc < N || (R && ldr < N)) return SN_ERR_ARG;
  if ((N != 64 && N != 128 && N != 256) || K % kBlockK != 0 || M >= 0x7fffffffLL - kBlockM) return SN_ERR_UNSUPPORTED;
  if (group_bias && (rows_per_group <= 0 || rows_per_group > 0x7fffffffLL || !aligned16(group_bias))) return SN_ERR_ARG;
  if (lda % 4 || ldb % 4 || ldc % 4 || (R && ldr % 4) || !aligned16(A) || !aligned16(B) || !aligned16(C) ||
      (R && !aligned16(R)) || (bias && !aligned16(bias)) || (rscale && !aligned16(rscale)))
    return SN_ERR_UNSUPPORTED;
  if ((flags & SN_GEMM_ELU_BWD_LEFT) && !R) return SN_ERR_ARG;
  return SN_OK;
}

}  // namespace gemm
}  // namespace sn

SN_API int sn_gemm_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, const float* R,
                            int64_t ldr, const float* rscale, const float* group_bias, int64_t rows_per_group, float* C,
                            int64_t ldc, int64_t M, int64_t N, int64_t K, int flags, void* ws, size_t ws_bytes,
                            sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  const int rc = check_gemm_args(A, lda, B, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags);
  if (rc != SN_OK || M == 0) return rc;
  if (flags & SN_GEMM_SINGLE_PASS)
    return launch_gemm(A, lda, B, nullptr, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags,
                       (cudaStream_t)stream);
  if (!ws || ws_bytes < sn_gemm_tf32_ws_bytes(N, K)) return SN_ERR_WORKSPACE;
  float* hi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  float* lo = hi + N * K;
  split_weights_kernel<<<(unsigned)ceil_div(N * K, 256), 256, 0, (cudaStream_t)stream>>>(B, ldb, hi, lo, (int)N, (int)K);
  return launch_gemm(A, lda, hi, lo, K, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags,
                     (cudaStream_t)stream);
}

SN_API int sn_gemm_tf32_presplit_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                                     const float* bias, const float* R, int64_t ldr, const float* rscale,
                                     const float* group_bias, int64_t rows_per_group, float* C, int64_t ldc, int64_t M,
                                     int64_t N, int64_t K, int flags, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  const int rc = check_gemm_args(A, lda, B_hi, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags);
  if (rc != SN_OK || M == 0) return rc;
  if (!B_lo || !aligned16(B_lo) || (flags & SN_GEMM_SINGLE_PASS)) return SN_ERR_ARG;
  return launch_gemm(A, lda, B_hi, B_lo, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags,
                     (cudaStream_t)stream);
}

SN_API int sn_gemm_tf32_presplit_elubwd_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                                            const float* bias, const float* R, int64_t ldr, const float* rscale,
                                            const float* group_bias, int64_t rows_per_group, const float* row_scale,
                                            const float* R2, int64_t ldr2, float* C, int64_t ldc, int64_t M, int64_t N,
                                            int64_t K, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  const int rc = check_gemm_args(A, lda, B_hi, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, 0);
  if (rc != SN_OK || M == 0) return rc;
  if (!B_lo || !aligned16(B_lo) || !R) return SN_ERR_ARG;
  if (R2 && (ldr2 < N || ldr2 % 4 || !aligned16(R2))) return SN_ERR_UNSUPPORTED;
  return launch_gemm(A, lda, B_hi, B_lo, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, 0,
                     (cudaStream_t)stream, nullptr, 0, nullptr, false, row_scale, R2, ldr2, true);
}

SN_API int sn_gemm_nt_wide_tf32_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb, float* C,
                                    int64_t ldc, int64_t M, int64_t N, int64_t K, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  if (M < 0 || N <= 0 || K <= 0) return SN_ERR_ARG;
  if (M == 0) return SN_OK;
  if (!A || !B_hi || !B_lo || !C || lda < K || ldb < K || ldc < N) return SN_ERR_ARG;
  if (K % 4 || N % 4 || N < 128 || K < kBlockK || K > kATmemStages * kBlockK || lda % 4 || ldb % 4 || ldc % 4 || !aligned16(A) || !aligned16(B_hi) ||
      !aligned16(B_lo) || !aligned16(C) || M >= 0x7fffffffLL - kBlockM || N >= 0x7fffffffLL - 128)
    return SN_ERR_UNSUPPORTED;
  return launch_gemm(A, lda, B_hi, B_lo, ldb, nullptr, nullptr, 0, nullptr, nullptr, 0, C, ldc, M, N, K, 0, (cudaStream_t)stream,
                     nullptr, 0, nullptr, true);
}

SN_API size_t sn_gemm_act_ws_bytes(int64_t N) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return N <= 0 ? 0 : (size_t)sms * 2 * (size_t)N * sizeof(float);
}

SN_API int sn_gemm_tf32_presplit_act_f32(const float* A, int64_t lda, const float* B_hi, const float* B_lo, int64_t ldb,
                                         const float* bias, const float* R, int64_t ldr, const float* rscale,
                                         const float* group_bias, int64_t rows_per_group, float* C, int64_t ldc, float* C_act,
                                         int64_t ldc_act, float* act_mean, float* act_var, int64_t M, int64_t N, int64_t K,
                                         int flags, void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  if (!C_act || ldc_act < N || (act_mean == nullptr) != (act_var == nullptr)) return SN_ERR_ARG;
  // the raw output is optional here: validate the shapes against the activated destination when it is absent
  const int rc = check_gemm_args(A, lda, B_hi, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C ? C : C_act,
                                 C ? ldc : ldc_act, M, N, K, flags);
  if (rc != SN_OK || M == 0) return rc;
  if (!B_lo || !aligned16(B_lo) || (flags & (SN_GEMM_SINGLE_PASS | SN_GEMM_ELU_BWD_LEFT | SN_GEMM_LEGACY_SS))) return SN_ERR_ARG;
  if (ldc_act % 4 || !aligned16(C_act)) return SN_ERR_UNSUPPORTED;
  float* partial = nullptr;
  if (act_mean) {
    if (!ws || ws_bytes < sn_gemm_act_ws_bytes(N) || !aligned16(ws)) return SN_ERR_WORKSPACE;
    partial = reinterpret_cast<float*>(ws);
  }
  const int rc2 = launch_gemm(A, lda, B_hi, B_lo, ldb, bias, R, ldr, rscale, group_bias, rows_per_group, C, ldc, M, N, K, flags,
                              (cudaStream_t)stream, C_act, ldc_act, partial);
  if (rc2 != SN_OK || !act_mean) return rc2;
  return launch_colstats_final(partial, (int)gemm_grid(M), M, (int)N, nullptr, act_mean, act_var, (cudaStream_t)stream);
}

SN_API int sn_split_tf32_f32(const float* X, int64_t ldx, int64_t rows, int64_t cols, float* hi, float* lo, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::gemm;
  if (rows <= 0 || cols <= 0 || !X || !hi || !lo || ldx < cols || rows * cols > 0x7fffffffLL) return SN_ERR_ARG;
  split_weights_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(X, ldx, hi, lo, (int)rows, (int)cols);
  return launch_status();
}
