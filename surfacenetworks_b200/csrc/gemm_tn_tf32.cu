// gemm_tn_tf32.cu -- the weight-gradient product of a stage on the tensor cores:
//
//   G[M x N] = A[R x M]^T * B[R x N]            (reduction over the R = B*V or B*F rows; M = 128, N <= 256)
//
// In the backward of Linear(BatchNorm(Z)) (reference GraphConv1x1, src/utils/utils_pt.py:91-104) this one product,
// G = dY^T Z, yields dW and -- together with colsum(dY) -- everything BatchNorm's backward needs (see fused.py).
// The reference gets it from autograd as an fp32 SIMT GEMM plus two BatchNorm reduction passes.
//
// Both operands are "MN-major" for the tensor core (the contraction index R is the slow axis of dY and Z), so the
// tiles are loaded as 32-row x 32-float TMA boxes (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) and described to tcgen05 with
// MN-major SWIZZLE_128B_BASE32B descriptors (LBO = 4 KB between 32-wide MN atoms, SBO = 512 B between 4-row K atoms).  Split-K: every CTA owns a
// contiguous range of rows, accumulates its [128 x N] partial in TMEM (3xTF32: hi/lo split of both tiles in shared
// memory) and writes it to a workspace; a second kernel adds the partials in a fixed order (deterministic).
#include <cuda.h>

#include "common.cuh"

SN_API size_t sn_gemm_tn_tf32_ws_bytes(int64_t R, int64_t N);

namespace sn {
namespace gemm_tn {

constexpr int kM = 128;
constexpr int kBlockK = 32;            // rows (contraction) per stage
constexpr int kUmmaK = 8;
constexpr int kThreads = 512;          // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-11 split, warps 12-15 epilogue
constexpr int kBoxBytes = 32 * 32 * 4; // one TMA box: 32 rows x 128 bytes

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
constexpr int kPrefetchDist = 4;       // k-blocks of L2 prefetch distance
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// MN-major tf32 operands have exactly one legal shared-memory layout on sm_100: SWIZZLE_128B_BASE32B
// (CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available smem layout").  Its atom is 4 K-rows of
// 128 bytes (32 MN elements), 32-byte chunks XOR-ed with the row index mod 4 -- the pattern TMA produces with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  One MMA (K = 8) spans two atoms: SBO = 512 bytes; the next 32-wide MN atom
// is the next TMA box: LBO = 4096 bytes.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)(kBoxBytes >> 4) << 16;         // leading byte offset
  d |= (uint64_t)(512 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                        // layout type: SWIZZLE_128B_BASE32B
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
  float* partial;        // [grid][kM][N]
  int R, N;
  int kb_per_cta;        // 32-row blocks per CTA
  int n_kb;
  int split;
  int l2_prefetch;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_ss_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N;
  const int a_boxes = kM / 32, b_boxes = N / 32;
  const uint32_t a_bytes = a_boxes * kBoxBytes, b_bytes = b_boxes * kBoxBytes;
  const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);        // A, A_lo, B, B_lo
  constexpr int kStages = 2;
  __shared__ uint64_t full_bar[kStages], ready_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.x * p.kb_per_cta;
  const int kb1 = min(kb0 + p.kb_per_cta, p.n_kb);
  const int my_kb = max(kb1 - kb0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(ready_bar + s, 8);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem_cols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1u);
        unsigned char* sa = smem + (size_t)stage * stage_bytes;
        unsigned char* sb = sa + 2 * a_bytes;
        mbar_arrive_expect_tx(full_bar + stage, a_bytes + b_bytes);
        for (int i = 0; i < a_boxes; ++i) tma_load_2d(sa + i * kBoxBytes, &map_a, i * 32, kb * kBlockK, full_bar + stage);
        for (int i = 0; i < b_boxes; ++i) tma_load_2d(sb + i * kBoxBytes, &map_b, i * 32, kb * kBlockK, full_bar + stage);
        // two-stage ring (the hi/lo twins fill shared memory): too shallow for DRAM latency, so the k-blocks this CTA
        // loads kPrefetchDist iterations from now are pulled into L2 here
        if (p.l2_prefetch && kb + kPrefetchDist < kb1) {
          for (int i = 0; i < a_boxes; ++i) tma_prefetch_l2_2d(&map_a, i * 32, (kb + kPrefetchDist) * kBlockK);
          for (int i = 0; i < b_boxes; ++i) tma_prefetch_l2_2d(&map_b, i * 32, (kb + kPrefetchDist) * kBlockK);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // D = F32, A = B = TF32, both MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(kM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < my_kb; ++kb) {
      mbar_wait(ready_bar + stage, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sal = sa + a_bytes, sb = sa + 2 * a_bytes, sbl = sb + b_bytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          const uint32_t off = k * 1024;                       // 8 K rows = one 1 KB group inside every box
          const uint64_t da = umma_desc_mn_sw128(sa + off), db = umma_desc_mn_sw128(sb + off);
          umma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
          if (p.split) {
            umma_tf32(tmem_base, da, umma_desc_mn_sw128(sbl + off), idesc, true);
            umma_tf32(tmem_base, umma_desc_mn_sw128(sal + off), db, idesc, true);
          }
        }
        umma_commit(empty_bar + stage);
        if (kb == my_kb - 1) umma_commit(&done_bar);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 4 && warp < 12) {
    // hi / lo split of both tiles (elementwise, layout agnostic): 256 threads
    const int t = threadIdx.x - 128;
    int stage = 0;
    uint32_t phase = 0;
    const int n_a = a_bytes / 16, n_b = b_bytes / 16;
    for (int kb = 0; kb < my_kb; ++kb) {
      mbar_wait(full_bar + stage, phase);
      unsigned char* sa = smem + (size_t)stage * stage_bytes;
      unsigned char* sb = sa + 2 * a_bytes;
      for (int i = t; i < n_a + n_b; i += 256) {
        unsigned char* src = i < n_a ? sa + (size_t)i * 16 : sb + (size_t)(i - n_a) * 16;
        const uint32_t lo_off = i < n_a ? a_bytes : b_bytes;
        const float4 x = *reinterpret_cast<float4*>(src);
        float4 hi;
        hi.x = to_tf32(x.x); hi.y = to_tf32(x.y); hi.z = to_tf32(x.z); hi.w = to_tf32(x.w);
        *reinterpret_cast<float4*>(src) = hi;
        if (p.split) *reinterpret_cast<float4*>(src + lo_off) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(ready_bar + stage);
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= 12) {
    // epilogue: this CTA's partial [128 x N] -> workspace
    const int ew = warp - 12;
    const int row = ew * 32 + lane;
    float* out = p.partial + ((size_t)blockIdx.x * kM + row) * N;
    if (my_kb > 0) {
      mbar_wait(&done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    } else {
      for (int c0 = 0; c0 < N; c0 += 4) *reinterpret_cast<float4*>(out + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

// ====================================================================================================================
// Round-2 kernel: dY (the A operand, M = 128 columns of dY) goes through TENSOR MEMORY, and colsum(dY) comes for free
// ====================================================================================================================
// The round-1 kernel above split BOTH tiles in shared memory (48 KB read + 96 KB written per 32-row k-block) and fed both
// operands to the tensor core from shared memory (144 KB of operand reads): 336 KB of shared-memory traffic per k-block
// against 48 KB of HBM traffic, two pipeline stages.  Here the dY tile lands un-swizzled ([32 rows][128 columns]);
// thread m of the four A-split warps reads column m (32 conflict-free LDS.32), which is exactly row m of the tensor
// core's A operand, and writes hi / lo into tensor memory (tcgen05.st): A never returns to shared memory.  While the
// column passes through its registers the thread also adds it up: colsum(dY) -- which the BatchNorm backward needs next
// to G (fused.py) and which used to be a separate HBM pass over dY (sn_colstats_f32) -- costs one FADD per element.
// Z (the B operand, MN-major) is still split in shared memory by eight warps.
constexpr int kATmemStages = 4;
constexpr int kATmemCol0 = 256;
constexpr int kMaxAStages = 4;         // shared-memory dY ring (16 KB stages)
constexpr int kMaxRawStages = 6;
constexpr int kMaxLoStages = 4;

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_tf32_ts_lohi(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t bdesc_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 d;\nsetp.ne.b32 p, %5, 0;\nmov.b64 d, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct ParamsTS {
  float* partial;        // [grid][kM][N]
  float* csum_partial;   // [grid][kM]   per-CTA column sums of A (dY)
  int R, N;
  int kb_per_cta;
  int n_kb;
  int raw_stages;        // Z tiles as TMA lands them (= the hi operand: the tensor core truncates fp32 to tf32)
  int lo_stages;         // Z lo tiles written by the split warps
  int a_stages;          // dY tiles
  int split;
};

// Z (the B operand) needs no hi copy: tcgen05 reads an fp32 container as tf32 by TRUNCATING the low 13 mantissa bits
// (measured: tests/test_gpu_gemm.py::test_tf32_operand_truncation_probe), so the TMA-landed tile IS the hi operand and
// the split warps only write lo = tf32(x - trunc(x)) into a second ring.  The raw ring (3 stages) is deeper than the lo
// ring (2): TMA runs ahead of the split, which the in-place hi/lo scheme of the first TS version could not (ncu: the
// split warps sat on the `full` barrier of a two-stage ring).
//
// Roles: warp 0 = dY producer (TMA), warp 3 = Z producer (TMA), warp 1 = MMA issuer, warp 2 = TMEM allocation,
// warps 4-11 = dY: shared memory -> column sums, hi / lo -> tensor memory (warps w and w + 4 share the TMEM lane quarter
// (w - 4) % 4 and take the two 16-row halves of the 32-row k-block: the per-k-block latency chain LDS -> cvt ->
// tcgen05.st -> wait::st of a single warp per quarter was the kernel's period, ~2600 clk against 1090 clk of MMA time);
// warps 4-7 then run the epilogue; warps 12-15 = Z lo tiles.
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const ParamsTS p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N;
  const int b_boxes = N / 32;
  constexpr uint32_t a_bytes = kBlockK * kM * 4;               // 16 KB: [32 rows][128 columns], no swizzle
  const uint32_t b_bytes = (uint32_t)b_boxes * kBoxBytes;
  unsigned char* raw_ring = smem;                              // 1024-byte aligned boxes first
  unsigned char* lo_ring = raw_ring + (size_t)p.raw_stages * b_bytes;
  unsigned char* a_ring = lo_ring + (size_t)p.lo_stages * b_bytes;
  __shared__ uint64_t a_full[kMaxAStages], a_free[kMaxAStages], a_ready[kATmemStages], a_tfree[kATmemStages];
  __shared__ uint64_t raw_full[kMaxRawStages], raw_free[kMaxRawStages], lo_ready[kMaxLoStages], lo_free[kMaxLoStages], done_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_csum[kM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.x * p.kb_per_cta;
  const int kb1 = min(kb0 + p.kb_per_cta, p.n_kb);
  const int my_kb = max(kb1 - kb0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(a_full + s, 1);
      mbar_init(a_free + s, 8);
    }
    for (int s = 0; s < kATmemStages; ++s) {
      mbar_init(a_ready + s, 8);
      mbar_init(a_tfree + s, 1);
    }
    for (int s = 0; s < p.raw_stages; ++s) {
      mbar_init(raw_full + s, 1);
      mbar_init(raw_free + s, 1);
    }
    for (int s = 0; s < p.lo_stages; ++s) {
      mbar_init(lo_ready + s, 4);
      mbar_init(lo_free + s, 1);
    }
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(a_free + stage, phase ^ 1u);
        mbar_arrive_expect_tx(a_full + stage, a_bytes);
        tma_load_2d(a_ring + (size_t)stage * a_bytes, &map_a, 0, kb * kBlockK, a_full + stage);
        if (++stage == p.a_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(raw_free + stage, phase ^ 1u);
        unsigned char* sb = raw_ring + (size_t)stage * b_bytes;
        mbar_arrive_expect_tx(raw_full + stage, b_bytes);
        for (int i = 0; i < b_boxes; ++i) tma_load_2d(sb + i * kBoxBytes, &map_b, i * 32, kb * kBlockK, raw_full + stage);
        if (++stage == p.raw_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // D = F32, A = B = TF32, A K-major (tensor memory), B MN-major (bit 16), N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
      const uint64_t d0 = umma_desc_mn_sw128(smem_u32(raw_ring));
      const uint32_t d_raw0 = (uint32_t)d0, d_hi = (uint32_t)(d0 >> 32);
      const uint32_t d_lo0 = (uint32_t)umma_desc_mn_sw128(smem_u32(lo_ring));
      const uint32_t stage_units = b_bytes >> 4;
      const uint32_t split = (uint32_t)p.split;
      int rs = 0, ls = 0;
      uint32_t rph = 0, lph = 0;
      for (int kb = 0; kb < my_kb; ++kb) {
        const uint32_t at = (uint32_t)kb % kATmemStages;
        mbar_wait(a_ready + at, ((uint32_t)kb / kATmemStages) & 1u);
        mbar_wait(raw_full + rs, rph);
        if (split) mbar_wait(lo_ready + ls, lph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ta = tmem_base + (uint32_t)(kATmemCol0 + at * 64);
        const uint32_t dh = d_raw0 + (uint32_t)rs * stage_units, dl = d_lo0 + (uint32_t)ls * stage_units;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // 8 K rows = one 1 KB group inside every box = 64 descriptor units
          umma_tf32_ts_lohi(tmem_base, ta + k * kUmmaK, dh + 64 * k, d_hi, idesc, (uint32_t)((kb | k) != 0));   // hi * hi
          if (split) {
            umma_tf32_ts_lohi(tmem_base, ta + k * kUmmaK, dl + 64 * k, d_hi, idesc, 1u);                          // hi * lo
            umma_tf32_ts_lohi(tmem_base, ta + 32 + k * kUmmaK, dh + 64 * k, d_hi, idesc, 1u);                     // lo * hi
          }
        }
        umma_commit(raw_free + rs);
        if (split) umma_commit(lo_free + ls);
        umma_commit(a_tfree + at);
        if (kb == my_kb - 1) umma_commit(&done_bar);
        if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
        if (++ls == p.lo_stages) { ls = 0; lph ^= 1u; }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ dY: column m of the tile = row m of the A operand
    const int q = (warp - 4) & 3, kh = (warp - 4) >> 2;        // TMEM lane quarter, 16-row half of the k-block
    const int m = q * 32 + lane;                               // 0..127 = TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_ring_s = smem_u32(a_ring);
    float csum = 0.f;
    int as = 0;
    uint32_t aph = 0;
    for (int kb = 0; kb < my_kb; ++kb) {
      mbar_wait(a_full + as, aph);
      const uint32_t col = a_ring_s + (uint32_t)as * a_bytes + (uint32_t)m * 4u + (uint32_t)(kh * 16) * (kM * 4);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {                           // lanes read consecutive words of one row: conflict-free
        const float x = lds_f1(col + (uint32_t)k * (kM * 4));
        csum += x;                                             // rows past R are zero-filled by TMA
        const float h = to_tf32(x);
        hi[k] = __float_as_uint(h);
        lo[k] = __float_as_uint(to_tf32(x - h));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_free + as);
      const uint32_t at = (uint32_t)kb % kATmemStages;
      mbar_wait(a_tfree + at, (((uint32_t)kb / kATmemStages) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem_base + lane_base + (uint32_t)(kATmemCol0 + at * 64 + kh * 16);
      tmem_st16(ta, hi);
      if (p.split) tmem_st16(ta + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready + at);
      if (++as == p.a_stages) { as = 0; aph ^= 1u; }
    }
    // column sums: the two k-halves of a column are added in a fixed order (lower half first)
    if (kh == 1) s_csum[m] = csum;
    asm volatile("bar.sync 1, 256;" ::: "memory");             // the eight dY warps only
    if (kh == 1) goto tn_done;
    csum += s_csum[m];
    // ------------------------------------------------------------------ epilogue: this CTA's partial [128 x N] -> workspace
    p.csum_partial[(size_t)blockIdx.x * kM + m] = csum;
    float* out = p.partial + ((size_t)blockIdx.x * kM + m) * N;
    if (my_kb > 0) {
      mbar_wait(&done_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    } else {
      for (int c0 = 0; c0 < N; c0 += 4) *reinterpret_cast<float4*>(out + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (warp >= 12 && p.split) {
    // ------------------------------------------------------------------ Z: lo = tf32(x - trunc(x)) into the lo ring
    const int t = threadIdx.x - 384;                           // 0..127
    const uint32_t raw_s = smem_u32(raw_ring), lo_s = smem_u32(lo_ring);
    const int n_f4 = (int)(b_bytes / 16);
    int rs = 0, ls = 0;
    uint32_t rph = 0, lph = 0;
    for (int kb = 0; kb < my_kb; ++kb) {
      mbar_wait(raw_full + rs, rph);
      mbar_wait(lo_free + ls, lph ^ 1u);
      const uint32_t src = raw_s + (uint32_t)rs * b_bytes, dst = lo_s + (uint32_t)ls * b_bytes;
      for (int i0 = t; i0 < n_f4; i0 += 4 * 128) {             // four independent 16-byte loads in flight per thread
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * 128 < n_f4) x[u] = lds_f4(src + (uint32_t)(i0 + u * 128) * 16u);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + u * 128 < n_f4) {
            float4 l;
            l.x = to_tf32(x[u].x - trunc_tf32(x[u].x)); l.y = to_tf32(x[u].y - trunc_tf32(x[u].y));
            l.z = to_tf32(x[u].z - trunc_tf32(x[u].z)); l.w = to_tf32(x[u].w - trunc_tf32(x[u].w));
            sts_f4(dst + (uint32_t)(i0 + u * 128) * 16u, l);
          }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) mbar_arrive(lo_ready + ls);
      if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
      if (++ls == p.lo_stages) { ls = 0; lph ^= 1u; }
    }
  }

tn_done:
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// G[m][n] = sum over CTAs of partial[cta][m][n] and colsum[m] = sum over CTAs of csum[cta][m], both in a fixed order.
// CTA = 32 output quads (128 consecutive outputs, float4 each) x 8 partial groups: group j adds the partial rows j, j + 8, ...
// with up to six independent 16-byte loads in flight (a warp reads 512 contiguous bytes of one partial row), the eight group
// sums are added in order.  (First version: four lanes per scalar output, 32-byte segments, nine dependent batches: 9 us for
// 19 MB; the launch is latency, not bandwidth.)
__global__ void __launch_bounds__(256)
reduce_partials4_kernel(const float* __restrict__ partial, const float* __restrict__ csum_partial, int n_partials, int MN,
                        float* __restrict__ G, int64_t ldg, int N, float* __restrict__ colsum, int M) {
  __shared__ float4 red[8][32];
  const int q = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int n_gq = MN / 4;                                   // output quads of G (N % 4 == 0)
  const int gq = blockIdx.x * 32 + q;                        // this thread's quad: of G, then of the column sums
  const bool is_g = gq < n_gq, is_c = !is_g && colsum != nullptr && gq < n_gq + M / 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (is_g || is_c) {
    const float4* src = is_g ? reinterpret_cast<const float4*>(partial) + gq
                             : reinterpret_cast<const float4*>(csum_partial) + (gq - n_gq);
    const size_t stride = is_g ? (size_t)n_gq : (size_t)(M / 4);
    int c = j;
    for (; c + 5 * 8 < n_partials; c += 6 * 8) {
      float4 v[6];
#pragma unroll
      for (int u = 0; u < 6; ++u) v[u] = __ldg(src + (size_t)(c + u * 8) * stride);
#pragma unroll
      for (int u = 0; u < 6; ++u) acc = add4(acc, v[u]);
    }
    for (; c < n_partials; c += 8) acc = add4(acc, __ldg(src + (size_t)c * stride));
  }
  red[j][q] = acc;
  __syncthreads();
  if (j == 0 && (is_g || is_c)) {
    float4 tot = red[0][q];
#pragma unroll
    for (int k = 1; k < 8; ++k) tot = add4(tot, red[k][q]);
    if (is_g) {
      const int gi = gq * 4;
      *reinterpret_cast<float4*>(G + (int64_t)(gi / N) * ldg + (gi % N)) = tot;
    } else {
      *reinterpret_cast<float4*>(colsum + (gq - n_gq) * 4) = tot;
    }
  }
}

// G[m][n] = sum over CTAs of partial[cta][m][n], fixed order
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_partials, int MN, float* __restrict__ G,
                                       int64_t ldg, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float acc = 0.f;
  for (int c = 0; c < n_partials; ++c) acc += partial[(size_t)c * MN + i];
  G[(int64_t)(i / N) * ldg + (i % N)] = acc;
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
static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// dY tile for the TS kernel: one un-swizzled box of 32 rows x 128 columns (row-major in shared memory)
static bool make_map_plain(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kM, (cuuint32_t)kBlockK};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int grid_for(int64_t n_kb, int* kb_per_cta) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t per = (n_kb + sms - 1) / sms;
  if (per < 1) per = 1;
  *kb_per_cta = (int)per;
  return (int)((n_kb + per - 1) / per);
}

static int launch_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* G, int64_t ldg, float* colsum_A,
                     int64_t R, int64_t M, int64_t N, int flags, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (R <= 0 || M <= 0 || N <= 0 || !A || !B || !G || lda < M || ldb < N || ldg < N) return SN_ERR_ARG;
  if (M != kM || N % 32 != 0 || N > 256 || N < 32 || R >= 0x7fffffffLL - 64) return SN_ERR_UNSUPPORTED;
  if (lda % 4 || ldb % 4 || ldg % 4 || !aligned16(A) || !aligned16(B) || !aligned16(G) || (colsum_A && !aligned16(colsum_A)))
    return SN_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < sn_gemm_tn_tf32_ws_bytes(R, N)) return SN_ERR_WORKSPACE;
  CUtensorMap map_a, map_b;
  if (!make_map(&map_b, B, R, N, ldb)) return SN_ERR_UNSUPPORTED;
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int n_kb = (int)((R + kBlockK - 1) / kBlockK);
  int kb_per_cta;
  const int grid = grid_for(n_kb, &kb_per_cta);
  const int split = (flags & SN_GEMM_SINGLE_PASS) ? 0 : 1;
  const int l2_prefetch = (flags & SN_GEMM_NO_L2_PREFETCH) ? 0 : 1;
  const int MN = (int)(kM * N);
  float* csum_partial = partial + (size_t)grid * MN;

  if (flags & SN_GEMM_LEGACY_SS) {       // round-1 kernel (both operands split in and read from shared memory)
    if (colsum_A) return SN_ERR_UNSUPPORTED;
    if (!make_map(&map_a, A, R, M, lda)) return SN_ERR_UNSUPPORTED;
    Params p;
    p.partial = partial;
    p.R = (int)R; p.N = (int)N;
    p.n_kb = n_kb; p.kb_per_cta = kb_per_cta;
    p.split = split; p.l2_prefetch = l2_prefetch;
    const size_t stage_bytes = 2 * ((size_t)(kM / 32) * kBoxBytes + (size_t)(N / 32) * kBoxBytes);
    const size_t smem = 2 * stage_bytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gemm_tn_ss_kernel<<<grid, kThreads, smem, st>>>(map_a, map_b, p);
    reduce_partials_kernel<<<(MN + 255) / 256, 256, 0, st>>>(p.partial, grid, MN, G, ldg, (int)N);
    return launch_status();
  }

  if (!make_map_plain(&map_a, A, R, M, lda)) return SN_ERR_UNSUPPORTED;
  ParamsTS p;
  p.partial = partial;
  p.csum_partial = csum_partial;
  p.R = (int)R; p.N = (int)N;
  p.n_kb = n_kb; p.kb_per_cta = kb_per_cta;
  p.split = split;
  int dev = 0, smem_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  // shared memory: lo ring 2 stages, dY ring 3 stages, the rest (up to 6 stages) to the raw Z ring that TMA fills
  const size_t a_stage = (size_t)kBlockK * kM * 4, b_stage = (size_t)(N / 32) * kBoxBytes;
  p.lo_stages = 2;
  p.a_stages = 3;
  const size_t budget = (size_t)smem_optin - 2048 - 1024;
  int raw_stages = (int)((budget - p.lo_stages * b_stage - p.a_stages * a_stage) / b_stage);
  if (raw_stages > kMaxRawStages) raw_stages = kMaxRawStages;
  if (raw_stages < 2) return SN_ERR_UNSUPPORTED;
  p.raw_stages = raw_stages;
  if (raw_stages >= 5 && p.a_stages < kMaxAStages) p.a_stages = kMaxAStages;
  const size_t smem = (size_t)(p.raw_stages + p.lo_stages) * b_stage + p.a_stages * a_stage + 1024;
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  gemm_tn_ts_kernel<<<grid, kThreads, smem, st>>>(map_a, map_b, p);
  const int out_quads = (MN + (colsum_A ? (int)kM : 0)) / 4;
  reduce_partials4_kernel<<<(out_quads + 31) / 32, 256, 0, st>>>(partial, csum_partial, grid, MN, G, ldg, (int)N, colsum_A, (int)kM);
  return launch_status();
}

}  // namespace gemm_tn
}  // namespace sn

SN_API size_t sn_gemm_tn_tf32_ws_bytes(int64_t R, int64_t N) {
  using namespace sn::gemm_tn;
  if (R <= 0 || N <= 0) return 0;
  int per;
  const int grid = grid_for((R + kBlockK - 1) / kBlockK, &per);
  return (size_t)grid * kM * (size_t)(N + 1) * sizeof(float) + 256;      // partial products + partial column sums
}

SN_API int sn_gemm_tn_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* G, int64_t ldg, int64_t R,
                               int64_t M, int64_t N, int flags, void* ws, size_t ws_bytes, sn_stream_t stream) {
  return sn::gemm_tn::launch_tn(A, lda, B, ldb, G, ldg, nullptr, R, M, N, flags, ws, ws_bytes, (cudaStream_t)stream);
}

SN_API int sn_gemm_tn_colsum_tf32_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* G, int64_t ldg,
                                      float* colsum_A, int64_t R, int64_t M, int64_t N, int flags, void* ws, size_t ws_bytes,
                                      sn_stream_t stream) {
  if (!colsum_A) return SN_ERR_ARG;
  return sn::gemm_tn::launch_tn(A, lda, B, ldb, G, ldg, colsum_A, R, M, N, flags, ws, ws_bytes, (cudaStream_t)stream);
}
