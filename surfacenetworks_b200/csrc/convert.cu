// convert.cu -- format layer: torch-style COO (int64) -> CSR32 -> BSR4.
//
// Replaces the reference's COO->batched-CSR kernel src/utils/cuda/batch_csr.cu:13-47 (launcher
// batch_csr.py:28-59).  Differences that matter:
//   * empty rows ANYWHERE are handled (the reference zero-fills col_ptr and only writes non-empty rows,
//     so an interior empty row corrupts its predecessor's range -- batch_csr.py:48-49, batch_csr.cu:36-42);
//   * 32-bit indices (half the index bytes of the reference's int64 col_ind / col_ptr);
//   * unsorted input is accepted (counting sort by row + per-row ordering by (col, input position));
//   * the 3-D [B, R, C] batched layout is flattened to the block-diagonal 2-D operator so one SpMM launch
//     covers the whole batch (what sparse_diag_cat does on the CPU, utils_pt.py:41-53).
// The conversion runs once per operator; the per-step hot path only sees CSR32 / BSR4.
#include "common.cuh"

namespace sn {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t flat_index(const int64_t* __restrict__ batch, const int64_t* __restrict__ idx,
                                              int64_t i, int64_t per_batch) {
  int64_t v = idx[i];
  if (batch) v += batch[i] * per_batch;
  return v;
}

// ---------------------------------------------------------------- block-wide exclusive scan helpers
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /*[32]*/, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += n;
  }
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += n;
    }
    smem[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  total = smem[32];
  int res = inc - v + smem[wid];
  __syncthreads();
  return res;
}

// phase 1: per-tile sums
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int* __restrict__ in, int64_t n, int* __restrict__ tile_sums) {
  __shared__ int sm[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_exclusive_scan(s, sm, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// phase 2: one CTA scans the tile sums in place (exclusive), any count
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets(int* __restrict__ tile_sums, int n_tiles) {
  __shared__ int sm[33];
  int carry = 0;
  for (int base = 0; base < n_tiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    const int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sm, total);
    if (i < n_tiles) tile_sums[i] = ex + carry;
    carry += total;
  }
}
// phase 3: exclusive scan inside each tile + tile offset.  out may alias in.  out[n] (one past) = total.
__global__ void __launch_bounds__(kScanThreads) scan_apply(const int* __restrict__ in, int64_t n, const int* __restrict__ tile_offsets,
                                                            int* __restrict__ out) {
  __shared__ int sm[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int ex = block_exclusive_scan(s, sm, total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
    if (base + i == n - 1) out[n] = ex;
  }
}

int64_t exclusive_scan_tiles(int64_t n) { return ceil_div(n, kScanTile); }

// Exclusive scan of counts[0..n) into out[0..n], out[n] = total.  tile_ws: exclusive_scan_tiles(n) ints.
int exclusive_scan(const int* counts, int64_t n, int* out, int* tile_ws, cudaStream_t st) {
  const int64_t tiles = ceil_div(n, kScanTile);
  if (tiles > 0x7fffffffLL) return SN_ERR_OVERFLOW;
  scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, st>>>(counts, n, tile_ws);
  scan_tile_offsets<<<1, kScanThreads, 0, st>>>(tile_ws, (int)tiles);
  scan_apply<<<(unsigned)tiles, kScanThreads, 0, st>>>(counts, n, tile_ws, out);
  return launch_status();
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- COO -> CSR32, sorted input
__global__ void fill_i32(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256)
coo_sorted_to_csr(const int64_t* __restrict__ batch, const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                  const float* __restrict__ val, int64_t nnz, int64_t rpb, int64_t cpb, int64_t n_rows,
                  int32_t* __restrict__ rowptr, int32_t* __restrict__ colind, float* __restrict__ out_val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int64_t r = flat_index(batch, row, i, rpb);
  const int64_t rprev = i > 0 ? flat_index(batch, row, i - 1, rpb) : -1;
  colind[i] = (int32_t)flat_index(batch, col, i, cpb);
  out_val[i] = val[i];
  // first entry of row r also opens every empty row between the previous entry's row and r
  for (int64_t rr = rprev + 1; rr <= r; ++rr) rowptr[rr] = (int32_t)i;
  if (i == nnz - 1)
    for (int64_t rr = r + 1; rr <= n_rows; ++rr) rowptr[rr] = (int32_t)nnz;
}

// ---------------------------------------------------------------- COO -> CSR32, arbitrary order
__global__ void __launch_bounds__(256)
coo_row_histogram(const int64_t* __restrict__ batch, const int64_t* __restrict__ row, int64_t nnz, int64_t rpb,
                  int* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) atomicAdd(counts + flat_index(batch, row, i, rpb), 1);
}
__global__ void __launch_bounds__(256)
coo_scatter_perm(const int64_t* __restrict__ batch, const int64_t* __restrict__ row, int64_t nnz, int64_t rpb,
                 const int32_t* __restrict__ rowptr, int* __restrict__ cursor, int32_t* __restrict__ perm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int64_t r = flat_index(batch, row, i, rpb);
  const int pos = rowptr[r] + atomicAdd(cursor + r, 1);
  perm[pos] = (int32_t)i;
}
// order each row's entries by (col, input position): makes the result independent of atomic order
__global__ void __launch_bounds__(128)
coo_sort_rows(const int64_t* __restrict__ batch, const int64_t* __restrict__ col, int64_t cpb, int64_t n_rows,
              const int32_t* __restrict__ rowptr, int32_t* __restrict__ perm) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int s = rowptr[r], e = rowptr[r + 1];
  auto key_less = [&](int32_t a, int32_t b) {
    const int64_t ca = flat_index(batch, col, a, cpb), cb = flat_index(batch, col, b, cpb);
    return ca < cb || (ca == cb && a < b);
  };
  const int n = e - s;
  int32_t* p = perm + s;
  if (n <= 32) {  // insertion sort
    for (int i = 1; i < n; ++i) {
      const int32_t x = p[i];
      int j = i - 1;
      while (j >= 0 && key_less(x, p[j])) {
        p[j + 1] = p[j];
        --j;
      }
      p[j + 1] = x;
    }
  } else {  // in-place heapsort for the rare long row
    auto sift = [&](int root, int size) {
      while (true) {
        int child = 2 * root + 1;
        if (child >= size) break;
        if (child + 1 < size && key_less(p[child], p[child + 1])) ++child;
        if (!key_less(p[root], p[child])) break;
        const int32_t tmp = p[root];
        p[root] = p[child];
        p[child] = tmp;
        root = child;
      }
    };
    for (int i = n / 2 - 1; i >= 0; --i) sift(i, n);
    for (int size = n - 1; size > 0; --size) {
      const int32_t tmp = p[0];
      p[0] = p[size];
      p[size] = tmp;
      sift(0, size);
    }
  }
}
__global__ void __launch_bounds__(256)
coo_gather(const int64_t* __restrict__ batch, const int64_t* __restrict__ col, const float* __restrict__ val, int64_t nnz,
           int64_t cpb, const int32_t* __restrict__ perm, int32_t* __restrict__ colind, float* __restrict__ out_val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int32_t src = perm[i];
  colind[i] = (int32_t)flat_index(batch, col, src, cpb);
  out_val[i] = val[src];
}

// ---------------------------------------------------------------- CSR32 -> BSR4
// One thread per block-row walks the four scalar rows with a 4-way merge on (col >> 2).
template <bool FILL>
__global__ void __launch_bounds__(128)
csr_to_bsr4_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind, const float* __restrict__ val,
                   int64_t n_brows, int* __restrict__ counts, const int32_t* __restrict__ browptr,
                   int32_t* __restrict__ bcolind, float* __restrict__ bval) {
  const int64_t br = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (br >= n_brows) return;
  int p[4], e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p[i] = rowptr[4 * br + i];
    e[i] = rowptr[4 * br + i + 1];
  }
  int out = FILL ? browptr[br] : 0;
  int n = 0;
  while (true) {
    int cur = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p[i] < e[i]) cur = min(cur, colind[p[i]] >> 2);
    if (cur == 0x7fffffff) break;
    float blk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) blk[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      while (p[i] < e[i] && (colind[p[i]] >> 2) == cur) {
        if (FILL) {
          const int qq = colind[p[i]] & 3;
          const float v = val[p[i]];
#pragma unroll
          for (int s = 0; s < 4; ++s)  // static indexing keeps blk[] in registers
            if (s == qq) blk[4 * s + ((i - s) & 3)] += v;  // column q, slot (p - q) mod 4
        }
        ++p[i];
      }
    }
    if (FILL) {
      bcolind[out] = cur;
      float4* dst = reinterpret_cast<float4*>(bval + (int64_t)out * 16);
#pragma unroll
      for (int s = 0; s < 4; ++s) dst[s] = make_float4(blk[4 * s], blk[4 * s + 1], blk[4 * s + 2], blk[4 * s + 3]);
      ++out;
    }
    ++n;
  }
  if (!FILL) counts[br] = n;
}

}  // namespace sn

// ==================================================================================== C ABI
SN_API size_t sn_coo_to_csr32_ws_bytes(int64_t nnz, int64_t n_rows) {
  using namespace sn;
  if (nnz < 0 || n_rows < 0) return 0;
  // counts/cursor [n_rows+1] + perm [nnz] + scan tiles
  return align_up((size_t)(n_rows + 1) * 4, 256) + align_up((size_t)nnz * 4, 256) +
         align_up((size_t)(ceil_div(n_rows + 1, kScanTile) + 1) * 4, 256) + 256;
}

SN_API int sn_coo_to_csr32(const int64_t* batch, const int64_t* row, const int64_t* col, const float* val,
                           int64_t nnz, int64_t rows_per_batch, int64_t cols_per_batch, int64_t n_rows,
                           int64_t n_cols, int flags, int32_t* rowptr, int32_t* colind, float* out_val,
                           void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (nnz < 0 || n_rows < 0 || n_cols < 0 || !rowptr) return SN_ERR_ARG;
  if (nnz > 0 && (!row || !col || !val || !colind || !out_val)) return SN_ERR_ARG;
  if (batch && (rows_per_batch <= 0 || cols_per_batch <= 0)) return SN_ERR_ARG;
  if (nnz > 0x7fffffffLL || n_rows >= 0x7fffffffLL || n_cols > 0x7fffffffLL) return SN_ERR_OVERFLOW;
  cudaStream_t st = (cudaStream_t)stream;
  if (nnz == 0) {
    fill_i32<<<(unsigned)ceil_div(n_rows + 1, 256), 256, 0, st>>>(rowptr, n_rows + 1, 0);
    return launch_status();
  }
  const unsigned grid = (unsigned)ceil_div(nnz, 256);
  if (flags & SN_COO_SORTED) {
    coo_sorted_to_csr<<<grid, 256, 0, st>>>(batch, row, col, val, nnz, rows_per_batch, cols_per_batch, n_rows,
                                              rowptr, colind, out_val);
    return launch_status();
  }
  if (ws_bytes < sn_coo_to_csr32_ws_bytes(nnz, n_rows) || !ws) return SN_ERR_WORKSPACE;
  char* w = (char*)ws;
  int* counts = (int*)w;
  w += align_up((size_t)(n_rows + 1) * 4, 256);
  int32_t* perm = (int32_t*)w;
  w += align_up((size_t)nnz * 4, 256);
  int* tiles = (int*)w;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(n_rows + 1) * 4, st);
  if (e != cudaSuccess) return (int)e;
  coo_row_histogram<<<grid, 256, 0, st>>>(batch, row, nnz, rows_per_batch, counts);
  int rc = exclusive_scan(counts, n_rows, rowptr, tiles, st);
  if (rc != SN_OK) return rc;
  e = cudaMemsetAsync(counts, 0, (size_t)(n_rows + 1) * 4, st);
  if (e != cudaSuccess) return (int)e;
  coo_scatter_perm<<<grid, 256, 0, st>>>(batch, row, nnz, rows_per_batch, rowptr, counts, perm);
  coo_sort_rows<<<(unsigned)ceil_div(n_rows, 128), 128, 0, st>>>(batch, col, cols_per_batch, n_rows, rowptr, perm);
  coo_gather<<<grid, 256, 0, st>>>(batch, col, val, nnz, cols_per_batch, perm, colind, out_val);
  return launch_status();
}

SN_API size_t sn_csr32_to_bsr4_ws_bytes(int64_t n_rows) {
  using namespace sn;
  if (n_rows < 0) return 0;
  const int64_t nb = n_rows / 4;
  return align_up((size_t)(nb + 1) * 4, 256) + align_up((size_t)(ceil_div(nb + 1, kScanTile) + 1) * 4, 256) + 256;
}

SN_API int sn_csr32_to_bsr4_count(const int32_t* rowptr, const int32_t* colind, int64_t n_rows, int32_t* browptr,
                                  void* ws, size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  if (n_rows < 0 || !rowptr || !browptr) return SN_ERR_ARG;
  if (n_rows % 4 != 0) return SN_ERR_UNSUPPORTED;
  if (n_rows >= 0x7fffffffLL) return SN_ERR_OVERFLOW;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nb = n_rows / 4;
  if (nb == 0) {
    fill_i32<<<1, 32, 0, st>>>(browptr, 1, 0);
    return launch_status();
  }
  if (!ws || ws_bytes < sn_csr32_to_bsr4_ws_bytes(n_rows)) return SN_ERR_WORKSPACE;
  char* w = (char*)ws;
  int* counts = (int*)w;
  w += align_up((size_t)(nb + 1) * 4, 256);
  int* tiles = (int*)w;
  csr_to_bsr4_kernel<false><<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(rowptr, colind, nullptr, nb, counts, nullptr,
                                                                         nullptr, nullptr);
  return exclusive_scan(counts, nb, browptr, tiles, st);
}

SN_API int sn_csr32_to_bsr4_fill(const int32_t* rowptr, const int32_t* colind, const float* val, int64_t n_rows,
                                 const int32_t* browptr, int32_t* bcolind, float* bval, sn_stream_t stream) {
  using namespace sn;
  if (n_rows < 0 || !rowptr || !browptr) return SN_ERR_ARG;
  if (n_rows % 4 != 0) return SN_ERR_UNSUPPORTED;
  const int64_t nb = n_rows / 4;
  if (nb == 0) return SN_OK;
  if (!colind || !val || !bcolind || !bval) return SN_ERR_ARG;
  if (!aligned16(bval)) return SN_ERR_UNSUPPORTED;
  csr_to_bsr4_kernel<true><<<(unsigned)ceil_div(nb, 128), 128, 0, (cudaStream_t)stream>>>(
      rowptr, colind, val, nb, nullptr, browptr, bcolind, bval);
  return launch_status();
}
