// assemble.cu -- GPU-resident batch assembly.
//
// The reference rebuilds the block-diagonal batch operator on the CPU every training step
// (sparse_diag_cat(...).coalesce(), src/utils/utils_pt.py:41-53, 1.6 s at the ARAP size) and uploads it
// (src/as_rigid_as_possible/main.py:172-183).  Here every mesh's operator is converted once and stays on the GPU;
// a batch is assembled by two small kernels that concatenate the per-mesh CSR32 / BSR4 arrays with row / column /
// block offsets and pad every mesh to the batch's common size (padded rows are empty) -- bit-identical to converting
// the host-assembled block-diagonal matrix.
#include "common.cuh"

namespace sn {

// per-mesh part table (device): 6 int64 per part
//   [0] rowptr pointer  [1] colind pointer  [2] value pointer  [3] rows in this part  [4] entries (nnz / blocks)
//   [5] entry offset of this part in the batch (exclusive prefix sum of [4])
constexpr int kPartWords = 6;

__global__ void __launch_bounds__(256)
assemble_rowptr_kernel(const int64_t* __restrict__ parts, int n_parts, int rows_pad, int64_t total_entries,
                       int32_t* __restrict__ rowptr_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_rows = (int64_t)n_parts * rows_pad;
  if (r > n_rows) return;
  if (r == n_rows) {
    rowptr_out[r] = (int32_t)total_entries;
    return;
  }
  const int i = (int)(r / rows_pad), lr = (int)(r - (int64_t)i * rows_pad);
  const int64_t* p = parts + (size_t)i * kPartWords;
  const int32_t* rp = reinterpret_cast<const int32_t*>(p[0]);
  const int64_t local = lr < p[3] ? (int64_t)rp[lr] : p[4];      // rows past the mesh's own count are empty
  rowptr_out[r] = (int32_t)(p[5] + local);
}

// VALS floats per entry (1 = CSR scalar, 16 = BSR 4x4 block); one thread per entry and float4 (or float) of values
template <int VALS>
__global__ void __launch_bounds__(256)
assemble_entries_kernel(const int64_t* __restrict__ parts, int n_parts, int cols_pad, int64_t total_entries,
                        int32_t* __restrict__ colind_out, float* __restrict__ val_out) {
  constexpr int TPE = VALS == 1 ? 1 : VALS / 4;                  // threads per entry
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = t / TPE;
  const int sub = (int)(t - k * TPE);
  if (k >= total_entries) return;
  int lo = 0, hi = n_parts - 1;                                  // last part whose entry offset is <= k
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (parts[(size_t)mid * kPartWords + 5] <= k) lo = mid; else hi = mid - 1;
  }
  const int64_t* p = parts + (size_t)lo * kPartWords;
  const int64_t lk = k - p[5];
  if (sub == 0) colind_out[k] = reinterpret_cast<const int32_t*>(p[1])[lk] + lo * cols_pad;
  if (VALS == 1) {
    val_out[k] = reinterpret_cast<const float*>(p[2])[lk];
  } else {
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p[2]) + lk * VALS) + sub;
    reinterpret_cast<float4*>(val_out + k * VALS)[sub] = *src;
  }
}

}  // namespace sn

SN_API int sn_assemble_block_diag(const int64_t* parts, int64_t n_parts, int64_t rows_pad, int64_t cols_pad,
                                  int64_t total_entries, int vals_per_entry, int32_t* rowptr_out, int32_t* colind_out,
                                  float* val_out, sn_stream_t stream) {
  using namespace sn;
  if (n_parts <= 0 || rows_pad <= 0 || cols_pad <= 0 || total_entries < 0 || !parts || !rowptr_out) return SN_ERR_ARG;
  if (vals_per_entry != 1 && vals_per_entry != 16) return SN_ERR_UNSUPPORTED;
  if (n_parts * rows_pad >= 0x7fffffffLL || n_parts * cols_pad >= 0x7fffffffLL || total_entries >= 0x7fffffffLL)
    return SN_ERR_OVERFLOW;
  if (total_entries > 0 && (!colind_out || !val_out)) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = n_parts * rows_pad;
  assemble_rowptr_kernel<<<(unsigned)ceil_div(n_rows + 1, 256), 256, 0, st>>>(parts, (int)n_parts, (int)rows_pad,
                                                                              total_entries, rowptr_out);
  if (total_entries > 0) {
    if (vals_per_entry == 1)
      assemble_entries_kernel<1><<<(unsigned)ceil_div(total_entries, 256), 256, 0, st>>>(parts, (int)n_parts, (int)cols_pad,
                                                                                         total_entries, colind_out, val_out);
    else
      assemble_entries_kernel<16><<<(unsigned)ceil_div(total_entries * 4, 256), 256, 0, st>>>(
          parts, (int)n_parts, (int)cols_pad, total_entries, colind_out, val_out);
  }
  return launch_status();
}
