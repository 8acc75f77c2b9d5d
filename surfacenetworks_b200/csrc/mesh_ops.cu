// mesh_ops.cu -- GPU construction of the three mesh operators, straight into the formats the SpMM kernels read
// (SURVEY.md 8(f) row f3): the step in front of the hot path.
//
//   Dirac      D  [4F x 4V]  block(f, j) = -Q(0, V[j+1] - V[j+2]) / (2 A_f)            reference src/utils/mesh.py:35-58
//   adjoint    D* [4V x 4F]  block*(j, f) = block(f, j)^T A_f / A_v[j], A_v = sum A_f/3  reference src/utils/mesh.py:44-45,59
//   Laplacian  L  [V x V]    A^-1 (diag(colsum W) - W), cotangent weights W            reference src/utils/mesh.py:17-26,
//                            67-80,102-112, src/utils/graph.py:40-49, as_rigid_as_possible/add_laplacian.py:50-56
//
// The reference builds them offline with dense O(V^2) / O(F V) numpy temporaries (1 GB at 2k vertices) and stores scipy
// pickles.  Here a batch of meshes (padded to v_pad vertices / f_pad faces, padding faces marked by a negative index)
// becomes the block-diagonal BSR4 / CSR32 batch operators in a handful of O(F) kernels.  All geometry is evaluated in
// fp64 with the reference's operation order (explicit round-to-nearest intrinsics, no FMA contraction) and rounded to
// fp32 once at the end, exactly like `.astype('float32')` in the reference recipe -- tests compare against operators
// built by the reference's own code.  Deterministic: no floating-point atomics; sums run in ascending face order.
//
// Bound: latency (a few hundred KB per mesh); runs once per batch of new geometry, not per layer.
#include "common.cuh"

namespace sn {
namespace mesh {

constexpr int kMaxInc = 64;      // incident faces per vertex handled in registers / local memory (Delaunay meshes stay
                                 // below 20); vertices above it (poles of UV spheres, cone apexes) take the same code
                                 // over global scratch -- any valence is supported, like the reference

__device__ __forceinline__ double sqdist3(const double* a, const double* b) {
  const double d0 = __dsub_rn(a[0], b[0]), d1 = __dsub_rn(a[1], b[1]), d2 = __dsub_rn(a[2], b[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));   // mesh.py:24
}

struct FaceGeom {
  double l01, l12, l20, area;
};
// edge lengths and Heron area with the reference's 1e-6 floor (mesh.py:17-26, 67-80)
__device__ __forceinline__ FaceGeom face_geom(const double* P, int i0, int i1, int i2) {
  FaceGeom g;
  g.l01 = __dsqrt_rn(sqdist3(P + 3 * i0, P + 3 * i1));
  g.l12 = __dsqrt_rn(sqdist3(P + 3 * i1, P + 3 * i2));
  g.l20 = __dsqrt_rn(sqdist3(P + 3 * i2, P + 3 * i0));
  const double s = __ddiv_rn(__dadd_rn(__dadd_rn(g.l01, g.l12), g.l20), 2.0);
  const double prod = __dmul_rn(__dmul_rn(__dmul_rn(s, __dsub_rn(s, g.l01)), __dsub_rn(s, g.l12)), __dsub_rn(s, g.l20));
  g.area = prod > 0.0 ? __dsqrt_rn(prod) : 1e-6;
  return g;
}

__device__ __forceinline__ bool face_valid(const int32_t* f, int v_pad) {
  return f[0] >= 0 && f[1] >= 0 && f[2] >= 0 && f[0] < v_pad && f[1] < v_pad && f[2] < v_pad;
}

// pass 1 over faces: area, incidence counts per vertex, blocks per face row (3 or 0)
__global__ void __launch_bounds__(256)
face_pass_kernel(const double* __restrict__ V, const int32_t* __restrict__ F, int n_meshes, int v_pad, int f_pad,
                 double* __restrict__ area, int* __restrict__ vcount, int* __restrict__ fblocks) {
  const int64_t bf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bf >= (int64_t)n_meshes * f_pad) return;
  const int b = (int)(bf / f_pad);
  const int32_t* f = F + 3 * bf;
  if (!face_valid(f, v_pad)) {
    area[bf] = -1.0;
    fblocks[bf] = 0;
    return;
  }
  const double* P = V + (int64_t)b * v_pad * 3;
  area[bf] = face_geom(P, f[0], f[1], f[2]).area;
  fblocks[bf] = 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) atomicAdd(vcount + (int64_t)b * v_pad + f[c], 1);
}

// pass 2 over faces: scatter (face, corner) keys into the per-vertex incidence lists (order fixed later by a sort)
__global__ void __launch_bounds__(256)
incidence_fill_kernel(const int32_t* __restrict__ F, const double* __restrict__ area, int n_meshes, int v_pad, int f_pad,
                      const int* __restrict__ vptr, int* __restrict__ cursor, int* __restrict__ inc) {
  const int64_t bf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bf >= (int64_t)n_meshes * f_pad || area[bf] < 0.0) return;
  const int b = (int)(bf / f_pad);
  const int32_t* f = F + 3 * bf;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int64_t v = (int64_t)b * v_pad + f[c];
    inc[vptr[v] + atomicAdd(cursor + v, 1)] = (int)(bf * 4 + c);
  }
}

// sorts each vertex's incidence list ascending (= face order, then corner: the accumulation order of np.add.at in the
// reference) and computes the vertex area A_v = sum A_f / 3 (mesh.py:44-45)
__global__ void __launch_bounds__(128)
vertex_sort_kernel(const int* __restrict__ vptr, int* __restrict__ inc, const double* __restrict__ area, int64_t n_vert,
                   double* __restrict__ varea, int* __restrict__ status) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int k0 = vptr[v], n = vptr[v + 1] - k0;
  if (n > kMaxInc) {                     // rare: sort in place in global memory (status reports the largest valence seen)
    atomicMax(status, n);
    for (int i = 1; i < n; ++i) {
      const int x = inc[k0 + i];
      int j = i - 1;
      while (j >= 0 && inc[k0 + j] > x) {
        inc[k0 + j + 1] = inc[k0 + j];
        --j;
      }
      inc[k0 + j + 1] = x;
    }
    double a = 0.0;
    for (int i = 0; i < n; ++i) a = __dadd_rn(a, __ddiv_rn(area[inc[k0 + i] >> 2], 3.0));
    varea[v] = a;
    return;
  }
  int key[kMaxInc];
  for (int i = 0; i < n; ++i) key[i] = inc[k0 + i];
  for (int i = 1; i < n; ++i) {          // insertion sort: lists are short
    const int x = key[i];
    int j = i - 1;
    while (j >= 0 && key[j] > x) {
      key[j + 1] = key[j];
      --j;
    }
    key[j + 1] = x;
  }
  double a = 0.0;
  for (int i = 0; i < n; ++i) {
    inc[k0 + i] = key[i];
    a = __dadd_rn(a, __ddiv_rn(area[key[i] >> 2], 3.0));
  }
  varea[v] = a;
}

// real 4x4 matrix of the pure quaternion (0, e), mesh.py:28-33
__device__ __forceinline__ void quaternion_block(const double* e, double (&Q)[4][4]) {
  const double b = e[0], c = e[1], d = e[2];
  Q[0][0] = 0.0; Q[0][1] = -b;  Q[0][2] = -c;  Q[0][3] = -d;
  Q[1][0] = b;   Q[1][1] = 0.0; Q[1][2] = -d;  Q[1][3] = c;
  Q[2][0] = c;   Q[2][1] = d;   Q[2][2] = 0.0; Q[2][3] = -b;
  Q[3][0] = d;   Q[3][1] = -c;  Q[3][2] = b;   Q[3][3] = 0.0;
}
// mat = -Q(V[c+1] - V[c+2]) / (2 A_f) for corner c of face f (mesh.py:47-58)
__device__ __forceinline__ void dirac_block(const double* P, const int32_t* f, int c, double af, double (&M)[4][4]) {
  const double* pa = P + 3 * f[(c + 1) % 3];
  const double* pb = P + 3 * f[(c + 2) % 3];
  const double e[3] = {__dsub_rn(pa[0], pb[0]), __dsub_rn(pa[1], pb[1]), __dsub_rn(pa[2], pb[2])};
  double Q[4][4];
  quaternion_block(e, Q);
  const double den = __dmul_rn(2.0, af);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) M[p][q] = __ddiv_rn(-Q[p][q], den);
}
// rotated column-major block storage of the SpMM kernels: out[4 q + s] = B[(q + s) % 4][q]
__device__ __forceinline__ void store_block(float* out, const double (&B)[4][4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(out + 4 * q) = make_float4((float)B[q][q], (float)B[(q + 1) & 3][q],
                                                          (float)B[(q + 2) & 3][q], (float)B[(q + 3) & 3][q]);
}

// D: one thread per face row; its three blocks in ascending vertex order
__global__ void __launch_bounds__(128)
dirac_rows_kernel(const double* __restrict__ V, const int32_t* __restrict__ F, const double* __restrict__ area,
                  const double* __restrict__ varea, int n_meshes, int v_pad, int f_pad, const int* __restrict__ browptr,
                  int32_t* __restrict__ bcolind, float* __restrict__ bval, int32_t* __restrict__ at_bcolind,
                  float* __restrict__ at_bval) {
  const int64_t bf = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bf >= (int64_t)n_meshes * f_pad || area[bf] < 0.0) return;
  const int b = (int)(bf / f_pad);
  const int32_t* f = F + 3 * bf;
  const double* P = V + (int64_t)b * v_pad * 3;
  int order[3] = {0, 1, 2};
#pragma unroll
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && f[order[j - 1]] > f[order[j]]; --j) {
      const int t = order[j];
      order[j] = order[j - 1];
      order[j - 1] = t;
    }
  const int k0 = browptr[bf];
  for (int i = 0; i < 3; ++i) {
    const int c = order[i];
    double M[4][4];
    dirac_block(P, f, c, area[bf], M);
    bcolind[k0 + i] = b * v_pad + f[c];
    store_block(bval + (int64_t)(k0 + i) * 16, M);
    if (at_bval != nullptr) {        // (D*)^T has D's structure: block (f, j) = block*(j, f)^T = M A_f / A_v[j]
      const double av = varea[(int64_t)b * v_pad + f[c]];
      double T[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) T[p][q] = __ddiv_rn(__dmul_rn(M[p][q], area[bf]), av);
      at_bcolind[k0 + i] = b * v_pad + f[c];
      store_block(at_bval + (int64_t)(k0 + i) * 16, T);
    }
  }
}

// D*: one thread per vertex row; blocks in ascending face order
__global__ void __launch_bounds__(128)
adjoint_rows_kernel(const double* __restrict__ V, const int32_t* __restrict__ F, const double* __restrict__ area,
                    const double* __restrict__ varea, int v_pad, int f_pad, int64_t n_vert, const int* __restrict__ vptr,
                    const int* __restrict__ inc, int32_t* __restrict__ bcolind, float* __restrict__ bval,
                    int32_t* __restrict__ dt_bcolind, float* __restrict__ dt_bval) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int b = (int)(v / v_pad);
  const double* P = V + (int64_t)b * v_pad * 3;
  const int k0 = vptr[v], n = vptr[v + 1] - k0;
  const double av = varea[v];
  for (int i = 0; i < n; ++i) {
    const int key = inc[k0 + i];
    const int64_t bf = key >> 2;
    const int c = key & 3;
    const double af = area[bf];
    double M[4][4], T[4][4];
    dirac_block(P, F + 3 * bf, c, af, M);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) T[p][q] = __ddiv_rn(__dmul_rn(M[q][p], af), av);   // mat^T * A_f / A_v, mesh.py:59
    bcolind[k0 + i] = (int)bf;
    store_block(bval + (int64_t)(k0 + i) * 16, T);
    if (dt_bval != nullptr) {        // D^T has D*'s structure: block (j, f) = block(f, j)^T
      double Mt[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) Mt[p][q] = M[q][p];
      dt_bcolind[k0 + i] = (int)bf;
      store_block(dt_bval + (int64_t)(k0 + i) * 16, Mt);
    }
  }
}

// One Laplacian row (add_laplacian.py:50-56): gathers the cotangent contributions of the incident faces, merges them per
// neighbour in face order, drops exact zeros like csr_matrix(dense) does (mesh.py:112) and hands (column, value)
// pairs in ascending column order to `emit`.  Returns the number of entries.
template <typename Emit>
__device__ __forceinline__ int laplacian_row(const double* __restrict__ P, const int32_t* __restrict__ F,
                                             const double* __restrict__ area, const int* __restrict__ inc, int k0, int n,
                                             int vi /*local vertex index*/, int* nb, double* wij, double* wji, Emit emit) {
  // nb: neighbour (local vertex index); wij / wji: contributions to W[i, j] / W[j, i] (column i, feeds the degree
  // d_i = sum_j W[j, i]); 2 n entries each -- thread-local arrays up to kMaxInc faces, global scratch above
  double A = 0.0;
  int m = 0;
  for (int t = 0; t < n; ++t) {
    const int key = inc[k0 + t];
    const int64_t bf = key >> 2;
    const int c = key & 3;
    const int32_t* f = F + 3 * bf;
    const FaceGeom g = face_geom(P, f[0], f[1], f[2]);
    double sq[3][3];
    sq[0][1] = sq[1][0] = __dmul_rn(g.l01, g.l01);
    sq[1][2] = sq[2][1] = __dmul_rn(g.l12, g.l12);
    sq[2][0] = sq[0][2] = __dmul_rn(g.l20, g.l20);
    const double den = __dadd_rn(__dmul_rn(8.0, g.area), 1e-6);
    const double x = __ddiv_rn(__ddiv_rn(g.area, 3.0), 4.0);     // mesh.py:110, once per permutation: twice per corner
    A = __dadd_rn(__dadd_rn(A, x), x);
    // the two permutations (i, j, k) that start at this corner, in itertools.permutations order (ascending j)
    const int o1 = c == 0 ? 1 : 0, o2 = c == 2 ? 1 : 2;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int pj = u == 0 ? o1 : o2, pk = u == 0 ? o2 : o1;
      // mesh.py:109  W[i,j] += (-l_ij^2 + l_jk^2 + l_ki^2) / (8 a + 1e-6)
      const double cij = __ddiv_rn(__dadd_rn(__dadd_rn(-sq[c][pj], sq[pj][pk]), sq[pk][c]), den);
      const double cji = __ddiv_rn(__dadd_rn(__dadd_rn(-sq[pj][c], sq[c][pk]), sq[pk][pj]), den);
      // stable insertion by neighbour index: equal neighbours keep face order
      const int j = f[pj];
      int pos = m;
      while (pos > 0 && nb[pos - 1] > j) {
        nb[pos] = nb[pos - 1];
        wij[pos] = wij[pos - 1];
        wji[pos] = wji[pos - 1];
        --pos;
      }
      nb[pos] = j;
      wij[pos] = cij;
      wji[pos] = cji;
      ++m;
    }
  }
  const double ainv = __ddiv_rn(1.0, __dadd_rn(A, 1e-9));      // add_laplacian.py:53
  // degree: column sum of W in ascending row order (graph.py:44), exact zeros of W dropped first (no effect on the sum)
  double d = 0.0;
  for (int s = 0; s < m;) {
    double w = wji[s];
    int e = s + 1;
    while (e < m && nb[e] == nb[s]) w = __dadd_rn(w, wji[e++]);
    d = __dadd_rn(d, w);
    s = e;
  }
  int count = 0;
  bool diag_done = false;
  auto emit_diag = [&]() {
    if (d != 0.0) {
      emit(count, vi, (float)__dmul_rn(ainv, d));
      ++count;
    }
    diag_done = true;
  };
  for (int s = 0; s < m;) {
    double w = wij[s];
    int e = s + 1;
    while (e < m && nb[e] == nb[s]) w = __dadd_rn(w, wij[e++]);
    if (!diag_done && nb[s] > vi) emit_diag();
    if (w != 0.0) {
      emit(count, nb[s], (float)__dmul_rn(ainv, -w));
      ++count;
    }
    s = e;
  }
  if (!diag_done && n > 0) emit_diag();
  return count;
}

__global__ void __launch_bounds__(128)
laplacian_count_kernel(const double* __restrict__ V, const int32_t* __restrict__ F, const double* __restrict__ area,
                       int v_pad, int f_pad, int64_t n_vert, const int* __restrict__ vptr, const int* __restrict__ inc,
                       int* __restrict__ rowcount, int* s_nb, double* s_wij, double* s_wji) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int b = (int)(v / v_pad);
  const int k0 = vptr[v], n = vptr[v + 1] - k0;
  int cnt = 0;
  if (n > 0 && n <= kMaxInc) {
    int nb[2 * kMaxInc];
    double wij[2 * kMaxInc], wji[2 * kMaxInc];
    cnt = laplacian_row(V + (int64_t)b * v_pad * 3, F, area, inc, k0, n, (int)(v - (int64_t)b * v_pad), nb, wij, wji,
                        [](int, int, float) {});
  } else if (n > 0) {
    cnt = laplacian_row(V + (int64_t)b * v_pad * 3, F, area, inc, k0, n, (int)(v - (int64_t)b * v_pad), s_nb + 2 * (int64_t)k0,
                        s_wij + 2 * (int64_t)k0, s_wji + 2 * (int64_t)k0, [](int, int, float) {});
  }
  rowcount[v] = cnt;
}

__global__ void __launch_bounds__(128)
laplacian_fill_kernel(const double* __restrict__ V, const int32_t* __restrict__ F, const double* __restrict__ area,
                      int v_pad, int f_pad, int64_t n_vert, const int* __restrict__ vptr, const int* __restrict__ inc,
                      const int* __restrict__ rowptr, int32_t* __restrict__ colind, float* __restrict__ val, int* s_nb,
                      double* s_wij, double* s_wji) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vert) return;
  const int b = (int)(v / v_pad);
  const int k0 = vptr[v], n = vptr[v + 1] - k0;
  if (n <= 0) return;
  const int r0 = rowptr[v];
  const int cbase = b * v_pad;
  auto emit = [&](int i, int j, float x) {
    colind[r0 + i] = cbase + j;
    val[r0 + i] = x;
  };
  if (n <= kMaxInc) {
    int nb[2 * kMaxInc];
    double wij[2 * kMaxInc], wji[2 * kMaxInc];
    laplacian_row(V + (int64_t)b * v_pad * 3, F, area, inc, k0, n, (int)(v - (int64_t)b * v_pad), nb, wij, wji, emit);
  } else {
    laplacian_row(V + (int64_t)b * v_pad * 3, F, area, inc, k0, n, (int)(v - (int64_t)b * v_pad), s_nb + 2 * (int64_t)k0,
                  s_wij + 2 * (int64_t)k0, s_wji + 2 * (int64_t)k0, emit);
  }
}

struct Workspace {
  double* area;     // [n f_pad]   (-1: padding face)
  double* varea;    // [n v_pad]
  int* vcount;      // [n v_pad + 1]  -> vptr after the scan
  int* cursor;      // [n v_pad + 1]  (reused as the Laplacian row counts)
  int* fblocks;     // [n f_pad + 1]
  int* inc;         // [3 n f_pad]
  int* tiles;       // scan scratch
  int* s_nb;        // [6 n f_pad]  Laplacian rows of vertices with more than kMaxInc incident faces (2 entries per incidence)
  double* s_wij;    // [6 n f_pad]
  double* s_wji;    // [6 n f_pad]
  size_t bytes;
};
inline size_t align_up(size_t v) { return (v + 255) / 256 * 256; }
inline Workspace carve(void* ws, int64_t n, int64_t v_pad, int64_t f_pad) {
  Workspace w;
  char* p = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes);
    return r;
  };
  const int64_t nv = n * v_pad, nf = n * f_pad;
  w.area = reinterpret_cast<double*>(take(sizeof(double) * nf));
  w.varea = reinterpret_cast<double*>(take(sizeof(double) * nv));
  w.vcount = reinterpret_cast<int*>(take(sizeof(int) * (nv + 1)));
  w.cursor = reinterpret_cast<int*>(take(sizeof(int) * (nv + 1)));
  w.fblocks = reinterpret_cast<int*>(take(sizeof(int) * (nf + 1)));
  w.inc = reinterpret_cast<int*>(take(sizeof(int) * 3 * nf));
  w.tiles = reinterpret_cast<int*>(take(sizeof(int) * (exclusive_scan_tiles(nv > nf ? nv : nf) + 1)));
  w.s_nb = reinterpret_cast<int*>(take(sizeof(int) * 6 * nf));
  w.s_wij = reinterpret_cast<double*>(take(sizeof(double) * 6 * nf));
  w.s_wji = reinterpret_cast<double*>(take(sizeof(double) * 6 * nf));
  w.bytes = off;
  return w;
}

inline int check_args(const double* V, const int32_t* F, int64_t n, int64_t v_pad, int64_t f_pad, int32_t* status,
                      void* ws, size_t ws_bytes) {
  if (n < 0 || v_pad <= 0 || f_pad <= 0) return SN_ERR_ARG;
  if (n * f_pad * 4 >= 0x7fffffffLL || n * v_pad >= 0x7fffffffLL) return SN_ERR_OVERFLOW;
  if (n == 0) return SN_OK;
  if (!V || !F || !status || !ws) return SN_ERR_ARG;
  if (ws_bytes < carve(nullptr, n, v_pad, f_pad).bytes) return SN_ERR_WORKSPACE;
  return SN_OK;
}

// shared front end: areas, sorted incidence lists (vptr / inc), vertex areas; D's row pointers when asked
inline int front_end(const double* V, const int32_t* F, int64_t n, int64_t v_pad, int64_t f_pad, int32_t* status,
                     const Workspace& w, int32_t* d_browptr, cudaStream_t st) {
  const int64_t nv = n * v_pad, nf = n * f_pad;
  cudaMemsetAsync(w.vcount, 0, sizeof(int) * (nv + 1), st);
  cudaMemsetAsync(w.cursor, 0, sizeof(int) * (nv + 1), st);
  cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  face_pass_kernel<<<(unsigned)ceil_div(nf, 256), 256, 0, st>>>(V, F, (int)n, (int)v_pad, (int)f_pad, w.area, w.vcount,
                                                                 w.fblocks);
  int rc = exclusive_scan(w.vcount, nv, w.vcount, w.tiles, st);
  if (rc != SN_OK) return rc;
  if (d_browptr) {
    rc = exclusive_scan(w.fblocks, nf, d_browptr, w.tiles, st);
    if (rc != SN_OK) return rc;
  }
  incidence_fill_kernel<<<(unsigned)ceil_div(nf, 256), 256, 0, st>>>(F, w.area, (int)n, (int)v_pad, (int)f_pad, w.vcount,
                                                                      w.cursor, w.inc);
  vertex_sort_kernel<<<(unsigned)ceil_div(nv, 128), 128, 0, st>>>(w.vcount, w.inc, w.area, nv, w.varea, status);
  return launch_status();
}

}  // namespace mesh
}  // namespace sn

SN_API size_t sn_mesh_ws_bytes(int64_t n_meshes, int64_t v_pad, int64_t f_pad) {
  if (n_meshes <= 0 || v_pad <= 0 || f_pad <= 0) return 0;
  return sn::mesh::carve(nullptr, n_meshes, v_pad, f_pad).bytes;
}

SN_API int sn_mesh_dirac_bsr4(const double* V, const int32_t* F, int64_t n_meshes, int64_t v_pad, int64_t f_pad,
                              int32_t* d_browptr, int32_t* d_bcolind, float* d_bval, int32_t* da_browptr,
                              int32_t* da_bcolind, float* da_bval, int32_t* dt_bcolind, float* dt_bval,
                              int32_t* dat_bcolind, float* dat_bval, int32_t* status, void* ws, size_t ws_bytes,
                              sn_stream_t stream) {
  using namespace sn;
  using namespace sn::mesh;
  int rc = check_args(V, F, n_meshes, v_pad, f_pad, status, ws, ws_bytes);
  if (rc != SN_OK || n_meshes == 0) return rc;
  if (!d_browptr || !d_bcolind || !d_bval || !da_browptr || !da_bcolind || !da_bval) return SN_ERR_ARG;
  if (!aligned16(d_bval) || !aligned16(da_bval) || !aligned16(dt_bval) || !aligned16(dat_bval)) return SN_ERR_UNSUPPORTED;
  if ((dt_bval == nullptr) != (dt_bcolind == nullptr) || (dat_bval == nullptr) != (dat_bcolind == nullptr)) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const Workspace w = carve(ws, n_meshes, v_pad, f_pad);
  const int64_t nv = n_meshes * v_pad, nf = n_meshes * f_pad;
  rc = front_end(V, F, n_meshes, v_pad, f_pad, status, w, d_browptr, st);
  if (rc != SN_OK) return rc;
  dirac_rows_kernel<<<(unsigned)ceil_div(nf, 128), 128, 0, st>>>(V, F, w.area, w.varea, (int)n_meshes, (int)v_pad,
                                                                  (int)f_pad, d_browptr, d_bcolind, d_bval, dat_bcolind,
                                                                  dat_bval);
  cudaMemcpyAsync(da_browptr, w.vcount, sizeof(int) * (nv + 1), cudaMemcpyDeviceToDevice, st);
  adjoint_rows_kernel<<<(unsigned)ceil_div(nv, 128), 128, 0, st>>>(V, F, w.area, w.varea, (int)v_pad, (int)f_pad, nv,
                                                                    w.vcount, w.inc, da_bcolind, da_bval, dt_bcolind,
                                                                    dt_bval);
  return launch_status();
}

SN_API int sn_mesh_laplacian_csr(const double* V, const int32_t* F, int64_t n_meshes, int64_t v_pad, int64_t f_pad,
                                 int32_t* rowptr, int32_t* colind, float* val, int32_t* status, void* ws,
                                 size_t ws_bytes, sn_stream_t stream) {
  using namespace sn;
  using namespace sn::mesh;
  int rc = check_args(V, F, n_meshes, v_pad, f_pad, status, ws, ws_bytes);
  if (rc != SN_OK || n_meshes == 0) return rc;
  if (!rowptr || !colind || !val) return SN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const Workspace w = carve(ws, n_meshes, v_pad, f_pad);
  const int64_t nv = n_meshes * v_pad;
  rc = front_end(V, F, n_meshes, v_pad, f_pad, status, w, nullptr, st);
  if (rc != SN_OK) return rc;
  laplacian_count_kernel<<<(unsigned)ceil_div(nv, 128), 128, 0, st>>>(V, F, w.area, (int)v_pad, (int)f_pad, nv, w.vcount,
                                                                       w.inc, w.cursor, w.s_nb, w.s_wij, w.s_wji);
  rc = exclusive_scan(w.cursor, nv, rowptr, w.tiles, st);
  if (rc != SN_OK) return rc;
  laplacian_fill_kernel<<<(unsigned)ceil_div(nv, 128), 128, 0, st>>>(V, F, w.area, (int)v_pad, (int)f_pad, nv, w.vcount,
                                                                      w.inc, rowptr, colind, val, w.s_nb, w.s_wij, w.s_wji);
  return launch_status();
}
