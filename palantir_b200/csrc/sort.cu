// Locality ordering of the cells for the pseudotime stage.
//
// The multiscale diffusion space is close to a one-dimensional curve.
// Sorting the cells along its first component makes (a) the exact low-dimensional kNN
// prunable (a reference tile whose first-coordinate gap already exceeds the current
// k-th distance cannot contribute) and (b) the shortest-path frontier touch a narrow,
// L2-resident index window.  The sort itself is CUB's device radix sort on the
// order-preserving bit pattern of the FP64 key (plumbing; the kernels that use the
// order are hand-written).
#include <cstdlib>
#include "common.cuh"
#include <cub/cub.cuh>

namespace pb200 {

__global__ void sort_keys_kernel(const double* __restrict__ data, long long n, int m, int col,
                                 unsigned long long* __restrict__ keys, int* __restrict__ iota) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long u = (unsigned long long)__double_as_longlong(data[i * m + col]);
  u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
  keys[i] = u;
  iota[i] = (int)i;
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long x) {
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

// Index along the 3-D Hilbert curve of a point with 21-bit coordinates (Skilling's axes-to-transpose
// construction: undo the excess rotations/reflections level by level, then Gray-encode), bits
// interleaved with axis 0 most significant.  Unlike the Morton curve it has no jumps: consecutive
// indices are always neighbouring cells, so 128-point tiles have compact bounding boxes.
__device__ __forceinline__ unsigned long long hilbert_key3(unsigned int x0, unsigned int x1, unsigned int x2) {
  unsigned int X[3] = {x0, x1, x2};
  const unsigned int M = 1u << 20;
  for (unsigned int Q = M; Q > 1; Q >>= 1) {
    const unsigned int P = Q - 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const unsigned int t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  unsigned int t = 0;
  for (unsigned int Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// 63-bit space-filling-curve key of three projections (structure of arrays proj[3][n]), isotropic cells
__global__ void morton_keys_kernel(const double* __restrict__ proj, long long n, double lo0, double lo1, double lo2,
                                   double inv_cell, int hilbert, unsigned long long* __restrict__ keys,
                                   int* __restrict__ iota) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lo[3] = {lo0, lo1, lo2};
  unsigned long long key = 0;
  unsigned int c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double q = (proj[(long long)a * n + i] - lo[a]) * inv_cell;
    q = q < 0.0 ? 0.0 : (q > 2097151.0 ? 2097151.0 : q);
    c[a] = (unsigned int)q;
    key |= spread3((unsigned long long)c[a]) << (2 - a);
  }
  if (hilbert) key = hilbert_key3(c[0], c[1], c[2]);
  keys[i] = key;
  iota[i] = (int)i;
}

// lo[t][a] / hi[t][a] = bounding box (3 projections) of the 128 points of work-order tile t;
// pure padding tiles get lo = +inf, hi = -inf
__global__ void tile_boxes_kernel(const double* __restrict__ proj, const int* __restrict__ perm, long long n,
                                  long long ntiles, double* __restrict__ lo, double* __restrict__ hi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= ntiles) return;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int j = lane; j < 128; j += 32) {
    const long long i = t * 128 + j;
    if (i < n) {
      const long long src = perm ? perm[i] : i;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double v = proj[(long long)a * n + src];
        mn[a] = fmin(mn[a], v);
        mx[a] = fmax(mx[a], v);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if (lane == 0) { lo[t * 3 + a] = mn[a]; hi[t * 3 + a] = mx[a]; }
  }
}

__global__ void invert_perm_kernel(const int* __restrict__ perm, long long n, int* __restrict__ inv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = (int)i;
}

// out[i][:] = a[perm[i]][:]
__global__ void permute_rows_kernel(const double* __restrict__ a, long long n, int m, const int* __restrict__ perm,
                                    double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * m) return;
  const long long i = e / m;
  out[e] = a[(long long)perm[i] * m + (e - i * m)];
}

// out[perm[i]][:] = a[i][:]
__global__ void unpermute_rows_kernel(const double* __restrict__ a, long long n, int m, const int* __restrict__ perm,
                                      double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * m) return;
  const long long i = e / m;
  out[(long long)perm[i] * m + (e - i * m)] = a[e];
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Bytes of scratch pb200_argsort_col_f64 needs for n keys.
long long pb200_argsort_scratch_bytes(long long n) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n);
  const size_t al = 256;
  size_t tot = (temp + al - 1) / al * al;
  tot += ((size_t)n * 8 + al - 1) / al * al * 2;   // keys in / out
  tot += ((size_t)n * 4 + al - 1) / al * al;       // iota
  return (long long)tot;
}

// perm[i] = index of the i-th smallest data[:, col] (stable); inv[perm[i]] = i.
int pb200_argsort_col_f64(const double* data, long long n, int m, int col, int* perm, int* inv, void* scratch,
                          long long scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n >= 2147483647LL) { set_error_msg("pb200_argsort_col_f64: n too large"); return -1; }
  if (scratch_bytes < pb200_argsort_scratch_bytes(n)) { set_error_msg("pb200_argsort_col_f64: scratch too small"); return -1; }
  const size_t al = 256;
  char* p = (char*)scratch;
  unsigned long long* kin = (unsigned long long*)p;  p += ((size_t)n * 8 + al - 1) / al * al;
  unsigned long long* kout = (unsigned long long*)p; p += ((size_t)n * 8 + al - 1) / al * al;
  int* iota = (int*)p;                               p += ((size_t)n * 4 + al - 1) / al * al;
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, kin, kout, iota, perm, (int)n);
  sort_keys_kernel<<<div_up(n, 256), 256, 0, st>>>(data, n, m, col, kin, iota);
  PB_CHECK_LAUNCH("sort_keys_kernel");
  PB_CHECK(cub::DeviceRadixSort::SortPairs((void*)p, temp, kin, kout, iota, perm, (int)n, 0, 64, st));
  invert_perm_kernel<<<div_up(n, 256), 256, 0, st>>>(perm, n, inv);
  PB_CHECK_LAUNCH("invert_perm_kernel");
  return 0;
}

// Morton (Z-curve) order of n points given three projections proj[3][n] (structure of arrays):
// cell size 1/inv_cell, origin (lo0, lo1, lo2), 21 bits per axis.  Scratch as pb200_argsort_col_f64.
int pb200_morton_order(const double* proj, long long n, double lo0, double lo1, double lo2, double inv_cell,
                       int* perm, int* inv, void* scratch, long long scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n >= 2147483647LL) { set_error_msg("pb200_morton_order: n too large"); return -1; }
  if (scratch_bytes < pb200_argsort_scratch_bytes(n)) { set_error_msg("pb200_morton_order: scratch too small"); return -1; }
  const size_t al = 256;
  char* p = (char*)scratch;
  unsigned long long* kin = (unsigned long long*)p;  p += ((size_t)n * 8 + al - 1) / al * al;
  unsigned long long* kout = (unsigned long long*)p; p += ((size_t)n * 8 + al - 1) / al * al;
  int* iota = (int*)p;                               p += ((size_t)n * 4 + al - 1) / al * al;
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, kin, kout, iota, perm, (int)n);
  static int curve = -1;   // PB200_CURVE=hilbert: Hilbert instead of Z-order (measured at C3: 13.3 % vs 13.7 % of tiles visited, 143 vs 146 ms; not the default)
  if (curve < 0) { const char* e = getenv("PB200_CURVE"); curve = (e && e[0] == 'h') ? 1 : 0; }
  morton_keys_kernel<<<div_up(n, 256), 256, 0, st>>>(proj, n, lo0, lo1, lo2, inv_cell, curve, kin, iota);
  PB_CHECK_LAUNCH("morton_keys_kernel");
  PB_CHECK(cub::DeviceRadixSort::SortPairs((void*)p, temp, kin, kout, iota, perm, (int)n, 0, 64, st));
  invert_perm_kernel<<<div_up(n, 256), 256, 0, st>>>(perm, n, inv);
  PB_CHECK_LAUNCH("invert_perm_kernel");
  return 0;
}

// box_lo / box_hi [ntiles][3]
int pb200_tile_boxes(const double* proj, const int* perm, long long n, long long ntiles, double* box_lo,
                     double* box_hi, void* stream) {
  tile_boxes_kernel<<<div_up(ntiles, 8), 256, 0, (cudaStream_t)stream>>>(proj, perm, n, ntiles, box_lo, box_hi);
  PB_CHECK_LAUNCH("tile_boxes_kernel");
  return 0;
}

int pb200_permute_rows_f64(const double* a, long long n, int m, const int* perm, double* out, void* stream) {
  if (n <= 0 || m <= 0) return 0;
  permute_rows_kernel<<<div_up(n * m, 256), 256, 0, (cudaStream_t)stream>>>(a, n, m, perm, out);
  PB_CHECK_LAUNCH("permute_rows_kernel");
  return 0;
}

int pb200_unpermute_rows_f64(const double* a, long long n, int m, const int* perm, double* out, void* stream) {
  if (n <= 0 || m <= 0) return 0;
  unpermute_rows_kernel<<<div_up(n * m, 256), 256, 0, (cudaStream_t)stream>>>(a, n, m, perm, out);
  PB_CHECK_LAUNCH("unpermute_rows_kernel");
  return 0;
}

}  // extern "C"
