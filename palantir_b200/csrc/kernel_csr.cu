// Adaptive Gaussian kernel, symmetrisation and normalisation as CSR kernels.
//
// Replaces the numpy/scipy.sparse work of src/palantir/utils.py:408-417 (self
// strip, sigma_i = distance to the floor(knn/3)-th neighbour, W = exp(-d/sigma)),
// :440 (K = W + W^T, union pattern, sorted column indices as scipy's canonical
// CSR), :442-446 (alpha normalisation) and :478-480 (T = D^-1 K).  The same
// symmetrisation with "min" instead of "+" builds the undirected shortest-path
// graph scipy's dijkstra(directed=False) walks at core.py:362.
//
// Symmetrisation without a global sort: (1) histogram of in-degrees, (2) scan of
// the per-row upper bound k + indeg, (3) every arc (i,j,w) is written twice into
// a row-bucketed scratch (out-list of i, in-list of j), (4) one warp per row
// sorts its bucket by column (rank by counting in shared memory), merges the at
// most two entries per column and compacts in place, (5) scan of the merged row
// lengths gives indptr, (6) a copy pass writes the final int32/FP64 CSR arrays.
#include "common.cuh"

namespace pb200 {

// ---------------------------------------------------------------------------
// exclusive scan of int32 -> int32/int64 (three small kernels, no temp alloc)
// ---------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__global__ void scan_block_sums(const int* __restrict__ in, long long n, long long* __restrict__ bsum) {
  __shared__ long long red[32];
  const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  long long v = i < n ? in[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) bsum[blockIdx.x] = v;
  }
}

__global__ void scan_of_sums(long long* __restrict__ bsum, int nb) {
  // single block, sequential over chunks of 1024
  __shared__ long long s[SCAN_BLOCK];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    const int i = base + threadIdx.x;
    const long long v = i < nb ? bsum[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
      const long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) bsum[i] = carry + s[threadIdx.x] - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 0) carry += s[SCAN_BLOCK - 1];
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;  // grand total
}

template <typename OutT>
__global__ void scan_finish(const int* __restrict__ in, long long n, const long long* __restrict__ bsum,
                            OutT* __restrict__ out) {
  __shared__ long long s[SCAN_BLOCK];
  const long long i = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const long long v = i < n ? in[i] : 0;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
    const long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) out[i] = (OutT)(bsum[blockIdx.x] + s[threadIdx.x] - v);
  if (i == n - 1) out[n] = (OutT)(bsum[blockIdx.x] + s[threadIdx.x]);
}

// out has n+1 entries; bsum scratch needs div_up(n,1024)+1 int64.
template <typename OutT>
static int exclusive_scan(const int* in, long long n, OutT* out, long long* bsum, cudaStream_t st) {
  const int nb = div_up(n, SCAN_BLOCK);
  scan_block_sums<<<nb, SCAN_BLOCK, 0, st>>>(in, n, bsum);
  PB_CHECK_LAUNCH("scan_block_sums");
  scan_of_sums<<<1, SCAN_BLOCK, 0, st>>>(bsum, nb);
  PB_CHECK_LAUNCH("scan_of_sums");
  scan_finish<OutT><<<nb, SCAN_BLOCK, 0, st>>>(in, n, bsum, out);
  PB_CHECK_LAUNCH("scan_finish");
  return 0;
}

// ---------------------------------------------------------------------------
// adaptive weights
// ---------------------------------------------------------------------------
__global__ void self_first_kernel(const int* __restrict__ idx, long long n, int kc, int* __restrict__ not_self) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i * kc] != (int)i) atomicOr(not_self, 1);
}

__global__ void adaptive_weights_kernel(const double* __restrict__ dist, long long n, int kc, int col0, int ka,
                                        double* __restrict__ w) {
  const int keff = kc - col0;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * keff) return;
  const long long i = e / keff;
  const int s = (int)(e - i * keff);
  const double sigma = dist[i * kc + col0 + ka - 1];
  const double q = dist[i * kc + col0 + s] / sigma;
  w[e] = exp(-q);
}

// ---------------------------------------------------------------------------
// symmetrisation
// ---------------------------------------------------------------------------
__global__ void indeg_kernel(const int* __restrict__ idx, long long n, int kc, int col0, int drop_self,
                             int* __restrict__ indeg) {
  const int keff = kc - col0;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * keff) return;
  const long long i = e / keff;
  const int j = idx[i * kc + col0 + (int)(e - i * keff)];
  if (drop_self && j == (int)i) return;
  atomicAdd(&indeg[j], 1);
}

__global__ void row_ub_kernel(const int* __restrict__ idx, long long n, int kc, int col0, int drop_self,
                              const int* __restrict__ indeg, int* __restrict__ ub) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int out = kc - col0;
  if (drop_self)
    for (int s = col0; s < kc; ++s) out -= (idx[i * kc + s] == (int)i);
  ub[i] = out + indeg[i];
}

// Each row bucket [off[i], off[i+1]) receives its out-arcs first (in neighbour
// order) then its in-arcs (atomic cursor).  `cursor` must be zeroed.
__global__ void scatter_arcs_kernel(const int* __restrict__ idx, const double* __restrict__ w, long long n,
                                    int kc, int col0, int drop_self, const long long* __restrict__ off,
                                    const int* __restrict__ indeg, const int* __restrict__ ub,
                                    int* __restrict__ cursor, int* __restrict__ tcol,
                                    double* __restrict__ tval) {
  const int keff = kc - col0;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * keff) return;
  const long long i = e / keff;
  const int s = (int)(e - i * keff);
  const int j = idx[i * kc + col0 + s];
  if (drop_self && j == (int)i) return;
  const double v = w[e];
  // out-list slot: position among the kept arcs of row i
  int slot = s;
  if (drop_self)
    for (int t = 0; t < s; ++t) slot -= (idx[i * kc + col0 + t] == (int)i);
  tcol[off[i] + slot] = j;
  tval[off[i] + slot] = v;
  const int out_j = ub[j] - indeg[j];
  const int pos = atomicAdd(&cursor[j], 1);
  tcol[off[j] + out_j + pos] = (int)i;
  tval[off[j] + out_j + pos] = v;
}

constexpr int SYM_WARPS = 4;
constexpr int SYM_CAP = 192;  // bucket entries sorted through shared memory per warp (12 KB per block: full occupancy);
                              // longer rows are sorted in place in global memory

// mode 0: sum duplicates (W + W^T); mode 1: min (undirected shortest-path graph).
__global__ void __launch_bounds__(SYM_WARPS * 32) merge_rows_kernel(long long n, const long long* __restrict__ off,
                                                                    int* __restrict__ tcol,
                                                                    double* __restrict__ tval, int mode,
                                                                    int* __restrict__ row_nnz) {
  __shared__ int s_col[SYM_WARPS][SYM_CAP];
  __shared__ int s_col2[SYM_WARPS][SYM_CAP];
  __shared__ double s_val2[SYM_WARPS][SYM_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * SYM_WARPS + warp;
  if (row >= n) return;
  const long long b = off[row];
  const int len = (int)(off[row + 1] - b);
  if (len <= SYM_CAP) {
    int* c1 = s_col[warp];
    int* c2 = s_col2[warp];
    double* v2 = s_val2[warp];
    for (int e = lane; e < len; e += 32) c1[e] = tcol[b + e];
    __syncwarp();
    for (int e = lane; e < len; e += 32) {
      const int c = c1[e];
      int r = 0;
      for (int f = 0; f < len; ++f) {
        const int o = c1[f];
        r += (o < c) || (o == c && f < e);
      }
      c2[r] = c; v2[r] = tval[b + e];
    }
    __syncwarp();
    // merge equal neighbours, warp-wide prefix count of heads
    int base = 0;
    for (int e0 = 0; e0 < len; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < len;
      const bool head = valid && (e == 0 || c2[e] != c2[e - 1]);
      const unsigned m = __ballot_sync(0xffffffffu, head);
      if (head) {
        double v = v2[e];
        if (e + 1 < len && c2[e + 1] == c2[e]) v = mode == 0 ? v + v2[e + 1] : fmin(v, v2[e + 1]);
        const int o = base + __popc(m & ((1u << lane) - 1u));
        tcol[b + o] = c2[e];
        tval[b + o] = v;
      }
      base += __popc(m);
    }
    if (lane == 0) row_nnz[row] = base;
  } else {
    // hub rows (the symmetrised kNN graph of noisy high-dimensional data has them: 1283 rows above 768
    // entries, the longest 3948, at 1 M cells): warp-parallel bitonic sort IN PLACE in global memory, in
    // the formulation whose compare-exchanges all put the smaller key at the lower index (first step of
    // every merge pairs i with i ^ (2k-1)), so the virtual +inf padding beyond len never moves.
    int np2 = 1;
    while (np2 < len) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        const int mask = (j == (k >> 1)) ? (k - 1) : j;
        for (int i = lane; i < np2; i += 32) {
          const int pj = i ^ mask;
          if (pj > i && pj < len) {
            const int ci = tcol[b + i], cj = tcol[b + pj];
            if (cj < ci) {
              const double vi = tval[b + i], vj = tval[b + pj];
              tcol[b + i] = cj; tcol[b + pj] = ci;
              tval[b + i] = vj; tval[b + pj] = vi;
            }
          }
        }
        __syncwarp();
      }
    }
    // merge equal neighbours in place: heads are written at or below their own position, one 32-entry
    // chunk after the other, every lane having read its inputs before any lane of the chunk writes
    int base = 0;
    for (int e0 = 0; e0 < len; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < len;
      const int c = valid ? tcol[b + e] : 0;
      const bool head = valid && (e == 0 || tcol[b + e - 1] != c);
      double v = valid ? tval[b + e] : 0.0;
      if (head && e + 1 < len && tcol[b + e + 1] == c) v = mode == 0 ? v + tval[b + e + 1] : fmin(v, tval[b + e + 1]);
      const unsigned m = __ballot_sync(0xffffffffu, head);
      __syncwarp();
      if (head) {
        const int o = base + __popc(m & ((1u << lane) - 1u));
        tcol[b + o] = c;
        tval[b + o] = v;
      }
      __syncwarp();
      base += __popc(m);
    }
    if (lane == 0) row_nnz[row] = base;
  }
}

__global__ void compact_rows_kernel(long long n, const long long* __restrict__ off, const int* __restrict__ indptr,
                                    const int* __restrict__ tcol, const double* __restrict__ tval,
                                    int* __restrict__ indices, double* __restrict__ data) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= n) return;
  const long long src = off[row];
  const int dst = indptr[row];
  const int len = indptr[row + 1] - dst;
  for (int e = lane; e < len; e += 32) {
    indices[dst + e] = tcol[src + e];
    data[dst + e] = tval[src + e];
  }
}

// ---------------------------------------------------------------------------
// row sums (sequential per row: scipy's csr_matvec order), scaling
// ---------------------------------------------------------------------------
__global__ void csr_row_sums_kernel(const int* __restrict__ indptr, const double* __restrict__ data, long long n,
                                    double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int e = indptr[i]; e < indptr[i + 1]; ++e) s += data[e];
  out[i] = s;
}

// op 0: v[i] = v != 0 ? 1/v : v ;  op 1: v != 0 ? v^p : v ; op 2: v > 0 ? 1/sqrt(v) : 0 ; op 3: sqrt(v) ;
// op 4: v < p ? 0 : v
__global__ void vec_transform_kernel(double* __restrict__ v, long long n, int op, double p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = v[i];
  if (op == 0) v[i] = x != 0.0 ? 1.0 / x : x;
  else if (op == 1) v[i] = x != 0.0 ? pow(x, p) : x;
  else if (op == 2) v[i] = x > 0.0 ? 1.0 / sqrt(x) : 0.0;
  else if (op == 3) v[i] = sqrt(x);
  else v[i] = x < p ? 0.0 : x;
}

// counts entries (i,j,v) that have no mirror (j,i,v') within tolerance; needs sorted columns
__global__ void csr_sym_check_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                     const double* __restrict__ data, long long n, double tol,
                                     int* __restrict__ bad) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= n) return;
  for (int e = indptr[row] + lane; e < indptr[row + 1]; e += 32) {
    const int j = indices[e];
    const double v = data[e];
    int lo = indptr[j], hi = indptr[j + 1] - 1;
    bool ok = false;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const int c = indices[mid];
      if (c == (int)row) {
        const double o = data[mid];
        ok = fabs(o - v) <= tol * fmax(fabs(o), fabs(v));
        break;
      }
      if (c < (int)row) lo = mid + 1; else hi = mid - 1;
    }
    if (!ok && v != 0.0) atomicAdd(bad, 1);
  }
}

// out[e] = (left[i] * data[e]) * right[col[e]]   (either factor may be NULL)
__global__ void csr_scale_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                 const double* __restrict__ data, long long n, const double* __restrict__ left,
                                 const double* __restrict__ right, double* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= n) return;
  const double l = left ? left[row] : 1.0;
  for (int e = indptr[row] + lane; e < indptr[row + 1]; e += 32) {
    double v = data[e];
    if (left) v = l * v;
    if (right) v = v * right[indices[e]];
    out[e] = v;
  }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// flag_out (device int) = 1 iff some row's first neighbour is not itself.
int pb200_knn_self_first(const int* idx, long long n, int kc, int* flag_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PB_CHECK(cudaMemsetAsync(flag_out, 0, sizeof(int), st));
  self_first_kernel<<<div_up(n, 256), 256, 0, st>>>(idx, n, kc, flag_out);
  PB_CHECK_LAUNCH("self_first_kernel");
  return 0;
}

// w[n][kc-col0] = exp(-(dist[i][col0+s] / dist[i][col0+ka-1]))
int pb200_adaptive_weights(const double* dist, long long n, int kc, int col0, int ka, double* w, void* stream) {
  const int keff = kc - col0;
  if (keff < 1 || ka < 1 || ka > keff) { set_error_msg("pb200_adaptive_weights: bad ka"); return -1; }
  adaptive_weights_kernel<<<div_up(n * keff, 256), 256, 0, (cudaStream_t)stream>>>(dist, n, kc, col0, ka, w);
  PB_CHECK_LAUNCH("adaptive_weights_kernel");
  return 0;
}

// Phase 1 of symmetrisation.  Scratch: indeg[n], ub[n], cursor[n] (int), off[n+1] (int64),
// bsum[div_up(n,1024)+2] (int64), tcol/tval[2*n*(kc-col0)], row_nnz[n].
// Output: indptr[n+1] (int32).  nnz = indptr[n] (read it back, allocate, call phase 2).
int pb200_symmetrize_phase1(const int* idx, const double* w, long long n, int kc, int col0, int mode,
                            int drop_self, int* indeg, int* ub, int* cursor, long long* off, long long* bsum,
                            int* tcol, double* tval, int* row_nnz, int* indptr, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int keff = kc - col0;
  if (keff < 1) { set_error_msg("pb200_symmetrize_phase1: no neighbours"); return -1; }
  if (2 * n * keff >= 2147483647LL) { set_error_msg("pb200_symmetrize_phase1: nnz overflows int32"); return -1; }
  PB_CHECK(cudaMemsetAsync(indeg, 0, sizeof(int) * n, st));
  PB_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int) * n, st));
  const int ge = div_up(n * keff, 256), gn = div_up(n, 256);
  indeg_kernel<<<ge, 256, 0, st>>>(idx, n, kc, col0, drop_self, indeg);
  PB_CHECK_LAUNCH("indeg_kernel");
  row_ub_kernel<<<gn, 256, 0, st>>>(idx, n, kc, col0, drop_self, indeg, ub);
  PB_CHECK_LAUNCH("row_ub_kernel");
  int rc = exclusive_scan<long long>(ub, n, off, bsum, st);
  if (rc) return rc;
  scatter_arcs_kernel<<<ge, 256, 0, st>>>(idx, w, n, kc, col0, drop_self, off, indeg, ub, cursor, tcol, tval);
  PB_CHECK_LAUNCH("scatter_arcs_kernel");
  merge_rows_kernel<<<div_up(n, SYM_WARPS), SYM_WARPS * 32, 0, st>>>(n, off, tcol, tval, mode, row_nnz);
  PB_CHECK_LAUNCH("merge_rows_kernel");
  rc = exclusive_scan<int>(row_nnz, n, indptr, bsum, st);
  return rc;
}

int pb200_symmetrize_phase2(long long n, const long long* off, const int* indptr, const int* tcol,
                            const double* tval, int* indices, double* data, void* stream) {
  compact_rows_kernel<<<div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(n, off, indptr, tcol, tval, indices, data);
  PB_CHECK_LAUNCH("compact_rows_kernel");
  return 0;
}

int pb200_csr_row_sums(const int* indptr, const double* data, long long n, double* out, void* stream) {
  csr_row_sums_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(indptr, data, n, out);
  PB_CHECK_LAUNCH("csr_row_sums_kernel");
  return 0;
}

int pb200_vec_transform(double* v, long long n, int op, double p, void* stream) {
  vec_transform_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(v, n, op, p);
  PB_CHECK_LAUNCH("vec_transform_kernel");
  return 0;
}

int pb200_csr_is_symmetric(const int* indptr, const int* indices, const double* data, long long n, double tol,
                           int* flag_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PB_CHECK(cudaMemsetAsync(flag_out, 0, sizeof(int), st));
  csr_sym_check_kernel<<<div_up(n, 8), 256, 0, st>>>(indptr, indices, data, n, tol, flag_out);
  PB_CHECK_LAUNCH("csr_sym_check_kernel");
  return 0;
}

int pb200_csr_scale(const int* indptr, const int* indices, const double* data, long long n, const double* left,
                    const double* right, double* out, void* stream) {
  csr_scale_kernel<<<div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, n, left, right, out);
  PB_CHECK_LAUNCH("csr_scale_kernel");
  return 0;
}

}  // extern "C"
