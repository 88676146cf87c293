// Exact kNN in PCA space, stage 2: FP64 re-rank of the FP32 candidates, and the
// exact FP64 brute-force fallback for rows the error bound cannot certify.
//
// The returned distances restate what sklearn's EuclideanArgKmin returns for the
// call at src/palantir/utils.py:404-407: FP64 expanded form
//     d2 = max((-2 x.y + |x|^2) + |y|^2, 0)     on inputs upcast to FP64,
// dist = sqrt(d2); for float32 input the square root is taken in float32
// (sqrtf((float)d2)) and stored as a double.  Ranking is by
// (d2, index): the lowest index wins exact ties, as the reference's strict-'<'
// heap does.
#include "common.cuh"

namespace pb200 {

__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

template <typename T>
__device__ __forceinline__ double dist_out(double d2) {
  return sqrt(d2);
}
template <>
__device__ __forceinline__ double dist_out<float>(double d2) {
  return (double)sqrtf((float)d2);
}

// One warp per query row; lane handles candidates lane and lane+32 (K <= 64).
template <typename T>
__global__ void __launch_bounds__(128) knn_rerank_kernel(
    const T* __restrict__ x, const int* __restrict__ perm, const double* __restrict__ xn64,
    const double* __restrict__ xcn, const double* __restrict__ rmax2, long long n, int d, int q0, int nq,
    const int* __restrict__ cand,
    const float* __restrict__ thr32, int kk, int k, double eps_coef, int* __restrict__ out_idx,
    double* __restrict__ out_dist,
    unsigned char* __restrict__ flag, int* __restrict__ n_flagged) {
  extern __shared__ double xs[];  // [4][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + warp;
  if (q >= nq) return;
  const long long i = (long long)q0 + q;            // work-order position of the query
  const long long isrc = perm ? perm[i] : i;        // its row in x
  double* xi = xs + (size_t)warp * d;
  for (int c = lane; c < d; c += 32) xi[c] = (double)x[isrc * d + c];
  __syncwarp();
  const double xni = xn64[i];

  double d2v[2];
  int idv[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int slot = lane + 32 * h;
    int j = slot < kk ? cand[(long long)q * kk + slot] : -1;
    double d2 = INFINITY;
    if (j >= 0) {
      const int jsrc = perm ? perm[j] : j;
      const T* y = x + (long long)jsrc * d;
      double dot = 0.0;
      for (int c = 0; c < d; ++c) dot = fma(xi[c], (double)y[c], dot);
      d2 = (-2.0 * dot + xni) + xn64[j];
      d2 = d2 > 0.0 ? d2 : 0.0;
      j = jsrc;                                      // report / tie-break on the original index
    } else {
      j = 0x7fffffff;
    }
    d2v[h] = d2;
    idv[h] = j;
  }
  int r[2] = {0, 0};
  for (int s = 0; s < 32; ++s) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double od = __shfl_sync(0xffffffffu, d2v[h], s);
      const int oi = __shfl_sync(0xffffffffu, idv[h], s);
      r[0] += key_less(od, oi, d2v[0], idv[0]);
      r[1] += key_less(od, oi, d2v[1], idv[1]);
    }
  }
  // certificate: every excluded reference p has FP32 score >= thr32, hence true
  // centred score >= thr32 - eps; the kept top-k is exact if its k-th true
  // centred score (d2_k - |x_i - mu|^2) stays below that.
  const float th = thr32[q];
  const double xc = xcn[i];
  const double R2 = *rmax2;
  const double u = 5.9604644775390625e-08;  // 2^-24
  // |s32 - s| <= u (eps_coef |x||y| + 4 R^2): eps_coef = 2 (d + 8) for the FP32 FMA chain of the SIMT
  // kernel, 2 (16 + 4 n_mma) for the 3xTF32 tensor-core kernel (operand split, dropped lo.lo term and
  // up to 2 ulp of accumulation error per MMA)
  const double eps_base = 1.05 * u * (eps_coef * sqrt(xc * R2) + 4.0 * R2);
  bool bad = false;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (r[h] < k && idv[h] != 0x7fffffff) {
      out_idx[(long long)q * k + r[h]] = idv[h];
      out_dist[(long long)q * k + r[h]] = dist_out<T>(d2v[h]);
    }
    if (r[h] == k - 1 && isfinite(th)) {
      const double eps = eps_base + 8.0 * u * sqrt(R2 * d2v[h]);
      if (!(d2v[h] - xc < (double)th - eps)) bad = true;
    }
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    flag[q] = bad ? 1 : 0;
    if (bad) atomicAdd(n_flagged, 1);
  }
}

// Exact FP64 brute force for a list of query rows: one CTA per row, distances to
// every reference cached in `scratch` (n doubles per resident CTA), then k rounds
// of block-wide lexicographic arg-min.
template <typename T>
__global__ void __launch_bounds__(512) knn_exact_rows_kernel(
    const T* __restrict__ x, const int* __restrict__ perm, const double* __restrict__ xn64, long long n, int d,
    const int* __restrict__ rows, int n_rows, int q0, int k, double* __restrict__ scratch, int* __restrict__ out_idx,
    double* __restrict__ out_dist) {
  extern __shared__ double xs[];  // [d]
  __shared__ double red_d[16];
  __shared__ int red_i[16];
  __shared__ double last_d;
  __shared__ int last_i;
  double* sc = scratch + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  for (int rr = blockIdx.x; rr < n_rows; rr += gridDim.x) {
    const int q = rows[rr];              // row index relative to q0
    const long long i = (long long)q0 + q;
    const long long isrc = perm ? perm[i] : i;
    __syncthreads();
    for (int c = tid; c < d; c += 512) xs[c] = (double)x[isrc * d + c];
    __syncthreads();
    const double xni = xn64[i];
    for (long long j = tid; j < n; j += 512) {
      const T* y = x + (perm ? (long long)perm[j] : j) * d;
      double dot = 0.0;
      for (int c = 0; c < d; ++c) dot = fma(xs[c], (double)y[c], dot);
      double d2 = (-2.0 * dot + xni) + xn64[j];
      sc[j] = d2 > 0.0 ? d2 : 0.0;
    }
    if (tid == 0) { last_d = -1.0; last_i = -1; }
    __syncthreads();
    for (int round = 0; round < k; ++round) {
      const double ld = last_d;
      const int li = last_i;
      double bd = INFINITY;
      int bi = 0x7fffffff;
      for (long long j = tid; j < n; j += 512) {
        const double v = sc[j];
        const int jo = perm ? perm[j] : (int)j;      // rank and report by the original index
        if (key_less(ld, li, v, jo) && key_less(v, jo, bd, bi)) { bd = v; bi = jo; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (key_less(od, oi, bd, bi)) { bd = od; bi = oi; }
      }
      if ((tid & 31) == 0) { red_d[tid >> 5] = bd; red_i[tid >> 5] = bi; }
      __syncthreads();
      if (tid < 32) {
        bd = tid < 16 ? red_d[tid] : INFINITY;
        bi = tid < 16 ? red_i[tid] : 0x7fffffff;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (key_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        if (tid == 0) {
          last_d = bd; last_i = bi;
          if (bi != 0x7fffffff) {
            out_idx[(long long)q * k + round] = bi;
            out_dist[(long long)q * k + round] = dist_out<T>(bd);
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void compact_flags_kernel(const unsigned char* __restrict__ flag, int nq, int* __restrict__ rows,
                                     int* __restrict__ counter) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && flag[q]) rows[atomicAdd(counter, 1)] = q;
}

// out_*[perm[i]][:] = in_*[i][:]  (neighbour table rows from work order back to input order)
__global__ void knn_unpermute_kernel(const int* __restrict__ idx, const double* __restrict__ dist, long long n,
                                     int k, const int* __restrict__ perm, int* __restrict__ oidx,
                                     double* __restrict__ odist) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * k) return;
  const long long i = e / k;
  const long long o = (long long)perm[i] * k + (e - i * k);
  oidx[o] = idx[e];
  odist[o] = dist[e];
}

}  // namespace pb200

using namespace pb200;

extern "C" {

int pb200_knn_rerank_f64(const void* x, int is_f64, const int* perm, const double* xn64, const double* xcn,
                         const double* rmax2, long long n, int d, int q0, int nq, const int* cand,
                         const float* thr32, int k_keep, int k, double eps_coef, int* out_idx, double* out_dist,
                         unsigned char* flag, int* n_flagged, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (k_keep > 64 || k > k_keep || k < 1) {
    set_error_msg("pb200_knn_rerank_f64: need 1 <= k <= k_keep <= 64");
    return -1;
  }
  if (nq <= 0) return 0;
  PB_CHECK(cudaMemsetAsync(n_flagged, 0, sizeof(int), st));
  const int blocks = div_up(nq, 4);
  const size_t smem = sizeof(double) * 4 * (size_t)d;
  if (is_f64)
    knn_rerank_kernel<double><<<blocks, 128, smem, st>>>((const double*)x, perm, xn64, xcn, rmax2, n, d, q0, nq,
                                                          cand, thr32, k_keep, k, eps_coef, out_idx, out_dist,
                                                          flag, n_flagged);
  else
    knn_rerank_kernel<float><<<blocks, 128, smem, st>>>((const float*)x, perm, xn64, xcn, rmax2, n, d, q0, nq,
                                                         cand, thr32, k_keep, k, eps_coef, out_idx, out_dist,
                                                         flag, n_flagged);
  PB_CHECK_LAUNCH("knn_rerank_kernel");
  return 0;
}

// rows_out must hold nq ints; counter one int (zeroed here).
int pb200_knn_flagged_rows(const unsigned char* flag, int nq, int* rows_out, int* counter, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PB_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), st));
  if (nq <= 0) return 0;
  compact_flags_kernel<<<div_up(nq, 256), 256, 0, st>>>(flag, nq, rows_out, counter);
  PB_CHECK_LAUNCH("compact_flags_kernel");
  return 0;
}

// scratch: n_ctas * n doubles.  rows are relative to q0 (like out_idx/out_dist rows).
int pb200_knn_exact_rows_f64(const void* x, int is_f64, const int* perm, const double* xn64, long long n, int d,
                             const int* rows, int n_rows, int q0, int k, double* scratch, int n_ctas,
                             int* out_idx, double* out_dist, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows <= 0) return 0;
  if (n_ctas < 1) { set_error_msg("pb200_knn_exact_rows_f64: n_ctas < 1"); return -1; }
  const int grid = n_rows < n_ctas ? n_rows : n_ctas;
  const size_t smem = sizeof(double) * (size_t)d;
  if (is_f64)
    knn_exact_rows_kernel<double><<<grid, 512, smem, st>>>((const double*)x, perm, xn64, n, d, rows, n_rows, q0,
                                                            k, scratch, out_idx, out_dist);
  else
    knn_exact_rows_kernel<float><<<grid, 512, smem, st>>>((const float*)x, perm, xn64, n, d, rows, n_rows, q0,
                                                           k, scratch, out_idx, out_dist);
  PB_CHECK_LAUNCH("knn_exact_rows_kernel");
  return 0;
}

int pb200_knn_unpermute(const int* idx, const double* dist, long long n, int k, const int* perm, int* out_idx,
                        double* out_dist, void* stream) {
  knn_unpermute_kernel<<<div_up(n * k, 256), 256, 0, (cudaStream_t)stream>>>(idx, dist, n, k, perm, out_idx,
                                                                            out_dist);
  PB_CHECK_LAUNCH("knn_unpermute_kernel");
  return 0;
}

}  // extern "C"
