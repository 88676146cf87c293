// Min-max scaling, boundary cells and max-min waypoint sampling.
//
// Replaces sklearn.preprocessing.minmax_scale (src/palantir/core.py:176-180), the
// idxmax/idxmin boundary cells (core.py:186) and the numpy farthest-point loop of
// _max_min_sampling (core.py:291-308).  All arithmetic is FP64 with the same
// operation order as the CPU code so the sampled waypoints are identical:
//   scale = 1/(max-min) (1 when the range is < 10 eps), min_ = -min*scale,
//   x' = x*scale + min_            (separate multiply and add, MinMaxScaler)
//   min_dists = min(min_dists, |v - v[wp]|);  next wp = first arg-max.
// The m columns are independent chains and advance together: one launch per
// sampling round updates every column and elects the next waypoint of each with a
// last-block-done reduction (no host round trip).
#include "common.cuh"

namespace pb200 {

struct ValIdx {
  double v;
  long long i;
};

// larger value wins; on ties the lower index
__device__ __forceinline__ bool better_max(double v, long long i, double bv, long long bi) {
  return v > bv || (v == bv && i < bi);
}
__device__ __forceinline__ bool better_min(double v, long long i, double bv, long long bi) {
  return v < bv || (v == bv && i < bi);
}

template <bool IS_MAX>
__device__ __forceinline__ void block_best(double& v, long long& i, double* sv, long long* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (IS_MAX ? better_max(ov, oi, v, i) : better_min(ov, oi, v, i)) { v = ov; i = oi; }
  }
  if (lane == 0) { sv[warp] = v; si[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    v = lane < nw ? sv[lane] : (IS_MAX ? -INFINITY : INFINITY);
    i = lane < nw ? si[lane] : 0x7fffffffffffffffLL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (IS_MAX ? better_max(ov, oi, v, i) : better_min(ov, oi, v, i)) { v = ov; i = oi; }
    }
  }
  __syncthreads();
}

// per (block, column) partial min / max with first-occurrence indices.  grid (nblk, m)
__global__ void __launch_bounds__(256) col_minmax_partial(const double* __restrict__ a, long long n, int m,
                                                          double* __restrict__ pmin, long long* __restrict__ pimin,
                                                          double* __restrict__ pmax,
                                                          long long* __restrict__ pimax) {
  __shared__ double sv[8];
  __shared__ long long si[8];
  const int c = blockIdx.y;
  double mn = INFINITY, mx = -INFINITY;
  long long imn = 0x7fffffffffffffffLL, imx = 0x7fffffffffffffffLL;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double v = a[i * m + c];
    if (better_min(v, i, mn, imn)) { mn = v; imn = i; }
    if (better_max(v, i, mx, imx)) { mx = v; imx = i; }
  }
  block_best<false>(mn, imn, sv, si);
  if (threadIdx.x == 0) { pmin[(long long)c * gridDim.x + blockIdx.x] = mn; pimin[(long long)c * gridDim.x + blockIdx.x] = imn; }
  block_best<true>(mx, imx, sv, si);
  if (threadIdx.x == 0) { pmax[(long long)c * gridDim.x + blockIdx.x] = mx; pimax[(long long)c * gridDim.x + blockIdx.x] = imx; }
}

__global__ void __launch_bounds__(256) col_minmax_final(const double* __restrict__ pmin,
                                                        const long long* __restrict__ pimin,
                                                        const double* __restrict__ pmax,
                                                        const long long* __restrict__ pimax, int nblk,
                                                        double* __restrict__ omin, long long* __restrict__ oimin,
                                                        double* __restrict__ omax, long long* __restrict__ oimax) {
  __shared__ double sv[8];
  __shared__ long long si[8];
  const int c = blockIdx.x;
  double mn = INFINITY, mx = -INFINITY;
  long long imn = 0x7fffffffffffffffLL, imx = 0x7fffffffffffffffLL;
  for (int b = threadIdx.x; b < nblk; b += 256) {
    const double v0 = pmin[(long long)c * nblk + b];
    const long long i0 = pimin[(long long)c * nblk + b];
    if (better_min(v0, i0, mn, imn)) { mn = v0; imn = i0; }
    const double v1 = pmax[(long long)c * nblk + b];
    const long long i1 = pimax[(long long)c * nblk + b];
    if (better_max(v1, i1, mx, imx)) { mx = v1; imx = i1; }
  }
  block_best<false>(mn, imn, sv, si);
  if (threadIdx.x == 0) { omin[c] = mn; oimin[c] = imn; }
  block_best<true>(mx, imx, sv, si);
  if (threadIdx.x == 0) { omax[c] = mx; oimax[c] = imx; }
}

// MinMaxScaler arithmetic (sklearn/preprocessing/_data.py: scale_, min_, X *= scale_; X += min_)
__global__ void minmax_apply_kernel(const double* __restrict__ a, long long n, int m,
                                    const double* __restrict__ mn, const double* __restrict__ mx,
                                    double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * m) return;
  const int c = (int)(e % m);
  double range = mx[c] - mn[c];
  if (range < 10.0 * 2.220446049250313e-16) range = 1.0;
  const double scale = 1.0 / range;
  const double shift = 0.0 - __dmul_rn(mn[c], scale);
  out[e] = __dadd_rn(__dmul_rn(a[e], scale), shift);
}

// column-major copy vt[c][i] = a[i][c]
__global__ void transpose_cols_kernel(const double* __restrict__ a, long long n, int m, double* __restrict__ vt) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * m) return;
  const long long i = e / m;
  const int c = (int)(e - i * m);
  vt[(long long)c * n + i] = a[e];
}

// One sampling round for all columns.  grid (nblk, m).  cur[c] is the waypoint chosen in
// the previous round; min_dists is updated against it and the first arg-max becomes
// cur[c] (written by the last block to finish) and out[c][round].
__global__ void __launch_bounds__(256) maxmin_round_kernel(const double* __restrict__ vt, long long n, int m,
                                                           double* __restrict__ md, long long* __restrict__ cur,
                                                           double* __restrict__ pv, long long* __restrict__ pi,
                                                           unsigned int* __restrict__ done, int round, int iters,
                                                           long long* __restrict__ out) {
  __shared__ double sv[8];
  __shared__ long long si[8];
  __shared__ bool is_last;
  const int c = blockIdx.y;
  const double* v = vt + (long long)c * n;
  double* mdc = md + (long long)c * n;
  const double ref = v[cur[c]];
  double bv = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double dnew = fabs(v[i] - ref);
    const double d = round == 0 ? dnew : fmin(mdc[i], dnew);
    mdc[i] = d;
    if (better_max(d, i, bv, bi)) { bv = d; bi = i; }
  }
  block_best<true>(bv, bi, sv, si);
  if (threadIdx.x == 0) {
    pv[(long long)c * gridDim.x + blockIdx.x] = bv;
    pi[(long long)c * gridDim.x + blockIdx.x] = bi;
    __threadfence();
    const unsigned int t = atomicAdd(&done[c], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  bv = -INFINITY;
  bi = 0x7fffffffffffffffLL;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += 256) {
    const double v0 = ((volatile double*)pv)[(long long)c * gridDim.x + b];
    const long long i0 = ((volatile long long*)pi)[(long long)c * gridDim.x + b];
    if (better_max(v0, i0, bv, bi)) { bv = v0; bi = i0; }
  }
  block_best<true>(bv, bi, sv, si);
  if (threadIdx.x == 0) {
    cur[c] = bi;
    out[(long long)c * iters + round + 1] = bi;
    done[c] = 0;
  }
}

__global__ void gather_rows_f64_kernel(const double* __restrict__ a, int m, const long long* __restrict__ rows,
                                       long long nr, double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nr * m) return;
  const long long r = e / m;
  out[e] = a[rows[r] * m + (e - r * m)];
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Column minima / maxima of a[n][m] with first-occurrence positions.
// scratch: pval[2*m*nblk] doubles, pidx[2*m*nblk] int64 with nblk = min(1024, ceil(n/256)).
int pb200_col_minmax(const double* a, long long n, int m, double* omin, long long* oimin, double* omax,
                     long long* oimax, double* pval, long long* pidx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = (int)(div_up(n, 256) < 1024 ? div_up(n, 256) : 1024);
  double* pmin = pval;
  double* pmax = pval + (long long)m * nblk;
  long long* pimin = pidx;
  long long* pimax = pidx + (long long)m * nblk;
  col_minmax_partial<<<dim3(nblk, m), 256, 0, st>>>(a, n, m, pmin, pimin, pmax, pimax);
  PB_CHECK_LAUNCH("col_minmax_partial");
  col_minmax_final<<<m, 256, 0, st>>>(pmin, pimin, pmax, pimax, nblk, omin, oimin, omax, oimax);
  PB_CHECK_LAUNCH("col_minmax_final");
  return 0;
}

int pb200_minmax_apply(const double* a, long long n, int m, const double* mn, const double* mx, double* out,
                       void* stream) {
  minmax_apply_kernel<<<div_up(n * m, 256), 256, 0, (cudaStream_t)stream>>>(a, n, m, mn, mx, out);
  PB_CHECK_LAUNCH("minmax_apply_kernel");
  return 0;
}

// Farthest-point sampling in each column.  first[m]: initial waypoint per column (the
// reference draws them with np.random.randint).  out[m][iters] (out[c][0] = first[c]).
// scratch: vt[m*n], md[m*n] doubles; cur[m] int64; pv[m*nblk] doubles, pi[m*nblk] int64
// (nblk = min(592, ceil(n/256))); done[m] uint32.
int pb200_maxmin_sampling(const double* data, long long n, int m, const long long* first, int iters,
                          long long* out, double* vt, double* md, long long* cur, double* pv, long long* pi,
                          unsigned int* done, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (iters < 1) return 0;
  const int nblk = (int)(div_up(n, 256) < 4 * kSMs ? div_up(n, 256) : 4 * kSMs);
  transpose_cols_kernel<<<div_up(n * m, 256), 256, 0, st>>>(data, n, m, vt);
  PB_CHECK_LAUNCH("transpose_cols_kernel");
  PB_CHECK(cudaMemcpyAsync(cur, first, sizeof(long long) * m, cudaMemcpyDeviceToDevice, st));
  PB_CHECK(cudaMemcpy2DAsync(out, sizeof(long long) * iters, first, sizeof(long long), sizeof(long long), m,
                             cudaMemcpyDeviceToDevice, st));
  PB_CHECK(cudaMemsetAsync(done, 0, sizeof(unsigned int) * m, st));
  for (int r = 0; r + 1 < iters; ++r) {
    maxmin_round_kernel<<<dim3(nblk, m), 256, 0, st>>>(vt, n, m, md, cur, pv, pi, done, r, iters, out);
    PB_CHECK_LAUNCH("maxmin_round_kernel");
  }
  return 0;
}

int pb200_gather_rows_f64(const double* a, int m, const long long* rows, long long nr, double* out, void* stream) {
  if (nr <= 0) return 0;
  gather_rows_f64_kernel<<<div_up(nr * m, 256), 256, 0, (cudaStream_t)stream>>>(a, m, rows, nr, out);
  PB_CHECK_LAUNCH("gather_rows_f64_kernel");
  return 0;
}

}  // extern "C"
