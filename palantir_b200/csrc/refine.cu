// Pseudotime refinement and fate projection on the nW x N shortest-path matrix.
//
// Replaces the dense numpy stage of _compute_pseudotime (src/palantir/core.py:368-397)
// and the projection + per-row entropy of run_palantir (core.py:225-230).  The
// reference materialises W, signs and P (several nW x N FP64 temporaries, 5 GB RSS at
// 100k cells); here D is the only nW x N array and every pass streams it once:
//   sdv    = std(D) * 1.06 * (nW N)^(-1/5)                (two reduction passes)
//   w_sj   = exp(-0.5 (D_sj / sdv)^2),  colsum_j = sum_s w_sj
//   new_j  = sum_s (D_sj * sign_sj + t_s) * (w_sj / colsum_j)
//   fate_j = sum_s (w_sj / colsum_j) * B_s,   entropy_j = H(fate_j / sum fate_j)
// One thread owns a cell (column j) and walks the waypoint rows in order, so on one
// GPU the accumulation order equals numpy's axis-0 reduction.  With the rows of D
// sharded across GPUs the same kernels produce partial sums that the host side
// all-reduces.
#include "common.cuh"

namespace pb200 {

constexpr int RB = 256;
constexpr long long RCHUNK = 4096;

// partial[b] = sum over chunk b of (x - shift)^power  (power 1 or 2)
__global__ void __launch_bounds__(RB) sum_pow_kernel(const double* __restrict__ x, long long count, double shift,
                                                     int power, double* __restrict__ partial) {
  __shared__ double red[RB / 32];
  double s = 0.0;
  const long long base = (long long)blockIdx.x * RCHUNK;
#pragma unroll 4
  for (int t = 0; t < RCHUNK / RB; ++t) {
    const long long i = base + threadIdx.x + (long long)t * RB;
    if (i < count) {
      const double v = x[i] - shift;
      s += power == 1 ? v : v * v;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int q = 0; q < RB / 32; ++q) tot += red[q];
    partial[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(1024) sum_final_kernel(const double* __restrict__ partial, long long nb,
                                                         double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long b = threadIdx.x; b < nb; b += 1024) s += partial[b];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = red[threadIdx.x];
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s;
  }
}

__device__ __forceinline__ double wp_weight(double d, double sdv) {
  const double r = d / sdv;
  return exp(-0.5 * (r * r));
}

__global__ void __launch_bounds__(RB) colsum_kernel(const double* __restrict__ D, int S, long long n, double sdv,
                                                    double* __restrict__ colsum) {
  const long long j = (long long)blockIdx.x * RB + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int r = 0; r < S; ++r) s += wp_weight(D[(long long)r * n + j], sdv);
  colsum[j] = s;
}

// out[j] = sum_s (D_sj * sign + t_add[s]) * (w_sj / colsum[j]);
// sign = -1 where pt[j] < t_mask[s], except on local row `plus_row` (the start waypoint).
__global__ void __launch_bounds__(RB) refine_step_kernel(const double* __restrict__ D, int S, long long n,
                                                         double sdv, const double* __restrict__ colsum,
                                                         const double* __restrict__ pt,
                                                         const double* __restrict__ t_mask,
                                                         const double* __restrict__ t_add, int plus_row,
                                                         double* __restrict__ out) {
  extern __shared__ double sh[];  // t_mask[S], t_add[S]
  for (int r = threadIdx.x; r < S; r += RB) { sh[r] = t_mask[r]; sh[S + r] = t_add[r]; }
  __syncthreads();
  const long long j = (long long)blockIdx.x * RB + threadIdx.x;
  if (j >= n) return;
  const double cs = colsum[j], tj = pt[j];
  double acc = 0.0;
  for (int r = 0; r < S; ++r) {
    const double d = D[(long long)r * n + j];
    const double w = wp_weight(d, sdv) / cs;
    const double sgn = (r != plus_row && tj < sh[r]) ? -1.0 : 1.0;
    const double p = __dadd_rn(__dmul_rn(d, sgn), sh[S + r]);
    acc = __dadd_rn(acc, __dmul_rn(p, w));
  }
  out[j] = acc;
}

// fates[j][t] = sum_s (w_sj / colsum_j) * B[s][t]
template <int NT>
__global__ void __launch_bounds__(RB) project_kernel(const double* __restrict__ D, int S, long long n, double sdv,
                                                     const double* __restrict__ colsum,
                                                     const double* __restrict__ B, int nt, int t0,
                                                     double* __restrict__ fates) {
  extern __shared__ double bs[];  // [chunk][NT]
  const long long j = (long long)blockIdx.x * RB + threadIdx.x;
  double acc[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) acc[t] = 0.0;
  const double cs = j < n ? colsum[j] : 1.0;
  constexpr int SCHUNK = 512;
  for (int s0 = 0; s0 < S; s0 += SCHUNK) {
    const int sc = S - s0 < SCHUNK ? S - s0 : SCHUNK;
    __syncthreads();
    for (int e = threadIdx.x; e < sc * NT; e += RB) {
      const int r = e / NT, t = e - r * NT;
      bs[e] = (t0 + t < nt) ? B[(long long)(s0 + r) * nt + t0 + t] : 0.0;
    }
    __syncthreads();
    if (j < n) {
      for (int r = 0; r < sc; ++r) {
        const double w = wp_weight(D[(long long)(s0 + r) * n + j], sdv) / cs;
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t] = fma(w, bs[r * NT + t], acc[t]);
      }
    }
  }
  if (j < n) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (t0 + t < nt) fates[j * nt + t0 + t] = acc[t];
  }
}

// scipy.stats.entropy of each row: p = row / sum(row); H = -sum p log p (0 log 0 = 0)
__global__ void row_entropy_kernel(const double* __restrict__ f, long long n, int nt, double* __restrict__ ent) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double tot = 0.0;
  for (int t = 0; t < nt; ++t) tot += f[j * nt + t];
  double h = 0.0;
  for (int t = 0; t < nt; ++t) {
    const double p = f[j * nt + t] / tot;
    if (p > 0.0) h += -p * log(p);
    else if (p < 0.0) h = -INFINITY;
  }
  ent[j] = h;
}

// partial[b][0..2] = sum (x-mx)(y-my), sum (x-mx)^2, sum (y-my)^2 over chunk b
__global__ void __launch_bounds__(RB) pearson_partial_kernel(const double* __restrict__ x,
                                                             const double* __restrict__ y, long long n, double mx,
                                                             double my, double* __restrict__ partial) {
  __shared__ double red[3][RB / 32];
  double a = 0, b = 0, c = 0;
  const long long base = (long long)blockIdx.x * RCHUNK;
  for (int t = 0; t < RCHUNK / RB; ++t) {
    const long long i = base + threadIdx.x + (long long)t * RB;
    if (i < n) {
      const double u = x[i] - mx, v = y[i] - my;
      a = fma(u, v, a); b = fma(u, u, b); c = fma(v, v, c);
    }
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double tot = 0.0;
    for (int q = 0; q < RB / 32; ++q) tot += red[threadIdx.x][q];
    partial[(long long)threadIdx.x * gridDim.x + blockIdx.x] = tot;
  }
}

__global__ void shift_scale_kernel(double* __restrict__ x, long long n, double sub, double div) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = (x[i] - sub) / div;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// out[0] = sum_i (x[i] - shift)^power, power in {1, 2}.  scratch: partial[count/4096 + 1]
int pb200_sum_pow_f64(const double* x, long long count, double shift, int power, double* partial, double* out,
                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long nb = (count + RCHUNK - 1) / RCHUNK;
  if (nb > 2147483647LL) { set_error_msg("pb200_sum_pow_f64: too many blocks"); return -1; }
  sum_pow_kernel<<<(int)nb, RB, 0, st>>>(x, count, shift, power, partial);
  PB_CHECK_LAUNCH("sum_pow_kernel");
  sum_final_kernel<<<1, 1024, 0, st>>>(partial, nb, out);
  PB_CHECK_LAUNCH("sum_final_kernel");
  return 0;
}

int pb200_waypoint_colsum(const double* D, int S, long long n, double sdv, double* colsum, void* stream) {
  colsum_kernel<<<div_up(n, RB), RB, 0, (cudaStream_t)stream>>>(D, S, n, sdv, colsum);
  PB_CHECK_LAUNCH("colsum_kernel");
  return 0;
}

int pb200_refine_step(const double* D, int S, long long n, double sdv, const double* colsum, const double* pt,
                      const double* t_mask, const double* t_add, int plus_row, double* out, void* stream) {
  const size_t smem = sizeof(double) * 2 * (size_t)S;
  if (smem > 200 * 1024) { set_error_msg("pb200_refine_step: too many waypoint rows per call (<= 12800)"); return -1; }
  PB_CHECK(cudaFuncSetAttribute(refine_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  refine_step_kernel<<<div_up(n, RB), RB, smem, (cudaStream_t)stream>>>(D, S, n, sdv, colsum, pt, t_mask, t_add,
                                                                       plus_row, out);
  PB_CHECK_LAUNCH("refine_step_kernel");
  return 0;
}

// fates[n][nt] = W_norm^T B, B[S][nt] row-major
int pb200_project_fates(const double* D, int S, long long n, double sdv, const double* colsum, const double* B,
                        int nt, double* fates, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < nt; t0 += 8) {
    const size_t smem = sizeof(double) * 512 * 8;
    project_kernel<8><<<div_up(n, RB), RB, smem, st>>>(D, S, n, sdv, colsum, B, nt, t0, fates);
    PB_CHECK_LAUNCH("project_kernel");
  }
  return 0;
}

int pb200_row_entropy(const double* fates, long long n, int nt, double* ent, void* stream) {
  row_entropy_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(fates, n, nt, ent);
  PB_CHECK_LAUNCH("row_entropy_kernel");
  return 0;
}

// out[0..2] = sum (x-mx)(y-my), sum (x-mx)^2, sum (y-my)^2.  scratch: partial[3 * (n/4096+1)]
int pb200_pearson_sums(const double* x, const double* y, long long n, double mx, double my, double* partial,
                       double* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = div_up(n, RCHUNK);
  pearson_partial_kernel<<<nb, RB, 0, st>>>(x, y, n, mx, my, partial);
  PB_CHECK_LAUNCH("pearson_partial_kernel");
  for (int q = 0; q < 3; ++q) {
    sum_final_kernel<<<1, 1024, 0, st>>>(partial + (long long)q * nb, nb, out + q);
    PB_CHECK_LAUNCH("sum_final_kernel");
  }
  return 0;
}

int pb200_shift_scale(double* x, long long n, double sub, double div, void* stream) {
  shift_scale_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, sub, div);
  PB_CHECK_LAUNCH("shift_scale_kernel");
  return 0;
}

}  // extern "C"
