// Waypoint Markov chain and FP64 dense absorption solve.
//
// Replaces _construct_markov_chain (src/palantir/core.py:523-563: pseudotime-directed
// kNN graph on the waypoints, symmetric-sigma Gaussian affinity, row-stochastic T) and
// the SuperLU solve of the absorbing chain in _differentiation_entropy
// (core.py:699-737: absorbing rows, Q = T[trans,trans], (I - Q) B = T[trans,abs]).
// The chain has nW ~ 10^3 states, so it is held dense in FP64 (nW^2 * 8 B = 11.5 MB at
// 1200 waypoints, 200 MB at 5000) and solved by Gauss-Jordan elimination with partial
// pivoting, two kernels per pivot column, no host synchronisation.
#include "common.cuh"

namespace pb200 {

// aff[i][c] for the kept arcs (0 for dropped ones), rowsum[i]
__global__ void markov_affinity_kernel(const int* __restrict__ ind, const double* __restrict__ dist, int S,
                                       int knn, int ka, const double* __restrict__ pt, double* __restrict__ aff,
                                       double* __restrict__ rowsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const double sig_i = dist[(long long)i * knn + ka];
  const double cut = pt[i] - sig_i;
  double tot = 0.0;
  for (int c = 0; c < knn; ++c) {
    const int j = ind[(long long)i * knn + c];
    const double z = dist[(long long)i * knn + c];
    double a = 0.0;
    // kept iff the neighbour is not further back than sigma_i in pseudotime; explicit
    // zeros (self, duplicates) vanish through scipy's find() in the reference
    if (pt[j] >= cut && z != 0.0) {
      const double sig_j = dist[(long long)j * knn + ka];
      const double z2 = z * z;
      a = exp(-z2 / (sig_i * sig_i) * 0.5 - z2 / (sig_j * sig_j) * 0.5);
    }
    aff[(long long)i * knn + c] = a;
    tot += a;
  }
  rowsum[i] = tot;
}

// dense T[S][S] (zero-initialised by the caller): T[i][ind] = aff / rowsum[i]
__global__ void markov_dense_kernel(const int* __restrict__ ind, const double* __restrict__ aff,
                                    const double* __restrict__ rowsum, int S, int knn, double* __restrict__ T) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)S * knn) return;
  const int i = (int)(e / knn);
  const double a = aff[e];
  if (a != 0.0) T[(long long)i * S + ind[e]] = a / rowsum[i];
}

// A[nt][nt] = I - T[trans][trans];  R[nt][na] = T[trans][abs]
__global__ void absorb_system_kernel(const double* __restrict__ T, int S, const int* __restrict__ trans, int nt,
                                     const int* __restrict__ absn, int na, double* __restrict__ A,
                                     double* __restrict__ R) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tot = (long long)nt * (nt + na);
  if (e >= tot) return;
  const int r = (int)(e / (nt + na));
  const int c = (int)(e - (long long)r * (nt + na));
  const int gi = trans[r];
  if (c < nt) {
    const double q = T[(long long)gi * S + trans[c]];
    A[(long long)r * nt + c] = (r == c ? 1.0 : 0.0) - q;
  } else {
    R[(long long)r * na + (c - nt)] = T[(long long)gi * S + absn[c - nt]];
  }
}

// ---- Gauss-Jordan with partial pivoting ----
// step k, kernel 1 (one block): pivot search in column k (rows >= k), row swap, scale pivot row.
__global__ void __launch_bounds__(1024) gj_pivot_kernel(double* __restrict__ A, double* __restrict__ B, int n,
                                                        int r, int k, int* __restrict__ singular) {
  __shared__ double sv[32];
  __shared__ int si[32];
  __shared__ int s_p;
  __shared__ double s_piv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double bv = -1.0;
  int bi = 0x7fffffff;
  for (int i = k + tid; i < n; i += 1024) {
    const double v = fabs(A[(long long)i * n + k]);
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    bv = sv[lane]; bi = si[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      s_p = bi;
      s_piv = A[(long long)bi * n + k];
      if (!(bv > 0.0)) atomicExch(singular, 1);
    }
  }
  __syncthreads();
  const int p = s_p;
  const double piv = s_piv;
  // swap rows k and p on columns >= k and on B, scaling the new pivot row by 1/piv
  for (int j = k + tid; j < n; j += 1024) {
    const double a_p = A[(long long)p * n + j];
    if (p != k) A[(long long)p * n + j] = A[(long long)k * n + j];
    A[(long long)k * n + j] = a_p / piv;
  }
  for (int j = tid; j < r; j += 1024) {
    const double b_p = B[(long long)p * r + j];
    if (p != k) B[(long long)p * r + j] = B[(long long)k * r + j];
    B[(long long)k * r + j] = b_p / piv;
  }
}

// step k, kernel 2: every row i != k: row_i -= a_ik * row_k on columns > k and on B.
// grid (ceil((n-k-1+r)/256), n)
__global__ void __launch_bounds__(256) gj_update_kernel(double* __restrict__ A, double* __restrict__ B, int n,
                                                        int r, int k) {
  const int i = blockIdx.y;
  if (i == k) return;
  const double f = A[(long long)i * n + k];
  if (f == 0.0) return;
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int na = n - k - 1;
  if (c < na) {
    const int j = k + 1 + c;
    A[(long long)i * n + j] = fma(-f, A[(long long)k * n + j], A[(long long)i * n + j]);
  } else if (c - na < r) {
    const int j = c - na;
    B[(long long)i * r + j] = fma(-f, B[(long long)k * r + j], B[(long long)i * r + j]);
  }
}

// y = M^T x for a dense row-major M[S][S] (stationary-vector power iteration)
__global__ void __launch_bounds__(256) dense_matvec_t_kernel(const double* __restrict__ M, int S,
                                                             const double* __restrict__ x,
                                                             double* __restrict__ y) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= S) return;
  double s = 0.0;
  for (int i = 0; i < S; ++i) s = fma(M[(long long)i * S + j], x[i], s);
  y[j] = s;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Dense waypoint transition matrix T[S][S] (caller zero-fills T).  ind/dist [S][knn] is the
// self-including neighbour table of the waypoints, pt[S] their pseudotime, ka = min(knn/3-1, 30).
// scratch: aff[S*knn], rowsum[S]
int pb200_markov_chain_dense(const int* ind, const double* dist, int S, int knn, int ka, const double* pt,
                             double* aff, double* rowsum, double* T, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (ka < 0 || ka >= knn) { set_error_msg("pb200_markov_chain_dense: adaptive index out of range"); return -1; }
  markov_affinity_kernel<<<div_up(S, 128), 128, 0, st>>>(ind, dist, S, knn, ka, pt, aff, rowsum);
  PB_CHECK_LAUNCH("markov_affinity_kernel");
  markov_dense_kernel<<<div_up((long long)S * knn, 256), 256, 0, st>>>(ind, aff, rowsum, S, knn, T);
  PB_CHECK_LAUNCH("markov_dense_kernel");
  return 0;
}

int pb200_absorb_system(const double* T, int S, const int* trans, int nt, const int* absn, int na, double* A,
                        double* R, void* stream) {
  if (nt <= 0) return 0;
  absorb_system_kernel<<<div_up((long long)nt * (nt + na), 256), 256, 0, (cudaStream_t)stream>>>(T, S, trans, nt,
                                                                                                absn, na, A, R);
  PB_CHECK_LAUNCH("absorb_system_kernel");
  return 0;
}

// Solve A X = B in place (A[n][n] destroyed, B[n][r] <- X).  singular[0] set to 1 on a zero pivot.
int pb200_dense_solve_f64(double* A, double* B, int n, int r, int* singular, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PB_CHECK(cudaMemsetAsync(singular, 0, sizeof(int), st));
  for (int k = 0; k < n; ++k) {
    gj_pivot_kernel<<<1, 1024, 0, st>>>(A, B, n, r, k, singular);
    PB_CHECK_LAUNCH("gj_pivot_kernel");
    const int cols = n - k - 1 + r;
    if (cols > 0 && n > 1) {
      gj_update_kernel<<<dim3(div_up(cols, 256), n), 256, 0, st>>>(A, B, n, r, k);
      PB_CHECK_LAUNCH("gj_update_kernel");
    }
  }
  return 0;
}

int pb200_dense_matvec_t(const double* M, int S, const double* x, double* y, void* stream) {
  dense_matvec_t_kernel<<<div_up(S, 256), 256, 0, (cudaStream_t)stream>>>(M, S, x, y);
  PB_CHECK_LAUNCH("dense_matvec_t_kernel");
  return 0;
}

}  // extern "C"
