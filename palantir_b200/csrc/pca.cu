// Truncated SVD without centring for run_pca (src/palantir/utils.py:52-118: scanpy's
// sc.pp.pca(zero_center=False) = sklearn TruncatedSVD): the O(N) parts on the device.
//
//   gram:     C[G][G] += X^T X over a block of rows (FP64 accumulation; only tiles with j >= i are computed,
//             the host mirrors them) -- the right singular vectors are the eigenvectors of C, a G x G problem
//             (G = highly variable genes, a few thousand) solved on the host like ARPACK's own small problems;
//   project:  Y[N][k] = X V[G][k] in FP64 (the scores U S).
//
// One generic shared-memory-tiled FP64 kernel serves both: OUT(i, j) (+)= sum_k A(i, k) B(k, j) with arbitrary
// strides, 64 x 64 output tiles, 16-deep k slabs, 4 x 4 outputs per thread.  These are HBM / FP64-pipe bound
// SIMT kernels (FP64 has no tcgen05 path); they run once per call, upstream of the hot path.
#include "common.cuh"

namespace pb200 {

constexpr int PT = 64;   // output tile
constexpr int PK = 16;   // k slab

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) tile_gemm_f64_kernel(const TA* __restrict__ a, long long sa_i, long long sa_k,
                                                            const TB* __restrict__ b, long long sb_k, long long sb_j,
                                                            double* __restrict__ out, long long ldo, long long m,
                                                            long long nn, long long kk, long long k_per_slab,
                                                            int upper_only, int atomic) {
  __shared__ double As[PK][PT + 1];
  __shared__ double Bs[PK][PT + 1];
  const long long i0 = (long long)blockIdx.y * PT, j0 = (long long)blockIdx.x * PT;
  if (upper_only && j0 + PT <= i0) return;          // strictly below the diagonal: the host mirrors the upper part
  const long long k_beg = (long long)blockIdx.z * k_per_slab;
  const long long k_end = k_beg + k_per_slab < kk ? k_beg + k_per_slab : kk;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 x 16 threads, 4 x 4 outputs each
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  // loaders: the fastest-varying thread index follows the unit-stride dimension of each operand
  const bool a_i_fast = sa_i <= sa_k, b_j_fast = sb_j <= sb_k;
  for (long long k0 = k_beg; k0 < k_end; k0 += PK) {
    for (int t = threadIdx.x; t < PK * PT; t += 256) {
      const int ii = a_i_fast ? t % PT : t / PK, kq = a_i_fast ? t / PT : t % PK;
      const long long gi = i0 + ii, gk = k0 + kq;
      As[kq][ii] = (gi < m && gk < k_end) ? (double)a[gi * sa_i + gk * sa_k] : 0.0;
    }
    for (int t = threadIdx.x; t < PK * PT; t += 256) {
      const int jj = b_j_fast ? t % PT : t / PK, kq = b_j_fast ? t / PT : t % PK;
      const long long gj = j0 + jj, gk = k0 + kq;
      Bs[kq][jj] = (gj < nn && gk < k_end) ? (double)b[gk * sb_k + gj * sb_j] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kq = 0; kq < PK; ++kq) {
      double av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) av[u] = As[kq][ty + 16 * u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bv[v] = Bs[kq][tx + 16 * v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long gi = i0 + ty + 16 * u;
    if (gi >= m) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const long long gj = j0 + tx + 16 * v;
      if (gj >= nn) continue;
      if (atomic) atomicAdd(&out[gi * ldo + gj], acc[u][v]);
      else out[gi * ldo + gj] = acc[u][v];
    }
  }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// C[g][g] += X^T X for the n rows of x (row-major, leading dimension ldx; FP32 or FP64); only the tiles on or
// above the diagonal are touched.  C must be zeroed before the first block of rows.
int pb200_pca_gram(const void* x, int is_f64, long long n, int g, long long ldx, double* C, void* stream) {
  if (n <= 0 || g <= 0) return 0;
  const int tiles = div_up(g, PT);
  long long slabs = (long long)(4 * kSMs) / ((long long)tiles * (tiles + 1) / 2) + 1;
  if (slabs > div_up(n, 4 * PK)) slabs = div_up(n, 4 * PK);
  if (slabs < 1) slabs = 1;
  if (slabs > 65535) slabs = 65535;
  long long per = (n + slabs - 1) / slabs;
  per = (per + PK - 1) / PK * PK;
  dim3 grid(tiles, tiles, (unsigned)div_up(n, per));
  if (is_f64)
    tile_gemm_f64_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const double*)x, 1, ldx, (const double*)x, ldx, 1, C, g, g, g, n, per, 1, 1);
  else
    tile_gemm_f64_kernel<float, float><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float*)x, 1, ldx, (const float*)x, ldx, 1, C, g, g, g, n, per, 1, 1);
  PB_CHECK_LAUNCH("tile_gemm_f64_kernel(gram)");
  return 0;
}

// Y[n][k] = X[n][g] V[g][k] (V row-major FP64), FP64 result.
int pb200_pca_project(const void* x, int is_f64, long long n, int g, long long ldx, const double* V, int k,
                      double* Y, void* stream) {
  if (n <= 0 || g <= 0 || k <= 0) return 0;
  dim3 grid(div_up(k, PT), (unsigned)div_up(n, PT), 1);
  if (grid.y > 65535) { set_error_msg("pb200_pca_project: too many rows for one call (split the rows)"); return -1; }
  if (is_f64)
    tile_gemm_f64_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const double*)x, ldx, 1, V, k, 1, Y, k, n, k, g, g, 0, 0);
  else
    tile_gemm_f64_kernel<float, double><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const float*)x, ldx, 1, V, k, 1, Y, k, n, k, g, g, 0, 0);
  PB_CHECK_LAUNCH("tile_gemm_f64_kernel(project)");
  return 0;
}

}  // extern "C"
