// Symmetric Lanczos eigensolver pieces: CSR SpMV/SpMM and the dense vector
// kernels of a fully re-orthogonalised Lanczos recurrence.
//
// Replaces scipy.sparse.linalg.eigs (ARPACK) at src/palantir/utils.py:484.  The
// reference asks for the largest-magnitude eigenpairs of the row-stochastic
// T = D^-1 K; T is similar to the symmetric S = D^-1/2 K D^-1/2, so the solver
// runs symmetric Lanczos on S (eigenvalues identical, eigenvectors of T are
// D^-1/2 u) -- SURVEY.md 7a step 4.
//
// The recurrence runs on the device with no host round trip per step: alpha and
// beta stay in device memory and every kernel reads the scalars it needs from
// there.  Each step applies the three-term recurrence (w = S v_j - alpha_j v_j -
// beta_j v_{j-1}) and then one full classical Gram-Schmidt pass against the whole
// basis (h = V^T w, w -= V h): after the recurrence w is already of size beta_{j+1},
// so the pass only removes round-off-level components and there is no cancellation
// that would call for a second pass.  All basis vectors are read with coalesced
// row-contiguous loads; all reductions are two-stage and deterministic (no FP atomics).
#include "common.cuh"

namespace pb200 {

// ---------------------------------------------------------------------------
// SpMV: y = A x, CSR with FP64 values / int32 columns.  L lanes cooperate on a
// row (L = 8, 16 or 32 chosen from the mean row length); the value and column
// streams are read coalesced, x is gathered through L2.  HBM-bound:
// 12 B/nnz + 4 B/row (indptr) + 8 B/row (y) + x once.
// ---------------------------------------------------------------------------
constexpr int SPMV_LONG_ROW = 128;   // rows above this many entries go to csr_spmv_long_rows_kernel when a list is given

// Rows [row0, row1) of the product (the whole matrix by default; a rank's slice when the eigensolver is row-sharded).
template <int L>
__global__ void __launch_bounds__(256) csr_spmv_kernel(const int* __restrict__ indptr,
                                                       const int* __restrict__ indices,
                                                       const double* __restrict__ vals,
                                                       const double* __restrict__ x, double* __restrict__ y,
                                                       long long row0, long long row1, int skip_long) {
  const int lane = threadIdx.x & (L - 1);
  const long long row = row0 + ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (row >= row1) return;  // whole sub-warp leaves together
  const int beg = indptr[row], end = indptr[row + 1];
  if (skip_long && end - beg > SPMV_LONG_ROW) return;
  double s0 = 0.0, s1 = 0.0;
  int e = beg + lane;
  for (; e + L < end; e += 2 * L) {
    const int c0 = indices[e], c1 = indices[e + L];
    const double v0 = vals[e], v1 = vals[e + L];
    s0 = fma(v0, __ldg(x + c0), s0);
    s1 = fma(v1, __ldg(x + c1), s1);
  }
  if (e < end) s0 = fma(vals[e], __ldg(x + indices[e]), s0);
  double s = s0 + s1;
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, L);
  if (lane == 0) y[row] = s;
}

// ---------------------------------------------------------------------------
// SpMV over the NONZERO stream ("CSR-stream"): a CTA takes ST_TILE consecutive nonzeros whatever rows they belong to,
// so the hubs of the symmetrised kNN graph (rows of up to ~4000 entries at 1 M cells, 13 % of the nonzeros in 5 % of
// the rows) cost what their nonzeros cost and need no kernel of their own.  Phase 1: every thread loads four
// (column, value) pairs -- coalesced, aligned to the tile, not to row starts -- and issues its four x gathers
// back to back; the products go to shared memory.  Phase 2: groups of 16 lanes sum the row segments of the tile
// (lane-strided partial sums, then a shuffle tree: the order is fixed, results are reproducible).  A row that lies
// inside one tile is written to y; a row cut by a tile boundary leaves its pieces in tail[tile] (the piece that
// starts the row) and head[tile] (the pieces that continue it), which csr_stream_fixup_kernel adds in tile order.
// tile_lo[b] = first row that STARTS at or after nonzero b * ST_TILE (csr_stream_tiles_kernel, once per matrix).
// ---------------------------------------------------------------------------
constexpr int ST_TILE = 1024;     // nonzeros per CTA (4 per thread)
constexpr int ST_THREADS = 256;

__global__ void csr_stream_tiles_kernel(const int* __restrict__ indptr, long long n, long long nnz, int n_tiles,
                                        int* __restrict__ tile_lo) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_tiles) return;
  if (b == n_tiles) { tile_lo[b] = (int)n; return; }
  const long long target = (long long)b * ST_TILE;
  // first r in [0, n] with indptr[r] >= target
  long long lo = 0, hi = n;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)indptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  tile_lo[b] = (int)lo;
}

__global__ void __launch_bounds__(ST_THREADS) csr_spmv_stream_kernel(const int* __restrict__ indptr,
                                                                     const int* __restrict__ indices,
                                                                     const double* __restrict__ vals,
                                                                     const double* __restrict__ x, double* __restrict__ y,
                                                                     long long nnz, const int* __restrict__ tile_lo,
                                                                     double* __restrict__ head, double* __restrict__ tail,
                                                                     int tile0, long long row0, long long row1) {
  __shared__ double prod[ST_TILE];
  const int b = tile0 + blockIdx.x;
  const long long base = (long long)b * ST_TILE;
  long long end = base + ST_TILE;
  if (end > nnz) end = nnz;
  const int lo = tile_lo[b], hi = tile_lo[b + 1];          // rows starting in this tile: [lo, hi)
  {
    int c[4];
    double v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const long long e = base + threadIdx.x + t * ST_THREADS;
      const bool ok = e < end;
      c[t] = ok ? __ldg(indices + e) : 0;
      v[t] = ok ? __ldg(vals + e) : 0.0;
    }
    double xg[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) xg[t] = __ldg(x + c[t]);
#pragma unroll
    for (int t = 0; t < 4; ++t) prod[threadIdx.x + t * ST_THREADS] = v[t] * xg[t];
  }
  // the nonzero at `base` belongs to a row that started in an earlier tile unless row `lo` starts exactly there
  // (tile_lo = first row with indptr >= base, so indptr[lo - 1] < base whenever indptr[lo] > base)
  const int has_prev = (long long)indptr[lo] > base ? 1 : 0;
  const int nseg = has_prev + (hi - lo);
  const int lane = threadIdx.x & 15, grp = threadIdx.x >> 4;
  const unsigned half_mask = 0xffffu << (threadIdx.x & 16);   // the 16 lanes of this group (groups of a warp may differ in trip count)
  __syncthreads();
  for (int sgm = grp; sgm < nseg; sgm += ST_THREADS / 16) {
    const bool prev = has_prev && sgm == 0;
    const long long r = prev ? lo - 1 : lo + sgm - has_prev;
    const long long rb = prev ? base : (long long)indptr[r];
    const long long re_full = (long long)indptr[r + 1];
    const long long re = re_full < end ? re_full : end;
    double sum = 0.0;
    for (long long e = rb + lane; e < re; e += 16) sum += prod[e - base];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(half_mask, sum, o, 16);
    if (lane == 0) {
      if (prev) head[b] = sum;
      else if (re_full > end) tail[b] = sum;
      else if (r >= row0 && r < row1) y[r] = sum;
    }
  }
}

// rows cut by tile boundaries: y[r] = tail[first tile] + head[next tiles...] in tile order.  One thread per tile.
__global__ void csr_stream_fixup_kernel(const int* __restrict__ indptr, long long nnz, const int* __restrict__ tile_lo,
                                        const double* __restrict__ head, const double* __restrict__ tail,
                                        double* __restrict__ y, int tile0, int tile1, long long row0, long long row1) {
  const int b = tile0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= tile1) return;
  const int lo = tile_lo[b], hi = tile_lo[b + 1];
  if (hi <= lo) return;
  const long long r = hi - 1;                                // last row starting in this tile
  const long long re = (long long)indptr[r + 1];
  long long end = (long long)(b + 1) * ST_TILE;
  if (re <= end) return;                                     // it also ends here
  if (r < row0 || r >= row1) return;
  double acc = tail[b];
  for (int c = b + 1;; ++c) {
    acc += head[c];
    end += ST_TILE;
    if (re <= end) break;
  }
  y[r] = acc;
}

// ---------------------------------------------------------------------------
// SpMV with the x window staged in shared memory, for matrices whose rows are in a
// locality-preserving order (the kNN stage's Morton order: 70 % of the kernel's columns lie within
// +-12 k of the row at 1 M cells).  The plain kernel above is bound by L1 gather wavefronts
// (one 32 B sector per nonzero, ~2 cycles each per SM); here a persistent CTA walks a contiguous row
// range in blocks of SW_R rows and keeps x[r0 - H, r0 + R + H) in a 192 KB ring in shared memory
// (position = column mod SW_C), so those gathers become shared-memory reads; the R new columns of the
// next block are fetched with cp.async while the current block is computed; columns outside the
// window fall back to the global gather.
// ---------------------------------------------------------------------------
constexpr int SW_THREADS = 512;
constexpr int SW_R = 1024;                 // rows per block
constexpr int SW_C = 24576;                // ring capacity in doubles (192 KB)
constexpr int SW_H = (SW_C - 2 * SW_R) / 2;   // halo on each side of the block: 11264 columns

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(SW_THREADS, 1) csr_spmv_window_kernel(const int* __restrict__ indptr,
                                                                        const int* __restrict__ indices,
                                                                        const double* __restrict__ vals,
                                                                        const double* __restrict__ x,
                                                                        double* __restrict__ y, long long n,
                                                                        long long rows_per_cta) {
  extern __shared__ __align__(16) double xs[];
  const int tid = threadIdx.x, lane16 = tid & 15, grp = tid >> 4;       // 32 row groups of 16 lanes
  const long long rb = (long long)blockIdx.x * rows_per_cta;
  long long re = rb + rows_per_cta;
  if (re > n) re = n;
  if (rb >= n) return;
  // first window
  {
    long long lo = rb - SW_H, hi = rb + SW_R + SW_H;
    if (lo < 0) lo = 0;
    if (hi > n) hi = n;
    for (long long c = lo + tid; c < hi; c += SW_THREADS) cp_async8(&xs[c % SW_C], x + c);
  }
  for (long long r0 = rb; r0 < re; r0 += SW_R) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();                                     // window of this block complete, previous block done
    {
      // columns that only the NEXT block needs: they overwrite the slots of columns below this block's window
      long long lo = r0 + SW_R + SW_H, hi = r0 + 2 * SW_R + SW_H;
      if (hi > n) hi = n;
      if (r0 + SW_R < re)
        for (long long c = lo + tid; c < hi; c += SW_THREADS) cp_async8(&xs[c % SW_C], x + c);
    }
    // window [wlo, whi) of this block; ring position of column c = c - ring0, minus SW_C when that is >= SW_C
    const int wlo = (int)(r0 - SW_H), whi = (int)(r0 + SW_R + SW_H);
    const int ring0 = (int)((r0 >= SW_H ? (r0 - SW_H) / SW_C : 0) * SW_C);
    long long rend = r0 + SW_R;
    if (rend > re) rend = re;
    // A group of 16 lanes takes FOUR consecutive rows at a time, four entries per lane and row (64 slots
    // per row; longer rows loop), and issues every value / column load of the batch before the first
    // gather, and every gather (one shared-memory and one global load per slot, the unused one at a
    // harmless address: no branches) before the first multiply: 16 warps per SM have to cover the
    // HBM latency with loads in flight, not with occupancy.
    int bnd[5], bnd_next[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const long long rr = r0 + 4 * grp + u < rend ? r0 + 4 * grp + u : rend;
      bnd_next[u] = __ldg(indptr + rr);
    }
    for (long long rbase = r0 + 4 * grp; rbase < rend; rbase += 4 * (SW_THREADS / 16)) {
#pragma unroll
      for (int u = 0; u < 5; ++u) bnd[u] = bnd_next[u];
      {
        const long long nb = rbase + 4 * (SW_THREADS / 16);     // the group's next batch: extents in flight meanwhile
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const long long rr = nb + u < rend ? nb + u : rend;
          bnd_next[u] = __ldg(indptr + rr);
        }
      }
      int c[4][4];
      double v[4][4], xg[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = bnd[u] + lane16 + 16 * t;
          const bool ok = e < bnd[u + 1];
          c[u][t] = ok ? __ldg(indices + e) : -1;
          v[u][t] = ok ? __ldg(vals + e) : 0.0;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int cc = c[u][t];
          const bool inwin = cc >= wlo && cc < whi;
          const bool outwin = cc >= 0 && !inwin;
          int pos = cc - ring0;
          pos = pos >= SW_C ? pos - SW_C : pos;
          xg[u][t] = __ldg(x + (outwin ? cc : 0));
          const double xsv = xs[inwin ? pos : 0];
          v[u][t] *= inwin ? xsv : 1.0;
          c[u][t] = outwin ? 1 : 0;
        }
      double acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u] = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[u] += c[u][t] ? v[u][t] * xg[u][t] : v[u][t];
        // rows with more than 64 entries (hubs of the symmetrised graph): plain loop for the rest
        for (int e = bnd[u] + 64 + lane16; e < bnd[u + 1]; e += 16) {
          const int cc = __ldg(indices + e);
          acc[u] = fma(__ldg(vals + e), __ldg(x + cc), acc[u]);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o, 16);
      if (lane16 < 4 && rbase + lane16 < rend)
        y[rbase + lane16] = lane16 == 0 ? acc[0] : (lane16 == 1 ? acc[1] : (lane16 == 2 ? acc[2] : acc[3]));
    }
  }
}

// Relabel a CSR matrix: new row r' = old row perm[r'], new column = inv[old column]
// (indptr_new = prefix sums of the permuted row lengths, supplied by the caller).  Columns inside a
// row keep the old order (not sorted by new label; the SpMV does not need it).
__global__ void __launch_bounds__(256) csr_permute_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                                          const double* __restrict__ vals, long long n,
                                                          const int* __restrict__ perm, const int* __restrict__ inv,
                                                          const int* __restrict__ indptr_new,
                                                          int* __restrict__ indices_new, double* __restrict__ vals_new) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + warp;
  if (r >= n) return;
  const int old = perm[r];
  const int b = indptr[old], len = indptr[old + 1] - b, nb = indptr_new[r];
  for (int t = lane; t < len; t += 32) {
    indices_new[nb + t] = inv[indices[b + t]];
    vals_new[nb + t] = vals[b + t];
  }
}

// The long rows of the same product, one warp per row, eight gathers in flight per lane.  The symmetrised
// kNN graph of noisy high-dimensional data has hubs (longest row 3948 at 1 M cells, 5 % of the rows above
// 128 entries holding 13 % of the nonzeros): walked by 16 lanes they are a third of the SpMV's time.
__global__ void __launch_bounds__(256) csr_spmv_long_rows_kernel(const int* __restrict__ indptr,
                                                                 const int* __restrict__ indices,
                                                                 const double* __restrict__ vals,
                                                                 const double* __restrict__ x, double* __restrict__ y,
                                                                 const int* __restrict__ rows, int n_rows) {
  const int lane = threadIdx.x & 31;
  const int w = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (w >= n_rows) return;
  const int row = rows[w];
  const int beg = indptr[row], end = indptr[row + 1];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int e = beg + lane;
  for (; e + 96 < end; e += 128) {
    const int c0 = indices[e], c1 = indices[e + 32], c2 = indices[e + 64], c3 = indices[e + 96];
    const double v0 = vals[e], v1 = vals[e + 32], v2 = vals[e + 64], v3 = vals[e + 96];
    s0 = fma(v0, __ldg(x + c0), s0);
    s1 = fma(v1, __ldg(x + c1), s1);
    s2 = fma(v2, __ldg(x + c2), s2);
    s3 = fma(v3, __ldg(x + c3), s3);
  }
  for (; e < end; e += 32) s0 = fma(vals[e], __ldg(x + indices[e]), s0);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s;
}

// SpMM for MAGIC: Y[n][g] (+)= A X[n][g], FP32 dense operands, FP64 CSR values
// converted on the fly (the reference casts T^t to float32, utils.py:791).
// One warp per row; lanes stride over the g columns (coalesced X row reads).
__global__ void __launch_bounds__(256) csr_spmm_f32_kernel(const int* __restrict__ indptr,
                                                           const int* __restrict__ indices,
                                                           const double* __restrict__ vals,
                                                           const float* __restrict__ x, float* __restrict__ y,
                                                           long long n, int g, long long ldx, long long ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= n) return;
  const int beg = indptr[row], end = indptr[row + 1];
  for (int c0 = lane; c0 < g; c0 += 128) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int e = beg; e < end; ++e) {
      const float v = (float)vals[e];
      const float* xr = x + (long long)indices[e] * ldx;
      a0 = fmaf(v, xr[c0], a0);
      if (c0 + 32 < g) a1 = fmaf(v, xr[c0 + 32], a1);
      if (c0 + 64 < g) a2 = fmaf(v, xr[c0 + 64], a2);
      if (c0 + 96 < g) a3 = fmaf(v, xr[c0 + 96], a3);
    }
    float* yr = y + row * ldy;
    yr[c0] = a0;
    if (c0 + 32 < g) yr[c0 + 32] = a1;
    if (c0 + 64 < g) yr[c0 + 64] = a2;
    if (c0 + 96 < g) yr[c0 + 96] = a3;
  }
}

// Gene-tiled SpMM for MAGIC (utils.py:791-804) at scale: Y[:, g0:g0+GT] = A X[:, g0:g0+GT], one launch per tile of GT
// gene columns, one warp per row, the row's CSR entries read once per tile (coalesced, FP32 values converted once
// beforehand), every lane holding 8 columns of the tile (two 16-byte loads per nonzero, 4 nonzeros in flight).
// With the rows in a locality-preserving order the X rows a neighbourhood of output rows needs (~40 k rows x 1 KB)
// stay in L2 while the tile is swept, so HBM sees X and Y once per tile: 16.6 GB per product at 1 M cells x 2000
// genes.  What bounds the kernel is the gather itself: nnz x G x 4 B = 424 GB per product from L2 to the SMs
// (every (nonzero, gene) pair is a load; the rows of a neighbourhood share almost no neighbours, so shared memory
// cannot hold anything worth re-using).
template <typename TX, int VEC>     // VEC elements per 16-byte load: 4 (float) or 2 (double)
__global__ void __launch_bounds__(256) csr_spmm_tile_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                                            const float* __restrict__ vals, const TX* __restrict__ x,
                                                            TX* __restrict__ y, long long n, int g, long long ldx,
                                                            long long ldy, int g0) {
  constexpr int GT = 64 * VEC;                  // columns per tile: 256 (float) / 128 (double)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= n) return;
  const int beg = indptr[row], end = indptr[row + 1];
  const int ca = g0 + lane * VEC, cb = g0 + (lane + 32) * VEC;       // first column of this lane's two vectors
  const bool fa = ca + VEC <= g, fb = cb + VEC <= g;                 // whole vector inside the matrix
  TX acc[2 * VEC];
#pragma unroll
  for (int k = 0; k < 2 * VEC; ++k) acc[k] = (TX)0;
  struct alignas(16) V16 { TX v[VEC]; };
  auto load = [&](const TX* rowp, int c0, bool full, V16& out) {
    if (full) {
      out = *reinterpret_cast<const V16*>(rowp + c0);
    } else {
#pragma unroll
      for (int k = 0; k < VEC; ++k) out.v[k] = (c0 + k < g) ? rowp[c0 + k] : (TX)0;
    }
  };
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int my_c = e < end ? indices[e] : 0;
    const float my_v = e < end ? vals[e] : 0.f;
    const int cnt = end - e0 < 32 ? end - e0 : 32;
    for (int j0 = 0; j0 < cnt; j0 += 4) {
      V16 xa[4], xb[4];
      TX v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u < cnt ? j0 + u : cnt - 1;             // clamped: a repeated entry gets weight 0 below
        const int c = __shfl_sync(0xffffffffu, my_c, j);
        const float vv = __shfl_sync(0xffffffffu, my_v, j);
        v[u] = j0 + u < cnt ? (TX)vv : (TX)0;
        const TX* rowp = x + (long long)c * ldx;
        load(rowp, ca, fa, xa[u]);
        load(rowp, cb, fb, xb[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          acc[k] = fma(v[u], xa[u].v[k], acc[k]);
          acc[VEC + k] = fma(v[u], xb[u].v[k], acc[VEC + k]);
        }
      }
    }
  }
  TX* yr = y + row * ldy;
  if (fa) {
    V16 o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = acc[k];
    *reinterpret_cast<V16*>(yr + ca) = o;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) if (ca + k < g) yr[ca + k] = acc[k];
  }
  if (fb) {
    V16 o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = acc[VEC + k];
    *reinterpret_cast<V16*>(yr + cb) = o;
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) if (cb + k < g) yr[cb + k] = acc[VEC + k];
  }
  (void)GT;
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// out[i][:] = (in[i][:] < thr) ? 0 : in[i][:] over an [n][ld] matrix (MAGIC's clip, utils.py:815-821), in place
template <typename TX>
__global__ void clip_below_kernel(TX* __restrict__ a, long long count, TX thr) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count && a[i] < thr) a[i] = (TX)0;
}

// ---------------------------------------------------------------------------
// dense vector kernels
// ---------------------------------------------------------------------------
constexpr int VB = 256;        // threads
constexpr int VCHUNK = 2048;   // elements per block (8 per thread)

// sum of partial[0..nblk) with the summation order of reduce_partials_kernel at 256 threads; all threads get it
__device__ __forceinline__ double block_sum_partials(const double* partial, int nblk, double* red) {
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += VB) s += partial[b];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int q = 0; q < VB / 32; ++q) tot += red[q];
  __syncthreads();
  return tot;
}

// partial[c * nblk + b] = sum_{i in chunk b} V[c][i] * w[i],  c = 0..m-1
template <typename TV>
__global__ void __launch_bounds__(VB) multi_dot_kernel(const TV* __restrict__ V, long long ldv, int m,
                                                       const double* __restrict__ w, long long n,
                                                       double* __restrict__ partial, const int* __restrict__ gate = nullptr) {
  __shared__ double red[VB / 32];
  if (gate && *gate == 0) return;
  const long long base = (long long)blockIdx.x * VCHUNK;
  double wr[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const long long i = base + threadIdx.x + (long long)t * VB;
    wr[t] = i < n ? w[i] : 0.0;
  }
  for (int c = 0; c < m; ++c) {
    const TV* v = V + (long long)c * ldv;
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const long long i = base + threadIdx.x + (long long)t * VB;
      if (i < n) s = fma((double)v[i], wr[t], s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < VB / 32; ++q) tot += red[q];
      partial[(long long)c * gridDim.x + blockIdx.x] = tot;
    }
    __syncthreads();
  }
}

// h[c] = sum_b partial[c][b]; optionally hacc[c] += h[c]
// fix_j >= 0: additionally alpha_fix[fix_j] += h[fix_j] (the Gram-Schmidt pass's correction of alpha_j)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nblk, int m,
                                       double* __restrict__ h, double* __restrict__ hacc,
                                       const int* __restrict__ gate = nullptr, int fix_j = -1,
                                       double* __restrict__ alpha_fix = nullptr) {
  const int c = blockIdx.x;
  if (c >= m) return;
  if (gate && *gate == 0) return;
  __shared__ double red[8];
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) s += partial[(long long)c * nblk + b];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += red[q];
    h[c] = tot;
    if (hacc) hacc[c] += tot;
    if (c == fix_j) alpha_fix[c] += tot;
  }
}

// w[i] -= sum_c h[c] V[c][i]
template <typename TV>
__global__ void __launch_bounds__(VB) multi_axpy_kernel(const TV* __restrict__ V, long long ldv, int m,
                                                        const double* __restrict__ h, double* __restrict__ w,
                                                        long long n, const int* __restrict__ gate = nullptr) {
  extern __shared__ double hs[];
  if (gate && *gate == 0) return;
  for (int c = threadIdx.x; c < m; c += VB) hs[c] = h[c];
  __syncthreads();
  const long long base = (long long)blockIdx.x * VCHUNK;
  double acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.0;
  for (int c = 0; c < m; ++c) {
    const TV* v = V + (long long)c * ldv;
    const double hc = hs[c];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const long long i = base + threadIdx.x + (long long)t * VB;
      if (i < n) acc[t] = fma(hc, (double)v[i], acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const long long i = base + threadIdx.x + (long long)t * VB;
    if (i < n) w[i] -= acc[t];
  }
}

// out[i] = w[i] / sqrt(*nrm2)    (also records sqrt(*nrm2) into *beta_out if given; out32 = FP32 shadow copy)
__global__ void scale_by_norm_kernel(const double* __restrict__ w, const double* __restrict__ nrm2,
                                     double* __restrict__ out, float* __restrict__ out32, long long n,
                                     double* __restrict__ beta_out) {
  const double b = sqrt(*nrm2);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && beta_out) *beta_out = b;
  if (i < n) {
    const double v = b > 0.0 ? w[i] / b : 0.0;
    out[i] = v;
    if (out32) out32[i] = (float)v;
  }
}

// h2[0] = beta[j] (0 when j == 0), h2[1] = alpha[j]
__global__ void recurrence_coeffs_kernel(const double* __restrict__ alpha, const double* __restrict__ beta, int j,
                                         double* __restrict__ h2) {
  if (threadIdx.x == 0) {
    h2[0] = j > 0 ? beta[j] : 0.0;
    h2[1] = alpha[j];
  }
}

// Partial reorthogonalisation (Simon 1984, as in PROPACK): om[r][k], r = j mod 3, estimates v_j . v_k.  After the
// three-term recurrence of step j (alpha[j] known, nrm2 = |w|^2 of the unnormalised v_{j+1}) this computes the
// estimates for v_{j+1}, and raises gate[0] when one of them exceeds `thresh` (or when the previous step did:
// the reorthogonalisation is always done on two consecutive vectors, gate[1] carries the second).  A raised gate
// makes the gated Gram-Schmidt kernels that follow do their work; the estimates are then reset to eps.
// nrm2_partials != NULL (blockDim.x == 256): |w|^2 = their sum (as reduce_partials_kernel<<<1, 256>>> forms it),
// also stored to *nrm2 for the kernels that follow.
__global__ void lanczos_omega_kernel(const double* __restrict__ alpha, const double* __restrict__ beta, int j,
                                     double* nrm2, double* __restrict__ om, long long omstride,
                                     double eps, double thresh, int* __restrict__ gate, int* __restrict__ counters,
                                     const double* nrm2_partials = nullptr, int nblk = 0) {
  __shared__ double red[8];
  double n2;
  if (nrm2_partials) {
    n2 = block_sum_partials(nrm2_partials, nblk, red);
    if (threadIdx.x == 0) *nrm2 = n2;
  } else {
    n2 = *nrm2;
  }
  const double bj1 = sqrt(n2);
  const double* cur = om + (long long)(j % 3) * omstride;            // omega_{j, .}
  const double* prv = om + (long long)((j + 2) % 3) * omstride;      // omega_{j-1, .}
  double* nxt = om + (long long)((j + 1) % 3) * omstride;            // omega_{j+1, .}
  const double aj = alpha[j], bj = j > 0 ? beta[j] : 0.0;
  double mx = 0.0;
  for (int k = threadIdx.x; k < j; k += blockDim.x) {
    double num = beta[k + 1] * cur[k + 1] + (alpha[k] - aj) * cur[k] - bj * prv[k];
    if (k > 0) num += beta[k] * cur[k - 1];
    const double noise = eps * (beta[k + 1] + bj1);
    num += num >= 0.0 ? noise : -noise;
    const double o = bj1 > 0.0 ? num / bj1 : 0.0;
    nxt[k] = o;
    mx = fmax(mx, fabs(o));
  }
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  __shared__ int s_do;
  if (threadIdx.x == 0) {
    double m2 = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m2 = fmax(m2, red[q]);
    const int forced = gate[1];
    const int go = (m2 > thresh || forced || !(bj1 > 0.0)) ? 1 : 0;
    gate[0] = go;
    gate[1] = (go && !forced) ? 1 : 0;
    if (go) counters[0] += 1;
    nxt[j] = j > 0 && bj1 > 0.0 ? eps * beta[1] / bj1 : eps;
    nxt[j + 1] = 1.0;
    s_do = go;
  }
  __syncthreads();
  if (s_do)
    for (int k = threadIdx.x; k <= j; k += blockDim.x) nxt[k] = eps;
}

// tmp2[0] = |w|^2 recomputed only when the gate is up (w changed); alpha[j] += h[j] likewise
__global__ void lanczos_fix_alpha_kernel(const double* __restrict__ h, int j, double* __restrict__ alpha,
                                         const int* __restrict__ gate) {
  if (threadIdx.x == 0 && *gate) alpha[j] += h[j];
}

// ---------------------------------------------------------------------------
// Fused forms of the per-step vector kernels (pb200_lanczos_post_pro): the same arithmetic in the same order as the
// separate kernels above -- every block re-derives a scalar from the per-block partial sums exactly as
// reduce_partials_kernel<<<1, 256>>> does -- so that a step is 7 launches after the SpMV instead of 14.
// ---------------------------------------------------------------------------
// alpha_j = sum of pa[] (the partial sums of v_j . w);  w -= beta_j v_{j-1} + alpha_j v_j;  pb[block] = partial |w|^2
__global__ void __launch_bounds__(VB) lanczos_recur_kernel(const double* __restrict__ vj, const double* __restrict__ vjm1,
                                                           const double* pa, int nblk, double* alpha,
                                                           const double* __restrict__ beta, int j, double* w,
                                                           long long n, double* pb) {
  __shared__ double red[VB / 32];
  const double a = block_sum_partials(pa, nblk, red);
  const double bj = j > 0 ? beta[j] : 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0) alpha[j] = a;
  const long long base = (long long)blockIdx.x * VCHUNK;
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const long long i = base + threadIdx.x + (long long)t * VB;
    if (i < n) {
      double acc = 0.0;
      if (vjm1) acc = fma(bj, vjm1[i], acc);
      acc = fma(a, vj[i], acc);
      const double wn = w[i] - acc;
      w[i] = wn;
      s = fma(wn, wn, s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < VB / 32; ++q) tot += red[q];
    pb[blockIdx.x] = tot;
  }
}

// gated: w -= sum_c h[c] V[c], then pb[block] = partial |w|^2 of the result
__global__ void __launch_bounds__(VB) multi_axpy_nrm2_kernel(const double* __restrict__ V, long long ldv, int m,
                                                             const double* __restrict__ h, double* w, long long n,
                                                             double* pb, const int* __restrict__ gate) {
  extern __shared__ double hs[];
  __shared__ double red[VB / 32];
  if (*gate == 0) return;
  for (int c = threadIdx.x; c < m; c += VB) hs[c] = h[c];
  __syncthreads();
  const long long base = (long long)blockIdx.x * VCHUNK;
  double acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.0;
  for (int c = 0; c < m; ++c) {
    const double* v = V + (long long)c * ldv;
    const double hc = hs[c];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const long long i = base + threadIdx.x + (long long)t * VB;
      if (i < n) acc[t] = fma(hc, v[i], acc[t]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const long long i = base + threadIdx.x + (long long)t * VB;
    if (i < n) {
      const double wn = w[i] - acc[t];
      w[i] = wn;
      s = fma(wn, wn, s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < VB / 32; ++q) tot += red[q];
    pb[blockIdx.x] = tot;
  }
}

// v_{j+1} = w / beta_{j+1}, beta_{j+1} = sqrt(|w|^2): |w|^2 from tmp (the omega kernel's) or, when the gate is up (w was
// re-orthogonalised), from the partial sums pb[] of multi_axpy_nrm2_kernel
__global__ void __launch_bounds__(VB) lanczos_scale_kernel(const double* __restrict__ w, double* tmp, const double* pb,
                                                           int nblk, const int* __restrict__ gate,
                                                           double* __restrict__ out, long long n, double* beta_out) {
  __shared__ double red[VB / 32];
  double nrm2;
  if (*gate) {
    nrm2 = block_sum_partials(pb, nblk, red);
  } else {
    nrm2 = *tmp;
  }
  const double b = sqrt(nrm2);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *beta_out = b;
  if (i < n) out[i] = b > 0.0 ? w[i] / b : 0.0;
}

// Arnoldi bookkeeping: H[0..j][j] = h + h2 (two Gram-Schmidt passes), H[j+1][j] = sqrt(nrm2); H column-major, ldh rows
__global__ void arnoldi_store_kernel(const double* __restrict__ h, const double* __restrict__ h2, int j,
                                     const double* __restrict__ nrm2, double* __restrict__ H, int ldh) {
  for (int c = threadIdx.x; c <= j; c += blockDim.x) H[(long long)j * ldh + c] = h[c] + h2[c];
  if (threadIdx.x == 0) H[(long long)j * ldh + j + 1] = sqrt(*nrm2);
}

__global__ void vec_mul_kernel(const double* __restrict__ a, const double* __restrict__ b,
                               double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}

// U[i][c] = scale[i] * sum_j V[j][i] * Y[j][c],   c < kk <= 16, Y row-major m x kk
template <int KK>
__global__ void __launch_bounds__(256) ritz_combine_kernel(const double* __restrict__ V, long long ldv, int m,
                                                           const double* __restrict__ Y,
                                                           const double* __restrict__ scale,
                                                           double* __restrict__ U, long long n, int kk) {
  extern __shared__ double ys[];  // m * kk
  for (int e = threadIdx.x; e < m * kk; e += 256) ys[e] = Y[e];
  __syncthreads();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  double acc[KK];
#pragma unroll
  for (int c = 0; c < KK; ++c) acc[c] = 0.0;
  for (int j = 0; j < m; ++j) {
    const double v = V[(long long)j * ldv + i];
#pragma unroll
    for (int c = 0; c < KK; ++c)
      if (c < kk) acc[c] = fma(v, ys[j * kk + c], acc[c]);
  }
  const double sc = scale ? scale[i] : 1.0;
#pragma unroll
  for (int c = 0; c < KK; ++c)
    if (c < kk) U[i * kk + c] = sc * acc[c];
}

// partial[b][c] = sum_{i in block b} A[i][c]^2
__global__ void __launch_bounds__(256) col_sumsq_kernel(const double* __restrict__ A, long long n, int kk,
                                                        double* __restrict__ partial) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) s[c] = 0.0;
  const long long i0 = (long long)blockIdx.x * 4096;
  for (int t = 0; t < 16; ++t) {
    const long long i = i0 + threadIdx.x + t * 256;
    if (i < n)
      for (int c = 0; c < kk; ++c) { const double v = A[i * kk + c]; s[c] = fma(v, v, s[c]); }
  }
  for (int c = 0; c < kk; ++c) {
    const double t = warp_sum(s[c]);
    if (lane == 0) red[warp][c] = t;
  }
  __syncthreads();
  if (threadIdx.x < kk) {
    double tot = 0.0;
    for (int q = 0; q < 8; ++q) tot += red[q][threadIdx.x];
    partial[(long long)threadIdx.x * gridDim.x + blockIdx.x] = tot;
  }
}

__global__ void col_scale_kernel(double* __restrict__ A, long long n, int kk, const double* __restrict__ sumsq) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * kk) return;
  const int c = (int)(e % kk);
  A[e] = A[e] / sqrt(sumsq[c]);
}

}  // namespace pb200

using namespace pb200;

static int launch_spmv_window(const int* indptr, const int* indices, const double* vals, const double* x, double* y,
                              long long n, cudaStream_t st) {
  static bool attr_set = false;
  const int smem = SW_C * (int)sizeof(double);
  if (!attr_set) {
    PB_CHECK(cudaFuncSetAttribute(csr_spmv_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  // contiguous row ranges, a multiple of the block height, one CTA per SM
  long long rows_per_cta = (n + kSMs - 1) / kSMs;
  rows_per_cta = (rows_per_cta + SW_R - 1) / SW_R * SW_R;
  const int grid = (int)((n + rows_per_cta - 1) / rows_per_cta);
  csr_spmv_window_kernel<<<grid, SW_THREADS, smem, st>>>(indptr, indices, vals, x, y, n, rows_per_cta);
  PB_CHECK_LAUNCH("csr_spmv_window_kernel");
  return 0;
}

// y[row0..row1) = (A x)[row0..row1).  long_rows (optional): the rows above SPMV_LONG_ROW entries INSIDE that range.
static int launch_spmv_rows(const int* indptr, const int* indices, const double* vals, const double* x, double* y,
                            long long n, long long nnz, long long row0, long long row1, cudaStream_t st,
                            const int* long_rows, int n_long) {
  const double mean = n > 0 ? (double)nnz / (double)n : 0.0;
  const int skip = (long_rows != nullptr && n_long > 0) ? 1 : 0;
  const long long nr = row1 - row0;
  if (nr <= 0) return 0;
  if (skip) {
    // the long rows first: they are the longest-running warps
    csr_spmv_long_rows_kernel<<<div_up((long long)n_long * 32, 256), 256, 0, st>>>(indptr, indices, vals, x, y, long_rows,
                                                                                 n_long);
    PB_CHECK_LAUNCH("csr_spmv_long_rows_kernel");
  }
  if (mean > 24.0) {
    csr_spmv_kernel<16><<<div_up(nr * 16, 256), 256, 0, st>>>(indptr, indices, vals, x, y, row0, row1, skip);
  } else if (mean > 12.0) {
    csr_spmv_kernel<8><<<div_up(nr * 8, 256), 256, 0, st>>>(indptr, indices, vals, x, y, row0, row1, skip);
  } else {
    csr_spmv_kernel<4><<<div_up(nr * 4, 256), 256, 0, st>>>(indptr, indices, vals, x, y, row0, row1, skip);
  }
  PB_CHECK_LAUNCH("csr_spmv_kernel");
  return 0;
}

static int launch_spmv(const int* indptr, const int* indices, const double* vals, const double* x, double* y,
                       long long n, long long nnz, cudaStream_t st, int windowed = 0, const int* long_rows = nullptr,
                       int n_long = 0) {
  if (windowed && n >= 4 * SW_C) return launch_spmv_window(indptr, indices, vals, x, y, n, st);
  return launch_spmv_rows(indptr, indices, vals, x, y, n, nnz, 0, n, st, long_rows, n_long);
}

static int nrm2_into(const double* w, long long n, double* partial, double* out, cudaStream_t st) {
  const int nblk = div_up(n, VCHUNK);
  multi_dot_kernel<double><<<nblk, VB, 0, st>>>(w, 0, 1, w, n, partial);
  PB_CHECK_LAUNCH("multi_dot_kernel(nrm2)");
  reduce_partials_kernel<<<1, 256, 0, st>>>(partial, nblk, 1, out, nullptr);
  PB_CHECK_LAUNCH("reduce_partials_kernel(nrm2)");
  return 0;
}

extern "C" {

int pb200_csr_spmv_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                       const double* x, double* y, void* stream) {
  return launch_spmv(indptr, indices, vals, x, y, n, nnz, (cudaStream_t)stream);
}

// y = A x for a matrix whose rows are in a locality-preserving order (see csr_spmv_window_kernel);
// same result as pb200_csr_spmv_f64 up to the summation order inside a row.
int pb200_csr_spmv_window_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                              const double* x, double* y, void* stream) {
  return launch_spmv(indptr, indices, vals, x, y, n, nnz, (cudaStream_t)stream, 1);
}

// Relabelled copy of a CSR matrix: row r' = row perm[r'], columns mapped through inv; indptr_new given.
int pb200_csr_permute(const int* indptr, const int* indices, const double* vals, long long n, const int* perm,
                      const int* inv, const int* indptr_new, int* indices_new, double* vals_new, void* stream) {
  if (n <= 0) return 0;
  csr_permute_kernel<<<div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(indptr, indices, vals, n, perm, inv, indptr_new,
                                                                   indices_new, vals_new);
  PB_CHECK_LAUNCH("csr_permute_kernel");
  return 0;
}

// y = A x with the rows of more than pb200_spmv_long_row() entries listed in long_rows (int32 [n_long]): those are
// walked by a whole warp each, the others by the kernel of pb200_csr_spmv_f64.
int pb200_spmv_long_row(void) { return SPMV_LONG_ROW; }
int pb200_csr_spmv_split_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                             const int* long_rows, int n_long, const double* x, double* y, void* stream) {
  return launch_spmv(indptr, indices, vals, x, y, n, nnz, (cudaStream_t)stream, 0, long_rows, n_long);
}

// Rows [row0, row1) of y = A x (a rank's slice when the eigensolver is sharded by rows; y is indexed by the global
// row).  long_rows: the rows above pb200_spmv_long_row() entries inside the range.
int pb200_csr_spmv_rows_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                            long long row0, long long row1, const int* long_rows, int n_long, const double* x,
                            double* y, void* stream) {
  if (row0 < 0 || row1 > n || row0 > row1) { set_error_msg("pb200_csr_spmv_rows_f64: need 0 <= row0 <= row1 <= n"); return -1; }
  return launch_spmv_rows(indptr, indices, vals, x, y, n, nnz, row0, row1, (cudaStream_t)stream, long_rows, n_long);
}

// CSR-stream SpMV (csr_spmv_stream_kernel): nonzeros per tile; tiles of a matrix = ceil(nnz / tile).
int pb200_csr_spmv_stream_tile(void) { return ST_TILE; }

// tile_lo[0 .. n_tiles] for pb200_csr_spmv_stream_f64 (once per matrix; n_tiles = ceil(nnz / tile), n_tiles + 1 ints).
int pb200_csr_spmv_stream_prepare(const int* indptr, long long n, long long nnz, int* tile_lo, void* stream) {
  const int n_tiles = (int)div_up(nnz, ST_TILE);
  csr_stream_tiles_kernel<<<div_up(n_tiles + 1, 256), 256, 0, (cudaStream_t)stream>>>(indptr, n, nnz, n_tiles, tile_lo);
  PB_CHECK_LAUNCH("csr_stream_tiles_kernel");
  return 0;
}

static int launch_spmv_stream(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                              const int* tile_lo, double* head, double* tail, const double* x, double* y,
                              long long row0, long long row1, int tile0, int tile1, cudaStream_t st) {
  if (tile1 <= tile0) return 0;
  csr_spmv_stream_kernel<<<tile1 - tile0, ST_THREADS, 0, st>>>(indptr, indices, vals, x, y, nnz, tile_lo, head, tail,
                                                              tile0, row0, row1);
  PB_CHECK_LAUNCH("csr_spmv_stream_kernel");
  csr_stream_fixup_kernel<<<div_up(tile1 - tile0, 256), 256, 0, st>>>(indptr, nnz, tile_lo, head, tail, y, tile0, tile1,
                                                                     row0, row1);
  PB_CHECK_LAUNCH("csr_stream_fixup_kernel");
  return 0;
}

// y[row0..row1) = (A x)[row0..row1) over the nonzero stream: tiles [tile0, tile1) must cover the nonzeros of those rows
// (tile0 = indptr[row0] / tile, tile1 = ceil(indptr[row1] / tile); the whole matrix: rows [0, n), tiles [0, n_tiles)).
// head / tail: n_tiles doubles of scratch each.  Rows without entries are not written (zero y first if there are any).
int pb200_csr_spmv_stream_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                              const int* tile_lo, double* head, double* tail, const double* x, double* y,
                              long long row0, long long row1, int tile0, int tile1, void* stream) {
  const int n_tiles = (int)div_up(nnz, ST_TILE);
  if (row0 < 0 || row1 > n || row0 > row1 || tile0 < 0 || tile1 > n_tiles) {
    set_error_msg("pb200_csr_spmv_stream_f64: bad row or tile range");
    return -1;
  }
  return launch_spmv_stream(indptr, indices, vals, n, nnz, tile_lo, head, tail, x, y, row0, row1, tile0, tile1,
                            (cudaStream_t)stream);
}

int pb200_csr_spmm_f32(const int* indptr, const int* indices, const double* vals, long long n, const float* x,
                       long long ldx, float* y, long long ldy, int g, void* stream) {
  csr_spmm_f32_kernel<<<div_up(n, 8), 256, 0, (cudaStream_t)stream>>>(indptr, indices, vals, x, y, n, g, ldx, ldy);
  PB_CHECK_LAUNCH("csr_spmm_f32_kernel");
  return 0;
}

// Gene-tiled SpMM (MAGIC at scale): y[n][ldy] = A x[n][ldx] for the first g columns; vals32: the CSR values in FP32
// (pb200_cast_f64_f32); x / y FP32 (is_f64 = 0) or FP64; ldx, ldy multiples of 4 (FP32) / 2 (FP64) elements and
// both bases 16-byte aligned.  One launch per tile of 256 (FP32) / 128 (FP64) columns.
int pb200_csr_spmm_tiled(const int* indptr, const int* indices, const float* vals32, long long n, const void* x,
                         long long ldx, void* y, long long ldy, int g, int is_f64, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0 || g <= 0) return 0;
  const int vec = is_f64 ? 2 : 4;
  if (ldx % vec != 0 || ldy % vec != 0 || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) {
    set_error_msg("pb200_csr_spmm_tiled: leading dimensions / bases must be 16-byte aligned");
    return -1;
  }
  const int gt = 64 * vec;
  for (int g0 = 0; g0 < g; g0 += gt) {
    if (is_f64)
      csr_spmm_tile_kernel<double, 2><<<div_up(n, 8), 256, 0, st>>>(indptr, indices, vals32, (const double*)x, (double*)y,
                                                                   n, g, ldx, ldy, g0);
    else
      csr_spmm_tile_kernel<float, 4><<<div_up(n, 8), 256, 0, st>>>(indptr, indices, vals32, (const float*)x, (float*)y, n,
                                                                  g, ldx, ldy, g0);
    PB_CHECK_LAUNCH("csr_spmm_tile_kernel");
  }
  return 0;
}

int pb200_cast_f64_f32(const double* in, float* out, long long n, void* stream) {
  if (n <= 0) return 0;
  cast_f64_f32_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n);
  PB_CHECK_LAUNCH("cast_f64_f32_kernel");
  return 0;
}

// a[i] = 0 where a[i] < thr (count elements, FP32 or FP64)
int pb200_clip_below(void* a, long long count, double thr, int is_f64, void* stream) {
  if (count <= 0) return 0;
  if (is_f64) clip_below_kernel<double><<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>((double*)a, count, thr);
  else        clip_below_kernel<float><<<div_up(count, 256), 256, 0, (cudaStream_t)stream>>>((float*)a, count, (float)thr);
  PB_CHECK_LAUNCH("clip_below_kernel");
  return 0;
}

int pb200_vec_mul(const double* a, const double* b, double* out, long long n, void* stream) {
  vec_mul_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  PB_CHECK_LAUNCH("vec_mul_kernel");
  return 0;
}

// V[0] = w / |w| (and its FP32 shadow V32[0] when given).   scratch: partial[div_up(n,2048)], tmp[1]
int pb200_lanczos_init(const double* w, long long n, double* V0, float* V32_0, double* partial, double* tmp,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc = nrm2_into(w, n, partial, tmp, st);
  if (rc) return rc;
  scale_by_norm_kernel<<<div_up(n, 256), 256, 0, st>>>(w, tmp, V0, V32_0, n, nullptr);
  PB_CHECK_LAUNCH("scale_by_norm_kernel");
  return 0;
}

// Runs Lanczos steps j = j0 .. j0+nsteps-1 on the symmetric CSR matrix.
// In: V[0..j0] (rows of length ldv >= n) orthonormal.  Out: alpha[j], beta[j+1], V[j+1].
// V32 (optional, same leading dimension): FP32 shadow of the basis, kept up to date here and used for
// the Gram-Schmidt pass -- what that pass removes are round-off-level components (|h| <~ 1e-13 after
// the three-term recurrence), so the 6e-8 relative error of the shadow changes w by < 1e-20 while the
// pass reads half the bytes.  The recurrence itself, alpha, beta and the Ritz vectors use the FP64 basis.
// long_rows / n_long (optional): rows above pb200_spmv_long_row() entries, walked by a warp each.
// windowed != 0: the matrix rows are in a locality-preserving order; use the shared-memory-window SpMV.
// Scratch: w[n], h[mcap], partial[mcap * div_up(n,2048)], tmp[1].
int pb200_lanczos_steps(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                        const int* long_rows, int n_long, double* V, long long ldv, float* V32, int windowed, int j0,
                        int nsteps, double* alpha, double* beta, double* w, double* h, double* partial, double* tmp,
                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = div_up(n, VCHUNK);
  for (int j = j0; j < j0 + nsteps; ++j) {
    const int m = j + 1;
    int rc = launch_spmv(indptr, indices, vals, V + (long long)j * ldv, w, n, nnz, st, windowed, long_rows, n_long);
    if (rc) return rc;
    // three-term recurrence: alpha_j = v_j . w;  w -= alpha_j v_j + beta_j v_{j-1}
    multi_dot_kernel<double><<<nblk, VB, 0, st>>>(V + (long long)j * ldv, ldv, 1, w, n, partial);
    PB_CHECK_LAUNCH("multi_dot_kernel(alpha)");
    reduce_partials_kernel<<<1, 256, 0, st>>>(partial, nblk, 1, alpha + j, nullptr);
    PB_CHECK_LAUNCH("reduce_partials_kernel(alpha)");
    recurrence_coeffs_kernel<<<1, 32, 0, st>>>(alpha, beta, j, h);
    PB_CHECK_LAUNCH("recurrence_coeffs_kernel");
    if (j > 0)
      multi_axpy_kernel<double><<<nblk, VB, sizeof(double) * 2, st>>>(V + (long long)(j - 1) * ldv, ldv, 2, h, w, n);
    else
      multi_axpy_kernel<double><<<nblk, VB, sizeof(double) * 1, st>>>(V, ldv, 1, h + 1, w, n);
    PB_CHECK_LAUNCH("multi_axpy_kernel(recurrence)");
    // one full Gram-Schmidt pass against V[0..j]; its v_j component corrects alpha_j
    if (V32) multi_dot_kernel<float><<<nblk, VB, 0, st>>>(V32, ldv, m, w, n, partial);
    else     multi_dot_kernel<double><<<nblk, VB, 0, st>>>(V, ldv, m, w, n, partial);
    PB_CHECK_LAUNCH("multi_dot_kernel");
    reduce_partials_kernel<<<m, 256, 0, st>>>(partial, nblk, m, h, nullptr);
    PB_CHECK_LAUNCH("reduce_partials_kernel");
    reduce_partials_kernel<<<1, 32, 0, st>>>(h + j, 1, 1, tmp, alpha + j);
    PB_CHECK_LAUNCH("reduce_partials_kernel(alpha)");
    if (V32) multi_axpy_kernel<float><<<nblk, VB, sizeof(double) * m, st>>>(V32, ldv, m, h, w, n);
    else     multi_axpy_kernel<double><<<nblk, VB, sizeof(double) * m, st>>>(V, ldv, m, h, w, n);
    PB_CHECK_LAUNCH("multi_axpy_kernel");
    rc = nrm2_into(w, n, partial, tmp, st);
    if (rc) return rc;
    scale_by_norm_kernel<<<div_up(n, 256), 256, 0, st>>>(w, tmp, V + (long long)(j + 1) * ldv,
                                                         V32 ? V32 + (long long)(j + 1) * ldv : nullptr, n, beta + j + 1);
    PB_CHECK_LAUNCH("scale_by_norm_kernel");
  }
  return 0;
}

// Everything of Lanczos step j that follows the product w = S v_j, with partial reorthogonalisation (see
// pb200_lanczos_steps_pro): alpha_j, the three-term recurrence, |w|^2, the omega recurrence and its gate, the gated
// Gram-Schmidt pass over V[0..j], v_{j+1} and beta_{j+1}.  Seven launches (fused kernels, same arithmetic as the
// separate ones of pb200_lanczos_steps).
static int lanczos_post_pro(long long n, double* V, long long ldv, int j, double* alpha, double* beta, double* w,
                            double* h, double* partial, double* tmp, double* om, long long omstride, int* gate,
                            int* counters, double eps, double thresh, cudaStream_t st) {
  const int nblk = div_up(n, VCHUNK);
  const int m = j + 1;
  double* pa = partial;             // partial sums of v_j . w
  double* pb = partial + nblk;      // partial sums of |w|^2 (needs partial[>= 2 nblk]: mcap >= 2)
  const double* vj = V + (long long)j * ldv;
  multi_dot_kernel<double><<<nblk, VB, 0, st>>>(vj, ldv, 1, w, n, pa);
  PB_CHECK_LAUNCH("multi_dot_kernel(alpha)");
  lanczos_recur_kernel<<<nblk, VB, 0, st>>>(vj, j > 0 ? vj - ldv : nullptr, pa, nblk, alpha, beta, j, w, n, pb);
  PB_CHECK_LAUNCH("lanczos_recur_kernel");
  lanczos_omega_kernel<<<1, 256, 0, st>>>(alpha, beta, j, tmp, om, omstride, eps, thresh, gate, counters, pb, nblk);
  PB_CHECK_LAUNCH("lanczos_omega_kernel");
  // gated: one Gram-Schmidt pass against V[0..j], alpha correction, new norm.  The FP64 basis, not an FP32
  // shadow: here the pass has to bring the overlaps down from ~1e-8 to round-off (the estimates are reset to
  // eps afterwards), and a shadow vector's 6e-8 rounding error, dotted with a w of norm beta, puts ~1e-9 back
  multi_dot_kernel<double><<<nblk, VB, 0, st>>>(V, ldv, m, w, n, partial, gate);
  PB_CHECK_LAUNCH("multi_dot_kernel(gated)");
  reduce_partials_kernel<<<m, 256, 0, st>>>(partial, nblk, m, h, nullptr, gate, j, alpha);
  PB_CHECK_LAUNCH("reduce_partials_kernel(gated)");
  multi_axpy_nrm2_kernel<<<nblk, VB, sizeof(double) * m, st>>>(V, ldv, m, h, w, n, pb, gate);
  PB_CHECK_LAUNCH("multi_axpy_nrm2_kernel(gated)");
  lanczos_scale_kernel<<<div_up(n, VB), VB, 0, st>>>(w, tmp, pb, nblk, gate, V + (long long)(j + 1) * ldv, n,
                                                    beta + j + 1);
  PB_CHECK_LAUNCH("lanczos_scale_kernel");
  return 0;
}

// The same recurrence with PARTIAL reorthogonalisation: the Gram-Schmidt pass over the whole basis runs only when
// the omega recurrence (lanczos_omega_kernel) says that orthogonality is about to drop below `thresh`
// (~sqrt(machine eps)), on two consecutive steps each time; in between only the three-term recurrence runs.
// Semi-orthogonality at that level keeps the Ritz values accurate to round-off and the Ritz vectors to ~1e-8
// (Simon 1984).  om: 3 * (mcap + 2) doubles, om[0] = 1 before the first step; gate: 2 ints, zero before the first step;
// counters[0] counts the reorthogonalised steps.  V32 is not used (kept for the ABI).  Everything else as
// pb200_lanczos_steps.
int pb200_lanczos_steps_pro(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                            const int* long_rows, int n_long, double* V, long long ldv, float* V32, int windowed,
                            int j0, int nsteps, double* alpha, double* beta, double* w, double* h, double* partial,
                            double* tmp, double* om, long long omstride, int* gate, int* counters, double eps,
                            double thresh, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  (void)V32;
  for (int j = j0; j < j0 + nsteps; ++j) {
    int rc = launch_spmv(indptr, indices, vals, V + (long long)j * ldv, w, n, nnz, st, windowed, long_rows, n_long);
    if (rc) return rc;
    rc = lanczos_post_pro(n, V, ldv, j, alpha, beta, w, h, partial, tmp, om, omstride, gate, counters, eps, thresh, st);
    if (rc) return rc;
  }
  return 0;
}

// Step j of pb200_lanczos_steps_pro WITHOUT its product: the caller has put w = S v_j (all n entries) in place --
// the eigensolver sharded by rows computes its slice with pb200_csr_spmv_rows_f64 and all-gathers w in between.
int pb200_lanczos_post_pro(long long n, double* V, long long ldv, int j, double* alpha, double* beta, double* w,
                           double* h, double* partial, double* tmp, double* om, long long omstride, int* gate,
                           int* counters, double eps, double thresh, void* stream) {
  return lanczos_post_pro(n, V, ldv, j, alpha, beta, w, h, partial, tmp, om, omstride, gate, counters, eps, thresh,
                          (cudaStream_t)stream);
}

// Arnoldi steps j = j0 .. j0+nsteps-1 for a general (non-symmetric) CSR matrix, e.g. T = D^-1 K of a kernel that
// is not symmetric (utils.py:477-484 accepts any kernel): w = A v_j, two classical Gram-Schmidt passes against
// V[0..j], H[:, j] (column-major, ldh >= mcap + 1 rows), v_{j+1} = w / |w|.  Scratch: w[n], h[mcap], h2[mcap],
// partial[mcap * div_up(n, 2048)], tmp[1].
int pb200_arnoldi_steps(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                        double* V, long long ldv, int j0, int nsteps, double* H, int ldh, double* w, double* h,
                        double* h2, double* partial, double* tmp, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = div_up(n, VCHUNK);
  for (int j = j0; j < j0 + nsteps; ++j) {
    const int m = j + 1;
    int rc = launch_spmv(indptr, indices, vals, V + (long long)j * ldv, w, n, nnz, st);
    if (rc) return rc;
    for (int pass = 0; pass < 2; ++pass) {
      double* hh = pass ? h2 : h;
      multi_dot_kernel<double><<<nblk, VB, 0, st>>>(V, ldv, m, w, n, partial);
      PB_CHECK_LAUNCH("multi_dot_kernel(arnoldi)");
      reduce_partials_kernel<<<m, 256, 0, st>>>(partial, nblk, m, hh, nullptr);
      PB_CHECK_LAUNCH("reduce_partials_kernel(arnoldi)");
      multi_axpy_kernel<double><<<nblk, VB, sizeof(double) * m, st>>>(V, ldv, m, hh, w, n);
      PB_CHECK_LAUNCH("multi_axpy_kernel(arnoldi)");
    }
    rc = nrm2_into(w, n, partial, tmp, st);
    if (rc) return rc;
    arnoldi_store_kernel<<<1, 256, 0, st>>>(h, h2, j, tmp, H, ldh);
    PB_CHECK_LAUNCH("arnoldi_store_kernel");
    scale_by_norm_kernel<<<div_up(n, 256), 256, 0, st>>>(w, tmp, V + (long long)(j + 1) * ldv, nullptr, n, nullptr);
    PB_CHECK_LAUNCH("scale_by_norm_kernel");
  }
  return 0;
}

// U[n][kk] = diag(scale) * V[0..m-1]^T-combination with Y (m x kk, row-major, device).
int pb200_ritz_vectors(const double* V, long long ldv, int m, const double* Y, int kk, const double* scale,
                       double* U, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kk < 1 || kk > 32) { set_error_msg("pb200_ritz_vectors: need 1 <= kk <= 32"); return -1; }
  const size_t smem = sizeof(double) * (size_t)m * kk;
  if (smem > 200 * 1024) { set_error_msg("pb200_ritz_vectors: m*kk too large for shared memory"); return -1; }
  if (kk <= 16) {
    PB_CHECK(cudaFuncSetAttribute(ritz_combine_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ritz_combine_kernel<16><<<div_up(n, 256), 256, smem, st>>>(V, ldv, m, Y, scale, U, n, kk);
  } else {
    PB_CHECK(cudaFuncSetAttribute(ritz_combine_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ritz_combine_kernel<32><<<div_up(n, 256), 256, smem, st>>>(V, ldv, m, Y, scale, U, n, kk);
  }
  PB_CHECK_LAUNCH("ritz_combine_kernel");
  return 0;
}

// A[n][kk] columns scaled to unit L2 norm.  scratch: partial[kk * div_up(n,4096)], sumsq[kk]
int pb200_normalize_columns(double* A, long long n, int kk, double* partial, double* sumsq, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kk < 1 || kk > 32) { set_error_msg("pb200_normalize_columns: need 1 <= kk <= 32"); return -1; }
  const int nblk = div_up(n, 4096);
  col_sumsq_kernel<<<nblk, 256, 0, st>>>(A, n, kk, partial);
  PB_CHECK_LAUNCH("col_sumsq_kernel");
  reduce_partials_kernel<<<kk, 256, 0, st>>>(partial, nblk, kk, sumsq, nullptr);
  PB_CHECK_LAUNCH("reduce_partials_kernel(cols)");
  col_scale_kernel<<<div_up(n * kk, 256), 256, 0, st>>>(A, n, kk, sumsq);
  PB_CHECK_LAUNCH("col_scale_kernel");
  return 0;
}

}  // extern "C"
