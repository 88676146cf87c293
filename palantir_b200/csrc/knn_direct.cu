// Exact FP64 kNN with direct differences, for low-dimensional spaces (m <= 16).
//
// Replaces sklearn's kd-tree NearestNeighbors at src/palantir/core.py:355-356
// (multiscale space, all cells) and core.py:525-526 (waypoints), and the
// kd-tree branch of utils.py:404-407 when the PCA input has <= 15 columns.
// The kd-tree's leaf distance is rdist = sum_c (x_c - y_c)^2 accumulated in
// dimension order with separate multiply and add, then sqrt; the kernel does the
// same in FP64 with contraction disabled (__dmul_rn / __dadd_rn), so the edge
// weights are bit-equal.  Self is included (distance 0).
// Ranking is by (distance, index).
//
// One thread owns one query row: coordinates in registers, its k-best list in
// local memory (L1-resident; after warm-up an insertion happens ~k ln(N/k) times
// per row), reference points streamed through shared memory in 128-point tiles
// and read as warp-wide broadcasts.
#include "common.cuh"

namespace pb200 {

constexpr int KD_ROWS = 128;  // query rows per CTA (one per thread)
constexpr int KD_TILE = 128;  // reference points per shared-memory tile
constexpr int KD_KMAX = 64;

template <int M>
__global__ void __launch_bounds__(KD_ROWS) knn_direct_kernel(const double* __restrict__ data, long long n,
                                                             const double* __restrict__ query,
                                                             const int* __restrict__ qrows, int q0, int nq,
                                                             int k, int* __restrict__ out_idx,
                                                             double* __restrict__ out_dist) {
  __shared__ double tile[KD_TILE][M];
  const int tid = threadIdx.x;
  const int q = blockIdx.x * KD_ROWS + tid;
  const bool active = q < nq;
  double x[M];
  {
    // query coordinates: either rows q0.. of `query`, or gathered rows qrows[q]
    long long src = active ? (qrows ? (long long)qrows[q] : (long long)q0 + q) : 0;
#pragma unroll
    for (int c = 0; c < M; ++c) x[c] = query[src * M + c];
  }
  double bd[KD_KMAX];
  int bi[KD_KMAX];
  for (int s = 0; s < k; ++s) { bd[s] = INFINITY; bi[s] = -1; }
  double worst = INFINITY;

  for (long long t0 = 0; t0 < n; t0 += KD_TILE) {
    __syncthreads();
    for (int e = tid; e < KD_TILE * M; e += KD_ROWS) {
      const long long g = t0 * M + e;
      (&tile[0][0])[e] = g < n * M ? data[g] : 0.0;
    }
    __syncthreads();
    const int lim = (int)((n - t0) < KD_TILE ? (n - t0) : KD_TILE);
    if (active) {
#pragma unroll 4
      for (int j = 0; j < lim; ++j) {
        double r = 0.0;
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const double df = __dsub_rn(x[c], tile[j][c]);
          r = __dadd_rn(r, __dmul_rn(df, df));
        }
        if (r < worst) {
          // sorted insertion (ascending); ties keep the earlier (lower) index first
          int pos = k - 1;
          while (pos > 0 && bd[pos - 1] > r) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
          }
          bd[pos] = r;
          bi[pos] = (int)(t0 + j);
          worst = bd[k - 1];
        }
      }
    }
  }
  if (active) {
    for (int s = 0; s < k; ++s) {
      out_idx[(long long)q * k + s] = bi[s];
      out_dist[(long long)q * k + s] = sqrt(bd[s]);
    }
  }
}

// Same search on cells SORTED by their first coordinate, with exact pruning: the CTA
// walks the 128-point reference tiles outwards from its own position and stops on a side
// once the squared first-coordinate gap to that tile is >= the largest current k-th
// distance of its 128 rows.  The bound is safe in floating point: rounding is monotone, so
// fl((x0_q - x0_p)^2) >= fl(gap^2) for every point p beyond the tile edge, and adding the
// other (non-negative) terms cannot decrease the rounded sum.  Indices are positions in the
// sorted order.
template <int M>
__global__ void __launch_bounds__(KD_ROWS) knn_sorted_kernel(const double* __restrict__ data, long long n, int q0,
                                                             int nq, int k, int* __restrict__ out_idx,
                                                             double* __restrict__ out_dist,
                                                             unsigned long long* __restrict__ tiles_scanned) {
  __shared__ double tile[KD_TILE][M];
  __shared__ double s_red[KD_ROWS / 32];
  __shared__ double s_maxworst;
  const int tid = threadIdx.x;
  const long long qb0 = (long long)q0 + (long long)blockIdx.x * KD_ROWS;
  const long long q = qb0 + tid;
  const bool active = (long long)blockIdx.x * KD_ROWS + tid < nq;
  double x[M];
  {
    const long long src = active ? q : qb0;
#pragma unroll
    for (int c = 0; c < M; ++c) x[c] = data[src * M + c];
  }
  long long q_last = qb0 + KD_ROWS - 1;
  if (q_last > (long long)q0 + nq - 1) q_last = (long long)q0 + nq - 1;
  const double xq_lo = data[qb0 * M], xq_hi = data[q_last * M];
  double bd[KD_KMAX];
  int bi[KD_KMAX];
  for (int s = 0; s < k; ++s) { bd[s] = INFINITY; bi[s] = -1; }
  double worst = INFINITY;
  if (tid == 0) s_maxworst = INFINITY;
  __syncthreads();
  const long long ntiles = (n + KD_TILE - 1) / KD_TILE;
  long long right = qb0 / KD_TILE, left = right - 1;
  bool r_done = false, l_done = false;
  unsigned long long scanned = 0;
  int turn = 0;
  while (!(r_done && l_done)) {
    const bool go_right = (turn == 0 && !r_done) || l_done;
    turn ^= 1;
    long long t;
    if (go_right) {
      if (right >= ntiles) { r_done = true; continue; }
      double gap = data[right * KD_TILE * M] - xq_hi;
      gap = gap > 0.0 ? gap : 0.0;
      if (__dmul_rn(gap, gap) >= s_maxworst) { r_done = true; continue; }
      t = right++;
    } else {
      if (left < 0) { l_done = true; continue; }
      long long lastp = left * KD_TILE + KD_TILE - 1;
      if (lastp > n - 1) lastp = n - 1;
      double gap = xq_lo - data[lastp * M];
      gap = gap > 0.0 ? gap : 0.0;
      if (__dmul_rn(gap, gap) >= s_maxworst) { l_done = true; continue; }
      t = left--;
    }
    const long long t0 = t * KD_TILE;
    __syncthreads();
    for (int e = tid; e < KD_TILE * M; e += KD_ROWS) {
      const long long g = t0 * M + e;
      (&tile[0][0])[e] = g < n * M ? data[g] : 0.0;
    }
    __syncthreads();
    ++scanned;
    const int lim = (int)((n - t0) < KD_TILE ? (n - t0) : KD_TILE);
    if (active) {
#pragma unroll 4
      for (int j = 0; j < lim; ++j) {
        double r = 0.0;
#pragma unroll
        for (int c = 0; c < M; ++c) {
          const double df = __dsub_rn(x[c], tile[j][c]);
          r = __dadd_rn(r, __dmul_rn(df, df));
        }
        if (r < worst) {
          int pos = k - 1;
          while (pos > 0 && bd[pos - 1] > r) {
            bd[pos] = bd[pos - 1];
            bi[pos] = bi[pos - 1];
            --pos;
          }
          bd[pos] = r;
          bi[pos] = (int)(t0 + j);
          worst = bd[k - 1];
        }
      }
    }
    double w = active ? worst : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = w;
    __syncthreads();
    if (tid == 0) {
      double mw = s_red[0];
      for (int qq = 1; qq < KD_ROWS / 32; ++qq) mw = fmax(mw, s_red[qq]);
      s_maxworst = mw;
    }
    __syncthreads();
  }
  if (active) {
    const long long row = (long long)blockIdx.x * KD_ROWS + tid;
    for (int s = 0; s < k; ++s) {
      out_idx[row * k + s] = bi[s];
      out_dist[row * k + s] = sqrt(bd[s]);
    }
  }
  if (tid == 0 && tiles_scanned) atomicAdd(tiles_scanned, scanned);
}


// Per-tile bounding boxes (all m coordinates) of 128-point tiles of row-major data.
__global__ void __launch_bounds__(128) tile_boxes_md_kernel(const double* __restrict__ data, long long n, int m,
                                                            double* __restrict__ box_lo, double* __restrict__ box_hi) {
  const long long t = blockIdx.x;
  const long long i = t * KD_TILE + threadIdx.x;
  __shared__ double s_lo[4][16], s_hi[4][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < m; ++c) {
    double lo = INFINITY, hi = -INFINITY;
    if (i < n) lo = hi = data[i * m + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { s_lo[warp][c] = lo; s_hi[warp][c] = hi; }
  }
  __syncthreads();
  if (threadIdx.x < m) {
    const int c = threadIdx.x;
    box_lo[t * m + c] = fmin(fmin(s_lo[0][c], s_lo[1][c]), fmin(s_lo[2][c], s_lo[3][c]));
    box_hi[t * m + c] = fmax(fmax(s_hi[0][c], s_hi[1][c]), fmax(s_hi[2][c], s_hi[3][c]));
  }
}

// outward sequence of tiles around tile q: q, q-1, q+1, q-2, ... then the remaining side
__device__ __forceinline__ long long kd_tile_of(long long s, long long q, long long T) {
  const long long R = T - q, L = q;
  const long long mn = R < L ? R : L;
  if (s < 2 * mn) return (s & 1) ? q - 1 - (s >> 1) : q + (s >> 1);
  const long long rest = s - 2 * mn;
  return R > L ? q + mn + rest : q - 1 - mn - rest;
}

// Same search on cells in ANY locality-preserving order (here: the Morton curve of the leading
// coordinates) with exact pruning by bounding boxes: a 128-point reference tile is skipped when the
// squared gap between its box and the box of the CTA's 128 query rows is >= the largest current k-th
// squared distance of those rows, and a single row skips a tile whose box is no closer than its own
// k-th distance.  The gap is accumulated exactly like a distance (coordinate order, separate
// multiply and add, round-to-nearest), and every operation in that chain is monotone, so the
// computed gap never exceeds the computed distance to any point of the box: nothing is lost.
template <int M>
__global__ void __launch_bounds__(KD_ROWS) knn_boxed_kernel(const double* __restrict__ data, long long n, int q0,
                                                            int nq, int k, const double* __restrict__ box_lo,
                                                            const double* __restrict__ box_hi,
                                                            int* __restrict__ out_idx, double* __restrict__ out_dist,
                                                            unsigned long long* __restrict__ tiles_scanned) {
  __shared__ double tile[KD_TILE][M];
  __shared__ double s_qlo[M], s_qhi[M], s_tlo[M], s_thi[M];
  __shared__ double s_red[KD_ROWS / 32];
  __shared__ double s_maxworst;
  __shared__ int s_acc[KD_ROWS];
  __shared__ int s_wcnt[KD_ROWS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long qb0 = (long long)q0 + (long long)blockIdx.x * KD_ROWS;
  const long long q = qb0 + tid;
  const bool active = (long long)blockIdx.x * KD_ROWS + tid < nq;
  const long long ntiles = (n + KD_TILE - 1) / KD_TILE;
  const long long qtile = qb0 / KD_TILE;
  double x[M];
  {
    const long long src = active ? q : qb0;
#pragma unroll
    for (int c = 0; c < M; ++c) x[c] = data[src * M + c];
  }
  if (tid < M) { s_qlo[tid] = box_lo[qtile * M + tid]; s_qhi[tid] = box_hi[qtile * M + tid]; }
  // the k best so far, UNSORTED: a new candidate overwrites the current worst entry and the new worst is found
  // by one scan of the k entries (independent loads) -- keeping the list sorted meant shifting it entry by
  // entry through a chain of dependent local-memory loads, 90 % of the kernel's time; sorted once at the end
  double bd[KD_KMAX];
  int bi[KD_KMAX];
  for (int s = 0; s < k; ++s) { bd[s] = INFINITY; bi[s] = -1; }
  double worst = INFINITY;
  int wpos = 0;
  if (tid == 0) s_maxworst = INFINITY;
  __syncthreads();
  unsigned long long scanned = 0;
  for (long long base = 0; base < ntiles; base += KD_ROWS) {
    // ---- 128 candidate tiles of the outward sequence, one per thread ----
    const long long sidx = base + tid;
    long long cand = -1;
    bool ok = false;
    if (sidx < ntiles) {
      cand = kd_tile_of(sidx, qtile, ntiles);
      double g2 = 0.0;
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const double d1 = __dsub_rn(box_lo[cand * M + c], s_qhi[c]);
        const double d2 = __dsub_rn(s_qlo[c], box_hi[cand * M + c]);
        double gp = d1 > d2 ? d1 : d2;
        gp = gp > 0.0 ? gp : 0.0;
        g2 = __dadd_rn(g2, __dmul_rn(gp, gp));
      }
      ok = g2 < s_maxworst;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < KD_ROWS / 32; ++w) { off += w < warp ? s_wcnt[w] : 0; total += s_wcnt[w]; }
    if (ok) s_acc[off + __popc(bal & ((1u << lane) - 1u))] = (int)cand;
    __syncthreads();
    // ---- scan the accepted tiles (bound re-checked: it tightens as the scan proceeds) ----
    for (int a = 0; a < total; ++a) {
      const long long t = s_acc[a];
      const long long t0 = t * KD_TILE;
      __syncthreads();
      if (tid < M) { s_tlo[tid] = box_lo[t * M + tid]; s_thi[tid] = box_hi[t * M + tid]; }
      __syncthreads();
      double gq = 0.0, gr = 0.0;      // gap of the tile box to the CTA's box and to this row
#pragma unroll
      for (int c = 0; c < M; ++c) {
        const double d1 = __dsub_rn(s_tlo[c], s_qhi[c]);
        const double d2 = __dsub_rn(s_qlo[c], s_thi[c]);
        double gp = d1 > d2 ? d1 : d2;
        gp = gp > 0.0 ? gp : 0.0;
        gq = __dadd_rn(gq, __dmul_rn(gp, gp));
        const double e1 = __dsub_rn(s_tlo[c], x[c]);
        const double e2 = __dsub_rn(x[c], s_thi[c]);
        double ep = e1 > e2 ? e1 : e2;
        ep = ep > 0.0 ? ep : 0.0;
        gr = __dadd_rn(gr, __dmul_rn(ep, ep));
      }
      if (!(gq < s_maxworst)) continue;          // uniform: every thread reads the same values
      for (int e = tid; e < KD_TILE * M; e += KD_ROWS) {
        const long long g = t0 * M + e;
        (&tile[0][0])[e] = g < n * M ? data[g] : 0.0;
      }
      __syncthreads();
      ++scanned;
      const int lim = (int)((n - t0) < KD_TILE ? (n - t0) : KD_TILE);
      if (active && gr < worst) {
        // Two phases per tile.  (A) all 128 distances, the candidates below the current k-th distance only
        // marked in a bit mask: no divergence.  (B) each lane inserts its marked candidates (distance
        // recomputed, re-tested against the k-th distance it has by then): the warp runs as many insertion
        // passes as its busiest lane has candidates, not one pass per candidate of any lane -- with the
        // insertion inside the scan the shifting of the k-entry lists was 90 % of the kernel.
        for (int w0 = 0; w0 < lim; w0 += 32) {
          unsigned int mask = 0;
#pragma unroll 8
          for (int jj = 0; jj < 32; ++jj) {
            const int j = w0 + jj;
            double r = 0.0;
#pragma unroll
            for (int c = 0; c < M; ++c) {
              const double df = __dsub_rn(x[c], tile[j < lim ? j : 0][c]);
              r = __dadd_rn(r, __dmul_rn(df, df));
            }
            mask |= (j < lim && r < worst) ? (1u << jj) : 0u;
          }
          while (mask) {
            const int jj = __ffs(mask) - 1;
            mask &= mask - 1;
            const int j = w0 + jj;
            double r = 0.0;
#pragma unroll
            for (int c = 0; c < M; ++c) {
              const double df = __dsub_rn(x[c], tile[j][c]);
              r = __dadd_rn(r, __dmul_rn(df, df));
            }
            if (r < worst) {
              bd[wpos] = r;
              bi[wpos] = (int)(t0 + j);
              double mx = bd[0];
              int mp = 0;
              for (int s2 = 1; s2 < k; ++s2) {
                const double v2 = bd[s2];
                if (v2 > mx) { mx = v2; mp = s2; }
              }
              worst = mx;
              wpos = mp;
            }
          }
        }
      }
      double w = active ? worst : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
      if (lane == 0) s_red[warp] = w;
      __syncthreads();
      if (tid == 0) {
        double mw = s_red[0];
        for (int qq = 1; qq < KD_ROWS / 32; ++qq) mw = fmax(mw, s_red[qq]);
        s_maxworst = mw;
      }
    }
    __syncthreads();
  }
  if (active) {
    // ascending by (distance, index)
    for (int e = 1; e < k; ++e) {
      const double d = bd[e];
      const int ix = bi[e];
      int f = e - 1;
      while (f >= 0 && (bd[f] > d || (bd[f] == d && bi[f] > ix))) { bd[f + 1] = bd[f]; bi[f + 1] = bi[f]; --f; }
      bd[f + 1] = d; bi[f + 1] = ix;
    }
    const long long row = (long long)blockIdx.x * KD_ROWS + tid;
    for (int s = 0; s < k; ++s) {
      out_idx[row * k + s] = bi[s];
      out_dist[row * k + s] = sqrt(bd[s]);
    }
  }
  if (tid == 0 && tiles_scanned) atomicAdd(tiles_scanned, scanned);
}

template <int M>
static int launch_boxed(const double* data, long long n, int q0, int nq, int k, const double* box_lo,
                        const double* box_hi, int* out_idx, double* out_dist, unsigned long long* tiles_scanned,
                        cudaStream_t st) {
  knn_boxed_kernel<M><<<div_up(nq, KD_ROWS), KD_ROWS, 0, st>>>(data, n, q0, nq, k, box_lo, box_hi, out_idx, out_dist,
                                                             tiles_scanned);
  PB_CHECK_LAUNCH("knn_boxed_kernel");
  return 0;
}

template <int M>
static int launch_sorted(const double* data, long long n, int q0, int nq, int k, int* out_idx, double* out_dist,
                         unsigned long long* tiles_scanned, cudaStream_t st) {
  knn_sorted_kernel<M><<<div_up(nq, KD_ROWS), KD_ROWS, 0, st>>>(data, n, q0, nq, k, out_idx, out_dist, tiles_scanned);
  PB_CHECK_LAUNCH("knn_sorted_kernel");
  return 0;
}

template <int M>
static int launch_direct(const double* data, long long n, const double* query, const int* qrows, int q0, int nq,
                         int k, int* out_idx, double* out_dist, cudaStream_t st) {
  knn_direct_kernel<M><<<div_up(nq, KD_ROWS), KD_ROWS, 0, st>>>(data, n, query, qrows, q0, nq, k, out_idx,
                                                                out_dist);
  PB_CHECK_LAUNCH("knn_direct_kernel");
  return 0;
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// data: n x m row-major FP64 (reference set).  Queries are rows q0..q0+nq-1 of
// `query` (same m), or, when qrows != NULL, rows qrows[0..nq-1] of `query`.
int pb200_knn_direct_f64(const double* data, long long n, int m, const double* query, const int* qrows, int q0,
                         int nq, int k, int* out_idx, double* out_dist, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m < 1 || m > 16 || k < 1 || k > KD_KMAX || k > n) {
    set_error_msg("pb200_knn_direct_f64: need 1 <= m <= 16 and 1 <= k <= min(64, n)");
    return -1;
  }
  if (nq <= 0) return 0;
  switch (m) {
#define PB_CASE(M) case M: return launch_direct<M>(data, n, query, qrows, q0, nq, k, out_idx, out_dist, st);
    PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8)
    PB_CASE(9) PB_CASE(10) PB_CASE(11) PB_CASE(12) PB_CASE(13) PB_CASE(14) PB_CASE(15) PB_CASE(16)
#undef PB_CASE
  }
  return -1;
}

// Same result on data sorted ascending by column 0 (rows q0..q0+nq-1 are the queries, q0 % 128 == 0);
// indices refer to the sorted order.  tiles_scanned (optional, device) accumulates the number of
// 128-point tiles visited.
int pb200_knn_sorted_f64(const double* data, long long n, int m, int q0, int nq, int k, int* out_idx,
                         double* out_dist, unsigned long long* tiles_scanned, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m < 1 || m > 16 || k < 1 || k > KD_KMAX || k > n || q0 % KD_TILE != 0) {
    set_error_msg("pb200_knn_sorted_f64: need 1 <= m <= 16, 1 <= k <= min(64, n), q0 % 128 == 0");
    return -1;
  }
  if (nq <= 0) return 0;
  switch (m) {
#define PB_CASE(M) case M: return launch_sorted<M>(data, n, q0, nq, k, out_idx, out_dist, tiles_scanned, st);
    PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8)
    PB_CASE(9) PB_CASE(10) PB_CASE(11) PB_CASE(12) PB_CASE(13) PB_CASE(14) PB_CASE(15) PB_CASE(16)
#undef PB_CASE
  }
  return -1;
}

// Bounding boxes box_lo / box_hi [ceil(n/128)][m] of the 128-row tiles of data (n x m, row-major).
int pb200_tile_boxes_md(const double* data, long long n, int m, double* box_lo, double* box_hi, void* stream) {
  if (m < 1 || m > 16) { set_error_msg("pb200_tile_boxes_md: need 1 <= m <= 16"); return -1; }
  if (n <= 0) return 0;
  tile_boxes_md_kernel<<<div_up(n, KD_TILE), KD_TILE, 0, (cudaStream_t)stream>>>(data, n, m, box_lo, box_hi);
  PB_CHECK_LAUNCH("tile_boxes_md_kernel");
  return 0;
}

// Same result as pb200_knn_direct_f64 on cells stored in a locality-preserving order (any order is
// correct; a space-filling curve makes the boxes small), with exact bounding-box pruning; box_lo /
// box_hi from pb200_tile_boxes_md; rows q0..q0+nq-1 are the queries (q0 % 128 == 0); indices refer
// to the stored order.
int pb200_knn_boxed_f64(const double* data, long long n, int m, int q0, int nq, int k, const double* box_lo,
                        const double* box_hi, int* out_idx, double* out_dist, unsigned long long* tiles_scanned,
                        void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m < 1 || m > 16 || k < 1 || k > KD_KMAX || k > n || q0 % KD_TILE != 0) {
    set_error_msg("pb200_knn_boxed_f64: need 1 <= m <= 16, 1 <= k <= min(64, n), q0 % 128 == 0");
    return -1;
  }
  if (nq <= 0) return 0;
  switch (m) {
#define PB_CASE(M) case M: return launch_boxed<M>(data, n, q0, nq, k, box_lo, box_hi, out_idx, out_dist, tiles_scanned, st);
    PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8)
    PB_CASE(9) PB_CASE(10) PB_CASE(11) PB_CASE(12) PB_CASE(13) PB_CASE(14) PB_CASE(15) PB_CASE(16)
#undef PB_CASE
  }
  return -1;
}

}  // extern "C"
