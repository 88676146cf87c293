// Cell-decomposed multi-source shortest paths (exact up to the association of the path sums).
//
// Replaces scipy.sparse.csgraph.dijkstra(adj, directed=False, indices=waypoints) at
// src/palantir/core.py:362 for large graphs.  The graph Palantir runs this on is the kNN graph
// of a trajectory: long and thin, so a per-source label-correcting run is a chain of ~4*10^4
// dependent rounds of a few dozen vertices each at 1 M cells (sssp.cu: 260 ms for ONE source,
// whatever the number of sources).  The same thinness means small vertex separators, and that is
// what this file uses:
//
//   1. partition   the vertices are grown into P cells around P seeds (one CTA per seed, all
//                  resident, label-correcting on a packed (float distance, cell) key with
//                  atomicMin: an approximate graph-Voronoi partition; ANY partition is valid, a
//                  good one only makes the boundaries small).  Cell boundaries = vertices with an
//                  arc into another cell: ~55 per cell of ~4000 at 1 M cells.
//   2. tables      for every cell c and boundary vertex b: the distances T_c[b][v] to every vertex
//                  v of the cell INSIDE the cell (cells_local_sssp_kernel: one CTA per (c, b),
//                  distances / work lists in shared memory, ~200 shallow rounds; all ~2*10^4 runs
//                  are independent).  Source-independent: done once for all waypoints.
//   3. overlay     graph on the boundary vertices: the arcs that cross cells plus, inside each
//                  cell, the clique T_c[b][b'].  Cells whose boundary is too large for that to pay
//                  (|b|^2 > arcs of the cell: blobs where the data is not thin) are "dissolved":
//                  all their vertices and arcs join the overlay as they are.  Every source gets a
//                  virtual overlay vertex with arcs T(src -> boundary of its cell).  The per-source
//                  kernel of sssp.cu runs on this graph: ~6*10^4 vertices and a few hundred rounds
//                  instead of 10^6 and ~4*10^4.
//   4. expansion   D[s][v] = min_b ( Dov[s][b] + T_c[b][v] ) for every cell: a dense min-plus
//                  product, FP64-pipe bound (cells_expand_kernel), written straight into the
//                  output rows (the only pass over the nW x N matrix).
//
// Every shortest path decomposes into segments that stay inside one cell between boundary vertices,
// so the result equals Dijkstra's up to the order in which the arc weights of a path are added
// (relative differences ~1e-15; tests allow 1e-12).
#include "common.cuh"
#include <cub/cub.cuh>

namespace pb200 {

struct CellInfo {
  int voff;      // first vertex (cell order)
  int size;      // vertices
  int nb;        // boundary vertices (they come first inside the cell)
  int ovoff;     // first overlay id: nb ids when kept, size ids when dissolved
  int kept;      // 1: tables + clique, 0: dissolved into the overlay
  int pad;
  long long toff;  // offset (doubles) of T_c[nb][size] in the table buffer
};

// ------------------------------------------------------------------------------------------
// 1. partition
// ------------------------------------------------------------------------------------------
__global__ void cells_seed_kernel(const int* __restrict__ seeds, int P, unsigned long long* __restrict__ key) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < P) atomicMin(&key[seeds[c]], (unsigned long long)c);   // distance 0
}

__device__ int g_cells_overflow;

constexpr int GROW_THREADS = 128;
constexpr int GROW_LPV = 8;
constexpr int GROW_FINE = 4;      // the key carries the distance in steps of q / 16 ...

// key[v] = (distance to the cell's seed in fixed point, step q/16) << 32 | cell.  A vertex changes hands -- and is
// expanded again -- only when the offer is at least one COARSE step q better than what it holds (q = the mean arc
// weight): with exact comparisons the near-ties of a thin graph turn every last-digit improvement into a wave of
// re-expansions through everything downstream (16 ms at 1 M cells, against 250 breadth-first rounds).  Any
// partition is valid for the decomposition; only the size of the boundaries depends on its quality.
__global__ void __launch_bounds__(GROW_THREADS) cells_grow_kernel(
    const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ wts,
    const int* __restrict__ seeds, int P, float inv_fine, int unit, unsigned long long* __restrict__ key,
    int* __restrict__ q_all, int qcap) {
  __shared__ int s_cnt[2];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, grp = tid / GROW_LPV, sub = tid % GROW_LPV;
  constexpr int NG = GROW_THREADS / GROW_LPV;
  int* const q0 = q_all + (size_t)c * 2 * (size_t)qcap;
  int* const q1 = q0 + qcap;
  if (tid == 0) {
    const int s = seeds[c];
    // a seed that lost its own vertex to an identical seed of lower id grows nothing
    s_cnt[0] = ((unsigned int)key[s] == (unsigned int)c) ? 1 : 0;
    s_cnt[1] = 0;
    q0[0] = s;
  }
  __syncthreads();
  int p = 0;
  for (;;) {
    int cnt = s_cnt[p];
    if (cnt == 0) break;
    if (cnt > qcap) cnt = qcap;
    const int* qa = p ? q1 : q0;
    int* qb = p ? q0 : q1;
    for (int i = grp; i < cnt; i += NG) {
      const int u = qa[i];
      const unsigned long long ku = __ldcg(key + u);
      if ((unsigned int)ku != (unsigned int)c) continue;       // taken over by a closer seed since it was pushed
      const unsigned int du = (unsigned int)(ku >> 32);
      const int beg = indptr[u], end = indptr[u + 1];
      for (int e = beg + sub; e < end; e += GROW_LPV) {
        const int v = indices[e];
        const float wf = unit ? (float)(1 << GROW_FINE) : (float)wts[e] * inv_fine;
        const unsigned int nd = du + (unsigned int)fminf(wf + 0.5f, 1.0e9f);
        const unsigned long long cand = ((unsigned long long)nd << 32) | (unsigned int)c;
        unsigned long long kv = __ldcg(key + v);
        // accepted only when better by a coarse step AND by 1/32 of what the vertex holds: an improvement of a few
        // per cent near a seed would otherwise be handed down through everything behind it, once per such event
        for (;;) {
          const unsigned int od = (unsigned int)(kv >> 32);
          const unsigned int margin = (od >> 5) > (1u << GROW_FINE) ? (od >> 5) : (1u << GROW_FINE);
          if (nd >= od || od - nd < margin) break;
          const unsigned long long seen = atomicCAS(key + v, kv, cand);
          if (seen == kv) {
            const int pos = atomicAdd(&s_cnt[p ^ 1], 1);
            if (pos < qcap) qb[pos] = v;
            else g_cells_overflow = 1;
            break;
          }
          kv = seen;
        }
      }
    }
    __syncthreads();
    if (tid == 0) s_cnt[p] = 0;
    p ^= 1;
    __syncthreads();
  }
}

__global__ void cells_extract_kernel(const unsigned long long* __restrict__ key, long long n, int P,
                                     int* __restrict__ cell) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const unsigned long long k = key[v];
  cell[v] = (k == ~0ull) ? P : (int)(unsigned int)k;
}

// boundary flags, per-cell counts (size, boundary vertices, arcs) and the 64-bit sort key
// (cell, interior, vertex) that puts the boundary vertices of a cell first
__global__ void cells_stats_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, long long n,
                                   const int* __restrict__ cell, int* __restrict__ size, int* __restrict__ nb,
                                   long long* __restrict__ arcs, unsigned long long* __restrict__ sortkey,
                                   int* __restrict__ iota) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int c = -1, isb = 0, deg = 0;
  if (v < n) {
    c = cell[v];
    const int beg = indptr[v], end = indptr[v + 1];
    deg = end - beg;
    for (int e = beg; e < end; ++e) isb |= (cell[indices[e]] != c);
    sortkey[v] = ((unsigned long long)c << 33) | ((unsigned long long)(isb ? 0 : 1) << 32) | (unsigned long long)v;
    iota[v] = (int)v;
  }
  // warp-aggregated histogram updates (consecutive vertices mostly share their cell)
  const unsigned m = __match_any_sync(0xffffffffu, c);
  const unsigned bm = __ballot_sync(0xffffffffu, isb);
  const int leader = __ffs(m) - 1;
  long long dsum = 0;
  for (unsigned rest = m; rest; rest &= rest - 1) {
    const int src = __ffs(rest) - 1;
    dsum += __shfl_sync(m, deg, src);
  }
  if (lane == leader && c >= 0) {
    atomicAdd(&size[c], __popc(m));
    atomicAdd(&nb[c], __popc(m & bm));
    atomicAdd((unsigned long long*)&arcs[c], (unsigned long long)dsum);
  }
}

__global__ void cells_invert_kernel(const int* __restrict__ perm, long long n, int* __restrict__ inv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = (int)i;
}

__global__ void cells_lens_kernel(const int* __restrict__ indptr, const int* __restrict__ perm, long long n,
                                  int* __restrict__ lens) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int o = perm[i]; lens[i] = indptr[o + 1] - indptr[o]; }
  if (i == n) lens[i] = 0;
}

// indices / weights of the relabelled graph (row r = old row perm[r], columns through inv), the cell of every new
// position, and -- arcs that stay inside the cell first -- nint[r] = number of such arcs in row r
__global__ void __launch_bounds__(256) cells_relabel_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                                            const double* __restrict__ wts, long long n,
                                                            const int* __restrict__ perm, const int* __restrict__ inv,
                                                            const int* __restrict__ cell, const int* __restrict__ indptr2,
                                                            int* __restrict__ indices2, double* __restrict__ wts2,
                                                            int* __restrict__ cell2, int* __restrict__ nint) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const unsigned gmask = 0xffu << (lane & 24);
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool live = r < n;                       // whole 8-lane groups are live or not
  if (!live) return;
  const int old = perm[r];
  const int c = cell[old];
  const int b = indptr[old], len = indptr[old + 1] - b, nb2 = indptr2[r];
  int n_in = 0;
  for (int t0 = 0; t0 < len; t0 += 8) {
    const int t = t0 + sub;
    const bool in = t < len && cell[indices[b + t]] == c;
    n_in += __popc(__ballot_sync(gmask, in));
  }
  int k_in = 0, k_out = 0;
  for (int t0 = 0; t0 < len; t0 += 8) {
    const int t = t0 + sub;
    const bool ok = t < len;
    int v = 0;
    bool in = false;
    if (ok) { v = indices[b + t]; in = cell[v] == c; }
    const unsigned m_in = (__ballot_sync(gmask, ok && in) >> (lane & 24)) & 0xffu;
    const unsigned m_out = (__ballot_sync(gmask, ok && !in) >> (lane & 24)) & 0xffu;
    if (ok) {
      const unsigned below = (1u << sub) - 1u;
      const int dst = in ? nb2 + k_in + __popc(m_in & below) : nb2 + n_in + k_out + __popc(m_out & below);
      indices2[dst] = inv[v];
      wts2[dst] = wts[b + t];
    }
    k_in += __popc(m_in);
    k_out += __popc(m_out);
  }
  if (sub == 0) { cell2[r] = c; nint[r] = n_in; }
}

// ------------------------------------------------------------------------------------------
// 2. distances inside a cell (tables T_c and the source rows)
// ------------------------------------------------------------------------------------------
constexpr int LOC_THREADS = 128;
constexpr int LOC_LPV = 8;
constexpr int LOC_UNR = 5;
constexpr int LOC_LCAP = 512;      // vertices expanded per round
constexpr int LOC_MAX_CELL = 16384;

// One work item = one (cell, source vertex inside it): delta-stepping restricted to the arcs that stay inside the
// cell (the first nint[v] of every row).  The distances are the table row itself (global memory, L2-resident while
// the item runs: FP64 atomicMin on the bit pattern); only the bitmap of the vertices whose distance changed since
// they were last expanded and the round's list live in shared memory (4 KB), so 12 items run per SM -- with the
// distances in shared memory (8 B per vertex: 3 items of 8192 vertices per SM) the stage took 29 ms at 1 M cells,
// the same ~350 rounds per item costing more instructions in idle warps than in relaxations.
// A round = (A) each warp scans blocks of 32 bitmap words -- one coalesced load finds the few non-empty ones, the
// frontier being a run of neighbouring vertices -- and moves the vertices below the threshold into the list;
// (B) groups of 8 lanes relax the arcs of the listed vertices.  An empty list raises the threshold to the
// smallest pending distance + delta.
__global__ void __launch_bounds__(LOC_THREADS, 12) cells_local_sssp_kernel(
    const int* __restrict__ indptr, const int* __restrict__ nint, const int* __restrict__ indices,
    const double* __restrict__ wts, const CellInfo* __restrict__ cells, const int* __restrict__ item_cell,
    const int* __restrict__ item_src, const long long* __restrict__ item_out, int n_items, double delta,
    double* __restrict__ T) {
  constexpr int NG = LOC_THREADS / LOC_LPV;
  __shared__ unsigned int pend[LOC_MAX_CELL / 32];
  __shared__ int lv[LOC_LCAP];
  __shared__ int s_n[2];
  __shared__ unsigned long long s_min[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LOC_LPV, sub = tid % LOC_LPV;
  const unsigned long long INF_BITS = 0x7ff0000000000000ull;

  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    const CellInfo ci = cells[item_cell[it]];
    const int size = ci.size, voff = ci.voff;
    const int nwords = (size + 31) >> 5;
    double* dist = T + item_out[it];
    for (int i = tid; i < size; i += LOC_THREADS) dist[i] = INFINITY;
    for (int i = tid; i < nwords; i += LOC_THREADS) pend[i] = 0u;
    __syncthreads();
    if (tid == 0) {
      const int src = item_src[it];
      dist[src] = 0.0;
      pend[src >> 5] = 1u << (src & 31);
      s_n[0] = 0; s_n[1] = 0;
      s_min[0] = INF_BITS; s_min[1] = INF_BITS;
    }
    __syncthreads();
    double thr = delta;
    int cur = 0;
    for (;;) {
      // ---- (A) pending vertices below the threshold -> list; smallest pending distance above it ----
      double mn = INFINITY;
      for (int w0 = warp * 32; w0 < nwords; w0 += (LOC_THREADS / 32) * 32) {
        const unsigned int word = (w0 + lane < nwords) ? pend[w0 + lane] : 0u;
        unsigned int nz = __ballot_sync(0xffffffffu, word != 0u);
        unsigned int keep = word;
        while (nz) {
          const int k = __ffs(nz) - 1;
          nz &= nz - 1u;
          const unsigned int bits = __shfl_sync(0xffffffffu, word, k);
          const int v = ((w0 + k) << 5) + lane;
          const bool isset = (bits >> lane) & 1u;
          const double d = isset ? __ldcg(dist + v) : INFINITY;
          const bool sel = isset && d < thr;
          const unsigned int selm = __ballot_sync(0xffffffffu, sel);
          int base = 0;
          if (selm) {
            if (lane == 0) base = atomicAdd(&s_n[cur], __popc(selm));
            base = __shfl_sync(0xffffffffu, base, 0);
          }
          const int slot = base + __popc(selm & ((1u << lane) - 1u));
          const bool took = sel && slot < LOC_LCAP;
          if (took) lv[slot] = v;
          const unsigned int tookm = __ballot_sync(0xffffffffu, took);
          if (lane == k) keep &= ~tookm;
          if (isset && !took) mn = fmin(mn, d);
        }
        if (keep != word) pend[w0 + lane] = keep;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0 && mn < INFINITY) atomicMin(&s_min[cur], (unsigned long long)__double_as_longlong(mn));
      __syncthreads();
      int n = s_n[cur];
      const unsigned long long mnb = s_min[cur];
      if (tid == 0) { s_n[cur ^ 1] = 0; s_min[cur ^ 1] = INF_BITS; }
      cur ^= 1;
      if (n == 0) {
        if (mnb == INF_BITS) break;
        thr = __longlong_as_double((long long)mnb) + delta;
        __syncthreads();
        continue;
      }
      if (n > LOC_LCAP) n = LOC_LCAP;
      // ---- (B) relax the arcs of the listed vertices: a group of 8 lanes per vertex, two vertices' distances and
      // row extents fetched together ----
      for (int i0 = grp; i0 < n; i0 += 2 * NG) {
        const int i1 = i0 + NG;
        const int u0 = lv[i0], u1 = i1 < n ? lv[i1] : -1;
        const int b0 = indptr[voff + u0], n0 = nint[voff + u0];
        const double d0 = __ldcg(dist + u0);
        const int b1 = u1 >= 0 ? indptr[voff + u1] : 0, n1 = u1 >= 0 ? nint[voff + u1] : 0;
        const double d1 = u1 >= 0 ? __ldcg(dist + u1) : 0.0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h && u1 < 0) break;
          const int beg = h ? b1 : b0, end = beg + (h ? n1 : n0);
          const double du = h ? d1 : d0;
          for (int e0 = beg + sub; e0 < end; e0 += LOC_UNR * LOC_LPV) {
            int vl[LOC_UNR];
            double nd[LOC_UNR], dv[LOC_UNR];
#pragma unroll
            for (int k = 0; k < LOC_UNR; ++k) {
              const int e = e0 + k * LOC_LPV;
              const bool ok = e < end;
              vl[k] = ok ? indices[e] - voff : -1;
              nd[k] = ok ? du + wts[e] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < LOC_UNR; ++k) dv[k] = vl[k] >= 0 ? __ldcg(dist + vl[k]) : 0.0;
            double oldv[LOC_UNR];
#pragma unroll
            for (int k = 0; k < LOC_UNR; ++k)
              oldv[k] = (vl[k] >= 0 && nd[k] < dv[k]) ? atomic_min_pos_f64(dist + vl[k], nd[k]) : 0.0;
#pragma unroll
            for (int k = 0; k < LOC_UNR; ++k)
              if (vl[k] >= 0 && nd[k] < dv[k] && nd[k] < oldv[k]) atomicOr(&pend[vl[k] >> 5], 1u << (vl[k] & 31));
          }
        }
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// 3. overlay graph
// ------------------------------------------------------------------------------------------
// pos_ov[p] = overlay id of the vertex at (cell-order) position p, -1 for the interior of a kept cell
__global__ void cells_pos_ov_kernel(const CellInfo* __restrict__ cells, const int* __restrict__ cell2, long long n,
                                    int* __restrict__ pos_ov, int* __restrict__ ov_pos) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const CellInfo ci = cells[cell2[p]];
  const int local = (int)p - ci.voff;
  int x = -1;
  if (!ci.kept || local < ci.nb) x = ci.ovoff + local;
  pos_ov[p] = x;
  if (x >= 0) ov_pos[x] = (int)p;
}

// degree of every overlay row: rows [0, n_ov) are overlay vertices, rows [n_ov, n_ov + S) the sources
__global__ void cells_overlay_count_kernel(const int* __restrict__ indptr, const int* __restrict__ nint,
                                           const CellInfo* __restrict__ cells, const int* __restrict__ cell2,
                                           const int* __restrict__ pos_ov, const int* __restrict__ ov_pos, int n_ov,
                                           const int* __restrict__ src_pos, int S, int* __restrict__ deg) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > n_ov + S) return;
  if (x == n_ov + S) { deg[x] = 0; return; }
  if (x >= n_ov) {
    const int p = src_pos[x - n_ov];
    deg[x] = pos_ov[p] >= 0 ? 1 : cells[cell2[p]].nb;
    return;
  }
  const int p = ov_pos[x];
  const int c = cell2[p];
  const CellInfo ci = cells[c];
  const int full = indptr[p + 1] - indptr[p];
  deg[x] = ci.kept ? ci.nb - 1 + full - nint[p] : full;
}

// one warp per overlay row
__global__ void __launch_bounds__(256) cells_overlay_fill_kernel(
    const int* __restrict__ indptr, const int* __restrict__ nint, const int* __restrict__ indices,
    const double* __restrict__ wts,
    const CellInfo* __restrict__ cells, const int* __restrict__ cell2, const int* __restrict__ pos_ov,
    const int* __restrict__ ov_pos, int n_ov, const int* __restrict__ src_pos, const long long* __restrict__ src_toff,
    int S, const double* __restrict__ T, const int* __restrict__ ov_indptr, int* __restrict__ ov_indices,
    double* __restrict__ ov_wts) {
  const int lane = threadIdx.x & 31;
  const int x = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (x >= n_ov + S) return;
  int o = ov_indptr[x];
  if (x >= n_ov) {
    const int j = x - n_ov;
    const int p = src_pos[j];
    if (pos_ov[p] >= 0) {
      if (lane == 0) { ov_indices[o] = pos_ov[p]; ov_wts[o] = 0.0; }
      return;
    }
    const CellInfo ci = cells[cell2[p]];
    const double* row = T + src_toff[j];                  // distances inside the cell from the source; boundary first
    for (int b = lane; b < ci.nb; b += 32) { ov_indices[o + b] = ci.ovoff + b; ov_wts[o + b] = row[b]; }
    return;
  }
  const int p = ov_pos[x];
  const int c = cell2[p];
  const CellInfo ci = cells[c];
  const int beg = indptr[p], end = indptr[p + 1];
  if (!ci.kept) {
    for (int e = beg + lane; e < end; e += 32) { ov_indices[o + e - beg] = pos_ov[indices[e]]; ov_wts[o + e - beg] = wts[e]; }
    return;
  }
  const int b = p - ci.voff;
  const double* row = T + ci.toff + (long long)b * ci.size;
  for (int t = lane; t < ci.nb - 1; t += 32) {
    const int b2 = t < b ? t : t + 1;                     // every boundary vertex of the cell but this one
    ov_indices[o + t] = ci.ovoff + b2;
    ov_wts[o + t] = row[b2];
  }
  o += ci.nb - 1;
  // arcs that leave the cell: the tail of the row
  const int cb = beg + nint[p];
  for (int e = cb + lane; e < end; e += 32) { ov_indices[o + e - cb] = pos_ov[indices[e]]; ov_wts[o + e - cb] = wts[e]; }
}

// ------------------------------------------------------------------------------------------
// 4. expansion: D[s][voff + v] = min_b Dov[s][ovoff + b] + T_c[b][v]
// ------------------------------------------------------------------------------------------
constexpr int EXP_THREADS = 256;   // columns per tile
constexpr int EXP_RS = 32;         // sources per CTA
constexpr int EXP_KA = 32;         // boundary vertices per chunk of the source operand
constexpr int EXP_KT = 16;         // boundary vertices per stage of the table operand

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A thread owns one column of the cell and 32 sources: acc[r] = min_k A[k][r] + T[k][col].  The A chunk (the
// sources' overlay distances to the boundary vertices) is shared by the CTA through shared memory; the T values
// are private to the thread's column, so each thread prefetches its own through a two-stage cp.async buffer
// (16 values ahead) -- with the plain load in the loop 53 % of the stall samples were waits for it.  Blocks that
// share a T tile (same columns, other sources) are adjacent in the grid, so the tile comes from L2.
__global__ void __launch_bounds__(EXP_THREADS, 2) cells_expand_kernel(
    const CellInfo* __restrict__ cells, const int2* __restrict__ tiles, int n_stiles, const double* __restrict__ T,
    const double* __restrict__ Dov, long long ldov, int S, double* __restrict__ D, long long n) {
  __shared__ __align__(16) double A[EXP_KA][EXP_RS + 2];   // +2: row stride 68 words, 4-way instead of 32-way store conflicts
  extern __shared__ __align__(16) unsigned char exp_smem[];     // Ts[2][EXP_KT][EXP_THREADS] doubles (64 KB)
  double (*Ts)[EXP_KT][EXP_THREADS] = reinterpret_cast<double (*)[EXP_KT][EXP_THREADS]>(exp_smem);
  const int2 tile = tiles[blockIdx.x / n_stiles];
  const int s0 = (blockIdx.x % n_stiles) * EXP_RS;
  const CellInfo ci = cells[tile.x];
  const int col = tile.y + threadIdx.x;          // local column
  const bool active = col < ci.size;
  const int nb = ci.nb;
  double acc[EXP_RS];
#pragma unroll
  for (int r = 0; r < EXP_RS; ++r) acc[r] = INFINITY;
  const double* Tc = T + ci.toff + (active ? col : 0);
  const int n_stage = (nb + EXP_KT - 1) / EXP_KT;
  auto prefetch = [&](int st) {
    if (active) {
      const int b0 = st * EXP_KT;
#pragma unroll
      for (int k = 0; k < EXP_KT; ++k)
        if (b0 + k < nb) cp_async8(&Ts[st & 1][k][threadIdx.x], Tc + (long long)(b0 + k) * ci.size);
    }
    cp_async_commit();
  };
  if (n_stage > 0) prefetch(0);
  for (int st = 0; st < n_stage; ++st) {
    const int b0 = st * EXP_KT;
    if ((b0 % EXP_KA) == 0) {
      __syncthreads();
      // A[k][r] = Dov[s0 + r][ovoff + b0 + k]
      for (int t = threadIdx.x; t < EXP_KA * EXP_RS; t += EXP_THREADS) {
        const int r = t / EXP_KA, k = t % EXP_KA;
        double v = INFINITY;
        if (b0 + k < nb && s0 + r < S) v = Dov[(long long)(s0 + r) * ldov + ci.ovoff + b0 + k];
        A[k][r] = v;
      }
      __syncthreads();
    }
    if (st + 1 < n_stage) { prefetch(st + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    if (active) {
      const int kc = nb - b0 < EXP_KT ? nb - b0 : EXP_KT;
      const int ka = b0 % EXP_KA;
      for (int k = 0; k < kc; ++k) {
        const double t = Ts[st & 1][k][threadIdx.x];
#pragma unroll
        for (int r = 0; r < EXP_RS; r += 2) {
          const double2 a = *reinterpret_cast<const double2*>(&A[ka + k][r]);
          const double c0 = a.x + t, c1 = a.y + t;
          acc[r] = c0 < acc[r] ? c0 : acc[r];
          acc[r + 1] = c1 < acc[r + 1] ? c1 : acc[r + 1];
        }
      }
    }
  }
  if (active) {
    double* out = D + ci.voff + col;
#pragma unroll
    for (int r = 0; r < EXP_RS; ++r)
      if (s0 + r < S) out[(long long)(s0 + r) * n] = acc[r];
  }
}

// dissolved cells: their vertices are overlay vertices, D[s][p] = Dov[s][pos_ov[p]]
__global__ void cells_expand_copy_kernel(const int* __restrict__ pos_list, int n_list, const int* __restrict__ pos_ov,
                                         const double* __restrict__ Dov, long long ldov, int S,
                                         double* __restrict__ D, long long n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  const int p = pos_list[i];
  const int x = pos_ov[p];
  for (int s = blockIdx.y; s < S; s += gridDim.y) D[(long long)s * n + p] = Dov[(long long)s * ldov + x];
}

// a source inside a kept cell also reaches its cell-mates without leaving the cell
__global__ void cells_expand_src_kernel(const CellInfo* __restrict__ cells, const int* __restrict__ cell2,
                                        const int* __restrict__ src_pos, const long long* __restrict__ src_toff,
                                        const int* __restrict__ pos_ov, int S, const double* __restrict__ T,
                                        double* __restrict__ D, long long n) {
  const int j = blockIdx.x;
  const int p = src_pos[j];
  const CellInfo ci = cells[cell2[p]];
  if (!ci.kept || src_toff[j] < 0) return;
  const double* row = T + src_toff[j];
  double* out = D + (long long)j * n + ci.voff;
  for (int v = threadIdx.x; v < ci.size; v += blockDim.x) {
    const double a = row[v], b = out[v];
    out[v] = a < b ? a : b;
  }
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// Bytes of scratch pb200_cells_order / pb200_cells_relabel need for n vertices.
long long pb200_cells_scratch_bytes(long long n) {
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (const int*)nullptr, (int*)nullptr, (int)n + 1);
  const size_t al = 256;
  size_t tot = ((t1 > t2 ? t1 : t2) + al - 1) / al * al;
  tot += ((size_t)n * 8 + al - 1) / al * al * 2;   // keys in / out
  tot += ((size_t)(n + 1) * 4 + al - 1) / al * al; // iota / lens
  return (long long)tot;
}

// Partition: key[n] must be filled with 0xff bytes by the caller; seeds[P] distinct vertices; step: the coarse
// distance step (the mean arc weight); unit != 0: every arc counts 1 (breadth-first cells); q: P * 2 * qcap ints.
// Afterwards key[v] = (fixed-point distance << 32 | cell) or ~0 for vertices no seed reaches.
int pb200_cells_partition(const int* indptr, const int* indices, const double* wts, long long n, const int* seeds,
                          int P, double step, int unit, unsigned long long* key, int* q, int qcap, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (P <= 0 || qcap <= 0) { set_error_msg("pb200_cells_partition: P and qcap must be positive"); return -1; }
  if (!(step > 0.0)) { set_error_msg("pb200_cells_partition: step must be > 0"); return -1; }
  cells_seed_kernel<<<div_up(P, 128), 128, 0, st>>>(seeds, P, key);
  PB_CHECK_LAUNCH("cells_seed_kernel");
  cells_grow_kernel<<<P, GROW_THREADS, 0, st>>>(indptr, indices, wts, seeds, P, (float)((1 << GROW_FINE) / step), unit,
                                               key, q, qcap);
  PB_CHECK_LAUNCH("cells_grow_kernel");
  return 0;
}

// 1 if a work list of the partition overflowed since the last call (the partition is then still valid --
// every vertex keeps the best key it was given -- but possibly not converged); clears the flag.
int pb200_cells_overflowed(void* stream) {
  int h = 0;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&h, g_cells_overflow, sizeof(int)) != cudaSuccess) return -1;
  if (h) { const int z = 0; cudaMemcpyToSymbol(g_cells_overflow, &z, sizeof(int)); }
  return h;
}

// cell[v] (P = not reached), per-cell size / nb / arcs (P + 1 entries each, zeroed by the caller), and the cell
// order: perm[i] = vertex at position i (cells in id order, boundary vertices first inside a cell), inv = inverse.
int pb200_cells_order(const int* indptr, const int* indices, long long n, const unsigned long long* key, int P,
                      int* cell, int* size, int* nb, long long* arcs, int* perm, int* inv, void* scratch,
                      long long scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n >= 2147483647LL) { set_error_msg("pb200_cells_order: n too large"); return -1; }
  if (scratch_bytes < pb200_cells_scratch_bytes(n)) { set_error_msg("pb200_cells_order: scratch too small"); return -1; }
  const size_t al = 256;
  char* p = (char*)scratch;
  unsigned long long* kin = (unsigned long long*)p;  p += ((size_t)n * 8 + al - 1) / al * al;
  unsigned long long* kout = (unsigned long long*)p; p += ((size_t)n * 8 + al - 1) / al * al;
  int* iota = (int*)p;                               p += ((size_t)(n + 1) * 4 + al - 1) / al * al;
  cells_extract_kernel<<<div_up(n, 256), 256, 0, st>>>(key, n, P, cell);
  PB_CHECK_LAUNCH("cells_extract_kernel");
  cells_stats_kernel<<<div_up(n, 256), 256, 0, st>>>(indptr, indices, n, cell, size, nb, arcs, kin, iota);
  PB_CHECK_LAUNCH("cells_stats_kernel");
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, kin, kout, iota, perm, (int)n);
  int bits = 33;
  while ((1 << (bits - 33)) <= P) ++bits;
  PB_CHECK(cub::DeviceRadixSort::SortPairs((void*)p, temp, kin, kout, iota, perm, (int)n, 0, bits, st));
  cells_invert_kernel<<<div_up(n, 256), 256, 0, st>>>(perm, n, inv);
  PB_CHECK_LAUNCH("cells_invert_kernel");
  return 0;
}

// The graph relabelled into cell order (indptr2 [n+1], indices2 / wts2 [nnz]; in every row the arcs that stay inside
// the cell come first, nint[p] of them) and cell2[p] = cell of position p.
int pb200_cells_relabel(const int* indptr, const int* indices, const double* wts, long long n, const int* perm,
                        const int* inv, const int* cell, int* indptr2, int* indices2, double* wts2, int* cell2,
                        int* nint, void* scratch, long long scratch_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (scratch_bytes < pb200_cells_scratch_bytes(n)) { set_error_msg("pb200_cells_relabel: scratch too small"); return -1; }
  const size_t al = 256;
  char* p = (char*)scratch;
  int* lens = (int*)p; p += ((size_t)(n + 1) * 4 + al - 1) / al * al;
  cells_lens_kernel<<<div_up(n + 1, 256), 256, 0, st>>>(indptr, perm, n, lens);
  PB_CHECK_LAUNCH("cells_lens_kernel");
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, lens, indptr2, (int)n + 1);
  PB_CHECK(cub::DeviceScan::ExclusiveSum((void*)p, temp, lens, indptr2, (int)n + 1, st));
  cells_relabel_kernel<<<div_up(n * 8, 256), 256, 0, st>>>(indptr, indices, wts, n, perm, inv, cell, indptr2,
                                                           indices2, wts2, cell2, nint);
  PB_CHECK_LAUNCH("cells_relabel_kernel");
  return 0;
}

// Largest cell the in-cell kernel takes (its bitmap is static shared memory).
int pb200_cells_max_cell(void) { return LOC_MAX_CELL; }

// In-cell shortest paths for n_items (cell, local source) pairs: T[item_out[i] + v] = distance inside the cell.
// The graph is the cell-ordered one; cells: device array of CellInfo (32 bytes each, see sssp_cells.cu).
int pb200_cells_local_sssp(const int* indptr2, const int* nint, const int* indices2, const double* wts2,
                           const void* cells, const int* item_cell, const int* item_src, const long long* item_out,
                           int n_items, double delta, double* T, void* stream) {
  if (n_items <= 0) return 0;
  const int grid = n_items < 12 * kSMs ? n_items : 12 * kSMs;
  cells_local_sssp_kernel<<<grid, LOC_THREADS, 0, (cudaStream_t)stream>>>(
      indptr2, nint, indices2, wts2, (const CellInfo*)cells, item_cell, item_src, item_out, n_items, delta, T);
  PB_CHECK_LAUNCH("cells_local_sssp_kernel");
  return 0;
}

// pos_ov[n] / ov_pos[n_ov] maps and the degree of every overlay row (deg: n_ov + S + 1 ints, last = 0).
int pb200_cells_overlay_count(const int* indptr2, const int* nint, long long n, const void* cells,
                              const int* cell2, int n_ov, const int* src_pos, int S, int* pos_ov, int* ov_pos,
                              int* deg, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cells_pos_ov_kernel<<<div_up(n, 256), 256, 0, st>>>((const CellInfo*)cells, cell2, n, pos_ov, ov_pos);
  PB_CHECK_LAUNCH("cells_pos_ov_kernel");
  cells_overlay_count_kernel<<<div_up(n_ov + S + 1, 256), 256, 0, st>>>(indptr2, nint, (const CellInfo*)cells,
                                                                       cell2, pos_ov, ov_pos, n_ov, src_pos, S, deg);
  PB_CHECK_LAUNCH("cells_overlay_count_kernel");
  return 0;
}

// ov_indptr = exclusive scan of deg (n_ov + S + 1 entries); scratch as pb200_cells_scratch_bytes(n_ov + S).
int pb200_cells_overlay_scan(const int* deg, int rows_plus_1, int* ov_indptr, void* scratch, long long scratch_bytes,
                             void* stream) {
  size_t temp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, temp, deg, ov_indptr, rows_plus_1);
  if ((long long)temp > scratch_bytes) { set_error_msg("pb200_cells_overlay_scan: scratch too small"); return -1; }
  PB_CHECK(cub::DeviceScan::ExclusiveSum(scratch, temp, deg, ov_indptr, rows_plus_1, (cudaStream_t)stream));
  return 0;
}

int pb200_cells_overlay_fill(const int* indptr2, const int* nint, const int* indices2, const double* wts2, const void* cells,
                             const int* cell2, const int* pos_ov, const int* ov_pos, int n_ov, const int* src_pos,
                             const long long* src_toff, int S, const double* T, const int* ov_indptr,
                             int* ov_indices, double* ov_wts, void* stream) {
  const long long rows = (long long)n_ov + S;
  if (rows <= 0) return 0;
  cells_overlay_fill_kernel<<<div_up(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      indptr2, nint, indices2, wts2, (const CellInfo*)cells, cell2, pos_ov, ov_pos, n_ov, src_pos, src_toff, S, T,
      ov_indptr, ov_indices, ov_wts);
  PB_CHECK_LAUNCH("cells_overlay_fill_kernel");
  return 0;
}

// D[S][n] (cell order) from the overlay distances Dov[S][ldov] and the tables.  tiles: (cell, first column) pairs of
// 256 columns covering every kept cell; dis_pos: positions of the vertices of dissolved cells.
int pb200_cells_expand(const void* cells, const int* cell2, const void* tiles, int n_tiles, const int* dis_pos,
                       int n_dis, const int* pos_ov, const int* src_pos, const long long* src_toff, const double* T,
                       const double* Dov, long long ldov, int S, double* D, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (S <= 0) return 0;
  if (n_tiles > 0) {
    const int n_stiles = div_up(S, EXP_RS);
    const long long blocks = (long long)n_tiles * n_stiles;
    if (blocks >= 2147483647LL) { set_error_msg("pb200_cells_expand: grid too large"); return -1; }
    const int smem = 2 * EXP_KT * EXP_THREADS * (int)sizeof(double);
    PB_CHECK(cudaFuncSetAttribute(cells_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cells_expand_kernel<<<(unsigned int)blocks, EXP_THREADS, smem, st>>>((const CellInfo*)cells, (const int2*)tiles,
                                                                      n_stiles, T, Dov, ldov, S, D, n);
    PB_CHECK_LAUNCH("cells_expand_kernel");
  }
  if (n_dis > 0) {
    dim3 grid(div_up(n_dis, 256), S < 64 ? S : 64);
    cells_expand_copy_kernel<<<grid, 256, 0, st>>>(dis_pos, n_dis, pos_ov, Dov, ldov, S, D, n);
    PB_CHECK_LAUNCH("cells_expand_copy_kernel");
  }
  cells_expand_src_kernel<<<S, 256, 0, st>>>((const CellInfo*)cells, cell2, src_pos, src_toff, pos_ov, S, T, D, n);
  PB_CHECK_LAUNCH("cells_expand_src_kernel");
  return 0;
}

}  // extern "C"
