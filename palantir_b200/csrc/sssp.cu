// Batched multi-source shortest paths over the undirected kNN graph.
//
// Replaces scipy.sparse.csgraph.dijkstra(adj, directed=False, indices=waypoints) at
// src/palantir/core.py:362 (one binary-heap Dijkstra per source on one CPU thread,
// 66 % of the reference's wall time at 100k cells).
//
// One CTA per source (persistent CTAs loop over the sources), near/far
// delta-stepping: distances live in the output row D[s][:] (FP64, atomicMin on the
// bit pattern), the "near" work lists and the "far" pile are per-CTA global buffers
// that stay L2-resident, compact membership bitmaps (N/8 bytes each, so the working
// set of all resident sources fits L2) de-duplicate pushes, and only
// __syncthreads separates rounds -- no grid-wide
// synchronisation, no host round trips.  The graph Palantir runs this on is long and
// thin (shortest-path trees tens of thousands of hops deep at 1 M cells, SURVEY.md
// 7b-2), so a source needs ~3*10^4 rounds of a few dozen vertices each: the kernel is
// bound by the latency of one round, not by bandwidth.  A round is therefore flat:
// the frontier's (dist, row extent) are fetched by one thread per vertex, a block
// scan of the degrees assigns ONE ARC PER THREAD, and all arcs of the round are
// relaxed in a single wave of dependent loads.
//
// Any label-correcting order converges to the same fixed point as Dijkstra,
// dist[v] = min_u fl(dist[u] + w_uv), because FP64 addition is monotone; the
// results are therefore bit-identical to a heap-based Dijkstra on the same graph.
#include "common.cuh"

namespace pb200 {

// Variant with a GROUP of LPV lanes per frontier vertex (LPV = 8: up to THREADS/8 vertices of a
// round are expanded at once, each lane issuing the loads of four of its arcs back to back), the
// popped vertex clearing its own membership bit (no clearing pass) and double-buffered counters
// (two block barriers per round).  Distances are re-read when the vertex is expanded, so later
// vertices of a round see the improvements made by earlier ones (fewer re-relaxations than a
// snapshot of the whole frontier).
// HASH: membership in the next near list is tracked in a small shared-memory table instead of a global
// bitmap: slot v mod 1024 holds (round << 32 | v) of the last push that hashed there; a push finding its own
// tag is a duplicate of this round and is dropped.  A collision only lets a duplicate through (expanded twice
// next round: harmless), it can never drop a vertex, because tags of different rounds differ.  The test
// costs a shared-memory atomic instead of an L2 one -- waiting for that result was 19 % of the stall samples.
// The far pile keeps its exact global bitmap (its entries live for many rounds).
__device__ int g_sssp_overflow;

template <int THREADS, int LPV, int UNR = 4, int MINB = 0, bool HASH = false>
__global__ void __launch_bounds__(THREADS, MINB) sssp_group_kernel(
    const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ wts, int n,
    const long long* __restrict__ sources, int n_sources, double delta, double* __restrict__ D,
    int* __restrict__ q_all, unsigned int* __restrict__ bits_all, long long* __restrict__ stats) {
  constexpr int NW = THREADS / 32;
  constexpr int NG = THREADS / LPV;
  constexpr int TAB = 1024;
  __shared__ unsigned long long s_tab[HASH ? TAB : 1];
  __shared__ unsigned int s_ftab[HASH ? TAB : 1];     // far pile: vertex last pushed with this hash (see below)
  __shared__ int s_cnt[2];
  __shared__ int s_cnt_far;
  __shared__ double s_thr;
  __shared__ double s_red[NW];
  __shared__ int s_wtot[NW];
  __shared__ int s_wtot2[NW];
  __shared__ int s_base_keep, s_base_near;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPV, sub = tid % LPV;
  int* qbuf[2];
  qbuf[0] = q_all + (size_t)blockIdx.x * 3 * (size_t)n;
  qbuf[1] = qbuf[0] + n;
  int* far = qbuf[1] + n;
  const int nwords = (n + 31) >> 5;
  unsigned int* bnear[2];
  bnear[0] = bits_all + (size_t)blockIdx.x * 3 * (size_t)nwords;
  bnear[1] = bnear[0] + nwords;
  unsigned int* bfar = bnear[1] + nwords;

  for (int s = blockIdx.x; s < n_sources; s += gridDim.x) {
    double* dist = D + (size_t)s * (size_t)n;
    const int src = (int)sources[s];
    for (int i = tid; i < n; i += THREADS) dist[i] = INFINITY;
    if (HASH) {
      for (int i = tid; i < TAB; i += THREADS) { s_tab[i] = ~0ull; s_ftab[i] = ~0u; }
    } else {
      for (int i = tid; i < 3 * nwords; i += THREADS) bnear[0][i] = 0u;
    }
    __syncthreads();
    int p = 0;
    if (tid == 0) {
      dist[src] = 0.0;
      qbuf[0][0] = src;
      s_cnt[0] = 1; s_cnt[1] = 0; s_cnt_far = 0;
      s_thr = delta;
    }
    __syncthreads();
    long long rounds = 0, relax = 0, advances = 0;

    for (;;) {
      for (;;) {
        int cnt = s_cnt[p];
        if (cnt == 0) break;
        if (HASH && cnt > n) cnt = n;        // overflow of the list was flagged when it happened
        ++rounds;
        const double thr = s_thr;
        const int* qa = qbuf[p];
        int* qb = qbuf[p ^ 1];
        unsigned int* bcur = bnear[p];
        unsigned int* bnext = bnear[p ^ 1];
        // (prefetching the next vertex of the group was measured: its id and row extent one vertex ahead
        // changes nothing, its distance one vertex ahead is stale often enough to raise the re-relaxations
        // from 1.5 to 1.8 per arc and the time by a quarter)
        for (int i = grp; i < cnt; i += NG) {
          const int u = qa[i];
          const double du = __ldcg(dist + u);
          const int beg = indptr[u], end = indptr[u + 1];
          if (!HASH && sub == 0) atomicAnd(&bcur[u >> 5], ~(1u << (u & 31)));   // popped
          for (int e0 = beg + sub; e0 < end; e0 += UNR * LPV) {
            int vv[UNR];
            double nd[UNR], dv[UNR];
            bool ok[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
              const int e = e0 + k * LPV;
              ok[k] = e < end;
              vv[k] = ok[k] ? indices[e] : 0;
              nd[k] = ok[k] ? du + wts[e] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) dv[k] = ok[k] ? __ldcg(dist + vv[k]) : 0.0;
            // the (rare) updates in three batched levels -- all atomicMin of the four arcs, then all
            // membership atomicOr, then the queue stores -- so that their L2 round trips overlap
            // instead of following one another (they were a third of the kernel's stall samples)
            bool imp[UNR];
            double oldv[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) imp[k] = ok[k] && nd[k] < dv[k];
#pragma unroll
            for (int k = 0; k < UNR; ++k) oldv[k] = imp[k] ? atomic_min_pos_f64(dist + vv[k], nd[k]) : 0.0;
            unsigned int prev[UNR];
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
              imp[k] = imp[k] && nd[k] < oldv[k];
              const unsigned int bit = 1u << (vv[k] & 31);
              if (HASH && nd[k] < thr) {
                const unsigned long long tag = ((unsigned long long)rounds << 32) | (unsigned int)vv[k];
                prev[k] = imp[k] ? (atomicExch(&s_tab[vv[k] & (TAB - 1)], tag) == tag ? 1u : 0u) : 1u;
              } else if (HASH) {
                // far pile: a vertex enters it once (its distance only falls, the threshold only rises), so the
                // tag is the vertex itself, cleared when it moves to a near list; an evicted tag lets a second
                // copy into the pile, which is expanded twice at most
                prev[k] = imp[k] ? (atomicExch(&s_ftab[vv[k] & (TAB - 1)], (unsigned int)vv[k]) == (unsigned int)vv[k] ? 1u : 0u)
                                 : 1u;
              } else {
                unsigned int* bm = nd[k] < thr ? bnext : bfar;
                prev[k] = imp[k] ? (atomicOr(&bm[vv[k] >> 5], bit) & bit) : 1u;
              }
            }
#pragma unroll
            for (int k = 0; k < UNR; ++k) {
              if (!prev[k]) {
                if (nd[k] < thr) {
                  const int pos = atomicAdd(&s_cnt[p ^ 1], 1);
                  if (!HASH || pos < n) qb[pos] = vv[k];
                  else g_sssp_overflow = 1;
                } else {
                  const int pos = atomicAdd(&s_cnt_far, 1);
                  if (!HASH || pos < n) far[pos] = vv[k];
                  else g_sssp_overflow = 1;
                }
              }
            }
          }
          if (sub == 0) relax += end - beg;
        }
        __syncthreads();
        if (tid == 0) s_cnt[p] = 0;  // this buffer is the push target of the round after next
        p ^= 1;
        __syncthreads();
      }
      // ---------------- advance the threshold ----------------
      int nfar = s_cnt_far;
      if (nfar == 0) break;
      if (HASH && nfar > n) nfar = n;
      ++advances;
      double mn = INFINITY;
      for (int i = tid; i < nfar; i += THREADS) mn = fmin(mn, __ldcg(dist + far[i]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) s_red[warp] = mn;
      __syncthreads();
      if (tid == 0) {
        double m2 = INFINITY;
        for (int w = 0; w < NW; ++w) m2 = fmin(m2, s_red[w]);
        s_thr = m2 + delta;
        s_base_keep = 0;
        s_base_near = 0;
      }
      __syncthreads();
      const double thr2 = s_thr;
      int* qa = qbuf[p];
      unsigned int* bcur = bnear[p];
      for (int c0 = 0; c0 < nfar; c0 += THREADS) {
        const int i = c0 + tid;
        int v = -1;
        bool to_near = false, keep = false;
        if (i < nfar) {
          v = far[i];
          to_near = __ldcg(dist + v) < thr2;
          keep = !to_near;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        const unsigned mnr = __ballot_sync(0xffffffffu, to_near);
        if (lane == 0) { s_wtot[warp] = __popc(mk); s_wtot2[warp] = __popc(mnr); }
        __syncthreads();
        int offk = 0, offn = 0;
        for (int w = 0; w < warp; ++w) { offk += s_wtot[w]; offn += s_wtot2[w]; }
        const int bk = s_base_keep, bn = s_base_near;
        __syncthreads();
        if (keep) far[bk + offk + __popc(mk & ((1u << lane) - 1u))] = v;
        if (to_near) {
          qa[bn + offn + __popc(mnr & ((1u << lane) - 1u))] = v;
          if (HASH) {
            atomicCAS(&s_ftab[v & (TAB - 1)], (unsigned int)v, ~0u);
          } else {
            atomicAnd(&bfar[v >> 5], ~(1u << (v & 31)));
            atomicOr(&bcur[v >> 5], 1u << (v & 31));
          }
        }
        if (tid == THREADS - 1) {
          s_base_keep = bk + offk + __popc(mk);
          s_base_near = bn + offn + __popc(mnr);
        }
        __syncthreads();
      }
      if (tid == 0) { s_cnt_far = s_base_keep; s_cnt[p] = s_base_near; }
      __syncthreads();
    }
    if (stats) {
      double r = (double)relax;
      r = warp_sum(r);
      if (lane == 0) s_red[warp] = r;
      __syncthreads();
      if (tid == 0) {
        double tot = 0;
        for (int w = 0; w < NW; ++w) tot += s_red[w];
        stats[(size_t)s * 3 + 0] = rounds;
        stats[(size_t)s * 3 + 1] = (long long)tot;
        stats[(size_t)s * 3 + 2] = advances;
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// Arc-record variant.  Each arc carries the row start of its head vertex, {v, indptr[v], w}
// (16 B), and the work lists hold (vertex, row start) pairs.  A popped vertex can therefore issue
// its arc loads at once -- in parallel with its own distance and row end instead of after an
// indptr lookup -- and a push prefetches the pushed vertex's arcs into L2, so that the next
// round finds them there instead of in HBM.  One dependent memory level fewer per round, and the
// longest one (HBM) shortened: the round latency is what bounds this kernel.
// ---------------------------------------------------------------------------------------------
struct __align__(16) SsspArc {
  int v;
  int beg;
  double w;
};

__global__ void sssp_build_arcs_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                       const double* __restrict__ wts, long long nnz, long long nnz_pad,
                                       SsspArc* __restrict__ arcs) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz_pad) return;
  SsspArc a;
  if (e < nnz) { a.v = indices[e]; a.beg = indptr[a.v]; a.w = wts[e]; }
  else { a.v = 0; a.beg = 0; a.w = INFINITY; }
  arcs[e] = a;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1280 / THREADS) sssp_arcrec_kernel(
    const int* __restrict__ indptr, const SsspArc* __restrict__ arcs, int n, const long long* __restrict__ sources,
    int n_sources, double delta, double* __restrict__ D, int2* __restrict__ q_all,
    unsigned int* __restrict__ bits_all, long long* __restrict__ stats) {
  constexpr int LPV = 8;
  constexpr int NW = THREADS / 32;
  constexpr int NG = THREADS / LPV;
  __shared__ int s_cnt[2];
  __shared__ int s_cnt_far;
  __shared__ double s_thr;
  __shared__ double s_red[NW];
  __shared__ int s_wtot[NW];
  __shared__ int s_wtot2[NW];
  __shared__ int s_base_keep, s_base_near;
  constexpr int TAB = 1024;                         // near-list membership: see sssp_group_kernel<..., HASH>
  __shared__ unsigned long long s_tab[TAB];
  __shared__ unsigned int s_ftab[TAB];              // far-pile membership, same scheme
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / LPV, sub = tid % LPV;
  int2* qbuf[2];
  qbuf[0] = q_all + (size_t)blockIdx.x * 3 * (size_t)n;
  qbuf[1] = qbuf[0] + n;
  int2* far = qbuf[1] + n;

  for (int s = blockIdx.x; s < n_sources; s += gridDim.x) {
    double* dist = D + (size_t)s * (size_t)n;
    const int src = (int)sources[s];
    for (int i = tid; i < n; i += THREADS) dist[i] = INFINITY;
    for (int i = tid; i < TAB; i += THREADS) { s_tab[i] = ~0ull; s_ftab[i] = ~0u; }
    __syncthreads();
    int p = 0;
    if (tid == 0) {
      dist[src] = 0.0;
      qbuf[0][0] = make_int2(src, indptr[src]);
      s_cnt[0] = 1; s_cnt[1] = 0; s_cnt_far = 0;
      s_thr = delta;
    }
    __syncthreads();
    long long rounds = 0, relax = 0, advances = 0;

    for (;;) {
      for (;;) {
        int cnt = s_cnt[p];
        if (cnt == 0) break;
        if (cnt > n) cnt = n;                       // overflow of the list was flagged when it happened
        ++rounds;
        const double thr = s_thr;
        const int2* qa = qbuf[p];
        int2* qb = qbuf[p ^ 1];
        for (int i = grp; i < cnt; i += NG) {
          const int2 ub = qa[i];
          const int u = ub.x, beg = ub.y;
          // level 1: own distance, row end and the first 32 arcs, all independent
          SsspArc ar[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) ar[k] = arcs[beg + sub + k * LPV];      // padded: always in bounds
          const double du = __ldcg(dist + u);
          const int end = indptr[u + 1];
          for (int e0 = beg + sub;;) {
            bool ok[4];
            double nd[4], dv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ok[k] = e0 + k * LPV < end;
              nd[k] = du + ar[k].w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) dv[k] = ok[k] ? __ldcg(dist + ar[k].v) : 0.0;     // level 2
            // levels 3-5 batched over the four arcs: every atomicMin, then every membership atomicOr, then the stores
            bool imp[4];
            double oldv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) imp[k] = ok[k] && nd[k] < dv[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) oldv[k] = imp[k] ? atomic_min_pos_f64(dist + ar[k].v, nd[k]) : 0.0;   // level 3
            unsigned int prev[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              imp[k] = imp[k] && nd[k] < oldv[k];
              if (nd[k] < thr) {
                const unsigned long long tag = ((unsigned long long)rounds << 32) | (unsigned int)ar[k].v;
                prev[k] = imp[k] ? (atomicExch(&s_tab[ar[k].v & (TAB - 1)], tag) == tag ? 1u : 0u) : 1u;   // level 4
              } else {
                prev[k] = imp[k] ? (atomicExch(&s_ftab[ar[k].v & (TAB - 1)], (unsigned int)ar[k].v) == (unsigned int)ar[k].v ? 1u : 0u)
                                 : 1u;
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (!prev[k]) {
                if (nd[k] < thr) {
                  const int pos = atomicAdd(&s_cnt[p ^ 1], 1);
                  if (pos < n) qb[pos] = make_int2(ar[k].v, ar[k].beg);
                  else g_sssp_overflow = 1;
                  const char* row = reinterpret_cast<const char*>(arcs + ar[k].beg);
                  prefetch_l2(row); prefetch_l2(row + 128); prefetch_l2(row + 256);
                  prefetch_l2(row + 384); prefetch_l2(row + 512);
                } else {
                  const int pos = atomicAdd(&s_cnt_far, 1);
                  if (pos < n) far[pos] = make_int2(ar[k].v, ar[k].beg);
                  else g_sssp_overflow = 1;
                }
              }
            }
            e0 += 4 * LPV;
            if (e0 - sub >= end) break;                 // uniform within the group
#pragma unroll
            for (int k = 0; k < 4; ++k) ar[k] = arcs[e0 + k * LPV];
          }
          if (sub == 0) relax += end - beg;
        }
        __syncthreads();
        if (tid == 0) s_cnt[p] = 0;  // this buffer is the push target of the round after next
        p ^= 1;
        __syncthreads();
      }
      // ---------------- advance the threshold ----------------
      int nfar = s_cnt_far;
      if (nfar == 0) break;
      if (nfar > n) nfar = n;
      ++advances;
      double mn = INFINITY;
      for (int i = tid; i < nfar; i += THREADS) mn = fmin(mn, __ldcg(dist + far[i].x));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if (lane == 0) s_red[warp] = mn;
      __syncthreads();
      if (tid == 0) {
        double m2 = INFINITY;
        for (int w = 0; w < NW; ++w) m2 = fmin(m2, s_red[w]);
        s_thr = m2 + delta;
        s_base_keep = 0;
        s_base_near = 0;
      }
      __syncthreads();
      const double thr2 = s_thr;
      int2* qa = qbuf[p];
      for (int c0 = 0; c0 < nfar; c0 += THREADS) {
        const int i = c0 + tid;
        int2 vb = make_int2(-1, 0);
        bool to_near = false, keep = false;
        if (i < nfar) {
          vb = far[i];
          to_near = __ldcg(dist + vb.x) < thr2;
          keep = !to_near;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        const unsigned mnr = __ballot_sync(0xffffffffu, to_near);
        if (lane == 0) { s_wtot[warp] = __popc(mk); s_wtot2[warp] = __popc(mnr); }
        __syncthreads();
        int offk = 0, offn = 0;
        for (int w = 0; w < warp; ++w) { offk += s_wtot[w]; offn += s_wtot2[w]; }
        const int bk = s_base_keep, bn = s_base_near;
        __syncthreads();
        if (keep) far[bk + offk + __popc(mk & ((1u << lane) - 1u))] = vb;
        if (to_near) {
          qa[bn + offn + __popc(mnr & ((1u << lane) - 1u))] = vb;
          atomicCAS(&s_ftab[vb.x & (TAB - 1)], (unsigned int)vb.x, ~0u);
          prefetch_l2(arcs + vb.y);
        }
        if (tid == THREADS - 1) {
          s_base_keep = bk + offk + __popc(mk);
          s_base_near = bn + offn + __popc(mnr);
        }
        __syncthreads();
      }
      if (tid == 0) { s_cnt_far = s_base_keep; s_cnt[p] = s_base_near; }
      __syncthreads();
    }
    if (stats) {
      double r = (double)relax;
      r = warp_sum(r);
      if (lane == 0) s_red[warp] = r;
      __syncthreads();
      if (tid == 0) {
        double tot = 0;
        for (int w = 0; w < NW; ++w) tot += s_red[w];
        stats[(size_t)s * 3 + 0] = rounds;
        stats[(size_t)s * 3 + 1] = (long long)tot;
        stats[(size_t)s * 3 + 2] = advances;
      }
    }
    __syncthreads();
  }
}

// Variant with one WARP per frontier vertex (lanes = arcs) and a bit-clearing pass per round.
__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }

}  // namespace pb200

using namespace pb200;

extern "C" {

// CTAs (= concurrently processed sources) pb200_sssp_multi launches; scratch is sized per CTA.  Up to 9 are
// resident per SM (128 threads, 56 registers); beyond 16 per SM the CTAs loop over the sources.
int pb200_sssp_num_ctas(int n_sources) {
  const int cap = 16 * kSMs;
  return n_sources < cap ? n_sources : cap;
}

// D[n_sources][n] = shortest-path distances from sources[s] over the symmetric CSR graph
// (indptr/indices/wts; every undirected edge stored in both directions).  Unreachable = +inf.
// scratch: q[n_ctas * 3 * n] int, bits[n_ctas * 3 * ceil(n/32)] uint32 with
// n_ctas = pb200_sssp_num_ctas(n_sources); stats (optional) [n_sources][3] int64 =
// rounds, arcs relaxed, threshold advances.
// exact_membership = 0: list membership in shared-memory tag tables (the fast kernel; a hash collision can let a
// duplicate into a list, so a list may in theory outgrow its n entries: pb200_sssp_overflowed tells, never observed);
// exact_membership != 0: global bitmaps (no such case; the re-run after an overflow).
int pb200_sssp_multi(const int* indptr, const int* indices, const double* wts, long long n, const long long* sources,
                     int n_sources, double delta, int exact_membership, double* D, int* q, unsigned int* bits,
                     long long* stats, void* stream) {
  if (n_sources <= 0) return 0;
  if (n >= 1073741823LL) { set_error_msg("pb200_sssp_multi: n too large"); return -1; }
  if (!(delta > 0.0)) { set_error_msg("pb200_sssp_multi: delta must be > 0"); return -1; }
  const int grid = pb200_sssp_num_ctas(n_sources);
  // 128 threads per source, 8 lanes per frontier vertex, five arcs in flight per lane (degree <= 40 in one pass)
  if (exact_membership)
    sssp_group_kernel<128, 8, 5, 9, false><<<grid, 128, 0, (cudaStream_t)stream>>>(indptr, indices, wts, (int)n, sources,
                                                                                   n_sources, delta, D, q, bits, stats);
  else
    sssp_group_kernel<128, 8, 5, 9, true><<<grid, 128, 0, (cudaStream_t)stream>>>(indptr, indices, wts, (int)n, sources,
                                                                                  n_sources, delta, D, q, bits, stats);
  PB_CHECK_LAUNCH("sssp_group_kernel");
  return 0;
}

// 1 if a near list of the shared-memory-table variant ever outgrew its buffer since the last call (results
// must then be recomputed with another variant; n entries per list make this a theoretical case); clears the flag.
int pb200_sssp_overflowed(void* stream) {
  int h = 0;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&h, g_sssp_overflow, sizeof(int)) != cudaSuccess) return -1;
  if (h) { const int z = 0; cudaMemcpyToSymbol(g_sssp_overflow, &z, sizeof(int)); }
  return h;
}

// Arc records {v, indptr[v], w} (16 B each, nnz + 64 entries) for pb200_sssp_multi_arcrec.
int pb200_sssp_build_arcs(const int* indptr, const int* indices, const double* wts, long long nnz, void* arcs,
                          void* stream) {
  const long long pad = nnz + 64;
  sssp_build_arcs_kernel<<<div_up(pad, 256), 256, 0, (cudaStream_t)stream>>>(indptr, indices, wts, nnz, pad,
                                                                           (SsspArc*)arcs);
  PB_CHECK_LAUNCH("sssp_build_arcs_kernel");
  return 0;
}

// Number of resident CTAs for the arc-record kernel with `threads` (128, 256 or 512) threads per source.
int pb200_sssp_arcrec_ctas(int n_sources, int threads) {
  const int per_sm = 1280 / threads > 0 ? 1280 / threads : 1;      // 48 registers per thread
  const int cap = per_sm * kSMs;
  return n_sources < cap ? n_sources : cap;
}

// Same result as pb200_sssp_multi from arc records.  scratch: q[n_ctas*3*n] pairs of int (8 B each),
// bits[n_ctas*3*ceil(n/32)] uint32 with n_ctas = pb200_sssp_arcrec_ctas(n_sources, threads).
int pb200_sssp_multi_arcrec(const int* indptr, const void* arcs, long long n, const long long* sources,
                            int n_sources, double delta, int threads, double* D, void* q, unsigned int* bits,
                            long long* stats, void* stream) {
  if (n_sources <= 0) return 0;
  if (n >= 1073741823LL) { set_error_msg("pb200_sssp_multi_arcrec: n too large"); return -1; }
  if (!(delta > 0.0)) { set_error_msg("pb200_sssp_multi_arcrec: delta must be > 0"); return -1; }
  const int grid = pb200_sssp_arcrec_ctas(n_sources, threads);
  cudaStream_t st = (cudaStream_t)stream;
  if (threads == 128)
    sssp_arcrec_kernel<128><<<grid, 128, 0, st>>>(indptr, (const SsspArc*)arcs, (int)n, sources, n_sources, delta, D,
                                                  (int2*)q, bits, stats);
  else if (threads == 256)
    sssp_arcrec_kernel<256><<<grid, 256, 0, st>>>(indptr, (const SsspArc*)arcs, (int)n, sources, n_sources, delta, D,
                                                  (int2*)q, bits, stats);
  else if (threads == 512)
    sssp_arcrec_kernel<512><<<grid, 512, 0, st>>>(indptr, (const SsspArc*)arcs, (int)n, sources, n_sources, delta, D,
                                                  (int2*)q, bits, stats);
  else { set_error_msg("pb200_sssp_multi_arcrec: threads must be 128, 256 or 512"); return -1; }
  PB_CHECK_LAUNCH("sssp_arcrec_kernel");
  return 0;
}

}  // extern "C"
