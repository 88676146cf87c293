// Graph utilities for the pseudotime stage: connected components (reachability from
// the start cell) and the small kernels of the reconnect loop.
//
// Replaces the networkx reachability pass of _connect_graph
// (src/palantir/core.py:806-817; 19.5 s of pure-Python graph building at 100k cells
// in the reference).  A cell is unreachable from the start cell over the undirected
// kNN graph iff it lies in another connected component, so one lock-free union-find
// pass over the arcs answers the question for all cells.
#include "common.cuh"

namespace pb200 {

__device__ __forceinline__ int uf_find(volatile int* parent, int x) {
  for (;;) {
    const int p = parent[x];
    if (p == x) return x;
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving (benign race)
    x = p;
  }
}

__global__ void uf_init_kernel(int* parent, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = (int)i;
}

__global__ void uf_hook_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, long long n,
                               int* parent) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (u >= n) return;
  for (int e = indptr[u] + lane; e < indptr[u + 1]; e += 32) {
    const int v = indices[e];
    if (v <= (int)u) continue;  // each undirected edge once (the graph is symmetric)
    int a = (int)u, b = v;
    for (;;) {
      a = uf_find(parent, a);
      b = uf_find(parent, b);
      if (a == b) break;
      const int hi = a > b ? a : b, lo = a > b ? b : a;
      const int old = atomicCAS(&parent[hi], hi, lo);
      if (old == hi) break;
      a = hi; b = lo;
    }
  }
}

// Final labels.  The walk must not write: a path-halving store that was decided before another
// thread published its final label would overwrite that label with an inner node of the tree
// (seen as 1-6 "unreachable" cells per million, different on every run).  Read-only walks see
// either the old parent or the root, and both lead to the root.
__global__ void uf_compress_kernel(int* parent, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  volatile int* vp = parent;
  int x = (int)i;
  for (;;) {
    const int p = vp[x];
    if (p == x) break;
    x = p;
  }
  vp[i] = x;
}

__global__ void count_other_label_kernel(const int* __restrict__ label, long long n, long long ref,
                                         int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = label[ref];
  const bool other = i < n && label[i] != r;
  const unsigned m = __ballot_sync(0xffffffffu, other);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, __popc(m));
}

// sklearn.metrics.pairwise.euclidean_distances of one point against every row:
// sqrt(max(|x|^2 - 2 x.y + |y|^2, 0)) in FP64 (core.py:824-827, reconnect loop).
__global__ void point_dists_expanded_kernel(const double* __restrict__ data, long long n, int m, long long p,
                                            double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xx = 0.0, yy = 0.0, xy = 0.0;
  for (int c = 0; c < m; ++c) {
    const double a = data[p * m + c], b = data[i * m + c];
    xx = fma(a, a, xx);
    yy = fma(b, b, yy);
    xy = fma(a, b, xy);
  }
  double d2 = -2.0 * xy;
  d2 += xx;
  d2 += yy;
  out[i] = sqrt(d2 > 0.0 ? d2 : 0.0);
}

}  // namespace pb200

using namespace pb200;

extern "C" {

// label[i] = representative (smallest-index root) of i's connected component.
int pb200_graph_components(const int* indptr, const int* indices, long long n, int* label, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  uf_init_kernel<<<div_up(n, 256), 256, 0, st>>>(label, n);
  PB_CHECK_LAUNCH("uf_init_kernel");
  uf_hook_kernel<<<div_up(n, 8), 256, 0, st>>>(indptr, indices, n, label);
  PB_CHECK_LAUNCH("uf_hook_kernel");
  uf_compress_kernel<<<div_up(n, 256), 256, 0, st>>>(label, n);
  PB_CHECK_LAUNCH("uf_compress_kernel");
  return 0;
}

// count[0] = number of i with label[i] != label[ref]
int pb200_count_other_label(const int* label, long long n, long long ref, int* count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PB_CHECK(cudaMemsetAsync(count, 0, sizeof(int), st));
  count_other_label_kernel<<<div_up(n, 256), 256, 0, st>>>(label, n, ref, count);
  PB_CHECK_LAUNCH("count_other_label_kernel");
  return 0;
}

int pb200_point_dists_expanded(const double* data, long long n, int m, long long p, double* out, void* stream) {
  point_dists_expanded_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(data, n, m, p, out);
  PB_CHECK_LAUNCH("point_dists_expanded_kernel");
  return 0;
}

}  // extern "C"
