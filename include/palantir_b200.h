/* palantir_b200 C-ABI: the drop-in boundary for Palantir's diffusion-map and
 * fate-probability hot path on B200 (sm_100a).
 *
 * The reference (dpeerlab/Palantir v1.4.4) is pure Python and has no FFI of its
 * own: every entry point below replaces one third-party CPU library call site in
 * the reference (cited per function as src/palantir/<file>:<line>); the Python
 * mirror in palantir_b200/{utils,core}.py binds them through ctypes exactly as a
 * maintainer of the reference would.
 *
 * Conventions: plain pointers and sizes, no framework types.  Every pointer is a
 * DEVICE pointer unless its name ends in _host.  Matrices are row-major and
 * contiguous.  `stream` is a cudaStream_t.  Every function returns 0 on success;
 * otherwise a CUDA error code (> 0) or -1 (bad argument), with a message in
 * pb200_last_error().  Nothing here allocates device memory: the caller owns
 * every buffer, scratch included (sizes are stated per function).
 * The parser in palantir_b200/_native.py reads the prototypes of this file:
 * one prototype per statement, types limited to int / long long / double /
 * pointers.
 */
#ifndef PALANTIR_B200_H
#define PALANTIR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

const char* pb200_last_error(void);
int pb200_abi_version(void);
int pb200_set_device(int dev);
int pb200_device_sync(void);
/* kernel launches issued by the library since load (bench.py's gpu_launches) */
long long pb200_launch_count(void);

/* ---- exact kNN in PCA space: sklearn NearestNeighbors(brute).kneighbors, utils.py:404-407 ---- */

/* mean[d] = column means of x[n][d] (float32 or float64 input). */
int pb200_col_means(const void* x, int is_f64, long long n, int d, double* mean, void* stream);
/* proj[r] = (float)(x[r] - mean) . v   (FP64 accumulate; the FP32-rounded centred points are what the
 * candidate kernel works on).  With out[c] = sum_r w[r] (x[r][c] - mean[c]) this is one step of the power
 * iteration that finds the leading principal axis used as sort / pruning direction. */
int pb200_project_rows(const void* x, int is_f64, long long n, int d, const double* mean, const double* v,
                       double* proj, void* stream);
int pb200_weighted_col_sums(const void* x, int is_f64, long long n, int d, const double* mean, const double* w,
                            double* out, void* stream);

/* Centred FP32 copy xt[d_pad * ld] of the points in WORK order (work point i = input row perm[i]; perm may be
 * NULL) in the tiled k-major layout the candidate kernel streams (point i = 128*tile + il, coordinate
 * c = 16*kc + k at ((tile*(d_pad/16) + kc)*16 + k)*128 + il), FP32 score norms yn[ld] (+inf on padding),
 * FP64 centred norms xcn[n], FP64 row norms of the original data xn64[n] (all in work order), max centred
 * norm^2 rmax2[1].  d <= 1024, d_pad % 16 == 0, d_pad - d < 16, ld % 256 == 0. */
int pb200_knn_prepare(const void* x, int is_f64, long long n, int d, int d_pad, long long ld, const double* mean,
                      const int* perm, float* xt, float* yn, double* xcn, double* xn64, double* rmax2,
                      void* stream);
/* box_lo / box_hi [ntiles][3]: bounding box of the three projections proj[3][n] (structure of arrays,
 * indexed by input row) of every 128-point work-order tile; padding tiles get (+inf, -inf). */
int pb200_tile_boxes(const double* proj, const int* perm, long long n, long long ntiles, double* box_lo,
                     double* box_hi, void* stream);

/* FP32 candidate generation: for work-order points q0..q0+nq-1 (q0 % 256 == 0) keep the k_keep (<= 48)
 * smallest scores |y|^2 - 2 x.y over all n_pad points.  With box_lo / box_hi (pb200_tile_boxes: boxes of the
 * projections on three orthonormal directions; NULL = none) reference tiles that provably cannot
 * contribute are skipped.  out_idx/out_score [nq][k_keep] ascending (work-order indices), out_thr[nq] =
 * k_keep-th score (lower bound for every excluded reference; +inf if nothing excluded).
 * scratch: lists[ceil(nq/256)*256*64] 64-bit words; tiles_visited: optional counter. */
int pb200_knn_candidates_f32(const float* xt, long long n_pad, const float* yn, const double* xcn,
                             const double* box_lo, const double* box_hi, int q0, int nq, int d, int d_pad,
                             int k_keep, unsigned long long* lists, int* out_idx, float* out_score, float* out_thr,
                             unsigned long long* tiles_visited, void* stream);

/* FP64 re-rank with the reference's expanded-form distance; candidates are work-order indices, x is the
 * input matrix (row of work point j = perm[j]; perm may be NULL); out_idx [nq][k] holds INPUT-order indices,
 * rows are in work order, ascending by (distance, input index); flag[nq] = 1 where the FP32 error bound
 * cannot certify the row (|s32 - s| <= 2^-24 (eps_coef |x||y| + 4 R^2); eps_coef = 2 (d + 8) for the FP32
 * kernel); n_flagged[1]. */
int pb200_knn_rerank_f64(const void* x, int is_f64, const int* perm, const double* xn64, const double* xcn,
                         const double* rmax2, long long n, int d, int q0, int nq, const int* cand,
                         const float* thr32, int k_keep, int k, double eps_coef, int* out_idx, double* out_dist,
                         unsigned char* flag, int* n_flagged, void* stream);

/* Tensor-core (tcgen05, 3xTF32) candidate generation: same outputs as pb200_knn_candidates_f32 for d <= 128.
 * xtc: TF32 hi/lo operand copy (pb200_knn_tc_operand_floats floats) written by pb200_knn_prepare_tc; q0 % 128 == 0;
 * lists: ceil(nq/128)*128*pb200_knn_tc_list_cap() 64-bit words.  The re-rank certificate uses
 * eps_coef = 2 (16 + 4 * 3 * ceil(d/8)). */
long long pb200_knn_tc_operand_floats(long long n_pad, int d);
int pb200_knn_tc_list_cap(void);
int pb200_knn_prepare_tc(const void* x, int is_f64, long long n, int d, long long n_pad, const double* mean,
                         const int* perm, float* xtc, void* stream);
int pb200_knn_candidates_tc(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                            const double* box_lo, const double* box_hi, int q0, int nq, int d, int k_keep,
                            unsigned long long* lists, int* out_idx, float* out_score, float* out_thr,
                            unsigned long long* tiles_visited, void* stream);

/* the same kernel on a sample of query tiles (qtiles[n_list], each < ceil(n / 128)): counts[i] = reference tiles the
 * CTA of qtiles[i] visited (visit_limit = 0), or the tiles visited up to visit_limit plus those that still pass the
 * box test against the bound reached by then (an estimate at visit_limit tiles per CTA) -- the ranks of a sharded
 * run cut the query rows where the cumulative counts are equal.
 * Scratch outputs: out_idx / out_score [n_list * 128 * k_keep], out_thr [n_list * 128] */
int pb200_knn_tc_visit_counts(const float* xtc, long long n_pad, const float* yn, const double* xcn,
                              const double* box_lo, const double* box_hi, long long n, int d, int k_keep,
                              const int* qtiles, int n_list, int visit_limit, int* out_idx, float* out_score,
                              float* out_thr, int* counts, void* stream);

/* rows_out[0..counter) = indices (relative to q0) of flagged rows. */
/* TF32 tensor peak probe: blocks CTAs x iters bare tcgen05.mma.kind::tf32 128x128x8 instructions (262144 FLOP each);
 * timed by the caller; the roofline denominator of pb200_knn_candidates_tc (bench.py) */
int pb200_tf32_peak(int blocks, int iters, void* stream);
int pb200_tf32_peak_rot(int blocks, int iters, int rot, void* stream);   /* rot 0 / 1 / 3: 1 / 2 / 4 accumulators */
int pb200_knn_flagged_rows(const unsigned char* flag, int nq, int* rows_out, int* counter, void* stream);

/* Exact FP64 brute force for the listed rows.  scratch: n_ctas * n doubles. */
int pb200_knn_exact_rows_f64(const void* x, int is_f64, const int* perm, const double* xn64, long long n, int d,
                             const int* rows, int n_rows, int q0, int k, double* scratch, int n_ctas,
                             int* out_idx, double* out_dist, void* stream);
/* out[perm[i]][:] = in[i][:]: neighbour-table rows from work order back to input order */
int pb200_knn_unpermute(const int* idx, const double* dist, long long n, int k, const int* perm, int* out_idx,
                        double* out_dist, void* stream);

/* Roofline probe: blocks x 1024 threads x iters x 16 FFMAs. */
int pb200_fma_peak_f32(float* scratch, int blocks, int iters, void* stream);

/* ---- exact kNN, direct FP64 differences (kd-tree equivalent): core.py:355-356, core.py:525-526 ---- */
int pb200_knn_direct_f64(const double* data, long long n, int m, const double* query, const int* qrows, int q0,
                         int nq, int k, int* out_idx, double* out_dist, void* stream);

/* Same search on cells sorted ascending by column 0 with exact first-coordinate pruning; indices are
 * positions in the sorted order; tiles_scanned (optional) counts visited 128-point tiles. */
int pb200_knn_sorted_f64(const double* data, long long n, int m, int q0, int nq, int k, int* out_idx,
                         double* out_dist, unsigned long long* tiles_scanned, void* stream);
/* Same search on cells stored in any locality-preserving order (Morton curve of the leading coordinates),
 * with exact pruning by the bounding boxes of the 128-row tiles (pb200_tile_boxes_md: box_lo / box_hi
 * [ceil(n/128)][m]); indices are positions in the stored order. */
int pb200_tile_boxes_md(const double* data, long long n, int m, double* box_lo, double* box_hi, void* stream);
int pb200_knn_boxed_f64(const double* data, long long n, int m, int q0, int nq, int k, const double* box_lo,
                        const double* box_hi, int* out_idx, double* out_dist, unsigned long long* tiles_scanned,
                        void* stream);
/* locality ordering: perm[i] = index of the i-th smallest data[:, col]; inv[perm[i]] = i */
long long pb200_argsort_scratch_bytes(long long n);
int pb200_argsort_col_f64(const double* data, long long n, int m, int col, int* perm, int* inv, void* scratch,
                          long long scratch_bytes, void* stream);
/* Morton order of n points from three projections proj[3][n]: 21 bits per axis, origin lo*, cell 1/inv_cell */
int pb200_morton_order(const double* proj, long long n, double lo0, double lo1, double lo2, double inv_cell,
                       int* perm, int* inv, void* scratch, long long scratch_bytes, void* stream);
/* out[i] = a[perm[i]]  /  out[perm[i]] = a[i]   (rows of m doubles) */
int pb200_permute_rows_f64(const double* a, long long n, int m, const int* perm, double* out, void* stream);
int pb200_unpermute_rows_f64(const double* a, long long n, int m, const int* perm, double* out, void* stream);

/* ---- adaptive kernel / CSR construction: utils.py:408-417, 440-446, 478-480 ---- */
int pb200_knn_self_first(const int* idx, long long n, int kc, int* flag_out, void* stream);
int pb200_adaptive_weights(const double* dist, long long n, int kc, int col0, int ka, double* w, void* stream);

/* Symmetrise the arc list (idx[n][kc], w[n][kc-col0]); mode 0: K = W + W^T, mode 1: min(w_ij, w_ji).
 * Scratch: indeg[n], ub[n], cursor[n] (int), off[n+1], bsum[n/1024+2] (long long),
 * tcol/tval[2*n*(kc-col0)], row_nnz[n].  Output indptr[n+1]; nnz = indptr[n]. */
int pb200_symmetrize_phase1(const int* idx, const double* w, long long n, int kc, int col0, int mode,
                            int drop_self, int* indeg, int* ub, int* cursor, long long* off, long long* bsum,
                            int* tcol, double* tval, int* row_nnz, int* indptr, void* stream);
int pb200_symmetrize_phase2(long long n, const long long* off, const int* indptr, const int* tcol,
                            const double* tval, int* indices, double* data, void* stream);

int pb200_csr_row_sums(const int* indptr, const double* data, long long n, double* out, void* stream);
/* op 0: 1/v (v != 0); op 1: v^p (v != 0); op 2: 1/sqrt(v) (v > 0, else 0); op 3: sqrt(v);
 * op 4: 0 where v < p (PResults clips fate probabilities below 0.01, src/palantir/presults.py:34) */
int pb200_vec_transform(double* v, long long n, int op, double p, void* stream);
/* out[e] = (left[row] * data[e]) * right[col]; left / right may be NULL */
int pb200_csr_scale(const int* indptr, const int* indices, const double* data, long long n, const double* left,
                    const double* right, double* out, void* stream);
/* flag_out[0] = number of entries (i,j,v) without a matching (j,i,v') with |v-v'| <= tol*max(|v|,|v'|) */
int pb200_csr_is_symmetric(const int* indptr, const int* indices, const double* data, long long n, double tol,
                           int* flag_out, void* stream);

/* ---- eigensolver: scipy.sparse.linalg.eigs (ARPACK), utils.py:484 ---- */
int pb200_csr_spmv_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                       const double* x, double* y, void* stream);
int pb200_csr_spmm_f32(const int* indptr, const int* indices, const double* vals, long long n, const float* x,
                       long long ldx, float* y, long long ldy, int g, void* stream);
/* gene-tiled SpMM for MAGIC at scale (utils.py:791-804): FP32 CSR values (pb200_cast_f64_f32), x / y FP32 or FP64
 * with 16-byte aligned rows, one launch per tile of 256 / 128 columns; pb200_clip_below = the clip of utils.py:815-821 */
int pb200_csr_spmm_tiled(const int* indptr, const int* indices, const float* vals32, long long n, const void* x,
                         long long ldx, void* y, long long ldy, int g, int is_f64, void* stream);
int pb200_cast_f64_f32(const double* in, float* out, long long n, void* stream);
int pb200_clip_below(void* a, long long count, double thr, int is_f64, void* stream);
int pb200_spmv_long_row(void);
int pb200_csr_spmv_split_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                             const int* long_rows, int n_long, const double* x, double* y, void* stream);
/* rows [row0, row1) of y = A x (y indexed by global row): a rank's slice when the eigensolver is sharded by rows.
 * long_rows: the rows above pb200_spmv_long_row() entries inside the range */
int pb200_csr_spmv_rows_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                            long long row0, long long row1, const int* long_rows, int n_long, const double* x,
                            double* y, void* stream);
/* SpMV over the nonzero stream: a CTA per pb200_csr_spmv_stream_tile() consecutive nonzeros, rows summed inside the
 * tile, rows cut by tile boundaries completed from head / tail (n_tiles doubles each; n_tiles = ceil(nnz / tile)).
 * tile_lo: n_tiles + 1 ints from pb200_csr_spmv_stream_prepare (once per matrix).  Rows [row0, row1) are written;
 * tiles [tile0, tile1) must cover their nonzeros (tile0 = indptr[row0] / tile, tile1 = ceil(indptr[row1] / tile)) */
int pb200_csr_spmv_stream_tile(void);
int pb200_csr_spmv_stream_prepare(const int* indptr, long long n, long long nnz, int* tile_lo, void* stream);
int pb200_csr_spmv_stream_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                              const int* tile_lo, double* head, double* tail, const double* x, double* y,
                              long long row0, long long row1, int tile0, int tile1, void* stream);
int pb200_csr_spmv_window_f64(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                              const double* x, double* y, void* stream);
/* relabelled copy: row r' = row perm[r'], columns mapped through inv (inv[perm[i]] = i); indptr_new = prefix sums
 * of the permuted row lengths (caller) */
int pb200_csr_permute(const int* indptr, const int* indices, const double* vals, long long n, const int* perm,
                      const int* inv, const int* indptr_new, int* indices_new, double* vals_new, void* stream);
int pb200_vec_mul(const double* a, const double* b, double* out, long long n, void* stream);
/* V0 = w / |w| (V32_0: optional FP32 shadow copy).  scratch: partial[n/2048+1], tmp[1] */
int pb200_lanczos_init(const double* w, long long n, double* V0, float* V32_0, double* partial, double* tmp,
                       void* stream);
/* Lanczos steps j0..j0+nsteps-1 with full re-orthogonalisation (three-term recurrence + one Gram-Schmidt pass); alpha[j], beta[j+1], V[j+1] written.
 * V32 (optional, leading dimension ldv): FP32 shadow of the basis, maintained here and read by the Gram-Schmidt pass
 * (it only removes round-off-level components); NULL = the pass reads the FP64 basis.
 * long_rows / n_long (optional): the rows above pb200_spmv_long_row() entries (hubs of the symmetrised kNN graph);
 * the SpMV walks each of them with a whole warp (pb200_csr_spmv_split_f64).
 * windowed != 0: rows are in a locality-preserving order (pb200_csr_permute with the kNN stage's Morton order) and
 * the SpMV keeps the x window of each row block in shared memory (pb200_csr_spmv_window_f64).
 * scratch: w[n], h[j0+nsteps], partial[(j0+nsteps) * (n/2048+1)], tmp[1] */
int pb200_lanczos_steps(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                        const int* long_rows, int n_long, double* V, long long ldv, float* V32, int windowed, int j0,
                        int nsteps, double* alpha, double* beta, double* w, double* h, double* partial, double* tmp,
                        void* stream);
/* U[n][kk] = diag(scale) * sum_j V[j][:] Y[j][kk]  (scale may be NULL), kk <= 32 */
/* the same steps with partial reorthogonalisation (Gram-Schmidt passes only when the omega recurrence asks for them);
 * om: 3 * omstride doubles (omstride >= mcap + 2, om[0] = 1, rest 0 before the first step), gate: 2 ints (zero),
 * counters: 1 int (reorthogonalised steps) */
int pb200_lanczos_steps_pro(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                            const int* long_rows, int n_long, double* V, long long ldv, float* V32, int windowed,
                            int j0, int nsteps, double* alpha, double* beta, double* w, double* h, double* partial,
                            double* tmp, double* om, long long omstride, int* gate, int* counters, double eps,
                            double thresh, void* stream);
/* step j of pb200_lanczos_steps_pro without its product: w = S v_j (all n entries) is already in place -- the
 * eigensolver sharded by rows computes its slice with pb200_csr_spmv_rows_f64 and all-gathers w in between */
int pb200_lanczos_post_pro(long long n, double* V, long long ldv, int j, double* alpha, double* beta, double* w,
                           double* h, double* partial, double* tmp, double* om, long long omstride, int* gate,
                           int* counters, double eps, double thresh, void* stream);
/* Arnoldi steps for a kernel that is not symmetric (utils.py:477-484 takes any kernel): H column-major [ldh][mcap] */
int pb200_arnoldi_steps(const int* indptr, const int* indices, const double* vals, long long n, long long nnz,
                        double* V, long long ldv, int j0, int nsteps, double* H, int ldh, double* w, double* h,
                        double* h2, double* partial, double* tmp, void* stream);
int pb200_ritz_vectors(const double* V, long long ldv, int m, const double* Y, int kk, const double* scale,
                       double* U, long long n, void* stream);
/* unit-L2 columns.  scratch: partial[kk * (n/4096+1)], sumsq[kk] */
int pb200_normalize_columns(double* A, long long n, int kk, double* partial, double* sumsq, void* stream);

/* ---- run_palantir: scaling, boundaries, waypoint sampling (core.py:176-186, 252-312) ---- */
/* scratch: pval[2*m*nblk] doubles, pidx[2*m*nblk] long long, nblk = min(1024, ceil(n/256)) */
int pb200_col_minmax(const double* a, long long n, int m, double* omin, long long* oimin, double* omax,
                     long long* oimax, double* pval, long long* pidx, void* stream);
int pb200_minmax_apply(const double* a, long long n, int m, const double* mn, const double* mx, double* out,
                       void* stream);
/* out[m][iters]; scratch: vt[m*n], md[m*n], cur[m], pv[m*592], pi[m*592], done[m] */
int pb200_maxmin_sampling(const double* data, long long n, int m, const long long* first, int iters,
                          long long* out, double* vt, double* md, long long* cur, double* pv, long long* pi,
                          unsigned int* done, void* stream);
int pb200_gather_rows_f64(const double* a, int m, const long long* rows, long long nr, double* out, void* stream);

/* ---- run_pca (utils.py:52-118: scanpy pca with zero_center=False = truncated SVD without centring) ---- */
int pb200_pca_gram(const void* x, int is_f64, long long n, int g, long long ldx, double* C, void* stream);
int pb200_pca_project(const void* x, int is_f64, long long n, int g, long long ldx, const double* V, int k,
                      double* Y, void* stream);

/* ---- reachability (networkx pass of _connect_graph, core.py:806-817) ---- */
int pb200_graph_components(const int* indptr, const int* indices, long long n, int* label, void* stream);
int pb200_count_other_label(const int* label, long long n, long long ref, int* count, void* stream);
int pb200_point_dists_expanded(const double* data, long long n, int m, long long p, double* out, void* stream);

/* ---- multi-source shortest paths: scipy.sparse.csgraph.dijkstra(directed=False), core.py:362 ---- */
int pb200_sssp_num_ctas(int n_sources);
/* 1 if a work list of the shared-memory-table kernel outgrew its buffer since the last call; clears the flag */
int pb200_sssp_overflowed(void* stream);
/* D[n_sources][n]; scratch: q[n_ctas*3*n] int, bits[n_ctas*3*ceil(n/32)] uint32; stats [n_sources][3] or NULL;
 * exact_membership != 0: global bitmaps instead of shared-memory tag tables (the re-run after an overflow) */
int pb200_sssp_multi(const int* indptr, const int* indices, const double* wts, long long n, const long long* sources,
                     int n_sources, double delta, int exact_membership, double* D, int* q, unsigned int* bits,
                     long long* stats, void* stream);

/* Arc-record form of the same search: arcs[nnz + 64] of {int v; int indptr[v]; double w} (16 B each); work lists
 * hold (vertex, row start) pairs.  scratch: q[n_ctas*3*n] int pairs, bits[n_ctas*3*ceil(n/32)] uint32 with
 * n_ctas = pb200_sssp_arcrec_ctas(n_sources, threads); threads in {128, 256, 512} per source. */
int pb200_sssp_build_arcs(const int* indptr, const int* indices, const double* wts, long long nnz, void* arcs,
                          void* stream);
int pb200_sssp_arcrec_ctas(int n_sources, int threads);
int pb200_sssp_multi_arcrec(const int* indptr, const void* arcs, long long n, const long long* sources,
                            int n_sources, double delta, int threads, double* D, void* q, unsigned int* bits,
                            long long* stats, void* stream);

/* ---- the same search decomposed over cells (sssp_cells.cu), for large thin graphs: partition into cells around
 * seeds, in-cell distance tables from every boundary vertex, per-source search on the overlay graph of the boundary
 * vertices (pb200_sssp_multi on it), dense min-plus expansion into D.  Replaces core.py:362 like pb200_sssp_multi;
 * equal to it up to the association of the path sums (~1e-15 relative).  `cells` is a device array of 32-byte
 * records {int voff, size, nb, ovoff, kept, pad; long long toff} written by the host from the per-cell counts. ---- */
long long pb200_cells_scratch_bytes(long long n);
int pb200_cells_partition(const int* indptr, const int* indices, const double* wts, long long n, const int* seeds,
                          int P, double step, int unit, unsigned long long* key, int* q, int qcap, void* stream);
int pb200_cells_overflowed(void* stream);
int pb200_cells_order(const int* indptr, const int* indices, long long n, const unsigned long long* key, int P,
                      int* cell, int* size, int* nb, long long* arcs, int* perm, int* inv, void* scratch,
                      long long scratch_bytes, void* stream);
int pb200_cells_relabel(const int* indptr, const int* indices, const double* wts, long long n, const int* perm,
                        const int* inv, const int* cell, int* indptr2, int* indices2, double* wts2, int* cell2,
                        int* nint, void* scratch, long long scratch_bytes, void* stream);
int pb200_cells_max_cell(void);
int pb200_cells_local_sssp(const int* indptr2, const int* nint, const int* indices2, const double* wts2,
                           const void* cells, const int* item_cell, const int* item_src, const long long* item_out,
                           int n_items, double delta, double* T, void* stream);
int pb200_cells_overlay_count(const int* indptr2, const int* nint, long long n, const void* cells,
                              const int* cell2, int n_ov, const int* src_pos, int S, int* pos_ov, int* ov_pos,
                              int* deg, void* stream);
int pb200_cells_overlay_scan(const int* deg, int rows_plus_1, int* ov_indptr, void* scratch, long long scratch_bytes,
                             void* stream);
int pb200_cells_overlay_fill(const int* indptr2, const int* nint, const int* indices2, const double* wts2,
                             const void* cells,
                             const int* cell2, const int* pos_ov, const int* ov_pos, int n_ov, const int* src_pos,
                             const long long* src_toff, int S, const double* T, const int* ov_indptr,
                             int* ov_indices, double* ov_wts, void* stream);
int pb200_cells_expand(const void* cells, const int* cell2, const void* tiles, int n_tiles, const int* dis_pos,
                       int n_dis, const int* pos_ov, const int* src_pos, const long long* src_toff, const double* T,
                       const double* Dov, long long ldov, int S, double* D, long long n, void* stream);

/* ---- pseudotime refinement and projection (core.py:368-397, 225-230) ---- */
int pb200_sum_pow_f64(const double* x, long long count, double shift, int power, double* partial, double* out,
                      void* stream);
int pb200_waypoint_colsum(const double* D, int S, long long n, double sdv, double* colsum, void* stream);
int pb200_refine_step(const double* D, int S, long long n, double sdv, const double* colsum, const double* pt,
                      const double* t_mask, const double* t_add, int plus_row, double* out, void* stream);
int pb200_project_fates(const double* D, int S, long long n, double sdv, const double* colsum, const double* B,
                        int nt, double* fates, void* stream);
int pb200_row_entropy(const double* fates, long long n, int nt, double* ent, void* stream);
int pb200_pearson_sums(const double* x, const double* y, long long n, double mx, double my, double* partial,
                       double* out, void* stream);
int pb200_shift_scale(double* x, long long n, double sub, double div, void* stream);

/* ---- waypoint Markov chain and absorption solve (core.py:523-563, 699-737) ---- */
int pb200_markov_chain_dense(const int* ind, const double* dist, int S, int knn, int ka, const double* pt,
                             double* aff, double* rowsum, double* T, void* stream);
int pb200_absorb_system(const double* T, int S, const int* trans, int nt, const int* absn, int na, double* A,
                        double* R, void* stream);
int pb200_dense_solve_f64(double* A, double* B, int n, int r, int* singular, void* stream);
int pb200_dense_matvec_t(const double* M, int S, const double* x, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PALANTIR_B200_H */
