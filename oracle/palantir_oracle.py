"""CPU oracle for the Palantir diffusion-map / fate-probability hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``palantir_b200/`` imports this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.  The product path runs on the
CUDA extension and fails loudly when it is missing.

What this is: an array-level restatement (numpy in / numpy out, positions
instead of pandas labels) of the reference's algorithm, calling the *same*
third-party CPU libraries at the same call sites the reference does -- the
reference is 100 % Python and all of its arithmetic lives in scikit-learn,
scipy and networkx, none of which is vendored or pinned by the reference
(``pyproject.toml:29-46``).  The versions that define the oracle are the ones in
this image: numpy 2.3.5, scipy 1.18.1, scikit-learn 1.9.0, pandas 3.0.2,
networkx 3.6.1.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the reference's own
source (``/root/reference/src/palantir``, loaded by ``oracle/ref_shim.py``) on
seeded inputs, checks that this restatement reproduces it, and commits the
reference outputs as fixtures under ``tests/golden/``; the CPU tests re-check
the restatement against those fixtures on every run (bit for bit), and, where the
reference checkout is present, against the reference source itself on further
inputs (``tests/test_oracle_golden.py::test_oracle_vs_reference_source_live``:
bit for bit up to the pseudotime; the fate projection ``np.dot(W.T, B)`` goes
through BLAS, whose summation order depends on the operands' memory layout, so
fates and entropy agree to a few ulp there).

Each function cites the reference ``file:line`` it follows.
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix, eye as sp_eye, find as sp_find
from scipy.sparse import csgraph
from scipy.sparse.linalg import eigs, splu
from scipy.stats import entropy as _entropy, norm as _norm, pearsonr
from sklearn.metrics import pairwise_distances
from sklearn.neighbors import NearestNeighbors
from sklearn.preprocessing import minmax_scale as _sk_minmax

# ----------------------------------------------------------------------------
# diffusion maps (reference: src/palantir/utils.py)
# ----------------------------------------------------------------------------


def truncated_svd_pca(x, n_comps: int):
    """What ``sc.pp.pca(ad, n_comps=n_comps, zero_center=False)`` computes (the call of run_pca, utils.py:92-104):
    scanpy's zero_center=False branch is sklearn ``TruncatedSVD(n_components, random_state=0, algorithm="arpack")``
    (scanpy/preprocessing/_pca/__init__.py, pinned scanpy>=1.6).  scanpy itself is not installed in the build
    image, so this restatement is PARITY UNPINNED against the reference's own function; it is anchored on the
    sklearn estimator scanpy delegates to.  Returns (X_pca, components, variance, variance_ratio)."""
    from sklearn.decomposition import TruncatedSVD

    svd = TruncatedSVD(n_components=n_comps, random_state=0, algorithm="arpack")
    scores = svd.fit_transform(x)
    return scores, svd.components_, svd.explained_variance_, svd.explained_variance_ratio_


def knn_with_self(x: np.ndarray, n_neighbors: int, n_jobs: int = -1):
    """(dist, ind) of the ``n_neighbors`` nearest rows of ``x`` to every row of
    ``x`` (self included), as ``sklearn.neighbors.NearestNeighbors`` returns them.
    Call site: utils.py:404-407 (PCA space, brute force for d > 15) and
    core.py:355-356 / core.py:525-526 (multiscale space, kd-tree)."""
    nn = NearestNeighbors(n_neighbors=n_neighbors, metric="euclidean", n_jobs=n_jobs).fit(x)
    return nn.kneighbors(x)


def knn_expanded_f64(x: np.ndarray, n_neighbors: int):
    """Pure-numpy statement of what sklearn's brute ``EuclideanArgKmin`` computes
    (small inputs only): FP64 ``|x|^2 - 2 x.y + |y|^2`` on inputs upcast to FP64,
    clamped at 0, k smallest with the lowest index winning exact ties; the
    returned distance is ``sqrt`` of that expanded value -- taken in float32
    when the input is float32."""
    x64 = np.asarray(x, dtype=np.float64)
    sq = np.einsum("ij,ij->i", x64, x64)
    d2 = sq[:, None] - 2.0 * (x64 @ x64.T) + sq[None, :]
    np.maximum(d2, 0.0, out=d2)
    order = np.argsort(d2, axis=1, kind="stable")[:, :n_neighbors]
    d2k = np.take_along_axis(d2, order, axis=1)
    if np.asarray(x).dtype == np.float32:
        dist = np.sqrt(d2k.astype(np.float32)).astype(np.float64)
    else:
        dist = np.sqrt(d2k)
    return dist, order


def adaptive_kernel(x: np.ndarray, knn: int = 30, alpha: float = 0.0, n_jobs: int = -1) -> csr_matrix:
    """Adaptive anisotropic kernel, ``backend="sklearn"`` branch.
    utils.py:390,403-417 (kNN, self strip, sigma = distance to the floor(knn/3)-th
    neighbour, ``exp(-d/sigma)``), :440 (``W + W.T``), :442-446 (alpha)."""
    n = x.shape[0]
    dist, ind = knn_with_self(x, min(knn + 1, n), n_jobs)
    return kernel_from_knn(dist, ind, knn, alpha)


def kernel_from_knn(dist: np.ndarray, ind: np.ndarray, knn: int, alpha: float = 0.0) -> csr_matrix:
    """Same as :func:`adaptive_kernel` from a given neighbour table (stage-isolated
    parity: both sides are fed the same kNN).  utils.py:408-417,440-446."""
    n = dist.shape[0]
    if dist.shape[1] > 1 and np.all(ind[:, 0] == np.arange(n)):
        dist, ind = dist[:, 1:], ind[:, 1:]
    k_eff = dist.shape[1]
    ka = max(1, min(int(np.floor(knn / 3)), k_eff))
    sigma = dist[:, ka - 1]
    scaled = dist.ravel() / np.repeat(sigma, k_eff)
    w = csr_matrix((np.exp(-scaled), (np.repeat(np.arange(n), k_eff), ind.ravel())), shape=(n, n))
    kern = w + w.T
    if alpha > 0:
        deg = np.ravel(kern.sum(axis=1))
        nz = deg != 0
        deg[nz] = deg[nz] ** (-alpha)
        s = csr_matrix((deg, (range(n), range(n))), shape=(n, n))
        kern = s.dot(kern).dot(s)
    return kern


def transition_matrix(kernel: csr_matrix) -> csr_matrix:
    """Row-stochastic ``T = D^-1 K``; zero rows stay zero.  utils.py:477-480."""
    n = kernel.shape[0]
    deg = np.ravel(kernel.sum(axis=1))
    nz = deg != 0
    deg[nz] = 1 / deg[nz]
    return csr_matrix((deg, (range(n), range(n))), shape=(n, n)).dot(kernel)


def diffusion_eigs(kernel: csr_matrix, n_components: int = 10, seed=0, tol: float = 1e-4):
    """Top eigenpairs of T.  utils.py:482-493 (ARPACK, ``which='LM'``, start vector
    ``np.random.rand(N)`` after ``np.random.seed(seed)``, real part, descending,
    unit L2 columns).  ``tol=1e-4`` is the reference's value ("as-is"); a tight
    tolerance gives the converged answer the GPU solver is compared with."""
    t = transition_matrix(kernel)
    np.random.seed(seed)
    v0 = np.random.rand(min(t.shape))
    lam, vec = eigs(t, n_components, tol=tol, maxiter=1000, v0=v0)
    lam, vec = np.real(lam), np.real(vec)
    order = np.argsort(lam)[::-1]
    lam, vec = lam[order], vec[:, order]
    for c in range(vec.shape[1]):
        vec[:, c] = vec[:, c] / np.linalg.norm(vec[:, c])
    return t, vec, lam


def pick_n_eigs(lam: np.ndarray) -> int:
    """Eigen-gap rule of utils.py:893-900."""
    lam = np.ravel(lam)
    gaps = lam[:-1] - lam[1:]
    order = np.argsort(gaps)
    n_eigs = int(order[-1] + 1)
    if n_eigs < 3:
        n_eigs = int(order[-2] + 1)
    if n_eigs < 3:
        n_eigs = 3
    return n_eigs


def multiscale_space(vec: np.ndarray, lam: np.ndarray, n_eigs=None) -> np.ndarray:
    """``V[:, 1:n_eigs] * lam/(1-lam)``.  utils.py:893-906."""
    if n_eigs is None:
        n_eigs = pick_n_eigs(lam)
    use = list(range(1, n_eigs))
    ev = np.ravel(np.asarray(lam)[use])
    return vec[:, use] * (ev / (1 - ev))


def magic_impute(t: csr_matrix, x: np.ndarray, n_steps: int = 3, clip_threshold: float = 1e-2):
    """``(T**n_steps).astype(float32) @ X`` with the small-value clip.  utils.py:791-821
    (dense X, ``sparse=False`` flavour: values below the threshold become 0)."""
    tp = (t ** n_steps).astype(np.float32)
    out = np.asarray(tp.dot(x))
    out[out < clip_threshold] = 0
    return out


# ----------------------------------------------------------------------------
# run_palantir (reference: src/palantir/core.py)
# ----------------------------------------------------------------------------


def minmax_scale(ms: np.ndarray) -> np.ndarray:
    """core.py:176-180 (``sklearn.preprocessing.minmax_scale``)."""
    return _sk_minmax(ms)


def boundary_cells(data: np.ndarray) -> np.ndarray:
    """Positions of the per-column arg-max and arg-min cells, sorted, unique.
    core.py:186 (``idxmax`` / ``idxmin`` take the first occurrence)."""
    return np.unique(np.concatenate([np.argmax(data, axis=0), np.argmin(data, axis=0)]))


def pick_start_cell(data: np.ndarray, boundaries: np.ndarray, early: int) -> int:
    """Boundary cell nearest (sklearn ``pairwise_distances``) to the early cell.
    core.py:187-190."""
    d = pairwise_distances(data[boundaries, :], data[early, :].reshape(1, -1))
    return int(boundaries[int(np.argmin(np.ravel(d)))])


def max_min_sampling(data: np.ndarray, num_waypoints: int, seed=None) -> np.ndarray:
    """Per-column farthest-point sampling.  core.py:275-310: ``int(nW/m)`` picks
    per column, first pick ``np.random.randint(N)`` from the global MT19937 seeded
    once, then repeated ``argmax`` of the running minimum 1-D distance (lowest
    index on ties).  Returns positions, first occurrences kept, in pick order."""
    n, m = data.shape
    floor_wp = max(3, m)
    if num_waypoints < floor_wp:
        warnings.warn(
            f"num_waypoints={num_waypoints} is too small for {m} components; "
            f"using {floor_wp} instead for stable sampling.", UserWarning)
        num_waypoints = floor_wp
    per_col = int(num_waypoints / m)
    if seed is not None:
        np.random.seed(seed)
    picks = []
    for c in range(m):
        v = data[:, c]
        cur = np.random.randint(n)
        picks.append(cur)
        md = np.abs(v - v[cur])
        for _ in range(1, per_col):
            cur = int(np.argmax(md))
            picks.append(cur)
            md = np.minimum(md, np.abs(v - v[cur]))
    picks = np.asarray(picks, dtype=np.int64)
    _, first = np.unique(picks, return_index=True)
    return picks[np.sort(first)]


def assemble_waypoints(sampled: np.ndarray, boundaries: np.ndarray, terminals, start: int) -> np.ndarray:
    """core.py:203-209: sorted union of sampled, boundary and terminal cells, minus
    the start cell, start cell prepended (positions stand in for labels; with
    zero-padded names label order equals position order)."""
    wp = np.union1d(sampled, boundaries)
    if terminals is not None:
        wp = np.union1d(wp, np.asarray(list(terminals), dtype=np.int64))
    wp = wp[wp != start]
    return np.concatenate([[start], wp]).astype(np.int64)


def knn_graph_with_self(data: np.ndarray, knn: int, n_jobs: int = -1) -> csr_matrix:
    """core.py:355-356 ``kneighbors_graph(data, mode="distance")``: ``knn`` arcs per
    row, self included as an explicit zero."""
    nn = NearestNeighbors(n_neighbors=knn, metric="euclidean", n_jobs=n_jobs).fit(data)
    return nn.kneighbors_graph(data, mode="distance")


def connect_graph(adj: csr_matrix, data: np.ndarray, start: int) -> csr_matrix:
    """core.py:806-842: while some cell is unreachable from ``start`` over the
    undirected graph, add an arc from the farthest reachable cell to its nearest
    (sklearn euclidean) unreachable cell.  The reference walks the graph with
    networkx; so does this."""
    import networkx as nx

    n = data.shape[0]

    def reach():
        g = nx.Graph(adj)
        return nx.single_source_dijkstra_path_length(g, start)

    dist = reach()
    if len(dist) < n:
        warnings.warn(
            "Some of the cells were unreachable. Consider increasing the k for \n \
            nearest neighbor graph construction.")
    while len(dist) < n:
        keys = np.fromiter(dist.keys(), dtype=np.int64, count=len(dist))
        vals = np.fromiter(dist.values(), dtype=np.float64, count=len(dist))
        far = int(keys[int(np.argmax(vals))])           # first maximum in networkx visiting order
        missing = np.setdiff1d(np.arange(n), keys)
        dd = np.ravel(pairwise_distances(data[far, :].reshape(1, -1), data[missing, :]))
        j = int(np.argmin(dd))
        adj[far, int(missing[j])] = dd[j]
        dist = reach()
    return adj


def refine_pseudotime(dmat: np.ndarray, wp_idx: np.ndarray, start_row: int, max_iterations: int = 25):
    """core.py:368-397 on the dense ``nW x N`` shortest-path matrix: Gaussian
    waypoint weights with the Silverman-style bandwidth, then the perspective
    iteration until Pearson > 0.9999, then shift/scale to [0, 1]."""
    flat = dmat.ravel()
    sdv = np.std(flat) * 1.06 * len(flat) ** (-1 / 5)
    w = np.exp(-0.5 * np.power(dmat / sdv, 2))
    w = w / w.sum(axis=0, keepdims=True)
    pt = dmat[start_row, :].copy()
    it, done = 1, False
    corrs = []
    while not done and it < max_iterations:
        t_wp = pt[wp_idx][:, None]
        sign = np.where(pt[None, :] < t_wp, -1.0, 1.0)
        t_wp[start_row] = 0.0
        sign[start_row, :] = 1.0
        new = np.sum((dmat * sign + t_wp) * w, axis=0)
        corr = pearsonr(pt, new)[0]
        corrs.append(corr)
        done = corr > 0.9999
        pt = new
        it += 1
    pt = pt - np.min(pt)
    pt = pt / np.max(pt)
    return pt, w, sdv, corrs


def compute_pseudotime(data: np.ndarray, start: int, knn: int, wp_idx: np.ndarray,
                       n_jobs: int = -1, max_iterations: int = 25):
    """core.py:353-401: kNN graph -> connect -> multi-source Dijkstra (undirected)
    -> :func:`refine_pseudotime`.  Returns ``(pseudotime[N], W[nW,N], D[nW,N])``."""
    adj = knn_graph_with_self(data, knn, n_jobs)
    adj = connect_graph(adj, data, start)
    dmat = np.asarray(csgraph.dijkstra(adj, directed=False, indices=wp_idx), dtype=float)
    start_row = int(np.flatnonzero(wp_idx == start)[0])
    pt, w, _, _ = refine_pseudotime(dmat, wp_idx, start_row, max_iterations)
    return pt, w, dmat


def markov_chain(wp_data: np.ndarray, knn: int, pt_wp: np.ndarray, n_jobs: int = -1) -> csr_matrix:
    """core.py:523-563: waypoint kNN (self included), sigma = dist[:, min(knn//3 - 1, 30)],
    keep i->j iff pt_j >= pt_i - sigma_i, affinity exp(-z^2/(2 sigma_i^2) - z^2/(2 sigma_j^2))
    (zero distances vanish through ``find``), row-normalise."""
    n = wp_data.shape[0]
    dist, ind = knn_with_self(wp_data, knn, n_jobs)
    ka = int(np.min([int(np.floor(knn / 3)) - 1, 30]))
    sig = np.ravel(dist[:, ka])
    keep = pt_wp[ind] >= (pt_wp - sig)[:, None]
    r, c = np.nonzero(keep)
    g = csr_matrix((dist[r, c], (r, ind[r, c])), shape=(n, n))
    x, y, z = sp_find(g)
    aff = np.exp(-(z ** 2) / (sig[x] ** 2) * 0.5 - (z ** 2) / (sig[y] ** 2) * 0.5)
    w = csr_matrix((aff, (x, y)), [n, n])
    deg = np.ravel(w.sum(axis=1))
    x, y, z = sp_find(w)
    return csr_matrix((z / deg[x], (x, y)), [n, n])


def terminal_states_from_chain(t: csr_matrix, wp_data: np.ndarray, pt_wp: np.ndarray, v0=None) -> np.ndarray:
    """core.py:590-649, positions into the waypoint list.  ``v0`` makes the ARPACK
    start reproducible (the reference leaves it to scipy's entropy-seeded default)."""
    bnd = boundary_cells(wp_data)
    n = min(t.shape)
    if n <= 2:
        vals, vecs = np.linalg.eig(t.T.toarray())
        ranks = np.abs(np.real(vecs[:, np.argsort(vals)[-1]]))
    else:
        vals, vecs = eigs(t.T, 1, maxiter=n * 50, v0=v0)
        ranks = np.abs(np.real(vecs[:, 0]))
    med = np.median(ranks)
    cutoff = _norm.ppf(0.9999, loc=med, scale=np.median(np.abs(ranks - med)))
    sel = np.flatnonzero(ranks > cutoff)
    sub = t[sel, :][:, sel]
    n_comp, lab = csgraph.connected_components(sub, directed=False)
    cand = []
    for c in range(n_comp):
        members = sel[lab == c]
        cand.append(int(members[int(np.argmax(pt_wp[members]))]))
    if not cand:
        return np.array([], dtype=np.int64)
    dd = pairwise_distances(wp_data[bnd, :], wp_data[cand, :])
    return np.unique(bnd[np.argmin(dd, axis=0)])


def identify_terminal_states(ms: np.ndarray, early: int, knn: int = 30, num_waypoints: int = 1200,
                             n_jobs: int = -1, max_iterations: int = 25, seed: int = 20):
    """core.py:404-491 on positions: scale, boundary / start cell, waypoints (no terminal cells joined),
    pseudotime, waypoint Markov chain, :func:`terminal_states_from_chain`.  Returns
    ``(terminal cell positions, excluded boundary cell positions)`` -- the boundary cells of the WAYPOINT
    matrix that are neither terminal nor the start cell (core.py:488-490)."""
    data = minmax_scale(np.asarray(ms, dtype=np.float64))
    bnd = boundary_cells(data)
    start = pick_start_cell(data, bnd, early)
    wp = assemble_waypoints(max_min_sampling(data, int(num_waypoints), seed), bnd, None, start)
    pt, _, _ = compute_pseudotime(data, start, knn, wp, n_jobs, max_iterations)
    wp_data = data[wp, :]
    chain = markov_chain(wp_data, knn, pt[wp], n_jobs)
    term = wp[terminal_states_from_chain(chain, wp_data, pt[wp])]
    wb = wp[boundary_cells(wp_data)]
    excluded = np.setdiff1d(np.setdiff1d(wb, term), [start])
    return term, excluded


def absorption_probabilities(t: csr_matrix, abs_states: np.ndarray):
    """core.py:694-765 on positions: absorbing rows, ``(I-Q) B = R`` by SuperLU,
    clamp, entropy, identity rows appended for the absorbing states.
    Returns ``(row_positions, entropy, B)`` with transient rows first (ascending
    position) then the absorbing ones; columns follow ``abs_states`` order."""
    n = t.shape[0]
    abs_states = np.asarray(abs_states, dtype=np.int64)
    t = t.tolil()
    t[abs_states, :] = 0
    t[abs_states, abs_states] = 1
    t = t.tocsr()
    trans = np.setdiff1d(np.arange(n), abs_states)
    if trans.size == 0:
        return abs_states, np.zeros(abs_states.size), np.eye(abs_states.size)
    q = t[trans, :][:, trans]
    i_q = sp_eye(q.shape[0], format="csc") - q.tocsc()
    r = t[trans, :][:, abs_states].toarray()
    b = splu(i_q).solve(r)
    b[b < 0] = 0
    ent = _entropy(b, axis=1)
    rows = np.concatenate([trans, abs_states])
    return rows, np.concatenate([ent, np.zeros(abs_states.size)]), np.vstack([b, np.eye(abs_states.size)])


def run_palantir(ms: np.ndarray, early: int, terminal=None, knn: int = 30, num_waypoints=1200,
                 n_jobs: int = -1, scale_components: bool = True, use_early_cell_as_start: bool = False,
                 max_iterations: int = 25, seed: int = 20, row_apply_entropy: bool = True):
    """Array-level restatement of core.py:61-249 (+ presults.py:34).

    ``terminal`` is a sequence of cell positions (or None for auto-detection);
    ``num_waypoints`` an int or an array of pre-made waypoint positions.
    ``row_apply_entropy`` keeps the reference's per-row pandas ``apply`` for the
    final entropy (core.py:230) so that a timing of this function is a timing of
    the reference's algorithm, cost profile included.
    Returns a dict: pseudotime[N], entropy[N], fate_probs[N,nT] (after the 0.01
    clip), fate_probs_raw (before it), terminal (column order = waypoint order),
    waypoints, start."""
    ms = np.asarray(ms, dtype=np.float64)
    data = minmax_scale(ms) if scale_components else ms.copy()
    bnd = boundary_cells(data)
    start = early if use_early_cell_as_start else pick_start_cell(data, bnd, early)
    if isinstance(num_waypoints, (int, np.integer)):
        sampled = max_min_sampling(data, int(num_waypoints), seed)
    else:
        sampled = np.asarray(num_waypoints, dtype=np.int64)
    wp = assemble_waypoints(sampled, bnd, terminal, start)
    pt, w, dmat = compute_pseudotime(data, start, knn, wp, n_jobs, max_iterations)

    wp_data = data[wp, :]
    chain = markov_chain(wp_data, knn, pt[wp], n_jobs)
    if terminal is None:
        abs_pos = terminal_states_from_chain(chain, wp_data, pt[wp])
    else:
        abs_pos = np.flatnonzero(np.isin(wp, np.asarray(list(terminal))))
    if abs_pos.size == 0:
        fates = np.zeros((ms.shape[0], 0))
        ent = np.zeros(ms.shape[0])
        return dict(pseudotime=pt, entropy=ent, fate_probs=fates, fate_probs_raw=fates.copy(),
                    terminal=wp[abs_pos], waypoints=wp, start=start, W=w, D=dmat)
    rows, _, b = absorption_probabilities(chain, abs_pos)
    b_in_wp_order = np.empty_like(b)
    b_in_wp_order[rows, :] = b
    fates = np.dot(w.T, b_in_wp_order)
    if row_apply_entropy:
        ent = pd.DataFrame(fates).apply(_entropy, axis=1).values
    else:
        ent = _entropy(fates, axis=1)
    raw = fates.copy()
    fates = fates.copy()
    fates[fates < 0.01] = 0
    return dict(pseudotime=pt, entropy=ent, fate_probs=fates, fate_probs_raw=raw,
                terminal=wp[abs_pos], waypoints=wp, start=start, W=w, D=dmat)


def run_diffusion_maps(x: np.ndarray, n_components: int = 10, knn: int = 30, alpha: float = 0.0,
                       seed=0, tol: float = 1e-4, n_jobs: int = -1):
    """utils.py:563-577 on arrays: kernel -> T, eigenpairs."""
    kern = adaptive_kernel(x, knn, alpha, n_jobs)
    t, vec, lam = diffusion_eigs(kern, n_components, seed, tol)
    return dict(kernel=kern, T=t, EigenVectors=vec, EigenValues=lam)
