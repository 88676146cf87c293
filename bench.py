#!/usr/bin/env python
"""Benchmark of the Palantir hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c2|c4|c5]

One "step" is one pass of the hot path (run_diffusion_maps -> determine_multiscale_space ->
run_palantir) over one synthetic branching-trajectory matrix.  The default configuration is
BASELINE.json config 3 (1,000,000 cells x 100 PCs, knn 30, 10 diffusion components, 1200 waypoints);
--config c1 is the tutorial-sized case (4142 x 100, 500 waypoints), c2 config 2 (100 k x 50), c4 config 4 (4 M x 50, 5000 waypoints: meant for 8 GPUs), c5
config 5 (run_magic_imputation, t = 3, 1 M cells x 2000 genes).

`value`   : cells/s with the PCA matrix already resident in HBM and all results left there
            (palantir_b200.pipeline.run_device), CUDA-event timed over exactly K steps, max over ranks.
`e2e`     : the same metric through the public API (palantir_b200.utils / .core) with ordinary (pageable) host
            pandas inputs and host scipy / pandas outputs: every H2D and D2H copy is inside the timed region.
`roofline`: the kernel with the largest share of the step against its bound; `roofline_*`: the other kernels
            BASELINE.json names (SpMV and kernel build against HBM, shortest paths as arcs/s and bytes floor).
`cpu_baseline`: the reference's algorithm (oracle port: the same sklearn / scipy / networkx calls at the same
            call sites, pinned bit for bit against the reference source) on the host cores: one full step of
            config 2, its stage times, and the config-3 figure extrapolated from them (labelled as such).
`same_config_as_reference_arm`: this arm's value / e2e on the bounded sample `--impl reference` times for the
            same --steps / --warmup, so that the two lines can be compared like for like.
`--impl reference`: times that CPU path alone and prints its own line.

With N > 1 (torchrun, one process per GPU, NCCL) the job is split: kNN query rows (cut at equal work) and
shortest-path sources are sharded, neighbour lists are all-gathered, the eigensolver's product is sharded by rows
(one all-gather of an N-vector per step) and the per-cell refinement sums all-reduced ("scaling": "strong" -- the
job size is fixed).  `traffic` figures are per-launch DRAM bytes read from the committed ncu summaries (NCU_TRAFFIC_C3).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

UNIT = "cells/s"
KNN, NCOMP = 30, 10
CONFIGS = {
    # name: (cells, PCs, waypoints)
    "c1": (4_142, 100, 500),          # the size of the reference's tutorial data set (config 1; synthetic stand-in)
    "c2": (100_000, 50, 1200),
    "c3": (1_000_000, 100, 1200),
    "c4": (4_000_000, 50, 5000),
}
MAGIC_GENES, MAGIC_STEPS = 2000, 3


def _metric(cfg):
    if cfg == "c5":
        return "cells/s, run_magic_imputation t=3 at 1M cells x 2000 genes"
    n = CONFIGS[cfg][0]
    label = {4_142: "4142", 100_000: "100k", 1_000_000: "1M", 4_000_000: "4M"}[n]
    return f"cells/s, run_diffusion_maps + determine_multiscale_space + run_palantir at {label} cells"


def _workload_name(n, d, nwp):
    return (f"synthetic branching trajectory, {n} cells x {d} PCs float32, knn={KNN}, "
            f"{NCOMP} diffusion components, {nwp} waypoints, 3 explicit terminal states")


# --------------------------------------------------------------------------------------
# CPU arm (the reference's algorithm through the oracle port)
# --------------------------------------------------------------------------------------
# bounded samples of the workload's generator for --impl reference: (cells, waypoints, seconds per step on the
# 8-core build container with d = 100).  The arm takes the largest one that fits its time budget.
REF_LADDER = [(100_000, 1200, 150.0), (50_000, 600, 45.0), (25_000, 300, 14.0), (12_000, 150, 5.0), (6_000, 80, 2.0)]
REF_BUDGET_S = 200.0


def _ladder_choice(steps, warmup):
    per_step = REF_BUDGET_S / max(1, steps + warmup)
    for n, wp, sec in REF_LADDER:
        if sec <= per_step:
            return n, wp
    return REF_LADDER[-1][:2]


def _cpu_step(n_cells, d, n_wp, stage_times=None):
    """One CPU step: the oracle end to end on the first n_cells cells of the workload's generator.  Returns
    seconds; with stage_times (a dict) the kNN / eigensolver / Dijkstra / rest split is recorded."""
    from oracle import palantir_oracle as po
    from palantir_b200.datasets import synth_branching

    x, info = synth_branching(n_cells, d, seed=0)
    timers = {"knn": 0.0, "eigs": 0.0, "dijkstra": 0.0}
    saved = (po.knn_with_self, po.eigs, po.csgraph.dijkstra)

    def wrap(name, fn):
        def inner(*a, **k):
            t = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                timers[name] += time.perf_counter() - t
        return inner

    class _Csgraph:
        def __getattr__(self, item):
            return getattr(saved_csgraph, item)

    saved_csgraph = po.csgraph
    proxy = _Csgraph()
    proxy.dijkstra = wrap("dijkstra", saved_csgraph.dijkstra)
    if stage_times is not None:
        po.knn_with_self = wrap("knn", saved[0])
        po.eigs = wrap("eigs", saved[1])
        po.csgraph = proxy
    t0 = time.perf_counter()
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            res = po.run_diffusion_maps(x, NCOMP, KNN)
            ms = po.multiscale_space(res["EigenVectors"], res["EigenValues"])
            po.run_palantir(ms, info["early"], terminal=list(info["terminals"].values()), num_waypoints=n_wp, knn=KNN)
    finally:
        po.knn_with_self, po.eigs, po.csgraph = saved[0], saved[1], saved_csgraph
    total = time.perf_counter() - t0
    if stage_times is not None:
        stage_times.update(timers, total=total, rest=total - sum(timers.values()))
    return total


def _cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([os.cpu_count() or 1] + [p.get("num_threads", 1) for p in threadpool_info()])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def _extrapolate_c3(st, n2=100_000, d2=50, wp2=1200, n3=1_000_000, d3=100, wp3=1200):
    """Config-3 CPU seconds from the measured config-2 stage times (the reference cannot finish a
    1 M-cell run in a benchmark): brute kNN ~ N^2 d, ARPACK ~ nnz ~ N, Dijkstra ~ nW N log N, the rest (row-wise
    entropy apply, graph building, refinement) ~ N."""
    import math

    f_knn = (n3 / n2) ** 2 * (d3 / d2)
    f_lin = n3 / n2
    f_dij = (wp3 / wp2) * (n3 / n2) * math.log(n3) / math.log(n2)
    sec = st["knn"] * f_knn + st["eigs"] * f_lin + st["dijkstra"] * f_dij + st["rest"] * f_lin
    return sec, {"knn": f_knn, "eigs": f_lin, "dijkstra": f_dij, "rest": f_lin}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n, d, nwp = CONFIGS[args.config if args.config in CONFIGS else "c3"]
    n_s, wp = _ladder_choice(args.steps, args.warmup)
    for _ in range(args.warmup):
        _cpu_step(n_s, d, wp)
    t = [_cpu_step(n_s, d, wp) for _ in range(args.steps)]
    sec = sum(t) / len(t)
    val = n_s / sec
    sample = (f"oracle port (sklearn/scipy/networkx at the reference's own call sites, pinned bit for bit against the "
              f"reference source) end to end on the first {n_s} cells x {d} PCs of the workload's generator with {wp} "
              f"waypoints -- the largest sample whose {args.steps}+{args.warmup} steps fit {REF_BUDGET_S:.0f} s; the "
              f"b200 arm reports its own numbers on this same sample under 'same_config_as_reference_arm'")
    return {
        "impl": "reference", "metric": _metric(args.config if args.config in CONFIGS else "c3"), "value": val,
        "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_name(n, d, nwp), "timed_sample": sample,
                   "sample_cells": n_s, "sample_waypoints": wp},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": _cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            load = sorted(s for s, pw in zip(sm, power) if pw > 0.5 * max(power)) or sorted(sm)
            out = {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power)}
        return out


STAGES = ["knn_total", "knn_balance", "knn_candidates", "knn_rerank", "kernel_build", "eigensolver", "lanczos_steps", "ritz_vectors",
          "maxmin_sampling", "knn_direct", "graph_build", "sssp_partition", "sssp_tables", "sssp_overlay", "sssp",
          "sssp_expand", "refine", "markov_chain", "absorption_solve", "project"]
SSSP_STAGES = ("sssp_partition", "sssp_tables", "sssp_overlay", "sssp", "sssp_expand")

# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` captures of
# these kernels on the config-3 workload at one GPU (profiles/): kernel -> (bytes, file)
NCU_TRAFFIC_C3 = {
    "knn_candidates_tc_kernel": (1.706e11, "profiles/r2_knn_tc_ncu.md"),
    "csr_spmv_stream_kernel": (6.617e8, "profiles/r2_spmv_ncu.md"),
    "sssp_cells_stage": (1.931e10, "profiles/r2_sssp_cells_ncu.md"),     # grow + local + overlay search + expansion
}


def _time_device(fn, reps):
    import torch

    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_gpu_arm(args):
    import numpy as np
    import pandas as pd
    import torch
    import torch.distributed as dist

    from palantir_b200 import _device as dv
    from palantir_b200 import _native, core, pipeline, utils
    from palantir_b200 import config as palantir_config
    from palantir_b200.datasets import synth_branching, synth_names

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, d, nwp = CONFIGS[args.config]
    if args.cells:
        n = args.cells
    if args.pcs:
        d = args.pcs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(n_cells, n_pcs, n_wp, steps, warmup, e2e_steps, profile):
        """value (device-resident), stage profile and e2e (public API, pageable host input) on one workload."""
        x, info = synth_branching(n_cells, n_pcs, seed=0)
        names = synth_names(n_cells)
        term_pos = list(info["terminals"].values())
        x_dev = dv.to_device(x)
        out = None
        for _ in range(warmup):
            out = pipeline.run_device(x_dev, info["early"], term_pos, KNN, NCOMP, 0.0, n_wp)
        barrier()
        if profile:
            dv.PROFILE = {}
        l0 = _native.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = pipeline.run_device(x_dev, info["early"], term_pos, KNN, NCOMP, 0.0, n_wp)
        e1.record()
        barrier()
        launches = _native.launch_count() - l0
        sssp_info = dict(core.LAST_INFO.get("sssp", {}))
        prof = {}
        if profile:
            dv.flush_profile()
            vec = torch.tensor([dv.PROFILE.get(k, 0.0) / steps for k in STAGES], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(vec, op=dist.ReduceOp.MAX)        # a stage costs what its slowest rank takes
            prof = {k: float(v) for k, v in zip(STAGES, vec.tolist()) if v > 0.0}
            dv.PROFILE = None
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = float(ms.item())
        del x_dev
        # ---- end to end through the public API: ordinary host objects in, ordinary host objects out ----
        df = pd.DataFrame(x, index=names, copy=False)            # pageable memory, as a caller's frame is
        term = {names[v]: k for k, v in info["terminals"].items()}

        def api_step():
            with contextlib.redirect_stdout(io.StringIO()):
                res = utils.run_diffusion_maps(df, n_components=NCOMP, knn=KNN, kernel_backend="sklearn")
                msd = utils.determine_multiscale_space(res)
                pr = core.run_palantir(msd, names[info["early"]], terminal_states=term, num_waypoints=n_wp, knn=KNN)
            return res, msd, pr

        res = msd = pr = None
        e2e_sec, h2d, d2h = None, 0, 0
        if e2e_steps > 0:
            api_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                res = msd = pr = None                  # release the previous step's (page-locked) result buffers
                res, msd, pr = api_step()
            barrier()
            t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_sec = float(t.item())
            kern, tmat = res["kernel"], res["T"]
            # PCA matrix, eigenvectors (determine_multiscale_space) and multiscale matrix (run_palantir) go up
            # again with every call; the Lanczos start vector is drawn on the host
            h2d = x.nbytes + msd.values.nbytes + n_cells * 8
            sparse_bytes = 0 if kern is None else (kern.data.nbytes + kern.indices.nbytes + kern.indptr.nbytes +
                                                   tmat.data.nbytes + tmat.indices.nbytes + tmat.indptr.nbytes)
            d2h = (sparse_bytes + res["EigenVectors"].values.nbytes + pr.pseudotime.values.nbytes +
                   pr.entropy.values.nbytes + pr.branch_probs.values.nbytes)
        # size-independent properties of the last device-resident result (not timed)
        from scipy.stats import spearmanr

        pt = out["pseudotime"].cpu().numpy()
        fp = out["fate_probs"].cpu().numpy()
        stride = max(1, n_cells // 200_000)
        sanity = {"pseudotime_range": [float(pt.min()), float(pt.max())],
                  "spearman_vs_latent_time": float(spearmanr(pt[::stride], info["time"][::stride])[0]),
                  "fate_rows_sum_max_dev": float(np.abs(fp.sum(1) - 1.0).max()) if fp.shape[1] else None,
                  "own_fate_of_terminal_cells": [float(fp[p].max()) for p in term_pos] if fp.shape[1] else None,
                  "knn_rows_needing_exact_fallback": int((out.get("knn") or {}).get("flagged", 0)),
                  "eigenvalues": [float(v) for v in np.asarray(out["eigenvalues"])[:4]]}
        return dict(ms_step=ms_step, prof=prof, launches=launches, sssp_info=sssp_info, out=out, e2e_sec=e2e_sec,
                    h2d=int(h2d), d2h=int(d2h), n=n_cells, d=n_pcs, nwp=n_wp, sanity=sanity)

    if world > 1:
        palantir_config.HOST_SPARSE_ON_ALL_RANKS = False      # rank 0 materialises the two sparse matrices
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    main = measure(n, d, nwp, args.steps, args.warmup, max(1, args.e2e_steps), True)
    clocks = sampler.stop() if rank == 0 else {}
    out = main["out"]
    prof = main["prof"]
    ms_step = main["ms_step"]

    # the same arm on the bounded sample the reference arm times for these --steps / --warmup
    n_ref, wp_ref = _ladder_choice(args.steps, args.warmup)
    same = None
    if world == 1 and not args.no_same_config:
        sm = measure(n_ref, d, wp_ref, 5, 3, 5, False)
        same = {"workload": _workload_name(n_ref, d, wp_ref), "value": n_ref / (sm["ms_step"] / 1e3), "unit": UNIT,
                "ms_per_step": sm["ms_step"], "e2e": {"value": n_ref / sm["e2e_sec"], "ms_per_step": sm["e2e_sec"] * 1e3},
                "note": "bench.py --impl reference times the CPU path on exactly this sample"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return None

    # ---- rooflines ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = ("MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks
               else "B200_PROFILING.md fallback 6650 GB/s (of fallback)")
    # TF32 tensor peak: bare tcgen05.mma.kind::tf32 128x128x8 instructions, one CTA per SM (pb200_tf32_peak)
    it = 200_000
    tf32_ms = _time_device(lambda: dv.call("pb200_tf32_peak", 148, it), 3)
    tf32_peak = 148 * it * 262144 / tf32_ms / 1e9                         # TFLOP/s
    scr = dv.empty((16,), dv.F32)
    fp32_ms = _time_device(lambda: dv.call("pb200_fma_peak_f32", scr, 148 * 8, 16384), 2)
    fp32_peak = 148 * 8 * 1024 * 16384 * 16 * 2 / fp32_ms / 1e9          # TFLOP/s
    nq_rank = -(-n // world)
    knn_ms = prof.get("knn_candidates", float("nan"))
    knn_alg_tf = 2.0 * nq_rank * n * d / knn_ms / 1e9 if knn_ms == knn_ms else None
    kinfo = out["knn"] or {}
    visited = kinfo.get("tiles_visited_frac") or 1.0
    engine = kinfo.get("engine", "simt")
    if engine == "tc":
        d8 = -(-d // 8) * 8
        knn_tf = 3.0 * 2.0 * nq_rank * n * d8 * visited / knn_ms / 1e9 if knn_alg_tf else None
        traffic = NCU_TRAFFIC_C3["knn_candidates_tc_kernel"] if (n == 1_000_000 and d == 100 and world == 1) else (None, None)
        knn_roof = {"bound": "tensor", "kernel": "knn_candidates_tc_kernel (tcgen05.mma kind::tf32, 3xTF32 split)",
                    "achieved": knn_tf, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": knn_tf / tf32_peak if knn_tf else None,
                    "traffic": traffic[0], "traffic_source": traffic[1],
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                    "peak_source": "measured in this run: pb200_tf32_peak, 148 CTAs x bare tcgen05.mma.cta_group::1."
                                   "kind::tf32 128x128x8 (MEASURED_PEAKS.json has no TF32 figure; its bf16_tflops / 2 = "
                                   f"{float(peaks.get('bf16_tflops', 1590.0)) / 2:.0f})",
                    "algorithmic": "3 x 2*Nq*Nr*d8 TF32 FLOP per launch x fraction of 128x128 tiles visited (exact "
                                   "projection pruning skips the rest); 'effective' = 2*Nq*Nr*d / time, all tiles counted",
                    "tiles_visited_frac": visited, "effective_tflops": knn_alg_tf, "ms_per_launch": knn_ms,
                    "share_of_step": knn_ms / ms_step, "fp32_simt_probe_tflops": fp32_peak}
    else:
        knn_tf = knn_alg_tf * visited if knn_alg_tf else None
        knn_roof = {"bound": "fp32 (SIMT FFMA; schema value would be 'tensor')", "kernel": "knn_candidates_kernel",
                    "achieved": knn_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": (knn_tf / fp32_peak) if knn_tf else None, "traffic": None,
                    "peak_source": "FFMA probe pb200_fma_peak_f32 timed in this run",
                    "tiles_visited_frac": visited, "effective_tflops": knn_alg_tf, "ms_per_launch": knn_ms,
                    "share_of_step": knn_ms / ms_step}

    # SpMV timed alone on the operator of the last step, in the row order the eigensolver uses
    kd = out["kernel"]
    sp_indptr, sp_indices, sp_vals = kd.indptr, kd.indices, kd.data
    spmv_order = "input order"
    if kd.order is not None and dv.LANCZOS_WINDOWED:
        pl = kd.order.long()
        inv = torch.empty_like(kd.order)
        inv[pl] = torch.arange(n, dtype=kd.order.dtype, device=kd.order.device)
        sp_indptr = torch.zeros((n + 1,), dtype=torch.int32, device="cuda")
        sp_indptr[1:] = torch.cumsum((kd.indptr[1:] - kd.indptr[:-1])[pl], 0)
        sp_indices, sp_vals = dv.empty((kd.nnz,), dv.I32), dv.empty((kd.nnz,), dv.F64)
        dv.call("pb200_csr_permute", kd.indptr, kd.indices, kd.data, n, kd.order, inv, sp_indptr, sp_indices, sp_vals)
        spmv_order = "Morton order of the kNN stage"
    xs = torch.rand(n, dtype=torch.float64, device="cuda")
    ys = torch.empty_like(xs)
    long_rows = torch.nonzero((sp_indptr[1:] - sp_indptr[:-1]) > int(_native.load().pb200_spmv_long_row()))
    long_rows = long_rows.to(torch.int32).reshape(-1).contiguous()
    spmv_rows_ms = _time_device(lambda: dv.call("pb200_csr_spmv_split_f64", sp_indptr, sp_indices, sp_vals, n, kd.nnz,
                                                long_rows if long_rows.numel() else None, int(long_rows.numel()), xs, ys),
                                50)
    spmv_ms, spmv_kernel = spmv_rows_ms, "csr_spmv_kernel<16> + csr_spmv_long_rows_kernel (rows above 128 entries)"
    if dv.SPMV_STREAM:
        # the product the eigensolver runs: one CTA per 1024 consecutive nonzeros
        tile = int(_native.load().pb200_csr_spmv_stream_tile())
        n_tiles = -(-kd.nnz // tile)
        tile_lo = dv.empty((n_tiles + 1,), dv.I32)
        st_head, st_tail = dv.empty((n_tiles,), dv.F64), dv.empty((n_tiles,), dv.F64)
        dv.call("pb200_csr_spmv_stream_prepare", sp_indptr, n, kd.nnz, tile_lo)
        spmv_ms = _time_device(lambda: dv.call("pb200_csr_spmv_stream_f64", sp_indptr, sp_indices, sp_vals, n, kd.nnz,
                                               tile_lo, st_head, st_tail, xs, ys, 0, n, 0, n_tiles), 50)
        spmv_kernel = "csr_spmv_stream_kernel (a CTA per 1024 consecutive nonzeros) + csr_stream_fixup_kernel"
    spmv_bytes = kd.nnz * 12 + (n + 1) * 4 + n * 8 + n * 8
    spmv_gbs = spmv_bytes / spmv_ms / 1e6
    # kernel build: read idx + dist, write K (values, columns, row pointers)
    kb_ms = prof.get("kernel_build")
    kb_bytes = n * (KNN + 1) * 12 + kd.nnz * 12 + (n + 1) * 4
    # shortest paths: arcs/s and the bytes floor of SURVEY.md 8d (every arc once per source, the output once)
    sssp_info = main["sssp_info"]
    sssp_ms = sum(prof.get(k, 0.0) for k in SSSP_STAGES)
    n_src, n_arcs = sssp_info.get("sources", 0), sssp_info.get("arcs", 0)
    sssp_floor_bytes = n_src * n_arcs * 12 + n_src * n * 8
    sssp_roof = None
    if sssp_ms > 0:
        sssp_roof = {"bound": "hbm", "kernel": "cell-decomposed search (cells_grow / cells_local_sssp / sssp_group on the "
                     "overlay / cells_expand)" if sssp_info.get("cells") else "sssp_group_kernel (per source)",
                     "achieved": sssp_floor_bytes / sssp_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                     "frac": sssp_floor_bytes / sssp_ms / 1e6 / hbm_peak, "peak_source": hbm_src,
                     "algorithmic": "sources x arcs x 12 B + sources x N x 8 B (SURVEY.md 8d bytes floor of a per-source "
                                    "search; the cell decomposition reads far fewer bytes than that floor)",
                     "arcs_per_s": n_src * n_arcs / (sssp_ms / 1e3), "ms": sssp_ms, "share_of_step": sssp_ms / ms_step,
                     "parts_ms": {k: prof.get(k) for k in SSSP_STAGES if k in prof},
                     "overlay_rounds_mean": sssp_info.get("rounds_mean"), "sources": n_src, "arcs": n_arcs,
                     "cells": sssp_info.get("cells"),
                     "traffic": (NCU_TRAFFIC_C3["sssp_cells_stage"][0]
                                 if (sssp_info.get("cells") and n == 1_000_000 and d == 100 and world == 1) else None),
                     "traffic_source": (NCU_TRAFFIC_C3["sssp_cells_stage"][1]
                                        if (sssp_info.get("cells") and n == 1_000_000 and d == 100 and world == 1) else None)}

    # the kernel with the largest share of the step is the line's `roofline`
    shares = {"knn": knn_ms if knn_ms == knn_ms else 0.0, "sssp": sssp_ms,
              "spmv": spmv_ms * (utils.LAST_INFO.get("lanczos") or {}).get("steps", 0)}
    dominant = max(shares, key=shares.get)
    spmv_roof = {"bound": "hbm", "kernel": spmv_kernel,
                 "achieved": spmv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_gbs / hbm_peak,
                 "traffic": (NCU_TRAFFIC_C3["csr_spmv_stream_kernel"][0]
                             if (dv.SPMV_STREAM and n == 1_000_000 and d == 100) else None),
                 "traffic_source": (NCU_TRAFFIC_C3["csr_spmv_stream_kernel"][1]
                                    if (dv.SPMV_STREAM and n == 1_000_000 and d == 100) else None),
                 "peak_source": hbm_src, "algorithmic": "12*nnz + 4*(N+1) + 16*N bytes per launch",
                 "ms_per_launch": spmv_ms, "row_order": spmv_order, "row_kernels_ms_per_launch": spmv_rows_ms,
                 "share_of_step": shares["spmv"] / ms_step,
                 "note": "bound by the load/store unit: 53 M scattered 8-byte gathers of x, one address per cycle per "
                         "SM = 0.18 ms at 148 SMs x 1.965 GHz, whatever the CSR stream costs"}
    roofline = {"knn": knn_roof, "sssp": sssp_roof, "spmv": spmv_roof}[dominant]

    # ---- CPU baseline (rank 0, N == 1 only): one full step of config 2 + the config-3 extrapolation ----
    cpu = None
    if world == 1 and not args.no_cpu:
        st = {}
        n2, d2, wp2 = CONFIGS["c2"]
        sec = _cpu_step(n2, d2, wp2, st)
        c3_sec, factors = _extrapolate_c3(st)
        cpu = {"value": n2 / sec, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
               "sample": (f"one full step of config 2 ({n2} cells x {d2} PCs, {wp2} waypoints) through the oracle port "
                          f"({sec:.1f} s); the b200 arm on the same config is under 'config2'"),
               "stages_s": {k: round(v, 2) for k, v in st.items()},
               "config3_extrapolated": {"value": CONFIGS["c3"][0] / c3_sec, "unit": UNIT, "seconds": c3_sec,
                                        "how": "EXTRAPOLATED, not measured: config-2 stage times x complexity ratios "
                                               "(brute kNN ~ N^2 d, ARPACK ~ N, Dijkstra ~ nW N log N, rest ~ N)",
                                        "factors": factors}}
    config2 = None
    if world == 1 and not args.no_cpu and args.config == "c3":
        n2, d2, wp2 = CONFIGS["c2"]
        m2 = measure(n2, d2, wp2, 5, 3, 5, False)
        config2 = {"workload": _workload_name(n2, d2, wp2), "value": n2 / (m2["ms_step"] / 1e3), "unit": UNIT,
                   "ms_per_step": m2["ms_step"], "e2e": {"value": n2 / m2["e2e_sec"], "ms_per_step": m2["e2e_sec"] * 1e3}}

    line = {
        "metric": _metric(args.config), "value": n / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32x3 kNN candidates (f64 re-rank); f64 everything else", "data": "synthetic",
        "config": {"workload": _workload_name(n, d, nwp),
                   "l2": "inputs larger than L2 (PCA matrix 400 MB, kernel 0.6 GB, distance matrix 9.6 GB at config 3)",
                   "parallelism": f"kNN rows and SSSP sources sharded over {world} GPU(s)"
                   + ("; e2e: kernel and T materialised on the host by rank 0 only "
                      "(config.HOST_SPARSE_ON_ALL_RANKS = False)" if world > 1 else "")},
        "e2e": {"value": n / main["e2e_sec"], "unit": UNIT, "h2d_bytes_per_step": main["h2d"],
                "d2h_bytes_per_step": main["d2h"], "ms_per_step": main["e2e_sec"] * 1e3, "steps": max(1, args.e2e_steps),
                "input": "pageable pandas DataFrame; scipy / pandas results on the host"},
        "gpu_launches": int(main["launches"]),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_knn": knn_roof if dominant != "knn" else None,
        "roofline_spmv": spmv_roof if dominant != "spmv" else None,
        "roofline_sssp": sssp_roof if dominant != "sssp" else None,
        "roofline_kernel_build": None if not kb_ms else {
            "bound": "hbm", "kernel": "adaptive_weights + symmetrize (merge_rows & co.)", "achieved": kb_bytes / kb_ms / 1e6,
            "peak": hbm_peak, "unit": "GB/s", "frac": kb_bytes / kb_ms / 1e6 / hbm_peak, "ms": kb_ms,
            "algorithmic": "N*(k+1)*12 B read + nnz*12 + 4*(N+1) B written", "peak_source": hbm_src},
        "sanity": main["sanity"],
        "memory_gib": {"peak_allocated_rank0": torch.cuda.max_memory_allocated() / 2 ** 30,
                       "distance_rows_per_gpu": main["sssp_info"].get("sources", 0) * n * 8 / 2 ** 30},
        "stages_ms": {k: round(v, 3) for k, v in sorted(prof.items())},
        "stages_note": "max over ranks per stage",
        "lanczos": {k: v for k, v in (utils.LAST_INFO.get("lanczos") or {}).items() if k != "history"},
        "same_config_as_reference_arm": same,
        "config2": config2,
        "cpu_baseline": cpu,
    }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def _quiet_stdout():
    """Route everything the run writes to fd 1 (NCCL's version banner, library prints) to stderr;
    returns a file object on the real stdout for the single JSON line."""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--cells", type=int, default=0, help="override the configuration's cell count")
    ap.add_argument("--pcs", type=int, default=0, help="override the configuration's PC count")
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the end-to-end (public API) measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / config2 legs")
    ap.add_argument("--no-same-config", action="store_true", help="skip the reference-arm-sample leg")
    args = ap.parse_args()
    out = _quiet_stdout()
    if args.config == "c5":
        import bench_c5

        line = bench_c5.run(args)
    else:
        line = run_reference_arm(args) if args.impl == "reference" else run_gpu_arm(args)
    if line is not None:
        out.write(json.dumps(line) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
