#!/usr/bin/env python
"""Benchmark of the Palantir hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells C] [--pcs D]

One "step" is one pass of the hot path (run_diffusion_maps -> determine_multiscale_space ->
run_palantir) over one synthetic branching-trajectory matrix (BASELINE.json config 3:
1,000,000 cells x 100 PCs, knn 30, 10 diffusion components, 1200 waypoints).

`value`  : cells/s with the PCA matrix already resident in HBM and all results left there
           (palantir_b200.pipeline.run_device), CUDA-event timed, max over ranks.
`e2e`    : cells/s through the public API (palantir_b200.utils / .core) with host pandas /
           numpy inputs and host scipy / pandas outputs: H2D and D2H inside the timed region.
`roofline`: the dominant kernel (FP32 kNN candidate generation) against an FFMA probe
           measured in the same run; `roofline_spmv` the Lanczos SpMV against the measured
           HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (same third-party calls as the reference) on a bounded sample.
`--impl reference`: times that CPU path alone, on the host cores, and prints its own line.

With N > 1 (torchrun, one process per GPU, NCCL) the 1M-cell job is split: kNN query rows
and shortest-path sources are sharded, neighbour lists are all-gathered and the per-cell
refinement sums all-reduced ("scaling": "strong" -- the job size is fixed).
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

METRIC = "cells/s, run_diffusion_maps + determine_multiscale_space + run_palantir at 1M cells"
UNIT = "cells/s"
KNN, NCOMP, NWP = 30, 10, 1200


def _workload_name(n, d):
    return (f"synthetic branching trajectory, {n} cells x {d} PCs float32, knn={KNN}, "
            f"{NCOMP} diffusion components, {NWP} waypoints, 3 explicit terminal states")


# --------------------------------------------------------------------------------------
# CPU arm (the reference's algorithm through the oracle port)
# --------------------------------------------------------------------------------------
def _cpu_sample(n_sample, d, n_wp):
    """One bounded CPU step: the oracle end to end on the first n_sample cells of the workload's
    generator.  Returns seconds."""
    import numpy as np

    from oracle import palantir_oracle as po
    from palantir_b200.datasets import synth_branching

    x, info = synth_branching(n_sample, d, seed=0)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = po.run_diffusion_maps(x, NCOMP, KNN)
        ms = po.multiscale_space(res["EigenVectors"], res["EigenValues"])
        po.run_palantir(ms, info["early"], terminal=list(info["terminals"].values()), num_waypoints=n_wp,
                        knn=KNN)
    return time.perf_counter() - t0


def _cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([os.cpu_count() or 1] + [p.get("num_threads", 1) for p in threadpool_info()])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


CPU_SAMPLE_CELLS, CPU_SAMPLE_WP = 12000, 150


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n_s, wp = CPU_SAMPLE_CELLS, CPU_SAMPLE_WP
    for _ in range(args.warmup):
        _cpu_sample(n_s, args.pcs, wp)
    t = [_cpu_sample(n_s, args.pcs, wp) for _ in range(args.steps)]
    sec = sum(t) / len(t)
    val = n_s / sec
    sample = (f"oracle port (sklearn/scipy/networkx, the reference's own call sites) end to end on a "
              f"{n_s}-cell x {args.pcs}-PC sample of the workload's generator with {wp} waypoints; CPU cost grows "
              f"super-linearly in cells (kNN ~N^2, Dijkstra ~nW*N), so this flatters the CPU at 1M cells")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_name(args.cells, args.pcs), "timed_sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": _cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


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
    n, d = args.cells, args.pcs
    x, info = synth_branching(n, d, seed=0)
    names = synth_names(n)
    term_pos = list(info["terminals"].values())
    x_dev = dv.to_device(x)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return pipeline.run_device(x_dev, info["early"], term_pos, KNN, NCOMP, 0.0, NWP)

    # ---- device-resident throughput (`value`) ----
    for _ in range(args.warmup):
        out = device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dv.PROFILE = {}
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = device_step()
    e1.record()
    barrier()
    launches = _native.launch_count() - l0
    sssp_info = dict(core.LAST_INFO.get("sssp", {}))
    dv.flush_profile()
    prof = {k: v / args.steps for k, v in dv.PROFILE.items()}
    dv.PROFILE = None
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item())
    clocks = sampler.stop() if rank == 0 else {}

    # ---- end to end through the public API (`e2e`) ----
    # the step's input matrix lies in page-locked host memory (bench contract); everything else the API
    # touches on the host is ordinary pandas / scipy / numpy
    if world > 1:
        palantir_config.HOST_SPARSE_ON_ALL_RANKS = False      # rank 0 materialises the two sparse matrices
    x_host = torch.from_numpy(x).pin_memory().numpy()
    df = pd.DataFrame(x_host, index=names, copy=False)
    term = {names[v]: k for k, v in info["terminals"].items()}

    def api_step():
        with contextlib.redirect_stdout(io.StringIO()):
            res = utils.run_diffusion_maps(df, n_components=NCOMP, knn=KNN, kernel_backend="sklearn")
            msd = utils.determine_multiscale_space(res)
            pr = core.run_palantir(msd, names[info["early"]], terminal_states=term, num_waypoints=NWP, knn=KNN)
        return res, msd, pr

    api_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 2))
    res = msd = pr = None
    for _ in range(n_e2e):
        res = msd = pr = None                  # release the previous step's (page-locked) result buffers
        res, msd, pr = api_step()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_sec = float(e2e_s.item())
    kern, tmat = res["kernel"], res["T"]
    h2d = x.nbytes + msd.values.nbytes + n * 8                      # PCA matrix, multiscale matrix, Lanczos start
    sparse_bytes = 0 if kern is None else (kern.data.nbytes + kern.indices.nbytes + kern.indptr.nbytes +
                                           tmat.data.nbytes + tmat.indices.nbytes + tmat.indptr.nbytes)
    d2h = (sparse_bytes + res["EigenVectors"].values.nbytes + pr.pseudotime.values.nbytes +
           pr.entropy.values.nbytes + pr.branch_probs.values.nbytes)      # rank 0's bytes are the ones printed

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
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    scr = dv.empty((16,), dv.F32)
    for _ in range(2):
        dv.call("pb200_fma_peak_f32", scr, 148 * 8, 2048)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(); dv.call("pb200_fma_peak_f32", scr, 148 * 8, 16384); p1.record(); torch.cuda.synchronize()
    fp32_peak = 148 * 8 * 1024 * 16384 * 16 * 2 / p0.elapsed_time(p1) / 1e9          # TFLOP/s
    nq_rank = -(-n // world)
    knn_ms = prof.get("knn_candidates", float("nan"))
    knn_alg_tf = 2.0 * nq_rank * n * d / knn_ms / 1e9 if knn_ms == knn_ms else None
    kinfo = out["knn"] or {}
    visited = kinfo.get("tiles_visited_frac") or 1.0
    engine = kinfo.get("engine", "simt")
    if engine == "tc":
        # executed tensor work: three TF32 MMAs per product, coordinates padded to a multiple of 8
        d8 = -(-d // 8) * 8
        knn_tf = 3.0 * 2.0 * nq_rank * n * d8 * visited / knn_ms / 1e9 if knn_alg_tf else None
        knn_peak = float(peaks.get("bf16_tflops", 1590.0)) / 2.0
        knn_roof = {"bound": "tensor", "kernel": "knn_candidates_tc_kernel (tcgen05.mma kind::tf32, 3xTF32 split)",
                    "achieved": knn_tf, "peak": knn_peak, "unit": "TFLOP/s", "frac": knn_tf / knn_peak if knn_tf else None,
                    # DRAM bytes per launch from the ncu --set full capture of this kernel on this workload
                    # (profiles/r1_knn_candidates_ncu.md, knn_tc_final: 151.0 GB read + 0.35 GB written)
                    "traffic": 1.513e11 if (n == 1_000_000 and d == 100 and world == 1) else None,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the BF16 rate; of measured)"
                                    if "bf16_tflops" in peaks else "B200_PROFILING.md fallback 1590/2 TF/s (of fallback)"),
                    "algorithmic": "3 x 2*Nq*Nr*d8 TF32 FLOP per launch x fraction of 128x128 tiles visited (exact "
                                   "projection pruning skips the rest); 'effective' = 2*Nq*Nr*d / time, all tiles counted",
                    "tiles_visited_frac": visited, "effective_tflops": knn_alg_tf, "ms_per_launch": knn_ms,
                    "fp32_simt_probe_tflops": fp32_peak}
    else:
        knn_tf = knn_alg_tf * visited if knn_alg_tf else None          # FLOPs actually executed
        knn_roof = {"bound": "fp32 (SIMT FFMA; schema value would be 'tensor')", "kernel": "knn_candidates_kernel",
                    "achieved": knn_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": (knn_tf / fp32_peak) if knn_tf else None, "traffic": None,
                    "peak_source": "FFMA probe pb200_fma_peak_f32 timed in this run (MEASURED_PEAKS.json has no FP32 "
                                   "figure)",
                    "algorithmic": "2*Nq*Nr*d FLOP per launch x fraction of 256x128 tiles visited (exact projection "
                                   "pruning skips the rest); 'effective' counts the skipped tiles too",
                    "tiles_visited_frac": visited, "effective_tflops": knn_alg_tf, "ms_per_launch": knn_ms}
    # SpMV timed alone on the operator of the last step, in the row order the eigensolver uses (relabelled
    # into the kNN stage's Morton order when that order is known)
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

    def spmv():
        dv.call("pb200_csr_spmv_split_f64", sp_indptr, sp_indices, sp_vals, n, kd.nnz,
                long_rows if long_rows.numel() else None, int(long_rows.numel()), xs, ys)

    for _ in range(5):
        spmv()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    s0.record()
    for _ in range(reps):
        spmv()
    s1.record(); torch.cuda.synchronize()
    spmv_ms = s0.elapsed_time(s1) / reps
    spmv_bytes = kd.nnz * 12 + (n + 1) * 4 + n * 8 + n * 8
    spmv_gbs = spmv_bytes / spmv_ms / 1e6
    sssp_ms = prof.get("sssp")
    arcs_s = (sssp_info.get("sources", 0) * sssp_info.get("arcs", 0) / (sssp_ms / 1e3)) if sssp_ms else None

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        sec = _cpu_sample(CPU_SAMPLE_CELLS, d, CPU_SAMPLE_WP)
        cpu = {"value": CPU_SAMPLE_CELLS / sec, "unit": UNIT, "cores": _cpu_threads(), "kind": "port",
               "sample": (f"oracle port end to end on a {CPU_SAMPLE_CELLS}-cell x {d}-PC sample with {CPU_SAMPLE_WP} "
                          f"waypoints ({sec:.1f} s); the CPU cost is super-linear in cells, so cells/s at 1M is lower")}

    line = {
        "metric": METRIC, "value": n / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 kNN candidates; f64 everything else", "data": "synthetic",
        "config": {"workload": _workload_name(n, d), "l2": "inputs larger than L2 (PCA matrix 400 MB, kernel 0.6 GB, "
                   "distance matrix 9.6 GB)", "parallelism": f"kNN rows and SSSP sources sharded over {world} GPU(s)"
                   + ("; e2e: kernel and T materialised on the host by rank 0 only (config.HOST_SPARSE_ON_ALL_RANKS = False)" if world > 1 else "")},
        "e2e": {"value": n / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_sec * 1e3, "steps": n_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": knn_roof,
        "roofline_spmv": {"bound": "hbm", "kernel": "csr_spmv_kernel<16> + csr_spmv_long_rows_kernel (rows above 128 entries)", "achieved": spmv_gbs, "peak": hbm_peak,
                          "unit": "GB/s", "frac": spmv_gbs / hbm_peak, "traffic": None, "peak_source": hbm_src,
                          "algorithmic": "12*nnz + 4*(N+1) + 16*N bytes per launch", "ms_per_launch": spmv_ms,
                          "row_order": spmv_order},
        "sssp": {"arcs_per_s": arcs_s, "ms": sssp_ms, **{k: v for k, v in sssp_info.items() if k != "knn"}},
        "stages_ms": {k: round(v, 3) for k, v in sorted(prof.items())},
        "lanczos": utils.LAST_INFO.get("lanczos"),
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
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1_000_000)
    ap.add_argument("--pcs", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    out = _quiet_stdout()
    line = run_reference_arm(args) if args.impl == "reference" else run_gpu_arm(args)
    if line is not None:
        out.write(json.dumps(line) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
