"""bench.py --config c5: run_magic_imputation, t = 3, on 1 M cells x 2000 genes (BASELINE.json config 5).

The diffusion operator T comes from one run of the diffusion-map stage on the config-3 input (1 M cells x 100 PCs);
the expression matrix is `rng.random((N, G), float32)`.  A step = the three successive SpMMs + clip.

`value`: cells/s with X resident in HBM (one device-resident step = palantir_b200.utils.magic_device on this
         rank's gene columns); with N ranks the gene columns are split, T replicated, no collective.
`e2e`  : palantir_b200.utils.run_magic_imputation(X_host, dm_res, sparse=False): host ndarray in, host ndarray out.
`roofline`: the SpMM against HBM by the bytes floor of SURVEY.md 8d, t*(2*N*G*4 + nnz*8) per step; the kernel's real
         bound is the gather traffic from L2 (nnz*G*4 B per product), reported beside it.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def run(args):
    import pandas as pd
    import torch
    import torch.distributed as dist

    import bench
    from palantir_b200 import _device as dv
    from palantir_b200 import _dist, _native, utils
    from palantir_b200 import config as palantir_config
    from palantir_b200.datasets import synth_branching, synth_names

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return None
        return {"impl": "reference", "unavailable": "config 5 (T^3 at 1 M cells: ~2e9 stored entries) does not fit the "
                "reference's CPU path; SURVEY.md 8d marks it infeasible"}
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        palantir_config.HOST_SPARSE_ON_ALL_RANKS = False
    n = args.cells or 1_000_000
    d = args.pcs or 100
    G, T_STEPS = bench.MAGIC_GENES, bench.MAGIC_STEPS
    x, info = synth_branching(n, d, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):
        res = utils.run_diffusion_maps(pd.DataFrame(x, index=synth_names(n)), n_components=10, knn=30,
                                       kernel_backend="sklearn")
    if world > 1 and rank != 0:
        # rank 0 alone brought T home; the other ranks rebuild the operator from its broadcast parts
        res["T"] = None
    parts = [None, None, None]
    if rank == 0:
        t = res["T"]
        parts = [torch.from_numpy(t.indptr.astype(np.int32)).cuda(), torch.from_numpy(t.indices.astype(np.int32)).cuda(),
                 torch.from_numpy(t.data).cuda()]
    if world > 1:
        meta = torch.tensor([parts[1].numel() if rank == 0 else 0], dtype=torch.int64, device="cuda")
        dist.broadcast(meta, 0)
        nnz = int(meta.item())
        if rank != 0:
            parts = [torch.empty(n + 1, dtype=torch.int32, device="cuda"), torch.empty(nnz, dtype=torch.int32, device="cuda"),
                     torch.empty(nnz, dtype=torch.float64, device="cuda")]
        for p in parts:
            dist.broadcast(p, 0)
        if rank != 0:
            from scipy.sparse import csr_matrix

            res["T"] = csr_matrix((parts[2].cpu().numpy(), parts[1].cpu().numpy(), parts[0].cpu().numpy()), shape=(n, n))
    nnz = int(parts[1].numel())
    del parts
    operator = utils._magic_operator(res["T"], res["EigenVectors"].values)
    bounds = _dist.shard_bounds(G, world)
    g_mine = bounds[rank + 1] - bounds[rank]
    ld = dv.round_up(g_mine, 4)
    xd = torch.rand((n, ld), dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return utils.magic_device(operator, xd, g_mine, T_STEPS, 1e-2)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _native.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item())
    clocks = sampler.stop() if rank == 0 else {}
    del xd
    torch.cuda.empty_cache()

    # ---- end to end: host ndarray in, host ndarray out (every rank holds the whole X, imputes its columns) ----
    e2e_steps = max(1, min(args.e2e_steps, 2))
    xh = np.random.default_rng(0).random((n, G), dtype=np.float32)
    utils.run_magic_imputation(xh[:, :256 * world], dm_res=res, n_steps=T_STEPS, sparse=False)      # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = utils.run_magic_imputation(xh, dm_res=res, n_steps=T_STEPS, sparse=False)
    barrier()
    t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sec = float(t.item())
    if rank != 0:
        dist.barrier()
        dist.destroy_process_group()
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    floor_bytes = T_STEPS * (2.0 * n * g_mine * 4 + nnz * 8)               # per rank
    gather_bytes = T_STEPS * float(nnz) * g_mine * 4
    line = {
        "metric": bench._metric("c5"), "value": n / (ms_step / 1e3), "unit": bench.UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"run_magic_imputation t={T_STEPS}: diffusion operator of {n} cells x {d} PCs (knn 30, "
                               f"{nnz} stored entries), X = uniform random {n} x {G} float32",
                   "l2": "X (8 GB) and the operator (0.4 GB) are larger than L2",
                   "parallelism": f"gene columns split over {world} GPU(s), operator replicated, no collective on the "
                                  "device path (e2e: all-gather of the imputed column shards)"},
        "e2e": {"value": n / e2e_sec, "unit": bench.UNIT, "h2d_bytes_per_step": int(n * (bounds[1] - bounds[0]) * 4),
                "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
                "input": "pageable numpy float32 matrix; dense ndarray result on the host (rank 0)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "csr_spmm_tile_kernel<float, 4> (256-gene tiles, warp per row)",
                     "achieved": floor_bytes / ms_step / 1e6, "peak": hbm_peak, "unit": "GB/s",
                     "frac": floor_bytes / ms_step / 1e6 / hbm_peak,
                     "algorithmic": "t * (2*N*G*4 + nnz*8) bytes per step per GPU (SURVEY.md 8d floor: X read and Y "
                                    "written once per product, FP32 operator once)",
                     "gather_traffic_gbs": gather_bytes / ms_step / 1e6,
                     "gather_note": "every (nonzero, gene) pair is a 4-byte load served by L2: nnz*G*4 B per product; "
                                    "this, not HBM, bounds the kernel",
                     "traffic": None},
        "cpu_baseline": None,
    }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line
