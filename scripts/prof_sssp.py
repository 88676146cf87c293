"""SSSP timing on the benchmark graph with tuning knobs (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv, _device_core as dc, _native
n = 1_000_000
rng = np.random.default_rng(0)
if "--bench" in sys.argv:
    # the benchmark's own graph: multiscale space of the synthetic 1M x 100 workload
    import pandas as pd
    from palantir_b200 import pipeline, utils
    from palantir_b200.datasets import synth_branching
    x, info = synth_branching(n, 100)
    out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
    data_t = dc.minmax_scale(out["multiscale"])
    perm, inv, work = dc.morton_order(data_t)
    idx, dist, st = dc.knn_table_boxed(work, 30)
    g = dc.undirected_graph(idx, dist)
    src = dv.to_device(inv.cpu().numpy().astype(np.int64)[out["waypoints"]])
    lib = _native.load()
    mean_w = float(g.data.mean().item())
    print("bench graph arcs", g.nnz, "mean w", mean_w)
else:
  pass
if "--bench" not in sys.argv:
    # a thick noisy curve in 6-D, like the min-max scaled multiscale space
    t = np.sort(rng.random(n))
    data = np.stack([t, np.sin(3 * t) * 0.5 + 0.5, np.cos(5 * t) * 0.5 + 0.5, t ** 2, np.sqrt(t), 0.5 + 0.3 * np.sin(9 * t)], 1)
    data += rng.normal(size=data.shape) * 2e-3
    dev = dv.to_device(np.ascontiguousarray(data))
    perm, inv, work = dc.morton_order(dev)
    idx, dist, st = dc.knn_table_boxed(work, 30)
    g = dc.undirected_graph(idx, dist)
    src = dv.to_device(rng.choice(n, 1198, replace=False).astype(np.int64))
    lib = _native.load()
    mean_w = float(g.data.mean().item())
    print("arcs", g.nnz, "mean w", mean_w)
def run(tag, mult=4.0, nsrc=None):
    sr = src if nsrc is None else src[:nsrc].contiguous()
    torch.cuda.synchronize(); t0 = time.time()
    D, stats = dc.sssp(g, sr, delta=mult * mean_w, want_stats=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    s = stats.double().mean(0).cpu().numpy()
    print(f"{tag} sources {sr.numel()} delta {mult}x: {dt*1e3:.1f} ms rounds {s[0]:.0f} relax/arcs {s[1]/g.nnz:.2f} advances {s[2]:.0f}", flush=True)
    return D
dc.SSSP_MODE = "csr"; lib.pb200_sssp_set_variant(4)
Dref = run("csr v4 (128 thr)")
if "--one" in sys.argv:
    sys.exit(0)
if "--small-delta" in sys.argv:
    dc.SSSP_MODE = "csr"; lib.pb200_sssp_set_variant(12)
    for mult in (1.5, 2.0, 2.5, 3.0, 4.0, 6.0):
        run("csr v12", mult)
    sys.exit(0)
if "--arcrec128" in sys.argv:
    dc.SSSP_MODE = "arcrec"; dc.SSSP_THREADS = 128
    for mult in (2.5, 4.0):
        D = run("arcrec 128", mult)
        print("   identical:", bool(torch.equal(D, Dref)))
    dc.SSSP_MODE = "csr"; dc.SSSP_THREADS = None; lib.pb200_sssp_set_variant(12)
    run("csr v12", 2.5)
    sys.exit(0)
if "--sources" in sys.argv:
    for ns in (599, 300, 150, 1):
        dc.SSSP_MODE = "arcrec"; dc.SSSP_THREADS = None
        D = run("arcrec auto", 2.5, nsrc=ns)
        print("   identical to the reference rows:", bool(torch.equal(D, Dref[:ns])))
        dc.SSSP_MODE = "csr"; lib.pb200_sssp_set_variant(12)
        D = run("csr v12", 2.5, nsrc=ns)
        print("   identical to the reference rows:", bool(torch.equal(D, Dref[:ns])))
    sys.exit(0)
if "--delta" in sys.argv:
    dc.SSSP_MODE = "arcrec"; dc.SSSP_THREADS = None
    for ns in (1198, 599, 300, 150):
        for mult in (4.0, 8.0, 16.0, 32.0, 64.0):
            run("auto", mult, nsrc=ns)
    sys.exit(0)
if "--variants" in sys.argv:
    for v in (12, 8, 12, 8):
        lib.pb200_sssp_set_variant(v)
        D = run(f"csr variant {v}")
        print("   identical:", bool(torch.equal(D, Dref)), flush=True)
        del D
    sys.exit(0)
run("csr v4 (128 thr)", nsrc=150)
dc.SSSP_MODE = "arcrec"
for thr in (128, 256, 512):
    dc.SSSP_THREADS = thr
    for mult in (4.0, 8.0):
        D = run(f"arcrec {thr} thr", mult)
    print("   identical to csr result:", bool(torch.equal(D, Dref)))
dc.SSSP_THREADS = None
for ns in (599, 300, 150):
    run("arcrec auto", nsrc=ns)
