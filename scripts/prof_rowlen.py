"""Row-length distribution of the kernel matrix and of the shortest-path graph at benchmark size (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
k = out["kernel"]
ln = (k.indptr[1:] - k.indptr[:-1]).cpu().numpy()
print("kernel rows: mean", ln.mean(), "max", ln.max(), "quantiles 50/99/99.9/99.99 %:", np.quantile(ln, [.5, .99, .999, .9999]))
print("rows with more than 64 / 128 / 256 / 384 merged entries:", (ln > 64).sum(), (ln > 128).sum(), (ln > 256).sum(), (ln > 384).sum())
print("upper bound before merging (30 + in-degree): rows above 256 / 384 / 768:", ((ln + 7) > 256).sum(), ((ln + 7) > 384).sum(), ((ln + 7) > 768).sum())
# is the SpMV limited by the longest rows?  time it on the real matrix and on one whose rows are cut at 128 entries
import time
def timeit(fn, reps=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
xv = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(xv)
t_full = timeit(lambda: dv.call("pb200_csr_spmv_f64", k.indptr, k.indices, k.data, n, k.nnz, xv, y))
# truncated copy: same arrays, row i covers [indptr[i], min(indptr[i+1], indptr[i] + 128)) -- needs a CSR with gaps, so
# build a compact one
lens = torch.clamp(k.indptr[1:] - k.indptr[:-1], max=128)
ip2 = torch.zeros(n + 1, dtype=torch.int32, device="cuda"); ip2[1:] = torch.cumsum(lens, 0)
pos = torch.arange(int(ip2[-1].item()), device="cuda")
row_of = torch.repeat_interleave(torch.arange(n, device="cuda"), lens.long())
src = k.indptr[:-1].long()[row_of] + (pos - ip2[:-1].long()[row_of])
ind2, val2 = k.indices[src].contiguous(), k.data[src].contiguous()
nnz2 = int(ip2[-1].item())
t_cut = timeit(lambda: dv.call("pb200_csr_spmv_f64", ip2, ind2, val2, n, nnz2, xv, y))
print(f"SpMV full {t_full:.3f} ms ({k.nnz} nnz); rows cut at 128 entries {t_cut:.3f} ms ({nnz2} nnz) -> per-nnz {t_full/k.nnz*1e9:.3f} vs {t_cut/nnz2*1e9:.3f} fs")
