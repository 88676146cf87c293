"""SpMV (as the eigensolver uses it) timing at benchmark size (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import pipeline, _device as dv, _native
from palantir_b200.datasets import synth_branching
n = 1_000_000
x, info = synth_branching(n, 100)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()))
k = out["kernel"]
pl = k.order.long(); inv = torch.empty_like(k.order); inv[pl] = torch.arange(n, dtype=k.order.dtype, device="cuda")
ip = torch.zeros(n + 1, dtype=torch.int32, device="cuda"); ip[1:] = torch.cumsum((k.indptr[1:] - k.indptr[:-1])[pl], 0)
ind, val = dv.empty((k.nnz,), dv.I32), dv.empty((k.nnz,), dv.F64)
dv.call("pb200_csr_permute", k.indptr, k.indices, k.data, n, k.order, inv, ip, ind, val)
thr = int(_native.load().pb200_spmv_long_row())
lr = torch.nonzero((ip[1:] - ip[:-1]) > thr).to(torch.int32).reshape(-1).contiguous()
xv = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(xv); y2 = torch.empty_like(xv)
def timeit(fn, reps=50):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t0 = timeit(lambda: dv.call("pb200_csr_spmv_f64", ip, ind, val, n, k.nnz, xv, y))
t1 = timeit(lambda: dv.call("pb200_csr_spmv_split_f64", ip, ind, val, n, k.nnz, lr, int(lr.numel()), xv, y2))
alg = 12 * k.nnz + 4 * (n + 1) + 16 * n
print(f"long-row threshold {thr}: {lr.numel()} long rows; plain {t0:.3f} ms, split {t1:.3f} ms = {alg/t1/1e6:.0f} GB/s; max diff {float((y-y2).abs().max()):.2e}")
