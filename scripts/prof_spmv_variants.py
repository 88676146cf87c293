"""SpMV at benchmark size (dev tool): the row kernels against the kernel over the nonzero stream, whole matrix and
row ranges (what a rank computes when the eigensolver is sharded by rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv, _native
from palantir_b200.datasets import synth_branching
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x, info = synth_branching(n, 100)
xd = dv.to_device(x)
idx, dist, st = dv.knn_like_sklearn(xd, 31)
order = st.pop("perm", None)
k = dv.adaptive_kernel(idx, dist, 30, 0.0)
k.order = order
tvals, svals, dis, dsq = dv.transition_and_symmetric(k)
pl = order.long(); inv = torch.empty_like(order); inv[pl] = torch.arange(n, dtype=order.dtype, device="cuda")
ip = torch.zeros(n + 1, dtype=torch.int32, device="cuda"); ip[1:] = torch.cumsum((k.indptr[1:] - k.indptr[:-1])[pl], 0)
ind, val = dv.empty((k.nnz,), dv.I32), dv.empty((k.nnz,), dv.F64)
dv.call("pb200_csr_permute", k.indptr, k.indices, svals, n, order, inv, ip, ind, val)
thr = int(_native.load().pb200_spmv_long_row())
lr = torch.nonzero((ip[1:] - ip[:-1]) > thr).to(torch.int32).reshape(-1).contiguous()
nl = int(lr.numel())
xv = torch.rand(n, dtype=torch.float64, device="cuda")
def timeit(fn, reps=100):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
alg = 12 * k.nnz + 4 * (n + 1) + 16 * n
ys = {}
y = torch.zeros_like(xv)
t = timeit(lambda: dv.call("pb200_csr_spmv_rows_f64", ip, ind, val, n, k.nnz, 0, n, lr, nl, xv, y))
ys[0] = y
print(f"row kernels (16 lanes per row + a warp per hub row): {t:.4f} ms = {alg / t / 1e6:.0f} GB/s")
lens = (ip[1:] - ip[:-1]).long()
ll = lens[lr.long()]
print(f"long rows {nl}: nnz {int(ll.sum())} of {k.nnz}; mean {float(ll.float().mean()):.1f}, max {int(ll.max())}; "
      f"quantiles {[int(v) for v in torch.quantile(ll.float(), torch.tensor([0.5, 0.9, 0.99, 0.999], device=ll.device))]}")
# nonzero-stream kernel
tile = int(_native.load().pb200_csr_spmv_stream_tile())
n_tiles = -(-k.nnz // tile)
tile_lo = dv.empty((n_tiles + 1,), dv.I32)
head, tail = dv.zeros((n_tiles,), dv.F64), dv.zeros((n_tiles,), dv.F64)
dv.call("pb200_csr_spmv_stream_prepare", ip, n, k.nnz, tile_lo)
ystr = torch.full_like(xv, float("nan"))
t = timeit(lambda: dv.call("pb200_csr_spmv_stream_f64", ip, ind, val, n, k.nnz, tile_lo, head, tail, xv, ystr, 0, n, 0, n_tiles))
rel = float(((ystr - ys[0]).abs() / ys[0].abs().clamp_min(1e-300)).max())
print(f"stream kernel: {t:.4f} ms = {alg / t / 1e6:.0f} GB/s; max rel |y - y0| {rel:.2e}; nan {int(torch.isnan(ystr).sum())}")
hrow = (n // 2 + 15) // 16 * 16
e_mid = int(ip[hrow].item())
y2 = torch.full_like(xv, float("nan"))
dv.call("pb200_csr_spmv_stream_f64", ip, ind, val, n, k.nnz, tile_lo, head, tail, xv, y2, 0, hrow, 0, -(-e_mid // tile))
dv.call("pb200_csr_spmv_stream_f64", ip, ind, val, n, k.nnz, tile_lo, head, tail, xv, y2, hrow, n, e_mid // tile, n_tiles)
print("stream kernel, two row ranges: identical to one call:", bool(torch.equal(y2, ystr)), " nan", int(torch.isnan(y2).sum()))
