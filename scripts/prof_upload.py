"""Dev probe: pageable host -> device upload, plain .to(device) against the staged pipeline of _device.to_device."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from palantir_b200 import _device as dv
x = np.random.default_rng(0).random((1_000_000, 100), dtype=np.float32)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = torch.from_numpy(x).to("cuda"); torch.cuda.synchronize()
    t1 = time.perf_counter()
    b = dv.to_device(x); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"plain {1e3*(t1-t0):.1f} ms, staged {1e3*(t2-t1):.1f} ms, equal {bool(torch.equal(a, b))}")
xf = np.asfortranarray(x[:, :48].astype(np.float64))
t0 = time.perf_counter(); c = dv.to_device(xf); torch.cuda.synchronize(); print(f"fortran 384 MB staged {1e3*(time.perf_counter()-t0):.1f} ms, equal {bool(torch.equal(c.cpu(), torch.from_numpy(np.ascontiguousarray(xf))))}")
# staging threads / chunk size of the pipeline
for bufs, chunk_mb in ((4, 16), (8, 16), (8, 8), (12, 8), (16, 4), (6, 32), (16, 8)):
    dv._UPLOAD_BUFS, dv._UPLOAD_CHUNK = bufs, chunk_mb << 20
    dv._upload_state.clear()
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        b = dv.to_device(x); torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print(f"{bufs} staging buffers/threads x {chunk_mb} MB: {min(ts[1:]):.1f} ms (runs {[round(t, 1) for t in ts]})")
print("host cores", os.cpu_count())
