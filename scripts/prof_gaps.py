"""GPU idle gaps inside one public-API step at benchmark size (dev tool; torch.profiler / CUPTI)."""
import sys, os, io, contextlib, warnings, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np, pandas as pd, torch
from torch.profiler import profile, ProfilerActivity
from palantir_b200 import utils, core
from palantir_b200.datasets import synth_branching, synth_names
n = 1_000_000
x, info = synth_branching(n, 100)
names = synth_names(n)
df = pd.DataFrame(x, index=names)
term = {names[v]: k for k, v in info["terminals"].items()}
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        res = utils.run_diffusion_maps(df, n_components=10, knn=30, kernel_backend="sklearn")
        ms = utils.determine_multiscale_space(res)
        pr = core.run_palantir(ms, names[info["early"]], terminal_states=term, num_waypoints=1200)
    return res, ms, pr
step(); step(); torch.cuda.synchronize()
t0 = time.time()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print("wall under profiler", time.time() - t0)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ev)
print("gpu activities", len(iv))
start, end = iv[0][0], max(e[1] for e in iv)
busy = 0.0; cur_s, cur_e = iv[0][0], iv[0][1]; gaps = []; last_name = iv[0][2]
for s, e, nm in iv[1:]:
    if s > cur_e:
        gaps.append((s - cur_e, last_name, nm, cur_e - start))
        busy += cur_e - cur_s; cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
    if e >= cur_e: last_name = nm
busy += cur_e - cur_s
print(f"span {1e-3*(end-start):.1f} ms, busy {1e-3*busy:.1f} ms, idle {1e-3*(end-start-busy):.1f} ms in {len(gaps)} gaps")
for g, a, b, at in sorted(gaps, reverse=True)[:28]:
    print(f"  {1e-3*g:7.2f} ms idle at t={1e-3*at:7.1f} ms  after {a[:50]:50s} before {b[:50]}")
agg = {}
for s, e, nm in iv:
    key = nm if nm.startswith("Mem") else "kernels"
    a = agg.setdefault(key, [0.0, 0]); a[0] += e - s; a[1] += 1
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {1e-3*t:8.1f} ms  x{c:5d}  {k}")
big = sorted(((e - s, nm, s - start) for s, e, nm in iv if nm.startswith("Mem")), reverse=True)[:14]
for d, nm, at in big:
    print(f"     {1e-3*d:7.2f} ms at t={1e-3*at:7.1f}  {nm}")
