"""One device-resident step of the hot path for `ncu` (launch list / full capture): a warm-up step outside the
profiled range, then one step between cudaProfilerStart / Stop.  Use with `ncu --profile-from-start off` and a
`-k regex:` filter for the kernels wanted (the ~3000 short vector kernels of the eigensolver make an unfiltered
capture of a step take tens of minutes).

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      -k regex:'knn_|cells_|sssp_|csr_spmv|merge_rows|symmetrize|gj_|refine|colsum|project|maxmin' \
      --log-file gpurun_out/launches.csv python scripts/ncu_step.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from palantir_b200 import _device as dv, pipeline
from palantir_b200.datasets import synth_branching

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
x, info = synth_branching(n, d)
xd = dv.to_device(x)
term = list(info["terminals"].values())
pipeline.run_device(xd, info["early"], term)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pipeline.run_device(xd, info["early"], term)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"step: {e0.elapsed_time(e1):.1f} ms (under a profiler this is not a benchmark number)")
