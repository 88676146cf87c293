"""Dev probe: bare tcgen05.mma.kind::tf32 throughput with 1, 2 and 4 accumulators in rotation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from palantir_b200 import _device as dv
for rot in (0, 1, 3):
    it = 200_000
    for _ in range(2):
        dv.call("pb200_tf32_peak_rot", 148, it, rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dv.call("pb200_tf32_peak_rot", 148, it, rot); e1.record(); torch.cuda.synchronize()
    print(f"accumulators {rot + 1}: {148 * it * 262144 / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s")
