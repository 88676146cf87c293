"""Multi-GPU parity: the sharded run (torchrun, NCCL) must reproduce the single-GPU results.

  python scripts/mgpu_check.py --save gpurun_out/mgpu_ref.npz
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/mgpu_check.py --compare gpurun_out/mgpu_ref.npz
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from palantir_b200 import _device as dv, pipeline
from palantir_b200.datasets import synth_branching

ap = argparse.ArgumentParser()
ap.add_argument("--save"); ap.add_argument("--compare"); ap.add_argument("--cells", type=int, default=200_000)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, info = synth_branching(a.cells, 50)
out = pipeline.run_device(dv.to_device(x), info["early"], list(info["terminals"].values()), num_waypoints=600)
res = dict(pt=out["pseudotime"].cpu().numpy(), fates=out["fate_probs"].cpu().numpy(), ent=out["entropy"].cpu().numpy(),
           lam=np.asarray(out["eigenvalues"]), wp=out["waypoints"], kdata=out["kernel"].data.cpu().numpy(),
           kidx=out["kernel"].indices.cpu().numpy())
if a.save and rank == 0:
    np.savez(a.save, **res); print("saved", a.save)
if a.compare:
    ref = np.load(a.compare)
    ok = True
    for k in res:
        same = np.array_equal(res[k], ref[k])
        err = float(np.max(np.abs(res[k].astype(np.float64) - ref[k].astype(np.float64)))) if res[k].shape == ref[k].shape else float("nan")
        print(f"rank {rank} {k}: identical={same} max|diff|={err:.3e}", flush=True)
        ok &= (same or err < 1e-9)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if ok else 1)
if world > 1:
    dist.destroy_process_group()
