"""torchrun target: time the slab-decomposed coupled step (development aid).  args: nx [nsteps]"""
import os, sys, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import torch.distributed as dist
logging.disable(logging.CRITICAL)
from niwqg_b200 import CoupledModel, slab
from cases import lamb_params

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
kw, U0, k0 = lamb_params(nx, True, 10 ** 9, 1)
kw["tmax"] = 1e30; kw["twrite"] = 10 ** 9
m = slab.make_model(CoupledModel.Model, dist=dist, **kw)
lo, hi = slab.rows_of(rank, world, nx)
rng = np.random.RandomState(rank)
m.set_q(1e-5 * rng.randn(hi - lo, nx))
m.set_phi((np.ones((hi - lo, nx)) + 1j) * 0.14)
h = m._h
h.step(2); h.sync(); dist.barrier()
ms = h.time_steps(n) / n
t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
h.profile(True); h.step(n); prof = h.profile(False)
if rank == 0:
    print("slab nx=%d on %d GPUs: %.3f ms/step  %.3e pt-steps/s (%.1f%% of %d x 3392B roofline)"
          % (nx, world, ms, nx * nx / (ms * 1e-3), 100 * 3392 * nx * nx / (ms * 1e-3) / 6544e9 / world, world))
    for k, (tt, c) in prof.items():
        if c:
            print("   %-8s %8.3f ms/step  %4d launches/step  %.4f ms each" % (k, tt / n, c // n, tt / c))
dist.barrier(); dist.destroy_process_group()
