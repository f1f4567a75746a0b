"""Where the time of dist.mll_grad_sharded goes beyond the local device call (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as tdist
rank, lrank, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
tdist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
from bobe_b200 import GP, dist
from oracle import gp_oracle as O
X, y = O.synthetic_training_set(2000, 16)
gp = GP(X, y, kernel="matern", lengthscales=np.ones(16), device=torch.device("cuda", lrank))
ref = O.OracleGP(X, y, kernel="matern", lengthscales=np.ones(16))
x0 = O.synthetic_restarts(ref, 8 * ws)
lo, hi = dist.shard_bounds(8 * ws, rank, ws)
def sync():
    tdist.barrier(); torch.cuda.synchronize()
for _ in range(4): dist.mll_grad_sharded(gp, x0)
sync(); t0 = time.perf_counter()
for _ in range(10): dist.mll_grad_sharded(gp, x0)
sync(); t_all = (time.perf_counter() - t0) / 10 * 1e3
sync(); t0 = time.perf_counter()
for _ in range(10): v, g = gp.neg_mll_and_grad_batched(x0[lo:hi])
torch.cuda.synchronize(); t_loc = (time.perf_counter() - t0) / 10 * 1e3
both = torch.as_tensor(np.concatenate([v[:, None], g], axis=1))
sync(); t0 = time.perf_counter()
for _ in range(10): dist.allgather_rows(both, 8 * ws)
sync(); t_g = (time.perf_counter() - t0) / 10 * 1e3
print(f"rank {rank}/{ws}: sharded call {t_all:.3f} ms | local neg_mll_and_grad_batched {t_loc:.3f} ms | allgather_rows alone {t_g:.3f} ms", flush=True)
tdist.destroy_process_group()
