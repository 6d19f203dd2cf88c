"""Are the data-parallel replicas BIT-identical after K fused steps?  (one process per GPU, torchrun)
Each rank trains on its own batches; the flat main / target weights are compared across ranks as integers
(all-reduce MAX == all-reduce MIN).  DDRL_DP_NVLS=1 checks the multimem.ld_reduce exchange: the NVSwitch's
order of additions must be the same for every requesting rank."""
import os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ddrl_b200 import Learner
D, A, hid, B = 24, 4, (256, 256), 1024
space = SimpleNamespace(high=np.ones(A, np.float32))
opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                      lr=1e-3, polyak=0.995, seed=0, batch_size=B)
L = Learner(opt, "learner", device=local)
assert L.connect_peers()
dev = torch.device("cuda", local)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for it in range(K):
    batch = dict(obs1=torch.randn(B, D, device=dev, generator=gen), obs2=torch.randn(B, D, device=dev, generator=gen),
                 acts=torch.rand(B, A, device=dev, generator=gen) * 2 - 1, rews=torch.randn(B, device=dev, generator=gen),
                 done=torch.zeros(B, device=dev))
    L.train(batch)
torch.cuda.synchronize()
same = True
for which in ("main", "target"):
    w = L.get_flat_weights(which).view(torch.int32).to(torch.int64)
    hi, lo = w.clone(), w.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    same = same and bool(torch.equal(hi, lo))
    if rank == 0:
        print(f"{which}: {int((hi != lo).sum())} of {w.numel()} weights differ between the {world} replicas after {K} steps "
              f"(nvls={L.nvls}, comm_error={L.comm_error()}, finite={bool(torch.isfinite(L.get_flat_weights(which)).all())})", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if same else 3)
