"""Phase time line of the fused data-parallel optimiser kernel (k_adam_dp), one process per GPU:
    DDRL_DP_TRACE=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_trace.py"""
import ctypes as C, os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ddrl_b200 import Learner, _native
D, A, hid, B = 24, 4, (256, 256), 1024
space = SimpleNamespace(high=np.ones(A, np.float32))
opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                      lr=1e-3, polyak=0.995, seed=0, batch_size=B)
L = Learner(opt, "learner", device=local)
assert L.connect_peers()
dev = torch.device("cuda", local)
batch = dict(obs1=torch.randn(B, D, device=dev), obs2=torch.randn(B, D, device=dev), acts=torch.rand(B, A, device=dev) * 2 - 1,
             rews=torch.randn(B, device=dev), done=torch.zeros(B, device=dev))
# steady state: 300 steps queued without any host synchronisation, then the stamps of the LAST step (two-kernel exchange:
# slots 0 reduce start, 1 flag published, 2 optimiser start, 3 all flags seen, 4 optimiser end, 6 / 7 this / previous prologue start)
if os.environ.get("DDRL_DP_V1", "0") == "1":
    for it in range(300):
        L.train(batch)
    out = (C.c_uint64 * 8)()
    _native.check(_native.lib().ddrl_sac_dp_trace(L._h, out))
    t = np.array(list(out), dtype=np.float64)
    rel = lambda i: t[i] - t[6]
    print(f"rank {rank}/{world} last step (ns after its prologue start): reduce kernel start {rel(0):.0f}, flag published {rel(1):.0f}, "
          f"optimiser kernel start {rel(2):.0f}, all flags seen {rel(3):.0f}, optimiser end {rel(4):.0f}; step period {t[6] - t[7]:.0f}; "
          f"absolute prologue start {t[6]:.0f}, publish {t[1]:.0f}, flags seen {t[3]:.0f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
rows = []
exs = []
for it in range(60):
    L.train(batch)
    if it >= 20:
        out = (C.c_uint64 * 8)()
        _native.check(_native.lib().ddrl_sac_dp_trace(L._h, out))
        t = np.array(list(out)[:6], dtype=np.float64)
        rows.append(t[1:] - t[:-1])
        ex = np.array(list(out)[6:8], dtype=np.float64)
        if ex[0] > 0:
            exs.append([ex[0] - t[0], ex[1] - t[0]])
        dist.barrier()
r = np.median(np.array(rows), axis=0)
print(f"rank {rank}/{world} k_adam_dp CTA 0 (median ns): split-K partials summed into the exchange slot {r[0]:.0f}, publish + wait for all "
      f"ranks' flags {r[1]:.0f}, all-peer read + adam {r[2]:.0f}, total {r[:3].sum():.0f}", flush=True)
if exs:
    e = np.median(np.array(exs), axis=0)
    print(f"rank {rank}: side-stream publish kernel started {e[0]:.0f} ns and ended {e[1]:.0f} ns relative to the start of the final kernel", flush=True)
dist.barrier()
dist.destroy_process_group()
