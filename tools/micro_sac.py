"""Scratch micro-benchmark of the SAC1 step (GPU box only)."""
import sys, os, json
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import Learner

def flops(D, A, h1, h2, B):
    Lpi = 2 * B * (D * h1 + h1 * h2 + 2 * h2 * A)
    Lq = 2 * B * ((D + A) * h1 + h1 * h2 + h2)
    return 5 * Lpi + 10 * Lq

res = []
for name, D, A, hid, B in [("C1", 8, 2, (256, 256), 256), ("C2", 24, 4, (256, 256), 1024), ("C3", 376, 17, (256, 256), 4096),
                           ("refdefault", 24, 4, (400, 300), 256)]:
    space = SimpleNamespace(high=np.ones(A, np.float32))
    opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                          lr=1e-3, polyak=0.995, seed=0, batch_size=B)
    L = Learner(opt, "learner")
    dev = torch.device("cuda")
    batch = dict(obs1=torch.randn(B, D, device=dev), obs2=torch.randn(B, D, device=dev), acts=torch.rand(B, A, device=dev) * 2 - 1,
                 rews=torch.randn(B, device=dev), done=torch.zeros(B, device=dev))
    for _ in range(5): L.train(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n): L.train(batch)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n * 1e-3
    f = flops(D, A, hid[0], hid[1], B)
    res.append(dict(cfg=name, us=t * 1e6, updates_s=1 / t, Mtrans_s=B / t / 1e6, TFLOPs=f / t / 1e12))
    print(res[-1], flush=True)
    del L
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_sac.json"), "w"), indent=1)
