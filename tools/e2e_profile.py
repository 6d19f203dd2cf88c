"""Host-side time of each call of the end-to-end step (GPU box only): where the e2e microseconds go."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from bench import CONFIGS, make_opt
from ddrl_b200 import Learner, ReplayBuffer

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
D, A, B = cfg["D"], cfg["A"], cfg["B"]
dev = torch.device("cuda", 0)
rb = ReplayBuffer(D, A, 1_000_000, device=0, seed=3)
n = 250_000
for _ in range(4):
    rb.store_batch(torch.randn((n, D), device=dev), torch.rand((n, A), device=dev), torch.randn(n, device=dev),
                   torch.randn((n, D), device=dev), torch.zeros(n, device=dev))
L = Learner(make_opt(cfg, seed=7), "learner", device=0)
g = np.random.Generator(np.random.PCG64(5))
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
new = [pin(g.standard_normal((B, D), dtype=np.float32)), pin(g.uniform(-1, 1, (B, A)).astype(np.float32)),
       pin(g.standard_normal(B, dtype=np.float32)), pin(g.standard_normal((B, D), dtype=np.float32)),
       pin((g.random(B) < 0.01).astype(np.float32))]
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
N_IT = 300
for it in range(N_IT + 20):
    if it == 20:
        T.clear(); torch.cuda.synchronize(); t_all = time.perf_counter()
    t = time.perf_counter(); batch = rb.sample_batch(B); tick("sample_batch->numpy", t)
    t = time.perf_counter(); out = L.train(batch); tick("train(numpy) enqueue", t)
    t = time.perf_counter(); rb.store_batch(*new); tick("store_batch(host)", t)
    t = time.perf_counter(); sc = out["scalars"].cpu(); tick("losses.cpu()", t)
torch.cuda.synchronize()
tot = time.perf_counter() - t_all
for k, v in T.items():
    print(f"{k:28s} {v / N_IT * 1e6:8.1f} us")
print(f"{'total':28s} {tot / N_IT * 1e6:8.1f} us")
# device-resident loop, host time per call
T.clear()
for it in range(N_IT + 20):
    if it == 20:
        T.clear(); torch.cuda.synchronize(); t_all = time.perf_counter()
    t = time.perf_counter(); batch = rb.sample_batch(B, device=True); tick("sample_batch(device) enqueue", t)
    t = time.perf_counter(); out = L.train(batch); tick("train(device) enqueue", t)
t = time.perf_counter(); torch.cuda.synchronize(); tick("final sync", t)
tot = time.perf_counter() - t_all
for k, v in T.items():
    print(f"{k:28s} {v / N_IT * 1e6:8.1f} us")
print(f"{'total device loop':28s} {tot / N_IT * 1e6:8.1f} us")
