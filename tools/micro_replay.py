"""Scratch micro-benchmark of the replay kernels (GPU box only)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import ReplayBuffer

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

res = []
for name, D, A, cap, B in [("C1", 8, 2, 1_000_000, 256), ("C2", 24, 4, 1_000_000, 1024), ("C3", 376, 17, 2_000_000, 4096)]:
    dev = torch.device("cuda")
    rb = ReplayBuffer(D, A, cap, seed=1)
    chunk = 250_000
    for lo in range(0, cap, chunk):
        n = min(chunk, cap - lo)
        rb.store_batch(torch.randn(n, D, device=dev), torch.rand(n, A, device=dev), torch.randn(n, device=dev),
                       torch.randn(n, D, device=dev), torch.zeros(n, device=dev))
    row = 4 * (2 * D + A + 2)
    for nb in (1, 16, 256, 2048):
        if nb * B * row > 6e9: continue
        t = timeit(lambda: rb.sample_many(nb, B))
        res.append(dict(cfg=name, op="sample", n_batches=nb, rows=nb * B, us=t * 1e6, Mtrans_s=nb * B / t / 1e6, GBs=2 * row * nb * B / t / 1e9))
        print(res[-1], flush=True)
    n = 1 << 20
    o, a, r, o2, d = torch.randn(n, D, device=dev), torch.rand(n, A, device=dev), torch.randn(n, device=dev), torch.randn(n, D, device=dev), torch.zeros(n, device=dev)
    for m in (256, 65536, n):
        if m > cap: continue
        t = timeit(lambda: rb.store_batch(o[:m], a[:m], r[:m], o2[:m], d[:m]))
        res.append(dict(cfg=name, op="store", rows=m, us=t * 1e6, Mtrans_s=m / t / 1e6, GBs=2 * row * m / t / 1e9))
        print(res[-1], flush=True)
    del rb
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_replay.json"), "w"), indent=1)
