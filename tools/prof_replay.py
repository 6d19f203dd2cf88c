"""sample_many / store_batch launches for ncu (GPU box only): python tools/prof_replay.py [C1|C2|C3] [n_batches]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import ReplayBuffer
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
D, A, cap, B = {"C1": (8, 2, 1_000_000, 256), "C2": (24, 4, 1_000_000, 1024), "C3": (376, 17, 2_000_000, 4096)}[cfg]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else (2048 if cfg != "C3" else 128)
dev = torch.device("cuda")
rb = ReplayBuffer(D, A, cap, seed=1)
for lo in range(0, cap, 250_000):
    n = min(250_000, cap - lo)
    rb.store_batch(torch.randn(n, D, device=dev), torch.rand(n, A, device=dev), torch.randn(n, device=dev),
                   torch.randn(n, D, device=dev), torch.zeros(n, device=dev))
for _ in range(4):
    out = rb.sample_many(nb, B)
torch.cuda.synchronize()
print("done")
