"""N-step sequence sample launches for ncu (GPU box only)."""
import sys, os
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import NStepReplayBuffer
opt = SimpleNamespace(Ln=8, obs_shape=(24,), act_shape=(4,), buffer_size=1_000_000, batch_size=1024, num_buffers=1)
rb = NStepReplayBuffer(opt, seed=3)
g = np.random.Generator(np.random.PCG64(0))
for lo in range(0, 1_000_000, 100_000):
    rb.store_batch(g.standard_normal((100_000, 9, 24), dtype=np.float32), g.standard_normal((100_000, 8, 4), dtype=np.float32),
                   g.standard_normal((100_000, 8), dtype=np.float32), np.zeros((100_000, 8), np.float32))
for _ in range(4):
    rb.sample_batch(262144, device=True)
torch.cuda.synchronize()
print("done")
