"""Gradient error of the full-size C3 step vs the float64 oracle, both GEMM modes, and of a float32 oracle (GPU box only)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import Learner
from test_sac_gpu import build_pair, rel
from oracle.sac1_oracle import SAC1Oracle, conditioned_params, make_batch
D, A, B, scale, hidden = 376, 17, 4096, 0.4, (256, 256)
params = conditioned_params(D, A, hidden, seed=300 + D)
batch, noise = make_batch(D, A, B, seed=400 + D)
o32 = SAC1Oracle(D, A, hidden=hidden, gamma=0.99, polyak=0.995, lr=1e-3, alpha=0.2, act_scale=scale, params=params, dtype=torch.float32)
for gemm in ("tc", "ffma"):
    learner, oracle = build_pair(Learner, D, A, hidden, B, params, gemm=gemm, act_scale=scale)
    want_g = oracle.flat_grads(batch, noise)
    want = oracle.step(batch, noise)
    got = learner.train(batch, noise=noise, split=True, sync_outputs=True)
    g = learner.get_flat_weights("grad").cpu().numpy()
    sc = got["scalars"].cpu().numpy()
    print(gemm, "grad rel", rel(g, want_g), "losses rel", [abs(sc[i] - float(want[k])) / abs(float(want[k])) for i, k in enumerate(("pi_loss", "q1_loss", "q2_loss"))],
          "q1 rel", rel(got["q1"].cpu().numpy(), want["q1"]), flush=True)
g32 = o32.flat_grads(batch, noise)
print("float32 oracle grad rel", rel(g32, want_g))
