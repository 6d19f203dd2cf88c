"""A few launches of ONE phase of the SAC1 step for ncu (GPU box only): python tools/prof_stage.py [C1|C2|C3] [stage] [reps]
(stage numbering of ddrl_sac_debug_stage)."""
import sys, os, ctypes as C
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import Learner, _native
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
D, A, hid, B = {"C1": (8, 2, (256, 256), 256), "C2": (24, 4, (256, 256), 1024), "C3": (376, 17, (256, 256), 4096)}[cfg]
space = SimpleNamespace(high=np.ones(A, np.float32))
opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                      lr=1e-3, polyak=0.995, seed=0, batch_size=B)
os.environ["DDRL_NO_GRAPH"] = "1"
L = Learner(opt, "learner")
dev = torch.device("cuda")
batch = dict(obs1=torch.randn(B, D, device=dev), obs2=torch.randn(B, D, device=dev), acts=torch.rand(B, A, device=dev) * 2 - 1,
             rews=torch.randn(B, device=dev), done=torch.zeros(B, device=dev))
L.train(batch)
torch.cuda.synchronize()
s = torch.cuda.current_stream()
_native.check(_native.lib().ddrl_sac_debug_stage(L._h, B, stage, reps, C.c_void_p(s.cuda_stream)))
torch.cuda.synchronize()
print("done")
