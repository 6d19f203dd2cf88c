"""Real (CUDA-event) time of each GEMM stage of the SAC1 step, launched back to back (GPU box only)."""
import sys, os, ctypes as C
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import Learner, _native
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
D, A, hid, B = {"C1": (8, 2, (256, 256), 256), "C2": (24, 4, (256, 256), 1024), "C3": (376, 17, (256, 256), 4096)}[cfg]
space = SimpleNamespace(high=np.ones(A, np.float32))
opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                      lr=1e-3, polyak=0.995, seed=0, batch_size=B)
L = Learner(opt, "learner")
dev = torch.device("cuda")
batch = dict(obs1=torch.randn(B, D, device=dev), obs2=torch.randn(B, D, device=dev), acts=torch.rand(B, A, device=dev) * 2 - 1,
             rews=torch.randn(B, device=dev), done=torch.zeros(B, device=dev))
L.train(batch); torch.cuda.synchronize()
lib = _native.lib(); s = torch.cuda.current_stream()
names = ["L1", "L2", "QL1", "QL2", "BQ", "BP", "BP3", "prologue", "heads", "qheads", "pbwd", "adam", "sideBQ", "sideBP", "sideBP3"]
for st, nm in enumerate(names):
    reps = 50
    _native.check(lib.ddrl_sac_debug_stage(L._h, B, st, 5, C.c_void_p(s.cuda_stream)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.check(lib.ddrl_sac_debug_stage(L._h, B, st, reps, C.c_void_p(s.cuda_stream)))
    e1.record(); torch.cuda.synchronize()
    t_stream = e0.elapsed_time(e1) / reps * 1e3
    _native.check(lib.ddrl_sac_debug_stage(L._h, B, st, -reps, C.c_void_p(s.cuda_stream)))      # builds the graph + warm-up
    torch.cuda.synchronize()
    e0.record()
    _native.check(lib.ddrl_sac_debug_stage(L._h, B, st, -reps, C.c_void_p(s.cuda_stream)))
    e1.record(); torch.cuda.synchronize()
    print(f"{cfg} stage {nm:8s}: {t_stream:8.1f} us as stream launches, {e0.elapsed_time(e1)/reps*1e3:8.1f} us as graph nodes", flush=True)
