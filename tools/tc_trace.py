"""Phase time line of one gemm_grouped_tc launch (GPU box only): python tools/tc_trace.py [C1|C2|C3] [stage]"""
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
D, A, hid, B = {"C1": (8, 2, (256, 256), 256), "C2": (24, 4, (256, 256), 1024), "C3": (376, 17, (256, 256), 4096)}[cfg]
space = SimpleNamespace(high=np.ones(A, np.float32))
opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                      lr=1e-3, polyak=0.995, seed=0, batch_size=B)
L = Learner(opt, "learner")
dev = torch.device("cuda")
batch = dict(obs1=torch.randn(B, D, device=dev), obs2=torch.randn(B, D, device=dev), acts=torch.rand(B, A, device=dev) * 2 - 1,
             rews=torch.randn(B, device=dev), done=torch.zeros(B, device=dev))
for _ in range(3):
    L.train(batch)
torch.cuda.synchronize()
lib = _native.lib(); s = torch.cuda.current_stream()
tr = torch.zeros((1024, 32), dtype=torch.int64, device=dev)
tiles = C.c_int()
names = ["start", "setup", "1st stage", "MMAs issued", "acc done", "epi warp0", "all warps"]
for rep in range(3):
    _native.check(lib.ddrl_sac_debug_stage(L._h, B, stage, 3, C.c_void_p(s.cuda_stream)))      # warm caches like the real step
    _native.check(lib.ddrl_sac_trace_stage(L._h, B, stage, C.c_void_p(tr.data_ptr()), 1024, C.byref(tiles), C.c_void_p(s.cuda_stream)))
    torch.cuda.synchronize()
t = tr[: tiles.value, :7].cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
print(f"{cfg} stage {stage}: {tiles.value} tiles; ns relative to the first CTA start (median over CTAs / max)")
for i, nm in enumerate(names):
    col = t[:, i] - t0
    print(f"  {nm:12s} median {np.median(col):8.0f}   max {col.max():8.0f}")
full = tr[: tiles.value].cpu().numpy().astype(np.float64)
if full[:, 8:].any():
    print("  fused kernel, ns after the layer-1 accumulator (median over CTAs): k-block operands ready (MMA issuer) / converted (q0 warp)")
    ref = full[:, 3]
    for kb in range(8):
        if full[:, 8 + kb].any():
            print(f"    kb {kb}: ready {np.median(full[:, 8 + kb] - ref):7.0f}   converted {np.median(full[:, 16 + kb] - ref):7.0f}")
if full[:, 24:].any():
    for cg in (0, 1):
        c = 24 + 4 * cg
        print("  converter group %d, second k-block (median ns): ld+wait %.0f, convert+st issue %.0f, wait::st %.0f" % (
            cg, np.median(full[:, c + 1] - full[:, c]), np.median(full[:, c + 2] - full[:, c + 1]), np.median(full[:, c + 3] - full[:, c + 2])))
print("  per-CTA durations (median ns): setup %.0f, load latency %.0f, mainloop issue %.0f, MMA drain %.0f, epilogue %.0f, join %.0f" % tuple(
    np.median(t[:, i + 1] - t[:, i]) for i in range(6)))
# per tile: first stage landed -> accumulator complete (the MMA main loop), in launch order: the problems of a stage are laid
# out one after the other, so populations with different operand layouts (dgrad: K-major B; wgrad: MN-major A and B) show
if os.environ.get("TC_TRACE_TILES"):
    ml = t[:, 4] - t[:, 2]
    print("  main loop per tile (ns), 16 tiles per line:")
    for i in range(0, len(ml), 16):
        print("   ", " ".join(f"{x:5.0f}" for x in ml[i:i + 16]))
