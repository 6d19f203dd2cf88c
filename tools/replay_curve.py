"""Gather / store throughput against rows per launch (GPU box only): python tools/replay_curve.py C1|C2|C3 [tag]
Device time of the kernels alone: `reps` native calls back to back between one pair of CUDA events, outputs
preallocated.  Environment knobs of the ring are read at creation: DDRL_ROW_ALIGN, DDRL_GATHER_MODE (0 auto, 1 bulk-async,
2 register kernels), DDRL_BULK_MIN_BYTES."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import ReplayBuffer, _native

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
tag = sys.argv[2] if len(sys.argv) > 2 else "default"
D, A, cap, B = {"C1": (8, 2, 1_000_000, 256), "C2": (24, 4, 1_000_000, 1024), "C3": (376, 17, 2_000_000, 4096)}[cfg]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda")
rb = ReplayBuffer(D, A, cap, seed=1)
for lo in range(0, cap, 250_000):
    n = min(250_000, cap - lo)
    rb.store_batch(torch.randn(n, D, device=dev), torch.rand(n, A, device=dev), torch.randn(n, device=dev),
                   torch.randn(n, D, device=dev), torch.zeros(n, device=dev))
row = 4 * (2 * D + A + 2)
lib = _native.lib()
s = torch.cuda.current_stream()
nbs = [nb for nb in (1, 8, 64, 512, 2048) if nb * B * row <= 3.2e9]
nmax = max(nbs) * B
f32 = dict(dtype=torch.float32, device=dev)
o = [torch.empty((nmax, D), **f32), torch.empty((nmax, D), **f32), torch.empty((nmax, A), **f32), torch.empty(nmax, **f32), torch.empty(nmax, **f32)]
res = []
for nb in nbs:
    reps = max(10, min(400, 4096 // nb))
    for i in range(3 + reps):
        if i == 3:
            e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
        _native.check(lib.ddrl_rb_sample(rb.native_handle, B, nb, None, 77, i, 0, *[C.c_void_p(t.data_ptr()) for t in o], None,
                                         C.c_void_p(s.cuda_stream)))
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / reps
    gbs = 2 * row * nb * B / t / 1e9
    res.append(dict(cfg=cfg, tag=tag, op="sample", batches_per_launch=nb, rows=nb * B, us=t * 1e6, GBs=gbs, frac=gbs / peak))
    print(res[-1], flush=True)
src = [torch.randn((nmax, D), **f32), torch.rand((nmax, A), **f32), torch.randn(nmax, **f32), torch.randn((nmax, D), **f32), torch.zeros(nmax, **f32)]
for nb in nbs:
    m = min(nb * B, cap)
    reps = max(10, min(400, 4096 // nb))
    args = [t[:m] for t in src]
    for i in range(3 + reps):
        if i == 3:
            e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
        _native.check(lib.ddrl_rb_store_batch(rb.native_handle, *[C.c_void_p(t.data_ptr()) for t in args], m, 0, C.c_void_p(s.cuda_stream)))
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / reps
    gbs = 2 * row * m / t / 1e9
    res.append(dict(cfg=cfg, tag=tag, op="store", rows=m, us=t * 1e6, GBs=gbs, frac=gbs / peak))
    print(res[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"replay_curve_{cfg}_{tag}.json"), "w") as f:
    json.dump(res, f, indent=1)
