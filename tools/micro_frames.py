"""Config C4 (Atari-shaped uint8 84x84x4 frame replay, one 125 000-transition shard of the 1e6 / 8 GPUs, batch 512):
gather throughput of the naive layout (obs1 + obs2 stored per transition) and the frame-deduplicated layout
(stack rebuilt from 5 consecutive frames), plus the N-step sequence ring (N3).  GPU box only."""
import json, os, sys
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import NStepReplayBuffer
from ddrl_b200.frames import FrameReplayBuffer

dev = torch.device("cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=30, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

res = []
shard, B, fb = 125_000, 512, 84 * 84
# dedup: one frame per transition
rb = FrameReplayBuffer((84, 84), 4, shard, mode="dedup", seed=1)
z25 = torch.zeros(25_000, device=dev)
for lo in range(0, shard, 25_000):
    fr = torch.randint(0, 256, (25_000, fb), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rb.store_frames(fr, z25, z25, z25); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
res.append(dict(cfg="C4 dedup store_frames", rows=25_000, us=t * 1e6, GBs=2 * 25_000 * (fb + 12) / t / 1e9, frac=2 * 25_000 * (fb + 12) / t / 1e9 / peak))
print(res[-1], flush=True)
for n_small in (256, 2048):
    t = timeit(lambda: rb.store_frames(fr[:n_small], z25[:n_small], z25[:n_small], z25[:n_small]))
    res.append(dict(cfg="C4 dedup store_frames", rows=n_small, us=t * 1e6, GBs=2 * n_small * (fb + 12) / t / 1e9, frac=2 * n_small * (fb + 12) / t / 1e9 / peak))
    print(res[-1], flush=True)
for nb in (512, 8192):
    t = timeit(lambda: rb.sample_batch(nb))
    bytes_alg = nb * (5 * fb + 12 + 2 * 4 * fb + 12)      # read 5 frames + scalars, write two stacks + scalars
    res.append(dict(cfg="C4 dedup", batch=nb, us=t * 1e6, Mtrans_s=nb / t / 1e6, GBs=bytes_alg / t / 1e9, frac=bytes_alg / t / 1e9 / peak))
    print(res[-1], flush=True)
del rb
# naive: stacked obs1 + obs2 per transition (56 448 B + scalars per row) -- 30 000 rows = 1.7 GB is enough to defeat L2
rbn = FrameReplayBuffer((84, 84), 4, 30_000, mode="naive", seed=1)
for lo in range(0, 30_000, 5_000):
    o = torch.randint(0, 256, (5_000, 4, 84, 84), dtype=torch.uint8, device=dev)
    rbn.store_batch(o, torch.zeros(5_000, device=dev), torch.zeros(5_000, device=dev), o, torch.zeros(5_000, device=dev))
for nb in (512, 8192):
    t = timeit(lambda: rbn.sample_batch(nb))
    bytes_alg = nb * 2 * (2 * 4 * fb + 12)
    res.append(dict(cfg="C4 naive", batch=nb, us=t * 1e6, Mtrans_s=nb / t / 1e6, GBs=bytes_alg / t / 1e9, frac=bytes_alg / t / 1e9 / peak))
    print(res[-1], flush=True)
del rbn
# N3: D=24, A=4, Ln=8 sequences (1056-byte rows)
opt = SimpleNamespace(Ln=8, obs_shape=(24,), act_shape=(4,), buffer_size=1_000_000, batch_size=1024, num_buffers=1)
nb_ = NStepReplayBuffer(opt, seed=3)
g = np.random.Generator(np.random.PCG64(0))
for lo in range(0, 1_000_000, 100_000):
    nb_.store_batch(g.standard_normal((100_000, 9, 24), dtype=np.float32), g.standard_normal((100_000, 8, 4), dtype=np.float32),
                    g.standard_normal((100_000, 8), dtype=np.float32), np.zeros((100_000, 8), np.float32))
dobs, dact, drew, ddone = (torch.randn((100_000, 9, 24), device=dev), torch.randn((100_000, 8, 4), device=dev), torch.randn((100_000, 8), device=dev),
                           torch.zeros((100_000, 8), device=dev))
for n_ in (1024, 100_000):
    t = timeit(lambda: nb_.store_batch(dobs[:n_], dact[:n_], drew[:n_], ddone[:n_]))
    bytes_alg = n_ * 2 * 4 * nb_.used
    res.append(dict(cfg="N3 nstep store_batch (device rows)", rows=n_, us=t * 1e6, GBs=bytes_alg / t / 1e9, frac=bytes_alg / t / 1e9 / peak))
    print(res[-1], flush=True)
for B_ in (1024, 262144):
    t = timeit(lambda: nb_.sample_batch(B_, device=True))
    bytes_alg = B_ * 2 * 4 * nb_.used
    res.append(dict(cfg="N3 nstep", batch=B_, us=t * 1e6, Mtrans_s=B_ / t / 1e6, GBs=bytes_alg / t / 1e9, frac=bytes_alg / t / 1e9 / peak))
    print(res[-1], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "micro_frames.json"), "w"), indent=1)
