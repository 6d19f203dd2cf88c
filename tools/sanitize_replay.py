"""Small invocations of every replay kernel family for compute-sanitizer (GPU box only): narrow / bulk / wide-TMA gathers,
staged and generic stores, frame ring store + TMA stack gather, N-step segment store + gather, concurrent producer."""
import os, sys, threading
from types import SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "distributed-drl_b200")]
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from ddrl_b200 import NStepReplayBuffer, ReplayBuffer
from ddrl_b200.frames import FrameReplayBuffer

dev = torch.device("cuda")
g = np.random.Generator(np.random.PCG64(0))
for mode in os.environ.get("SAN_MODES", "0,1,2,3").split(","):
    os.environ["DDRL_GATHER_MODE"] = mode
    for D, A, cap, B in ((8, 2, 3000, 256), (24, 4, 3000, 512), (376, 17, 600, 128), (5, 3, 100, 37)):
        rb = ReplayBuffer(D, A, cap, seed=1)
        n = cap + 57
        rb.store_batch(g.standard_normal((n, D)), g.uniform(-1, 1, (n, A)), g.standard_normal(n), g.standard_normal((n, D)), g.random(n) < 0.1)
        rb.store_batch(torch.randn(64, D, device=dev), torch.rand(64, A, device=dev), torch.randn(64, device=dev), torch.randn(64, D, device=dev),
                       torch.zeros(64, device=dev))
        rb.sample_batch(B); rb.sample_batch(B, device=True); rb.sample_many(4, B)
        rb.sample_batch(16, idxs=np.arange(16))
os.environ["DDRL_GATHER_MODE"] = "0"
fb = FrameReplayBuffer((84, 84), 4, 64, mode="dedup", seed=1)
fb.store_frames(g.integers(0, 256, (100, 84 * 84), dtype=np.uint8), np.zeros(100), np.zeros(100), np.zeros(100))
fb.sample_batch(32); fb.sample_batch(8, idxs=np.arange(10, 18))
fn = FrameReplayBuffer((84, 84), 4, 16, mode="naive", seed=1)
o = g.integers(0, 256, (20, 4, 84, 84), dtype=np.uint8)
fn.store_batch(o, np.zeros(20), np.zeros(20), o, np.zeros(20)); fn.sample_batch(8)
opt = SimpleNamespace(Ln=3, obs_shape=(5,), act_shape=(3,), buffer_size=32, batch_size=16, num_buffers=1)
nb = NStepReplayBuffer(opt, seed=2)
nb.store_batch(g.standard_normal((50, 4, 5)), g.standard_normal((50, 3, 3)), g.standard_normal((50, 3)), np.zeros((50, 3)))
nb.sample_batch(device=True)
rb = ReplayBuffer(12, 3, 512, seed=3)
rb.store_batch(torch.randn(256, 12, device=dev), torch.rand(256, 3, device=dev), torch.randn(256, device=dev), torch.randn(256, 12, device=dev), torch.zeros(256, device=dev))
def prod():
    with torch.cuda.stream(torch.cuda.Stream()):
        for _ in range(10):
            rb.store_batch(np.zeros((32, 12), np.float32), np.zeros((32, 3), np.float32), np.zeros(32, np.float32), np.zeros((32, 12), np.float32), np.zeros(32, np.float32))
t = threading.Thread(target=prod); t.start()
for _ in range(10):
    rb.sample_batch(64, device=True)
t.join()
torch.cuda.synchronize()
print("sanitize_replay done")
