"""Frame-stack replay for Atari-shaped observations (BASELINE.json config 4: uint8 84x84x4, batch 512).

The reference has no uint8 frame buffer: its dqn-family ring
(algos/dqn/train.py:37-80) stores float32 vectors and stacks frames env-side
(algos/trading_env.py:289-325).  Two layouts are offered with that ring's call surface
(store(obs, act, rew, next_obs, done) / sample_batch(batch_size) -> dict(obs1, obs2, acts, rews, done)):

  FrameReplayBuffer(mode="naive")  stores the stacked obs1 and obs2 of every transition (2*28224 B per
      row) — the packed-row ring of ddrl_b200.ReplayBuffer with the uint8 rows viewed as 4-byte words
      (a bit copy, so gathers are exact by construction);
  FrameReplayBuffer(mode="dedup")  stores ONE new frame per transition and rebuilds the two stacks at
      sample time from stack+1 consecutive frames (ddrl_fb_sample_stack): 8x less HBM per transition.
      Like most Atari replays it ignores episode boundaries inside a stack; sampled windows never cross the write head.

ShardedFrameReplayBuffer (ddrl_b200.dist) is the configured C4 shape: 1e6 transitions as one shard per GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .replay import ReplayBuffer


class FrameReplayBuffer:
    def __init__(self, frame_shape=(84, 84), stack=4, size=100_000, *, mode="dedup", device=None, seed=None, rng_stream=0):
        if mode not in ("naive", "dedup"):
            raise ValueError(mode)
        if not torch.cuda.is_available():
            raise RuntimeError("ddrl_b200.FrameReplayBuffer needs a CUDA device (no CPU fallback)")
        self.mode, self.stack = mode, int(stack)
        self.frame_shape = tuple(int(x) for x in frame_shape)
        self.frame_bytes = int(np.prod(self.frame_shape))
        self.obs_bytes = self.frame_bytes * self.stack
        self.max_size = int(size)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._dev = torch.device("cuda", self.device)
        self._lib = N.lib()
        if mode == "naive":
            if self.obs_bytes % 4:
                raise ValueError("stacked observation bytes must be a multiple of 4")
            self._rb = ReplayBuffer(self.obs_bytes // 4, 1, size, device=self.device, flavor="dqn", seed=seed,
                                    rng_stream=rng_stream)
        else:
            if self.frame_bytes % 16:
                raise ValueError("frame bytes must be a multiple of 16")
            self.frames = torch.zeros((self.max_size, self.frame_bytes), dtype=torch.uint8, device=self._dev)
            self.act = torch.zeros(self.max_size, dtype=torch.float32, device=self._dev)
            self.rew = torch.zeros_like(self.act)
            self.done = torch.zeros_like(self.act)
            self.ptr = self.size = self.steps = self.sample_times = 0
            self._seed = int(seed) if seed is not None else None
            self._rng_stream, self._counter = int(rng_stream), 0

    # ---- store -----------------------------------------------------------------------------------
    def _u8(self, x, n, width):
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.uint8)) if not isinstance(x, torch.Tensor) else x
        return t.reshape(n, width).to(self._dev, torch.uint8).contiguous()

    def store_batch(self, obs, act, rew, next_obs, done):
        """naive: obs/next_obs are the stacked uint8 observations [n, stack, H, W].
        dedup: only next_obs[:, -1] (the newest frame) is stored; `obs` is ignored after the first row
        of an episode has seeded the ring (call seed_frames for that)."""
        n = int(np.asarray(rew).shape[0]) if not isinstance(rew, torch.Tensor) else int(rew.shape[0])
        f = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float32)).to(self._dev) if not isinstance(v, torch.Tensor) else v.to(self._dev, torch.float32)
        if self.mode == "naive":
            o = self._u8(obs, n, self.obs_bytes).view(torch.float32)
            o2 = self._u8(next_obs, n, self.obs_bytes).view(torch.float32)
            return self._rb.store_batch(o, f(act).reshape(n, 1), f(rew), o2, f(done))
        newest = next_obs.reshape(n, self.stack, self.frame_bytes)[:, -1] if isinstance(next_obs, torch.Tensor) else \
            np.asarray(next_obs).reshape(n, self.stack, self.frame_bytes)[:, -1]
        self.store_frames(newest, act, rew, done)

    def store_frames(self, frames, act, rew, done):
        """dedup layout: append n frames (frame t+1 of each transition) and the transition scalars of
        the step that produced them; transition i = (stack ending at frame i) -> (stack ending at i+1),
        so act/rew/done of that step are stored at slot i = position of the previous frame.  One launch of
        fb_store_frames (ddrl_fb_store_frames); host arrays are uploaded first, CUDA tensors are consumed in place."""
        fr = self._u8(frames, -1, self.frame_bytes)
        n = int(fr.shape[0])
        if n == 0:
            return
        f = lambda v: (torch.as_tensor(np.asarray(v, dtype=np.float32)).reshape(-1).to(self._dev) if not isinstance(v, torch.Tensor)
                       else v.reshape(-1).to(self._dev, torch.float32)).contiguous()
        a, r, d = f(act), f(rew), f(done)
        if not (a.numel() == r.numel() == d.numel() == n):
            raise ValueError("store_frames: frames / act / rew / done disagree on n")
        s = torch.cuda.current_stream(self.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        N.check(self._lib.ddrl_fb_store_frames(self.device, p(self.frames), self.frame_bytes, self.max_size, self.ptr,
                                               p(self.act), p(self.rew), p(self.done), p(fr), p(a), p(r), p(d), n,
                                               C.c_void_p(s.cuda_stream)))
        for t in (fr, a, r, d):
            t.record_stream(s)
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)
        self.steps += n

    def store(self, obs, act, rew, next_obs, done):
        one = lambda x: np.asarray(x)[None]
        self.store_batch(one(obs), one(act), one(rew), one(next_obs), one(done))

    # ---- sample ------------------------------------------------------------------------------------
    def sample_batch(self, batch_size=512, *, idxs=None, return_idxs=False):
        B = int(batch_size)
        shape = (B, self.stack) + self.frame_shape
        if self.mode == "naive":
            out = self._rb.sample_batch(B, idxs=idxs, device=True, return_idxs=return_idxs)
            out["obs1"] = out["obs1"].view(torch.uint8).reshape(shape)
            out["obs2"] = out["obs2"].view(torch.uint8).reshape(shape)
            return out
        if self.size < self.stack + 1:
            raise ValueError("high <= 0")
        # two allocations for the five outputs (allocator calls dominate the host time of a 512-row batch)
        obs = torch.empty((2, B) + (self.stack,) + self.frame_shape, dtype=torch.uint8, device=self._dev)
        sc = torch.empty((3, B), dtype=torch.float32, device=self._dev)
        oi = torch.empty(B, dtype=torch.int64, device=self._dev) if return_idxs else None
        di = None
        if idxs is not None:
            di = torch.as_tensor(idxs, dtype=torch.int64).reshape(-1).to(self._dev).contiguous()
        if self._seed is None:
            self._seed = int(np.random.randint(0, 2 ** 31 - 1))
        s = torch.cuda.current_stream(self.device)
        p = lambda t: t.data_ptr() if t is not None else None
        oldest = self.ptr if self.size == self.max_size else 0     # Philox draws ages from the oldest frame: no window
        ob, sb = obs.data_ptr(), sc.data_ptr()                     # crosses the write head
        N.check(self._lib.ddrl_fb_sample_stack(self.device, self.frames.data_ptr(), self.frame_bytes, self.stack, self.max_size,
                                               self.size, oldest, self.act.data_ptr(), self.rew.data_ptr(), self.done.data_ptr(),
                                               B, p(di), self._seed, self._counter, self._rng_stream, ob, ob + B * self.obs_bytes,
                                               sb, sb + 4 * B, sb + 8 * B, p(oi), s.cuda_stream))
        if idxs is None:
            self._counter += 1
        self.sample_times += 1
        o1, o2 = obs.unbind(0)
        oa, orw, od = sc.unbind(0)
        out = dict(obs1=o1, obs2=o2, acts=oa, rews=orw, done=od)
        if return_idxs:
            out["idxs"] = oi
        return out

    def get_counts(self):
        if self.mode == "naive":
            return self._rb.get_counts()
        return self.sample_times, self.steps, self.size
