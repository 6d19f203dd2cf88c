"""TEST INFRASTRUCTURE ONLY — numpy restatement of the frame-stack replay of BASELINE.json config 4.

The reference has no uint8 frame replay; this oracle states the semantics the
CUDA path must match, built from the reference's ring (algos/dqn/train.py:37-80: FIFO ring, scalar
action, uniform sampling with replacement):
  naive : the same five-array ring with dtype=uint8 observation rows (ReplayRingOracle, flavor "dqn");
  dedup : a ring of single frames; transition i = (frames[i-S+1..i]) -> (frames[i-S+2..i+1]) with the
          scalars of the step that produced frame i+1 stored at slot i; no episode-boundary handling.
PARITY UNPINNED by the reference for the dedup layout (nothing to pin against); the naive layout
inherits the pinning of ReplayRingOracle.
"""
import numpy as np


class FrameRingOracle:
    def __init__(self, frame_bytes, stack, size):
        self.frames = np.zeros((size, frame_bytes), np.uint8)
        self.act = np.zeros(size, np.float32)
        self.rew = np.zeros(size, np.float32)
        self.done = np.zeros(size, np.float32)
        self.ptr, self.size, self.max_size, self.stack = 0, 0, size, stack

    def store_frames(self, frames, act, rew, done):
        for fr, a, r, d in zip(frames, act, rew, done):
            self.frames[self.ptr] = fr
            prev = (self.ptr - 1) % self.max_size
            self.act[prev], self.rew[prev], self.done[prev] = a, r, d
            self.ptr = (self.ptr + 1) % self.max_size
            self.size = min(self.size + 1, self.max_size)

    def drawn_indices(self, ages):
        """Ring positions of the transitions whose AGE (0 = the oldest complete window) is given: the sampler draws ages
        uniformly over [0, size - S) so that no window crosses the write head once the ring has wrapped."""
        oldest = self.ptr if self.size == self.max_size else 0
        return (oldest + (self.stack - 1) + np.asarray(ages)) % self.max_size

    def sample_batch(self, idxs):
        S, cap = self.stack, self.max_size
        idxs = np.asarray(idxs)
        win = lambda off: np.stack([self.frames[(idxs - (S - 1) + off + k) % cap] for k in range(S)], axis=1)
        return dict(obs1=win(0), obs2=win(1), acts=self.act[idxs], rews=self.rew[idxs], done=self.done[idxs])
