"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's N-step sequence ring
(algos/sac1/sac_ray.py:34-83, float32 observation layout `opt.obs_shape == (115,)` generalised to any shape).

Pinned: tests/test_oracle_nstep.py drives this class and the reference's own class (loaded by
oracle/ref_extract.py from /root/reference) with the same stores and the same injected index stream and
compares all four arrays and the counters; tests/golden/nstep_*.npz holds outputs generated from the reference
class by oracle/make_golden.py.
"""
from __future__ import annotations

import numpy as np


class NStepRingOracle:
    def __init__(self, opt):
        self.opt = opt
        # sac_ray.py:45-48
        self.buffer_o = np.zeros((opt.buffer_size, opt.Ln + 1) + tuple(opt.obs_shape), dtype=np.float32)
        self.buffer_a = np.zeros((opt.buffer_size, opt.Ln) + tuple(opt.act_shape), dtype=np.float32)
        self.buffer_r = np.zeros((opt.buffer_size, opt.Ln), dtype=np.float32)
        self.buffer_d = np.zeros((opt.buffer_size, opt.Ln), dtype=np.float32)
        self.ptr, self.size, self.max_size = 0, 0, opt.buffer_size      # :49
        self.steps, self.sample_times = 0, 0                            # :50

    def store(self, o_queue, a_r_d_queue, worker_index=0):
        obs, = np.stack(o_queue, axis=1)                                # :54
        self.buffer_o[self.ptr] = np.array(list(obs), dtype=np.float32)  # :59
        # :61 `a, r, d, = np.stack(a_r_d_queue, axis=1)` — column extraction; written with zip because numpy >= 1.24
        # refuses the ragged (array, float, bool) tuples a vector action produces (identical for scalar actions,
        # which is what the pinning test drives through the reference's own class)
        a, r, d = zip(*a_r_d_queue)
        self.buffer_a[self.ptr] = np.array(list(a), dtype=np.float32)
        self.buffer_r[self.ptr] = np.array(list(r), dtype=np.float32)
        self.buffer_d[self.ptr] = np.array(list(d), dtype=np.float32)
        self.ptr = (self.ptr + 1) % self.max_size                       # :66
        self.size = min(self.size + 1, self.max_size)
        self.steps += 1 * self.opt.num_buffers                          # :69

    def sample_batch(self, idxs=None):
        if idxs is None:
            idxs = np.random.randint(0, self.size, size=self.opt.batch_size)   # :73
        self.sample_times += 1 * self.opt.num_buffers                   # :75
        return dict(obs=self.buffer_o[idxs], acts=self.buffer_a[idxs], rews=self.buffer_r[idxs], done=self.buffer_d[idxs])

    def get_counts(self):
        return self.sample_times, self.steps, self.size                 # :82-83


def make_sequences(opt, n, seed):
    """n synthetic (o_queue, a_r_d_queue) pairs in the shape worker_rollout builds them (sac_ray.py:282-300):
    o_queue = Ln+1 entries (obs,), a_r_d_queue = Ln entries (action, reward, done)."""
    g = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(n):
        oq = [(g.standard_normal(tuple(opt.obs_shape)).astype(np.float32),) for _ in range(opt.Ln + 1)]
        aq = [(g.uniform(-1, 1, tuple(opt.act_shape)).astype(np.float32), float(g.standard_normal()), bool(g.random() < 0.1))
              for _ in range(opt.Ln)]
        out.append((oq, aq))
    return out


def pack_rows(obs, acts, rews, done, Ln, D, A):
    """Expected ring rows of ddrl_b200.NStepReplayBuffer: [n, Ln+1, ...], [n, Ln, ...], [n, Ln], [n, Ln] -> packed float32
    rows [n, row_f] = [obs | acts | rews | done | 0-pad to 4 floats] (numpy assignment casts, as the reference's
    `buffer[ptr] = np.array(..., dtype=np.float32)` does, sac_ray.py:52-68)."""
    widths = [(Ln + 1) * D, Ln * A, Ln, Ln]
    o = [0, widths[0], widths[0] + widths[1], widths[0] + widths[1] + Ln]
    row_f = (sum(widths) + 3) // 4 * 4
    n = int(np.asarray(rews).shape[0])
    rows = np.zeros((n, row_f), dtype=np.float32)
    rows[:, o[0]:o[0] + widths[0]] = np.asarray(obs).reshape(n, -1)
    rows[:, o[1]:o[1] + widths[1]] = np.asarray(acts).reshape(n, -1)
    rows[:, o[2]:o[2] + Ln] = np.asarray(rews).reshape(n, -1)
    rows[:, o[3]:o[3] + Ln] = np.asarray(done).reshape(n, -1)
    return rows
