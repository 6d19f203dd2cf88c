"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the replay ring,
the parameter server and the index stream.

What is restated, and from where (paths relative to /root/reference):

* ring construction            example/dsac.py:20-27   (= example/sac.py:15-21, algos/sac1/sac1.py:34-41)
* store (one transition)       example/dsac.py:29-37   (algos/sac1/sac1.py:43-51)
* sample_batch                 example/dsac.py:39-45   (algos/sac1/sac1.py:53-60: default 128, sample_times += 1)
* get_counts                   example/dsac.py:47-48 / algos/sac1/sac1.py:62-63 / algos/dqn/train.py:75-76
* dqn-family scalar action     algos/dqn/train.py:43-53 (acts_buf is 1-D)
* ParameterServer              algos/sac1/sac1.py:66-100 (= example/dsac.py:51-73 + weights_file restore)

Pinning: tests/test_oracle_replay.py drives this class and the reference classes (extracted from
the reference files by oracle/ref_extract.py, in the build container where /root/reference exists)
with the same store sequence and the same injected index stream and requires bit-identical ring
arrays, counters and sampled batches; the same comparison is frozen into tests/golden/replay_*.npz
(made by oracle/make_golden.py from the *reference* classes) so that it also runs on the GPU box.

The Philox4x32-10 generator below is NOT part of the reference (the reference draws indices with
numpy's global MT19937, example/dsac.py:40).  It restates the published Random123 algorithm
(Salmon et al., SC'11) and is pinned by Random123's known-answer vectors; it exists so that the
device-side index generator can be checked bit-exactly instead of only statistically.
"""
from __future__ import annotations

import pickle

import numpy as np

# --------------------------------------------------------------------------------------------
# Replay ring
# --------------------------------------------------------------------------------------------


class ReplayRingOracle:
    """FIFO ring of transitions held as five separate arrays, exactly like the reference.

    ``flavor`` selects which reference variant's counters are mimicked:
      "dsac"  -> example/dsac.py   (rollout_steps; get_counts() -> int)
      "sac1"  -> algos/sac1/sac1.py (steps, sample_times; get_counts() -> (sample_times, steps, size))
      "dqn"   -> algos/dqn/train.py (scalar action; get_counts() -> (learner_steps, actor_steps, size))
    """

    def __init__(self, obs_dim, act_dim, size, flavor="sac1", obs_dtype=np.float32):
        size = int(size)
        self.flavor = flavor
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.obs1_buf = np.zeros([size, obs_dim], dtype=obs_dtype)
        self.obs2_buf = np.zeros([size, obs_dim], dtype=obs_dtype)
        if flavor == "dqn":
            self.acts_buf = np.zeros(size, dtype=np.float32)  # algos/dqn/train.py:48
        else:
            self.acts_buf = np.zeros([size, act_dim], dtype=np.float32)
        self.rews_buf = np.zeros(size, dtype=np.float32)
        self.done_buf = np.zeros(size, dtype=np.float32)
        self.ptr, self.size, self.max_size = 0, 0, size
        self.steps = 0          # rollout_steps / steps / actor_steps
        self.sample_times = 0   # sample_times / learner_steps (not counted by the dsac flavor)

    # -- one transition (numpy assignment performs the dtype cast: f64->f32 RNE, bool->0/1) -----
    def store(self, obs, act, rew, next_obs, done):
        p = self.ptr
        self.obs1_buf[p] = obs
        self.obs2_buf[p] = next_obs
        self.acts_buf[p] = act
        self.rews_buf[p] = rew
        self.done_buf[p] = done
        self.ptr = (p + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)
        self.steps += 1

    # -- n transitions: defined as n sequential store() calls (the contract of rb_store_batch) --
    def store_batch(self, obs, act, rew, next_obs, done):
        n = len(rew)
        for i in range(n):
            self.store(obs[i], act[i], rew[i], next_obs[i], done[i])

    def draw_indices(self, batch_size):
        """The reference's own draw: global legacy RandomState, uniform on [0, size), with
        replacement, int64 (example/dsac.py:40).  Raises ValueError when the ring is empty."""
        return np.random.randint(0, self.size, size=batch_size)

    def sample_batch(self, batch_size=128, idxs=None):
        if idxs is None:
            idxs = self.draw_indices(batch_size)
        if self.flavor != "dsac":
            self.sample_times += 1
        return dict(obs1=self.obs1_buf[idxs],
                    obs2=self.obs2_buf[idxs],
                    acts=self.acts_buf[idxs],
                    rews=self.rews_buf[idxs],
                    done=self.done_buf[idxs])

    def get_counts(self):
        if self.flavor == "dsac":
            return self.steps
        return self.sample_times, self.steps, self.size


# --------------------------------------------------------------------------------------------
# Parameter server  (algos/sac1/sac1.py:66-100)
# --------------------------------------------------------------------------------------------


class ParameterServerOracle:
    def __init__(self, keys, values, weights_file=""):
        if weights_file:
            with open(weights_file, "rb") as f:
                self.weights = pickle.load(f)
        else:
            self.weights = {k: np.array(v, copy=True) for k, v in zip(keys, values)}

    def push(self, keys, values):
        for k, v in zip(keys, values):
            self.weights[k] = np.array(v, copy=True)

    def pull(self, keys):
        return [self.weights[k] for k in keys]

    def get_weights(self):
        return self.weights

    def save_weights(self, name):
        with open(name + "weights.pickle", "wb") as f:
            pickle.dump(self.weights, f)


# --------------------------------------------------------------------------------------------
# Philox4x32-10 (Random123) and the index stream derived from it
# --------------------------------------------------------------------------------------------

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  All arguments broadcastable uint32-valued arrays.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0           # < 2^64, exact in uint64
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK32
        hi1, lo1 = p1 >> _S32, p1 & _MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _mulhi64(a, b):
    """High 64 bits of the 128-bit product of uint64 arrays a and scalar b (pure integer)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.uint64(b)
    a_lo, a_hi = a & _MASK32, a >> _S32
    b_lo, b_hi = b & _MASK32, b >> _S32
    ll = a_lo * b_lo
    lh = a_lo * b_hi
    hl = a_hi * b_lo
    hh = a_hi * b_hi
    mid = (ll >> _S32) + (lh & _MASK32) + (hl & _MASK32)
    return hh + (lh >> _S32) + (hl >> _S32) + (mid >> _S32)


def philox_indices(n, size, seed, counter, stream=0):
    """Index stream of the CUDA sampler (csrc/replay.cu: philox_index):
    for ordinal i in [0, n):  (x0,x1,_,_) = philox4x32_10(ctr=(i, counter_lo, counter_hi, stream),
    key=(seed_lo, seed_hi));  u = x1<<32 | x0;  idx = floor(u * size / 2^64)  (Lemire multiply-shift,
    bias < size / 2^64).  Uniform on [0, size), with replacement, int64 — the distribution of
    the reference's np.random.randint(0, size, n) (example/dsac.py:40), not its bit stream."""
    if size <= 0:
        raise ValueError("high <= 0")
    i = np.arange(n, dtype=np.uint64)
    x0, x1, _, _ = philox4x32_10(i & _MASK32, counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF,
                                 stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = (x1.astype(np.uint64) << _S32) | x0.astype(np.uint64)
    return _mulhi64(u, size).astype(np.int64)


def philox_normals(n, seed, counter, stream=0):
    """Standard-normal stream of the CUDA learner (csrc/sac.cu: philox_normal4): ordinal j yields
    four normals by two Box-Muller pairs from (x0,x1) and (x2,x3):
        u = (x + 0.5) * 2^-32  in (0,1);  r = sqrt(-2 ln u_a);  z = r*cos(2 pi u_b), r*sin(2 pi u_b).
    Returned as float64 [n] (n padded up to a multiple of 4 internally); device evaluates in fp32,
    so comparisons are approximate (statistical), not bit-exact."""
    m = (n + 3) // 4
    j = np.arange(m, dtype=np.uint64)
    x = philox4x32_10(j & _MASK32, counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF,
                      stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [(xi.astype(np.float64) + 0.5) * 2.0 ** -32 for xi in x]
    r0 = np.sqrt(-2.0 * np.log(u[0]))
    r1 = np.sqrt(-2.0 * np.log(u[2]))
    z = np.stack([r0 * np.cos(2 * np.pi * u[1]), r0 * np.sin(2 * np.pi * u[1]),
                  r1 * np.cos(2 * np.pi * u[3]), r1 * np.sin(2 * np.pi * u[3])], axis=1)
    return z.reshape(-1)[:n]
