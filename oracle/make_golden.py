"""TEST INFRASTRUCTURE ONLY — regenerate tests/golden/replay_*.npz and ps_*.npz from the
REFERENCE's own classes (oracle/ref_extract.py; needs /root/reference, i.e. the build container).

    python -m oracle.make_golden

Each replay fixture records a seeded store sequence (mixed input dtypes, as gym would hand them
over: float64 observations, float32 actions, Python-float rewards, bool dones), the reference
ring state afterwards, an index stream and the reference's sampled batch for it.  The fixtures are
self-generated: the reference ships no tests or golden vectors of its own.
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_extract

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: (variant, obs_dim, act_dim, capacity, n_store, batch, seed)
REPLAY_CASES = {
    "c1_wrap":      ("sac1", 8, 2, 50, 137, 64, 1001),     # LunarLander-shaped rows, wraps 2.7x
    "c2_wrap":      ("sac1", 24, 4, 64, 100, 128, 1002),   # BipedalWalker-shaped rows
    "c3_wide":      ("sac1", 376, 17, 16, 40, 32, 1003),   # Humanoid-shaped rows (odd act_dim)
    "odd_exact":    ("dsac", 3, 1, 7, 7, 16, 1004),        # D%4!=0, exactly full, ptr back at 0
    "partial":      ("sac", 5, 3, 20, 5, 33, 1005),        # size < capacity: only [0,5) sampled
    "single_row":   ("sac1", 8, 2, 1, 3, 4, 1006),         # capacity 1
}


def make_inputs(obs_dim, act_dim, n, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    obs = g.standard_normal((n, obs_dim))                      # float64, cast on store
    nxt = g.standard_normal((n, obs_dim))
    act = g.uniform(-1, 1, (n, act_dim)).astype(np.float32)
    rew = g.standard_normal(n)                                  # float64 -> python float
    done = g.random(n) < 0.1                                    # bool
    return obs, act, rew, nxt, done


def run_reference(variant, obs_dim, act_dim, capacity, n_store, batch, seed):
    cls = ref_extract.reference_replay(variant)
    buf = cls(obs_dim, act_dim, capacity)
    obs, act, rew, nxt, done = make_inputs(obs_dim, act_dim, n_store, seed)
    for i in range(n_store):
        buf.store(obs[i], act[i], float(rew[i]), nxt[i], bool(done[i]))
    g = np.random.Generator(np.random.PCG64(seed + 7))
    idxs = g.integers(0, buf.size, size=batch, dtype=np.int64)
    # inject the index stream exactly where the reference draws it (example/dsac.py:40)
    saved = np.random.randint
    np.random.randint = lambda lo, hi, size=None: idxs
    try:
        out = buf.sample_batch(batch)
    finally:
        np.random.randint = saved
    counts = np.atleast_1d(np.asarray(buf.get_counts(), dtype=np.int64)) if hasattr(buf, "get_counts") \
        else np.zeros(0, np.int64)
    return dict(
        variant=np.array(variant), obs_dim=obs_dim, act_dim=act_dim, capacity=capacity,
        in_obs=obs, in_act=act, in_rew=rew, in_next=nxt, in_done=done, idxs=idxs,
        ring_obs1=buf.obs1_buf, ring_obs2=buf.obs2_buf, ring_acts=buf.acts_buf,
        ring_rews=buf.rews_buf, ring_done=buf.done_buf,
        ptr=buf.ptr, size=buf.size, counts=counts,
        out_obs1=out["obs1"], out_obs2=out["obs2"], out_acts=out["acts"],
        out_rews=out["rews"], out_done=out["done"])


def run_reference_ps(seed=2001):
    cls = ref_extract.reference_ps("sac1")
    g = np.random.Generator(np.random.PCG64(seed))
    keys = ["main/pi/dense/kernel", "main/pi/dense/bias", "main/q1/dense/kernel"]
    v0 = [g.standard_normal((4, 3)).astype(np.float32), g.standard_normal(3).astype(np.float32),
          g.standard_normal((5, 3)).astype(np.float32)]
    v1 = [g.standard_normal((4, 3)).astype(np.float32), g.standard_normal((5, 3)).astype(np.float32)]
    ps = cls(keys, v0)
    ps.push([keys[0], keys[2]], v1)
    pulled = ps.pull([keys[2], keys[1], keys[0]])
    return dict(keys=np.array(keys), init0=v0[0], init1=v0[1], init2=v0[2], push0=v1[0], push2=v1[1],
                pull_k2=pulled[0], pull_k1=pulled[1], pull_k0=pulled[2])


def run_reference_nstep():
    """tests/golden/nstep_scalar_act.npz: the reference's N-step ring (algos/sac1/sac_ray.py:34-83) driven with
    the sequences and index stream of tests/test_oracle_nstep.py::drive."""
    from types import SimpleNamespace
    from oracle.nstep_oracle import make_sequences
    cls = ref_extract.load_reference_class("algos/sac1/sac_ray.py", "ReplayBuffer")
    opt = SimpleNamespace(Ln=8, obs_shape=(115,), act_shape=(), buffer_size=37, batch_size=64, num_buffers=3)
    buf = cls(opt)
    for oq, aq in make_sequences(opt, 50, 5):
        buf.store(oq, aq, 0)
    idx = np.random.Generator(np.random.PCG64(6)).integers(0, 37, 64)
    saved = np.random.randint
    try:
        np.random.randint = lambda lo, hi, size: idx
        out = buf.sample_batch()
    finally:
        np.random.randint = saved
    return dict(idx=idx, obs=out["obs"], acts=out["acts"], rews=out["rews"], done=out["done"],
                counts=np.array(buf.get_counts()))


def main():
    if not ref_extract.reference_available():
        raise SystemExit("reference tree not found; golden vectors can only be regenerated where it is mounted")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, case in REPLAY_CASES.items():
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"replay_{name}.npz"), **run_reference(*case))
        print("wrote", name)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ps_sac1.npz"), **run_reference_ps())
    print("wrote ps_sac1")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "nstep_scalar_act.npz"), **run_reference_nstep())
    print("wrote nstep_scalar_act")


if __name__ == "__main__":
    main()
