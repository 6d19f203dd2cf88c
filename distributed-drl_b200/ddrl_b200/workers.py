"""worker_train / worker_rollout / worker_test — the reference's role tasks (algos/sac1/sac1.py:133-252,
3-argument forms example/dsac.py:76-177), written against the `ray` calling convention so that they run
under real Ray or under ddrl_b200.ray_shim.  Same control flow: the learner pulls weights, trains on
sampled batches and pushes every `push_freq` (300, sac1.py:149) steps; a rollout worker acts, stores
one transition per env step, throttles on steps / sample_times > a_l_ratio (sac1.py:203-207) and pulls
weights at every episode end.

Differences, all additive: the loops stop when `opt.total_steps` learner steps / env steps are reached
or `opt.stop` (a threading.Event) is set, so tests terminate; the learner's batches stay on the GPU
(`sample_batch(B, device=True)` — the role of the reference's Cache prefetch process, sac1.py:103-130);
`env_fn(opt)` builds the environment (gym is not a dependency).
"""
from __future__ import annotations

import time

import numpy as np

from .learner import Actor, Learner


def _ray():
    try:
        import ray  # noqa: F401
        return ray
    except Exception:
        from . import ray_shim
        return ray_shim


def _stopped(opt):
    ev = getattr(opt, "stop", None)
    return ev is not None and ev.is_set()


def worker_train(ps, replay_buffer, opt, learner_index=0):
    ray = _ray()
    agent = Learner(opt, job="learner")
    keys = agent.get_weights()[0]
    weights = ray.get(ps.pull.remote(keys))
    agent.set_weights(keys, weights)
    push_freq = int(getattr(opt, "push_freq", 300))
    total = int(getattr(opt, "total_learner_steps", 0))
    # the replay actor lives in this process: sample on the GPU through its handle when possible
    direct = getattr(replay_buffer, "_obj", None)
    cnt = 1
    while not _stopped(opt):
        if direct is not None and hasattr(direct, "sample_batch"):
            with replay_buffer._lock:
                batch = direct.sample_batch(opt.batch_size, device=True)
        else:
            batch = ray.get(replay_buffer.sample_batch.remote(opt.batch_size))
        agent.train(batch)
        if cnt % push_freq == 0:
            k, v = agent.get_weights()
            ps.push.remote(k, v)
        if total and cnt >= total:
            k, v = agent.get_weights()
            ps.push.remote(k, v)
            break
        cnt += 1
    return cnt


def worker_rollout(ps, replay_buffer, opt, worker_index=0):
    ray = _ray()
    env = opt.env_fn(opt)
    agent = Actor(opt, job="worker")
    keys = agent.get_weights()[0]
    o, r, d, ep_ret, ep_len = env.reset(), 0, False, 0, 0
    weights = ray.get(ps.pull.remote(keys))
    agent.set_weights(keys, weights)
    total = int(getattr(opt, "total_env_steps", 0))
    t = 0
    while not _stopped(opt):
        if t > opt.start_steps:
            a = agent.get_action(o)
        else:
            a = env.action_space.sample()
        t += 1
        o2, r, d, _ = env.step(a)
        ep_ret += r
        ep_len += 1
        d = False if ep_len == opt.max_ep_len else d      # time-limit dones are stored as 0 (sac1.py:192)
        replay_buffer.store.remote(o, a, r, o2, d)
        o = o2
        if d or (ep_len == opt.max_ep_len):
            sample_times, steps, _ = ray.get(replay_buffer.get_counts.remote())
            while sample_times > 0 and steps / sample_times > opt.a_l_ratio and not _stopped(opt):
                sample_times, steps, _ = ray.get(replay_buffer.get_counts.remote())
                time.sleep(0.01)
            weights = ray.get(ps.pull.remote(keys))
            agent.set_weights(keys, weights)
            o, r, d, ep_ret, ep_len = env.reset(), 0, False, 0, 0
        if total and t >= total:
            break
    return t


def worker_test(ps, replay_buffer, opt, rounds=1):
    ray = _ray()
    agent = Actor(opt, job="main")
    keys, _ = agent.get_weights()
    env = opt.env_fn(opt)
    out = []
    for _ in range(rounds):
        weights = ray.get(ps.pull.remote(keys))
        agent.set_weights(keys, weights)
        ep_ret = agent.test(env, replay_buffer, n=int(getattr(opt, "test_episodes", 2)))
        sample_times, steps, size = ray.get(replay_buffer.get_counts.remote())
        out.append((ep_ret, sample_times, steps, size))
    return out


class SyntheticEnv:
    """Tiny deterministic continuous-control environment with the gym call surface (reset/step/
    action_space.sample/high) used where gym is unavailable: x' = 0.9 x + 0.1 B a, reward = -|x|^2."""

    class _Space:
        def __init__(self, dim, high, rng):
            self.shape, self.high, self.low, self._rng = (dim,), np.full(dim, high, np.float32), np.full(dim, -high, np.float32), rng

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(np.float32)

    def __init__(self, obs_dim, act_dim, act_high=1.0, seed=0, horizon=50):
        self.rng = np.random.default_rng(seed)
        self.obs_dim, self.act_dim, self.horizon = obs_dim, act_dim, horizon
        self.B = self.rng.standard_normal((obs_dim, act_dim)) / np.sqrt(act_dim)
        self.action_space = self._Space(act_dim, act_high, self.rng)
        self.t, self.x = 0, None

    def reset(self):
        self.t = 0
        self.x = self.rng.standard_normal(self.obs_dim)
        return self.x.copy()

    def step(self, a):
        self.x = 0.9 * self.x + 0.1 * self.B @ np.asarray(a, dtype=np.float64)
        self.t += 1
        r = -float(self.x @ self.x) / self.obs_dim
        return self.x.copy(), r, self.t >= self.horizon and False, {}
