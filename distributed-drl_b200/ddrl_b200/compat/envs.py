"""Synthetic environments with the gym call surface the reference drivers use (gym itself is not a dependency):
`gym.make(name)`, `env.reset() -> obs`, `env.step(a) -> (obs, reward, done, info)`, `env.observation_space.shape`,
`env.action_space.shape / .high / .low / .sample()`.  The dynamics are a tiny deterministic linear system
(x' = 0.9 x + 0.1 B a, reward = -|x|^2 / dim): enough for the drivers' control flow, throttling and episode logic."""
from __future__ import annotations

import numpy as np

# observation / action dimensions and action bound of the environments the reference's scripts name
ENV_SHAPES = {
    "LunarLanderContinuous-v2": (8, 2, 1.0),
    "BipedalWalker-v2": (24, 4, 1.0),
    "BipedalWalkerHardcore-v2": (24, 4, 1.0),
    "Pendulum-v0": (3, 1, 2.0),
    "Humanoid-v2": (376, 17, 0.4),
}


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, rng=None):
        shape = tuple(shape) if shape is not None else np.shape(low)
        self.shape, self.dtype = shape, dtype
        self.low = np.full(shape, low, dtype) if np.isscalar(low) else np.asarray(low, dtype)
        self.high = np.full(shape, high, dtype) if np.isscalar(high) else np.asarray(high, dtype)
        self._rng = rng or np.random.default_rng(0)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)


class SyntheticEnv:
    def __init__(self, obs_dim, act_dim, act_high=1.0, seed=0, horizon=50):
        self.rng = np.random.default_rng(seed)
        self.obs_dim, self.act_dim, self.horizon = obs_dim, act_dim, horizon
        self.B = self.rng.standard_normal((obs_dim, act_dim)) / np.sqrt(act_dim)
        self.action_space = Box(-act_high, act_high, (act_dim,), rng=self.rng)
        self.observation_space = Box(-np.inf, np.inf, (obs_dim,), rng=self.rng)
        self.t, self.x = 0, None

    def reset(self):
        self.t = 0
        self.x = self.rng.standard_normal(self.obs_dim)
        return self.x.copy()

    def step(self, a):
        self.x = 0.9 * self.x + 0.1 * self.B @ np.asarray(a, dtype=np.float64).reshape(-1)
        self.t += 1
        r = -float(self.x @ self.x) / self.obs_dim
        return self.x.copy(), r, self.t >= self.horizon, {}

    def render(self, *a, **k):
        return None

    def close(self):
        pass


_made = [0]


def make(name, **_):
    if name not in ENV_SHAPES:
        raise KeyError(f"ddrl_b200.compat.envs: no synthetic stand-in registered for gym environment {name!r}")
    d, a, hi = ENV_SHAPES[name]
    _made[0] += 1
    return SyntheticEnv(d, a, act_high=hi, seed=_made[0])
