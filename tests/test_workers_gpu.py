"""GPU: the reference's driver flow (algos/sac1/sac1.py:255-280) end to end on the ray stand-in:
ParameterServer + ReplayBuffer actors, rollout workers storing one transition per env step, a learner
sampling / training / pushing weights, the tester pulling them — with a synthetic environment."""
import threading
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_actor_matches_oracle_policy_and_driver_flow_runs():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import Actor, Learner, ParameterServer, ReplayBuffer, ray_shim as ray
    from ddrl_b200.workers import SyntheticEnv, worker_rollout, worker_test, worker_train
    from oracle.sac1_oracle import SAC1Oracle, conditioned_params

    D, A, hid = 6, 2, (32, 32)
    space = SimpleNamespace(high=np.full(A, 0.5, np.float32), shape=(A,))
    opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                          lr=1e-3, polyak=0.995, seed=0, batch_size=64, start_steps=50, max_ep_len=25, a_l_ratio=4,
                          push_freq=20, total_learner_steps=120, total_env_steps=600, test_episodes=2,
                          env_fn=lambda o: SyntheticEnv(D, A, act_high=0.5, seed=1), stop=threading.Event())

    # --- Actor.get_action against the oracle's policy (deterministic and with injected noise)
    params = conditioned_params(D, A, hid, seed=2)
    actor = Actor(opt, "worker")
    actor.set_weights(list(params), list(params.values()))
    ora = SAC1Oracle(D, A, hidden=hid, params=params, act_scale=0.5)
    obs = np.random.default_rng(0).standard_normal((9, D)).astype(np.float32)
    eps = np.random.default_rng(1).standard_normal((9, A)).astype(np.float32)
    mu, pi, _ = ora.policy(ora.main, torch.tensor(obs, dtype=torch.float64), torch.tensor(eps, dtype=torch.float64))
    got_mu = actor.get_actions(obs, deterministic=True).cpu().numpy()
    got_pi = actor.get_actions(obs, noise=eps).cpu().numpy()
    assert np.abs(got_mu - 0.5 * mu.detach().numpy()).max() <= 1e-5
    assert np.abs(got_pi - 0.5 * pi.detach().numpy()).max() <= 1e-5
    one = actor.get_action(obs[0], deterministic=True)
    assert one.shape == (A,) and np.allclose(one, got_mu[0])
    keys, vals = actor.get_weights()
    assert len(keys) == 8 and all("/pi/" in k for k in keys)          # the actor only knows the policy

    # --- driver flow
    ray.init()
    net = Learner(opt, job="main")
    all_keys, all_values = net.get_weights()
    ps = ray.remote(ParameterServer).remote(all_keys, all_values)
    rb = ray.remote(ReplayBuffer).remote(obs_dim=D, act_dim=A, size=5000)
    rollouts = [ray.remote(worker_rollout).remote(ps, rb, opt, i) for i in range(2)]
    import time
    while ray.get(rb.get_counts.remote())[1] < 200:                  # the reference sleeps 5 s here (sac1.py:274)
        time.sleep(0.01)
    train = ray.remote(worker_train).remote(ps, rb, opt, 0)
    steps_done = ray.get(train, timeout=120)
    opt.stop.set()
    ray.get(rollouts, timeout=60)
    assert steps_done == 120
    sample_times, steps, size = ray.get(rb.get_counts.remote())
    assert sample_times == 120 and steps == size and steps >= 200
    pushed = ray.get(ps.pull.remote(all_keys))
    assert any(not np.array_equal(a, b) for a, b in zip(pushed, all_values))     # the learner's weights arrived
    res = ray.get(ray.remote(worker_test).remote(ps, rb, opt))
    assert len(res) == 1 and np.isfinite(res[0][0])
