"""GPU: Actor.get_action / get_actions (algos/sac1/actor_learner.py:195-197, vectorised: SURVEY.md row N4) against
the oracle's policy.  The driver flow itself (worker_train / worker_rollout / worker_test / Cache / __main__) is
exercised with the reference's OWN code in tests/test_compat_gpu.py."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_actor_matches_oracle_policy():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import Actor
    from oracle.sac1_oracle import SAC1Oracle, conditioned_params

    D, A, hid = 6, 2, (32, 32)
    space = SimpleNamespace(high=np.full(A, 0.5, np.float32), shape=(A,))
    opt = SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=hid, action_space=space), alpha=0.2, gamma=0.99,
                          lr=1e-3, polyak=0.995, seed=0, batch_size=64, max_ep_len=25)
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
    # an Actor view of a Learner shares its weights (example/model.py's Model both trains and acts)
    from ddrl_b200 import Learner
    L = Learner(opt, "model")
    L.set_weights(list(params), list(params.values()))
    view = Actor.from_learner(L)
    assert np.allclose(view.get_actions(obs, deterministic=True).cpu().numpy(), got_mu, atol=1e-6)
