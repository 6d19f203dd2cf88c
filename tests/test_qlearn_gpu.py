"""GPU parity of the DDQN / SQN learner steps (csrc/qlearn.cu through ddrl_b200.DQNLearner / SQNLearner) against the
float64 oracle (oracle/qlearn_oracle.py, parity unpinned — TensorFlow 1.x): losses and Q values to 1e-5, gradients to 2e-5
of their maximum, updated main / target weights to 1e-5 of max|w| where the gradient is not epsilon-dominated (the same
statement of the bar as tests/test_sac_gpu.py; plain FFMA arithmetic, no tensor-core rounding)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.qlearn_oracle import DDQNOracle, SQNOracle, init_q_params, make_q_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def QL():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import DQNLearner, SQNLearner
    return dict(ddqn=DQNLearner, sqn=SQNLearner)


def opt_of(D, nA, hidden, B, **kw):
    return SimpleNamespace(obs_dim=D, act_dim=nA, hidden_size=list(hidden), gamma=0.99, lr=1e-3, polyak=0.995, seed=0, batch_size=B,
                           alpha=kw.get("alpha", 0.1))


@pytest.mark.parametrize("kind,D,nA,hidden,B", [("ddqn", 8, 4, (64, 48), 96), ("ddqn", 115, 3, (400, 300), 256),
                                                ("sqn", 8, 4, (64, 48), 96), ("sqn", 24, 18, (400, 300), 512)])
def test_steps_match_oracle(QL, kind, D, nA, hidden, B):
    n_nets = 1 if kind == "ddqn" else 2
    params = init_q_params(D, nA, hidden, n_nets, seed=11)
    learner = QL[kind](opt_of(D, nA, hidden, B, alpha=0.2), "learner")
    assert learner.names == list(params) and [tuple(s) for s in learner.shapes.values()] == [v.shape for v in params.values()]
    learner.set_weights(list(params), list(params.values()))
    oracle = (DDQNOracle if kind == "ddqn" else SQNOracle)(params, alpha=0.2)
    for it in range(3):
        batch = make_q_batch(D, nA, B, seed=20 + it)
        want = oracle.step(batch)
        got = learner.train(batch, cnt=it, sync_outputs=True)
        loss = got["loss"].cpu().numpy()
        for k in range(n_nets + 1):
            assert abs(loss[k] - want["losses"][k]) <= 1e-5 * abs(want["losses"][k]), (it, k, loss[k], want["losses"][k])
        for k in range(n_nets):
            q = got["q"][k].cpu().numpy()
            assert np.abs(q - want["q"][k]).max() <= 1e-5 * max(1.0, np.abs(want["q"][k]).max()), (it, k)
        g_got = learner.get_flat_weights("grad").cpu().numpy().astype(np.float64)
        g_want = np.concatenate([v.reshape(-1) for v in want["grads"].values()])
        # first step: same weights on both sides; later steps also carry the (bounded, see below) divergence of the weights
        assert np.abs(g_got - g_want).max() <= (2e-5 if it == 0 else 1e-4) * np.abs(g_want).max(), it
        for which in ("main", "target"):
            w_got, w_want = learner.get_flat_weights(which).cpu().numpy().astype(np.float64), oracle.flat(which)
            err = np.abs(w_got - w_want)
            if it == 0:
                strong = np.abs(g_want) > 1e-4 * np.abs(g_want).max()
                assert err[strong].max() <= 1e-5 * np.abs(w_want).max(), (it, which)
                assert err.max() <= 5e-5 * np.abs(w_want).max(), (it, which)      # epsilon-dominated Adam entries
            else:       # the epsilon-dominated entries keep their first-step offset and add one per step
                assert err.max() <= 5e-5 * (it + 1) * np.abs(w_want).max(), (it, which)


def test_weight_surface_and_actor_side(QL):
    D, nA, hidden = 6, 5, (32, 32)
    learner = QL["sqn"](opt_of(D, nA, hidden, 64), "learner")
    keys, values = learner.get_weights()
    assert keys == learner.names and len(keys) == 12 and values[0].shape == (D, 32) and values[-1].shape == (nA,)
    new = [v + 0.25 for v in values[:2]]
    learner.set_weights(keys[:2], new)                                # subset assignment; target re-initialised from main
    k2, v2 = learner.get_weights()
    assert np.array_equal(v2[0], new[0]) and np.array_equal(v2[2], values[2])
    assert torch.equal(learner.get_flat_weights("main"), learner.get_flat_weights("target"))
    obs = np.random.Generator(np.random.PCG64(1)).standard_normal((7, D)).astype(np.float32)
    q = learner.q_values(obs, net=1).cpu().numpy()
    w = dict(zip(k2, v2))
    h = np.maximum(obs @ w["main/q2/dense/kernel"] + w["main/q2/dense/bias"], 0)
    h = np.maximum(h @ w["main/q2/dense_1/kernel"] + w["main/q2/dense_1/bias"], 0)
    assert np.allclose(q, h @ w["main/q2/dense_2/kernel"] + w["main/q2/dense_2/bias"], rtol=1e-5, atol=1e-5)
    assert 0 <= learner.get_action(obs[0], deterministic=True) < nA and 0 <= learner.get_action(obs[0]) < nA


def test_dqn_learner_trains_from_the_dqn_flavour_ring(QL):
    """The dqn-family loop (algos/dqn/train.py:146-160): batch = replay_buffer.sample_batch(B); agent.train(batch, cnt)."""
    from ddrl_b200 import ReplayBuffer
    D, nA, B = 10, 3, 128
    rb = ReplayBuffer(D, None, 2000, flavor="dqn", seed=3)
    g = np.random.Generator(np.random.PCG64(2))
    n = 1500
    rb.store_batch(g.standard_normal((n, D)), g.integers(0, nA, n), g.standard_normal(n), g.standard_normal((n, D)), g.random(n) < 0.05)
    learner = QL["ddqn"](opt_of(D, nA, (64, 64), B), "learner")
    first = None
    fixed = rb.sample_batch(B, device=True)
    for cnt in range(200):
        out = learner.train(fixed, cnt)
        if first is None:
            first = float(out["loss"][0])
    assert np.isfinite(float(out["loss"][0])) and float(out["loss"][0]) < first      # the loss on a fixed batch goes down
    out = learner.train(rb.sample_batch(B), 0, sync_outputs=True)                     # host dict, scalar actions
    assert out["q"].shape == (1, B, nA) and torch.isfinite(out["q"]).all()
