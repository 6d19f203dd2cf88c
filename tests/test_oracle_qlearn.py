"""CPU self-checks of oracle/qlearn_oracle.py (parity unpinned: TensorFlow 1.x is not installable): analytic gradients
against central finite differences in float64, the TF1 Adam recurrence on a scalar, the stop-gradient / polyak structure."""
import numpy as np
import pytest
import torch

from oracle.qlearn_oracle import DDQNOracle, SQNOracle, init_q_params, make_q_batch


@pytest.mark.parametrize("cls,n_nets", [(DDQNOracle, 1), (SQNOracle, 2)])
def test_gradients_match_finite_differences(cls, n_nets):
    D, nA, hidden, B = 5, 4, (7, 6), 16
    params = init_q_params(D, nA, hidden, n_nets, seed=1)
    batch = make_q_batch(D, nA, B, seed=2)
    ora = cls(params, alpha=0.3)
    b = {k: torch.tensor(v, dtype=torch.float64) for k, v in batch.items()}
    ls, _, backup = ora.losses(b)
    backup = backup.detach()                      # tf.stop_gradient: held fixed while a weight is perturbed
    grads = torch.autograd.grad(sum(ls), list(ora.main.values()))
    g = np.random.Generator(np.random.PCG64(3))
    for (name, w), gr in zip(ora.main.items(), grads):
        for _ in range(3):
            idx = tuple(int(g.integers(0, s)) for s in w.shape)
            eps = 1e-6
            with torch.no_grad():
                old = float(w[idx])
                w[idx] = old + eps
                lp = float(sum(ora.losses(b, backup)[0]))
                w[idx] = old - eps
                lm = float(sum(ora.losses(b, backup)[0]))
                w[idx] = old
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - float(gr[idx])) <= 1e-6 * max(1.0, abs(fd)), (name, idx, fd, float(gr[idx]))


def test_target_is_not_differentiated_and_polyak_uses_updated_weights():
    D, nA, hidden, B = 4, 3, (5, 5), 8
    params = init_q_params(D, nA, hidden, 1, seed=4)
    ora = DDQNOracle(params, lr=1e-2, polyak=0.9)
    before_t = ora.flat("target").copy()
    out = ora.step(make_q_batch(D, nA, B, seed=5))
    after_m, after_t = ora.flat("main"), ora.flat("target")
    assert np.allclose(after_t, 0.9 * before_t + 0.1 * after_m, rtol=0, atol=1e-15)     # polyak of the UPDATED main
    # first TF1-Adam step in closed form: m = 0.1 g, v = 0.001 g^2, lr_t = lr sqrt(0.001) / 0.1
    #   |dw| = lr |g| / (|g| + eps / sqrt(0.001))
    g = np.concatenate([v.reshape(-1) for v in out["grads"].values()])
    moved = np.abs(after_m - np.concatenate([np.asarray(params[k], np.float64).reshape(-1) for k in ora.names]))
    assert np.allclose(moved, 1e-2 * np.abs(g) / (np.abs(g) + 1e-8 / np.sqrt(0.001)), rtol=1e-9, atol=1e-18)


def test_sqn_entropy_term_sign_and_value():
    """core.py:42 `logp_pi = sum(exp(pi_log) * pi_log)` is MINUS the entropy; the backup subtracts alpha times it, i.e. adds
    alpha * H.  With a uniform softmax (equal Q values) H = log(nA)."""
    D, nA, hidden = 3, 5, (4, 4)
    params = init_q_params(D, nA, hidden, 2, seed=6)
    for k in params:                                   # zero networks: q == 0 everywhere
        params[k] = np.zeros_like(params[k])
    ora = SQNOracle(params, alpha=0.2, gamma=0.5)
    batch = make_q_batch(D, nA, 6, seed=7)
    batch["done"][:] = 0.0
    b = {k: torch.tensor(v, dtype=torch.float64) for k, v in batch.items()}
    losses, _, _ = ora.losses(b)
    backup = batch["rews"] + 0.5 * (0.0 + 0.2 * np.log(nA))
    assert np.isclose(float(losses[0]), 0.5 * np.mean(backup.astype(np.float64) ** 2), rtol=1e-12)
