"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of one update of the reference's discrete-action learners.

    DDQNOracle  algos/dqn/actor_learner.py:27-67 + algos/dqn/core.py:15-63
    SQNOracle   algos/sqn/actor_learner.py:27-73 + algos/sqn/core.py:30-79
Graph: q = mlp(x, hidden + [n_actions], relu, None) per network (tf.layers.dense: x @ kernel + bias); the losses of the
cited lines; tf.train.AdamOptimizer (TF1: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), epsilon outside the square root of v) on
the main variables, then v_targ <- polyak v_targ + (1 - polyak) v_main with the UPDATED main variables (control dependency).

PARITY UNPINNED: the arithmetic lives in TensorFlow 1.x (not installable here) and the reference holds no tests or golden
vectors for it; the oracle is checked by finite differences (tests/test_oracle_qlearn.py) and read against the cited lines."""
from collections import OrderedDict

import numpy as np
import torch


def q_names(n_nets):
    out = []
    for q in ("q1", "q2")[:n_nets]:
        for s in ("dense", "dense_1", "dense_2"):
            out += [f"main/{q}/{s}/kernel", f"main/{q}/{s}/bias"]
    return out


def init_q_params(obs_dim, n_actions, hidden, n_nets, seed, bias_scale=0.1):
    g = np.random.Generator(np.random.PCG64(seed))
    h1, h2 = hidden
    shapes = [(obs_dim, h1), (h1,), (h1, h2), (h2,), (h2, n_actions), (n_actions,)] * n_nets
    out = OrderedDict()
    for n, s in zip(q_names(n_nets), shapes):
        lim = np.sqrt(6.0 / (s[0] + s[1])) if len(s) == 2 else bias_scale
        out[n] = g.uniform(-lim, lim, s).astype(np.float32)
    return out


def make_q_batch(obs_dim, n_actions, B, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    return dict(obs1=g.standard_normal((B, obs_dim)).astype(np.float32), obs2=g.standard_normal((B, obs_dim)).astype(np.float32),
                acts=g.integers(0, n_actions, B).astype(np.float32), rews=g.standard_normal(B).astype(np.float32),
                done=(g.random(B) < 0.1).astype(np.float32))


class _QOracle:
    n_nets = 1

    def __init__(self, params, gamma=0.99, polyak=0.995, lr=1e-3, alpha=0.1, dtype=torch.float64):
        self.dtype, self.gamma, self.polyak, self.lr, self.alpha = dtype, gamma, polyak, lr, alpha
        self.names = list(params)
        self.main = OrderedDict((k, torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True)) for k, v in params.items())
        self.target = OrderedDict((k, v.detach().clone()) for k, v in self.main.items())      # target_init
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.main.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.main.items())
        self.t = 0

    def mlp(self, w, net, x):          # core.mlp: relu hidden layers, linear output
        p = f"main/q{net + 1}/"
        h = torch.relu(x @ w[p + "dense/kernel"] + w[p + "dense/bias"])
        h = torch.relu(h @ w[p + "dense_1/kernel"] + w[p + "dense_1/bias"])
        return h @ w[p + "dense_2/kernel"] + w[p + "dense_2/bias"]

    def losses(self, batch, backup=None):
        """-> ([loss per network], [q(x) per network], backup).  `backup` given: use it instead of recomputing it (the
        reference wraps it in tf.stop_gradient; finite-difference checks must hold it fixed the same way)."""
        raise NotImplementedError

    def step(self, batch):
        b = {k: torch.tensor(np.asarray(v), dtype=self.dtype) for k, v in batch.items()}
        losses, qs, _ = self.losses(b)
        total = sum(losses)
        grads = torch.autograd.grad(total, list(self.main.values()))
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        lr_t = self.lr * np.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for (k, w), g in zip(self.main.items(), grads):
                self.m[k] = b1 * self.m[k] + (1 - b1) * g
                self.v[k] = b2 * self.v[k] + (1 - b2) * g * g
                w -= lr_t * self.m[k] / (torch.sqrt(self.v[k]) + eps)
                self.target[k] = self.polyak * self.target[k] + (1 - self.polyak) * w
        return dict(losses=[float(l) for l in losses] + [float(total)], q=[q.detach().numpy() for q in qs],
                    grads=OrderedDict((k, g.numpy()) for k, g in zip(self.main, grads)))

    def flat(self, which="main"):
        src = self.main if which == "main" else self.target
        return np.concatenate([src[k].detach().numpy().reshape(-1) for k in self.names])


class DDQNOracle(_QOracle):
    n_nets = 1

    def losses(self, b, backup=None):
        q = self.mlp(self.main, 0, b["obs1"])                                   # actor_learner.py:31
        q_x2 = self.mlp(self.main, 0, b["obs2"])
        q_next = self.mlp(self.target, 0, b["obs2"])                            # :35
        a = b["acts"].to(torch.int64)
        q_value = q.gather(1, a[:, None])[:, 0]                                 # :41-42
        best = q_x2.argmax(dim=1)                                               # :45 online argmax
        q_target = q_next.gather(1, best[:, None])[:, 0]                        # :46 target value
        q_backup = (b["rews"] + self.gamma * (1 - b["done"]) * q_target).detach() if backup is None else backup     # :52
        return [0.5 * ((q_backup - q_value) ** 2).mean()], [q], q_backup        # :55


class SQNOracle(_QOracle):
    n_nets = 2

    def losses(self, b, backup=None):
        q1, q2 = self.mlp(self.main, 0, b["obs1"]), self.mlp(self.main, 1, b["obs1"])     # core.py:61-75
        q1_x2 = self.mlp(self.main, 0, b["obs2"])                               # core.py:67
        pi_log = torch.log_softmax(q1_x2 / self.alpha, dim=1)                   # core.py:32
        entropy_x2 = (pi_log.exp() * pi_log).sum(dim=1)                         # core.py:42 ("exact entropy", negative)
        q1_mu_ = self.mlp(self.target, 0, b["obs2"]).max(dim=1).values          # q at its own argmax (core.py:64-65)
        q2_mu_ = self.mlp(self.target, 1, b["obs2"]).max(dim=1).values
        a = b["acts"].to(torch.int64)
        q1_a, q2_a = q1.gather(1, a[:, None])[:, 0], q2.gather(1, a[:, None])[:, 0]       # actor_learner.py:43-45
        v_backup = (torch.minimum(q1_mu_, q2_mu_) - self.alpha * entropy_x2).detach()     # :48-51
        q_backup = b["rews"] + self.gamma * (1 - b["done"]) * v_backup if backup is None else backup        # :52
        return [0.5 * ((q_backup - q1_a) ** 2).mean(), 0.5 * ((q_backup - q2_a) ** 2).mean()], [q1, q2], q_backup   # :56-58
