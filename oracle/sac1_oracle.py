"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — torch-CPU restatement of one SAC1 learner
update, float64 by default (ground truth) or float32 (the CPU baseline that bench.py times).

PARITY UNPINNED BY THE REFERENCE: the arithmetic of this path lives in TensorFlow 1.x graph mode
(tf.layers.dense, tf.train.AdamOptimizer, tf.random_normal ...), a third-party dependency that the
reference neither vendors nor pins (no requirements file; era Aug-2019 => TF 1.13/1.14) and that
cannot be installed here.  The reference has no tests or golden vectors for it.  This file follows
the reference's own graph-construction code line by line and TF1's published op definitions:

  dense / mlp                         algos/sac1/core.py:15-18    (tf.layers.dense: x @ kernel + bias)
  gaussian_likelihood                 algos/sac1/core.py:30-32
  clip_but_pass_gradient              algos/sac1/core.py:35-38
  mlp_gaussian_policy                 algos/sac1/core.py:49-79    (LOG_STD_MIN/MAX :45-46)
  apply_squashing_func                algos/sac1/core.py:82-87
  mlp_actor_critic (x / x2 wiring)    algos/sac1/core.py:91-121
  main / target graphs                algos/sac1/actor_learner.py:27-38
  entropy alpha ('auto')              algos/sac1/actor_learner.py:46-55   (reference-intended)
  targets and losses                  algos/sac1/actor_learner.py:58-69
  optimisers and their order          algos/sac1/actor_learner.py:73-81
  polyak after both Adam steps        algos/sac1/actor_learner.py:85-87
  fetches (pre-update values)         algos/sac1/actor_learner.py:96-101
  set_weights -> target_init          algos/sac1/actor_learner.py:104-105,125-127
  get_weights ("main" keys only)      algos/sac1/actor_learner.py:129-133
  TF1 AdamOptimizer                   tensorflow/python/training/adam.py (1.x): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
                                      var -= lr_t * m / (sqrt(v) + eps)   ("epsilon hat" form)

Self-checks that do not need TensorFlow live in tests/test_oracle_sac1.py: finite-difference
gradients of both losses, the Adam recurrence against a scalar hand computation, and structural
properties (Q weights untouched by the pi-loss, polyak over the policy too, pre-update fetches).
The three tf.random_normal draws of a step cannot be reproduced; oracle and kernel both take the
noise as an input ([3, B, A]: eps1 for pi(x), eps2 for pi(x2), eps3 for the target policy on x2).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

LOG_STD_MAX = 2.0
LOG_STD_MIN = -20.0
EPS = 1e-8


def param_names():
    names = []
    for i, suffix in enumerate(("dense", "dense_1", "dense_2", "dense_3")):
        names += [f"main/pi/{suffix}/kernel", f"main/pi/{suffix}/bias"]
    for q in ("q1", "q2"):
        for suffix in ("dense", "dense_1", "dense_2"):
            names += [f"main/{q}/{suffix}/kernel", f"main/{q}/{suffix}/bias"]
    return names


def param_shapes(obs_dim, act_dim, hidden):
    h1, h2 = hidden
    D, A = obs_dim, act_dim
    shapes = OrderedDict()
    pi = [(D, h1), (h1,), (h1, h2), (h2,), (h2, A), (A,), (h2, A), (A,)]
    q = [(D + A, h1), (h1,), (h1, h2), (h2,), (h2, 1), (1,)]
    for n, s in zip(param_names(), pi + q + q):
        shapes[n] = s
    return shapes


def init_params(obs_dim, act_dim, hidden, seed, dtype=np.float32):
    """Glorot-uniform kernels / zero biases (TF1 tf.layers.dense defaults) from a numpy PCG64 stream.
    Returns OrderedDict name -> ndarray in canonical order (pi, q1, q2)."""
    g = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    for n, s in param_shapes(obs_dim, act_dim, hidden).items():
        if len(s) == 2:
            lim = math.sqrt(6.0 / (s[0] + s[1]))
            out[n] = g.uniform(-lim, lim, s).astype(dtype)
        else:
            out[n] = np.zeros(s, dtype=dtype)
    return out


def conditioned_params(obs_dim, act_dim, hidden, seed, dtype=np.float32):
    """init_params with a log_std head typical of a trained policy (log_std ~ -2 +- 0.5 instead of
    the Glorot-init range [-17, -1]).  Why parity tests use it: the reference graph evaluates
    `(pi - mu) / (exp(log_std) + 1e-8)` with pi = mu + eps*std (algos/sac1/core.py:31,76-78) and
    `log(1 - tanh(u)**2 + 1e-6)` (core.py:86) in float32; when std << ulp(mu) or |u| >~ 5 these
    cancel catastrophically and the float32 result is dominated by rounding (it depends on the exact
    tanh/exp implementation of the TF build), so no implementation — the reference on another GPU
    included — reproduces it to 1e-5.  In the conditioned regime float32 and float64 agree to ~1e-7."""
    p = init_params(obs_dim, act_dim, hidden, seed, dtype)
    p["main/pi/dense_3/kernel"] = (p["main/pi/dense_3/kernel"] * 0.1).astype(dtype)
    p["main/pi/dense_3/bias"][:] = 0.75
    g = np.random.Generator(np.random.PCG64(seed + 1))
    for n in p:
        if n.endswith("bias") and "dense_3" not in n:
            p[n] = g.uniform(-0.05, 0.05, p[n].shape).astype(dtype)
    return p


class TF1Adam:
    """tf.train.AdamOptimizer(learning_rate) with TF1 defaults beta1=0.9 beta2=0.999 epsilon=1e-8."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.t = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, self.m, self.v):
                m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                p.sub_(lr_t * m / (v.sqrt() + self.eps))


def clip_but_pass_gradient(x, l=-1.0, u=1.0):
    clip_up = (x > u).to(x.dtype)
    clip_low = (x < l).to(x.dtype)
    return x + ((u - x) * clip_up + (l - x) * clip_low).detach()


def gaussian_likelihood(x, mu, log_std):
    pre_sum = -0.5 * (((x - mu) / (torch.exp(log_std) + EPS)) ** 2 + 2 * log_std + math.log(2 * math.pi))
    return pre_sum.sum(dim=1)


class SAC1Oracle:
    def __init__(self, obs_dim, act_dim, hidden=(400, 300), gamma=0.99, polyak=0.995, lr=1e-3, alpha=0.2,
                 act_scale=1.0, dtype=torch.float64, params=None, seed=0):
        self.D, self.A, self.hidden = obs_dim, act_dim, tuple(hidden)
        self.gamma, self.polyak, self.lr, self.act_scale = gamma, polyak, lr, act_scale
        self.dtype = dtype
        self.names = param_names()
        if params is None:
            params = init_params(obs_dim, act_dim, hidden, seed)
        self.main = OrderedDict((n, torch.tensor(np.asarray(params[n]), dtype=dtype).clone().requires_grad_(True))
                                for n in self.names)
        self.target = OrderedDict((n, self.main[n].detach().clone()) for n in self.names)
        self.pi_names = [n for n in self.names if "/pi/" in n]
        self.q_names = [n for n in self.names if "/q1/" in n or "/q2/" in n]
        self.opt_pi = TF1Adam([self.main[n] for n in self.pi_names], lr)
        self.opt_q = TF1Adam([self.main[n] for n in self.q_names], lr)
        self.auto_alpha = alpha == "auto"
        if self.auto_alpha:
            self.log_alpha = torch.zeros((), dtype=dtype, requires_grad=True)
            self.target_entropy = -float(act_dim)
            self.opt_alpha = TF1Adam([self.log_alpha], lr)
        else:
            self.alpha = float(alpha)

    # ---- weights contract --------------------------------------------------------------------
    def get_weights(self):
        keys = list(self.names)
        return keys, [self.main[k].detach().numpy().copy() for k in keys]

    def set_weights(self, keys, values):
        with torch.no_grad():
            for k, v in zip(keys, values):
                self.main[k].copy_(torch.as_tensor(np.asarray(v), dtype=self.dtype))
            for k in self.names:                      # target_init over all 20 tensors
                self.target[k].copy_(self.main[k])

    # ---- networks ----------------------------------------------------------------------------
    @staticmethod
    def _dense(x, W, prefix, layer):
        return x @ W[f"{prefix}/{layer}/kernel"] + W[f"{prefix}/{layer}/bias"]

    def policy(self, W, x, eps, scope="main"):
        """returns (mu, pi, logp_pi) after squashing, BEFORE action scaling."""
        p = f"main/pi"                     # names are stored under 'main/...' for both scopes
        net = torch.relu(self._dense(x, W, p, "dense"))
        net = torch.relu(self._dense(net, W, p, "dense_1"))
        mu = self._dense(net, W, p, "dense_2")
        log_std = torch.tanh(self._dense(net, W, p, "dense_3"))
        log_std = LOG_STD_MIN + 0.5 * (LOG_STD_MAX - LOG_STD_MIN) * (log_std + 1)
        std = torch.exp(log_std)
        pi = mu + eps * std
        logp_pi = gaussian_likelihood(pi, mu, log_std)
        mu = torch.tanh(mu)
        pi = torch.tanh(pi)
        logp_pi = logp_pi - torch.log(clip_but_pass_gradient(1 - pi ** 2, l=0.0, u=1.0) + 1e-6).sum(dim=1)
        return mu, pi, logp_pi

    def qf(self, W, q, x, a):
        p = f"main/{q}"
        h = torch.relu(self._dense(torch.cat([x, a], dim=-1), W, p, "dense"))
        h = torch.relu(self._dense(h, W, p, "dense_1"))
        return self._dense(h, W, p, "dense_2").squeeze(1)

    # ---- forward graph (algos/sac1/actor_learner.py:27-69) -------------------------------------
    def forward(self, batch, noise):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=self.dtype)
        x, x2, a, r, d = (t(batch[k]) for k in ("obs1", "obs2", "acts", "rews", "done"))
        e1, e2, e3 = (t(noise[i]) for i in range(3))
        W, Wt = self.main, self.target
        alpha = torch.exp(self.log_alpha) if self.auto_alpha else self.alpha
        s = self.act_scale
        _, pi1, logp1 = self.policy(W, x, e1)
        _, _, logp2 = self.policy(W, x2, e2)
        a1 = pi1 * s
        q1, q2 = self.qf(W, "q1", x, a), self.qf(W, "q2", x, a)
        q1_pi = self.qf(W, "q1", x, a1)
        with torch.no_grad():
            _, pi3, _ = self.policy(Wt, x2, e3)
            a3 = pi3 * s
            q1_pi_t, q2_pi_t = self.qf(Wt, "q1", x2, a3), self.qf(Wt, "q2", x2, a3)
        min_q = torch.minimum(q1_pi_t, q2_pi_t)
        alpha_c = alpha.detach() if self.auto_alpha else alpha
        v_backup = (min_q - alpha_c * logp2).detach()
        q_backup = r + self.gamma * (1 - d) * v_backup
        pi_loss = (alpha_c * logp1 - q1_pi).mean()
        q1_loss = 0.5 * ((q_backup - q1) ** 2).mean()
        q2_loss = 0.5 * ((q_backup - q2) ** 2).mean()
        out = dict(pi_loss=pi_loss, q1_loss=q1_loss, q2_loss=q2_loss, q1=q1, q2=q2, logp_pi=logp1,
                   value_loss=q1_loss + q2_loss, q_backup=q_backup, logp_pi2=logp2, q1_pi=q1_pi, a1=a1)
        if self.auto_alpha:
            out["alpha_loss"] = (-self.log_alpha * (logp1 + self.target_entropy).detach()).mean()
            out["alpha"] = alpha.detach()
        else:
            out["alpha"] = torch.tensor(self.alpha, dtype=self.dtype)
        return out

    def gradients(self, batch, noise):
        f = self.forward(batch, noise)
        pi_params = [self.main[n] for n in self.pi_names]
        q_params = [self.main[n] for n in self.q_names]
        g_pi = torch.autograd.grad(f["pi_loss"], pi_params, retain_graph=True)
        g_q = torch.autograd.grad(f["value_loss"], q_params, retain_graph=self.auto_alpha)
        g_alpha = torch.autograd.grad(f["alpha_loss"], [self.log_alpha])[0] if self.auto_alpha else None
        return f, g_pi, g_q, g_alpha

    # ---- one update (algos/sac1/actor_learner.py:73-101,135-142) -------------------------------
    def step(self, batch, noise, grad_transform=None):
        """grad_transform(g_pi, g_q) -> (g_pi, g_q): hook used by the multi-GPU test to average
        gradients across simulated ranks."""
        f, g_pi, g_q, g_alpha = self.gradients(batch, noise)
        if grad_transform is not None:
            g_pi, g_q = grad_transform(g_pi, g_q)
        self.opt_pi.step([self.main[n] for n in self.pi_names], g_pi)
        self.opt_q.step([self.main[n] for n in self.q_names], g_q)
        with torch.no_grad():
            for n in self.names:        # polyak over all 20 tensors, policy included, post-Adam weights
                self.target[n].mul_(self.polyak).add_(self.main[n], alpha=1 - self.polyak)
        if self.auto_alpha:
            self.opt_alpha.step([self.log_alpha], [g_alpha])
        keys = ("pi_loss", "q1_loss", "q2_loss", "q1", "q2", "logp_pi", "alpha")
        return {k: f[k].detach().numpy().copy() for k in keys}

    def flat_grads(self, batch, noise):
        _, g_pi, g_q, _ = self.gradients(batch, noise)
        return torch.cat([g.reshape(-1) for g in list(g_pi) + list(g_q)]).detach().numpy()

    def flat(self, which="main"):
        W = self.main if which == "main" else self.target
        return torch.cat([W[n].detach().reshape(-1) for n in self.names]).numpy().copy()


def make_batch(obs_dim, act_dim, batch, seed, dtype=np.float32):
    """Synthetic transitions with the distributions of SURVEY.md §8(d)."""
    g = np.random.Generator(np.random.PCG64(seed))
    return dict(obs1=g.standard_normal((batch, obs_dim)).astype(dtype),
                obs2=g.standard_normal((batch, obs_dim)).astype(dtype),
                acts=g.uniform(-1, 1, (batch, act_dim)).astype(dtype),
                rews=g.standard_normal(batch).astype(dtype),
                done=(g.random(batch) < 0.01).astype(dtype)), \
        g.standard_normal((3, batch, act_dim)).astype(dtype)
