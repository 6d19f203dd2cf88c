"""The discrete-action learners of the reference's dqn / sqn families on the GPU.

    DQNLearner(opt, job)   algos/dqn/actor_learner.py:19-130  (double DQN: online argmax, target-network value)
    SQNLearner(opt, job)   algos/sqn/actor_learner.py:19-125  (two soft Q networks, min-double-Q + entropy backup)
with the reference's call surface: .train(batch, cnt) on the dict the dqn-family ReplayBuffer returns (obs1, obs2, acts as
action indices, rews, done), .get_weights() -> (keys, values) of the main variables in TF creation order,
.set_weights(keys, values) (also re-initialises the target networks, like the reference's `target_init`).
opt: obs_dim, act_dim (number of actions), hidden_size (two hidden layers, reference default [400, 300]), gamma, lr,
polyak, seed (+ alpha for SQN).  Actors: .q_values(obs) / .get_action(o, deterministic) on the main network.

All arithmetic runs in libddrl_b200 (csrc/qlearn.cu): grouped fp32 GEMMs + one loss kernel + one Adam/polyak pass; there
is no CPU path."""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _native as N

_KEYS_BATCH = ("obs1", "obs2", "acts", "rews", "done")


def q_param_names(n_nets):
    """TF1 variable names in creation order: tf.layers.dense default naming inside main/q1 (and main/q2)."""
    names = []
    for q in ("q1", "q2")[:n_nets]:
        for suffix in ("dense", "dense_1", "dense_2"):
            names += [f"main/{q}/{suffix}/kernel", f"main/{q}/{suffix}/bias"]
    return names


def q_param_shapes(obs_dim, n_actions, hidden, n_nets):
    h1, h2 = hidden
    one = [(obs_dim, h1), (h1,), (h1, h2), (h2,), (h2, n_actions), (n_actions,)]
    return OrderedDict(zip(q_param_names(n_nets), one * n_nets))


class _QLearner(object):
    N_NETS = 1

    def __init__(self, opt, job="learner", *, device=None, max_batch=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ddrl_b200 Q learners need a CUDA device (no CPU fallback)")
        self.opt = opt
        self._lib = N.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._dev = torch.device("cuda", self.device)
        hidden = tuple(getattr(opt, "hidden_size", (400, 300)))
        if len(hidden) != 2:
            raise ValueError("two hidden layers (the reference's default [400, 300] shape) are supported")
        self.obs_dim, self.n_actions, self.hidden = int(opt.obs_dim), int(opt.act_dim), (int(hidden[0]), int(hidden[1]))
        self.max_batch = int(max_batch or getattr(opt, "batch_size", 256) or 256)
        self.shapes = q_param_shapes(self.obs_dim, self.n_actions, self.hidden, self.N_NETS)
        self.names = list(self.shapes)
        h = C.c_void_p()
        N.check(self._lib.ddrl_ql_create(self.device, self.obs_dim, self.n_actions, self.hidden[0], self.hidden[1], self.max_batch,
                                         self.N_NETS, float(getattr(opt, "gamma", 0.99)), float(getattr(opt, "polyak", 0.995)),
                                         float(getattr(opt, "lr", 1e-3)), float(getattr(opt, "alpha", 0.1)), C.byref(h)))
        self._h = h
        self.P = int(self._lib.ddrl_ql_param_count(self._h))
        assert self.P == sum(int(np.prod(s)) for s in self.shapes.values())
        # tf.layers.dense defaults: glorot_uniform kernels, zero biases
        g = np.random.Generator(np.random.PCG64(int(getattr(opt, "seed", 0))))
        init = []
        for s in self.shapes.values():
            if len(s) == 2:
                lim = math.sqrt(6.0 / (s[0] + s[1]))
                init.append(g.uniform(-lim, lim, s).astype(np.float32))
            else:
                init.append(np.zeros(s, np.float32))
        self.set_weights(self.names, init)
        self._loss = torch.zeros(self.N_NETS + 1, dtype=torch.float32, device=self._dev)
        self._q = None
        self.steps = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ddrl_ql_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device)

    # ---- weights (reference: ray TensorFlowVariables keyed by variable name) ----------------------------------------
    def get_flat_weights(self, which="main"):
        out = torch.empty(self.P, dtype=torch.float32, device=self._dev)
        idx = {"main": 0, "target": 1, "adam_m": 2, "adam_v": 3, "grad": 4}[which]
        N.check(self._lib.ddrl_ql_get_weights(self._h, out.data_ptr(), idx, self._stream().cuda_stream))
        return out

    def get_weights(self):
        flat = self.get_flat_weights("main").cpu().numpy()
        vals, o = [], 0
        for s in self.shapes.values():
            n = int(np.prod(s))
            vals.append(flat[o:o + n].reshape(s).copy())
            o += n
        return list(self.names), vals

    def set_weights(self, variable_names, weights):
        """Assign by name (a subset is allowed, like TensorFlowVariables.set_weights), then target <- main."""
        cur = dict(zip(*self.get_weights())) if set(variable_names) != set(self.names) else {}
        cur.update({k: np.asarray(v, dtype=np.float32) for k, v in zip(variable_names, weights)})
        unknown = set(cur) - set(self.names)
        if unknown:
            raise KeyError(f"unknown variables {sorted(unknown)}")
        flat = np.concatenate([cur[k].reshape(-1) for k in self.names]).astype(np.float32)
        for k in self.names:
            if tuple(cur[k].shape) != tuple(self.shapes[k]):
                raise ValueError(f"{k}: shape {cur[k].shape}, expected {self.shapes[k]}")
        t = torch.from_numpy(flat).to(self._dev)
        s = self._stream()
        N.check(self._lib.ddrl_ql_set_weights(self._h, t.data_ptr(), 1, s.cuda_stream))
        t.record_stream(s)

    # ---- train ----------------------------------------------------------------------------------------------------
    def train(self, batch, cnt=0, sync_outputs=False):
        """One update (reference: sess.run(step_ops, feed_dict)).  Returns dict(loss=[per-network losses..., their sum]
        (device tensor), q=[n_nets, B, n_actions] the main networks' Q values of obs1) without synchronising."""
        ts = []
        for k, width in zip(_KEYS_BATCH, (self.obs_dim, self.obs_dim, 1, 1, 1)):
            v = batch[k]
            if not (isinstance(v, torch.Tensor) and v.is_cuda):
                v = torch.as_tensor(np.ascontiguousarray(np.asarray(v), dtype=np.float32)).to(self._dev)
            ts.append(v.to(torch.float32).reshape(-1, width).contiguous())
        B = int(ts[3].shape[0])
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        if self._q is None or self._q.shape[1] != B:
            self._q = torch.empty((self.N_NETS, B, self.n_actions), dtype=torch.float32, device=self._dev)
        s = self._stream()
        N.check(self._lib.ddrl_ql_step(self._h, *[t.data_ptr() for t in ts], B, self._loss.data_ptr(), self._q.data_ptr(),
                                       s.cuda_stream))
        for t in ts:
            t.record_stream(s)
        self.steps += 1
        if sync_outputs:
            s.synchronize()
        return dict(loss=self._loss, q=self._q)

    # ---- act ------------------------------------------------------------------------------------------------------
    def q_values(self, obs, net=0):
        o = torch.as_tensor(np.asarray(obs, dtype=np.float32)).to(self._dev) if not isinstance(obs, torch.Tensor) else obs.to(self._dev, torch.float32)
        o = o.reshape(-1, self.obs_dim).contiguous()
        out = torch.empty((o.shape[0], self.n_actions), dtype=torch.float32, device=self._dev)
        s = self._stream()
        N.check(self._lib.ddrl_ql_forward(self._h, o.data_ptr(), int(o.shape[0]), int(net), out.data_ptr(), s.cuda_stream))
        o.record_stream(s)
        return out


class DQNLearner(_QLearner):
    """algos/dqn/actor_learner.py Learner (double DQN)."""
    N_NETS = 1

    def get_action(self, o, deterministic=True, epsilon=0.0):
        if not deterministic and np.random.rand() < epsilon:
            return int(np.random.randint(self.n_actions))
        return int(self.q_values(np.asarray(o).reshape(1, -1)).argmax(dim=1)[0])


class SQNLearner(_QLearner):
    """algos/sqn/actor_learner.py Learner (soft Q network: softmax(q1 / alpha) policy)."""
    N_NETS = 2

    def get_action(self, o, deterministic=False):
        q = self.q_values(np.asarray(o).reshape(1, -1))[0]
        if deterministic:
            return int(q.argmax())                                  # mu = argmax pi_log (core.py:33)
        p = torch.softmax(q / float(getattr(self.opt, "alpha", 0.1)), dim=0)
        return int(torch.multinomial(p, 1)[0])                      # pi ~ multinomial(pi_log) (core.py:38)
