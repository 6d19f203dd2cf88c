"""Multi-GPU plumbing for the learner path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box; gloo in CPU tests).

What the reference does: several replay-buffer actors with a random shard per
call (algos/sac1/sac_ray.py:246, :137-141), a ParameterServer that learners overwrite by pickle +
RPC (algos/sac1/sac1.py:86-92, :149-151), and no gradient exchange at all (actor_learner.py:144-148
are empty stubs).  Here:
  * replay is sharded one ring per rank (ShardMap); "local" sampling = the reference's behaviour,
    "global" sampling draws (shard, row) uniformly and reads remote rows over NVLink P2P;
  * learners are synchronous data-parallel replicas: the flat gradient buffer is all-reduced
    (Learner.train does it) so that an N-rank step equals one step on the concatenated batch;
  * the ParameterServer push/pull across ranks is ONE broadcast of the flat weight buffer.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world_size(group=None):
    return dist.get_world_size(group) if is_dist() else 1


def rank(group=None):
    return dist.get_rank(group) if is_dist() else 0


def allreduce_mean_(t, group=None):
    """In-place mean over ranks (SUM then scale: identical on NCCL and gloo)."""
    if world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(world_size(group))
    return t


def broadcast_flat_(t, src=0, group=None):
    if world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


class ShardMap:
    """Global transition index <-> (shard, row) for a replay of `total_rows` split evenly over
    `n_shards` rings (capacity per shard = ceil(total / n)); shard s owns global rows
    [s*cap, (s+1)*cap)."""

    def __init__(self, total_rows, n_shards):
        self.n = int(n_shards)
        self.cap = -(-int(total_rows) // self.n)
        self.total = self.cap * self.n

    def locate(self, gidx):
        gidx = np.asarray(gidx, dtype=np.int64)
        return gidx // self.cap, gidx % self.cap

    def global_index(self, shard, row):
        return np.asarray(shard, dtype=np.int64) * self.cap + np.asarray(row, dtype=np.int64)

    def pick_shard(self, rng=np.random):
        """The reference's policy: one uniformly random shard per store / sample call
        (algos/sac1/sac_ray.py:246 `np.random.choice(opt.num_buffers, 1)`, :137-141)."""
        return int(rng.randint(0, self.n))


class DistributedParameterServer:
    """ParameterServer whose push/pull crosses ranks as one broadcast of a flat float32 buffer.

    Same call surface as the reference's (algos/sac1/sac1.py:66-100): push(keys, values) on the
    source rank stages new values; sync() (collective: every rank calls it at the same step, e.g.
    every 300 learner steps like sac1.py:149) broadcasts; pull(keys) returns the local replica.
    `device` is where the flat buffer lives: the learner's GPU (NCCL) or "cpu" (gloo tests)."""

    def __init__(self, keys, values, src=0, device="cpu", group=None):
        self.keys = list(keys)
        self.shapes = OrderedDict((k, tuple(np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v).shape))
                                  for k, v in zip(keys, values))
        self.offsets, o = OrderedDict(), 0
        for k, s in self.shapes.items():
            n = int(np.prod(s)) if len(s) else 1
            self.offsets[k] = (o, n)
            o += n
        self.flat = torch.zeros(o, dtype=torch.float32, device=device)
        self.src, self.group = src, group
        self.version = 0
        self.push(keys, values)
        self.sync()

    def push(self, keys, values):
        for k, v in zip(keys, values):
            o, n = self.offsets[k]
            t = v.detach() if isinstance(v, torch.Tensor) else torch.from_numpy(np.array(v, dtype=np.float32, copy=True))
            self.flat[o:o + n].copy_(t.reshape(-1).to(self.flat.device, torch.float32))

    def push_flat(self, flat):
        """Device-to-device push of a learner's flat weight vector (Learner.get_flat_weights())."""
        self.flat.copy_(flat.to(self.flat.device, torch.float32))

    def sync(self):
        broadcast_flat_(self.flat, self.src, self.group)
        self.version += 1

    def pull(self, keys):
        host = self.flat.detach().cpu().numpy()
        return [host[self.offsets[k][0]:self.offsets[k][0] + self.offsets[k][1]].reshape(self.shapes[k]).copy()
                for k in keys]

    def pull_flat(self):
        return self.flat

    def get_weights(self):
        return OrderedDict(zip(self.keys, self.pull(self.keys)))

    def save_weights(self, name):
        import pickle
        with open(name + "weights.pickle", "wb") as f:
            pickle.dump(dict(self.get_weights()), f)
