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


class ShardedReplayBuffer:
    """One replay ring per rank (GPU) of a node + the two sampling modes of SURVEY.md §8(e).

    local  : each learner samples its own shard — the reference's behaviour (one shard per call,
             algos/sac1/sac_ray.py:137-141), no communication at all;
    global : indices are uniform over ALL stored transitions of ALL shards; rows that live on another
             GPU are read by the gather kernel straight from the peer's ring over NVLink (CUDA IPC
             mapping, ddrl_rb_peer_attach), still without a data-path collective.
    `connect()` is the only collective (an all_gather of the 64-byte IPC handles); `refresh_sizes()`
    all-gathers the shard fill counts (call it after store phases; sizes are constant once full)."""

    def __init__(self, obs_dim, act_dim, total_size, group=None, device=None, **kw):
        import ctypes as C
        from .replay import ReplayBuffer
        self._C = C
        self.group = group
        self.world, self.rank = world_size(group), rank(group)
        if self.world > 8:
            raise ValueError("one node: at most 8 shards")
        self.map = ShardMap(total_size, self.world)
        kw.setdefault("rng_stream", self.rank)
        self.local = ReplayBuffer(obs_dim, act_dim, self.map.cap, device=device, **kw)
        self.sizes = [0] * self.world
        self._connected = False

    # local pass-throughs (the reference call surface)
    def store(self, *a):
        return self.local.store(*a)

    def store_batch(self, *a):
        return self.local.store_batch(*a)

    def get_counts(self):
        return self.local.get_counts()

    def connect(self):
        from . import _native as N
        C = self._C
        lib = N.lib()
        self.local.flush()
        buf = (C.c_ubyte * 64)()
        N.check(lib.ddrl_rb_ipc_export(self.local._h, buf))
        mine = bytes(buf)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=self.group)
        else:
            handles = [mine]
        for s, h in enumerate(handles):
            if s == self.rank:
                N.check(lib.ddrl_rb_peer_attach(self.local._h, self.world, s, None))
            else:
                N.check(lib.ddrl_rb_peer_attach(self.local._h, self.world, s, C.c_char_p(h)))
        self._connected = True
        self.refresh_sizes()

    def refresh_sizes(self):
        self.local.flush()
        torch.cuda.synchronize(self.local.device)     # our own stores are visible before peers learn the size
        mine = int(self.local.size)
        if self.world > 1:
            out = [None] * self.world
            dist.all_gather_object(out, mine, group=self.group)
            self.sizes = [int(x) for x in out]
        else:
            self.sizes = [mine]
        return self.sizes

    def sample_batch(self, batch_size=128, *, mode="local", idxs=None, device=True, return_idxs=False):
        if mode == "local":
            return self.local.sample_batch(batch_size, idxs=idxs, device=device, return_idxs=return_idxs)
        if mode != "global":
            raise ValueError(mode)
        if not self._connected:
            raise RuntimeError("call connect() on every rank before global sampling")
        from . import _native as N
        C = self._C
        rb = self.local
        rb.flush()
        n = int(batch_size)
        if sum(self.sizes) == 0:
            raise ValueError("high <= 0")
        dev = torch.device("cuda", rb.device)
        D, A = rb.obs_dim, rb.act_dim
        f32 = dict(dtype=torch.float32, device=dev)
        out = dict(obs1=torch.empty((n, D), **f32), obs2=torch.empty((n, D), **f32), acts=torch.empty((n, A), **f32),
                   rews=torch.empty(n, **f32), done=torch.empty(n, **f32))
        o_idx = torch.empty(n, dtype=torch.int64, device=dev) if return_idxs else None
        d_idx = None
        if idxs is not None:
            d_idx = torch.as_tensor(idxs, dtype=torch.int64).reshape(-1).to(dev).contiguous()
        sizes = (C.c_int64 * self.world)(*self.sizes)
        s = torch.cuda.current_stream(rb.device)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        seed = 0 if idxs is not None else rb._philox_seed()
        N.check(N.lib().ddrl_rb_sample_global(rb._h, n, 1, sizes, p(d_idx), seed, rb._counter, rb._rng_stream,
                                              p(out["obs1"]), p(out["obs2"]), p(out["acts"]), p(out["rews"]),
                                              p(out["done"]), p(o_idx), C.c_void_p(s.cuda_stream)))
        if idxs is None:
            rb._counter += 1
        if return_idxs:
            out["idxs"] = o_idx
        if not device:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out


class ShardedFrameReplayBuffer:
    """BASELINE.json config 4 as configured: an Atari-shaped uint8 frame replay of `total_size` transitions sharded over
    the GPUs of a node (1e6 transitions over 8 B200 = 125 000 per shard, 0.9 GB deduplicated / 7 GB naive per GPU).

    The path shards with NO data-path collective: every rank owns a FrameReplayBuffer of ceil(total / world) slots, its
    rollout producers store into it and its learner samples from it — the reference's one-shard-per-call behaviour
    (algos/sac1/sac_ray.py:137-141).  `get_counts(global_=True)` is the only collective (an all-reduce of three integers)."""

    def __init__(self, frame_shape=(84, 84), stack=4, total_size=1_000_000, *, mode="dedup", group=None, device=None,
                 seed=None):
        from .frames import FrameReplayBuffer
        self.group = group
        self.world, self.rank = world_size(group), rank(group)
        self.map = ShardMap(total_size, self.world)
        self.local = FrameReplayBuffer(frame_shape, stack, self.map.cap, mode=mode, device=device, seed=seed,
                                       rng_stream=self.rank)

    def store(self, *a):
        return self.local.store(*a)

    def store_batch(self, *a):
        return self.local.store_batch(*a)

    def store_frames(self, *a):
        return self.local.store_frames(*a)

    def sample_batch(self, batch_size=512, **kw):
        return self.local.sample_batch(batch_size, **kw)

    def get_counts(self, global_=False):
        c = self.local.get_counts()
        if not global_ or self.world == 1:
            return c
        t = torch.tensor(list(c), dtype=torch.int64, device=self.local._dev if dist.get_backend(self.group) == "nccl" else "cpu")
        dist.all_reduce(t, group=self.group)
        return tuple(int(x) for x in t.tolist())
