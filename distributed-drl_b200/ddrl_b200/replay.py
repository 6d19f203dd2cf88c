"""ReplayBuffer — the reference's replay API over an HBM-resident ring.

Mirrors (paths relative to the reference tree):
    ReplayBuffer(obs_dim, act_dim, size)            example/dsac.py:20-27, algos/sac1/sac1.py:34-41
    .store(obs, act, rew, next_obs, done)           example/dsac.py:29-37
    .sample_batch(batch_size)                       example/dsac.py:39-45  (default 128: sac1.py:53)
    .get_counts()                                   algos/sac1/sac1.py:62-63 / example/dsac.py:47-48
    .save() / .load()  (.npy + buffer_infos)        algos/dqn/train.py:82-108

All data lives in GPU memory (libddrl_b200, csrc/replay.cu); this file is host-side marshalling
only.  Additive fast paths that the reference does not have: store_batch (vectorised producers),
sample_many (several batches per launch), device=True outputs (no D2H), injected index streams.
"""
from __future__ import annotations

import ctypes as C
import functools
import threading
import os

import numpy as np
import torch

from . import _native as N


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def check_indices(idxs, size):
    """Injected row indices with numpy's fancy-indexing contract (the reference's `buf[idxs]`): negative indices wrap,
    anything outside [-size, size) raises IndexError — the device kernels dereference what they are given.  Accepts
    numpy arrays / sequences (checked on the host) and torch tensors (one min/max reduction where they live)."""
    if isinstance(idxs, torch.Tensor):
        t = idxs.reshape(-1).to(torch.int64)
        if t.numel() == 0:
            return t
        lo, hi = int(t.min()), int(t.max())
        if lo < -size or hi >= size:
            raise IndexError(f"index {hi if hi >= size else lo} is out of bounds for axis 0 with size {size}")
        return torch.where(t < 0, t + size, t) if lo < 0 else t
    a = np.asarray(idxs).reshape(-1).astype(np.int64, copy=False)
    if a.size == 0:
        return a
    lo, hi = int(a.min()), int(a.max())
    if lo < -size or hi >= size:
        raise IndexError(f"index {hi if hi >= size else lo} is out of bounds for axis 0 with size {size}")
    return np.where(a < 0, a + size, a) if lo < 0 else a


class HostBatch(dict):
    """sample_batch()'s host result: the reference's dict of five numpy arrays.  The arrays are views of ONE
    pinned block (`block`, the layout ddrl_rb_sample_host wrote), which lets Learner.train() feed the batch
    back to the GPU with a single H2D copy instead of five."""
    block = None      # pinned torch.uint8 tensor backing the arrays
    n = 0             # rows


class PendingBatch:
    """A sample_batch_async() result: `wait()` returns the HostBatch once its D2H copy has completed."""

    def __init__(self, event, batch):
        self._event, self._batch = event, batch

    def ready(self):
        return self._event.query()

    def wait(self):
        self._event.synchronize()
        return self._batch


class Cache:
    """The learner-side prefetcher of the reference (algos/sac1/sac1.py:103-130: a process that keeps Queue(10) filled with
    `sample_batch(opt.batch_size)` results so that `batch = cache.q1.get()` never waits for the replay actor, and that
    forwards `cache.q2.put(agent.get_weights())` to the parameter server).  Same surface — Cache(replay_buffer), .start(),
    .end(), .q1.get(), .q2.put((keys, values)) — without a process or a host queue: `depth` samples are kept IN FLIGHT on a
    dedicated CUDA stream (gather kernel + D2H copy into pinned blocks), so batch k+1 crosses PCIe while the learner's
    stream runs update k.  Batches are drawn up to `depth` calls ahead of the stores that precede their consumption — the
    reference's queue is up to 10 batches stale in the same way."""

    class _Q1:
        def __init__(self, cache):
            self._c = cache

        def get(self):
            return self._c._next()

        def qsize(self):
            return sum(1 for p in self._c._pending if p.ready())

        def empty(self):
            return self.qsize() == 0

    class _Q2:
        def __init__(self, cache):
            self._c = cache

        def put(self, kv):
            if self._c.ps is not None:
                keys, values = kv
                self._c.ps.push(keys, values)

        def qsize(self):
            return 0

        def empty(self):
            return True

    def __init__(self, replay_buffer, batch_size=None, *, depth=2, ps=None):
        from collections import deque
        self.replay_buffer, self.ps = replay_buffer, ps
        self.batch_size = batch_size
        self.depth = max(1, int(depth))
        self._pending = deque()
        self._stream = None
        self.q1, self.q2 = Cache._Q1(self), Cache._Q2(self)

    def start(self, batch_size=None):
        if batch_size is not None:
            self.batch_size = batch_size
        if self.batch_size is None:
            raise ValueError("Cache needs a batch size (the reference reads the global opt.batch_size)")
        self._stream = torch.cuda.Stream(device=self.replay_buffer.device)
        self._fill()

    def _fill(self):
        while len(self._pending) < self.depth:
            self._pending.append(self.replay_buffer.sample_batch_async(self.batch_size, stream=self._stream))

    def _next(self):
        if self._stream is None:
            self.start()
        batch = self._pending.popleft().wait()
        self._fill()
        return batch

    def end(self):
        if self._stream is not None:
            self._stream.synchronize()
        self._pending.clear()


def _locked(fn):
    """Serialise a method on the buffer's lock: producers (store / store_batch from rollout threads, each on its own CUDA
    stream) and the learner (sample_batch / train_from_buffer) may call one buffer concurrently, as the reference's
    workers call one Ray actor (algos/sac1/sac1.py:195); the host-side staging and counters are shared state."""
    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        with self._lock:
            return fn(self, *a, **k)
    return wrapper


class ReplayBuffer:
    """Drop-in for the reference ReplayBuffer.

    Extra keyword arguments (all optional, defaults keep the reference call pattern working):
      device        CUDA device index (default: torch.cuda.current_device())
      flavor        "sac1" | "dsac" | "dqn": which reference variant's get_counts()/acts shape to mimic
      index_source  "philox": indices drawn on the GPU (Philox4x32-10 keyed by `seed`)
                    "numpy" : indices drawn on the host with np.random.randint exactly where the
                              reference draws them (bit-identical batches to the reference under the
                              same np.random.seed), then injected
      seed          Philox key; default: drawn once from numpy's global RandomState at first use, so
                    `np.random.seed(opt.seed)` (algos/sac1/actor_learner.py:24) makes runs repeatable
      rng_stream    Philox sub-stream (rank of the shard in multi-GPU runs)
      stage_rows    rows of pinned host staging for single-transition store()
    """

    def __init__(self, obs_dim, act_dim, size, *, device=None, flavor="sac1", index_source="philox",
                 seed=None, rng_stream=0, stage_rows=1024):
        if not torch.cuda.is_available():
            raise RuntimeError("ddrl_b200.ReplayBuffer needs a CUDA device (no CPU fallback)")
        if flavor not in ("sac1", "dsac", "dqn"):
            raise ValueError(f"unknown flavor {flavor!r}")
        if index_source not in ("philox", "numpy"):
            raise ValueError(f"unknown index_source {index_source!r}")
        self._lib = N.lib()
        self._lock = threading.RLock()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.flavor = flavor
        self.obs_dim = int(obs_dim)
        self._scalar_act = flavor == "dqn" or act_dim in (0, None)
        self.act_dim = 1 if self._scalar_act else int(act_dim)
        self.max_size = int(size)
        self.index_source = index_source
        self._seed = None if seed is None else int(seed) & (2 ** 64 - 1)
        self._rng_stream = int(rng_stream) & 0xFFFFFFFF
        self._counter = 0
        h = C.c_void_p()
        N.check(self._lib.ddrl_rb_create(self.device, self.obs_dim, self.act_dim, self.max_size, C.byref(h)))
        self._h = h
        # pinned staging for store(): two sets, alternated so a set is never rewritten while its
        # H2D copy may still be in flight
        self._stage_rows = max(1, int(stage_rows))
        self._stages = [self._new_stage() for _ in range(2)]
        self._cur = 0
        self._staged = 0

    # ------------------------------------------------------------------------------------------
    def _new_stage(self):
        S, D, A = self._stage_rows, self.obs_dim, self.act_dim
        pin = dict(dtype=torch.float32, pin_memory=True)
        t = dict(obs=torch.empty((S, D), **pin), nxt=torch.empty((S, D), **pin),
                 act=torch.empty((S, A), **pin), rew=torch.empty(S, **pin), done=torch.empty(S, **pin))
        st = {k: (v, v.numpy()) for k, v in t.items()}
        st["event"] = None
        return st

    def _stream(self):
        return torch.cuda.current_stream(self.device)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ddrl_rb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def close(self):
        self.__del__()

    # ------------------------------------------------------------------------------------------
    # store
    # ------------------------------------------------------------------------------------------
    @_locked
    def store(self, obs, act, rew, next_obs, done):
        """One transition (reference signature).  Arguments are captured by value at call time
        (numpy assignment into pinned staging performs the same casts as the reference's
        ``buf[ptr] = x``); rows reach the GPU ring in batches, at the latest before the next
        sample / export / get of ring state."""
        st = self._stages[self._cur]
        k = self._staged
        if k == 0 and st["event"] is not None:
            st["event"].synchronize()
            st["event"] = None
        st["obs"][1][k] = obs
        st["nxt"][1][k] = next_obs
        st["act"][1][k] = act
        st["rew"][1][k] = rew
        st["done"][1][k] = done
        self._staged = k + 1
        if self._staged == self._stage_rows:
            self.flush()

    @_locked
    def flush(self):
        """Push staged single-transition stores to the GPU ring."""
        k = self._staged
        if k == 0:
            return
        st = self._stages[self._cur]
        s = self._stream()
        N.check(self._lib.ddrl_rb_store_batch_host(
            self._h, _ptr(st["obs"][0]), _ptr(st["act"][0]), _ptr(st["rew"][0]), _ptr(st["nxt"][0]),
            _ptr(st["done"][0]), k, N.F32, C.c_void_p(s.cuda_stream)))
        ev = torch.cuda.Event()
        ev.record(s)
        st["event"] = ev
        self._staged = 0
        self._cur ^= 1

    @_locked
    def store_batch(self, obs, act, rew, next_obs, done):
        """n transitions at once (== n sequential store() calls in row order).  Accepts numpy
        arrays (copied through pinned staging by the library call) or torch tensors; CUDA tensors
        on this device are consumed in place with no host round-trip (float32 or float64, cast on the GPU)."""
        self.flush()
        arrs = [obs, act, rew, next_obs, done]
        if all(isinstance(a, torch.Tensor) and a.is_cuda for a in arrs):
            return self._store_device(*arrs)
        np_arrs = []
        for a in arrs:
            if isinstance(a, torch.Tensor):
                a = a.detach().cpu().numpy()
            np_arrs.append(np.asarray(a))
        n = int(np_arrs[2].shape[0])
        D, A = self.obs_dim, self.act_dim
        shapes = [(n, D), (n, A), (n,), (n, D), (n,)]
        s = self._stream()
        # host-side cast to float32 is numpy's own assignment cast, i.e. the reference's `buf[ptr] = x` semantics (a no-op
        # for contiguous float32 inputs); the library captures the arrays BY VALUE into its pinned staging at call time
        # and queues one H2D copy + the store kernel — nothing on the producer's call path waits for the GPU
        host = [np.ascontiguousarray(a.reshape(sh), dtype=np.float32) for a, sh in zip(np_arrs, shapes)]
        N.check(self._lib.ddrl_rb_store_batch_host_copy(
            self._h, *[a.ctypes.data for a in host], n, N.F32, s.cuda_stream))

    def _store_device(self, obs, act, rew, next_obs, done):
        n = int(rew.shape[0])
        D, A = self.obs_dim, self.act_dim
        dt = torch.float64 if obs.dtype == torch.float64 else torch.float32
        dev = torch.device("cuda", self.device)
        shapes = [(n, D), (n, A), (n,), (n, D), (n,)]
        ts = [t.to(device=dev, dtype=dt).reshape(sh).contiguous() for t, sh in zip((obs, act, rew, next_obs, done), shapes)]
        s = self._stream()
        N.check(self._lib.ddrl_rb_store_batch(self._h, *[_ptr(t) for t in ts], n,
                                              N.F64 if dt == torch.float64 else N.F32,
                                              C.c_void_p(s.cuda_stream)))
        # keep the inputs alive until the stream has consumed them
        for t in ts:
            t.record_stream(s)

    # ------------------------------------------------------------------------------------------
    # sample
    # ------------------------------------------------------------------------------------------
    def _philox_seed(self):
        if self._seed is None:
            # one draw from the same global RandomState the reference samples from
            self._seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        return self._seed

    def _counts_native(self):
        v = [C.c_int64() for _ in range(5)]
        N.check(self._lib.ddrl_rb_counts(self._h, *[C.byref(x) for x in v]))
        return [int(x.value) for x in v]  # ptr, size, capacity, steps, sample_times

    def _shape_out(self, n_batches, batch, many):
        lead = (n_batches, batch) if many else (batch,)
        return lead

    @_locked
    def _sample(self, batch_size, n_batches, idxs, device, many, return_idxs):
        self.flush()
        batch_size, n_batches = int(batch_size), int(n_batches)
        n = batch_size * n_batches
        if self.size == 0:
            # the reference's np.random.randint(0, 0, ...) raises exactly this
            raise ValueError("high <= 0")
        if idxs is None and self.index_source == "numpy":
            idxs = np.random.randint(0, self.size, size=n)
        elif idxs is not None:
            idxs = check_indices(idxs, self.max_size)      # `buf[idxs]` indexes the whole array, filled or not
        s = self._stream()
        D, A = self.obs_dim, self.act_dim
        lead = (n_batches, batch_size) if many else (batch_size,)
        act_shape = lead if self._scalar_act else lead + (A,)
        seed = 0 if idxs is not None else self._philox_seed()
        counter = self._counter
        if device:
            dev = torch.device("cuda", self.device)
            d_idx = None
            if idxs is not None:
                d_idx = torch.as_tensor(idxs, dtype=torch.int64).reshape(-1).to(dev, non_blocking=False).contiguous()
                if d_idx.numel() != n:
                    raise ValueError(f"idxs has {d_idx.numel()} entries, expected {n}")
            f32 = dict(dtype=torch.float32, device=dev)
            out = dict(obs1=torch.empty(lead + (D,), **f32), obs2=torch.empty(lead + (D,), **f32),
                       acts=torch.empty(act_shape, **f32), rews=torch.empty(lead, **f32),
                       done=torch.empty(lead, **f32))
            o_idx = torch.empty(lead, dtype=torch.int64, device=dev) if return_idxs else None
            N.check(self._lib.ddrl_rb_sample(
                self._h, batch_size, n_batches, _ptr(d_idx), seed, counter, self._rng_stream,
                _ptr(out["obs1"]), _ptr(out["obs2"]), _ptr(out["acts"]), _ptr(out["rews"]),
                _ptr(out["done"]), _ptr(o_idx), C.c_void_p(s.cuda_stream)))
            if d_idx is not None:
                d_idx.record_stream(s)
            if return_idxs:
                out["idxs"] = o_idx
        else:
            h_idx = None
            if idxs is not None:
                if isinstance(idxs, torch.Tensor):
                    idxs = idxs.detach().cpu().numpy()
                h_idx = np.ascontiguousarray(np.asarray(idxs).reshape(-1), dtype=np.int64)
                if h_idx.size != n:
                    raise ValueError(f"idxs has {h_idx.size} entries, expected {n}")
            nbytes = int(self._lib.ddrl_rb_sample_block_bytes(self._h, n))
            block = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            N.check(self._lib.ddrl_rb_sample_host(
                self._h, batch_size, n_batches, C.c_void_p(h_idx.ctypes.data) if h_idx is not None else None,
                seed, counter, self._rng_stream, _ptr(block), nbytes, C.c_void_p(s.cuda_stream)))
            out = self._host_batch(block, nbytes, n, lead, act_shape, return_idxs)
        if idxs is None:
            self._counter += 1
        return out

    def _host_batch(self, block, nbytes, n, lead, act_shape, return_idxs):
        D, A = self.obs_dim, self.act_dim
        raw = block.numpy()
        f = raw[: n * (2 * D + A + 2) * 4].view(np.float32)
        o = 0
        out = HostBatch()
        out.block, out.n = block, n
        for key, width, shape in (("obs1", D, lead + (D,)), ("obs2", D, lead + (D,)),
                                  ("acts", A, act_shape), ("rews", 1, lead), ("done", 1, lead)):
            out[key] = f[o:o + n * width].reshape(shape)
            o += n * width
        if return_idxs:
            out["idxs"] = raw[nbytes - n * 8:].view(np.int64).reshape(lead)
        return out

    @_locked
    def sample_batch_async(self, batch_size=128, *, stream=None):
        """sample_batch() issued AHEAD of its consumer: the gather and the D2H copy into a fresh pinned block are queued on
        `stream` (default: the current stream) and the call returns at once with a PendingBatch; `.wait()` blocks until
        the block has arrived and returns the usual dict of numpy arrays.  The index stream advances exactly as for
        sample_batch(), so a sequence of async calls returns the same batches as the same sequence of blocking calls —
        each drawn from the ring as it was when ITS turn came on the stream."""
        self.flush()
        B = int(batch_size)
        if self.size == 0:
            raise ValueError("high <= 0")
        idxs = np.random.randint(0, self.size, size=B) if self.index_source == "numpy" else None
        s = stream if stream is not None else self._stream()
        h_idx = np.ascontiguousarray(idxs, dtype=np.int64) if idxs is not None else None
        nbytes = int(self._lib.ddrl_rb_sample_block_bytes(self._h, B))
        block = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        N.check(self._lib.ddrl_rb_sample_host_async(
            self._h, B, 1, C.c_void_p(h_idx.ctypes.data) if h_idx is not None else None,
            0 if idxs is not None else self._philox_seed(), self._counter, self._rng_stream, _ptr(block), nbytes,
            C.c_void_p(s.cuda_stream)))
        if h_idx is not None:
            s.synchronize()              # the pageable index array was read by an asynchronous copy
        else:
            self._counter += 1
        ev = torch.cuda.Event()
        ev.record(s)
        act_shape = (B,) if self._scalar_act else (B, self.act_dim)
        return PendingBatch(ev, self._host_batch(block, nbytes, B, (B,), act_shape, False))

    def sample_batch(self, batch_size=128, *, idxs=None, device=False, return_idxs=False):
        """Reference signature.  Returns dict(obs1, obs2, acts, rews, done): float32 numpy arrays
        (fresh, C-contiguous) or, with device=True, CUDA tensors that never leave the GPU."""
        if self.flavor == "dqn" and batch_size is None:
            raise TypeError("dqn flavour: pass opt.batch_size explicitly")
        return self._sample(batch_size, 1, idxs, device, False, return_idxs)

    def sample_many(self, n_batches, batch_size=128, *, idxs=None, device=True, return_idxs=False):
        """n_batches sample_batch() calls in ONE kernel launch; outputs are stacked
        [n_batches, batch_size, ...].  Counts as n_batches samples (sample_times += n_batches)."""
        return self._sample(batch_size, n_batches, idxs, device, True, return_idxs)

    # ------------------------------------------------------------------------------------------
    # counters / state
    # ------------------------------------------------------------------------------------------
    @property
    def ptr(self):
        return (self._counts_native()[0] + self._staged) % self.max_size

    @property
    def size(self):
        return min(self._counts_native()[1] + self._staged, self.max_size)

    @property
    def steps(self):
        return self._counts_native()[3] + self._staged

    @property
    def sample_times(self):
        return self._counts_native()[4]

    rollout_steps = steps  # example/dsac.py naming

    @_locked
    def get_counts(self):
        p, size, cap, steps, samples = self._counts_native()
        steps += self._staged
        size = min(size + self._staged, cap)
        if self.flavor == "dsac":
            return steps                      # example/dsac.py:47-48
        return samples, steps, size           # algos/sac1/sac1.py:62-63 ; dqn: (learner_steps, actor_steps, size)

    @_locked
    def ring_arrays(self):
        """The reference's five arrays (obs1_buf, obs2_buf, acts_buf, rews_buf, done_buf) as numpy,
        unpacked from the GPU ring — for checkpoints and parity tests, not a hot path."""
        self.flush()
        dev = torch.device("cuda", self.device)
        cap, D, A = self.max_size, self.obs_dim, self.act_dim
        f32 = dict(dtype=torch.float32, device=dev)
        o1, o2 = torch.empty((cap, D), **f32), torch.empty((cap, D), **f32)
        oa, orw, od = torch.empty((cap, A), **f32), torch.empty(cap, **f32), torch.empty(cap, **f32)
        s = self._stream()
        N.check(self._lib.ddrl_rb_export(self._h, _ptr(o1), _ptr(o2), _ptr(oa), _ptr(orw), _ptr(od),
                                         C.c_void_p(s.cuda_stream)))
        s.synchronize()
        acts = oa.cpu().numpy()
        if self._scalar_act:
            acts = acts.reshape(cap)
        return dict(obs1_buf=o1.cpu().numpy(), obs2_buf=o2.cpu().numpy(), acts_buf=acts,
                    rews_buf=orw.cpu().numpy(), done_buf=od.cpu().numpy())

    @_locked
    def load_ring_arrays(self, obs1_buf, obs2_buf, acts_buf, rews_buf, done_buf, ptr, size, steps=0,
                         sample_times=0):
        self.flush()
        dev = torch.device("cuda", self.device)
        cap, D, A = self.max_size, self.obs_dim, self.act_dim
        def up(a, shape):
            return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))).to(dev)
        ts = [up(obs1_buf, (cap, D)), up(obs2_buf, (cap, D)), up(acts_buf, (cap, A)), up(rews_buf, (cap,)),
              up(done_buf, (cap,))]
        s = self._stream()
        N.check(self._lib.ddrl_rb_import(self._h, *[_ptr(t) for t in ts], int(ptr), int(size), int(steps),
                                         int(sample_times), C.c_void_p(s.cuda_stream)))
        s.synchronize()

    # -- algos/dqn/train.py:82-108 on-disk format -------------------------------------------------
    def save(self, checkpoint_dir, buffer_index=0):
        os.makedirs(checkpoint_dir, exist_ok=True)
        arrs = self.ring_arrays()
        for name, a in arrs.items():
            np.save(os.path.join(checkpoint_dir, f"{name}-{buffer_index}"), a)
        p, size, cap, steps, samples = self._counts_native()
        np.save(os.path.join(checkpoint_dir, f"buffer_infos-{buffer_index}"),
                np.array((p, size, cap, steps, samples)))

    def load(self, checkpoint_dir, buffer_index=0):
        ld = lambda n: np.load(os.path.join(checkpoint_dir, f"{n}-{buffer_index}.npy"))
        infos = ld("buffer_infos")
        if int(infos[2]) != self.max_size:
            raise ValueError(f"checkpoint capacity {int(infos[2])} != buffer capacity {self.max_size}")
        self.load_ring_arrays(ld("obs1_buf"), ld("obs2_buf"), ld("acts_buf"), ld("rews_buf"), ld("done_buf"),
                              ptr=int(infos[0]), size=int(infos[1]), steps=int(infos[3]),
                              sample_times=int(infos[4]))

    # handle for native consumers (fused sample->update)
    @property
    def native_handle(self):
        self.flush()
        return self._h
