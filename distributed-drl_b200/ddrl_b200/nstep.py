"""N-step sequence replay — the SQN_N_STEP ring of the reference's later-generation SAC driver
(algos/sac1/sac_ray.py:34-83; `Ln` from algos/sac1/hyperparams.py:84), HBM-resident.

Reference call surface (one Ray actor per buffer shard):
    ReplayBuffer(opt)                                   opt.Ln, opt.obs_shape, opt.act_shape, opt.buffer_size,
                                                        opt.batch_size, opt.num_buffers
    .store(o_queue, a_r_d_queue, worker_index)          o_queue: Ln+1 entries (obs,), a_r_d_queue: Ln entries (a, r, d)
    .sample_batch()                                     -> dict(obs [B, Ln+1, ...], acts [B, Ln, ...], rews [B, Ln], done [B, Ln])
    .get_counts()                                       -> (sample_times, steps, size); both counters advance by
                                                        opt.num_buffers per call (sac_ray.py:71-79)
Only the float32 observation layout is mirrored (the reference also has an lz4-packed string layout for image
observations, sac_ray.py:42-43, which is an encoding of the same rows).

One packed row per sequence [obs | acts | rews | done], padded to 4 floats; sample_batch is ONE launch of the
segmented gather (ddrl_seg_sample) that produces the four dense arrays of the reference's dict.  Additive:
store_batch (vectorised producers), device=True outputs, injected index streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N


def row_layout(Ln, D, A):
    """(widths, offsets, used floats, padded row floats) of the packed sequence row [obs | acts | rews | done]."""
    widths = [(Ln + 1) * D, Ln * A, Ln, Ln]
    offsets = [0, widths[0], widths[0] + widths[1], widths[0] + widths[1] + Ln]
    used = sum(widths)
    return widths, offsets, used, (used + 3) // 4 * 4


class NStepReplayBuffer:
    def __init__(self, opt, *, device=None, seed=None, rng_stream=0, index_source="philox"):
        if not torch.cuda.is_available():
            raise RuntimeError("ddrl_b200.NStepReplayBuffer needs a CUDA device (no CPU fallback)")
        if index_source not in ("philox", "numpy"):
            raise ValueError(index_source)
        self.opt = opt
        self._lib = N.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._dev = torch.device("cuda", self.device)
        self.Ln = int(opt.Ln)
        self.obs_shape, self.act_shape = tuple(opt.obs_shape), tuple(opt.act_shape)
        self.D, self.A = int(np.prod(self.obs_shape)), int(np.prod(self.act_shape)) if self.act_shape else 1
        self.widths, self.offsets, self.used, self.row_f = row_layout(self.Ln, self.D, self.A)
        self.max_size = int(opt.buffer_size)
        self.batch_size = int(opt.batch_size)
        self.num_buffers = int(getattr(opt, "num_buffers", 1))
        self.ring = torch.zeros((self.max_size, self.row_f), dtype=torch.float32, device=self._dev)
        self.ptr = self.size = self.steps = self.sample_times = 0
        self.index_source = index_source
        self._seed = None if seed is None else int(seed) & (2 ** 64 - 1)
        self._rng_stream, self._counter = int(rng_stream) & 0xFFFFFFFF, 0

    # ---- store -----------------------------------------------------------------------------------
    def store_batch(self, obs, acts, rews, done):
        """n sequences at once (== n store() calls in row order); only the last `capacity` survive.  The four dense
        arrays ([n, Ln+1, ...], [n, Ln, ...], [n, Ln], [n, Ln]; numpy or CUDA tensors) are packed into ring rows by ONE
        launch of seg_store_rows (ddrl_seg_store); host arrays are cast to float32 (numpy assignment semantics, as the
        reference's `buffer[ptr] = np.array(..., dtype=np.float32)`) and uploaded, CUDA tensors are consumed in place."""
        ins = []
        for x, w in zip((obs, acts, rews, done), self.widths):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                t = x.to(self._dev, torch.float32).reshape(-1, w).contiguous()
            else:
                x = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
                t = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, w), dtype=np.float32)).to(self._dev)
            ins.append(t)
        n = int(ins[2].shape[0])
        if any(int(t.shape[0]) != n for t in ins):
            raise ValueError("store_batch: obs / acts / rews / done disagree on n")
        if n == 0:
            return
        s = torch.cuda.current_stream(self.device)
        off = (C.c_int * 4)(*self.offsets)
        wid = (C.c_int * 4)(*self.widths)
        ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in ins])
        N.check(self._lib.ddrl_seg_store(self.device, C.c_void_p(self.ring.data_ptr()), self.row_f, self.max_size, self.ptr, 4,
                                         off, wid, ptrs, n, C.c_void_p(s.cuda_stream)))
        for t in ins:
            t.record_stream(s)
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)
        self.steps += n * self.num_buffers

    def store(self, o_queue, a_r_d_queue, worker_index=0):
        obs, = np.stack(o_queue, axis=1)                               # sac_ray.py:55
        a, r, d = zip(*a_r_d_queue)
        self.store_batch(np.asarray(list(obs), dtype=np.float32)[None], np.asarray(list(a), dtype=np.float32)[None],
                         np.asarray(list(r), dtype=np.float32)[None], np.asarray(list(d), dtype=np.float32)[None])

    # ---- sample -----------------------------------------------------------------------------------
    def sample_batch(self, batch_size=None, *, idxs=None, device=False, return_idxs=False):
        B = self.batch_size if batch_size is None else int(batch_size)
        if self.size == 0:
            raise ValueError("high <= 0")
        if idxs is None and self.index_source == "numpy":
            idxs = np.random.randint(0, self.size, size=B)             # sac_ray.py:74
        f32 = dict(dtype=torch.float32, device=self._dev)
        outs = [torch.empty((B, w), **f32) for w in self.widths]
        o_idx = torch.empty(B, dtype=torch.int64, device=self._dev) if return_idxs else None
        d_idx = None
        if idxs is not None:
            d_idx = torch.as_tensor(np.asarray(idxs), dtype=torch.int64).reshape(-1).to(self._dev).contiguous()
            if d_idx.numel() != B:
                raise ValueError(f"idxs has {d_idx.numel()} entries, expected {B}")
        if self._seed is None:
            self._seed = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        s = torch.cuda.current_stream(self.device)
        off = (C.c_int * 4)(*self.offsets)
        wid = (C.c_int * 4)(*self.widths)
        ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in outs])
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        N.check(self._lib.ddrl_seg_sample(self.device, C.c_void_p(self.ring.data_ptr()), self.row_f, self.size, 4, off, wid,
                                          ptrs, B, p(d_idx), self._seed if idxs is None else 0, self._counter,
                                          self._rng_stream, p(o_idx), C.c_void_p(s.cuda_stream)))
        if idxs is None:
            self._counter += 1
        self.sample_times += self.num_buffers
        out = dict(obs=outs[0].view((B, self.Ln + 1) + self.obs_shape), acts=outs[1].view((B, self.Ln) + self.act_shape),
                   rews=outs[2], done=outs[3])
        if return_idxs:
            out["idxs"] = o_idx
        if not device:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out

    def prefetch(self, depth=8, batch_size=None):
        """Bounded prefetch of the learner loop (the reference's Cache thread + Queue(maxsize) in front of the sequence
        buffers, algos/sac1/sac_ray.py:86-145): an iterator that keeps up to `depth` sampled batches queued on the GPU —
        `depth` sample_batch launches are issued ahead of the consumer, on the current stream, and each batch is handed
        out in sampling order; batches never leave HBM, so no thread and no host queue are needed."""
        from collections import deque
        q = deque()
        while True:
            while len(q) < max(1, int(depth)):
                q.append(self.sample_batch(batch_size, device=True))
            yield q.popleft()

    def get_counts(self):
        return self.sample_times, self.steps, self.size
