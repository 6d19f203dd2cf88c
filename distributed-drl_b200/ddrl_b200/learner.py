"""Learner — the reference's SAC1 learner API over the CUDA step in libddrl_b200 (csrc/sac.cu).

Mirrors algos/sac1/actor_learner.py:
    Learner(opt, job)                  :20-123   (opt fields read: seed, obs_dim, act_dim,
                                                   ac_kwargs["action_space"].high[0], alpha, gamma, lr, polyak)
    .train(batch)                      :135-142  one sess.run(step_ops)
    .get_weights() -> (keys, values)   :129-133  all "main" variables, TF1 names
    .set_weights(keys, values)         :125-127  assign + target_init
and example/model.py's Model(args).train(replay_buffer, args) (:92-101).

Additions: hidden sizes are a parameter (opt.ac_kwargs["hidden_sizes"], default core.py:91's
(400, 300)); batches may already be on the GPU (ReplayBuffer.sample_batch(device=True)); with
torch.distributed initialised the step becomes data-parallel (gradient all-reduce over NCCL between
the backward and the optimiser), which the reference stubs out (:144-148).
"""
from __future__ import annotations

import collections
import ctypes as C
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _native as N

_KEYS_BATCH = ("obs1", "obs2", "acts", "rews", "done")


def param_names():
    """TF1 variable names in creation order (tf.layers.dense default naming inside the
    main/pi, main/q1, main/q2 scopes): dense_2 is the mu head, dense_3 the log_std head."""
    names = []
    for suffix in ("dense", "dense_1", "dense_2", "dense_3"):
        names += [f"main/pi/{suffix}/kernel", f"main/pi/{suffix}/bias"]
    for q in ("q1", "q2"):
        for suffix in ("dense", "dense_1", "dense_2"):
            names += [f"main/{q}/{suffix}/kernel", f"main/{q}/{suffix}/bias"]
    return names


def param_shapes(obs_dim, act_dim, hidden):
    h1, h2 = hidden
    D, A = int(obs_dim), int(act_dim)
    pi = [(D, h1), (h1,), (h1, h2), (h2,), (h2, A), (A,), (h2, A), (A,)]
    q = [(D + A, h1), (h1,), (h1, h2), (h2,), (h2, 1), (1,)]
    return OrderedDict(zip(param_names(), pi + q + q))


def glorot_init(obs_dim, act_dim, hidden, seed):
    """tf.layers.dense defaults: glorot_uniform kernels, zero biases (one-off, host side)."""
    g = np.random.Generator(np.random.PCG64(int(seed)))
    out = OrderedDict()
    for n, s in param_shapes(obs_dim, act_dim, hidden).items():
        if len(s) == 2:
            lim = math.sqrt(6.0 / (s[0] + s[1]))
            out[n] = g.uniform(-lim, lim, s).astype(np.float32)
        else:
            out[n] = np.zeros(s, dtype=np.float32)
    return out


class _DevPtr:
    """Zero-copy torch view of a device buffer owned by the native handle."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _opt_get(opt, name, default=None):
    if isinstance(opt, dict):
        return opt.get(name, default)
    return getattr(opt, name, default)


class Learner(object):
    def __init__(self, opt, job="learner", *, device=None, max_batch=None, process_group=None, gemm=None):
        """gemm: None / "tc" (default: tcgen05 tensor cores with the 3xTF32 split, fp32-class accuracy) | "ffma"
        (plain fp32 FFMA tiles).  Both hold the 1e-5 bar of tests/test_sac_gpu.py::test_one_step_matches_oracle."""
        if not torch.cuda.is_available():
            raise RuntimeError("ddrl_b200.Learner needs a CUDA device (no CPU fallback)")
        self.opt = opt
        self.job = job
        self._lib = N.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._dev = torch.device("cuda", self.device)
        ac = _opt_get(opt, "ac_kwargs", {}) or {}
        self.obs_dim, self.act_dim = int(_opt_get(opt, "obs_dim")), int(_opt_get(opt, "act_dim"))
        self.hidden = tuple(int(h) for h in ac.get("hidden_sizes", (400, 300)))
        if len(self.hidden) != 2:
            raise ValueError("two hidden layers (the reference's mlp_actor_critic shape) are supported")
        space = ac.get("action_space", None)
        self.act_scale = float(space.high[0]) if space is not None else float(_opt_get(opt, "act_scale", 1.0))
        alpha = _opt_get(opt, "alpha", 0.2)
        self.auto_alpha = alpha == "auto"
        self.alpha = -1.0 if self.auto_alpha else float(alpha)
        gamma = _opt_get(opt, "gamma", 0.99)
        if isinstance(gamma, (tuple, list)):      # example/dsac.py:198's trailing comma makes gamma a tuple
            gamma = gamma[0]
        self.gamma, self.polyak, self.lr = float(gamma), float(_opt_get(opt, "polyak", 0.995)), float(_opt_get(opt, "lr", 1e-3))
        self.seed = int(_opt_get(opt, "seed", 0) or 0)
        self.max_batch = int(max_batch or _opt_get(opt, "batch_size", 256) or 256)
        h = C.c_void_p()
        if gemm not in (None, "tc", "ffma"):
            raise ValueError(f"gemm must be None, 'tc' or 'ffma', not {gemm!r}")
        mode = N.GEMM_AUTO if gemm is None else N.GEMM_TC if gemm == "tc" else N.GEMM_FFMA
        N.check(self._lib.ddrl_sac_create(self.device, self.obs_dim, self.act_dim, self.hidden[0], self.hidden[1],
                                          self.max_batch, self.gamma, self.polyak, self.lr, self.alpha,
                                          self.act_scale, mode, C.byref(h)))
        self._h = h
        self.names = param_names()
        self.shapes = param_shapes(self.obs_dim, self.act_dim, self.hidden)
        self.P = int(self._lib.ddrl_sac_param_count(self._h))
        assert self.P == sum(int(np.prod(s)) for s in self.shapes.values())
        gp, cnt, ap = C.c_void_p(), C.c_int64(), C.c_void_p()
        N.check(self._lib.ddrl_sac_grad_buffer(self._h, C.byref(gp), C.byref(cnt), C.byref(ap)))
        self._grad = torch.as_tensor(_DevPtr(gp.value, int(cnt.value)), device=self._dev)
        self._alpha_stat = torch.as_tensor(_DevPtr(ap.value, 1), device=self._dev)
        self._pg = process_group
        self._fused = False
        self._outs = None
        self.nvls = False
        self._hscal, self._hscal_event, self._hscal_ring = None, None, None
        self._via_host_key, self._via_host_nbytes = None, 0
        self._inflight = collections.deque()
        self._stage = {}
        self.steps = 0
        init = glorot_init(self.obs_dim, self.act_dim, self.hidden, self.seed)
        self.set_weights(list(init.keys()), list(init.values()))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ddrl_sac_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device)

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self._pg)
        return 1

    def connect_peers(self, group=None):
        """Fused data-parallel mode (collective; call on every rank of one node before the first train()):
        exchanges the CUDA IPC handles of the ranks' gradient buffers, after which the all-reduce between
        compute_grads and apply_grads is done INSIDE the optimiser kernel over NVLink peer memory (no NCCL call
        on the step path).  Without it, train() uses torch.distributed.all_reduce (NCCL)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2:
            return False
        # NVLS first (measured on 8 B200: 116.4 vs 118.4 us per C2 step, replicas bit-identical after 200 steps —
        # tools/dp_replica_check.py); CUDA IPC + peer reads when multicast is unavailable, with DDRL_DP_NVLS=0, or for the
        # two-kernel form of the exchange
        if (os.environ.get("DDRL_DP_NVLS", "1") != "0" and os.environ.get("DDRL_DP_V1", "0") != "1"
                and self._connect_symmetric(group, world, rank)):
            return True
        buf = (C.c_ubyte * 64)()
        N.check(self._lib.ddrl_sac_comm_export(self._h, buf))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(buf), group=group)
        blob = b"".join(handles)
        N.check(self._lib.ddrl_sac_comm_attach(self._h, world, rank, C.c_char_p(blob)))
        self._pg = group
        self._fused = True
        dist.barrier(group)
        return True

    def _connect_symmetric(self, group, world, rank):
        """Exchange buffers from torch's symmetric-memory allocator (device memory plumbing): every rank's buffer mapped in
        every process AND, where the NVSwitch supports it, one multicast (NVLS) mapping of all of them, through which the
        optimiser kernel reads the sum of the ranks' gradients with multimem.ld_reduce.  Returns False (after agreeing with
        the other ranks) when the allocator or the multicast mapping is unavailable; the caller then uses CUDA IPC."""
        import torch.distributed as dist
        ok, hdl, t = 1, None, None
        try:
            import torch.distributed._symmetric_memory as symm
            n = int(self._lib.ddrl_sac_comm_bytes(self._h))
            t = symm.empty((n + 3) // 4, dtype=torch.float32, device=self._dev)
            t.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
            if int(hdl.multicast_ptr) == 0:
                ok = 0
        except Exception:       # noqa: BLE001  (allocator / driver without fabric or multicast support)
            ok = 0
        flag = torch.tensor([ok], device=self._dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)        # all ranks take the same path
        if int(flag.item()) == 0:
            return False
        ptrs = (C.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
        N.check(self._lib.ddrl_sac_comm_attach_ptrs(self._h, world, rank, ptrs, C.c_void_p(int(hdl.multicast_ptr))))
        self._symm = (t, hdl)       # keeps the allocation and its mappings alive
        self._pg = group
        self._fused = True
        self.nvls = True
        dist.barrier(group)
        return True

    def comm_error(self):
        e = C.c_int()
        N.check(self._lib.ddrl_sac_comm_error(self._h, C.byref(e)))
        return int(e.value)

    # ---- weights -------------------------------------------------------------------------------
    def _flat_from(self, keys, values):
        """flat float32 vector in canonical order, starting from the current weights so that a
        subset of keys can be assigned (the reference assigns by name)."""
        given = dict(zip(keys, values))
        unknown = [k for k in given if k not in self.shapes]
        if unknown:
            raise KeyError(f"unknown variable names {unknown[:3]}")
        if len(given) == len(self.names):
            base = None
        else:
            base = self.get_flat_weights().cpu().numpy()
        parts, o = [], 0
        for n in self.names:
            size = int(np.prod(self.shapes[n]))
            if n in given:
                v = given[n]
                v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
                if tuple(v.shape) != tuple(self.shapes[n]):
                    raise ValueError(f"{n}: shape {v.shape} != {self.shapes[n]}")
                parts.append(np.ascontiguousarray(v, dtype=np.float32).reshape(-1))
            else:
                parts.append(base[o:o + size])
            o += size
        return np.concatenate(parts)

    def set_weights(self, variable_names, weights):
        flat = torch.from_numpy(self._flat_from(variable_names, weights)).to(self._dev)
        self.set_flat_weights(flat)

    def set_flat_weights(self, flat, also_target=True):
        """flat: CUDA float32 [P] in canonical order (device-to-device; used by the NCCL parameter
        broadcast).  Re-initialises the target like the reference's set_weights."""
        flat = flat.to(self._dev, torch.float32).contiguous()
        assert flat.numel() == self.P
        s = self._stream()
        N.check(self._lib.ddrl_sac_set_weights(self._h, C.c_void_p(flat.data_ptr()), 1 if also_target else 0,
                                               C.c_void_p(s.cuda_stream)))
        flat.record_stream(s)

    def get_flat_weights(self, which="main"):
        code = {"main": 0, "target": 1, "adam_m": 2, "adam_v": 3, "grad": 4}[which]
        out = torch.empty(self.P, dtype=torch.float32, device=self._dev)
        N.check(self._lib.ddrl_sac_get_weights(self._h, C.c_void_p(out.data_ptr()), code,
                                               C.c_void_p(self._stream().cuda_stream)))
        return out

    def _split(self, flat_np):
        vals, o = [], 0
        for n in self.names:
            size = int(np.prod(self.shapes[n]))
            vals.append(flat_np[o:o + size].reshape(self.shapes[n]).copy())
            o += size
        return vals

    def get_weights(self):
        flat = self.get_flat_weights("main").cpu().numpy()
        return list(self.names), self._split(flat)

    def get_target_weights(self):
        return list(self.names), self._split(self.get_flat_weights("target").cpu().numpy())

    # ---- training ------------------------------------------------------------------------------
    def _to_device(self, batch):
        """numpy batches go through pinned staging (one H2D per array); CUDA tensors pass through."""
        out = []
        s = self._stream()
        D, A = self.obs_dim, self.act_dim
        blk = getattr(batch, "block", None)
        if blk is not None and self._views_of_block(batch, blk):
            # our own ReplayBuffer.sample_batch() host result: five views of one pinned block -> ONE H2D copy
            n = int(batch.n)
            nb = n * (2 * D + A + 2) * 4
            ent = self._stage.get(("block", nb))
            if ent is None:
                dev = torch.empty(nb, dtype=torch.uint8, device=self._dev)
                f = dev.view(torch.float32)
                views, o = [], 0
                for width, shape in ((D, (n, D)), (D, (n, D)), (A, (n, A)), (1, (n,)), (1, (n,))):
                    views.append(f[o:o + n * width].view(shape))
                    o += n * width
                ent = (dev, views)
                self._stage[("block", nb)] = ent
            dev, views = ent
            dev.copy_(blk[:nb], non_blocking=True)
            self._keep = blk    # (torch's pinned-memory allocator also defers reuse until the copy has run)
            return views
        if all(isinstance(batch[k], np.ndarray) for k in _KEYS_BATCH):
            # generic host dict: pack into one pinned block (numpy assignment casts), ONE H2D copy
            n = int(np.asarray(batch["rews"]).shape[0])
            nf = n * (2 * D + A + 2)
            key = ("pack", nf)
            st = self._stage.get(key)
            if st is None:
                st = [(torch.empty(nf, dtype=torch.float32, pin_memory=True), [None]) for _ in range(2)] + [0]
                self._stage[key] = st
                self._stage[("packdev", nf)] = torch.empty(nf, dtype=torch.float32, device=self._dev)
            pin, ev = st[st[2]]
            st[2] ^= 1
            if ev[0] is not None:
                ev[0].synchronize()
            hp = pin.numpy()
            dev = self._stage[("packdev", nf)]
            o = 0
            for k, width, shape in (("obs1", D, (n, D)), ("obs2", D, (n, D)), ("acts", A, (n, A)), ("rews", 1, (n,)),
                                    ("done", 1, (n,))):
                np.copyto(hp[o:o + n * width].reshape(shape), np.asarray(batch[k]).reshape(shape), casting="unsafe")
                out.append(dev[o:o + n * width].view(shape))
                o += n * width
            dev.copy_(pin, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s)
            ev[0] = e
            return out
        for k in _KEYS_BATCH:
            v = batch[k]
            if isinstance(v, torch.Tensor) and v.is_cuda:
                out.append(v.to(torch.float32).contiguous())
                continue
            a = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            a = np.ascontiguousarray(a, dtype=np.float32)
            key = (k, a.shape)
            st = self._stage.get(key)
            if st is None:
                st = (torch.empty(a.shape, dtype=torch.float32, pin_memory=True),
                      torch.empty(a.shape, dtype=torch.float32, device=self._dev), [None])
                self._stage[key] = st
            pin, dev, ev = st
            if ev[0] is not None:
                ev[0].synchronize()
            pin.numpy()[...] = a
            dev.copy_(pin, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s)
            ev[0] = e
            out.append(dev)
        return out

    def _views_of_block(self, batch, blk):
        """True when the five entries of a HostBatch still ARE the views of its pinned block that sample_batch()
        returned (same address, dtype, layout).  A caller may rebind an entry (batch["rews"] = batch["rews"] * scale,
        n-step rewards, clipping ...): such a batch must take the generic packing path, not the stale block."""
        D, A = self.obs_dim, self.act_dim
        n = int(getattr(batch, "n", 0))
        if n <= 0 or blk.numel() < n * (2 * D + A + 2) * 4:
            return False
        addr = blk.data_ptr()
        for k, width in (("obs1", D), ("obs2", D), ("acts", A), ("rews", 1), ("done", 1)):
            v = batch.get(k)
            if not (isinstance(v, np.ndarray) and v.dtype == np.float32 and v.size == n * width and v.flags.c_contiguous
                    and v.ctypes.data == addr):
                return False
            addr += 4 * n * width
        return True

    def _noise_seed(self):
        """Philox key of the on-GPU policy noise.  Data-parallel replicas share opt.seed (it also seeds the initial
        weights), so the rank is folded in: every rank draws its own eps1/eps2/eps3 for its shard of the batch."""
        world = self._world()
        if world > 1:
            import torch.distributed as dist
            return (self.seed + 0x9E3779B97F4A7C15 * (dist.get_rank(self._pg) + 1)) & 0xFFFFFFFFFFFFFFFF
        return self.seed

    def train(self, batch, noise=None, sync_outputs=False, split=False):
        """One SAC1 update on `batch` (dict obs1/obs2/acts/rews/done).  `noise` ([3,B,A]) injects the
        three normal draws (parity tests); default: drawn on the GPU.  Returns the reference's fetch
        list as a dict of CUDA tensors (pi_loss, q1_loss, q2_loss, alpha in `scalars`; q1, q2,
        logp_pi), valid until the next call, without synchronising."""
        blk = getattr(batch, "block", None)
        if (blk is not None and noise is None and not split and (self._world() == 1 or self._fused)
                and self._views_of_block(batch, blk)):
            return self._train_host_block(batch, blk, sync_outputs)
        x, x2, a, r, d = self._to_device(batch)
        B = int(r.shape[0])
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch} (pass max_batch= to Learner)")
        if self._outs is None or self._outs["q1"].shape[0] != B:
            f = dict(dtype=torch.float32, device=self._dev)
            self._outs = dict(scalars=torch.zeros(4, **f), q1=torch.empty(B, **f), q2=torch.empty(B, **f),
                              logp_pi=torch.empty(B, **f))
        o = self._outs
        nz = None
        if noise is not None:
            nz = torch.as_tensor(noise, dtype=torch.float32).to(self._dev).contiguous()
            assert tuple(nz.shape) == (3, B, self.act_dim)
        s = self._stream()
        sp = C.c_void_p(s.cuda_stream)
        args = [C.c_void_p(t.data_ptr()) for t in (x, x2, a, r, d)]
        nzp = C.c_void_p(nz.data_ptr()) if nz is not None else None
        outs = [C.c_void_p(o[k].data_ptr()) for k in ("scalars", "q1", "q2", "logp_pi")]
        world = self._world()
        if world == 1 and split:       # same arithmetic through the data-parallel entry points
            N.check(self._lib.ddrl_sac_compute_grads(self._h, *args, B, nzp, self.seed, 1.0, *outs, sp))
            N.check(self._lib.ddrl_sac_apply_grads(self._h, B, sp))
        elif world == 1:
            N.check(self._lib.ddrl_sac_step(self._h, *args, B, nzp, self.seed, *outs, sp))
        elif self._fused:
            # gradients are exchanged over NVLink peer memory inside the optimiser kernel: one graph, no NCCL call
            N.check(self._lib.ddrl_sac_step_dp(self._h, *args, B, nzp, self._noise_seed(), 1.0 / world, *outs, sp))
        else:
            import torch.distributed as dist
            N.check(self._lib.ddrl_sac_compute_grads(self._h, *args, B, nzp, self._noise_seed(), 1.0 / world, *outs, sp))
            dist.all_reduce(self._grad, op=dist.ReduceOp.SUM, group=self._pg)
            if self.auto_alpha:
                dist.all_reduce(self._alpha_stat, op=dist.ReduceOp.SUM, group=self._pg)
                self._alpha_stat.div_(world)
            N.check(self._lib.ddrl_sac_apply_grads(self._h, B, sp))
        for t in (x, x2, a, r, d):
            t.record_stream(s)
        if nz is not None:
            nz.record_stream(s)
        self.steps += 1
        if sync_outputs:
            s.synchronize()
        return o

    def _train_host_block(self, batch, blk, sync_outputs, B=None):
        """train() on a HostBatch that still IS the pinned block sample_batch() returned: one native call
        (ddrl_sac_step_host) queues the H2D copy of the block, the update, and the D2H copy of the four scalars into a
        pinned host array — no torch op, no wait.  `losses()` reads the scalars back."""
        B = int(batch.n) if B is None else B
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch} (pass max_batch= to Learner)")
        if self._outs is None or self._outs["q1"].shape[0] != B:
            f = dict(dtype=torch.float32, device=self._dev)
            self._outs = dict(scalars=torch.zeros(4, **f), q1=torch.empty(B, **f), q2=torch.empty(B, **f),
                              logp_pi=torch.empty(B, **f))
        if self._hscal_ring is None:
            self._hscal_ring = torch.zeros((8, 4), dtype=torch.float32, pin_memory=True)
        self._hscal = self._hscal_ring[self.steps % 8]          # a deferred read (losses_async) may lag a few steps
        o = self._outs
        s = self._stream()
        world = self._world()
        N.check(self._lib.ddrl_sac_step_host(
            self._h, C.c_void_p(blk.data_ptr()), B, self._noise_seed() if world > 1 else self.seed,
            C.c_void_p(self._hscal.data_ptr()), *[C.c_void_p(o[k].data_ptr()) for k in ("q1", "q2", "logp_pi")],
            C.c_void_p(s.cuda_stream)))
        ev = torch.cuda.Event()
        ev.record(s)
        # the pinned block must stay allocated until its copy has run: keep it until a later step's event has fired
        self._inflight.append((ev, blk))
        while len(self._inflight) > 1 and self._inflight[0][0].query():
            self._inflight.popleft()
        self._hscal_event = ev
        self.steps += 1
        if sync_outputs:
            s.synchronize()
        out = dict(o)
        out["scalars"] = _LazyScalars(self)
        return out

    def train_via_host(self, replay_buffer, batch_size):
        """example/model.py:92-101 `Model.train(replay_buffer, args)` with the reference's data flow kept: the batch is
        sampled INTO HOST MEMORY (what `ray.get(replay_buffer.sample_batch.remote(B))` delivers) and fed back from host
        memory (feed_dict), but as two queued native calls with no wait in between — gather + D2H into a pinned block, then
        H2D of that block + the update + D2H of the four scalars.  Returns train()'s dict plus `batch`: the sampled host
        batch (numpy views of the pinned block), valid like the scalars once `losses()` has returned."""
        rb = getattr(replay_buffer, "local", replay_buffer)
        B = int(batch_size)
        if (getattr(rb, "index_source", None) != "philox" or getattr(rb, "device", None) != self.device
                or (self._world() > 1 and not self._fused) or getattr(rb, "_scalar_act", False)):
            return self.train(replay_buffer.sample_batch(B))
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch} (pass max_batch= to Learner)")
        with rb._lock:
            rb.flush()
            if rb.size == 0:
                raise ValueError("high <= 0")
            s = self._stream()
            key = (B, rb.obs_dim, rb.act_dim)
            if self._via_host_key != key:
                self._via_host_key = key
                self._via_host_nbytes = int(self._lib.ddrl_rb_sample_block_bytes(rb._h, B))
            nbytes = self._via_host_nbytes
            blk = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            N.check(self._lib.ddrl_rb_sample_host_async(rb._h, B, 1, None, rb._philox_seed(), rb._counter, rb._rng_stream,
                                                        blk.data_ptr(), nbytes, s.cuda_stream))
            rb._counter += 1
        # the update is queued right behind the gather; the numpy views of the block are built afterwards, off the GPU's
        # critical path
        out = self._train_host_block(None, blk, False, B=B)
        out["batch"] = rb._host_batch(blk, nbytes, B, (B,), (B, self.act_dim), False)
        return out

    def losses_async(self):
        """A handle on the losses of the last train(): a callable that waits for THAT step and returns its four scalars (on
        the host-block path each step gets its own pinned slot, so the handle stays valid while later steps are queued)."""
        ev, slot = self._hscal_event, self._hscal
        if ev is None:                      # the last train() took another path: snapshot its device scalars on the stream
            snap = self._outs["scalars"].clone()
            return lambda: snap.cpu().numpy()

        def result():
            ev.synchronize()
            return slot.numpy().copy()
        return result

    def losses(self):
        """(pi_loss, q1_loss, q2_loss, alpha) of the last train() that took the host-block path, as a numpy array: waits for
        that step only (an event), not for the stream."""
        if self._hscal_event is None:
            raise RuntimeError("losses(): no host-block train() call yet")
        self._hscal_event.synchronize()
        return self._hscal.numpy().copy()

    def train_from_buffer(self, replay_buffer, batch_size, noise=None, sync_outputs=False):
        """`batch = replay_buffer.sample_batch(B); agent.train(batch)` as ONE native call (example/model.py:92-101's
        Model.train(replay_buffer, args) shape): the step's first kernel gathers the batch straight from the ring with
        the buffer's own Philox index stream, so the rows are exactly those sample_batch(B) would have returned at this
        point and the batch never exists as separate arrays.  Falls back to sample_batch + train for buffers that
        draw indices with numpy, live on another device, or are sharded wrappers."""
        rb = getattr(replay_buffer, "local", replay_buffer)
        B = int(batch_size)
        world = self._world()
        if (getattr(rb, "index_source", None) != "philox" or getattr(rb, "device", None) != self.device
                or (world > 1 and not self._fused) or getattr(rb, "_scalar_act", False)):
            return self.train(replay_buffer.sample_batch(B, device=True), noise=noise, sync_outputs=sync_outputs)
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch} (pass max_batch= to Learner)")
        if rb.size == 0:
            raise ValueError("high <= 0")        # what the reference's np.random.randint(0, 0, ...) raises
        if self._outs is None or self._outs["q1"].shape[0] != B:
            f = dict(dtype=torch.float32, device=self._dev)
            self._outs = dict(scalars=torch.zeros(4, **f), q1=torch.empty(B, **f), q2=torch.empty(B, **f),
                              logp_pi=torch.empty(B, **f))
        o = self._outs
        nz = None
        if noise is not None:
            nz = torch.as_tensor(noise, dtype=torch.float32).to(self._dev).contiguous()
            assert tuple(nz.shape) == (3, B, self.act_dim)
        s = self._stream()
        outs = [C.c_void_p(o[k].data_ptr()) for k in ("scalars", "q1", "q2", "logp_pi")]
        N.check(self._lib.ddrl_sac_step_from_buffer(
            self._h, rb.native_handle, B, rb._philox_seed(), rb._counter, rb._rng_stream,
            C.c_void_p(nz.data_ptr()) if nz is not None else None, self._noise_seed(), *outs, C.c_void_p(s.cuda_stream)))
        rb._counter += 1
        if nz is not None:
            nz.record_stream(s)
        self.steps += 1
        if sync_outputs:
            s.synchronize()
        return o

    def state(self):
        t_pi, t_q, t_a, la = C.c_int(), C.c_int(), C.c_int(), C.c_float()
        N.check(self._lib.ddrl_sac_state(self._h, C.byref(t_pi), C.byref(t_q), C.byref(t_a), C.byref(la),
                                         C.c_void_p(self._stream().cuda_stream)))
        return dict(t_pi=t_pi.value, t_q=t_q.value, t_alpha=t_a.value, log_alpha=la.value)

    # reference stubs (actor_learner.py:144-148)
    def compute_gradients(self, x, y):
        pass

    def apply_gradients(self, gradients):
        pass


class _LazyScalars:
    """`train()`'s `scalars` entry on the host-block path: behaves like the 4-float tensor of the other paths for the two
    things callers do with it — `.cpu()` (waits for that step's event and returns a CPU tensor) and indexing."""

    def __init__(self, learner):
        self._l = learner

    def cpu(self):
        return torch.from_numpy(self._l.losses())

    def numpy(self):
        return self._l.losses()

    def __getitem__(self, i):
        return self.cpu()[i]

    def tolist(self):
        return self._l.losses().tolist()


class Actor(object):
    """The reference's rollout-side policy (algos/sac1/actor_learner.py:151-229): holds the main
    policy only, `set_weights` does NOT touch a target net, `get_action(o, deterministic)` returns
    one action.  Inference runs on the GPU (ddrl_sac_act); `get_actions` is the vectorised form for
    many environments per call."""

    def __init__(self, opt, job="worker", *, device=None):
        self._learner = Learner(opt, job, device=device, max_batch=1)
        self.opt = opt
        self.names = [n for n in self._learner.names if "/pi/" in n]
        self._calls = 0
        self._obs_stage = {}

    @classmethod
    def from_learner(cls, learner):
        """An Actor view of an existing Learner's main policy (example/model.py's Model both trains and acts)."""
        self = cls.__new__(cls)
        self._learner, self.opt = learner, learner.opt
        self.names = [n for n in learner.names if "/pi/" in n]
        self._calls, self._obs_stage = 0, {}
        return self

    def set_weights(self, variable_names, weights):
        # assign by name; the rollout policy has no target network to re-initialise
        L = self._learner
        flat = torch.from_numpy(L._flat_from(variable_names, weights)).to(L._dev)
        L.set_flat_weights(flat, also_target=False)

    def get_weights(self):
        keys, values = self._learner.get_weights()
        keep = [i for i, k in enumerate(keys) if "/pi/" in k]       # variables reachable from self.pi
        return [keys[i] for i in keep], [values[i] for i in keep]

    def get_actions(self, obs, deterministic=False, noise=None):
        L = self._learner
        o = torch.as_tensor(np.asarray(obs, dtype=np.float32)).reshape(-1, L.obs_dim).to(L._dev).contiguous() \
            if not (isinstance(obs, torch.Tensor) and obs.is_cuda) else obs.to(torch.float32).reshape(-1, L.obs_dim).contiguous()
        n = int(o.shape[0])
        out = torch.empty((n, L.act_dim), dtype=torch.float32, device=L._dev)
        nz = None
        if noise is not None:
            nz = torch.as_tensor(noise, dtype=torch.float32).to(L._dev).contiguous()
        s = L._stream()
        self._calls += 1
        N.check(L._lib.ddrl_sac_act(L._h, C.c_void_p(o.data_ptr()), n, 1 if deterministic else 0,
                                    C.c_void_p(nz.data_ptr()) if nz is not None else None, L.seed + 7919, self._calls,
                                    C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream)))
        o.record_stream(s)
        return out

    def get_action(self, o, deterministic=False):
        return self.get_actions(np.asarray(o).reshape(1, -1), deterministic)[0].cpu().numpy()

    def test(self, test_env, replay_buffer=None, n=25):
        rew = []
        for _ in range(n):
            o, d, ep_ret, ep_len = test_env.reset(), False, 0.0, 0
            while not (d or ep_len == _opt_get(self.opt, "max_ep_len", 1000)):
                o, r, d, _ = test_env.step(self.get_action(o, True))
                ep_ret += r
                ep_len += 1
            rew.append(ep_ret)
        return sum(rew) / n
