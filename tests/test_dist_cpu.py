"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path.
  * data-parallel algebra: per-rank gradients on half batches, averaged by all-reduce, give the
    same update as one step on the concatenated batch (the oracle is the arithmetic here; on the
    GPU box the same all-reduce sits between ddrl_sac_compute_grads and ddrl_sac_apply_grads);
  * DistributedParameterServer push/sync/pull == the reference ParameterServer semantics;
  * ShardMap index arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddrl_b200.dist import DistributedParameterServer, ShardMap, allreduce_mean_
from oracle.sac1_oracle import SAC1Oracle, conditioned_params, make_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run_world(fn, world=2):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


D, A, HID, B = 6, 2, (16, 16), 32


def _dp_step(rank, world):
    torch.set_num_threads(1)
    params = conditioned_params(D, A, HID, seed=3)
    o = SAC1Oracle(D, A, hidden=HID, params=params)
    batch, noise = make_batch(D, A, B, seed=4)
    lo, hi = rank * B // world, (rank + 1) * B // world
    local = {k: v[lo:hi] for k, v in batch.items()}
    lnoise = noise[:, lo:hi]

    def avg(g_pi, g_q):
        return [allreduce_mean_(g.clone()) for g in g_pi], [allreduce_mean_(g.clone()) for g in g_q]

    o.step(local, lnoise, grad_transform=avg)
    return o.flat("main"), o.flat("target")


def test_two_rank_step_equals_concatenated_batch():
    got = run_world(_dp_step, 2)
    params = conditioned_params(D, A, HID, seed=3)
    single = SAC1Oracle(D, A, hidden=HID, params=params)
    batch, noise = make_batch(D, A, B, seed=4)
    single.step(batch, noise)
    for main, target in got:
        assert np.allclose(main, single.flat("main"), rtol=0, atol=1e-12)
        assert np.allclose(target, single.flat("target"), rtol=0, atol=1e-12)
    assert np.array_equal(got[0][0], got[1][0])          # replicas stay bit-identical


def _ps(rank, world):
    keys = ["main/pi/dense/kernel", "main/pi/dense/bias"]
    init = [np.full((3, 2), float(rank + 1), np.float32), np.full(2, 10.0 * (rank + 1), np.float32)]
    ps = DistributedParameterServer(keys, init, src=0)
    first = ps.pull(keys)                                   # every rank sees rank 0's initial values
    new = [np.full((3, 2), 7.0 + rank, np.float32)]
    ps.push([keys[0]], new)                                 # rank 1's push is local until ...
    new[0][:] = -1                                          # (values are copied on push)
    ps.sync()                                               # ... the collective: source rank wins
    second = ps.pull([keys[1], keys[0]])
    return [a.copy() for a in first], [a.copy() for a in second], ps.version


def test_distributed_parameter_server():
    out = run_world(_ps, 2)
    for first, second, version in out:
        assert np.all(first[0] == 1.0) and np.all(first[1] == 10.0)
        assert np.all(second[0] == 10.0) and np.all(second[1] == 7.0) and second[1].shape == (3, 2)
        assert version == 2


def test_shard_map():
    m = ShardMap(1_000_000, 8)
    assert m.cap == 125_000 and m.total == 1_000_000
    s, r = m.locate([0, 124_999, 125_000, 999_999])
    assert list(s) == [0, 0, 1, 7] and list(r) == [0, 124_999, 0, 124_999]
    assert list(m.global_index(s, r)) == [0, 124_999, 125_000, 999_999]
    m2 = ShardMap(10, 3)
    assert m2.cap == 4 and m2.total == 12
    np.random.seed(0)
    picks = {m.pick_shard() for _ in range(200)}
    assert picks == set(range(8))
