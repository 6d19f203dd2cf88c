"""GPU, >= 2 devices, NCCL: the data-parallel SAC1 step (compute_grads -> all-reduce -> apply_grads)
equals the single-GPU step on the concatenated batch, replicas stay bit-identical, and the
ParameterServer broadcast delivers the learner's flat weights to every rank."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.sac1_oracle import conditioned_params, make_batch

pytestmark = pytest.mark.gpu

D, A, HID, B = 24, 4, (128, 128), 256


def _opt(batch):
    space = SimpleNamespace(high=np.ones(A, np.float32), shape=(A,))
    return SimpleNamespace(obs_dim=D, act_dim=A, ac_kwargs=dict(hidden_sizes=HID, action_space=space), alpha=0.2,
                           gamma=0.99, lr=1e-3, polyak=0.995, seed=0, batch_size=batch)


def _worker(rank, world, port, ret, fused=False):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ddrl_b200 import Learner
    from ddrl_b200.dist import DistributedParameterServer
    params = conditioned_params(D, A, HID, seed=3)
    L = Learner(_opt(B // world), "learner", device=rank)
    L.set_weights(list(params), list(params.values()))
    if fused:
        assert L.connect_peers()
    for it in range(3):
        batch, noise = make_batch(D, A, B, seed=40 + it)
        lo, hi = rank * B // world, (rank + 1) * B // world
        L.train({k: v[lo:hi] for k, v in batch.items()}, noise=noise[:, lo:hi].copy())
    keys, values = L.get_weights()
    ps = DistributedParameterServer(keys, values, src=0, device=torch.device("cuda", rank))
    if rank == 0:
        ps.push_flat(L.get_flat_weights() * 0 + 3.0)
    ps.sync()
    ret[rank] = (L.get_flat_weights("main").cpu().numpy(), L.get_flat_weights("target").cpu().numpy(),
                 float(ps.pull_flat().min()), float(ps.pull_flat().max()), L.comm_error(), bool(L.nvls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["nccl", "peer-fused", "peer-fused-two-kernels", "peer-fused-nvls"])
def test_two_gpu_step_equals_single_gpu_on_concatenated_batch(mode, monkeypatch):
    """peer-fused (Learner.connect_peers, no NCCL call on the step path): ONE kernel sums the split-K partials into this
    rank's exchange slot, publishes a flag on every peer, waits for all flags, reads every peer's gradient over NVLink peer
    memory and applies Adam; peer-fused-nvls: the same kernel fetches the sum of all ranks' gradients with
    multimem.ld_reduce over an NVLS multicast mapping of the exchange buffers (the NVSwitch adds); peer-fused-two-kernels (DDRL_DP_V1=1): the same as a reduce + publish kernel followed by the
    optimiser kernel; nccl: torch.distributed.all_reduce between compute_grads and apply_grads."""
    import torch.multiprocessing as mp
    import __graft_entry__
    __graft_entry__.build()
    fused = mode != "nccl"
    if mode == "peer-fused-two-kernels":
        monkeypatch.setenv("DDRL_DP_V1", "1")       # inherited by the spawned ranks
    # peer-fused-nvls: multimem.ld_reduce over a multicast mapping (the default when the NVSwitch offers it; falls back to
    # IPC peer reads otherwise); "peer-fused" pins the IPC form
    monkeypatch.setenv("DDRL_DP_NVLS", "1" if mode == "peer-fused-nvls" else "0")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret, fused), nprocs=2, join=True)
    from ddrl_b200 import Learner
    params = conditioned_params(D, A, HID, seed=3)
    single = Learner(_opt(B), "learner", device=0)
    single.set_weights(list(params), list(params.values()))
    for it in range(3):
        batch, noise = make_batch(D, A, B, seed=40 + it)
        single.train(batch, noise=noise)
    want_m = single.get_flat_weights("main").cpu().numpy()
    want_t = single.get_flat_weights("target").cpu().numpy()
    assert np.array_equal(ret[0][0], ret[1][0]) and np.array_equal(ret[0][1], ret[1][1])
    for r in (0, 1):
        # same arithmetic up to the summation order of the batch reduction (2 x 128 rows vs 256 rows)
        assert np.abs(ret[r][0] - want_m).max() <= 2e-5 * np.abs(want_m).max()
        assert np.abs(ret[r][1] - want_t).max() <= 2e-5 * np.abs(want_t).max()
        assert ret[r][2] == 3.0 and ret[r][3] == 3.0
        assert ret[r][4] == 0            # no peer time-out seen by the fused kernel
    print(f"{mode}: NVLS multicast in use: {ret[0][5]}")
    assert ret[0][5] == ret[1][5] and (ret[0][5] is False or mode == "peer-fused-nvls")


def _global_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ddrl_b200.dist import ShardedReplayBuffer
    from oracle.make_golden import make_inputs
    Dd, Aa, total = 24, 4, 600                       # 300 rows per shard
    srb = ShardedReplayBuffer(Dd, Aa, total, device=rank, seed=9)
    n_local = 300 if rank == 0 else 220              # ragged fill: shard 1 is not full
    obs, act, rew, nxt, done = make_inputs(Dd, Aa, n_local, seed=50 + rank)
    srb.store_batch(obs, act, rew, nxt, done)
    srb.connect()
    dist.barrier()
    g = np.random.Generator(np.random.PCG64(3))
    idx = g.integers(0, 520, 257)                    # global indices over 300 + 220 stored rows
    out = srb.sample_batch(257, mode="global", idxs=idx, device=False)
    drawn = srb.sample_batch(64, mode="global", device=True, return_idxs=True)
    torch.cuda.synchronize()
    dist.barrier()
    ret[rank] = (out, srb.sizes, drawn["idxs"].cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_global_uniform_gather_reads_peer_shards_over_p2p():
    import torch.multiprocessing as mp
    import __graft_entry__
    __graft_entry__.build()
    from oracle.make_golden import make_inputs
    from oracle.replay_oracle import ReplayRingOracle
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_global_worker, args=(2, port, ret), nprocs=2, join=True)
    # oracle: the concatenation of both shards' stored rows (300 + 220)
    ora = ReplayRingOracle(24, 4, 520)
    for r, n in ((0, 300), (1, 220)):
        obs, act, rew, nxt, done = make_inputs(24, 4, n, seed=50 + r)
        ora.store_batch(obs, act, rew, nxt, done)
    g = np.random.Generator(np.random.PCG64(3))
    idx = g.integers(0, 520, 257)
    want = ora.sample_batch(257, idxs=idx)
    for r in (0, 1):
        out, sizes, drawn = ret[r]
        assert sizes == [300, 220]
        for k in want:
            assert np.array_equal(out[k].view(np.uint32), want[k].view(np.uint32)), (r, k)
        assert drawn.min() >= 0 and drawn.max() < 520


def _frames_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ddrl_b200.dist import ShardedFrameReplayBuffer
    from oracle.frames_oracle import FrameRingOracle
    from oracle.replay_oracle import philox_indices
    H, W, S, total, n = 16, 16, 4, 60, 45 + 10 * rank            # shard capacity 30: both shards wrap
    g = np.random.Generator(np.random.PCG64(70 + rank))
    frames = g.integers(0, 256, (n, H * W), dtype=np.uint8)
    act, rew, done = g.integers(0, 5, n).astype(np.float32), g.standard_normal(n).astype(np.float32), (g.random(n) < 0.1).astype(np.float32)
    srb = ShardedFrameReplayBuffer((H, W), S, total, mode="dedup", device=rank, seed=21)
    ora = FrameRingOracle(H * W, S, srb.map.cap)
    srb.store_frames(frames, act, rew, done)
    ora.store_frames(frames, act, rew, done)
    got = srb.sample_batch(64, return_idxs=True)
    idx = got["idxs"].cpu().numpy()
    want = ora.sample_batch(idx)
    ok = np.array_equal(idx, ora.drawn_indices(philox_indices(64, ora.size - S, 21, 0, rank)))      # rng_stream = rank
    ok = ok and all(np.array_equal(got[k].cpu().numpy().reshape(64, S, -1), want[k]) for k in ("obs1", "obs2"))
    ok = ok and all(np.array_equal(got[k].cpu().numpy(), want[k]) for k in ("acts", "rews", "done"))
    ret[rank] = (bool(ok), srb.get_counts(), srb.get_counts(global_=True), srb.map.cap)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_frame_replay_two_gpus():
    """C4 shape over 2 GPUs: one frame ring per rank, local Philox sampling on sub-stream = rank, bit-exact against the
    per-shard oracle; the only collective is the counter all-reduce of get_counts(global_=True)."""
    import torch.multiprocessing as mp
    import __graft_entry__
    __graft_entry__.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_frames_worker, args=(2, port, ret), nprocs=2, join=True)
    for r in (0, 1):
        ok, local, glob, cap = ret[r]
        assert ok and cap == 30
        assert local == (1, 45 + 10 * r, 30)
        assert glob == (2, 100, 60)
