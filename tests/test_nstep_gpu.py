"""GPU parity of the N-step sequence ring (ddrl_b200.NStepReplayBuffer -> ddrl_seg_sample) against the numpy
oracle (oracle/nstep_oracle.py, itself pinned to the reference class): bit-exact float32 for every output."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.nstep_oracle import NStepRingOracle, make_sequences, pack_rows
from oracle.replay_oracle import philox_indices

pytestmark = pytest.mark.gpu
KEYS = ("obs", "acts", "rews", "done")


@pytest.fixture(scope="module")
def NB():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import NStepReplayBuffer
    return NStepReplayBuffer


def same(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return a.dtype == np.float32 and a.shape == b.shape and np.array_equal(a.view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.parametrize("obs_shape,act_shape,Ln,cap,n,B", [
    ((115,), (), 8, 37, 50, 64),            # the reference's own float layout, scalar actions, wraps
    ((24,), (4,), 8, 500, 321, 256),        # N3 of SURVEY 8f: D=24, A=4, Ln=8 -> 1056-byte rows, partially filled
    ((5,), (3,), 3, 16, 40, 33),            # odd widths: segments not 16-byte aligned
    ((376,), (17,), 4, 64, 64, 128),        # Humanoid-shaped, exactly full
])
def test_sample_matches_oracle_bit_exact(NB, obs_shape, act_shape, Ln, cap, n, B):
    opt = SimpleNamespace(Ln=Ln, obs_shape=obs_shape, act_shape=act_shape, buffer_size=cap, batch_size=B, num_buffers=2)
    rb, ora = NB(opt, seed=5), NStepRingOracle(opt)
    for oq, aq in make_sequences(opt, n, 9):
        rb.store(oq, aq, 0)
        ora.store(oq, aq, 0)
    assert rb.get_counts() == ora.get_counts()
    idx = np.random.Generator(np.random.PCG64(1)).integers(0, ora.size, B)
    want = ora.sample_batch(idxs=idx)
    for device in (True, False):
        got = rb.sample_batch(idxs=idx, device=device)
        for k in KEYS:
            assert same(got[k], want[k]), (k, device)
        ora.sample_times += opt.num_buffers if device else 0      # two samples on the GPU side, one on the oracle + this
    assert rb.get_counts() == ora.get_counts()


def test_store_batch_equals_stores_and_philox_stream(NB):
    opt = SimpleNamespace(Ln=8, obs_shape=(24,), act_shape=(4,), buffer_size=300, batch_size=128, num_buffers=1)
    seqs = make_sequences(opt, 450, 3)                                  # 1.5 x capacity through store_batch in two calls
    obs = np.stack([np.stack([o[0] for o in oq]) for oq, _ in seqs])
    act = np.stack([np.stack([a for a, _, _ in aq]) for _, aq in seqs])
    rew = np.array([[r for _, r, _ in aq] for _, aq in seqs], dtype=np.float64)     # float64 input: cast like numpy assignment
    done = np.array([[d for _, _, d in aq] for _, aq in seqs])
    rb, ora = NB(opt, seed=0xABCDEF, rng_stream=2), NStepRingOracle(opt)
    rb.store_batch(obs[:200], act[:200], rew[:200], done[:200])
    rb.store_batch(obs[200:], act[200:], rew[200:], done[200:])
    for oq, aq in seqs:
        ora.store(oq, aq)
    assert (rb.ptr, rb.size) == (ora.ptr, ora.size)
    for call in range(3):
        got = rb.sample_batch(device=True, return_idxs=True)
        want_idx = philox_indices(128, 300, 0xABCDEF, call, 2)
        assert np.array_equal(got["idxs"].cpu().numpy(), want_idx)
        want = ora.sample_batch(idxs=want_idx)
        for k in KEYS:
            assert same(got[k], want[k]), (call, k)


def test_store_kernel_ring_rows_device_inputs_overflow_and_prefetch(NB):
    """seg_store_rows: ring rows bit-equal to the expected packed rows for host AND CUDA-tensor inputs, wrap-around, a single
    call with more rows than slots (only the last `capacity` survive, in order); prefetch() hands out sample_batch() in order."""
    opt = SimpleNamespace(Ln=3, obs_shape=(5,), act_shape=(3,), buffer_size=16, batch_size=8, num_buffers=1)
    seqs = make_sequences(opt, 70, 4)
    obs = np.stack([np.stack([o[0] for o in oq]) for oq, _ in seqs]).astype(np.float32)
    act = np.stack([np.stack([a for a, _, _ in aq]) for _, aq in seqs]).astype(np.float32)
    rew = np.array([[r for _, r, _ in aq] for _, aq in seqs], dtype=np.float32)
    done = np.array([[d for _, _, d in aq] for _, aq in seqs], dtype=np.float32)
    rows = pack_rows(obs, act, rew, done, 3, 5, 3)
    dev = torch.device("cuda")
    rb = NB(opt, seed=1)
    rb.store_batch(obs[:10], act[:10], rew[:10], done[:10])                                        # host arrays
    rb.store_batch(*[torch.from_numpy(x[10:21]).to(dev) for x in (obs, act, rew, done)])           # CUDA tensors, wraps
    want = np.zeros((16, rows.shape[1]), np.float32)
    for i in range(21):
        want[i % 16] = rows[i]
    assert (rb.ptr, rb.size, rb.steps) == (21 % 16, 16, 21)
    assert np.array_equal(rb.ring.cpu().numpy().view(np.uint32), want.view(np.uint32))
    rb.store_batch(obs[21:], act[21:], rew[21:], done[21:])                                        # 49 rows into 16 slots
    for i in range(21, 70):
        want[i % 16] = rows[i]
    assert (rb.ptr, rb.size, rb.steps) == (70 % 16, 16, 70)
    assert np.array_equal(rb.ring.cpu().numpy().view(np.uint32), want.view(np.uint32))
    ora_idx = [philox_indices(8, 16, 1, c, 0) for c in range(5)]
    it = rb.prefetch(depth=3, batch_size=8)
    for c in range(5):
        got = next(it)
        assert np.array_equal(got["rews"].cpu().numpy(), want[ora_idx[c]][:, 29:32]), c


def test_empty_raises_like_the_reference(NB):
    opt = SimpleNamespace(Ln=2, obs_shape=(4,), act_shape=(1,), buffer_size=8, batch_size=4, num_buffers=1)
    with pytest.raises(ValueError, match="high <= 0"):
        NB(opt).sample_batch()
