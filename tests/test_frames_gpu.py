"""GPU parity for the Atari-shaped frame replay (BASELINE.json config 4, synthetic extension): bit-exact
against the numpy oracles; index stream of the dedup sampler against the Philox restatement."""
import numpy as np
import pytest
import torch

from oracle.frames_oracle import FrameRingOracle
from oracle.replay_oracle import ReplayRingOracle, philox_indices

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def FB():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200.frames import FrameReplayBuffer
    return FrameReplayBuffer


def test_naive_uint8_rows_bit_exact(FB):
    H, W, S, cap, n, B = 84, 84, 4, 48, 70, 37
    g = np.random.Generator(np.random.PCG64(1004))
    obs = g.integers(0, 256, (n, S, H, W), dtype=np.uint8)
    nxt = g.integers(0, 256, (n, S, H, W), dtype=np.uint8)
    act = g.integers(0, 6, n).astype(np.float32)
    rew, done = g.standard_normal(n).astype(np.float32), (g.random(n) < 0.1).astype(np.float32)
    fb = FB((H, W), S, cap, mode="naive")
    ora = ReplayRingOracle(S * H * W, 1, cap, flavor="dqn", obs_dtype=np.uint8)
    fb.store_batch(obs[:30], act[:30], rew[:30], nxt[:30], done[:30])
    fb.store_batch(obs[30:], act[30:], rew[30:], nxt[30:], done[30:])
    ora.store_batch(obs.reshape(n, -1), act, rew, nxt.reshape(n, -1), done)
    idx = g.integers(0, cap, B)
    got, want = fb.sample_batch(B, idxs=idx), ora.sample_batch(B, idxs=idx)
    assert got["obs1"].dtype == torch.uint8 and got["obs1"].shape == (B, S, H, W)
    assert np.array_equal(got["obs1"].cpu().numpy().reshape(B, -1), want["obs1"])
    assert np.array_equal(got["obs2"].cpu().numpy().reshape(B, -1), want["obs2"])
    for k in ("acts", "rews", "done"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    assert fb.get_counts() == (1, n, cap)


@pytest.mark.parametrize("H,W,S,cap,n", [(84, 84, 4, 64, 50), (84, 84, 4, 40, 100), (16, 16, 3, 20, 57)])
def test_dedup_stack_gather_bit_exact(FB, H, W, S, cap, n):
    g = np.random.Generator(np.random.PCG64(7))
    frames = g.integers(0, 256, (n, H * W), dtype=np.uint8)
    act = g.integers(0, 6, n).astype(np.float32)
    rew, done = g.standard_normal(n).astype(np.float32), (g.random(n) < 0.1).astype(np.float32)
    fb, ora = FB((H, W), S, cap, mode="dedup", seed=11), FrameRingOracle(H * W, S, cap)
    fb.store_frames(frames[:13], act[:13], rew[:13], done[:13])
    fb.store_frames(frames[13:], act[13:], rew[13:], done[13:])
    ora.store_frames(frames, act, rew, done)
    assert (fb.ptr, fb.size) == (ora.ptr, ora.size)
    wrapped = n > cap
    lo, hi = (0, cap) if wrapped else (S - 1, n - 1)        # a wrapped ring is valid everywhere but at the seam
    idx = g.integers(lo, hi, 65)
    got, want = fb.sample_batch(65, idxs=idx), ora.sample_batch(idx)
    assert got["obs1"].shape == (65, S, H, W)
    for k in ("obs1", "obs2"):
        assert np.array_equal(got[k].cpu().numpy().reshape(65, S, -1), want[k]), k
    for k in ("acts", "rews", "done"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    # overlapping stacks: obs2[:, :-1] is obs1[:, 1:]
    assert torch.equal(got["obs2"][:, :-1], got["obs1"][:, 1:])
    # Philox-drawn indices stay inside the valid window and follow the restated stream
    got = fb.sample_batch(256, return_idxs=True)
    drawn = got["idxs"].cpu().numpy()
    ages = philox_indices(256, fb.size - S, 11, 0, 0)
    assert np.array_equal(drawn, ora.drawn_indices(ages))
    # no drawn window touches the write head: frames i-S+1 .. i+1 are S+1 consecutive AGES, all older than the newest frame
    assert ages.min() >= 0 and ages.max() + S <= fb.size - 1
    want = ora.sample_batch(drawn)
    for k in ("obs1", "obs2"):
        assert np.array_equal(got[k].cpu().numpy().reshape(256, S, -1), want[k]), k
    for k in ("acts", "rews", "done"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


def test_dedup_store_device_tensors_and_overflow(FB):
    """store_frames from CUDA tensors, more frames than slots in one call (only the last `capacity` survive)."""
    H, W, S, cap, n = 16, 16, 3, 24, 61
    g = np.random.Generator(np.random.PCG64(9))
    frames = g.integers(0, 256, (n, H * W), dtype=np.uint8)
    act, rew, done = g.integers(0, 4, n).astype(np.float32), g.standard_normal(n).astype(np.float32), (g.random(n) < 0.2).astype(np.float32)
    fb, ora = FB((H, W), S, cap, mode="dedup", seed=3), FrameRingOracle(H * W, S, cap)
    dev = torch.device("cuda")
    fb.store_frames(torch.from_numpy(frames[:5]).to(dev), torch.from_numpy(act[:5]).to(dev), torch.from_numpy(rew[:5]).to(dev),
                    torch.from_numpy(done[:5]).to(dev))
    fb.store_frames(torch.from_numpy(frames[5:]).to(dev), torch.from_numpy(act[5:]).to(dev), torch.from_numpy(rew[5:]).to(dev),
                    torch.from_numpy(done[5:]).to(dev))
    ora.store_frames(frames, act, rew, done)
    assert (fb.ptr, fb.size) == (ora.ptr, ora.size)
    assert np.array_equal(fb.frames.cpu().numpy(), ora.frames)
    for k in ("act", "rew", "done"):
        assert np.array_equal(getattr(fb, k).cpu().numpy(), getattr(ora, k)), k


def test_dedup_needs_a_full_stack(FB):
    fb = FB((16, 16), 4, 32, mode="dedup")
    fb.store_frames(np.zeros((3, 256), np.uint8), np.zeros(3), np.zeros(3), np.zeros(3))
    with pytest.raises(ValueError):
        fb.sample_batch(4)
