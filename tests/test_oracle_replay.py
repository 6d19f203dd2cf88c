"""CPU: pin the numpy oracle (oracle/replay_oracle.py) against the reference's own classes —
live where /root/reference is mounted, and through the committed golden vectors everywhere."""
import glob
import os

import numpy as np
import pytest

from oracle import ref_extract
from oracle.make_golden import REPLAY_CASES, make_inputs
from oracle.replay_oracle import (ParameterServerOracle, ReplayRingOracle, philox4x32_10,
                                  philox_indices, philox_normals)

FLAVOR = {"sac": "sac1", "dsac": "dsac", "sac1": "sac1"}


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def replay_golden_files(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "replay_*.npz")))


def drive_oracle(g):
    variant = str(g["variant"])
    buf = ReplayRingOracle(int(g["obs_dim"]), int(g["act_dim"]), int(g["capacity"]), flavor=FLAVOR[variant])
    n = len(g["in_rew"])
    for i in range(n):
        buf.store(g["in_obs"][i], g["in_act"][i], float(g["in_rew"][i]), g["in_next"][i], bool(g["in_done"][i]))
    return buf


@pytest.mark.parametrize("name", sorted(REPLAY_CASES))
def test_oracle_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"replay_{name}.npz"))
    buf = drive_oracle(g)
    assert (buf.ptr, buf.size) == (int(g["ptr"]), int(g["size"]))
    for k in ("obs1", "obs2", "acts", "rews", "done"):
        ring = getattr(buf, f"{k}_buf")
        assert ring.dtype == np.float32
        assert np.array_equal(bits(ring), bits(g[f"ring_{k}"])), k
    out = buf.sample_batch(len(g["idxs"]), idxs=g["idxs"])
    for k in ("obs1", "obs2", "acts", "rews", "done"):
        assert out[k].dtype == np.float32 and out[k].flags["C_CONTIGUOUS"]
        assert np.array_equal(bits(out[k]), bits(g[f"out_{k}"])), k
    if str(g["variant"]) == "sac1":
        assert tuple(np.atleast_1d(buf.get_counts())) == tuple(g["counts"])
    elif str(g["variant"]) == "dsac":
        # the golden counts were taken after the reference's sample_batch: dsac counts stores only
        assert buf.get_counts() == int(g["counts"][0])


def test_all_golden_files_are_covered(golden_dir):
    names = {os.path.basename(f)[len("replay_"):-4] for f in replay_golden_files(golden_dir)}
    assert names == set(REPLAY_CASES)


def test_store_batch_equals_sequential_stores():
    obs, act, rew, nxt, done = make_inputs(6, 3, 45, 77)
    a = ReplayRingOracle(6, 3, 20)
    b = ReplayRingOracle(6, 3, 20)
    for i in range(45):
        a.store(obs[i], act[i], rew[i], nxt[i], done[i])
    b.store_batch(obs[:30], act[:30], rew[:30], nxt[:30], done[:30])
    b.store_batch(obs[30:], act[30:], rew[30:], nxt[30:], done[30:])
    assert (a.ptr, a.size, a.steps) == (b.ptr, b.size, b.steps) == (5, 20, 45)
    for k in ("obs1", "obs2", "acts", "rews", "done"):
        assert np.array_equal(getattr(a, k + "_buf"), getattr(b, k + "_buf"))


def test_empty_ring_raises_like_reference():
    buf = ReplayRingOracle(4, 2, 10)
    with pytest.raises(ValueError):
        buf.sample_batch(8)


def test_default_index_draw_is_numpy_global_state():
    buf = ReplayRingOracle(4, 2, 10)
    obs, act, rew, nxt, done = make_inputs(4, 2, 10, 5)
    buf.store_batch(obs, act, rew, nxt, done)
    np.random.seed(123)
    expect = np.random.randint(0, 10, size=16)
    np.random.seed(123)
    out = buf.sample_batch(16)
    assert np.array_equal(out["rews"], buf.rews_buf[expect])


@pytest.mark.skipif(not ref_extract.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("variant", ["sac", "dsac", "sac1"])
def test_oracle_matches_live_reference(variant):
    cls = ref_extract.reference_replay(variant)
    D, A, cap, n = 11, 3, 37, 120
    ref = cls(D, A, cap)
    ora = ReplayRingOracle(D, A, cap, flavor=FLAVOR[variant])
    obs, act, rew, nxt, done = make_inputs(D, A, n, 4242)
    for i in range(n):
        ref.store(obs[i], act[i], float(rew[i]), nxt[i], bool(done[i]))
        ora.store(obs[i], act[i], float(rew[i]), nxt[i], bool(done[i]))
    assert (ref.ptr, ref.size, ref.max_size) == (ora.ptr, ora.size, ora.max_size)
    for k in ("obs1", "obs2", "acts", "rews", "done"):
        assert np.array_equal(bits(getattr(ref, k + "_buf")), bits(getattr(ora, k + "_buf")))
    # same global numpy state -> same index stream -> same batches
    np.random.seed(99)
    r = ref.sample_batch(50)
    np.random.seed(99)
    o = ora.sample_batch(50)
    for k in r:
        assert np.array_equal(bits(r[k]), bits(o[k]))
    if hasattr(ref, "get_counts"):
        assert np.array_equal(np.atleast_1d(ref.get_counts()), np.atleast_1d(ora.get_counts()))


def test_make_golden_is_reproducible(golden_dir):
    if not ref_extract.reference_available():
        pytest.skip("reference tree not mounted")
    from oracle.make_golden import run_reference
    for name, case in REPLAY_CASES.items():
        fresh = run_reference(*case)
        g = np.load(os.path.join(golden_dir, f"replay_{name}.npz"))
        for k in ("ring_obs1", "ring_acts", "out_obs2", "out_done", "idxs"):
            assert np.array_equal(fresh[k], g[k]), (name, k)


# ---- Philox ------------------------------------------------------------------------------------

def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = tuple(int(x) for x in philox4x32_10(*ctr, *key))
        assert got == want


def test_philox_indices_uniform_and_in_range():
    n, size = 200_000, 1000
    idx = philox_indices(n, size, seed=0x1234_5678_9ABC, counter=3, stream=1)
    assert idx.dtype == np.int64 and idx.min() >= 0 and idx.max() < size
    counts = np.bincount(idx, minlength=size)
    chi2 = ((counts - n / size) ** 2 / (n / size)).sum()
    assert 800 < chi2 < 1220          # 999 dof: mean 999, sd 44.7 -> +-4.5 sd
    # with replacement: duplicates must occur
    assert len(np.unique(idx[:2000])) < 2000
    # distinct (counter, stream, seed) give distinct streams; same arguments repeat
    assert np.array_equal(idx[:100], philox_indices(100, size, 0x1234_5678_9ABC, 3, 1))
    assert not np.array_equal(idx[:100], philox_indices(100, size, 0x1234_5678_9ABC, 4, 1))
    assert not np.array_equal(idx[:100], philox_indices(100, size, 0x1234_5678_9ABC, 3, 2))
    with pytest.raises(ValueError):
        philox_indices(4, 0, 1, 1)
    assert np.all(philox_indices(50, 1, 5, 5) == 0)
    big = philox_indices(1000, 10_000_000, 7, 0)
    assert big.max() < 10_000_000 and big.max() > 9_000_000


def test_philox_normals_moments():
    z = philox_normals(400_000, seed=11, counter=2)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3) < 0.1


# ---- ParameterServer -----------------------------------------------------------------------------

def test_ps_oracle_matches_golden(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "ps_sac1.npz"))
    keys = [str(k) for k in g["keys"]]
    init = [g["init0"].copy(), g["init1"].copy(), g["init2"].copy()]
    ps = ParameterServerOracle(keys, init)
    init[0][:] = 777.0                      # must not alias
    push = [g["push0"].copy(), g["push2"].copy()]
    ps.push([keys[0], keys[2]], push)
    push[0][:] = -1.0
    got = ps.pull([keys[2], keys[1], keys[0]])
    assert np.array_equal(got[0], g["pull_k2"]) and np.array_equal(got[1], g["pull_k1"]) and \
        np.array_equal(got[2], g["pull_k0"])
    ps.save_weights(str(tmp_path / "x_"))
    ps2 = ParameterServerOracle([], [], weights_file=str(tmp_path / "x_weights.pickle"))
    assert set(ps2.get_weights()) == set(keys)
    assert np.array_equal(ps2.pull([keys[0]])[0], g["pull_k0"])
