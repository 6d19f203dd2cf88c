"""N-step sequence ring: the numpy oracle is pinned to the reference's own class
(algos/sac1/sac_ray.py:34-83, loaded by ast from /root/reference when present) and to the committed golden
vectors generated from that class."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import ref_extract
from oracle.nstep_oracle import NStepRingOracle, make_sequences

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "nstep_scalar_act.npz")


def ref_opt():
    # the float32 layout of the reference is selected by obs_shape == (115,) (sac_ray.py:42); scalar actions keep
    # the reference's np.stack(a_r_d_queue) legal under numpy >= 1.24
    return SimpleNamespace(Ln=8, obs_shape=(115,), act_shape=(), buffer_size=37, batch_size=64, num_buffers=3)


def drive(buf, opt, n, seed):
    for oq, aq in make_sequences(opt, n, seed):
        buf.store(oq, aq, 0)
    g = np.random.Generator(np.random.PCG64(seed + 1))
    idx = g.integers(0, min(n, opt.buffer_size), opt.batch_size)
    return idx


@pytest.mark.skipif(not ref_extract.reference_available(), reason="/root/reference not present")
def test_oracle_matches_live_reference_class():
    Ref = ref_extract.load_reference_class("algos/sac1/sac_ray.py", "ReplayBuffer")
    opt = ref_opt()
    ref, ora = Ref(opt), NStepRingOracle(opt)
    idx = drive(ref, opt, 50, 5)            # 50 > 37: wraps
    drive(ora, opt, 50, 5)
    for name in ("buffer_o", "buffer_a", "buffer_r", "buffer_d"):
        assert np.array_equal(getattr(ref, name), getattr(ora, name)), name
    assert (ref.ptr, ref.size, ref.steps) == (ora.ptr, ora.size, ora.steps)
    saved = np.random.randint
    try:
        np.random.randint = lambda lo, hi, size: idx        # inject the index stream into the reference's draw
        want = ref.sample_batch()
    finally:
        np.random.randint = saved
    got = ora.sample_batch(idxs=idx)
    for k in ("obs", "acts", "rews", "done"):
        assert np.array_equal(want[k], got[k]) and want[k].dtype == got[k].dtype == np.float32, k
    assert ref.get_counts() == ora.get_counts() == (3, 150, 37)


def test_oracle_matches_golden():
    z = np.load(GOLDEN)
    opt = ref_opt()
    ora = NStepRingOracle(opt)
    idx = drive(ora, opt, 50, 5)
    assert np.array_equal(idx, z["idx"])
    got = ora.sample_batch(idxs=idx)
    for k in ("obs", "acts", "rews", "done"):
        assert np.array_equal(got[k], z[k]), k
    assert tuple(z["counts"]) == ora.get_counts()


def test_vector_actions_and_empty():
    opt = SimpleNamespace(Ln=3, obs_shape=(5,), act_shape=(2,), buffer_size=4, batch_size=6, num_buffers=1)
    ora = NStepRingOracle(opt)
    with pytest.raises(ValueError):
        np.random.seed(0)
        ora.sample_batch()                   # np.random.randint(0, 0, ...) -> "high <= 0"
    seqs = make_sequences(opt, 6, 1)
    for oq, aq in seqs:
        ora.store(oq, aq)
    assert ora.buffer_a.shape == (4, 3, 2) and ora.size == 4 and ora.ptr == 2
    assert np.array_equal(ora.buffer_o[1], np.stack([x[0] for x in seqs[5][0]]))    # slot 1 holds the 6th sequence


def test_packed_row_layout_round_trips_the_oracle_arrays():
    """Row layout of ddrl_b200.NStepReplayBuffer (no GPU): the packed row [obs | acts | rews | done] the GPU tests expect in
    the ring holds exactly the oracle's four arrays at the segment offsets ddrl_seg_store / ddrl_seg_sample are given,
    float64 / bool inputs cast like numpy assignment."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "distributed-drl_b200"))
    from ddrl_b200.nstep import row_layout
    from oracle.nstep_oracle import pack_rows
    opt = SimpleNamespace(Ln=3, obs_shape=(5,), act_shape=(2,), buffer_size=8, batch_size=4, num_buffers=1)
    ora = NStepRingOracle(opt)
    seqs = make_sequences(opt, 6, 2)
    for oq, aq in seqs:
        ora.store(oq, aq)
    widths, off, used, row_f = row_layout(3, 5, 2)
    assert (widths, off, used, row_f) == ([20, 6, 3, 3], [0, 20, 26, 29], 32, 32)
    obs = np.stack([np.stack([o[0] for o in oq]) for oq, _ in seqs]).astype(np.float64)
    act = np.stack([np.stack([a for a, _, _ in aq]) for _, aq in seqs])
    rew = np.array([[r for _, r, _ in aq] for _, aq in seqs], dtype=np.float64)
    done = np.array([[d for _, _, d in aq] for _, aq in seqs])
    rows = pack_rows(obs, act, rew, done, 3, 5, 2)
    assert rows.dtype == np.float32 and rows.shape == (6, 32)
    assert np.array_equal(rows[:, 0:20].reshape(6, 4, 5), ora.buffer_o[:6])
    assert np.array_equal(rows[:, 20:26].reshape(6, 3, 2), ora.buffer_a[:6])
    assert np.array_equal(rows[:, 26:29], ora.buffer_r[:6]) and np.array_equal(rows[:, 29:32], ora.buffer_d[:6])
