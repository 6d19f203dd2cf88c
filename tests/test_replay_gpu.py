"""GPU parity: ddrl_b200.ReplayBuffer (CUDA, through the C ABI) vs the reference numpy ring —
golden vectors generated from the reference classes, the numpy oracle on seeded inputs, and
size-independent properties at BASELINE.json's full sizes.  Bar: BIT-EXACT (uint32 views)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle.make_golden import REPLAY_CASES, make_inputs
from oracle.replay_oracle import ReplayRingOracle, philox_indices

pytestmark = pytest.mark.gpu

KEYS = ("obs1", "obs2", "acts", "rews", "done")
FLAVOR = {"sac": "sac1", "dsac": "dsac", "sac1": "sac1"}


def bits(a):
    if isinstance(a, torch.Tensor):
        a = a.cpu().numpy()
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same(a, b):
    return np.array_equal(bits(a), bits(b))


@pytest.fixture(scope="module")
def RB():
    import __graft_entry__
    __graft_entry__.build()
    from ddrl_b200 import ReplayBuffer
    return ReplayBuffer


@pytest.mark.parametrize("name", sorted(REPLAY_CASES))
@pytest.mark.parametrize("device_out", [False, True])
def test_golden_vectors(RB, golden_dir, name, device_out):
    g = np.load(os.path.join(golden_dir, f"replay_{name}.npz"))
    variant = str(g["variant"])
    rb = RB(int(g["obs_dim"]), int(g["act_dim"]), int(g["capacity"]), flavor=FLAVOR[variant], stage_rows=16)
    for i in range(len(g["in_rew"])):
        rb.store(g["in_obs"][i], g["in_act"][i], float(g["in_rew"][i]), g["in_next"][i], bool(g["in_done"][i]))
    assert (rb.ptr, rb.size, rb.max_size) == (int(g["ptr"]), int(g["size"]), int(g["capacity"]))
    ring = rb.ring_arrays()
    for k in KEYS:
        assert same(ring[f"{k}_buf"], g[f"ring_{k}"]), k
    out = rb.sample_batch(len(g["idxs"]), idxs=g["idxs"], device=device_out)
    assert set(out) == set(KEYS)
    for k in KEYS:
        o = out[k].cpu().numpy() if device_out else out[k]
        assert o.dtype == np.float32 and o.flags["C_CONTIGUOUS"] and o.shape == g[f"out_{k}"].shape
        assert same(o, g[f"out_{k}"]), k
    if variant == "sac1":
        assert tuple(rb.get_counts()) == tuple(int(x) for x in g["counts"])
    elif variant == "dsac":
        assert rb.get_counts() == int(g["counts"][0])


@pytest.mark.parametrize("D,A,cap,n", [(8, 2, 1000, 2500), (24, 4, 333, 1000), (376, 17, 64, 200),
                                       (3, 1, 17, 40), (5, 3, 64, 10), (1, 1, 9, 30), (130, 6, 50, 75),
                                       (17, 6, 40, 100)])
def test_store_batch_and_sample_vs_oracle(RB, D, A, cap, n):
    obs, act, rew, nxt, done = make_inputs(D, A, n, seed=D * 1000 + A)
    rb, ora = RB(D, A, cap), ReplayRingOracle(D, A, cap)
    cuts = [0, n // 3, n // 3 + 1, n]          # ragged batches, including a 1-row batch
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        rb.store_batch(obs[lo:hi], act[lo:hi], rew[lo:hi], nxt[lo:hi], done[lo:hi])
        ora.store_batch(obs[lo:hi], act[lo:hi], rew[lo:hi], nxt[lo:hi], done[lo:hi])
    assert (rb.ptr, rb.size, rb.steps) == (ora.ptr, ora.size, ora.steps)
    ring = rb.ring_arrays()
    for k in KEYS:
        assert same(ring[f"{k}_buf"], getattr(ora, f"{k}_buf")), k
    g = np.random.Generator(np.random.PCG64(5))
    for B in (1, 31, 32, 33, 257):
        idxs = g.integers(0, ora.size, B)
        got, want = rb.sample_batch(B, idxs=idxs), ora.sample_batch(B, idxs=idxs)
        for k in KEYS:
            assert same(got[k], want[k]), (B, k)
    assert rb.get_counts() == ora.get_counts()


def test_store_batch_larger_than_capacity(RB):
    D, A, cap, n = 8, 2, 100, 350
    obs, act, rew, nxt, done = make_inputs(D, A, n, 3)
    rb, ora = RB(D, A, cap), ReplayRingOracle(D, A, cap)
    rb.store_batch(obs[:7], act[:7], rew[:7], nxt[:7], done[:7])
    ora.store_batch(obs[:7], act[:7], rew[:7], nxt[:7], done[:7])
    rb.store_batch(obs[7:], act[7:], rew[7:], nxt[7:], done[7:])        # 343 rows into 100 slots
    ora.store_batch(obs[7:], act[7:], rew[7:], nxt[7:], done[7:])
    assert (rb.ptr, rb.size, rb.steps) == (ora.ptr, ora.size, ora.steps)
    ring = rb.ring_arrays()
    for k in KEYS:
        assert same(ring[f"{k}_buf"], getattr(ora, f"{k}_buf")), k


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_device_inputs_cast_like_numpy(RB, dtype):
    D, A, cap, n = 24, 4, 300, 450
    obs, act, rew, nxt, done = make_inputs(D, A, n, 9)
    rb, ora = RB(D, A, cap), ReplayRingOracle(D, A, cap)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    host = [np.asarray(a).astype(npdt) for a in (obs, act, rew, nxt, done)]
    ora.store_batch(host[0], host[1], host[2], host[3], host[4])
    dev = [torch.from_numpy(a).cuda() for a in host]
    rb.store_batch(dev[0], dev[1], dev[2], dev[3], dev[4])
    ring = rb.ring_arrays()
    for k in KEYS:
        assert same(ring[f"{k}_buf"], getattr(ora, f"{k}_buf")), k


def test_by_value_capture(RB):
    rb, ora = RB(4, 2, 10), ReplayRingOracle(4, 2, 10)
    o = np.zeros(4)
    for i in range(5):
        o[:] = i                      # the caller mutates its observation buffer in place
        rb.store(o, np.ones(2) * i, i, o + 1, False)
        ora.store(o.copy(), np.ones(2) * i, i, o + 1, False)
    o[:] = -99
    assert same(rb.ring_arrays()["obs1_buf"], ora.obs1_buf)


def test_philox_stream_is_bit_exact_and_batches_match(RB):
    D, A, cap = 24, 4, 5000
    obs, act, rew, nxt, done = make_inputs(D, A, cap, 21)
    rb = RB(D, A, cap, seed=0xDEADBEEF12345, rng_stream=3)
    ora = ReplayRingOracle(D, A, cap)
    rb.store_batch(obs, act, rew, nxt, done)
    ora.store_batch(obs, act, rew, nxt, done)
    for call, B in enumerate((256, 1000, 33)):
        got = rb.sample_batch(B, device=True, return_idxs=True)
        want_idx = philox_indices(B, cap, 0xDEADBEEF12345, call, 3)
        assert np.array_equal(got["idxs"].cpu().numpy(), want_idx)
        want = ora.sample_batch(B, idxs=want_idx)
        for k in KEYS:
            assert same(got[k], want[k]), (call, k)


def test_numpy_index_source_reproduces_reference_stream(RB):
    D, A, cap = 8, 2, 400
    obs, act, rew, nxt, done = make_inputs(D, A, cap, 8)
    rb = RB(D, A, cap, index_source="numpy")
    ora = ReplayRingOracle(D, A, cap)
    rb.store_batch(obs, act, rew, nxt, done)
    ora.store_batch(obs, act, rew, nxt, done)
    np.random.seed(2024)
    want = [ora.sample_batch(64) for _ in range(3)]
    np.random.seed(2024)
    got = [rb.sample_batch(64) for _ in range(3)]
    for w, g_ in zip(want, got):
        for k in KEYS:
            assert same(g_[k], w[k])


def test_sample_many_equals_sequential_batches(RB):
    D, A, cap = 24, 4, 2000
    obs, act, rew, nxt, done = make_inputs(D, A, cap, 31)
    a, b = RB(D, A, cap, seed=77), RB(D, A, cap, seed=77)
    for rb in (a, b):
        rb.store_batch(obs, act, rew, nxt, done)
    many = a.sample_many(5, 128, return_idxs=True)
    assert many["obs1"].shape == (5, 128, D) and many["rews"].shape == (5, 128)
    idx = philox_indices(5 * 128, cap, 77, 0, 0).reshape(5, 128)
    assert np.array_equal(many["idxs"].cpu().numpy(), idx)
    ora = ReplayRingOracle(D, A, cap)
    ora.store_batch(obs, act, rew, nxt, done)
    for j in range(5):
        want = ora.sample_batch(128, idxs=idx[j])
        for k in KEYS:
            assert same(many[k][j], want[k])
    assert a.get_counts()[0] == 5


def test_empty_ring_raises_value_error(RB):
    rb = RB(8, 2, 16)
    with pytest.raises(ValueError):
        rb.sample_batch(4)
    from ddrl_b200 import _native as N
    out = [torch.empty(4 * 8, device="cuda") for _ in range(5)]
    rc = N.lib().ddrl_rb_sample(rb._h, 4, 1, None, 1, 0, 0, *[C.c_void_p(t.data_ptr()) for t in out], None, None)
    assert rc == N.EEMPTY


def test_dqn_flavor_scalar_action(RB):
    rb = RB(6, 1, 20, flavor="dqn")
    g = np.random.Generator(np.random.PCG64(0))
    for i in range(25):
        rb.store(g.standard_normal(6), int(g.integers(0, 3)), 1.0, g.standard_normal(6), False)
    out = rb.sample_batch(7)
    assert out["acts"].shape == (7,) and set(np.unique(out["acts"])) <= {0.0, 1.0, 2.0}
    assert rb.ring_arrays()["acts_buf"].shape == (20,)
    assert rb.get_counts() == (1, 25, 20)


def test_save_load_dqn_npy_format(RB, tmp_path):
    D, A, cap, n = 8, 2, 64, 100
    obs, act, rew, nxt, done = make_inputs(D, A, n, 13)
    rb = RB(D, A, cap)
    rb.store_batch(obs, act, rew, nxt, done)
    rb.sample_batch(4)
    rb.save(str(tmp_path), 2)
    infos = np.load(tmp_path / "buffer_infos-2.npy")
    assert list(infos) == [n % cap, cap, cap, n, 1]          # algos/dqn/train.py:88
    assert np.load(tmp_path / "obs1_buf-2.npy").shape == (cap, D)
    rb2 = RB(D, A, cap)
    rb2.load(str(tmp_path), 2)
    assert (rb2.ptr, rb2.size, rb2.steps, rb2.sample_times) == (rb.ptr, rb.size, rb.steps, rb.sample_times)
    r1, r2 = rb.ring_arrays(), rb2.ring_arrays()
    for k in r1:
        assert same(r1[k], r2[k])


# ---- full-size properties (BASELINE.json configs) -------------------------------------------------

@pytest.mark.parametrize("D,A,cap,B", [(8, 2, 1_000_000, 256), (24, 4, 1_000_000, 1024), (376, 17, 1_000_000, 4096)])
def test_full_size_round_trip_and_gather(RB, D, A, cap, B):
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(D)
    obs = torch.randn((cap, D), device=dev, generator=gen)
    nxt = torch.randn((cap, D), device=dev, generator=gen)
    act = torch.rand((cap, A), device=dev, generator=gen) * 2 - 1
    rew = torch.randn(cap, device=dev, generator=gen)
    done = (torch.rand(cap, device=dev, generator=gen) < 0.01).float()
    rb = RB(D, A, cap, seed=5)
    half = cap // 2 + 17
    rb.store_batch(obs[:half], act[:half], rew[:half], nxt[:half], done[:half])
    rb.store_batch(obs[half:], act[half:], rew[half:], nxt[half:], done[half:])
    assert (rb.ptr, rb.size) == (0, cap)
    out = rb.sample_many(8, B, return_idxs=True)
    idx = out["idxs"].reshape(-1)
    assert int(idx.min()) >= 0 and int(idx.max()) < cap
    # gather == torch fancy indexing of the inputs (bit-exact): encode -> sample -> compare
    assert torch.equal(out["obs1"].reshape(-1, D), obs[idx])
    assert torch.equal(out["obs2"].reshape(-1, D), nxt[idx])
    assert torch.equal(out["acts"].reshape(-1, A), act[idx])
    assert torch.equal(out["rews"].reshape(-1), rew[idx])
    assert torch.equal(out["done"].reshape(-1), done[idx])
    # index stream == oracle's Philox restatement at full size
    assert np.array_equal(idx.cpu().numpy(), philox_indices(8 * B, cap, 5, 0, 0))
    # identity gather of the whole ring returns the inputs (store -> export round trip)
    full = rb.sample_batch(cap, idxs=torch.arange(cap, device=dev), device=True)
    assert torch.equal(full["obs1"], obs) and torch.equal(full["obs2"], nxt) and torch.equal(full["acts"], act)
    assert torch.equal(full["rews"], rew) and torch.equal(full["done"], done)
    # overwrite the first rows again (wrap) and check FIFO replacement
    rb.store_batch(obs[-100:], act[-100:], rew[-100:], nxt[-100:], done[-100:])
    first = rb.sample_batch(100, idxs=torch.arange(100, device=dev), device=True)
    assert torch.equal(first["obs1"], obs[-100:]) and rb.ptr == 100


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("D,A,cap,B,nb", [(24, 4, 200_000, 1024, 32), (376, 17, 50_000, 4096, 2), (8, 2, 300_000, 256, 512)])
def test_every_gather_kernel_with_philox_indices(RB, monkeypatch, mode, D, A, cap, B, nb):
    """DDRL_GATHER_MODE picks the kernel family (0 auto, 1 bulk-async + register drain, 2 register kernels, 3 TMA-only for
    rows wider than 512 B): every one must return, for the Philox-drawn index stream, exactly the rows torch indexing does."""
    monkeypatch.setenv("DDRL_GATHER_MODE", str(mode))
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(100 + D)
    obs, nxt = torch.randn((cap, D), device=dev, generator=gen), torch.randn((cap, D), device=dev, generator=gen)
    act, rew = torch.rand((cap, A), device=dev, generator=gen), torch.randn(cap, device=dev, generator=gen)
    done = (torch.rand(cap, device=dev, generator=gen) < 0.05).float()
    rb = RB(D, A, cap, seed=77, rng_stream=3)
    rb.store_batch(obs, act, rew, nxt, done)
    for call in range(2):
        out = rb.sample_many(nb, B, return_idxs=True)
        idx = out["idxs"].reshape(-1)
        assert np.array_equal(idx.cpu().numpy(), philox_indices(nb * B, cap, 77, call, 3))
        assert torch.equal(out["obs1"].reshape(-1, D), obs[idx]) and torch.equal(out["obs2"].reshape(-1, D), nxt[idx])
        assert torch.equal(out["acts"].reshape(-1, A), act[idx])
        assert torch.equal(out["rews"].reshape(-1), rew[idx]) and torch.equal(out["done"].reshape(-1), done[idx])
    small = rb.sample_batch(B, device=True, return_idxs=True)            # one plain batch through the same family
    assert torch.equal(small["obs1"], obs[small["idxs"]]) and torch.equal(small["rews"], rew[small["idxs"]])


def test_c3_ring_at_the_configured_1e7_rows(RB):
    """BASELINE.json configs[2]: obs 376, act 17, replay 1e7 rows (30.9 GB ring), batch 4096.  Inputs are regenerated
    chunk by chunk from a counter-based recipe, so the check costs no second copy of the ring."""
    D, A, cap, B, chunk = 376, 17, 10_000_000, 4096, 500_000
    free, _ = torch.cuda.mem_get_info()
    if free < 48e9:
        pytest.skip("needs ~45 GB of free HBM")
    dev = torch.device("cuda")

    def make(lo, n):
        gen = torch.Generator(device=dev).manual_seed(7_000 + lo // chunk)
        return (torch.randn((n, D), device=dev, generator=gen), torch.rand((n, A), device=dev, generator=gen) * 2 - 1,
                torch.randn(n, device=dev, generator=gen), torch.randn((n, D), device=dev, generator=gen),
                (torch.rand(n, device=dev, generator=gen) < 0.01).float())

    rb = RB(D, A, cap, seed=9)
    for lo in range(0, cap, chunk):
        rb.store_batch(*make(lo, chunk))
    assert (rb.ptr, rb.size) == (0, cap) and rb.get_counts()[1] == cap
    out = rb.sample_many(16, B, return_idxs=True)
    idx = out["idxs"].reshape(-1)
    assert np.array_equal(idx.cpu().numpy(), philox_indices(16 * B, cap, 9, 0, 0))
    assert int(idx.max()) > 9_000_000                   # the draw really spans the 1e7 rows
    o1, o2, ac, rw, dn = (out[k].reshape(16 * B, -1) for k in KEYS)
    checked = 0
    for lo in range(0, cap, chunk):
        sel = ((idx >= lo) & (idx < lo + chunk)).nonzero().reshape(-1)
        if sel.numel() == 0:
            continue
        obs, act, rew, nxt, done = make(lo, chunk)
        j = idx[sel] - lo
        assert torch.equal(o1[sel], obs[j]) and torch.equal(o2[sel], nxt[j]) and torch.equal(ac[sel], act[j])
        assert torch.equal(rw[sel, 0], rew[j]) and torch.equal(dn[sel, 0], done[j])
        checked += int(sel.numel())
    assert checked == 16 * B


def test_concurrent_producers_and_learner_equal_some_serial_order(RB):
    """BASELINE config 5 / algos/sac1/sac1.py:195: many rollout workers store while the learner samples.  Four producer
    threads, each on its own CUDA stream, store tagged batches (host arrays and CUDA tensors) into one buffer while the main
    thread samples on a fifth stream.  Afterwards the ring must equal SOME serial order of the calls: every batch occupies
    consecutive slots in its own row order, batches of one producer appear in call order, counters add up; and every row
    a concurrent sample returned is a whole row some call stored (no torn rows)."""
    import threading
    D, A, cap, nprod, calls, nrows = 12, 3, 4096, 4, 60, 64
    rb = RB(D, A, cap, seed=1)
    dev = torch.device("cuda")

    def rows(key0, n):                      # row fields are all functions of ONE key: a torn row is detectable
        key = np.arange(key0, key0 + n, dtype=np.float32)
        return (np.repeat(key[:, None], D, 1), np.repeat(-key[:, None], A, 1), key * 2, np.repeat(key[:, None] + 0.5, D, 1),
                (key % 2).astype(np.float32))

    rb.store_batch(*rows(1_000_000, 256))             # something to sample from the start
    errors, stop = [], threading.Event()

    def producer(p):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for c in range(calls):
                    batch = rows(p * 100_000 + c * nrows, nrows)
                    if (p + c) % 2:
                        batch = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in batch]
                    rb.store_batch(*batch)
        except BaseException as e:          # noqa: BLE001
            errors.append(e)

    sampled = []
    threads = [threading.Thread(target=producer, args=(p,)) for p in range(nprod)]
    with torch.cuda.stream(torch.cuda.Stream()):
        for t in threads:
            t.start()
        while any(t.is_alive() for t in threads):
            sampled.append(rb.sample_batch(128, device=True))
        for t in threads:
            t.join()
    torch.cuda.synchronize()
    assert not errors, errors
    total = 256 + nprod * calls * nrows
    assert rb.get_counts()[1:] == (total, cap) and rb.ptr == total % cap
    for s in sampled:                       # no torn rows
        k = s["rews"] / 2
        assert torch.equal(s["obs1"], k[:, None].expand(-1, D)) and torch.equal(s["obs2"], (k + 0.5)[:, None].expand(-1, D))
        assert torch.equal(s["acts"], (-k)[:, None].expand(-1, A)) and torch.equal(s["done"], k % 2)
    ring = rb.ring_arrays()
    key = ring["rews_buf"] / 2
    assert np.array_equal(ring["obs1_buf"], np.repeat(key[:, None], D, 1)) and np.array_equal(ring["acts_buf"], np.repeat(-key[:, None], A, 1))
    # walk the ring from the oldest slot: whole batches, consecutive keys inside a batch, per-producer call order
    order = np.roll(key, -rb.ptr).astype(np.int64)
    last_call = {}
    i = 0
    while i < cap and order[i] >= 1_000_000:         # tail of the seed rows that survived (cap < total: they are gone)
        i += 1
    first = True
    while i < cap:
        p, c, j = order[i] // 100_000, (order[i] % 100_000) // nrows, (order[i] % 100_000) % nrows
        n = nrows - j
        assert first or j == 0, "a batch does not start at its first row"      # only the oldest batch may be cut by the wrap
        run = order[i:i + n]
        assert np.array_equal(run, order[i] + np.arange(len(run))), "rows of one batch are not consecutive"
        assert c > last_call.get(p, -1), "batches of one producer out of call order"
        last_call[p] = c
        i += len(run)
        first = False


def test_host_sample_into_pageable_and_pinned_blocks_agree(RB):
    """ddrl_rb_sample_host through the raw C ABI: a pinned block is written by the gather kernel itself (device mapping of
    the block), a pageable numpy block goes through device staging + cudaMemcpyAsync; both must hold the same bytes, and
    DDRL_ZERO_COPY=0 must not change them."""
    from ddrl_b200 import _native
    lib = _native.lib()
    D, A, cap, B = 24, 4, 5000, 1000
    g = np.random.Generator(np.random.PCG64(31))
    rows = [g.standard_normal((cap, D), dtype=np.float32), g.uniform(-1, 1, (cap, A)).astype(np.float32),
            g.standard_normal(cap, dtype=np.float32), g.standard_normal((cap, D), dtype=np.float32), (g.random(cap) < 0.1).astype(np.float32)]
    blocks = []
    for zero_copy in ("1", "0"):
        os.environ["DDRL_ZERO_COPY"] = zero_copy
        try:
            rb = RB(D, A, cap, seed=8)
        finally:
            os.environ.pop("DDRL_ZERO_COPY", None)
        rb.store_batch(*rows)
        n = int(lib.ddrl_rb_sample_block_bytes(rb.native_handle, B))
        pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        pageable = np.empty(n, np.uint8)
        s = torch.cuda.current_stream()
        for ptr in (pinned.data_ptr(), pageable.ctypes.data):
            _native.check(lib.ddrl_rb_sample_host(rb.native_handle, B, 1, None, 8, 5, 0, ptr, n, s.cuda_stream))
        assert np.array_equal(pinned.numpy(), pageable)
        blocks.append(pageable.copy())
        want = philox_indices(B, cap, 8, 5, 0)
        assert np.array_equal(pageable[n - 8 * B:].view(np.int64), want)
        assert np.array_equal(pageable[:B * D * 4].view(np.float32).reshape(B, D), rows[0][want])
    assert np.array_equal(blocks[0], blocks[1])
