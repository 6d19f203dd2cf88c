"""CPU: the in-process ray stand-in keeps the semantics the reference relies on."""
import threading
import time

import numpy as np
import pytest

from ddrl_b200 import ray_shim as ray
from ddrl_b200.ps import ParameterServer


def test_actor_calls_are_serialised_and_by_value():
    @ray.remote
    class Counter:
        def __init__(self):
            self.v, self.inside, self.max_inside, self.rows = 0, 0, 0, []

        def bump(self, arr):
            self.inside += 1
            self.max_inside = max(self.max_inside, self.inside)
            time.sleep(0.001)
            self.v += 1
            self.rows.append(arr)
            self.inside -= 1
            return self.v

        def state(self):
            return self.v, self.max_inside, self.rows

    c = Counter.remote()
    buf = np.zeros(2)

    @ray.remote
    def hammer(handle, k):
        for i in range(20):
            buf[:] = k
            handle.bump.remote(buf)
        return k

    refs = [hammer.remote(c, k) for k in range(4)]
    assert sorted(ray.get(refs)) == [0, 1, 2, 3]
    v, max_inside, rows = ray.get(c.state.remote())
    assert v == 80 and max_inside == 1
    buf[:] = -5
    assert all(r[0] >= 0 for r in rows)          # captured by value at call time


def test_get_wait_put_and_exceptions():
    @ray.remote(num_gpus=1, max_calls=1)
    def slow(x):
        time.sleep(0.05)
        return x * 2

    @ray.remote
    def boom():
        raise KeyError("nope")

    r = slow.remote(21)
    ready, pending = ray.wait([r], timeout=0.0)
    assert ready == [] and pending == [r]
    assert ray.get(r) == 42
    with pytest.raises(KeyError):
        ray.get(boom.remote())
    assert ray.get(ray.put({"a": 1})) == {"a": 1}
    ray.init(resources={"node0": 256})


def test_parameter_server_as_actor():
    PS = ray.remote(ParameterServer)
    ps = PS.remote(["k"], [np.ones(3, np.float32)])
    ps.push.remote(["k"], [np.full(3, 2.0, np.float32)])
    assert np.all(ray.get(ps.pull.remote(["k"]))[0] == 2.0)
    ps2 = PS._remote(args=[["k"], [np.zeros(1, np.float32)]], resources={"node0": 1})
    assert ray.get(ps2.get_weights.remote())["k"].shape == (1,)


def test_actor_lock_is_fifo_and_reentrant():
    """A Ray actor serves its mailbox in arrival order: the shim's per-actor lock hands over to waiters first-in
    first-out (a learner's sample_batch is not starved by producers re-acquiring the lock) and is re-entrant."""
    import threading
    from ddrl_b200.ray_shim import _FifoLock
    lock, order, started = _FifoLock(), [], []

    def waiter(i):
        started.append(i)
        with lock:
            with lock:                      # re-entrant
                order.append(i)
                time.sleep(0.002)

    with lock:
        threads = []
        for i in range(6):
            t = threading.Thread(target=waiter, args=(i,))
            t.start()
            while len(started) <= i or len(lock._waiters) <= i:      # thread i is queued before thread i + 1 starts
                time.sleep(0.0005)
            threads.append(t)
    for t in threads:
        t.join()
    assert order == list(range(6))
