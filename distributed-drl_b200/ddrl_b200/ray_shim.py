"""In-process stand-in for the subset of Ray the reference drivers use:
`ray.init`, `@ray.remote` on classes and functions, `.remote(...)`, `._remote(args=, resources=)`,
`ray.get`, `ray.wait`, `ray.put`.

Semantics kept from Ray: an actor executes ONE method at a time (a lock per actor); arguments of an
actor call are captured by value at call time (numpy arrays are copied, like pickling into the
object store); remote functions run concurrently (daemon threads — the reference's workers are
infinite loops); `ray.get` re-raises the task's exception.  There is no process boundary: state
lives on this process's GPU, which is the point of the one-process-per-GPU design.

Two additions for running the reference's own driver scripts (ddrl_b200.compat.run_reference_script):
`SUBSTITUTE` maps the NAME of a class decorated with @ray.remote to the class to instantiate instead (the scripts
define ReplayBuffer / ParameterServer inline; this is how the B200 ones are dropped in without editing them), and
`stop()` makes every later `.remote()`, `ray.get`, `ray.wait` raise `Stopped` so the scripts' infinite loops end.
"""
from __future__ import annotations

import collections
import copy
import threading
import time

import numpy as np


class Stopped(BaseException):
    """Raised inside tasks and in the driver once stop() has been called (BaseException: the reference's loops have
    no handler that could swallow it)."""


SUBSTITUTE = {}          # class name -> class used instead of the decorated one
ACTORS = []              # (class name, ActorHandle) of every actor created since reset()
TASKS = []               # ObjectRef of every remote function call since reset()
_STOP = threading.Event()


def stop():
    _STOP.set()


def stopped():
    return _STOP.is_set()


def reset():
    _STOP.clear()
    SUBSTITUTE.clear()
    del ACTORS[:]
    del TASKS[:]


def _check_stop():
    if _STOP.is_set():
        raise Stopped()


class ObjectRef:
    def __init__(self):
        self._ev = threading.Event()
        self._val = None
        self._exc = None

    def _set(self, val=None, exc=None):
        self._val, self._exc = val, exc
        self._ev.set()

    def ready(self):
        return self._ev.is_set()


def _by_value(a):
    if isinstance(a, np.ndarray):
        return a.copy()
    if isinstance(a, (list, tuple)) and any(isinstance(x, np.ndarray) for x in a):
        return type(a)(_by_value(x) for x in a)
    return a


class _FifoLock:
    """Re-entrant lock that serves waiters in ARRIVAL order — a Ray actor executes its mailbox first-in first-out, so a
    learner's sample_batch call is never starved by rollout workers hammering store (threading.RLock makes no such promise
    and, under the GIL, lets a tight producer loop re-acquire it for seconds)."""

    def __init__(self):
        self._mu = threading.Lock()
        self._waiters = collections.deque()
        self._owner, self._depth = None, 0

    def __enter__(self):
        me = threading.get_ident()
        with self._mu:
            if self._owner == me:
                self._depth += 1
                return self
            if self._owner is None and not self._waiters:
                self._owner, self._depth = me, 1
                return self
            ev = threading.Event()
            self._waiters.append((me, ev))
        ev.wait()
        return self

    def __exit__(self, *exc):
        with self._mu:
            self._depth -= 1
            if self._depth:
                return False
            if self._waiters:
                self._owner, ev = self._waiters.popleft()
                self._depth = 1
                ev.set()
            else:
                self._owner = None
        return False


class _ActorMethod:
    def __init__(self, actor, name):
        self._actor, self._name = actor, name

    def remote(self, *args, **kwargs):
        _check_stop()
        ref = ObjectRef()
        args = tuple(_by_value(a) for a in args)
        with self._actor._lock:
            try:
                ref._set(getattr(self._actor._obj, self._name)(*args, **kwargs))
            except BaseException as e:  # noqa: BLE001
                ref._set(exc=e)
        return ref


class ActorHandle:
    def __init__(self, obj):
        self._obj = obj
        self._lock = _FifoLock()

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _ActorMethod(self, name)


class _RemoteClass:
    def __init__(self, cls):
        self._cls = cls

    def remote(self, *args, **kwargs):
        _check_stop()
        cls = SUBSTITUTE.get(self._cls.__name__, self._cls)
        h = ActorHandle(cls(*args, **kwargs))
        ACTORS.append((self._cls.__name__, h))
        return h

    def _remote(self, args=(), kwargs=None, **_resources):
        return self.remote(*args, **(kwargs or {}))

    def options(self, **_):
        return self


class _RemoteFunction:
    def __init__(self, fn):
        self._fn = fn

    def remote(self, *args, **kwargs):
        _check_stop()
        ref = ObjectRef()
        ref.name = self._fn.__name__

        def run():
            try:
                ref._set(self._fn(*args, **kwargs))
            except Stopped:
                ref._set(None)
            except BaseException as e:  # noqa: BLE001
                ref._set(exc=e)

        t = threading.Thread(target=run, daemon=True, name=f"ray-task-{self._fn.__name__}")
        t.start()
        ref._thread = t
        TASKS.append(ref)
        return ref

    def _remote(self, args=(), kwargs=None, **_resources):
        return self.remote(*args, **(kwargs or {}))

    def options(self, **_):
        return self


def remote(*dargs, **dkwargs):
    """@ray.remote, @ray.remote(num_gpus=1, max_calls=1), @ray.remote(num_cpus=2) ..."""
    def wrap(x):
        return _RemoteClass(x) if isinstance(x, type) else _RemoteFunction(x)
    if len(dargs) == 1 and not dkwargs and callable(dargs[0]):
        return wrap(dargs[0])
    return wrap


def init(*_a, **_k):
    return {}


def shutdown():
    pass


def put(x):
    ref = ObjectRef()
    ref._set(copy.deepcopy(x))
    return ref


def get(ref, timeout=None):
    if isinstance(ref, (list, tuple)):
        return [get(r, timeout) for r in ref]
    t0 = time.time()
    while not ref._ev.wait(0.05):
        _check_stop()
        if timeout is not None and time.time() - t0 >= timeout:
            raise TimeoutError("ray_shim.get timed out")
    if ref._exc is not None:
        raise ref._exc
    return ref._val


def wait(refs, num_returns=1, timeout=None):
    t0 = time.time()
    while True:
        _check_stop()
        ready = [r for r in refs if r.ready()]
        if len(ready) >= num_returns or (timeout is not None and time.time() - t0 >= timeout):
            return ready[:num_returns] if len(ready) >= num_returns else ready, [r for r in refs if r not in ready]
        time.sleep(0.005)
