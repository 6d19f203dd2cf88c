"""Stand-in modules that let the reference's own driver scripts (example/dsac.py, algos/sac1/sac1.py) execute
UNCHANGED against ddrl_b200 (BASELINE.json north_star).

The scripts import, at module top, packages that are not dependencies of this repository (ray, tensorflow 1.x, gym,
spinup) and sibling modules of their own directory (model, actor_learner, hyperparams, core).  `install()` puts
stand-ins for exactly those names into sys.modules:
    ray             ddrl_b200.ray_shim (in-process actors / tasks); with substitute=True the script's inline
                    `@ray.remote class ReplayBuffer / ParameterServer` are instantiated as ddrl_b200.ReplayBuffer /
                    ddrl_b200.ParameterServer — same constructor and method signatures, data in HBM
    actor_learner   Actor, Learner                       -> ddrl_b200.Actor / Learner (algos/sac1/actor_learner.py:19-229)
    model           Model(args)                          -> the same learner behind example/model.py:12-118's surface
    hyperparams     HyperParameters(env_name, total_epochs, num_workers, a_l_ratio), Wrapper — the generation the
                    script itself calls (algos/sac1/sac1.py:259; the tree's hyperparams.py belongs to sac_ray.py, SURVEY D-2)
    gym             make() / spaces.Box over synthetic environments of the named shapes (compat/envs.py)
    tensorflow      tf.app.flags only (all the scripts themselves use of it)
    spinup          EpochLogger / setup_logger_kwargs / core.get_vars as inert objects
    multiprocessing (algos/sac1/sac1.py's Cache, :103-130) Process -> thread, Queue -> queue.Queue: the replay ring
                    lives on this process's GPU, a forked child could not touch it
`run_reference_script()` executes a script as __main__ under those stand-ins for a bounded time (the scripts' loops
never return) and reports what the actors saw.  Nothing here reads oracle/ or /root/reference: the caller passes the
path of the script.
"""
from __future__ import annotations

import queue as _queue
import runpy
import sys
import threading
import time as _time
import types
from types import SimpleNamespace

import numpy as np

from .. import ray_shim
from . import envs


# ------------------------------------------------------------------------------------------------
# model / actor_learner / hyperparams
# ------------------------------------------------------------------------------------------------
class Model(object):
    """example/model.py's surface over the SAC1 learner: Model(args).train(replay_buffer, args) /
    get_weights / set_weights / get_action / test_agent."""

    def __init__(self, args):
        from ..learner import Actor, Learner
        self._learner = Learner(args, "model", max_batch=int(getattr(args, "batch_size", 256) or 256))
        self._actor = Actor.from_learner(self._learner)

    def set_weights(self, variable_names, weights):
        self._learner.set_weights(variable_names, weights)

    def get_weights(self):
        return self._learner.get_weights()

    def get_action(self, o, deterministic=False):
        return self._actor.get_action(o, deterministic)

    def train(self, replay_buffer, args):
        # example/model.py:92-101: a blocking sample_batch RPC, then one update
        batch = ray_shim.get(replay_buffer.sample_batch.remote(args.batch_size))
        self._learner.train(batch)

    def test_agent(self, test_env, args, n=10):
        test_ret = []
        for _ in range(n):
            o, d, ep_ret, ep_len = test_env.reset(), False, 0, 0
            while not (d or (ep_len == args.max_ep_len)):
                o, r, d, _ = test_env.step(self.get_action(o, True))
                ep_ret += r
                ep_len += 1
            test_ret.append(ep_ret)
        return sum(test_ret) / len(test_ret)


class Wrapper(object):
    """Action repeat + observation / action noise + reward scale (algos/sac1/hyperparams.py:107-134, without the
    BipedalWalker dimensions hard-wired there)."""

    def __init__(self, env, obs_noise, act_noise, reward_scale, action_repeat=3):
        self._env, self.obs_noise, self.act_noise = env, obs_noise, act_noise
        self.reward_scale, self.action_repeat = reward_scale, action_repeat
        self.action_space, self.observation_space = env.action_space, env.observation_space
        self._rng = np.random.default_rng(0)

    def reset(self):
        o = self._env.reset()
        return o + self.obs_noise * self._rng.standard_normal(o.shape)

    def step(self, action):
        action = np.asarray(action) + self.act_noise * self._rng.standard_normal(np.shape(action))
        r = 0.0
        for _ in range(self.action_repeat):
            o, r1, d, info = self._env.step(action)
            r += r1
            if d:
                break
        return o + self.obs_noise * self._rng.standard_normal(o.shape), r * self.reward_scale, d, info


class HyperParameters(object):
    """The option object algos/sac1/sac1.py builds at :259 and reads in worker_* (field list: SURVEY.md Appendix C)."""

    def __init__(self, env_name, total_epochs, num_workers, a_l_ratio):
        self.env_name, self.total_epochs, self.num_workers, self.a_l_ratio = env_name, total_epochs, num_workers, a_l_ratio
        env = envs.make(env_name)
        self.obs_dim, self.act_dim = env.observation_space.shape[0], env.action_space.shape[0]
        self.ac_kwargs = dict(hidden_sizes=(400, 300), action_space=env.action_space)      # core.py:91's default net
        self.action_space = env.action_space
        self.alpha, self.gamma, self.lr, self.polyak = 0.1, 0.997, 5e-5, 0.995           # hyperparams.py:31-95
        self.replay_size, self.batch_size = int(1e6), 256
        self.start_steps, self.max_ep_len, self.steps_per_epoch = int(5e4), 1000, 5000
        self.obs_noise, self.act_noise, self.reward_scale = 0, 0.3, 5
        self.seed, self.gpu_fraction, self.summary_dir = 0, 0.3, "./tboard_ray"
        self.num_learners = 1


# ------------------------------------------------------------------------------------------------
# multiprocessing stand-in (Cache): threads and in-process queues that honour ray_shim.stop()
# ------------------------------------------------------------------------------------------------
class _Queue(object):
    def __init__(self, maxsize=0):
        self._q = _queue.Queue(maxsize)

    def put(self, item):
        while True:
            ray_shim._check_stop()
            try:
                return self._q.put(item, timeout=0.05)
            except _queue.Full:
                pass

    def get(self):
        while True:
            ray_shim._check_stop()
            try:
                return self._q.get(timeout=0.05)
            except _queue.Empty:
                pass

    def empty(self):
        return self._q.empty()

    def qsize(self):
        return self._q.qsize()


class _Process(object):
    def __init__(self, target=None, args=(), kwargs=None, **_):
        def run():
            try:
                target(*args, **(kwargs or {}))
            except ray_shim.Stopped:
                pass
        self._t = threading.Thread(target=run, daemon=True, name="compat-process")

    def start(self):
        self._t.start()

    def join(self, timeout=None):
        self._t.join(0 if timeout else timeout)      # Cache.start() joins for 10 s only to let the child start

    def terminate(self):
        pass


# ------------------------------------------------------------------------------------------------
# sys.modules plumbing
# ------------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def stand_ins():
    """name -> module object for every import the reference drivers make that this repository does not provide."""
    from ..learner import Actor, Learner
    flags_ns = SimpleNamespace()

    def _define(name, default, _help=""):
        setattr(flags_ns, name, default)
    flags = SimpleNamespace(FLAGS=flags_ns, DEFINE_string=_define, DEFINE_integer=_define, DEFINE_float=_define,
                            DEFINE_boolean=_define, DEFINE_bool=_define)
    tf = _module("tensorflow", app=SimpleNamespace(flags=flags), set_random_seed=lambda *_a, **_k: None)

    class EpochLogger(object):
        def __init__(self, **_k):
            self.rows = []

        def save_config(self, *_a, **_k):
            pass

        def store(self, **_k):
            pass

        def log_tabular(self, key, val=None, **_k):
            self.rows.append((key, val))

        def dump_tabular(self):
            self.rows = []

    core = _module("spinup.algos.sac.core", get_vars=lambda scope: [], placeholders=lambda *a: tuple(None for _ in a))
    sac = _module("spinup.algos.sac", core=core)
    algos = _module("spinup.algos", sac=sac)
    logx = _module("spinup.utils.logx", EpochLogger=EpochLogger)
    run_utils = _module("spinup.utils.run_utils", setup_logger_kwargs=lambda exp_name, seed=None, **_k: dict())
    utils = _module("spinup.utils", logx=logx, run_utils=run_utils)
    spinup = _module("spinup", algos=algos, utils=utils)
    spaces = _module("gym.spaces", Box=envs.Box)
    gym = _module("gym", make=envs.make, spaces=spaces)
    import multiprocessing as _real_mp
    # anything but Process / Queue falls through to the real module (other libraries may import it meanwhile)
    mp = _module("multiprocessing", Process=_Process, Queue=_Queue, __getattr__=lambda name: getattr(_real_mp, name))
    ray_experimental = _module("ray.experimental", tf_utils=_module("ray.experimental.tf_utils"))
    ray_shim.experimental = ray_experimental
    return {
        "ray": ray_shim, "ray.experimental": ray_experimental, "ray.experimental.tf_utils": ray_experimental.tf_utils,
        "tensorflow": tf, "gym": gym, "gym.spaces": spaces,
        "spinup": spinup, "spinup.algos": algos, "spinup.algos.sac": sac, "spinup.algos.sac.core": core,
        "spinup.utils": utils, "spinup.utils.logx": logx, "spinup.utils.run_utils": run_utils,
        "model": _module("model", Model=Model),
        "actor_learner": _module("actor_learner", Actor=Actor, Learner=Learner),
        "hyperparams": _module("hyperparams", HyperParameters=HyperParameters, Wrapper=Wrapper),
        "core": _module("core"),
        "multiprocessing": mp,
    }


def run_reference_script(path, argv=(), budget_s=8.0, time_scale=0.01, substitute=True, until=None):
    """Execute the reference driver at `path` as __main__, unchanged, for at most `budget_s` seconds of wall clock — or
    until `until()` (polled every 50 ms) returns true, so that a test can wait for "the learner has pushed weights"
    instead of guessing how long that takes on a loaded machine.

    time.sleep() calls of the script are scaled by `time_scale` (the drivers sleep 5-20 s between launching their
    roles).  After the budget every Ray call raises ray_shim.Stopped, which ends the worker loops and the driver's
    final ray.wait.  Returns dict(actors={class name: underlying object}, tasks=[(function name, exception or None)],
    globals=<the script's namespace>, error=<exception that ended __main__ early, if any>)."""
    from ..ps import ParameterServer
    from ..replay import ReplayBuffer
    ray_shim.reset()
    if substitute:
        ray_shim.SUBSTITUTE.update(ReplayBuffer=ReplayBuffer, ParameterServer=ParameterServer)
    mods = stand_ins()
    saved = {k: sys.modules.get(k) for k in mods}
    saved_argv, saved_sleep = sys.argv, _time.sleep
    sys.modules.update(mods)
    # every Ray role is a thread of THIS process here: with CPython's default 5 ms switch interval each hand-over of an
    # actor's lock to a waiting thread costs up to 5 ms whenever another role (worker_test's evaluation loop) is computing,
    # which throttles the learner to a few updates per second; 0.2 ms keeps the roles interleaved like separate processes
    saved_switch = sys.getswitchinterval()
    sys.setswitchinterval(2e-4)
    sys.argv = [path] + list(argv)
    _time.sleep = lambda s: saved_sleep(min(float(s) * time_scale, 0.2))
    cancel = threading.Event()

    def watch():
        t_end = _time.monotonic() + budget_s
        while not cancel.is_set() and _time.monotonic() < t_end:
            try:
                if until is not None and until():
                    break
            except Exception:      # noqa: BLE001  (the condition may look at objects the script has not created yet)
                pass
            saved_sleep(0.05)
        if not cancel.is_set():
            ray_shim.stop()

    timer = threading.Thread(target=watch, daemon=True)
    timer.start()
    ns, err = {}, None
    try:
        ns = runpy.run_path(path, run_name="__main__")
    except ray_shim.Stopped:
        pass
    except BaseException as e:  # noqa: BLE001
        err = e
    finally:
        ray_shim.stop()
        cancel.set()
        _time.sleep = saved_sleep
        sys.setswitchinterval(saved_switch)
        sys.argv = saved_argv
        for ref in list(ray_shim.TASKS):
            t = getattr(ref, "_thread", None)
            if t is not None:
                t.join(timeout=20)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    actors = {}
    for name, h in ray_shim.ACTORS:
        actors.setdefault(name, h._obj)
    tasks = [(getattr(r, "name", "?"), r._exc) for r in ray_shim.TASKS]
    return dict(actors=actors, tasks=tasks, globals=ns, error=err)
