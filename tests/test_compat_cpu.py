"""CPU: the reference's own driver scripts (materialised under oracle/_ref by oracle/materialize_ref.py) execute
UNCHANGED under the stand-in modules of ddrl_b200.compat.  Here the GPU classes are replaced by tiny fakes (no GPU in
this container) and the scripts keep their own inline numpy ReplayBuffer / ParameterServer (substitute=False): what
is proven is the plumbing — imports, flags, the ray calling convention, Cache's process / queues, the bounded run.
tests/test_compat_gpu.py runs the same scripts against the real ddrl_b200 classes."""
import os
import sys

import numpy as np
import pytest

from oracle import materialize_ref


@pytest.fixture(scope="module")
def ref_scripts():
    if materialize_ref.available():
        materialize_ref.materialize()
    if materialize_ref.manifest() is None:
        pytest.skip("oracle/_ref has not been materialised (needs /root/reference once, in the build container)")
    return materialize_ref


class FakeLearner(object):
    names = ["main/pi/dense/kernel", "main/q1/dense/kernel"]
    trained = 0

    def __init__(self, opt, job="learner", **_k):
        self.opt = opt
        self.w = {n: np.zeros((2, 2), np.float32) for n in self.names}

    def get_weights(self):
        return list(self.w), [v.copy() for v in self.w.values()]

    def set_weights(self, keys, values):
        self.w.update(zip(keys, values))

    def train(self, batch):
        assert set(batch) >= {"obs1", "obs2", "acts", "rews", "done"}
        for v in self.w.values():
            v += 1.0
        FakeLearner.trained += 1


class FakeActor(FakeLearner):
    @classmethod
    def from_learner(cls, learner):
        return cls(learner.opt)

    def get_action(self, o, deterministic=False):
        return np.zeros(int(getattr(self.opt, "act_dim", 2)), np.float32)

    def test(self, test_env, replay_buffer=None, n=25):
        return 0.0


@pytest.mark.parametrize("script,argv", [("example_dsac.py", []), ("algos_sac1_sac1.py", ["--env_name", "BipedalWalker-v2"])])
def test_reference_driver_runs_unchanged_on_the_stand_ins(ref_scripts, script, argv, monkeypatch):
    import ddrl_b200.learner as learner_mod
    from ddrl_b200 import compat
    monkeypatch.setattr(learner_mod, "Learner", FakeLearner)
    monkeypatch.setattr(learner_mod, "Actor", FakeActor)
    FakeLearner.trained = 0
    path = ref_scripts.path(script)
    sha = ref_scripts.manifest()[script]["sha256"]
    # runs until the learner has trained past its second parameter push (a loaded machine only makes that take longer)
    out = compat.run_reference_script(path, argv, budget_s=60.0, time_scale=0.002, substitute=False,
                                      until=lambda: FakeLearner.trained > 650)
    print(f"{script}: reference sha256 {sha}; tasks {[(n, type(e).__name__ if e else None) for n, e in out['tasks']]}")
    assert out["error"] is None, repr(out["error"])
    names = [n for n, _ in out["tasks"]]
    assert names.count("worker_rollout") >= 1 and "worker_train" in names and "worker_test" in names
    assert all(e is None for n, e in out["tasks"] if n != "worker_test"), out["tasks"]   # worker_test: see SURVEY D-3
    rb, ps = out["actors"]["ReplayBuffer"], out["actors"]["ParameterServer"]
    assert type(rb).__module__ != "ddrl_b200.replay"            # the script's own inline class (substitute=False)
    assert rb.size > 0 and FakeLearner.trained > 300            # rollouts stored, the learner trained ...
    assert float(ps.weights["main/pi/dense/kernel"].max()) >= 300.0     # ... and pushed every 300 updates
    assert "multiprocessing" in sys.modules and hasattr(sys.modules["multiprocessing"], "get_context")  # restored
