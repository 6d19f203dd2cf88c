"""GPU: the reference's own driver scripts — example/dsac.py and algos/sac1/sac1.py, byte-identical copies
materialised under oracle/_ref by oracle/materialize_ref.py — executed UNCHANGED as __main__ against ddrl_b200:
their inline `@ray.remote class ReplayBuffer / ParameterServer` are instantiated as ddrl_b200.ReplayBuffer /
ParameterServer (ray stand-in substitution by class name), `Model` / `Actor` / `Learner` resolve to the CUDA learner,
`Cache` prefetches through in-process queues, rollout workers store one transition per env step, the learner trains on
sampled batches and pushes weights every 300 updates, the tester pulls them (BASELINE.json north_star: "so dsac.py and
the algos/ scripts run unchanged").  Environments are synthetic (gym is not a dependency)."""
import numpy as np
import pytest

from oracle import materialize_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_scripts():
    import __graft_entry__
    __graft_entry__.build()
    if materialize_ref.manifest() is None:
        pytest.skip("oracle/_ref has not been materialised (needs /root/reference once, in the build container)")
    return materialize_ref


@pytest.mark.parametrize("script,argv,flavor", [("example_dsac.py", [], "dsac"),
                                                ("algos_sac1_sac1.py", ["--env_name", "BipedalWalker-v2"], "sac1")])
def test_reference_driver_runs_unchanged_against_ddrl_b200(ref_scripts, script, argv, flavor):
    import torch
    from ddrl_b200 import ParameterServer, ReplayBuffer, _native, compat
    path = ref_scripts.path(script)
    sha = ref_scripts.manifest()[script]["sha256"]
    before = _native.launch_count()
    def trained_enough():         # the learner has sampled past its second parameter push
        rbs = [h._obj for n, h in compat.ray_shim.ACTORS if n == "ReplayBuffer"]
        return bool(rbs) and rbs[0].sample_times > 650

    out = compat.run_reference_script(path, argv, budget_s=60.0, time_scale=0.002, substitute=True, until=trained_enough)
    torch.cuda.synchronize()
    print(f"{script}: reference file sha256 {sha}; tasks {[(n, type(e).__name__ if e else None) for n, e in out['tasks']]}")
    for n, e in out["tasks"]:
        if e is not None:
            print(f"  task {n} raised {type(e).__name__}: {e}")
    assert out["error"] is None, repr(out["error"])
    names = [n for n, _ in out["tasks"]]
    assert names.count("worker_rollout") >= 1 and "worker_train" in names and "worker_test" in names
    assert all(e is None for n, e in out["tasks"] if n != "worker_test"), out["tasks"]   # worker_test: see SURVEY D-3
    rb, ps = out["actors"]["ReplayBuffer"], out["actors"]["ParameterServer"]
    assert isinstance(rb, ReplayBuffer) and isinstance(ps, ParameterServer)       # the B200 classes were dropped in
    counts = rb.get_counts()
    steps = counts if isinstance(counts, int) else counts[1]
    assert steps > 0 and rb.size > 0                                             # rollout workers stored transitions
    assert rb.sample_times >= 300 and ps.version >= 1                            # the learner sampled, trained, pushed
    keys = [k for k in ps.get_weights() if "main/pi" in k]
    assert len(keys) == 8 and all(np.isfinite(ps.get_weights()[k]).all() for k in keys)
    assert _native.launch_count() - before > 1000                                # ... on the GPU
