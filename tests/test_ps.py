"""CPU: ParameterServer (product, host dict semantics) against the reference-generated golden
vectors (algos/sac1/sac1.py:66-100)."""
import os

import numpy as np
import torch

from ddrl_b200.ps import ParameterServer


def test_ps_matches_reference_golden(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "ps_sac1.npz"))
    keys = [str(k) for k in g["keys"]]
    init = [g["init0"].copy(), g["init1"].copy(), g["init2"].copy()]
    ps = ParameterServer(keys, init)
    init[0][:] = 777.0                     # values are copied on the way in (example/dsac.py:54-56)
    push = [g["push0"].copy(), g["push2"].copy()]
    ps.push([keys[0], keys[2]], push)
    push[1][:] = 0.0                       # ... and on push (example/dsac.py:60)
    got = ps.pull([keys[2], keys[1], keys[0]])
    for a, k in zip(got, ("pull_k2", "pull_k1", "pull_k0")):
        assert a.dtype == np.float32 and np.array_equal(a, g[k])
    assert list(ps.get_weights()) == keys
    ps.save_weights(str(tmp_path / "run_"))
    ps2 = ParameterServer([], [], weights_file=str(tmp_path / "run_weights.pickle"))
    for k in keys:
        assert np.array_equal(ps2.get_weights()[k], ps.get_weights()[k])


def test_ps_accepts_torch_values():
    ps = ParameterServer(["a"], [torch.arange(6, dtype=torch.float32).reshape(2, 3)])
    v = ps.pull(["a"])[0]
    assert isinstance(v, np.ndarray) and v.shape == (2, 3) and v[1, 2] == 5.0
