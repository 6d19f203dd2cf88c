"""The reference's role tasks — worker_train / worker_rollout / worker_test (algos/sac1/sac1.py:133-252,
example/dsac.py:76-177) — are NOT re-implemented here: ddrl_b200.compat.run_reference_script executes the reference's
own functions, unchanged, against this package (tests/test_compat_gpu.py).  This module only keeps the synthetic
environment under its old name."""
from .compat.envs import SyntheticEnv  # noqa: F401
