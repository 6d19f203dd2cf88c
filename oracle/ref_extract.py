"""TEST INFRASTRUCTURE ONLY — load the reference's OWN classes from /root/reference.

The reference modules cannot be imported (tensorflow / ray / gym / spinup are imported at module
top and are absent), but the ReplayBuffer and ParameterServer class bodies need only numpy and
pickle.  We parse the file with ``ast``, take the ClassDef, drop its ``@ray.remote`` decorator and
exec it.  No reference source is copied into this repository; this only works where
/root/reference exists (the build container) and is used to (a) pin oracle/replay_oracle.py and
(b) generate tests/golden/*.npz (oracle/make_golden.py).  It never runs on the GPU box.
"""
from __future__ import annotations

import ast
import os
import pickle

import numpy as np

REFERENCE_ROOT = os.environ.get("DDRL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "example", "dsac.py"))


def load_reference_class(relpath: str, name: str):
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path, "r") as f:
        tree = ast.parse(f.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ast.fix_missing_locations(mod)
            ns = {"np": np, "pickle": pickle, "print": lambda *a, **k: None}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise LookupError(f"{name} not found in {path}")


# (file, flavor) of every tuple-row ReplayBuffer variant with the (obs_dim, act_dim, size) ctor
REPLAY_VARIANTS = {
    "sac": "example/sac.py",
    "dsac": "example/dsac.py",
    "sac1": "algos/sac1/sac1.py",
}
PS_VARIANTS = {
    "dsac": "example/dsac.py",
    "sac1": "algos/sac1/sac1.py",
}


def reference_replay(variant="sac1"):
    return load_reference_class(REPLAY_VARIANTS[variant], "ReplayBuffer")


def reference_ps(variant="sac1"):
    return load_reference_class(PS_VARIANTS[variant], "ParameterServer")
