"""TEST INFRASTRUCTURE ONLY — materialise the reference's OWN Python under oracle/_ref/ (git-ignored).

/root/reference exists only in the build container; the GPU box receives a snapshot of this repository.  This recipe
copies, at build time, (a) the two driver scripts of the hot path whole and (b) the stand-alone classes the oracle is
pinned against, from where they lie under /root/reference into oracle/_ref/ — which is listed in .gitignore, so no
reference source ever enters the history, but NOT in .gpurunignore, so the files travel to the GPU box like a built
.so.  Consumers: tests/ (the reference's worker_train / worker_rollout / Cache / __main__ executed unchanged on the
stand-in modules of ddrl_b200.compat) and bench.py's reference arm / cpu_baseline (the reference's numpy ReplayBuffer
timed on the box's host cores).  Nothing in the product package reads oracle/.

    python -m oracle.materialize_ref          # writes oracle/_ref/*, prints the manifest
"""
from __future__ import annotations

import ast
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("DDRL_REFERENCE_ROOT", "/root/reference")

# whole scripts (executed with runpy by ddrl_b200.compat.run_reference_script)
SCRIPTS = {
    "example_dsac.py": "example/dsac.py",
    "algos_sac1_sac1.py": "algos/sac1/sac1.py",
}
# (output module, source file, [top-level class names]) — class bodies only need numpy / pickle
CLASSES = [
    ("ref_replay_sac.py", "example/sac.py", ["ReplayBuffer"]),
    ("ref_replay_dsac.py", "example/dsac.py", ["ReplayBuffer", "ParameterServer"]),
    ("ref_replay_sac1.py", "algos/sac1/sac1.py", ["ReplayBuffer", "ParameterServer"]),
    ("ref_replay_nstep.py", "algos/sac1/sac_ray.py", ["ReplayBuffer"]),
]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "example", "dsac.py"))


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def materialize(verbose=False):
    """Returns the manifest dict; raises if the reference tree is absent."""
    if not available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not present: oracle/_ref can only be (re)built in the build container")
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for out, rel in SCRIPTS.items():
        src = os.path.join(REFERENCE_ROOT, rel)
        with open(src, "r") as f:
            text = f.read()
        with open(os.path.join(OUT, out), "w") as f:
            f.write(text)                        # byte-identical copy: "the reference's scripts, unchanged"
        manifest[out] = dict(source=rel, sha256=_sha(src), kind="script")
    for out, rel, names in CLASSES:
        src = os.path.join(REFERENCE_ROOT, rel)
        with open(src, "r") as f:
            text = f.read()
        tree = ast.parse(text, filename=src)
        parts = [f'"""Extracted verbatim by oracle/materialize_ref.py from {rel} (sha256 {_sha(src)}); decorators dropped."""',
                 "import numpy as np", "import pickle", ""]
        spans = {}
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in names:
                seg = ast.get_source_segment(text, node)          # the class statement without its decorators
                parts += [seg, "", ""]
                spans[node.name] = [node.lineno, node.end_lineno]
        missing = [n for n in names if n not in spans]
        if missing:
            raise LookupError(f"{missing} not found in {src}")
        with open(os.path.join(OUT, out), "w") as f:
            f.write("\n".join(parts))
        manifest[out] = dict(source=rel, sha256=_sha(src), kind="classes", lines=spans)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print(json.dumps(manifest, indent=1, sort_keys=True))
    return manifest


def manifest():
    path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        return json.load(f)


def path(name):
    p = os.path.join(OUT, name)
    return p if os.path.isfile(p) else None


if __name__ == "__main__":
    materialize(verbose=True)
