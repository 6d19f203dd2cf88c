"""TEST INFRASTRUCTURE ONLY — CPU restatements of the reference's learner-side data path.

Nothing in the product package (``distributed-drl_b200/``) may import from here.  The only
allowed importers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``, and there only as the checker / the CPU arm that is
being compared against — never as the thing measured as the product or shipped.

Modules
-------
replay_oracle   numpy restatement of ReplayBuffer / ParameterServer (pinned against the reference
                classes extracted with ``ast`` from /root/reference, see ref_extract.py and
                tests/golden/) plus the Philox4x32-10 index stream the CUDA sampler uses.
sac1_oracle     torch-CPU (float64 / float32) restatement of the SAC1 learner step
                (algos/sac1/core.py + algos/sac1/actor_learner.py).  PARITY UNPINNED by the
                reference: the arithmetic lives in TensorFlow 1.x which is not vendored, not pinned
                and not installable here; the reference holds no tests or golden vectors.
ref_extract     (this container only) loads the reference's own classes from /root/reference.
make_golden     regenerates tests/golden/*.npz from the reference classes.
"""
