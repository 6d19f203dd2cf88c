"""ddrl_b200 — B200-native learner-side data path of createamind/Distributed-DRL.

Public surface (same names and call pattern as the reference's inline classes):
    ReplayBuffer(obs_dim, act_dim, size).store / sample_batch / get_counts
    Cache(replay_buffer).start / q1.get / q2.put / end              (the learner's prefetcher, algos/sac1/sac1.py:103-130)
    ParameterServer(keys, values[, weights_file]).push / pull / get_weights / save_weights
    Learner(opt, job).train / get_weights / set_weights
    NStepReplayBuffer(opt).store / sample_batch / get_counts        (algos/sac1/sac_ray.py sequence ring)
    DQNLearner(opt, job) / SQNLearner(opt, job).train / get_weights / set_weights   (algos/dqn, algos/sqn learners)
All numerics run in libddrl_b200.so (hand-written sm_100a CUDA, C ABI in include/ddrl_b200.h).
"""
from .replay import Cache, ReplayBuffer  # noqa: F401
from .ps import ParameterServer  # noqa: F401
from .learner import Actor, Learner  # noqa: F401
from .nstep import NStepReplayBuffer  # noqa: F401
from .qlearn import DQNLearner, SQNLearner  # noqa: F401
from . import _native  # noqa: F401

__all__ = ["ReplayBuffer", "Cache", "ParameterServer", "Learner", "Actor", "NStepReplayBuffer", "DQNLearner", "SQNLearner"]
