"""ParameterServer — the reference's weight store (algos/sac1/sac1.py:66-100 = example/dsac.py:51-73
plus pickle restore).

Semantics kept: values are COPIED on the way in (the reference copies because Ray object-store
arrays are read-only / shared, example/dsac.py:54-56,60); pull(keys) returns a list aligned with
`keys`; get_weights() returns the dict; save_weights(name) writes name + "weights.pickle" holding
{name: float32 ndarray} — the format algos/sac1/render_test.py:32-38 loads.

B200 side: values may be CUDA tensors (a Learner's device-resident weights); they are kept on the
GPU as one flat float32 buffer so that, across ranks, push/pull is a single NCCL broadcast of that
buffer over NVLink (see dist.py: broadcast_parameters) instead of pickle + RPC.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict

import numpy as np
import torch


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().to("cpu", copy=True).numpy()
    return np.array(v, copy=True)


class ParameterServer(object):
    def __init__(self, keys, values, weights_file=""):
        if weights_file:
            # reference: try/except -> print + exit(); we raise the underlying error instead
            with open(weights_file, "rb") as pickle_in:
                self.weights = pickle.load(pickle_in)
        else:
            self.weights = OrderedDict((k, _to_numpy(v)) for k, v in zip(keys, values))
        self.version = 0

    def push(self, keys, values):
        for key, value in zip(keys, values):
            self.weights[key] = _to_numpy(value)
        self.version += 1

    def pull(self, keys):
        return [self.weights[key] for key in keys]

    def get_weights(self):
        return self.weights

    def save_weights(self, name):
        with open(name + "weights.pickle", "wb") as pickle_out:
            pickle.dump(dict(self.weights), pickle_out)
