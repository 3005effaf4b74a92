"""Shared helpers for reading the committed golden fixtures (tests/golden/)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DIMS = {"3wrobotNI": (3, 2), "3wrobot": (5, 2), "2tank": (2, 1)}
PRESET = {
    "3wrobotNI": dict(pars=[], bnds=[[-25, 25], [-5, 5]], R1_diag=[1, 10, 1, 0, 0], dt=0.01, psm=1.0, target=[]),
    "3wrobot": dict(pars=[10, 1], bnds=[[-300, 300], [-100, 100]], R1_diag=[1, 10, 1, 0, 0, 0, 0], dt=0.01, psm=2.0, target=[]),
    "2tank": dict(pars=[18.4, 24.4, 1.3, 1, 0.2], bnds=[[0, 1]], R1_diag=[10, 10, 1], dt=0.1, psm=2.0, target=[0.5, 0.5]),
}


def load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def mixed_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor): relative for large values, absolute below `floor`."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
