"""Synthetic workload of bench.py / the sharding tests: initial states and
candidate action sequences keyed by GLOBAL environment index (Philox counter-based streams,
one per block of ``BLOCK`` environments), so an environment gets the same inputs no matter
how many ranks the batch is sharded over.  numpy only: shared by the CUDA arm and the CPU arm.
"""
from __future__ import annotations

import numpy as np

BLOCK = 1024        # environments per Philox stream (== rcognita_b200.shard.BLOCK)

STATE_BOX = {       # plot boxes of the presets / visuals.py:644
    "3wrobotNI": ([-10.0, -10.0, -np.pi], [10.0, 10.0, np.pi]),
    "3wrobot": ([-10.0, -10.0, -np.pi, -1.0, -1.0], [10.0, 10.0, np.pi, 1.0, 1.0]),
    "2tank": ([-2.0, -2.0], [2.0, 2.0]),
}


def _stream(seed: int, block: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed), int(block)]))


def synthetic_states(system: str, lo: int, hi: int, seed: int = 0) -> np.ndarray:
    """Initial states [hi-lo, n] ~ U(box) for global environments lo..hi-1."""
    if lo % BLOCK or hi % BLOCK:
        raise ValueError(f"shard boundaries must be multiples of {BLOCK}")
    blo, bhi = np.asarray(STATE_BOX[system][0]), np.asarray(STATE_BOX[system][1])
    out = [(_stream(seed, b).uniform(blo, bhi, size=(BLOCK, blo.size))) for b in range(lo // BLOCK, hi // BLOCK)]
    return np.concatenate(out, axis=0) if out else np.zeros((0, blo.size))


def synthetic_candidates(ctrl_bnds, Nactor: int, C: int, seed: int = 1, env_range=None) -> np.ndarray:
    """Candidate action sequences ~ U(action_sqn_min, action_sqn_max) (controllers.py:970-971):
    a shared table [C, Nactor*m] (``env_range`` None) or per-environment sets
    [hi-lo, C, Nactor*m] keyed by global environment index."""
    b = np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2)
    lo_sq, hi_sq = np.tile(b[:, 0], Nactor), np.tile(b[:, 1], Nactor)
    if env_range is None:
        return _stream(seed, 0).uniform(lo_sq, hi_sq, size=(C, lo_sq.size))
    lo, hi = env_range
    if lo % BLOCK or hi % BLOCK:
        raise ValueError(f"shard boundaries must be multiples of {BLOCK}")
    out = [_stream(seed, 1 + blk).uniform(lo_sq, hi_sq, size=(BLOCK, C, lo_sq.size)).astype(np.float64)
           for blk in range(lo // BLOCK, hi // BLOCK)]
    return np.concatenate(out, axis=0)




def make_workload(args, lo: int, hi: int):
    """(initial states [hi-lo, 3], candidates) of bench.py's 3wrobot_NI workload for global
    environments lo..hi-1; ``args`` carries nactor / cands / shared_cands."""
    bnds = [[-25.0, 25.0], [-5.0, 5.0]]
    x0 = synthetic_states("3wrobotNI", lo, hi, seed=0)
    if args.shared_cands:
        cand = synthetic_candidates(bnds, args.nactor, args.cands, seed=1)
    else:
        cand = synthetic_candidates(bnds, args.nactor, args.cands, seed=1, env_range=(lo, hi))
    return x0, cand
