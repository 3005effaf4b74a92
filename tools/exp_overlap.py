#!/usr/bin/env python3
"""Experiment (round 2): hiding the latency-bound `rcg_rk45_advance` launch behind the HBM-bound `rcg_actor_cost`
launch of BASELINE config 2 with `engine.PipelinedLoop` (P environment blocks on P streams).  Prints ms per control
interval for every (P, stagger) variant and checks the final state
against the single-stream engine bit for bit.

    python tools/exp_overlap.py [--envs 65536] [--steps 240] [--variants 1,2,2n,3,4]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench_workload import make_workload  # noqa: E402
from rcognita_b200.engine import ClosedLoopEngine, PipelinedLoop  # noqa: E402

KW = dict(ctrl_bnds=[[-25.0, 25.0], [-5.0, 5.0]], mode="MPC", dt=0.01, t1=100.0, R1=[1.0, 10.0, 1.0, 0.0, 0.0],
          action_init=[-2.5, -0.5])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--cands", type=int, default=256)
    ap.add_argument("--nactor", type=int, default=6)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--shared-cands", action="store_true")
    ap.add_argument("--variants", default="1,2,2n,3,4")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    x0, cand = make_workload(args, 0, args.envs)
    W = 5
    eng = ClosedLoopEngine("3wrobotNI", x0, cand, Nactor=args.nactor, **KW)
    for _ in range(W + args.steps):
        eng.run_interval()
    ref = eng.results()
    del eng
    for name in args.variants.split(","):
        stagger = not name.endswith("n")
        P = int(name.rstrip("n"))
        loop = PipelinedLoop("3wrobotNI", x0, cand, nchunks=P, stagger=stagger, Nactor=args.nactor, **KW)
        for _ in range(W):
            loop.step()
        loop.synchronize()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loop.step()
        loop.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res = loop.results()
        print(json.dumps({"P": P, "stagger": stagger, "ms_per_interval": ms, "evals_per_s": args.envs * args.cands / (ms * 1e-3),
                          "bit_identical": all(np.array_equal(res[k], ref[k], equal_nan=True) for k in ref)}), flush=True)
        del loop


if __name__ == "__main__":
    main()
