#!/usr/bin/env python3
"""Dump the critic least-squares problems (buffers, w_prev, fitted w, cost) met in the closed loops of
BASELINE configs 3 / 4 for a few thousand lanes, for offline analysis of the fit's iteration counts."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_workload import synthetic_candidates, synthetic_states  # noqa: E402
from configs import CFG  # noqa: E402
from rcognita_b200.engine import ClosedLoopEngine  # noqa: E402

out = {}
for name in ("config3", "config4"):
    c = CFG[name]
    E = 2048
    x0 = synthetic_states(c["system"], 0, E, seed=0)
    cand = synthetic_candidates(c["bnds"], c["N"], 256, seed=1)
    eng = ClosedLoopEngine(c["system"], x0, cand, pars=c["pars"], ctrl_bnds=c["bnds"], mode=c["mode"], Nactor=c["N"],
                           dt=c["dt"], pred_step_size=c["dt"] * c["psm"], t1=1e9, R1=c["R1"], observation_target=c["target"],
                           critic_struct=c["cs"], critic_fit=True, Ncritic=4, buffer_size=10, action_init=c["a_init"])
    for k in range(1, 181):
        wprev_before = eng.w_prev.clone()
        eng.run_interval()
        if k in (3, 8, 12, 20, 40, 100, 180):
            out[f"{name}_k{k}_obs_buf"] = eng.obs_buf.cpu().numpy()
            out[f"{name}_k{k}_act_buf"] = eng.act_buf.cpu().numpy()
            out[f"{name}_k{k}_w_prev"] = wprev_before.cpu().numpy()
            out[f"{name}_k{k}_w"] = eng.w.cpu().numpy()
            out[f"{name}_k{k}_Jc"] = eng.Jc.cpu().numpy()
            out[f"{name}_k{k}_flag"] = eng.critic_flag.cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/critic_problems.npz", **out)
print("saved", len(out), "arrays")
