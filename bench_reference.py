"""The UNMODIFIED reference (rcognita v0.1.2, installed into ``baseline/_ref`` by
``python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``)
timed on the host cores: the baseline BASELINE.md section 4 asks to have reported beside the GPU number.

Nothing of this repo runs on this path: the objects are the reference's own ``Sys3WRobotNI`` / ``CtrlOptPred`` /
``Simulator`` built exactly like ``presets/main_3wrobot_NI.py:214-316`` and driven by the loop of ``:415-440``; the
arithmetic is numpy / scipy (RK45, SLSQP).  GUI-only imports of the package (matplotlib, mpldatacursor, svgpath2mpl --
absent from the image) are stubbed before ``import rcognita``; that does not touch the path.

Two measurements, each on every host core at once (``multiprocessing``, one environment per worker, its own seeded
initial state) for a fixed wall budget:
  preset    the preset-faithful closed loop: SLSQP ``_actor_optimizer`` on ``_actor_cost`` (config 1 parameters):
            accepted ``sim_step`` calls / s and ``_actor_cost`` calls / s;
  candidates  the closed loop with the candidate / arg-min controller of configs 2-4 (``_actor_cost`` on every row of a
            256-row table + ``np.argmin``): the same work per sample as the GPU arm does per environment.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "rcognita"))


def _import_reference():
    import types
    import warnings
    warnings.simplefilter("ignore")
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpldatacursor", "svgpath2mpl"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpldatacursor"].datacursor = lambda *a, **k: None
    sys.modules["svgpath2mpl"].parse_path = lambda *a, **k: None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from rcognita import controllers, simulator, systems
    return systems, simulator, controllers


def _worker(job):
    kind, seed, budget_s, nactor, ncand = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    systems, simulator, controllers = _import_reference()
    rng = np.random.default_rng(seed)
    dt, t1 = 0.01, 1e9
    ctrl_bnds = np.array([[-25.0, 25.0], [-5.0, 5.0]])
    state_init = np.array([rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-np.pi, np.pi)])
    my_sys = systems.Sys3WRobotNI(sys_type="diff_eqn", dim_state=3, dim_input=2, dim_output=3, dim_disturb=2, pars=[],
                                  ctrl_bnds=ctrl_bnds, is_dyn_ctrl=0, is_disturb=0,
                                  pars_disturb=np.array([[200 * dt, 200 * dt], [0, 0], [0.3, 0.3]]))
    ctrl = controllers.CtrlOptPred(2, 3, "MPC", ctrl_bnds=ctrl_bnds, action_init=[], t0=0, sampling_time=dt, Nactor=nactor,
                                   pred_step_size=dt, sys_rhs=my_sys._state_dyn, sys_out=my_sys.out, state_sys=state_init,
                                   prob_noise_pow=False, is_est_model=0, model_est_stage=1.0, model_est_period=dt,
                                   buffer_size=10, model_order=5, model_est_checks=0, gamma=1, Ncritic=4, critic_period=dt,
                                   critic_struct="quad-nomix", stage_obj_struct="quadratic",
                                   stage_obj_pars=[np.diag(np.array([1.0, 10.0, 1.0, 0.0, 0.0]))], observation_target=[])
    sim = simulator.Simulator(sys_type="diff_eqn", closed_loop_rhs=my_sys.closed_loop_rhs, sys_out=my_sys.out,
                              state_init=state_init, disturb_init=np.array([0, 0]), action_init=np.zeros(2), t0=0, t1=t1,
                              dt=dt, max_step=dt / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0, is_dyn_ctrl=0)
    calls = [0]
    orig_cost = ctrl._actor_cost

    def counted(action_sqn, observation):
        calls[0] += 1
        return orig_cost(action_sqn, observation)

    ctrl._actor_cost = counted
    if kind == "candidates":
        table = rng.uniform(ctrl.action_sqn_min, ctrl.action_sqn_max, size=(ncand, nactor * 2))

        def argmin_actor(observation):
            J = [ctrl._actor_cost(u, observation) for u in table]
            return table[int(np.argmin(J)), :2].copy()

        ctrl._actor_optimizer = argmin_actor
    steps = samples = 0
    t_begin = time.perf_counter()
    while time.perf_counter() - t_begin < budget_s:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        before = ctrl.ctrl_clock
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, "MPC")
        samples += int(ctrl.ctrl_clock != before)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        steps += 1
    return steps, samples, calls[0], time.perf_counter() - t_begin


def measure(cores: int, budget_s: float = 6.0, nactor: int = 6, ncand: int = 256) -> dict:
    """Runs both closed loops on ``cores`` worker processes for ``budget_s`` seconds each."""
    if not available():
        return {"unavailable": "baseline/_ref/rcognita is not installed"}
    out = {"cores": cores, "kind": "reference", "budget_s_per_loop": budget_s,
           "what": "unmodified rcognita v0.1.2 from baseline/_ref (numpy/scipy), one environment per core"}
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for kind in ("preset", "candidates"):
            res = pool.map(_worker, [(kind, 1000 + i, budget_s, nactor, ncand) for i in range(cores)])
            wall = max(r[3] for r in res)
            out[kind] = {"env_steps_per_s": sum(r[0] for r in res) / wall, "actor_cost_evals_per_s": sum(r[2] for r in res) / wall,
                         "controller_samples_per_s": sum(r[1] for r in res) / wall, "wall_s": wall,
                         "controller": ("scipy SLSQP on _actor_cost (presets/main_3wrobot_NI.py, ctrl_mode=MPC, Nactor=%d)" % nactor)
                         if kind == "preset" else ("_actor_cost on %d candidate rows + np.argmin per sample" % ncand)}
    return out


if __name__ == "__main__":
    import json
    n = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, len(os.sched_getaffinity(0)))
    print(json.dumps(measure(n, float(sys.argv[2]) if len(sys.argv) > 2 else 4.0)))
