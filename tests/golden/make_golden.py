#!/usr/bin/env python3
"""Generate golden vectors from the LIVE reference (rcognita v0.1.2 at /root/reference).

Run in the build container only (``python tests/golden/make_golden.py``); the GPU box has
no /root/reference, so the outputs (``tests/golden/*.json`` / ``*.npz``) are committed.
The reference is imported unmodified with its GUI-only imports stubbed;
objects are built exactly like presets/main_3wrobot_NI.py:214-316 (and the 3wrobot / 2tank
presets) do.  Nothing here is product code.
"""
import json
import os
import sys
import types
import warnings

import numpy as np

warnings.simplefilter("ignore")
for _name in ["matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpldatacursor", "svgpath2mpl"]:
    sys.modules[_name] = types.ModuleType(_name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["mpldatacursor"].datacursor = lambda *a, **k: None
sys.modules["svgpath2mpl"].parse_path = lambda *a, **k: None
REF = os.environ.get("RCOGNITA_REF", "/root/reference")
sys.path.insert(0, REF)
from rcognita import controllers, simulator, systems  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# Preset constants (presets/main_3wrobot_NI.py:186-211, main_3wrobot.py:207-215, main_2tank.py:187-211)
SYSTEMS = {
    "3wrobotNI": dict(cls="Sys3WRobotNI", n=3, m=2, pars=[], bnds=[[-25, 25], [-5, 5]],
                      x0=[5, 5, -3 * np.pi / 4], R1_diag=[1, 10, 1, 0, 0], dt=0.01, psm=1.0, target=[]),
    "3wrobot": dict(cls="Sys3WRobot", n=5, m=2, pars=[10, 1], bnds=[[-300, 300], [-100, 100]],
                    x0=[5, 5, -3 * np.pi / 4, 0.3, -0.2], R1_diag=[1, 10, 1, 0, 0, 0, 0], dt=0.01, psm=2.0,
                    target=[]),
    "2tank": dict(cls="Sys2Tank", n=2, m=1, pars=[18.4, 24.4, 1.3, 1, 0.2], bnds=[[0, 1]],
                  x0=[2, -2], R1_diag=[10, 10, 1], dt=0.1, psm=2.0, target=[0.5, 0.5]),
}


def make_sys(name):
    cfg = SYSTEMS[name]
    cls = getattr(systems, cfg["cls"])
    return cls(sys_type="diff_eqn", dim_state=cfg["n"], dim_input=cfg["m"], dim_output=cfg["n"],
               dim_disturb=2 if cfg["m"] == 2 else 1, pars=list(cfg["pars"]),
               ctrl_bnds=np.array(cfg["bnds"], dtype=float), is_dyn_ctrl=0, is_disturb=0,
               pars_disturb=[])


def make_ctrl(name, my_sys, mode, Nactor, critic_struct="quad-nomix", gamma=1.0, R1=None, R2=None,
              stage_obj_struct="quadratic", target=None, state_sys=None, pred_step=None, buffer_size=10,
              Ncritic=4, action_init=(), critic_period=None):
    cfg = SYSTEMS[name]
    R1 = np.diag(np.array(cfg["R1_diag"], dtype=float)) if R1 is None else R1
    pars = [R1] if R2 is None else [R1, R2]
    tgt = cfg["target"] if target is None else target          # python LIST (numpy-2 quirk)
    return controllers.CtrlOptPred(
        cfg["m"], cfg["n"], mode, ctrl_bnds=np.array(cfg["bnds"], dtype=float), action_init=action_init,
        t0=0, sampling_time=cfg["dt"], Nactor=Nactor,
        pred_step_size=cfg["dt"] * cfg["psm"] if pred_step is None else pred_step,
        sys_rhs=my_sys._state_dyn, sys_out=my_sys.out,
        state_sys=np.array(cfg["x0"], dtype=float) if state_sys is None else state_sys,
        prob_noise_pow=False, is_est_model=0, model_est_stage=1.0, model_est_period=cfg["dt"],
        buffer_size=buffer_size, model_order=5, model_est_checks=0, gamma=gamma, Ncritic=Ncritic,
        critic_period=cfg["dt"] if critic_period is None else critic_period, critic_struct=critic_struct, stage_obj_struct=stage_obj_struct,
        stage_obj_pars=pars, observation_target=tgt)


def make_sim(name, my_sys, t1, x0=None):
    cfg = SYSTEMS[name]
    x0 = np.array(cfg["x0"], dtype=float) if x0 is None else np.array(x0, dtype=float)
    return simulator.Simulator(sys_type="diff_eqn", closed_loop_rhs=my_sys.closed_loop_rhs, sys_out=my_sys.out,
                               state_init=x0, disturb_init=np.array([0, 0]), action_init=np.zeros(cfg["m"]),
                               t0=0, t1=t1, dt=cfg["dt"], max_step=cfg["dt"] / 2, first_step=1e-6,
                               atol=1e-5, rtol=1e-3, is_disturb=0, is_dyn_ctrl=0)


def L(a):
    return np.asarray(a, dtype=float).tolist()


# --------------------------------------------------------------------------- function level
def gen_functions():
    out = {}
    for name, cfg in SYSTEMS.items():
        n, m, p = cfg["n"], cfg["m"], cfg["n"] + cfg["m"]
        rng = np.random.default_rng(1000 + n * 10 + m)
        my_sys = make_sys(name)
        bn = np.array(cfg["bnds"], dtype=float)
        cases = {"state_dyn": [], "closed_loop_rhs": [], "stage_obj": [], "critic": [], "critic_cost": [],
                 "actor_cost": []}
        for _ in range(8):
            st = rng.uniform(-6, 6, size=n)
            ac = rng.uniform(bn[:, 0], bn[:, 1])
            cases["state_dyn"].append(dict(state=L(st), action=L(ac), out=L(my_sys._state_dyn([], st, ac))))
            ac2 = rng.uniform(2 * bn[:, 0] - 1, 2 * bn[:, 1] + 1)          # some out of bounds -> clip
            my_sys.receive_action(ac2.copy())
            rhs = my_sys.closed_loop_rhs(0.0, st)
            cases["closed_loop_rhs"].append(dict(state=L(st), action=L(ac2), out=L(rhs),
                                                 action_clipped=L(my_sys.action)))
        A = rng.normal(size=(p, p)); R1_dense = A @ A.T
        B = rng.normal(size=(p, p)); R2_dense = B @ B.T
        variants = [
            dict(tag="diag", R1=np.diag(np.array(cfg["R1_diag"], dtype=float)), R2=None, struct="quadratic", target=cfg["target"]),
            dict(tag="dense", R1=R1_dense, R2=None, struct="quadratic", target=cfg["target"]),
            dict(tag="biquad", R1=R1_dense, R2=R2_dense, struct="biquadratic", target=cfg["target"]),
            dict(tag="dense_target", R1=R1_dense, R2=None, struct="quadratic", target=L(rng.uniform(-1, 1, size=n))),
        ]
        for v in variants:
            ctrl = make_ctrl(name, my_sys, "MPC", 3, R1=v["R1"], R2=v["R2"], stage_obj_struct=v["struct"], target=v["target"])
            for _ in range(4):
                ob = rng.uniform(-6, 6, size=n); ac = rng.uniform(bn[:, 0], bn[:, 1])
                cases["stage_obj"].append(dict(R1=L(v["R1"]), R2=None if v["R2"] is None else L(v["R2"]),
                                               struct=v["struct"], target=list(v["target"]), obs=L(ob), act=L(ac),
                                               out=float(ctrl.stage_obj(ob, ac))))
        for cs in ["quad-lin", "quadratic", "quad-nomix", "quad-mix"]:
            for tgt in ([], L(rng.uniform(-1, 1, size=n))):
                for gamma in (1.0, 0.9):
                    ctrl = make_ctrl(name, my_sys, "RQL", 4, critic_struct=cs, gamma=gamma, target=tgt)
                    dimc = ctrl.dim_critic
                    w = rng.uniform(0, 2, size=dimc); w_prev = rng.uniform(0, 2, size=dimc)
                    ob = rng.uniform(-6, 6, size=n); ac = rng.uniform(bn[:, 0], bn[:, 1])
                    cases["critic"].append(dict(critic_struct=cs, target=list(tgt), obs=L(ob), act=L(ac), w=L(w),
                                                dim_critic=int(dimc), out=float(ctrl._critic(ob, ac, w))))
                    ctrl.observation_buffer = rng.normal(size=(10, n))
                    ctrl.action_buffer = rng.uniform(bn[:, 0], bn[:, 1], size=(10, m))
                    ctrl.w_critic_prev = w_prev
                    cases["critic_cost"].append(dict(critic_struct=cs, target=list(tgt), gamma=gamma, Ncritic=int(ctrl.Ncritic),
                                                     R1_diag=cfg["R1_diag"], obs_buf=L(ctrl.observation_buffer),
                                                     act_buf=L(ctrl.action_buffer), w=L(w), w_prev=L(w_prev),
                                                     out=float(ctrl._critic_cost(w))))
        for mode in ["MPC", "RQL", "SQL"]:
            for cs in (["quad-nomix"] if mode == "MPC" else ["quad-lin", "quadratic", "quad-nomix", "quad-mix"]):
                for (N, gamma, dense, tgt) in [(6, 1.0, False, cfg["target"]), (3, 0.9, True, cfg["target"]),
                                               (10, 0.95, False, L(rng.uniform(-1, 1, size=n))), (1, 1.0, False, cfg["target"])]:
                    R1 = R1_dense if dense else None
                    x_sys = rng.uniform(-6, 6, size=n)
                    ob = x_sys + rng.normal(size=n) * 0.01
                    ctrl = make_ctrl(name, my_sys, mode, N, critic_struct=cs, gamma=gamma, R1=R1, target=tgt, state_sys=x_sys)
                    ctrl.w_critic = rng.uniform(0, 2, size=ctrl.dim_critic)
                    U = rng.uniform(ctrl.action_sqn_min, ctrl.action_sqn_max, size=(6, N * m))
                    U[4] = U[1]                                       # exact tie -> first index must win
                    J = [float(ctrl._actor_cost(u, ob)) for u in U]
                    cases["actor_cost"].append(dict(mode=mode, critic_struct=cs, N=N, gamma=gamma,
                                                    R1=L(R1 if dense else np.diag(np.array(cfg["R1_diag"], dtype=float))),
                                                    target=list(tgt), pred_step=float(ctrl.pred_step_size),
                                                    state_sys=L(x_sys), obs=L(ob), w=L(ctrl.w_critic), cand=L(U), J=J,
                                                    argmin=int(np.argmin(J))))
        out[name] = dict(pars=cfg["pars"], bnds=cfg["bnds"], cases=cases)
    with open(os.path.join(HERE, "functions.json"), "w") as fh:
        json.dump(out, fh)
    print("functions.json", {k: {c: len(v) for c, v in d["cases"].items()} for k, d in out.items()})


# --------------------------------------------------------------------------- integrator only (App. A.3)
SCHEDULES = {
    "3wrobotNI": dict(t1=0.25, sched=[[-2.5, -0.5], [25, 5], [-25, 5], [10, -5]]),
    "3wrobot": dict(t1=0.25, sched=[[-30, -10], [300, 100], [-300, 100], [120, -100]]),
    "2tank": dict(t1=2.5, sched=[[0.5], [1.0], [0.0], [0.3]]),
}


def gen_integrator():
    out = {}
    for name, sc in SCHEDULES.items():
        for variant in ("inbounds", "outofbounds"):
            my_sys = make_sys(name)
            sim = make_sim(name, my_sys, sc["t1"])
            sched = np.array(sc["sched"], dtype=float) * (1.0 if variant == "inbounds" else 1.7)
            rows = []
            k = 0
            while True:
                sim.sim_step()
                k += 1
                my_sys.receive_action(sched[(k // 5) % 4].copy())
                s = sim.ODE_solver
                rows.append([s.t] + L(s.y) + L(s.f) + [float(s.h_abs), int(s.nfev)])
                if s.status != "running":
                    break
            out[f"{name}:{variant}"] = dict(system=name, t1=sc["t1"], sched=L(sched), rows=rows,
                                            status=sim.ODE_solver.status)
            print("integrator", name, variant, len(rows), "steps, nfev", rows[-1][-1])
    with open(os.path.join(HERE, "integrator.json"), "w") as fh:
        json.dump(out, fh)


# --------------------------------------------------------------------------- closed loop, candidate/arg-min controller (App. A.4)
def closed_loop(name, mode, Nactor, t1, critic_struct="quad-nomix", gamma=1.0, w_fixed=None, C=256, seed=1,
                x0=None, action_init=()):
    cfg = SYSTEMS[name]
    my_sys = make_sys(name)
    x0 = np.array(cfg["x0"], dtype=float) if x0 is None else np.array(x0, dtype=float)
    ctrl = make_ctrl(name, my_sys, mode, Nactor, critic_struct=critic_struct, gamma=gamma, state_sys=x0,
                     action_init=action_init)
    sim = make_sim(name, my_sys, t1, x0=x0)
    Ctab = np.random.default_rng(seed).uniform(ctrl.action_sqn_min, ctrl.action_sqn_max, size=(C, Nactor * cfg["m"]))
    picks = []

    def opt(observation):
        J = [ctrl._actor_cost(u, observation) for u in Ctab]
        i = int(np.argmin(J))
        picks.append([i, float(J[i])])
        return Ctab[i, :cfg["m"]].copy()

    ctrl._actor_optimizer = opt
    if w_fixed is not None:
        ctrl._critic_optimizer = lambda: np.array(w_fixed, dtype=float)
    rows = []
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        npk = len(picks)
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, mode)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        rows.append([t] + L(state_full) + L(action) + [float(ctrl.accum_obj_val), int(len(picks) > npk)])
        if t >= t1:
            break
    return dict(system=name, mode=mode, Nactor=Nactor, t1=t1, critic_struct=critic_struct, gamma=gamma,
                w_fixed=None if w_fixed is None else L(w_fixed), C=C, seed=seed, x0=L(x0),
                action_init=L(ctrl.action_min / 10) if len(action_init) == 0 else L(action_init),
                cand=L(Ctab), rows=rows, picks=picks, nfev=int(sim.ODE_solver.nfev))


def gen_closed_loop():
    out = {}
    out["NI_MPC_N6"] = closed_loop("3wrobotNI", "MPC", 6, 2.0)
    out["NI_MPC_N6_x1"] = closed_loop("3wrobotNI", "MPC", 6, 1.0, x0=[-3.0, 7.5, 1.1], seed=7, C=64)
    out["3wrobot_RQL_N10"] = closed_loop("3wrobot", "RQL", 10, 1.0, critic_struct="quadratic",
                                         w_fixed=np.arange(1, 29) / 10.0)
    out["2tank_SQL_N8"] = closed_loop("2tank", "SQL", 8, 20.0, critic_struct="quad-nomix", w_fixed=[11.0, 11.0, 1.0],
                                      action_init=0.5 * np.ones(1))
    for k, v in out.items():
        print("closed_loop", k, len(v["rows"]), "steps", len(v["picks"]), "samples, nfev", v["nfev"],
              "accum", v["rows"][-1][-2])
    with open(os.path.join(HERE, "closed_loop.json"), "w") as fh:
        json.dump(out, fh)




# --------------------------------------------------------------------------- closed loop WITH the reference's own critic refit (a18)
def closed_loop_refit(name, mode, Nactor, t1, critic_struct, gamma=1.0, C=64, seed=1, x0=None, action_init=(),
                      critic_period=None, buffer_size=10, Ncritic=4):
    """The RQL/SQL branch of compute_action (controllers.py:1455-1479) run by the UNMODIFIED reference, SLSQP
    `_critic_optimizer` included; only `_actor_optimizer` is the candidate/arg-min stand-in of App. A.4.  Every
    refit is recorded (solver time, FIFO buffers after the push, w_critic_prev before the fit, the fitted weights,
    `_critic_cost` at w_critic_init and at the fit) so that a loop which LOADS the recorded weights instead of fitting must
    reproduce buffers, critic-clock firings, arg-min picks, trajectory and accumulated objective."""
    cfg = SYSTEMS[name]
    my_sys = make_sys(name)
    x0 = np.array(cfg["x0"], dtype=float) if x0 is None else np.array(x0, dtype=float)
    ctrl = make_ctrl(name, my_sys, mode, Nactor, critic_struct=critic_struct, gamma=gamma, state_sys=x0,
                     action_init=action_init, critic_period=critic_period, buffer_size=buffer_size, Ncritic=Ncritic)
    sim = make_sim(name, my_sys, t1, x0=x0)
    Ctab = np.random.default_rng(seed).uniform(ctrl.action_sqn_min, ctrl.action_sqn_max, size=(C, Nactor * cfg["m"]))
    picks, fits = [], []

    def opt(observation):
        J = [ctrl._actor_cost(u, observation) for u in Ctab]
        i = int(np.argmin(J))
        picks.append([i, float(J[i])])
        return Ctab[i, :cfg["m"]].copy()

    ctrl._actor_optimizer = opt
    orig_fit = ctrl._critic_optimizer
    now = [0.0]

    def fit():
        w_prev = np.array(ctrl.w_critic_prev, dtype=float)
        w = orig_fit()
        fits.append(dict(t=float(now[0]), obs_buf=L(ctrl.observation_buffer), act_buf=L(ctrl.action_buffer), w_prev=L(w_prev),
                         w=L(w), J_init=float(ctrl._critic_cost(ctrl.w_critic_init)), J_fit=float(ctrl._critic_cost(w))))
        return w

    ctrl._critic_optimizer = fit
    rows = []
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        now[0] = t
        npk, nft = len(picks), len(fits)
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, mode)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        rows.append([t] + L(state_full) + L(action) + [float(ctrl.accum_obj_val), int(len(picks) > npk), int(len(fits) > nft)])
        if t >= t1:
            break
    return dict(system=name, mode=mode, Nactor=Nactor, t1=t1, critic_struct=critic_struct, gamma=gamma, C=C, seed=seed,
                x0=L(x0), action_init=L(ctrl.action_min / 10) if len(action_init) == 0 else L(action_init),
                critic_period=float(ctrl.critic_period), buffer_size=int(buffer_size), Ncritic=int(ctrl.Ncritic),
                cand=L(Ctab), rows=rows, picks=picks, fits=fits, nfev=int(sim.ODE_solver.nfev),
                w_final=L(ctrl.w_critic), w_prev_final=L(ctrl.w_critic_prev),
                obs_buf_final=L(ctrl.observation_buffer), act_buf_final=L(ctrl.action_buffer))


def gen_closed_loop_refit():
    out = {}
    # BASELINE configs 3 and 4 at E = 1
    out["3wrobot_RQL_quadratic_N10"] = closed_loop_refit("3wrobot", "RQL", 10, 0.6, "quadratic")
    out["2tank_SQL_nomix_N8"] = closed_loop_refit("2tank", "SQL", 8, 12.0, "quad-nomix", action_init=0.5 * np.ones(1))
    # critic clock slower than the sampling clock: the `w_critic = w_critic_prev` branch (:1478-1479); gamma < 1
    out["NI_RQL_quadlin_N5_period3"] = closed_loop_refit("3wrobotNI", "RQL", 5, 0.5, "quad-lin", gamma=0.95,
                                                         critic_period=0.03, x0=[2.0, -3.0, 0.7])
    out["NI_SQL_quadmix_N3_Ncritic6"] = closed_loop_refit("3wrobotNI", "SQL", 3, 0.4, "quad-mix", critic_period=0.02,
                                                          buffer_size=8, Ncritic=6, x0=[-1.0, 2.0, -2.0])
    for k, v in out.items():
        print("closed_loop_refit", k, len(v["rows"]), "steps", len(v["picks"]), "samples", len(v["fits"]), "fits, accum",
              v["rows"][-1][-3], "J_fit range", min(f["J_fit"] for f in v["fits"]), max(f["J_fit"] for f in v["fits"]))
    with open(os.path.join(HERE, "closed_loop_refit.json"), "w") as fh:
        json.dump(out, fh)



# --------------------------------------------------------------------------- disturbance lanes (is_disturb = 1), function level
def gen_disturb():
    """`_state_dyn(t, state, action, disturb)` with a disturbance and `_disturb_dyn(t, disturb)` of the UNMODIFIED reference
    (systems.py:316-318, :373-376, :341-343, :390-392, :421-424).  Two workarounds, neither touches the arithmetic:
    `disturb` is passed as a Python LIST (under numpy 2 `ndarray != []` raises, which is also why the reference's own
    closed loop cannot run with is_disturb = 1 and no loop-level golden exists), and `systems.randn` is replaced by a
    function that replays recorded draws (the reference takes them from numpy's global stream)."""
    out = {}
    rng = np.random.default_rng(2024)
    for name, cfg in SYSTEMS.items():
        n, m = cfg["n"], cfg["m"]
        nd = 2 if m == 2 else 1
        dt = cfg["dt"]
        pars_disturb = np.array([[200 * dt, 150 * dt], [0.0, 0.1], [0.3, 0.45]])[:, :nd] if nd == 2 else np.array([[0.2], [0.0], [0.3]])
        cls = getattr(systems, cfg["cls"])
        my_sys = cls(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n, dim_disturb=nd, pars=list(cfg["pars"]),
                     ctrl_bnds=np.array(cfg["bnds"], dtype=float), is_dyn_ctrl=0, is_disturb=1, pars_disturb=pars_disturb)
        assert my_sys._dim_full_state == n + nd
        bn = np.array(cfg["bnds"], dtype=float)
        cases = []
        for _ in range(24):
            state = rng.uniform(-3, 3, size=n)
            action = rng.uniform(bn[:, 0], bn[:, 1])
            disturb = rng.normal(size=nd) * 2.0
            z = rng.normal(size=nd)
            d_state = my_sys._state_dyn([], state, action, disturb=list(disturb))
            it = iter(z)
            systems.randn = lambda: float(next(it))
            d_dist = my_sys._disturb_dyn([], list(disturb))
            cases.append(dict(state=L(state), action=L(action), disturb=L(disturb), z=L(z), d_state=L(d_state), d_disturb=L(d_dist)))
        out[name] = dict(pars=list(cfg["pars"]), bnds=cfg["bnds"], pars_disturb=L(pars_disturb), dim_disturb=nd,
                         dim_full_state=int(my_sys._dim_full_state), cases=cases)
        print("disturb", name, len(cases), "cases; d_disturb[0]", cases[0]["d_disturb"])
    with open(os.path.join(HERE, "disturb.json"), "w") as fh:
        json.dump(out, fh)

# --------------------------------------------------------------------------- critic fit (reference SLSQP as the bar)
def gen_critic_fit():
    """Reference `_critic_optimizer` (SLSQP, controllers.py:1248-1271) on seeded buffers: the fitted cost is
    the bar the batched bounded-least-squares fit must reach (the minimiser itself is not unique)."""
    out = []
    for name, cfg in SYSTEMS.items():
        n, m = cfg["n"], cfg["m"]
        bn = np.array(cfg["bnds"], dtype=float)
        my_sys = make_sys(name)
        rng = np.random.default_rng(4242 + n)
        for cs in ["quad-lin", "quadratic", "quad-nomix", "quad-mix"]:
            for regime in ["random", "trajectory", "cold", "random_wprev"]:
                for gamma in (1.0, 0.9):
                    ctrl = make_ctrl(name, my_sys, "RQL", 4, critic_struct=cs, gamma=gamma)
                    if regime in ("random", "random_wprev"):
                        ob = rng.normal(size=(10, n)); ac = rng.uniform(bn[:, 0], bn[:, 1], size=(10, m))
                    elif regime == "trajectory":       # consecutive samples of a smooth run: nearly collinear rows
                        x = rng.uniform(-5, 5, size=n); v = rng.normal(size=n) * 0.05
                        ob = np.array([x + k * v for k in range(10)])
                        a0 = rng.uniform(bn[:, 0], bn[:, 1])
                        ac = np.array([a0 * (1 - 0.02 * k) for k in range(10)])
                    else:                              # first samples of an episode: zero rows at the top
                        ob = np.zeros((10, n)); ac = np.zeros((10, m))
                        ob[2:] = rng.normal(size=(8, n)); ac[2:] = rng.uniform(bn[:, 0], bn[:, 1], size=(8, m))
                    ctrl.observation_buffer, ctrl.action_buffer = ob, ac
                    if regime == "random_wprev":
                        ctrl.w_critic_prev = rng.uniform(0, 3, size=ctrl.dim_critic)
                    w_ref = ctrl._critic_optimizer()
                    out.append(dict(system=name, critic_struct=cs, regime=regime, gamma=gamma, Ncritic=int(ctrl.Ncritic),
                                    target=list(cfg["target"]), R1_diag=cfg["R1_diag"], obs_buf=L(ob), act_buf=L(ac),
                                    w_prev=L(ctrl.w_critic_prev), w_init=L(ctrl.w_critic_init), Wmin=float(ctrl.Wmin[0]),
                                    Wmax=float(ctrl.Wmax[0]), w_ref=L(w_ref), J_ref=float(ctrl._critic_cost(w_ref)),
                                    J_init=float(ctrl._critic_cost(ctrl.w_critic_init))))
    with open(os.path.join(HERE, "critic_fit.json"), "w") as fh:
        json.dump(out, fh)
    print("critic_fit.json", len(out), "cases; J_ref range", min(c["J_ref"] for c in out), max(c["J_ref"] for c in out))

# --------------------------------------------------------------------------- actor optimiser (reference SLSQP as the bar)
def gen_actor_opt():
    """Reference `_actor_optimizer` (SLSQP on `_actor_cost`, controllers.py:1330-1427) on seeded problems.  The
    method returns only the first action; `controllers.minimize` is wrapped to also capture the full minimiser
    and its cost (the call itself is the reference's, unmodified).  The minimum found is the bar for the batched
    projected-gradient optimiser (the minimiser need not be unique: the last action never enters an MPC cost)."""
    out = []
    rec = {}
    orig_min = controllers.minimize

    def wrapped(fun, x0, **kw):
        res = orig_min(fun, x0, **kw)
        rec["res"] = res
        return res

    controllers.minimize = wrapped
    horizons = {"3wrobotNI": 6, "3wrobot": 10, "2tank": 8}
    try:
        for name, cfg in SYSTEMS.items():
            n, m, p = cfg["n"], cfg["m"], cfg["n"] + cfg["m"]
            my_sys = make_sys(name)
            rng = np.random.default_rng(777 + n)
            A = rng.normal(size=(p, p)); R1_dense = A @ A.T / p
            N0 = horizons[name]
            plan = []
            for scale in (6.0, 6.0, 1.0, 0.2, 0.05, 0.01):                      # far from / near the goal
                plan.append(dict(mode="MPC", cs="quad-nomix", N=N0, gamma=1.0, dense=False, scale=scale))
            plan.append(dict(mode="MPC", cs="quad-nomix", N=3, gamma=0.9, dense=True, scale=2.0))
            plan.append(dict(mode="MPC", cs="quad-nomix", N=12, gamma=0.95, dense=False, scale=0.5))
            for mode in ("RQL", "SQL"):
                for cs in ("quad-lin", "quadratic", "quad-nomix", "quad-mix"):
                    plan.append(dict(mode=mode, cs=cs, N=N0, gamma=1.0, dense=False, scale=3.0))
                    plan.append(dict(mode=mode, cs=cs, N=5, gamma=0.9, dense=False, scale=0.3))
            for pl in plan:
                tgt = cfg["target"]
                x_sys = rng.uniform(-1, 1, size=n) * pl["scale"] + (np.array(tgt) if len(tgt) else 0.0)
                ob = x_sys + rng.normal(size=n) * 0.005 * pl["scale"]
                R1 = R1_dense if pl["dense"] else None
                ctrl = make_ctrl(name, my_sys, pl["mode"], pl["N"], critic_struct=pl["cs"], gamma=pl["gamma"], R1=R1,
                                 state_sys=x_sys)
                if pl["cs"] in ("quad-lin", "quad-mix"):
                    ctrl.w_critic = rng.uniform(-1, 2, size=ctrl.dim_critic)
                else:
                    ctrl.w_critic = rng.uniform(0, 2, size=ctrl.dim_critic)
                a1 = ctrl._actor_optimizer(ob)
                res = rec["res"]
                x_ref = np.asarray(res.x, dtype=float)
                out.append(dict(system=name, mode=pl["mode"], critic_struct=pl["cs"], N=pl["N"], gamma=pl["gamma"],
                                R1=L(R1_dense if pl["dense"] else np.diag(np.array(cfg["R1_diag"], dtype=float))),
                                target=list(tgt), pred_step=float(ctrl.pred_step_size), state_sys=L(x_sys), obs=L(ob),
                                w=L(ctrl.w_critic), x_init=L(ctrl.action_sqn_init), x_ref=L(x_ref),
                                J_ref=float(ctrl._actor_cost(x_ref, ob)),
                                J_init=float(ctrl._actor_cost(np.asarray(ctrl.action_sqn_init, dtype=float), ob)),
                                nit=int(res.nit), nfev=int(res.nfev), status=int(res.status), first_action=L(a1)))
    finally:
        controllers.minimize = orig_min
    with open(os.path.join(HERE, "actor_opt.json"), "w") as fh:
        json.dump(out, fh)
    print("actor_opt.json", len(out), "cases; nfev", min(c["nfev"] for c in out), "..", max(c["nfev"] for c in out),
          "status", sorted(set(c["status"] for c in out)))

# --------------------------------------------------------------------------- nominal parking controller (NI)
def gen_nominal():
    """CtrlNominal3WRobotNI (controllers.py:1758-1956), the default ctrl_mode of presets/main_3wrobot_NI.py:
    actions for seeded observations (incl. the xNI[0] = xNI[1] = 0 branch and out-of-bounds results) and the
    preset-default closed loop (ctrl_gain 0.5 -- presets/main_3wrobot_NI.py:231 -- dt 0.01, t1 3)."""
    name = "3wrobotNI"
    cfg = SYSTEMS[name]
    bn = np.array(cfg["bnds"], dtype=float)
    rng = np.random.default_rng(99)
    cases = []
    for gain in (0.5, 10.0):
        nom = controllers.CtrlNominal3WRobotNI(ctrl_gain=gain, ctrl_bnds=bn, t0=0, sampling_time=cfg["dt"])
        obs_list = [rng.uniform([-10, -10, -np.pi], [10, 10, np.pi]) for _ in range(24)]
        obs_list += [rng.uniform([-0.1, -0.1, -0.1], [0.1, 0.1, 0.1]) for _ in range(6)]
        obs_list += [np.array([0.0, 1.5, 0.0]), np.array([0.0, -0.3, 0.0])]            # xNI[0] = xNI[1] = 0
        for ob in obs_list:
            act = nom.compute_action(1.0 + len(cases), ob)          # clock always fires: sampling_time elapsed
            cases.append(dict(gain=gain, obs=L(ob), action=L(act)))
    my_sys = make_sys(name)
    sim = make_sim(name, my_sys, 3.0)
    nom = controllers.CtrlNominal3WRobotNI(ctrl_gain=0.5, ctrl_bnds=bn, t0=0, sampling_time=cfg["dt"])
    ctrl = make_ctrl(name, my_sys, "MPC", 3)                        # the preset's accumulator (upd_accum_obj)
    rows = []
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        action = controllers.ctrl_selector(t, observation, None, nom, ctrl, "nominal")
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        rows.append([t] + L(state_full) + L(action) + [float(ctrl.accum_obj_val)])
        if t >= 3.0:
            break
    with open(os.path.join(HERE, "nominal.json"), "w") as fh:
        json.dump(dict(cases=cases, episode=dict(gain=0.5, t1=3.0, x0=L(cfg["x0"]), rows=rows,
                                                 nfev=int(sim.ODE_solver.nfev))), fh)
    print("nominal.json", len(cases), "cases;", len(rows), "episode steps; final", rows[-1])

# --------------------------------------------------------------------------- config 1: preset-faithful SLSQP episode (App. A.2)
def gen_config1():
    name = "3wrobotNI"
    cfg = SYSTEMS[name]
    my_sys = make_sys(name)
    x0 = np.array(cfg["x0"], dtype=float)
    ctrl = make_ctrl(name, my_sys, "MPC", 6, state_sys=x0)
    sim = make_sim(name, my_sys, 10.0)
    ncost = [0]
    orig = ctrl._actor_cost

    def counted(a, o):
        ncost[0] += 1
        return orig(a, o)

    ctrl._actor_cost = counted
    rows = []
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, "MPC")
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        s = sim.ODE_solver
        rows.append([t] + L(state_full) + L(action) + [float(ctrl.stage_obj(observation, action)),
                                                       float(ctrl.accum_obj_val), float(s.h_abs), float(s.nfev)])
        if t >= 10.0:
            break
    rows = np.array(rows)
    np.savez_compressed(os.path.join(HERE, "config1_slsqp_episode.npz"), rows=rows,
                        columns=np.array(["t", "x", "y", "theta", "a0", "a1", "stage_obj", "accum_obj", "h_abs", "nfev"]),
                        actor_cost_calls=np.array(ncost[0]))
    print("config1", rows.shape, "final", rows[-1], "actor_cost calls", ncost[0])


if __name__ == "__main__":
    which = sys.argv[1:] or ["functions", "integrator", "closed_loop", "closed_loop_refit", "disturb", "config1", "critic_fit",
                             "actor_opt", "nominal"]
    if "functions" in which:
        gen_functions()
    if "integrator" in which:
        gen_integrator()
    if "closed_loop" in which:
        gen_closed_loop()
    if "closed_loop_refit" in which:
        gen_closed_loop_refit()
    if "disturb" in which:
        gen_disturb()
    if "config1" in which:
        gen_config1()
    if "critic_fit" in which:
        gen_critic_fit()
    if "actor_opt" in which:
        gen_actor_opt()
    if "nominal" in which:
        gen_nominal()
