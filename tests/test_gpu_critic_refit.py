"""SURVEY.md section 8 row a18, the RQL / SQL branch of ``CtrlOptPred.compute_action`` WITH critic refit
(rcognita/controllers.py:1455-1479, :1248-1271), pinned to the live reference and to the CPU oracle:

* ``tests/golden/closed_loop_refit.json`` is the unmodified reference's loop with its own SLSQP ``_critic_optimizer``
  (BASELINE configs 3 and 4 at E = 1, plus a slow critic clock and a deeper critic stack).  With the fit replaced by
  "load the recorded weights", the drop-in class loop AND the fused engine must reproduce the reference's FIFO buffers,
  critic-clock firings, ``w_critic_prev`` hand-over, arg-min picks, trajectory and accumulated objective.
* ``rcg_critic_fit`` (generic, K <= 3 one-lane and warp-per-environment kernels) against ``oracle.critic_fit``, the
  scalar C restatement of the same algorithm, on the 96 recorded problems and on every in-loop problem above.
* a 64-environment closed loop with the product's own refit against the oracle's loop with the restated fit.
"""
import os

import numpy as np
import pytest

from golden_util import DIMS, PRESET, load, mixed_err, rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CASES = ["3wrobot_RQL_quadratic_N10", "2tank_SQL_nomix_N8", "NI_RQL_quadlin_N5_period3", "NI_SQL_quadmix_N3_Ncritic6"]
SYS_CLS = {"3wrobotNI": "Sys3WRobotNI", "3wrobot": "Sys3WRobot", "2tank": "Sys2Tank"}


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


def w_bounds(cs):
    return (-1e3, 1e3) if cs in ("quad-lin", "quad-mix") else (0.0, 1e3)


@pytest.mark.parametrize("key", CASES)
def test_class_loop_with_recorded_weights_reproduces_reference(cuda, key):
    """presets/main_3wrobot_NI.py:415-440 verbatim over the drop-in classes, one environment, numpy in / numpy out."""
    from rcognita_b200 import controllers, simulator, systems
    g = load("closed_loop_refit.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    cfg = PRESET[name]
    bnds = np.array(cfg["bnds"], dtype=float)
    x0 = np.array(g["x0"])
    my_sys = getattr(systems, SYS_CLS[name])(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n,
                                              dim_disturb=2 if m == 2 else 1, pars=list(cfg["pars"]), ctrl_bnds=bnds,
                                              is_dyn_ctrl=0, is_disturb=0, pars_disturb=[])
    ctrl = controllers.CtrlOptPred(m, n, g["mode"], ctrl_bnds=bnds, action_init=g["action_init"], t0=0, sampling_time=cfg["dt"],
                                   Nactor=g["Nactor"], pred_step_size=cfg["dt"] * cfg["psm"], sys_rhs=my_sys._state_dyn,
                                   sys_out=my_sys.out, state_sys=x0, prob_noise_pow=False, is_est_model=0,
                                   buffer_size=g["buffer_size"], gamma=g["gamma"], Ncritic=g["Ncritic"],
                                   critic_period=g["critic_period"], critic_struct=g["critic_struct"],
                                   stage_obj_struct="quadratic", stage_obj_pars=[np.diag(np.array(cfg["R1_diag"], dtype=float))],
                                   observation_target=cfg["target"], candidates=np.array(g["cand"]), actor="candidates")
    sim = simulator.Simulator(sys_type="diff_eqn", closed_loop_rhs=my_sys.closed_loop_rhs, sys_out=my_sys.out,
                              state_init=x0, disturb_init=np.array([0, 0]), action_init=np.zeros(m), t0=0, t1=g["t1"],
                              dt=cfg["dt"], max_step=cfg["dt"] / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0,
                              is_dyn_ctrl=0)
    fits = g["fits"]
    nfit = [0]
    seen = []

    def recorded_fit(mask=None):
        w = ctrl._w_critic.clone()
        if mask is not None and int(mask[0].item()):
            f = fits[nfit[0]]
            # what the reference's optimiser saw at this refit: buffers after the push, w_critic_prev before the fit
            seen.append((ctrl.observation_buffer.copy(), ctrl.action_buffer.copy(), ctrl.w_critic_prev.copy()))
            w[:, 0] = torch.as_tensor(np.array(f["w"]), device=w.device)
            nfit[0] += 1
        return w

    ctrl._critic_optimizer = recorded_fit
    rows = np.array(g["rows"])
    got = []
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        k0, f0 = int(ctrl.num_samples[0].item()), nfit[0]
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, g["mode"])
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        got.append([t] + list(state_full) + list(np.atleast_1d(action)) + [ctrl.accum_obj_val,
                   int(ctrl.num_samples[0].item()) - k0, nfit[0] - f0])
        if t >= g["t1"]:
            break
    got = np.array(got, dtype=np.float64)
    assert got.shape == rows.shape
    assert np.max(np.abs(got[:, 0] - rows[:, 0])) <= 1e-15 * g["t1"]
    assert mixed_err(got[:, 1:1 + n], rows[:, 1:1 + n], 1e-2) <= 1e-9
    assert np.array_equal(got[:, 1 + n:1 + n + m], rows[:, 1 + n:1 + n + m]), "applied actions differ"
    assert rel_err(got[:, 1 + n + m], rows[:, 1 + n + m]) <= 1e-9
    assert np.array_equal(got[:, -2], rows[:, -2]), "controller sampling steps differ"
    assert np.array_equal(got[:, -1], rows[:, -1]), "critic-clock firings differ"
    assert nfit[0] == len(fits) == len(seen)
    for (ob, ab, wp), f in zip(seen, fits):
        assert mixed_err(ob, f["obs_buf"], 1e-2) <= 1e-9
        assert np.array_equal(ab, np.array(f["act_buf"]))
        assert np.array_equal(wp, np.array(f["w_prev"]))
    assert np.array_equal(ctrl.w_critic, np.array(g["w_final"])) and np.array_equal(ctrl.w_critic_prev, np.array(g["w_prev_final"]))


@pytest.mark.parametrize("key", CASES)
def test_engine_with_recorded_weights_reproduces_reference(cuda, key):
    """The fused engine (rk45_advance + push_buffers + ctrl_sample + [fit] + actor_cost per control interval) with its
    `_critic_optimizer` replaced by the recorded weights: trajectory ring of the environment == the reference's rows."""
    from rcognita_b200.engine import ClosedLoopEngine
    g = load("closed_loop_refit.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    cfg = PRESET[name]
    rows = np.array(g["rows"])
    eng = ClosedLoopEngine(name, [g["x0"]], np.array(g["cand"]), pars=cfg["pars"], ctrl_bnds=cfg["bnds"], mode=g["mode"],
                           Nactor=g["Nactor"], dt=cfg["dt"], pred_step_size=cfg["dt"] * cfg["psm"], t1=g["t1"],
                           R1=cfg["R1_diag"], observation_target=cfg["target"], critic_struct=g["critic_struct"],
                           gamma=g["gamma"], critic_fit=True, Ncritic=g["Ncritic"], buffer_size=g["buffer_size"],
                           critic_period=g["critic_period"], action_init=g["action_init"], log_every=1,
                           log_capacity=len(rows) + 8)
    fits = g["fits"]
    nfit = [0]
    seen = []

    def recorded_fit():
        if int(eng.critic_flag[0].item()):
            f = fits[nfit[0]]
            seen.append((eng.obs_buf[:, :, 0].cpu().numpy(), eng.act_buf[:, :, 0].cpu().numpy(), eng.w_prev[:, 0].cpu().numpy()))
            w = torch.as_tensor(np.array(f["w"]), device=eng.w.device)
            eng.w[:, 0] = w
            eng.w_prev[:, 0] = w
            nfit[0] += 1

    eng._critic_optimizer = recorded_fit
    eng.run()
    tr = eng.trajectory(0)                      # reference logger column order
    if name == "2tank":                         # t, h1, h2, p, stage_obj, accum_obj -> t, state, action, accum
        got = np.concatenate([tr[:, 0:3], tr[:, 3:4], tr[:, 5:6]], axis=1)
    else:                                       # t, state[n], stage_obj, accum_obj, action[m]
        got = np.concatenate([tr[:, 0:1 + n], tr[:, 3 + n:3 + n + m], tr[:, 2 + n:3 + n]], axis=1)
    assert got.shape[0] == rows.shape[0]
    assert np.max(np.abs(got[:, 0] - rows[:, 0])) <= 1e-15 * g["t1"]
    assert mixed_err(got[:, 1:1 + n], rows[:, 1:1 + n], 1e-2) <= 1e-9
    assert np.array_equal(got[:, 1 + n:1 + n + m], rows[:, 1 + n:1 + n + m]), "applied actions differ"
    assert rel_err(got[:, 1 + n + m], rows[:, 1 + n + m]) <= 1e-9
    res = eng.results()
    assert int(res["nsamples"][0]) == len(g["picks"]) and int(res["nfits"][0]) == len(fits) == nfit[0]
    for (ob, ab, wp), f in zip(seen, fits):
        assert mixed_err(ob, f["obs_buf"], 1e-2) <= 1e-9
        assert np.array_equal(ab, np.array(f["act_buf"]))
        assert np.array_equal(wp, np.array(f["w_prev"]))
    assert np.array_equal(res["w_critic"][0], np.array(g["w_final"]))
    assert int(res["argmin"][0]) == g["picks"][-1][0]


def _fit_problems():
    """(config key, problem) for every recorded fit: the 96 seeded problems + the reference's in-loop refits."""
    out = {}
    for c in load("critic_fit.json"):
        key = (c["system"], c["critic_struct"], c["gamma"], c["Ncritic"], 10, tuple(c["target"]), tuple(c["R1_diag"]))
        out.setdefault(key, []).append((np.array(c["obs_buf"]), np.array(c["act_buf"]), np.array(c["w_prev"]), c["J_ref"], c["J_init"]))
    for g in load("closed_loop_refit.json").values():
        cfg = PRESET[g["system"]]
        key = (g["system"], g["critic_struct"], g["gamma"], g["Ncritic"], g["buffer_size"], tuple(cfg["target"]), tuple(cfg["R1_diag"]))
        for f in g["fits"]:
            out.setdefault(key, []).append((np.array(f["obs_buf"]), np.array(f["act_buf"]), np.array(f["w_prev"]), f["J_fit"], f["J_init"]))
    return out


@pytest.mark.parametrize("variant", ["default", "one_phase", "budget24_then_rest"])
def test_critic_fit_kernels_against_oracle_fit(cuda, variant):
    """Every kernel behind rcg_critic_fit against the scalar restatement of the same algorithm: fitted cost to 1e-6
    relative (of max(J, 1e-9 J_init)) and never above the reference's SLSQP cost."""
    import oracle
    from rcognita_b200 import _C, ops
    env = {"default": {}, "one_phase": {"RCG_FIT_ONE_PHASE": "1"}, "budget24_then_rest": {}}[variant]
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        total, worse_than_ref, off = 0, [], []
        for key, probs in _fit_problems().items():
            name, cs, gamma, Ncritic, bufsz, target, R1d = key
            n, m = DIMS[name]
            obj = _C.make_objective(n, m, mode="RQL", Nactor=4, gamma=gamma, Ncritic=Ncritic, buffer_size=bufsz, critic_struct=cs,
                                    R1=list(R1d), observation_target=list(target))
            oc = oracle.make_ctrl(n, m, mode="RQL", Nactor=4, gamma=gamma, Ncritic=Ncritic, buffer_size=bufsz, critic_struct=cs,
                                  R1=list(R1d), observation_target=list(target))
            E = len(probs)
            ob = torch.as_tensor(np.stack([p[0] for p in probs], axis=2), device=cuda).contiguous()        # [L, n, E]
            ab = torch.as_tensor(np.stack([p[1] for p in probs], axis=2), device=cuda).contiguous()
            wp = torch.as_tensor(np.stack([p[2] for p in probs], axis=1), device=cuda).contiguous()        # [dimc, E]
            dimc = wp.shape[0]
            w = torch.zeros((dimc, E), dtype=torch.float64, device=cuda)
            w_init = torch.ones((dimc,), dtype=torch.float64, device=cuda)
            Jc = torch.zeros((E,), dtype=torch.float64, device=cuda)
            lo, hi = w_bounds(cs)
            if variant == "budget24_then_rest":
                ops.critic_fit(obj, n, m, ob, ab, wp, lo, hi, w, w_init=w_init, max_evals=24, Jc_out=Jc)
                Jb = Jc.clone()
                ops.critic_fit(obj, n, m, ob, ab, wp, lo, hi, w, w_init=w_init, Jc_out=Jc)
                # (1e-3: the budgeted call backtracks, the converged one searches exactly -- two descent paths, same bar)
                assert bool((Jc <= Jb * (1 + 1e-3) + 1e-300).all()), "the budgeted fit must not beat the converged one"
            else:
                ops.critic_fit(obj, n, m, ob, ab, wp, lo, hi, w, w_init=w_init, Jc_out=Jc)
            wh, Jh = w.cpu().numpy(), Jc.cpu().numpy()
            assert wh.min() >= lo and wh.max() <= hi
            for e, (o_, a_, wp_, J_ref, J_init) in enumerate(probs):
                # the checker mirrors the line search of the path taken: exact on the two-phase path (K <= 3, >= 10 weights,
                # run to convergence), Armijo backtracking in the one-lane kernels
                w_o, J_o, _ = oracle.critic_fit(oc, n, m, o_, a_, wp_, lo, hi, ls_mode=0 if variant == "one_phase" else None)
                J_chk = oracle.critic_cost(oc, n, m, o_, a_, wh[:, e].copy(), wp_)
                floor = 1e-9 * abs(J_init) + 1e-18
                assert abs(J_chk - Jh[e]) <= 1e-6 * max(abs(J_chk), floor), (key, e, J_chk, Jh[e])
                total += 1
                if abs(Jh[e] - J_o) > 1e-6 * max(abs(J_o), floor):
                    off.append((key[:2], e, Jh[e], J_o))
                if not Jh[e] <= J_ref * (1 + 1e-6) + 1e-9 * abs(J_init):
                    worse_than_ref.append((key[:2], e, Jh[e], J_ref))
        assert total >= 280
        assert not worse_than_ref, worse_than_ref[:5]
        assert not off, (len(off), total, off[:8])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("name,mode,cs,N,t1,per", [("3wrobot", "RQL", "quadratic", 10, 0.3, None), ("2tank", "SQL", "quad-nomix", 8, 6.0, None),
                                                  ("3wrobotNI", "RQL", "quad-mix", 5, 0.3, 0.03)])
def test_engine_refit_loop_against_oracle_loop(cuda, name, mode, cs, N, t1, per):
    """BASELINE configs 3 / 4 (64 environments) with the product's own critic refit against the oracle's loop with the
    restated fit: identical step / sample / refit counts and solver times, states, returns and weights to 1e-6."""
    import oracle
    from rcognita_b200.engine import ClosedLoopEngine
    n, m = DIMS[name]
    cfg = PRESET[name]
    E = 64
    rng = np.random.default_rng(5)
    box = {"3wrobotNI": ([-5, -5, -3], [5, 5, 3]), "3wrobot": ([-5, -5, -3, -1, -1], [5, 5, 3, 1, 1]), "2tank": ([-2, -2], [2, 2])}[name]
    x0 = rng.uniform(box[0], box[1], size=(E, n))
    b = np.array(cfg["bnds"], dtype=float)
    cand = rng.uniform(np.tile(b[:, 0], N), np.tile(b[:, 1], N), size=(64, N * m))
    a_init = [0.5] if name == "2tank" else list(b[:, 0] / 10)
    eng = ClosedLoopEngine(name, x0, cand, pars=cfg["pars"], ctrl_bnds=cfg["bnds"], mode=mode, Nactor=N, dt=cfg["dt"],
                           pred_step_size=cfg["dt"] * cfg["psm"], t1=t1, R1=cfg["R1_diag"], observation_target=cfg["target"],
                           critic_struct=cs, critic_fit=True, Ncritic=4, buffer_size=10, critic_period=per, action_init=a_init)
    eng.run()
    got = eng.results()
    s = oracle.make_sys(name, cfg["pars"], cfg["bnds"])
    c = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=cfg["dt"] * cfg["psm"], Ncritic=4, buffer_size=10,
                         critic_struct=cs, R1=cfg["R1_diag"], observation_target=cfg["target"])
    ref = oracle.closed_loop_critic(c, s, x0, cand, a_init, cfg["dt"], 0.0, t1, cfg["dt"] / 2, 10, w_bounds(cs), critic_period=per)
    assert np.array_equal(got["nsteps"], ref["nsteps"])
    assert np.array_equal(got["nsamples"], ref["nsamples"])
    assert np.array_equal(got["nfits"], ref["nfits"])
    assert got["nfits"].min() >= 5
    assert np.array_equal(got["t"], ref["t"])
    # Solver times, step / sample / refit counts are exact; the trajectory and the returns must agree in EVERY environment
    # to 1e-6.  The fitted cost is compared to 1e-3 everywhere and 1e-6 in the median: on the isolated problems of
    # test_critic_fit_kernels_against_oracle_fit the kernels agree with the restatement to 1e-6, but inside the loop many
    # problems are infeasible (the critic diverges in the reference too: J_c ~ 1e13 in closed_loop_refit.json) and end at the
    # iteration caps, where the kernels' fused multiply-adds (grouped differently from the scalar restatement, butterfly
    # sums in the warp kernel) leave the iterates a few 1e-5 apart.  The weights themselves are not unique (3 rows, up to 35
    # unknowns); their agreement is printed, not asserted.
    y_err = np.max(np.abs(got["y"] - ref["y"]) / np.maximum(np.abs(ref["y"]), 1e-2), axis=1)
    a_err = np.abs(got["accum"] - ref["accum"]) / np.abs(ref["accum"])
    J_err = np.abs(got["Jc"] - ref["Jc"]) / np.maximum(np.abs(ref["Jc"]), 1e-9 * np.max(np.abs(ref["Jc"])) + 1e-300)
    w_err = np.max(np.abs(got["w_critic"] - ref["w_critic"]) / np.maximum(np.abs(ref["w_critic"]), 1e-3), axis=1)
    msg = dict(y=float(y_err.max()), accum=float(a_err.max()), Jc_max=float(J_err.max()), Jc_p90=float(np.percentile(J_err, 90)),
               Jc_frac_1e6=float(np.mean(J_err <= 1e-6)), w_frac_1e6=float(np.mean(w_err <= 1e-6)),
               w_p50=float(np.median(w_err)), w_p90=float(np.percentile(w_err, 90)), w_max=float(w_err.max()))
    print("refit loop vs oracle loop:", name, msg)
    assert y_err.max() <= 1e-6, msg
    assert a_err.max() <= 1e-6, msg
    assert J_err.max() <= 1e-3 and np.median(J_err) <= 1e-6, msg
