"""When the reference tree is mounted (/root/reference -- the build container only; the GPU box does not have it), import it
UNMODIFIED (GUI imports stubbed exactly like tests/golden/make_golden.py) and compare the oracle with it on FRESH random
inputs, and check that the committed fixtures are what the live reference produces today.  Skipped elsewhere: the
committed fixtures under tests/golden/ then carry the pin."""
import importlib.util
import json
import os

import numpy as np
import pytest

import oracle
from golden_util import DIMS, GOLDEN, PRESET

REF = os.environ.get("RCOGNITA_REF", "/root/reference")
pytestmark = [pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rcognita")), reason="reference tree not mounted"),
              pytest.mark.filterwarnings("ignore")]


@pytest.fixture(scope="module")
def mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)           # stubs the GUI modules and imports rcognita from REF; generates nothing
    return mod


@pytest.mark.parametrize("name", ["3wrobotNI", "3wrobot", "2tank"])
def test_oracle_equals_live_reference_on_fresh_inputs(mg, name):
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(20261017 + n)
    my_sys = mg.make_sys(name)
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    bn = np.array(P["bnds"], dtype=float)
    for mode, cs, N, gamma in (("MPC", "quad-nomix", 6, 1.0), ("RQL", "quadratic", 10, 0.9), ("SQL", "quad-lin", 4, 1.0),
                               ("RQL", "quad-mix", 7, 0.95), ("SQL", "quad-nomix", 8, 1.0)):
        x_sys = rng.uniform(-5, 5, size=n)
        ob = x_sys + 0.01 * rng.normal(size=n)
        ctrl = mg.make_ctrl(name, my_sys, mode, N, critic_struct=cs, gamma=gamma, state_sys=x_sys)
        ct = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=float(ctrl.pred_step_size), gamma=gamma,
                              critic_struct=cs, R1=np.diag(np.array(P["R1_diag"], dtype=float)), observation_target=P["target"],
                              Ncritic=int(ctrl.Ncritic), buffer_size=10)
        w = rng.uniform(0, 2, size=ctrl.dim_critic)
        ctrl.w_critic = w
        for _ in range(5):
            st, ac = rng.uniform(-6, 6, size=n), rng.uniform(bn[:, 0], bn[:, 1])
            assert np.allclose(oracle.state_dyn(s, st, ac), my_sys._state_dyn([], st, ac), rtol=1e-14, atol=0)
            assert abs(oracle.stage_obj(ct, n, m, st, ac) - ctrl.stage_obj(st, ac)) <= 1e-12 * abs(ctrl.stage_obj(st, ac))
            q = ctrl._critic(st, ac, w)
            assert abs(oracle.critic(ct, n, m, st, ac, w) - q) <= 1e-12 * max(abs(q), 1e-9)
            u = rng.uniform(ctrl.action_sqn_min, ctrl.action_sqn_max)
            J = float(ctrl._actor_cost(u, ob))
            assert abs(oracle.actor_cost(ct, s, u, ob, x_sys, w) - J) <= 1e-11 * max(abs(J), 1e-9)
        ctrl.observation_buffer = rng.normal(size=(10, n))
        ctrl.action_buffer = rng.uniform(bn[:, 0], bn[:, 1], size=(10, m))
        ctrl.w_critic_prev = rng.uniform(0, 2, size=ctrl.dim_critic)
        Jc = float(ctrl._critic_cost(w))
        got = oracle.critic_cost(ct, n, m, ctrl.observation_buffer, ctrl.action_buffer, w, ctrl.w_critic_prev)
        assert abs(got - Jc) <= 1e-11 * max(abs(Jc), 1e-9)


def test_oracle_rk45_follows_live_scipy_step_for_step(mg):
    """Sys3WRobot under a jumping action schedule (rejections, FSAL staleness): nfev (i.e. every accept / reject
    decision) exact, t / h_abs / y of every solver step to 1e-12 -- the oracle's sin, cos and err**-0.2 are fully specified
    functions (so that CPU and GPU agree bit for bit) and differ from this machine's libm in the last bit."""
    name = "3wrobot"
    P = PRESET[name]
    my_sys = mg.make_sys(name)
    sim = mg.make_sim(name, my_sys, 0.3, x0=[1.0, -2.0, 0.7, 0.4, -0.3])
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    r = oracle.RK45(s, [1.0, -2.0, 0.7, 0.4, -0.3], 0.0, 0.3, P["dt"] / 2)
    rng = np.random.default_rng(3)
    k = 0
    while True:
        sim.sim_step()
        r.step()
        k += 1
        so = sim.ODE_solver
        assert r.nfev == so.nfev, k
        assert abs(r.t - so.t) <= 1e-12 * so.t and abs(r.h_abs - so.h_abs) <= 1e-12 * so.h_abs, k
        assert np.max(np.abs(r.y - so.y) / np.maximum(np.abs(so.y), 1e-3)) <= 1e-12
        if k % 4 == 0:
            a = rng.uniform([-500, -150], [500, 150])           # partly outside the bounds: clipped in place
            my_sys.receive_action(a.copy())
            r.receive_action(a.copy())
        if so.status != "running":
            assert r.status == so.status
            break
    assert k > 50


def test_nominal_controller_oracle_equals_live_reference(mg):
    bn = np.array(PRESET["3wrobotNI"]["bnds"], dtype=float)
    nom = mg.controllers.CtrlNominal3WRobotNI(ctrl_gain=0.5, ctrl_bnds=bn, t0=0, sampling_time=0.01)
    s = oracle.make_sys("3wrobotNI", [], bn)
    rng = np.random.default_rng(11)
    for k in range(200):
        ob = rng.uniform([-10, -10, -np.pi], [10, 10, np.pi]) * rng.choice([1.0, 0.01])
        ref = np.array(nom.compute_action(1.0 + k, ob), dtype=float)
        got = oracle.nominal_ni(0.5, s, ob)
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9)) <= 1e-12


def test_committed_fixtures_are_what_the_live_reference_produces(mg):
    """Spot check: regenerate a slice of functions.json / nominal.json from the live reference and compare with the
    committed files bit for bit (the generator is seeded)."""
    fn = json.load(open(os.path.join(GOLDEN, "functions.json")))
    for name in ("3wrobotNI", "3wrobot", "2tank"):
        my_sys = mg.make_sys(name)
        for c in fn[name]["cases"]["state_dyn"]:
            assert mg.L(my_sys._state_dyn([], np.array(c["state"]), np.array(c["action"]))) == c["out"]
        c = fn[name]["cases"]["actor_cost"][0]
        ctrl = mg.make_ctrl(name, my_sys, c["mode"], c["N"], critic_struct=c["critic_struct"], gamma=c["gamma"],
                            target=c["target"], state_sys=np.array(c["state_sys"]))
        ctrl.w_critic = np.array(c["w"])
        assert [float(ctrl._actor_cost(np.array(u), np.array(c["obs"]))) for u in c["cand"]] == c["J"]
    nom_cases = json.load(open(os.path.join(GOLDEN, "nominal.json")))["cases"]
    bn = np.array(PRESET["3wrobotNI"]["bnds"], dtype=float)
    for c in nom_cases[:10]:
        nom = mg.controllers.CtrlNominal3WRobotNI(ctrl_gain=c["gain"], ctrl_bnds=bn, t0=0, sampling_time=0.01)
        assert mg.L(nom.compute_action(1.0, np.array(c["obs"]))) == c["action"]


@pytest.mark.parametrize("name,mode,N,t1,seed", [("3wrobotNI", "MPC", 5, 0.6, 5), ("2tank", "MPC", 4, 4.0, 6), ("3wrobot", "MPC", 3, 0.3, 7)])
def test_oracle_closed_loop_equals_live_reference_loop_on_fresh_inputs(mg, name, mode, N, t1, seed):
    """The reference's own loop (Simulator + CtrlOptPred, `_actor_optimizer` replaced by its `_actor_cost` on a candidate
    table + np.argmin -- SURVEY App. A.4) from a FRESH random start and table against orc_closed_loop: same number of
    solver steps and controller samples, final state and accumulated objective to 1e-9."""
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-3, 3, size=n)
    g = mg.closed_loop(name, mode, N, t1, C=24, seed=seed, x0=x0)
    rows = np.array(g["rows"])
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=P["dt"] * P["psm"], R1=P["R1_diag"],
                          observation_target=P["target"])
    ref = oracle.closed_loop(ct, s, x0[None, :], np.array(g["cand"]), g["action_init"], P["dt"], 0.0, t1, P["dt"] / 2)
    assert int(ref["nsteps"][0]) == rows.shape[0] and int(ref["nsamples"][0]) == len(g["picks"])
    assert abs(ref["t"][0] - rows[-1, 0]) <= 1e-12 * t1
    assert np.max(np.abs(ref["y"][0] - rows[-1, 1:1 + n]) / np.maximum(np.abs(rows[-1, 1:1 + n]), 1e-2)) <= 1e-9
    assert abs(ref["accum"][0] - rows[-1, 1 + n + m]) <= 1e-9 * abs(rows[-1, 1 + n + m])


@pytest.mark.parametrize("name,mode,cs,N,scale,seed", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, 4.0, 11), ("3wrobotNI", "RQL", "quadratic", 5, 0.3, 12),
    ("3wrobot", "MPC", "quad-nomix", 5, 2.0, 13), ("3wrobot", "SQL", "quad-nomix", 6, 0.5, 14),
    ("2tank", "SQL", "quad-lin", 8, 1.0, 15), ("2tank", "MPC", "quad-nomix", 10, 0.2, 16)])
def test_restated_minimiser_reaches_live_slsqp_minimum_on_fresh_problems(mg, name, mode, cs, N, scale, seed):
    """The reference's `_actor_optimizer` (SLSQP) run live on a FRESH problem, against the oracle's restatement of the
    batched minimiser from the same start point: minimum reached or beaten (SLSQP's own tolerance as slack), feasible."""
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(seed)
    tgt = np.array(P["target"]) if len(P["target"]) else 0.0
    x_sys = rng.uniform(-1, 1, size=n) * scale + tgt
    ob = x_sys + rng.normal(size=n) * 0.005 * scale
    my_sys = mg.make_sys(name)
    ctrl = mg.make_ctrl(name, my_sys, mode, N, critic_struct=cs, state_sys=x_sys)
    w = rng.uniform(0, 2, size=ctrl.dim_critic)
    ctrl.w_critic = w
    rec = {}
    orig = mg.controllers.minimize

    def wrapped(fun, x0, **kw):
        rec["res"] = orig(fun, x0, **kw)
        return rec["res"]
    mg.controllers.minimize = wrapped
    try:
        ctrl._actor_optimizer(ob)
    finally:
        mg.controllers.minimize = orig
    J_ref = float(ctrl._actor_cost(np.asarray(rec["res"].x, dtype=float), ob))
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=float(ctrl.pred_step_size), critic_struct=cs,
                          R1=np.diag(np.array(P["R1_diag"], dtype=float)), observation_target=P["target"])
    x, J, iters, nfev = oracle.actor_opt(ct, s, np.asarray(ctrl.action_sqn_init, dtype=float), ob, x_sys, w, max_iter=300,
                                         pg_tol=1e-7, f_tol=1e-12)
    b = np.array(P["bnds"], dtype=float)
    assert np.all(x >= np.tile(b[:, 0], N)) and np.all(x <= np.tile(b[:, 1], N))
    assert abs(J - float(ctrl._actor_cost(x, ob))) <= 1e-9 * max(abs(J), 1e-9)
    assert J <= J_ref + 1e-7 * max(abs(J_ref), 1.0), (J, J_ref, iters, rec["res"].nfev)
