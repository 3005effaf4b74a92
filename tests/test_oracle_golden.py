"""Pins the CPU oracle (oracle/rcg_oracle.c) against outputs of the LIVE reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py from the
unmodified rcognita v0.1.2 + scipy 1.18.1 (the reference ships no tests of its own).
Tolerances: costs 1e-12 relative (north-star bar is 1e-9); integrator state 1e-9 relative
(bar 1e-6), step times t and nfev EXACT.
"""
import numpy as np
import pytest

import oracle
from golden_util import DIMS, PRESET, load, mixed_err, rel_err

COST_RTOL = 1e-12
SYSTEMS = ["3wrobotNI", "3wrobot", "2tank"]


@pytest.fixture(scope="module")
def fn():
    return load("functions.json")


def _sys(name, d=None):
    d = d or PRESET[name]
    return oracle.make_sys(name, d["pars"], d["bnds"])


@pytest.mark.parametrize("name", SYSTEMS)
def test_state_dyn_and_closed_loop_rhs(fn, name):
    s = _sys(name, fn[name])
    for c in fn[name]["cases"]["state_dyn"]:
        assert rel_err(oracle.state_dyn(s, c["state"], c["action"]), c["out"]) <= 4e-16
    for c in fn[name]["cases"]["closed_loop_rhs"]:
        rhs, clipped = oracle.closed_loop_rhs(s, c["state"], c["action"])
        assert rel_err(rhs, c["out"]) <= 4e-16
        assert np.array_equal(clipped[: s.m], np.array(c["action_clipped"]))


@pytest.mark.parametrize("name", SYSTEMS)
def test_stage_obj(fn, name):
    n, m = DIMS[name]
    for c in fn[name]["cases"]["stage_obj"]:
        ctrl = oracle.make_ctrl(n, m, R1=c["R1"], R2=c["R2"], stage_obj_struct=c["struct"], observation_target=c["target"])
        assert rel_err(oracle.stage_obj(ctrl, n, m, c["obs"], c["act"]), c["out"]) <= COST_RTOL


@pytest.mark.parametrize("name", SYSTEMS)
def test_critic_and_critic_cost(fn, name):
    n, m = DIMS[name]
    for c in fn[name]["cases"]["critic"]:
        ctrl = oracle.make_ctrl(n, m, critic_struct=c["critic_struct"], observation_target=c["target"])
        assert oracle.dim_critic(c["critic_struct"], n, m) == c["dim_critic"]
        assert rel_err(oracle.critic(ctrl, n, m, c["obs"], c["act"], c["w"]), c["out"]) <= COST_RTOL
    for c in fn[name]["cases"]["critic_cost"]:
        ctrl = oracle.make_ctrl(n, m, mode="RQL", critic_struct=c["critic_struct"], gamma=c["gamma"], Ncritic=4,
                                buffer_size=10, R1=c["R1_diag"], observation_target=c["target"])
        assert ctrl.Ncritic == c["Ncritic"]
        got = oracle.critic_cost(ctrl, n, m, c["obs_buf"], c["act_buf"], c["w"], c["w_prev"])
        assert rel_err(got, c["out"]) <= 1e-11       # difference of O(1e3) terms squared


@pytest.mark.parametrize("name", SYSTEMS)
def test_actor_cost_and_argmin(fn, name):
    n, m = DIMS[name]
    s = _sys(name, fn[name])
    for c in fn[name]["cases"]["actor_cost"]:
        ctrl = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                                critic_struct=c["critic_struct"], R1=c["R1"], observation_target=c["target"])
        J, am = oracle.actor_cost_table(ctrl, s, c["cand"], c["obs"], c["state_sys"], c["w"])
        assert rel_err(J, c["J"]) <= COST_RTOL, (c["mode"], c["critic_struct"], c["N"])
        assert am == c["argmin"]
        assert J[4] == J[1]                     # duplicated candidate -> exact tie
        one = oracle.actor_cost(ctrl, s, c["cand"][2], c["obs"], c["state_sys"], c["w"])
        assert one == J[2]


def test_deterministic_elementary_functions():
    """orc_sincos / orc_pow_m02 (the functions the CUDA fp64 path reproduces bit for bit): < 1 ulp
    against libm over the path's argument range, quadrant logic, special values; x ** -0.2
    equal to libm's pow except on rare last-bit cases (where libm is the one mis-rounding)."""
    import math
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-30, 30, 20000), rng.uniform(-1e4, 1e4, 5000), rng.normal(size=5000) * 1e-4,
                         np.arange(-16, 17) * (np.pi / 4), [0.0, -0.0, 1e-300, 99999.0]])
    for x in xs:
        s, c = oracle.sincos(x)
        assert abs(s - math.sin(x)) <= 1.0 * np.spacing(abs(math.sin(x))) + 1e-300, x
        assert abs(c - math.cos(x)) <= 1.0 * np.spacing(abs(math.cos(x))) + 1e-300, x
    assert oracle.sincos(0.0) == (0.0, 1.0)
    assert all(np.isnan(v) for v in oracle.sincos(float("nan"))) and all(np.isnan(v) for v in oracle.sincos(float("inf")))
    assert oracle.sincos(3e7) == (math.sin(3e7), math.cos(3e7))          # huge arguments: libm fallback
    es = np.concatenate([rng.uniform(0, 1, 20000) ** 3 * 5 + 1e-12, [1.0, 0.59049, 1e-30, 1e30]])
    mism = 0
    for e in es:
        got, ref = oracle.pow_m02(e), float(e) ** -0.2
        assert abs(got - ref) <= np.spacing(ref), e
        mism += got != ref
    assert mism <= 0.005 * len(es)
    assert oracle.pow_m02(1.0) == 1.0 and oracle.pow_m02(float("inf")) == 0.0 and np.isnan(oracle.pow_m02(float("nan")))


def test_argmin_semantics():
    assert oracle.argmin([3.0, 1.0, 1.0, 2.0]) == 1           # first minimum
    assert oracle.argmin([3.0, np.nan, 0.0, np.nan]) == 1     # np.argmin: first NaN wins
    assert oracle.argmin([np.inf, np.inf]) == 0
    for _ in range(20):
        x = np.random.default_rng(_).normal(size=37)
        assert oracle.argmin(x) == int(np.argmin(x))


def test_survey_appendix_a1_known_answers():
    """SURVEY.md App. A.1 values (captured from the live reference during the survey)."""
    x0 = [5, 5, -3 * np.pi / 4]
    s = _sys("3wrobotNI")
    assert rel_err(oracle.state_dyn(s, x0, [25, 5]), [-17.677669529663685, -17.67766952966369, 5.0]) <= 4e-16
    ctrl = oracle.make_ctrl(3, 2, mode="MPC", Nactor=6, pred_step_size=0.01, R1=[1, 10, 1, 0, 0])
    assert rel_err(oracle.stage_obj(ctrl, 3, 2, x0, [25, 5]), 280.55165247561274) <= COST_RTOL
    assert rel_err(oracle.actor_cost(ctrl, s, [-2.5, -0.5] * 6, x0, x0), 1712.8592297260025) <= COST_RTOL
    lo, hi = np.tile([-25.0, -5.0], 6), np.tile([25.0, 5.0], 6)
    U = np.random.default_rng(0).uniform(lo, hi, size=(4, 12))
    J, am = oracle.actor_cost_table(ctrl, s, U, x0, x0)
    assert rel_err(J, [1683.274365086458, 1576.183747793757, 1563.836057710278, 1605.7555990810986]) <= COST_RTOL
    assert am == 2
    ctrl9 = oracle.make_ctrl(3, 2, mode="MPC", Nactor=6, pred_step_size=0.01, gamma=0.9, R1=[1, 10, 1, 0, 0])
    assert rel_err(oracle.actor_cost(ctrl9, s, U[0], x0, x0), 1315.6103918095434) <= COST_RTOL
    rb = np.random.default_rng(1)
    obs_buf = rb.normal(size=(10, 3)); act_buf = rb.uniform(-1, 1, size=(10, 2))
    table = {"quad-lin": (20, 1217.0970145208614, 1740.3482209252315, 2005.1583306576185, 63.287885007843386),
             "quadratic": (15, 1147.3381646032076, 1708.1267238783219, 1891.0533070732733, 80.4956769728908),
             "quad-nomix": (5, 271.66549574268385, 1528.9552042195323, 543.0573288416265, 122.06204560680067),
             "quad-mix": (11, 758.9387307329714, 1723.160023213367, 1329.304409163996, 113.92951122558704)}
    for cs, (dimc, q, jr, js, jc) in table.items():
        assert oracle.dim_critic(cs, 3, 2) == dimc
        w = np.arange(1, dimc + 1) / 10
        cr = oracle.make_ctrl(3, 2, mode="RQL", Nactor=6, pred_step_size=0.01, critic_struct=cs, R1=[1, 10, 1, 0, 0],
                              Ncritic=4, buffer_size=10)
        cq = oracle.make_ctrl(3, 2, mode="SQL", Nactor=6, pred_step_size=0.01, critic_struct=cs, R1=[1, 10, 1, 0, 0],
                              Ncritic=4, buffer_size=10)
        assert rel_err(oracle.critic(cr, 3, 2, x0, [25, 5], w), q) <= COST_RTOL
        assert rel_err(oracle.actor_cost(cr, s, U[0], x0, x0, w), jr) <= COST_RTOL
        assert rel_err(oracle.actor_cost(cq, s, U[0], x0, x0, w), js) <= COST_RTOL
        assert rel_err(oracle.critic_cost(cr, 3, 2, obs_buf, act_buf, w, np.ones(dimc)), jc) <= 1e-11


@pytest.mark.parametrize("key", ["3wrobotNI:inbounds", "3wrobotNI:outofbounds", "3wrobot:inbounds",
                                 "3wrobot:outofbounds", "2tank:inbounds", "2tank:outofbounds"])
def test_rk45_integrator_traces(key):
    """SURVEY.md App. A.3 protocol: sim_step(); k += 1; receive_action(schedule[(k//5) % 4])."""
    g = load("integrator.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    d = PRESET[name]
    s = _sys(name)
    x0 = {"3wrobotNI": [5, 5, -3 * np.pi / 4], "3wrobot": [5, 5, -3 * np.pi / 4, 0.3, -0.2], "2tank": [2, -2]}[name]
    r = oracle.RK45(s, x0, 0.0, g["t1"], d["dt"] / 2, 1e-6, 1e-3, 1e-5)
    sched = np.array(g["sched"])
    rows = np.array(g["rows"])
    k = 0
    n_t_mismatch = 0
    while True:
        r.step()
        k += 1
        r.receive_action(sched[(k // 5) % 4])
        ref = rows[k - 1]
        n_t_mismatch += int(r.t != ref[0])
        assert abs(r.t - ref[0]) <= 1e-13 * max(abs(ref[0]), 1e-6), (k, r.t, ref[0])
        assert mixed_err(r.y, ref[1:1 + n], floor=1e-3) <= 1e-9, k
        assert mixed_err(r.f, ref[1 + n:1 + 2 * n], floor=1e-3) <= 1e-9, k
        assert rel_err(r.h_abs, ref[1 + 2 * n]) <= 1e-9, k
        assert r.nfev == int(ref[2 + 2 * n]), k           # same accept/reject pattern
        if r.status != "running":
            break
    assert k == len(rows) and r.status == g["status"]
    # last-bit differences of t appear only downstream of err**-0.2 steps (3wrobot: rejections)
    # (numpy's BLAS dot sums K.T @ E in a different order than the scalar oracle -> err differs by ~1e-16)
    # -> h after a non-max step differs in its last bits, and the offset in t then persists); NI/2tank in-bounds: exact
    assert n_t_mismatch == 0 or key not in ("3wrobotNI:inbounds", "2tank:inbounds"), n_t_mismatch
    with pytest.raises(RuntimeError):
        r.step()


@pytest.mark.parametrize("key", ["NI_MPC_N6", "NI_MPC_N6_x1", "3wrobot_RQL_N10", "2tank_SQL_N8"])
def test_closed_loop_candidate_controller(key):
    """SURVEY.md App. A.4 protocol against the reference's own closed loop."""
    g = load("closed_loop.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    d = PRESET[name]
    s = _sys(name)
    ctrl = oracle.make_ctrl(n, m, mode=g["mode"], Nactor=g["Nactor"], pred_step_size=d["dt"] * d["psm"], gamma=g["gamma"],
                            critic_struct=g["critic_struct"], R1=d["R1_diag"], observation_target=d["target"])
    rows = np.array(g["rows"])
    out = oracle.closed_loop(ctrl, s, [g["x0"]], np.array(g["cand"]), g["action_init"], d["dt"], 0.0, g["t1"], d["dt"] / 2,
                             w_critic=g["w_fixed"], traj_cap=len(rows) + 10)
    tr = out["traj"]
    assert tr.shape[0] == rows.shape[0] == out["nsteps"][0]
    assert out["nsamples"][0] == len(g["picks"])
    assert out["nfev"][0] == g["nfev"]
    assert np.max(np.abs(tr[:, 0] - rows[:, 0])) <= 1e-15 * g["t1"]                  # t
    assert mixed_err(tr[:, 1:1 + n], rows[:, 1:1 + n], floor=1e-2) <= 1e-9           # state (bar: 1e-6)
    assert np.array_equal(tr[:, 1 + n:1 + n + m], rows[:, 1 + n:1 + n + m])          # selected actions, bit-exact
    assert rel_err(tr[:, 1 + n + m], rows[:, 1 + n + m]) <= 1e-10                    # accum_obj
    sampled = rows[:, -1] > 0
    picks = np.array(g["picks"])
    assert np.array_equal(tr[sampled, 2 + n + m].astype(int), picks[:, 0].astype(int))   # arg-min indices
    assert rel_err(tr[sampled, 3 + n + m], picks[:, 1]) <= 1e-11


def test_config1_slsqp_episode_replay(golden_dir):
    """Config 1: replay the reference's recorded SLSQP actions through the
    oracle integrator; 2004 steps, nfev 12025, final state/accum of BASELINE.md."""
    import os
    z = np.load(os.path.join(golden_dir, "config1_slsqp_episode.npz"))
    rows = z["rows"]
    assert rows.shape[0] == 2004 and int(rows[-1, 9]) == 12025
    assert rel_err(rows[-1, 7], 71.31881380046568) <= 1e-12
    s = _sys("3wrobotNI")
    ctrl = oracle.make_ctrl(3, 2, mode="MPC", Nactor=6, pred_step_size=0.01, R1=[1, 10, 1, 0, 0])
    r = oracle.RK45(s, [5, 5, -3 * np.pi / 4], 0.0, 10.0, 0.005, 1e-6, 1e-3, 1e-5)
    acc = 0.0
    for k in range(rows.shape[0]):
        r.step()
        assert r.t == rows[k, 0], k
        r.receive_action(rows[k, 4:6])
        acc += oracle.stage_obj(ctrl, 3, 2, r.y, rows[k, 4:6]) * 0.01
        assert mixed_err(r.y, rows[k, 1:4], floor=1e-2) <= 1e-9, k
    assert r.nfev == 12025 and r.status == "finished"
    assert rel_err(acc, rows[-1, 7]) <= 1e-10


# ------------------------------------------------------------------ actor optimiser (oracle/rcg_oracle_opt.c)

def _opt_case(c):
    name = c["system"]
    n, m = DIMS[name]
    s = oracle.make_sys(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                          critic_struct=c["critic_struct"], R1=np.array(c["R1"]), observation_target=c["target"])
    w = c["w"] if c["mode"] != "MPC" else None
    return s, ct, w


def test_actor_opt_golden_costs_are_the_references():
    """The oracle's _actor_cost reproduces the live reference's cost at SLSQP's minimiser and at the start."""
    for c in load("actor_opt.json"):
        s, ct, w = _opt_case(c)
        for x, J in ((c["x_ref"], c["J_ref"]), (c["x_init"], c["J_init"])):
            got = oracle.actor_cost(ct, s, x, c["obs"], c["state_sys"], w)
            assert abs(got - J) <= 1e-9 * max(abs(J), 1e-6), (c["system"], c["mode"], got, J)


def test_actor_grad_matches_finite_differences_of_the_reference_cost():
    """Adjoint gradient vs Richardson-extrapolated central differences of orc_actor_cost (which is pinned to the
    live reference): this is the derivative SLSQP approximates by forward differences."""
    rng = np.random.default_rng(0)
    for c in load("actor_opt.json"):
        s, ct, w = _opt_case(c)
        x = np.array(c["x_ref"]) + rng.normal(size=len(c["x_ref"])) * 0.1
        J, g = oracle.actor_grad(ct, s, x, c["obs"], c["state_sys"], w)
        f = lambda z: oracle.actor_cost(ct, s, z, c["obs"], c["state_sys"], w)       # noqa: E731
        assert J == f(x)
        gf = np.zeros_like(g)
        for i in range(x.size):
            h = 1e-3 * max(1.0, abs(x[i]))

            def cd(hh):
                xp, xm = x.copy(), x.copy()
                xp[i] += hh
                xm[i] -= hh
                return (f(xp) - f(xm)) / (2 * hh)
            gf[i] = (4 * cd(h / 2) - cd(h)) / 3
        assert np.max(np.abs(g - gf)) <= 1e-7 * max(np.max(np.abs(gf)), 1e-6), (c["system"], c["mode"], c["critic_struct"])


def test_actor_opt_oracle_reaches_the_reference_slsqp_minimum():
    """From the reference's start point (action_sqn_init) the restated minimiser ends at or below the cost the live
    reference's SLSQP reached, stays inside the box, and never goes above the start cost."""
    for c in load("actor_opt.json"):
        s, ct, w = _opt_case(c)
        x, J, iters, nfev = oracle.actor_opt(ct, s, c["x_init"], c["obs"], c["state_sys"], w, max_iter=300,
                                             pg_tol=1e-7, f_tol=1e-12)
        b = np.array(PRESET[c["system"]]["bnds"], dtype=float)
        lo, hi = np.tile(b[:, 0], c["N"]), np.tile(b[:, 1], c["N"])
        assert np.all(x >= lo) and np.all(x <= hi)
        assert J <= c["J_init"] + 1e-12 * abs(c["J_init"])
        assert J <= c["J_ref"] + 1e-7 * max(abs(c["J_ref"]), 1.0), (c["system"], c["mode"], c["critic_struct"], c["N"], J, c["J_ref"])
        assert J == oracle.actor_cost(ct, s, x, c["obs"], c["state_sys"], w)


def test_actor_opt_hybrid_reaches_the_slsqp_minimum_with_bounded_iterations():
    """Gauss-Newton (iLQR) sweeps + L-BFGS hand-over (groundwork for the next optimiser kernel): at or
    below the live reference's SLSQP cost on every golden problem, inside the box, and the slowest problem needs an
    order of magnitude fewer dependent iterations than L-BFGS alone (which runs up to its 300-iteration cap)."""
    worst, total_lbfgs = 0, 0
    for c in load("actor_opt.json"):
        s, ct, w = _opt_case(c)
        x, J, sweeps, iters = oracle.actor_opt_hybrid(ct, s, c["x_init"], c["obs"], c["state_sys"], w, max_sweeps=25,
                                                      max_iter=300, pg_tol=1e-7, f_tol=1e-12)
        _, _, it_alone, _ = oracle.actor_opt(ct, s, c["x_init"], c["obs"], c["state_sys"], w, max_iter=300, pg_tol=1e-7,
                                             f_tol=1e-12)
        b = np.array(PRESET[c["system"]]["bnds"], dtype=float)
        lo, hi = np.tile(b[:, 0], c["N"]), np.tile(b[:, 1], c["N"])
        assert np.all(x >= lo) and np.all(x <= hi)
        assert J <= c["J_ref"] + 1e-7 * max(abs(c["J_ref"]), 1.0), (c["system"], c["mode"], c["critic_struct"], c["N"], J, c["J_ref"])
        assert J == oracle.actor_cost(ct, s, x, c["obs"], c["state_sys"], w)
        assert sweeps <= 25
        worst = max(worst, sweeps + iters)
        total_lbfgs = max(total_lbfgs, it_alone)
    assert worst <= 40 and total_lbfgs >= 5 * worst, (worst, total_lbfgs)
