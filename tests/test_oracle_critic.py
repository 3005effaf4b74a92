"""Pins the critic side of the CPU oracle (oracle/rcg_oracle_critic.c) against the LIVE reference.

* ``closed_loop_refit.json``: the unmodified reference's RQL / SQL loop with its own SLSQP ``_critic_optimizer``
  (controllers.py:1455-1479, :1248-1271; only ``_actor_optimizer`` is the candidate/arg-min stand-in).  The oracle loop
  LOADS the recorded fitted weights instead of fitting; everything else -- FIFO pushes, critic-clock firings, the
  ``w_critic_prev`` hand-over, arg-min picks, trajectory, accumulated objective -- must then match to 1e-9.
* ``critic_fit.json``: the oracle's restated fit (the product's algorithm in scalar C) reaches the reference's SLSQP cost
  on every recorded problem, and on every in-loop problem recorded above.
"""
import numpy as np
import pytest

import oracle
from golden_util import DIMS, PRESET, load, mixed_err, rel_err

CASES = ["3wrobot_RQL_quadratic_N10", "2tank_SQL_nomix_N8", "NI_RQL_quadlin_N5_period3", "NI_SQL_quadmix_N3_Ncritic6"]


def build(g):
    name = g["system"]
    n, m = DIMS[name]
    d = PRESET[name]
    s = oracle.make_sys(name, d["pars"], d["bnds"])
    c = oracle.make_ctrl(n, m, mode=g["mode"], Nactor=g["Nactor"], pred_step_size=d["dt"] * d["psm"], gamma=g["gamma"],
                         Ncritic=g["Ncritic"], buffer_size=g["buffer_size"], critic_struct=g["critic_struct"],
                         R1=d["R1_diag"], observation_target=d["target"])
    wb = (-1e3, 1e3) if g["critic_struct"] in ("quad-lin", "quad-mix") else (0.0, 1e3)
    return name, n, m, d, s, c, wb


@pytest.mark.parametrize("key", CASES)
def test_refit_loop_replaying_reference_weights(key):
    g = load("closed_loop_refit.json")[key]
    name, n, m, d, s, c, wb = build(g)
    w_rec = np.array([f["w"] for f in g["fits"]])
    rows = np.array(g["rows"])
    out = oracle.closed_loop_critic(c, s, [g["x0"]], np.array(g["cand"]), g["action_init"], d["dt"], 0.0, g["t1"], d["dt"] / 2,
                                    g["buffer_size"], wb, critic_period=g["critic_period"], w_replay=w_rec,
                                    traj_cap=len(rows) + 8)
    tr = out["traj"]
    assert len(tr) == len(rows)
    assert np.max(np.abs(tr[:, 0] - rows[:, 0])) <= 1e-15 * g["t1"], "solver times differ"   # numpy's pow vs the correctly rounded x**-0.2
    assert mixed_err(tr[:, 1:1 + n], rows[:, 1:1 + n], 1e-2) <= 1e-9
    assert np.array_equal(tr[:, 1 + n:1 + n + m], rows[:, 1 + n:1 + n + m]), "applied actions differ"
    assert rel_err(tr[:, 1 + n + m], rows[:, 1 + n + m]) <= 1e-9, "accumulated objective"
    assert np.array_equal(tr[:, 4 + n + m].astype(int), rows[:, -2].astype(int)), "controller sampling steps differ"
    fired = np.diff(np.concatenate([[0], tr[:, 5 + n + m]])).astype(int)
    assert np.array_equal(fired, rows[:, -1].astype(int)), "critic-clock firings differ"
    samp = tr[:, 4 + n + m] == 1
    assert np.array_equal(tr[samp, 2 + n + m].astype(int), np.array([p[0] for p in g["picks"]])), "arg-min picks differ"
    assert rel_err(tr[samp, 3 + n + m], [p[1] for p in g["picks"]]) <= 1e-9
    assert int(out["nfits"][0]) == len(g["fits"])
    assert mixed_err(out["obs_buf"][0], g["obs_buf_final"], 1e-2) <= 1e-9
    assert np.array_equal(out["act_buf"][0], np.array(g["act_buf_final"]))
    assert np.array_equal(out["w_critic"][0], np.array(g["w_final"]))


@pytest.mark.parametrize("key", CASES)
def test_fit_reaches_slsqp_cost_on_in_loop_problems(key):
    """Every refit the reference performed in the loop, handed to the restated fit: J_c(fit) <= J_c(SLSQP)."""
    g = load("closed_loop_refit.json")[key]
    name, n, m, d, s, c, wb = build(g)
    worse = []
    for i, f in enumerate(g["fits"]):
        w, J, evals = oracle.critic_fit(c, n, m, f["obs_buf"], f["act_buf"], f["w_prev"], wb[0], wb[1])
        assert np.all(w >= wb[0]) and np.all(w <= wb[1])
        Jchk = oracle.critic_cost(c, n, m, f["obs_buf"], f["act_buf"], w, f["w_prev"])
        assert abs(Jchk - J) <= 1e-9 * max(abs(J), 1e-300) + 1e-12 * abs(f["J_init"])
        assert J <= f["J_init"] * (1 + 1e-12)
        if not J <= f["J_fit"] * (1 + 1e-6) + 1e-9 * abs(f["J_init"]):
            worse.append((i, J, f["J_fit"]))
    assert not worse, worse[:5]


def test_fit_reaches_slsqp_cost_on_recorded_problems():
    cases = load("critic_fit.json")
    for k, g in enumerate(cases):
        name = g["system"]
        n, m = DIMS[name]
        c = oracle.make_ctrl(n, m, mode="RQL", Nactor=4, gamma=g["gamma"], Ncritic=g["Ncritic"] , buffer_size=10,
                             critic_struct=g["critic_struct"], R1=g["R1_diag"], observation_target=g["target"])
        w, J, evals = oracle.critic_fit(c, n, m, g["obs_buf"], g["act_buf"], g["w_prev"], g["Wmin"], g["Wmax"], w_init=g["w_init"])
        assert J <= g["J_ref"] * (1 + 1e-6) + 1e-9 * abs(g["J_init"]), (k, name, g["critic_struct"], g["regime"], J, g["J_ref"])
        assert J <= g["J_init"] * (1 + 1e-12)


def test_exact_line_search_matches_armijo_quality_with_fewer_passes():
    """The two line searches of the fit (Armijo backtracking in the one-lane kernels; unit step or exact minimiser along
    the Newton direction on the two-phase path) on every refit the reference performed in the config-3 loop: both stay at
    or below the reference's SLSQP cost, the exact search is never more than 1e-3 above the backtracking one (it is 6 % BELOW
    on one problem) and equal to 1e-6 in the median, and it never needs more than a quarter of the backtracking version's worst
    case."""
    g = load("closed_loop_refit.json")["3wrobot_RQL_quadratic_N10"]
    name, n, m, d, s, c, wb = build(g)
    ev = {0: [], 1: []}
    rel = []
    for f in g["fits"]:
        out = {}
        for mode in (0, 1):
            w, J, evals = oracle.critic_fit(c, n, m, f["obs_buf"], f["act_buf"], f["w_prev"], wb[0], wb[1], ls_mode=mode)
            assert J <= f["J_fit"] * (1 + 1e-6) + 1e-9 * abs(f["J_init"])
            ev[mode].append(evals)
            out[mode] = J
        rel.append((out[1] - out[0]) / max(abs(out[0]), 1e-9 * abs(f["J_init"]) + 1e-300))
    assert max(rel) <= 1e-3 and np.median(np.abs(rel)) <= 1e-6
    assert max(ev[1]) <= max(32, max(ev[0]) // 4) or max(ev[0]) <= 64, (max(ev[0]), max(ev[1]))
    assert np.mean(ev[1]) <= np.mean(ev[0]) * 1.05
