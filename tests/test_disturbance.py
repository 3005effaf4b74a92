"""Disturbance lanes, ``System(is_disturb=1)`` (rcognita/systems.py:228-231, :247-248, :316-318,
:373-376, :325-345, :384-394, :421-424).

Function level is pinned to the LIVE reference (``tests/golden/disturb.json``: ``_state_dyn`` with a list-valued ``disturb``,
``_disturb_dyn`` under a patched ``randn``).  The reference's own closed loop cannot run with ``is_disturb=1`` under numpy 2
(``ndarray != []`` raises at systems.py:316 / :373) and draws from numpy's global stream, so the disturbed LOOP has no
reference golden: there the CUDA kernels are compared with the CPU oracle's restatement (scipy's RK45 on the full state +
the specified per-environment normal stream), bit for bit.
"""
import numpy as np
import pytest

import oracle
from golden_util import DIMS, load, rel_err

SYSTEMS = ["3wrobotNI", "3wrobot", "2tank"]
SYS_CLS = {"3wrobotNI": "Sys3WRobotNI", "3wrobot": "Sys3WRobot", "2tank": "Sys2Tank"}


# ------------------------------------------------------------------------------------------- CPU: oracle vs live reference
@pytest.mark.parametrize("name", SYSTEMS)
def test_oracle_disturbed_functions_match_reference(name):
    g = load("disturb.json")[name]
    n, m = DIMS[name]
    s = oracle.make_sys(name, g["pars"], g["bnds"])
    d = oracle.make_dist(g["pars_disturb"])
    assert g["dim_full_state"] == n + g["dim_disturb"] == n + oracle.DIST_DIM[s.sys_id]
    for c in g["cases"]:
        # numpy's sin / cos and the specified sincos may differ in the last bit (same bar as test_oracle_golden.py)
        assert rel_err(oracle.state_dyn_disturbed(s, c["state"], c["action"], c["disturb"]), c["d_state"]) <= 4e-16
        assert np.array_equal(oracle.disturb_dyn(s, d, c["disturb"], c["z"]), np.array(c["d_disturb"]))
        full, clipped = oracle.closed_loop_rhs_disturbed(s, d, c["state"] + c["disturb"], c["action"], z=c["z"])
        assert rel_err(full[:n], c["d_state"]) <= 4e-16 and np.array_equal(full[n:], np.array(c["d_disturb"]))
        assert np.array_equal(clipped[:m], np.array(c["action"]))


def test_normal_stream_is_standard_normal_and_keyed():
    z = np.array([oracle.normal2(11, e, c) for e in range(400) for c in range(50)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert abs(((z - z.mean()) ** 4).mean() / z.var() ** 2 - 3) < 0.15            # kurtosis
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.03
    assert not np.array_equal(oracle.normal2(11, 3, 5), oracle.normal2(12, 3, 5))
    assert not np.array_equal(oracle.normal2(11, 3, 5), oracle.normal2(11, 4, 5))
    assert not np.array_equal(oracle.normal2(11, 3, 5), oracle.normal2(11, 3, 6))
    assert np.array_equal(oracle.normal2(11, (1 << 33) + 3, 5), oracle.normal2(11, (1 << 33) + 3, 5))
    xs = np.random.default_rng(0).uniform(1e-19, 1, 4000)
    assert max(abs(oracle.det_log(x) - np.log(x)) / abs(np.log(x)) for x in xs) <= 4e-16


def test_oracle_disturbed_rk45_mean_reverts():
    """Sanity of the restated loop: with sigma = 0 the disturbance decays like exp(-tau t) and the state feels it."""
    s = oracle.make_sys("3wrobotNI", [], [[-25, 25], [-5, 5]])
    d = oracle.make_dist([[0.0, 0.0], [0.0, 0.0], [0.3, 0.45]], seed=1)
    r = oracle.RK45Disturbed(s, d, [1.0, 2.0, 0.3, 0.5, -0.4], 0.0, 1.0, 0.005)
    while r.status == "running":
        r.step()
    y = r.y
    assert abs(y[3] - 0.5 * np.exp(-0.3)) < 1e-4 and abs(y[4] + 0.4 * np.exp(-0.45)) < 1e-4
    assert abs(y[0] - (1.0 + 0.5 / 0.3 * (1 - np.exp(-0.3)))) < 1e-3            # zero action: dx/dt = disturb[0]


# ------------------------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@gpu
@pytest.mark.parametrize("name", SYSTEMS)
def test_kernels_match_reference_at_function_level(cuda, name):
    """rcg_rhs_disturbed with the draws GIVEN == the reference's _state_dyn(.., disturb) / _disturb_dyn, through the
    drop-in System methods too."""
    from rcognita_b200 import _C, ops, systems
    g = load("disturb.json")[name]
    n, m = DIMS[name]
    nd = g["dim_disturb"]
    cases = g["cases"]
    E = len(cases)
    sysd = _C.make_system(name, g["pars"], g["bnds"])
    distd = _C.make_disturb(g["pars_disturb"], seed=5)
    y = torch.as_tensor(np.array([c["state"] + c["disturb"] for c in cases]).T.copy(), device=cuda)
    a = torch.as_tensor(np.array([c["action"] for c in cases]).T.copy(), device=cuda)
    z = torch.zeros((2, E), dtype=torch.float64, device=cuda)
    z[:nd] = torch.as_tensor(np.array([c["z"] for c in cases]).T.copy(), device=cuda)
    out = ops.rhs_disturbed(sysd, distd, y, a.clone(), normals=z).cpu().numpy().T
    want = np.array([c["d_state"] + c["d_disturb"] for c in cases])
    assert np.max(np.abs(out - want) / np.maximum(np.abs(want), 1e-300)) <= 4e-16
    my_sys = getattr(systems, SYS_CLS[name])(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n, dim_disturb=nd,
                                              pars=list(g["pars"]), ctrl_bnds=np.array(g["bnds"], dtype=float), is_dyn_ctrl=0,
                                              is_disturb=1, pars_disturb=np.array(g["pars_disturb"]))
    assert my_sys._dim_full_state == g["dim_full_state"]
    c = cases[3]
    got = my_sys._state_dyn([], np.array(c["state"]), np.array(c["action"]), disturb=list(c["disturb"]))
    assert np.max(np.abs(got - np.array(c["d_state"])) / np.maximum(np.abs(np.array(c["d_state"])), 1e-300)) <= 4e-16
    got = my_sys._disturb_dyn([], np.array(c["disturb"]), normals=np.array(c["z"]))
    assert np.max(np.abs(got - np.array(c["d_disturb"]))) <= 4e-16 * max(1.0, np.max(np.abs(c["d_disturb"])))
    batch = my_sys._state_dyn([], np.array([k["state"] for k in cases]), np.array([k["action"] for k in cases]),
                              disturb=np.array([k["disturb"] for k in cases]))
    assert np.max(np.abs(batch - want[:, :n]) / np.maximum(np.abs(want[:, :n]), 1e-300)) <= 4e-16


@gpu
def test_normal_stream_matches_oracle_bit_for_bit(cuda):
    from rcognita_b200 import _C, ops
    E = 1000
    for seed, off, call in ((3, 0, 0), (3, 12345, 7), ((1 << 40) + 9, (1 << 34) + 5, 123456)):
        distd = _C.make_disturb([[1, 1], [0, 0], [1, 1]], seed=seed, env_offset=off)
        z = ops.disturb_normals(distd, E, call, device=cuda).cpu().numpy()
        want = np.array([oracle.normal2(seed, off + e, call) for e in range(E)]).T
        assert np.array_equal(z, want), (seed, off, call)
    calls = torch.arange(E, dtype=torch.int32, device=cuda) * 3
    z = ops.disturb_normals(_C.make_disturb([[1, 1], [0, 0], [1, 1]], seed=8), E, calls).cpu().numpy()
    assert np.array_equal(z, np.array([oracle.normal2(8, e, 3 * e) for e in range(E)]).T)
    big = ops.disturb_normals(_C.make_disturb([[1, 1], [0, 0], [1, 1]], seed=1), 1 << 20, 4, device=cuda)
    assert abs(float(big.mean())) < 3e-3 and abs(float(big.std()) - 1) < 3e-3


@gpu
@pytest.mark.parametrize("name", SYSTEMS)
def test_disturbed_rk45_matches_oracle_bit_for_bit(cuda, name):
    """Simulator.sim_step of a disturbed system (rcg_rk45_step_disturbed, through the drop-in classes, action changed every
    few steps, some out of bounds) against the oracle lane by lane: identical t, h_abs, nfev, status and full state on
    every step -- the noisy right-hand side makes the error estimate reject often, all of it reproduced."""
    from rcognita_b200 import simulator, systems
    g = load("disturb.json")[name]
    n, m = DIMS[name]
    nd = g["dim_disturb"]
    bn = np.array(g["bnds"], dtype=float)
    rng = np.random.default_rng(17)
    E, off, seed = 48, 4096, 21
    x0 = rng.uniform(-2, 2, size=(E, n))
    q0 = rng.normal(size=nd) * 0.3
    my_sys = getattr(systems, SYS_CLS[name])(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n, dim_disturb=nd,
                                              pars=list(g["pars"]), ctrl_bnds=bn, is_dyn_ctrl=0, is_disturb=1,
                                              pars_disturb=np.array(g["pars_disturb"]), seed=seed, env_offset=off)
    dt = 0.01
    sim = simulator.Simulator("diff_eqn", my_sys.closed_loop_rhs, my_sys.out, x0, disturb_init=q0, action_init=np.zeros(m), t0=0,
                              t1=0.25, dt=dt, max_step=dt / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=1, is_dyn_ctrl=0)
    s = oracle.make_sys(name, g["pars"], g["bnds"])
    d = oracle.make_dist(g["pars_disturb"], seed=seed)
    lanes = [oracle.RK45Disturbed(s, d, np.concatenate([x0[e], q0]), 0.0, 0.25, dt / 2, env=off + e) for e in range(E)]
    assert np.array_equal(sim._f.cpu().numpy().T, np.array([r.f for r in lanes]))          # RHS call 0 at construction
    rejections = 0
    for k in range(400):
        running = [r.status == "running" for r in lanes]
        if not any(running):
            break
        sim.sim_step()
        for e, r in enumerate(lanes):
            if running[e]:
                before = r.nfev
                r.step()
                rejections += (r.nfev - before) // 6 - 1
        t, state, observation, state_full = sim.get_sim_step_data()
        assert np.array_equal(np.asarray(t), np.array([r.t for r in lanes])), k
        assert np.array_equal(state_full, np.array([r.y for r in lanes])), k
        assert np.array_equal(state, state_full[:, :n]) and state_full.shape[1] == n + nd
        assert np.array_equal(sim._h.cpu().numpy(), np.array([r.h_abs for r in lanes]))
        assert np.array_equal(sim._nfev.cpu().numpy(), np.array([r.nfev for r in lanes]))
        assert list(sim.ODE_solver.status) == [r.status for r in lanes]
        if k % 3 == 0:
            act = rng.uniform(bn[:, 0] * 1.3, bn[:, 1] * 1.3, size=(E, m))
            my_sys.receive_action(act)
            for e, r in enumerate(lanes):
                r.receive_action(act[e])
    assert k >= 50                                           # at least fifty solver steps compared on every lane
    if name != "3wrobot":                                    # (force / moment noise keeps Sys3WRobot's steps tiny: 400 steps < t1)
        assert not any(r.status == "running" for r in lanes)
    if name != "2tank":
        assert rejections > 0


@gpu
def test_disturbed_advance_is_sharding_invariant_and_equals_stepping(cuda):
    """rcg_rk45_advance_disturbed (fused steps up to the sampling event) == repeated rcg_rk45_step_disturbed with the
    clock test on the host, and a batch split in two with env_offset reproduces the unsplit run bit for bit."""
    from rcognita_b200 import _C, ops
    name, n, m, nd = "3wrobotNI", 3, 2, 2
    E = 2048
    rng = np.random.default_rng(3)
    sysd = _C.make_system(name, [], [[-25, 25], [-5, 5]])
    sol = _C.make_solver(1.0, 0.005, 1e-3, 1e-5)
    obj = _C.make_objective(n, m, mode="MPC", Nactor=3, pred_step_size=0.01, R1=[1, 10, 1, 0, 0])
    pd = [[2.0, 1.5], [0.0, 0.1], [0.3, 0.45]]
    y0 = np.concatenate([rng.uniform(-3, 3, size=(n, E)), 0.2 * rng.normal(size=(nd, E))], axis=0)
    a0 = rng.uniform([-25, -5], [25, 5], size=(E, m)).T.copy()

    def run(lo, hi, fused):
        Es = hi - lo
        distd = _C.make_disturb(pd, seed=77, env_offset=lo)
        dev = dict(dtype=torch.float64, device=cuda)
        y = torch.as_tensor(y0[:, lo:hi].copy(), device=cuda)
        a = torch.as_tensor(a0[:, lo:hi].copy(), device=cuda)
        f = torch.empty_like(y)
        t, h = torch.zeros(Es, **dev), torch.full((Es,), 1e-6, **dev)
        clock, accum = torch.zeros(Es, **dev), torch.zeros(Es, **dev)
        status = torch.zeros(Es, dtype=torch.int32, device=cuda)
        nfev = torch.ones(Es, dtype=torch.int32, device=cuda)
        nsteps = torch.zeros(Es, dtype=torch.int32, device=cuda)
        flag = torch.zeros(Es, dtype=torch.int32, device=cuda)
        state_sys = torch.zeros((n, Es), **dev)
        ops.rhs_disturbed(sysd, distd, y, a, call=None, out=f)
        if fused:
            for _ in range(3):
                ops.rk45_advance_disturbed(sysd, distd, sol, obj, y, f, t, h, status, a, clock, 0.01, 1 << 30, nfev,
                                           state_sys=state_sys, accum=accum, sample_flag=flag, nsteps=nsteps)
        else:
            for _ in range(3):
                todo = torch.ones(Es, dtype=torch.bool, device=cuda)
                while bool(todo.any()):
                    # step only the lanes that have not reached their sampling event: park the others
                    st_save = status.clone()
                    status[~todo] = _C.FINISHED
                    yprev = y.clone()
                    ops.rk45_step_disturbed(sysd, distd, sol, y, f, t, h, status, a, nfev)
                    status.copy_(torch.where(todo, status, st_save))
                    nsteps += todo.int()
                    fired = todo & (t - clock >= 0.01)
                    clock.copy_(torch.where(fired, t, clock))
                    so = ops.stage_obj(obj, n, m, y[:n].contiguous(), a)
                    accum += torch.where(todo & ~fired, so * 0.01, torch.zeros_like(so))
                    state_sys.copy_(torch.where(fired[None, :], yprev[:n], torch.where(todo[None, :], y[:n], state_sys)))
                    todo = todo & ~fired & (status == _C.RUNNING)
        return [v.cpu().numpy() for v in (y, f, t, h, nfev, nsteps, clock, accum, state_sys)]

    whole = run(0, E, True)
    parts = [run(0, 1024, True), run(1024, E, True)]
    for w, p0, p1 in zip(whole, parts[0], parts[1]):
        assert np.array_equal(w, np.concatenate([p0, p1], axis=-1))
    stepped = run(0, 256, False)
    for w, s_ in zip(whole, stepped):
        assert np.array_equal(w[..., :256], s_)
    assert whole[5].min() >= 6                              # several steps per interval were taken


@gpu
def test_engine_with_disturbance_equals_class_loop_and_is_sharding_invariant(cuda):
    """ClosedLoopEngine(pars_disturb=...) -- MPC over a candidate table on disturbed Sys3WRobotNI environments, fused
    rk45_advance_disturbed + actor launches -- takes exactly the steps of the reference-style loop over the drop-in classes
    (System(is_disturb=1) / Simulator(is_disturb=1) / CtrlOptPred), and a batch split in two with env_offset reproduces it."""
    from rcognita_b200 import controllers, simulator, systems
    from rcognita_b200.engine import ClosedLoopEngine
    name, n, m, nd = "3wrobotNI", 3, 2, 2
    bn = np.array([[-25, 25], [-5, 5]], dtype=float)
    pd = np.array([[0.4, 0.3], [0.0, 0.1], [0.3, 0.45]])
    rng = np.random.default_rng(8)
    E, N, C_, t1, dt, seed = 96, 5, 48, 0.12, 0.01, 31
    x0 = rng.uniform([-5, -5, -3], [5, 5, 3], size=(E, n))
    q0 = np.array([0.2, -0.1])
    cand = rng.uniform(np.tile(bn[:, 0], N), np.tile(bn[:, 1], N), size=(C_, N * m))
    kw = dict(ctrl_bnds=bn, mode="MPC", Nactor=N, dt=dt, t1=t1, R1=[1, 10, 1, 0, 0], pars_disturb=pd, disturb_init=q0, seed=seed)
    eng = ClosedLoopEngine(name, x0, cand, **kw)
    eng.run()
    got = eng.results()
    assert got["disturb"].shape == (E, nd) and (got["status"] == 1).all()

    parts = [ClosedLoopEngine(name, x0[a:b], cand, env_offset=a, **kw) for a, b in ((0, 32), (32, E))]
    for p in parts:
        p.run()
    for k in ("y", "disturb", "t", "accum", "nsteps", "nsamples", "argmin", "nfev"):
        assert np.array_equal(got[k], np.concatenate([p.results()[k] for p in parts], axis=0)), k

    my_sys = systems.Sys3WRobotNI(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n, dim_disturb=nd, pars=[], ctrl_bnds=bn,
                                  is_dyn_ctrl=0, is_disturb=1, pars_disturb=pd, seed=seed)
    ctrl = controllers.CtrlOptPred(m, n, "MPC", ctrl_bnds=bn, action_init=[], t0=0, sampling_time=dt, Nactor=N, pred_step_size=dt,
                                   sys_rhs=my_sys._state_dyn, sys_out=my_sys.out, state_sys=x0, gamma=1, stage_obj_struct="quadratic",
                                   stage_obj_pars=[np.diag([1.0, 10, 1, 0, 0])], observation_target=[], candidates=cand)
    sim = simulator.Simulator("diff_eqn", my_sys.closed_loop_rhs, my_sys.out, x0, disturb_init=q0, action_init=np.zeros(m), t0=0, t1=t1,
                              dt=dt, max_step=dt / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=1, is_dyn_ctrl=0)
    nsteps = np.zeros(E, dtype=np.int64)
    for _ in range(100000):
        running = np.array([st == "running" for st in sim.ODE_solver.status])
        if not running.any():
            break
        sim.sim_step()
        nsteps += running
        t, state, observation, state_full = sim.get_sim_step_data()
        keep_a, keep_acc = ctrl._action_curr.clone(), ctrl._accum.clone()
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, "MPC")
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        if not running.all():                      # finished lanes are no longer driven (the reference loop breaks at t >= t1)
            done = torch.as_tensor(~running, device=cuda)
            ctrl._action_curr[:, done] = keep_a[:, done]
            ctrl._accum[done] = keep_acc[done]
    assert np.array_equal(nsteps, got["nsteps"])
    assert np.array_equal(sim._t.cpu().numpy(), got["t"])
    assert np.array_equal(sim._y.t().cpu().numpy(), np.concatenate([got["y"], got["disturb"]], axis=1))
    assert np.max(np.abs(ctrl._accum.cpu().numpy() - got["accum"]) / np.abs(got["accum"])) <= 1e-12
