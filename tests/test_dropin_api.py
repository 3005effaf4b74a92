"""The drop-in Python surface (rcognita_b200.systems / simulator / controllers) driven exactly like the
reference's headless main loop (presets/main_3wrobot_NI.py:415-440), against the live-reference goldens,
plus the batched critic fit (rcg_critic_fit) against the reference's SLSQP result.

CPU part (no marker): constructor signatures are the reference's, foreign callables are rejected.
"""
import inspect

import numpy as np
import pytest

from golden_util import DIMS, PRESET, load, mixed_err, rel_err

# rcognita/controllers.py:811-837, rcognita/simulator.py:71-85, rcognita/systems.py:69-79 (names, order, defaults)
REF_CTRLOPTPRED_ARGS = [
    ("dim_input", None), ("dim_output", None), ("mode", "MPC"), ("ctrl_bnds", []), ("action_init", []), ("t0", 0),
    ("sampling_time", 0.1), ("Nactor", 1), ("pred_step_size", 0.1), ("sys_rhs", []), ("sys_out", []),
    ("state_sys", []), ("prob_noise_pow", 1), ("is_est_model", 0), ("model_est_stage", 1), ("model_est_period", 0.1),
    ("buffer_size", 20), ("model_order", 3), ("model_est_checks", 0), ("gamma", 1), ("Ncritic", 4),
    ("critic_period", 0.1), ("critic_struct", "quad-nomix"), ("stage_obj_struct", "quadratic"),
    ("stage_obj_pars", []), ("observation_target", [])]
REF_SIMULATOR_ARGS = [
    ("sys_type", None), ("closed_loop_rhs", None), ("sys_out", None), ("state_init", None), ("disturb_init", []),
    ("action_init", []), ("t0", 0), ("t1", 1), ("dt", 1e-2), ("max_step", 0.5e-2), ("first_step", 1e-6),
    ("atol", 1e-5), ("rtol", 1e-3), ("is_disturb", 0), ("is_dyn_ctrl", 0)]
REF_SYSTEM_ARGS = [
    ("sys_type", None), ("dim_state", None), ("dim_input", None), ("dim_output", None), ("dim_disturb", None),
    ("pars", []), ("ctrl_bnds", []), ("is_dyn_ctrl", 0), ("is_disturb", 0), ("pars_disturb", [])]


def _sig(fn):
    ps = list(inspect.signature(fn).parameters.values())[1:]
    return [(p.name, None if p.default is inspect.Parameter.empty else p.default) for p in ps]


def test_constructor_signatures_are_the_references():
    from rcognita_b200 import controllers, simulator, systems
    assert _sig(controllers.CtrlOptPred.__init__)[:len(REF_CTRLOPTPRED_ARGS)] == REF_CTRLOPTPRED_ARGS
    assert _sig(simulator.Simulator.__init__) == REF_SIMULATOR_ARGS
    assert _sig(systems.System.__init__)[:len(REF_SYSTEM_ARGS)] == REF_SYSTEM_ARGS
    assert list(inspect.signature(controllers.ctrl_selector).parameters) == [
        "t", "observation", "action_manual", "ctrl_nominal", "ctrl_benchmarking", "mode"]
    for cls, methods in ((controllers.CtrlOptPred, ["compute_action", "receive_sys_state", "stage_obj", "upd_accum_obj",
                                                    "reset", "_actor_cost", "_critic", "_critic_cost", "_actor_optimizer",
                                                    "_critic_optimizer"]),
                         (simulator.Simulator, ["sim_step", "get_sim_step_data", "reset"]),
                         (systems.System, ["_state_dyn", "out", "receive_action", "closed_loop_rhs"])):
        for mname in methods:
            assert callable(getattr(cls, mname)), (cls, mname)
    assert systems.Sys3WRobotNI.name == "3wrobotNI" and systems.Sys3WRobot.name == "3wrobot" and systems.Sys2Tank.name == "2tank"


def test_ctrl_selector_dispatch():
    from rcognita_b200.controllers import ctrl_selector

    class Stub:
        def __init__(self, v):
            self.v = v

        def compute_action(self, t, obs):
            return (self.v, t, obs)

    assert ctrl_selector(1.0, "o", "manual-action", Stub("n"), Stub("b"), "manual") == "manual-action"
    assert ctrl_selector(1.0, "o", None, Stub("n"), Stub("b"), "nominal") == ("n", 1.0, "o")
    for mode in ("MPC", "RQL", "SQL", "JACS"):
        assert ctrl_selector(2.0, "o", None, Stub("n"), Stub("b"), mode) == ("b", 2.0, "o")


def test_out_of_scope_controllers_raise():
    from rcognita_b200 import controllers
    for cls in (controllers.CtrlRLStab, controllers.CtrlNominal3WRobot):
        with pytest.raises(NotImplementedError):
            cls()


# ------------------------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")
gpu = pytest.mark.gpu

SYS_CLS = {"3wrobotNI": "Sys3WRobotNI", "3wrobot": "Sys3WRobot", "2tank": "Sys2Tank"}


@pytest.fixture(scope="module")
def rb():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    torch.cuda.set_device(0)
    from rcognita_b200 import controllers, simulator, systems
    return systems, simulator, controllers


def build_objects(rb, name, mode, Nactor, x0, t1, cand, critic_struct="quad-nomix", gamma=1.0, action_init=(),
                  buffer_size=10, Ncritic=4):
    """Same construction as presets/main_3wrobot_NI.py:214-316 (and the 3wrobot / 2tank presets)."""
    systems, simulator, controllers = rb
    cfg = PRESET[name]
    n, m = DIMS[name]
    bnds = np.array(cfg["bnds"], dtype=float)
    my_sys = getattr(systems, SYS_CLS[name])(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n,
                                              dim_disturb=2 if m == 2 else 1, pars=list(cfg["pars"]), ctrl_bnds=bnds,
                                              is_dyn_ctrl=0, is_disturb=0, pars_disturb=[])
    ctrl = controllers.CtrlOptPred(m, n, mode, ctrl_bnds=bnds, action_init=action_init, t0=0, sampling_time=cfg["dt"],
                                   Nactor=Nactor, pred_step_size=cfg["dt"] * cfg["psm"], sys_rhs=my_sys._state_dyn,
                                   sys_out=my_sys.out, state_sys=x0, prob_noise_pow=False, is_est_model=0,
                                   model_est_stage=1.0, model_est_period=cfg["dt"], buffer_size=buffer_size, model_order=5,
                                   model_est_checks=0, gamma=gamma, Ncritic=Ncritic, critic_period=cfg["dt"],
                                   critic_struct=critic_struct, stage_obj_struct="quadratic",
                                   stage_obj_pars=[np.diag(np.array(cfg["R1_diag"], dtype=float))],
                                   observation_target=cfg["target"], candidates=cand)
    sim = simulator.Simulator(sys_type="diff_eqn", closed_loop_rhs=my_sys.closed_loop_rhs, sys_out=my_sys.out,
                              state_init=x0, disturb_init=np.array([0, 0]), action_init=np.zeros(m), t0=0, t1=t1,
                              dt=cfg["dt"], max_step=cfg["dt"] / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0,
                              is_dyn_ctrl=0)
    return my_sys, ctrl, sim


@gpu
@pytest.mark.parametrize("key", ["NI_MPC_N6", "NI_MPC_N6_x1", "3wrobot_RQL_N10", "2tank_SQL_N8"])
def test_reference_main_loop_single_env(rb, key):
    """One environment, numpy in / numpy out, the loop body of presets/main_3wrobot_NI.py:415-440 verbatim:
    every row [t, state, action, accum_obj] of the live-reference run must be reproduced."""
    _, _, controllers = rb
    g = load("closed_loop.json")[key]
    name, mode = g["system"], g["mode"]
    n, m = DIMS[name]
    rows = np.array(g["rows"])
    my_sys, ctrl, sim = build_objects(rb, name, mode, g["Nactor"], np.array(g["x0"]), g["t1"], np.array(g["cand"]),
                                      critic_struct=g["critic_struct"], gamma=g["gamma"], action_init=g["action_init"])
    if g["w_fixed"] is not None:       # the golden run pinned the critic weights (make_golden.py closed_loop)
        wf = torch.as_tensor(np.array(g["w_fixed"]), device="cuda")[:, None]
        ctrl._critic_optimizer = lambda mask=None: wf.expand(-1, ctrl.num_envs).contiguous()
    k = 0
    nsamp = 0
    while True:
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, mode)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        ref = rows[k]
        assert t == ref[0] or abs(t - ref[0]) <= 1e-15 * g["t1"], k
        assert isinstance(state, np.ndarray) and state.shape == (n,) and action.shape == (m,)
        assert mixed_err(state_full, ref[1:1 + n], floor=1e-2) <= 1e-9, k
        assert np.array_equal(action, ref[1 + n:1 + n + m]), k
        assert rel_err(ctrl.accum_obj_val, ref[1 + n + m]) <= 1e-9, k
        nsamp += int(ref[-1])
        k += 1
        if t >= g["t1"]:
            break
    assert k == len(rows)
    assert int(ctrl.num_samples[0].item()) == len(g["picks"]) == nsamp
    assert sim.ODE_solver.status == "finished" and int(sim.ODE_solver.nfev) == g["nfev"]
    with pytest.raises(RuntimeError):
        sim.sim_step()                         # scipy base.py:189-191


@gpu
def test_reference_main_loop_batched_tensors(rb):
    """The same loop with E environments as CUDA tensors ([E, n] rows): lanes that start from the golden
    initial state reproduce the golden rows while other lanes run their own trajectories."""
    _, _, controllers = rb
    g = load("closed_loop.json")["NI_MPC_N6"]
    name, mode = g["system"], g["mode"]
    n, m = DIMS[name]
    rows = np.array(g["rows"])
    E = 33
    rng = np.random.default_rng(5)
    x0 = np.stack([rng.uniform(-10, 10, E), rng.uniform(-10, 10, E), rng.uniform(-np.pi, np.pi, E)], 1)
    x0[[0, 17, 32]] = g["x0"]
    x0_t = torch.as_tensor(x0, device="cuda")
    my_sys, ctrl, sim = build_objects(rb, name, mode, g["Nactor"], x0_t, 0.5, np.array(g["cand"]))
    for k in range(60):
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        assert isinstance(t, torch.Tensor) and t.shape == (E,) and state.shape == (E, n) and state.is_cuda
        action = controllers.ctrl_selector(t, observation, None, None, ctrl, mode)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        ref = rows[k]
        for lane in (0, 17, 32):
            assert abs(t[lane].item() - ref[0]) <= 1e-15
            assert mixed_err(state_full[lane].cpu().numpy(), ref[1:1 + n], floor=1e-2) <= 1e-9, (k, lane)
            assert np.array_equal(action[lane].cpu().numpy(), ref[1 + n:1 + n + m]), (k, lane)
            assert rel_err(ctrl.accum_obj_val[lane].item(), ref[1 + n + m]) <= 1e-9
    assert len(torch.unique(state_full[:, 0])) > 3


@gpu
def test_function_level_methods_match_goldens(rb):
    """stage_obj / _critic / _critic_cost / _actor_cost called like the reference's methods (App. A.1 values)."""
    name = "3wrobotNI"
    x0 = np.array([5, 5, -3 * np.pi / 4])
    U = np.random.default_rng(0).uniform(np.tile([-25.0, -5.0], 6), np.tile([25.0, 5.0], 6), size=(4, 12))
    my_sys, ctrl, _ = build_objects(rb, name, "MPC", 6, x0, 1.0, U)
    assert rel_err(my_sys._state_dyn([], x0, np.array([25.0, 5.0])), [-17.677669529663685, -17.67766952966369, 5.0]) <= 1e-12
    assert rel_err(ctrl.stage_obj(x0, np.array([25.0, 5.0])), 280.55165247561274) <= 1e-12
    assert rel_err(ctrl._actor_cost(ctrl.action_sqn_init, x0), 1712.8592297260025) <= 1e-9
    J = ctrl._actor_cost(U, x0)
    assert rel_err(J, [1683.274365086458, 1576.183747793757, 1563.836057710278, 1605.7555990810986]) <= 1e-9
    assert int(np.argmin(J)) == 2
    rbuf = np.random.default_rng(1)
    obs_buf = rbuf.normal(size=(10, 3)); act_buf = rbuf.uniform(-1, 1, size=(10, 2))
    for cs, dimc, q_ref, jc_ref in [("quad-lin", 20, 1217.0970145208614, 63.287885007843386),
                                    ("quadratic", 15, 1147.3381646032076, 80.4956769728908),
                                    ("quad-nomix", 5, 271.66549574268385, 122.06204560680067),
                                    ("quad-mix", 11, 758.9387307329714, 113.92951122558704)]:
        _, c2, _ = build_objects(rb, name, "RQL", 6, x0, 1.0, U, critic_struct=cs)
        assert c2.dim_critic == dimc
        w = np.arange(1, dimc + 1) / 10
        assert rel_err(c2._critic(x0, np.array([25.0, 5.0]), w), q_ref) <= 1e-12
        c2._obs_buf.copy_(torch.as_tensor(obs_buf, device="cuda")[:, :, None])
        c2._act_buf.copy_(torch.as_tensor(act_buf, device="cuda")[:, :, None])
        assert rel_err(c2._critic_cost(w), jc_ref) <= 1e-9


@gpu
def test_rql_sql_compute_action_buffers_and_refit(rb):
    """compute_action in RQL mode: FIFO pushes use the PREVIOUS action and the current observation
    (controllers.py:1463-1464), only sampling lanes push, refits keep the weights inside [Wmin, Wmax] and never
    increase _critic_cost relative to w_critic_init."""
    name = "2tank"
    E = 64
    rng = np.random.default_rng(3)
    x0 = rng.uniform(-2, 2, size=(E, 2))
    cand = rng.uniform(0, 1, size=(32, 8))
    my_sys, ctrl, sim = build_objects(rb, name, "SQL", 8, torch.as_tensor(x0, device="cuda"), 5.0, cand, action_init=[0.5])
    from rcognita_b200.controllers import ctrl_selector
    prev_action = ctrl.action_curr.clone()
    pushes = 0
    for k in range(120):
        sim.sim_step()
        t, state, observation, state_full = sim.get_sim_step_data()
        before = ctrl.num_samples.clone()
        buf_before = ctrl._obs_buf.clone()
        action = ctrl_selector(t, observation, None, None, ctrl, "SQL").clone()
        fired = (ctrl.num_samples - before).bool()
        if fired.any():
            pushes += 1
            # newest row of the fired lanes = (observation, previous action); older rows shifted up by one
            assert torch.equal(ctrl._obs_buf[-1][:, fired], observation.t()[:, fired])
            assert torch.equal(ctrl._act_buf[-1][:, fired], prev_action.t()[:, fired])
            assert torch.equal(ctrl._obs_buf[:-1][:, :, fired], buf_before[1:][:, :, fired])
        assert torch.equal(ctrl._obs_buf[:, :, ~fired], buf_before[:, :, ~fired])
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        prev_action = action
    assert pushes >= 20
    w = ctrl._w_critic
    assert float(w.min()) >= 0.0 and float(w.max()) <= 1e3
    J_fit = ctrl._critic_cost(ctrl.w_critic)
    J_init = ctrl._critic_cost(torch.ones_like(ctrl.w_critic))
    assert bool((J_fit <= J_init * (1 + 1e-12) + 1e-300).all())
    assert bool((w != 1.0).any())


@gpu
def test_critic_fit_reaches_reference_slsqp_cost(rb):
    """rcg_critic_fit vs the reference's `_critic_optimizer` (SLSQP) on the committed buffers: the fitted cost,
    re-evaluated by the CPU oracle's _critic_cost, is <= the reference's fitted cost; weights stay in the box."""
    import oracle
    from rcognita_b200 import _C, ops
    cases = load("critic_fit.json")
    worse = []
    for c in cases:
        n, m = DIMS[c["system"]]
        obj = _C.make_objective(n, m, mode="RQL", Nactor=4, gamma=c["gamma"], Ncritic=c["Ncritic"], buffer_size=10,
                                critic_struct=c["critic_struct"], R1=c["R1_diag"], observation_target=c["target"])
        E = 3                                   # the same problem on three lanes, the middle one masked out
        ob = torch.as_tensor(np.array(c["obs_buf"]), device="cuda")[:, :, None].expand(-1, -1, E).contiguous()
        ac = torch.as_tensor(np.array(c["act_buf"]), device="cuda")[:, :, None].expand(-1, -1, E).contiguous()
        wp = torch.as_tensor(np.array(c["w_prev"]), device="cuda")[:, None].expand(-1, E).contiguous()
        w = torch.full((len(c["w_init"]), E), 7.0, dtype=torch.float64, device="cuda")
        w_init = torch.as_tensor(np.array(c["w_init"]), device="cuda")
        mask = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
        Jc = torch.full((E,), -1.0, dtype=torch.float64, device="cuda")
        wp0 = wp.clone()
        ops.critic_fit(obj, n, m, ob, ac, wp, c["Wmin"], c["Wmax"], w, w_init=w_init, mask=mask, Jc_out=Jc)
        assert torch.equal(wp, wp0)                                                          # update_prev off
        wh = w.cpu().numpy()
        assert np.all(wh[:, 1] == 7.0) and Jc[1].item() == -1.0                              # masked lane untouched
        assert np.array_equal(wh[:, 0], wh[:, 2])
        assert wh.min() >= c["Wmin"] and wh.max() <= c["Wmax"]
        oc = oracle.make_ctrl(n, m, mode="RQL", Nactor=4, gamma=c["gamma"], critic_struct=c["critic_struct"],
                              R1=c["R1_diag"], observation_target=c["target"], Ncritic=c["Ncritic"], buffer_size=10)
        J_fit = oracle.critic_cost(oc, n, m, np.array(c["obs_buf"]), np.array(c["act_buf"]), wh[:, 0].copy(),
                                   np.array(c["w_prev"]))
        assert abs(J_fit - Jc[0].item()) <= 1e-6 * max(J_fit, 1e-9 * c["J_init"]) + 1e-18, (J_fit, Jc[0].item())
        assert J_fit <= c["J_init"] * (1 + 1e-12)
        if not J_fit <= c["J_ref"] * (1 + 1e-6) + 1e-9 * c["J_init"]:
            worse.append((c["system"], c["critic_struct"], c["regime"], c["gamma"], J_fit, c["J_ref"]))
    assert not worse, worse


@gpu
@pytest.mark.parametrize("name,mode,cs,N,t1", [("2tank", "SQL", "quad-nomix", 8, 4.0), ("3wrobot", "RQL", "quadratic", 10, 0.25),
                                               ("3wrobotNI", "RQL", "quad-mix", 5, 0.3)])
def test_engine_with_critic_fit_equals_class_loop(rb, name, mode, cs, N, t1):
    """BASELINE configs 3/4 (RQL / SQL with critic buffer fitting): the fused ClosedLoopEngine (rk45_advance +
    push_buffers + critic_fit + actor_cost per control interval) and the reference-style loop over the drop-in
    classes take identical decisions: same step counts, times, actions, states, returns and critic weights."""
    from rcognita_b200.controllers import ctrl_selector
    from rcognita_b200.engine import ClosedLoopEngine
    n, m = DIMS[name]
    cfg = PRESET[name]
    E = 40
    rng = np.random.default_rng(21)
    box = {"3wrobotNI": ([-5, -5, -3], [5, 5, 3]), "3wrobot": ([-5, -5, -3, -1, -1], [5, 5, 3, 1, 1]), "2tank": ([-2, -2], [2, 2])}[name]
    x0 = rng.uniform(box[0], box[1], size=(E, n))
    b = np.array(cfg["bnds"], dtype=float)
    cand = rng.uniform(np.tile(b[:, 0], N), np.tile(b[:, 1], N), size=(48, N * m))
    a_init = [0.5] if name == "2tank" else []
    eng = ClosedLoopEngine(name, x0, cand, pars=cfg["pars"], ctrl_bnds=cfg["bnds"], mode=mode, Nactor=N, dt=cfg["dt"],
                           pred_step_size=cfg["dt"] * cfg["psm"], t1=t1, R1=cfg["R1_diag"], observation_target=cfg["target"],
                           critic_struct=cs, critic_fit=True, Ncritic=4, buffer_size=10, action_init=a_init)
    eng.run()
    got = eng.results()
    assert np.all(got["nfits"] == got["nsamples"]) and got["nfits"].min() >= 5       # critic_period == sampling_time
    lo_w = -1e3 if cs in ("quad-lin", "quad-mix") else 0.0
    assert got["w_critic"].min() >= lo_w and got["w_critic"].max() <= 1e3 and np.any(got["w_critic"] != 1.0)

    my_sys, ctrl, sim = build_objects(rb, name, mode, N, torch.as_tensor(x0, device="cuda"), t1, cand, critic_struct=cs,
                                      action_init=a_init)
    nsteps = torch.zeros(E, dtype=torch.int64, device="cuda")
    for _ in range(100000):
        running = torch.as_tensor([s == "running" for s in sim.ODE_solver.status], device="cuda")
        if not bool(running.any()):
            break
        sim.sim_step()
        nsteps += running
        t, state, observation, state_full = sim.get_sim_step_data()
        if not bool(running.all()):
            # finished lanes must not be driven any further (the reference loop breaks at t >= t1)
            keep_a, keep_acc = ctrl._action_curr.clone(), ctrl._accum.clone()
        action = ctrl_selector(t, observation, None, None, ctrl, mode)
        my_sys.receive_action(action)
        ctrl.receive_sys_state(my_sys._state)
        ctrl.upd_accum_obj(observation, action)
        if not bool(running.all()):
            ctrl._action_curr[:, ~running] = keep_a[:, ~running]
            ctrl._accum[~running] = keep_acc[~running]
    assert np.array_equal(nsteps.cpu().numpy(), got["nsteps"])
    assert np.array_equal(ctrl.num_samples.cpu().numpy(), got["nsamples"])
    assert np.array_equal(sim._t.cpu().numpy(), got["t"])
    assert np.array_equal(sim._y.t().cpu().numpy(), got["y"])
    assert rel_err(ctrl._accum.cpu().numpy(), got["accum"]) <= 1e-12       # fused vs separate accumulation kernels
    assert np.array_equal(ctrl._w_critic.t().cpu().numpy(), got["w_critic"])


@gpu
def test_simulator_masked_reset(rb):
    """Simulator.reset(mask=...): the selected environments restart exactly like a freshly constructed simulator
    (state, t0, first_step, f(t0, y0) with the lane's current action), the others are left untouched bit for bit."""
    import oracle
    from rcognita_b200 import simulator, systems
    bn = np.array([[-25, 25], [-5, 5]], dtype=float)
    rng = np.random.default_rng(4)
    E = 40
    x0 = np.stack([rng.uniform(-5, 5, E), rng.uniform(-5, 5, E), rng.uniform(-3, 3, E)], 1)

    def make():
        sy = systems.Sys3WRobotNI(sys_type="diff_eqn", dim_state=3, dim_input=2, dim_output=3, dim_disturb=2, pars=[],
                                  ctrl_bnds=bn, is_dyn_ctrl=0, is_disturb=0, pars_disturb=[])
        sim = simulator.Simulator("diff_eqn", sy.closed_loop_rhs, sy.out, x0, disturb_init=[], action_init=np.zeros(2), t0=0,
                                  t1=1.0, dt=0.01, max_step=0.005, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0, is_dyn_ctrl=0)
        return sy, sim
    sy, sim = make()
    act = rng.uniform(bn[:, 0] * 1.4, bn[:, 1] * 1.4, size=(E, 2))        # some outside the bounds
    for _ in range(7):
        sim.sim_step()
        sy.receive_action(act)
    before = [t.clone() for t in (sim._y, sim._f, sim._t, sim._h, sim._status, sim._nfev)]
    mask = np.zeros(E, dtype=bool)
    mask[[1, 5, 6, 33]] = True
    sim.reset(mask=mask)
    keep = torch.as_tensor(~mask, device="cuda")
    for a, b in zip(before, (sim._y, sim._f, sim._t, sim._h, sim._status, sim._nfev)):
        assert torch.equal(a[..., keep], b[..., keep])
    m = torch.as_tensor(mask, device="cuda")
    assert torch.equal(sim._y[:, m].t().cpu(), torch.as_tensor(x0[mask]))
    assert bool((sim._t[m] == 0).all()) and bool((sim._h[m] == 1e-6).all()) and bool((sim._nfev[m] == 1).all())
    s = oracle.make_sys("3wrobotNI", [], bn)
    for e in np.flatnonzero(mask):
        f = oracle.closed_loop_rhs(s, x0[e], act[e].copy())
        f = f[0] if isinstance(f, tuple) else f
        assert np.max(np.abs(sim._f[:, e].cpu().numpy() - np.asarray(f))) <= 1e-12 * max(1.0, np.max(np.abs(f)))
    for _ in range(3):
        sim.sim_step()                                                   # restarted and running lanes step on together
    assert bool((sim._status == 0).all())
