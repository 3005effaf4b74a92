"""GPU parity tests: librcg_b200.so (through its C ABI / the torch launchers) against the CPU
oracle and the committed live-reference golden vectors.

Tolerances (BASELINE.json north_star): _actor_cost / _critic_cost 1e-9 relative in fp64;
RK45 closed-loop trajectories 1e-6 relative (we assert 1e-9, and exact step times / counts);
arg-min indices bit-exact.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from golden_util import DIMS, PRESET, load, mixed_err, rel_err  # noqa: E402

COST_RTOL = 1e-9
SYSTEMS = ["3wrobotNI", "3wrobot", "2tank"]
X0 = {"3wrobotNI": [5, 5, -3 * np.pi / 4], "3wrobot": [5, 5, -3 * np.pi / 4, 0.3, -0.2], "2tank": [2, -2]}


@pytest.fixture(scope="module")
def rb():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import rcognita_b200
    from rcognita_b200 import _C, ops
    torch.cuda.set_device(0)
    return rcognita_b200, _C, ops


@pytest.fixture(scope="module")
def fn():
    return load("functions.json")


def dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dtype or torch.float64)


def soa(rows):
    """[E, d] row layout (reference) -> [d, E] component-major device tensor."""
    return dev(np.asarray(rows, dtype=np.float64).T.copy())


def random_states(name, E, seed=0):
    rng = np.random.default_rng(seed)
    if name == "3wrobotNI":
        return np.stack([rng.uniform(-10, 10, E), rng.uniform(-10, 10, E), rng.uniform(-np.pi, np.pi, E)], 1)
    if name == "3wrobot":
        return np.stack([rng.uniform(-10, 10, E), rng.uniform(-10, 10, E), rng.uniform(-np.pi, np.pi, E),
                         rng.uniform(-1, 1, E), rng.uniform(-1, 1, E)], 1)
    return np.stack([rng.uniform(-2, 2, E), rng.uniform(-2, 2, E)], 1)


def random_cands(name, shape_prefix, N, seed=1):
    b = np.array(PRESET[name]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    return np.random.default_rng(seed).uniform(lo, hi, size=tuple(shape_prefix) + (N * b.shape[0],))


# ------------------------------------------------------------------ function level vs golden

@pytest.mark.parametrize("name", SYSTEMS)
def test_state_dyn_and_rhs_golden(rb, fn, name):
    _, _C, ops = rb
    d = fn[name]
    sysd = _C.make_system(name, d["pars"], d["bnds"])
    cases = d["cases"]["state_dyn"]
    out = ops.state_dyn(sysd, soa([c["state"] for c in cases]), soa([c["action"] for c in cases]))
    assert rel_err(out.T.cpu().numpy(), [c["out"] for c in cases]) <= 1e-14
    cases = d["cases"]["closed_loop_rhs"]
    act = soa([c["action"] for c in cases])
    out = ops.rhs(sysd, soa([c["state"] for c in cases]), act)
    assert rel_err(out.T.cpu().numpy(), [c["out"] for c in cases]) <= 1e-14
    assert np.array_equal(act.T.cpu().numpy(), np.array([c["action_clipped"] for c in cases]))   # in-place clip


def test_deterministic_sincos_bit_exact(rb):
    """The fp64 kernels' sin/cos must equal the oracle's deterministic sincos BIT FOR BIT (it feeds
    the adaptive step size and with it the controller's sampling pattern).  Sys3WRobotNI with
    action (1, 0): _state_dyn = (cos theta, sin theta, 0)."""
    _, _C, ops = rb
    rng = np.random.default_rng(5)
    th = np.concatenate([rng.uniform(-30, 30, 60000), rng.uniform(-1e4, 1e4, 20000), rng.normal(size=20000) * 1e-3,
                         np.arange(-16, 17) * (np.pi / 4), [0.0, 1e-300, 99999.0, 1.0e5]])
    E = th.size
    sysd = _C.make_system("3wrobotNI", [], [])
    state = np.zeros((3, E)); state[2] = th
    act = np.zeros((2, E)); act[0] = 1.0
    out = ops.state_dyn(sysd, dev(state), dev(act)).cpu().numpy()
    ref = np.array([oracle.sincos(x) for x in th])
    assert np.array_equal(out[0], ref[:, 1]) and np.array_equal(out[1], ref[:, 0]) and np.all(out[2] == 0.0)


@pytest.mark.parametrize("name", SYSTEMS)
def test_stage_obj_critic_golden(rb, fn, name):
    _, _C, ops = rb
    n, m = DIMS[name]
    for c in fn[name]["cases"]["stage_obj"]:
        obj = _C.make_objective(n, m, R1=c["R1"], R2=c["R2"], stage_obj_struct=c["struct"], observation_target=c["target"])
        got = ops.stage_obj(obj, n, m, soa([c["obs"]]), soa([c["act"]])).item()
        assert rel_err(got, c["out"]) <= COST_RTOL
    for c in fn[name]["cases"]["critic"]:
        obj = _C.make_objective(n, m, critic_struct=c["critic_struct"], observation_target=c["target"])
        assert _C.dim_critic(c["critic_struct"], n, m) == c["dim_critic"]
        got = ops.critic(obj, n, m, soa([c["obs"]]), soa([c["act"]]), dev(c["w"])).item()
        assert rel_err(got, c["out"]) <= COST_RTOL
        got = ops.critic(obj, n, m, soa([c["obs"]] * 3), soa([c["act"]] * 3), soa([c["w"]] * 3), w_per_env=True)
        assert rel_err(got.cpu().numpy(), [c["out"]] * 3) <= COST_RTOL


@pytest.mark.parametrize("name", SYSTEMS)
def test_critic_cost_golden(rb, fn, name):
    _, _C, ops = rb
    n, m = DIMS[name]
    for c in fn[name]["cases"]["critic_cost"]:
        obj = _C.make_objective(n, m, mode="RQL", critic_struct=c["critic_struct"], gamma=c["gamma"], Ncritic=4,
                                buffer_size=10, R1=c["R1_diag"], observation_target=c["target"])
        E, W = 3, 2
        ob = dev(np.repeat(np.asarray(c["obs_buf"])[:, :, None], E, axis=2))          # [L, n, E]
        ab = dev(np.repeat(np.asarray(c["act_buf"])[:, :, None], E, axis=2))
        w = dev(np.broadcast_to(np.asarray(c["w"])[:, None, None], (len(c["w"]), E, W)).copy())
        wp = dev(np.repeat(np.asarray(c["w_prev"])[:, None], E, axis=1))
        got = ops.critic_cost(obj, n, m, ob, ab, w, wp).cpu().numpy()
        assert rel_err(got, np.full((E, W), c["out"])) <= COST_RTOL


@pytest.mark.parametrize("name", SYSTEMS)
@pytest.mark.parametrize("per_env", [False, True])
def test_actor_cost_golden(rb, fn, name, per_env):
    _, _C, ops = rb
    n, m = DIMS[name]
    d = fn[name]
    sysd = _C.make_system(name, d["pars"], d["bnds"])
    for c in d["cases"]["actor_cost"]:
        obj = _C.make_objective(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                                critic_struct=c["critic_struct"], R1=c["R1"], observation_target=c["target"])
        cand = np.asarray(c["cand"])                                                  # [C, N*m]
        C_ = cand.shape[0]
        E = 2 if per_env else 1
        if per_env:
            cdev = dev(np.tile(cand.T[:, None, :], (1, E, 1)).reshape(cand.shape[1], E * C_))
        else:
            cdev = dev(cand.T.copy())
        w = None if c["w"] is None else dev(c["w"])
        J, am, Jmin = ops.actor_cost(sysd, obj, soa([c["state_sys"]] * E), soa([c["obs"]] * E), cdev, per_env, C_,
                                     w_critic=w)
        J = J.cpu().numpy()
        for e in range(E):
            assert rel_err(J[e], c["J"]) <= COST_RTOL, (c["mode"], c["critic_struct"], c["N"])
            assert int(am[e]) == c["argmin"]
            assert J[e, 4] == J[e, 1]                                                # duplicated candidate: exact tie
            assert Jmin[e].item() == J[e, c["argmin"]]


# ------------------------------------------------------------------ vs the oracle, seeded random lanes

@pytest.mark.parametrize("name,mode,cs,N", [
    ("3wrobotNI", "MPC", "quad-nomix", 6), ("3wrobotNI", "RQL", "quad-lin", 5), ("3wrobotNI", "SQL", "quad-mix", 7),
    ("3wrobot", "RQL", "quadratic", 10), ("3wrobot", "SQL", "quad-nomix", 4), ("3wrobot", "MPC", "quad-nomix", 12),
    ("2tank", "SQL", "quad-nomix", 8), ("2tank", "RQL", "quad-mix", 3), ("2tank", "MPC", "quad-nomix", 1),
])
@pytest.mark.parametrize("C_,per_env,w_per_env", [(256, False, False), (48, True, True), (1000, False, True), (7, True, False)])
def test_actor_cost_vs_oracle(rb, name, mode, cs, N, C_, per_env, w_per_env):
    _, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    E = 37
    gamma = 0.95
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    R1 = np.diag(p["R1_diag"]).astype(float)
    if cs == "quad-mix":                                     # also exercise a dense R1
        R1 = R1 + 0.01 * np.arange((n + m) ** 2).reshape(n + m, n + m)
    kw = dict(mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], gamma=gamma, critic_struct=cs, R1=R1,
              observation_target=p["target"])
    obj = _C.make_objective(n, m, **kw)
    s = oracle.make_sys(name, p["pars"], p["bnds"])
    c = oracle.make_ctrl(n, m, **kw)
    obs = random_states(name, E, 3)
    xs = obs + 0.01 * np.random.default_rng(4).normal(size=obs.shape)
    cand = random_cands(name, (E, C_) if per_env else (C_,), N, 5)
    dimc = _C.dim_critic(cs, n, m)
    w = np.random.default_rng(6).uniform(0, 2, size=(E, dimc) if w_per_env else (dimc,))
    mask = np.ones(E, dtype=np.int32)
    mask[5] = 0
    mask[E - 1] = 0
    if per_env:
        cdev = dev(np.transpose(cand, (2, 0, 1)).reshape(N * m, E * C_))
    else:
        cdev = dev(cand.T.copy())
    action_out = torch.full((m, E), -777.0, device="cuda", dtype=torch.float64)
    accum = torch.full((E,), 0.5, device="cuda", dtype=torch.float64)
    J, am, Jmin = ops.actor_cost(sysd, obj, soa(xs), soa(obs), cdev, per_env, C_,
                                 w_critic=dev(w.T.copy() if w_per_env else w), w_per_env=w_per_env,
                                 mask=dev(mask, torch.int32), action_out=action_out, accum=accum, sampling_time=0.01)
    J, am, Jmin = J.cpu().numpy(), am.cpu().numpy(), Jmin.cpu().numpy()
    action_out, accum = action_out.T.cpu().numpy(), accum.cpu().numpy()
    for e in range(E):
        if not mask[e]:
            assert am[e] == -1 and np.isnan(Jmin[e]) and np.all(action_out[e] == -777.0) and accum[e] == 0.5
            continue
        tab = cand[e] if per_env else cand
        Jr, ar = oracle.actor_cost_table(c, s, tab, obs[e], xs[e], w[e] if w_per_env else w)
        assert rel_err(J[e], Jr) <= COST_RTOL
        assert am[e] == int(np.argmin(J[e]))                 # bit-exact arg-min of the GPU's own costs
        if am[e] != ar:                                      # oracle may differ only on a sub-tolerance near-tie
            assert abs(Jr[am[e]] - Jr[ar]) <= COST_RTOL * abs(Jr[ar])
        assert Jmin[e] == J[e, am[e]]
        assert np.array_equal(action_out[e], tab[am[e], :m])
        assert rel_err(accum[e], 0.5 + oracle.stage_obj(c, n, m, obs[e], tab[am[e], :m]) * 0.01) <= COST_RTOL


@pytest.mark.parametrize("name,mode,cs,N,C_", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, 256), ("3wrobotNI", "RQL", "quad-lin", 5, 96), ("3wrobotNI", "RQL", "quad-mix", 10, 32),
    ("3wrobot", "RQL", "quadratic", 10, 256), ("3wrobot", "MPC", "quad-nomix", 7, 64), ("3wrobot", "RQL", "quad-nomix", 3, 160),
    ("3wrobotNI", "MPC", "quad-nomix", 6, 16), ("3wrobotNI", "RQL", "quadratic", 4, 8),      # several environments per warp
])
def test_actor_cost_shared_table_kernel(rb, name, mode, cs, N, C_):
    """actor_cost_tab_kernel (shared candidate table on a robot, the presets' lean objective: the candidate part of the
    heading is tabulated once per launch, csrc/actor_tab.cuh) against the oracle and against actor_cost_kernel on the same
    inputs: costs to 1e-12 relative (the sums of the heading are formed in a different order, ~1e-15), arg-min of the
    kernel's own costs, action hand-over, accumulated objective, masked environments untouched, a candidate with a non-finite
    action costs NaN and wins the arg-min like np.argmin -- and the dispatch rule (batches of >= 1,024 environments)."""
    rcg, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    E = 1100
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    kw = dict(mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], gamma=1.0, critic_struct=cs, R1=p["R1_diag"])
    obj = _C.make_objective(n, m, **kw)
    s = oracle.make_sys(name, p["pars"], p["bnds"])
    c = oracle.make_ctrl(n, m, **kw)
    obs = random_states(name, E, 13)
    xs = obs + 0.01 * np.random.default_rng(14).normal(size=obs.shape)
    cand = random_cands(name, (C_,), N, 15)
    dimc = _C.dim_critic(cs, n, m)
    w = np.random.default_rng(16).uniform(0, 2, size=(E, dimc))
    mask = np.ones(E, dtype=np.int32)
    mask[[5, 700, E - 1]] = 0

    def run(table):
        cdev = dev(table.T.copy())
        action_out = torch.full((m, E), -777.0, device="cuda", dtype=torch.float64)
        accum = torch.full((E,), 0.5, device="cuda", dtype=torch.float64)
        J, am, Jmin = ops.actor_cost(sysd, obj, soa(xs), soa(obs), cdev, False, C_, w_critic=dev(w.T.copy()), w_per_env=True,
                                     mask=dev(mask, torch.int32), action_out=action_out, accum=accum, sampling_time=0.01)
        return J.cpu().numpy(), am.cpu().numpy(), Jmin.cpu().numpy(), action_out.T.cpu().numpy(), accum.cpu().numpy()

    J, am, Jmin, act, accum = run(cand)
    assert rcg.last_actor_kernel() == "actor_cost_tab_kernel"
    os.environ["RCG_ACTOR_NO_TABLE"] = "1"
    try:
        J2, am2, Jmin2, act2, accum2 = run(cand)
        assert rcg.last_actor_kernel() == "actor_cost_kernel"
    finally:
        os.environ.pop("RCG_ACTOR_NO_TABLE", None)
    live = mask != 0
    assert np.max(np.abs(J[live] - J2[live]) / np.maximum(np.abs(J2[live]), 1e-300)) <= 1e-12
    assert np.mean(am[live] == am2[live]) >= 0.999                                   # a flip needs a 1e-15 near-tie
    for e in list(range(0, E, 97)) + [5, 700, E - 1]:
        if not mask[e]:
            assert am[e] == -1 and np.isnan(Jmin[e]) and np.all(act[e] == -777.0) and accum[e] == 0.5
            continue
        Jr, ar = oracle.actor_cost_table(c, s, cand, obs[e], xs[e], w[e])
        assert rel_err(J[e], Jr) <= COST_RTOL
        assert am[e] == int(np.argmin(J[e])) and Jmin[e] == J[e, am[e]]
        if am[e] != ar:
            assert abs(Jr[am[e]] - Jr[ar]) <= COST_RTOL * abs(Jr[ar])
        assert np.array_equal(act[e], cand[am[e], :m])
        assert rel_err(accum[e], 0.5 + oracle.stage_obj(c, n, m, obs[e], cand[am[e], :m]) * 0.01) <= COST_RTOL
    # a non-finite action anywhere in a sequence -- also in the last stage, which never enters the rollout -- costs NaN
    bad = cand.copy()
    bad[3, (N - 1) * m] = np.inf
    bad[C_ - 2, 1] = np.nan
    Jb, amb, _, _, _ = run(bad)
    assert np.isnan(Jb[live][:, 3]).all() and np.isnan(Jb[live][:, C_ - 2]).all() and np.all(amb[live] == 3)
    keep = np.ones(C_, dtype=bool)
    keep[[3, C_ - 2]] = False
    assert np.array_equal(Jb[live][:, keep], J[live][:, keep])
    # small batches stay on the direct kernel (one launch instead of two)
    few = soa(xs[:64])
    ops.actor_cost(sysd, obj, few, few, dev(cand.T.copy()), False, C_, w_critic=dev(w[:64].T.copy()), w_per_env=True, want_J=False)
    assert rcg.last_actor_kernel() == "actor_cost_kernel"


@pytest.mark.parametrize("name,mode,cs,N,E,C_", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, 4098, 256), ("3wrobotNI", "MPC", "quad-nomix", 6, 3000, 96),
    ("3wrobotNI", "MPC", "quad-nomix", 5, 2502, 8), ("3wrobot", "RQL", "quadratic", 10, 1500, 64),
    ("2tank", "SQL", "quad-nomix", 8, 6000, 32), ("2tank", "MPC", "quad-nomix", 3, 7001, 2),
    ("3wrobotNI", "SQL", "quad-lin", 6, 1000, 1024),
    # horizons without a compile-time specialisation (anything outside 3..10): the runtime-horizon TMA kernel (boxes of 4
    # stages; the last box of Nactor = 13, 50, 1 reaches past the array and is zero-filled by the tensor map); Nactor = 7
    # has had its own instantiation since round 2
    ("3wrobotNI", "MPC", "quad-nomix", 7, 3001, 256), ("3wrobot", "RQL", "quadratic", 20, 700, 64),
    ("2tank", "SQL", "quad-nomix", 1, 5000, 32), ("3wrobotNI", "SQL", "quad-lin", 50, 300, 16),
    ("3wrobot", "MPC", "quad-nomix", 13, 1203, 8), ("2tank", "RQL", "quad-mix", 12, 900, 96),
])
def test_actor_tma_kernel_matches_direct_kernel(rb, name, mode, cs, N, E, C_):
    """The TMA-staged kernel (per-env candidates, compile-time horizon, C a multiple of 32 or a power of two
    below 32) against the direct-load kernel: same source arithmetic (only FMA contraction may differ between the
    two compilations: <= 1e-12 relative), identical arg-min up to such ties, with a ragged mask, many environment
    groups per warp, several environments per warp (C < 32) and a last partial warp."""
    import os
    _, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], gamma=0.97, critic_struct=cs,
                            R1=p["R1_diag"], observation_target=p["target"])
    obs = random_states(name, E, 11)
    xs = obs + 0.01 * np.random.default_rng(12).normal(size=obs.shape)
    g = torch.Generator(device="cuda").manual_seed(13)
    b = torch.tensor(p["bnds"], device="cuda", dtype=torch.float64)
    cdev = torch.empty((N * m, E * C_), device="cuda", dtype=torch.float64)
    for k in range(N * m):
        j = k % m
        cdev[k] = b[j, 0] + (b[j, 1] - b[j, 0]) * torch.rand((E * C_,), device="cuda", dtype=torch.float64, generator=g)
    dimc = _C.dim_critic(cs, n, m)
    w = dev(np.random.default_rng(14).uniform(0, 2, size=(dimc, E)))
    mask = (np.random.default_rng(15).uniform(size=E) < 0.8).astype(np.int32)
    outs = []
    for no_pipe in (False, True):
        if no_pipe:
            os.environ["RCG_ACTOR_NO_TMA"] = "1"
        else:
            os.environ.pop("RCG_ACTOR_NO_TMA", None)
        try:
            action_out = torch.full((m, E), -777.0, device="cuda", dtype=torch.float64)
            accum = torch.full((E,), 0.25, device="cuda", dtype=torch.float64)
            J = torch.full((E, C_), -1.0, device="cuda", dtype=torch.float64)
            _, am, Jmin = ops.actor_cost(sysd, obj, soa(xs), soa(obs), cdev, True, C_, w_critic=w, w_per_env=True,
                                         mask=dev(mask, torch.int32), J_out=J, action_out=action_out, accum=accum,
                                         sampling_time=0.01)
            torch.cuda.synchronize()
            outs.append([t.cpu().numpy() for t in (J, am, Jmin, action_out, accum)])
        finally:
            os.environ.pop("RCG_ACTOR_NO_TMA", None)
    (J, am, Jmin, act, acc), (J2, am2, Jmin2, act2, acc2) = outs
    assert rel_err(J, J2) <= 1e-12 and rel_err(acc, acc2) <= 1e-12
    diff = np.flatnonzero(am != am2)
    for e in diff:                              # only exact-to-rounding ties may pick differently
        assert abs(J2[e, am[e]] - J2[e, am2[e]]) <= 1e-12 * abs(J2[e, am2[e]])
    assert len(diff) <= max(1, E // 500)
    same = am == am2
    assert np.array_equal(act[:, same], act2[:, same])
    on = mask.astype(bool)
    assert np.all(am[~on] == -1) and np.all(J[~on] == -1.0)
    assert np.array_equal(am[on], np.argmin(J[on], axis=1))
    # spot-check a few lanes against the oracle
    s = oracle.make_sys(name, p["pars"], p["bnds"])
    c = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], gamma=0.97, critic_struct=cs,
                         R1=p["R1_diag"], observation_target=p["target"])
    ch = cdev.cpu().numpy().reshape(N * m, E, C_)
    wh = w.cpu().numpy()
    for e in list(np.flatnonzero(on)[:3]) + [int(np.flatnonzero(on)[-1])]:
        Jr, _ = oracle.actor_cost_table(c, s, ch[:, e, :].T.copy(), obs[e], xs[e], wh[:, e].copy())
        assert rel_err(J[e], Jr) <= COST_RTOL


@pytest.mark.parametrize("pred_step", [0.01, 0.02, 0.06, 0.2])
@pytest.mark.parametrize("per_env", [False, True])
def test_actor_cost_heading_rotation_paths(rb, pred_step, per_env):
    """The predictor advances (sin, cos) of the heading by rotation: truncated polynomials for |h * omega| <= 1/8 (the
    presets), the full small-angle polynomials up to pi/4 (pred_step 0.06: most candidates), a full sincos beyond
    (0.2).  All three paths must give the reference's cost to 1e-9, in the TMA-staged and the direct kernel."""
    _, _C, ops = rb
    name, N, C_, E = "3wrobotNI", 6, 64, 300
    p = PRESET[name]
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    kw = dict(mode="MPC", Nactor=N, pred_step_size=pred_step, R1=p["R1_diag"])
    obj = _C.make_objective(3, 2, **kw)
    s = oracle.make_sys(name, p["pars"], p["bnds"]); c = oracle.make_ctrl(3, 2, **kw)
    obs = random_states(name, E, 81)
    cand = random_cands(name, (E, C_) if per_env else (C_,), N, 82)
    cd = dev(cand.transpose(2, 0, 1).reshape(2 * N, E * C_).copy()) if per_env else dev(cand.T.copy())
    J, am, _ = ops.actor_cost(sysd, obj, soa(obs), soa(obs), cd, per_env, C_)
    J, am = J.cpu().numpy(), am.cpu().numpy()
    for e in range(0, E, 7):
        Jr, ar = oracle.actor_cost_table(c, s, cand[e] if per_env else cand, obs[e], obs[e])
        assert rel_err(J[e], Jr) <= COST_RTOL, (e, pred_step)
        assert am[e] == ar or abs(Jr[am[e]] - Jr[ar]) <= 1e-12 * abs(Jr[ar])


def test_argmin_ties_and_nan(rb):
    """np.argmin semantics: first minimum wins; NaN counts as minimal (first NaN wins)."""
    _, _C, ops = rb
    p = PRESET["3wrobotNI"]
    sysd = _C.make_system("3wrobotNI", p["pars"], p["bnds"])
    N, C_ = 6, 700
    obj = _C.make_objective(3, 2, mode="MPC", Nactor=N, pred_step_size=0.01, R1=p["R1_diag"])
    base = random_cands("3wrobotNI", (1,), N, 9)[0]
    cand = np.tile(base, (C_, 1))
    # R1 puts zero weight on actions and a[N-1] never enters the dynamics: candidates differing only
    # in the last action tie EXACTLY -> index 0 must win.
    cand[:, -2:] = np.random.default_rng(2).uniform(-1, 1, size=(C_, 2))
    x = soa([X0["3wrobotNI"]])
    J, am, _ = ops.actor_cost(sysd, obj, x, x, dev(cand.T.copy()), False, C_)
    assert np.unique(J.cpu().numpy()).size == 1 and int(am[0]) == 0
    cand2 = random_cands("3wrobotNI", (C_,), N, 10)
    cand2[300:, :] = cand2[299, :]                       # the minimum's duplicates come later
    J, am, _ = ops.actor_cost(sysd, obj, x, x, dev(cand2.T.copy()), False, C_)
    assert int(am[0]) == int(np.argmin(J.cpu().numpy()[0]))
    cand3 = cand2.copy()
    cand3[613, 0] = np.nan
    cand3[401, 2] = np.nan
    J, am, Jmin = ops.actor_cost(sysd, obj, x, x, dev(cand3.T.copy()), False, C_)
    assert int(am[0]) == 401 and np.isnan(Jmin[0].item())
    # a non-finite value in the LAST action never enters the dynamics and carries zero weight, yet the reference's
    # 0 * a * a makes the cost NaN: the lean kernels (zero action weights dropped) must keep that, per-env candidates too
    s = oracle.make_sys("3wrobotNI", p["pars"], p["bnds"])
    ct = oracle.make_ctrl(3, 2, mode="MPC", Nactor=N, pred_step_size=0.01, R1=p["R1_diag"])
    for bad in (np.nan, np.inf, -np.inf):
        cand4 = cand2[:640].copy()
        cand4[77, -1] = bad
        cand4[500, -2] = bad
        assert np.isnan(oracle.actor_cost(ct, s, cand4[77], X0["3wrobotNI"], X0["3wrobotNI"]))
        for per_env in (False, True):
            J, am, Jmin = ops.actor_cost(sysd, obj, x, x, dev(cand4.T.copy()), per_env, 640)
            Jh = J.cpu().numpy()[0]
            assert np.isnan(Jh[77]) and np.isnan(Jh[500]) and np.isfinite(np.delete(Jh, [77, 500])).all()
            assert int(am[0]) == 77 and np.isnan(Jmin[0].item())


@pytest.mark.parametrize("name", SYSTEMS)
def test_critic_cost_vs_oracle(rb, name):
    _, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    E, W, L = 19, 5, 10
    rng = np.random.default_rng(11)
    for cs in ["quad-lin", "quadratic", "quad-nomix", "quad-mix"]:
        kw = dict(mode="RQL", critic_struct=cs, gamma=0.9, Ncritic=4, buffer_size=L, R1=p["R1_diag"],
                  observation_target=p["target"])
        obj = _C.make_objective(n, m, **kw)
        c = oracle.make_ctrl(n, m, **kw)
        dimc = _C.dim_critic(cs, n, m)
        ob = rng.normal(size=(E, L, n)); ab = rng.uniform(-1, 1, size=(E, L, m))
        w = rng.uniform(0, 3, size=(E, W, dimc)); wp = rng.uniform(0, 3, size=(E, dimc))
        got = ops.critic_cost(obj, n, m, dev(np.transpose(ob, (1, 2, 0)).copy()), dev(np.transpose(ab, (1, 2, 0)).copy()),
                              dev(np.transpose(w, (2, 0, 1)).copy()), dev(wp.T.copy())).cpu().numpy()
        for e in range(E):
            for k in range(W):
                assert rel_err(got[e, k], oracle.critic_cost(c, n, m, ob[e], ab[e], w[e, k], wp[e])) <= COST_RTOL


def test_push_buffers(rb):
    _, _C, ops = rb
    n, m, L, E = 3, 2, 10, 9
    rng = np.random.default_rng(0)
    ob = rng.normal(size=(L, n, E)); ab = rng.normal(size=(L, m, E))
    o = rng.normal(size=(n, E)); a = rng.normal(size=(m, E))
    mask = (np.arange(E) % 3 != 0).astype(np.int32)
    obd, abd = dev(ob), dev(ab)
    ops.push_buffers(n, m, obd, abd, dev(o), dev(a), dev(mask, torch.int32))
    exp_o = ob.copy(); exp_a = ab.copy()
    for e in range(E):
        if mask[e]:                                                # utilities.push_vec
            exp_o[:, :, e] = np.vstack([ob[1:, :, e], o[:, e]])
            exp_a[:, :, e] = np.vstack([ab[1:, :, e], a[:, e]])
    assert np.array_equal(obd.cpu().numpy(), exp_o) and np.array_equal(abd.cpu().numpy(), exp_a)


# ------------------------------------------------------------------ RK45 integrator

def _sim_state(ops, _C, sysd, x0_rows, t0=0.0, first_step=1e-6):
    E = len(x0_rows)
    y = soa(x0_rows)
    n, m = _C.SYS_DIMS[sysd.sys_id]
    action = torch.zeros((m, E), device="cuda", dtype=torch.float64)
    f = ops.rhs(sysd, y, action)
    t = torch.full((E,), t0, device="cuda", dtype=torch.float64)
    h = torch.full((E,), first_step, device="cuda", dtype=torch.float64)
    status = torch.zeros((E,), device="cuda", dtype=torch.int32)
    nfev = torch.ones((E,), device="cuda", dtype=torch.int32)
    return y, f, t, h, status, nfev, action


@pytest.mark.parametrize("key", ["3wrobotNI:inbounds", "3wrobotNI:outofbounds", "3wrobot:inbounds",
                                 "3wrobot:outofbounds", "2tank:inbounds", "2tank:outofbounds"])
def test_rk45_step_golden_traces(rb, key):
    """SURVEY.md App. A.3 protocol against the live reference's recorded solver traces, with 33
    identical lanes (a full warp + 1): every lane must reproduce t, y, f, h_abs and nfev per step."""
    _, _C, ops = rb
    g = load("integrator.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    d = PRESET[name]
    E = 33
    sysd = _C.make_system(name, d["pars"], d["bnds"])
    sol = _C.make_solver(g["t1"], d["dt"] / 2)
    y, f, t, h, status, nfev, action = _sim_state(ops, _C, sysd, [X0[name]] * E)
    sched = np.array(g["sched"]); rows = np.array(g["rows"])
    for k in range(1, len(rows) + 1):
        ops.rk45_step(sysd, sol, y, f, t, h, status, action, nfev=nfev)
        action.copy_(dev(sched[(k // 5) % 4])[:, None].expand(m, E))
        ref = rows[k - 1]
        tt = t.cpu().numpy()
        assert np.all(tt == tt[0])
        assert abs(tt[0] - ref[0]) <= 1e-13 * max(abs(ref[0]), 1e-6), (k, tt[0], ref[0])
        yy = y.cpu().numpy(); ff = f.cpu().numpy()
        assert np.all(yy == yy[:, :1]) and np.all(ff == ff[:, :1])
        assert mixed_err(yy[:, 0], ref[1:1 + n], floor=1e-3) <= 1e-9, k
        assert mixed_err(ff[:, 0], ref[1 + n:1 + 2 * n], floor=1e-3) <= 1e-9, k
        assert rel_err(h.cpu().numpy()[0], ref[1 + 2 * n]) <= 1e-9, k
        assert np.all(nfev.cpu().numpy() == int(ref[2 + 2 * n])), k
    st = status.cpu().numpy()
    assert np.all(st == {"running": 0, "finished": 1, "failed": 2}[g["status"]])
    # stepping a finished lane is a no-op (scipy raises; the batched kernel skips non-running lanes)
    y_before = y.clone()
    ops.rk45_step(sysd, sol, y, f, t, h, status, action, nfev=nfev)
    assert torch.equal(y, y_before)


@pytest.mark.parametrize("name", SYSTEMS)
def test_rk45_step_vs_oracle_random_lanes(rb, name):
    """Different lanes, lane-specific action jumps (-> lane-specific rejections): every lane must
    match its own scalar oracle solver, including nfev and exact t."""
    _, _C, ops = rb
    n, m = DIMS[name]
    d = PRESET[name]
    E, steps = 70, 60
    x0 = random_states(name, E, 21)
    sysd = _C.make_system(name, d["pars"], d["bnds"])
    s = oracle.make_sys(name, d["pars"], d["bnds"])
    t1 = d["dt"] * 12
    sol = _C.make_solver(t1, d["dt"] / 2)
    y, f, t, h, status, nfev, action = _sim_state(ops, _C, sysd, x0)
    solvers = [oracle.RK45(s, x0[e], 0.0, t1, d["dt"] / 2) for e in range(E)]
    b = np.array(d["bnds"], dtype=float)
    rng = np.random.default_rng(22)
    for k in range(steps):
        ops.rk45_step(sysd, sol, y, f, t, h, status, action, nfev=nfev)
        for r in solvers:
            if r.status == "running":
                r.step()
        if k % 4 == 3:                                       # new (partly out-of-bounds) actions per lane
            a = rng.uniform(1.5 * b[:, 0], 1.5 * b[:, 1], size=(E, m))
            action.copy_(soa(a))
            for e, r in enumerate(solvers):
                r.receive_action(a[e])
        tt, yy, hh, nn, ss = t.cpu().numpy(), y.cpu().numpy(), h.cpu().numpy(), nfev.cpu().numpy(), status.cpu().numpy()
        for e, r in enumerate(solvers):
            assert tt[e] == r.t, (k, e)
            assert mixed_err(yy[:, e], r.y, floor=1e-3) <= 1e-9, (k, e)
            assert rel_err(hh[e], r.h_abs) <= 1e-9
            assert nn[e] == r.nfev
            assert _C.STATUS_NAMES[int(ss[e])] == r.status
    assert np.all(status.cpu().numpy() == _C.FINISHED)


# ------------------------------------------------------------------ closed loop (engine)

def _engine_for(g, name, x0_rows, cand, dtype=None):
    from rcognita_b200.engine import ClosedLoopEngine
    d = PRESET[name]
    return ClosedLoopEngine(name, x0_rows, cand, pars=d["pars"], ctrl_bnds=d["bnds"], mode=g["mode"], Nactor=g["Nactor"],
                            dt=d["dt"], pred_step_size=d["dt"] * d["psm"], t1=g["t1"], gamma=g["gamma"],
                            R1=d["R1_diag"], observation_target=d["target"], critic_struct=g["critic_struct"],
                            w_critic=g["w_fixed"], action_init=g["action_init"],
                            dtype=dtype or torch.float64)


@pytest.mark.parametrize("key", ["NI_MPC_N6", "NI_MPC_N6_x1", "3wrobot_RQL_N10", "2tank_SQL_N8"])
def test_closed_loop_golden(rb, key):
    """the live reference's closed loop with its own _actor_cost on a candidate
    table + np.argmin, vs the fused GPU loop (rk45_advance + actor_cost), sampled at every
    controller sample: t, state, picked index, J_min; and the episode totals."""
    _, _C, ops = rb
    g = load("closed_loop.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    rows = np.array(g["rows"]); picks = np.array(g["picks"])
    E = 5
    eng = _engine_for(g, name, [g["x0"]] * E, np.array(g["cand"]))
    sampled_rows = rows[rows[:, -1] > 0]
    k = 0
    while not eng.all_done():
        eng.run_interval()
        fl = eng.sample_flag.cpu().numpy()
        assert np.all(fl == fl[0])
        if fl[0]:
            ref = sampled_rows[k]
            assert np.all(eng.t.cpu().numpy() == ref[0]) or abs(eng.t[0].item() - ref[0]) <= 1e-15 * g["t1"]
            yy = eng.y.cpu().numpy()
            assert np.all(yy == yy[:, :1])
            assert mixed_err(yy[:, 0], ref[1:1 + n], floor=1e-2) <= 1e-9, k
            assert np.all(eng.argmin.cpu().numpy() == int(picks[k, 0])), k
            assert rel_err(eng.Jmin.cpu().numpy(), np.full(E, picks[k, 1])) <= COST_RTOL
            assert np.array_equal(eng.action.cpu().numpy()[:, 0], ref[1 + n:1 + n + m])
            assert rel_err(eng.accum[0].item(), ref[1 + n + m]) <= 1e-9
            k += 1
    res = eng.results()
    assert k == len(picks)
    assert np.all(res["nsteps"] == len(rows)) and np.all(res["nsamples"] == len(picks)) and np.all(res["nfev"] == g["nfev"])
    assert mixed_err(res["y"][0], rows[-1, 1:1 + n], floor=1e-2) <= 1e-9
    assert rel_err(res["accum"], np.full(E, rows[-1, 1 + n + m])) <= 1e-9
    assert np.all(res["status"] == _C.FINISHED)


@pytest.mark.parametrize("name,mode,cs,N,t1,per_env", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, 0.5, True), ("3wrobotNI", "MPC", "quad-nomix", 6, 0.5, False),
    ("3wrobot", "RQL", "quadratic", 10, 0.3, False), ("2tank", "SQL", "quad-nomix", 8, 5.0, True),
])
def test_closed_loop_vs_oracle_many_envs(rb, name, mode, cs, N, t1, per_env):
    """Different environments desynchronise (lane-specific sampling events, rejections): every
    env must match its own scalar oracle episode: exact step/sample counts and final time,
    state and accumulated objective to 1e-9."""
    _, _C, ops = rb
    from rcognita_b200.engine import ClosedLoopEngine
    n, m = DIMS[name]
    p = PRESET[name]
    E, C_ = 200, 64
    x0 = random_states(name, E, 31)
    cand = random_cands(name, (E, C_) if per_env else (C_,), N, 32)
    dimc = _C.dim_critic(cs, n, m)
    w = None if mode == "MPC" else np.random.default_rng(33).uniform(0, 2, size=dimc)
    b = np.array(p["bnds"], dtype=float)
    a0 = b[:, 0] / 10 if name != "2tank" else np.array([0.5])
    eng = ClosedLoopEngine(name, x0, cand, pars=p["pars"], ctrl_bnds=p["bnds"], mode=mode, Nactor=N, dt=p["dt"],
                           pred_step_size=p["dt"] * p["psm"], t1=t1, R1=p["R1_diag"], observation_target=p["target"],
                           critic_struct=cs, w_critic=w, action_init=a0)
    eng.run()
    got = eng.results()
    s = oracle.make_sys(name, p["pars"], p["bnds"])
    c = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], critic_struct=cs, R1=p["R1_diag"],
                         observation_target=p["target"])
    ref = oracle.closed_loop(c, s, x0, cand, a0, p["dt"], 0.0, t1, p["dt"] / 2, w_critic=w)
    assert np.array_equal(got["nsteps"], ref["nsteps"])
    assert np.array_equal(got["nsamples"], ref["nsamples"])
    assert np.array_equal(got["nfev"], ref["nfev"])
    assert np.array_equal(got["t"], ref["t"])
    assert mixed_err(got["y"], ref["y"], floor=1e-2) <= 1e-9
    assert rel_err(got["accum"], ref["accum"]) <= 1e-9
    assert np.all(got["status"] == _C.FINISHED)


def test_lane_permutation_and_batch_size_invariance(rb):
    """No cross-environment arithmetic: permuting lanes permutes results bit-exactly, and a lane run
    alone (E=1) gives the same bits as inside a big batch (the basis of shard-count invariance)."""
    _, _C, ops = rb
    from rcognita_b200.engine import ClosedLoopEngine
    name, N, C_, E = "3wrobotNI", 6, 256, 4096 + 17
    p = PRESET[name]
    x0 = random_states(name, E, 41)
    cand = random_cands(name, (C_,), N, 42)
    kw = dict(ctrl_bnds=p["bnds"], mode="MPC", Nactor=N, dt=0.01, t1=0.15, R1=p["R1_diag"])
    a = ClosedLoopEngine(name, x0, cand, **kw); a.run(); ra = a.results()
    perm = np.random.default_rng(43).permutation(E)
    b = ClosedLoopEngine(name, x0[perm], cand, **kw); b.run(); rb_ = b.results()
    for k in ("y", "t", "accum", "nsteps", "nsamples", "argmin", "Jmin"):
        assert np.array_equal(ra[k][perm], rb_[k], equal_nan=True), k
    one = ClosedLoopEngine(name, x0[1234:1235], cand, **kw); one.run(); r1 = one.results()
    for k in ("y", "t", "accum", "nsteps", "nsamples", "argmin", "Jmin"):
        assert np.array_equal(ra[k][1234:1235], r1[k], equal_nan=True), k


def test_fp32_twin_tolerance(rb):
    """_f32 kernels: costs within 2e-5 relative of the fp64 oracle; arg-min equal or a near-tie."""
    _, _C, ops = rb
    name, N, C_, E = "2tank", 8, 256, 64
    n, m = DIMS[name]
    p = PRESET[name]
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    kw = dict(mode="SQL", Nactor=N, pred_step_size=0.2, critic_struct="quad-nomix", R1=p["R1_diag"],
              observation_target=p["target"])
    obj = _C.make_objective(n, m, **kw)
    s = oracle.make_sys(name, p["pars"], p["bnds"]); c = oracle.make_ctrl(n, m, **kw)
    obs = random_states(name, E, 51); cand = random_cands(name, (C_,), N, 52); w = np.array([11.0, 11.0, 1.0])
    f32 = torch.float32
    J, am, Jmin = ops.actor_cost(sysd, obj, soa(obs).to(f32), soa(obs).to(f32), dev(cand.T.copy()).to(f32), False, C_,
                                 w_critic=dev(w).to(f32))
    J, am = J.cpu().numpy().astype(np.float64), am.cpu().numpy()
    for e in range(E):
        Jr, ar = oracle.actor_cost_table(c, s, cand, obs[e], obs[e], w)
        assert rel_err(J[e], Jr) <= 2e-5
        assert am[e] == ar or abs(Jr[am[e]] - Jr[ar]) <= 2e-5 * abs(Jr[ar])


@pytest.mark.parametrize("name,mode,N,C_", [("3wrobotNI", "MPC", 6, 64), ("3wrobotNI", "MPC", 7, 64), ("3wrobot", "RQL", 20, 32),
                                           ("2tank", "SQL", 13, 96)])
def test_fp32_twin_per_env_candidates(rb, name, mode, N, C_):
    """_f32 twins of the TMA-staged kernels (specialised and runtime horizons, lean and general objectives) with
    per-environment candidates: costs within 5e-5 relative of the fp64 kernels on the same inputs."""
    _, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    E = 517
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], critic_struct="quad-nomix",
                            R1=p["R1_diag"], observation_target=p["target"])
    obs = soa(random_states(name, E, 61))
    g = torch.Generator(device="cuda").manual_seed(62)
    b = torch.tensor(p["bnds"], device="cuda", dtype=torch.float64)
    cand = torch.empty((N * m, E * C_), device="cuda", dtype=torch.float64)
    for k in range(N * m):
        j = k % m
        cand[k] = b[j, 0] + (b[j, 1] - b[j, 0]) * torch.rand((E * C_,), device="cuda", dtype=torch.float64, generator=g)
    w = dev(np.random.default_rng(63).uniform(0.5, 2, size=n + m))
    J64, am64, _ = ops.actor_cost(sysd, obj, obs, obs, cand, True, C_, w_critic=w)
    f32 = torch.float32
    J32, am32, _ = ops.actor_cost(sysd, obj, obs.to(f32), obs.to(f32), cand.to(f32), True, C_, w_critic=w.to(f32))
    rel = ((J32.double() - J64).abs() / J64.abs().clamp_min(1e-12)).max().item()
    assert rel <= 5e-5, rel
    assert (am32 == am64).double().mean().item() >= 0.98


def test_errors_are_loud(rb):
    rcognita_b200, _C, ops = rb
    p = PRESET["3wrobotNI"]
    sysd = _C.make_system("3wrobotNI", p["pars"], p["bnds"])
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.state_dyn(sysd, torch.zeros((3, 4), dtype=torch.float64), torch.zeros((2, 4), dtype=torch.float64))
    with pytest.raises(ValueError):
        _C.make_system("pendulum")
    sol = _C.make_solver(1.0, -1.0)
    y, f, t, h, status, nfev, action = _sim_state(ops, _C, sysd, [X0["3wrobotNI"]])
    with pytest.raises(RuntimeError, match="max_step"):
        ops.rk45_step(sysd, sol, y, f, t, h, status, action)


@pytest.mark.parametrize("name,mode,kw", [
    ("3wrobotNI", "MPC", dict(actor="candidates")), ("3wrobotNI", "MPC", dict(actor="opt", opt_start="argmin", log_every=2, log_capacity=16)),
    ("2tank", "SQL", dict(critic_fit=True)),
])
def test_checkpoint_resume_is_bit_identical(rb, name, mode, kw):
    """state_dict() after 7 intervals, 9 more; a fresh engine that loads the snapshot and runs the same 9 intervals ends
    in exactly the same state (lane state, counters, critic weights, trajectory ring)."""
    from rcognita_b200.engine import ClosedLoopEngine
    p = PRESET[name]
    n, m = DIMS[name]
    E, N = 257, 6
    x0 = random_states(name, E, 31)
    cand = random_cands(name, (64,), N, 32)
    args = dict(pars=p["pars"], ctrl_bnds=p["bnds"], mode=mode, Nactor=N, dt=p["dt"], pred_step_size=p["dt"] * p["psm"], t1=1e6,
                R1=p["R1_diag"], observation_target=p["target"], **kw)
    a = ClosedLoopEngine(name, x0, cand, **args)
    for _ in range(7):
        a.run_interval()
    sd = a.state_dict()
    for _ in range(9):
        a.run_interval()
    b = ClosedLoopEngine(name, x0, cand, **args)
    b.load_state_dict(sd)
    for _ in range(9):
        b.run_interval()
    assert torch.equal(a._blob, b._blob) and a.intervals == b.intervals == 16
    if kw.get("critic_fit"):
        assert torch.equal(a.w, b.w) and torch.equal(a.obs_buf, b.obs_buf) and torch.equal(a.nfits, b.nfits)
    if a.log is not None:
        assert torch.equal(a.log.count, b.log.count)
        assert np.array_equal(a.trajectory(5), b.trajectory(5))


def test_empty_batch_is_a_no_op(rb):
    """E = 0 (an empty shard): every entry point returns success without touching a pointer."""
    _, _C, ops = rb
    name = "3wrobotNI"
    p = PRESET[name]
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    sol = _C.make_solver(1.0, 0.005)
    obj = _C.make_objective(3, 2, mode="RQL", Nactor=6, pred_step_size=0.01, R1=p["R1_diag"], buffer_size=10)
    z = lambda *shape, dt=torch.float64: torch.zeros(shape, dtype=dt, device="cuda")       # noqa: E731
    i32 = torch.int32
    y, a = z(3, 0), z(2, 0)
    ops.rhs(sysd, y, a)
    ops.state_dyn(sysd, y, a)
    ops.rk45_step(sysd, sol, y, z(3, 0), z(0), z(0), z(0, dt=i32), a)
    ops.rk45_advance(sysd, sol, obj, y, z(3, 0), z(0), z(0), z(0, dt=i32), a, z(0), 0.01, 4, state_sys=z(3, 0), accum=z(0),
                     sample_flag=z(0, dt=i32))
    J, am, jm = ops.actor_cost(sysd, obj, y, y, z(12, 8), False, 8, w_critic=z(5))
    assert J.shape == (0, 8) and am.numel() == 0
    ops.stage_obj(obj, 3, 2, y, a)
    ops.critic(obj, 3, 2, y, a, z(5))
    ops.critic_cost(obj, 3, 2, z(10, 3, 0), z(10, 2, 0), z(5, 0, 1), z(5, 0))
    ops.critic_fit(obj, 3, 2, z(10, 3, 0), z(10, 2, 0), z(5, 0), 0.0, 1e3, z(5, 0), w_init=z(5))
    ops.ctrl_sample(z(0), z(0), 0.01)
    ops.push_buffers(3, 2, z(10, 3, 0), z(10, 2, 0), y, a)
    ops.nominal_ni(sysd, y, 0.5, a)
    # round-2 entry points: disturbance lanes, fp32 critic cost
    distd = _C.make_disturb([[1, 1], [0, 0], [0.3, 0.3]], seed=1)
    yf = z(5, 0)
    ops.rhs_disturbed(sysd, distd, yf, a)
    assert ops.disturb_normals(distd, 0, 3, device="cuda").shape == (2, 0)
    ops.rk45_step_disturbed(sysd, distd, sol, yf, z(5, 0), z(0), z(0), z(0, dt=i32), a, z(0, dt=i32))
    ops.rk45_advance_disturbed(sysd, distd, sol, obj, yf, z(5, 0), z(0), z(0), z(0, dt=i32), a, z(0), 0.01, 4, z(0, dt=i32),
                               state_sys=z(3, 0), accum=z(0), sample_flag=z(0, dt=i32))
    f32 = torch.float32
    assert ops.critic_cost(obj, 3, 2, z(10, 3, 0, dt=f32), z(10, 2, 0, dt=f32), z(5, 0, 1, dt=f32), z(5, 0, dt=f32)).shape == (0, 1)


@pytest.mark.parametrize("name,cs", [("2tank", "quad-nomix"), ("3wrobot", "quadratic"), ("3wrobotNI", "quad-lin")])
def test_critic_cost_f32_twin(rb, name, cs):
    """rcg_critic_cost_f32 against the fp64 kernel on the same buffers: stated fp32 tolerance 2e-5 of the cost scale
    (the sum of the squared TD terms' magnitudes: J_c is a difference of O(Q) terms, squared)."""
    from rcognita_b200 import _C, ops
    n, m = DIMS[name]
    p = PRESET[name]
    E, W = 300, 3
    rng = np.random.default_rng(9)
    obj = _C.make_objective(n, m, mode="RQL", Nactor=4, gamma=0.95, Ncritic=4, buffer_size=10, critic_struct=cs,
                            R1=p["R1_diag"], observation_target=p["target"])
    dimc = _C.dim_critic(cs, n, m)
    b = np.array(p["bnds"], dtype=float)
    ob = torch.as_tensor(rng.normal(size=(10, n, E)), device="cuda")
    ab = torch.as_tensor(rng.uniform(b[:, 0], b[:, 1], size=(E, 10, m)).transpose(1, 2, 0).copy(), device="cuda")
    w = torch.as_tensor(rng.uniform(0, 2, size=(dimc, E, W)), device="cuda")
    wp = torch.as_tensor(rng.uniform(0, 2, size=(dimc, E)), device="cuda")
    J64 = ops.critic_cost(obj, n, m, ob, ab, w, wp)
    J32 = ops.critic_cost(obj, n, m, ob.float(), ab.float(), w.float(), wp.float())
    assert J32.dtype == torch.float32 and J32.shape == J64.shape
    # scale of the terms that are differenced: |Q| of the largest row, squared
    q = torch.stack([ops.critic(obj, n, m, ob[k].contiguous(), ab[k].contiguous(), wp, w_per_env=True).abs() for k in range(4)]).max(0).values
    scale = torch.maximum(J64.abs(), (q * q)[:, None])
    assert float(((J32.double() - J64).abs() / scale).max()) <= 2e-5


@pytest.mark.parametrize("graph,actor", [(False, "candidates"), (True, "candidates"), (True, "opt")])
def test_host_staged_loop_equals_resident_loop(rb, graph, actor):
    """engine.HostStagedLoop (the caller owns state and time in pinned host memory and reads back state, time, action,
    accumulated objective, status and arg-min; solver internals stay on the device; several environment blocks on their
    own streams, optionally one CUDA graph per block) takes exactly the same steps as the device-resident
    ClosedLoopEngine: identical state, clocks, counters and picks on every lane after every control interval."""
    from rcognita_b200.engine import ClosedLoopEngine, HostStagedLoop
    name, N, C_, E = "3wrobotNI", 6, 64, 1000
    p = PRESET[name]
    x0 = random_states(name, E, 71)
    cand = random_cands(name, (E, C_), N, 72)
    kw = dict(ctrl_bnds=p["bnds"], mode="MPC", Nactor=N, dt=p["dt"], t1=0.4, R1=p["R1_diag"], actor=actor)
    ref = ClosedLoopEngine(name, x0, cand, **kw)
    loop = HostStagedLoop(name, x0, cand, nchunks=3, graph=graph, **kw)
    if graph:
        ref.run_interval()                      # capture() itself advances the loop by one warm-up interval
    assert set(ClosedLoopEngine.HOST_OUT_FIELDS) <= set(loop.hosts[0]) and "f" not in loop.hosts[0] and "h_abs" not in loop.hosts[0]
    for k in range(12):
        ref.run_interval()
        h2d, d2h = loop.step()
        assert 0 < h2d < d2h
        assert h2d <= E * (3 * 8 + 8) + 3 * 2 * 256 and d2h <= E * (3 * 8 + 8 + 2 * 8 + 8 + 3 * 4) + 3 * 7 * 256
        for f in ClosedLoopEngine.HOST_OUT_FIELDS:                       # valid on return, no device access
            assert np.array_equal(loop.host_field(f).numpy(), getattr(ref, f).cpu().numpy(), equal_nan=True), (k, f)
        for f in ("f", "state_sys", "h_abs", "ctrl_clock", "nsteps", "nsamples", "nfev", "Jmin"):
            assert np.array_equal(loop.field(f).cpu().numpy(), getattr(ref, f).cpu().numpy(), equal_nan=True), (k, f)


@pytest.mark.parametrize("graph", [False, True])
def test_host_staged_pipelined_run_equals_resident_loop(rb, graph):
    """HostStagedLoop.run: the blocks are pipelined (block b's interval k+1 is enqueued as soon as its interval-k results
    are on the host); the callback sees every block's results after every interval, and they are the resident loop's."""
    from rcognita_b200.engine import ClosedLoopEngine, HostStagedLoop
    name, N, C_, E, K = "3wrobotNI", 6, 32, 640, 9
    p = PRESET[name]
    x0 = random_states(name, E, 73)
    cand = random_cands(name, (E, C_), N, 74)
    kw = dict(ctrl_bnds=p["bnds"], mode="MPC", Nactor=N, dt=p["dt"], t1=0.4, R1=p["R1_diag"])
    ref = ClosedLoopEngine(name, x0, cand, **kw)
    loop = HostStagedLoop(name, x0, cand, nchunks=4, graph=graph, **kw)
    if graph:
        ref.run_interval()
    want = []
    for k in range(K):
        ref.run_interval()
        want.append({f: getattr(ref, f).cpu().numpy().copy() for f in ClosedLoopEngine.HOST_OUT_FIELDS})
    seen = set()

    def on_block(b, k, host):
        a, z = loop.bounds[b], loop.bounds[b + 1]
        for f in ClosedLoopEngine.HOST_OUT_FIELDS:
            assert np.array_equal(host[f].numpy(), want[k][f][..., a:z], equal_nan=True), (b, k, f)
        seen.add((b, k))

    h2d, d2h = loop.run(K, on_block=on_block)
    assert seen == {(b, k) for b in range(loop.nchunks) for k in range(K)}
    assert h2d == K * loop.h2d_bytes and d2h == K * loop.d2h_bytes


def test_last_actor_kernel_names_the_dispatched_variant(rb):
    """rcg_last_actor_kernel: per-environment candidates with a specialised horizon -> the TMA-staged kernel, a runtime
    horizon -> its runtime twin, a shared table -> the direct kernel (what bench.py reports as roofline.kernel)."""
    import rcognita_b200
    from rcognita_b200 import _C, ops
    name, n, m = "3wrobotNI", 3, 2
    p = PRESET[name]
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    E, C_ = 64, 32
    x = torch.as_tensor(random_states(name, E, 5).T.copy(), device="cuda")
    for N, per_env, want in ((6, True, "actor_cost_tma_kernel"), (7, True, "actor_cost_tma_kernel"), (12, True, "actor_cost_tma_rt_kernel"),
                             (6, False, "actor_cost_kernel")):
        obj = _C.make_objective(n, m, mode="MPC", Nactor=N, pred_step_size=0.01, R1=p["R1_diag"])
        cand = torch.as_tensor(random_cands(name, (E * C_,) if per_env else (C_,), N, 6).T.copy(), device="cuda")
        ops.actor_cost(sysd, obj, x, x, cand, per_env, C_, want_J=False)
        assert rcognita_b200.last_actor_kernel() == want


@pytest.mark.parametrize("nchunks,stagger,per_env", [(2, True, True), (3, False, True), (4, True, False)])
def test_pipelined_loop_equals_single_engine(rb, nchunks, stagger, per_env):
    """engine.PipelinedLoop (environment blocks on their own streams, rk45_advance of one block beside the actor launch
    of another) is bit-identical to one ClosedLoopEngine over the whole batch, with and without events around the actor
    launches."""
    from rcognita_b200.engine import ClosedLoopEngine, PipelinedLoop
    name, N, C_, E = "3wrobotNI", 6, 64, 4096
    p = PRESET[name]
    x0 = random_states(name, E, 75)
    cand = random_cands(name, (E, C_) if per_env else (C_,), N, 76)
    kw = dict(ctrl_bnds=p["bnds"], mode="MPC", Nactor=N, dt=p["dt"], t1=0.3, R1=p["R1_diag"])
    ref = ClosedLoopEngine(name, x0, cand, **kw)
    loop = PipelinedLoop(name, x0, cand, nchunks=nchunks, stagger=stagger, **kw)
    assert loop.nchunks == nchunks and loop.bounds[-1] == E and all(b % 1024 == 0 for b in loop.bounds)
    for k in range(40):
        if k == 20:
            loop.actor_events = []
        ref.run_interval()
        loop.step()
    assert len(loop.actor_events) == 20 * nchunks
    torch.cuda.synchronize()
    assert all(a.elapsed_time(b) >= 0 for a, b in loop.actor_events)
    got, want = loop.results(), ref.results()
    assert got["status"].max() >= 1                      # the episode ended inside the test
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
