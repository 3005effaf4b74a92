"""Parity at BASELINE.json's FULL sizes (configs 2, 3, 4) through size-independent properties plus oracle spot checks
on randomly sampled environments -- the oracle cannot run 1.7e7 evaluations, but it can run the sampled lanes of the
very launch that processed all of them.

  config 2: 3wrobot_NI MPC Nactor=6, 65,536 envs x 256 per-environment candidates (the TMA-staged kernel)
  config 3: 3wrobot RQL 'quadratic' critic Nactor=10, 1,048,576 envs x 256 shared candidates, per-env weights
  config 4: 2tank SQL 'quad-nomix' Nactor=8, 262,144 envs, critic buffer fitting
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from golden_util import DIMS, PRESET  # noqa: E402

F64 = torch.float64


@pytest.fixture(scope="module")
def rb():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import rcognita_b200
    from rcognita_b200 import _C, ops
    torch.cuda.set_device(0)
    return rcognita_b200, _C, ops


def _states(name, E, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    box = {"3wrobotNI": ([-10, -10, -np.pi], [10, 10, np.pi]), "3wrobot": ([-10, -10, -np.pi, -1, -1], [10, 10, np.pi, 1, 1]),
           "2tank": ([-2, -2], [2, 2])}[name]
    lo, hi = (torch.tensor(v, device="cuda", dtype=F64) for v in box)
    return lo[:, None] + (hi - lo)[:, None] * torch.rand((lo.numel(), E), device="cuda", dtype=F64, generator=g)


def _cands(name, N, ncol, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    b = torch.tensor(PRESET[name]["bnds"], device="cuda", dtype=F64)
    m = b.shape[0]
    cand = torch.empty((N * m, ncol), device="cuda", dtype=F64)
    for k in range(N * m):
        j = k % m
        cand[k] = b[j, 0] + (b[j, 1] - b[j, 0]) * torch.rand((ncol,), device="cuda", dtype=F64, generator=g)
    return cand


def test_config2_full_size_actor_cost(rb):
    """65,536 x 256 per-environment candidates in ONE launch: (i) 48 sampled environments against the oracle (all
    256 costs to 1e-9, arg-min exact); (ii) reversing every environment's candidate order mirrors the arg-min and
    leaves J_min bit-identical; (iii) the TMA-staged kernel and the direct-load kernel agree bit for bit."""
    _, _C, ops = rb
    name, N, E, C = "3wrobotNI", 6, 65536, 256
    n, m = DIMS[name]
    sysd = _C.make_system(name, [], PRESET[name]["bnds"])
    kw = dict(mode="MPC", Nactor=N, pred_step_size=0.01, R1=np.diag(PRESET[name]["R1_diag"]).astype(float))
    obj = _C.make_objective(n, m, **kw)
    x = _states(name, E, 0)
    cand = _cands(name, N, E * C, 1)                                  # [12, E*C], 1.6 GB
    J, am, jm = ops.actor_cost(sysd, obj, x, x, cand, True, C)
    s = oracle.make_sys(name, [], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, **kw)
    rng = np.random.default_rng(0)
    xs = x.cpu().numpy()
    for e in rng.choice(E, size=48, replace=False):
        tab = cand[:, e * C:(e + 1) * C].t().cpu().numpy()
        Jo, ao = oracle.actor_cost_table(ct, s, tab, xs[:, e], xs[:, e])
        assert np.max(np.abs(J[e].cpu().numpy() - Jo) / np.abs(Jo)) <= 1e-9
        assert am[e].item() == ao and abs(jm[e].item() - Jo[ao]) <= 1e-9 * abs(Jo[ao])
    # (ii) mirror the candidate order inside every environment
    rev = cand.view(N * m, E, C).flip(2).reshape(N * m, E * C).contiguous()
    _, am_r, jm_r = ops.actor_cost(sysd, obj, x, x, rev, True, C, want_J=False)
    assert torch.equal(jm_r, jm)
    Js = J.sort(dim=1).values
    unique_min = Js[:, 0] < Js[:, 1]
    assert unique_min.float().mean().item() > 0.99
    assert torch.equal(am_r[unique_min], (C - 1 - am)[unique_min])
    del rev
    # (iii) direct-load kernel on the same inputs
    os.environ["RCG_ACTOR_NO_TMA"] = "1"
    try:
        J2, am2, jm2 = ops.actor_cost(sysd, obj, x, x, cand, True, C)
    finally:
        del os.environ["RCG_ACTOR_NO_TMA"]
    assert torch.equal(J2, J) and torch.equal(am2, am) and torch.equal(jm2, jm)


def test_config2_full_size_closed_loop_sampled_lanes(rb):
    """The fused loop on 65,536 environments x 256 per-environment candidates for 6 control intervals: 24 sampled
    lanes equal the oracle's closed loop (same step / sample counts and solver times, states to 1e-9)."""
    from rcognita_b200.engine import ClosedLoopEngine
    name, N, E, C = "3wrobotNI", 6, 65536, 256
    n, m = DIMS[name]
    x = _states(name, E, 3)
    cand = _cands(name, N, E * C, 4).view(N * m, E, C).permute(1, 2, 0)            # [E, C, L] view
    t1 = 0.055
    eng = ClosedLoopEngine(name, x.t(), cand, ctrl_bnds=PRESET[name]["bnds"], mode="MPC", Nactor=N, dt=0.01, t1=t1,
                           R1=PRESET[name]["R1_diag"])
    eng.run()
    got = eng.results()
    assert np.all(got["status"] == 1)
    s = oracle.make_sys(name, [], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, mode="MPC", Nactor=N, pred_step_size=0.01, R1=PRESET[name]["R1_diag"])
    lanes = np.random.default_rng(1).choice(E, size=24, replace=False)
    x0 = x.t()[lanes].cpu().numpy()
    cd = cand[lanes].cpu().numpy()
    ref = oracle.closed_loop(ct, s, x0, cd, [-2.5, -0.5], 0.01, 0.0, t1, 0.005)
    assert np.array_equal(got["nsteps"][lanes], ref["nsteps"]) and np.array_equal(got["nsamples"][lanes], ref["nsamples"])
    assert np.array_equal(got["t"][lanes], ref["t"])
    assert np.max(np.abs(got["y"][lanes] - ref["y"]) / np.maximum(np.abs(ref["y"]), 1e-2)) <= 1e-9
    assert np.max(np.abs(got["accum"][lanes] - ref["accum"]) / np.abs(ref["accum"])) <= 1e-9


def test_config3_full_size_actor_cost_per_env_weights(rb):
    """1,048,576 environments x 256 shared candidates, RQL with the 28-weight 'quadratic' critic and per-environment
    weights: 32 sampled environments against the oracle; a lane permutation of the whole batch permutes the results."""
    _, _C, ops = rb
    name, N, E, C = "3wrobot", 10, 1 << 20, 256
    n, m = DIMS[name]
    P = PRESET[name]
    sysd = _C.make_system(name, P["pars"], P["bnds"])
    kw = dict(mode="RQL", Nactor=N, pred_step_size=0.02, critic_struct="quadratic", R1=np.diag(P["R1_diag"]).astype(float))
    obj = _C.make_objective(n, m, **kw)
    x = _states(name, E, 5)
    cand = _cands(name, N, C, 6)
    g = torch.Generator(device="cuda").manual_seed(7)
    W = torch.rand((28, E), device="cuda", dtype=F64, generator=g) * 2
    _, am, jm = ops.actor_cost(sysd, obj, x, x, cand, False, C, w_critic=W, w_per_env=True, want_J=False)
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, **kw)
    tab = cand.t().cpu().numpy()
    for e in np.random.default_rng(2).choice(E, size=32, replace=False):
        xe = x[:, e].cpu().numpy()
        Jo, ao = oracle.actor_cost_table(ct, s, tab, xe, xe, W[:, e].cpu().numpy())
        assert am[e].item() == ao and abs(jm[e].item() - Jo[ao]) <= 1e-9 * abs(Jo[ao])
    perm = torch.randperm(E, device="cuda", generator=g)
    _, am_p, jm_p = ops.actor_cost(sysd, obj, x[:, perm].contiguous(), x[:, perm].contiguous(), cand, False, C,
                                   w_critic=W[:, perm].contiguous(), w_per_env=True, want_J=False)
    assert torch.equal(am_p, am[perm]) and torch.equal(jm_p, jm[perm])


def test_config3_full_size_rk45_sampled_lanes(rb):
    """1,048,576 Sys3WRobot lanes, 12 sim_steps with per-lane actions (some out of bounds): 64 sampled lanes equal
    the oracle step for step (t, h_abs bit-exact; y, f to 1e-12); times increase strictly on every lane."""
    _, _C, ops = rb
    name, E = "3wrobot", 1 << 20
    n, m = DIMS[name]
    P = PRESET[name]
    sysd = _C.make_system(name, P["pars"], P["bnds"])
    sol = _C.make_solver(1.0, 0.005)
    y = _states(name, E, 8)
    g = torch.Generator(device="cuda").manual_seed(9)
    b = torch.tensor(P["bnds"], device="cuda", dtype=F64)
    act = (b[:, :1] + (b[:, 1:] - b[:, :1]) * torch.rand((m, E), device="cuda", dtype=F64, generator=g)) * 1.3
    lanes = np.random.default_rng(3).choice(E, size=64, replace=False)
    y0, a0 = y[:, lanes].t().cpu().numpy(), act[:, lanes].t().cpu().numpy()
    zero = torch.zeros_like(act)
    f = ops.rhs(sysd, y, zero)                                        # RK45.__init__ with System.action = 0
    t = torch.zeros(E, device="cuda", dtype=F64)
    h = torch.full((E,), 1e-6, device="cuda", dtype=F64)
    st = torch.zeros(E, device="cuda", dtype=torch.int32)
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    refs = []
    for k in range(64):
        r = oracle.RK45(s, y0[k], 0.0, 1.0, 0.005)
        r.receive_action(a0[k])
        refs.append(r)
    t_prev = t.clone()
    for step in range(12):
        ops.rk45_step(sysd, sol, y, f, t, h, st, act)
        assert bool((t > t_prev).all())
        t_prev = t.clone()
        for r in refs:
            r.step()
    tt, hh, yy, ff = t[lanes].cpu().numpy(), h[lanes].cpu().numpy(), y[:, lanes].t().cpu().numpy(), f[:, lanes].t().cpu().numpy()
    for k, r in enumerate(refs):
        assert tt[k] == r.t and hh[k] == r.h_abs
        assert np.max(np.abs(yy[k] - r.y) / np.maximum(np.abs(r.y), 1e-3)) <= 1e-12
        assert np.max(np.abs(ff[k] - r.f) / np.maximum(np.abs(r.f), 1e-3)) <= 1e-12
    assert float(act.abs().max()) <= 300.0                             # clipped in place by closed_loop_rhs


def test_config4_full_size_critic_fit_properties(rb):
    """262,144 critic fits (2tank, 'quad-nomix', 3 weights, Ncritic=4): the fitted cost equals _critic_cost at the
    fitted weights, is <= the cost at w_critic_init on EVERY lane, the weights stay inside [Wmin, Wmax], refitting
    from the result does not raise the cost (idempotence), and 32 sampled lanes match the oracle's _critic_cost."""
    _, _C, ops = rb
    name, E = "2tank", 262144
    n, m = DIMS[name]
    P = PRESET[name]
    kw = dict(mode="SQL", Nactor=8, Ncritic=4, buffer_size=10, critic_struct="quad-nomix", R1=np.diag(P["R1_diag"]).astype(float),
              observation_target=P["target"])
    obj = _C.make_objective(n, m, **kw)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = _states(name, E, 10)
    v = 0.05 * torch.randn((n, E), device="cuda", dtype=F64, generator=g)
    obs_buf = torch.stack([x + k * v for k in range(10)]).contiguous()
    act_buf = torch.rand((10, m, E), device="cuda", dtype=F64, generator=g)
    w_prev = torch.ones((3, E), device="cuda", dtype=F64)
    w_init = torch.ones((3,), device="cuda", dtype=F64)
    w = torch.empty((3, E), device="cuda", dtype=F64)
    Jc = torch.empty((E,), device="cuda", dtype=F64)
    ops.critic_fit(obj, n, m, obs_buf, act_buf, w_prev, 0.0, 1e3, w, w_init=w_init, Jc_out=Jc)
    J0 = ops.critic_cost(obj, n, m, obs_buf, act_buf, w_init[:, None, None].expand(3, E, 1).contiguous(), w_prev)[:, 0]
    J1 = ops.critic_cost(obj, n, m, obs_buf, act_buf, w[:, :, None].contiguous(), w_prev)[:, 0]
    assert bool((w >= 0.0).all()) and bool((w <= 1e3).all())
    assert bool((Jc <= J0 * (1 + 1e-12)).all())
    assert bool(((Jc - J1).abs() <= 1e-6 * torch.maximum(J1, 1e-9 * J0) + 1e-18).all())
    w2 = w.clone()
    Jc2 = torch.empty_like(Jc)
    ops.critic_fit(obj, n, m, obs_buf, act_buf, w_prev, 0.0, 1e3, w2, w_init=None, Jc_out=Jc2)      # start from w
    assert bool((Jc2 <= Jc * (1 + 1e-12) + 1e-300).all())
    ct = oracle.make_ctrl(n, m, mode="SQL", Nactor=8, Ncritic=4, buffer_size=10, critic_struct="quad-nomix",
                          R1=np.diag(P["R1_diag"]).astype(float), observation_target=P["target"])
    for e in np.random.default_rng(4).choice(E, size=32, replace=False):
        Jo = oracle.critic_cost(ct, n, m, obs_buf[:, :, e].cpu().numpy(), act_buf[:, :, e].cpu().numpy(),
                                w[:, e].cpu().numpy(), np.ones(3))
        assert abs(J1[e].item() - Jo) <= 1e-9 * max(abs(Jo), 1e-12)
