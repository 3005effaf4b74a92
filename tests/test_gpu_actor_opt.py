"""GPU parity tests of the batched actor optimiser (rcg_actor_grad / rcg_actor_opt / rcg_gather_sqn), the stand-in
for CtrlOptPred._actor_optimizer (rcognita/controllers.py:1330-1427).

Bars: the adjoint gradient equals the oracle's analytic gradient to 1e-9 (relative to the gradient's largest entry)
and the cost equals orc_actor_cost to 1e-9; the minimum found from the reference's own start point
(action_sqn_init) is <= the live reference's SLSQP minimum on every committed golden problem (1e-7 relative slack
= SLSQP's own tolerance); the CUDA iteration and the oracle's restatement of it end at the same cost; the result is
feasible, never above the start cost, and invariant to batching.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from golden_util import DIMS, PRESET, load  # noqa: E402


@pytest.fixture(scope="module")
def rb():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import rcognita_b200
    from rcognita_b200 import _C, ops
    torch.cuda.set_device(0)
    return rcognita_b200, _C, ops


@pytest.fixture(autouse=True)
def _default_optimiser_variant():
    """Every test starts and ends on the default rcg_actor_opt variant (four lanes per problem where instantiated)."""
    import rcognita_b200
    rcognita_b200.actor_opt_lanes(0)
    yield
    rcognita_b200.actor_opt_lanes(0)


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=dtype)


def _descr(_C, c):
    name = c["system"]
    n, m = DIMS[name]
    sysd = _C.make_system(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    obj = _C.make_objective(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                            critic_struct=c["critic_struct"], R1=np.array(c["R1"]), observation_target=c["target"])
    s = oracle.make_sys(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                          critic_struct=c["critic_struct"], R1=np.array(c["R1"]), observation_target=c["target"])
    return n, m, sysd, obj, s, ct


GOLD = load("actor_opt.json")


@pytest.mark.parametrize("lanes", [4, 1])
@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_actor_opt_reaches_the_reference_slsqp_minimum(rb, idx, lanes):
    """From action_sqn_init (the reference's start point) the batched optimiser must end at a cost <= SLSQP's -- both
    kernels behind rcg_actor_opt (four lanes per problem = the default, and one lane per problem)."""
    rcg, _C, ops = rb
    rcg.actor_opt_lanes(lanes)
    c = GOLD[idx]
    n, m, sysd, obj, s, ct = _descr(_C, c)
    L = c["N"] * m
    # lane 0: the golden problem; lanes 1..: the same problem from perturbed starts (exercise divergence)
    E = 5
    rng = np.random.default_rng(idx)
    b = np.array(PRESET[c["system"]]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], c["N"]), np.tile(b[:, 1], c["N"])
    starts = np.stack([np.array(c["x_init"])] + [rng.uniform(lo, hi) for _ in range(E - 1)], axis=1)      # [L, E]
    state = dev(np.tile(np.array(c["state_sys"])[:, None], (1, E)))
    obs = dev(np.tile(np.array(c["obs"])[:, None], (1, E)))
    w = dev(c["w"]) if c["mode"] != "MPC" else None
    sqn = dev(starts)
    J0, _ = ops.actor_grad(sysd, obj, state, obs, sqn, w_critic=w)
    J, iters, nfev = ops.actor_opt(sysd, obj, state, obs, sqn, w_critic=w, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
    specialised = 3 <= c["N"] <= 10 and np.count_nonzero(np.array(c["R1"]) - np.diag(np.diagonal(np.array(c["R1"])))) == 0
    specialised = specialised and c["system"] != "2tank"     # Sys2Tank stays on the one-lane kernel
    assert rcg.last_actor_opt_kernel() == ("actor_opt_quad_kernel" if lanes == 4 and specialised else "actor_opt_kernel")
    J, x = J.cpu().numpy(), sqn.cpu().numpy()
    assert abs(J0[0].item() - c["J_init"]) <= 1e-9 * max(abs(c["J_init"]), 1.0)
    # feasible, monotone
    assert np.all(x >= lo[:, None]) and np.all(x <= hi[:, None])
    assert np.all(J <= J0.cpu().numpy() + 1e-12 * np.abs(J0.cpu().numpy()))
    # the returned cost is _actor_cost of the returned sequence (oracle = the reference's arithmetic)
    for e in range(E):
        Jo = oracle.actor_cost(ct, s, x[:, e], c["obs"], c["state_sys"], c["w"] if c["mode"] != "MPC" else None)
        assert abs(J[e] - Jo) <= 1e-9 * max(abs(Jo), 1e-3), (e, J[e], Jo)
    # the bar: the live reference's SLSQP minimum from the same start
    tol = 1e-7 * max(abs(c["J_ref"]), 1.0)
    assert J[0] <= c["J_ref"] + tol, f"J={J[0]!r} > SLSQP {c['J_ref']!r} (iters {iters[0].item()}, nfev {nfev[0].item()})"
    # CUDA iteration vs the oracle's restatement of the same algorithm.  The two evaluate the heading trigonometry
    # differently (rotation recurrence vs sincos per stage, ~1e-15 apart), so in flat ill-conditioned valleys the
    # iterates separate after some dozens of steps and stop at different points of the valley floor (or at the
    # iteration cap): there the costs only agree to a few per cent of max(|J|, 1), and what is asserted is that BOTH
    # meet the reference's bar; where the problem is well conditioned (<= 30 iterations) they agree to 1e-9.
    _, Jorc, it_o, _ = oracle.actor_opt(ct, s, c["x_init"], c["obs"], c["state_sys"],
                                        c["w"] if c["mode"] != "MPC" else None, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
    assert Jorc <= c["J_ref"] + tol
    well_conditioned = max(iters[0].item(), it_o) <= 30
    assert abs(J[0] - Jorc) <= (1e-9 if well_conditioned else 5e-2) * max(abs(Jorc), 1.0), (J[0], Jorc, iters[0].item(), it_o)


@pytest.mark.parametrize("name,mode,cs,N,dense", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, False), ("3wrobotNI", "RQL", "quad-lin", 6, False),
    ("3wrobotNI", "SQL", "quad-mix", 7, False), ("3wrobotNI", "MPC", "quad-nomix", 4, True),
    ("3wrobot", "RQL", "quadratic", 10, False), ("3wrobot", "SQL", "quad-lin", 5, False),
    ("3wrobot", "MPC", "quad-nomix", 12, False), ("3wrobot", "RQL", "quad-mix", 3, True),
    ("2tank", "SQL", "quad-nomix", 8, False), ("2tank", "RQL", "quadratic", 9, False),
    ("2tank", "MPC", "quad-nomix", 8, True), ("2tank", "SQL", "quad-lin", 64, False),
])
def test_actor_grad_vs_oracle(rb, name, mode, cs, N, dense):
    """Adjoint gradient == the oracle's analytic gradient (itself pinned to finite differences of the reference's
    cost in tests/test_oracle_golden.py), specialised and generic kernels, S > 1, per-environment weights."""
    _, _C, ops = rb
    n, m = DIMS[name]
    p = n + m
    rng = np.random.default_rng(abs(hash((name, mode, cs, N))) % 2**32)
    E, S = 37, 4
    R1 = np.diag(PRESET[name]["R1_diag"]).astype(float)
    if dense:
        A = rng.normal(size=(p, p)); R1 = A @ A.T / p
    tgt = PRESET[name]["target"]
    sysd = _C.make_system(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    kw = dict(mode=mode, Nactor=N, pred_step_size=PRESET[name]["dt"] * PRESET[name]["psm"], gamma=0.95,
              critic_struct=cs, R1=R1, observation_target=tgt)
    obj = _C.make_objective(n, m, **kw)
    s = oracle.make_sys(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, **kw)
    dimc = oracle.dim_critic(cs, n, m)
    b = np.array(PRESET[name]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    x_sys = rng.uniform(-3, 3, size=(E, n))
    ob = x_sys + rng.normal(size=(E, n)) * 0.01
    W = rng.uniform(-1, 2, size=(E, dimc))
    U = rng.uniform(lo, hi, size=(E, S, N * m))
    sqn = dev(U.transpose(2, 0, 1).reshape(N * m, E * S))
    J, g = ops.actor_grad(sysd, obj, dev(x_sys.T.copy()), dev(ob.T.copy()), sqn, S=S,
                          w_critic=dev(W.T.copy()) if mode != "MPC" else None, w_per_env=True)
    J, g = J.cpu().numpy().reshape(E, S), g.cpu().numpy().reshape(N * m, E, S)
    for e in range(0, E, 3):
        for k in range(S):
            Jo, go = oracle.actor_grad(ct, s, U[e, k], ob[e], x_sys[e], W[e] if mode != "MPC" else None)
            assert abs(J[e, k] - Jo) <= 1e-9 * max(abs(Jo), 1e-3)
            assert np.max(np.abs(g[:, e, k] - go)) <= 1e-9 * max(np.max(np.abs(go)), 1e-3)


def test_actor_opt_multi_start_mask_and_action_handover(rb):
    """S starts per environment: best_out / Jmin_out = np.argmin over the starts' final costs, action_out = first
    action of the best minimiser, accum += stage_obj * sampling_time; masked environments are untouched."""
    _, _C, ops = rb
    name, N = "3wrobotNI", 6
    n, m = DIMS[name]
    E, S = 70, 8
    rng = np.random.default_rng(3)
    sysd = _C.make_system(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    kw = dict(mode="MPC", Nactor=N, pred_step_size=0.01, gamma=1.0, R1=np.diag(PRESET[name]["R1_diag"]).astype(float))
    obj = _C.make_objective(n, m, **kw)
    ct = oracle.make_ctrl(n, m, **kw)
    b = np.array(PRESET[name]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    x_sys = np.stack([rng.uniform(-1, 1, E), rng.uniform(-1, 1, E), rng.uniform(-np.pi, np.pi, E)], 1) * 0.3
    U = rng.uniform(lo, hi, size=(E, S, N * m))
    sqn = dev(U.transpose(2, 0, 1).reshape(N * m, E * S))
    before = sqn.clone()
    mask = torch.ones(E, dtype=torch.int32, device="cuda")
    mask[5::7] = 0
    best = torch.full((E,), -7, dtype=torch.int32, device="cuda")
    Jmin = torch.full((E,), -7.0, dtype=torch.float64, device="cuda")
    act = torch.full((m, E), -7.0, dtype=torch.float64, device="cuda")
    accum = torch.zeros(E, dtype=torch.float64, device="cuda")
    st = dev(x_sys.T.copy())
    J, iters, nfev = ops.actor_opt(sysd, obj, st, st, sqn, S=S, mask=mask, max_iter=60, best_out=best, Jmin_out=Jmin,
                                   action_out=act, accum=accum, sampling_time=0.01)
    J = J.cpu().numpy().reshape(E, S)
    x = sqn.cpu().numpy().reshape(N * m, E, S)
    mk = mask.cpu().numpy().astype(bool)
    assert np.array_equal(sqn.reshape(N * m, E, S)[:, ~mask.bool()].cpu().numpy(),
                          before.reshape(N * m, E, S)[:, ~mask.bool()].cpu().numpy())
    assert np.all(best.cpu().numpy()[~mk] == -7) and np.all(act.cpu().numpy()[:, ~mk] == -7.0)
    for e in np.nonzero(mk)[0]:
        bi = int(np.argmin(J[e]))
        assert best[e].item() == bi
        assert Jmin[e].item() == J[e, bi]
        assert np.array_equal(act[:, e].cpu().numpy(), x[:m, e, bi])
        so = oracle.stage_obj(ct, n, m, x_sys[e], x[:m, e, bi])
        assert abs(accum[e].item() - so * 0.01) <= 1e-12 * max(abs(so), 1.0)
    assert np.all(accum.cpu().numpy()[~mk] == 0.0)


def test_actor_opt_batch_invariance_and_gather(rb):
    """The minimiser of an environment does not depend on its neighbours (batch of 1 == lane of a big batch, bit
    for bit), and rcg_gather_sqn picks the arg-min candidate as the start point."""
    _, _C, ops = rb
    name, N, C_ = "3wrobot", 10, 64
    n, m = DIMS[name]
    E = 300
    rng = np.random.default_rng(11)
    sysd = _C.make_system(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    obj = _C.make_objective(n, m, mode="RQL", Nactor=N, pred_step_size=0.02, gamma=1.0, critic_struct="quadratic",
                            R1=np.diag(PRESET[name]["R1_diag"]).astype(float))
    b = np.array(PRESET[name]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    x_sys = rng.uniform(-2, 2, size=(E, n))
    st = dev(x_sys.T.copy())
    w = dev(rng.uniform(0, 2, size=oracle.dim_critic("quadratic", n, m)))
    cand = dev(rng.uniform(lo, hi, size=(C_, N * m)).T.copy())                       # shared table [L, C]
    _, am, _ = ops.actor_cost(sysd, obj, st, st, cand, False, C_, w_critic=w, want_J=False)
    sqn = torch.empty((N * m, E), dtype=torch.float64, device="cuda")
    ops.gather_sqn(cand, False, C_, am, sqn)
    assert torch.equal(sqn, cand[:, am.long()])
    ref = sqn.clone()
    J, iters, _ = ops.actor_opt(sysd, obj, st, st, sqn, w_critic=w, max_iter=40)
    for e in (0, 17, 299):
        one = ref[:, e:e + 1].clone()
        J1, it1, _ = ops.actor_opt(sysd, obj, st[:, e:e + 1].contiguous(), st[:, e:e + 1].contiguous(), one, w_critic=w,
                                   max_iter=40)
        assert torch.equal(one[:, 0], sqn[:, e]) and J1[0].item() == J[e].item() and it1[0].item() == iters[e].item()


def test_actor_opt_errors_are_loud(rb):
    _, _C, ops = rb
    name = "2tank"
    n, m = DIMS[name]
    sysd = _C.make_system(name, PRESET[name]["pars"], PRESET[name]["bnds"])
    obj = _C.make_objective(n, m, mode="SQL", Nactor=8, R1=np.eye(3))
    st = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    sqn = torch.zeros((8, 4 * 3), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="power of two"):
        ops.actor_opt(sysd, obj, st, st, sqn, S=3, w_critic=torch.ones(3, dtype=torch.float64, device="cuda"))
    sqn = torch.zeros((8, 4), dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="w_critic"):
        ops.actor_opt(sysd, obj, st, st, sqn)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.actor_opt(sysd, obj, st.cpu(), st, sqn, w_critic=torch.ones(3, dtype=torch.float64, device="cuda"))


def _oracle_loop_with_optimizer(s, ct, x0, action_init, dt, t1, N, m, max_iter, w=None):
    """The headless main loop (presets/main_3wrobot_NI.py:415-440) for one environment with the oracle's pieces and
    the restated minimiser as _actor_optimizer, started from action_sqn_init at every sample like the reference."""
    r = oracle.RK45(s, x0, 0.0, t1, dt / 2)
    sqn_init = np.tile(action_init, N)
    action, state_sys, clock, accum, steps, samples = np.array(action_init, dtype=float), np.array(x0, dtype=float), 0.0, 0.0, 0, 0
    n = len(x0)
    while True:
        r.step()
        steps += 1
        t, obs = r.t, r.y
        if t - clock >= dt:
            clock = t
            x, _, _, _ = oracle.actor_opt(ct, s, sqn_init, obs, state_sys, w, max_iter=max_iter, pg_tol=1e-7, f_tol=1e-12)
            action = x[:m].copy()
            samples += 1
        r.receive_action(action)
        state_sys = r.y
        accum += oracle.stage_obj(ct, n, m, obs, action) * dt
        if t >= t1:
            return r.y, accum, steps, samples


@pytest.mark.parametrize("lanes", [4, 1])
def test_closed_loop_with_optimizer_vs_oracle(rb, lanes):
    """Engine with actor='opt', opt_start='init' (the reference's protocol: every sample minimises from
    action_sqn_init) against the same loop assembled from the oracle: identical step and sample counts,
    trajectories and accumulated objective to 1e-6 relative (north star: 1e-6 on trajectories).

    The last environment starts next to the goal, where the minimiser of a sample is not a continuous function of the
    state (two valley points 3 % apart in cost trade places under a 1e-12 change of the observation: measured with BOTH
    kernels on identical inputs, which then agree to the last digit).  The one-lane kernel happens to stay on the
    oracle's branch for the whole episode and is held to 1e-6 there as before; the four-lane kernel, whose inner
    products round differently in the 13th digit, leaves it at the 11th sample, so for it that environment is checked
    sample by sample instead: every solve of the episode is repeated by the oracle on the engine's own inputs."""
    rcg, _C, ops = rb
    from rcognita_b200.engine import ClosedLoopEngine
    rcg.actor_opt_lanes(lanes)
    name, N, dt, t1 = "3wrobotNI", 6, 0.01, 0.4
    n, m = DIMS[name]
    rng = np.random.default_rng(21)
    E = 6
    x0 = np.stack([rng.uniform(-6, 6, E), rng.uniform(-6, 6, E), rng.uniform(-np.pi, np.pi, E)], 1)
    x0[-1] = [0.02, -0.01, 0.03]                                   # near the goal: interior minimisers
    R1 = PRESET[name]["R1_diag"]
    eng = ClosedLoopEngine(name, x0, None, ctrl_bnds=PRESET[name]["bnds"], mode="MPC", Nactor=N, dt=dt, t1=t1, R1=R1,
                           actor="opt", opt_start="init", opt_iters=300)
    solves = []
    real = ops.actor_opt

    def recording(sysd, obj, state_sys, obs, sqn, **kw):
        rec = {"state_sys": state_sys.clone(), "obs": obs.clone(), "mask": kw["mask"].clone()}
        out = real(sysd, obj, state_sys, obs, sqn, **kw)
        rec.update(action=kw["action_out"].clone(), Jmin=kw["Jmin_out"].clone())
        solves.append(rec)
        return out
    ops.actor_opt = recording
    try:
        eng.run()
    finally:
        ops.actor_opt = real
    assert rcg.last_actor_opt_kernel() == ("actor_opt_quad_kernel" if lanes == 4 else "actor_opt_kernel")
    got = eng.results()
    s = oracle.make_sys(name, [], PRESET[name]["bnds"])
    ct = oracle.make_ctrl(n, m, mode="MPC", Nactor=N, pred_step_size=dt, R1=R1)
    for e in range(E):
        y, accum, steps, samples = _oracle_loop_with_optimizer(s, ct, x0[e], [-2.5, -0.5], dt, t1, N, m, 300)
        assert got["nsteps"][e] == steps and got["nsamples"][e] == samples
        if lanes == 4 and e == E - 1:
            continue
        assert np.max(np.abs(got["y"][e] - y) / np.maximum(np.abs(y), 1e-2)) <= 1e-6, (e, got["y"][e], y)
        assert abs(got["accum"][e] - accum) <= 1e-6 * abs(accum)
    if lanes == 4:
        e, checked = E - 1, 0
        for rec in solves:
            if rec["mask"][e].item() == 0:
                continue
            xo, Jo, _, _ = oracle.actor_opt(ct, s, np.tile([-2.5, -0.5], N), rec["obs"][:, e].cpu().numpy(),
                                            rec["state_sys"][:, e].cpu().numpy(), None, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
            assert abs(rec["Jmin"][e].item() - Jo) <= 1e-9 * max(abs(Jo), 1e-3), (checked, rec["Jmin"][e].item(), Jo)
            assert np.max(np.abs(rec["action"][:, e].cpu().numpy() - xo[:m])) <= 1e-5, (checked, rec["action"][:, e], xo[:m])
            checked += 1
        assert checked == got["nsamples"][e]


@pytest.mark.parametrize("name,mode,cs,N,scale", [
    ("3wrobotNI", "MPC", "quad-nomix", 6, 1.0), ("3wrobotNI", "MPC", "quad-nomix", 6, 0.01), ("3wrobotNI", "SQL", "quad-lin", 5, 0.3),
    ("3wrobotNI", "RQL", "quad-mix", 9, 0.3), ("3wrobot", "RQL", "quadratic", 10, 1.0), ("3wrobot", "MPC", "quad-nomix", 7, 0.1),
    ("3wrobot", "SQL", "quad-nomix", 4, 1.0), ("3wrobot", "SQL", "quadratic", 8, 0.3), ("3wrobotNI", "RQL", "quadratic", 3, 1.0),
    ("3wrobotNI", "MPC", "quad-nomix", 10, 0.2)])
def test_four_lane_kernel_equals_one_lane_kernel(rb, name, mode, cs, N, scale):
    """The two kernels behind rcg_actor_opt run the same iteration (the parallel line search accepts the point sequential
    halving would have accepted; only the summation order of the inner products differs): on identical inputs -- distinct
    observation and predictor state, per-environment weights, a mask, far from and next to the goal -- costs agree to 1e-9
    wherever the iteration is well conditioned, within the golden test's valley tolerance elsewhere; skipped environments
    keep their start point in both."""
    rcg, _C, ops = rb
    n, m = DIMS[name]
    p = PRESET[name]
    E = 777
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{name}/{mode}/{cs}/{N}".encode()))
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=0.02, gamma=0.97, critic_struct=cs, R1=np.diag(p["R1_diag"]).astype(float))
    box = {"3wrobotNI": [10, 10, np.pi], "3wrobot": [10, 10, np.pi, 1, 1]}[name]
    x = rng.uniform(-1, 1, size=(E, n)) * np.array(box) * scale
    st = dev(x.T.copy())
    ob = dev((x + 1e-3 * rng.standard_normal(x.shape)).T.copy())
    dimc = oracle.dim_critic(cs, n, m)
    w = None if mode == "MPC" else dev(rng.uniform(0.1, 2.0, size=(dimc, E)))
    mask = dev((rng.uniform(size=E) > 0.2).astype(np.int32), dtype=torch.int32)
    b = np.array(p["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    start = dev(rng.uniform(lo, hi, size=(E, N * m)).T.copy())
    res = {}
    for lanes in (4, 1):
        rcg.actor_opt_lanes(lanes)
        sqn = start.clone()
        J, it, nf = ops.actor_opt(sysd, obj, st, ob, sqn, w_critic=w, w_per_env=w is not None, mask=mask, max_iter=200,
                                  pg_tol=1e-7, f_tol=1e-12)
        assert rcg.last_actor_opt_kernel() == ("actor_opt_quad_kernel" if lanes == 4 else "actor_opt_kernel")
        res[lanes] = (J.cpu().numpy(), it.cpu().numpy(), nf.cpu().numpy(), sqn.cpu().numpy())
    mk = mask.cpu().numpy() != 0
    (J4, it4, nf4, x4), (J1, it1, nf1, x1) = res[4], res[1]
    assert np.array_equal(x4[:, ~mk], start.cpu().numpy()[:, ~mk]) and np.array_equal(x1[:, ~mk], start.cpu().numpy()[:, ~mk])
    rel = np.abs(J4[mk] - J1[mk]) / np.maximum(np.abs(J1[mk]), 1.0)
    dx = np.max(np.abs(x4[:, mk] - x1[:, mk]) / (hi - lo)[:, None], axis=0)
    easy = np.maximum(it4[mk], it1[mk]) <= 30                  # "well conditioned" as in the golden test above
    stats = dict(frac_same=float((rel <= 1e-9).mean()), worst_rel=float(rel.max()), frac_easy=float(easy.mean()),
                 rel_easy=float(rel[easy].max()), dx_easy=float(dx[easy].max()), mean4=float(J4[mk].mean()), mean1=float(J1[mk].mean()))
    # measured on a B200: all ten shapes agree to <= 3e-12 on every problem that stops within 30 iterations; where random
    # starts and random critic weights drive some problems to the iteration cap (3wrobotNI RQL quad-mix N=9, 3wrobot RQL
    # quadratic N=10), 7-11 % of the problems end at valley points <= 6e-3 apart in cost, with equal batch means to 1e-5
    assert rel[easy].max() <= 1e-9 and dx[easy].max() <= 1e-3, stats
    assert rel.max() <= 5e-2 and (rel <= 1e-9).mean() >= 0.85, stats
    assert abs(stats["mean4"] - stats["mean1"]) <= 1e-4 * max(abs(stats["mean1"]), 1.0), stats


def test_config1_episode_with_optimizer_matches_the_reference_slsqp_controller(rb):
    """BASELINE config 1 (presets/main_3wrobot_NI.py, MPC, Nactor=6, dt=0.01, t1=10 from [5, 5, -3pi/4]): the live
    reference with its SLSQP controller accumulates 71.3188 over 2004 solver steps
    (tests/golden/config1_slsqp_episode.npz).  The batched optimiser as the controller must do as well: within 1e-3
    relative from the reference's own start point, and no worse when warm-started from the arg-min candidate;
    enumerate-and-argmin alone (256 random candidates) cannot (76.4)."""
    from rcognita_b200.engine import ClosedLoopEngine
    g = np.load(__import__("os").path.join(__import__("golden_util").GOLDEN, "config1_slsqp_episode.npz"))
    ref_accum, ref_steps = float(g["rows"][-1][7]), g["rows"].shape[0]
    name, N = "3wrobotNI", 6
    b = np.array(PRESET[name]["bnds"], dtype=float)
    cand = np.random.default_rng(1).uniform(np.tile(b[:, 0], N), np.tile(b[:, 1], N), size=(256, 2 * N))
    res = {}
    for tag, kw in (("init", dict(actor="opt", opt_start="init")), ("argmin", dict(actor="opt", opt_start="argmin")),
                    ("candidates", dict(actor="candidates"))):
        eng = ClosedLoopEngine(name, np.array([[5, 5, -3 * np.pi / 4]]), cand, ctrl_bnds=PRESET[name]["bnds"], mode="MPC",
                               Nactor=N, dt=0.01, t1=10.0, R1=PRESET[name]["R1_diag"], **kw)
        eng.run()
        r = eng.results()
        res[tag] = (float(r["accum"][0]), int(r["nsteps"][0]), r["y"][0])
    assert abs(res["init"][0] - ref_accum) <= 1e-3 * ref_accum, res
    assert res["argmin"][0] <= ref_accum * (1 + 1e-4), res
    assert abs(res["init"][1] - ref_steps) <= 0.01 * ref_steps and abs(res["argmin"][1] - ref_steps) <= 0.01 * ref_steps
    assert np.max(np.abs(res["argmin"][2])) <= 0.02 and np.max(np.abs(res["init"][2])) <= 0.02      # parked at the origin
    assert res["candidates"][0] > ref_accum * 1.03


def test_nominal_ni_controller_golden(rb):
    """CtrlNominal3WRobotNI (controllers.py:1758-1956) against actions of the live reference (both gains, the
    xNI[0] = xNI[1] = 0 branch, clipped results) and against the oracle on a large random batch; the class keeps
    the reference's clock / zero-order-hold behaviour."""
    _, _C, ops = rb
    from rcognita_b200 import controllers
    g = load("nominal.json")["cases"]
    bn = PRESET["3wrobotNI"]["bnds"]
    sysd = _C.make_system("3wrobotNI", [], bn)
    for gain in (0.5, 10.0):
        cs = [c for c in g if c["gain"] == gain]
        obs = dev(np.array([c["obs"] for c in cs]).T.copy())
        act = torch.zeros((2, len(cs)), dtype=torch.float64, device="cuda")
        ops.nominal_ni(sysd, obs, gain, act)
        ref = np.array([c["action"] for c in cs]).T
        assert np.max(np.abs(act.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-6)) <= 1e-12
    rng = np.random.default_rng(8)
    E = 20000
    ob = rng.uniform([-10, -10, -np.pi], [10, 10, np.pi], size=(E, 3))
    ob[::7] *= 1e-3
    s = oracle.make_sys("3wrobotNI", [], bn)
    act = torch.zeros((2, E), dtype=torch.float64, device="cuda")
    mask = torch.ones(E, dtype=torch.int32, device="cuda")
    mask[::5] = 0
    ops.nominal_ni(sysd, dev(ob.T.copy()), 0.5, act, mask=mask)
    got = act.cpu().numpy()
    assert np.all(got[:, ::5] == 0.0)
    for e in list(range(1, 400)) + list(range(7, E, 997)):
        if e % 5 == 0:
            continue
        ref = oracle.nominal_ni(0.5, s, ob[e])
        assert np.max(np.abs(got[:, e] - ref) / np.maximum(np.abs(ref), 1e-6)) <= 1e-11, (e, got[:, e], ref)
    # class-level: clock and hold (reference: compute_action :1907-1927)
    nom = controllers.CtrlNominal3WRobotNI(ctrl_gain=0.5, ctrl_bnds=np.array(bn, dtype=float), t0=0, sampling_time=0.01)
    a0 = nom.compute_action(0.005, ob[1])
    assert np.array_equal(a0, np.zeros(2))                          # clock has not fired: action_curr = zeros
    a1 = nom.compute_action(0.0101, ob[1])
    assert np.max(np.abs(a1 - oracle.nominal_ni(0.5, s, ob[1]))) <= 1e-11 and nom.ctrl_clock == 0.0101
    a2 = nom.compute_action(0.015, ob[2])
    assert np.array_equal(a2, a1)                                   # held
    nom.reset(0)
    assert np.array_equal(nom.action_curr, np.zeros(2))


def test_fused_engine_nominal_episode_golden(rb):
    """The fused engine (rk45_advance + nominal kernel per control interval) on the preset-default nominal episode
    of the live reference: same step count and solver times, final state and accumulated objective to 1e-6."""
    from rcognita_b200.engine import ClosedLoopEngine
    g = load("nominal.json")["episode"]
    ref = np.array(g["rows"])
    eng = ClosedLoopEngine("3wrobotNI", np.array([g["x0"], g["x0"]]), None, ctrl_bnds=PRESET["3wrobotNI"]["bnds"], mode="MPC",
                           Nactor=3, dt=0.01, t1=g["t1"], R1=PRESET["3wrobotNI"]["R1_diag"], actor="nominal", ctrl_gain=g["gain"])
    eng.run()
    r = eng.results()
    assert r["nsteps"][0] == ref.shape[0] and r["t"][0] == ref[-1, 0] and r["nfev"][0] == g["nfev"]
    assert np.max(np.abs(r["y"][0] - ref[-1, 1:4])) <= 1e-6 * np.max(np.abs(ref[-1, 1:4]))
    assert abs(r["accum"][0] - ref[-1, 6]) <= 1e-6 * ref[-1, 6]
    assert np.array_equal(r["y"][0], r["y"][1])


@pytest.mark.parametrize("key", ["NI_MPC_N6_x1", "3wrobot_RQL_N10", "2tank_SQL_N8"])
def test_device_trajectory_ring_golden(rb, key):
    """log_every=1: the device-side ring of environment 0 holds one row per solver step, equal to the rows the live
    reference's loop produced (t bit-exact; state, action, accum to 1e-9), in the reference's column order; the
    stage_obj column equals the oracle's stage_obj of that row; other environments log independently."""
    from rcognita_b200.engine import ClosedLoopEngine
    g = load("closed_loop.json")[key]
    name = g["system"]
    n, m = DIMS[name]
    P = PRESET[name]
    ref = np.array(g["rows"])                                   # t, state[n], action[m], accum, sampled
    x0 = np.array([g["x0"], g["x0"], np.array(g["x0"]) * 0.9])
    kw = dict(pars=P["pars"], ctrl_bnds=P["bnds"], mode=g["mode"], Nactor=g["Nactor"], dt=P["dt"],
              pred_step_size=P["dt"] * P["psm"], t1=g["t1"], R1=P["R1_diag"], gamma=g["gamma"], critic_struct=g["critic_struct"],
              observation_target=P["target"], action_init=g["action_init"], w_critic=g["w_fixed"])
    eng = ClosedLoopEngine(name, x0, np.array(g["cand"]), log_every=1, log_capacity=ref.shape[0] + 8, **kw)
    eng.run()
    tr = eng.trajectory(0)
    assert tr.shape == (ref.shape[0], 1 + n + 2 + m)
    so_col, acc_col = (1 + n, 2 + n) if name != "2tank" else (4, 5)
    act_cols = list(range(3 + n, 3 + n + m)) if name != "2tank" else [3]
    assert np.array_equal(tr[:, 0], ref[:, 0])
    assert np.max(np.abs(tr[:, 1:1 + n] - ref[:, 1:1 + n]) / np.maximum(np.abs(ref[:, 1:1 + n]), 1e-2)) <= 1e-9
    assert np.max(np.abs(tr[:, act_cols] - ref[:, 1 + n:1 + n + m])) <= 1e-9 * np.max(np.abs(ref[:, 1 + n:1 + n + m]))
    assert np.max(np.abs(tr[:, acc_col] - ref[:, 1 + n + m]) / np.maximum(np.abs(ref[:, 1 + n + m]), 1e-2)) <= 1e-9
    ct = oracle.make_ctrl(n, m, mode=g["mode"], Nactor=g["Nactor"], R1=P["R1_diag"], observation_target=P["target"])
    for r in tr[:: max(1, len(tr) // 40)]:
        so = oracle.stage_obj(ct, n, m, r[1:1 + n], r[act_cols])
        assert abs(r[so_col] - so) <= 1e-9 * max(abs(so), 1e-6)
    assert np.array_equal(eng.trajectory(1), tr)
    tr2 = eng.trajectory(2)
    assert tr2.shape[0] == int(eng.results()["nsteps"][2]) and not np.array_equal(tr2[-1, 1:1 + n], tr[-1, 1:1 + n])


def test_device_trajectory_ring_decimation_and_wraparound(rb):
    """log_every=3 keeps the steps whose per-lane index is a multiple of 3; a ring shorter than the episode keeps
    the most recent rows in order."""
    from rcognita_b200.engine import ClosedLoopEngine
    g = load("closed_loop.json")["NI_MPC_N6_x1"]
    P = PRESET["3wrobotNI"]
    ref = np.array(g["rows"])
    kw = dict(ctrl_bnds=P["bnds"], mode="MPC", Nactor=6, dt=0.01, t1=g["t1"], R1=P["R1_diag"], action_init=g["action_init"])
    eng = ClosedLoopEngine("3wrobotNI", np.array([g["x0"]] * 2), np.array(g["cand"]), log_every=3, log_capacity=20, **kw)
    eng.run()
    tr = eng.trajectory(1)
    want = ref[2::3][-20:]                                       # steps 3, 6, 9, ... (1-based), the last 20 of them
    assert tr.shape[0] == 20 and np.array_equal(tr[:, 0], want[:, 0])
    assert int(eng.log.count[0].item()) == ref.shape[0] // 3


def test_closed_loop_with_optimizer_vs_oracle_2tank_sql(rb):
    """Same as above for Sys2Tank, SQL with a fixed 'quad-nomix' critic and an observation target (the general,
    non-lean objective; 1-D action box [0, 1]): engine with actor='opt' vs the loop assembled from the oracle."""
    from rcognita_b200.engine import ClosedLoopEngine
    name, N, dt, t1 = "2tank", 8, 0.1, 3.0
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(5)
    E = 5
    x0 = rng.uniform(-2, 2, size=(E, n))
    w = np.array([11.0, 11.0, 1.0])
    kw = dict(mode="SQL", Nactor=N, pred_step_size=dt * P["psm"], critic_struct="quad-nomix", R1=P["R1_diag"],
              observation_target=P["target"])
    eng = ClosedLoopEngine(name, x0, None, pars=P["pars"], ctrl_bnds=P["bnds"], dt=dt, t1=t1, w_critic=w, action_init=[0.5],
                           actor="opt", opt_start="init", opt_iters=300, **kw)
    eng.run()
    got = eng.results()
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, **kw)
    for e in range(E):
        y, accum, steps, samples = _oracle_loop_with_optimizer(s, ct, x0[e], [0.5], dt, t1, N, m, 300, w=w)
        assert got["nsteps"][e] == steps and got["nsamples"][e] == samples
        assert np.max(np.abs(got["y"][e] - y) / np.maximum(np.abs(y), 1e-2)) <= 1e-6, (e, got["y"][e], y)
        assert abs(got["accum"][e] - accum) <= 1e-6 * abs(accum)


def test_closed_loop_with_optimizer_sample_by_sample_3wrobot_rql(rb):
    """Sys3WRobot, RQL with a fixed 'quadratic' critic, Nactor = 10 (BASELINE config 3's shapes; the four-lane kernel): every
    solve of a closed-loop episode with actor='opt' is repeated by the oracle's restatement on the engine's own inputs --
    cost to 1e-8, first action to 1e-4 of the box -- and the engine's bookkeeping around the solves (which environments
    sample, what reaches `action` and `accum`) is checked against the recorded solves."""
    rcg, _C, ops = rb
    from rcognita_b200.engine import ClosedLoopEngine
    name, N, t1 = "3wrobot", 10, 0.25
    n, m = DIMS[name]
    P = PRESET[name]
    dt = P["dt"]
    rng = np.random.default_rng(17)
    E = 4
    x0 = rng.uniform(-1, 1, size=(E, n)) * np.array([3.0, 3.0, 2.0, 0.5, 0.5])
    dimc = oracle.dim_critic("quadratic", n, m)
    Wm = rng.uniform(0.0, 1.0, size=(n + m, n + m))
    Wm = Wm @ Wm.T + np.diag([1.0, 1.0, 1.0, 0.5, 0.5, 1e-3, 1e-2])       # positive definite critic: well-conditioned solves
    w = np.array([Wm[i, j] * (1.0 if i == j else 2.0) for i in range(n + m) for j in range(i, n + m)])
    assert w.size == dimc
    R1 = [1.0, 10.0, 1.0, 0.1, 0.1, 1e-4, 1e-3]
    kw = dict(mode="RQL", Nactor=N, pred_step_size=dt * P["psm"], critic_struct="quadratic", R1=R1)
    a_init = list(np.array(P["bnds"], dtype=float)[:, 0] / 10)
    eng = ClosedLoopEngine(name, x0, None, pars=P["pars"], ctrl_bnds=P["bnds"], dt=dt, t1=t1, w_critic=w, action_init=a_init,
                           actor="opt", opt_start="init", opt_iters=300, **kw)
    solves = []
    real = ops.actor_opt

    def recording(sysd, obj, state_sys, obs, sqn, **k):
        rec = {"state_sys": state_sys.clone(), "obs": obs.clone(), "mask": k["mask"].clone(), "accum0": k["accum"].clone()}
        out = real(sysd, obj, state_sys, obs, sqn, **k)
        rec.update(action=k["action_out"].clone(), Jmin=k["Jmin_out"].clone(), accum1=k["accum"].clone())
        solves.append(rec)
        return out
    ops.actor_opt = recording
    try:
        eng.run()
    finally:
        ops.actor_opt = real
    assert rcg.last_actor_opt_kernel() == "actor_opt_quad_kernel"
    got = eng.results()
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, **kw)
    rngbox = np.array(P["bnds"], dtype=float)
    checked = np.zeros(E, dtype=int)
    for rec in solves:
        for e in range(E):
            if rec["mask"][e].item() == 0:
                assert rec["accum1"][e].item() == rec["accum0"][e].item()
                continue
            ob, st = rec["obs"][:, e].cpu().numpy(), rec["state_sys"][:, e].cpu().numpy()
            xo, Jo, _, _ = oracle.actor_opt(ct, s, np.tile(a_init, N), ob, st, w, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
            assert abs(rec["Jmin"][e].item() - Jo) <= 1e-8 * max(abs(Jo), 1.0), (checked, rec["Jmin"][e].item(), Jo)
            act = rec["action"][:, e].cpu().numpy()
            # (the cost pins the solve; along the flat directions of the double integrator the stop test |P(x - g) - x| <= 1e-7
            #  leaves the first action free to a few 1e-6 of the box: measured 4e-6)
            assert np.max(np.abs(act - xo[:m]) / (rngbox[:, 1] - rngbox[:, 0])) <= 1e-4, (checked, act, xo[:m])
            so = oracle.stage_obj(ct, n, m, ob, act)
            assert abs((rec["accum1"][e].item() - rec["accum0"][e].item()) - so * dt) <= 1e-12 * max(abs(so * dt), 1.0)
            checked[e] += 1
    assert np.array_equal(checked, got["nsamples"]) and checked.min() >= 20


def test_actor_opt_edge_cases(rb):
    """Nactor = 1 (no Euler step at all), one environment with 32 starts, max_iter = 0 (the clipped start point and
    its cost come back), starts outside the box (projected first), and an empty batch."""
    _, _C, ops = rb
    name = "3wrobotNI"
    n, m = DIMS[name]
    bn = PRESET[name]["bnds"]
    sysd = _C.make_system(name, [], bn)
    s = oracle.make_sys(name, [], bn)
    R1 = np.diag([1.0, 10.0, 1.0, 0.5, 0.25])                        # non-zero action weights: interior minimiser
    x = dev(np.array([[1.0], [-2.0], [0.3]]))
    # Nactor = 1: J = stage_obj(obs, a) -> minimiser a = 0
    obj1 = _C.make_objective(n, m, mode="MPC", Nactor=1, pred_step_size=0.01, R1=R1)
    sq = dev(np.array([[7.0], [-3.0]]))
    J, it, nf = ops.actor_opt(sysd, obj1, x, x, sq, max_iter=50)
    assert torch.all(sq.abs() <= 1e-6) and abs(J[0].item() - (1 + 40 + 0.09)) <= 1e-9
    # 32 starts of one environment, some far outside the box; max_iter = 0 returns the projected starts
    N = 6
    kw = dict(mode="MPC", Nactor=N, pred_step_size=0.01, R1=R1)
    obj = _C.make_objective(n, m, **kw)
    ct = oracle.make_ctrl(n, m, **kw)
    rng = np.random.default_rng(0)
    U = rng.uniform(-60, 60, size=(32, N * m))
    lo, hi = np.tile([-25.0, -5.0], N), np.tile([25.0, 5.0], N)
    sq = dev(U.T.copy())
    J0, it0, _ = ops.actor_opt(sysd, obj, x, x, sq, S=32, max_iter=0)
    assert np.array_equal(sq.cpu().numpy(), np.clip(U, lo, hi).T) and int(it0.max()) == 0
    for k in (0, 13, 31):
        Jo = oracle.actor_cost(ct, s, np.clip(U[k], lo, hi), [1.0, -2.0, 0.3], [1.0, -2.0, 0.3])
        assert abs(J0[k].item() - Jo) <= 1e-9 * abs(Jo)
    best = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    Jmin = torch.zeros(1, dtype=torch.float64, device="cuda")
    J1, it1, _ = ops.actor_opt(sysd, obj, x, x, sq, S=32, max_iter=300, best_out=best, Jmin_out=Jmin)
    assert bool((J1 <= J0 + 1e-12).all()) and best[0].item() == int(np.argmin(J1.cpu().numpy())) and Jmin[0].item() == J1.min().item()
    # strictly convex objective: all 32 starts end at the same minimum
    assert float(J1.max() - J1.min()) <= 1e-6 * float(J1.min())
    # empty batch: a no-op
    e0 = torch.zeros((n, 0), dtype=torch.float64, device="cuda")
    ops.actor_opt(sysd, obj, e0, e0, torch.zeros((N * m, 0), dtype=torch.float64, device="cuda"), max_iter=5)


@pytest.mark.timeout(120)
def test_four_lane_kernel_degenerate_inputs(rb):
    """The four-lane kernel on the inputs that must not hang or leak into their neighbours: a non-finite state in one
    environment (its line search gives up, everybody else's result is bit-identical to the batch without it), an all-zero
    mask (nothing is touched), an unbounded box (has_bnds = 0: strictly convex objective, interior minimiser equal to the
    one-lane kernel's), one single problem."""
    rcg, _C, ops = rb
    name, N = "3wrobot", 10
    n, m = DIMS[name]
    p = PRESET[name]
    rng = np.random.default_rng(3)
    E = 97
    R1 = np.diag([1.0, 10.0, 1.0, 0.1, 0.1, 1e-4, 1e-3])
    obj = _C.make_objective(n, m, mode="MPC", Nactor=N, pred_step_size=0.05, R1=R1)
    sysd = _C.make_system(name, p["pars"], p["bnds"])
    x = rng.uniform(-1, 1, size=(E, n)) * np.array([3, 3, np.pi, 1, 1])
    start = np.zeros((N * m, E))
    st = dev(x.T.copy())
    sq = dev(start)
    J, it, nf = ops.actor_opt(sysd, obj, st, st, sq, max_iter=60)
    assert rcg.last_actor_opt_kernel() == "actor_opt_quad_kernel"
    xb = x.copy()
    xb[40, 1] = np.nan
    xb[41, 3] = np.inf
    stb = dev(xb.T.copy())
    sqb = dev(start)
    Jb, itb, nfb = ops.actor_opt(sysd, obj, stb, stb, sqb, max_iter=60)
    keep = np.ones(E, dtype=bool)
    keep[[40, 41]] = False
    assert torch.equal(sqb[:, keep], sq[:, keep]) and torch.equal(Jb[keep], J[keep]) and torch.equal(itb[keep], it[keep])
    assert not np.isfinite(Jb.cpu().numpy()[[40, 41]]).any() and np.array_equal(sqb.cpu().numpy()[:, [40, 41]], start[:, [40, 41]])
    # nothing selected: nothing written
    sq0 = dev(start + 0.25)
    J0 = torch.full((E,), -7.0, dtype=torch.float64, device="cuda")
    ops.actor_opt(sysd, obj, st, st, sq0, mask=torch.zeros(E, dtype=torch.int32, device="cuda"), J_out=J0, max_iter=60)
    assert bool((sq0 == 0.25).all()) and bool((J0 == -7.0).all())
    # unbounded box, one problem
    free = _C.make_system(name, p["pars"], None)
    one = dev(x[:1].T.copy())
    res = {}
    for lanes in (4, 1):
        rcg.actor_opt_lanes(lanes)
        sq1 = dev(start[:, :1])
        J1, it1, _ = ops.actor_opt(free, obj, one, one, sq1, max_iter=300)
        res[lanes] = (J1[0].item(), sq1.cpu().numpy()[:, 0], it1[0].item())
    assert abs(res[4][0] - res[1][0]) <= 1e-9 * max(abs(res[1][0]), 1.0) and res[4][0] < J[0].item() + 1e-9
    assert np.max(np.abs(res[4][1] - res[1][1])) <= 1e-5 * max(np.max(np.abs(res[1][1])), 1.0)


# ---- Gauss-Newton (iLQR) pre-pass: rcg_actor_ilqr --------------------------------------------------------------------
def test_ilqr_presweeps_then_opt_reach_the_slsqp_minimum_without_a_tail(rb):
    """rcg_actor_ilqr followed by rcg_actor_opt on the 72 problems recorded from the live reference: at or below SLSQP's
    minimum, the sweeps never raise the cost, the sweep counts are the CPU checker's (same algorithm: the core is also
    pinned on the host by tests/test_ilqr_core_host.py), and the slowest problem needs <= 48 dependent iterations (the
    quasi-Newton iteration alone runs into its 300-iteration cap on the 3wrobot problems)."""
    _, _C, ops = rb
    worst, same = 0, 0
    for c in GOLD:
        n, m, sysd, obj, s, ct = _descr(_C, c)
        wl = c["w"] if c["mode"] != "MPC" else None
        state, obs = dev(np.array(c["state_sys"])[:, None]), dev(np.array(c["obs"])[:, None])
        w = dev(c["w"]) if c["mode"] != "MPC" else None
        sqn = dev(np.array(c["x_init"])[:, None])
        sweeps = ops.actor_ilqr(sysd, obj, state, obs, sqn, w_critic=w, max_sweeps=25, pg_tol=1e-7)
        x_mid = sqn.cpu().numpy()[:, 0].copy()
        J_mid = oracle.actor_cost(ct, s, x_mid, c["obs"], c["state_sys"], wl)
        assert J_mid <= c["J_init"] + 1e-12 * max(abs(c["J_init"]), 1.0)
        J, iters, _ = ops.actor_opt(sysd, obj, state, obs, sqn, w_critic=w, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
        key = (c["system"], c["mode"], c["critic_struct"], c["N"])
        assert J[0].item() <= c["J_ref"] + 1e-7 * max(abs(c["J_ref"]), 1.0), (key, J[0].item(), c["J_ref"])
        _, _, swo, _ = oracle.actor_opt_hybrid(ct, s, c["x_init"], c["obs"], c["state_sys"], wl)
        same += int(sweeps[0].item() == swo)
        worst = max(worst, sweeps[0].item() + iters[0].item())
    assert worst <= 48, worst
    assert same >= len(GOLD) - 3, same            # device sincos differs from libm in the last bits: decisions may flip rarely


def test_ilqr_batched_layout_masks_and_starts(rb):
    """E environments x S starts with per-environment critic weights and a mask: every column equals the CPU checker's
    sweeps on that problem alone (cost after the sweeps to 1e-9), masked environments are untouched."""
    _, _C, ops = rb
    c = next(g for g in GOLD if g["system"] == "3wrobot" and g["mode"] == "RQL" and g["critic_struct"] == "quad-nomix")
    n, m, sysd, obj, s, ct = _descr(_C, c)
    E, S, L = 67, 4, c["N"] * m
    rng = np.random.default_rng(5)
    b = np.array(PRESET[c["system"]]["bnds"], dtype=float)
    lo, hi = np.tile(b[:, 0], c["N"]), np.tile(b[:, 1], c["N"])
    states = np.array(c["state_sys"])[:, None] + rng.normal(size=(n, E)) * 0.3
    W = np.abs(np.array(c["w"])[:, None] * (1.0 + 0.2 * rng.normal(size=(len(c["w"]), E))))
    starts = rng.uniform(lo[:, None], hi[:, None], size=(L, E * S)) * 0.5
    mask = (rng.uniform(size=E) < 0.8).astype(np.int32)
    sqn = dev(starts)
    sweeps = ops.actor_ilqr(sysd, obj, dev(states), dev(states), sqn, S=S, w_critic=dev(W), w_per_env=True,
                            mask=dev(mask, torch.int32), max_sweeps=25, pg_tol=1e-7).cpu().numpy()
    x = sqn.cpu().numpy()
    agree = 0
    for e in range(E):
        for k in range(S):
            p = e * S + k
            if not mask[e]:
                assert np.array_equal(x[:, p], starts[:, p]) and sweeps[p] == 0
                continue
            xo, Jo, swo, _ = oracle.actor_opt_hybrid(ct, s, starts[:, p], states[:, e], states[:, e], W[:, e], max_sweeps=25,
                                                     max_iter=0, pg_tol=1e-7)
            Jg = oracle.actor_cost(ct, s, x[:, p], states[:, e], states[:, e], W[:, e])
            J0 = oracle.actor_cost(ct, s, starts[:, p], states[:, e], states[:, e], W[:, e])
            assert np.all(x[:, p] >= lo) and np.all(x[:, p] <= hi) and Jg <= J0 + 1e-12 * max(abs(J0), 1.0)
            agree += int(sweeps[p] == swo and abs(Jg - Jo) <= 1e-9 * max(abs(Jo), 1.0))
    assert agree >= 0.97 * int(mask.sum()) * S, (agree, int(mask.sum()) * S)


def test_ilqr_closed_loop_engine_option(rb):
    """ClosedLoopEngine(actor='opt', opt_presweeps=25) on Sys3WRobot MPC (Nactor = 10, zero control weights: the
    ill-conditioned, non-convex case where the quasi-Newton iteration alone runs into its iteration cap at most samples).
    At the first sample both engines solve the same 64 problems: the minimum found with the sweeps is at or below the one
    found without in (nearly) every environment (host study: 64 of 64, up to 21 % lower).  Later samples see different
    states (a different local minimum was applied), so the episode is only required to stay finite and inside the box.
    Plus the argument checks of rcg_actor_ilqr."""
    rcognita_b200, _C, ops = rb
    from rcognita_b200.engine import ClosedLoopEngine
    P = PRESET["3wrobot"]
    rng = np.random.default_rng(2)
    E = 64
    x0 = np.array([5.0, 5.0, 2.4, 0.0, 0.0])[None, :] + rng.normal(size=(E, 5)) * np.array([0.5, 0.5, 0.2, 0.0, 0.0])
    first, out = [], []
    for pre in (0, 25):
        eng = ClosedLoopEngine("3wrobot", x0, None, pars=P["pars"], ctrl_bnds=P["bnds"], mode="MPC", Nactor=10, dt=0.05,
                               pred_step_size=0.1, t1=0.5, R1=np.diag([10.0, 10.0, 1.0, 0.0, 0.0, 0.0, 0.0]), actor="opt",
                               opt_start="init", opt_presweeps=pre)
        eng.run_interval()
        first.append(eng.Jmin.cpu().numpy().copy())
        for _ in range(11):
            eng.run_interval()
        out.append(eng.accum.cpu().numpy().copy())
        act = eng.action.cpu().numpy()
        b = np.array(P["bnds"], dtype=float)
        assert np.all(act >= b[:, :1]) and np.all(act <= b[:, 1:])
    assert np.all(np.isfinite(out[0])) and np.all(np.isfinite(out[1]))
    assert np.sum(first[1] <= first[0] * (1.0 + 1e-5)) >= E - 3, (first[0], first[1])
    assert first[1].mean() <= first[0].mean()
    # argument checks
    n, m, sysd, obj, s, ct = _descr(_C, GOLD[0])
    empty = torch.zeros((n, 0), dtype=torch.float64, device="cuda")
    sq = torch.zeros((GOLD[0]["N"] * m, 0), dtype=torch.float64, device="cuda")
    w = dev(GOLD[0]["w"]) if GOLD[0]["mode"] != "MPC" else None
    assert ops.actor_ilqr(sysd, obj, empty, empty, sq, w_critic=w).numel() == 0
    with pytest.raises(RuntimeError):
        ops.actor_ilqr(sysd, obj, empty, empty, sq, S=3, w_critic=w)
