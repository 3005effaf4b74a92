"""Property tests (hypothesis): invariants of the path that hold for ANY input -- the partitioner, np.argmin semantics,
the adjoint gradient, and (on the GPU) lane-permutation / batch-split invariance of the kernels: no arithmetic ever
crosses environments, so a lane's result may not depend on where it sits in the batch."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from golden_util import DIMS, PRESET

SYSTEMS = ["3wrobotNI", "3wrobot", "2tank"]


@settings(max_examples=200, deadline=None, derandomize=True)
@given(nb=st.integers(1, 4096), world=st.integers(1, 64))
def test_shard_range_partitions_the_batch(nb, world):
    from rcognita_b200 import shard
    E = nb * shard.BLOCK
    cuts = [shard.shard_range(E, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == E
    sizes = []
    for (lo, hi), (lo2, _) in zip(cuts, cuts[1:] + [(E, E)]):
        assert lo <= hi == lo2 and lo % shard.BLOCK == 0
        sizes.append(hi - lo)
    assert max(sizes) - min(sizes) <= shard.BLOCK and sizes == sorted(sizes, reverse=True)


@settings(max_examples=300, deadline=None, derandomize=True)
@given(st.lists(st.one_of(st.floats(-1e6, 1e6), st.just(float("nan")), st.just(float("inf")), st.just(-0.0), st.just(0.0)),
                min_size=1, max_size=70))
def test_oracle_argmin_is_numpy_argmin(vals):
    J = np.array(vals, dtype=np.float64)
    assert oracle.argmin(J) == int(np.argmin(J))


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(name=st.sampled_from(SYSTEMS), mode=st.sampled_from(["MPC", "RQL", "SQL"]),
       cs=st.sampled_from(["quad-lin", "quadratic", "quad-nomix", "quad-mix"]), N=st.integers(1, 9),
       gamma=st.sampled_from([1.0, 0.9]), seed=st.integers(0, 2**31 - 1))
def test_adjoint_gradient_is_the_derivative_of_the_cost(name, mode, cs, N, gamma, seed):
    """orc_actor_grad (the algorithm of the CUDA reverse sweep) against central differences of orc_actor_cost."""
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(seed)
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, mode=mode, Nactor=N, pred_step_size=P["dt"] * P["psm"], gamma=gamma, critic_struct=cs,
                          R1=np.diag(P["R1_diag"]).astype(float) + 0.1 * np.eye(n + m), observation_target=P["target"])
    b = np.array(P["bnds"], dtype=float)
    x = rng.uniform(np.tile(b[:, 0], N), np.tile(b[:, 1], N))
    xs = rng.uniform(-3, 3, size=n)
    ob = xs + 0.01 * rng.normal(size=n)
    w = rng.uniform(-1, 2, size=oracle.dim_critic(cs, n, m))
    J, g = oracle.actor_grad(ct, s, x, ob, xs, w)
    f = lambda z: oracle.actor_cost(ct, s, z, ob, xs, w)          # noqa: E731
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
    assert np.max(np.abs(g - gf)) <= 1e-6 * max(np.max(np.abs(gf)), 1e-3)


# ------------------------------------------------------------------------------------------- GPU
torch = pytest.importorskip("torch")


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(name=st.sampled_from(SYSTEMS), mode=st.sampled_from(["MPC", "RQL", "SQL"]), N=st.sampled_from([1, 3, 6, 7, 10, 13]),
       E=st.integers(1, 700), C=st.sampled_from([1, 2, 8, 32, 33, 96, 256]), per_env=st.booleans(), seed=st.integers(0, 2**31 - 1))
def test_actor_cost_lane_permutation_and_split_invariance(name, mode, N, E, C, per_env, seed):
    """rcg_actor_cost on a batch, on a random permutation of its environments, and on its two halves: the same costs,
    arg-min and J_min per environment.  A permutation keeps the launch geometry, so it must be bit-identical; halving the
    batch can select another kernel variant (the TMA-staged kernels need an even number of candidate columns), and two
    variants are separate compilations whose FMA contraction may differ: costs then agree to 1e-12 and the arg-min up to
    such ties."""
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from rcognita_b200 import _C, ops
    n, m = DIMS[name]
    P = PRESET[name]
    rng = np.random.default_rng(seed)
    sysd = _C.make_system(name, P["pars"], P["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=P["dt"] * P["psm"], critic_struct="quadratic",
                            R1=P["R1_diag"], observation_target=P["target"])
    b = np.array(P["bnds"], dtype=float)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda", dtype=torch.float64)   # noqa: E731
    x = rng.uniform(-4, 4, size=(E, n))
    W = rng.uniform(0, 2, size=(E, _C.dim_critic("quadratic", n, m)))
    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
    cand = rng.uniform(lo, hi, size=(E, C, N * m)) if per_env else rng.uniform(lo, hi, size=(C, N * m))

    def run(idx):
        xs = dev(x[idx].T)
        cd = dev(cand[idx].transpose(2, 0, 1).reshape(N * m, -1)) if per_env else dev(cand.T)
        J, am, jm = ops.actor_cost(sysd, obj, xs, xs, cd, per_env, C, w_critic=dev(W[idx].T), w_per_env=True)
        return J.cpu().numpy(), am.cpu().numpy(), jm.cpu().numpy()

    full = run(np.arange(E))
    perm = rng.permutation(E)
    pj, pa, pm = run(perm)
    assert np.array_equal(pj, full[0][perm]) and np.array_equal(pa, full[1][perm]) and np.array_equal(pm, full[2][perm], equal_nan=True)
    if E >= 2:
        h = E // 2
        a, b2 = run(np.arange(h)), run(np.arange(h, E))
        J2, am2 = np.concatenate([a[0], b2[0]]), np.concatenate([a[1], b2[1]])
        assert np.max(np.abs(J2 - full[0]) / np.maximum(np.abs(full[0]), 1e-300)) <= 1e-12
        for e in np.flatnonzero(am2 != full[1]):
            assert abs(full[0][e, am2[e]] - full[0][e, full[1][e]]) <= 1e-12 * abs(full[0][e, full[1][e]])
