"""The per-problem algorithm of rcg_actor_ilqr (rcognita_b200/csrc/actor_ilqr_core.cuh, one __host__ __device__ function)
compiled for the HOST by tests/hostcheck/ilqr_host.cu and checked against the CPU checker's restatement
(oracle/rcg_oracle_opt.c: orc_actor_opt_hybrid) on the problems recorded from the live reference
(tests/golden/actor_opt.json: CtrlOptPred._actor_optimizer's SLSQP results).  No GPU: the kernel wrapper (indexing,
strides, masks) is covered by tests/test_gpu_actor_opt.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import DIMS, PRESET, load
from rcognita_b200 import _C

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "ilqr_host.cu")
OUT = os.path.join(HERE, "hostcheck", "_build", "ilqr_host.so")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    core = os.path.join(HERE, "..", "rcognita_b200", "csrc", "actor_ilqr_core.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(core)):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler", "-fPIC", "-o", OUT, SRC],
                       check=True)
    lib = C.CDLL(OUT)
    lib.ilqr_host.argtypes = [C.POINTER(_C.RcgSystem), C.POINTER(_C.RcgObjective), dp, dp, dp, dp, C.c_int, C.c_double]
    lib.ilqr_host.restype = C.c_int
    return lib


def _case(c, stage="quadratic"):
    name = c["system"]
    n, m = DIMS[name]
    P = PRESET[name]
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                          critic_struct=c["critic_struct"], R1=np.array(c["R1"]), observation_target=c["target"])
    rs = _C.make_system(name, P["pars"], P["bnds"])
    ro = _C.make_objective(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                           critic_struct=c["critic_struct"], R1=np.array(c["R1"]), R2=np.array(c["R1"]),
                           stage_obj_struct=stage, observation_target=c["target"] or ())
    w = c["w"] if c["mode"] != "MPC" else None
    return s, ct, rs, ro, w


def _run(host, rs, ro, c, w, max_sweeps=25, pg_tol=1e-7):
    U = np.array(c["x_init"], dtype=np.float64).copy()
    x0 = np.array(c["state_sys"], dtype=np.float64)
    ob = np.array(c["obs"], dtype=np.float64)
    wa = np.array(w if w is not None else [0.0], dtype=np.float64)
    sw = host.ilqr_host(C.byref(rs), C.byref(ro), x0.ctypes.data_as(dp), ob.ctypes.data_as(dp), wa.ctypes.data_as(dp),
                        U.ctypes.data_as(dp), int(max_sweeps), float(pg_tol))
    return U, sw


def test_core_reproduces_the_checker_on_every_recorded_problem(host):
    total, worst = 0, 0
    for c in load("actor_opt.json"):
        s, ct, rs, ro, w = _case(c)
        U, sw = _run(host, rs, ro, c, w)
        b = np.array(PRESET[c["system"]]["bnds"], dtype=float)
        assert np.all(U >= np.tile(b[:, 0], c["N"])) and np.all(U <= np.tile(b[:, 1], c["N"]))
        J_mid = oracle.actor_cost(ct, s, U, c["obs"], c["state_sys"], w)
        assert J_mid <= c["J_init"] + 1e-12 * max(abs(c["J_init"]), 1.0)              # the sweeps never raise the cost
        x, J, it, _ = oracle.actor_opt(ct, s, U, c["obs"], c["state_sys"], w, max_iter=300, pg_tol=1e-7, f_tol=1e-12)
        _, Jo, swo, ito = oracle.actor_opt_hybrid(ct, s, c["x_init"], c["obs"], c["state_sys"], w)
        key = (c["system"], c["mode"], c["critic_struct"], c["N"])
        assert sw == swo, key
        assert abs(J - Jo) <= 1e-9 * max(abs(Jo), 1.0), (key, J, Jo)
        assert J <= c["J_ref"] + 1e-7 * max(abs(c["J_ref"]), 1.0), (key, J, c["J_ref"])   # the live reference's SLSQP minimum
        total += sw
        worst = max(worst, sw + it)
    assert worst <= 40 and total > 0


def test_core_leaves_stationary_starts_and_non_quadratic_costs_alone(host):
    cases = load("actor_opt.json")
    c = cases[0]
    s, ct, rs, ro, w = _case(c)
    x, J, _, _ = oracle.actor_opt(ct, s, c["x_init"], c["obs"], c["state_sys"], w, max_iter=300, pg_tol=1e-9, f_tol=0.0)
    c2 = dict(c, x_init=list(x))
    U, sw = _run(host, rs, ro, c2, w, pg_tol=1e-6)
    assert sw == 0 and np.array_equal(U, x)
    _, _, rs, ro_bi, w = _case(c, stage="biquadratic")
    U, sw = _run(host, rs, ro_bi, c, w)
    assert sw == 0 and np.array_equal(U, np.array(c["x_init"], dtype=np.float64))
    U, sw = _run(host, rs, ro, c, w, max_sweeps=0)
    assert sw == 0 and np.array_equal(U, np.array(c["x_init"], dtype=np.float64))


def test_core_on_random_problems_of_every_system_mode_and_critic(host):
    """240 seeded random problems (3 systems x MPC/RQL/SQL x 4 critic structures, random states, starts, positive or
    indefinite weights): the sweeps never leave the box and never raise the cost; sweep count and cost agree with the
    checker's restatement (libm vs the checker's own sincos can flip a line-search decision: >= 97 % identical)."""
    rng = np.random.default_rng(11)
    structs = ["quad-lin", "quadratic", "quad-nomix", "quad-mix"]
    total = same = 0
    for name in ("3wrobotNI", "3wrobot", "2tank"):
        n, m = DIMS[name]
        P = PRESET[name]
        b = np.array(P["bnds"], dtype=float)
        for mode in ("MPC", "RQL", "SQL"):
            for cs in structs if mode != "MPC" else structs[:1]:
                for rep in range(8 if mode != "MPC" else 16):
                    N = int(rng.integers(1, 9))
                    R1 = np.diag(rng.uniform(0.0, 10.0, size=n + m) * (rng.uniform(size=n + m) < 0.8))
                    kw = dict(mode=mode, Nactor=N, pred_step_size=float(rng.choice([0.01, 0.05, 0.1])),
                              gamma=float(rng.choice([1.0, 0.95])), critic_struct=cs, R1=R1)
                    tgt = list(P["target"]) if name == "2tank" else []
                    s = oracle.make_sys(name, P["pars"], P["bnds"])
                    ct = oracle.make_ctrl(n, m, observation_target=tgt, **kw)
                    rs = _C.make_system(name, P["pars"], P["bnds"])
                    ro = _C.make_objective(n, m, observation_target=tgt or (), **kw)
                    dimc = _C.dim_critic(cs, n, m)
                    w = rng.uniform(0.0, 20.0, size=dimc) if cs in ("quadratic", "quad-nomix") else rng.normal(size=dimc) * 5.0
                    x0 = rng.normal(size=n) * (3.0 if name != "2tank" else 0.5)
                    lo, hi = np.tile(b[:, 0], N), np.tile(b[:, 1], N)
                    start = rng.uniform(lo, hi) * rng.choice([0.0, 0.3, 1.0])
                    c = dict(x_init=list(start), state_sys=list(x0), obs=list(x0))
                    wl = w if mode != "MPC" else None
                    U, sw = _run(host, rs, ro, c, wl)
                    J0 = oracle.actor_cost(ct, s, np.clip(start, lo, hi), x0, x0, wl)
                    J1 = oracle.actor_cost(ct, s, U, x0, x0, wl)
                    assert np.all(U >= lo) and np.all(U <= hi)
                    assert J1 <= J0 + 1e-9 * max(abs(J0), 1.0), (name, mode, cs, N, J0, J1)
                    xo, Jo, swo, _ = oracle.actor_opt_hybrid(ct, s, start, x0, x0, wl, max_sweeps=25, max_iter=0, pg_tol=1e-7)
                    total += 1
                    same += int(sw == swo and abs(J1 - Jo) <= 1e-9 * max(abs(Jo), 1.0))
    assert total == 240 and same >= 0.97 * total, (same, total)
