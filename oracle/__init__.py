"""CPU ORACLE -- test infrastructure, NOT product code.

ctypes front-end of ``oracle/librcg_oracle.so`` (``rcg_oracle.c``): a scalar plain-C
restatement of rcognita's hot path (``System._state_dyn`` / ``closed_loop_rhs``, scipy's
``RK45`` as driven by ``Simulator.sim_step``, ``CtrlOptPred.stage_obj`` / ``_critic`` /
``_critic_cost`` / ``_actor_cost``) with reference file:line citations in the C source.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  ``rcognita_b200`` never does.

Parity status: the reference has no tests or golden vectors of its own, so the oracle is
pinned against outputs of the live reference captured by ``tests/golden/make_golden.py``
(see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librcg_oracle.so")

MAX_N, MAX_M, MAX_P, MAX_W = 5, 2, 7, 35

SYS_IDS = {"3wrobotNI": 0, "3wrobot": 1, "2tank": 2}
SYS_DIMS = {0: (3, 2), 1: (5, 2), 2: (2, 1)}
MODES = {"MPC": 0, "RQL": 1, "SQL": 2}
CRITIC_STRUCTS = {"quad-lin": 0, "quadratic": 1, "quad-nomix": 2, "quad-mix": 3}
STAGE_STRUCTS = {"quadratic": 0, "biquadratic": 1}
STATUS = {0: "running", 1: "finished", 2: "failed"}


class SysT(C.Structure):
    _fields_ = [("sys_id", C.c_int), ("n", C.c_int), ("m", C.c_int), ("has_bnds", C.c_int),
                ("pars", C.c_double * 8), ("lo", C.c_double * MAX_M), ("hi", C.c_double * MAX_M)]


class CtrlT(C.Structure):
    _fields_ = [("mode", C.c_int), ("critic_struct", C.c_int), ("stage_struct", C.c_int),
                ("has_target", C.c_int), ("Nactor", C.c_int), ("Ncritic", C.c_int),
                ("gamma", C.c_double), ("pred_step_size", C.c_double),
                ("R1", C.c_double * (MAX_P * MAX_P)), ("R2", C.c_double * (MAX_P * MAX_P)),
                ("target", C.c_double * MAX_N)]


class Rk45T(C.Structure):
    _fields_ = [("t", C.c_double), ("t_bound", C.c_double), ("h_abs", C.c_double),
                ("max_step", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
                ("y", C.c_double * MAX_N), ("f", C.c_double * MAX_N),
                ("status", C.c_int), ("nfev", C.c_long)]


class DistT(C.Structure):
    _fields_ = [("sigma", C.c_double * 2), ("mu", C.c_double * 2), ("tau", C.c_double * 2), ("seed", C.c_ulonglong)]


class Rk45dT(C.Structure):
    _fields_ = [("t", C.c_double), ("t_bound", C.c_double), ("h_abs", C.c_double), ("max_step", C.c_double),
                ("rtol", C.c_double), ("atol", C.c_double), ("y", C.c_double * 7), ("f", C.c_double * 7),
                ("status", C.c_int), ("nfull", C.c_int), ("nfev", C.c_long), ("env", C.c_ulonglong)]


class EnvT(C.Structure):
    _fields_ = [("r", Rk45T), ("sys_action", C.c_double * MAX_M), ("action_curr", C.c_double * MAX_M),
                ("state_sys", C.c_double * MAX_N), ("ctrl_clock", C.c_double), ("accum", C.c_double),
                ("Jbest", C.c_double), ("steps", C.c_int), ("samples", C.c_int), ("best", C.c_int),
                ("done", C.c_int)]


def build(force: bool = False) -> str:
    """Compile ``librcg_oracle.so`` (gcc; OpenMP if the toolchain has it)."""
    srcs = [os.path.join(_HERE, "rcg_oracle.c"), os.path.join(_HERE, "rcg_oracle_opt.c"),
            os.path.join(_HERE, "rcg_oracle_critic.c"), os.path.join(_HERE, "rcg_oracle_disturb.c")]
    hdr = os.path.join(_HERE, "rcg_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(f) for f in srcs + [hdr])):
        return _LIB_PATH
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    base = [gcc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11", "-shared",
            "-o", _LIB_PATH, *srcs, "-lm"]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base[:1] + extra + base[1:], capture_output=True, text=True)
        if r.returncode == 0:
            return _LIB_PATH
    raise RuntimeError("oracle build failed:\n" + r.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_dim_critic.restype = C.c_int
        L.orc_state_dyn.argtypes = [C.POINTER(SysT), dp, dp, dp]
        L.orc_closed_loop_rhs.argtypes = [C.POINTER(SysT), dp, dp, dp]
        L.orc_rk45_init.argtypes = [C.POINTER(Rk45T), C.POINTER(SysT), dp, dp] + [C.c_double] * 6
        L.orc_rk45_step.argtypes = [C.POINTER(Rk45T), C.POINTER(SysT), dp]
        L.orc_rk45_step.restype = C.c_int
        L.orc_stage_obj.argtypes = [C.POINTER(CtrlT), C.c_int, C.c_int, dp, dp]
        L.orc_stage_obj.restype = C.c_double
        L.orc_critic.argtypes = [C.POINTER(CtrlT), C.c_int, C.c_int, dp, dp, dp]
        L.orc_critic.restype = C.c_double
        L.orc_critic_cost.argtypes = [C.POINTER(CtrlT), C.c_int, C.c_int, dp, dp, dp, dp]
        L.orc_critic_cost.restype = C.c_double
        L.orc_actor_cost.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), dp, dp, dp, dp]
        L.orc_actor_cost.restype = C.c_double
        L.orc_actor_grad.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), dp, dp, dp, dp, dp]
        L.orc_actor_grad.restype = C.c_double
        L.orc_actor_opt.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), dp, dp, dp, dp, C.c_int, C.c_double, C.c_double,
                                    ip, ip]
        L.orc_actor_opt.restype = C.c_double
        L.orc_actor_opt_set_lanes.argtypes = [C.c_int]
        L.orc_actor_opt_set_lanes.restype = None
        L.orc_actor_opt_hybrid.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), dp, dp, dp, dp, C.c_int, C.c_int, C.c_double,
                                           C.c_double, ip, ip]
        L.orc_actor_opt_hybrid.restype = C.c_double
        L.orc_actor_opt_batch.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), C.c_int, dp, dp, dp, C.c_int, C.c_double,
                                          C.c_double, C.c_int, dp]
        L.orc_actor_opt_batch.restype = C.c_longlong
        L.orc_nominal_ni.argtypes = [C.c_double, C.POINTER(SysT), dp, dp]
        L.orc_nominal_ni.restype = None
        L.orc_argmin.argtypes = [dp, C.c_int]
        L.orc_argmin.restype = C.c_int
        L.orc_actor_cost_table.argtypes = [C.POINTER(CtrlT), C.POINTER(SysT), C.c_int, dp, dp, dp, dp, dp, ip]
        L.orc_closed_loop.argtypes = ([C.POINTER(CtrlT), C.POINTER(SysT), C.c_int, dp, C.c_int, dp, C.c_int, dp, dp]
                                      + [C.c_double] * 7 + [C.c_int, C.c_int, dp, dp, dp, ip, ip,
                                                            C.POINTER(C.c_long), dp, C.c_int, ip,
                                                            C.POINTER(C.c_longlong)])
        L.orc_closed_loop.restype = C.c_longlong
        L.orc_env_init.argtypes = [C.POINTER(EnvT), C.POINTER(SysT), C.c_int, dp, dp] + [C.c_double] * 6
        L.orc_env_init.restype = None
        L.orc_env_interval.argtypes = [C.POINTER(EnvT), C.POINTER(CtrlT), C.POINTER(SysT), C.c_int, C.c_int, dp,
                                       C.c_int, dp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_env_interval.restype = C.c_longlong
        L.orc_critic_fit.argtypes = [C.POINTER(CtrlT), C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, dp, C.c_int, ip]
        L.orc_critic_fit.restype = C.c_double
        L.orc_critic_fit_ls.argtypes = [C.POINTER(CtrlT), C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, dp, C.c_int, C.c_int, ip]
        L.orc_critic_fit_ls.restype = C.c_double
        L.orc_closed_loop_critic.argtypes = ([C.POINTER(CtrlT), C.POINTER(SysT), C.c_int, dp, C.c_int, dp, C.c_int, dp, C.c_int]
                                             + [C.c_double] * 10 + [C.c_int, C.c_int, dp, C.c_int, dp, dp, dp, ip, ip, ip, dp, dp,
                                                                    dp, dp, dp, C.c_int, ip])
        L.orc_closed_loop_critic.restype = C.c_longlong
        L.orc_state_dyn_disturbed.argtypes = [C.POINTER(SysT), dp, dp, dp, dp]
        L.orc_state_dyn_disturbed.restype = None
        L.orc_disturb_dyn.argtypes = [C.POINTER(SysT), C.POINTER(DistT), dp, dp, dp]
        L.orc_disturb_dyn.restype = None
        L.orc_log.argtypes = [C.c_double]
        L.orc_log.restype = C.c_double
        L.orc_normal2.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_uint, dp]
        L.orc_normal2.restype = None
        L.orc_closed_loop_rhs_disturbed.argtypes = [C.POINTER(SysT), C.POINTER(DistT), dp, dp, C.c_ulonglong, C.c_uint, dp, dp]
        L.orc_closed_loop_rhs_disturbed.restype = None
        L.orc_rk45d_init.argtypes = [C.POINTER(Rk45dT), C.POINTER(SysT), C.POINTER(DistT), C.c_ulonglong, dp, dp] + [C.c_double] * 6
        L.orc_rk45d_init.restype = None
        L.orc_rk45d_step.argtypes = [C.POINTER(Rk45dT), C.POINTER(SysT), C.POINTER(DistT), dp]
        L.orc_rk45d_step.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_has_openmp.restype = C.c_int
        L.orc_sincos.argtypes = [C.c_double, dp, dp]
        L.orc_sincos.restype = None
        L.orc_pow_m02.argtypes = [C.c_double]
        L.orc_pow_m02.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    """float64 C-contiguous array + its double* (keeps the array alive via the tuple)."""
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr, arr.ctypes.data_as(C.POINTER(C.c_double))


def _null():
    return C.POINTER(C.c_double)()


def make_sys(name, pars=(), ctrl_bnds=None) -> SysT:
    """``name`` is System.name of the reference: '3wrobotNI' | '3wrobot' | '2tank'."""
    s = SysT()
    s.sys_id = SYS_IDS[name]
    s.n, s.m = SYS_DIMS[s.sys_id]
    for i, p in enumerate(pars):
        s.pars[i] = float(p)
    b = np.zeros((0, 2)) if ctrl_bnds is None else np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2)
    s.has_bnds = int(b.size > 0 and bool(b.any()))
    for k in range(b.shape[0]):
        s.lo[k], s.hi[k] = b[k, 0], b[k, 1]
    return s


def make_ctrl(n, m, mode="MPC", Nactor=1, pred_step_size=0.1, gamma=1.0, Ncritic=4, buffer_size=20,
              critic_struct="quad-nomix", stage_obj_struct="quadratic", R1=None, R2=None,
              observation_target=()) -> CtrlT:
    c = CtrlT()
    p = n + m
    c.mode, c.critic_struct, c.stage_struct = MODES[mode], CRITIC_STRUCTS[critic_struct], STAGE_STRUCTS[stage_obj_struct]
    c.Nactor = int(Nactor)
    c.Ncritic = int(min(Ncritic, buffer_size - 1))
    c.gamma, c.pred_step_size = float(gamma), float(pred_step_size)
    for name, R in (("R1", R1), ("R2", R2)):
        if R is not None:
            R = np.asarray(R, dtype=np.float64)
            if R.ndim == 1:
                R = np.diag(R)
            assert R.shape == (p, p)
            flat = R.reshape(-1)
            arr = getattr(c, name)
            for i in range(p * p):
                arr[i] = flat[i]
    tgt = np.asarray(observation_target, dtype=np.float64).reshape(-1)
    c.has_target = int(tgt.size > 0)
    for i in range(tgt.size):
        c.target[i] = tgt[i]
    return c


def dim_critic(critic_struct, n, m):
    return lib().orc_dim_critic(CRITIC_STRUCTS[critic_struct], n, m)


def state_dyn(s, state, action):
    st, stp = _d(state)
    ac, acp = _d(np.atleast_1d(action))
    out = np.zeros(s.n)
    lib().orc_state_dyn(C.byref(s), stp, acp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def closed_loop_rhs(s, y, action):
    """Returns (rhs, clipped_action); the input ``action`` is not modified."""
    yy, yp = _d(y)
    ac = np.array(np.atleast_1d(action), dtype=np.float64)
    out = np.zeros(s.n)
    lib().orc_closed_loop_rhs(C.byref(s), yp, ac.ctypes.data_as(C.POINTER(C.c_double)),
                              out.ctypes.data_as(C.POINTER(C.c_double)))
    return out, ac


class RK45:
    """One scipy-RK45-like solver lane bound to a system and a mutable action array."""

    def __init__(self, s, y0, t0, t_bound, max_step, first_step=1e-6, rtol=1e-3, atol=1e-5, action=None):
        self.s = s
        self.r = Rk45T()
        self.action = np.zeros(MAX_M) if action is None else np.array(action, dtype=np.float64)
        y0a, y0p = _d(y0)
        lib().orc_rk45_init(C.byref(self.r), C.byref(s), y0p, self._ap(), t0, t_bound, max_step,
                            first_step, rtol, atol)

    def _ap(self):
        return self.action.ctypes.data_as(C.POINTER(C.c_double))

    def receive_action(self, action):
        a = np.atleast_1d(np.asarray(action, dtype=np.float64))
        self.action[: a.size] = a

    def step(self):
        rc = lib().orc_rk45_step(C.byref(self.r), C.byref(self.s), self._ap())
        if rc < 0:
            raise RuntimeError("Attempt to step on a failed or finished solver.")
        return rc

    t = property(lambda self: self.r.t)
    h_abs = property(lambda self: self.r.h_abs)
    nfev = property(lambda self: self.r.nfev)
    status = property(lambda self: STATUS[self.r.status])
    y = property(lambda self: np.array(self.r.y[: self.s.n]))
    f = property(lambda self: np.array(self.r.f[: self.s.n]))


def stage_obj(c, n, m, obs, act):
    o, op = _d(obs)
    a, ap = _d(np.atleast_1d(act))
    return lib().orc_stage_obj(C.byref(c), n, m, op, ap)


def critic(c, n, m, obs, act, w):
    o, op = _d(obs)
    a, ap = _d(np.atleast_1d(act))
    ww, wp = _d(w)
    return lib().orc_critic(C.byref(c), n, m, op, ap, wp)


def critic_cost(c, n, m, obs_buf, act_buf, w, w_prev):
    ob, obp = _d(obs_buf)
    ab, abp = _d(act_buf)
    ww, wp = _d(w)
    wv, wvp = _d(w_prev)
    return lib().orc_critic_cost(C.byref(c), n, m, obp, abp, wp, wvp)


def actor_cost(c, s, action_sqn, observation, state_sys, w_critic=None):
    a, ap = _d(action_sqn)
    o, op = _d(observation)
    x, xp = _d(state_sys)
    if w_critic is None:
        w, wp = None, _null()
    else:
        w, wp = _d(w_critic)
    return lib().orc_actor_cost(C.byref(c), C.byref(s), ap, op, xp, wp)


def actor_grad(c, s, action_sqn, observation, state_sys, w_critic=None):
    """(J, dJ/d action_sqn) by the adjoint of the Euler rollout (rcg_oracle_opt.c)."""
    a, ap = _d(action_sqn)
    o, op = _d(observation)
    x, xp = _d(state_sys)
    w, wp = (None, _null()) if w_critic is None else _d(w_critic)
    g = np.zeros(a.size)
    J = lib().orc_actor_grad(C.byref(c), C.byref(s), ap, op, xp, wp, g.ctypes.data_as(C.POINTER(C.c_double)))
    return J, g


def actor_opt(c, s, action_sqn_init, observation, state_sys, w_critic=None, max_iter=100, pg_tol=1e-6, f_tol=1e-9, lanes=1):
    """Bounded minimiser of ``_actor_cost`` (projected L-BFGS, rcg_oracle_opt.c).  Returns (x, J, iters, nfev).
    ``lanes``: summation order of the inner products (1 = index order, the one-lane kernel; 4 = strided partial sums combined
    pairwise, the four-lanes-per-problem kernel)."""
    lib().orc_actor_opt_set_lanes(int(lanes))
    a = np.array(action_sqn_init, dtype=np.float64).reshape(-1).copy()
    o, op = _d(observation)
    x, xp = _d(state_sys)
    w, wp = (None, _null()) if w_critic is None else _d(w_critic)
    it, nf = C.c_int(0), C.c_int(0)
    J = lib().orc_actor_opt(C.byref(c), C.byref(s), a.ctypes.data_as(C.POINTER(C.c_double)), op, xp, wp,
                            int(max_iter), float(pg_tol), float(f_tol), C.byref(it), C.byref(nf))
    lib().orc_actor_opt_set_lanes(1)
    return a, J, it.value, nf.value


def actor_opt_hybrid(c, s, action_sqn_init, observation, state_sys, w_critic=None, max_sweeps=25, max_iter=300,
                     pg_tol=1e-7, f_tol=1e-12):
    """Control-limited Gauss-Newton (iLQR) sweeps, then the L-BFGS iteration from there (rcg_oracle_opt.c): groundwork
    for the next optimiser kernel.  Returns (x, J, sweeps, lbfgs_iters)."""
    a = np.array(action_sqn_init, dtype=np.float64).reshape(-1).copy()
    o, op = _d(observation)
    x, xp = _d(state_sys)
    w, wp = (None, _null()) if w_critic is None else _d(w_critic)
    sw, it = C.c_int(0), C.c_int(0)
    J = lib().orc_actor_opt_hybrid(C.byref(c), C.byref(s), a.ctypes.data_as(C.POINTER(C.c_double)), op, xp, wp, int(max_sweeps),
                                   int(max_iter), float(pg_tol), float(f_tol), C.byref(sw), C.byref(it))
    return a, J, sw.value, it.value


def actor_opt_batch(c, s, x_init, states, w_critic=None, max_iter=300, pg_tol=1e-7, f_tol=1e-12, nthreads=0):
    """orc_actor_opt for every row of ``states`` [E, n] (observation = state_sys), OpenMP over the problems.
    Returns (J [E], gradient evaluations)."""
    xi, xip = _d(np.asarray(x_init, dtype=np.float64).reshape(-1))
    st, stp = _d(np.atleast_2d(states))
    w, wp = (None, _null()) if w_critic is None else _d(w_critic)
    J = np.zeros(st.shape[0])
    g = lib().orc_actor_opt_batch(C.byref(c), C.byref(s), st.shape[0], xip, stp, wp, int(max_iter), float(pg_tol), float(f_tol),
                                  int(nthreads), J.ctypes.data_as(C.POINTER(C.c_double)))
    return J, int(g)


def nominal_ni(ctrl_gain, s, observation):
    """CtrlNominal3WRobotNI: action for one observation (clock-free part of compute_action, with clipping)."""
    o, op = _d(observation)
    a = np.zeros(2)
    lib().orc_nominal_ni(float(ctrl_gain), C.byref(s), op, a.ctypes.data_as(C.POINTER(C.c_double)))
    return a


def actor_cost_table(c, s, cand, observation, state_sys, w_critic=None):
    """J[C] and np.argmin(J) for a candidate table [C, Nactor*m]."""
    cd, cp = _d(cand)
    o, op = _d(observation)
    x, xp = _d(state_sys)
    if w_critic is None:
        w, wp = None, _null()
    else:
        w, wp = _d(w_critic)
    ncand = cd.shape[0]
    J = np.zeros(ncand)
    am = C.c_int(-1)
    lib().orc_actor_cost_table(C.byref(c), C.byref(s), ncand, cp, op, xp, wp,
                               J.ctypes.data_as(C.POINTER(C.c_double)), C.byref(am))
    return J, am.value


def argmin(J):
    j, jp = _d(J)
    return lib().orc_argmin(jp, j.size)


def closed_loop(c, s, state_init, cand, action_init, sampling_time, t0, t1, max_step, first_step=1e-6,
                rtol=1e-3, atol=1e-5, w_critic=None, max_steps_per_env=1 << 30, nthreads=0, traj_cap=0):
    """Closed loop with the enumerate-and-argmin controller for E envs.

    ``state_init`` [E, n]; ``cand`` [C, N*m] (shared) or [E, C, N*m] (per env).
    Returns a dict of per-env results (+ ``traj`` of env 0 if ``traj_cap`` > 0).
    """
    x0, x0p = _d(np.atleast_2d(state_init))
    E = x0.shape[0]
    cd, cp = _d(cand)
    per_env = int(cd.ndim == 3)
    ncand = cd.shape[-2]
    ai, aip = _d(np.atleast_1d(action_init))
    if w_critic is None:
        w, wp = None, _null()
    else:
        w, wp = _d(w_critic)
    n, m = s.n, s.m
    yf = np.zeros((E, n)); tf = np.zeros(E); acc = np.zeros(E)
    nst = np.zeros(E, dtype=np.int32); nsa = np.zeros(E, dtype=np.int32); nfe = np.zeros(E, dtype=np.int64)
    ncol = 1 + n + m + 3
    traj = np.zeros((max(traj_cap, 1), ncol))
    rows = C.c_int(0)
    evals = C.c_longlong(0)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    total = lib().orc_closed_loop(
        C.byref(c), C.byref(s), E, x0p, ncand, cp, per_env, wp, aip, sampling_time, t0, t1, max_step,
        first_step, rtol, atol, int(max_steps_per_env), int(nthreads),
        yf.ctypes.data_as(dp), tf.ctypes.data_as(dp), acc.ctypes.data_as(dp), nst.ctypes.data_as(ip),
        nsa.ctypes.data_as(ip), nfe.ctypes.data_as(C.POINTER(C.c_long)),
        traj.ctypes.data_as(dp) if traj_cap > 0 else _null(), int(traj_cap), C.byref(rows), C.byref(evals))
    return {"y": yf, "t": tf, "accum": acc, "nsteps": nst, "nsamples": nsa, "nfev": nfe,
            "traj": traj[: rows.value], "total_steps": int(total), "total_evals": int(evals.value)}


def critic_fit(c, n, m, obs_buf, act_buf, w_prev, w_min, w_max, w_init=None, max_evals=0, ls_mode=None):
    """Bounded least-squares stand-in for ``_critic_optimizer`` (rcg_oracle_critic.c: the algorithm of the product's
    ``rcg_critic_fit`` restated in scalar C).  Buffers ``[buffer_size, n]`` / ``[buffer_size, m]`` (row 0 = oldest).
    Returns (w, J_c(w), dual passes)."""
    ob, obp = _d(obs_buf)
    ab, abp = _d(act_buf)
    wv, wvp = _d(w_prev)
    D = dim_critic({v: k for k, v in CRITIC_STRUCTS.items()}[c.critic_struct], n, m)
    wi = np.ones(D) if w_init is None else np.asarray(w_init, dtype=np.float64)
    wi, wip = _d(wi)
    w = np.zeros(D)
    ev = C.c_int(0)
    if ls_mode is None:      # like rcg_critic_fit: exact line search on the two-phase path (K <= 3, >= 10 weights, to convergence)
        J = lib().orc_critic_fit(C.byref(c), n, m, obp, abp, wvp, float(w_min), float(w_max), wip,
                                 w.ctypes.data_as(C.POINTER(C.c_double)), int(max_evals), C.byref(ev))
    else:
        J = lib().orc_critic_fit_ls(C.byref(c), n, m, obp, abp, wvp, float(w_min), float(w_max), wip,
                                    w.ctypes.data_as(C.POINTER(C.c_double)), int(max_evals), int(ls_mode), C.byref(ev))
    return w, J, ev.value


def closed_loop_critic(c, s, state_init, cand, action_init, sampling_time, t0, t1, max_step, buffer_size, w_bounds,
                       critic_period=None, first_step=1e-6, rtol=1e-3, atol=1e-5, w_replay=None,
                       max_steps_per_env=1 << 30, nthreads=0, traj_cap=0):
    """RQL / SQL closed loop INCLUDING the critic branch of ``compute_action`` (controllers.py:1455-1479) with the
    enumerate-and-argmin actor: FIFO buffers, critic clock, refit by :func:`critic_fit` -- or, with ``w_replay``
    ``[F, dimc]``, fit number j LOADS ``w_replay[j]`` (the reference's recorded SLSQP results).  ``traj`` rows of env 0:
    [t, y(n), action(m), accum, argmin, Jmin, sampled, nfits]."""
    x0, x0p = _d(np.atleast_2d(state_init))
    E = x0.shape[0]
    cd, cp = _d(cand)
    per_env = int(cd.ndim == 3)
    ncand = cd.shape[-2]
    ai, aip = _d(np.atleast_1d(action_init))
    n, m = s.n, s.m
    D = lib().orc_dim_critic(c.critic_struct, n, m)
    if w_replay is None:
        wr, wrp, nrep = None, _null(), 0
    else:
        wr, wrp = _d(np.atleast_2d(w_replay))
        nrep = wr.shape[0]
    yf = np.zeros((E, n)); tf = np.zeros(E); acc = np.zeros(E)
    nst = np.zeros(E, dtype=np.int32); nsa = np.zeros(E, dtype=np.int32); nft = np.zeros(E, dtype=np.int32)
    wf = np.zeros((E, D)); Jc = np.zeros(E)
    obf = np.zeros((E, buffer_size, n)); abf = np.zeros((E, buffer_size, m))
    ncol = 1 + n + m + 5
    traj = np.zeros((max(traj_cap, 1), ncol))
    rows = C.c_int(0)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    total = lib().orc_closed_loop_critic(
        C.byref(c), C.byref(s), E, x0p, ncand, cp, per_env, aip, int(buffer_size), float(w_bounds[0]), float(w_bounds[1]),
        float(sampling_time), float(sampling_time if critic_period is None else critic_period), float(t0), float(t1),
        float(max_step), float(first_step), float(rtol), float(atol), int(max_steps_per_env), int(nthreads), wrp, nrep,
        yf.ctypes.data_as(dp), tf.ctypes.data_as(dp), acc.ctypes.data_as(dp), nst.ctypes.data_as(ip), nsa.ctypes.data_as(ip),
        nft.ctypes.data_as(ip), wf.ctypes.data_as(dp), Jc.ctypes.data_as(dp), obf.ctypes.data_as(dp), abf.ctypes.data_as(dp),
        traj.ctypes.data_as(dp) if traj_cap > 0 else _null(), int(traj_cap), C.byref(rows))
    if total < 0:
        raise ValueError("buffer_size exceeds the oracle's ORC_MAX_BUF")
    return {"y": yf, "t": tf, "accum": acc, "nsteps": nst, "nsamples": nsa, "nfits": nft, "w_critic": wf, "Jc": Jc,
            "obs_buf": obf, "act_buf": abf, "traj": traj[: rows.value], "total_steps": int(total)}


DIST_DIM = {0: 2, 1: 2, 2: 1}


def make_dist(pars_disturb, seed=0) -> DistT:
    """``pars_disturb = [sigma_disturb, mu_disturb, tau_disturb]`` (each ``[dim_disturb]``) + the key of the draw stream."""
    d = DistT()
    p = np.zeros((3, 2))
    for i, row in enumerate(pars_disturb):
        r = np.atleast_1d(np.asarray(row, dtype=np.float64))
        p[i, : r.size] = r[:2]
    for k in range(2):
        d.sigma[k], d.mu[k], d.tau[k] = p[0, k], p[1, k], p[2, k]
    d.seed = int(seed)
    return d


def state_dyn_disturbed(s, state, action, disturb):
    st, stp = _d(state)
    ac, acp = _d(np.atleast_1d(action))
    q, qp = _d(np.resize(np.atleast_1d(np.asarray(disturb, dtype=np.float64)), 2))
    out = np.zeros(s.n)
    lib().orc_state_dyn_disturbed(C.byref(s), stp, acp, qp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def disturb_dyn(s, dist, disturb, z):
    """``_disturb_dyn`` given the draws ``z`` (one ``randn()`` per component)."""
    nd = DIST_DIM[s.sys_id]
    q, qp = _d(np.resize(np.atleast_1d(np.asarray(disturb, dtype=np.float64)), 2))
    zz, zp = _d(np.resize(np.atleast_1d(np.asarray(z, dtype=np.float64)), 2))
    out = np.zeros(2)
    lib().orc_disturb_dyn(C.byref(s), C.byref(dist), qp, zp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[:nd]


def closed_loop_rhs_disturbed(s, dist, state_full, action, env=0, call=0, z=None):
    """(rhs_full, clipped action) of ``closed_loop_rhs`` with ``is_disturb = 1``; ``z`` given = patched ``randn``."""
    nd = DIST_DIM[s.sys_id]
    y, yp = _d(state_full)
    ac = np.array(np.atleast_1d(action), dtype=np.float64)
    zz, zp = (None, _null()) if z is None else _d(np.resize(np.atleast_1d(np.asarray(z, dtype=np.float64)), 2))
    out = np.zeros(s.n + 2)
    lib().orc_closed_loop_rhs_disturbed(C.byref(s), C.byref(dist), yp, ac.ctypes.data_as(C.POINTER(C.c_double)), int(env), int(call),
                                        zp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out[: s.n + nd], ac


def normal2(seed, env, call):
    """The two standard-normal draws of RHS call ``call`` of global environment ``env`` (the specified stream)."""
    z = np.zeros(2)
    lib().orc_normal2(int(seed), int(env), int(call), z.ctypes.data_as(C.POINTER(C.c_double)))
    return z


def det_log(x):
    return lib().orc_log(float(x))


class RK45Disturbed:
    """scipy-RK45-like solver lane on the full state ``[state, disturb]`` of a disturbed system."""

    def __init__(self, s, dist, y0_full, t0, t_bound, max_step, env=0, first_step=1e-6, rtol=1e-3, atol=1e-5, action=None):
        self.s, self.dist = s, dist
        self.r = Rk45dT()
        self.action = np.zeros(MAX_M) if action is None else np.array(action, dtype=np.float64)
        y0a, y0p = _d(y0_full)
        lib().orc_rk45d_init(C.byref(self.r), C.byref(s), C.byref(dist), int(env), y0p, self._ap(), t0, t_bound, max_step,
                             first_step, rtol, atol)

    def _ap(self):
        return self.action.ctypes.data_as(C.POINTER(C.c_double))

    def receive_action(self, action):
        a = np.atleast_1d(np.asarray(action, dtype=np.float64))
        self.action[: a.size] = a

    def step(self):
        rc = lib().orc_rk45d_step(C.byref(self.r), C.byref(self.s), C.byref(self.dist), self._ap())
        if rc < 0:
            raise RuntimeError("Attempt to step on a failed or finished solver.")
        return rc

    t = property(lambda self: self.r.t)
    h_abs = property(lambda self: self.r.h_abs)
    nfev = property(lambda self: self.r.nfev)
    status = property(lambda self: STATUS[self.r.status])
    y = property(lambda self: np.array(self.r.y[: self.r.nfull]))
    f = property(lambda self: np.array(self.r.f[: self.r.nfull]))


def sincos(x):
    """The oracle's deterministic (sin, cos) -- the spec the CUDA fp64 path reproduces bit for bit."""
    s, c = C.c_double(), C.c_double()
    lib().orc_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def pow_m02(x):
    """The oracle's correctly rounded x ** -0.2 (scipy rk.py:153, :162 ERROR_EXPONENT)."""
    return lib().orc_pow_m02(float(x))


def num_threads() -> int:
    """Threads the OpenMP loops of the oracle will use (1 without OpenMP)."""
    return int(lib().orc_num_threads())


def has_openmp() -> bool:
    """True if the oracle was compiled with OpenMP (its batch loops then honour the ``nthreads`` argument)."""
    return bool(lib().orc_has_openmp())


class EnvBatch:
    """E resumable closed-loop environments (``orc_env_t``): ``interval()`` advances every live
    environment up to and including its next controller sample, like one
    ``ClosedLoopEngine.run_interval`` of the product -- used interval by interval in tests and as
    the timed unit of bench.py's CPU arms."""

    def __init__(self, c, s, state_init, cand, action_init, sampling_time, t0, t1, max_step, first_step=1e-6,
                 rtol=1e-3, atol=1e-5, w_critic=None):
        self.c, self.s = c, s
        self.x0, x0p = _d(np.atleast_2d(state_init))
        self.E = self.x0.shape[0]
        self.cand, self._cp = _d(cand)
        self.per_env = int(self.cand.ndim == 3)
        self.C = self.cand.shape[-2]
        ai, aip = _d(np.atleast_1d(action_init))
        self.w, self._wp = (None, _null()) if w_critic is None else _d(w_critic)
        self.sampling_time, self.t1 = float(sampling_time), float(t1)
        self.envs = (EnvT * self.E)()
        lib().orc_env_init(self.envs, C.byref(s), self.E, x0p, aip, t0, t1, max_step, first_step, rtol, atol)

    def interval(self, nthreads=0):
        """Returns (accepted solver steps, _actor_cost evaluations) of this interval."""
        ev = C.c_longlong(0)
        steps = lib().orc_env_interval(self.envs, C.byref(self.c), C.byref(self.s), self.E, self.C, self._cp,
                                       self.per_env, self._wp, self.sampling_time, self.t1, int(nthreads),
                                       C.byref(ev))
        return int(steps), int(ev.value)

    def snapshot(self):
        n, m = self.s.n, self.s.m
        v = self.envs
        return {
            "t": np.array([v[e].r.t for e in range(self.E)]),
            "y": np.array([[v[e].r.y[i] for i in range(n)] for e in range(self.E)]),
            "action": np.array([[v[e].action_curr[j] for j in range(m)] for e in range(self.E)]),
            "accum": np.array([v[e].accum for e in range(self.E)]),
            "argmin": np.array([v[e].best for e in range(self.E)], dtype=np.int32),
            "Jmin": np.array([v[e].Jbest for e in range(self.E)]),
            "nsteps": np.array([v[e].steps for e in range(self.E)], dtype=np.int32),
            "nsamples": np.array([v[e].samples for e in range(self.E)], dtype=np.int32),
            "nfev": np.array([v[e].r.nfev for e in range(self.E)], dtype=np.int64),
            "done": np.array([v[e].done for e in range(self.E)], dtype=np.int32),
        }
