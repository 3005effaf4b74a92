"""ctypes binding of ``librcg_b200.so`` (C ABI declared in ``include/rcg.h``).

The library is the product's only compute path: there is no CPU or PyTorch fallback.  If the
shared object is missing this module raises at import, and every compute entry point returns
an error (raised here as ``RuntimeError``) when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librcg_b200.so")

MAX_N, MAX_M, MAX_P, MAX_W, MAX_NACTOR = 5, 2, 7, 35, 64

SYS_IDS = {"3wrobotNI": 0, "3wrobot": 1, "2tank": 2}
SYS_DIMS = {0: (3, 2), 1: (5, 2), 2: (2, 1)}
MODES = {"MPC": 0, "RQL": 1, "SQL": 2}
CRITIC_STRUCTS = {"quad-lin": 0, "quadratic": 1, "quad-nomix": 2, "quad-mix": 3}
STAGE_STRUCTS = {"quadratic": 0, "biquadratic": 1}
RUNNING, FINISHED, FAILED = 0, 1, 2
STATUS_NAMES = {RUNNING: "running", FINISHED: "finished", FAILED: "failed"}


class RcgSystem(C.Structure):
    _fields_ = [("sys_id", C.c_int32), ("has_bnds", C.c_int32), ("pars", C.c_double * 8),
                ("lo", C.c_double * MAX_M), ("hi", C.c_double * MAX_M)]


class RcgObjective(C.Structure):
    _fields_ = [("mode", C.c_int32), ("critic_struct", C.c_int32), ("stage_struct", C.c_int32),
                ("r_is_diag", C.c_int32), ("has_target", C.c_int32), ("Nactor", C.c_int32),
                ("Ncritic", C.c_int32), ("buffer_size", C.c_int32), ("gamma", C.c_double),
                ("pred_step_size", C.c_double), ("gamma_pow", C.c_double * MAX_NACTOR),
                ("R1", C.c_double * (MAX_P * MAX_P)), ("R2", C.c_double * (MAX_P * MAX_P)),
                ("target", C.c_double * MAX_N)]


class RcgDisturb(C.Structure):
    _fields_ = [("sigma", C.c_double * 2), ("mu", C.c_double * 2), ("tau", C.c_double * 2), ("seed", C.c_uint64),
                ("env_offset", C.c_int64)]


DIST_DIM = {0: 2, 1: 2, 2: 1}          # dim_disturb per system (the presets' values)


class RcgLog(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("count", C.c_void_p), ("capacity", C.c_int32), ("every", C.c_int32)]


class RcgSolver(C.Structure):
    _fields_ = [("t_bound", C.c_double), ("max_step", C.c_double), ("rtol", C.c_double), ("atol", C.c_double)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m rcognita_b200.build` "
            "(or __graft_entry__.build()). rcognita_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sysp, objp, solp = C.POINTER(RcgSystem), C.POINTER(RcgObjective), C.POINTER(RcgSolver)
    L.rcg_version.restype = C.c_int
    L.rcg_last_error_string.restype = C.c_char_p
    L.rcg_device_count.restype = C.c_int
    L.rcg_last_actor_kernel.restype = C.c_char_p
    L.rcg_last_actor_opt_kernel.restype = C.c_char_p
    L.rcg_actor_opt_lanes.argtypes = [i32]
    L.rcg_actor_opt_lanes.restype = C.c_int
    L.rcg_dim_state.argtypes = [i32]
    L.rcg_dim_input.argtypes = [i32]
    L.rcg_dim_critic.argtypes = [i32, i32, i32]
    L.rcg_launch_count.restype = C.c_int64
    L.rcg_reset_launch_count.restype = None
    for name in ("rcg_rhs", "rcg_rhs_f32", "rcg_state_dyn"):
        getattr(L, name).argtypes = [sysp, i64, vp, vp, vp, vp]
    for name in ("rcg_rk45_step", "rcg_rk45_step_f32"):
        getattr(L, name).argtypes = [sysp, solp, i64] + [vp] * 7 + [vp]
    for name in ("rcg_rk45_advance", "rcg_rk45_advance_f32"):
        getattr(L, name).argtypes = [sysp, solp, objp, i64] + [vp] * 9 + [dbl, i32, vp, vp, vp, vp, vp]
    L.rcg_rk45_advance_logged.argtypes = ([sysp, solp, objp, i64] + [vp] * 9 + [dbl, i32, vp, vp, vp, vp,
                                                                                C.POINTER(RcgLog), vp])
    distp = C.POINTER(RcgDisturb)
    L.rcg_rhs_disturbed.argtypes = [sysp, distp, i64, vp, vp, vp, vp, vp, i32, vp]
    L.rcg_disturb_normals.argtypes = [distp, i64, vp, i32, vp, vp]
    L.rcg_rk45_step_disturbed.argtypes = [sysp, distp, solp, i64] + [vp] * 7 + [vp]
    L.rcg_rk45_advance_disturbed.argtypes = [sysp, distp, solp, objp, i64] + [vp] * 9 + [dbl, i32, vp, vp, vp, vp, vp]
    L.rcg_log_rows.argtypes = [objp, i32, i32, i64, vp, vp, vp, vp, vp, vp, C.POINTER(RcgLog), vp]
    for name in ("rcg_actor_cost", "rcg_actor_cost_f32"):
        getattr(L, name).argtypes = [sysp, objp, i64, i32, vp, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, dbl, vp]
    L.rcg_actor_opt_workspace_bytes.argtypes = [sysp, objp, i64, i32]
    L.rcg_actor_opt_workspace_bytes.restype = C.c_int64
    L.rcg_actor_opt.argtypes = ([sysp, objp, i64, i32, vp, vp, vp, vp, i32, vp, i32, dbl, dbl, vp, i64]
                                + [vp] * 7 + [dbl, vp])
    L.rcg_actor_grad.argtypes = [sysp, objp, i64, i32, vp, vp, vp, vp, i32, vp, i64, vp, vp, vp]
    L.rcg_actor_ilqr_workspace_bytes.argtypes = [sysp, objp, i64, i32]
    L.rcg_actor_ilqr_workspace_bytes.restype = C.c_int64
    L.rcg_actor_ilqr.argtypes = [sysp, objp, i64, i32, vp, vp, vp, vp, i32, vp, i32, dbl, vp, i64, vp, vp]
    L.rcg_gather_sqn.argtypes = [i32, i64, i32, vp, i32, vp, vp, vp, vp]
    L.rcg_nominal_ni.argtypes = [sysp, i64, vp, dbl, vp, vp, objp, vp, dbl, vp]
    L.rcg_stage_obj.argtypes = [objp, i32, i32, i64, vp, vp, vp, vp, dbl, vp]
    L.rcg_critic.argtypes = [objp, i32, i32, i64, vp, vp, vp, i32, vp, vp]
    L.rcg_critic_cost.argtypes = [objp, i32, i32, i64, i32, vp, vp, vp, vp, vp, vp]
    L.rcg_critic_cost_f32.argtypes = [objp, i32, i32, i64, i32, vp, vp, vp, vp, vp, vp]
    L.rcg_critic_fit.argtypes = [objp, i32, i32, i64, vp, vp, vp, dbl, dbl, vp, vp, vp, i32, i32, vp, vp]
    L.rcg_ctrl_sample.argtypes = [i64, vp, vp, dbl, vp, vp, vp]
    L.rcg_push_buffers.argtypes = [i32, i32, i32, i64, vp, vp, vp, vp, vp, vp]
    return L


lib = _load()

EXPORTS = [
    "rcg_version", "rcg_last_error_string", "rcg_device_count", "rcg_dim_state", "rcg_dim_input", "rcg_dim_critic",
    "rcg_launch_count", "rcg_reset_launch_count", "rcg_last_actor_kernel", "rcg_rhs", "rcg_rhs_f32", "rcg_state_dyn", "rcg_rk45_step",
    "rcg_rk45_step_f32", "rcg_rk45_advance", "rcg_rk45_advance_f32", "rcg_rk45_advance_logged", "rcg_log_rows",
    "rcg_rhs_disturbed", "rcg_disturb_normals", "rcg_rk45_step_disturbed", "rcg_rk45_advance_disturbed", "rcg_actor_cost", "rcg_actor_cost_f32",
    "rcg_actor_opt_workspace_bytes", "rcg_actor_opt", "rcg_actor_opt_lanes", "rcg_last_actor_opt_kernel", "rcg_actor_grad", "rcg_actor_ilqr_workspace_bytes", "rcg_actor_ilqr", "rcg_gather_sqn", "rcg_nominal_ni", "rcg_stage_obj", "rcg_critic", "rcg_critic_cost", "rcg_critic_cost_f32", "rcg_critic_fit", "rcg_ctrl_sample", "rcg_push_buffers",
]


def last_error() -> str:
    return lib.rcg_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "librcg_b200") -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def make_system(name: str, pars=(), ctrl_bnds=None) -> RcgSystem:
    """Descriptor of a reference ``System`` (``name`` = ``System.name``)."""
    if name not in SYS_IDS:
        raise ValueError(f"unknown system name {name!r}; the engine implements {sorted(SYS_IDS)}")
    s = RcgSystem()
    s.sys_id = SYS_IDS[name]
    for i, p in enumerate(list(pars)[:8]):
        s.pars[i] = float(p)
    b = np.zeros((0, 2)) if ctrl_bnds is None else np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2)
    s.has_bnds = int(b.size > 0 and bool(b.any()))
    for k in range(min(b.shape[0], MAX_M)):
        s.lo[k], s.hi[k] = b[k, 0], b[k, 1]
    return s


def make_disturb(pars_disturb, seed=0, env_offset=0) -> RcgDisturb:
    """Descriptor of ``pars_disturb = [sigma_disturb, mu_disturb, tau_disturb]`` (rcognita/systems.py:303-306, :366-369) plus the
    key of the per-environment draw stream and the global index of lane 0."""
    d = RcgDisturb()
    rows = list(pars_disturb) if len(pars_disturb) else []
    p = np.zeros((3, 2))
    for i, row in enumerate(rows[:3]):
        r = np.atleast_1d(np.asarray(row, dtype=np.float64)).reshape(-1)
        p[i, : min(2, r.size)] = r[:2]
    for k in range(2):
        d.sigma[k], d.mu[k], d.tau[k] = p[0, k], p[1, k], p[2, k]
    d.seed, d.env_offset = int(seed), int(env_offset)
    return d


def make_objective(n: int, m: int, mode="MPC", Nactor=1, pred_step_size=0.1, gamma=1.0, Ncritic=4, buffer_size=20,
                   critic_struct="quad-nomix", stage_obj_struct="quadratic", R1=None, R2=None,
                   observation_target=()) -> RcgObjective:
    """Descriptor of the cost-relevant fields of a reference ``CtrlOptPred``."""
    o = RcgObjective()
    p = n + m
    o.mode, o.critic_struct, o.stage_struct = MODES[mode], CRITIC_STRUCTS[critic_struct], STAGE_STRUCTS[stage_obj_struct]
    if not 1 <= int(Nactor) <= MAX_NACTOR:
        raise ValueError(f"Nactor must be in [1, {MAX_NACTOR}]")
    o.Nactor = int(Nactor)
    o.buffer_size = int(buffer_size)
    o.Ncritic = int(min(Ncritic, buffer_size - 1))            # controllers.py:1015
    o.gamma, o.pred_step_size = float(gamma), float(pred_step_size)
    for k in range(MAX_NACTOR):
        o.gamma_pow[k] = math.pow(float(gamma), float(k))     # Python's gamma**k (controllers.py:1306)
    diag = True
    for name, R in (("R1", R1), ("R2", R2)):
        if R is None:
            continue
        R = np.asarray(R, dtype=np.float64)
        if R.ndim == 1:
            R = np.diag(R)
        if R.shape != (p, p):
            raise ValueError(f"{name} must have shape ({p}, {p})")
        diag = diag and bool(np.count_nonzero(R - np.diag(np.diag(R))) == 0)
        arr = getattr(o, name)
        flat = R.reshape(-1)
        for i in range(p * p):
            arr[i] = flat[i]
    o.r_is_diag = int(diag)
    tgt = np.asarray(observation_target, dtype=np.float64).reshape(-1)
    o.has_target = int(tgt.size > 0)
    for i in range(min(tgt.size, MAX_N)):
        o.target[i] = tgt[i]
    return o


def make_solver(t_bound, max_step, rtol=1e-3, atol=1e-5) -> RcgSolver:
    s = RcgSolver()
    s.t_bound, s.max_step, s.rtol, s.atol = float(t_bound), float(max_step), float(rtol), float(atol)
    return s


def dim_critic(critic_struct: str, n: int, m: int) -> int:
    return int(lib.rcg_dim_critic(CRITIC_STRUCTS[critic_struct], n, m))
