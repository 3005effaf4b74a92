"""Tensor-level launchers: CUDA tensors in struct-of-arrays layout -> ``librcg_b200.so``.

Layout: a per-lane vector quantity of dimension d is a contiguous ``[d, E]`` tensor
(component-major).  PyTorch only provides device memory and the stream; all arithmetic
happens in the library's hand-written sm_100a kernels.  No fallback exists: non-CUDA
tensors raise.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _C

_F64, _F32, _I32 = torch.float64, torch.float32, torch.int32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=None, shape=None, name="tensor", optional=False):
    if t is None:
        if optional:
            return C.c_void_p(0)
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (rcognita_b200 has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the current device's current stream: a tensor of another GPU would be an illegal address
        raise RuntimeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                           f"wrap the call in `with torch.cuda.device({t.device.index}):`")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return C.c_void_p(t.data_ptr())


def _suffix(dtype):
    if dtype == _F64:
        return ""
    if dtype == _F32:
        return "_f32"
    raise TypeError(f"unsupported dtype {dtype}")


def rhs(sysd, y, action, out=None):
    """``System.closed_loop_rhs``: clips ``action`` [m,E] in place, returns f [n,E]."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = y.shape[1]
    out = torch.empty_like(y) if out is None else out
    fn = getattr(_C.lib, "rcg_rhs" + _suffix(y.dtype))
    _C.check(fn(C.byref(sysd), E, _ptr(y, None, (n, E), "y"), _ptr(action, y.dtype, (m, E), "action"),
                _ptr(out, y.dtype, (n, E), "out"), _stream()), "rcg_rhs")
    return out


def state_dyn(sysd, state, action, out=None):
    """``System._state_dyn`` (no clipping)."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = state.shape[1]
    out = torch.empty_like(state) if out is None else out
    _C.check(_C.lib.rcg_state_dyn(C.byref(sysd), E, _ptr(state, _F64, (n, E), "state"),
                                  _ptr(action, _F64, (m, E), "action"), _ptr(out, _F64, (n, E), "out"), _stream()),
             "rcg_state_dyn")
    return out


def rk45_step(sysd, sol, y, f, t, h_abs, status, action, nfev=None):
    """``Simulator.sim_step``: one accepted RK45 step per running lane, in place."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = y.shape[1]
    fn = getattr(_C.lib, "rcg_rk45_step" + _suffix(y.dtype))
    _C.check(fn(C.byref(sysd), C.byref(sol), E, _ptr(y, None, (n, E), "y"), _ptr(f, y.dtype, (n, E), "f"),
                _ptr(t, _F64, (E,), "t"), _ptr(h_abs, _F64, (E,), "h_abs"), _ptr(status, _I32, (E,), "status"),
                _ptr(nfev, _I32, (E,), "nfev", optional=True), _ptr(action, y.dtype, (m, E), "action"), _stream()),
             "rcg_rk45_step")


def rhs_disturbed(sysd, distd, y_full, action, call=None, normals=None, out=None, clip=True):
    """``System.closed_loop_rhs`` with ``is_disturb = 1`` on the full state ``[n + nd, E]``: clips ``action`` in place
    (``clip``), ``out[:n] = _state_dyn(state, action, disturb)``, ``out[n:] = _disturb_dyn(disturb)`` with the draws
    ``normals`` ``[2, E]`` when given, else the environment's stream at RHS call number ``call`` ``[E]`` int32 (0 if None)."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    nf = n + _C.DIST_DIM[sysd.sys_id]
    E = y_full.shape[1]
    out = torch.empty_like(y_full) if out is None else out
    _C.check(_C.lib.rcg_rhs_disturbed(C.byref(sysd), C.byref(distd), E, _ptr(y_full, _F64, (nf, E), "y_full"),
                                      _ptr(action, _F64, (m, E), "action"), _ptr(call, _I32, (E,), "call", optional=True),
                                      _ptr(normals, _F64, (2, E), "normals", optional=True), _ptr(out, _F64, (nf, E), "out"),
                                      int(bool(clip)), _stream()), "rcg_rhs_disturbed")
    return out


def disturb_normals(distd, E, call=0, device=None):
    """The two standard-normal draws ``[2, E]`` of RHS call ``call`` (an int, or ``[E]`` int32) of every environment."""
    per_lane = isinstance(call, torch.Tensor)
    out = torch.empty((2, E), dtype=_F64, device=call.device if per_lane else device)
    _C.check(_C.lib.rcg_disturb_normals(C.byref(distd), E, _ptr(call, _I32, (E,), "call") if per_lane else C.c_void_p(0),
                                        0 if per_lane else int(call), _ptr(out), _stream()), "rcg_disturb_normals")
    return out


def rk45_step_disturbed(sysd, distd, sol, y_full, f_full, t, h_abs, status, action, nfev):
    """``Simulator.sim_step`` of a disturbed system: one accepted RK45 step of the full state per running lane."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    nf = n + _C.DIST_DIM[sysd.sys_id]
    E = y_full.shape[1]
    _C.check(_C.lib.rcg_rk45_step_disturbed(C.byref(sysd), C.byref(distd), C.byref(sol), E, _ptr(y_full, _F64, (nf, E), "y_full"),
                                            _ptr(f_full, _F64, (nf, E), "f_full"), _ptr(t, _F64, (E,), "t"),
                                            _ptr(h_abs, _F64, (E,), "h_abs"), _ptr(status, _I32, (E,), "status"),
                                            _ptr(nfev, _I32, (E,), "nfev"), _ptr(action, _F64, (m, E), "action"), _stream()),
             "rcg_rk45_step_disturbed")


def rk45_advance_disturbed(sysd, distd, sol, obj, y_full, f_full, t, h_abs, status, action, ctrl_clock, sampling_time, max_steps,
                           nfev, state_sys=None, accum=None, sample_flag=None, nsteps=None, nsamples=None):
    """``rk45_advance`` on the full state of a disturbed system (``state_sys`` / ``accum`` see the state rows only)."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    nf = n + _C.DIST_DIM[sysd.sys_id]
    E = y_full.shape[1]
    _C.check(_C.lib.rcg_rk45_advance_disturbed(
        C.byref(sysd), C.byref(distd), C.byref(sol), C.byref(obj), E, _ptr(y_full, _F64, (nf, E), "y_full"),
        _ptr(f_full, _F64, (nf, E), "f_full"), _ptr(t, _F64, (E,), "t"), _ptr(h_abs, _F64, (E,), "h_abs"),
        _ptr(status, _I32, (E,), "status"), _ptr(nfev, _I32, (E,), "nfev"), _ptr(nsteps, _I32, (E,), "nsteps", optional=True),
        _ptr(action, _F64, (m, E), "action"), _ptr(ctrl_clock, _F64, (E,), "ctrl_clock"), float(sampling_time), int(max_steps),
        _ptr(state_sys, _F64, (n, E), "state_sys", optional=True), _ptr(accum, _F64, (E,), "accum", optional=True),
        _ptr(sample_flag, _I32, (E,), "sample_flag", optional=True), _ptr(nsamples, _I32, (E,), "nsamples", optional=True),
        _stream()), "rcg_rk45_advance_disturbed")


class TrajectoryLog:
    """Device-side trajectory ring of ``rcg_rk45_advance_logged`` / ``rcg_log_rows``: ``rows`` [capacity, ncols, E]
    with columns (t, state[n], stage_obj, accum_obj, action[m]); ``count`` [E] rows written per lane."""

    def __init__(self, n, m, E, capacity, every=1, device=None):
        self.n, self.m, self.E, self.capacity, self.every = n, m, E, int(capacity), int(every)
        self.ncols = 1 + n + 2 + m
        self.rows = torch.full((self.capacity, self.ncols, E), float("nan"), dtype=_F64, device=device)
        self.count = torch.zeros((E,), dtype=_I32, device=device)
        self.desc = _C.RcgLog(self.rows.data_ptr(), self.count.data_ptr(), self.capacity, self.every)

    def reset(self):
        self.rows.fill_(float("nan"))
        self.count.zero_()

    def env_rows(self, e):
        """Chronological rows of environment ``e`` (oldest kept first) as a host array [rows, ncols]."""
        c = int(self.count[e].item())
        k = min(c, self.capacity)
        idx = [(c - k + i) % self.capacity for i in range(k)]
        return self.rows[idx, :, e].cpu().numpy()


def log_rows(obj, n, m, log, t, y, action, accum, nsteps=None, mask=None):
    """Append the current row of the masked lanes to the trajectory ring (``rcg_log_rows``)."""
    E = y.shape[1]
    _C.check(_C.lib.rcg_log_rows(C.byref(obj), n, m, E, _ptr(t, _F64, (E,), "t"), _ptr(y, _F64, (n, E), "y"),
                                 _ptr(action, _F64, (m, E), "action"), _ptr(accum, _F64, (E,), "accum"),
                                 _ptr(nsteps, _I32, (E,), "nsteps", optional=True),
                                 _ptr(mask, _I32, (E,), "mask", optional=True), C.byref(log.desc), _stream()),
             "rcg_log_rows")


def rk45_advance(sysd, sol, obj, y, f, t, h_abs, status, action, ctrl_clock, sampling_time, max_steps,
                 state_sys=None, accum=None, sample_flag=None, nfev=None, nsteps=None, nsamples=None, log=None):
    """Fused loop body between two controller samples (see ``rcg_rk45_advance`` in rcg.h); with ``log`` (a
    ``TrajectoryLog``) the held-action steps are appended to the trajectory ring."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = y.shape[1]
    dt = y.dtype
    if log is not None:
        if dt != _F64:
            raise TypeError("trajectory logging runs in fp64")
        _C.check(_C.lib.rcg_rk45_advance_logged(
            C.byref(sysd), C.byref(sol), C.byref(obj), E, _ptr(y, None, (n, E), "y"), _ptr(f, dt, (n, E), "f"),
            _ptr(t, _F64, (E,), "t"), _ptr(h_abs, _F64, (E,), "h_abs"), _ptr(status, _I32, (E,), "status"),
            _ptr(nfev, _I32, (E,), "nfev", optional=True), _ptr(nsteps, _I32, (E,), "nsteps"),
            _ptr(action, dt, (m, E), "action"), _ptr(ctrl_clock, _F64, (E,), "ctrl_clock"), float(sampling_time),
            int(max_steps), _ptr(state_sys, dt, (n, E), "state_sys", optional=True), _ptr(accum, dt, (E,), "accum"),
            _ptr(sample_flag, _I32, (E,), "sample_flag", optional=True),
            _ptr(nsamples, _I32, (E,), "nsamples", optional=True), C.byref(log.desc), _stream()), "rcg_rk45_advance_logged")
        return
    fn = getattr(_C.lib, "rcg_rk45_advance" + _suffix(dt))
    _C.check(fn(C.byref(sysd), C.byref(sol), C.byref(obj), E, _ptr(y, None, (n, E), "y"), _ptr(f, dt, (n, E), "f"),
                _ptr(t, _F64, (E,), "t"), _ptr(h_abs, _F64, (E,), "h_abs"), _ptr(status, _I32, (E,), "status"),
                _ptr(nfev, _I32, (E,), "nfev", optional=True), _ptr(nsteps, _I32, (E,), "nsteps", optional=True),
                _ptr(action, dt, (m, E), "action"), _ptr(ctrl_clock, _F64, (E,), "ctrl_clock"),
                float(sampling_time), int(max_steps), _ptr(state_sys, dt, (n, E), "state_sys", optional=True),
                _ptr(accum, dt, (E,), "accum", optional=True), _ptr(sample_flag, _I32, (E,), "sample_flag", optional=True),
                _ptr(nsamples, _I32, (E,), "nsamples", optional=True), _stream()), "rcg_rk45_advance")


def actor_cost(sysd, obj, state_sys, obs, cand, cand_per_env, C_, w_critic=None, w_per_env=False, mask=None,
               want_J=True, J_out=None, argmin_out=None, Jmin_out=None, action_out=None, accum=None,
               sampling_time=0.0):
    """``CtrlOptPred._actor_cost`` over E x C candidates + per-env ``np.argmin``.

    ``cand``: ``[N*m, C]`` (shared) or ``[N*m, E*C]`` (per env, element (k,j,e,c) at column e*C+c).
    Returns ``(J [E,C] or None, argmin [E] int32, Jmin [E])``.
    """
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = obs.shape[1]
    dt = obs.dtype
    L = obj.Nactor * m
    ncol = E * C_ if cand_per_env else C_
    if want_J and J_out is None:
        J_out = torch.empty((E, C_), dtype=dt, device=obs.device)
    if argmin_out is None:
        argmin_out = torch.full((E,), -1, dtype=_I32, device=obs.device)
    if Jmin_out is None:
        Jmin_out = torch.full((E,), float("nan"), dtype=dt, device=obs.device)
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    wshape = None if w_critic is None else ((dimc, E) if w_per_env else (dimc,))
    fn = getattr(_C.lib, "rcg_actor_cost" + _suffix(dt))
    _C.check(fn(C.byref(sysd), C.byref(obj), E, int(C_), _ptr(state_sys, dt, (n, E), "state_sys"),
                _ptr(obs, dt, (n, E), "obs"), _ptr(cand, dt, (L, ncol), "cand"), int(bool(cand_per_env)),
                _ptr(w_critic, dt, wshape, "w_critic", optional=True), int(bool(w_per_env)),
                _ptr(mask, _I32, (E,), "mask", optional=True),
                _ptr(J_out, dt, (E, C_), "J_out", optional=True), _ptr(argmin_out, _I32, (E,), "argmin_out"),
                _ptr(Jmin_out, dt, (E,), "Jmin_out"), _ptr(action_out, dt, (m, E), "action_out", optional=True),
                _ptr(accum, dt, (E,), "accum", optional=True), float(sampling_time), _stream()), "rcg_actor_cost")
    return J_out, argmin_out, Jmin_out


def _opt_workspace(sysd, obj, E, S, device, workspace=None):
    need = int(_C.lib.rcg_actor_opt_workspace_bytes(C.byref(sysd), C.byref(obj), E, S))
    if need < 0:
        raise ValueError("rcg_actor_opt_workspace_bytes: bad arguments")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need // 8, 1),), dtype=_F64, device=device)
    return workspace, need


opt_workspace = _opt_workspace          # public name: (tensor, bytes) of the scratch rcg_actor_opt / rcg_actor_grad need


def actor_grad(sysd, obj, state_sys, obs, sqn, S=1, w_critic=None, w_per_env=False, workspace=None):
    """``_actor_cost`` and its exact gradient for E x S action sequences: ``sqn`` ``[N*m, E*S]`` ->
    ``(J [E*S], grad [N*m, E*S])`` (adjoint of the Euler rollout; the reference's SLSQP uses forward differences)."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = obs.shape[1]
    L = obj.Nactor * m
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    wshape = None if w_critic is None else ((dimc, E) if w_per_env else (dimc,))
    ws, nbytes = _opt_workspace(sysd, obj, E, S, obs.device, workspace)
    J = torch.empty((E * S,), dtype=_F64, device=obs.device)
    g = torch.empty((L, E * S), dtype=_F64, device=obs.device)
    _C.check(_C.lib.rcg_actor_grad(C.byref(sysd), C.byref(obj), E, int(S), _ptr(state_sys, _F64, (n, E), "state_sys"),
                                   _ptr(obs, _F64, (n, E), "obs"), _ptr(sqn, _F64, (L, E * S), "sqn"),
                                   _ptr(w_critic, _F64, wshape, "w_critic", optional=True), int(bool(w_per_env)),
                                   _ptr(ws), nbytes, _ptr(J), _ptr(g), _stream()), "rcg_actor_grad")
    return J, g


def actor_opt(sysd, obj, state_sys, obs, sqn, S=1, w_critic=None, w_per_env=False, mask=None, max_iter=300,
              pg_tol=1e-7, f_tol=1e-12, workspace=None, J_out=None, iters_out=None, nfev_out=None, best_out=None,
              Jmin_out=None, action_out=None, accum=None, sampling_time=0.0, want_stats=True):
    """``CtrlOptPred._actor_optimizer``: bounded minimisation of ``_actor_cost`` from the E x S start points in
    ``sqn`` ``[N*m, E*S]`` (overwritten with the minimisers).  Returns ``(J [E*S], iters [E*S], nfev [E*S])``
    (``None`` each when ``want_stats`` is false and no output tensor is given)."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = obs.shape[1]
    L = obj.Nactor * m
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    wshape = None if w_critic is None else ((dimc, E) if w_per_env else (dimc,))
    ws, nbytes = _opt_workspace(sysd, obj, E, S, obs.device, workspace)
    if want_stats:
        J_out = torch.empty((E * S,), dtype=_F64, device=obs.device) if J_out is None else J_out
        iters_out = torch.zeros((E * S,), dtype=_I32, device=obs.device) if iters_out is None else iters_out
        nfev_out = torch.zeros((E * S,), dtype=_I32, device=obs.device) if nfev_out is None else nfev_out
    _C.check(_C.lib.rcg_actor_opt(C.byref(sysd), C.byref(obj), E, int(S), _ptr(state_sys, _F64, (n, E), "state_sys"),
                                  _ptr(obs, _F64, (n, E), "obs"), _ptr(sqn, _F64, (L, E * S), "sqn"),
                                  _ptr(w_critic, _F64, wshape, "w_critic", optional=True), int(bool(w_per_env)),
                                  _ptr(mask, _I32, (E,), "mask", optional=True), int(max_iter), float(pg_tol),
                                  float(f_tol), _ptr(ws), nbytes,
                                  _ptr(J_out, _F64, (E * S,), "J_out", optional=True),
                                  _ptr(iters_out, _I32, (E * S,), "iters_out", optional=True),
                                  _ptr(nfev_out, _I32, (E * S,), "nfev_out", optional=True),
                                  _ptr(best_out, _I32, (E,), "best_out", optional=True),
                                  _ptr(Jmin_out, _F64, (E,), "Jmin_out", optional=True),
                                  _ptr(action_out, _F64, (m, E), "action_out", optional=True),
                                  _ptr(accum, _F64, (E,), "accum", optional=True), float(sampling_time), _stream()),
             "rcg_actor_opt")
    return J_out, iters_out, nfev_out


def ilqr_workspace_bytes(sysd, obj, E, S=1):
    need = int(_C.lib.rcg_actor_ilqr_workspace_bytes(C.byref(sysd), C.byref(obj), int(E), int(S)))
    if need < 0:
        raise ValueError("rcg_actor_ilqr_workspace_bytes: bad arguments")
    return need


def actor_ilqr(sysd, obj, state_sys, obs, sqn, S=1, w_critic=None, w_per_env=False, mask=None, max_sweeps=25, pg_tol=1e-7,
               workspace=None, sweeps_out=None):
    """Gauss-Newton (iLQR) pre-pass of :func:`actor_opt`: at most ``max_sweeps`` control-limited Riccati sweeps per
    problem on ``sqn`` ``[N*m, E*S]`` in place (the cost never increases).  Returns ``sweeps [E*S]``."""
    n, m = _C.SYS_DIMS[sysd.sys_id]
    E = obs.shape[1]
    L = obj.Nactor * m
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    wshape = None if w_critic is None else ((dimc, E) if w_per_env else (dimc,))
    need = int(_C.lib.rcg_actor_ilqr_workspace_bytes(C.byref(sysd), C.byref(obj), E, int(S)))
    if need < 0:
        raise ValueError("rcg_actor_ilqr_workspace_bytes: bad arguments")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need // 8, 1),), dtype=_F64, device=obs.device)
    if sweeps_out is None:
        sweeps_out = torch.zeros((E * S,), dtype=_I32, device=obs.device)
    _C.check(_C.lib.rcg_actor_ilqr(C.byref(sysd), C.byref(obj), E, int(S), _ptr(state_sys, _F64, (n, E), "state_sys"),
                                   _ptr(obs, _F64, (n, E), "obs"), _ptr(sqn, _F64, (L, E * S), "sqn"),
                                   _ptr(w_critic, _F64, wshape, "w_critic", optional=True), int(bool(w_per_env)),
                                   _ptr(mask, _I32, (E,), "mask", optional=True), int(max_sweeps), float(pg_tol),
                                   _ptr(workspace), need, _ptr(sweeps_out, _I32, (E * S,), "sweeps_out"), _stream()),
             "rcg_actor_ilqr")
    return sweeps_out


def gather_sqn(cand, cand_per_env, C_, idx, out, mask=None):
    """Start points of the optimiser from an arg-min: ``out[:, e]`` = candidate ``idx[e]`` of environment e."""
    L, E = out.shape
    ncol = E * C_ if cand_per_env else C_
    _C.check(_C.lib.rcg_gather_sqn(int(L), E, int(C_), _ptr(cand, _F64, (L, ncol), "cand"), int(bool(cand_per_env)),
                                   _ptr(idx, _I32, (E,), "idx"), _ptr(mask, _I32, (E,), "mask", optional=True),
                                   _ptr(out, _F64, (L, E), "out"), _stream()), "rcg_gather_sqn")
    return out


def nominal_ni(sysd, obs, ctrl_gain, action, mask=None, obj=None, accum=None, sampling_time=0.0):
    """``CtrlNominal3WRobotNI``: writes the nominal parking action into ``action`` [2, E] for the masked lanes
    (+ fused ``upd_accum_obj`` when ``accum`` and ``obj`` are given)."""
    E = obs.shape[1]
    _C.check(_C.lib.rcg_nominal_ni(C.byref(sysd), E, _ptr(obs, _F64, (3, E), "obs"), float(ctrl_gain),
                                   _ptr(mask, _I32, (E,), "mask", optional=True), _ptr(action, _F64, (2, E), "action"),
                                   C.byref(obj) if obj is not None else None,
                                   _ptr(accum, _F64, (E,), "accum", optional=True), float(sampling_time), _stream()),
             "rcg_nominal_ni")
    return action


def stage_obj(obj, n, m, obs, act, out=None, accum=None, scale=0.0, want_out=True):
    """``CtrlOptPred.stage_obj`` (+ fused ``upd_accum_obj`` when ``accum`` is given)."""
    E = obs.shape[1]
    if want_out and out is None:
        out = torch.empty((E,), dtype=_F64, device=obs.device)
    _C.check(_C.lib.rcg_stage_obj(C.byref(obj), n, m, E, _ptr(obs, _F64, (n, E), "obs"), _ptr(act, _F64, (m, E), "act"),
                                  _ptr(out, _F64, (E,), "out", optional=True),
                                  _ptr(accum, _F64, (E,), "accum", optional=True), float(scale), _stream()),
             "rcg_stage_obj")
    return out


def critic(obj, n, m, obs, act, w, w_per_env=False, out=None):
    """``CtrlOptPred._critic``."""
    E = obs.shape[1]
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    out = torch.empty((E,), dtype=_F64, device=obs.device) if out is None else out
    _C.check(_C.lib.rcg_critic(C.byref(obj), n, m, E, _ptr(obs, _F64, (n, E), "obs"), _ptr(act, _F64, (m, E), "act"),
                               _ptr(w, _F64, (dimc, E) if w_per_env else (dimc,), "w"), int(bool(w_per_env)),
                               _ptr(out, _F64, (E,), "out"), _stream()), "rcg_critic")
    return out


def critic_cost(obj, n, m, obs_buf, act_buf, w, w_prev, out=None):
    """``CtrlOptPred._critic_cost`` for W weight vectors per env: ``w`` [dimc, E, W] -> Jc [E, W]."""
    L, _, E = obs_buf.shape
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    W = w.shape[2]
    dt = obs_buf.dtype                                     # fp64, or fp32 (rcg_critic_cost_f32: the tolerance report)
    out = torch.empty((E, W), dtype=dt, device=obs_buf.device) if out is None else out
    fn = getattr(_C.lib, "rcg_critic_cost" + _suffix(dt))
    _C.check(fn(C.byref(obj), n, m, E, W, _ptr(obs_buf, dt, (L, n, E), "obs_buf"), _ptr(act_buf, dt, (L, m, E), "act_buf"),
                _ptr(w, dt, (dimc, E, W), "w"), _ptr(w_prev, dt, (dimc, E), "w_prev"), _ptr(out, dt, (E, W), "out"), _stream()),
             "rcg_critic_cost")
    return out


def critic_fit(obj, n, m, obs_buf, act_buf, w_prev, w_min, w_max, w, w_init=None, mask=None, max_evals=0,
               update_prev=False, Jc_out=None):
    """``CtrlOptPred._critic_optimizer``: bounded least-squares fit of ``_critic_cost`` per environment.
    Starts from ``w_init`` [dimc] (shared) or, if None, from the content of ``w`` [dimc, E]; the fitted weights
    are written to ``w`` (and to ``w_prev`` when ``update_prev``) for the lanes with ``mask`` != 0.
    ``max_evals`` > 0 bounds the work per environment (best iterate so far is returned); 0 = to convergence."""
    L, _, E = obs_buf.shape
    dimc = _C.lib.rcg_dim_critic(obj.critic_struct, n, m)
    _C.check(_C.lib.rcg_critic_fit(C.byref(obj), n, m, E, _ptr(obs_buf, _F64, (L, n, E), "obs_buf"),
                                   _ptr(act_buf, _F64, (L, m, E), "act_buf"), _ptr(w_prev, _F64, (dimc, E), "w_prev"),
                                   float(w_min), float(w_max), _ptr(w_init, _F64, (dimc,), "w_init", optional=True),
                                   _ptr(w, _F64, (dimc, E), "w"), _ptr(mask, _I32, (E,), "mask", optional=True),
                                   int(max_evals), int(bool(update_prev)),
                                   _ptr(Jc_out, _F64, (E,), "Jc_out", optional=True), _stream()), "rcg_critic_fit")
    return w


def ctrl_sample(t, clock, period, in_mask=None, mask_out=None):
    """Clock test of ``CtrlOptPred.compute_action``: returns mask [E] int32, updates ``clock`` in place."""
    E = t.shape[0]
    if mask_out is None:
        mask_out = torch.empty((E,), dtype=_I32, device=t.device)
    _C.check(_C.lib.rcg_ctrl_sample(E, _ptr(t, _F64, (E,), "t"), _ptr(clock, _F64, (E,), "clock"), float(period),
                                    _ptr(in_mask, _I32, (E,), "in_mask", optional=True),
                                    _ptr(mask_out, _I32, (E,), "mask_out"), _stream()), "rcg_ctrl_sample")
    return mask_out


def push_buffers(n, m, obs_buf, act_buf, obs, act, mask=None):
    """``utilities.push_vec`` on the controller FIFO buffers for the masked lanes."""
    L, _, E = obs_buf.shape
    _C.check(_C.lib.rcg_push_buffers(n, m, L, E, _ptr(obs_buf, _F64, (L, n, E), "obs_buf"),
                                     _ptr(act_buf, _F64, (L, m, E), "act_buf"), _ptr(obs, _F64, (n, E), "obs"),
                                     _ptr(act, _F64, (m, E), "act"), _ptr(mask, _I32, (E,), "mask", optional=True),
                                     _stream()), "rcg_push_buffers")
