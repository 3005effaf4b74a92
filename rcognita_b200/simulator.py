"""Batched drop-in for ``rcognita.simulator.Simulator`` (rcognita/simulator.py:71-204).

``Simulator(sys_type, closed_loop_rhs, sys_out, state_init, ...)`` keeps the reference's
signature; ``state_init`` may be ``[n]`` (one environment, like the reference) or ``[E, n]``.
``sim_step()`` is one launch of ``rcg_rk45_step``: exactly one ACCEPTED scipy-RK45 step per
running environment (per-lane step-size control, FSAL derivative carried across action changes
like scipy's).  ``closed_loop_rhs`` must be the bound method of one of the
package's ``System`` subclasses: a kernel cannot call back into Python per lane, so the owning
system is recovered from ``closed_loop_rhs.__self__`` and anything else is rejected loudly.

Inputs given as numpy arrays come back as numpy (host copies each step); CUDA tensors come back
as CUDA tensors (views of the solver state, invalidated by the next ``sim_step``).
"""
from __future__ import annotations

import torch

from . import _C, ops
from .systems import System, from_soa, to_soa

_F64 = torch.float64


class _SolverView:
    """What the reference's loop reads from ``ODE_solver`` (a ``scipy.integrate.RK45``): ``t``, ``y``,
    ``f``, ``h_abs``, ``nfev``, ``status`` -- per environment."""

    def __init__(self, sim):
        self._sim = sim

    def _scalar(self, v):
        return v.item() if not self._sim._batched else (v.cpu().numpy() if self._sim._numpy_io else v)

    t = property(lambda self: self._scalar(self._sim._t[0] if not self._sim._batched else self._sim._t))
    h_abs = property(lambda self: self._scalar(self._sim._h[0] if not self._sim._batched else self._sim._h))
    nfev = property(lambda self: self._scalar(self._sim._nfev[0] if not self._sim._batched else self._sim._nfev))
    y = property(lambda self: from_soa(self._sim._y, self._sim._batched, self._sim._numpy_io))
    f = property(lambda self: from_soa(self._sim._f, self._sim._batched, self._sim._numpy_io))

    @property
    def status(self):
        st = self._sim._status.cpu().numpy()
        names = [_C.STATUS_NAMES[int(s)] for s in st]
        return names if self._sim._batched else names[0]


class Simulator:
    def __init__(self, sys_type, closed_loop_rhs, sys_out, state_init, disturb_init=[], action_init=[], t0=0, t1=1,
                 dt=1e-2, max_step=0.5e-2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0, is_dyn_ctrl=0):
        if sys_type != "diff_eqn":
            # simulator.py:170-185: discr_fnc / discr_prob are outside the hot path; anything else is invalid
            raise ValueError("Invalid system description" if sys_type not in ("discr_fnc", "discr_prob")
                             else f"sys_type {sys_type!r} is outside the B200 hot path (only 'diff_eqn')")
        if is_dyn_ctrl:
            raise NotImplementedError("is_dyn_ctrl is outside the B200 hot path")
        owner = getattr(closed_loop_rhs, "__self__", None)
        if not isinstance(owner, System) or closed_loop_rhs.__func__ is not System.closed_loop_rhs:
            raise TypeError("closed_loop_rhs must be the bound `closed_loop_rhs` of a rcognita_b200 System "
                            "(kernels cannot call arbitrary Python per lane; there is no CPU fallback)")
        self.sys = owner
        self.sys_type, self.closed_loop_rhs, self.sys_out, self.dt = sys_type, closed_loop_rhs, sys_out, dt
        if first_step <= 0:
            raise ValueError("`first_step` must be positive.")                   # scipy common.py:10-16
        if first_step > abs(t1 - t0):
            raise ValueError("`first_step` exceeds bounds.")
        if bool(is_disturb) != bool(owner.is_disturb):
            raise ValueError("Simulator(is_disturb=...) must agree with the System's is_disturb")
        self.is_disturb = int(bool(is_disturb))
        self._numpy_io = not isinstance(state_init, torch.Tensor)
        ns = owner.dim_state
        y0, self._batched = to_soa(state_init, ns, owner.device, "state_init")
        if self.is_disturb:                                  # simulator.py:100-101: state_full_init = [state_init, disturb_init]
            q0, _ = to_soa(disturb_init, owner.dim_disturb, owner.device, "disturb_init")
            y0 = torch.cat([y0, q0.expand(owner.dim_disturb, y0.shape[1])], dim=0).contiguous()
        n = y0.shape[0]                                       # rows the solver integrates
        self._y0 = y0.clone()
        self.E = E = y0.shape[1]
        self.dim_state = ns
        self.t0, self.t1, self.first_step = float(t0), float(t1), float(first_step)
        # simulator.py:150: max_step = dt/2 -- the constructor's own max_step argument is ignored there too
        self._sol = _C.make_solver(t1, dt / 2, rtol, atol)
        dev = owner.device
        self._y = torch.empty((n, E), dtype=_F64, device=dev)
        self._f = torch.empty((n, E), dtype=_F64, device=dev)
        self._t = torch.empty((E,), dtype=_F64, device=dev)
        self._h = torch.empty((E,), dtype=_F64, device=dev)
        self._status = torch.empty((E,), dtype=torch.int32, device=dev)
        self._nfev = torch.empty((E,), dtype=torch.int32, device=dev)
        self.state_full_init = state_init if not self.is_disturb else from_soa(self._y0, self._batched, self._numpy_io)
        self.ODE_solver = _SolverView(self)
        self._construct_solver()

    def _construct_solver(self):
        """RK45.__init__ (scipy rk.py:85-103): y = y0, t = t0, h_abs = first_step and
        f = fun(t0, y0) evaluated with the system's CURRENT action (zeros at construction)."""
        self._y.copy_(self._y0)
        self._t.fill_(self.t0)
        self._h.fill_(self.first_step)
        self._status.fill_(_C.RUNNING)
        self._nfev.fill_(1)
        sysobj = self.sys
        if sysobj.num_envs != self.E:
            a_old = sysobj._action_soa
            sysobj._resize(self.E, self._batched)
            if a_old.shape[1] == 1:
                sysobj._action_soa.copy_(a_old.expand(sysobj.dim_input, self.E))
        sysobj._batched, sysobj._numpy_io = self._batched, self._numpy_io
        if self.is_disturb:                                   # RHS call number 0 of every environment's draw stream
            ops.rhs_disturbed(sysobj._sysd, sysobj._distd, self._y, sysobj._action_soa, call=None, out=self._f)
            sysobj._state_soa = self._y[: self.dim_state]
        else:
            ops.rhs(sysobj._sysd, self._y, sysobj._action_soa, out=self._f)
            sysobj._state_soa = self._y

    # ---- reference interface ------------------------------------------------------------------
    def sim_step(self):
        """simulator.py:156-168 -> scipy RK45.step(): one accepted step per running environment.
        Raises like scipy (base.py:189-191) when no environment is running any more."""
        if self.E == 1 and int(self._status[0].item()) != _C.RUNNING:
            raise RuntimeError("Attempt to step on a failed or finished solver.")
        if self.is_disturb:
            ops.rk45_step_disturbed(self.sys._sysd, self.sys._distd, self._sol, self._y, self._f, self._t, self._h, self._status,
                                    self.sys._action_soa, self._nfev)
            self.sys._state_soa = self._y[: self.dim_state]
            return
        ops.rk45_step(self.sys._sysd, self._sol, self._y, self._f, self._t, self._h, self._status,
                      self.sys._action_soa, nfev=self._nfev)
        self.sys._state_soa = self._y                                           # systems.py:251

    @property
    def t(self):
        return self.ODE_solver.t

    @property
    def state_full(self):
        return from_soa(self._y, self._batched, self._numpy_io)

    @property
    def state(self):
        """simulator.py:167: ``state_full[0:dim_state]``."""
        if not self.is_disturb:
            return self.state_full
        return from_soa(self._y[: self.dim_state], self._batched, self._numpy_io)

    @property
    def observation(self):
        return self.sys_out(self.state)

    def get_sim_step_data(self):
        """simulator.py:187-195 -> (t, state, observation, state_full)."""
        state = self.state
        return self.t, state, self.sys_out(state), (state if not self.is_disturb else self.state_full)

    def reset(self, literal=False, mask=None):
        """Documented intent of ``Simulator.reset`` (multi-episode runs): restore state_full_init, t0,
        f(t0, y0) and h_abs = first_step.  ``literal=True`` reproduces what the reference's code
        actually does (simulator.py:197-201): only the clock and the status are rewound, the state,
        FSAL derivative and step size carry over.
        ``mask`` ([E] bool / int, numpy or tensor): re-initialise only those environments (a batch whose episodes end at
        different times restarts the finished lanes and leaves the others untouched); the restarted lanes get
        f = fun(t0, y0) with the system's current action of that lane, like a freshly constructed RK45."""
        if mask is None:
            if literal:
                self._t.fill_(self.t0)
                self._status.fill_(_C.RUNNING)
            else:
                self._construct_solver()
            return
        sel = torch.as_tensor(mask, device=self._y.device).reshape(-1) != 0
        if sel.numel() != self.E:
            raise ValueError(f"mask must have {self.E} entries")
        self._t.masked_fill_(sel, self.t0)
        self._status.masked_fill_(sel, _C.RUNNING)
        if literal:
            return
        if self.is_disturb:
            raise NotImplementedError("masked non-literal reset of a disturbed system")
        self._y.copy_(torch.where(sel[None, :], self._y0, self._y))
        self._h.masked_fill_(sel, self.first_step)
        self._nfev.masked_fill_(sel, 1)
        act = self.sys._action_soa.clone()                 # rcg_rhs clips in place: only the restarted lanes may be touched
        f0 = ops.rhs(self.sys._sysd, self._y, act)
        self._f.copy_(torch.where(sel[None, :], f0, self._f))
        self.sys._action_soa.copy_(torch.where(sel[None, :], act, self.sys._action_soa))
