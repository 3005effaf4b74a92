"""Batched drop-in for ``rcognita.systems`` (rcognita/systems.py): ``System`` and the three
concrete environments of the hot path, with the reference's constructor signatures, attribute
names and method names.  Every vector the reference holds as a numpy array of shape ``[d]``
is held here for ``E`` environments as a CUDA tensor in struct-of-arrays layout ``[d, E]``
(the storage of record); callers see it in the reference's row layout ``[E, d]`` as a
transposed *view*, or as ``[d]`` when the object was built for a single environment.

All arithmetic runs in ``librcg_b200.so`` (``rcg_state_dyn``, ``rcg_rhs``); there is no CPU path.
``is_disturb=1`` (rcognita/systems.py:228-231, :247-248, :325-345, :384-394) is supported: the full state is
``[state, disturb]`` and the normal draws of ``_disturb_dyn`` come from a per-environment counter-based stream
(``seed`` keyword) instead of numpy's global ``randn()``.  Out of scope: ``is_dyn_ctrl=1`` raises
(no preset enables it and the reference's dyn-ctrl branch is broken).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _C, ops

_F64 = torch.float64


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("rcognita_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def to_soa(x, d, device, name="array"):
    """Reference-layout input ([d], [E, d]; numpy / list / tensor) -> (SoA CUDA tensor [d, E], batched?).
    A ``.T`` view of a contiguous [d, E] CUDA tensor comes back without a copy."""
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=_F64)
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float64), device=device)
    batched = t.dim() == 2
    if t.dim() == 1:
        t = t[None, :]
    if t.dim() != 2 or t.shape[1] != d:
        raise ValueError(f"{name} must have shape [{d}] or [E, {d}], got {tuple(t.shape)}")
    return t.t().contiguous() if not t.t().is_contiguous() else t.t(), batched


def from_soa(t, batched, like_numpy):
    """SoA [d, E] -> the caller's layout: [E, d] view (batched) or [d]; numpy if the caller uses numpy."""
    out = t.t() if batched else t[:, 0]
    return out.cpu().numpy() if like_numpy else out


class System:
    """rcognita/systems.py:17-253.  New optional keyword ``device``; everything else verbatim."""

    name = None

    def __init__(self, sys_type, dim_state, dim_input, dim_output, dim_disturb, pars=[], ctrl_bnds=[],
                 is_dyn_ctrl=0, is_disturb=0, pars_disturb=[], device=None, seed=0, env_offset=0):
        if is_dyn_ctrl:
            raise NotImplementedError("is_dyn_ctrl is outside the B200 hot path (no preset enables it; broken upstream)")
        if self.name not in _C.SYS_IDS:
            raise ValueError(f"System.name {self.name!r} is not one of {sorted(_C.SYS_IDS)}: only the reference's "
                             "three systems have kernels (no CPU fallback for user-defined dynamics)")
        n, m = _C.SYS_DIMS[_C.SYS_IDS[self.name]]
        if (dim_state, dim_input, dim_output) != (n, m, n):
            raise ValueError(f"{self.name}: dimensions must be state={n}, input={m}, output={n}")
        self.sys_type = sys_type
        self.dim_state, self.dim_input, self.dim_output, self.dim_disturb = dim_state, dim_input, dim_output, dim_disturb
        self.pars = pars
        self.ctrl_bnds = np.asarray(ctrl_bnds, dtype=np.float64) if len(ctrl_bnds) else np.zeros((0, 2))
        self.is_dyn_ctrl, self.is_disturb, self.pars_disturb = is_dyn_ctrl, is_disturb, pars_disturb
        self.device = _device(device)
        self._sysd = _C.make_system(self.name, pars, self.ctrl_bnds)
        self._dim_full_state = dim_state
        self._distd = None
        if is_disturb:                                                              # systems.py:139-145
            nd = _C.DIST_DIM[_C.SYS_IDS[self.name]]
            if dim_disturb != nd:
                raise ValueError(f"{self.name}: dim_disturb must be {nd} (the presets' value) for the disturbance kernels")
            self._dim_full_state = dim_state + dim_disturb
            self._distd = _C.make_disturb(pars_disturb, seed=seed, env_offset=env_offset)
            if self.name != "2tank":
                self.sigma_disturb, self.mu_disturb, self.tau_disturb = pars_disturb[0], pars_disturb[1], pars_disturb[2]
            self._ncall = None                      # RHS calls made so far per environment: numbers the random draws
        # SoA storage; sized on first use (E is set by whoever hands us the first batch)
        self._E = 1
        self._batched = False
        self._numpy_io = True
        self._state_soa = torch.zeros((n, 1), dtype=_F64, device=self.device)
        self._action_soa = torch.zeros((m, 1), dtype=_F64, device=self.device)       # systems.py:134

    # ---- storage -----------------------------------------------------------------------------
    def _resize(self, E, batched):
        if E != self._E:
            n, m = self.dim_state, self.dim_input
            self._E = E
            self._state_soa = torch.zeros((n, E), dtype=_F64, device=self.device)
            self._action_soa = torch.zeros((m, E), dtype=_F64, device=self.device)
        self._batched = batched

    @property
    def num_envs(self):
        return self._E

    @property
    def _state(self):
        return from_soa(self._state_soa, self._batched, self._numpy_io)

    @property
    def action(self):
        return from_soa(self._action_soa, self._batched, self._numpy_io)

    # ---- reference interface ------------------------------------------------------------------
    def _state_dyn(self, t, state, action, disturb=[]):
        """``_state_dyn`` (systems.py:308-323, :370-382, :412-419), unclipped; batched over rows."""
        like_numpy = not isinstance(state, torch.Tensor)
        x, batched = to_soa(state, self.dim_state, self.device, "state")
        a, _ = to_soa(action, self.dim_input, self.device, "action")
        if a.shape[1] != x.shape[1]:
            a = a.expand(self.dim_input, x.shape[1]).contiguous()
        if self.is_disturb and not (isinstance(disturb, (list, tuple)) and len(disturb) == 0):    # systems.py:316, :373
            q, _ = to_soa(disturb, self.dim_disturb, self.device, "disturb")
            if q.shape[1] != x.shape[1]:
                q = q.expand(self.dim_disturb, x.shape[1])
            z = torch.zeros((2, x.shape[1]), dtype=_F64, device=self.device)
            full = ops.rhs_disturbed(self._sysd, self._distd, torch.cat([x, q], dim=0).contiguous(), a.clone(), normals=z, clip=False)
            return from_soa(full[: self.dim_state].contiguous(), batched, like_numpy)
        return from_soa(ops.state_dyn(self._sysd, x, a), batched, like_numpy)

    def _disturb_dyn(self, t, disturb, normals=None):
        """``_disturb_dyn`` (systems.py:325-345, :384-394, :421-424): ``-tau * (disturb + sigma * (randn() + mu))`` per
        component.  ``normals`` (``[nd]`` / ``[E, nd]``) stands in for the draws; by default every environment takes the
        next pair of its counter-based stream (the reference: numpy's global ``randn()``)."""
        if not self.is_disturb:
            raise RuntimeError("_disturb_dyn needs is_disturb=1")
        like_numpy = not isinstance(disturb, torch.Tensor)
        q, batched = to_soa(disturb, self.dim_disturb, self.device, "disturb")
        E = q.shape[1]
        if normals is None:
            if self._ncall is None or self._ncall.numel() != E:
                self._ncall = torch.zeros((E,), dtype=torch.int32, device=self.device)
            z = ops.disturb_normals(self._distd, E, self._ncall)
            self._ncall += 1
        else:
            zz, _ = to_soa(normals, self.dim_disturb, self.device, "normals")
            z = torch.zeros((2, E), dtype=_F64, device=self.device)
            z[: self.dim_disturb] = zz.expand(self.dim_disturb, E)
        full = torch.cat([torch.zeros((self.dim_state, E), dtype=_F64, device=self.device), q], dim=0).contiguous()
        a = torch.zeros((self.dim_input, E), dtype=_F64, device=self.device)
        out = ops.rhs_disturbed(self._sysd, self._distd, full, a, normals=z, clip=False)
        return from_soa(out[self.dim_state:].contiguous(), batched, like_numpy)

    def out(self, state, action=[]):
        """systems.py:185-198: the observation is the state itself."""
        return state

    def receive_action(self, action):
        """systems.py:200-211.  The reference keeps the caller's array and later clips it in place;
        here the action is copied into the system's SoA buffer (clipped in place by the next RHS)."""
        a, batched = to_soa(action, self.dim_input, self.device, "action")
        if a.shape[1] != self._E:
            if a.shape[1] == 1:
                a = a.expand(self.dim_input, self._E)
            else:
                self._resize(a.shape[1], batched)
        self._numpy_io = not isinstance(action, torch.Tensor)
        self._action_soa.copy_(a)

    def closed_loop_rhs(self, t, state_full):
        """systems.py:213-253: clips the stored action in place, returns ``_state_dyn`` and tracks ``_state``."""
        like_numpy = not isinstance(state_full, torch.Tensor)
        x, batched = to_soa(state_full, self._dim_full_state, self.device, "state_full")
        if self.is_disturb:
            E = x.shape[1]
            if E != self._E:
                a_old = self._action_soa
                self._resize(E, batched)
                if a_old.shape[1] == 1:
                    self._action_soa.copy_(a_old.expand(self.dim_input, E))
            if self._ncall is None or self._ncall.numel() != E:
                self._ncall = torch.zeros((E,), dtype=torch.int32, device=self.device)
            f = ops.rhs_disturbed(self._sysd, self._distd, x, self._action_soa, call=self._ncall)
            self._ncall += 1
            self._state_soa = x[: self.dim_state]
            self._batched, self._numpy_io = batched, like_numpy
            return from_soa(f, batched, like_numpy)
        if x.shape[1] != self._E:
            a_old = self._action_soa
            self._resize(x.shape[1], batched)
            if a_old.shape[1] == 1:
                self._action_soa.copy_(a_old.expand(self.dim_input, self._E))
        f = ops.rhs(self._sysd, x, self._action_soa)
        self._state_soa = x
        self._batched, self._numpy_io = batched, like_numpy
        return from_soa(f, batched, like_numpy)


class Sys3WRobot(System):
    """rcognita/systems.py:255-351: three-wheel robot with mass and inertia, pars = [m, I]."""
    name = "3wrobot"


class Sys3WRobotNI(System):
    """rcognita/systems.py:353-399: kinematic three-wheel robot (non-holonomic integrator)."""
    name = "3wrobotNI"


class Sys2Tank(System):
    """rcognita/systems.py:401-428: two-tank system, pars = [tau1, tau2, K1, K2, K3]."""
    name = "2tank"
