"""Batched drop-in for ``rcognita.controllers`` on the hot path: ``CtrlOptPred`` (MPC / RQL / SQL
predictive agent, rcognita/controllers.py:679-1493) and ``ctrl_selector`` (:40-63).

Constructor signature, attribute names and method names are the reference's.  Every per-agent
vector (``action_curr``, ``state_sys``, clocks, FIFO buffers, critic weights, ``accum_obj_val``)
is held for ``E`` environments in struct-of-arrays CUDA tensors; methods accept the reference's
shapes (``[n]``) or batched rows (``[E, n]``), numpy or CUDA tensors, and answer in kind.

What differs, and why: the reference minimises ``_actor_cost`` with scipy's SLSQP
(serial, finite-difference gradients -- not on the data-parallel path).  Here ``_actor_optimizer``
enumerates a table of candidate action sequences (``candidates=`` keyword, default: ``num_candidates``
sequences drawn uniformly from the action box with ``seed``), evaluates the reference's own
``_actor_cost`` for all E x C pairs in one ``rcg_actor_cost`` launch and takes ``np.argmin`` per
environment (first minimal index wins).  ``_critic_optimizer`` fits the critic weights with
``rcg_critic_fit`` (bounded least squares on ``_critic_cost``, which is quadratic in ``w``) instead
of SLSQP.  All arithmetic runs in ``librcg_b200.so``; there is no CPU path.

Out of scope (raise): ``is_est_model=1`` (needs ``sippy``; ``_estimate_model`` is dead code in the
reference), ``CtrlRLStab`` and the nominal parking controllers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _C, ops
from .systems import System, from_soa, to_soa

_F64 = torch.float64
_I32 = torch.int32


def ctrl_selector(t, observation, action_manual, ctrl_nominal, ctrl_benchmarking, mode):
    """rcognita/controllers.py:40-63: manual -> ``action_manual``; nominal -> the nominal controller;
    anything else -> the benchmarking controller."""
    if mode == 'manual':
        action = action_manual
    elif mode == 'nominal':
        action = ctrl_nominal.compute_action(t, observation)
    else:  # Controller for benchmakring
        action = ctrl_benchmarking.compute_action(t, observation)
    return action


def rep_mat(array, n, m):
    """rcognita/utilities.py:71-76 for the 1-D case the path uses: ``[a0, a1, a0, a1, ...]``."""
    return np.squeeze(np.tile(np.asarray(array, dtype=np.float64), (n, m)))


def structured_candidates(ctrl_bnds, Nactor, num_candidates=256, seed=1):
    """A candidate table ``[num_candidates, Nactor*m]`` for the enumerate-and-argmin actor that is not just noise:
    CONSTANT action sequences on a symmetric log-spaced grid of every input's range (fractions 1, 1/2, 1/5, 1/10, 1/20,
    1/50, 1/100 of each bound, both signs, and the mid-point) -- the minimisers of the presets' objectives are
    bang-bang far from the target and small near it -- filled up with uniform random sequences.  Measured on the
    3wrobot_NI preset: mean closed-loop objective 140.6 against 263.6 for 256 uniform random
    sequences and 136.4 for the batched optimiser, at the cost of the plain arg-min."""
    b = np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2)
    m = b.shape[0]
    mid, half = 0.5 * (b[:, 0] + b[:, 1]), 0.5 * (b[:, 1] - b[:, 0])
    frac = np.array([1, .5, .2, .1, .05, .02, .01])
    lev = np.concatenate([-frac, [0.0], frac[::-1]])
    while len(lev) ** m > num_candidates and len(lev) > 3:            # thin the grid until it fits
        lev = np.concatenate([lev[:len(lev) // 2][::2], [0.0], lev[len(lev) // 2 + 1:][::-1][::2][::-1]])
    grids = np.meshgrid(*[mid[j] + half[j] * lev for j in range(m)], indexing="ij")
    const = np.stack([g.reshape(-1) for g in grids], axis=1)                             # [len(lev)**m, m]
    tab = np.tile(const, (1, int(Nactor)))[:num_candidates]
    rng = np.random.default_rng(seed)
    fill = rng.uniform(np.tile(b[:, 0], Nactor), np.tile(b[:, 1], Nactor), size=(max(0, num_candidates - len(tab)), m * int(Nactor)))
    return np.concatenate([tab, fill], axis=0)


def _owner_system(fn, attr):
    owner = getattr(fn, "__self__", None)
    if not isinstance(owner, System) or getattr(fn, "__func__", None) is not getattr(System, attr):
        raise TypeError(f"`{attr}` must be the bound method of a rcognita_b200 System (the predictor runs inside "
                        "a CUDA kernel and cannot call arbitrary Python; there is no CPU fallback)")
    return owner


class CtrlOptPred:
    """rcognita/controllers.py:679-1493.  New keywords (all optional, after the reference's):
    ``actor='opt'`` (the default when no candidate table is given -- the reference's semantics): the batched bounded
    minimiser ``rcg_actor_opt`` (exact adjoint gradients, projected quasi-Newton, at most ``opt_iters`` iterations -- the
    reference's SLSQP has maxiter 300) inside the controller's own ``ctrl_bnds`` box, started from ``action_sqn_init``
    like the reference (``opt_start='init'``, the default without a table) or from the arg-min candidate
    (``opt_start='argmin'``); ``actor='candidates'`` (the default when ``candidates`` is given): enumerate-and-argmin
    over ``candidates`` ``[C, Nactor*m]`` (shared table), ``[E, C, Nactor*m]`` (per-environment sets), ``'structured'``
    (``structured_candidates``: constant sequences on a log-spaced grid + random fill) or ``'random'``
    (``num_candidates`` uniform random sequences drawn with ``seed``)."""

    def __init__(self, dim_input, dim_output, mode='MPC', ctrl_bnds=[], action_init=[], t0=0, sampling_time=0.1,
                 Nactor=1, pred_step_size=0.1, sys_rhs=[], sys_out=[], state_sys=[], prob_noise_pow=1,
                 is_est_model=0, model_est_stage=1, model_est_period=0.1, buffer_size=20, model_order=3,
                 model_est_checks=0, gamma=1, Ncritic=4, critic_period=0.1, critic_struct='quad-nomix',
                 stage_obj_struct='quadratic', stage_obj_pars=[], observation_target=[],
                 candidates=None, num_candidates=256, seed=1, actor=None, opt_start=None, opt_iters=300,
                 opt_pg_tol=1e-7, opt_f_tol=1e-12):
        if is_est_model:
            raise NotImplementedError("is_est_model=1 is outside the B200 hot path (needs sippy; dead code upstream)")
        if mode not in _C.MODES:
            raise ValueError(f"mode must be one of {sorted(_C.MODES)} (JACS / nominal controllers are out of scope)")
        self._sys = _owner_system(sys_rhs, "_state_dyn")
        if sys_out != [] and sys_out is not None:
            _owner_system(sys_out, "out")
        n, m = self._sys.dim_state, self._sys.dim_input
        if (dim_input, dim_output) != (m, n):
            raise ValueError(f"dim_input/dim_output must be {m}/{n} for system {self._sys.name!r}")
        self.device = self._sys.device
        self.dim_input, self.dim_output, self.mode = dim_input, dim_output, mode
        self.t0 = t0
        self.sampling_time = sampling_time
        self.Nactor, self.pred_step_size = Nactor, pred_step_size
        ctrl_bnds = np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2)
        if ctrl_bnds.shape[0] != m:
            raise ValueError(f"ctrl_bnds must have shape [{m}, 2]")
        self.action_min, self.action_max = ctrl_bnds[:, 0].copy(), ctrl_bnds[:, 1].copy()    # :968-969
        self.action_sqn_min = rep_mat(self.action_min, 1, Nactor).reshape(-1)                 # :970-971
        self.action_sqn_max = rep_mat(self.action_max, 1, Nactor).reshape(-1)
        if len(np.atleast_1d(action_init)) == 0:                                              # :973-978
            self._action_init = self.action_min / 10
        else:
            self._action_init = np.asarray(action_init, dtype=np.float64).reshape(m)
        self.action_sqn_init = rep_mat(self._action_init, 1, Nactor).reshape(-1)
        self.sys_rhs, self.sys_out = sys_rhs, sys_out
        self.is_est_model, self.is_prob_noise, self.prob_noise_pow = is_est_model, 1, prob_noise_pow
        self.model_est_stage, self.model_est_period = model_est_stage, model_est_period
        self.buffer_size, self.model_order, self.model_est_checks = buffer_size, model_order, model_est_checks
        self.gamma = gamma
        self.Ncritic = int(np.min([Ncritic, buffer_size - 1]))                                # :1015
        self.critic_period, self.critic_struct = critic_period, critic_struct
        self.stage_obj_struct, self.stage_obj_pars = stage_obj_struct, stage_obj_pars
        self.observation_target = observation_target
        if critic_struct not in _C.CRITIC_STRUCTS:
            raise ValueError(f"critic_struct must be one of {sorted(_C.CRITIC_STRUCTS)}")
        self.dim_critic = _C.dim_critic(critic_struct, n, m)                                  # :1024-1039
        lo_w = -1e3 if critic_struct in ('quad-lin', 'quad-mix') else 0.0
        self.Wmin, self.Wmax = lo_w * np.ones(self.dim_critic), 1e3 * np.ones(self.dim_critic)
        R1 = stage_obj_pars[0] if len(stage_obj_pars) > 0 else None
        R2 = stage_obj_pars[1] if len(stage_obj_pars) > 1 else None
        if R1 is None:
            raise ValueError("stage_obj_pars must hold at least R1")
        if stage_obj_struct == 'biquadratic' and R2 is None:
            raise IndexError("stage_obj_struct='biquadratic' needs stage_obj_pars = [R1, R2]")   # like the reference
        tgt = np.asarray(observation_target, dtype=np.float64).reshape(-1)
        self._obj = _C.make_objective(n, m, mode=mode, Nactor=Nactor, pred_step_size=pred_step_size, gamma=gamma,
                                      Ncritic=Ncritic, buffer_size=buffer_size, critic_struct=critic_struct,
                                      stage_obj_struct=stage_obj_struct, R1=R1, R2=R2, observation_target=tgt)
        # Defaults follow the reference: without a candidate table the actor MINIMISES _actor_cost from action_sqn_init
        # inside the controller's box (what SLSQP does at controllers.py:1383-1398); handing in a table selects
        # enumerate-and-argmin, the throughput form.
        if actor is None:
            actor = 'candidates' if candidates is not None else 'opt'
        if opt_start is None:
            opt_start = 'argmin' if candidates is not None else 'init'
        if actor not in ('candidates', 'opt') or opt_start not in ('argmin', 'init'):
            raise ValueError("actor must be 'candidates' or 'opt'; opt_start 'argmin' or 'init'")
        # The optimiser's box is the CONTROLLER's (Bounds(action_sqn_min, action_sqn_max), :1384), not the System's
        # clipping range: a descriptor of its own carries it (the predictor uses the unclipped _state_dyn, so only
        # the box differs).  has_bnds is forced: the reference bounds SLSQP even with all-zero bounds.
        self._sysd = _C.make_system(self._sys.name, self._sys.pars, ctrl_bnds)
        self._sysd.has_bnds = 1
        self.actor, self.opt_start = actor, opt_start
        self.opt_iters, self.opt_pg_tol, self.opt_f_tol = int(opt_iters), float(opt_pg_tol), float(opt_f_tol)
        # candidate action sequences of the enumerate-and-argmin actor
        L = Nactor * m
        if isinstance(candidates, str):
            if candidates not in ('structured', 'random'):
                raise ValueError("candidates must be an array, None, 'random' or 'structured'")
            candidates = structured_candidates(ctrl_bnds, Nactor, int(num_candidates), seed) if candidates == 'structured' else None
        if candidates is None:
            candidates = np.random.default_rng(seed).uniform(self.action_sqn_min, self.action_sqn_max,
                                                             size=(int(num_candidates), L))
        cand = candidates if isinstance(candidates, torch.Tensor) else torch.as_tensor(np.asarray(candidates, dtype=np.float64))
        cand = cand.to(device=self.device, dtype=_F64)
        if cand.dim() == 2 and cand.shape[1] == L:
            self.num_candidates, self._cand_per_env = cand.shape[0], False
            self._cand = cand.t().contiguous()                                     # [L, C]
        elif cand.dim() == 3 and cand.shape[2] == L:
            self.num_candidates, self._cand_per_env = cand.shape[1], True
            self._cand = cand.permute(2, 0, 1).reshape(L, -1).contiguous()         # [L, E*C]
            self._cand_E = cand.shape[0]
        else:
            raise ValueError(f"candidates must be [C, {L}] or [E, C, {L}]")
        self._E = 0
        self._batched, self._numpy_io = False, True
        E0 = 1
        if not (isinstance(state_sys, (list, tuple)) and len(state_sys) == 0):
            x, self._batched = to_soa(state_sys, n, self.device, "state_sys")
            self._numpy_io = not isinstance(state_sys, torch.Tensor)
            E0 = x.shape[1]
        self._alloc(E0)
        if E0 and not (isinstance(state_sys, (list, tuple)) and len(state_sys) == 0):
            self._state_sys.copy_(x)

    # ---- storage ------------------------------------------------------------------------------
    def _alloc(self, E):
        n, m, dev = self.dim_output, self.dim_input, self.device
        if self._cand_per_env and E != self._cand_E:
            raise ValueError(f"per-environment candidates were given for {self._cand_E} environments, got {E}")
        self._E = E
        self._state_sys = torch.zeros((n, E), dtype=_F64, device=dev)
        self._action_curr = torch.as_tensor(self._action_init, device=dev)[:, None].expand(m, E).contiguous()
        self._ctrl_clock = torch.full((E,), float(self.t0), dtype=_F64, device=dev)
        self._critic_clock = torch.full((E,), float(self.t0), dtype=_F64, device=dev)
        self._accum = torch.zeros((E,), dtype=_F64, device=dev)
        self._obs_buf = torch.zeros((self.buffer_size, n, E), dtype=_F64, device=dev)        # :980-981
        self._act_buf = torch.zeros((self.buffer_size, m, E), dtype=_F64, device=dev)
        self._w_critic = torch.ones((self.dim_critic, E), dtype=_F64, device=dev)
        self._w_critic_prev = torch.ones((self.dim_critic, E), dtype=_F64, device=dev)       # :1041-1042
        self._w_critic_init = torch.ones((self.dim_critic,), dtype=_F64, device=dev)
        self._mask = torch.zeros((E,), dtype=_I32, device=dev)
        self._cmask = torch.zeros((E,), dtype=_I32, device=dev)
        self._argmin = torch.full((E,), -1, dtype=_I32, device=dev)
        self._Jmin = torch.full((E,), float("nan"), dtype=_F64, device=dev)
        self.num_samples = torch.zeros((E,), dtype=torch.int64, device=dev)
        if self.actor == 'opt':
            L = self.Nactor * m
            self._sqn = torch.zeros((L, E), dtype=_F64, device=dev)
            self._sqn_init = torch.as_tensor(self.action_sqn_init, device=dev)[:, None].expand(L, E).contiguous()
            self._opt_ws, _ = ops._opt_workspace(self._sysd, self._obj, E, 1, dev)
            self.opt_iters_used = torch.zeros((E,), dtype=_I32, device=dev)
            self.opt_nfev_used = torch.zeros((E,), dtype=_I32, device=dev)

    def _ensure(self, E, batched):
        if E != self._E:
            if self._E > 1:
                raise ValueError(f"controller holds {self._E} environments, got a batch of {E}")
            old = self._state_sys
            self._alloc(E)
            self._state_sys.copy_(old.expand(self.dim_output, E))
        self._batched = self._batched or batched

    def _out(self, t):
        return from_soa(t, self._batched, self._numpy_io)

    def _vec(self, t):
        """per-environment scalars [E] -> python float (one environment) / numpy / tensor."""
        if not self._batched:
            return float(t[0].item())
        return t.cpu().numpy() if self._numpy_io else t

    @property
    def num_envs(self):
        return self._E

    # reference attribute names -------------------------------------------------------------------
    action_curr = property(lambda self: self._out(self._action_curr))
    state_sys = property(lambda self: self._out(self._state_sys))
    ctrl_clock = property(lambda self: self._vec(self._ctrl_clock))
    critic_clock = property(lambda self: self._vec(self._critic_clock))
    accum_obj_val = property(lambda self: self._vec(self._accum))
    w_critic = property(lambda self: self._out(self._w_critic))
    w_critic_prev = property(lambda self: self._out(self._w_critic_prev))
    w_critic_init = property(lambda self: self._out(self._w_critic_init[:, None].expand(-1, self._E)))
    action_buffer = property(lambda self: self._buf_out(self._act_buf))
    observation_buffer = property(lambda self: self._buf_out(self._obs_buf))

    def _buf_out(self, b):
        out = b.permute(2, 0, 1) if self._batched else b[:, :, 0]          # [E, L, d] / [L, d]
        return out.cpu().numpy() if self._numpy_io else out

    # ---- reference interface --------------------------------------------------------------------
    def reset(self, t0):
        """controllers.py:1046-1054: rewinds the clock and the current action; learned parameters stay."""
        self._ctrl_clock.fill_(float(t0))
        self._action_curr.copy_(torch.as_tensor(self.action_min / 10, device=self.device)[:, None].expand(-1, self._E))

    def receive_sys_state(self, state):
        """controllers.py:1056-1061."""
        x, batched = to_soa(state, self.dim_output, self.device, "state")
        self._ensure(x.shape[1], batched)
        self._state_sys.copy_(x)

    def stage_obj(self, observation, action):
        """controllers.py:1063-1084 (a.k.a. rcost)."""
        obs, act, batched, like_numpy = self._pair(observation, action)
        r = ops.stage_obj(self._obj, self.dim_output, self.dim_input, obs, act)
        return self._scalar_out(r, batched, like_numpy)

    def upd_accum_obj(self, observation, action):
        """controllers.py:1086-1093: accum_obj_val += stage_obj(observation, action) * sampling_time."""
        obs, act, batched, _ = self._pair(observation, action)
        self._ensure(obs.shape[1], batched)
        ops.stage_obj(self._obj, self.dim_output, self.dim_input, obs, act, accum=self._accum,
                      scale=float(self.sampling_time), want_out=False)

    def _critic(self, observation, action, w_critic):
        """controllers.py:1192-1214: w . phi(observation, action)."""
        obs, act, batched, like_numpy = self._pair(observation, action)
        w, w_per_env = self._weights(w_critic, obs.shape[1])
        q = ops.critic(self._obj, self.dim_output, self.dim_input, obs, act, w, w_per_env=w_per_env)
        return self._scalar_out(q, batched, like_numpy)

    def _critic_cost(self, w_critic):
        """controllers.py:1216-1245 on the controller's own buffers.  ``w_critic``: ``[dimc]`` (the same
        trial vector for every environment), ``[E, dimc]`` or ``[E, W, dimc]`` (W trial vectors each)."""
        like_numpy = not isinstance(w_critic, torch.Tensor)
        w = torch.as_tensor(np.asarray(w_critic, dtype=np.float64)) if like_numpy else w_critic
        w = w.to(device=self.device, dtype=_F64)
        E, dimc = self._E, self.dim_critic
        if w.dim() == 1:
            w3 = w[:, None, None].expand(dimc, E, 1)
        elif w.dim() == 2:
            w3 = w.t()[:, :, None]
        else:
            w3 = w.permute(2, 0, 1)
        Jc = ops.critic_cost(self._obj, self.dim_output, self.dim_input, self._obs_buf, self._act_buf,
                             w3.contiguous(), self._w_critic_prev)
        if w.dim() <= 2:
            Jc = Jc[:, 0]
            if w.dim() == 1 and not self._batched:
                return float(Jc[0].item())
        return Jc.cpu().numpy() if like_numpy else Jc

    def _critic_optimizer(self, mask=None):
        """controllers.py:1248-1271: minimiser of ``_critic_cost`` within [Wmin, Wmax], started from
        ``w_critic_init``.  ``_critic_cost`` is a linear least-squares objective in ``w``; the fit runs
        per environment in ``rcg_critic_fit`` (bounded least squares) instead of SLSQP.  Returns
        ``[dimc, E]`` SoA weights (only the lanes with ``mask`` != 0 are refitted)."""
        w = self._w_critic.clone()
        ops.critic_fit(self._obj, self.dim_output, self.dim_input, self._obs_buf, self._act_buf, self._w_critic_prev,
                       float(self.Wmin[0]), float(self.Wmax[0]), w, w_init=self._w_critic_init, mask=mask)
        return w

    def _actor_cost(self, action_sqn, observation):
        """controllers.py:1273-1328.  ``action_sqn``: ``[N*m]`` or a table ``[C, N*m]`` (-> ``[C]`` /
        ``[E, C]`` costs).  Uses the stored ``state_sys`` and ``w_critic`` like the reference."""
        like_numpy = not isinstance(observation, torch.Tensor)
        obs, batched = to_soa(observation, self.dim_output, self.device, "observation")
        self._ensure(obs.shape[1], batched)
        sq = torch.as_tensor(np.asarray(action_sqn, dtype=np.float64)) if not isinstance(action_sqn, torch.Tensor) else action_sqn
        sq = sq.to(device=self.device, dtype=_F64)
        single = sq.dim() == 1
        tab = (sq[None, :] if single else sq).t().contiguous()                 # [L, C]
        J, _, _ = ops.actor_cost(self._sysd, self._obj, self._state_sys, obs, tab, False, tab.shape[1],
                                 w_critic=self._w_critic if self.mode != 'MPC' else None, w_per_env=True)
        if single:
            J = J[:, 0]
        if not batched:
            J = J[0]
        if like_numpy:
            J = J.cpu().numpy()
            return float(J) if J.ndim == 0 else J
        return J

    def _actor_optimizer(self, observation, mask=None):
        """controllers.py:1330-1427 with SLSQP replaced by enumerate-and-argmin over the candidate table:
        one launch evaluates ``_actor_cost`` for E x C pairs, picks ``np.argmin`` per environment and writes
        the first action of the best sequence into ``action_curr`` for the lanes with ``mask`` != 0."""
        obs, batched = to_soa(observation, self.dim_output, self.device, "observation")
        self._ensure(obs.shape[1], batched)
        w = self._w_critic if self.mode != 'MPC' else None
        if self.actor == 'opt':
            # bounded minimisation of _actor_cost (what SLSQP does in the reference), one thread per environment
            if self.opt_start == 'argmin':
                ops.actor_cost(self._sysd, self._obj, self._state_sys, obs, self._cand, self._cand_per_env,
                               self.num_candidates, w_critic=w, w_per_env=True, mask=mask, want_J=False,
                               argmin_out=self._argmin, Jmin_out=self._Jmin)
                ops.gather_sqn(self._cand, self._cand_per_env, self.num_candidates, self._argmin, self._sqn, mask=mask)
            else:
                self._sqn.copy_(self._sqn_init)                                      # my_action_sqn_init (:1383)
            ops.actor_opt(self._sysd, self._obj, self._state_sys, obs, self._sqn, S=1, w_critic=w, w_per_env=True,
                          mask=mask, max_iter=self.opt_iters, pg_tol=self.opt_pg_tol, f_tol=self.opt_f_tol,
                          workspace=self._opt_ws, Jmin_out=self._Jmin, action_out=self._action_curr,
                          iters_out=self.opt_iters_used, nfev_out=self.opt_nfev_used, want_stats=False)
            return self._out(self._action_curr)
        ops.actor_cost(self._sysd, self._obj, self._state_sys, obs, self._cand, self._cand_per_env,
                       self.num_candidates, w_critic=w, w_per_env=True,
                       mask=mask, want_J=False, argmin_out=self._argmin, Jmin_out=self._Jmin,
                       action_out=self._action_curr)
        return self._out(self._action_curr)

    def compute_action(self, t, observation):
        """controllers.py:1429-1493.  ``t`` is the solver time (scalar, or ``[E]`` when lanes have their own
        clocks).  Lanes whose sampling clock fires get a new action; the others hold ``action_curr``."""
        self._numpy_io = not isinstance(observation, torch.Tensor)
        obs, batched = to_soa(observation, self.dim_output, self.device, "observation")
        self._ensure(obs.shape[1], batched)
        E = self._E
        if isinstance(t, torch.Tensor):
            tt = t.to(device=self.device, dtype=_F64).reshape(-1)
        else:
            tt = torch.as_tensor(np.asarray(t, dtype=np.float64).reshape(-1), device=self.device)
        if tt.numel() == 1 and E > 1:
            tt = tt.expand(E)
        tt = tt.contiguous()
        ops.ctrl_sample(tt, self._ctrl_clock, float(self.sampling_time), mask_out=self._mask)      # :1438-1442
        self.num_samples += self._mask
        if self.mode in ('RQL', 'SQL'):
            ops.push_buffers(self.dim_output, self.dim_input, self._obs_buf, self._act_buf, obs, self._action_curr,
                             mask=self._mask)                                                      # :1463-1464
            ops.ctrl_sample(tt, self._critic_clock, float(self.critic_period), in_mask=self._mask,
                            mask_out=self._cmask)                                                  # :1459-1468
            w_new = self._critic_optimizer(mask=self._cmask)
            refit = self._cmask.bool()[None, :]
            sampled = self._mask.bool()[None, :]
            # :1470-1479: refit lanes take the new weights (and remember them), the other sampling lanes
            # fall back to w_critic_prev; lanes that do not sample keep everything.
            self._w_critic = torch.where(refit, w_new, torch.where(sampled, self._w_critic_prev, self._w_critic))
            self._w_critic_prev = torch.where(refit, w_new, self._w_critic_prev)
        self._actor_optimizer(obs.t() if batched else obs[:, 0], mask=self._mask)                  # :1452 / :1489
        return self._out(self._action_curr)

    # ---- helpers --------------------------------------------------------------------------------
    def _pair(self, observation, action):
        like_numpy = not isinstance(observation, torch.Tensor)
        obs, batched = to_soa(observation, self.dim_output, self.device, "observation")
        act, _ = to_soa(action, self.dim_input, self.device, "action")
        if act.shape[1] != obs.shape[1]:
            act = act.expand(self.dim_input, obs.shape[1]).contiguous()
        return obs, act, batched, like_numpy

    def _weights(self, w, E):
        w = torch.as_tensor(np.asarray(w, dtype=np.float64)) if not isinstance(w, torch.Tensor) else w
        w = w.to(device=self.device, dtype=_F64)
        if w.dim() == 1:
            return w.contiguous(), False
        return w.t().contiguous(), True                                            # [E, dimc] -> [dimc, E]

    @staticmethod
    def _scalar_out(v, batched, like_numpy):
        if not batched:
            return float(v[0].item()) if like_numpy else v[0]
        return v.cpu().numpy() if like_numpy else v


class _OutOfScope:
    _what = ""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{self._what} is outside the B200 hot path (SURVEY.md section 2): use the reference's "
                                  "implementation for this controller")


class CtrlRLStab(_OutOfScope):
    _what = "CtrlRLStab (JACS)"


class CtrlNominal3WRobot(_OutOfScope):
    _what = "CtrlNominal3WRobot"


class CtrlNominal3WRobotNI:
    """rcognita/controllers.py:1758-1956: nominal parking controller of the non-holonomic integrator ("disassembled
    subgradients"), the default ``ctrl_mode`` of presets/main_3wrobot_NI.py.  Closed form per sample; batched like
    ``CtrlOptPred`` (``[3]`` or ``[E, 3]`` observations, numpy or CUDA tensors).  New optional keyword ``device``."""

    def __init__(self, ctrl_gain=10, ctrl_bnds=[], t0=0, sampling_time=0.1, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("CtrlNominal3WRobotNI needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctrl_gain = ctrl_gain
        self.ctrl_bnds = np.asarray(ctrl_bnds, dtype=np.float64).reshape(-1, 2) if len(np.atleast_1d(ctrl_bnds)) else np.zeros((0, 2))
        self.t0 = t0
        self.sampling_time = sampling_time
        self._sysd = _C.make_system("3wrobotNI", [], self.ctrl_bnds)
        self._sysd_unclipped = _C.make_system("3wrobotNI", [], None)
        self._batched, self._numpy_io = False, True
        self._alloc(1)

    def _alloc(self, E):
        self._E = E
        self._ctrl_clock = torch.full((E,), float(self.t0), dtype=_F64, device=self.device)
        self._action_curr = torch.zeros((2, E), dtype=_F64, device=self.device)               # np.zeros(2) (:1770)
        self._mask = torch.zeros((E,), dtype=_I32, device=self.device)

    def _ensure(self, E, batched):
        if E != self._E:
            if self._E > 1:
                raise ValueError(f"controller holds {self._E} environments, got a batch of {E}")
            self._alloc(E)
        self._batched = self._batched or batched

    action_curr = property(lambda self: from_soa(self._action_curr, self._batched, self._numpy_io))
    ctrl_clock = property(lambda self: (float(self._ctrl_clock[0].item()) if not self._batched else
                                        (self._ctrl_clock.cpu().numpy() if self._numpy_io else self._ctrl_clock)))

    def reset(self, t0):
        """:1772-1778."""
        self._ctrl_clock.fill_(float(t0))
        self._action_curr.zero_()

    def compute_action(self, t, observation):
        """:1907-1927: lanes whose clock fires get a new (clipped) action, the others hold ``action_curr``."""
        self._numpy_io = not isinstance(observation, torch.Tensor)
        obs, batched = to_soa(observation, 3, self.device, "observation")
        self._ensure(obs.shape[1], batched)
        E = self._E
        if isinstance(t, torch.Tensor):
            tt = t.to(device=self.device, dtype=_F64).reshape(-1)
        else:
            tt = torch.as_tensor(np.asarray(t, dtype=np.float64).reshape(-1), device=self.device)
        if tt.numel() == 1 and E > 1:
            tt = tt.expand(E)
        ops.ctrl_sample(tt.contiguous(), self._ctrl_clock, float(self.sampling_time), mask_out=self._mask)
        ops.nominal_ni(self._sysd, obs, self.ctrl_gain, self._action_curr, mask=self._mask)
        return from_soa(self._action_curr, self._batched, self._numpy_io)

    def compute_action_vanila(self, observation):
        """:1929-1941: no clock and -- like the reference -- no clipping."""
        self._numpy_io = not isinstance(observation, torch.Tensor)
        obs, batched = to_soa(observation, 3, self.device, "observation")
        self._ensure(obs.shape[1], batched)
        ops.nominal_ni(self._sysd_unclipped, obs, self.ctrl_gain, self._action_curr)
        return from_soa(self._action_curr, self._batched, self._numpy_io)

    def compute_LF(self, observation):
        raise NotImplementedError("compute_LF is a diagnostic outside the B200 hot path")
