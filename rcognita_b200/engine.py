"""Batched closed loop: E independent environments stepped by the fused kernels.

This is the device-resident form of the reference's headless main loop
(presets/main_3wrobot_NI.py:415-440) with ``_actor_optimizer`` replaced by
enumerate-and-argmin over candidate action sequences: per control
interval one ``rcg_rk45_advance`` launch (every lane integrates to its own next sampling
event) and one ``rcg_actor_cost`` launch (E x C ``_actor_cost`` evaluations, per-env arg-min,
action hand-over and ``upd_accum_obj``).  All state stays in HBM as ``[component, lane]``
tensors; nothing returns to the host until results are requested.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _C, ops


def _as_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=dtype, device=device)


class ClosedLoopEngine:
    """Closed loop of ``Simulator`` + ``CtrlOptPred`` (MPC, or RQL/SQL with fixed critic weights)
    for ``E`` environments with a candidate/arg-min actor.

    Parameters mirror the reference objects: ``system`` is ``System.name``; ``pars`` /
    ``ctrl_bnds`` as in ``System``; ``dt`` is both ``Simulator.dt`` (``max_step = dt/2``,
    rcognita/simulator.py:150) and ``CtrlOptPred.sampling_time``; ``candidates`` is a shared
    table ``[C, Nactor*m]`` (rows = action sequences like the reference's ``action_sqn``) or a
    per-environment set ``[E, C, Nactor*m]``.

    ``actor`` selects what stands in for ``_actor_optimizer`` (controllers.py:1330-1427): ``"candidates"`` =
    enumerate-and-argmin over the candidate set; ``"opt"`` = the batched bounded minimiser ``rcg_actor_opt``
    (exact adjoint gradients, projected quasi-Newton; at most ``opt_iters`` iterations per sample) started from the
    arg-min candidate (``opt_start="argmin"``) or, like the reference, from ``action_sqn_init`` every time
    (``opt_start="init"``; the candidate set is then unused and may be ``None``), optionally after at most
    ``opt_presweeps`` control-limited Gauss-Newton (iLQR) sweeps (``rcg_actor_ilqr``; 0 = off); ``"nominal"`` = the reference's
    ``CtrlNominal3WRobotNI`` with ``ctrl_gain`` (Sys3WRobotNI only; ``action_init`` defaults to zeros like the
    reference's ``action_curr``, candidates unused).

    ``pars_disturb = [sigma, mu, tau]`` switches the disturbance lanes on (``System(is_disturb=1)``, rcognita/systems.py:228-231,
    :247-248, :325-345, :384-394): the solver integrates ``[state, disturb]`` from ``[state_init, disturb_init]``, the normal
    draws of ``_disturb_dyn`` come from the counter-based stream of environment ``env_offset + e`` under key ``seed``; the
    controller still sees (and predicts with) the undisturbed state rows only, like the reference's ``sys_rhs([], state, action)``.

    ``log_every`` > 0 keeps a device-side trajectory ring for EVERY environment: one row
    (t, state, stage_obj, accum_obj, action -- the rows of rcognita/loggers.py) per ``log_every``-th solver step,
    the last ``log_capacity`` rows per environment (``trajectory(e)``, ``shard.gather_trajectories``).
    """

    def __init__(self, system, state_init, candidates, *, pars=(), ctrl_bnds=None, mode="MPC", Nactor=6, dt=0.01,
                 pred_step_size=None, t0=0.0, t1=10.0, first_step=1e-6, atol=1e-5, rtol=1e-3, gamma=1.0, R1=None,
                 R2=None, stage_obj_struct="quadratic", observation_target=(), critic_struct="quad-nomix",
                 w_critic=None, action_init=(), device=None, dtype=torch.float64, critic_fit=False, Ncritic=4,
                 buffer_size=10, critic_period=None, critic_fit_evals=0, actor="candidates", opt_start="argmin",
                 opt_iters=300, opt_pg_tol=1e-7, opt_f_tol=1e-12, opt_presweeps=0, ctrl_gain=0.5, log_every=0,
                 log_capacity=0, pars_disturb=None, disturb_init=(), seed=0, env_offset=0):
        if not torch.cuda.is_available():
            raise RuntimeError("ClosedLoopEngine needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.sysd = _C.make_system(system, pars, ctrl_bnds)
        self.n, self.m = _C.SYS_DIMS[self.sysd.sys_id]
        n, m = self.n, self.m
        self.sampling_time = float(dt)
        self.t0, self.t1 = float(t0), float(t1)
        self.first_step = float(first_step)
        if first_step <= 0:
            raise ValueError("`first_step` must be positive.")            # scipy common.py:10-16
        if first_step > abs(t1 - t0):
            raise ValueError("`first_step` exceeds bounds.")
        self.sol = _C.make_solver(t1, dt / 2, rtol, atol)
        self.distd = None
        self.nf = n                                                    # rows the solver integrates
        if pars_disturb is not None:
            if dtype != torch.float64 or log_every > 0:
                raise ValueError("disturbance lanes run in fp64 and without the trajectory ring")
            self.nd = _C.DIST_DIM[self.sysd.sys_id]
            self.nf = n + self.nd
            self.distd = _C.make_disturb(pars_disturb, seed=seed, env_offset=env_offset)
        self.obj = _C.make_objective(n, m, mode=mode, Nactor=Nactor,
                                     pred_step_size=dt if pred_step_size is None else pred_step_size, gamma=gamma,
                                     Ncritic=Ncritic, buffer_size=buffer_size, critic_struct=critic_struct, stage_obj_struct=stage_obj_struct, R1=R1, R2=R2,
                                     observation_target=observation_target)
        with torch.cuda.device(self.device):
            x0 = _as_dev(state_init, dtype, self.device)
            if x0.dim() == 1:
                x0 = x0[None, :]
            if x0.shape[1] != n:
                raise ValueError(f"state_init must be [E, {n}]")
            self.E = E = x0.shape[0]
            self.y0 = x0.t().contiguous()                                  # [n, E]
            if self.distd is not None:                                     # state_full_init = [state_init, disturb_init]
                q0 = _as_dev(np.zeros(self.nd) if len(disturb_init) == 0 else disturb_init, dtype, self.device)
                q0 = q0[None, :] if q0.dim() == 1 else q0
                if q0.shape[1] != self.nd or q0.shape[0] not in (1, E):
                    raise ValueError(f"disturb_init must be [{self.nd}] or [E, {self.nd}]")
                self.y0 = torch.cat([self.y0, q0.t().expand(self.nd, E)], dim=0).contiguous()
            L = Nactor * m
            if actor not in ("candidates", "opt", "nominal") or opt_start not in ("argmin", "init"):
                raise ValueError("actor must be 'candidates', 'opt' or 'nominal'; opt_start 'argmin' or 'init'")
            if actor == "nominal" and (system != "3wrobotNI" or dtype != torch.float64):
                raise ValueError("the nominal controller is CtrlNominal3WRobotNI (Sys3WRobotNI, fp64)")
            self.ctrl_gain = float(ctrl_gain)
            self.log_every, self.log_capacity = int(log_every), int(log_capacity)
            if self.log_every > 0 and (self.log_capacity < 1 or dtype != torch.float64):
                raise ValueError("trajectory logging needs log_capacity >= 1 and fp64")
            self.actor, self.opt_start = actor, opt_start
            self.opt_iters, self.opt_pg_tol, self.opt_f_tol = int(opt_iters), float(opt_pg_tol), float(opt_f_tol)
            self.opt_presweeps = int(opt_presweeps)
            self.ilqr_ws = None
            if actor == "opt" and dtype != torch.float64:
                raise ValueError("the actor optimiser runs in fp64")
            if candidates is None:
                if not ((actor == "opt" and opt_start == "init") or actor == "nominal"):
                    raise ValueError("candidates are required unless actor='opt' with opt_start='init' or actor='nominal'")
                candidates = np.zeros((1, L))
            cand = _as_dev(candidates, dtype, self.device)
            if cand.dim() == 2:
                if cand.shape[1] != L:
                    raise ValueError(f"candidates must be [C, {L}]")
                self.C = cand.shape[0]
                self.cand_per_env = False
                self.cand = cand.t().contiguous()                          # [L, C]
            elif cand.dim() == 3:
                if cand.shape[0] != E or cand.shape[2] != L:
                    raise ValueError(f"candidates must be [E, C, {L}]")
                self.C = cand.shape[1]
                self.cand_per_env = True
                self.cand = cand.permute(2, 0, 1).reshape(L, E * self.C).contiguous()   # [L, E*C]
            else:
                raise ValueError("candidates must be [C, N*m] or [E, C, N*m]")
            self.critic_fit = bool(critic_fit) and mode != "MPC"
            self.critic_fit_evals = int(critic_fit_evals)      # 0: fit to convergence; > 0: work bound per environment
            self.critic_period = float(dt if critic_period is None else critic_period)
            self.buffer_size = int(buffer_size)
            self.w_bounds = (-1e3, 1e3) if critic_struct in ("quad-lin", "quad-mix") else (0.0, 1e3)   # controllers.py:1024-1039
            if self.critic_fit:
                # CtrlOptPred with its critic: FIFO buffers, refit every critic_period, w_critic_init = ones
                if dtype != torch.float64:
                    raise ValueError("critic fitting runs in fp64")
                dimc = _C.dim_critic(critic_struct, n, m)
                self.w = torch.ones((dimc, E), dtype=dtype, device=self.device)
                self.w_per_env = True
            elif mode != "MPC":
                if w_critic is None:
                    raise ValueError("w_critic is required for RQL/SQL (or critic_fit=True)")
                w = _as_dev(w_critic, dtype, self.device)
                dimc = _C.dim_critic(critic_struct, n, m)
                if w.dim() == 1:
                    self.w, self.w_per_env = w.contiguous(), False
                else:
                    self.w, self.w_per_env = w.t().contiguous(), True      # [E, dimc] -> [dimc, E]
                assert self.w.shape[0] == dimc
            else:
                self.w, self.w_per_env = None, False
            lo = np.array([self.sysd.lo[k] for k in range(m)])
            a0 = lo / 10 if len(action_init) == 0 else np.asarray(action_init, dtype=np.float64).reshape(m)
            if actor == "nominal" and len(action_init) == 0:
                a0 = np.zeros(m)                                           # CtrlNominal3WRobotNI.action_curr (:1770)
            self.action_init = _as_dev(a0, dtype, self.device)             # controllers.py:973-978
            self._alloc()
            self.reset()

    def _alloc(self):
        n, m, E, dt, dev = self.n, self.m, self.E, self.dtype, self.device
        # All per-lane state lives in ONE device allocation (fields are typed views at 256-byte-aligned offsets) in
        # three segments: what a caller that owns its environments in HOST memory hands in every step (the state and
        # time that `Simulator.get_sim_step_data` returned to it, rcognita/simulator.py:187-195), what it reads back in
        # addition (action, accumulated objective, status, arg-min), and the solver / controller internals that never
        # leave the device (FSAL derivative, step size, state_sys, clocks, counters).  A host-staged step is ONE copy in
        # (segment 1) and ONE copy out (segments 1 + 2).
        nf = self.nf
        spec = [("y_full", (nf, E), dt), ("t", (E,), torch.float64),
                # read back by the host in addition
                ("action", (m, E), dt), ("accum", (E,), dt), ("status", (E,), torch.int32),
                ("sample_flag", (E,), torch.int32), ("argmin", (E,), torch.int32),
                # device-resident internals
                ("f", (nf, E), dt), ("state_sys", (n, E), dt), ("h_abs", (E,), torch.float64),
                ("ctrl_clock", (E,), torch.float64), ("nfev", (E,), torch.int32), ("nsteps", (E,), torch.int32),
                ("nsamples", (E,), torch.int32), ("Jmin", (E,), dt)]
        off, layout = 0, []
        for name, shape, dtype_ in spec:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype_).element_size()
            layout.append((name, shape, dtype_, off, nbytes))
            off = (off + nbytes + 255) // 256 * 256
            if name == "t":
                self._in_bytes = off
            if name == "argmin":
                self._io_bytes = off
        self._blob_bytes = off
        self._layout = layout
        self._blob = torch.empty((off,), dtype=torch.uint8, device=dev)
        for name, shape, dtype_, o, nbytes in layout:
            setattr(self, name, self._blob[o:o + nbytes].view(dtype_).view(shape))
        self.y = self.y_full[:n]                           # the state rows (all of y_full without disturbance lanes)
        if self.critic_fit:
            dimc = self.w.shape[0]
            self.obs_buf = torch.empty((self.buffer_size, n, E), dtype=dt, device=dev)
            self.act_buf = torch.empty((self.buffer_size, m, E), dtype=dt, device=dev)
            self.w_prev = torch.empty((dimc, E), dtype=dt, device=dev)
            self.w_init = torch.ones((dimc,), dtype=dt, device=dev)
            self.critic_clock = torch.empty((E,), dtype=torch.float64, device=dev)
            self.critic_flag = torch.empty((E,), dtype=torch.int32, device=dev)
            self.Jc = torch.empty((E,), dtype=dt, device=dev)
            self.nfits = torch.empty((E,), dtype=torch.int32, device=dev)
        self.log = ops.TrajectoryLog(n, m, E, self.log_capacity, self.log_every, dev) if self.log_every > 0 else None
        if self.actor == "opt":
            L = self.obj.Nactor * m
            self.sqn = torch.zeros((L, E), dtype=dt, device=dev)
            self.sqn_init = self.action_init.repeat(self.obj.Nactor)[:, None].expand(L, E).contiguous()   # rep_mat (:973-978)
            self.opt_ws, _ = ops._opt_workspace(self.sysd, self.obj, E, 1, dev)
            if self.opt_presweeps > 0:
                self.ilqr_ws = torch.empty((max(ops.ilqr_workspace_bytes(self.sysd, self.obj, E, 1) // 8, 1),),
                                           dtype=torch.float64, device=dev)
                self.ilqr_sweeps = torch.zeros((E,), dtype=torch.int32, device=dev)

    def reset(self):
        """Documented intent of ``Simulator.reset`` + ``CtrlOptPred.reset``: restore y0, t0,
        f(t0, y0) with zero action, h_abs = first_step, clocks and accumulators."""
        self.y_full.copy_(self.y0)
        self.state_sys.copy_(self.y0[:self.n])
        self.action.zero_()                                # System.action = zeros (systems.py:134)
        self.t.fill_(self.t0)
        self.h_abs.fill_(self.first_step)
        self.ctrl_clock.fill_(self.t0)
        self.accum.zero_()
        self.status.fill_(_C.RUNNING)
        self.nfev.fill_(1)
        self.nsteps.zero_()
        self.nsamples.zero_()
        self.sample_flag.zero_()
        self.argmin.fill_(-1)
        self.Jmin.fill_(float("nan"))
        if self.critic_fit:
            self.obs_buf.zero_()                           # controllers.py:980-981
            self.act_buf.zero_()
            self.w.fill_(1.0)                              # w_critic_prev = w_critic_init = ones (:1041-1042)
            self.w_prev.fill_(1.0)
            self.critic_clock.fill_(self.t0)
            self.critic_flag.zero_()
            self.Jc.zero_()
            self.nfits.zero_()
        if self.log is not None:
            self.log.reset()
        if self.distd is not None:                                        # RHS call number 0 of every draw stream
            ops.rhs_disturbed(self.sysd, self.distd, self.y_full, self.action, call=None, out=self.f)
        else:
            ops.rhs(self.sysd, self.y, self.action, out=self.f)           # RK45.__init__: f = fun(t0, y0)
        self.intervals = 0
        self._first_done = False
        self.actor_events = None        # set to a list to record (start, end) CUDA events per actor launch

    # -- the very first solver step runs with System.action = 0; only afterwards does the
    #    system receive the controller's initial action (main_3wrobot_NI.py:417-424).
    def _first_step(self):
        if self.distd is not None:
            ops.rk45_step_disturbed(self.sysd, self.distd, self.sol, self.y_full, self.f, self.t, self.h_abs, self.status,
                                    self.action, self.nfev)
        else:
            ops.rk45_step(self.sysd, self.sol, self.y, self.f, self.t, self.h_abs, self.status, self.action, nfev=self.nfev)
        self.nsteps += 1
        # compute_action's clock test (controllers.py:1440-1442): sets sample_flag and moves ctrl_clock where it fires
        ops.ctrl_sample(self.t, self.ctrl_clock, self.sampling_time, mask_out=self.sample_flag)
        flag = self.sample_flag
        self.action.copy_(self.action_init[:, None].expand(self.m, self.E))
        hold = torch.empty((self.E,), dtype=self.dtype, device=self.device)
        if self.dtype == torch.float64:
            ops.stage_obj(self.obj, self.n, self.m, self.y, self.action, out=hold)
        else:
            hold = ops.stage_obj(self.obj, self.n, self.m, self.y.double(), self.action.double()).to(self.dtype)
        self.accum += hold * self.sampling_time * (1 - flag).to(self.dtype)
        if bool(flag.any()):
            self.nsamples += self.sample_flag
            self._actor()                                   # state_sys is still y0 here
        self.state_sys.copy_(self.y)
        if self.log is not None:
            ops.log_rows(self.obj, self.n, self.m, self.log, self.t, self.y, self.action, self.accum, nsteps=self.nsteps)
        self._first_done = True

    def _actor(self):
        ev = None
        if self.actor_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        self._actor_launch()
        if ev is not None:
            ev[1].record()
            self.actor_events.append(ev)

    def _critic_update(self):
        """RQL/SQL part of compute_action for the sampling lanes (controllers.py:1455-1479): push (observation,
        previous action) into the FIFO buffers, test the critic clock, refit the critic where it fired (those
        lanes also remember the new weights as w_critic_prev), fall back to w_critic_prev elsewhere."""
        n, m = self.n, self.m
        ops.push_buffers(n, m, self.obs_buf, self.act_buf, self.y, self.action, mask=self.sample_flag)
        ops.ctrl_sample(self.t, self.critic_clock, self.critic_period, in_mask=self.sample_flag, mask_out=self.critic_flag)
        # every fit starts from w_critic_init = ones (:1264); refit lanes store the result as w_critic and
        # w_critic_prev (:1470-1471).  Sampling lanes whose critic clock did not fire take w_critic_prev (:1479),
        # which already equals their w_critic (both were written by their last refit, or are still ones).
        self._critic_optimizer()
        self.nfits += self.critic_flag

    def _critic_optimizer(self):
        """``CtrlOptPred._critic_optimizer`` (controllers.py:1248-1271) + the ``w_critic_prev = w_critic`` hand-over
        (:1470-1471) for the lanes whose critic clock fired (``critic_flag``): ``rcg_critic_fit`` in place of SLSQP."""
        ops.critic_fit(self.obj, self.n, self.m, self.obs_buf, self.act_buf, self.w_prev, self.w_bounds[0],
                       self.w_bounds[1], self.w, w_init=self.w_init, mask=self.critic_flag,
                       max_evals=self.critic_fit_evals, update_prev=True, Jc_out=self.Jc)

    def _actor_launch(self):
        if self.critic_fit:
            self._critic_update()
        if self.actor == "nominal":
            ops.nominal_ni(self.sysd, self.y, self.ctrl_gain, self.action, mask=self.sample_flag, obj=self.obj,
                           accum=self.accum, sampling_time=self.sampling_time)
            return
        if self.actor == "opt":
            if self.opt_start == "argmin":
                ops.actor_cost(self.sysd, self.obj, self.state_sys, self.y, self.cand, self.cand_per_env, self.C,
                               w_critic=self.w, w_per_env=self.w_per_env, mask=self.sample_flag, want_J=False,
                               argmin_out=self.argmin, Jmin_out=self.Jmin)
                ops.gather_sqn(self.cand, self.cand_per_env, self.C, self.argmin, self.sqn, mask=self.sample_flag)
            else:
                self.sqn.copy_(self.sqn_init)              # my_action_sqn_init (:1383), the same for every sample
            if self.opt_presweeps > 0:                     # Gauss-Newton (iLQR) sweeps first: rcg_actor_ilqr
                ops.actor_ilqr(self.sysd, self.obj, self.state_sys, self.y, self.sqn, S=1, w_critic=self.w,
                               w_per_env=self.w_per_env, mask=self.sample_flag, max_sweeps=self.opt_presweeps,
                               pg_tol=self.opt_pg_tol, workspace=self.ilqr_ws, sweeps_out=self.ilqr_sweeps)
            ops.actor_opt(self.sysd, self.obj, self.state_sys, self.y, self.sqn, S=1, w_critic=self.w,
                          w_per_env=self.w_per_env, mask=self.sample_flag, max_iter=self.opt_iters,
                          pg_tol=self.opt_pg_tol, f_tol=self.opt_f_tol, workspace=self.opt_ws, Jmin_out=self.Jmin,
                          action_out=self.action, accum=self.accum, sampling_time=self.sampling_time, want_stats=False)
            return
        ops.actor_cost(self.sysd, self.obj, self.state_sys, self.y, self.cand, self.cand_per_env, self.C,
                       w_critic=self.w, w_per_env=self.w_per_env, mask=self.sample_flag, want_J=False,
                       argmin_out=self.argmin, Jmin_out=self.Jmin, action_out=self.action, accum=self.accum,
                       sampling_time=self.sampling_time)

    def run_interval(self, max_steps=1 << 30):
        """One control interval for every running lane: advance to the next sampling event,
        evaluate all candidates, pick the arg-min action."""
        self.advance(max_steps)
        self.control()

    def advance(self, max_steps=1 << 30):
        """First half of a control interval: every running lane integrates to its next sampling event
        (``rcg_rk45_advance``: sim_step / receive_sys_state / upd_accum_obj with the held action)."""
        if not self._first_done:
            self._first_step()
        if self.distd is not None:
            ops.rk45_advance_disturbed(self.sysd, self.distd, self.sol, self.obj, self.y_full, self.f, self.t, self.h_abs,
                                       self.status, self.action, self.ctrl_clock, self.sampling_time, max_steps, self.nfev,
                                       state_sys=self.state_sys, accum=self.accum, sample_flag=self.sample_flag,
                                       nsteps=self.nsteps, nsamples=self.nsamples)
            return
        ops.rk45_advance(self.sysd, self.sol, self.obj, self.y, self.f, self.t, self.h_abs, self.status, self.action,
                         self.ctrl_clock, self.sampling_time, max_steps, state_sys=self.state_sys, accum=self.accum,
                         sample_flag=self.sample_flag, nfev=self.nfev, nsteps=self.nsteps, nsamples=self.nsamples,
                         log=self.log)

    def control(self):
        """Second half: ``compute_action`` of the sampling lanes (critic refit, actor, action hand-over)."""
        self._actor()
        if self.log is not None:            # the row of the sampling step: its action is known only now
            ops.log_rows(self.obj, self.n, self.m, self.log, self.t, self.y, self.action, self.accum, nsteps=self.nsteps,
                         mask=self.sample_flag)
        self.intervals += 1

    def all_done(self) -> bool:
        return not bool((self.status == _C.RUNNING).any())

    def run(self, max_intervals=None, check_every=32):
        """Run until every lane is finished/failed (or ``max_intervals``). Returns intervals run."""
        k = 0
        while max_intervals is None or k < max_intervals:
            self.run_interval()
            k += 1
            if k % check_every == 0 and self.all_done():
                break
        if max_intervals is None:
            while not self.all_done():
                self.run_interval()
                k += 1
        return k

    # -- checkpoint / resume: everything a continuation depends on, as host tensors
    _CKPT_EXTRA = ("obs_buf", "act_buf", "w", "w_prev", "critic_clock", "critic_flag", "Jc", "nfits", "sqn")

    def state_dict(self):
        """Snapshot of the run (lane state, counters, critic buffers / weights, optimiser sequences, trajectory ring):
        ``load_state_dict`` on an engine built with the same arguments continues bit-identically."""
        if not self._first_done:
            self._first_step()
        torch.cuda.synchronize(self.device)
        sd = {"blob": self._blob.cpu(), "intervals": self.intervals, "E": self.E, "system": int(self.sysd.sys_id)}
        for name in self._CKPT_EXTRA:
            v = getattr(self, name, None)
            if isinstance(v, torch.Tensor):
                sd[name] = v.cpu()
        if self.log is not None:
            sd["log_rows"], sd["log_count"] = self.log.rows.cpu(), self.log.count.cpu()
        return sd

    def load_state_dict(self, sd):
        if sd["E"] != self.E or sd["system"] != int(self.sysd.sys_id) or sd["blob"].numel() != self._blob.numel():
            raise ValueError("checkpoint was taken from an engine with a different system / batch / layout")
        self._blob.copy_(sd["blob"])
        for name in self._CKPT_EXTRA:
            if name in sd:
                getattr(self, name).copy_(sd[name])
        if self.log is not None and "log_rows" in sd:
            self.log.rows.copy_(sd["log_rows"])
            self.log.count.copy_(sd["log_count"])
        self.intervals = int(sd["intervals"])
        self._first_done = True

    def trajectory(self, e=0):
        """Logged rows of environment ``e`` in the reference's column order (rcognita/loggers.py:41-94), oldest
        kept row first: NI t,x,y,alpha,stage_obj,accum_obj,v,omega; 3wrobot t,x,y,alpha,v,omega,stage_obj,accum_obj,F,M;
        2tank t,h1,h2,p,stage_obj,accum_obj."""
        if self.log is None:
            raise RuntimeError("trajectory logging is off (log_every=0)")
        rows = self.log.env_rows(e)
        if self.sysd.sys_id == _C.SYS_IDS["2tank"]:
            rows = rows[:, [0, 1, 2, 5, 3, 4]]
        return rows

    def results(self):
        """Host copies in the reference's row layout: y [E,n], action [E,m], etc."""
        return {
            "y": self.y.t().contiguous().cpu().numpy(), "t": self.t.cpu().numpy(),
            "action": self.action.t().contiguous().cpu().numpy(), "accum": self.accum.cpu().numpy(),
            "status": self.status.cpu().numpy(), "nfev": self.nfev.cpu().numpy(),
            "nsteps": self.nsteps.cpu().numpy(), "nsamples": self.nsamples.cpu().numpy(),
            "argmin": self.argmin.cpu().numpy(), "Jmin": self.Jmin.cpu().numpy(),
            **({"disturb": self.y_full[self.n:].t().contiguous().cpu().numpy()} if self.distd is not None else {}),
            **({"w_critic": self.w.t().contiguous().cpu().numpy(), "nfits": self.nfits.cpu().numpy(),
                "Jc": self.Jc.cpu().numpy()} if self.critic_fit else {}),
        }


    # -- host-buffer form of one control interval: the caller owns its environments in PINNED HOST memory between
    #    calls (like the reference's Python loop does): it hands in the state and time `get_sim_step_data` gave it and
    #    reads back state, time, action, accumulated objective, status and arg-min; solver internals stay on the device.
    HOST_IN_FIELDS = ("y", "t")                    # ("y" = the state rows of "y_full"; with disturbance lanes the host holds y_full)
    HOST_OUT_FIELDS = ("y", "t", "action", "accum", "status", "sample_flag", "argmin")

    def make_host_state(self):
        """Pinned host mirror of the caller-visible part of the lane state (initialised from the current device
        state): a dict of typed views per field (``HOST_OUT_FIELDS``) plus the raw byte buffer under "_blob"."""
        if not self._first_done:
            self._first_step()
        blob = torch.empty((self._io_bytes,), dtype=torch.uint8).pin_memory()
        blob.copy_(self._blob[:self._io_bytes])
        host = {"_blob": blob}
        for name, shape, dtype_, o, nbytes in self._layout:
            if o + nbytes <= self._io_bytes:
                host[name] = blob[o:o + nbytes].view(dtype_).view(shape)
        host["y"] = host["y_full"][:self.n]
        return host

    def run_interval_host(self, host, sync=True):
        """H2D (y, t: one copy) -> rk45_advance + actor -> D2H (y, t, action, accum, status, sample_flag, argmin: one
        copy), all on the current stream; returns the bytes moved (h2d, d2h).  With ``sync`` the host buffers are valid
        on return."""
        self._blob[:self._in_bytes].copy_(host["_blob"][:self._in_bytes], non_blocking=True)
        self.run_interval()
        host["_blob"].copy_(self._blob[:self._io_bytes], non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return self._in_bytes, self._io_bytes


class _ChunkedLoop:
    """E environments split into ``nchunks`` contiguous blocks, each a ``ClosedLoopEngine`` on its own CUDA stream.
    Environments are independent, so the blocks never synchronise with each other: while one block's HBM-bound actor
    launch streams its candidates, the latency-bound ``rk45_advance`` launch of another block runs beside it."""

    def __init__(self, system, state_init, candidates, nchunks, device=None, align=1, **engine_kwargs):
        x0 = state_init if isinstance(state_init, torch.Tensor) else np.asarray(state_init, dtype=np.float64)
        if x0.ndim == 1:
            x0 = x0[None, :]
        E = x0.shape[0]
        nchunks = max(1, min(int(nchunks), max(1, E // max(1, align))))
        bounds = [((E * i) // nchunks) // align * align for i in range(nchunks)] + [E]
        per_env = candidates is not None and (candidates.ndim == 3 if not isinstance(candidates, torch.Tensor) else candidates.dim() == 3)
        self.bounds = bounds
        self.engines = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            cand = candidates[a:b] if per_env else candidates
            kw = dict(engine_kwargs)
            w = kw.get("w_critic")
            if w is not None and getattr(w, "ndim", 1) == 2:
                kw["w_critic"] = w[a:b]
            self.engines.append(ClosedLoopEngine(system, x0[a:b], cand, device=device, **kw))
        self.device = self.engines[0].device
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        torch.cuda.synchronize(self.device)
        self.E = E
        self.nchunks = len(self.engines)

    def synchronize(self):
        """The calling stream waits for every block (results read afterwards are those of the last enqueued step)."""
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)

    def field(self, name):
        """Device view of one lane field over all blocks (lane axis last), after ``synchronize``."""
        self.synchronize()
        return torch.cat([getattr(eng, name) for eng in self.engines], dim=-1)

    def results(self):
        self.synchronize()
        parts = [eng.results() for eng in self.engines]
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


class PipelinedLoop(_ChunkedLoop):
    """Device-resident closed loop of ``nchunks`` environment blocks on ``nchunks`` streams.  ``step()`` enqueues one
    control interval of every block and returns without synchronising; per block the order is rk45_advance -> actor,
    and block i+1's rk45_advance additionally waits for block i's (``stagger``), so that in the steady state every
    rk45_advance launch (FP64-latency-bound, a fraction of a wave) runs beside another block's actor launch
    (HBM-bound) instead of in front of it.  Results are bit-identical to one ``ClosedLoopEngine`` over the whole batch:
    the arithmetic of an environment does not depend on its block."""

    def __init__(self, system, state_init, candidates, nchunks=2, device=None, stagger=True, align=1024, **engine_kwargs):
        super().__init__(system, state_init, candidates, nchunks, device=device, align=align, **engine_kwargs)
        self.stagger = bool(stagger) and self.nchunks > 1
        self.actor_events = None            # set to a list to record (start, end) CUDA events per actor launch
        # rk45_advance goes to a HIGH-PRIORITY stream per block: its few hundred blocks are then placed as soon as
        # blocks of the running actor launch retire, instead of queueing behind the other block's pending actor blocks
        with torch.cuda.device(self.device):
            self.rk_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in self.engines]
        cur = torch.cuda.current_stream()
        for eng, st in zip(self.engines, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                eng._first_step()
        self.synchronize()

    def step(self):
        prev = None
        for eng, st, rk in zip(self.engines, self.streams, self.rk_streams):
            rk.wait_stream(st)                                  # this block's previous actor launch
            if prev is not None:
                rk.wait_stream(prev)                            # stagger: the previous block's rk45_advance
            with torch.cuda.stream(rk):
                eng.advance()
            if self.stagger:
                prev = rk
            st.wait_stream(rk)
            with torch.cuda.stream(st):
                if self.actor_events is not None:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record(st)
                    eng.control()
                    ev[1].record(st)
                    self.actor_events.append(ev)
                else:
                    eng.control()

    def run(self, intervals):
        for _ in range(int(intervals)):
            self.step()
        self.synchronize()


class HostStagedLoop(_ChunkedLoop):
    """The closed loop for a caller that owns its environments in HOST memory (like the reference's Python loop
    does).  Per block and control interval: ONE host->device copy of what the caller holds (state and time, the values
    ``Simulator.get_sim_step_data`` returned), rk45_advance + actor, ONE device->host copy of what the caller reads
    (state, time, action, accumulated objective, status, arg-min), replayed as one CUDA graph per block.  The solver
    and controller internals (FSAL derivative, step size, ``state_sys``, clocks, counters) and the candidate sets stay
    on the device.

    ``step()`` runs one control interval of every block and returns when all host buffers are valid.
    ``run(K, on_block=None)`` pipelines the blocks: block b's interval k+1 is enqueued as soon as ITS interval-k results
    are on the host (``on_block(b, k, host_views)`` is called at that point -- the caller's chance to read or modify
    its environments), while the other blocks' copies and kernels keep the GPU and both PCIe directions busy."""

    def __init__(self, system, state_init, candidates, nchunks=8, device=None, graph=True, **engine_kwargs):
        super().__init__(system, state_init, candidates, nchunks, device=device, **engine_kwargs)
        with torch.cuda.device(self.device):
            self.hosts = [eng.make_host_state() for eng in self.engines]
            self.done = [torch.cuda.Event() for _ in self.engines]
        torch.cuda.synchronize(self.device)
        self.graphs = None
        self.h2d_bytes = sum(eng._in_bytes for eng in self.engines)
        self.d2h_bytes = sum(eng._io_bytes for eng in self.engines)
        if graph:
            self.capture()

    def capture(self):
        """One CUDA graph per block (copy in, two kernels, copy out): a replay costs one launch instead of four
        Python-driven enqueues, which is what bounds the step once the batch is split finely."""
        graphs = []
        for eng, st, host in zip(self.engines, self.streams, self.hosts):
            with torch.cuda.stream(st):
                eng.run_interval_host(host, sync=False)          # warm-up outside capture (lazy per-kernel attributes)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                eng.run_interval_host(host, sync=False)
            graphs.append(g)
        self.graphs = graphs
        self.warmup_intervals = 1
        return self

    def _enqueue(self, b):
        st = self.streams[b]
        with torch.cuda.stream(st):
            if self.graphs is not None:
                self.graphs[b].replay()
                self.engines[b].intervals += 1
            else:
                self.engines[b].run_interval_host(self.hosts[b], sync=False)
            self.done[b].record(st)

    def step(self):
        for b in range(self.nchunks):
            self._enqueue(b)
        for ev in self.done:
            ev.synchronize()
        return self.h2d_bytes, self.d2h_bytes

    def run(self, intervals, on_block=None):
        for k in range(int(intervals)):
            for b in range(self.nchunks):
                if k > 0:
                    self.done[b].synchronize()                    # block b's previous results are on the host
                    if on_block is not None:
                        on_block(b, k - 1, self.hosts[b])
                self._enqueue(b)
        for b in range(self.nchunks):
            self.done[b].synchronize()
            if on_block is not None:
                on_block(b, int(intervals) - 1, self.hosts[b])
        return self.h2d_bytes * int(intervals), self.d2h_bytes * int(intervals)

    def host_field(self, name):
        """Concatenated host view of one caller-visible lane field over all blocks (lane axis last)."""
        return torch.cat([h[name] for h in self.hosts], dim=-1)
