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
                 log_capacity=0):
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
        # All per-lane state lives in ONE device allocation (fields are typed views at 256-byte-aligned offsets):
        # the lane fields first, the per-step results after them, so that a caller that owns its environments in
        # host memory moves the whole state with one copy in (lane part) and one copy out (lane + result part).
        es = torch.empty((), dtype=dt).element_size()
        spec = [("y", (n, E), dt), ("f", (n, E), dt), ("state_sys", (n, E), dt), ("action", (m, E), dt),
                ("t", (E,), torch.float64), ("h_abs", (E,), torch.float64), ("ctrl_clock", (E,), torch.float64),
                ("accum", (E,), dt), ("status", (E,), torch.int32),
                # results
                ("nfev", (E,), torch.int32), ("nsteps", (E,), torch.int32), ("nsamples", (E,), torch.int32),
                ("sample_flag", (E,), torch.int32), ("argmin", (E,), torch.int32), ("Jmin", (E,), dt)]
        off, layout = 0, []
        for name, shape, dtype_ in spec:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype_).element_size()
            layout.append((name, shape, dtype_, off, nbytes))
            off = (off + nbytes + 255) // 256 * 256
            if name == "status":
                self._lane_bytes = off
        self._blob_bytes = off
        self._layout = layout
        self._blob = torch.empty((off,), dtype=torch.uint8, device=dev)
        for name, shape, dtype_, o, nbytes in layout:
            setattr(self, name, self._blob[o:o + nbytes].view(dtype_).view(shape))
        del es
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
        self.y.copy_(self.y0)
        self.state_sys.copy_(self.y0)
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
        ops.rhs(self.sysd, self.y, self.action, out=self.f)               # RK45.__init__: f = fun(t0, y0)
        self.intervals = 0
        self._first_done = False
        self.actor_events = None        # set to a list to record (start, end) CUDA events per actor launch

    # -- the very first solver step runs with System.action = 0; only afterwards does the
    #    system receive the controller's initial action (main_3wrobot_NI.py:417-424).
    def _first_step(self):
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
        ops.critic_fit(self.obj, n, m, self.obs_buf, self.act_buf, self.w_prev, self.w_bounds[0], self.w_bounds[1],
                       self.w, w_init=self.w_init, mask=self.critic_flag, max_evals=self.critic_fit_evals, update_prev=True,
                       Jc_out=self.Jc)
        self.nfits += self.critic_flag

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
        if not self._first_done:
            self._first_step()
        ops.rk45_advance(self.sysd, self.sol, self.obj, self.y, self.f, self.t, self.h_abs, self.status, self.action,
                         self.ctrl_clock, self.sampling_time, max_steps, state_sys=self.state_sys, accum=self.accum,
                         sample_flag=self.sample_flag, nfev=self.nfev, nsteps=self.nsteps, nsamples=self.nsamples,
                         log=self.log)
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
            **({"w_critic": self.w.t().contiguous().cpu().numpy(), "nfits": self.nfits.cpu().numpy(),
                "Jc": self.Jc.cpu().numpy()} if self.critic_fit else {}),
        }


    # -- host-buffer form of one control interval: the lane state lives in PINNED HOST memory between calls (a
    #    caller that owns its environments on the host, like the reference's Python loop does); every call copies
    #    it in, runs the two kernels and copies state + results back.
    LANE_FIELDS = ("y", "f", "state_sys", "action", "t", "h_abs", "ctrl_clock", "accum", "status")
    RESULT_FIELDS = ("nfev", "nsteps", "nsamples", "sample_flag", "argmin", "Jmin")

    def make_host_state(self):
        """Pinned host mirror of the lane-state allocation (initialised from the current device state): a dict of
        typed views per field plus the raw byte buffer under "_blob"."""
        if not self._first_done:
            self._first_step()
        blob = torch.empty((self._blob_bytes,), dtype=torch.uint8).pin_memory()
        blob.copy_(self._blob)
        host = {"_blob": blob}
        for name, shape, dtype_, o, nbytes in self._layout:
            host[name] = blob[o:o + nbytes].view(dtype_).view(shape)
        return host

    def run_interval_host(self, host, sync=True):
        """H2D lane state (one copy) -> rk45_advance + actor_cost -> D2H lane state and results (one copy), all
        on the current stream; returns the bytes moved (h2d, d2h).  With ``sync`` the host buffers are valid on
        return."""
        nl = self._lane_bytes
        self._blob[:nl].copy_(host["_blob"][:nl], non_blocking=True)
        self.run_interval()
        host["_blob"].copy_(self._blob, non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return nl, self._blob_bytes


class HostStagedLoop:
    """The closed loop for a caller that owns its environments in HOST memory (like the reference's Python
    loop does): the batch is split into ``nchunks`` contiguous blocks of environments, each with its own
    ``ClosedLoopEngine`` and CUDA stream, so that the host->device copy of block i+1, the two kernels of block i
    and the device->host copy of block i-1 overlap (PCIe is full duplex, kernels of different blocks may share
    the GPU).  ``step()`` = one control interval of every environment, host buffers valid on return.
    Candidate tables stay resident on the device (they are parameters of the controller, not per-step inputs)."""

    def __init__(self, system, state_init, candidates, nchunks=4, device=None, **engine_kwargs):
        x0 = state_init if isinstance(state_init, torch.Tensor) else np.asarray(state_init, dtype=np.float64)
        E = x0.shape[0]
        nchunks = max(1, min(int(nchunks), E))
        bounds = [(E * i) // nchunks for i in range(nchunks + 1)]
        per_env = candidates.ndim == 3 if not isinstance(candidates, torch.Tensor) else candidates.dim() == 3
        self.engines, self.streams, self.hosts = [], [], []
        for a, b in zip(bounds[:-1], bounds[1:]):
            cand = candidates[a:b] if per_env else candidates
            eng = ClosedLoopEngine(system, x0[a:b], cand, device=device, **engine_kwargs)
            self.engines.append(eng)
        dev = self.engines[0].device
        with torch.cuda.device(dev):
            for eng in self.engines:
                self.streams.append(torch.cuda.Stream(device=dev))
                self.hosts.append(eng.make_host_state())
        torch.cuda.synchronize(dev)
        self.E = E

    def _enqueue(self):
        """Fork every block onto its stream, enqueue copy-in / kernels / copy-out, join back."""
        h2d = d2h = 0
        cur = torch.cuda.current_stream()
        for eng, st, host in zip(self.engines, self.streams, self.hosts):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                a, b = eng.run_interval_host(host, sync=False)
            h2d += a
            d2h += b
        for st in self.streams:
            cur.wait_stream(st)
        return h2d, d2h

    def capture(self):
        """Record one step (all blocks, all streams) into a CUDA graph: a replay costs one launch instead of
        4 x nchunks Python-driven copies and kernel launches, which is what bounds the step once the batch is split
        finely enough for the copies to hide behind the kernels."""
        self._enqueue()                                    # warm-up outside capture (lazy per-kernel attributes)
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._bytes = self._enqueue()
        self._graph = g
        return self

    def step(self):
        if getattr(self, "_graph", None) is not None:
            self._graph.replay()
            torch.cuda.current_stream().synchronize()
            for eng in self.engines:
                eng.intervals += 1
            return self._bytes
        out = self._enqueue()
        torch.cuda.current_stream().synchronize()
        return out

    def host_field(self, name):
        """Concatenated host view of one lane field over all blocks (lane axis last)."""
        return torch.cat([h[name] for h in self.hosts], dim=-1)
