"""The presets' argument set and headless main loop (presets/main_3wrobot_NI.py, main_3wrobot.py, main_2tank.py)
on top of the B200 engine.  Flag names, types and defaults are the reference's, including
the ``type=bool`` quirk (any non-empty string is True; pass '' for False).  New flags select the batch:
``--num_envs``, ``--num_candidates``, ``--seed``, ``--state_spread``, ``--candidate_table``.

Differences: ``--is_visualization`` has no effect beyond a notice (the matplotlib animators are outside the hot
path; the loop run is the headless one, presets/main_3wrobot_NI.py:411-462); ``--ctrl_mode nominal`` runs the
batched ``CtrlNominal3WRobotNI`` for Sys3WRobotNI and raises for Sys3WRobot, ``JACS`` raises (out-of-scope
controllers); MPC / RQL / SQL use the candidate/arg-min actor (``--actor candidates``, default) or the batched
bounded minimiser started from the arg-min candidate or from ``action_sqn_init`` (``--actor opt``,
``--opt_start argmin|init``, ``--opt_iters``) and the bounded-least-squares critic fit of
``rcognita_b200.controllers``.  With ``--Nruns`` > 1 every run restarts from ``state_init``
(the documented intent of the reference's reset; its own code raises NameError there).
"""
from __future__ import annotations

import argparse
import ast
import csv
import math
import operator
import os
import pathlib
from datetime import datetime

import numpy as np

# (flag, type, NI default, 3wrobot default, 2tank default); None = same default as the NI preset
_FLAGS = [
    ("--dt", float, 0.01, None, 0.1),
    ("--t1", float, 10.0, None, 100.0),
    ("--Nruns", int, 1, None, None),
    ("--is_log_data", bool, False, None, None),
    ("--is_visualization", bool, True, None, None),
    ("--is_print_sim_step", bool, True, None, None),
    ("--is_est_model", bool, False, None, None),
    ("--model_est_stage", float, 1.0, None, None),
    ("--model_est_period_multiplier", float, 1, None, None),
    ("--model_order", int, 5, None, None),
    ("--prob_noise_pow", float, False, None, None),
    ("--Nactor", int, 3, 5, 10),
    ("--pred_step_size_multiplier", float, 1.0, 2.0, 2.0),
    ("--buffer_size", int, 10, None, None),
    ("--Ncritic", int, 4, None, None),
    ("--gamma", float, 1.0, None, None),
    ("--critic_period_multiplier", float, 1.0, None, None),
]
_LISTS = {   # nargs='+' flags
    "--state_init": (str, ["5", "5", "-3*pi/4"], ["5", "5", "-3*pi/4", "0", "0"], ["2", "-2"]),
    "--action_manual": (float, [-5, -3], [-5, -3], [0.5]),
    "--R1_diag": (float, [1, 10, 1, 0, 0], [1, 10, 1, 0, 0, 0, 0], [10, 10, 1]),
    "--R2_diag": (float, [1, 10, 1, 0, 0], [1, 10, 1, 0, 0, 0, 0], [10, 10, 1]),
}
_COL = {"3wrobotNI": 2, "3wrobot": 3, "2tank": 4}

SYSTEMS = {
    "3wrobotNI": dict(cls="Sys3WRobotNI", logger="Logger3WRobotNI", n=3, m=2, dim_disturb=2, pars=[],
                      ctrl_bnds=[[-25, 25], [-5, 5]], target=[], action_init=[], modes=['manual', 'nominal', 'MPC', 'RQL', 'SQL', 'JACS'],
                      default_mode='nominal', columns=['t [s]', 'x [m]', 'y [m]', 'alpha [rad]', 'stage_obj', 'accum_obj', 'v [m/s]', 'omega [rad/s]']),
    "3wrobot": dict(cls="Sys3WRobot", logger="Logger3WRobot", n=5, m=2, dim_disturb=2, pars=[10, 1],
                    ctrl_bnds=[[-300, 300], [-100, 100]], target=[], action_init=[], modes=['manual', 'nominal', 'MPC', 'RQL', 'SQL', 'JACS'],
                    default_mode='nominal', columns=['t [s]', 'x [m]', 'y [m]', 'alpha [rad]', 'v [m/s]', 'omega [rad/s]', 'stage_obj', 'accum_obj', 'F [N]', 'M [N m]']),
    "2tank": dict(cls="Sys2Tank", logger="Logger2Tank", n=2, m=1, dim_disturb=1, pars=[18.4, 24.4, 1.3, 1, 0.2],
                  ctrl_bnds=[[0, 1]], target=[0.5, 0.5], action_init=[0.5], modes=['manual', 'MPC', 'RQL', 'SQL'],
                  default_mode='MPC', columns=['t [s]', 'h1', 'h2', 'p', 'stage_obj', 'accum_obj']),
}

_OPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
        ast.Pow: operator.pow, ast.USub: operator.neg, ast.UAdd: operator.pos}


def parse_number(expr: str) -> float:
    """'-3*pi/4' -> float.  The reference ``eval``s the string after replacing ``pi``; this accepts the same
    arithmetic (numbers, pi, + - * / **, parentheses) without evaluating arbitrary code."""
    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.Name) and node.id == "pi":
            return math.pi
        if isinstance(node, ast.BinOp) and type(node.op) in _OPS:
            return _OPS[type(node.op)](ev(node.left), ev(node.right))
        if isinstance(node, ast.UnaryOp) and type(node.op) in _OPS:
            return _OPS[type(node.op)](ev(node.operand))
        raise ValueError(f"unsupported expression in state_init: {expr!r}")
    return float(ev(ast.parse(str(expr), mode="eval")))


def make_parser(system: str) -> argparse.ArgumentParser:
    S = SYSTEMS[system]
    col = _COL[system]
    p = argparse.ArgumentParser(description=f"rcognita preset for {system} on the B200 engine (flags as in the reference preset)")
    p.add_argument('--ctrl_mode', metavar='ctrl_mode', type=str, choices=S["modes"], default=S["default_mode"],
                   help='Control mode: manual constant action, nominal (Sys3WRobotNI), MPC, RQL, SQL (JACS: out of scope here).')
    for flag, typ, d_ni, d_3w, d_2t in _FLAGS:
        d = (d_ni, d_3w, d_2t)[col - 2]
        p.add_argument(flag, type=typ, default=d_ni if d is None else d, help=f"as in the reference preset (default %(default)s)")
    for flag, (typ, d_ni, d_3w, d_2t) in _LISTS.items():
        p.add_argument(flag, type=typ, nargs='+', default=list((d_ni, d_3w, d_2t)[col - 2]),
                       help="as in the reference preset (default %(default)s)")
    p.add_argument('--stage_obj_struct', type=str, default='quadratic', choices=['quadratic', 'biquadratic'])
    p.add_argument('--critic_struct', type=str, default='quad-nomix', choices=['quad-lin', 'quadratic', 'quad-nomix', 'quad-mix'])
    p.add_argument('--actor_struct', type=str, default='quad-nomix', choices=['quad-lin', 'quadratic', 'quad-nomix'])
    # --- new: the batch ---
    p.add_argument('--num_envs', type=int, default=1, help='environments stepped in parallel (1 = the reference\'s shapes)')
    p.add_argument('--num_candidates', type=int, default=256, help='candidate action sequences of the arg-min actor')
    p.add_argument('--seed', type=int, default=1, help='seed of the candidate table (and of the initial-state spread)')
    p.add_argument('--candidate_table', type=str, default='random', choices=['random', 'structured'],
                   help='candidate action sequences: uniform random, or constant sequences on a log-spaced grid + random fill')
    p.add_argument('--actor', type=str, default='opt', choices=['candidates', 'opt'],
                   help='stand-in for the SLSQP actor: the batched bounded minimiser (default: the reference\'s semantics), or '
                        'arg-min over a candidate table (the throughput form)')
    p.add_argument('--opt_start', type=str, default='init', choices=['argmin', 'init'],
                   help='start point of --actor opt: action_sqn_init like the reference (default), or the arg-min candidate')
    p.add_argument('--opt_iters', type=int, default=300, help='iteration cap of --actor opt (reference SLSQP: maxiter 300)')
    p.add_argument('--state_spread', type=float, default=0.0,
                   help='std of the Gaussian spread of the initial states around state_init when num_envs > 1')
    return p


def build(system: str, args):
    """Objects of presets/main_3wrobot_NI.py:214-316: (my_sys, my_ctrl, my_simulator, my_logger, derived settings)."""
    from . import controllers, loggers, simulator, systems
    S = SYSTEMS[system]
    n, m = S["n"], S["m"]
    state_init = np.array([parse_number(v) for v in args.state_init], dtype=np.float64)
    if not args.t1 > args.dt > 0.0:
        raise AssertionError("t1 > dt > 0 is required")
    if state_init.size != n:
        raise AssertionError(f"state_init must have {n} entries")
    if args.ctrl_mode == 'JACS' or (args.ctrl_mode == 'nominal' and system != '3wrobotNI'):
        raise NotImplementedError(f"ctrl_mode {args.ctrl_mode!r} needs a controller outside the B200 hot path "
                                  "(CtrlNominal3WRobot, CtrlRLStab); use the reference for it")
    pred_step_size = args.dt * args.pred_step_size_multiplier
    model_est_period = args.dt * args.model_est_period_multiplier
    critic_period = args.dt * args.critic_period_multiplier
    R1, R2 = np.diag(np.array(args.R1_diag, dtype=float)), np.diag(np.array(args.R2_diag, dtype=float))
    ctrl_bnds = np.array(S["ctrl_bnds"], dtype=float)
    if args.num_envs > 1:
        rng = np.random.default_rng(args.seed)
        x0 = state_init[None, :] + args.state_spread * rng.normal(size=(args.num_envs, n))
        x0[0] = state_init
    else:
        x0 = state_init
    my_sys = getattr(systems, S["cls"])(sys_type="diff_eqn", dim_state=n, dim_input=m, dim_output=n, dim_disturb=S["dim_disturb"],
                                        pars=list(S["pars"]), ctrl_bnds=ctrl_bnds, is_dyn_ctrl=0, is_disturb=0, pars_disturb=[])
    mode = args.ctrl_mode if args.ctrl_mode in ('MPC', 'RQL', 'SQL') else 'MPC'
    my_ctrl = controllers.CtrlOptPred(m, n, mode, ctrl_bnds=ctrl_bnds, action_init=S["action_init"], t0=0, sampling_time=args.dt,
                                      Nactor=args.Nactor, pred_step_size=pred_step_size, sys_rhs=my_sys._state_dyn,
                                      sys_out=my_sys.out, state_sys=x0, prob_noise_pow=args.prob_noise_pow,
                                      is_est_model=args.is_est_model, model_est_stage=args.model_est_stage,
                                      model_est_period=model_est_period, buffer_size=args.buffer_size,
                                      model_order=args.model_order, model_est_checks=0, gamma=args.gamma, Ncritic=args.Ncritic,
                                      critic_period=critic_period, critic_struct=args.critic_struct,
                                      stage_obj_struct=args.stage_obj_struct, stage_obj_pars=[R1, R2][:1 if args.stage_obj_struct == 'quadratic' else 2],
                                      observation_target=S["target"], num_candidates=args.num_candidates, seed=args.seed,
                                      candidates=args.candidate_table if (args.actor == 'candidates' or args.opt_start == 'argmin') else None,
                                      actor=args.actor, opt_start=args.opt_start, opt_iters=args.opt_iters)
    my_sim = simulator.Simulator(sys_type="diff_eqn", closed_loop_rhs=my_sys.closed_loop_rhs, sys_out=my_sys.out,
                                 state_init=x0, disturb_init=[], action_init=np.zeros(m), t0=0, t1=args.t1, dt=args.dt,
                                 max_step=args.dt / 2, first_step=1e-6, atol=1e-5, rtol=1e-3, is_disturb=0, is_dyn_ctrl=0)
    my_logger = getattr(loggers, S["logger"])()
    my_ctrl_nominal = None
    if system == '3wrobotNI':          # presets/main_3wrobot_NI.py:235
        my_ctrl_nominal = controllers.CtrlNominal3WRobotNI(ctrl_gain=0.5, ctrl_bnds=ctrl_bnds, t0=0, sampling_time=args.dt)
    return my_sys, my_ctrl, my_sim, my_logger, dict(state_init=state_init, x0=x0, ctrl_nominal=my_ctrl_nominal)


def write_csv_header(datafile, system, args, state_init):
    """The 20 settings rows + the column row of presets/main_3wrobot_NI.py:336-358."""
    with open(datafile, 'w', newline='') as outfile:
        w = csv.writer(outfile)
        w.writerow(['System', system])
        w.writerow(['Controller', args.ctrl_mode])
        w.writerow(['dt', str(args.dt)])
        w.writerow(['state_init', str(state_init)])
        for key in ('is_est_model', 'model_est_stage', 'model_est_period_multiplier', 'model_order', 'prob_noise_pow', 'Nactor',
                    'pred_step_size_multiplier', 'buffer_size', 'stage_obj_struct', 'R1_diag', 'R2_diag', 'Ncritic', 'gamma',
                    'critic_period_multiplier', 'critic_struct', 'actor_struct'):
            w.writerow([key, str(getattr(args, key))])
        w.writerow(SYSTEMS[system]["columns"])


def _first(v):
    """Environment 0 of a batched quantity (or the quantity itself for one environment) as numpy / float."""
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return v


def run_headless(system, args, data_folder=None, quiet=False):
    """The headless loop of presets/main_3wrobot_NI.py:411-462.  Returns a summary dict (per-run final time,
    accumulated objective of every environment, datafiles)."""
    from . import controllers
    my_sys, my_ctrl, my_sim, my_logger, extra = build(system, args)
    batched = args.num_envs > 1
    if data_folder is None:
        data_folder = '../simdata' if os.path.basename(os.path.normpath(os.path.abspath(os.getcwd()))) == 'presets' else 'simdata'
    datafiles = []
    if args.is_log_data:
        pathlib.Path(data_folder).mkdir(parents=True, exist_ok=True)
        stamp = datetime.now().strftime("%Y-%m-%d__%Hh%Mm%Ss")
        for k in range(args.Nruns):
            f = f"{data_folder}/{my_sys.name}__{args.ctrl_mode}__{stamp}__run{k + 1:02d}.csv"
            datafiles.append(f)
            if not quiet:
                print('Logging data to:    ' + f)
            write_csv_header(f, system, args, extra["state_init"])
    if args.is_visualization and not quiet:
        print('[rcognita_b200] visualisation is outside the B200 hot path: running the headless loop')
    action_manual = np.array(args.action_manual, dtype=np.float64)
    if batched and args.ctrl_mode == 'manual':
        action_manual = np.tile(action_manual[None, :], (args.num_envs, 1))
    summary = {"runs": [], "datafiles": datafiles}
    for run in range(args.Nruns):
        if run > 0:
            my_sim.reset()
            my_ctrl.reset(0)
            if extra["ctrl_nominal"] is not None:
                extra["ctrl_nominal"].reset(0)
            my_ctrl._accum.zero_()
        nsteps = 0
        while True:
            if batched:          # lanes that have reached t1 are no longer stepped: freeze their accumulated objective too
                running = my_sim._status == 0
                accum_before = my_ctrl._accum.clone()
            my_sim.sim_step()
            t, state, observation, state_full = my_sim.get_sim_step_data()
            action = controllers.ctrl_selector(t, observation, action_manual, extra["ctrl_nominal"], my_ctrl, args.ctrl_mode)
            my_sys.receive_action(action)
            my_ctrl.receive_sys_state(my_sys._state)
            my_ctrl.upd_accum_obj(observation, action)
            if batched:
                my_ctrl._accum.copy_(my_ctrl._accum.where(running, accum_before))
            nsteps += 1
            if args.is_print_sim_step or args.is_log_data:
                stage_obj = my_ctrl.stage_obj(observation, action)
                accum_obj = my_ctrl.accum_obj_val
                s0 = _first(state_full)[0] if batched else _first(state_full)
                a0 = _first(action)[0] if batched else _first(action)
                row_t = float(_first(t)[0]) if batched else float(t)
                so = float(_first(stage_obj)[0]) if batched else float(stage_obj)
                ao = float(_first(accum_obj)[0]) if batched else float(accum_obj)
                if system == "3wrobotNI":
                    row = (row_t, s0[0], s0[1], s0[2], so, ao, a0)
                elif system == "3wrobot":
                    row = (row_t, s0[0], s0[1], s0[2], s0[3], s0[4], so, ao, a0)
                else:
                    row = (row_t, s0[0], s0[1], a0[0], so, ao)
                if args.is_print_sim_step and not quiet:
                    my_logger.print_sim_step(*row)
                if args.is_log_data:
                    my_logger.log_data_row(datafiles[run], *row)
            t_min = float(_first(t).min()) if batched else float(t)
            if t_min >= args.t1:
                if args.is_print_sim_step and not quiet:
                    print('.....................................Run {run:2d} done.....................................'.format(run=run + 1))
                break
        acc = my_ctrl.accum_obj_val
        summary["runs"].append({"steps": nsteps, "t": t_min, "accum_obj": np.atleast_1d(_first(acc)).astype(float).tolist()})
    return summary


def main(system: str, argv=None):
    args = make_parser(system).parse_args(argv)
    return run_headless(system, args)
