#!/usr/bin/env python3
"""Kernel-level timings (CUDA events) of the hot-path kernels on one GPU: the actor-cost sweep
(BASELINE.json configs[4]) and the RK45 step/advance kernels under a held action.  Prints one
JSON line per point; used for profiles/ and DESIGN.md, not for the bench contract."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcognita_b200 import _C, ops  # noqa: E402

PRESET = {
    "3wrobotNI": dict(pars=[], bnds=[[-25, 25], [-5, 5]], R1=[1, 10, 1, 0, 0], dt=0.01, psm=1.0, target=[]),
    "3wrobot": dict(pars=[10, 1], bnds=[[-300, 300], [-100, 100]], R1=[1, 10, 1, 0, 0, 0, 0], dt=0.01, psm=2.0, target=[]),
    "2tank": dict(pars=[18.4, 24.4, 1.3, 1, 0.2], bnds=[[0, 1]], R1=[10, 10, 1], dt=0.1, psm=2.0, target=[0.5, 0.5]),
}
BOX = {"3wrobotNI": ([-10, -10, -np.pi], [10, 10, np.pi]), "3wrobot": ([-10, -10, -np.pi, -1, -1], [10, 10, np.pi, 1, 1]),
       "2tank": ([-2, -2], [2, 2])}


WORLD, RANK = 1, 0          # set by main() under torchrun (config-5 sweep at 2/4/8 GPUs: the environments are sharded)


def time_it(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if WORLD > 1:                                   # device time, max over ranks
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def actor_point(system, mode, cs, N, E, C, per_env, dtype=torch.float64, iters=10):
    p = PRESET[system]
    n, m = _C.SYS_DIMS[_C.SYS_IDS[system]]
    sysd = _C.make_system(system, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], critic_struct=cs, R1=p["R1"],
                            observation_target=p["target"])
    g = torch.Generator(device="cuda").manual_seed(0)
    lo, hi = (torch.tensor(v, device="cuda", dtype=torch.float64) for v in BOX[system])
    x = (lo[:, None] + (hi - lo)[:, None] * torch.rand((n, E), device="cuda", dtype=torch.float64, generator=g)).to(dtype)
    b = torch.tensor(p["bnds"], device="cuda", dtype=torch.float64)
    ncol = E * C if per_env else C
    cand = torch.empty((N * m, ncol), device="cuda", dtype=dtype)
    for k in range(N * m):
        j = k % m
        cand[k] = (b[j, 0] + (b[j, 1] - b[j, 0]) * torch.rand((ncol,), device="cuda", dtype=torch.float64, generator=g)).to(dtype)
    dimc = _C.dim_critic(cs, n, m)
    w = None if mode == "MPC" else torch.rand((dimc,), device="cuda", dtype=torch.float64, generator=g).to(dtype)
    am = torch.empty((E,), device="cuda", dtype=torch.int32)
    jm = torch.empty((E,), device="cuda", dtype=dtype)
    act = torch.empty((m, E), device="cuda", dtype=dtype)
    inloop = bool(os.environ.get("RCG_KB_INLOOP"))          # the in-loop call shape: sampling mask + upd_accum_obj epilogue
    mask = torch.ones((E,), device="cuda", dtype=torch.int32) if inloop else None
    accum = torch.zeros((E,), device="cuda", dtype=dtype) if inloop else None
    fn = lambda: ops.actor_cost(sysd, obj, x, x, cand, per_env, C, w_critic=w, want_J=False, argmin_out=am, Jmin_out=jm,
                                action_out=act, mask=mask, accum=accum, sampling_time=0.01)
    ms = time_it(fn, iters)
    evals = E * C
    bytes_eval = (N * m * cand.element_size() + cand.element_size()) if per_env else cand.element_size()
    return {"kernel": "actor_cost", "system": system, "mode": mode, "critic": cs, "Nactor": N, "E": E, "C": C,
            "per_env_cands": bool(per_env), "dtype": str(dtype).split(".")[-1], "ms": ms, "evals_per_s": evals / ms * 1e3,
            "alg_GBps": evals * bytes_eval / ms * 1e-6}


def rk45_point(system, E, steps_per_launch, iters=5):
    p = PRESET[system]
    n, m = _C.SYS_DIMS[_C.SYS_IDS[system]]
    sysd = _C.make_system(system, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode="MPC", Nactor=1, R1=p["R1"], observation_target=p["target"])
    sol = _C.make_solver(1e9, p["dt"] / 2)
    g = torch.Generator(device="cuda").manual_seed(0)
    lo, hi = (torch.tensor(v, device="cuda", dtype=torch.float64) for v in BOX[system])
    y = lo[:, None] + (hi - lo)[:, None] * torch.rand((n, E), device="cuda", dtype=torch.float64, generator=g)
    b = torch.tensor(p["bnds"], device="cuda", dtype=torch.float64)
    action = b[:, :1] + (b[:, 1:] - b[:, :1]) * torch.rand((m, E), device="cuda", dtype=torch.float64, generator=g)
    f = ops.rhs(sysd, y, action)
    t = torch.zeros((E,), device="cuda", dtype=torch.float64)
    h = torch.full((E,), 1e-6, device="cuda", dtype=torch.float64)
    status = torch.zeros((E,), device="cuda", dtype=torch.int32)
    nsteps = torch.zeros((E,), device="cuda", dtype=torch.int32)
    clock = torch.zeros((E,), device="cuda", dtype=torch.float64)
    accum = torch.zeros((E,), device="cuda", dtype=torch.float64)
    ssys = y.clone()
    flag = torch.zeros((E,), device="cuda", dtype=torch.int32)
    if steps_per_launch == 1:
        fn = lambda: ops.rk45_step(sysd, sol, y, f, t, h, status, action)
    else:   # sampling_time huge: every lane takes exactly steps_per_launch accepted steps per launch
        fn = lambda: ops.rk45_advance(sysd, sol, obj, y, f, t, h, status, action, clock, 1e18, steps_per_launch,
                                      state_sys=ssys, accum=accum, sample_flag=flag, nsteps=nsteps)
    for _ in range(8):
        fn()                                                   # reach the max_step regime
    ms = time_it(fn, iters)
    steps = E * steps_per_launch
    bytes_step = ((2 * n + 2 + m) * 8 + (2 * n + 2) * 8)
    return {"kernel": "rk45_step" if steps_per_launch == 1 else "rk45_advance", "system": system, "E": E,
            "steps_per_launch": steps_per_launch, "ms": ms, "env_steps_per_s": steps / ms * 1e3,
            "alg_GBps_if_per_step_io": steps * bytes_step / ms * 1e-6}


def critic_fit_point(system, cs, E, iters=5):
    """rcg_critic_fit on trajectory-like buffers (consecutive samples of a smooth run) for E environments."""
    p = PRESET[system]
    n, m = _C.SYS_DIMS[_C.SYS_IDS[system]]
    obj = _C.make_objective(n, m, mode="RQL", Nactor=4, Ncritic=4, buffer_size=10, critic_struct=cs, R1=p["R1"],
                            observation_target=p["target"])
    g = torch.Generator(device="cuda").manual_seed(0)
    lo, hi = (torch.tensor(v, device="cuda", dtype=torch.float64) for v in BOX[system])
    x = lo[:, None] + (hi - lo)[:, None] * torch.rand((n, E), device="cuda", dtype=torch.float64, generator=g)
    v = 0.05 * torch.randn((n, E), device="cuda", dtype=torch.float64, generator=g)
    obs_buf = torch.stack([x + k * v for k in range(10)])
    b = torch.tensor(p["bnds"], device="cuda", dtype=torch.float64)
    a0 = b[:, :1] + (b[:, 1:] - b[:, :1]) * torch.rand((m, E), device="cuda", dtype=torch.float64, generator=g)
    act_buf = torch.stack([a0 * (1 - 0.02 * k) for k in range(10)])
    dimc = _C.dim_critic(cs, n, m)
    w_prev = torch.ones((dimc, E), device="cuda", dtype=torch.float64)
    w = torch.empty((dimc, E), device="cuda", dtype=torch.float64)
    w_init = torch.ones((dimc,), device="cuda", dtype=torch.float64)
    Jc = torch.empty((E,), device="cuda", dtype=torch.float64)
    lo_w = -1e3 if cs in ("quad-lin", "quad-mix") else 0.0
    fn = lambda: ops.critic_fit(obj, n, m, obs_buf, act_buf, w_prev, lo_w, 1e3, w, w_init=w_init, Jc_out=Jc)
    ms = time_it(fn, iters)
    J0 = ops.critic_cost(obj, n, m, obs_buf, act_buf, w_init[:, None, None].expand(dimc, E, 1).contiguous(), w_prev)[:, 0]
    return {"kernel": "critic_fit", "system": system, "critic": cs, "dim_critic": dimc, "E": E, "ms": ms,
            "fits_per_s": E / ms * 1e3, "median_cost_ratio_fit_over_init": float((Jc / J0.clamp_min(1e-300)).median().item())}


def actor_opt_point(system, mode, cs, N, E, state_scale=1.0, max_iter=300, iters=4, lanes=0):
    """rcg_actor_opt (the batched stand-in for _actor_optimizer) from action_sqn_init for E environments.  States
    are drawn from BOX scaled by `state_scale` around the target (small scale = near the goal = interior minimisers,
    more iterations).  Reports solves/s and gradient (forward + adjoint sweep) evaluations/s."""
    p = PRESET[system]
    n, m = _C.SYS_DIMS[_C.SYS_IDS[system]]
    sysd = _C.make_system(system, p["pars"], p["bnds"])
    obj = _C.make_objective(n, m, mode=mode, Nactor=N, pred_step_size=p["dt"] * p["psm"], critic_struct=cs, R1=p["R1"],
                            observation_target=p["target"])
    g = torch.Generator(device="cuda").manual_seed(0)
    lo, hi = (torch.tensor(v, device="cuda", dtype=torch.float64) for v in BOX[system])
    x = lo[:, None] + (hi - lo)[:, None] * torch.rand((n, E), device="cuda", dtype=torch.float64, generator=g)
    scale = torch.full((n, 1), state_scale, device="cuda", dtype=torch.float64)
    if system != "2tank":
        scale[2] = 1.0                                    # headings stay uniform on the circle
    x = x * scale
    if p["target"]:
        x = x + torch.tensor(p["target"], device="cuda", dtype=torch.float64)[:, None]
    b = np.array(p["bnds"], dtype=np.float64)
    init = torch.as_tensor(np.tile(b[:, 0] / 10, N), device="cuda")[:, None].expand(N * m, E).contiguous()
    dimc = _C.dim_critic(cs, n, m)
    w = None if mode == "MPC" else torch.rand((dimc,), device="cuda", dtype=torch.float64, generator=g)
    ws, _ = ops._opt_workspace(sysd, obj, E, 1, x.device)
    J = torch.empty((E,), device="cuda", dtype=torch.float64)
    it = torch.zeros((E,), device="cuda", dtype=torch.int32)
    nf = torch.zeros((E,), device="cuda", dtype=torch.int32)
    sqn = init.clone()

    import rcognita_b200
    rcognita_b200.actor_opt_lanes(lanes)                  # 0 = default (four lanes per problem where instantiated), 1 = one lane

    def fn():
        sqn.copy_(init)
        ops.actor_opt(sysd, obj, x, x, sqn, w_critic=w, max_iter=max_iter, workspace=ws, J_out=J, iters_out=it, nfev_out=nf)
    ms = time_it(fn, iters, warm=2)
    variant = rcognita_b200.last_actor_opt_kernel()
    rcognita_b200.actor_opt_lanes(0)
    ms_copy = time_it(lambda: sqn.copy_(init), iters, warm=1)
    ms -= ms_copy
    itc, nfc = it.cpu().numpy(), nf.cpu().numpy()
    J0, _ = ops.actor_grad(sysd, obj, x, x, init, w_critic=w)
    return {"kernel": "actor_opt", "variant": variant, "J_mean": float(J.mean().item()), "system": system, "mode": mode, "critic": cs, "Nactor": N, "E": E,
            "state_scale": state_scale, "max_iter": max_iter, "ms": ms, "solves_per_s": E / ms * 1e3,
            "iters_mean": float(itc.mean()), "iters_p50": float(np.median(itc)), "iters_p99": float(np.percentile(itc, 99)),
            "iters_max": int(itc.max()), "linesearch_evals_mean": float(nfc.mean()),
            "grad_evals_per_s": float(itc.sum() + E) / ms * 1e3, "cost_evals_per_s": float(nfc.sum()) / ms * 1e3,
            "median_cost_ratio_opt_over_init": float((J / J0).median().item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="headline", choices=["headline", "sweep", "sweep_multi", "rk45", "rk45ni", "critic", "opt", "all"])
    a = ap.parse_args()
    global WORLD, RANK
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        WORLD, RANK = dist.get_world_size(), dist.get_rank()
    pts = []
    if a.what == "sweep_multi":
        # BASELINE config 5 at 1/2/4/8 GPUs: E TOTAL environments sharded over the ranks (strong scaling of one launch;
        # no communication), a subset of the single-GPU sweep
        for N in (5, 10, 20, 50):
            for C in (64, 256):
                for E in (65536, 1048576, 4194304):
                    El = E // WORLD
                    per_env = El * C * N * 2 * 8 <= 4e9

                    def point(N=N, C=C, E=E, El=El, pe=per_env):
                        r = actor_point("3wrobotNI", "MPC", "quad-nomix", N, El, C, pe, iters=4)
                        r.update(E=E, E_per_gpu=El, n_gpus=WORLD, evals_per_s=E * C / r["ms"] * 1e3,
                                 alg_GBps=r["alg_GBps"] * WORLD, scaling="strong")
                        return r
                    pts.append(point)
    if a.what in ("headline", "all"):
        pts += [lambda: actor_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 256, True),
                lambda: actor_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 256, False),
                lambda: actor_point("3wrobot", "RQL", "quadratic", 10, 262144, 256, False),
                lambda: actor_point("2tank", "SQL", "quad-nomix", 8, 262144, 256, False),
                lambda: actor_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 256, True, torch.float32),
                lambda: actor_point("3wrobotNI", "MPC", "quad-nomix", 7, 65536, 256, True)]
    if a.what in ("sweep", "all"):
        for N in (5, 10, 20, 50):
            for C in (16, 64, 256, 1024):
                for E in (4096, 65536, 1048576, 4194304):
                    per_env = E * C * N * 2 * 8 <= 4e9
                    if E * C > 4.4e9:
                        continue
                    pts.append(lambda N=N, C=C, E=E, pe=per_env: actor_point("3wrobotNI", "MPC", "quad-nomix", N, E, C, pe, iters=5))
    if a.what in ("rk45", "all"):
        for system in ("3wrobotNI", "3wrobot", "2tank"):
            for spl in (1, 16, 256):
                pts.append(lambda s=system, k=spl: rk45_point(s, 1 << 20, k))
    if a.what == "rk45ni":
        pts += [lambda: rk45_point("3wrobotNI", 1 << 20, 1, iters=3), lambda: rk45_point("3wrobotNI", 1 << 20, 16, iters=3)]
    if a.what in ("critic", "all"):
        pts += [lambda: critic_fit_point("2tank", "quad-nomix", 262144), lambda: critic_fit_point("3wrobot", "quadratic", 1 << 20),
                lambda: critic_fit_point("3wrobotNI", "quad-lin", 262144)]
    if a.what in ("opt", "all"):
        for lanes in tuple(int(v) for v in os.environ.get('RCG_BENCH_LANES', '0,1').split(',')):   # 0 = default (four lanes per problem), 1 = one lane
            pts += [lambda lanes=lanes: actor_opt_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 1.0, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 0.05, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("3wrobotNI", "MPC", "quad-nomix", 6, 1 << 20, 0.05, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("3wrobotNI", "MPC", "quad-nomix", 7, 65536, 0.05, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("3wrobot", "RQL", "quadratic", 10, 262144, 1.0, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("3wrobot", "RQL", "quadratic", 10, 262144, 0.05, max_iter=100, lanes=lanes),
                    lambda lanes=lanes: actor_opt_point("2tank", "SQL", "quad-nomix", 8, 262144, 0.5, lanes=lanes)]
    for p in pts:
        r = p()
        if RANK == 0:
            print(json.dumps(r), flush=True)
    if WORLD > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
