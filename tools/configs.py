#!/usr/bin/env python3
"""BASELINE.json configs 3 and 4 on one GPU (closed loops with the critic in the loop) and the fp64-vs-fp32
tolerance report of config 4.  One JSON line per result; used for profiles/ and DESIGN.md, not for the bench
contract.

    python tools/configs.py config3 [--envs 1048576] [--t1 0.3]     # 3wrobot RQL, 'quadratic' critic, Nactor=10
    python tools/configs.py config4 [--envs 262144]  [--t1 10]      # 2tank SQL, critic buffer fitting, Nactor=8
    python tools/configs.py fp32    [--envs 262144]  [--t1 10]      # config 4: fp64 vs fp32 tolerance report
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_workload import synthetic_candidates, synthetic_states  # noqa: E402
from rcognita_b200 import _C, ops  # noqa: E402
from rcognita_b200.engine import ClosedLoopEngine  # noqa: E402

CFG = {
    "config3": dict(system="3wrobot", mode="RQL", cs="quadratic", N=10, dt=0.01, psm=2.0, pars=[10, 1],
                    bnds=[[-300, 300], [-100, 100]], R1=[1, 10, 1, 0, 0, 0, 0], target=[], a_init=[]),
    "config4": dict(system="2tank", mode="SQL", cs="quad-nomix", N=8, dt=0.1, psm=2.0, pars=[18.4, 24.4, 1.3, 1, 0.2],
                    bnds=[[0, 1]], R1=[10, 10, 1], target=[0.5, 0.5], a_init=[0.5]),
}


def dist_info():
    """(rank, world) under torchrun (one process per GPU, NCCL), else (0, 1)."""
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl")
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def timed_closed_loop(c, E, C, t1, critic_fit=True, dtype=torch.float64, w_critic=None, fit_evals=0, lo=0, hi=None):
    E = (E + 1023) // 1024 * 1024
    if hi is not None:                         # this rank's contiguous block of the global batch (shard.shard_range)
        x0 = synthetic_states(c["system"], lo, hi, seed=0)
        E = hi - lo
    else:
        x0 = synthetic_states(c["system"], 0, E, seed=0)
    cand = synthetic_candidates(c["bnds"], c["N"], C, seed=1)
    eng = ClosedLoopEngine(c["system"], x0, cand, pars=c["pars"], ctrl_bnds=c["bnds"], mode=c["mode"], Nactor=c["N"],
                           dt=c["dt"], pred_step_size=c["dt"] * c["psm"], t1=t1, R1=c["R1"], observation_target=c["target"],
                           critic_struct=c["cs"], critic_fit=critic_fit, w_critic=w_critic, Ncritic=4, buffer_size=10,
                           action_init=c["a_init"], dtype=dtype, critic_fit_evals=fit_evals)
    for _ in range(3):
        eng.run_interval()
    torch.cuda.synchronize()
    phases = {}
    if os.environ.get("RCG_PHASES"):
        # per-phase CUDA-event timing (adds events between the launches; the totals below then include that overhead)
        def wrap(obj, name, key):
            fn = getattr(obj, name)

            def timed(*a, **k):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **k)
                e1.record()
                phases.setdefault(key, []).append((e0, e1))
                return r
            setattr(obj, name, timed)
        import rcognita_b200.engine as engmod
        for nm in ("rk45_advance", "push_buffers", "ctrl_sample", "critic_fit", "actor_cost"):
            wrap(engmod.ops, nm, nm)
    s0, n0 = int(eng.nsteps.sum().item()), int(eng.nsamples.sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = eng.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps, samples = int(eng.nsteps.sum().item()) - s0, int(eng.nsamples.sum().item()) - n0
    if phases:
        print(json.dumps({"phases_ms_per_interval": {k: sum(a.elapsed_time(b) for a, b in v) / max(k_, 1)
                                                      for k, v in phases.items() for k_ in [len(v)]}}), flush=True)
    return eng, dict(E=E, C=C, intervals=k, ms=ms, env_steps_per_s=steps / ms * 1e3, actor_evals_per_s=samples * C / ms * 1e3,
                     critic_fits_per_s=(samples / ms * 1e3) if critic_fit else 0.0, ms_per_interval=ms / max(k, 1))


def run_config_sharded(name, a, rank, world):
    """Strong scaling (BASELINE config 3: the 1 M environments sharded over 1/2/4/8 GPUs): every rank runs its
    contiguous block, no per-step communication; time = max over ranks between two barriers; returns gathered."""
    import torch.distributed as dist
    from rcognita_b200 import shard
    c = CFG[name]
    Eg = (a.envs + 1023) // 1024 * 1024
    lo, hi = shard.shard_range(Eg, rank, world)
    dist.barrier()
    eng, r = timed_closed_loop(c, Eg, a.cands, a.t1, fit_evals=a.fit_evals, lo=lo, hi=hi)
    ms = torch.tensor([r["ms"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    counts = torch.tensor([int(eng.nsteps.sum().item()), int(eng.nsamples.sum().item())], dtype=torch.int64, device="cuda")
    ret, tot = shard.gather_returns(eng.accum, counts)
    if rank == 0:
        msv = float(ms.item())
        print(json.dumps(dict(config=name, n_gpus=world, scaling="strong", E_total=Eg, E_per_rank=hi - lo, C=a.cands, t1=a.t1,
                              ms_max_over_ranks=msv, intervals=r["intervals"], ms_per_interval=msv / max(r["intervals"], 1),
                              env_steps_total=int(tot[0].item()), mean_return=float(ret.mean().item()),
                              returns_gathered=int(ret.numel()))), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_config(name, a):
    rank, world = dist_info()
    if world > 1:
        return run_config_sharded(name, a, rank, world)
    c = CFG[name]
    eng, r = timed_closed_loop(c, a.envs, a.cands, a.t1, fit_evals=a.fit_evals)
    res = eng.results()
    r.update(config=name, critic_fit_evals=a.fit_evals, system=c["system"], mode=c["mode"], critic=c["cs"], Nactor=c["N"], t1=a.t1, dtype="f64",
             status_finished=int((res["status"] == _C.FINISHED).sum()), status_failed=int((res["status"] == _C.FAILED).sum()),
             mean_return=float(res["accum"].mean()), mean_Jc=float(res["Jc"].mean()),
             w_critic_mean=float(res["w_critic"].mean()), nfits_mean=float(res["nfits"].mean()))
    print(json.dumps(r), flush=True)


def run_config2_actors(a):
    """BASELINE config 2 (3wrobot_NI MPC Nactor=6, 65,536 envs x 256 per-environment candidates) with the three
    stand-ins for _actor_optimizer: arg-min over the candidates, the batched minimiser warm-started from the arg-min
    candidate, and the minimiser started from action_sqn_init like the reference.  Same initial states; reports
    time per control interval and the mean accumulated objective (closed-loop quality) at t1."""
    from bench_workload import synthetic_candidates as sc
    E = (a.envs + 1023) // 1024 * 1024
    bn = [[-25.0, 25.0], [-5.0, 5.0]]
    x0 = synthetic_states("3wrobotNI", 0, E, seed=0)
    cand = torch.as_tensor(sc(bn, 6, a.cands, seed=1, env_range=(0, E)), device="cuda")
    variants = [("candidates", dict(actor="candidates")), ("opt_argmin", dict(actor="opt", opt_start="argmin")),
                ("opt_init", dict(actor="opt", opt_start="init"))]
    if a.opt_sweep:          # stopping-rule sweep of the minimiser (pg_tol, f_tol, iteration cap)
        variants = [(f"opt_init pg={pg:g} f={ft:g} it={it}", dict(actor="opt", opt_start="init", opt_pg_tol=pg, opt_f_tol=ft, opt_iters=it))
                    for pg, ft, it in ((1e-7, 1e-12, 300), (1e-6, 1e-10, 300), (1e-5, 1e-9, 300), (1e-4, 1e-8, 300),
                                       (1e-3, 1e-7, 300), (1e-7, 1e-12, 20), (1e-7, 1e-12, 10), (1e-5, 1e-9, 20))]
    for tag, kw in variants:
        eng = ClosedLoopEngine("3wrobotNI", x0, cand, ctrl_bnds=bn, mode="MPC", Nactor=6, dt=0.01, t1=a.t1,
                               R1=[1, 10, 1, 0, 0], **kw)
        for _ in range(3):
            eng.run_interval()
        torch.cuda.synchronize()
        s0 = int(eng.nsamples.sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k = eng.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res = eng.results()
        samples = int(eng.nsamples.sum().item()) - s0
        print(json.dumps(dict(config="config2", actor=tag, E=E, C=a.cands, t1=a.t1, intervals=k, ms_per_interval=ms / max(k, 1),
                              controller_samples_per_s=samples / ms * 1e3, mean_return=float(res["accum"].mean()),
                              median_return=float(np.median(res["accum"])),
                              mean_final_distance=float(np.sqrt((res["y"][:, :2] ** 2).sum(1)).mean()))), flush=True)
        del eng


def run_env_steps(a):
    """Closed-loop env-steps/s of 3wrobot_NI (dt = 0.01) at scale with cheap controllers in the loop: the nominal parking
    controller, MPC over a small shared candidate table, and MPC with the batched optimiser -- the env-step side of the
    north-star target (the headline bench is dominated by its 256 candidate evaluations per environment and sample)."""
    E = (a.envs + 1023) // 1024 * 1024
    bn = [[-25.0, 25.0], [-5.0, 5.0]]
    x0 = torch.as_tensor(synthetic_states("3wrobotNI", 0, E, seed=0), device="cuda")
    cand16 = synthetic_candidates(bn, 6, 16, seed=1)
    for tag, cand, kw in (("nominal", None, dict(actor="nominal", ctrl_gain=0.5)),
                          ("mpc_16_shared_candidates", cand16, dict(actor="candidates")),
                          ("mpc_optimizer_pg1e-4", None, dict(actor="opt", opt_start="init", opt_pg_tol=1e-4, opt_f_tol=1e-8))):
        eng = ClosedLoopEngine("3wrobotNI", x0, cand, ctrl_bnds=bn, mode="MPC", Nactor=6, dt=0.01, t1=1e6, R1=[1, 10, 1, 0, 0], **kw)
        for _ in range(5):
            eng.run_interval()
        torch.cuda.synchronize()
        s0 = int(eng.nsteps.sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.intervals):
            eng.run_interval()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        steps = int(eng.nsteps.sum().item()) - s0
        print(json.dumps(dict(config="closed-loop env-steps", controller=tag, E=E, intervals=a.intervals,
                              ms_per_interval=ms / a.intervals, env_steps_per_s=steps / ms * 1e3,
                              failed=int((eng.status == _C.FAILED).sum().item()))), flush=True)
        del eng


def pct(x):
    x = np.asarray(x, dtype=np.float64)
    return {"p50": float(np.percentile(x, 50)), "p99": float(np.percentile(x, 99)), "max": float(x.max())}


def run_fp32_report(a):
    """Config 4 in fp64 and fp32 on identical inputs: (i) one E x C actor-cost launch (J, arg-min agreement),
    (ii) the closed loop with the critic weights pinned (trajectory / return divergence at t1), (iii) `_critic_cost`
    (J_c) in fp32 (rcg_critic_cost_f32) against fp64 on the FIFO buffers and weights the fp64 closed loop WITH critic
    refit holds at several times of the episode.  The fit itself runs in fp64 only (its Gram matrices need it)."""
    c = CFG["config4"]
    n, m = _C.SYS_DIMS[_C.SYS_IDS[c["system"]]]
    E = (a.envs + 1023) // 1024 * 1024
    C = a.cands
    x0 = synthetic_states(c["system"], 0, E, seed=0)
    cand = synthetic_candidates(c["bnds"], c["N"], C, seed=1)
    w = np.array([11.0, 11.0, 1.0])
    sysd = _C.make_system(c["system"], c["pars"], c["bnds"])
    obj = _C.make_objective(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["dt"] * c["psm"], critic_struct=c["cs"],
                            R1=c["R1"], observation_target=c["target"])
    x64 = torch.as_tensor(x0.T.copy(), device="cuda")
    c64 = torch.as_tensor(cand.T.copy(), device="cuda")
    w64 = torch.as_tensor(w, device="cuda")
    J64, am64, _ = ops.actor_cost(sysd, obj, x64, x64, c64, False, C, w_critic=w64)
    J32, am32, _ = ops.actor_cost(sysd, obj, x64.float(), x64.float(), c64.float(), False, C, w_critic=w64.float())
    relJ = ((J32.double() - J64).abs() / J64.abs().clamp_min(1e-300)).cpu().numpy().reshape(-1)
    agree = float((am64 == am32).double().mean().item())
    # where the arg-min differs: how much worse (in fp64 cost) is the fp32 pick
    idx = torch.arange(E, device="cuda")
    regret = ((J64[idx, am32.long()] - J64[idx, am64.long()]) / J64[idx, am64.long()].abs().clamp_min(1e-300)).cpu().numpy()
    e64, r64 = timed_closed_loop(c, E, C, a.t1, critic_fit=False, w_critic=w)
    e32, r32 = timed_closed_loop(c, E, C, a.t1, critic_fit=False, w_critic=w, dtype=torch.float32)
    y64, y32 = e64.results(), e32.results()
    rel_y = np.abs(y32["y"].astype(np.float64) - y64["y"]) / np.maximum(np.abs(y64["y"]), 1e-2)
    rel_acc = np.abs(y32["accum"].astype(np.float64) - y64["accum"]) / np.abs(y64["accum"])
    # (iii) J_c: fp64 closed loop with the critic refitted at every sample; at checkpoints evaluate _critic_cost at the
    # current weights (and at w_critic_init = ones) in both precisions on the same buffers
    eng = ClosedLoopEngine(c["system"], x0, cand, pars=c["pars"], ctrl_bnds=c["bnds"], mode=c["mode"], Nactor=c["N"], dt=c["dt"],
                           pred_step_size=c["dt"] * c["psm"], t1=a.t1, R1=c["R1"], observation_target=c["target"],
                           critic_struct=c["cs"], critic_fit=True, Ncritic=4, buffer_size=10, action_init=c["a_init"])
    total = int(round(a.t1 / c["dt"]))
    marks = sorted({max(8, total // 100), total // 10, total // 2, total - 2})
    jc = {}
    done = 0
    for mk in marks:
        for _ in range(mk - done):
            eng.run_interval()
        done = mk
        for tag, wt in (("fitted_w", eng.w), ("w_init", torch.ones_like(eng.w))):
            J64c = ops.critic_cost(obj, n, m, eng.obs_buf, eng.act_buf, wt[:, :, None].contiguous(), eng.w_prev)[:, 0]
            J32c = ops.critic_cost(obj, n, m, eng.obs_buf.float(), eng.act_buf.float(), wt.float()[:, :, None].contiguous(),
                                   eng.w_prev.float())[:, 0]
            scale = J64c.abs().clamp_min(1e-300)
            rel = ((J32c.double() - J64c).abs() / scale).cpu().numpy()
            # fitted costs are often ~0 (the 3-row problem is feasible): absolute error against the cost at w_init as well
            J0 = ops.critic_cost(obj, n, m, eng.obs_buf, eng.act_buf, torch.ones_like(eng.w)[:, :, None].contiguous(), eng.w_prev)[:, 0]
            rel0 = ((J32c.double() - J64c).abs() / J0.abs().clamp_min(1e-300)).cpu().numpy()
            jc[f"interval_{mk}_{tag}"] = {"rel_err": pct(rel[np.isfinite(rel)]), "err_over_Jc_at_w_init": pct(rel0[np.isfinite(rel0)]),
                                          "Jc_fp64_median": float(J64c.median().item())}
    del eng
    out = {"config": "config4-fp32-report", "E": E, "C": C, "t1": a.t1, "critic_cost_fp32_vs_fp64": jc,
           "actor_cost_rel_err_fp32_vs_fp64": pct(relJ), "argmin_agreement": agree, "argmin_regret_rel": pct(regret),
           "closed_loop_state_rel_err_at_t1": pct(rel_y.reshape(-1)), "closed_loop_return_rel_err": pct(rel_acc),
           "same_step_counts": float((y32["nsteps"] == y64["nsteps"]).mean()),
           "fp64": {k: r64[k] for k in ("ms", "env_steps_per_s", "actor_evals_per_s")},
           "fp32": {k: r32[k] for k in ("ms", "env_steps_per_s", "actor_evals_per_s")},
           "note": "the critic FIT runs in fp64 only; _critic_cost has an fp32 twin (rcg_critic_cost_f32)"}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["config2", "config3", "config4", "fp32", "envsteps"])
    ap.add_argument("--intervals", type=int, default=200)
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--cands", type=int, default=256)
    ap.add_argument("--t1", type=float, default=0.0)
    ap.add_argument("--opt-sweep", action="store_true", help="config2: sweep the minimiser's stopping rule")
    ap.add_argument("--fit-evals", type=int, default=0, help="work bound of the critic fit per environment (0 = to convergence)")
    a = ap.parse_args()
    if a.what == "envsteps":
        a.envs = a.envs or 1 << 20
        run_env_steps(a)
    elif a.what == "config2":
        a.envs, a.t1 = a.envs or 65536, a.t1 or 2.0
        run_config2_actors(a)
    elif a.what == "config3":
        a.envs, a.t1 = a.envs or 1 << 20, a.t1 or 0.3
        run_config("config3", a)
    elif a.what == "config4":
        a.envs, a.t1 = a.envs or 262144, a.t1 or 100.0            # t1 = 100
        run_config("config4", a)
    else:
        a.envs, a.t1 = a.envs or 262144, a.t1 or 100.0
        run_fp32_report(a)


if __name__ == "__main__":
    main()
