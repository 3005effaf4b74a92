#!/usr/bin/env python3
"""bench.py -- throughput of rcognita's hot path on B200 (and of its CPU restatement).

Workload (BASELINE.json configs[1]): 3wrobot_NI, MPC, Nactor=6, dt=0.01; 65,536 environments
PER GPU x 256 candidate action sequences per environment, RK45 closed loop.  One bench "step"
= one control interval of the whole batch: `rcg_rk45_advance` integrates every environment
to its next controller sample (scipy-faithful RK45, ~2-3 accepted steps) and `rcg_actor_cost`
evaluates E x C `_actor_cost` rollouts (candidate stream staged by TMA), takes the per-environment
arg-min and hands the action over.  `value` = `_actor_cost` evaluations per second over the whole closed loop (all GPUs),
`env_steps_per_s` the accepted RK45 steps per second of the same timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA)
    python bench.py --impl reference ...                           # CPU arm: the oracle port, all host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SYSTEM = "3wrobotNI"
BNDS = [[-25.0, 25.0], [-5.0, 5.0]]
R1_DIAG = [1.0, 10.0, 1.0, 0.0, 0.0]
DT = 0.01
ACTION_INIT = [-2.5, -0.5]
METRIC = "closed-loop actor-cost evals/s (+ env-steps/s), 3wrobot_NI MPC Nactor=6"
BYTES_PER_EVAL_PER_ENV_CAND = lambda N, m: N * m * 8 + 8      # candidate read + J (folded into arg-min) -- DESIGN.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU")
    ap.add_argument("--cands", type=int, default=256)
    ap.add_argument("--nactor", type=int, default=6)
    ap.add_argument("--shared-cands", action="store_true", help="one shared candidate table instead of per-env sets")
    ap.add_argument("--cpu-sample-envs", type=int, default=0, help="environments in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="environment blocks (streams) of the host-staged loop")
    ap.add_argument("--blocks", type=int, default=2, help="environment blocks (streams) of the device-resident pipelined loop")
    ap.add_argument("--no-extra", action="store_true", help="skip the env-steps and config-3 strong-scaling blocks")
    ap.add_argument("--envsteps-envs", type=int, default=1 << 20, help="environments per GPU of the env-steps block")
    ap.add_argument("--envsteps-intervals", type=int, default=100)
    ap.add_argument("--config3-envs", type=int, default=1 << 20, help="TOTAL environments of the config-3 strong-scaling block")
    ap.add_argument("--config3-t1", type=float, default=2.0)
    ap.add_argument("--no-reference-python", action="store_true", help="skip timing the unmodified reference from baseline/_ref")
    ap.add_argument("--reference-python-budget", type=float, default=6.0, help="seconds of wall budget per reference loop")
    ap.add_argument("--no-opt", action="store_true", help="skip the extra block that times the closed loop with the batched actor optimiser")
    ap.add_argument("--no-graph", action="store_true", help="drive the host-staged loop from Python instead of one CUDA graph per block")
    return ap.parse_args()


def workload_name(args):
    kind = "shared candidate table" if args.shared_cands else "per-env candidates"
    return (f"3wrobot_NI MPC Nactor={args.nactor} dt={DT}: {args.envs} envs/GPU x {args.cands} candidate action "
            f"sequences ({kind}), RK45 closed loop")


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_start=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ts, line in self.rows:
            if t_start is not None and not (t_start - 0.05 <= ts <= t_end + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms (oracle port)

def host_threads():
    """Host threads available to this process (torchrun exports OMP_NUM_THREADS=1 to its workers, which must not
    turn the CPU arm into a single-thread run: the oracle's OpenMP loops are sized explicitly with this)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample_envs(args, threads):
    """Environments in the CPU sample: 256 per thread, rounded up to whole Philox blocks of 1,024."""
    n = args.cpu_sample_envs or 256 * threads
    return max(1024, (n + 1023) // 1024 * 1024)


def cpu_closed_loop(args, sample_envs, steps, warmup, budget_s=None):
    """Times the CPU restatement of the same closed loop (oracle/, C + OpenMP on all host threads)
    on a bounded sample of the workload: `sample_envs` environments with the same distributions.
    Returns (evals/s, env-steps/s, ms/step, threads, steps run)."""
    import oracle
    from bench_workload import make_workload
    x0, cand = make_workload(args, 0, sample_envs)
    s = oracle.make_sys(SYSTEM, [], BNDS)
    c = oracle.make_ctrl(3, 2, mode="MPC", Nactor=args.nactor, pred_step_size=DT, R1=R1_DIAG)
    t1 = max(10.0, (steps + warmup + 16) * 3 * DT)
    batch = oracle.EnvBatch(c, s, x0, cand, ACTION_INIT, DT, 0.0, t1, DT / 2)
    threads = host_threads() if oracle.num_threads() >= 1 and oracle.has_openmp() else 1
    for _ in range(warmup):
        batch.interval(threads)
    tot_steps = tot_evals = done = 0
    t_begin = time.perf_counter()
    for _ in range(steps):
        st, ev = batch.interval(threads)
        tot_steps += st; tot_evals += ev; done += 1
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    el = time.perf_counter() - t_begin
    return tot_evals / el, tot_steps / el, 1e3 * el / done, threads, done


def bench_t1(args):
    """Episode length of both arms: long enough that no environment finishes inside the bench."""
    return max(10.0, (2 * (args.steps + max(args.warmup, 3)) + 32) * 3 * DT) + 120.0      # + the clock-sampling continuation


def bench_config(args, world):
    """The `config` object of the JSON line -- the same keys and values for both arms."""
    E, C, N = args.envs, args.cands, args.nactor
    return {"workload": workload_name(args), "envs_total": E * world, "candidates": C, "Nactor": N, "t1": bench_t1(args),
            "l2": f"per-launch candidate stream {E * C * N * 2 * 8 / 1e9:.2f} GB > 126 MB L2 (no flush needed)"
                  if not args.shared_cands else "shared table: cache-resident by design",
            "sharding": f"{world} x contiguous env blocks, no per-step communication"}


def reference_python_baseline(args, cores):
    """The UNMODIFIED reference (baseline/_ref, numpy/scipy) on the host cores: preset-faithful SLSQP loop and the
    candidate/arg-min loop, a fixed wall budget each (bench_reference.py)."""
    if args.no_reference_python:
        return None
    try:
        import bench_reference
        return bench_reference.measure(cores, budget_s=args.reference_python_budget, nactor=args.nactor, ncand=args.cands)
    except Exception as exc:                               # never let the optional block take the bench line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the path on all host threads, same config / metric: the
    oracle port (C + OpenMP; ~200x faster per core than the numpy/scipy original, so the ratio against it is the
    conservative one), with the unmodified Python reference from baseline/_ref timed beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sample = cpu_sample_envs(args, host_threads())
    evals_s, steps_s, ms, threads, done = cpu_closed_loop(args, sample, args.steps, args.warmup)
    cpu = {"value": evals_s, "unit": "evals/s", "cores": threads, "kind": "port", "env_steps_per_s": steps_s,
           "sample": f"{sample} of {args.envs} envs x {args.cands} candidates, one control interval per step "
                     f"(oracle/rcg_oracle.c, OpenMP)",
           "reference_python": reference_python_baseline(args, threads)}
    line = {
        "impl": "reference", "metric": METRIC, "value": evals_s, "unit": "evals/s", "env_steps_per_s": steps_s,
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, max(world, args.gpus)), "cpu_baseline": cpu,
        "e2e": {"value": evals_s, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- CUDA arm

def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPUs NVML reports as local to
    its GPU: with 8 ranks on one host the per-step H2D/D2H traffic of the host-staged loop otherwise crosses NUMA
    nodes.  Returns the previous affinity (restored before the CPU baseline), or None if NVML is unavailable."""
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, prev)
        return prev
    except Exception:
        return None


FP64_PEAK_TFLOPS = 36.9          # DFMA issue peak measured with tools/peaks.cu on this pool (profiles/r01_peaks.jsonl)
NI_STEP_FLOP = 272               # algorithmic flop per accepted 3wrobot_NI RK45 step
ACTOR_DRAM_BYTES_PER_EVAL = None  # filled from profiles/actor_cost_traffic.json (ncu dram__bytes per launch / evals)


def extra_env_steps(args, rank, world, dev, barrier, allreduce_max, allreduce_sum):
    """Closed-loop env-steps/s of 3wrobot_NI at 1,048,576 environments PER GPU with cheap controllers in the loop (the
    env-step half of BASELINE.json's metric): the reference's nominal parking controller and MPC over 16 shared
    candidates.  rk45_advance is FP64-pipe-bound: fraction of the measured DFMA peak at SURVEY section 8d's 272 flop/step."""
    import torch
    from rcognita_b200 import shard
    from rcognita_b200.engine import ClosedLoopEngine
    from bench_workload import synthetic_candidates, synthetic_states
    E = args.envsteps_envs
    lo, hi = shard.shard_range(E * world, rank, world)
    x0 = torch.as_tensor(synthetic_states(SYSTEM, lo, hi, seed=0), device=dev)
    out = {"envs_per_gpu": E, "intervals": args.envsteps_intervals, "scaling": "weak"}
    for tag, cand, kw in (("nominal", None, dict(actor="nominal", ctrl_gain=0.5)),
                          ("mpc_16_shared_candidates", synthetic_candidates(BNDS, args.nactor, 16, seed=1), dict(actor="candidates"))):
        eng = ClosedLoopEngine(SYSTEM, x0, cand, ctrl_bnds=BNDS, mode="MPC", Nactor=args.nactor, dt=DT, t1=1e6, R1=R1_DIAG,
                               device=dev, **kw)
        for _ in range(5):
            eng.run_interval()
        barrier()
        s0 = int(eng.nsteps.sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.envsteps_intervals):
            eng.run_interval()
        e1.record()
        barrier()
        ms = allreduce_max(e0.elapsed_time(e1))
        steps = allreduce_sum(int(eng.nsteps.sum().item()) - s0)
        sps = steps / (ms * 1e-3)
        out[tag] = {"env_steps_per_s": sps, "ms_per_interval": ms / args.envsteps_intervals,
                    "fp64_roofline": {"bound": "fp64 pipe", "flop_per_step": NI_STEP_FLOP, "achieved_tflops": sps * NI_STEP_FLOP / 1e12,
                                      "peak_tflops": FP64_PEAK_TFLOPS * world, "frac": sps * NI_STEP_FLOP / 1e12 / (FP64_PEAK_TFLOPS * world),
                                      "note": "a transcendental counts as ONE flop in SURVEY 8d; fp64 sincos/pow expand to ~45 "
                                              "instructions, so the pipe utilisation is ~3x this fraction (ncu: profiles/)"}}
        del eng
    return out


def extra_config3_strong(args, rank, world, dev, barrier, allreduce_max, allreduce_sum):
    """BASELINE config 3, STRONG scaling: 1,048,576 Sys3WRobot environments in total sharded over the ranks, RQL with the
    'quadratic' critic refitted at every sample, Nactor = 10, 256 shared candidates, whole episodes of t1 = 2.
    No per-step communication; per-environment returns gathered over NCCL at the end."""
    import torch
    from rcognita_b200 import shard
    from rcognita_b200.engine import ClosedLoopEngine
    from bench_workload import synthetic_candidates, synthetic_states
    Eg = args.config3_envs
    lo, hi = shard.shard_range(Eg, rank, world)
    bn = [[-300, 300], [-100, 100]]
    x0 = synthetic_states("3wrobot", lo, hi, seed=0)
    cand = synthetic_candidates(bn, 10, 256, seed=1)
    eng = ClosedLoopEngine("3wrobot", x0, cand, pars=[10, 1], ctrl_bnds=bn, mode="RQL", Nactor=10, dt=0.01, pred_step_size=0.02,
                           t1=args.config3_t1, R1=[1, 10, 1, 0, 0, 0, 0], critic_struct="quadratic", critic_fit=True, Ncritic=4,
                           buffer_size=10, device=dev)
    for _ in range(3):
        eng.run_interval()
    barrier()
    s0, n0 = int(eng.nsteps.sum().item()), int(eng.nsamples.sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = eng.run()
    e1.record()
    barrier()
    ms = allreduce_max(e0.elapsed_time(e1))
    steps = allreduce_sum(int(eng.nsteps.sum().item()) - s0)
    samples = allreduce_sum(int(eng.nsamples.sum().item()) - n0)
    cnt = torch.tensor([0], dtype=torch.int64, device=dev)
    returns, _ = shard.gather_returns(eng.accum, cnt)
    return {"scaling": "strong", "envs_total": Eg, "envs_per_gpu": hi - lo, "t1": args.config3_t1, "candidates": 256, "Nactor": 10,
            "critic": "quadratic (28 weights), refit every sample", "intervals": k, "ms_total": ms, "ms_per_interval": ms / max(k, 1),
            "env_steps_per_s": steps / (ms * 1e-3), "actor_evals_per_s": samples * 256 / (ms * 1e-3),
            "critic_fits_per_s": samples / (ms * 1e-3), "mean_return": float(returns.mean().item()),
            "returns_gathered": int(returns.numel())}


def run_b200(args):
    # Everything the libraries print (NCCL's version banner goes to stdout) is sent to stderr: stdout carries
    # exactly one JSON line.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import rcognita_b200
    from rcognita_b200 import shard
    from rcognita_b200.engine import ClosedLoopEngine, HostStagedLoop, PipelinedLoop
    from bench_workload import make_workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rcognita_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    prev_affinity = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    K, W = args.steps, max(args.warmup, 3)
    E, C, N = args.envs, args.cands, args.nactor
    lo, hi = shard.shard_range(E * world, rank, world)
    x0, cand = make_workload(args, lo, hi)
    t1 = bench_t1(args)
    kw = dict(ctrl_bnds=BNDS, mode="MPC", Nactor=N, dt=DT, t1=t1, R1=R1_DIAG, action_init=ACTION_INIT)
    cand = torch.as_tensor(cand, device=dev)                  # one upload, sliced by both loops
    loop = PipelinedLoop(SYSTEM, x0, cand, nchunks=args.blocks, device=dev, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce_max(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce_sum(x):
        t = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    # ---- device-resident closed loop: warm-up, then exactly K timed steps.  The clock sampler is started BEFORE the
    # warm-up so that nothing but the barrier sits between the last warm-up step and the first timed one (a 0.3 s pause
    # there lets the GPU fall back to its idle clocks, which a 6 ms timed region then spends ramping up).
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0.0)
    barrier()
    for _ in range(W):
        loop.step()
    loop.synchronize()
    barrier()
    steps0, samples0 = int(loop.field("nsteps").sum().item()), int(loop.field("nsamples").sum().item())
    barrier()
    rcognita_b200.reset_launch_count()
    loop.actor_events = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(K):
        loop.step()
    loop.synchronize()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = rcognita_b200.launch_count()
    actor_kernel_name = rcognita_b200.last_actor_kernel()       # what the library dispatched to, not a literal
    ms_total = ev0.elapsed_time(ev1)
    actor_ms = [a.elapsed_time(b) for a, b in loop.actor_events]
    # time during which at least one actor launch was running (the blocks' launches overlap each other and the other
    # block's rk45_advance): union of the [start, end] intervals on the clock of the timed region's first event
    spans = sorted((ev0.elapsed_time(a), ev0.elapsed_time(b)) for a, b in loop.actor_events)
    actor_union_ms, cur_a, cur_b = 0.0, None, None
    for a_, b_ in spans:
        if cur_b is None or a_ > cur_b:
            actor_union_ms += (cur_b - cur_a) if cur_b is not None else 0.0
            cur_a, cur_b = a_, b_
        else:
            cur_b = max(cur_b, b_)
    actor_union_ms += (cur_b - cur_a) if cur_b is not None else 0.0
    loop.actor_events = None
    d_steps = int(loop.field("nsteps").sum().item()) - steps0
    d_evals = (int(loop.field("nsamples").sum().item()) - samples0) * C
    clocks = None
    if sampler:
        # nvidia-smi cannot sample faster than ~10 Hz: a timed region shorter than half a second yields fewer than five
        # samples, which is not a measurement.  In that case the SAME loop simply keeps running (untimed) until the
        # sampler has seen it for 1.5 s, and the clocks line says which region it describes.
        if t_wall1 - t_wall0 < 0.6:
            t_c0 = time.time()
            while time.time() - t_c0 < 1.5:
                for _ in range(64):
                    loop.step()
                loop.synchronize()
                torch.cuda.synchronize()
            clocks = sampler.stop(t_c0, time.time())
            clocks["region"] = (f"untimed continuation of the same loop for 1.5 s (the timed region of {K} steps lasted "
                                f"{1e3 * (t_wall1 - t_wall0):.1f} ms: too short for 5 nvidia-smi samples)")
        else:
            clocks = sampler.stop(t_wall0, t_wall1)
            clocks["region"] = "timed region"
        if clocks.get("samples", 0) < 5:
            clocks["reasons"] = list(clocks.get("reasons", [])) + ["fewer than 5 samples: not a measurement"]

    ms_total = allreduce_max(ms_total)
    actor_ms_avg = allreduce_max(float(np.mean(actor_ms)))
    actor_union_ms = allreduce_max(actor_union_ms)
    cnt = torch.tensor([d_steps, d_evals, launches], dtype=torch.int64, device=dev)
    # end-of-run collectives (the only ones on the path): gather per-env returns, sum the counters
    returns, cnt = shard.gather_returns(loop.field("accum"), cnt)
    tot_steps, tot_evals, tot_launches = (int(v) for v in cnt.tolist())
    value = tot_evals / (ms_total * 1e-3)
    steps_per_s = tot_steps / (ms_total * 1e-3)
    evals_per_actor_launch = d_evals / max(1, K * loop.nchunks)
    del loop

    # ---- end to end: the caller owns state and time in pinned HOST memory, copied in and out every step
    e2e = None
    if not args.no_e2e:
        hloop = HostStagedLoop(SYSTEM, x0, cand, nchunks=args.e2e_chunks, device=dev, graph=not args.no_graph, **kw)
        hloop.run(3)
        barrier()
        s0 = int(hloop.field("nsamples").sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d, d2h = hloop.run(K)
        e1.record()
        barrier()
        ms_e2e = allreduce_max(e0.elapsed_time(e1))
        ev_e2e = allreduce_sum((int(hloop.field("nsamples").sum().item()) - s0) * C)
        e2e = {"value": ev_e2e / (ms_e2e * 1e-3), "unit": "evals/s",
               "h2d_bytes_per_step": int(h2d // K) * world, "d2h_bytes_per_step": int(d2h // K) * world,
               "ms_per_step": ms_e2e / K,
               "host_inputs": "state y[3][E] and solver time t[E] per environment (what Simulator.get_sim_step_data returned)",
               "host_outputs": "y, t, action[2][E], accum_obj[E], status, sample_flag, argmin",
               "device_resident": f"solver internals (f, h_abs, state_sys, clocks, counters) and the per-environment candidate "
                                  f"sets ({E * C * N * 2 * 8 / 1e9:.2f} GB per GPU: controller parameters, uploaded once)",
               "api": f"HostStagedLoop.run: {hloop.nchunks} environment blocks, each one pinned-host copy in + rk45_advance + "
                      "actor_cost + one copy out per control interval" + ("" if args.no_graph else " (one CUDA graph per block)")
                      + "; a block's next interval starts when ITS results are on the host"}
        del hloop
    del cand

    extra = None
    if not args.no_extra:
        # (an optional block that fails on every rank -- out of memory on a smaller device, say -- must not take the headline
        #  line down with it; its entry then says why)
        extra = {}
        for key, fn in (("env_steps", extra_env_steps), ("config3_strong", extra_config3_strong)):
            try:
                extra[key] = fn(args, rank, world, dev, barrier, allreduce_max, allreduce_sum)
            except Exception as exc:
                extra[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                torch.cuda.empty_cache()

    # ---- extra (N = 1): the same closed loop with the batched bounded minimiser standing in for the SLSQP
    #      _actor_optimizer instead of enumerate-and-argmin; reported beside the headline, not part of it
    actor_opt = None
    if world == 1 and not args.no_opt:
        try:
            K2 = min(K, 200)
            eng2 = ClosedLoopEngine(SYSTEM, x0, None, ctrl_bnds=BNDS, mode="MPC", Nactor=N, dt=DT, t1=t1, R1=R1_DIAG,
                                    action_init=ACTION_INIT, device=dev, actor="opt", opt_start="init", opt_pg_tol=1e-4,
                                    opt_f_tol=1e-8)
            for _ in range(W):
                eng2.run_interval()
            torch.cuda.synchronize()
            n0, st0 = int(eng2.nsamples.sum().item()), int(eng2.nsteps.sum().item())
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(K2):
                eng2.run_interval()
            a1.record()
            torch.cuda.synchronize()
            ms_opt = a0.elapsed_time(a1)
            actor_opt = {"ms_per_step": ms_opt / K2, "steps": K2,
                         "solves_per_s": (int(eng2.nsamples.sum().item()) - n0) / (ms_opt * 1e-3),
                         "env_steps_per_s": (int(eng2.nsteps.sum().item()) - st0) / (ms_opt * 1e-3),
                         "kernel": rcognita_b200.last_actor_opt_kernel(),      # variant rcg_actor_opt dispatched to (rcg_last_actor_opt_kernel)
                         "api": "ClosedLoopEngine(actor='opt', opt_start='init', opt_pg_tol=1e-4, opt_f_tol=1e-8): rcg_rk45_advance + "
                                "rcg_actor_opt (exact adjoint gradient, projected L-BFGS from action_sqn_init, one bounded "
                                "minimisation of _actor_cost per environment and control interval)"}
            del eng2
            if not args.no_cpu_baseline:
                # the checker's restatement of the same minimiser on the host cores (bounded sample), like cpu_baseline
                import oracle
                if prev_affinity:
                    os.sched_setaffinity(0, prev_affinity)          # all host cores, like the cpu_baseline leg below
                ns = min(E, 32768)
                so = oracle.make_sys(SYSTEM, [], BNDS)
                co = oracle.make_ctrl(3, 2, mode="MPC", Nactor=N, pred_step_size=DT, R1=R1_DIAG)
                xs = np.ascontiguousarray(np.asarray(x0[:ns], dtype=np.float64))
                sq0 = np.tile(np.asarray(ACTION_INIT, dtype=np.float64), N)
                oracle.actor_opt_batch(co, so, sq0, xs[:1024], pg_tol=1e-4, f_tol=1e-8)            # thread pool warm-up
                tc = time.time()
                oracle.actor_opt_batch(co, so, sq0, xs, pg_tol=1e-4, f_tol=1e-8)
                tc = time.time() - tc
                actor_opt["cpu_port"] = {"solves_per_s": ns / tc, "cores": oracle.num_threads(), "kind": "port",
                                         "sample": f"{ns} minimisations from action_sqn_init at the initial states "
                                                   "(oracle/rcg_oracle_opt.c, C + OpenMP)"}
        except Exception as exc:                               # optional block: never take the headline line down
            actor_opt = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (actor_cost_tma_kernel), measured live with CUDA events around every launch
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_eval = (N * 2 * 8 + 8) if not args.shared_cands else 8
    read_bytes_per_eval = (N * 2 * 8) if not args.shared_cands else 0
    # `achieved`: algorithmic bytes of all actor launches of the timed region / time during which an actor launch was
    # running (CUDA events on the launching streams).  With one block this is bytes per launch / launch duration; with
    # the blocks pipelined the launches overlap each other, so the per-launch duration alone would count the same
    # wall time twice (kept below as `per_launch`).
    evals_region = d_evals
    achieved = evals_region * bytes_per_eval / (actor_union_ms * 1e-3) / 1e9
    per_launch = evals_per_actor_launch * bytes_per_eval / (actor_ms_avg * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": actor_kernel_name, "kernel_source": "rcg_last_actor_kernel() after the timed region",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel: read + write)" if peaks else "fallback 6650 GB/s",
                "bytes_per_eval": bytes_per_eval,
                "bytes_per_eval_note": f"SURVEY 8d: {read_bytes_per_eval} B candidate read + 8 B cost; this launch folds the cost into "
                                       "the arg-min and never writes it, so frac_read / frac_dram below are the physical figures",
                "frac_read": achieved / peak_gbs * read_bytes_per_eval / bytes_per_eval,
                "evals_per_launch": evals_per_actor_launch, "launches_per_step": args.blocks,
                "kernel_ms_avg": actor_ms_avg, "kernel_busy_ms_per_step": actor_union_ms / K,
                "kernel_share_of_step": actor_union_ms / ms_total,
                "kernel_evals_per_s": evals_region / (actor_union_ms * 1e-3),
                "per_launch": {"achieved": per_launch, "frac": per_launch / peak_gbs,
                               "note": f"{args.blocks} environment blocks on {args.blocks} streams: a launch shares the GPU with the other "
                                       "block's actor launch and rk45_advance for part of its duration"},
                # the whole control interval against the same peak: algorithmic bytes of the step / ms_per_step
                "frac_step": (d_evals / K) * bytes_per_eval / (ms_total / K * 1e-3) / 1e9 / peak_gbs}
    prof = os.path.join(ROOT, "profiles", "actor_cost_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as fh:
                tr = json.load(fh)
            per_eval = float(tr["dram_bytes_per_launch"]) / float(tr["evals_per_launch"])
            roofline["traffic"] = per_eval * evals_per_actor_launch
            roofline["traffic_source"] = (f"ncu dram__bytes_read.sum + dram__bytes_write.sum = {tr['dram_bytes_per_launch']:.4g} B for "
                                          f"{int(tr['evals_per_launch'])} evals ({per_eval:.1f} B/eval, {tr.get('source', 'profiles/')})")
            roofline["frac_dram"] = per_eval * evals_region / (actor_union_ms * 1e-3) / 1e9 / peak_gbs
        except (OSError, ValueError, KeyError):
            pass

    cpu = None
    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)
    if world == 1 and not args.no_cpu_baseline:
        sample = cpu_sample_envs(args, host_threads())
        ev_s, st_s, ms, threads, done = cpu_closed_loop(args, sample, 400, 3, budget_s=15.0)
        cpu = {"value": ev_s, "unit": "evals/s", "cores": threads, "kind": "port", "env_steps_per_s": st_s,
               "sample": f"{sample} of {E} envs x {C} candidates, {done} control intervals "
                         f"(oracle/rcg_oracle.c, C + OpenMP, {threads} threads)",
               "reference_python": reference_python_baseline(args, threads)}

    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "env_steps_per_s": steps_per_s, "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(args, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": tot_launches, "roofline": roofline, "cpu_baseline": cpu,
        "extra": extra, "actor_optimizer": actor_opt, "mean_return_so_far": float(returns.mean().item()),
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.envs % 1024:
        raise SystemExit("--envs must be a multiple of 1024")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
