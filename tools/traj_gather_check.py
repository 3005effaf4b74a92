#!/usr/bin/env python3
"""Multi-GPU check of the end-of-run collectives (torchrun, NCCL): every rank runs its shard of a small closed loop with
the device-side trajectory ring, `shard.gather_trajectories` / `gather_returns` bring rings and returns to rank 0 in global
environment order, and rank 0 compares them with the same batch run unsharded on its own GPU -- bit for bit.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/traj_gather_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_workload import synthetic_candidates, synthetic_states  # noqa: E402
from rcognita_b200 import shard  # noqa: E402
from rcognita_b200.engine import ClosedLoopEngine  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl")
E = 3 * 1024                                        # uneven shards for world = 2
bn = [[-25.0, 25.0], [-5.0, 5.0]]
cand = synthetic_candidates(bn, 6, 64, seed=1)
kw = dict(ctrl_bnds=bn, mode="MPC", Nactor=6, dt=0.01, t1=0.25, R1=[1, 10, 1, 0, 0], log_every=2, log_capacity=16)


def run(lo, hi):
    eng = ClosedLoopEngine("3wrobotNI", synthetic_states("3wrobotNI", lo, hi, seed=0), cand, **kw)
    eng.run()
    return eng


lo, hi = shard.shard_range(E, rank, world)
eng = run(lo, hi)
rows, count = shard.gather_trajectories(eng.log.rows, eng.log.count)
counts = torch.tensor([int(eng.nsteps.sum().item())], dtype=torch.int64, device="cuda")
returns, tot = shard.gather_returns(eng.accum, counts)
if rank == 0:
    ref = run(0, E)
    ok = (torch.equal(rows, ref.log.rows) or bool(((rows == ref.log.rows) | (rows.isnan() & ref.log.rows.isnan())).all())) \
        and torch.equal(count, ref.log.count) and torch.equal(returns, ref.accum) and int(tot[0]) == int(ref.nsteps.sum().item())
    e = E - 5                                        # an environment owned by the last rank
    tr = shard.ring_rows(rows, count, e).cpu().numpy()
    print(json.dumps(dict(check="gather_trajectories + gather_returns over NCCL", world=world, E=E, ok=bool(ok),
                          ring_shape=list(rows.shape), rows_of_env=[e, int(tr.shape[0])], last_row_t=float(tr[-1, 0]),
                          mean_return=float(returns.mean().item()))), flush=True)
    assert ok
dist.barrier()
dist.destroy_process_group()
