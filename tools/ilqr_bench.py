"""rcg_actor_opt alone vs rcg_actor_ilqr + rcg_actor_opt on one batch of actor problems (device-resident, CUDA events).
Usage: python tools/ilqr_bench.py [E [f_tol]]   -> one line per configuration."""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from rcognita_b200 import _C, ops          # noqa: E402
from golden_util import DIMS, PRESET       # noqa: E402


def run(name, N, pred, R1, E, spread, center, presweeps, reps=3, f_tol=1e-12):
    n, m = DIMS[name]
    P = PRESET[name]
    sysd = _C.make_system(name, P["pars"], P["bnds"])
    obj = _C.make_objective(n, m, mode="MPC", Nactor=N, pred_step_size=pred, R1=R1)
    rng = np.random.default_rng(0)
    x = np.array(center)[:, None] + rng.normal(size=(n, E)) * np.array(spread)[:, None]
    st = torch.as_tensor(x, device="cuda")
    L = N * m
    ws_o, _ = ops.opt_workspace(sysd, obj, E, 1, st.device)
    ws_i = torch.empty((max(ops.ilqr_workspace_bytes(sysd, obj, E, 1) // 8, 1),), dtype=torch.float64, device="cuda")
    J = torch.empty((E,), dtype=torch.float64, device="cuda")
    it = torch.zeros((E,), dtype=torch.int32, device="cuda")
    nf = torch.zeros((E,), dtype=torch.int32, device="cuda")
    sw = torch.zeros((E,), dtype=torch.int32, device="cuda")
    best = []
    for r in range(reps + 1):
        sqn = torch.zeros((L, E), dtype=torch.float64, device="cuda")
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record()
        if presweeps:
            ops.actor_ilqr(sysd, obj, st, st, sqn, max_sweeps=presweeps, pg_tol=1e-7, workspace=ws_i, sweeps_out=sw)
        e1.record()
        ops.actor_opt(sysd, obj, st, st, sqn, max_iter=300, pg_tol=1e-7, f_tol=f_tol, workspace=ws_o, J_out=J, iters_out=it,
                      nfev_out=nf)
        e2.record()
        torch.cuda.synchronize()
        if r:
            best.append((e0.elapsed_time(e2), e0.elapsed_time(e1)))
    tot, pre = min(best)
    print(f"{name} N={N} E={E} presweeps={presweeps} f_tol={f_tol:g}: {tot:.2f} ms ({pre:.2f} ms sweeps) = {E / tot * 1e3:.3e} solves/s; "
          f"mean J {J.mean().item():.6f}; sweeps mean {sw.float().mean().item():.1f} max {sw.max().item()}; "
          f"iters mean {it.float().mean().item():.1f} max {it.max().item()}; nfev mean {nf.float().mean().item():.1f}", flush=True)


if __name__ == "__main__":
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    f_tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
    if f_tol != 1e-12:                     # tolerance study: Sys3WRobot only
        for pre in (0, 50):
            run("3wrobot", 10, 0.1, np.diag([10.0, 10.0, 1.0, 0, 0, 0, 0]), E, [0.5, 0.5, 0.2, 0.3, 0.3], [5.0, 5.0, 2.4, 0, 0], pre,
                f_tol=f_tol)
        sys.exit(0)
    for pre in (0, 25, 50):
        run("3wrobot", 10, 0.1, np.diag([10.0, 10.0, 1.0, 0, 0, 0, 0]), E, [0.5, 0.5, 0.2, 0.3, 0.3], [5.0, 5.0, 2.4, 0, 0], pre)
    for pre in (0, 25):
        run("3wrobotNI", 6, 0.01, PRESET["3wrobotNI"]["R1_diag"], E, [2.0, 2.0, 1.0], [0.0, 0.0, 0.0], pre)
