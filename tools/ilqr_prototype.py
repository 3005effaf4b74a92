#!/usr/bin/env python3
"""Offline numpy prototype of a control-limited Gauss-Newton (iLQR) sweep for CtrlOptPred._actor_cost, the candidate
successor of the projected L-BFGS minimiser.  Runs on the committed golden problems
(tests/golden/actor_opt.json: 72 problems solved by the live reference's SLSQP) and reports, per problem, the cost reached
and the number of sweeps next to the L-BFGS restatement of the oracle.  Analysis tool, not product code.

Per sweep: a reverse Riccati pass over the horizon (stage costs are exactly quadratic in (observation, action) for every
structure of the path; dynamics x+ = x + h f(x, a) linearised), with the box on the actions handled per stage by a clamped
Newton step (m <= 2: active-set enumeration), Levenberg-Marquardt regularisation on Q_aa, and a forward pass with
backtracking on the true cost.  State per lane: O(n^2 + N (m + m n)) instead of five N*m vectors plus 12 (s, y) pairs.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from golden_util import DIMS, PRESET  # noqa: E402


def dyn(name, pars, x, a):
    if name == "3wrobotNI":
        return np.array([a[0] * np.cos(x[2]), a[0] * np.sin(x[2]), a[1]])
    if name == "3wrobot":
        return np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), x[4], a[0] / pars[0], a[1] / pars[1]])
    t1, t2, K1, K2, K3 = pars
    return np.array([(-x[0] + K1 * a[0]) / t1, (-x[1] + K2 * x[0] + K3 * x[1] ** 2) / t2])


def dyn_jac(name, pars, x, a):
    n, m = DIMS[name]
    fx, fa = np.zeros((n, n)), np.zeros((n, m))
    if name == "3wrobotNI":
        fx[0, 2], fx[1, 2] = -a[0] * np.sin(x[2]), a[0] * np.cos(x[2])
        fa[0, 0], fa[1, 0], fa[2, 1] = np.cos(x[2]), np.sin(x[2]), 1.0
    elif name == "3wrobot":
        fx[0, 2], fx[1, 2] = -x[3] * np.sin(x[2]), x[3] * np.cos(x[2])
        fx[0, 3], fx[1, 3], fx[2, 4] = np.cos(x[2]), np.sin(x[2]), 1.0
        fa[3, 0], fa[4, 1] = 1 / pars[0], 1 / pars[1]
    else:
        t1, t2, K1, K2, K3 = pars
        fx[0, 0], fx[1, 0], fx[1, 1] = -1 / t1, K2 / t2, (-1 + 2 * K3 * x[1]) / t2
        fa[0, 0] = K1 / t1
    return fx, fa


def stage_quadratic(c, n, m, k, N):
    """(H, g0, shift_obs): stage term k = 1/2 z^T H z + g0^T z with z = [obs - shift, action] (exactly quadratic)."""
    p = n + m
    R1 = np.array(c["R1"])
    tgt = np.array(c["target"]) if len(c["target"]) else np.zeros(n)
    mode, cs = c["mode"], c["critic_struct"]
    use_critic = mode == "SQL" or (mode == "RQL" and k == N - 1)
    if not use_critic:
        gk = c["gamma"] ** k
        return gk * (R1 + R1.T), np.zeros(p), tgt
    w = np.array(c["w"])
    H, g0, shift = np.zeros((p, p)), np.zeros(p), tgt
    if cs in ("quad-lin", "quadratic"):
        idx = 0
        for i in range(p):
            for j in range(i, p):
                H[i, j] += w[idx]
                H[j, i] += w[idx]
                idx += 1
        if cs == "quad-lin":
            g0 = w[idx:idx + p].copy()
    elif cs == "quad-nomix":
        H = np.diag(2 * w[:p])
    else:                                        # quad-mix uses the RAW observation
        shift = np.zeros(n)
        idx = 0
        for i in range(n):
            H[i, i] = 2 * w[idx]; idx += 1
        for i in range(n):
            for j in range(m):
                H[i, n + j] += w[idx]; H[n + j, i] += w[idx]; idx += 1
        for j in range(m):
            H[n + j, n + j] = 2 * w[idx]; idx += 1
    return H, g0, shift


def box_newton(Qaa, Qa, lo, hi):
    """argmin 1/2 d^T Qaa d + Qa^T d, lo <= d <= hi (m <= 2): enumerate the active sets, keep the best feasible KKT point."""
    m = len(Qa)
    best, bestv, bestfree = None, np.inf, None
    for pat in np.ndindex(*([3] * m)):           # 0 free, 1 at lo, 2 at hi
        d = np.zeros(m)
        free = [j for j in range(m) if pat[j] == 0]
        for j in range(m):
            if pat[j] == 1: d[j] = lo[j]
            if pat[j] == 2: d[j] = hi[j]
        if free:
            A = Qaa[np.ix_(free, free)]
            rhs = -(Qa[free] + Qaa[np.ix_(free, [j for j in range(m) if j not in free])] @ d[[j for j in range(m) if j not in free]])
            try:
                d[free] = np.linalg.solve(A, rhs)
            except np.linalg.LinAlgError:
                continue
        if np.any(d < lo - 1e-12) or np.any(d > hi + 1e-12) or not np.all(np.isfinite(d)):
            continue
        v = 0.5 * d @ Qaa @ d + Qa @ d
        if v < bestv:
            best, bestv, bestfree = d, v, free
    if best is None:
        best, bestfree = np.clip(-Qa / np.maximum(np.diag(Qaa), 1e-12), lo, hi), []
    return best, bestfree


def ilqr(c, max_sweeps=60, tol=1e-9):
    name = c["system"]
    n, m = DIMS[name]
    P = PRESET[name]
    N, h = c["N"], c["pred_step"]
    b = np.array(P["bnds"], dtype=float)
    s = oracle.make_sys(name, P["pars"], P["bnds"])
    ct = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=N, pred_step_size=h, gamma=c["gamma"], critic_struct=c["critic_struct"],
                          R1=np.array(c["R1"]), observation_target=c["target"])
    w = c["w"] if c["mode"] != "MPC" else None
    obs0, x0 = np.array(c["obs"]), np.array(c["state_sys"])
    cost = lambda U: oracle.actor_cost(ct, s, U.reshape(-1), obs0, x0, w)          # noqa: E731
    quad = [stage_quadratic(c, n, m, k, N) for k in range(N)]
    U = np.clip(np.array(c["x_init"]).reshape(N, m), b[:, 0], b[:, 1])
    J = cost(U)
    mu, sweeps, evals = 1e-6, 0, 1
    lo_t, hi_t = np.tile(b[:, 0], N), np.tile(b[:, 1], N)

    def pg_norm(Uc):
        _, g = oracle.actor_grad(ct, s, Uc.reshape(-1), obs0, x0, w)
        u = Uc.reshape(-1)
        return np.max(np.abs(np.clip(u - g, lo_t, hi_t) - u))
    if pg_norm(U) <= 1e-7:                      # the start is already a stationary point of the box problem
        return U, J, 0, evals
    stalls = 0
    for sweeps in range(1, max_sweeps + 1):
        X = [x0]
        for k in range(N - 1):
            X.append(X[-1] + h * dyn(name, P["pars"], X[-1], U[k]))
        while True:
            Vx, Vxx = np.zeros(n), np.zeros((n, n))
            kff, Kfb = [None] * N, [None] * N
            ok = True
            for k in range(N - 1, -1, -1):
                H, g0, shift = quad[k]
                ob = obs0 if k == 0 else X[k]
                z = np.concatenate([ob - shift, U[k]])
                g = H @ z + g0
                cx, ca = (g[:n] if k > 0 else np.zeros(n)), g[n:]
                cxx = H[:n, :n] if k > 0 else np.zeros((n, n))
                cax = H[n:, :n] if k > 0 else np.zeros((m, n))
                caa = H[n:, n:]
                if k < N - 1:
                    fx, fa = dyn_jac(name, P["pars"], X[k], U[k])
                    A, B = np.eye(n) + h * fx, h * fa
                    Qx, Qa = cx + A.T @ Vx, ca + B.T @ Vx
                    Qxx, Qax, Qaa = cxx + A.T @ Vxx @ A, cax + B.T @ Vxx @ A, caa + B.T @ Vxx @ B
                else:
                    Qx, Qa, Qxx, Qax, Qaa = cx, ca, cxx, cax, caa
                Qaa_r = Qaa + mu * np.diag((b[:, 1] - b[:, 0]) ** -2 if False else np.ones(m))
                if np.any(np.linalg.eigvalsh(0.5 * (Qaa_r + Qaa_r.T)) <= 0):
                    ok = False
                    break
                d, free = box_newton(Qaa_r, Qa, b[:, 0] - U[k], b[:, 1] - U[k])
                K = np.zeros((m, n))
                if free:
                    K[free] = -np.linalg.solve(Qaa_r[np.ix_(free, free)], Qax[free])
                kff[k], Kfb[k] = d, K
                Vx = Qx + K.T @ Qaa @ d + K.T @ Qa + Qax.T @ d
                Vxx = Qxx + K.T @ Qaa @ K + K.T @ Qax + Qax.T @ K
                Vxx = 0.5 * (Vxx + Vxx.T)
            if ok:
                break
            mu = max(mu * 10, 1e-6)
            if mu > 1e12:
                return U, J, sweeps, evals
        alpha, improved = 1.0, False
        for _ in range(12):
            Un, x = U.copy(), x0.copy()
            for k in range(N):
                Un[k] = np.clip(U[k] + alpha * kff[k] + Kfb[k] @ (x - X[k]), b[:, 0], b[:, 1])
                if k < N - 1:
                    x = x + h * dyn(name, P["pars"], x, Un[k])
            Jn = cost(Un)
            evals += 1
            if Jn < J:
                improved = True
                break
            alpha *= 0.5
        if not improved:
            mu *= 10
            stalls += 1
            if mu > 1e12 or stalls >= 4:         # hand over (the hybrid continues with L-BFGS from here)
                break
            continue
        stalls = 0
        dJ = J - Jn
        U, J = Un, Jn
        mu = max(mu / 10, 1e-9)
        if dJ <= tol * max(abs(J), 1.0):
            break
    return U, J, sweeps, evals


def main():
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "actor_opt.json")))
    worse_i = worse_l = 0
    tot_i = tot_l = 0
    hyb = []
    for c in cases:
        name = c["system"]
        n, m = DIMS[name]
        P = PRESET[name]
        U, Ji, sw, ev = ilqr(c, max_sweeps=25)
        s = oracle.make_sys(name, P["pars"], P["bnds"])
        ct = oracle.make_ctrl(n, m, mode=c["mode"], Nactor=c["N"], pred_step_size=c["pred_step"], gamma=c["gamma"],
                              critic_struct=c["critic_struct"], R1=np.array(c["R1"]), observation_target=c["target"])
        # hybrid: polish / rescue with the L-BFGS iteration from where the sweeps stopped
        _, Jh, ith, _ = oracle.actor_opt(ct, s, U.reshape(-1), c["obs"], c["state_sys"], c["w"] if c["mode"] != "MPC" else None,
                                         max_iter=300, pg_tol=1e-7, f_tol=1e-12)
        hyb.append((sw, ith, Jh, c["J_ref"]))
        _, Jl, itl, nfl = oracle.actor_opt(ct, s, c["x_init"], c["obs"], c["state_sys"], c["w"] if c["mode"] != "MPC" else None,
                                           max_iter=300, pg_tol=1e-7, f_tol=1e-12)
        tol = 1e-7 * max(abs(c["J_ref"]), 1.0)
        worse_i += Ji > c["J_ref"] + tol
        worse_l += Jl > c["J_ref"] + tol
        tot_i += sw
        tot_l += itl
        flag = "" if Ji <= c["J_ref"] + tol else "  <-- iLQR above SLSQP"
        print(f"{name:9s} {c['mode']} {c['critic_struct']:10s} N={c['N']:2d} J_slsqp={c['J_ref']:14.8g} J_ilqr={Ji:14.8g} "
              f"({sw:3d} sweeps) J_lbfgs={Jl:14.8g} ({itl:3d} it){flag}")
    print(f"iLQR: {worse_i} of {len(cases)} above SLSQP, {tot_i} sweeps in total; L-BFGS: {worse_l} above, {tot_l} iterations")
    wh = sum(Jh > Jr + 1e-7 * max(abs(Jr), 1.0) for _, _, Jh, Jr in hyb)
    print(f"hybrid (<= 25 sweeps, then L-BFGS from there): {wh} above SLSQP, {sum(h[0] for h in hyb)} sweeps + "
          f"{sum(h[1] for h in hyb)} L-BFGS iterations; max L-BFGS iterations {max(h[1] for h in hyb)}")


if __name__ == "__main__":
    main()
