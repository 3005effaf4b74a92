#!/usr/bin/env python3
"""Offline numpy prototype of the critic fit (rcognita_b200/csrc/critic_fit.cu, K = 3 fast path) on the in-loop problems
dumped by tools/dump_critic_problems.py (gpurun_out/critic_problems.npz): reproduces the kernel's fitted costs, counts dual
evaluations per problem and per continuation stage, and compares variants (exact line search, stopping at the first
non-improving stage, warm-started Armijo step).  Analysis tool, not product code.  Usage: tools/critic_fit_prototype.py [key]"""
import numpy as np, sys, time
d = np.load('gpurun_out/critic_problems.npz')
N, M, P = 5, 2, 7
R1 = np.array([1, 10, 1, 0, 0, 0, 0.0])
def phi(o, a):
    chi = np.concatenate([o, a]); return np.array([chi[i]*chi[j] for i in range(P) for j in range(i, P)])
def problem(obs_buf, act_buf, w_prev, e, K=3, gamma=1.0):
    Phi = np.zeros((K, 28)); b = np.zeros(K)
    for r in range(K):
        k = K - r
        op, on, ap, an = obs_buf[k-1,:,e], obs_buf[k,:,e], act_buf[k-1,:,e], act_buf[k,:,e]
        Phi[r] = phi(op, ap)
        chi = np.concatenate([op, ap])
        b[r] = gamma * (phi(on, an) @ w_prev[:, e]) + (chi*R1) @ chi
    return Phi, b
def fit(Phi, b, lo=0.0, hi=1e3, mode="armijo", max_ls=40, max_newton=20, max_outer=5, mu_rel=1e-3, stop_on_fail=False, warm_ls=False):
    K, D = Phi.shape
    clip = lambda z: np.minimum(np.maximum(z, lo), hi)
    w0 = clip(np.ones(D)); wb = w0.copy()
    cost = lambda w: 0.5*np.sum((Phi@w - b)**2)
    Jbest = cost(w0); J0 = Jbest
    trace = np.sum(Phi*Phi); bb = b@b
    evals = 0; newtons = 0; ls_total = 0
    stats = []
    if not (trace > 0 and np.isfinite(trace) and np.isfinite(bb)): return wb, Jbest, evals, stats
    for outer in range(max_outer):
        mu = mu_rel*trace/K; mu_rel *= 1e-2
        lam = np.zeros(K)
        nn = 0; nls = 0
        a_prev = 1.0
        def dual(l):
            z = w0 + Phi.T@l
            psi = np.where(z < lo, lo*z - 0.5*lo*lo, np.where(z > hi, hi*z - 0.5*hi*hi, 0.5*z*z))
            return 0.5*mu*(l@l) - b@l + psi.sum()
        for it in range(max_newton):
            evals += 1; nn += 1
            z = w0 + Phi.T@lam; w = clip(z)
            free = (z > lo) & (z < hi)
            F = mu*lam - b + Phi@w
            H = mu*np.eye(K) + (Phi[:, free] @ Phi[:, free].T)
            if np.max(np.abs(F)) <= 1e-13*np.sqrt(bb): break
            try: dl = -np.linalg.solve(H, F)
            except np.linalg.LinAlgError: break
            slope = F@dl
            if not (slope < 0): break
            D0 = dual(lam)
            patA = (z <= lo).astype(int) - (z >= hi).astype(int)
            if mode == "armijo":
                a = min(1.0, 4.0*a_prev) if warm_ls else 1.0
                ok = False; same = False
                for ls in range(max_ls):
                    evals += 1; nls += 1
                    lt = lam + a*dl
                    if dual(lt) <= D0 + 1e-4*a*slope + 1e-14*abs(D0):
                        ok = True; lam = lt; a_prev = a
                        zt = w0 + Phi.T@lt
                        same = (a == 1.0) and np.array_equal((zt <= lo).astype(int) - (zt >= hi).astype(int), patA)
                        break
                    a *= 0.5
                if (not ok) or same: break
            else:   # exact line search on the piecewise-quadratic dual along dl, a in [0, amax]
                s = Phi.T@dl
                def dphi(a):
                    zt = z + a*s; wt = clip(zt)
                    return mu*((lam + a*dl)@dl) - b@dl + s@wt, mu*(dl@dl) + np.sum(s[(zt > lo) & (zt < hi)]**2)
                a = 1.0; alo, ahi = 0.0, None
                for k in range(30):
                    evals += 1; nls += 1
                    g1, g2 = dphi(a)
                    if abs(g1) <= 1e-12*abs(slope): break
                    if g1 > 0: ahi = a
                    else: alo = a
                    an = a - g1/g2
                    if ahi is not None and not (alo < an < ahi): an = 0.5*(alo + ahi)
                    elif ahi is None and an <= alo: an = 2*a
                    if abs(an - a) <= 1e-15*max(a, 1e-300): a = an; break
                    a = an
                lt = lam + a*dl
                zt = w0 + Phi.T@lt
                same = np.array_equal((zt <= lo).astype(int) - (zt >= hi).astype(int), patA) and abs(a - 1.0) < 1e-12
                lam = lt
                if same: break
        wn = clip(w0 + Phi.T@lam); Jn = cost(wn)
        stats.append((nn, nls, Jn))
        if Jn < Jbest:
            Jbest = Jn; w0 = wn.copy(); wb = wn.copy()
            if Jn <= 1e-12*J0 or Jn <= 1e-20*bb: break
        elif stop_on_fail: break
    return wb, Jbest, evals, stats
if __name__ == "__main__":
    key = sys.argv[1] if len(sys.argv) > 1 else "config3_k20"
    ob, ab, wp, Jc, fl = d[key+"_obs_buf"], d[key+"_act_buf"], d[key+"_w_prev"], d[key+"_Jc"], d[key+"_flag"]
    E = ob.shape[2]; lanes = np.flatnonzero(fl)[:400]
    res = {}
    for mode in ("armijo", "exact"):
        ev = []; Js = []
        t0 = time.time()
        for e in lanes:
            Phi, b = problem(ob, ab, wp, e)
            w, J, evals, stats = fit(Phi, b, mode=mode)
            ev.append(evals); Js.append(J)
        ev = np.array(ev); Js = np.array(Js)
        res[mode] = (ev, Js)
        print(key, mode, "lanes", len(lanes), "evals mean %.1f p50 %d p90 %d p99 %d max %d" % (ev.mean(), np.median(ev), np.percentile(ev, 90), np.percentile(ev, 99), ev.max()), "sum", ev.sum(), "time %.1fs" % (time.time()-t0))
        if mode == "armijo":
            rel = np.abs(Js - Jc[lanes]) / np.maximum(np.abs(Jc[lanes]), 1e-300)
            print("   proto vs GPU Jc: median rel diff %.2e, frac within 1e-6: %.3f" % (np.median(rel), np.mean(rel < 1e-6)))
    a, x = res["armijo"], res["exact"]
    ratio = x[1] / np.maximum(a[1], 1e-300)
    print("   exact/armijo cost ratio: median %.6f p1 %.6f p99 %.6f; exact better-or-equal(1e-9): %.3f" % (np.median(ratio), np.percentile(ratio, 1), np.percentile(ratio, 99), np.mean(x[1] <= a[1]*(1+1e-9))))

def analyse(key="config3_k40", top=6):
    ob, ab, wp, Jc, fl = d[key+"_obs_buf"], d[key+"_act_buf"], d[key+"_w_prev"], d[key+"_Jc"], d[key+"_flag"]
    lanes = np.flatnonzero(fl)[:400]
    out = []
    for e in lanes:
        Phi, b = problem(ob, ab, wp, e)
        w, J, evals, stats = fit(Phi, b)
        out.append((evals, e, stats, J, np.sum(b*b)*0.5, (w <= 0).sum(), (w >= 1e3).sum()))
    out.sort(key=lambda x: -x[0])
    for evals, e, stats, J, Jzero, nlo, nhi in out[:top]:
        print("lane", e, "evals", evals, "J", "%.4g" % J, "0.5|b|^2 %.4g" % Jzero, "at lo", nlo, "at hi", nhi)
        for s_ in stats: print("    newton %2d  ls %3d  Jn %.6g" % s_)
    tot = sum(o[0] for o in out)
    print("share of evals in top 5% lanes:", sum(o[0] for o in out[:len(out)//20]) / tot)
    # where do evals go overall: newton passes vs line-search passes
    nn = sum(s_[0] for o in out for s_ in o[2]); nl = sum(s_[1] for o in out for s_ in o[2])
    print("newton passes", nn, "line-search passes", nl)
