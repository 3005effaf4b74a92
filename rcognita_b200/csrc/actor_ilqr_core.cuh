// actor_ilqr_core.cuh -- control-limited Gauss-Newton (iLQR) sweeps over one _actor_cost problem
// (rcognita/controllers.py:1273-1328): the pre-pass of the actor optimiser for long horizons.
//
// Every stage term of _actor_cost is exactly quadratic in z = [observation - shift, action]:
//   stage_obj 'quadratic' (controllers.py:1063-1084): gamma**k * z^T R1 z;
//   the four critic structures (controllers.py:1192-1214): a quadratic form in chi (the raw observation for
//   'quad-mix'), plus a linear part for 'quad-lin'.
// A reverse Riccati pass over the horizon with the linearised Euler predictor (controllers.py:1296-1306,
// systems.py:308-323, :370-382, :412-419) therefore gives a Newton-like step from O(n^2) state per problem instead of
// the O(N m) vectors of a quasi-Newton memory (Sys3WRobot adds the second-order term of its heading/speed coupling: DDP).
// The box on the actions is handled per stage by a clamped Newton step
// (m <= 2: the 3^m active sets are enumerated in closed form); Q_aa is regularised (Levenberg-Marquardt) until it is
// positive definite at every stage; the forward pass backtracks on the cost itself.  A start that already passes the
// projected-gradient test costs one reverse pass; four failed forward passes in a row (the indefinite critics) end the
// sweeps.  rcg_actor_opt then iterates from the returned sequence.
//
// One function, __host__ __device__: the kernel in actor_ilqr.cu runs it with one thread per problem on strided global
// storage; tests/hostcheck/ compiles the same function for the host so that its decisions are checked without a GPU.
// 'biquadratic' stage costs are not quadratic in z: the function returns without touching the sequence.
#pragma once

#include "rcg_device.cuh"

namespace rcg {

template <int SYS>
__host__ __device__ inline void ilqr_dyn(const SysDev<double> &S, const double *x, const double *a, double *dx)
{
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {
        double sn, cs;
        sincos(x[2], &sn, &cs);
        dx[0] = a[0] * cs; dx[1] = a[0] * sn; dx[2] = a[1];
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {
        double sn, cs;
        sincos(x[2], &sn, &cs);
        dx[0] = x[3] * cs; dx[1] = x[3] * sn; dx[2] = x[4];
        dx[3] = (1.0 / S.pars[0]) * a[0]; dx[4] = (1.0 / S.pars[1]) * a[1];
    } else {
        const double t1 = S.pars[0], t2 = S.pars[1], K1 = S.pars[2], K2 = S.pars[3], K3 = S.pars[4];
        dx[0] = (1.0 / t1) * (-x[0] + K1 * a[0]);
        dx[1] = (1.0 / t2) * (-x[1] + K2 * x[0] + K3 * x[1] * x[1]);
    }
}

// A = I + h df/dx [n][n], B = h df/da [n][m] at (x, a)
template <int SYS>
__host__ __device__ inline void ilqr_lin(const SysDev<double> &S, double h, const double *x, const double *a, double *A,
                                         double *B)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    for (int i = 0; i < N * N; ++i) A[i] = 0.0;
    for (int i = 0; i < N * M; ++i) B[i] = 0.0;
    for (int i = 0; i < N; ++i) A[i * N + i] = 1.0;
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {
        double sn, cs;
        sincos(x[2], &sn, &cs);
        A[0 * N + 2] = -h * a[0] * sn; A[1 * N + 2] = h * a[0] * cs;
        B[0 * M + 0] = h * cs; B[1 * M + 0] = h * sn; B[2 * M + 1] = h;
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {
        double sn, cs;
        sincos(x[2], &sn, &cs);
        A[0 * N + 2] = -h * x[3] * sn; A[1 * N + 2] = h * x[3] * cs;
        A[0 * N + 3] = h * cs; A[1 * N + 3] = h * sn; A[2 * N + 4] = h;
        B[3 * M + 0] = h / S.pars[0]; B[4 * M + 1] = h / S.pars[1];
    } else {
        const double t1 = S.pars[0], t2 = S.pars[1], K1 = S.pars[2], K2 = S.pars[3], K3 = S.pars[4];
        A[0] += -h / t1; A[1 * N + 0] = h * K2 / t2; A[1 * N + 1] += h * (-1.0 + 2.0 * K3 * x[1]) / t2;
        B[0] = h * K1 / t1;
    }
}

__host__ __device__ inline double ilqr_clip(double v, double lo, double hi)
{
    v = (v < lo) ? lo : v;
    return (v > hi) ? hi : v;
}

// argmin 1/2 d^T Q d + q^T d over lo <= d <= hi, Q positive definite, M <= 2: the best feasible stationary point over
// the 3^M active sets.  fr[j] = 1 where the minimiser is interior in component j.
template <int M>
__host__ __device__ inline void ilqr_box_newton(const double *Q, const double *q, const double *lo, const double *hi,
                                                double *d, int *fr)
{
    double bestv = INFINITY;
    bool found = false;
    constexpr int NPAT = (M == 1) ? 3 : 9;
    for (int pat = 0; pat < NPAT; ++pat) {
        int st[2] = {pat % 3, pat / 3};
        double t[2] = {0.0, 0.0};
        for (int j = 0; j < M; ++j) t[j] = (st[j] == 1) ? lo[j] : (st[j] == 2) ? hi[j] : 0.0;
        if constexpr (M == 1) {
            if (st[0] == 0) t[0] = -q[0] / Q[0];
        } else {
            if (st[0] == 0 && st[1] == 0) {
                const double det = Q[0] * Q[3] - Q[1] * Q[2];
                t[0] = (-q[0] * Q[3] + q[1] * Q[1]) / det;
                t[1] = (-q[1] * Q[0] + q[0] * Q[2]) / det;
            } else if (st[0] == 0) {
                t[0] = -(q[0] + Q[1] * t[1]) / Q[0];
            } else if (st[1] == 0) {
                t[1] = -(q[1] + Q[2] * t[0]) / Q[3];
            }
        }
        bool feas = true;
        for (int j = 0; j < M; ++j) feas = feas && isfinite(t[j]) && t[j] >= lo[j] - 1e-12 && t[j] <= hi[j] + 1e-12;
        if (!feas) continue;
        double v = 0.0;
        for (int i = 0; i < M; ++i) {
            v += q[i] * t[i];
            for (int j = 0; j < M; ++j) v += 0.5 * t[i] * Q[i * M + j] * t[j];
        }
        if (v < bestv) {
            bestv = v;
            found = true;
            for (int j = 0; j < M; ++j) { d[j] = t[j]; fr[j] = (st[j] == 0); }
        }
    }
    if (!found)
        for (int j = 0; j < M; ++j) {
            const double qq = Q[j * M + j] > 1e-12 ? Q[j * M + j] : 1e-12;
            d[j] = ilqr_clip(-q[j] / qq, lo[j], hi[j]);
            fr[j] = 0;
        }
}

constexpr int kIlqrMaxBacktracks = 12;
constexpr int kIlqrMaxStalls = 4;

// per-problem workspace in doubles: rollout [NA][n], feed-forward [NA][m], gains [NA][m][n], trial sequence [NA][m]
__host__ __device__ inline int64_t ilqr_ws_per_problem(int na, int n, int m) { return (int64_t)na * (n + 2 * m + m * n); }

// Feeder concept (the source of problems; the loop below is PERSISTENT: a lane that finishes its problem asks for the
// next one, so the lanes of a warp stay busy although sweep counts differ between problems):
//   bool next(double *x0, double *ob0, double *w, double *&U, int64_t &us)  -- fetch the next problem of this lane: state,
//        observation, critic weights, U[i * us] = component i of its action sequence (in/out); false = no more work;
//   bool all_idle(bool idle)   -- true when every lane that shares this lane's control flow is out of work (warp vote);
//   void done(int sweeps)      -- the current problem is finished.
// ws[i * wss] = workspace double i of this lane (ilqr_ws_per_problem doubles).  The caller guarantees a 'quadratic' stage
// cost and max_sweeps >= 1.
template <int SYS, class Feeder>
__host__ __device__ inline void ilqr_run(const SysDev<double> &Sd, const ObjDev<double> &O, int mode, int cs_id, Feeder &feed,
                                         double *ws, int64_t wss, int max_sweeps, double pg_tol)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m, P = N + M;
    constexpr int DIMW = P * (P + 1) / 2 + P;
    const int NA = O.Nactor;
    const double h = O.pred_step_size;
    double x0[N], ob0[N], w[DIMW];
    double *U = nullptr;
    int64_t us = 0;

    double lo[M], hi[M];
    for (int j = 0; j < M; ++j) {
        lo[j] = Sd.has_bnds ? Sd.lo[j] : -INFINITY;
        hi[j] = Sd.has_bnds ? Sd.hi[j] : INFINITY;
    }
    // ---- the two quadratic forms of the horizon: stage objective (scaled by gamma**k) and critic ----
    double Hs[P * P], Hc[P * P], gc[P], shs[N], shc[N];
    for (int i = 0; i < P; ++i) {
        gc[i] = 0.0;
        for (int j = 0; j < P; ++j) { Hs[i * P + j] = O.R1[i * P + j] + O.R1[j * P + i]; Hc[i * P + j] = 0.0; }
    }
    for (int i = 0; i < N; ++i) { shs[i] = O.target[i]; shc[i] = O.target[i]; }
    auto critic_form = [&]() {                                 // per problem: the weights may differ between environments
        if (mode == RCG_MODE_MPC) return;
        for (int i = 0; i < P; ++i) {
            gc[i] = 0.0;
            for (int j = 0; j < P; ++j) Hc[i * P + j] = 0.0;
        }
        int k = 0;
        if (cs_id == RCG_CRITIC_QUAD_LIN || cs_id == RCG_CRITIC_QUADRATIC) {
            for (int i = 0; i < P; ++i)
                for (int j = i; j < P; ++j) { Hc[i * P + j] += w[k]; Hc[j * P + i] += w[k]; ++k; }
            if (cs_id == RCG_CRITIC_QUAD_LIN)
                for (int i = 0; i < P; ++i) gc[i] = w[k++];
        } else if (cs_id == RCG_CRITIC_QUAD_NOMIX) {
            for (int i = 0; i < P; ++i) Hc[i * P + i] = 2.0 * w[k++];
        } else {                                               // quad-mix: raw observation (controllers.py:1212)
            for (int i = 0; i < N; ++i) shc[i] = 0.0;
            for (int i = 0; i < N; ++i) Hc[i * P + i] = 2.0 * w[k++];
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < M; ++j) { Hc[i * P + N + j] += w[k]; Hc[(N + j) * P + i] += w[k]; ++k; }
            for (int j = 0; j < M; ++j) Hc[(N + j) * P + N + j] = 2.0 * w[k++];
        }
    };
    auto is_critic = [&](int k) { return mode == RCG_MODE_SQL || (mode == RCG_MODE_RQL && k == NA - 1); };
    // gradient (gz) and value of stage k at (ob, a)
    auto stage = [&](int k, const double *ob, const double *a, double *gz) -> double {
        const bool uc = is_critic(k);
        const double *H = uc ? Hc : Hs;
        const double *sh = uc ? shc : shs;
        const double gk = uc ? 1.0 : O.gamma_pow[k];
        double z[P];
        for (int i = 0; i < N; ++i) z[i] = ob[i] - sh[i];
        for (int j = 0; j < M; ++j) z[N + j] = a[j];
        double val = 0.0;
        for (int i = 0; i < P; ++i) {
            double acc = 0.0;
            for (int j = 0; j < P; ++j) acc += H[i * P + j] * z[j];
            acc *= gk;
            val += 0.5 * acc * z[i];
            if (uc) { acc += gc[i]; val += gc[i] * z[i]; }
            if (gz) gz[i] = acc;
        }
        return val;
    };

    // ---- strided storage ----
    auto Ux = [&](int i) -> double & { return U[(int64_t)i * us]; };
    auto Xr = [&](int k, int i) -> double & { return ws[((int64_t)k * N + i) * wss]; };
    auto kff = [&](int k, int j) -> double & { return ws[((int64_t)NA * N + k * M + j) * wss]; };
    auto Kfb = [&](int k, int j, int i) -> double & { return ws[((int64_t)NA * (N + M) + (k * M + j) * N + i) * wss]; };
    auto Un = [&](int i) -> double & { return ws[((int64_t)NA * (N + M + M * N) + i) * wss]; };

    // rollout of the current sequence: keeps the predictor states, returns the cost
    auto rollout = [&]() -> double {
        double st[N], dd[N], a[M], Jc = 0.0;
        for (int i = 0; i < N; ++i) st[i] = x0[i];
        for (int k = 0; k < NA; ++k) {
            for (int j = 0; j < M; ++j) a[j] = Ux(k * M + j);
            for (int i = 0; i < N; ++i) Xr(k, i) = st[i];
            Jc += stage(k, (k == 0) ? ob0 : st, a, nullptr);
            if (k + 1 < NA) {
                ilqr_dyn<SYS>(Sd, st, a, dd);
                for (int i = 0; i < N; ++i) st[i] = st[i] + h * dd[i];
            }
        }
        return Jc;
    };

    double J = 0.0, mu = 1e-6;
    int sweeps = 0, stalls = 0;
    bool need = true, idle = false;
    for (;;) {
        if (need) {
            need = false;
            if (!feed.next(x0, ob0, w, U, us)) {
                idle = true;
            } else {
                critic_form();
                for (int i = 0; i < NA * M; ++i) Ux(i) = ilqr_clip(Ux(i), lo[i % M], hi[i % M]);
                J = rollout();
                mu = 1e-6;
                sweeps = 0;
                stalls = 0;
            }
        }
        if (feed.all_idle(idle)) break;
        if (idle) continue;
        // ---- one pass = one sweep of the current problem ----
        bool fin = !(sweeps < max_sweeps);
        if (!fin) {
        // ---- reverse pass: gains and feed-forward steps; raise mu until every Q_aa is positive definite ----
        bool ok = false;
        double pg = 0.0;
        while (!ok) {
            double Vx[N], Vxx[N * N], lam[N];
            for (int i = 0; i < N; ++i) { Vx[i] = 0.0; lam[i] = 0.0; }
            for (int i = 0; i < N * N; ++i) Vxx[i] = 0.0;
            ok = true;
            pg = 0.0;
            for (int k = NA - 1; k >= 0; --k) {
                const bool uc = is_critic(k);
                const double *H = uc ? Hc : Hs;
                const double gk = uc ? 1.0 : O.gamma_pow[k];
                double xk[N], a[M], gz[P];
                for (int i = 0; i < N; ++i) xk[i] = Xr(k, i);
                for (int j = 0; j < M; ++j) a[j] = Ux(k * M + j);
                stage(k, (k == 0) ? ob0 : xk, a, gz);
                double Qx[N], Qa[M], Qxx[N * N], Qax[M * N], Qaa[M * M], ga[M], ln[N];
                for (int i = 0; i < N; ++i) {
                    Qx[i] = (k > 0) ? gz[i] : 0.0;
                    ln[i] = Qx[i];
                    for (int j = 0; j < N; ++j) Qxx[i * N + j] = (k > 0) ? gk * H[i * P + j] : 0.0;
                }
                for (int j = 0; j < M; ++j) {
                    Qa[j] = gz[N + j];
                    ga[j] = gz[N + j];
                    for (int i = 0; i < N; ++i) Qax[j * N + i] = (k > 0) ? gk * H[(N + j) * P + i] : 0.0;
                    for (int l = 0; l < M; ++l) Qaa[j * M + l] = gk * H[(N + j) * P + N + l];
                }
                if (k < NA - 1) {
                    double A[N * N], B[N * M], VA[N * N], VB[N * M];
                    ilqr_lin<SYS>(Sd, h, xk, a, A, B);
                    for (int i = 0; i < N; ++i) {
                        for (int j = 0; j < N; ++j) {
                            double acc = 0.0;
                            for (int l = 0; l < N; ++l) acc += Vxx[i * N + l] * A[l * N + j];
                            VA[i * N + j] = acc;
                        }
                        for (int j = 0; j < M; ++j) {
                            double acc = 0.0;
                            for (int l = 0; l < N; ++l) acc += Vxx[i * N + l] * B[l * M + j];
                            VB[i * M + j] = acc;
                        }
                    }
                    for (int i = 0; i < N; ++i) {
                        for (int l = 0; l < N; ++l) { Qx[i] += A[l * N + i] * Vx[l]; ln[i] += A[l * N + i] * lam[l]; }
                        for (int j = 0; j < N; ++j)
                            for (int l = 0; l < N; ++l) Qxx[i * N + j] += A[l * N + i] * VA[l * N + j];
                    }
                    if constexpr (SYS == RCG_SYS_3WROBOT) {
                        // second-order term of the predictor (DDP): V_x' . d2f/dx2 for the heading/speed coupling
                        // x' = x + h v cos(theta), y' = y + h v sin(theta).  Without it the linear model of a saturated
                        // manoeuvre is valid for a quarter step only (measured: alpha = 1/4 typical, 141 quasi-Newton
                        // iterations left after 25 sweeps; with it 9 after 28).
                        double sn, cs;
                        sincos(xk[2], &sn, &cs);
                        Qxx[2 * N + 2] += h * (Vx[0] * (-xk[3] * cs) + Vx[1] * (-xk[3] * sn));
                        const double c = h * (Vx[0] * (-sn) + Vx[1] * cs);
                        Qxx[2 * N + 3] += c;
                        Qxx[3 * N + 2] += c;
                    }
                    for (int j = 0; j < M; ++j) {
                        for (int l = 0; l < N; ++l) { Qa[j] += B[l * M + j] * Vx[l]; ga[j] += B[l * M + j] * lam[l]; }
                        for (int i = 0; i < N; ++i)
                            for (int l = 0; l < N; ++l) Qax[j * N + i] += B[l * M + j] * VA[l * N + i];
                        for (int q = 0; q < M; ++q)
                            for (int l = 0; l < N; ++l) Qaa[j * M + q] += B[l * M + j] * VB[l * M + q];
                    }
                }
                for (int i = 0; i < N; ++i) lam[i] = ln[i];
                for (int j = 0; j < M; ++j) {                      // projected-gradient norm of the current sequence
                    const double d = fabs(ilqr_clip(a[j] - ga[j], lo[j], hi[j]) - a[j]);
                    pg = (d > pg) ? d : pg;
                }
                double Qr[M * M];
                for (int j = 0; j < M * M; ++j) Qr[j] = Qaa[j];
                for (int j = 0; j < M; ++j) Qr[j * M + j] += mu;
                bool pd;
                if constexpr (M == 1) pd = Qr[0] > 0.0;
                else pd = Qr[0] > 0.0 && Qr[0] * Qr[3] - 0.25 * (Qr[1] + Qr[2]) * (Qr[1] + Qr[2]) > 0.0;
                if (!pd) { ok = false; break; }
                double dlo[M], dhi[M], d[M];
                int fr[M];
                for (int j = 0; j < M; ++j) { dlo[j] = lo[j] - a[j]; dhi[j] = hi[j] - a[j]; }
                ilqr_box_newton<M>(Qr, Qa, dlo, dhi, d, fr);
                double K[M * N];
                for (int i = 0; i < M * N; ++i) K[i] = 0.0;
                if constexpr (M == 1) {
                    if (fr[0]) for (int i = 0; i < N; ++i) K[i] = -Qax[i] / Qr[0];
                } else {
                    if (fr[0] && fr[1]) {
                        const double det = Qr[0] * Qr[3] - Qr[1] * Qr[2];
                        for (int i = 0; i < N; ++i) {
                            K[0 * N + i] = -(Qr[3] * Qax[0 * N + i] - Qr[1] * Qax[1 * N + i]) / det;
                            K[1 * N + i] = -(Qr[0] * Qax[1 * N + i] - Qr[2] * Qax[0 * N + i]) / det;
                        }
                    } else if (fr[0]) {
                        for (int i = 0; i < N; ++i) K[0 * N + i] = -Qax[0 * N + i] / Qr[0];
                    } else if (fr[1]) {
                        for (int i = 0; i < N; ++i) K[1 * N + i] = -Qax[1 * N + i] / Qr[3];
                    }
                }
                for (int j = 0; j < M; ++j) {
                    kff(k, j) = d[j];
                    for (int i = 0; i < N; ++i) Kfb(k, j, i) = K[j * N + i];
                }
                // V_x = Q_x + K^T (Q_aa d + Q_a) + Q_ax^T d;  V_xx = Q_xx + K^T (Q_aa K + Q_ax) + Q_ax^T K, symmetrised
                double Qd[M], QK[M * N];
                for (int j = 0; j < M; ++j) {
                    Qd[j] = 0.0;
                    for (int i = 0; i < N; ++i) QK[j * N + i] = 0.0;
                    for (int q = 0; q < M; ++q) {
                        Qd[j] += Qaa[j * M + q] * d[q];
                        for (int i = 0; i < N; ++i) QK[j * N + i] += Qaa[j * M + q] * K[q * N + i];
                    }
                }
                for (int i = 0; i < N; ++i) {
                    double v = Qx[i];
                    for (int j = 0; j < M; ++j) v += K[j * N + i] * (Qd[j] + Qa[j]) + Qax[j * N + i] * d[j];
                    Vx[i] = v;
                }
                for (int i = 0; i < N; ++i)
                    for (int l = 0; l < N; ++l) {
                        double v = Qxx[i * N + l];
                        for (int j = 0; j < M; ++j)
                            v += K[j * N + i] * (QK[j * N + l] + Qax[j * N + l]) + Qax[j * N + i] * K[j * N + l];
                        Vxx[i * N + l] = v;
                    }
                for (int i = 0; i < N; ++i)
                    for (int l = i + 1; l < N; ++l) {
                        const double v = 0.5 * (Vxx[i * N + l] + Vxx[l * N + i]);
                        Vxx[i * N + l] = v;
                        Vxx[l * N + i] = v;
                    }
            }
            if (!ok) {
                mu = (mu * 10.0 > 1e-6) ? mu * 10.0 : 1e-6;
                if (mu > 1e12) break;
            }
        }
        if (!ok || !(pg > pg_tol)) fin = true;
        }
        if (!fin) {
        ++sweeps;
        // ---- forward pass with backtracking on the cost ----
        double alpha = 1.0, Jn = J;
        bool improved = false;
        for (int bt = 0; bt < kIlqrMaxBacktracks && !improved; ++bt) {
            double st[N], dd[N], a[M];
            for (int i = 0; i < N; ++i) st[i] = x0[i];
            Jn = 0.0;
            for (int k = 0; k < NA; ++k) {
                for (int j = 0; j < M; ++j) {
                    double v = Ux(k * M + j) + alpha * kff(k, j);
                    for (int i = 0; i < N; ++i) v += Kfb(k, j, i) * (st[i] - Xr(k, i));
                    a[j] = ilqr_clip(v, lo[j], hi[j]);
                    Un(k * M + j) = a[j];
                }
                Jn += stage(k, (k == 0) ? ob0 : st, a, nullptr);
                if (k + 1 < NA) {
                    ilqr_dyn<SYS>(Sd, st, a, dd);
                    for (int i = 0; i < N; ++i) st[i] = st[i] + h * dd[i];
                }
            }
            if (Jn < J) improved = true;
            else alpha *= 0.5;
        }
        if (!improved) {
            mu *= 10.0;
            if (mu > 1e12 || ++stalls >= kIlqrMaxStalls) fin = true;
        } else {
            stalls = 0;
            const double dJ = J - Jn;
            for (int i = 0; i < NA * M; ++i) Ux(i) = Un(i);
            J = rollout();                                         // refresh the stored rollout for the next reverse pass
            mu = (mu / 10.0 > 1e-9) ? mu / 10.0 : 1e-9;
            const double scale = (fabs(J) > 1.0) ? fabs(J) : 1.0;
            if (dJ <= 1e-9 * scale) fin = true;
        }
        }
        if (fin) {
            feed.done(sweeps);
            need = true;
        }
    }
}

// One problem on the calling thread (host checks; E x S = 1).  Returns the number of sweeps.
struct IlqrSingleFeeder {
    const double *x0, *ob0, *w;
    double *U;
    int64_t us;
    int n, dimw, sweeps;
    bool taken;
    __host__ __device__ bool next(double *x, double *ob, double *ww, double *&Uo, int64_t &uso)
    {
        if (taken) return false;
        taken = true;
        for (int i = 0; i < n; ++i) { x[i] = x0[i]; ob[i] = ob0[i]; }
        for (int i = 0; i < dimw; ++i) ww[i] = w[i];
        Uo = U;
        uso = us;
        return true;
    }
    __host__ __device__ bool all_idle(bool idle) const { return idle; }
    __host__ __device__ void done(int s) { sweeps = s; }
};

__host__ __device__ inline int ilqr_dim_critic(int mode, int cs, int n, int m)
{
    const int p = n + m;
    if (mode == RCG_MODE_MPC) return 0;
    switch (cs) {
    case RCG_CRITIC_QUAD_LIN:   return p * (p + 1) / 2 + p;
    case RCG_CRITIC_QUADRATIC:  return p * (p + 1) / 2;
    case RCG_CRITIC_QUAD_NOMIX: return p;
    default:                    return n + n * m + m;
    }
}

template <int SYS>
__host__ __device__ inline int ilqr_presweeps(const SysDev<double> &Sd, const ObjDev<double> &O, int mode, int cs_id,
                                              const double *x0, const double *ob0, const double *w, double *U, int64_t us,
                                              double *ws, int64_t wss, int max_sweeps, double pg_tol)
{
    if (O.stage_struct != RCG_STAGE_QUADRATIC || max_sweeps <= 0) return 0;
    IlqrSingleFeeder f{x0, ob0, w, U, us, SysDim<SYS>::n, ilqr_dim_critic(mode, cs_id, SysDim<SYS>::n, SysDim<SYS>::m), 0, false};
    ilqr_run<SYS>(Sd, O, mode, cs_id, f, ws, wss, max_sweeps, pg_tol);
    return f.sweeps;
}

}  // namespace rcg
