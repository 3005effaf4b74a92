/*
 * rcg_oracle_opt.c -- CPU ORACLE (test infrastructure, NOT product code), part 2:
 * the analytic gradient of CtrlOptPred._actor_cost and the bounded minimiser that stands in for
 * CtrlOptPred._actor_optimizer (ref: rcognita/controllers.py:1330-1427, scipy SLSQP with
 * Bounds(action_sqn_min, action_sqn_max), tol 1e-7, maxiter 300).
 *
 * The reference differentiates _actor_cost by finite differences inside SLSQP; here the gradient is
 * the exact adjoint (reverse sweep) of the Euler rollout of controllers.py:1290-1296, and the
 * minimiser is a projected limited-memory quasi-Newton method (L-BFGS two-loop recursion on the free
 * variables, projected Armijo backtracking).  This is a restatement
 * of the ALGORITHM the CUDA path implements (rcognita_b200/csrc/actor_opt_impl.cuh), scalar and in
 * the reference's row-major [Nactor, m] layout; it is pinned two ways in tests/test_oracle_golden.py:
 * the gradient against central differences of orc_actor_cost (itself pinned to the live reference),
 * and the minimum against the live reference's SLSQP results (tests/golden/actor_opt.json).
 */
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rcg_oracle.h"

static double clipd(double v, double lo, double hi)
{
    v = (v < lo) ? lo : v;
    return (v > hi) ? hi : v;
}

/* d stage_obj / d chi (ref: controllers.py:1063-1084): quadratic chi (R1 + R1^T);
 * biquadratic adds 2 chi_i [(R2 + R2^T) chi^2]_i. */
static void stage_obj_grad(const orc_ctrl_t *c, int p, const double *chi, double *gchi)
{
    for (int i = 0; i < p; ++i) {
        double acc = 0.0;
        for (int j = 0; j < p; ++j) acc += (c->R1[i * p + j] + c->R1[j * p + i]) * chi[j];
        gchi[i] = acc;
    }
    if (c->stage_struct == ORC_STAGE_BIQUADRATIC) {
        for (int i = 0; i < p; ++i) {
            double acc = 0.0;
            for (int j = 0; j < p; ++j) acc += (c->R2[i * p + j] + c->R2[j * p + i]) * (chi[j] * chi[j]);
            gchi[i] += 2.0 * chi[i] * acc;
        }
    }
}

/* d _critic / d (obs, act) (ref: controllers.py:1192-1214; feature order of orc_critic). */
static void critic_grad(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act, const double *w,
                        double *gobs, double *gact)
{
    const int p = n + m;
    double chi[ORC_MAX_P], g[ORC_MAX_P];
    int k = 0;
    for (int i = 0; i < n; ++i) chi[i] = c->has_target ? obs[i] - c->target[i] : obs[i];
    for (int j = 0; j < m; ++j) chi[n + j] = act[j];
    for (int i = 0; i < p; ++i) g[i] = 0.0;
    switch (c->critic_struct) {
    case ORC_CRITIC_QUAD_LIN:
    case ORC_CRITIC_QUADRATIC:
        for (int i = 0; i < p; ++i)
            for (int j = i; j < p; ++j) {
                g[i] += w[k] * chi[j];
                g[j] += w[k] * chi[i];
                ++k;
            }
        if (c->critic_struct == ORC_CRITIC_QUAD_LIN)
            for (int i = 0; i < p; ++i) g[i] += w[k++];
        break;
    case ORC_CRITIC_QUAD_NOMIX:
        for (int i = 0; i < p; ++i) g[i] = 2.0 * w[k++] * chi[i];
        break;
    case ORC_CRITIC_QUAD_MIX:                       /* raw observation (:1212) */
        for (int i = 0; i < n; ++i) g[i] = 2.0 * w[k++] * obs[i];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) {
                g[i] += w[k] * act[j];
                g[n + j] += w[k] * obs[i];
                ++k;
            }
        for (int j = 0; j < m; ++j) g[n + j] += 2.0 * w[k++] * act[j];
        break;
    default:
        break;
    }
    for (int i = 0; i < n; ++i) gobs[i] = g[i];
    for (int j = 0; j < m; ++j) gact[j] = g[n + j];
}

/* lam <- lam + h * (d _state_dyn / d state)^T lam, and ga += h * (d _state_dyn / d action)^T lam, both at
 * (x, a) with lam = lam_{k+1} on entry (ref for the dynamics: systems.py:308-323, :370-382, :412-419). */
static void dyn_adjoint(const orc_sys_t *s, double h, const double *x, const double *a, double *lam, double *ga)
{
    if (s->sys_id == ORC_SYS_3WROBOT_NI) {
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        ga[0] += h * (cs * lam[0] + sn * lam[1]);
        ga[1] += h * lam[2];
        lam[2] += h * (a[0] * (cs * lam[1] - sn * lam[0]));
    } else if (s->sys_id == ORC_SYS_3WROBOT) {
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        ga[0] += h * ((1.0 / s->pars[0]) * lam[3]);
        ga[1] += h * ((1.0 / s->pars[1]) * lam[4]);
        const double l2 = lam[2] + h * (x[3] * (cs * lam[1] - sn * lam[0]));
        const double l3 = lam[3] + h * (cs * lam[0] + sn * lam[1]);
        const double l4 = lam[4] + h * lam[2];
        lam[2] = l2; lam[3] = l3; lam[4] = l4;
    } else {
        const double tau1 = s->pars[0], tau2 = s->pars[1], K1 = s->pars[2], K2 = s->pars[3], K3 = s->pars[4];
        ga[0] += h * ((1.0 / tau1) * K1 * lam[0]);
        const double l0 = lam[0] + h * (-(1.0 / tau1) * lam[0] + (1.0 / tau2) * K2 * lam[1]);
        const double l1 = lam[1] + h * ((1.0 / tau2) * (-1.0 + 2.0 * K3 * x[1]) * lam[1]);
        lam[0] = l0; lam[1] = l1;
    }
}

/* _actor_cost and its gradient w.r.t. action_sqn ([Nactor, m] row-major).  Returns the cost exactly as
 * orc_actor_cost computes it. */
double orc_actor_grad(const orc_ctrl_t *c, const orc_sys_t *s, const double *action_sqn, const double *observation,
                      const double *state_sys, const double *w_critic, double *grad)
{
    const int n = s->n, m = s->m, N = c->Nactor, p = n + m;
    const double h = c->pred_step_size;
    double X[ORC_MAX_NACTOR][ORC_MAX_N];       /* X[k] = predictor state before stage k's Euler step */
    double d[ORC_MAX_N], lam[ORC_MAX_N], chi[ORC_MAX_P], gchi[ORC_MAX_P], gobs[ORC_MAX_N], gact[ORC_MAX_M];
    for (int i = 0; i < n; ++i) X[0][i] = state_sys[i];
    for (int k = 1; k < N; ++k) {
        orc_state_dyn(s, X[k - 1], action_sqn + (k - 1) * m, d);
        for (int i = 0; i < n; ++i) X[k][i] = X[k - 1][i] + h * d[i];
    }
    const double J = orc_actor_cost(c, s, action_sqn, observation, state_sys, w_critic);
    for (int i = 0; i < n; ++i) lam[i] = 0.0;
    for (int k = N - 1; k >= 0; --k) {
        const double *obs = (k == 0) ? observation : X[k];      /* observation_sqn[0] = observation (:1290) */
        const double *a = action_sqn + k * m;
        double *ga = grad + k * m;
        /* lam currently holds lam_{k+1}: dJ/dX[k+1] (zero for k = N-1) */
        for (int j = 0; j < m; ++j) ga[j] = 0.0;
        if (k < N - 1) dyn_adjoint(s, h, X[k], a, lam, ga);     /* lam <- (I + h A_k)^T lam_{k+1} */
        const int use_critic = (c->mode == ORC_MODE_SQL) || (c->mode == ORC_MODE_RQL && k == N - 1);
        if (use_critic) {
            critic_grad(c, n, m, obs, a, w_critic, gobs, gact);
        } else {
            const double gk = pow(c->gamma, (double)k);
            for (int i = 0; i < n; ++i) chi[i] = c->has_target ? obs[i] - c->target[i] : obs[i];
            for (int j = 0; j < m; ++j) chi[n + j] = a[j];
            stage_obj_grad(c, p, chi, gchi);
            for (int i = 0; i < n; ++i) gobs[i] = gk * gchi[i];
            for (int j = 0; j < m; ++j) gact[j] = gk * gchi[n + j];
        }
        for (int j = 0; j < m; ++j) ga[j] += gact[j];
        if (k > 0)
            for (int i = 0; i < n; ++i) lam[i] += gobs[i];       /* stage 0 reads the fixed observation */
    }
    return J;
}

/* Projected limited-memory quasi-Newton minimisation of _actor_cost over the box [lo, hi]^Nactor (lo/hi = the
 * system's ctrl_bnds tiled like action_sqn_min/max, ref: controllers.py:968-971; unbounded if !has_bnds).
 * x [Nactor*m]: start point in, minimiser out.  Returns the cost at the returned point; *iters_out = accepted
 * iterations (= gradient evaluations after the first), *nfev_out = cost evaluations of the line searches.
 *
 *   x <- P(x); (J, g) at x
 *   repeat: binding set B = { i : (x_i <= lo_i and g_i > 0) or (x_i >= hi_i and g_i < 0) }, free set F = rest
 *           stop if |P(x - g) - x|_inf <= pg_tol
 *           d_F = -H g_F by the L-BFGS two-loop recursion over the last ORC_OPT_MEM pairs (s, y), every inner
 *                 product taken over F only (pairs with (s.y)_F <= 1e-10 |s|_F |y|_F are skipped), initial scaling
 *                 (s.y)_F / (y.y)_F of the newest usable pair (w / |P(x - g) - x|_inf when there is none, w = the
 *                 widest side of the box, 1 if unbounded: without curvature information the first trial step
 *                 spans the box and the backtracking finds the scale); d_B = 0
 *           if g.d >= 0: forget the pairs, d_F = -g_F w / |P(x - g) - x|_inf
 *           projected backtracking: x+ = P(x + lambda d), lambda = 1, 1/2, ... until
 *                 J(x+) <= J + 1e-4 min(g.(x+ - x), 0)              (monotone: the last iterate is the best)
 *           remember s = x+ - x, y = g+ - g
 *   stop also when the cost moved by <= f_tol * max(|J|, 1) twice in a row, when the line search fails, or after
 *   max_iter iterations. */
#define ORC_OPT_MEM 6
#define ORC_OPT_LMAX (ORC_MAX_NACTOR * ORC_MAX_M)

/* Summation order of the inner products.  rcg_actor_opt has two kernels behind it that run this same iteration: one lane
 * per problem (sums in index order: lanes = 1, the default here) and G = 4 lanes per problem (actor_opt_quad.cuh: lane r
 * sums the components r, r + G, ... in order, the G partial sums are combined by an xor-butterfly, i.e. pairwise).  The
 * minimiser of a flat problem moves with the rounding of these sums; orc_actor_opt_set_lanes lets the tests measure by
 * how much on the checker itself. */
static _Thread_local int orc_opt_lanes = 1;
void orc_actor_opt_set_lanes(int lanes) { orc_opt_lanes = (lanes == 2 || lanes == 4 || lanes == 8) ? lanes : 1; }

static double orc_sum_free(const double *u, const double *v, const int *fr, int L)
{
    const int G = orc_opt_lanes;
    double part[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < L; ++i)
        if (!fr || fr[i]) part[i % G] += u[i] * v[i];
    for (int st = 1; st < G; st <<= 1)
        for (int r = 0; r < G; r += 2 * st) part[r] += part[r + st];
    return part[0];
}

double orc_actor_opt(const orc_ctrl_t *c, const orc_sys_t *s, double *x, const double *observation,
                     const double *state_sys, const double *w_critic, int max_iter, double pg_tol, double f_tol,
                     int *iters_out, int *nfev_out)
{
    const int m = s->m, L = c->Nactor * m;
    double lo[ORC_MAX_M], hi[ORC_MAX_M];
    double g[ORC_OPT_LMAX], gn[ORC_OPT_LMAX], d[ORC_OPT_LMAX], xt[ORC_OPT_LMAX];
    static _Thread_local double S[ORC_OPT_MEM][ORC_OPT_LMAX], Y[ORC_OPT_MEM][ORC_OPT_LMAX];
    double al[ORC_OPT_MEM], sy[ORC_OPT_MEM];
    int fr[ORC_OPT_LMAX];
    int npairs = 0, head = 0;                   /* pairs live in slots (head - 1 - j) mod MEM, j = 0 newest */
    int iters = 0, nfev = 0, stall = 0;
    for (int j = 0; j < m; ++j) {
        lo[j] = s->has_bnds ? s->lo[j] : -INFINITY;
        hi[j] = s->has_bnds ? s->hi[j] : INFINITY;
    }
    /* first-step length when there is no curvature information: across the widest side of the box */
    double step0 = 1.0;
    if (s->has_bnds) {
        step0 = 0.0;
        for (int j = 0; j < m; ++j) step0 = fmax(step0, hi[j] - lo[j]);
    }
    for (int i = 0; i < L; ++i) x[i] = clipd(x[i], lo[i % m], hi[i % m]);
    double J = orc_actor_grad(c, s, x, observation, state_sys, w_critic, g);
    while (iters < max_iter) {
        double pg = 0.0;
        for (int i = 0; i < L; ++i) {
            const double l = lo[i % m], u = hi[i % m];
            fr[i] = !((x[i] <= l && g[i] > 0.0) || (x[i] >= u && g[i] < 0.0));
            pg = fmax(pg, fabs(clipd(x[i] - g[i], l, u) - x[i]));
        }
        if (!(pg > pg_tol)) break;
        /* two-loop recursion on the free set */
        for (int i = 0; i < L; ++i) d[i] = fr[i] ? g[i] : 0.0;
        double scale = step0 / pg;
        int have_scale = 0;
        for (int j = 0; j < npairs; ++j) {
            const int k = (head - 1 - j + 2 * ORC_OPT_MEM) % ORC_OPT_MEM;
            const double a = orc_sum_free(S[k], Y[k], fr, L), ss = orc_sum_free(S[k], S[k], fr, L);
            const double yy = orc_sum_free(Y[k], Y[k], fr, L), sq = orc_sum_free(S[k], d, fr, L);
            sy[j] = (a > 1e-10 * sqrt(ss * yy)) ? a : 0.0;
            if (sy[j] > 0.0) {
                al[j] = sq / sy[j];
                for (int i = 0; i < L; ++i) if (fr[i]) d[i] -= al[j] * Y[k][i];
                if (!have_scale) { scale = sy[j] / yy; have_scale = 1; }
            }
        }
        for (int i = 0; i < L; ++i) d[i] *= scale;
        for (int j = npairs - 1; j >= 0; --j) {
            if (!(sy[j] > 0.0)) continue;
            const int k = (head - 1 - j + 2 * ORC_OPT_MEM) % ORC_OPT_MEM;
            const double yr = orc_sum_free(Y[k], d, fr, L);
            const double b = yr / sy[j];
            for (int i = 0; i < L; ++i) if (fr[i]) d[i] += (al[j] - b) * S[k][i];
        }
        for (int i = 0; i < L; ++i) d[i] = -d[i];
        const double gd = orc_sum_free(g, d, NULL, L);
        if (!(gd < 0.0) || !isfinite(gd)) {
            npairs = 0;
            for (int i = 0; i < L; ++i) d[i] = fr[i] ? -g[i] * (step0 / pg) : 0.0;
        }
        /* projected backtracking line search */
        double lam = 1.0, Jt = J;
        int ok = 0;
        for (int bt = 0; bt < 40; ++bt) {
            double gs = 0.0;
            for (int i = 0; i < L; ++i) {
                xt[i] = clipd(x[i] + lam * d[i], lo[i % m], hi[i % m]);
                gs += g[i] * (xt[i] - x[i]);
            }
            Jt = orc_actor_cost(c, s, xt, observation, state_sys, w_critic);
            ++nfev;
            if (Jt <= J + 1e-4 * fmin(gs, 0.0)) { ok = 1; break; }       /* never accept an increase (gs can be >= 0 after the projection) */
            lam *= 0.5;
        }
        if (!ok) break;
        orc_actor_grad(c, s, xt, observation, state_sys, w_critic, gn);
        ++iters;
        for (int i = 0; i < L; ++i) { S[head][i] = xt[i] - x[i]; Y[head][i] = gn[i] - g[i]; }
        head = (head + 1) % ORC_OPT_MEM;
        if (npairs < ORC_OPT_MEM) ++npairs;
        const double Jold = J;
        J = Jt;
        memcpy(x, xt, sizeof(double) * (size_t)L);
        memcpy(g, gn, sizeof(double) * (size_t)L);
        if (fabs(Jold - J) <= f_tol * fmax(fmax(fabs(Jold), fabs(J)), 1.0)) {
            if (++stall >= 2) break;
        } else {
            stall = 0;
        }
    }
    if (iters_out) *iters_out = iters;
    if (nfev_out) *nfev_out = nfev;
    return J;
}

/* Batch driver of orc_actor_opt for the CPU baseline of bench.py's actor_optimizer block: E independent problems
 * (observation = state_sys = row e of `states` [E, n], the same start point x_init [Nactor*m] and weights for all),
 * OpenMP over the problems.  J_out [E]; returns the total number of gradient evaluations. */
long long orc_actor_opt_batch(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *x_init, const double *states,
                              const double *w_critic, int max_iter, double pg_tol, double f_tol, int nthreads,
                              double *J_out)
{
    const int n = s->n, L = c->Nactor * s->m;
    long long grads = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : grads)
    for (int e = 0; e < E; ++e) {
        double x[ORC_OPT_LMAX];
        int it = 0, nf = 0;
        memcpy(x, x_init, sizeof(double) * (size_t)L);
        J_out[e] = orc_actor_opt(c, s, x, states + (long)e * n, states + (long)e * n, w_critic, max_iter, pg_tol, f_tol, &it, &nf);
        grads += it + 1;
    }
    return grads;
}

/* ------------------------------------------------- CtrlNominal3WRobotNI (nominal parking controller) */

static double sgn(double v) { return (v > 0.0) ? 1.0 : (v < 0.0) ? -1.0 : v; }   /* np.sign: 0 -> 0, NaN -> NaN */

/* ref: rcognita/controllers.py:1758-1956, CtrlNominal3WRobotNI.compute_action_vanila + the clipping of
 * compute_action (:1915-1921): _Cart2NH (:1880-1893), _zeta (:1786-1834), _kappa (:1836-1853),
 * uNI = ctrl_gain * kappa, _NH2ctrl_Cart (:1895-1905), then np.clip to ctrl_bnds if ctrl_bnds.any().
 * Expressions are evaluated in Python's order; x**3 and |x|**(1/3) are libm pow() like numpy's scalar power. */
void orc_nominal_ni(double ctrl_gain, const orc_sys_t *s, const double *obs, double *action)
{
    const double xc = obs[0], yc = obs[1], alpha = obs[2];
    const double ca = cos(alpha), sa = sin(alpha);
    double x[3], zeta[3];
    x[0] = alpha;
    x[1] = xc * ca + yc * sa;
    x[2] = -2 * (yc * ca - xc * sa) - alpha * (xc * ca + yc * sa);
    const double a2 = fabs(x[2]);
    if (x[0] == 0 && x[1] == 0) {                                   /* nablaF with theta = 0 (:1814-1830) */
        const double st = x[0] * 1.0 + x[1] * 0.0 + sqrt(a2);
        zeta[0] = 4 * pow(x[0], 3) - 2 * pow(a2, 3) * 1.0 / pow(st, 3);
        zeta[1] = 4 * pow(x[1], 3) - 2 * pow(a2, 3) * 0.0 / pow(st, 3);
        zeta[2] = (3 * x[0] * 1.0 + 3 * x[1] * 0.0 + 2 * sqrt(a2)) * pow(x[2], 2) * sgn(x[2]) / pow(st, 3);
    } else {                                                        /* nablaL (:1804-1810) */
        const double r = sqrt(pow(x[0], 2) + pow(x[1], 2));
        const double sigma = r + sqrt(a2);
        zeta[0] = 4 * pow(x[0], 3) + pow(a2, 3) / pow(sigma, 3) * 1 / pow(r, 3) * 2 * x[0];
        zeta[1] = 4 * pow(x[1], 3) + pow(a2, 3) / pow(sigma, 3) * 1 / pow(r, 3) * 2 * x[1];
        zeta[2] = 3 * pow(a2, 2) * sgn(x[2]) + pow(a2, 3) / pow(sigma, 3) * 1 / sqrt(a2) * sgn(x[2]);
    }
    const double d0 = zeta[0] * 1.0 + zeta[1] * 0.0 + zeta[2] * x[1];          /* np.dot(zeta, G[:,0]) */
    const double d1 = zeta[0] * 0.0 + zeta[1] * 1.0 + zeta[2] * (-x[0]);       /* np.dot(zeta, G[:,1]) */
    const double k0 = -pow(fabs(d0), 1.0 / 3) * sgn(d0);
    const double k1 = -pow(fabs(d1), 1.0 / 3) * sgn(d1);
    const double u0 = ctrl_gain * k0, u1 = ctrl_gain * k1;
    action[0] = u1 + 1.0 / 2 * u0 * (x[2] + x[0] * x[1]);
    action[1] = u0;
    if (s->has_bnds)
        for (int k = 0; k < 2; ++k) action[k] = clipd(action[k], s->lo[k], s->hi[k]);
}

/* ------------------------------------------------- control-limited Gauss-Newton (iLQR) sweeps + L-BFGS hand-over
 *
 * Groundwork for the successor of the projected L-BFGS kernel (numpy prototype:
 * tools/ilqr_prototype.py).  Every stage term of _actor_cost is exactly quadratic in z = [obs - shift, action]
 * (stage_obj 'quadratic': z^T R1 z scaled by gamma**k; the four critic structures: a quadratic form in chi, or in the
 * raw observation for quad-mix, plus a linear part for quad-lin), so a reverse Riccati pass over the horizon with the
 * linearised Euler step (plus, for Sys3WRobot, the second-order term of its heading/speed coupling) gives a Newton-like step
 * with O(n^2) state; the box on the actions is handled per stage by a
 * clamped Newton step (m <= 2: the 3^m active sets are enumerated), Q_aa is regularised (Levenberg-Marquardt), the
 * forward pass backtracks on the TRUE cost (orc_actor_cost).  At most max_sweeps sweeps; a start that is already
 * stationary (projected-gradient test) costs none; four failed forward passes in a row -- the indefinite critics --
 * hand over to orc_actor_opt, which also polishes the result.  'biquadratic' stage costs are not quadratic: L-BFGS only. */
static void stage_quad(const orc_ctrl_t *c, int n, int m, int use_critic, double gk, const double *w, double *H, double *g0,
                       double *shift)
{
    const int p = n + m;
    memset(H, 0, sizeof(double) * ORC_MAX_P * ORC_MAX_P);
    memset(g0, 0, sizeof(double) * ORC_MAX_P);
    for (int i = 0; i < n; ++i) shift[i] = c->has_target ? c->target[i] : 0.0;
    if (!use_critic) {
        for (int i = 0; i < p; ++i)
            for (int j = 0; j < p; ++j) H[i * ORC_MAX_P + j] = gk * (c->R1[i * p + j] + c->R1[j * p + i]);
        return;
    }
    int k = 0;
    switch (c->critic_struct) {
    case ORC_CRITIC_QUAD_LIN:
    case ORC_CRITIC_QUADRATIC:
        for (int i = 0; i < p; ++i)
            for (int j = i; j < p; ++j) { H[i * ORC_MAX_P + j] += w[k]; H[j * ORC_MAX_P + i] += w[k]; ++k; }
        if (c->critic_struct == ORC_CRITIC_QUAD_LIN)
            for (int i = 0; i < p; ++i) g0[i] = w[k++];
        break;
    case ORC_CRITIC_QUAD_NOMIX:
        for (int i = 0; i < p; ++i) H[i * ORC_MAX_P + i] = 2.0 * w[k++];
        break;
    default:                                    /* quad-mix: raw observation */
        for (int i = 0; i < n; ++i) shift[i] = 0.0;
        for (int i = 0; i < n; ++i) H[i * ORC_MAX_P + i] = 2.0 * w[k++];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < m; ++j) { H[i * ORC_MAX_P + n + j] += w[k]; H[(n + j) * ORC_MAX_P + i] += w[k]; ++k; }
        for (int j = 0; j < m; ++j) H[(n + j) * ORC_MAX_P + n + j] = 2.0 * w[k++];
        break;
    }
}

/* (d f / d x, d f / d a) of _state_dyn (systems.py:308-323, :370-382, :412-419), row-major [n][n], [n][m]. */
static void dyn_jac(const orc_sys_t *s, const double *x, const double *a, double *fx, double *fa)
{
    const int n = s->n, m = s->m;
    memset(fx, 0, sizeof(double) * (size_t)(n * n));
    memset(fa, 0, sizeof(double) * (size_t)(n * m));
    if (s->sys_id == ORC_SYS_3WROBOT_NI) {
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        fx[0 * n + 2] = -a[0] * sn; fx[1 * n + 2] = a[0] * cs;
        fa[0 * m + 0] = cs; fa[1 * m + 0] = sn; fa[2 * m + 1] = 1.0;
    } else if (s->sys_id == ORC_SYS_3WROBOT) {
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        fx[0 * n + 2] = -x[3] * sn; fx[1 * n + 2] = x[3] * cs;
        fx[0 * n + 3] = cs; fx[1 * n + 3] = sn; fx[2 * n + 4] = 1.0;
        fa[3 * m + 0] = 1.0 / s->pars[0]; fa[4 * m + 1] = 1.0 / s->pars[1];
    } else {
        const double t1 = s->pars[0], t2 = s->pars[1], K1 = s->pars[2], K2 = s->pars[3], K3 = s->pars[4];
        fx[0] = -1.0 / t1; fx[1 * n + 0] = K2 / t2; fx[1 * n + 1] = (-1.0 + 2.0 * K3 * x[1]) / t2;
        fa[0] = K1 / t1;
    }
}

/* argmin 1/2 d^T Q d + q^T d, lo <= d <= hi, m <= 2, Q positive definite: best feasible KKT point over the 3^m
 * active sets.  free[j] = 1 where the minimiser is interior. */
static void box_newton(int m, const double *Q, const double *q, const double *lo, const double *hi, double *d, int *fr)
{
    double bestv = INFINITY;
    int found = 0;
    const int npat = (m == 1) ? 3 : 9;
    for (int pat = 0; pat < npat; ++pat) {
        int st[2] = {pat % 3, pat / 3};
        double t[2] = {0.0, 0.0};
        for (int j = 0; j < m; ++j) t[j] = (st[j] == 1) ? lo[j] : (st[j] == 2) ? hi[j] : 0.0;
        if (m == 1) {
            if (st[0] == 0) t[0] = -q[0] / Q[0];
        } else if (st[0] == 0 && st[1] == 0) {
            const double det = Q[0] * Q[3] - Q[1] * Q[2];
            t[0] = (-q[0] * Q[3] + q[1] * Q[1]) / det;
            t[1] = (-q[1] * Q[0] + q[0] * Q[2]) / det;
        } else if (st[0] == 0) {
            t[0] = -(q[0] + Q[1] * t[1]) / Q[0];
        } else if (st[1] == 0) {
            t[1] = -(q[1] + Q[2] * t[0]) / Q[3];
        }
        int feas = 1;
        for (int j = 0; j < m; ++j) feas = feas && isfinite(t[j]) && t[j] >= lo[j] - 1e-12 && t[j] <= hi[j] + 1e-12;
        if (!feas) continue;
        double v = 0.0;
        for (int i = 0; i < m; ++i) {
            v += q[i] * t[i];
            for (int j = 0; j < m; ++j) v += 0.5 * t[i] * Q[i * m + j] * t[j];
        }
        if (v < bestv) {
            bestv = v;
            found = 1;
            for (int j = 0; j < m; ++j) { d[j] = t[j]; fr[j] = (st[j] == 0); }
        }
    }
    if (!found)
        for (int j = 0; j < m; ++j) { d[j] = clipd(-q[j] / fmax(Q[j * m + j], 1e-12), lo[j], hi[j]); fr[j] = 0; }
}

double orc_actor_opt_hybrid(const orc_ctrl_t *c, const orc_sys_t *s, double *x, const double *observation,
                            const double *state_sys, const double *w_critic, int max_sweeps, int max_iter, double pg_tol,
                            double f_tol, int *sweeps_out, int *iters_out)
{
    const int n = s->n, m = s->m, N = c->Nactor, L = N * m, P = ORC_MAX_P;
    const double h = c->pred_step_size;
    double lo[ORC_MAX_M], hi[ORC_MAX_M];
    int sweeps = 0, iters = 0, nf = 0;
    for (int j = 0; j < m; ++j) {
        lo[j] = s->has_bnds ? s->lo[j] : -INFINITY;
        hi[j] = s->has_bnds ? s->hi[j] : INFINITY;
    }
    for (int i = 0; i < L; ++i) x[i] = clipd(x[i], lo[i % m], hi[i % m]);
    if (c->stage_struct == ORC_STAGE_QUADRATIC && max_sweeps > 0) {
        static _Thread_local double X[ORC_MAX_NACTOR][ORC_MAX_N], kff[ORC_MAX_NACTOR][ORC_MAX_M],
            Kfb[ORC_MAX_NACTOR][ORC_MAX_M][ORC_MAX_N], Un[ORC_OPT_LMAX], g[ORC_OPT_LMAX];
        double H[ORC_MAX_P * ORC_MAX_P], g0[ORC_MAX_P], shift[ORC_MAX_N];
        double J = orc_actor_grad(c, s, x, observation, state_sys, w_critic, g);
        double pg = 0.0;
        for (int i = 0; i < L; ++i) pg = fmax(pg, fabs(clipd(x[i] - g[i], lo[i % m], hi[i % m]) - x[i]));
        double mu = 1e-6;
        int stalls = 0;
        while (pg > pg_tol && sweeps < max_sweeps) {
            ++sweeps;
            for (int i = 0; i < n; ++i) X[0][i] = state_sys[i];
            for (int k = 1; k < N; ++k) {
                double dd[ORC_MAX_N];
                orc_state_dyn(s, X[k - 1], x + (k - 1) * m, dd);
                for (int i = 0; i < n; ++i) X[k][i] = X[k - 1][i] + h * dd[i];
            }
            int ok = 0;
            while (!ok) {                                  /* reverse pass; raise mu until every Q_aa is positive definite */
                double Vx[ORC_MAX_N] = {0}, Vxx[ORC_MAX_N * ORC_MAX_N] = {0};
                ok = 1;
                for (int k = N - 1; k >= 0 && ok; --k) {
                    const int use_critic = (c->mode == ORC_MODE_SQL) || (c->mode == ORC_MODE_RQL && k == N - 1);
                    stage_quad(c, n, m, use_critic, pow(c->gamma, (double)k), w_critic, H, g0, shift);
                    const double *ob = (k == 0) ? observation : X[k];
                    const double *a = x + k * m;
                    double z[ORC_MAX_P], gz[ORC_MAX_P];
                    for (int i = 0; i < n; ++i) z[i] = ob[i] - shift[i];
                    for (int j = 0; j < m; ++j) z[n + j] = a[j];
                    for (int i = 0; i < n + m; ++i) {
                        double acc = g0[i];
                        for (int j = 0; j < n + m; ++j) acc += H[i * P + j] * z[j];
                        gz[i] = acc;
                    }
                    double Qx[ORC_MAX_N], Qa[ORC_MAX_M], Qxx[ORC_MAX_N * ORC_MAX_N], Qax[ORC_MAX_M * ORC_MAX_N], Qaa[4];
                    for (int i = 0; i < n; ++i) {
                        Qx[i] = (k > 0) ? gz[i] : 0.0;
                        for (int j = 0; j < n; ++j) Qxx[i * n + j] = (k > 0) ? H[i * P + j] : 0.0;
                    }
                    for (int j = 0; j < m; ++j) {
                        Qa[j] = gz[n + j];
                        for (int i = 0; i < n; ++i) Qax[j * n + i] = (k > 0) ? H[(n + j) * P + i] : 0.0;
                        for (int l = 0; l < m; ++l) Qaa[j * m + l] = H[(n + j) * P + n + l];
                    }
                    if (k < N - 1) {
                        double fx[ORC_MAX_N * ORC_MAX_N], fa[ORC_MAX_N * ORC_MAX_M], A[ORC_MAX_N * ORC_MAX_N], B[ORC_MAX_N * ORC_MAX_M];
                        double VA[ORC_MAX_N * ORC_MAX_N], VB[ORC_MAX_N * ORC_MAX_M];
                        dyn_jac(s, X[k], a, fx, fa);
                        for (int i = 0; i < n; ++i)
                            for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 1.0 : 0.0) + h * fx[i * n + j];
                        for (int i = 0; i < n; ++i)
                            for (int j = 0; j < m; ++j) B[i * m + j] = h * fa[i * m + j];
                        for (int i = 0; i < n; ++i) {
                            for (int j = 0; j < n; ++j) {
                                double acc = 0.0;
                                for (int l = 0; l < n; ++l) acc += Vxx[i * n + l] * A[l * n + j];
                                VA[i * n + j] = acc;
                            }
                            for (int j = 0; j < m; ++j) {
                                double acc = 0.0;
                                for (int l = 0; l < n; ++l) acc += Vxx[i * n + l] * B[l * m + j];
                                VB[i * m + j] = acc;
                            }
                        }
                        for (int i = 0; i < n; ++i) {
                            for (int l = 0; l < n; ++l) Qx[i] += A[l * n + i] * Vx[l];
                            for (int j = 0; j < n; ++j)
                                for (int l = 0; l < n; ++l) Qxx[i * n + j] += A[l * n + i] * VA[l * n + j];
                        }
                        if (s->sys_id == ORC_SYS_3WROBOT) {          /* second-order term of the heading/speed coupling (DDP) */
                            double sn, cs;
                            orc_sincos(X[k][2], &sn, &cs);
                            Qxx[2 * n + 2] += h * (Vx[0] * (-X[k][3] * cs) + Vx[1] * (-X[k][3] * sn));
                            const double cc = h * (Vx[0] * (-sn) + Vx[1] * cs);
                            Qxx[2 * n + 3] += cc;
                            Qxx[3 * n + 2] += cc;
                        }
                        for (int j = 0; j < m; ++j) {
                            for (int l = 0; l < n; ++l) Qa[j] += B[l * m + j] * Vx[l];
                            for (int i = 0; i < n; ++i)
                                for (int l = 0; l < n; ++l) Qax[j * n + i] += B[l * m + j] * VA[l * n + i];
                            for (int q = 0; q < m; ++q)
                                for (int l = 0; l < n; ++l) Qaa[j * m + q] += B[l * m + j] * VB[l * m + q];
                        }
                    }
                    double Qr[4];
                    for (int j = 0; j < m * m; ++j) Qr[j] = Qaa[j];
                    for (int j = 0; j < m; ++j) Qr[j * m + j] += mu;
                    const int pd = (m == 1) ? (Qr[0] > 0.0)
                                            : (Qr[0] > 0.0 && Qr[0] * Qr[3] - 0.25 * (Qr[1] + Qr[2]) * (Qr[1] + Qr[2]) > 0.0);
                    if (!pd) { ok = 0; break; }
                    double dlo[ORC_MAX_M], dhi[ORC_MAX_M], d[ORC_MAX_M];
                    int fr[ORC_MAX_M];
                    for (int j = 0; j < m; ++j) { dlo[j] = lo[j] - a[j]; dhi[j] = hi[j] - a[j]; }
                    box_newton(m, Qr, Qa, dlo, dhi, d, fr);
                    double K[ORC_MAX_M * ORC_MAX_N] = {0};
                    if (m == 1) {
                        if (fr[0]) for (int i = 0; i < n; ++i) K[i] = -Qax[i] / Qr[0];
                    } else if (fr[0] && fr[1]) {
                        const double det = Qr[0] * Qr[3] - Qr[1] * Qr[2];
                        for (int i = 0; i < n; ++i) {
                            K[0 * n + i] = -(Qr[3] * Qax[0 * n + i] - Qr[1] * Qax[1 * n + i]) / det;
                            K[1 * n + i] = -(Qr[0] * Qax[1 * n + i] - Qr[2] * Qax[0 * n + i]) / det;
                        }
                    } else if (fr[0]) {
                        for (int i = 0; i < n; ++i) K[0 * n + i] = -Qax[0 * n + i] / Qr[0];
                    } else if (fr[1]) {
                        for (int i = 0; i < n; ++i) K[1 * n + i] = -Qax[1 * n + i] / Qr[3];
                    }
                    for (int j = 0; j < m; ++j) {
                        kff[k][j] = d[j];
                        for (int i = 0; i < n; ++i) Kfb[k][j][i] = K[j * n + i];
                    }
                    /* V_x = Q_x + K^T Q_aa d + K^T Q_a + Q_ax^T d;  V_xx = Q_xx + K^T Q_aa K + K^T Q_ax + Q_ax^T K */
                    double Qd[ORC_MAX_M] = {0}, QK[ORC_MAX_M * ORC_MAX_N] = {0};
                    for (int j = 0; j < m; ++j)
                        for (int q = 0; q < m; ++q) {
                            Qd[j] += Qaa[j * m + q] * d[q];
                            for (int i = 0; i < n; ++i) QK[j * n + i] += Qaa[j * m + q] * K[q * n + i];
                        }
                    for (int i = 0; i < n; ++i) {
                        double v = Qx[i];
                        for (int j = 0; j < m; ++j) v += K[j * n + i] * (Qd[j] + Qa[j]) + Qax[j * n + i] * d[j];
                        Vx[i] = v;
                    }
                    for (int i = 0; i < n; ++i)
                        for (int l = 0; l < n; ++l) {
                            double v = Qxx[i * n + l];
                            for (int j = 0; j < m; ++j) v += K[j * n + i] * (QK[j * n + l] + Qax[j * n + l]) + Qax[j * n + i] * K[j * n + l];
                            Vxx[i * n + l] = v;
                        }
                    for (int i = 0; i < n; ++i)
                        for (int l = i + 1; l < n; ++l) {
                            const double v = 0.5 * (Vxx[i * n + l] + Vxx[l * n + i]);
                            Vxx[i * n + l] = v; Vxx[l * n + i] = v;
                        }
                }
                if (!ok) {
                    mu = fmax(mu * 10.0, 1e-6);
                    if (mu > 1e12) break;
                }
            }
            if (!ok) break;
            /* forward pass with backtracking on the true cost */
            double alpha = 1.0, Jn = J;
            int improved = 0;
            for (int bt = 0; bt < 12 && !improved; ++bt) {
                double xs[ORC_MAX_N], dd[ORC_MAX_N];
                for (int i = 0; i < n; ++i) xs[i] = state_sys[i];
                for (int k = 0; k < N; ++k) {
                    for (int j = 0; j < m; ++j) {
                        double v = x[k * m + j] + alpha * kff[k][j];
                        for (int i = 0; i < n; ++i) v += Kfb[k][j][i] * (xs[i] - X[k][i]);
                        Un[k * m + j] = clipd(v, lo[j], hi[j]);
                    }
                    if (k < N - 1) {
                        orc_state_dyn(s, xs, Un + k * m, dd);
                        for (int i = 0; i < n; ++i) xs[i] = xs[i] + h * dd[i];
                    }
                }
                Jn = orc_actor_cost(c, s, Un, observation, state_sys, w_critic);
                if (Jn < J) improved = 1; else alpha *= 0.5;
            }
            if (!improved) {
                mu *= 10.0;
                if (mu > 1e12 || ++stalls >= 4) break;
                continue;
            }
            stalls = 0;
            const double dJ = J - Jn;
            memcpy(x, Un, sizeof(double) * (size_t)L);
            J = Jn;
            mu = fmax(mu / 10.0, 1e-9);
            if (dJ <= 1e-9 * fmax(fabs(J), 1.0)) break;
            orc_actor_grad(c, s, x, observation, state_sys, w_critic, g);
            pg = 0.0;
            for (int i = 0; i < L; ++i) pg = fmax(pg, fabs(clipd(x[i] - g[i], lo[i % m], hi[i % m]) - x[i]));
        }
    }
    const double J = orc_actor_opt(c, s, x, observation, state_sys, w_critic, max_iter, pg_tol, f_tol, &iters, &nf);
    if (sweeps_out) *sweeps_out = sweeps;
    if (iters_out) *iters_out = iters;
    return J;
}
