/*
 * rcg_oracle_critic.c -- CPU ORACLE (test infrastructure, NOT product code): the critic side of
 * CtrlOptPred.compute_action for RQL / SQL.
 *
 *   orc_critic_fit          bounded least-squares stand-in for CtrlOptPred._critic_optimizer
 *                           (ref: rcognita/controllers.py:1248-1271).  The reference runs scipy's SLSQP on
 *                           _critic_cost (:1216-1245), which is linear least squares in w with K = Ncritic-1
 *                           rows and 3..35 unknowns (minimiser not unique).  This is a scalar restatement of
 *                           the algorithm the product's CUDA kernels use (rcognita_b200/csrc/critic_fit.cu:
 *                           proximal-point continuation, each stage solved in the K-dimensional dual by a
 *                           semismooth Newton method with Armijo backtracking; best iterate returned).  It is
 *                           the checker for those kernels; the bar against the REFERENCE stays the fitted
 *                           cost <= SLSQP's on tests/golden/critic_fit.json.
 *   orc_env_iterate_critic  one iteration of the headless main loop (presets/main_3wrobot_NI.py:415-440)
 *                           INCLUDING the RQL/SQL branch of compute_action (controllers.py:1455-1479): FIFO
 *                           push of (action_curr, observation), critic clock, refit, w_critic_prev
 *                           hand-over -- pinned against tests/golden/closed_loop_refit.json (the unmodified
 *                           reference with its own SLSQP fit; this loop can LOAD recorded weights instead of
 *                           fitting, which makes everything else comparable to 1e-9).
 */
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rcg_oracle.h"

#define FIT_MAX_K 15

/* phi(observation, action) of _critic (ref: controllers.py:1204-1212; orc_critic has the same order). */
static int critic_phi(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act, double *phi)
{
    const int p = n + m;
    double chi[ORC_MAX_P];
    int k = 0;
    for (int i = 0; i < n; ++i) chi[i] = c->has_target ? obs[i] - c->target[i] : obs[i];
    for (int j = 0; j < m; ++j) chi[n + j] = act[j];
    switch (c->critic_struct) {
    case ORC_CRITIC_QUAD_LIN:
        for (int i = 0; i < p; ++i) for (int j = i; j < p; ++j) phi[k++] = chi[i] * chi[j];
        for (int i = 0; i < p; ++i) phi[k++] = chi[i];
        break;
    case ORC_CRITIC_QUADRATIC:
        for (int i = 0; i < p; ++i) for (int j = i; j < p; ++j) phi[k++] = chi[i] * chi[j];
        break;
    case ORC_CRITIC_QUAD_NOMIX:
        for (int i = 0; i < p; ++i) phi[k++] = chi[i] * chi[i];
        break;
    default:
        for (int i = 0; i < n; ++i) phi[k++] = obs[i] * obs[i];
        for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) phi[k++] = obs[i] * act[j];
        for (int j = 0; j < m; ++j) phi[k++] = act[j] * act[j];
        break;
    }
    return k;
}

typedef struct {
    int K, D;
    double lo, hi, mu;
    const double *Phi, *b, *w0;
} fit_ctx_t;

static double clipw(const fit_ctx_t *f, double z) { return z < f->lo ? f->lo : (z > f->hi ? f->hi : z); }

static double zj(const fit_ctx_t *f, const double *l, int j)
{
    double z = f->w0[j];
    for (int r = 0; r < f->K; ++r) z = fma(f->Phi[r * f->D + j], l[r], z);
    return z;
}

/* the convex dual objective of one proximal stage (its gradient is F below) */
static double dual(const fit_ctx_t *f, const double *l)
{
    double q = 0, s = 0;
    for (int r = 0; r < f->K; ++r) { q = fma(l[r], l[r], q); s = fma(-f->b[r], l[r], s); }
    s = fma(0.5 * f->mu, q, s);
    for (int j = 0; j < f->D; ++j) {
        const double z = zj(f, l, j);
        s += z < f->lo ? f->lo * z - 0.5 * f->lo * f->lo : (z > f->hi ? f->hi * z - 0.5 * f->hi * f->hi : 0.5 * z * z);
    }
    return s;
}

static double ls_cost(int K, int D, const double *Phi, const double *b, const double *w)
{
    double J = 0;
    for (int r = 0; r < K; ++r) {
        double s = -b[r];
        for (int j = 0; j < D; ++j) s = fma(Phi[r * D + j], w[j], s);
        J = fma(0.5 * s, s, J);
    }
    return J;
}

static int clip_state(double z, double lo, double hi)
{
    return (z > lo ? 1 : (z < lo ? -1 : 0)) + (z > hi ? 1 : (z < hi ? -1 : 0));
}

/* Bounded least-squares fit of _critic_cost (see the header comment).  obs_buf [buffer_size, n], act_buf
 * [buffer_size, m] row-major (row 0 = oldest); w_init NULL = start from w_out's content.  Constants are the
 * product's: mu_rel = 1e-3 shrinking 100x per stage, 5 stages, 20 Newton steps, 40 halvings.  max_evals <= 0: to
 * convergence.  Returns _critic_cost at the fitted weights; *evals_out (may be NULL) = dual passes spent. */
/* g(a) = d . grad D(lam + a d) along the Newton direction: c0 + c1 a + sum_j s_j clip(z_j + a s_j) -- piecewise linear and
 * increasing in a. */
static double ls_g(const fit_ctx_t *f, const double *z, const double *s, double c0, double c1, double a)
{
    double acc = 0;
    for (int j = 0; j < f->D; ++j) acc = fma(s[j], clipw(f, fma(a, s[j], z[j])), acc);
    return fma(c1, a, c0) + acc;
}

/* Exact minimiser of the dual along d (the product's warp-per-environment kernel does the same, critic_fit.cu): every weight
 * contributes up to two breakpoints a > 0 where z_j + a s_j crosses lo or hi; g is evaluated at every breakpoint, the
 * bracket [a_L, a_R] = (largest breakpoint with g < 0, smallest with g >= 0) is the linear piece that holds the root.
 * Returns the step, or a negative number if there is none. */
static double ls_exact(const fit_ctx_t *f, const double *lam, const double *dl, double slope)
{
    double z[ORC_MAX_W], s[ORC_MAX_W];
    double c0 = 0, c1 = 0;
    for (int r = 0; r < f->K; ++r) { c0 = fma(f->mu * lam[r] - f->b[r], dl[r], c0); c1 = fma(f->mu * dl[r], dl[r], c1); }
    for (int j = 0; j < f->D; ++j) {
        z[j] = zj(f, lam, j);
        double sj = 0;
        for (int r = 0; r < f->K; ++r) sj = fma(f->Phi[r * f->D + j], dl[r], sj);
        s[j] = sj;
    }
    double aL = 0.0, gL = slope, aR = INFINITY, gR = 0.0;
    for (int j = 0; j < f->D; ++j) {
        if (s[j] == 0.0) continue;
        for (int side = 0; side < 2; ++side) {
            const double a = ((side ? f->hi : f->lo) - z[j]) / s[j];
            if (!(a > 0.0) || !isfinite(a)) continue;
            const double g = ls_g(f, z, s, c0, c1, a);
            if (g < 0.0) { if (a > aL) { aL = a; gL = g; } }
            else if (a < aR) { aR = a; gR = g; }
        }
    }
    if (aR < INFINITY) {
        if (!(aR > aL)) return aR;
        return (gR > gL) ? aL + (aR - aL) * (-gL) / (gR - gL) : aR;
    }
    /* beyond the last breakpoint g is linear with the slope of its last piece */
    const double g1 = ls_g(f, z, s, c0, c1, aL + 1.0);
    const double sl = g1 - gL;
    return (sl > 0.0) ? aL - gL / sl : -1.0;
}

/* ls_mode 0: Armijo backtracking (one-lane kernels, any K); 1: unit step if it passes the Armijo test, else the exact
 * minimiser along the Newton direction (the two-phase path of rcg_critic_fit: K <= 3 and >= 10 weights). */
double orc_critic_fit_ls(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                         const double *w_prev, double lo, double hi, const double *w_init, double *w_out,
                         int max_evals, int ls_mode, int *evals_out);

double orc_critic_fit(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                      const double *w_prev, double lo, double hi, const double *w_init, double *w_out,
                      int max_evals, int *evals_out)
{
    const int exact = max_evals <= 0 && c->Ncritic - 1 <= 3 && orc_dim_critic(c->critic_struct, n, m) >= 10;
    return orc_critic_fit_ls(c, n, m, obs_buf, act_buf, w_prev, lo, hi, w_init, w_out, max_evals, exact, evals_out);
}

double orc_critic_fit_ls(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                         const double *w_prev, double lo, double hi, const double *w_init, double *w_out,
                         int max_evals, int ls_mode, int *evals_out)
{
    const int K = c->Ncritic - 1, D = orc_dim_critic(c->critic_struct, n, m);
    const int max_outer = 5, max_newton = 20, max_ls = 40;
    double mu_rel = 1e-3;
    double Phi[FIT_MAX_K * ORC_MAX_W], b[FIT_MAX_K], lam[FIT_MAX_K], lt[FIT_MAX_K], F[FIT_MAX_K], dl[FIT_MAX_K];
    double H[FIT_MAX_K * FIT_MAX_K];
    double w0[ORC_MAX_W], wb[ORC_MAX_W], wn[ORC_MAX_W];
    int st[ORC_MAX_W];
    double trace = 0, bb = 0;
    if (max_evals <= 0) max_evals = 0x7fffffff;
    if (evals_out) *evals_out = 0;
    if (K > FIT_MAX_K) return NAN;
    /* rows of the least-squares problem (ref: controllers.py:1230-1242), r = 0 is buffer row k = K */
    for (int r = 0; r < K; ++r) {
        const int k = K - r;
        const double *op = obs_buf + (long)(k - 1) * n, *on = obs_buf + (long)k * n;
        const double *ap = act_buf + (long)(k - 1) * m, *an = act_buf + (long)k * m;
        critic_phi(c, n, m, op, ap, Phi + r * D);
        for (int j = 0; j < D; ++j) trace = fma(Phi[r * D + j], Phi[r * D + j], trace);
        b[r] = c->gamma * orc_critic(c, n, m, on, an, w_prev) + orc_stage_obj(c, n, m, op, ap);
        bb = fma(b[r], b[r], bb);
    }
    fit_ctx_t f = {K, D, lo, hi, 0.0, Phi, b, w0};
    for (int j = 0; j < D; ++j) { w0[j] = clipw(&f, w_init ? w_init[j] : w_out[j]); wb[j] = w0[j]; }
    double Jbest = ls_cost(K, D, Phi, b, w0);
    const double J0 = Jbest;
    int evals = 0;
    if (K >= 1 && trace > 0 && isfinite(trace) && isfinite(bb)) {
        for (int outer = 0; outer < max_outer && evals < max_evals; ++outer) {
            f.mu = mu_rel * trace / K;
            mu_rel *= 1e-2;
            for (int r = 0; r < K; ++r) lam[r] = 0;
            for (int it = 0; it < max_newton && evals < max_evals; ++it) {
                ++evals;
                for (int r = 0; r < K; ++r) {
                    F[r] = f.mu * lam[r] - b[r];
                    for (int q = 0; q <= r; ++q) H[r * FIT_MAX_K + q] = (q == r) ? f.mu : 0.0;
                }
                for (int j = 0; j < D; ++j) {
                    const double z = zj(&f, lam, j), w = clipw(&f, z);
                    st[j] = clip_state(z, lo, hi);
                    for (int r = 0; r < K; ++r) F[r] = fma(Phi[r * D + j], w, F[r]);
                    if (z > lo && z < hi)
                        for (int r = 0; r < K; ++r)
                            for (int q = 0; q <= r; ++q)
                                H[r * FIT_MAX_K + q] = fma(Phi[r * D + j], Phi[q * D + j], H[r * FIT_MAX_K + q]);
                }
                double fmx = 0;
                for (int r = 0; r < K; ++r) fmx = fmx > fabs(F[r]) ? fmx : fabs(F[r]);
                if (fmx <= 1e-13 * sqrt(bb)) break;
                int spd = 1;                                   /* Cholesky H = L L^T, lower, in place */
                for (int r = 0; r < K && spd; ++r) {
                    for (int q = 0; q <= r; ++q) {
                        double s = H[r * FIT_MAX_K + q];
                        for (int cc = 0; cc < q; ++cc) s = fma(-H[r * FIT_MAX_K + cc], H[q * FIT_MAX_K + cc], s);
                        if (q == r) {
                            if (!(s > 0)) { spd = 0; break; }
                            H[r * FIT_MAX_K + r] = sqrt(s);
                        } else {
                            H[r * FIT_MAX_K + q] = s / H[q * FIT_MAX_K + q];
                        }
                    }
                }
                if (!spd) break;
                for (int r = 0; r < K; ++r) {
                    double s = -F[r];
                    for (int cc = 0; cc < r; ++cc) s = fma(-H[r * FIT_MAX_K + cc], dl[cc], s);
                    dl[r] = s / H[r * FIT_MAX_K + r];
                }
                for (int r = K - 1; r >= 0; --r) {
                    double s = dl[r];
                    for (int cc = r + 1; cc < K; ++cc) s = fma(-H[cc * FIT_MAX_K + r], dl[cc], s);
                    dl[r] = s / H[r * FIT_MAX_K + r];
                }
                double slope = 0;
                for (int r = 0; r < K; ++r) slope = fma(F[r], dl[r], slope);
                if (!(slope < 0)) break;
                const double D0 = dual(&f, lam);
                double a = 1.0;
                int ok = 0;
                for (int ls = 0; ls < max_ls && evals < max_evals; ++ls) {
                    ++evals;
                    for (int r = 0; r < K; ++r) lt[r] = fma(a, dl[r], lam[r]);
                    if (dual(&f, lt) <= D0 + 1e-4 * a * slope + 1e-14 * fabs(D0)) { ok = 1; break; }
                    if (ls_mode == 1) {                      /* the unit step failed: exact minimiser along dl */
                        evals += 3;
                        a = ls_exact(&f, lam, dl, slope);
                        if (a > 0.0 && isfinite(a)) {
                            for (int r = 0; r < K; ++r) lt[r] = fma(a, dl[r], lam[r]);
                            ok = 1;
                        }
                        break;
                    }
                    a *= 0.5;
                }
                if (!ok) break;
                for (int r = 0; r < K; ++r) lam[r] = lt[r];
                if (a == 1.0) {
                    int same = 1;
                    for (int j = 0; j < D; ++j) same = same && (clip_state(zj(&f, lam, j), lo, hi) == st[j]);
                    if (same) break;
                }
            }
            for (int j = 0; j < D; ++j) wn[j] = clipw(&f, zj(&f, lam, j));
            const double Jn = ls_cost(K, D, Phi, b, wn);
            if (Jn < Jbest) {
                Jbest = Jn;
                for (int j = 0; j < D; ++j) { w0[j] = wn[j]; wb[j] = wn[j]; }
                if (Jn <= 1e-12 * J0 || Jn <= 1e-20 * bb) break;
            }
        }
    }
    for (int j = 0; j < D; ++j) w_out[j] = wb[j];
    if (evals_out) *evals_out = evals;
    return Jbest;
}

/* utilities.push_vec (ref: rcognita/utilities.py:78-79): drop row 0, append at the bottom. */
static void push_vec(double *buf, int rows, int d, const double *v)
{
    memmove(buf, buf + d, (size_t)(rows - 1) * d * sizeof(double));
    for (int i = 0; i < d; ++i) buf[(long)(rows - 1) * d + i] = v[i];
}

/* CtrlOptPred.__init__ for the critic part (ref: controllers.py:980-981, :1041-1042): zero buffers,
 * w_critic = w_critic_prev = w_critic_init = ones; critic_clock = t0. */
void orc_critic_state_init(orc_critic_state_t *k, int dimc, double t0)
{
    memset(k, 0, sizeof(*k));
    for (int i = 0; i < dimc; ++i) k->w[i] = k->w_prev[i] = 1.0;
    k->critic_clock = t0;
}

/* ONE iteration of the headless main loop for one RQL/SQL environment, critic refit included.
 * ref: presets/main_3wrobot_NI.py:415-440; controllers.py:1429-1493 with the RQL/SQL branch :1455-1479:
 *   push action_curr (the PREVIOUS action) and the observation; if t - critic_clock >= critic_period: critic_clock = t,
 *   w_critic = fit, w_critic_prev = w_critic; else w_critic = w_critic_prev; then the actor (arg-min stand-in).
 * w_replay != NULL: instead of fitting, fit number j takes the weights w_replay[j*dimc ...] (the reference's recorded SLSQP
 * results).  Returns 1 if the controller sampled, 0 if it held, -1 if the environment is done. */
int orc_env_iterate_critic(orc_env_t *v, orc_critic_state_t *k, const orc_ctrl_t *c, const orc_sys_t *s, int C,
                           const double *tab, int buffer_size, double w_lo, double w_hi, double sampling_time,
                           double critic_period, double t1, const double *w_replay, int n_replay)
{
    const int n = s->n, m = s->m, L = c->Nactor * m, dimc = orc_dim_critic(c->critic_struct, n, m);
    int sampled = 0;
    double Jtab[4096];
    if (v->done) return -1;
    if (orc_rk45_step(&v->r, s, v->sys_action) != 0) { v->done = 1; return -1; }   /* sim_step */
    ++v->steps;
    const double t = v->r.t;
    const double *obs = v->r.y;
    if (t - v->ctrl_clock >= sampling_time) {                  /* controllers.py:1440-1442 */
        v->ctrl_clock = t;
        const double time_in_critic_period = t - k->critic_clock;                  /* :1457 */
        push_vec(k->act_buf, buffer_size, m, v->action_curr);                     /* :1460 */
        push_vec(k->obs_buf, buffer_size, n, obs);                                /* :1461 */
        if (time_in_critic_period >= critic_period) {                              /* :1463 */
            k->critic_clock = t;
            if (w_replay) {
                if (k->nfits < n_replay) memcpy(k->w, w_replay + (long)k->nfits * dimc, (size_t)dimc * sizeof(double));
                k->Jc = orc_critic_cost(c, n, m, k->obs_buf, k->act_buf, k->w, k->w_prev);
            } else {
                double ones[ORC_MAX_W];
                for (int i = 0; i < dimc; ++i) ones[i] = 1.0;                      /* w_critic_init (:1264) */
                k->Jc = orc_critic_fit(c, n, m, k->obs_buf, k->act_buf, k->w_prev, w_lo, w_hi, ones, k->w, 0, NULL);
            }
            memcpy(k->w_prev, k->w, (size_t)dimc * sizeof(double));                /* :1471 */
            ++k->nfits;
        } else {
            memcpy(k->w, k->w_prev, (size_t)dimc * sizeof(double));                /* :1479 */
        }
        for (int i0 = 0; i0 < C; i0 += 4096) {
            int cnt = C - i0 < 4096 ? C - i0 : 4096;
            int bi;
            orc_actor_cost_table(c, s, cnt, tab + (long)i0 * L, obs, v->state_sys, k->w, Jtab, &bi);
            if (i0 == 0 || (!isnan(v->Jbest) && (isnan(Jtab[bi]) || Jtab[bi] < v->Jbest))) {
                v->Jbest = Jtab[bi];
                v->best = i0 + bi;
            }
        }
        for (int j = 0; j < m; ++j) v->action_curr[j] = tab[(long)v->best * L + j];
        ++v->samples;
        sampled = 1;
    }
    for (int j = 0; j < m; ++j) v->sys_action[j] = v->action_curr[j];     /* receive_action       */
    for (int i = 0; i < n; ++i) v->state_sys[i] = v->r.y[i];              /* receive_sys_state    */
    v->accum += orc_stage_obj(c, n, m, obs, v->action_curr) * sampling_time;   /* upd_accum_obj   */
    if (t >= t1) v->done = 1;
    return sampled;
}

/* Whole RQL/SQL episodes with critic refit for E environments (OpenMP over environments).  Outputs (any may be
 * NULL): y_final [E,n], t_final [E], accum [E], nsteps [E], nsamples [E], nfits [E], w_final [E,dimc], Jc_final [E],
 * obs_buf_final [E,buffer_size,n], act_buf_final [E,buffer_size,m]; traj for environment 0: rows of
 * [t, y(n), action(m), accum, argmin, Jmin, sampled, nfits].  w_replay [n_replay, dimc] applies to every environment
 * (use E = 1). */
long long orc_closed_loop_critic(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *state_init, int C,
                                 const double *cand, int cand_per_env, const double *action_init, int buffer_size,
                                 double w_lo, double w_hi, double sampling_time, double critic_period, double t0, double t1,
                                 double max_step, double first_step, double rtol, double atol, int max_steps_per_env,
                                 int nthreads, const double *w_replay, int n_replay, double *y_final, double *t_final,
                                 double *accum, int *nsteps, int *nsamples, int *nfits, double *w_final, double *Jc_final,
                                 double *obs_buf_final, double *act_buf_final, double *traj, int traj_cap, int *traj_rows)
{
    const int n = s->n, m = s->m, L = c->Nactor * m, dimc = orc_dim_critic(c->critic_struct, n, m);
    long long total_steps = 0;
    if (traj_rows) *traj_rows = 0;
    if (buffer_size > ORC_MAX_BUF) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_steps)
    for (int e = 0; e < E; ++e) {
        orc_env_t v;
        orc_critic_state_t k;
        const double *tab = cand_per_env ? cand + (long)e * C * L : cand;
        orc_env_init(&v, s, 1, state_init + (long)e * n, action_init, t0, t1, max_step, first_step, rtol, atol);
        orc_critic_state_init(&k, dimc, t0);
        while (v.steps < max_steps_per_env) {
            int rc = orc_env_iterate_critic(&v, &k, c, s, C, tab, buffer_size, w_lo, w_hi, sampling_time, critic_period,
                                            t1, w_replay, n_replay);
            if (rc < 0) break;
            if (e == 0 && traj && v.steps <= traj_cap) {
                double *row = traj + (long)(v.steps - 1) * (1 + n + m + 5);
                row[0] = v.r.t;
                for (int i = 0; i < n; ++i) row[1 + i] = v.r.y[i];
                for (int j = 0; j < m; ++j) row[1 + n + j] = v.action_curr[j];
                row[1 + n + m] = v.accum;
                row[2 + n + m] = (double)v.best;
                row[3 + n + m] = v.Jbest;
                row[4 + n + m] = (double)rc;
                row[5 + n + m] = (double)k.nfits;
                if (traj_rows) *traj_rows = v.steps;
            }
            if (v.done) break;
        }
        if (y_final) for (int i = 0; i < n; ++i) y_final[(long)e * n + i] = v.r.y[i];
        if (t_final) t_final[e] = v.r.t;
        if (accum) accum[e] = v.accum;
        if (nsteps) nsteps[e] = v.steps;
        if (nsamples) nsamples[e] = v.samples;
        if (nfits) nfits[e] = k.nfits;
        if (w_final) memcpy(w_final + (long)e * dimc, k.w, (size_t)dimc * sizeof(double));
        if (Jc_final) Jc_final[e] = k.Jc;
        if (obs_buf_final) memcpy(obs_buf_final + (long)e * buffer_size * n, k.obs_buf, (size_t)buffer_size * n * sizeof(double));
        if (act_buf_final) memcpy(act_buf_final + (long)e * buffer_size * m, k.act_buf, (size_t)buffer_size * m * sizeof(double));
        total_steps += v.steps;
    }
    return total_steps;
}
