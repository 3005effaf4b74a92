/*
 * rcg_oracle_disturb.c -- CPU ORACLE (test infrastructure, NOT product code): disturbance lanes,
 * System(is_disturb = 1).
 *
 *   orc_state_dyn_disturbed   _state_dyn(t, state, action, disturb) with a disturbance
 *                             (ref: rcognita/systems.py:316-318 Sys3WRobot, :373-376 Sys3WRobotNI, :412-419 Sys2Tank).
 *   orc_disturb_dyn           _disturb_dyn GIVEN the draws z[k] = randn() (ref: systems.py:341-343, :390-392, :421-424).
 *   orc_closed_loop_rhs_disturbed   closed_loop_rhs on state_full = [state, disturb] (ref: systems.py:213-253).
 * These three are pinned to the live reference at function level (tests/golden/disturb.json: list-valued `disturb`,
 * patched randn) -- the reference's own closed loop with is_disturb = 1 raises under numpy 2 (`disturb != []` on an
 * ndarray, systems.py:316, :373), so no loop-level golden exists: PARITY OF THE DISTURBED LOOP IS UNPINNED beyond the
 * function level; below it is a restatement of scipy's RK45 on the full state (the same step logic as orc_rk45_step).
 *
 *   orc_normal2               the counter-based normal stream standing in for numpy's global randn(): Philox4x32-10
 *                             block (RHS call number, global environment index) under key seed, Box-Muller with the
 *                             specified log / sincos -- the same operation sequence as rcognita_b200/csrc/rcg_device.cuh,
 *                             bit for bit.
 *   orc_rk45d_*               scipy RK45 (rk.py:111-176) on the full state with the draws numbered by nfev.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "rcg_oracle.h"

static int dist_dim(int sys_id) { return sys_id == ORC_SYS_2TANK ? 1 : 2; }

void orc_state_dyn_disturbed(const orc_sys_t *s, const double *x, const double *a, const double *q, double *d)
{
    switch (s->sys_id) {
    case ORC_SYS_3WROBOT_NI: {
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        d[0] = a[0] * cs + q[0];            /* :374 */
        d[1] = a[0] * sn + q[0];            /* :375 -- disturb[0] again, literally */
        d[2] = a[1] + q[1];                 /* :376 */
        break;
    }
    case ORC_SYS_3WROBOT: {
        double m = s->pars[0], I = s->pars[1];
        double sn, cs;
        orc_sincos(x[2], &sn, &cs);
        d[0] = x[3] * cs;
        d[1] = x[3] * sn;
        d[2] = x[4];
        d[3] = 1 / m * (a[0] + q[0]);       /* :317 */
        d[4] = 1 / I * (a[1] + q[1]);       /* :318 */
        break;
    }
    default:
        orc_state_dyn(s, x, a, d);
        break;
    }
}

void orc_disturb_dyn(const orc_sys_t *s, const orc_dist_t *D, const double *q, const double *z, double *dq)
{
    if (s->sys_id == ORC_SYS_2TANK) {
        dq[0] = 0.0;                        /* :421-424 */
        return;
    }
    for (int k = 0; k < 2; ++k) dq[k] = -D->tau[k] * (q[k] + D->sigma[k] * (z[k] + D->mu[k]));   /* :343, :392 */
}

/* fdlibm / musl log for normal positive x, every operation rounded separately (-ffp-contract=off). */
double orc_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    uint32_t hx = (uint32_t)(bits >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int k = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    bits = ((uint64_t)hx << 32) | (bits & 0xffffffffull);
    double xm;
    memcpy(&xm, &bits, 8);
    const double f = xm - 1.0;
    const double hfsq = (0.5 * f) * f;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double dk = (double)k;
    double r = s * (hfsq + R) + dk * ln2_lo;
    r = r - hfsq;
    r = r + f;
    return r + dk * ln2_hi;
}

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

void orc_normal2(unsigned long long seed, unsigned long long env, unsigned int call, double *z)
{
    uint32_t c[4] = {call, 0u, (uint32_t)env, (uint32_t)(env >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = ((double)((((uint64_t)c[0] << 32) | c[1]) >> 11) + 0.5) * 1.1102230246251565e-16;
    const double u2 = ((double)((((uint64_t)c[2] << 32) | c[3]) >> 11) + 0.5) * 1.1102230246251565e-16;
    const double r = sqrt(-2.0 * orc_log(u1));
    double sn, cs;
    orc_sincos(6.283185307179586 * u2, &sn, &cs);
    z[0] = r * cs;
    z[1] = r * sn;
}

/* closed_loop_rhs on the full state (ref: systems.py:213-253, is_disturb = 1, is_dyn_ctrl = 0).  z != NULL: the draws are
 * given; else they come from the environment's stream at RHS call number `call`. */
void orc_closed_loop_rhs_disturbed(const orc_sys_t *s, const orc_dist_t *D, const double *y_full, double *action,
                                   unsigned long long env, unsigned int call, const double *z_given, double *rhs)
{
    const int n = s->n;
    double z[2] = {0.0, 0.0};
    if (s->has_bnds) {
        for (int k = 0; k < s->m; ++k) {
            double a = action[k];
            if (a < s->lo[k]) a = s->lo[k];
            if (a > s->hi[k]) a = s->hi[k];
            action[k] = a;
        }
    }
    orc_state_dyn_disturbed(s, y_full, action, y_full + n, rhs);
    if (z_given) { z[0] = z_given[0]; z[1] = z_given[1]; }
    else if (s->sys_id != ORC_SYS_2TANK) orc_normal2(D->seed, env, call, z);
    orc_disturb_dyn(s, D, y_full + n, z, rhs + n);
}

static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* Simulator.__init__ with is_disturb: state_full_init = [state_init, disturb_init] (ref: simulator.py:100-101), RK45 ctor
 * evaluates f = fun(t0, y0) = RHS call number 0 (scipy rk.py:97). */
void orc_rk45d_init(orc_rk45d_t *r, const orc_sys_t *s, const orc_dist_t *D, unsigned long long env, const double *y0_full,
                    double *action, double t0, double t_bound, double max_step, double first_step, double rtol, double atol)
{
    memset(r, 0, sizeof(*r));
    r->t = t0; r->t_bound = t_bound; r->max_step = max_step; r->rtol = rtol; r->atol = atol; r->h_abs = first_step;
    r->nfull = s->n + dist_dim(s->sys_id);
    r->env = env;
    for (int i = 0; i < r->nfull; ++i) r->y[i] = y0_full[i];
    orc_closed_loop_rhs_disturbed(s, D, r->y, action, env, 0u, NULL, r->f);
    r->nfev = 1;
    r->status = ORC_RUNNING;
}

/* scipy RK45.step on the full state: the logic of orc_rk45_step (base.py:179-212, rk.py:111-176, :61-71), RHS call
 * number = nfev at the time of the call. */
int orc_rk45d_step(orc_rk45d_t *r, const orc_sys_t *s, const orc_dist_t *D, double *action)
{
    const int n = r->nfull;
    if (r->status != ORC_RUNNING) return -1;
    if (r->t == r->t_bound) { r->status = ORC_FINISHED; return 0; }
    const double t = r->t;
    const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
    double h_abs;
    if (r->h_abs > r->max_step) h_abs = r->max_step;
    else if (r->h_abs < min_step) h_abs = min_step;
    else h_abs = r->h_abs;
    int step_rejected = 0;
    double K[7][ORC_MAX_NFULL], y_new[ORC_MAX_NFULL], ytmp[ORC_MAX_NFULL];
    double t_new, h;
    for (;;) {
        if (h_abs < min_step) { r->status = ORC_FAILED; return 1; }
        h = h_abs;
        t_new = t + h;
        if (t_new - r->t_bound > 0) t_new = r->t_bound;
        h = t_new - t;
        h_abs = fabs(h);
        for (int i = 0; i < n; ++i) K[0][i] = r->f[i];
        for (int st = 1; st < 6; ++st) {
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                for (int j = 0; j < st; ++j) acc += K[j][i] * RK_A[st][j];
                ytmp[i] = r->y[i] + acc * h;
            }
            orc_closed_loop_rhs_disturbed(s, D, ytmp, action, r->env, (unsigned int)(r->nfev + st - 1), NULL, K[st]);
        }
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
            y_new[i] = r->y[i] + h * acc;
        }
        orc_closed_loop_rhs_disturbed(s, D, y_new, action, r->env, (unsigned int)(r->nfev + 5), NULL, K[6]);
        r->nfev += 6;
        double sq = 0.0;
        for (int i = 0; i < n; ++i) {
            double scale = r->atol + fmax(fabs(r->y[i]), fabs(y_new[i])) * r->rtol;
            double acc = 0.0;
            for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_E[j];
            double e = acc * h / scale;
            sq += e * e;
        }
        double err = sqrt(sq) / sqrt((double)n);
        if (err < 1) {
            double factor;
            if (err == 0) factor = 10;
            else factor = fmin(10, 0.9 * orc_pow_m02(err));
            if (step_rejected) factor = fmin(1, factor);
            h_abs *= factor;
            break;
        } else {
            h_abs *= fmax(0.2, 0.9 * orc_pow_m02(err));
            step_rejected = 1;
        }
    }
    r->t = t_new;
    r->h_abs = h_abs;
    for (int i = 0; i < n; ++i) { r->y[i] = y_new[i]; r->f[i] = K[6][i]; }
    if (r->t - r->t_bound >= 0) r->status = ORC_FINISHED;
    return 0;
}
