/*
 * rcg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See rcg_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off: no FMA contraction, so every + and * rounds exactly like the
 * reference's numpy scalar arithmetic.
 *
 * "ref:" comments cite /root/reference (rcognita v0.1.2); "scipy:" comments cite
 * scipy/integrate/_ivp/ of scipy 1.18.1 (the RK45 the reference instantiates at
 * rcognita/simulator.py:150).
 */
#include "rcg_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------- deterministic elementary functions
 *
 * The reference takes sin/cos from numpy (a SIMD kernel or libm, depending on the CPU) and
 * err ** -0.2 from libm's pow; their last bits differ between libm builds, and through the
 * adaptive step size they decide the bits of the solver time t, hence on which solver step the
 * controller samples.  So that the CPU oracle and the CUDA kernels agree BIT FOR
 * BIT (same accepted-step times, same sampling pattern, on every lane), both implement the SAME
 * fully specified functions from IEEE +,-,*,fma only:
 *
 *  - orc_sincos: argument reduction by q = rint(x * 2/pi) with pi/2 split into three doubles,
 *    carried out in double-double (head rh, tail rl), then the published fdlibm / FreeBSD msun
 *    kernels k_sin.c / k_cos.c (Sun Microsystems, freely redistributable; < 1 ulp).
 *  - orc_pow_m02: x ** -0.2 as libm pow() followed by one correction step from the
 *    double-double residual y^5 * x - 1 (plus the term for 0.2 != 1/5 in binary): the correctly
 *    rounded result independent of libm's last bits (up to astronomically rare hard cases),
 *    which is also what glibc's pow returns in all but ~1e-5 of the calls.
 *
 * Both are pinned, like everything else here, by the live-reference golden vectors.
 */
static const double PIO2_1 = 1.5707963267948966e+00;   /* double(pi/2)                    */
static const double PIO2_2 = 6.123233995736766e-17;    /* double(pi/2 - PIO2_1)           */
static const double PIO2_3 = -1.4973849048591698e-33;  /* double(pi/2 - PIO2_1 - PIO2_2)  */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;

void orc_sincos(double x, double *sn, double *cs)
{
    static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                        S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                        S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                        C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                        C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    if (!(fabs(x) <= 1.0e5)) {           /* huge, inf, NaN: outside the regime of the path */
        *sn = sin(x);
        *cs = cos(x);
        return;
    }
    const double q = rint(x * TWO_OVER_PI);
    /* r = x - q*pi/2 in double-double */
    const double ph = q * PIO2_1, pl = fma(q, PIO2_1, -ph);
    const double r = x - ph;
    double t = fma(-q, PIO2_2, -pl);
    t = fma(-q, PIO2_3, t);
    const double rh = r + t;
    const double bb = rh - r;
    const double rl = (r - (rh - bb)) + (t - bb);           /* two-sum */
    /* kernels on [-pi/4, pi/4] */
    const double z = rh * rh, w = z * z, v = z * rh;
    const double ps = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    const double ks = rh - ((z * (0.5 * rl - v * ps) - rl) - v * S1);
    const double pc = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z, w1 = 1.0 - hz;
    const double kc = w1 + (((1.0 - w1) - hz) + (z * pc - rh * rl));
    switch (((long long)q) & 3) {
    case 0:  *sn = ks;  *cs = kc;  break;
    case 1:  *sn = kc;  *cs = -ks; break;
    case 2:  *sn = -ks; *cs = -kc; break;
    default: *sn = -kc; *cs = ks;  break;
    }
}

double orc_pow_m02(double x)
{
    if (!(x > 0.0 && x < INFINITY)) return pow(x, -0.2);
    const double y = pow(x, -0.2);
    const double ah = y * y, al = fma(y, y, -ah);                       /* y^2          */
    const double bh = ah * ah, bl = fma(ah, ah, -bh) + 2.0 * (ah * al); /* y^4          */
    const double ch = bh * y, cl = fma(bh, y, -ch) + bl * y;            /* y^5          */
    const double dh = ch * x, dl = fma(ch, x, -dh) + cl * x;            /* y^5 * x ~ 1  */
    const double g = (dh - 1.0) + dl;
    /* exponent is the double 0.2 = 1/5 + 1.1102230246251565e-17 */
    return y + (-y) * (0.2 * g + 1.1102230246251565e-17 * log(x));
}

/* ------------------------------------------------------------------ systems */

/* ref: rcognita/systems.py:308-323 (Sys3WRobot), :370-382 (Sys3WRobotNI),
 * :412-419 (Sys2Tank); is_disturb = 0 branch (presets hard-code it,
 * presets/main_3wrobot_NI.py:186). Operation order is the reference's. */
void orc_state_dyn(const orc_sys_t *s, const double *state, const double *action, double *d)
{
    switch (s->sys_id) {
    case ORC_SYS_3WROBOT_NI: {
        double sn, cs;
        orc_sincos(state[2], &sn, &cs);
        d[0] = action[0] * cs;
        d[1] = action[0] * sn;
        d[2] = action[1];
        break;
    }
    case ORC_SYS_3WROBOT: {
        double m = s->pars[0], I = s->pars[1];
        double sn, cs;
        orc_sincos(state[2], &sn, &cs);
        d[0] = state[3] * cs;
        d[1] = state[3] * sn;
        d[2] = state[4];
        d[3] = 1 / m * action[0];      /* Python: (1/m) * F */
        d[4] = 1 / I * action[1];
        break;
    }
    case ORC_SYS_2TANK: {
        double tau1 = s->pars[0], tau2 = s->pars[1], K1 = s->pars[2], K2 = s->pars[3], K3 = s->pars[4];
        d[0] = 1 / (tau1) * (-state[0] + K1 * action[0]);
        d[1] = 1 / (tau2) * (-state[1] + K2 * state[0] + K3 * (state[1] * state[1]));
        break;
    }
    default:
        break;
    }
}

/* ref: rcognita/systems.py:213-253 with is_disturb = is_dyn_ctrl = 0: clip the stored
 * action IN PLACE (:241-243), then _state_dyn (:245). */
void orc_closed_loop_rhs(const orc_sys_t *s, const double *y, double *action, double *rhs)
{
    if (s->has_bnds) {
        for (int k = 0; k < s->m; ++k) {     /* np.clip = minimum(maximum(a, lo), hi) */
            double a = action[k];
            if (a < s->lo[k]) a = s->lo[k];
            if (a > s->hi[k]) a = s->hi[k];
            action[k] = a;
        }
    }
    orc_state_dyn(s, y, action, rhs);
}

/* ------------------------------------------------------------------- RK45 */

/* scipy: rk.py:538-553 (class RK45 tableau). */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* scipy: rk.py:85-103 (RungeKutta.__init__ with first_step given) as called from
 * ref: rcognita/simulator.py:150.  f = fun(t0, y0) is evaluated HERE, with whatever the
 * system's stored action is at construction (zeros, ref: systems.py:134). */
void orc_rk45_init(orc_rk45_t *r, const orc_sys_t *s, const double *y0, double *action,
                   double t0, double t_bound, double max_step, double first_step,
                   double rtol, double atol)
{
    memset(r, 0, sizeof(*r));
    r->t = t0;
    r->t_bound = t_bound;
    r->max_step = max_step;
    r->rtol = rtol;
    r->atol = atol;
    r->h_abs = first_step;
    for (int i = 0; i < s->n; ++i) r->y[i] = y0[i];
    orc_closed_loop_rhs(s, r->y, action, r->f);
    r->nfev = 1;
    r->status = ORC_RUNNING;
}

/* scipy: base.py:179-212 (OdeSolver.step), rk.py:111-176 (_step_impl), rk.py:61-71 (rk_step),
 * rk.py:105-109 + common.py:63-65 (RMS error norm).  Returns 0 on an accepted step or on
 * the t == t_bound corner case, -1 if called on a non-running solver (scipy raises
 * RuntimeError), 1 if the step failed (TOO_SMALL_STEP -> status 'failed'). */
int orc_rk45_step(orc_rk45_t *r, const orc_sys_t *s, double *action)
{
    const int n = s->n;
    if (r->status != ORC_RUNNING) return -1;
    if (r->t == r->t_bound) {               /* base.py:192-197 */
        r->status = ORC_FINISHED;
        return 0;
    }
    const double t = r->t;
    const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);   /* rk.py:118 */
    double h_abs;
    if (r->h_abs > r->max_step) h_abs = r->max_step;                 /* rk.py:120-125 */
    else if (r->h_abs < min_step) h_abs = min_step;
    else h_abs = r->h_abs;

    int step_rejected = 0;
    double K[7][ORC_MAX_N], y_new[ORC_MAX_N], ytmp[ORC_MAX_N];
    double t_new, h;
    for (;;) {
        if (h_abs < min_step) {             /* rk.py:131-132 */
            r->status = ORC_FAILED;
            return 1;
        }
        h = h_abs;                          /* direction = +1 */
        t_new = t + h;
        if (t_new - r->t_bound > 0) t_new = r->t_bound;   /* rk.py:137-138 */
        h = t_new - t;
        h_abs = fabs(h);

        /* rk_step, rk.py:61-71.  K[0] = f is the FSAL derivative carried over from the
         * previous accepted step -- NOT recomputed even if the action changed since. */
        for (int i = 0; i < n; ++i) K[0][i] = r->f[i];
        for (int st = 1; st < 6; ++st) {
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;                          /* np.dot(K[:s].T, a[:s]) */
                for (int j = 0; j < st; ++j) acc += K[j][i] * RK_A[st][j];
                ytmp[i] = r->y[i] + acc * h;               /* y + dy, dy = dot * h   */
            }
            orc_closed_loop_rhs(s, ytmp, action, K[st]);
        }
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;                              /* np.dot(K[:-1].T, B)    */
            for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
            y_new[i] = r->y[i] + h * acc;
        }
        orc_closed_loop_rhs(s, y_new, action, K[6]);
        r->nfev += 6;

        double sq = 0.0;
        for (int i = 0; i < n; ++i) {
            double scale = r->atol + fmax(fabs(r->y[i]), fabs(y_new[i])) * r->rtol;   /* rk.py:146 */
            double acc = 0.0;                              /* np.dot(K.T, E) * h     */
            for (int j = 0; j < 7; ++j) acc += K[j][i] * RK_E[j];
            double e = acc * h / scale;
            sq += e * e;
        }
        double err = sqrt(sq) / sqrt((double)n);           /* common.py:65           */

        if (err < 1) {                                     /* rk.py:149-160          */
            double factor;
            if (err == 0) factor = 10;
            else factor = fmin(10, 0.9 * orc_pow_m02(err));
            if (step_rejected) factor = fmin(1, factor);
            h_abs *= factor;
            break;
        } else {                                           /* rk.py:161-164          */
            h_abs *= fmax(0.2, 0.9 * orc_pow_m02(err));
            step_rejected = 1;
        }
    }
    r->t = t_new;
    r->h_abs = h_abs;
    for (int i = 0; i < n; ++i) { r->y[i] = y_new[i]; r->f[i] = K[6][i]; }
    if (r->t - r->t_bound >= 0) r->status = ORC_FINISHED;  /* base.py:207-208        */
    return 0;
}

/* -------------------------------------------------------------- controller */

/* ref: rcognita/controllers.py:1024-1039 */
int orc_dim_critic(int critic_struct, int n, int m)
{
    int p = n + m;
    switch (critic_struct) {
    case ORC_CRITIC_QUAD_LIN:   return (p + 1) * p / 2 + p;
    case ORC_CRITIC_QUADRATIC:  return (p + 1) * p / 2;
    case ORC_CRITIC_QUAD_NOMIX: return p;
    case ORC_CRITIC_QUAD_MIX:   return n + n * m + m;
    default: return 0;
    }
}

static void make_chi(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act, double *chi)
{
    /* ref: controllers.py:1069-1072, :1200-1203 */
    for (int i = 0; i < n; ++i) chi[i] = c->has_target ? obs[i] - c->target[i] : obs[i];
    for (int j = 0; j < m; ++j) chi[n + j] = act[j];
}

static double quad_form(const double *x, const double *R, int p)
{
    /* x @ R @ x evaluated left to right: v = x @ R (row vector), then v @ x. */
    double v[ORC_MAX_P];
    for (int j = 0; j < p; ++j) {
        double acc = 0.0;
        for (int i = 0; i < p; ++i) acc += x[i] * R[i * p + j];
        v[j] = acc;
    }
    double out = 0.0;
    for (int j = 0; j < p; ++j) out += v[j] * x[j];
    return out;
}

/* ref: rcognita/controllers.py:1063-1084 (stage_obj, a.k.a. rcost). */
double orc_stage_obj(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act)
{
    const int p = n + m;
    double chi[ORC_MAX_P];
    make_chi(c, n, m, obs, act, chi);
    if (c->stage_struct == ORC_STAGE_QUADRATIC) {
        return quad_form(chi, c->R1, p);                               /* :1078-1079 */
    } else {
        double chi2[ORC_MAX_P];
        for (int i = 0; i < p; ++i) chi2[i] = chi[i] * chi[i];
        return quad_form(chi2, c->R2, p) + quad_form(chi, c->R1, p);   /* :1082      */
    }
}

/* ref: rcognita/controllers.py:1192-1214 (_critic); feature order of uptria2vec is
 * row-major i <= j (ref: rcognita/utilities.py:81-96); np.kron(obs, act)[i*m+j] = obs_i*act_j.
 * NB quad-mix uses the RAW observation, not the target-shifted chi (:1212). */
double orc_critic(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act, const double *w)
{
    const int p = n + m;
    double chi[ORC_MAX_P], phi[ORC_MAX_W];
    int k = 0;
    make_chi(c, n, m, obs, act, chi);
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
    case ORC_CRITIC_QUAD_MIX:
        for (int i = 0; i < n; ++i) phi[k++] = obs[i] * obs[i];
        for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) phi[k++] = obs[i] * act[j];
        for (int j = 0; j < m; ++j) phi[k++] = act[j] * act[j];
        break;
    default:
        break;
    }
    double q = 0.0;
    for (int i = 0; i < k; ++i) q += w[i] * phi[i];
    return q;
}

/* ref: rcognita/controllers.py:1216-1245 (_critic_cost). obs_buf [buffer_size, n] and
 * act_buf [buffer_size, m] row-major; rows 0..Ncritic-1 are read, i.e. the OLDEST rows of the
 * bottom-pushed FIFO (push_vec, ref: rcognita/utilities.py:78-79). */
double orc_critic_cost(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                       const double *w, const double *w_prev)
{
    double Jc = 0.0;
    for (int k = c->Ncritic - 1; k > 0; --k) {
        const double *o_prev = obs_buf + (k - 1) * n, *o_next = obs_buf + k * n;
        const double *a_prev = act_buf + (k - 1) * m, *a_next = act_buf + k * m;
        double q_prev = orc_critic(c, n, m, o_prev, a_prev, w);
        double q_next = orc_critic(c, n, m, o_next, a_next, w_prev);
        double e = q_prev - c->gamma * q_next - orc_stage_obj(c, n, m, o_prev, a_prev);
        Jc += 1.0 / 2 * (e * e);
    }
    return Jc;
}

/* ref: rcognita/controllers.py:1273-1328 (_actor_cost), is_est_model = 0 branch.
 * action_sqn is [Nactor, m] row-major (np.reshape at :1282).  The Euler predictor calls the
 * UNCLIPPED _state_dyn (sys_rhs = my_sys._state_dyn, presets/main_3wrobot_NI.py:247) and starts
 * from state_sys while observation_sqn[0] = observation (:1290-1291). */
double orc_actor_cost(const orc_ctrl_t *c, const orc_sys_t *s, const double *action_sqn,
                      const double *observation, const double *state_sys, const double *w_critic)
{
    const int n = s->n, m = s->m, N = c->Nactor;
    double obs_sqn[ORC_MAX_NACTOR][ORC_MAX_N];
    double state[ORC_MAX_N], d[ORC_MAX_N];
    for (int i = 0; i < n; ++i) { obs_sqn[0][i] = observation[i]; state[i] = state_sys[i]; }
    for (int k = 1; k < N; ++k) {
        orc_state_dyn(s, state, action_sqn + (k - 1) * m, d);
        for (int i = 0; i < n; ++i) {
            state[i] = state[i] + c->pred_step_size * d[i];           /* Euler, :1294 */
            obs_sqn[k][i] = state[i];                                 /* sys_out = identity */
        }
    }
    double J = 0.0;
    if (c->mode == ORC_MODE_MPC) {
        for (int k = 0; k < N; ++k)
            J += pow(c->gamma, (double)k) * orc_stage_obj(c, n, m, obs_sqn[k], action_sqn + k * m);   /* :1305-1306 */
    } else if (c->mode == ORC_MODE_RQL) {
        for (int k = 0; k < N - 1; ++k)
            J += pow(c->gamma, (double)k) * orc_stage_obj(c, n, m, obs_sqn[k], action_sqn + k * m);   /* :1308-1309 */
        J += orc_critic(c, n, m, obs_sqn[N - 1], action_sqn + (N - 1) * m, w_critic);                 /* :1310 */
    } else {
        for (int k = 0; k < N; ++k)
            J += orc_critic(c, n, m, obs_sqn[k], action_sqn + k * m, w_critic);                       /* :1312-1326 */
    }
    return J;
}

/* np.argmin: first minimal index; NaN counts as minimal (first NaN wins). */
int orc_argmin(const double *J, int count)
{
    int best = 0;
    if (count <= 0) return -1;
    if (isnan(J[0])) return 0;
    for (int i = 1; i < count; ++i) {
        if (isnan(J[i])) return i;
        if (J[i] < J[best]) best = i;
    }
    return best;
}

/* The protocol of SURVEY.md App. A.4: the reference's own _actor_cost on every row of a
 * candidate table, then np.argmin. */
void orc_actor_cost_table(const orc_ctrl_t *c, const orc_sys_t *s, int C, const double *cand,
                          const double *observation, const double *state_sys, const double *w_critic,
                          double *J_out, int *argmin_out)
{
    const int L = c->Nactor * s->m;
    for (int i = 0; i < C; ++i)
        J_out[i] = orc_actor_cost(c, s, cand + (long)i * L, observation, state_sys, w_critic);
    if (argmin_out) *argmin_out = orc_argmin(J_out, C);
}

/* ------------------------------------------------- resumable per-environment closed loop */

/* Construction of the reference objects for one environment: Simulator.__init__
 * (ref: rcognita/simulator.py:150 -> RK45 ctor with System.action = zeros, systems.py:134) and
 * CtrlOptPred.__init__ (action_curr = action_init, ctrl_clock = t0, controllers.py:973-990). */
void orc_env_init(orc_env_t *envs, const orc_sys_t *s, int E, const double *state_init,
                  const double *action_init, double t0, double t1, double max_step, double first_step,
                  double rtol, double atol)
{
    for (int e = 0; e < E; ++e) {
        orc_env_t *v = envs + e;
        memset(v, 0, sizeof(*v));
        for (int j = 0; j < s->m; ++j) v->action_curr[j] = action_init[j];
        for (int i = 0; i < s->n; ++i) v->state_sys[i] = state_init[(long)e * s->n + i];
        orc_rk45_init(&v->r, s, state_init + (long)e * s->n, v->sys_action, t0, t1, max_step, first_step, rtol, atol);
        v->ctrl_clock = t0;
        v->best = -1;
        v->Jbest = NAN;
    }
}

/* ONE iteration of the headless main loop for one environment.
 * ref: presets/main_3wrobot_NI.py:415-440 + controllers.py:1429-1493 (compute_action, MPC /
 * fixed-critic RQL,SQL) + :1056 receive_sys_state + :1086 upd_accum_obj, with
 * _actor_optimizer replaced by enumerate-and-argmin.
 * Returns 1 if the controller sampled in this iteration, 0 if it held, -1 if the env is done. */
static int env_iterate(orc_env_t *v, const orc_ctrl_t *c, const orc_sys_t *s, int C, const double *tab,
                       const double *w_critic, double sampling_time, double t1)
{
    const int n = s->n, m = s->m, L = c->Nactor * m;
    int sampled = 0;
    double Jtab[4096];
    if (v->done) return -1;
    if (orc_rk45_step(&v->r, s, v->sys_action) != 0) { v->done = 1; return -1; }   /* sim_step */
    ++v->steps;
    const double t = v->r.t;
    const double *obs = v->r.y;                                /* out() = identity            */
    if (t - v->ctrl_clock >= sampling_time) {                  /* controllers.py:1440-1442    */
        v->ctrl_clock = t;
        for (int i0 = 0; i0 < C; i0 += 4096) {
            int cnt = C - i0 < 4096 ? C - i0 : 4096;
            int bi;
            orc_actor_cost_table(c, s, cnt, tab + (long)i0 * L, obs, v->state_sys, w_critic, Jtab, &bi);
            /* strict '<' keeps the first minimum across chunks; NaN wins once */
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
    if (t >= t1) v->done = 1;                                  /* main_3wrobot_NI.py:440      */
    return sampled;
}

/* Every environment that is not done iterates the main loop up to and including its next
 * controller sample (or its end).  Returns the accepted solver steps taken; *evals (may be
 * NULL) receives the number of _actor_cost evaluations. */
long long orc_env_interval(orc_env_t *envs, const orc_ctrl_t *c, const orc_sys_t *s, int E, int C,
                           const double *cand, int cand_per_env, const double *w_critic,
                           double sampling_time, double t1, int nthreads, long long *evals_out)
{
    const int L = c->Nactor * s->m;
    long long total_steps = 0, evals = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total_steps, evals)
    for (int e = 0; e < E; ++e) {
        orc_env_t *v = envs + e;
        const double *tab = cand_per_env ? cand + (long)e * C * L : cand;
        for (;;) {
            int rc = env_iterate(v, c, s, C, tab, w_critic, sampling_time, t1);
            if (rc < 0) break;
            ++total_steps;
            if (rc == 1) { evals += C; break; }
            if (v->done) break;
        }
    }
    if (evals_out) *evals_out = evals;
    return total_steps;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

/* Whole episodes (the closed loop above run to t1) for E environments, OpenMP-parallel. */
long long orc_closed_loop(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *state_init,
                          int C, const double *cand, int cand_per_env, const double *w_critic,
                          const double *action_init, double sampling_time,
                          double t0, double t1, double max_step, double first_step,
                          double rtol, double atol, int max_steps_per_env, int nthreads,
                          double *y_final, double *t_final, double *accum, int *nsteps,
                          int *nsamples, long *nfev, double *traj, int traj_cap, int *traj_rows,
                          long long *total_evals)
{
    const int n = s->n, m = s->m, L = c->Nactor * m;
    long long total_steps = 0, evals = 0;
    if (traj_rows) *traj_rows = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_steps, evals)
    for (int e = 0; e < E; ++e) {
        orc_env_t v;
        const double *tab = cand_per_env ? cand + (long)e * C * L : cand;
        orc_env_init(&v, s, 1, state_init + (long)e * n, action_init, t0, t1, max_step, first_step, rtol, atol);
        while (v.steps < max_steps_per_env) {
            int rc = env_iterate(&v, c, s, C, tab, w_critic, sampling_time, t1);
            if (rc < 0) break;
            if (rc == 1) evals += C;
            if (e == 0 && traj && v.steps <= traj_cap) {
                double *row = traj + (long)(v.steps - 1) * (1 + n + m + 3);
                row[0] = v.r.t;
                for (int i = 0; i < n; ++i) row[1 + i] = v.r.y[i];
                for (int j = 0; j < m; ++j) row[1 + n + j] = v.action_curr[j];
                row[1 + n + m] = v.accum;
                row[2 + n + m] = (double)v.best;
                row[3 + n + m] = v.Jbest;
                if (traj_rows) *traj_rows = v.steps;
            }
            if (v.done) break;
        }
        if (y_final) for (int i = 0; i < n; ++i) y_final[(long)e * n + i] = v.r.y[i];
        if (t_final) t_final[e] = v.r.t;
        if (accum) accum[e] = v.accum;
        if (nsteps) nsteps[e] = v.steps;
        if (nsamples) nsamples[e] = v.samples;
        if (nfev) nfev[e] = v.r.nfev;
        total_steps += v.steps;
    }
    if (total_evals) *total_evals = evals;
    return total_steps;
}
