/*
 * rcg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, plain-C restatement of rcognita's hot path (reference v0.1.2 at
 * /root/reference) and of scipy.integrate.RK45 (scipy 1.18.1, not vendored by the
 * reference; de-facto spec = scipy/integrate/_ivp/rk.py).  Every function cites the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use this library.  The product
 * (rcognita_b200/) never links, imports or calls it.
 *
 * Parity status: the reference ships NO tests or golden vectors, so this
 * oracle is pinned against outputs of the live reference generated in the build
 * container by tests/golden/make_golden.py (fixtures committed under tests/golden/).
 *
 * Vectors follow the reference's conventions: shape [n] row vectors, buffers [L, n]
 * row-major, action sequences [Nactor, dim_input] row-major.
 */
#ifndef RCG_ORACLE_H
#define RCG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_N 5   /* dim_state  <= 5 (Sys3WRobot)          */
#define ORC_MAX_M 2   /* dim_input  <= 2                       */
#define ORC_MAX_P 7   /* n + m                                 */
#define ORC_MAX_W 35  /* dim_critic <= 35 (quad-lin, p = 7)    */
#define ORC_MAX_NACTOR 64

enum { ORC_SYS_3WROBOT_NI = 0, ORC_SYS_3WROBOT = 1, ORC_SYS_2TANK = 2 };
enum { ORC_MODE_MPC = 0, ORC_MODE_RQL = 1, ORC_MODE_SQL = 2 };
enum { ORC_CRITIC_QUAD_LIN = 0, ORC_CRITIC_QUADRATIC = 1, ORC_CRITIC_QUAD_NOMIX = 2, ORC_CRITIC_QUAD_MIX = 3 };
enum { ORC_STAGE_QUADRATIC = 0, ORC_STAGE_BIQUADRATIC = 1 };
enum { ORC_RUNNING = 0, ORC_FINISHED = 1, ORC_FAILED = 2 };

/* rcognita/systems.py:69-145 -- the fields of System that the path reads. */
typedef struct {
    int sys_id;             /* ORC_SYS_*  (System.name, systems.py:301,363,410) */
    int n, m;               /* dim_state (= dim_output), dim_input              */
    int has_bnds;           /* ctrl_bnds.any()  (systems.py:241)                */
    double pars[8];         /* 3wrobot [m, I]; 2tank [tau1,tau2,K1,K2,K3]       */
    double lo[ORC_MAX_M];   /* ctrl_bnds[:,0]                                   */
    double hi[ORC_MAX_M];   /* ctrl_bnds[:,1]                                   */
} orc_sys_t;

/* rcognita/controllers.py:811-1042 -- the fields of CtrlOptPred that the costs read. */
typedef struct {
    int mode;               /* ORC_MODE_*                                       */
    int critic_struct;      /* ORC_CRITIC_*                                     */
    int stage_struct;       /* ORC_STAGE_*                                      */
    int has_target;         /* observation_target != []                         */
    int Nactor;
    int Ncritic;            /* already clipped to buffer_size-1 (:1015)         */
    double gamma;
    double pred_step_size;
    double R1[ORC_MAX_P * ORC_MAX_P];   /* row-major [p,p], stage_obj_pars[0]   */
    double R2[ORC_MAX_P * ORC_MAX_P];   /* stage_obj_pars[1] (biquadratic only) */
    double target[ORC_MAX_N];
} orc_ctrl_t;

/* scipy RK45 instance state as used by rcognita/simulator.py:150-168. */
typedef struct {
    double t, t_bound, h_abs, max_step, rtol, atol;
    double y[ORC_MAX_N], f[ORC_MAX_N];
    int status;             /* ORC_RUNNING / FINISHED / FAILED                  */
    long nfev;
} orc_rk45_t;

/* One environment of the headless closed loop (presets/main_3wrobot_NI.py:415-440): the solver,
 * System.action, and the CtrlOptPred fields the loop mutates. */
typedef struct {
    orc_rk45_t r;
    double sys_action[ORC_MAX_M];   /* System.action (clipped in place by closed_loop_rhs) */
    double action_curr[ORC_MAX_M];  /* CtrlOptPred.action_curr                              */
    double state_sys[ORC_MAX_N];    /* CtrlOptPred.state_sys (one solver step behind)       */
    double ctrl_clock, accum, Jbest;
    int steps, samples, best, done;
} orc_env_t;

/* Deterministic sin/cos and x ** -0.2 shared (as a specification) with the CUDA kernels. */
void   orc_sincos(double x, double *sn, double *cs);
double orc_pow_m02(double x);
int    orc_dim_critic(int critic_struct, int n, int m);
void   orc_state_dyn(const orc_sys_t *s, const double *state, const double *action, double *dstate);
void   orc_closed_loop_rhs(const orc_sys_t *s, const double *y, double *action, double *rhs);
void   orc_rk45_init(orc_rk45_t *r, const orc_sys_t *s, const double *y0, double *action,
                     double t0, double t_bound, double max_step, double first_step,
                     double rtol, double atol);
int    orc_rk45_step(orc_rk45_t *r, const orc_sys_t *s, double *action);
double orc_stage_obj(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act);
double orc_critic(const orc_ctrl_t *c, int n, int m, const double *obs, const double *act, const double *w);
double orc_critic_cost(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                       const double *w, const double *w_prev);
double orc_actor_cost(const orc_ctrl_t *c, const orc_sys_t *s, const double *action_sqn,
                      const double *observation, const double *state_sys, const double *w_critic);
int    orc_argmin(const double *J, int count);
void   orc_actor_cost_table(const orc_ctrl_t *c, const orc_sys_t *s, int C, const double *cand,
                            const double *observation, const double *state_sys, const double *w_critic,
                            double *J_out, int *argmin_out);

/* Closed loop of presets/main_3wrobot_NI.py:415-440 with the optimiser replaced by
 * enumerate-and-argmin over a candidate table, for E independent
 * environments, OpenMP-parallel over environments.  cand_per_env: 0 -> cand is one shared
 * [C, N*m] table, 1 -> cand is [E, C, N*m].  w_critic is a fixed [dim_critic] vector (or NULL
 * for MPC).  Outputs (any may be NULL): y_final [E,n], t_final [E], accum [E], nsteps [E],
 * nsamples [E], nfev [E]; traj (optional) records for env 0 up to traj_cap rows of
 * [t, y(n), action(m), accum, argmin, Jmin]. Returns total accepted steps over all envs. */
long long orc_closed_loop(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *state_init,
                          int C, const double *cand, int cand_per_env, const double *w_critic,
                          const double *action_init, double sampling_time,
                          double t0, double t1, double max_step, double first_step,
                          double rtol, double atol, int max_steps_per_env, int nthreads,
                          double *y_final, double *t_final, double *accum, int *nsteps,
                          int *nsamples, long *nfev, double *traj, int traj_cap, int *traj_rows,
                          long long *total_evals);

/* Resumable form of the same loop (bench cpu_baseline / reference arm, interval-level tests):
 * orc_env_init builds E environments; orc_env_interval advances every live environment up to
 * and including its next controller sample. */
void      orc_env_init(orc_env_t *envs, const orc_sys_t *s, int E, const double *state_init,
                       const double *action_init, double t0, double t1, double max_step, double first_step,
                       double rtol, double atol);
long long orc_env_interval(orc_env_t *envs, const orc_ctrl_t *c, const orc_sys_t *s, int E, int C,
                           const double *cand, int cand_per_env, const double *w_critic,
                           double sampling_time, double t1, int nthreads, long long *evals_out);
/* rcg_oracle_opt.c: analytic gradient of _actor_cost (adjoint of the Euler rollout) and the bounded
 * minimiser standing in for CtrlOptPred._actor_optimizer (ref: controllers.py:1330-1427). */
double orc_actor_grad(const orc_ctrl_t *c, const orc_sys_t *s, const double *action_sqn, const double *observation,
                      const double *state_sys, const double *w_critic, double *grad);
void orc_actor_opt_set_lanes(int lanes);
double orc_actor_opt(const orc_ctrl_t *c, const orc_sys_t *s, double *x, const double *observation,
                     const double *state_sys, const double *w_critic, int max_iter, double pg_tol, double f_tol,
                     int *iters_out, int *nfev_out);
double orc_actor_opt_hybrid(const orc_ctrl_t *c, const orc_sys_t *s, double *x, const double *observation,
                            const double *state_sys, const double *w_critic, int max_sweeps, int max_iter, double pg_tol,
                            double f_tol, int *sweeps_out, int *iters_out);
long long orc_actor_opt_batch(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *x_init, const double *states,
                              const double *w_critic, int max_iter, double pg_tol, double f_tol, int nthreads,
                              double *J_out);
/* rcg_oracle_critic.c: the critic side of compute_action for RQL / SQL (ref: controllers.py:1248-1271, :1455-1479). */
#define ORC_MAX_BUF 32
typedef struct {
    double obs_buf[ORC_MAX_BUF * ORC_MAX_N];    /* CtrlOptPred.observation_buffer [buffer_size, n], row 0 = oldest */
    double act_buf[ORC_MAX_BUF * ORC_MAX_M];    /* CtrlOptPred.action_buffer [buffer_size, m]                      */
    double w[ORC_MAX_W], w_prev[ORC_MAX_W];     /* w_critic, w_critic_prev                                         */
    double critic_clock, Jc;
    int nfits;
} orc_critic_state_t;
double orc_critic_fit(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                      const double *w_prev, double lo, double hi, const double *w_init, double *w_out,
                      int max_evals, int *evals_out);
double orc_critic_fit_ls(const orc_ctrl_t *c, int n, int m, const double *obs_buf, const double *act_buf,
                         const double *w_prev, double lo, double hi, const double *w_init, double *w_out,
                         int max_evals, int ls_mode, int *evals_out);
void   orc_critic_state_init(orc_critic_state_t *k, int dimc, double t0);
int    orc_env_iterate_critic(orc_env_t *v, orc_critic_state_t *k, const orc_ctrl_t *c, const orc_sys_t *s, int C,
                              const double *tab, int buffer_size, double w_lo, double w_hi, double sampling_time,
                              double critic_period, double t1, const double *w_replay, int n_replay);
long long orc_closed_loop_critic(const orc_ctrl_t *c, const orc_sys_t *s, int E, const double *state_init, int C,
                                 const double *cand, int cand_per_env, const double *action_init, int buffer_size,
                                 double w_lo, double w_hi, double sampling_time, double critic_period, double t0, double t1,
                                 double max_step, double first_step, double rtol, double atol, int max_steps_per_env,
                                 int nthreads, const double *w_replay, int n_replay, double *y_final, double *t_final,
                                 double *accum, int *nsteps, int *nsamples, int *nfits, double *w_final, double *Jc_final,
                                 double *obs_buf_final, double *act_buf_final, double *traj, int traj_cap, int *traj_rows);
/* rcg_oracle_disturb.c: disturbance lanes, System(is_disturb = 1) (ref: systems.py:228-231, :247-248, :316-318, :373-376,
 * :325-345, :384-394); the normal draws come from a counter-based stream shared (as a specification) with the CUDA kernels. */
#define ORC_MAX_NFULL 7
typedef struct {
    double sigma[2], mu[2], tau[2];     /* pars_disturb = [sigma_disturb, mu_disturb, tau_disturb] */
    unsigned long long seed;
} orc_dist_t;
typedef struct {
    double t, t_bound, h_abs, max_step, rtol, atol;
    double y[ORC_MAX_NFULL], f[ORC_MAX_NFULL];
    int status, nfull;
    long nfev;
    unsigned long long env;
} orc_rk45d_t;
void   orc_state_dyn_disturbed(const orc_sys_t *s, const double *x, const double *a, const double *q, double *d);
void   orc_disturb_dyn(const orc_sys_t *s, const orc_dist_t *D, const double *q, const double *z, double *dq);
double orc_log(double x);
void   orc_normal2(unsigned long long seed, unsigned long long env, unsigned int call, double *z);
void   orc_closed_loop_rhs_disturbed(const orc_sys_t *s, const orc_dist_t *D, const double *y_full, double *action,
                                     unsigned long long env, unsigned int call, const double *z_given, double *rhs);
void   orc_rk45d_init(orc_rk45d_t *r, const orc_sys_t *s, const orc_dist_t *D, unsigned long long env, const double *y0_full,
                      double *action, double t0, double t_bound, double max_step, double first_step, double rtol, double atol);
int    orc_rk45d_step(orc_rk45d_t *r, const orc_sys_t *s, const orc_dist_t *D, double *action);
void   orc_nominal_ni(double ctrl_gain, const orc_sys_t *s, const double *obs, double *action);
int       orc_num_threads(void);
int       orc_has_openmp(void);

#ifdef __cplusplus
}
#endif
#endif
