/*
 * rcg.h -- C ABI of librcg_b200.so: the B200 (sm_100a) batched agent-environment engine
 * for rcognita's data-parallel hot path.
 *
 * The reference (AIDynamicAction/rcognita v0.1.2) is pure Python and has no FFI seam; its
 * boundary for this path is a set of Python methods.  Each entry point below replaces the
 * per-environment arithmetic of one of those methods for E independent environments
 * ("lanes") at once; the Python mirror classes in rcognita_b200/ (same names and
 * signatures as the reference) are the only intended callers and bind these symbols with
 * ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *  - All array arguments are DEVICE pointers (cudaMalloc / torch CUDA tensors); the two
 *    descriptor structs are HOST pointers, read during the call.
 *  - Struct-of-arrays, component-major: a per-lane vector quantity v of dimension d is
 *    stored as v[i * E + e] (i < d component, e < E lane), so a warp reads 32 consecutive
 *    lanes of one component with one coalesced 256-byte request.
 *  - fp64 entry points have no suffix; "_f32" twins take float arrays for states, costs
 *    and weights but keep every time-like quantity (t, h_abs, ctrl_clock) in fp64.
 *  - All launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream).  Return value: 0 on success, otherwise a cudaError_t or a
 *    negative RCG_E* code; rcg_last_error_string() describes the last failure of the
 *    calling thread.
 *  - Nothing here falls back to the CPU: without a CUDA device every compute entry point
 *    returns an error.
 */
#ifndef RCG_H
#define RCG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCG_VERSION 100          /* 0.1.0 */

#define RCG_MAX_N 5              /* dim_state = dim_output <= 5 (Sys3WRobot)  */
#define RCG_MAX_M 2              /* dim_input <= 2                            */
#define RCG_MAX_P 7              /* n + m                                     */
#define RCG_MAX_W 35             /* dim_critic <= 35 (quad-lin, p = 7)        */
#define RCG_MAX_NACTOR 64        /* prediction horizon Nactor <= 64           */

/* System.name of the reference (rcognita/systems.py:363, :301, :410). */
enum { RCG_SYS_3WROBOT_NI = 0, RCG_SYS_3WROBOT = 1, RCG_SYS_2TANK = 2 };
/* CtrlOptPred.mode (rcognita/controllers.py:1304-1326). */
enum { RCG_MODE_MPC = 0, RCG_MODE_RQL = 1, RCG_MODE_SQL = 2 };
/* CtrlOptPred.critic_struct (rcognita/controllers.py:1204-1212). */
enum { RCG_CRITIC_QUAD_LIN = 0, RCG_CRITIC_QUADRATIC = 1, RCG_CRITIC_QUAD_NOMIX = 2, RCG_CRITIC_QUAD_MIX = 3 };
/* CtrlOptPred.stage_obj_struct (rcognita/controllers.py:1076-1082). */
enum { RCG_STAGE_QUADRATIC = 0, RCG_STAGE_BIQUADRATIC = 1 };
/* scipy OdeSolver.status per lane ('running' | 'finished' | 'failed', scipy base.py:189-208). */
enum { RCG_RUNNING = 0, RCG_FINISHED = 1, RCG_FAILED = 2 };

enum { RCG_EINVAL = -1, RCG_ENODEV = -2 };

/* The fields of rcognita.systems.System the path reads (rcognita/systems.py:69-145). */
typedef struct rcg_system {
    int32_t sys_id;              /* RCG_SYS_*                                              */
    int32_t has_bnds;            /* ctrl_bnds.any()  (systems.py:241)                      */
    double  pars[8];             /* Sys3WRobot [m, I]; Sys2Tank [tau1, tau2, K1, K2, K3]   */
    double  lo[RCG_MAX_M];       /* ctrl_bnds[:, 0]                                        */
    double  hi[RCG_MAX_M];       /* ctrl_bnds[:, 1]                                        */
} rcg_system_t;

/* The fields of rcognita.controllers.CtrlOptPred that stage_obj / _critic / _critic_cost /
 * _actor_cost read (rcognita/controllers.py:811-1042). */
typedef struct rcg_objective {
    int32_t mode;                /* RCG_MODE_*                                             */
    int32_t critic_struct;       /* RCG_CRITIC_*                                           */
    int32_t stage_struct;        /* RCG_STAGE_*                                            */
    int32_t r_is_diag;           /* 1: only the diagonals of R1/R2 are non-zero            */
    int32_t has_target;          /* observation_target != []                               */
    int32_t Nactor;              /* prediction horizon                                     */
    int32_t Ncritic;             /* critic stack size, already min(Ncritic, buffer_size-1) */
    int32_t buffer_size;         /* rows of the observation/action FIFO buffers            */
    double  gamma;               /* discounting factor                                     */
    double  pred_step_size;      /* Euler predictor step                                   */
    double  gamma_pow[RCG_MAX_NACTOR];   /* gamma**k, k < Nactor, computed by the host with
                                            libm pow() exactly like Python's float.__pow__ */
    double  R1[RCG_MAX_P * RCG_MAX_P];   /* stage_obj_pars[0], row-major [p, p]            */
    double  R2[RCG_MAX_P * RCG_MAX_P];   /* stage_obj_pars[1] (biquadratic only)           */
    double  target[RCG_MAX_N];           /* observation_target                             */
} rcg_objective_t;

/* scipy RK45 settings as passed at rcognita/simulator.py:150. */
typedef struct rcg_solver {
    double t_bound;              /* t1                                                     */
    double max_step;             /* dt / 2 (Simulator ignores its own max_step argument)   */
    double rtol, atol;
} rcg_solver_t;

/* Device-side trajectory ring (the rows of rcognita/loggers.py:41-94 for every environment): rows is
 * [capacity][1 + n + 2 + m][E] = (t, state[n], stage_obj, accum_obj, action[m]) -- the reference's column order for
 * Sys3WRobotNI and Sys3WRobot; Sys2Tank logs the action before stage_obj (loggers.py:90-94, reordered by the host).
 * Lane e has written count[e] rows so far; row r of lane e sits at slot r % capacity.  Only solver steps whose
 * per-lane index (nsteps) is a multiple of `every` are logged. */
typedef struct rcg_log {
    double  *rows;
    int32_t *count;
    int32_t  capacity;
    int32_t  every;
} rcg_log_t;

/* Disturbance lanes, System(is_disturb = 1) (rcognita/systems.py:139-145, :228-231, :247-248, :325-345, :384-394): the
 * full state is [state, disturb] with dim_disturb = 2 for the two robots and 1 for Sys2Tank (the presets' values);
 * pars_disturb = [sigma_disturb, mu_disturb, tau_disturb].  The reference draws randn() from numpy's global stream once
 * per component and RHS call -- irreproducible by construction; here every environment has a counter-based stream:
 * Philox4x32-10 block (RHS call number, global environment index) under key `seed`, Box-Muller with specified log /
 * sincos, so that a run does not depend on sharding and is reproduced bit for bit by the CPU checker. */
typedef struct rcg_disturb {
    double   sigma[2], mu[2], tau[2];
    uint64_t seed;
    int64_t  env_offset;         /* global index of lane 0 of this call (rank offset when the batch is sharded) */
} rcg_disturb_t;

int         rcg_version(void);
const char *rcg_last_error_string(void);
/* Number of CUDA devices visible, or a negative RCG_E* code (never touches a kernel). */
int         rcg_device_count(void);
/* dim_state / dim_input / dim_critic helpers (controllers.py:1024-1039). */
int         rcg_dim_state(int32_t sys_id);
int         rcg_dim_input(int32_t sys_id);
int         rcg_dim_critic(int32_t critic_struct, int32_t n, int32_t m);
/* Kernels launched by this library since load / since the last reset (bench "gpu_launches"). */
int64_t     rcg_launch_count(void);
void        rcg_reset_launch_count(void);

/* System.closed_loop_rhs (rcognita/systems.py:213-253), is_disturb = is_dyn_ctrl = 0:
 * clips action[m][E] IN PLACE to ctrl_bnds, then f_out[n][E] = _state_dyn(y, action)
 * (systems.py:308-323, :370-382, :412-419).  Used by Simulator.__init__ to seed the FSAL
 * derivative f = fun(t0, y0) that scipy evaluates at construction (scipy rk.py:97). */
int rcg_rhs(const rcg_system_t *sys, int64_t E, const double *y, double *action, double *f_out, void *stream);
int rcg_rhs_f32(const rcg_system_t *sys, int64_t E, const float *y, float *action, float *f_out, void *stream);

/* System._state_dyn alone (no clipping): dstate[n][E] = _state_dyn(state, action). */
int rcg_state_dyn(const rcg_system_t *sys, int64_t E, const double *state, const double *action,
                  double *dstate, void *stream);

/* System.closed_loop_rhs with is_disturb = 1 on the full state y_full[n + nd][E] (rcognita/systems.py:213-253): clips
 * action IN PLACE when clip != 0, f_out[:n] = _state_dyn(t, state, action, disturb) (:316-318, :373-376), f_out[n:] =
 * _disturb_dyn(t, disturb) (:341-343, :390-392) = -tau * (disturb + sigma * (z + mu)).  The draws z: normals[2][E] when
 * given (parity with the reference under a patched randn()), else the environment's stream at RHS call number call[e]
 * (0 when call is NULL).  clip = 0 and normals given evaluates _state_dyn / _disturb_dyn alone. */
int rcg_rhs_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, int64_t E, const double *y_full, double *action,
                      const int32_t *call, const double *normals, double *f_out, int32_t clip, void *stream);
/* The two standard-normal draws of RHS call number call[e] (call_all when call is NULL) of every environment:
 * normals[2][E]. */
int rcg_disturb_normals(const rcg_disturb_t *dist, int64_t E, const int32_t *call, int32_t call_all, double *normals,
                        void *stream);
/* rcg_rk45_step / rcg_rk45_advance on the full state [n + nd][E] of a disturbed system; nfev[E] is required: it numbers
 * the RHS calls, i.e. the random draws (1 after Simulator.__init__, += 6 per attempt). */
int rcg_rk45_step_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, const rcg_solver_t *sol, int64_t E,
                            double *y_full, double *f_full, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                            double *action, void *stream);
int rcg_rk45_advance_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, const rcg_solver_t *sol,
                               const rcg_objective_t *obj, int64_t E, double *y_full, double *f_full, double *t, double *h_abs,
                               int32_t *status, int32_t *nfev, int32_t *nsteps, double *action, double *ctrl_clock,
                               double sampling_time, int32_t max_steps, double *state_sys, double *accum,
                               int32_t *sample_flag, int32_t *nsamples, void *stream);

/* Simulator.sim_step, 'diff_eqn' branch (rcognita/simulator.py:161-168) = scipy
 * RK45.step() (scipy base.py:179-212, rk.py:111-176, rk.py:61-71): exactly ONE accepted
 * Dormand-Prince step per RUNNING lane, with per-lane adaptive step control, FSAL
 * derivative f carried across calls (stale w.r.t. action changes, like scipy), in-place
 * clipping of action.  Lanes that are not RUNNING are left untouched.  nfev[E] (may be
 * NULL) is incremented by 6 per attempt. */
int rcg_rk45_step(const rcg_system_t *sys, const rcg_solver_t *sol, int64_t E,
                  double *y, double *f, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                  double *action, void *stream);
int rcg_rk45_step_f32(const rcg_system_t *sys, const rcg_solver_t *sol, int64_t E,
                      float *y, float *f, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                      float *action, void *stream);

/* Fused main-loop body of presets/main_3wrobot_NI.py:415-440 between two controller
 * samples: each RUNNING lane repeats { sim_step; receive_sys_state; upd_accum_obj } with
 * the held action until its own sampling event `t - ctrl_clock >= sampling_time`
 * (rcognita/controllers.py:1440-1442), for at most max_steps accepted steps.  On a sampling
 * event the lane stops with sample_flag = 1, ctrl_clock = t and state_sys = the state BEFORE
 * the last step (the one-step lag of receive_sys_state); the accumulated
 * objective of that step is left to rcg_actor_cost's epilogue, which knows the new action.
 * On non-sampling steps accum += stage_obj(y, action) * sampling_time (controllers.py:1093).
 * nsteps[E] (may be NULL) counts accepted steps, nsamples[E] (may be NULL) sampling events. */
int rcg_rk45_advance(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj,
                     int64_t E, double *y, double *f, double *t, double *h_abs, int32_t *status,
                     int32_t *nfev, int32_t *nsteps, double *action, double *ctrl_clock,
                     double sampling_time, int32_t max_steps, double *state_sys, double *accum,
                     int32_t *sample_flag, int32_t *nsamples, void *stream);
/* rcg_rk45_advance that also appends one row per logged held-action step to the trajectory ring (the row of a
 * sampling step needs the controller's new action: the caller appends it with rcg_log_rows after the controller). */
int rcg_rk45_advance_logged(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj,
                            int64_t E, double *y, double *f, double *t, double *h_abs, int32_t *status,
                            int32_t *nfev, int32_t *nsteps, double *action, double *ctrl_clock,
                            double sampling_time, int32_t max_steps, double *state_sys, double *accum,
                            int32_t *sample_flag, int32_t *nsamples, const rcg_log_t *log, void *stream);
/* Appends the current (t, y, stage_obj(y, action), accum, action) of the lanes with mask != 0 (all if NULL) whose
 * nsteps[e] is a multiple of log->every (every lane if nsteps is NULL). */
int rcg_log_rows(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *t, const double *y,
                 const double *action, const double *accum, const int32_t *nsteps, const int32_t *mask,
                 const rcg_log_t *log, void *stream);
int rcg_rk45_advance_f32(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj,
                         int64_t E, float *y, float *f, double *t, double *h_abs, int32_t *status,
                         int32_t *nfev, int32_t *nsteps, float *action, double *ctrl_clock,
                         double sampling_time, int32_t max_steps, float *state_sys, float *accum,
                         int32_t *sample_flag, int32_t *nsamples, void *stream);

/* CtrlOptPred._actor_cost (rcognita/controllers.py:1273-1328) for E environments x C
 * candidate action sequences, one thread per (environment, candidate), plus np.argmin over
 * the candidates of each environment (first minimal index wins, NaN counts as minimal).
 *   state_sys[n][E], obs[n][E]  -- predictor start state and current observation.
 *   cand                        -- cand_per_env = 0: one shared table [Nactor*m][C]
 *                                  (component-major = transpose of the reference's
 *                                  action_sqn rows); cand_per_env = 1: [Nactor*m][E*C],
 *                                  element (k, j, e, c) at ((k*m + j) * E + e) * C + c.
 *   w_critic                    -- w_per_env = 0: [dim_critic]; 1: [dim_critic][E]; may be
 *                                  NULL in MPC mode.
 *   mask[E] or NULL             -- environments with mask == 0 are skipped entirely.
 *   J_out[E*C] or NULL          -- every cost (J_out[e*C + c]).
 *   argmin_out[E], Jmin_out[E]  -- or NULL.
 *   action_out[m][E] or NULL    -- first action of the arg-min sequence
 *                                  (_actor_optimizer's return value, controllers.py:1427).
 *   accum[E] or NULL            -- += stage_obj(obs, action_best) * sampling_time
 *                                  (upd_accum_obj for the sampling step, controllers.py:1093). */
int rcg_actor_cost(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                   const double *state_sys, const double *obs, const double *cand, int32_t cand_per_env,
                   const double *w_critic, int32_t w_per_env, const int32_t *mask,
                   double *J_out, int32_t *argmin_out, double *Jmin_out, double *action_out,
                   double *accum, double sampling_time, void *stream);
/* Name of the kernel variant the calling thread's last rcg_actor_cost(_f32) dispatched to: "actor_cost_tma_kernel"
 * (per-environment candidates staged by TMA, compile-time horizon), "actor_cost_tma_rt_kernel" (runtime horizon) or
 * "actor_cost_kernel" (direct loads: shared tables, dense R, unaligned shapes).  For benchmarks and tests. */
const char *rcg_last_actor_kernel(void);
int rcg_actor_cost_f32(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                       const float *state_sys, const float *obs, const float *cand, int32_t cand_per_env,
                       const float *w_critic, int32_t w_per_env, const int32_t *mask,
                       float *J_out, int32_t *argmin_out, float *Jmin_out, float *action_out,
                       float *accum, double sampling_time, void *stream);

/* ---- CtrlOptPred._actor_optimizer (rcognita/controllers.py:1330-1427) -------------------------------------
 * The reference minimises _actor_cost over the action sequence with scipy's SLSQP (finite-difference gradients,
 * Bounds(action_sqn_min, action_sqn_max), tol 1e-7, maxiter 300).  The batched form: one thread per
 * (environment, start point); exact gradients by the adjoint of the Euler rollout; a projected limited-memory
 * quasi-Newton iteration inside the box given by sys->lo/hi tiled over the horizon (controllers.py:968-971).
 *   sqn [Nactor*m][E*S]         -- start points in, minimisers out; element (k, j, e, s) at ((k*m+j)*E + e)*S + s
 *                                  (the layout of per-environment candidates).  S: power of two <= 32.
 *   workspace                   -- device scratch of rcg_actor_opt_workspace_bytes() bytes: a work-queue counter (the
 *                                  kernel is a persistent grid whose lanes pull the next problem when theirs is
 *                                  done), the final cost per start, and the quasi-Newton memory of the resident
 *                                  threads.  Its content is meaningless between calls; concurrent calls on
 *                                  different streams need separate workspaces.
 *   mask[E] or NULL             -- environments with mask == 0 are skipped entirely.
 *   max_iter, pg_tol, f_tol     -- stop after max_iter accepted iterations, when |P(x - g) - x|_inf <= pg_tol, or
 *                                  when the cost moved by <= f_tol * max(|J|, 1) twice in a row.
 *   J_out[E*S], iters_out[E*S], nfev_out[E*S] or NULL -- per start: final cost, accepted iterations (= gradient
 *                                  evaluations - 1), cost evaluations spent in the line searches.
 *   best_out[E], Jmin_out[E]    -- arg-min over the S starts of an environment (np.argmin order), or NULL.
 *   action_out[m][E] or NULL    -- first action of the best minimiser (_actor_optimizer's return value, :1427).
 *   accum[E] or NULL            -- += stage_obj(obs, action_best) * sampling_time (upd_accum_obj, :1093).
 * The iteration is monotone: the returned cost is never above the cost of the (clipped) start point. */
int64_t rcg_actor_opt_workspace_bytes(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S);
int rcg_actor_opt(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                  const double *obs, double *sqn, const double *w_critic, int32_t w_per_env, const int32_t *mask,
                  int32_t max_iter, double pg_tol, double f_tol, double *workspace, int64_t workspace_bytes,
                  double *J_out, int32_t *iters_out, int32_t *nfev_out, int32_t *best_out, double *Jmin_out,
                  double *action_out, double *accum, double sampling_time, void *stream);

/* Kernel variant behind rcg_actor_opt.  Default (0): Sys3WRobotNI / Sys3WRobot horizons 3..10 with diagonal R run `actor_opt_quad_kernel` -- four
 * lanes per problem, quasi-Newton pairs / iterate / rollout in shared memory, distributed inner products, the line search's
 * trial points evaluated side by side; everything else runs `actor_opt_kernel` (one lane per problem, state in registers and
 * the workspace).  lanes = 1 forces the one-lane kernel, 4 or 0 restore the default; other values are ignored.  Returns the
 * previous setting.  Both variants run the same iteration; results agree up to the summation order of inner products.
 * rcg_last_actor_opt_kernel names the variant the calling thread's last rcg_actor_opt dispatched to. */
int         rcg_actor_opt_lanes(int32_t lanes);
const char *rcg_last_actor_opt_kernel(void);

/* _actor_cost and its exact gradient w.r.t. the action sequence for E x S sequences (the adjoint sweep the
 * optimiser uses; what SLSQP approximates by forward differences): J_out[E*S], grad_out[Nactor*m][E*S].
 * `workspace` as for rcg_actor_opt (only read when the horizon/cost structure has no specialised kernel). */
int rcg_actor_grad(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                   const double *obs, const double *sqn, const double *w_critic, int32_t w_per_env, double *workspace,
                   int64_t workspace_bytes, double *J_out, double *grad_out, void *stream);

/* Gauss-Newton (iLQR) pre-pass of rcg_actor_opt for long horizons and stiff predictors (Sys3WRobot): every stage term of
 * _actor_cost (controllers.py:1273-1328) is quadratic in [observation - shift, action], so a reverse Riccati pass with the
 * linearised Euler predictor (plus, for Sys3WRobot, the second-order term of its heading/speed coupling) gives a Newton-like
 * step from O(n^2) state per problem; control limits by a clamped Newton
 * step per stage, Levenberg-Marquardt regularisation, backtracking on the cost.  At most max_sweeps sweeps per problem; a
 * start that passes the projected-gradient test (pg_tol) is left untouched; four failed forward passes in a row stop.
 * The cost never increases.  sqn, mask, w_critic as for rcg_actor_opt (sqn is clipped to the box); 'biquadratic' stage
 * costs are not quadratic: sqn is left untouched and sweeps_out = 0.  Call rcg_actor_opt afterwards: it
 * finishes from the returned point (on the reference's 72 recorded problems the slowest one needs 38 dependent iterations
 * instead of 300).  workspace: rcg_actor_ilqr_workspace_bytes() bytes.  sweeps_out[E*S] or NULL. */
int64_t rcg_actor_ilqr_workspace_bytes(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S);
int rcg_actor_ilqr(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                   const double *obs, double *sqn, const double *w_critic, int32_t w_per_env, const int32_t *mask,
                   int32_t max_sweeps, double pg_tol, double *workspace, int64_t workspace_bytes, int32_t *sweeps_out,
                   void *stream);

/* Start points from an arg-min: sqn_out[i][e] = candidate idx[e] of environment e (cand laid out as for
 * rcg_actor_cost; L = Nactor*m rows), for the lanes with mask != 0 and 0 <= idx[e] < C. */
int rcg_gather_sqn(int32_t L, int64_t E, int32_t C, const double *cand, int32_t cand_per_env, const int32_t *idx,
                   const int32_t *mask, double *sqn_out, void *stream);

/* CtrlNominal3WRobotNI (rcognita/controllers.py:1758-1956), the nominal parking controller of Sys3WRobotNI and the
 * default ctrl_mode of presets/main_3wrobot_NI.py: action[2][E] = clip(NH2ctrl_Cart(ctrl_gain * kappa(Cart2NH(obs))))
 * for the lanes with mask != 0 (all lanes if mask is NULL; the sampling-clock test of compute_action is
 * rcg_ctrl_sample / the sample_flag of rcg_rk45_advance).  If accum != NULL (obj required):
 * accum[E] += stage_obj(obs, action) * sampling_time (upd_accum_obj, :1086-1093). */
int rcg_nominal_ni(const rcg_system_t *sys, int64_t E, const double *obs, double ctrl_gain, const int32_t *mask,
                   double *action, const rcg_objective_t *obj, double *accum, double sampling_time, void *stream);

/* CtrlOptPred.stage_obj (rcognita/controllers.py:1063-1084): out[E] = stage_obj(obs, act);
 * if accum != NULL additionally accum[E] += out * scale (upd_accum_obj, :1086-1093). */
int rcg_stage_obj(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs,
                  const double *act, double *out, double *accum, double scale, void *stream);

/* CtrlOptPred._critic (rcognita/controllers.py:1192-1214): out[E] = w . phi(obs, act). */
int rcg_critic(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs,
               const double *act, const double *w, int32_t w_per_env, double *out, void *stream);

/* CtrlOptPred._critic_cost (rcognita/controllers.py:1216-1245) for E environments x W weight
 * vectors per environment.  obs_buf [buffer_size][n][E], act_buf [buffer_size][m][E] hold the
 * FIFO buffers (row 0 = oldest, push_vec of rcognita/utilities.py:78-79); w [dimc][E*W]
 * (element (i, e, k) at (i*E + e)*W + k), w_prev [dimc][E]; Jc_out [E*W]. */
int rcg_critic_cost(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, int32_t W,
                    const double *obs_buf, const double *act_buf, const double *w, const double *w_prev,
                    double *Jc_out, void *stream);

/* fp32 twin of rcg_critic_cost (the fp64-vs-fp32 tolerance report of BASELINE config 4). */
int rcg_critic_cost_f32(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, int32_t W,
                        const float *obs_buf, const float *act_buf, const float *w, const float *w_prev,
                        float *Jc_out, void *stream);

/* CtrlOptPred._critic_optimizer (rcognita/controllers.py:1248-1271) for E environments: the minimiser of
 * _critic_cost(w) subject to w_min <= w_i <= w_max (the reference's Bounds(Wmin, Wmax), controllers.py:1024-1039,
 * uniform per structure), started from w_init [dimc] (w_critic_init, shared by all environments; NULL = start
 * from the values found in w) and written to w [dimc][E].  The reference runs SLSQP; here a bounded linear
 * least-squares solve per environment (critic_fit.cu) -- the fitted cost is <= the cost at the start point
 * and, on the committed goldens, <= the reference's.  Lanes with mask == 0 are left untouched.
 * update_prev != 0 additionally stores the result in w_prev [dimc][E] (controllers.py:1471).
 * max_evals > 0 bounds the dual evaluations spent per environment (the best iterate so far is returned when the
 * budget runs out); max_evals <= 0: run to convergence -- for critics with >= 10 weights in two launches: a budgeted
 * pass with one environment per lane, then the unfinished environments again with one warp each (scratch from the
 * stream-ordered allocator of `stream`).  Jc_out[E] (may be NULL) receives
 * _critic_cost at the returned weights. */
int rcg_critic_fit(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs_buf,
                   const double *act_buf, double *w_prev, double w_min, double w_max, const double *w_init,
                   double *w, const int32_t *mask, int32_t max_evals, int32_t update_prev, double *Jc_out,
                   void *stream);

/* The sampling-clock test of CtrlOptPred.compute_action (rcognita/controllers.py:1440-1442;
 * the critic clock of :1459-1468 uses the same form): for every lane with in_mask != 0 (all
 * lanes if in_mask is NULL) mask_out = (t - clock >= period) and clock = t where it fired;
 * other lanes get mask_out = 0. */
int rcg_ctrl_sample(int64_t E, const double *t, double *clock, double period, const int32_t *in_mask,
                    int32_t *mask_out, void *stream);

/* utilities.push_vec on the controller FIFO buffers for the lanes with mask != 0
 * (rcognita/controllers.py:1463-1464): rows shift up by one, the new row goes to the bottom. */
int rcg_push_buffers(int32_t n, int32_t m, int32_t buffer_size, int64_t E, double *obs_buf, double *act_buf,
                     const double *obs, const double *act, const int32_t *mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RCG_H */
