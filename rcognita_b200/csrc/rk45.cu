// rk45.cu -- batched closed-loop integration: System.closed_loop_rhs and a per-lane
// restatement of scipy.integrate.RK45 as Simulator.sim_step drives it
// (rcognita/simulator.py:150, :161-168; scipy/integrate/_ivp/rk.py:14-71, :111-176,
// base.py:179-212, common.py:63-65).
//
// One thread per environment ("lane"); y, f and the seven stage derivatives live in
// registers; per-lane adaptive step control (accept/reject loop) exactly as scipy's,
// including the FSAL derivative that is NOT refreshed when the action changes between
// steps.  This translation unit is compiled with -fmad=false: the step
// size feeds the fp64 time accumulation that decides on which solver step the controller
// samples (section 3.3), so every multiply and add must round separately like numpy's.
#include <cstdlib>
#include <type_traits>

#include "rcg_host.h"

namespace rcg {

struct SolverDev {
    double t_bound, max_step, rtol, atol;
};

// scipy rk.py:538-553 (class RK45).  The coefficients sit in constant memory so that every DMUL takes its
// coefficient as a c[bank][offset] operand: as literals each use costs two UMOVs to materialise the 64-bit
// immediate (21 % of the issued instructions in the first profile of this kernel).
struct Tableau {
    double a10, a20, a21, a30, a31, a32, a40, a41, a42, a43, a50, a51, a52, a53, a54;
    double b0, b1, b2, b3, b4, b5;
    double e0, e1, e2, e3, e4, e5, e6;
};
static __constant__ Tableau kRK = {
    1.0 / 5, 3.0 / 40, 9.0 / 40, 44.0 / 45, -56.0 / 15, 32.0 / 9, 19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561,
    -212.0 / 729, 9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656,
    35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84,
    -71.0 / 57600, 0.0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
#define RK_A10 kRK.a10
#define RK_A20 kRK.a20
#define RK_A21 kRK.a21
#define RK_A30 kRK.a30
#define RK_A31 kRK.a31
#define RK_A32 kRK.a32
#define RK_A40 kRK.a40
#define RK_A41 kRK.a41
#define RK_A42 kRK.a42
#define RK_A43 kRK.a43
#define RK_A50 kRK.a50
#define RK_A51 kRK.a51
#define RK_A52 kRK.a52
#define RK_A53 kRK.a53
#define RK_A54 kRK.a54
#define RK_B0 kRK.b0
#define RK_B1 kRK.b1
#define RK_B2 kRK.b2
#define RK_B3 kRK.b3
#define RK_B4 kRK.b4
#define RK_B5 kRK.b5
#define RK_E0 kRK.e0
#define RK_E1 kRK.e1
#define RK_E2 kRK.e2
#define RK_E3 kRK.e3
#define RK_E4 kRK.e4
#define RK_E5 kRK.e5
#define RK_E6 kRK.e6

// One scipy RK45.step() for one lane.  Returns false if the step failed (TOO_SMALL_STEP).
// On success t, h_abs, y, f are advanced and *attempts holds the number of rk_step calls.
// Full-state dimension: the state, followed by the disturbance when is_disturb (systems.py:139-145).
template <int SYS, bool DIST> struct FullDim { static constexpr int n = SysDim<SYS>::n + (DIST ? DistDim<SYS>::nd : 0); };

// The right-hand side the solver integrates.  DIST = false: _state_dyn of the state.  DIST = true:
// System.closed_loop_rhs with is_disturb (systems.py:228-231, :247-248): state rows from _state_dyn(.., disturb), disturbance
// rows from _disturb_dyn with the two normal draws of RHS call number `call` of this environment.
template <typename T, int SYS, bool DIST>
__device__ __forceinline__ void rhs_eval(const SysDev<T> &S, const DistDev &D, unsigned long long env, unsigned int call,
                                         const T *y, const T *a, T *out)
{
    if constexpr (!DIST) {
        state_dyn<T, SYS>(S, y, a, out);
    } else {
        constexpr int N = SysDim<SYS>::n;
        state_dyn_disturbed<SYS>(S, y, a, y + N, out);
        double z[2] = {0.0, 0.0};
        if constexpr (SYS != RCG_SYS_2TANK) det_normal2(D.seed, env, call, z);
        disturb_dyn<SYS>(D, y + N, z, out + N);
    }
}

template <typename T, int SYS, bool DIST = false>
__device__ __forceinline__ bool rk45_one_step(const SysDev<T> &S, const SolverDev &sol, const T *a,
                                              double &t, double &h_abs, T *y, T *f, int &attempts,
                                              const DistDev &D = DistDev{}, unsigned long long env = 0, unsigned int call0 = 0)
{
    constexpr int N = FullDim<SYS, DIST>::n;
    const double t0 = t;
    const double min_step = 10 * fabs(nextafter(t0, (double)INFINITY) - t0);   // rk.py:118
    double ha;
    if (h_abs > sol.max_step) ha = sol.max_step;                               // rk.py:120-125
    else if (h_abs < min_step) ha = min_step;
    else ha = h_abs;

    bool rejected = false;
    T K[7][N], yn[N], yt[N];
    double t_new;
    attempts = 0;
    for (;;) {
        if (ha < min_step) return false;                                       // rk.py:131-132
        double h = ha;
        t_new = t0 + h;
        if (t_new - sol.t_bound > 0) t_new = sol.t_bound;                      // rk.py:137-138
        h = t_new - t0;
        ha = fabs(h);
        const T hT = (T)h;
        ++attempts;

        // rk_step (rk.py:61-71): dy = np.dot(K[:s].T, a[:s]) * h, sums left to right from 0.
#pragma unroll
        for (int i = 0; i < N; ++i) K[0][i] = f[i];                            // FSAL, never refreshed
#pragma unroll
        for (int i = 0; i < N; ++i) yt[i] = y[i] + (T(0) + K[0][i] * T(RK_A10)) * hT;
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 0u, yt, a, K[1]);
#pragma unroll
        for (int i = 0; i < N; ++i) yt[i] = y[i] + ((T(0) + K[0][i] * T(RK_A20)) + K[1][i] * T(RK_A21)) * hT;
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 1u, yt, a, K[2]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yt[i] = y[i] + (((T(0) + K[0][i] * T(RK_A30)) + K[1][i] * T(RK_A31)) + K[2][i] * T(RK_A32)) * hT;
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 2u, yt, a, K[3]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yt[i] = y[i] + ((((T(0) + K[0][i] * T(RK_A40)) + K[1][i] * T(RK_A41)) + K[2][i] * T(RK_A42)) +
                            K[3][i] * T(RK_A43)) * hT;
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 3u, yt, a, K[4]);
#pragma unroll
        for (int i = 0; i < N; ++i)
            yt[i] = y[i] + (((((T(0) + K[0][i] * T(RK_A50)) + K[1][i] * T(RK_A51)) + K[2][i] * T(RK_A52)) +
                             K[3][i] * T(RK_A53)) + K[4][i] * T(RK_A54)) * hT;
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 4u, yt, a, K[5]);
        // y_new = y + h * np.dot(K[:-1].T, B)
#pragma unroll
        for (int i = 0; i < N; ++i)
            yn[i] = y[i] + hT * ((((((T(0) + K[0][i] * T(RK_B0)) + K[1][i] * T(RK_B1)) + K[2][i] * T(RK_B2)) +
                                   K[3][i] * T(RK_B3)) + K[4][i] * T(RK_B4)) + K[5][i] * T(RK_B5));
        rhs_eval<T, SYS, DIST>(S, D, env, call0 + 6u * (unsigned)(attempts - 1) + 5u, yn, a, K[6]);

        // error norm: rk.py:105-109, :146-147; common.py:63-65
        T sq = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const T scale = (T)sol.atol + fmax(fabs(y[i]), fabs(yn[i])) * (T)sol.rtol;
            const T acc = ((((((T(0) + K[0][i] * T(RK_E0)) + K[1][i] * T(RK_E1)) + K[2][i] * T(RK_E2)) +
                             K[3][i] * T(RK_E3)) + K[4][i] * T(RK_E4)) + K[5][i] * T(RK_E5)) + K[6][i] * T(RK_E6);
            const T e = acc * hT / scale;
            sq += e * e;
        }
        const double err = (double)(sqrt(sq) / sqrt((T)N));

        if (err < 1) {                                                         // rk.py:149-160
            double factor;
            if (err == 0) factor = 10;
            else factor = fmin(10.0, 0.9 * det_pow_m02(err));
            if (rejected) factor = fmin(1.0, factor);
            ha *= factor;
            break;
        } else {                                                               // rk.py:161-164
            ha *= fmax(0.2, 0.9 * det_pow_m02(err));
            rejected = true;
        }
    }
    t = t_new;
    h_abs = ha;
#pragma unroll
    for (int i = 0; i < N; ++i) { y[i] = yn[i]; f[i] = K[6][i]; }
    return true;
}

// CTRL = false: Simulator.sim_step (one accepted step per running lane).
// CTRL = true : the fused loop body between two controller samples (see rcg_rk45_advance).
// Trajectory ring of the logged variants (rcg_log_t): rows [capacity][1 + n + 2 + m][E] =
// (t, state, stage_obj, accum_obj, action) per logged solver step and lane, written at slot count % capacity.
struct LogDev {
    double *rows;
    int32_t *count;
    int capacity, every;
};

template <typename T, int N, int M>
__device__ __forceinline__ void log_row(const LogDev &G, int64_t E, int64_t e, int &cnt, double t, const T *y, T stage,
                                        T accum, const T *a)
{
    constexpr int NC = 1 + N + 2 + M;
    double *r = G.rows + ((int64_t)(cnt % G.capacity) * NC) * E + e;
    r[0] = t;
#pragma unroll
    for (int i = 0; i < N; ++i) r[(int64_t)(1 + i) * E] = (double)y[i];
    r[(int64_t)(1 + N) * E] = (double)stage;
    r[(int64_t)(2 + N) * E] = (double)accum;
#pragma unroll
    for (int j = 0; j < M; ++j) r[(int64_t)(3 + N + j) * E] = (double)a[j];
    ++cnt;
}

// Resident 128-thread blocks per SM the register allocation is held to: the kernel is a chain of dependent FP64
// operations per lane, so residency (warps to switch between) is what buys pipe utilisation.  NI fits 5 blocks
// (96 registers) and 2tank 6 (74) without spilling; 3wrobot (35 state/stage doubles more) spills beyond 4.
__host__ __device__ constexpr int rk45_min_blocks(int sys) { return sys == RCG_SYS_3WROBOT_NI ? 5 : sys == RCG_SYS_2TANK ? 6 : 4; }

template <typename T, int SYS, bool CTRL, bool RDIAG, bool LOG = false, int BLOCK = 128, bool DIST = false>
__global__ void __launch_bounds__(BLOCK, DIST ? 3 : rk45_min_blocks(SYS) * (128 / BLOCK))
rk45_kernel(const __grid_constant__ SysDev<T> S, const __grid_constant__ SolverDev sol,
            const __grid_constant__ ObjDev<T> O, int64_t E, T *__restrict__ y_g, T *__restrict__ f_g,
            double *__restrict__ t_g, double *__restrict__ h_g, int32_t *__restrict__ status_g,
            int32_t *__restrict__ nfev_g, int32_t *__restrict__ nsteps_g, T *__restrict__ action_g,
            double *__restrict__ clock_g, double sampling_time, int max_steps, T *__restrict__ state_sys_g,
            T *__restrict__ accum_g, int32_t *__restrict__ flag_g, int32_t *__restrict__ nsamples_g,
            const __grid_constant__ LogDev G = LogDev{nullptr, nullptr, 0, 0},
            const __grid_constant__ DistDev D = DistDev{})
{
    // NS: rows of the state proper (stage_obj, state_sys, log); N: rows the solver integrates (+ the disturbance)
    constexpr int NS = SysDim<SYS>::n, N = FullDim<SYS, DIST>::n, M = SysDim<SYS>::m;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int st = status_g[e];
    if (st != RCG_RUNNING) {
        if (CTRL && flag_g) flag_g[e] = 0;
        return;
    }
    T y[N], f[N], a[M], yprev[N];
    const unsigned int call_base = DIST ? (unsigned int)nfev_g[e] : 0u;      // RHS calls made so far (1 at construction)
#pragma unroll
    for (int i = 0; i < N; ++i) { y[i] = y_g[i * E + e]; f[i] = f_g[i * E + e]; yprev[i] = y[i]; }
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = action_g[j * E + e];
    clip_action<T, M>(S, a);                       // systems.py:241-243, in place
    double t = t_g[e], h_abs = h_g[e];
    double clock = CTRL ? clock_g[e] : 0.0;
    T acc = (CTRL && accum_g) ? accum_g[e] : T(0);
    int nf = 0, ns = 0, flag = 0;
    int log_cnt = 0, step0 = 0;
    if constexpr (LOG) { log_cnt = G.count[e]; step0 = nsteps_g[e]; }

    for (int it = 0; it < max_steps && st == RCG_RUNNING; ++it) {
        if (t == sol.t_bound) { st = RCG_FINISHED; break; }                    // base.py:192-197
#pragma unroll
        for (int i = 0; i < N; ++i) yprev[i] = y[i];
        int attempts;
        const bool ok = rk45_one_step<T, SYS, DIST>(S, sol, a, t, h_abs, y, f, attempts, D,
                                                    (unsigned long long)(D.env_offset + e), call_base + (unsigned int)nf);
        nf += 6 * attempts;
        if (!ok) { st = RCG_FAILED; break; }                                   // base.py:203-204
        ++ns;
        if (t - sol.t_bound >= 0) st = RCG_FINISHED;                           // base.py:207-208
        if constexpr (CTRL) {
            if (t - clock >= sampling_time) {                                  // controllers.py:1440-1442
                clock = t;
                flag = 1;
                break;
            }
            // held action: compute_action returns action_curr (:1492-1493); upd_accum_obj (:1093)
            const T so = stage_obj<T, NS, M, RDIAG>(O, y, a);
            acc += so * (T)sampling_time;
            if constexpr (LOG) {                   // the row of a held-action step; sampling steps: rcg_log_rows
                if ((step0 + ns) % G.every == 0) log_row<T, NS, M>(G, E, e, log_cnt, t, y, so, acc, a);
            }
        }
    }
    if constexpr (LOG) G.count[e] = log_cnt;

#pragma unroll
    for (int i = 0; i < N; ++i) { y_g[i * E + e] = y[i]; f_g[i * E + e] = f[i]; }
#pragma unroll
    for (int j = 0; j < M; ++j) action_g[j * E + e] = a[j];
    t_g[e] = t;
    h_g[e] = h_abs;
    status_g[e] = st;
    if (nfev_g) nfev_g[e] += nf;
    if (nsteps_g) nsteps_g[e] += ns;
    if constexpr (CTRL) {
        clock_g[e] = clock;
        if (accum_g) accum_g[e] = acc;
        if (flag_g) flag_g[e] = flag;
        if (nsamples_g && flag) nsamples_g[e] += 1;
        if (state_sys_g && ns > 0) {
            // sampling lanes: the predictor starts from the state BEFORE the last step
            // (receive_sys_state runs after compute_action, main_3wrobot_NI.py:421-424);
            // other lanes: receive_sys_state(y) has already happened for this step.
#pragma unroll
            for (int i = 0; i < NS; ++i) state_sys_g[i * E + e] = flag ? yprev[i] : y[i];
        }
    }
}

// The row of a step on which the controller sampled (or of any step driven from the host): the action and the
// accumulated objective are only known after the controller ran.
template <int N, int M, bool RDIAG>
__global__ void __launch_bounds__(256)
log_rows_kernel(const __grid_constant__ ObjDev<double> O, const __grid_constant__ LogDev G, int64_t E,
                const double *__restrict__ t_g, const double *__restrict__ y_g, const double *__restrict__ action_g,
                const double *__restrict__ accum_g, const int32_t *__restrict__ nsteps_g, const int32_t *__restrict__ mask_g)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || (mask_g && mask_g[e] == 0)) return;
    if (nsteps_g && nsteps_g[e] % G.every != 0) return;
    double y[N], a[M];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = y_g[i * E + e];
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = action_g[j * E + e];
    int cnt = G.count[e];
    log_row<double, N, M>(G, E, e, cnt, t_g[e], y, stage_obj<double, N, M, RDIAG>(O, y, a), accum_g[e], a);
    G.count[e] = cnt;
}

template <typename T, int SYS, bool CLIP>
__global__ void __launch_bounds__(256)
rhs_kernel(const __grid_constant__ SysDev<T> S, int64_t E, const T *__restrict__ y_g, T *action_g, T *__restrict__ f_g)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    T y[N], a[M], d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = y_g[i * E + e];
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = action_g[j * E + e];
    if constexpr (CLIP) {
        clip_action<T, M>(S, a);
#pragma unroll
        for (int j = 0; j < M; ++j) action_g[j * E + e] = a[j];
    }
    state_dyn<T, SYS>(S, y, a, d);
#pragma unroll
    for (int i = 0; i < N; ++i) f_g[i * E + e] = d[i];
}

static DistDev make_dist_dev(const rcg_disturb_t *d)
{
    DistDev D{};
    for (int k = 0; k < 2; ++k) { D.sigma[k] = d->sigma[k]; D.mu[k] = d->mu[k]; D.tau[k] = d->tau[k]; }
    D.seed = d->seed;
    D.env_offset = d->env_offset;
    return D;
}

// System.closed_loop_rhs with is_disturb = 1 on the full state [n + nd][E]; draws numbered by call_g (0 if NULL).
// GIVEN != 0: the draws are taken from z_g [2][E] instead (parity with the reference under a patched randn()).
template <int SYS>
__global__ void __launch_bounds__(256)
rhs_disturbed_kernel(const __grid_constant__ SysDev<double> S, const __grid_constant__ DistDev D, int64_t E,
                     const double *__restrict__ y_g, double *action_g, const int32_t *__restrict__ call_g,
                     const double *__restrict__ z_g, double *__restrict__ f_g, int clip)
{
    constexpr int N = SysDim<SYS>::n, ND = DistDim<SYS>::nd, M = SysDim<SYS>::m;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double y[N + ND], a[M], d[N + ND];
#pragma unroll
    for (int i = 0; i < N + ND; ++i) y[i] = y_g[i * E + e];
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = action_g[j * E + e];
    if (clip) {
        clip_action<double, M>(S, a);
#pragma unroll
        for (int j = 0; j < M; ++j) action_g[j * E + e] = a[j];
    }
    if (z_g) {
        double z[2] = {z_g[e], ND > 1 ? z_g[E + e] : 0.0};
        state_dyn_disturbed<SYS>(S, y, a, y + N, d);
        disturb_dyn<SYS>(D, y + N, z, d + N);
    } else {
        rhs_eval<double, SYS, true>(S, D, (unsigned long long)(D.env_offset + e), call_g ? (unsigned int)call_g[e] : 0u, y, a, d);
    }
#pragma unroll
    for (int i = 0; i < N + ND; ++i) f_g[i * E + e] = d[i];
}

__global__ void __launch_bounds__(256)
normals_kernel(const __grid_constant__ DistDev D, int64_t E, const int32_t *__restrict__ call_g, int32_t call, double *__restrict__ z_g)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double z[2];
    det_normal2(D.seed, (unsigned long long)(D.env_offset + e), call_g ? (unsigned int)call_g[e] : (unsigned int)call, z);
    z_g[e] = z[0];
    z_g[E + e] = z[1];
}

template <typename T, bool CLIP>
static int launch_rhs(const rcg_system_t *sys, int64_t E, const T *y, T *action, T *f_out, void *stream)
{
    RCG_REQUIRE(sys && (E <= 0 || (y && action && f_out)), "rcg_rhs: null argument");
    RCG_REQUIRE(sys_n(sys->sys_id) > 0, "rcg_rhs: unknown sys_id %d", sys->sys_id);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const SysDev<T> S = make_sys_dev<T>(sys);
    const unsigned grid = (unsigned)((E + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI: rhs_kernel<T, RCG_SYS_3WROBOT_NI, CLIP><<<grid, 256, 0, s>>>(S, E, y, action, f_out); break;
    case RCG_SYS_3WROBOT:    rhs_kernel<T, RCG_SYS_3WROBOT, CLIP><<<grid, 256, 0, s>>>(S, E, y, action, f_out); break;
    default:                 rhs_kernel<T, RCG_SYS_2TANK, CLIP><<<grid, 256, 0, s>>>(S, E, y, action, f_out); break;
    }
    return check_launch("rcg_rhs");
}

template <typename T, int SYS, bool CTRL>
static void launch_rk45_sys(bool rdiag, unsigned grid, cudaStream_t s, const SysDev<T> &S, const SolverDev &sol,
                            const ObjDev<T> &O, int64_t E, T *y, T *f, double *t, double *h_abs, int32_t *status,
                            int32_t *nfev, int32_t *nsteps, T *action, double *clock, double sampling_time,
                            int max_steps, T *state_sys, T *accum, int32_t *flag, int32_t *nsamples,
                            const LogDev *log = nullptr, const DistDev *dist = nullptr)
{
    if constexpr (std::is_same<T, double>::value) {
        if (dist) {                                // disturbance lanes: full state [n + dim_disturb][E]
            const LogDev nolog{nullptr, nullptr, 0, 0};
            if (rdiag)
                rk45_kernel<T, SYS, CTRL, true, false, 128, true><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                                                      action, clock, sampling_time, max_steps, state_sys,
                                                                                      accum, flag, nsamples, nolog, *dist);
            else
                rk45_kernel<T, SYS, CTRL, false, false, 128, true><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                                                       action, clock, sampling_time, max_steps, state_sys,
                                                                                       accum, flag, nsamples, nolog, *dist);
            return;
        }
    }
    if constexpr (CTRL) {
        if (log) {
            if (rdiag)
                rk45_kernel<T, SYS, true, true, true><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                                           action, clock, sampling_time, max_steps, state_sys,
                                                                           accum, flag, nsamples, *log);
            else
                rk45_kernel<T, SYS, true, false, true><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                                            action, clock, sampling_time, max_steps, state_sys,
                                                                            accum, flag, nsamples, *log);
            return;
        }
    }
    if (rdiag)
        rk45_kernel<T, SYS, CTRL, true><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps, action,
                                                             clock, sampling_time, max_steps, state_sys, accum, flag,
                                                             nsamples);
    else
        rk45_kernel<T, SYS, CTRL, false><<<grid, 128, 0, s>>>(S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps, action,
                                                              clock, sampling_time, max_steps, state_sys, accum, flag,
                                                              nsamples);
}

template <typename T, bool CTRL>
static int launch_rk45(const char *what, const rcg_system_t *sys, const rcg_solver_t *sol_h, const rcg_objective_t *obj,
                       int64_t E, T *y, T *f, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                       int32_t *nsteps, T *action, double *clock, double sampling_time, int max_steps,
                       T *state_sys, T *accum, int32_t *flag, int32_t *nsamples, void *stream,
                       const rcg_log_t *log_h = nullptr, const rcg_disturb_t *dist_h = nullptr)
{
    RCG_REQUIRE(sys && sol_h && (E <= 0 || (y && f && t && h_abs && status && action)), "%s: null argument", what);
    DistDev distd{};
    if (dist_h) {
        RCG_REQUIRE(!log_h && (nfev || E <= 0), "%s: disturbance lanes need nfev (it numbers the random draws) and no log", what);
        distd = make_dist_dev(dist_h);
    }
    const DistDev *distp = dist_h ? &distd : nullptr;
    LogDev logd{nullptr, nullptr, 0, 0};
    if (log_h) {
        RCG_REQUIRE(CTRL && (E <= 0 || (log_h->rows && log_h->count && nsteps && accum)), "%s: log needs rows, count, nsteps and accum", what);
        RCG_REQUIRE(log_h->capacity >= 1 && log_h->every >= 1, "%s: log capacity and every must be >= 1", what);
        logd = LogDev{log_h->rows, log_h->count, log_h->capacity, log_h->every};
    }
    const LogDev *logp = log_h ? &logd : nullptr;
    const int n = sys_n(sys->sys_id), m = sys_m(sys->sys_id);
    RCG_REQUIRE(n > 0, "%s: unknown sys_id %d", what, sys->sys_id);
    RCG_REQUIRE(sol_h->max_step > 0, "%s: `max_step` must be positive.", what);    // scipy common.py:18-23
    if (CTRL) {
        RCG_REQUIRE(obj && (clock || E <= 0), "%s: objective and ctrl_clock are required", what);
        RCG_REQUIRE(max_steps > 0, "%s: max_steps must be positive", what);
    }
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const SysDev<T> S = make_sys_dev<T>(sys);
    const SolverDev sol{sol_h->t_bound, sol_h->max_step, sol_h->rtol, sol_h->atol};
    ObjDev<T> O;
    bool rdiag = true;
    if (CTRL) {
        O = make_obj_dev<T>(obj, n, m);
        rdiag = obj->r_is_diag && is_diag(obj->R1, n + m) &&
                (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, n + m));
    } else {
        memset(&O, 0, sizeof(O));
    }
    const unsigned grid = (unsigned)((E + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI:
        launch_rk45_sys<T, RCG_SYS_3WROBOT_NI, CTRL>(rdiag, grid, s, S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                      action, clock, sampling_time, max_steps, state_sys, accum, flag, nsamples, logp, distp);
        break;
    case RCG_SYS_3WROBOT:
        launch_rk45_sys<T, RCG_SYS_3WROBOT, CTRL>(rdiag, grid, s, S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                   action, clock, sampling_time, max_steps, state_sys, accum, flag, nsamples, logp, distp);
        break;
    default:
        launch_rk45_sys<T, RCG_SYS_2TANK, CTRL>(rdiag, grid, s, S, sol, O, E, y, f, t, h_abs, status, nfev, nsteps,
                                                 action, clock, sampling_time, max_steps, state_sys, accum, flag, nsamples, logp, distp);
        break;
    }
    return check_launch(what);
}

}  // namespace rcg

extern "C" {

int rcg_rhs(const rcg_system_t *sys, int64_t E, const double *y, double *action, double *f_out, void *stream)
{
    return rcg::launch_rhs<double, true>(sys, E, y, action, f_out, stream);
}

int rcg_rhs_f32(const rcg_system_t *sys, int64_t E, const float *y, float *action, float *f_out, void *stream)
{
    return rcg::launch_rhs<float, true>(sys, E, y, action, f_out, stream);
}

int rcg_state_dyn(const rcg_system_t *sys, int64_t E, const double *state, const double *action, double *dstate,
                  void *stream)
{
    return rcg::launch_rhs<double, false>(sys, E, state, const_cast<double *>(action), dstate, stream);
}

int rcg_rk45_step(const rcg_system_t *sys, const rcg_solver_t *sol, int64_t E, double *y, double *f, double *t,
                  double *h_abs, int32_t *status, int32_t *nfev, double *action, void *stream)
{
    return rcg::launch_rk45<double, false>("rcg_rk45_step", sys, sol, nullptr, E, y, f, t, h_abs, status, nfev, nullptr,
                                           action, nullptr, 0.0, 1, nullptr, nullptr, nullptr, nullptr, stream);
}

int rcg_rk45_step_f32(const rcg_system_t *sys, const rcg_solver_t *sol, int64_t E, float *y, float *f, double *t,
                      double *h_abs, int32_t *status, int32_t *nfev, float *action, void *stream)
{
    return rcg::launch_rk45<float, false>("rcg_rk45_step_f32", sys, sol, nullptr, E, y, f, t, h_abs, status, nfev,
                                          nullptr, action, nullptr, 0.0, 1, nullptr, nullptr, nullptr, nullptr, stream);
}

int rcg_rk45_advance(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj, int64_t E,
                     double *y, double *f, double *t, double *h_abs, int32_t *status, int32_t *nfev, int32_t *nsteps,
                     double *action, double *ctrl_clock, double sampling_time, int32_t max_steps, double *state_sys,
                     double *accum, int32_t *sample_flag, int32_t *nsamples, void *stream)
{
    return rcg::launch_rk45<double, true>("rcg_rk45_advance", sys, sol, obj, E, y, f, t, h_abs, status, nfev, nsteps,
                                          action, ctrl_clock, sampling_time, max_steps, state_sys, accum, sample_flag,
                                          nsamples, stream);
}

int rcg_log_rows(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *t, const double *y,
                 const double *action, const double *accum, const int32_t *nsteps, const int32_t *mask,
                 const rcg_log_t *log, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(obj && log && (E <= 0 || (t && y && action && accum && log->rows && log->count)), "rcg_log_rows: null argument");
    RCG_REQUIRE(log->capacity >= 1 && log->every >= 1, "rcg_log_rows: log capacity and every must be >= 1");
    RCG_REQUIRE((n == 3 && m == 2) || (n == 5 && m == 2) || (n == 2 && m == 1), "rcg_log_rows: unsupported dims n=%d m=%d", n, m);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const ObjDev<double> O = make_obj_dev<double>(obj, n, m);
    const LogDev G{log->rows, log->count, log->capacity, log->every};
    const bool rd = obj->r_is_diag && is_diag(obj->R1, n + m) && (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, n + m));
    const unsigned grid = (unsigned)((E + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
#define RCG_LOG_CALL(NN, MM)                                                                                       \
    if (rd) log_rows_kernel<NN, MM, true><<<grid, 256, 0, s>>>(O, G, E, t, y, action, accum, nsteps, mask);         \
    else log_rows_kernel<NN, MM, false><<<grid, 256, 0, s>>>(O, G, E, t, y, action, accum, nsteps, mask);
    if (n == 3) { RCG_LOG_CALL(3, 2) } else if (n == 5) { RCG_LOG_CALL(5, 2) } else { RCG_LOG_CALL(2, 1) }
#undef RCG_LOG_CALL
    return check_launch("rcg_log_rows");
}

int rcg_rk45_advance_logged(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj, int64_t E,
                            double *y, double *f, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                            int32_t *nsteps, double *action, double *ctrl_clock, double sampling_time, int32_t max_steps,
                            double *state_sys, double *accum, int32_t *sample_flag, int32_t *nsamples,
                            const rcg_log_t *log, void *stream)
{
    RCG_REQUIRE(log, "rcg_rk45_advance_logged: null log descriptor");
    return rcg::launch_rk45<double, true>("rcg_rk45_advance_logged", sys, sol, obj, E, y, f, t, h_abs, status, nfev, nsteps,
                                          action, ctrl_clock, sampling_time, max_steps, state_sys, accum, sample_flag,
                                          nsamples, stream, log);
}

int rcg_rhs_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, int64_t E, const double *y_full, double *action,
                      const int32_t *call, const double *normals, double *f_out, int32_t clip, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(sys && dist && (E <= 0 || (y_full && action && f_out)), "rcg_rhs_disturbed: null argument");
    RCG_REQUIRE(sys_n(sys->sys_id) > 0, "rcg_rhs_disturbed: unknown sys_id %d", sys->sys_id);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const SysDev<double> S = make_sys_dev<double>(sys);
    const DistDev D = make_dist_dev(dist);
    const unsigned grid = (unsigned)((E + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI: rhs_disturbed_kernel<RCG_SYS_3WROBOT_NI><<<grid, 256, 0, s>>>(S, D, E, y_full, action, call, normals, f_out, clip); break;
    case RCG_SYS_3WROBOT:    rhs_disturbed_kernel<RCG_SYS_3WROBOT><<<grid, 256, 0, s>>>(S, D, E, y_full, action, call, normals, f_out, clip); break;
    default:                 rhs_disturbed_kernel<RCG_SYS_2TANK><<<grid, 256, 0, s>>>(S, D, E, y_full, action, call, normals, f_out, clip); break;
    }
    return check_launch("rcg_rhs_disturbed");
}

int rcg_disturb_normals(const rcg_disturb_t *dist, int64_t E, const int32_t *call, int32_t call_all, double *normals, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(dist && (E <= 0 || normals), "rcg_disturb_normals: null argument");
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    normals_kernel<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_dist_dev(dist), E, call, call_all, normals);
    return check_launch("rcg_disturb_normals");
}

int rcg_rk45_step_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, const rcg_solver_t *sol, int64_t E,
                            double *y_full, double *f_full, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                            double *action, void *stream)
{
    RCG_REQUIRE(dist, "rcg_rk45_step_disturbed: null disturbance descriptor");
    return rcg::launch_rk45<double, false>("rcg_rk45_step_disturbed", sys, sol, nullptr, E, y_full, f_full, t, h_abs, status, nfev,
                                           nullptr, action, nullptr, 0.0, 1, nullptr, nullptr, nullptr, nullptr, stream, nullptr, dist);
}

int rcg_rk45_advance_disturbed(const rcg_system_t *sys, const rcg_disturb_t *dist, const rcg_solver_t *sol,
                               const rcg_objective_t *obj, int64_t E, double *y_full, double *f_full, double *t, double *h_abs,
                               int32_t *status, int32_t *nfev, int32_t *nsteps, double *action, double *ctrl_clock,
                               double sampling_time, int32_t max_steps, double *state_sys, double *accum,
                               int32_t *sample_flag, int32_t *nsamples, void *stream)
{
    RCG_REQUIRE(dist, "rcg_rk45_advance_disturbed: null disturbance descriptor");
    return rcg::launch_rk45<double, true>("rcg_rk45_advance_disturbed", sys, sol, obj, E, y_full, f_full, t, h_abs, status, nfev,
                                          nsteps, action, ctrl_clock, sampling_time, max_steps, state_sys, accum, sample_flag,
                                          nsamples, stream, nullptr, dist);
}

int rcg_rk45_advance_f32(const rcg_system_t *sys, const rcg_solver_t *sol, const rcg_objective_t *obj, int64_t E,
                         float *y, float *f, double *t, double *h_abs, int32_t *status, int32_t *nfev,
                         int32_t *nsteps, float *action, double *ctrl_clock, double sampling_time, int32_t max_steps,
                         float *state_sys, float *accum, int32_t *sample_flag, int32_t *nsamples, void *stream)
{
    return rcg::launch_rk45<float, true>("rcg_rk45_advance_f32", sys, sol, obj, E, y, f, t, h_abs, status, nfev, nsteps,
                                         action, ctrl_clock, sampling_time, max_steps, state_sys, accum, sample_flag,
                                         nsamples, stream);
}

}  // extern "C"
