// actor.cu -- CtrlOptPred._actor_cost (rcognita/controllers.py:1273-1328) for E environments
// x C candidate action sequences with a fused per-environment np.argmin.
//
// Mapping: one thread per (environment, candidate); the Euler rollout state, the running
// cost and the current/next action live in registers; the horizon runs in order inside the
// thread so the cost is accumulated in the reference's order (k = 0..Nactor-1).  Candidate
// components are read straight from HBM in component-major order (consecutive threads =
// consecutive candidates -> 256-byte coalesced requests), one stage ahead of their use.
// A block covers whole environments (EPB = 256 / C of them when C < 256, one when C >= 256,
// looping over C in chunks of 256), so the arg-min never leaves the block: per-thread best
// -> shared memory -> one warp per environment finishes with lexicographic (J, index)
// shuffles (first minimum wins, NaN counts as minimal, like np.argmin).
#include "rcg_host.h"

namespace rcg {

constexpr int kActorThreads = 256;
constexpr int kMaxEnvPerBlock = 32;

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ T ld_cached(const T *p) { return __ldg(p); }

template <typename T>
struct SmemW {
    const T *w;
    __device__ __forceinline__ T operator()(int i) const { return w[i]; }
};

// One _actor_cost evaluation.  `cp` points at component 0 of this thread's candidate, `ld` is
// the distance between consecutive components.
template <typename T, int SYS, int MODE, int CS, bool RDIAG, bool STREAM>
__device__ __forceinline__ T actor_cost_lane(const SysDev<T> &S, const ObjDev<T> &O, const T *x0, const T *ob0,
                                             const T *__restrict__ cp, int64_t ld, const T *w_s)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    const int NA = O.Nactor;
    T state[N], obs[N], a[M], an[M], d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { state[i] = x0[i]; obs[i] = ob0[i]; }    // controllers.py:1290-1291
#pragma unroll
    for (int j = 0; j < M; ++j) a[j] = STREAM ? ld_stream(cp + j * ld) : ld_cached(cp + j * ld);
    T J = T(0);
    const SmemW<T> w{w_s};
    for (int k = 0; k < NA; ++k) {
        const bool last = (k + 1 == NA);
        if (!last) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const T *p = cp + ((int64_t)(k + 1) * M + j) * ld;
                an[j] = STREAM ? ld_stream(p) : ld_cached(p);
            }
        }
        if constexpr (MODE == RCG_MODE_MPC) {
            J += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG>(O, obs, a);                 // :1305-1306
        } else if constexpr (MODE == RCG_MODE_RQL) {
            if (!last) J += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG>(O, obs, a);      // :1308-1309
            else J += critic<T, N, M, CS>(O, obs, a, w);                                // :1310
        } else {
            J += critic<T, N, M, CS>(O, obs, a, w);                                     // :1312-1326
        }
        if (!last) {
            state_dyn<T, SYS>(S, state, a, d);                                          // unclipped sys_rhs
#pragma unroll
            for (int i = 0; i < N; ++i) {
                state[i] = state[i] + O.pred_step_size * d[i];                          // Euler, :1294
                obs[i] = state[i];                                                      // sys_out = identity
            }
#pragma unroll
            for (int j = 0; j < M; ++j) a[j] = an[j];
        }
    }
    return J;
}

template <typename T, int SYS, int MODE, int CS, bool RDIAG, bool STREAM>
__global__ void __launch_bounds__(kActorThreads)
actor_cost_kernel(const __grid_constant__ SysDev<T> S, const __grid_constant__ ObjDev<T> O, int64_t E, int C,
                  int epb, int seg, int64_t num_groups, const T *__restrict__ state_sys_g, const T *__restrict__ obs_g,
                  const T *__restrict__ cand_g, int cand_per_env, const T *__restrict__ w_g, int w_per_env,
                  const int32_t *__restrict__ mask_g, T *__restrict__ J_g, int32_t *__restrict__ argmin_g,
                  T *__restrict__ Jmin_g, T *__restrict__ action_g, T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    constexpr int DIMC = dim_critic_c(CS, N, M);
    __shared__ T s_J[kActorThreads];
    __shared__ int s_I[kActorThreads];
    __shared__ T s_w[(MODE == RCG_MODE_MPC) ? 1 : kMaxEnvPerBlock * DIMC];

    const int tid = threadIdx.x;
    const int slot = tid / seg, cl = tid - slot * seg;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t ld = cand_per_env ? E * (int64_t)C : (int64_t)C;

    for (int64_t g = blockIdx.x; g < num_groups; g += gridDim.x) {
        const int64_t e = g * epb + slot;
        const bool active = slot < epb && e < E && (mask_g == nullptr || mask_g[e] != 0);

        T x0[N], ob[N];
        if (active) {
#pragma unroll
            for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob[i] = obs_g[i * E + e]; }
        }
        if constexpr (MODE != RCG_MODE_MPC) {
            if (active)
                for (int i = cl; i < DIMC; i += seg) s_w[slot * DIMC + i] = w_per_env ? w_g[i * E + e] : w_g[i];
            __syncthreads();
        }

        T bestJ = T(0);
        int bestI = 0x7fffffff;           // "no candidate": loses against any real index
        if (active) {
            const T *cbase = cand_g + (cand_per_env ? e * (int64_t)C : 0);
            for (int c = cl; c < C; c += seg) {
                const T J = actor_cost_lane<T, SYS, MODE, CS, RDIAG, STREAM>(S, O, x0, ob, cbase + c, ld,
                                                                             s_w + slot * DIMC);
                if (J_g) J_g[e * (int64_t)C + c] = J;
                if (bestI == 0x7fffffff || argmin_better(J, c, bestJ, bestI)) { bestJ = J; bestI = c; }
            }
        }
        s_J[tid] = bestJ;
        s_I[tid] = bestI;
        __syncthreads();

        // one warp per environment slot finishes the arg-min
        for (int sl = warp; sl < epb; sl += kActorThreads / 32) {
            const int64_t es = g * epb + sl;
            if (es >= E || (mask_g != nullptr && mask_g[es] == 0)) continue;      // warp-uniform
            T bJ = T(0);
            int bI = 0x7fffffff;
            for (int i = lane; i < seg; i += 32) {
                const T J = s_J[sl * seg + i];
                const int I = s_I[sl * seg + i];
                if (I != 0x7fffffff && (bI == 0x7fffffff || argmin_better(J, I, bJ, bI))) { bJ = J; bI = I; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const T oJ = __shfl_xor_sync(0xffffffffu, bJ, off);
                const int oI = __shfl_xor_sync(0xffffffffu, bI, off);
                if (oI != 0x7fffffff && (bI == 0x7fffffff || argmin_better(oJ, oI, bJ, bI))) { bJ = oJ; bI = oI; }
            }
            if (lane == 0 && bI != 0x7fffffff) {
                if (argmin_g) argmin_g[es] = bI;
                if (Jmin_g) Jmin_g[es] = bJ;
                if (action_g || accum_g) {
                    // first action of the best sequence (_actor_optimizer returns action_sqn[:dim_input])
                    T act[M], obs_e[N];
                    const T *cb = cand_g + (cand_per_env ? es * (int64_t)C : 0) + bI;
#pragma unroll
                    for (int j = 0; j < M; ++j) act[j] = cb[j * ld];
                    if (action_g) {
#pragma unroll
                        for (int j = 0; j < M; ++j) action_g[j * E + es] = act[j];
                    }
                    if (accum_g) {        // upd_accum_obj of the sampling step (controllers.py:1093)
#pragma unroll
                        for (int i = 0; i < N; ++i) obs_e[i] = obs_g[i * E + es];
                        accum_g[es] += stage_obj<T, N, M, RDIAG>(O, obs_e, act) * sampling_time;
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, int SYS, int MODE, int CS>
static void launch_actor_mc(bool rdiag, bool stream_ld, unsigned grid, cudaStream_t s, const SysDev<T> &S,
                            const ObjDev<T> &O, int64_t E, int C, int epb, int seg, int64_t groups, const T *state_sys,
                            const T *obs, const T *cand, int cand_per_env, const T *w, int w_per_env,
                            const int32_t *mask, T *J, int32_t *am, T *Jmin, T *action, T *accum, T st)
{
#define RCG_LAUNCH_ACTOR(RD, SL)                                                                                   \
    actor_cost_kernel<T, SYS, MODE, CS, RD, SL><<<grid, kActorThreads, 0, s>>>(                                     \
        S, O, E, C, epb, seg, groups, state_sys, obs, cand, cand_per_env, w, w_per_env, mask, J, am, Jmin, action, \
        accum, st)
    if (rdiag) { if (stream_ld) RCG_LAUNCH_ACTOR(true, true); else RCG_LAUNCH_ACTOR(true, false); }
    else       { if (stream_ld) RCG_LAUNCH_ACTOR(false, true); else RCG_LAUNCH_ACTOR(false, false); }
#undef RCG_LAUNCH_ACTOR
}

template <typename T, int SYS, typename... Args>
static int launch_actor_sys(int mode, int cs, Args... args)
{
#define RCG_CASE_CS(MODE)                                                                          \
    switch (cs) {                                                                                  \
    case RCG_CRITIC_QUAD_LIN:   launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_LIN>(args...); break;   \
    case RCG_CRITIC_QUADRATIC:  launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUADRATIC>(args...); break;  \
    case RCG_CRITIC_QUAD_NOMIX: launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_NOMIX>(args...); break; \
    case RCG_CRITIC_QUAD_MIX:   launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_MIX>(args...); break;   \
    default: return RCG_EINVAL;                                                                    \
    }
    if (mode == RCG_MODE_MPC) {
        launch_actor_mc<T, SYS, RCG_MODE_MPC, RCG_CRITIC_QUAD_NOMIX>(args...);
    } else if (mode == RCG_MODE_RQL) {
        RCG_CASE_CS(RCG_MODE_RQL)
    } else if (mode == RCG_MODE_SQL) {
        RCG_CASE_CS(RCG_MODE_SQL)
    } else {
        return RCG_EINVAL;
    }
#undef RCG_CASE_CS
    return 0;
}

template <typename T>
static int launch_actor(const char *what, const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                        const T *state_sys, const T *obs, const T *cand, int32_t cand_per_env, const T *w_critic,
                        int32_t w_per_env, const int32_t *mask, T *J_out, int32_t *argmin_out, T *Jmin_out,
                        T *action_out, T *accum, double sampling_time, void *stream)
{
    RCG_REQUIRE(sys && obj && state_sys && obs && cand, "%s: null argument", what);
    const int n = sys_n(sys->sys_id), m = sys_m(sys->sys_id);
    RCG_REQUIRE(n > 0, "%s: unknown sys_id %d", what, sys->sys_id);
    RCG_REQUIRE(obj->mode >= RCG_MODE_MPC && obj->mode <= RCG_MODE_SQL, "%s: unknown mode %d", what, obj->mode);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "%s: unknown critic_struct %d", what,
                obj->critic_struct);
    RCG_REQUIRE(obj->Nactor >= 1 && obj->Nactor <= RCG_MAX_NACTOR, "%s: Nactor %d out of range [1, %d]", what,
                obj->Nactor, RCG_MAX_NACTOR);
    RCG_REQUIRE(C >= 1, "%s: C must be >= 1", what);
    RCG_REQUIRE(obj->mode == RCG_MODE_MPC || w_critic, "%s: w_critic is required in RQL/SQL mode", what);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;

    const SysDev<T> S = make_sys_dev<T>(sys);
    const ObjDev<T> O = make_obj_dev<T>(obj, n, m);
    const bool rdiag = obj->r_is_diag && is_diag(obj->R1, n + m) &&
                       (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, n + m));
    int seg = C < kActorThreads ? C : kActorThreads;
    int epb = kActorThreads / seg;
    if (epb > kMaxEnvPerBlock) epb = kMaxEnvPerBlock;
    const int64_t groups = (E + epb - 1) / epb;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t max_grid = (int64_t)sms * 64;
    const unsigned grid = (unsigned)(groups < max_grid ? groups : max_grid);
    cudaStream_t s = (cudaStream_t)stream;
    const bool stream_ld = cand_per_env != 0;
    int rc;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI:
        rc = launch_actor_sys<T, RCG_SYS_3WROBOT_NI>(obj->mode, obj->critic_struct, rdiag, stream_ld, grid, s, S, O, E,
                                                      (int)C, epb, seg, groups, state_sys, obs, cand, (int)cand_per_env,
                                                      w_critic, (int)w_per_env, mask, J_out, argmin_out, Jmin_out,
                                                      action_out, accum, (T)sampling_time);
        break;
    case RCG_SYS_3WROBOT:
        rc = launch_actor_sys<T, RCG_SYS_3WROBOT>(obj->mode, obj->critic_struct, rdiag, stream_ld, grid, s, S, O, E,
                                                   (int)C, epb, seg, groups, state_sys, obs, cand, (int)cand_per_env,
                                                   w_critic, (int)w_per_env, mask, J_out, argmin_out, Jmin_out,
                                                   action_out, accum, (T)sampling_time);
        break;
    default:
        rc = launch_actor_sys<T, RCG_SYS_2TANK>(obj->mode, obj->critic_struct, rdiag, stream_ld, grid, s, S, O, E, (int)C,
                                                 epb, seg, groups, state_sys, obs, cand, (int)cand_per_env, w_critic,
                                                 (int)w_per_env, mask, J_out, argmin_out, Jmin_out, action_out, accum,
                                                 (T)sampling_time);
        break;
    }
    if (rc) { set_error("%s: bad mode/critic_struct", what); return rc; }
    return check_launch(what);
}

}  // namespace rcg

extern "C" {

int rcg_actor_cost(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C, const double *state_sys,
                   const double *obs, const double *cand, int32_t cand_per_env, const double *w_critic,
                   int32_t w_per_env, const int32_t *mask, double *J_out, int32_t *argmin_out, double *Jmin_out,
                   double *action_out, double *accum, double sampling_time, void *stream)
{
    return rcg::launch_actor<double>("rcg_actor_cost", sys, obj, E, C, state_sys, obs, cand, cand_per_env, w_critic,
                                     w_per_env, mask, J_out, argmin_out, Jmin_out, action_out, accum, sampling_time,
                                     stream);
}

int rcg_actor_cost_f32(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                       const float *state_sys, const float *obs, const float *cand, int32_t cand_per_env,
                       const float *w_critic, int32_t w_per_env, const int32_t *mask, float *J_out,
                       int32_t *argmin_out, float *Jmin_out, float *action_out, float *accum, double sampling_time,
                       void *stream)
{
    return rcg::launch_actor<float>("rcg_actor_cost_f32", sys, obj, E, C, state_sys, obs, cand, cand_per_env, w_critic,
                                    w_per_env, mask, J_out, argmin_out, Jmin_out, action_out, accum, sampling_time,
                                    stream);
}

}  // extern "C"
