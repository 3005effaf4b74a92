// actor_impl.cuh -- CtrlOptPred._actor_cost (rcognita/controllers.py:1273-1328) for E environments
// x C candidate action sequences with a fused per-environment np.argmin.
//
// Mapping: one WARP per environment (or, for C < 32, one power-of-two lane segment per
// environment).  Each lane evaluates candidates lane, lane+32, ... in order: the Euler rollout
// state, the running cost and the whole action sequence of the candidate live in registers,
// and the horizon runs in order inside the thread, so the cost is accumulated in the
// reference's order (k = 0..Nactor-1).  Candidate components are read straight from HBM in
// component-major order (consecutive lanes = consecutive candidates -> 256-byte coalesced
// requests); with a compile-time horizon (NA > 0) all Nactor*m loads of a candidate are issued
// before the first use.  The arg-min never leaves the warp: per-lane running best, then
// lexicographic (J, index) xor-shuffles (first minimum wins, NaN counts as minimal, like
// np.argmin) -- no shared memory, no block barrier.  Environment-uniform data (state_sys,
// observation, critic weights, sin/cos of the initial heading) are loaded once per environment.
//
// Heading trigonometry of the two robots: the predictor needs sin/cos of theta_k for every
// stage.  theta_{k+1} = theta_k + delta with delta = h * omega_k, so (sin, cos) are advanced by
// the angle-addition formulas with a reduction-free polynomial sincos(delta) (|delta| <= pi/4,
// the usual case: h = 0.01..0.02 s); otherwise a full sincos(theta_{k+1}).  Each rotation adds
// <= 2 ulp, i.e. ~1e-15 over the longest horizon -- against the 1e-9 cost tolerance.
#pragma once

#include <cuda.h>
#include <math_constants.h>

#include "rcg_host.h"

namespace rcg {

constexpr int kActorThreads = 256;
constexpr int kActorWarps = kActorThreads / 32;

// sin/cos on [-pi/4, pi/4] without range reduction (fdlibm __kernel_sin / __kernel_cos minimax
// coefficients, |error| < 2^-57).  The coefficients sit in constant memory so that every DFMA
// takes its coefficient as a c[bank][offset] operand (no immediate-materialising moves).
static __constant__ double kSinC[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
static __constant__ double kCosC[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                                2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

__device__ __forceinline__ void sincos_small(double x, double *s, double *c)
{
    const double z = x * x;
    double ps = kSinC[0], pc = kCosC[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) {
        ps = fma(ps, z, kSinC[i]);
        pc = fma(pc, z, kCosC[i]);
    }
    *s = fma(x * z, ps, x);
    *c = fma(z * z, pc, fma(z, -0.5, 1.0));
}
// |x| <= 1/8 (the predictor's heading increment h * omega: 0.05 rad at the presets' bounds): the two highest-order
// terms of each polynomial are below 2^-70 relative and are dropped -- 4 of the 16 FP64 instructions of a rotation.
__device__ __forceinline__ void sincos_tiny(double x, double *s, double *c)
{
    const double z = x * x;
    double ps = kSinC[2], pc = kCosC[2];
#pragma unroll
    for (int i = 3; i < 6; ++i) {
        ps = fma(ps, z, kSinC[i]);
        pc = fma(pc, z, kCosC[i]);
    }
    *s = fma(x * z, ps, x);
    *c = fma(z * z, pc, fma(z, -0.5, 1.0));
}

static __device__ __noinline__ double2 sincos_full(double x)
{
    double2 r;
    sincos(x, &r.x, &r.y);
    return r;
}

// (s, c) = (sin, cos)(theta_old) on entry; (sin, cos)(theta_new) on exit, theta_new = theta_old + delta.
// TINY: try the truncated polynomials first (measured: 0.309 -> 0.304 ms on the NI headline launch; for 3wrobot the
// extra path costs registers under its 128-register cap and was slower, 2.64 -> 2.88 ms, so it is NI-only).
template <bool TINY = false>
__device__ __forceinline__ void rotate_trig(double theta_new, double delta, double &s, double &c)
{
    if (fabs(delta) <= 0.78539816339744830962) {
        double sd, cd;
        if (TINY && fabs(delta) <= 0.125) sincos_tiny(delta, &sd, &cd);
        else sincos_small(delta, &sd, &cd);
        const double cn = fma(c, cd, -(s * sd));
        const double sn = fma(s, cd, c * sd);
        s = sn;
        c = cn;
    } else {
        const double2 r = sincos_full(theta_new);   // also the NaN / inf path
        s = r.x;
        c = r.y;
    }
}
// fp32 twin: same rotation with single-precision minimax kernels (Cephes sinf/cosf coefficients,
// < 1 ulp on [-pi/4, pi/4]); the accumulated rotation error (~1e-7 per stage) is inside the fp32
// tolerance of the path (tests: 2e-5 relative on costs).
template <bool TINY = false>
__device__ __forceinline__ void rotate_trig(float theta_new, float delta, float &s, float &c)
{
    if (fabsf(delta) <= 0.78539816f) {
        const float z = delta * delta;
        const float sd = fmaf(delta * z, fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), delta);
        const float cd = fmaf(z * z, fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                              fmaf(z, -0.5f, 1.0f));
        const float cn = fmaf(c, cd, -(s * sd));
        const float sn = fmaf(s, cd, c * sd);
        s = sn;
        c = cn;
    } else {
        sincosf(theta_new, &s, &c);
    }
}

// One explicit-Euler predictor step, state += h * _state_dyn(state, a)  (controllers.py:1294 with
// sys_rhs = System._state_dyn, unclipped).  (s, c) caches sin/cos of the heading state[2].
template <typename T, int SYS>
__device__ __forceinline__ void euler_step(const SysDev<T> &S, T h, T *x, const T *a, T &s, T &c)
{
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {           // systems.py:370-382
        const T d = h * a[1];
        x[0] = x[0] + h * (a[0] * c);
        x[1] = x[1] + h * (a[0] * s);
        x[2] = x[2] + d;
        rotate_trig<true>(x[2], d, s, c);
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {       // systems.py:308-323
        const T d = h * x[4];
        x[0] = x[0] + h * (x[3] * c);
        x[1] = x[1] + h * (x[3] * s);
        x[2] = x[2] + d;
        x[3] = x[3] + h * ((T(1) / S.pars[0]) * a[0]);
        x[4] = x[4] + h * ((T(1) / S.pars[1]) * a[1]);
        rotate_trig<false>(x[2], d, s, c);
    } else {                                             // systems.py:412-419
        T d[2];
        state_dyn<T, SYS>(S, x, a, d);
        x[0] = x[0] + h * d[0];
        x[1] = x[1] + h * d[1];
    }
}

template <typename T, int DIMC>
struct RegW {
    const T *w;
    __device__ __forceinline__ T operator()(int i) const { return w[i]; }
};

// Template flag RDIAG of the actor kernels = "R1 diagonal AND stage_obj_struct == 'quadratic'" (every
// preset); the general kernels (RDIAG = false) take dense R1/R2 and branch on the structure.
// ActorEval holds the running state of one _actor_cost evaluation (rollout state, cached heading
// trigonometry, accumulated cost); stage(k) adds the cost of stage k and advances the predictor.
// Non-finite test on the integer pipe (the FP64 pipe is the busy one): exponent bits all ones.
__device__ __forceinline__ int nonfinite_bits(double v) { return ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) ? 1 : 0; }
__device__ __forceinline__ int nonfinite_bits(float v) { return ((__float_as_int(v) & 0x7f800000) == 0x7f800000) ? 1 : 0; }

// LEAN = the presets' objective shape, decided on the host: diagonal R1 whose ACTION entries are zero, no
// observation_target, gamma = 1.  Then  gamma**k * ((chi - 0) R chi)  reduces to the n observation terms: the target
// subtraction, the two zero-weight action terms and the multiplication by 1 are exact no-ops for finite inputs and
// are dropped -- 9 of the ~40 FP64 instructions of a 3wrobot_NI stage (the TMA kernel is co-limited by FP64 issue).
// The one thing 0 * a * a does for the reference is to turn a non-finite action into a NaN cost; the lean kernels
// keep that with an integer-pipe test on the loaded actions (`bad`), so results are identical bit for bit.
template <typename T, int SYS, int MODE, int CS, bool RDIAG, bool LEAN = false>
struct ActorEval {
    static constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    static constexpr int DIMC = dim_critic_c(CS, N, M);
    const SysDev<T> &S;
    const ObjDev<T> &O;
    const RegW<T, DIMC> w;
    T state[N], obs[N];
    T s, c, J, h;

    __device__ __forceinline__ ActorEval(const SysDev<T> &S_, const ObjDev<T> &O_, const T *x0, const T *ob0, T s0, T c0,
                                         const T *w_r)
        : S(S_), O(O_), w{w_r}, s(s0), c(c0), J(T(0)), h(O_.pred_step_size)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) { state[i] = x0[i]; obs[i] = ob0[i]; }    // controllers.py:1290-1291
    }

    __device__ __forceinline__ T stage_obj_lean(const T *ob) const
    {
        constexpr int P = N + M;
        T out = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) out += (ob[i] * O.R1[i * P + i]) * ob[i];
        return out;
    }

    __device__ __forceinline__ void stage(int k, bool last, const T *a)
    {
        if constexpr (MODE == RCG_MODE_MPC) {
            if constexpr (LEAN) J += stage_obj_lean(obs);
            else J += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, obs, a);     // :1305-1306
        } else if constexpr (MODE == RCG_MODE_RQL) {
            if (!last) {
                if constexpr (LEAN) J += stage_obj_lean(obs);
                else J += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, obs, a); // :1308-1309
            } else {
                J += critic<T, N, M, CS>(O, obs, a, w);                                 // :1310
            }
        } else {
            J += critic<T, N, M, CS>(O, obs, a, w);                                     // :1312-1326
        }
        if (!last) {
            euler_step<T, SYS>(S, h, state, a, s, c);                                   // :1292-1296
#pragma unroll
            for (int i = 0; i < N; ++i) obs[i] = state[i];                              // sys_out = identity
        }
    }
};

// One _actor_cost evaluation.  `cp` points at component 0 of this lane's candidate, `ld` is the
// distance between consecutive components.  NA > 0: compile-time horizon, fully unrolled.
template <typename T, int SYS, int MODE, int CS, bool RDIAG, int NA, bool LEAN = false>
__device__ __forceinline__ T actor_cost_lane(const SysDev<T> &S, const ObjDev<T> &O, const T *x0, const T *ob0, T s0,
                                             T c0, const T *__restrict__ cp, int64_t ld, const T *w_r)
{
    constexpr int M = SysDim<SYS>::m;
    ActorEval<T, SYS, MODE, CS, RDIAG, LEAN> ev(S, O, x0, ob0, s0, c0, w_r);
    if constexpr (NA > 0) {
        T a[NA][M];
        int bad = 0;
#pragma unroll
        for (int k = 0; k < NA; ++k)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                a[k][j] = __ldg(cp + (int64_t)(k * M + j) * ld);
                if constexpr (LEAN) bad |= nonfinite_bits(a[k][j]);
            }
#pragma unroll
        for (int k = 0; k < NA; ++k) ev.stage(k, k + 1 == NA, a[k]);
        if constexpr (LEAN) {
            if (bad) ev.J = (T)CUDART_NAN;                               // 0 * non-finite action = NaN in the reference
        }
    } else {
        // runtime horizon: chunks of CH stages, the next chunk's actions are in flight while the
        // current chunk is evaluated
        constexpr int CH = 4;
        const int na = O.Nactor;
        T a[CH][M], an[CH][M];
        int bad = 0;
#pragma unroll
        for (int i = 0; i < CH; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) a[i][j] = (i < na) ? __ldg(cp + (int64_t)(i * M + j) * ld) : T(0);
        for (int k0 = 0; k0 < na; k0 += CH) {
#pragma unroll
            for (int i = 0; i < CH; ++i)
#pragma unroll
                for (int j = 0; j < M; ++j)
                    an[i][j] = (k0 + CH + i < na) ? __ldg(cp + ((int64_t)(k0 + CH + i) * M + j) * ld) : T(0);
#pragma unroll
            for (int i = 0; i < CH; ++i)
                if (k0 + i < na) {
                    if constexpr (LEAN) bad |= nonfinite_bits(a[i][0]) | nonfinite_bits(a[i][M - 1]);
                    ev.stage(k0 + i, k0 + i + 1 == na, a[i]);
                }
#pragma unroll
            for (int i = 0; i < CH; ++i)
#pragma unroll
                for (int j = 0; j < M; ++j) a[i][j] = an[i][j];
        }
        if constexpr (LEAN) {
            if (bad) ev.J = (T)CUDART_NAN;
        }
    }
    return ev.J;
}

struct ActorArgs {
    int64_t E;
    int C, seg, seg_shift;          // lanes per environment (32, or the power of two >= C), log2(seg)
    int64_t num_groups;             // warps' worth of environments: ceil(E / (32 / seg))
    int cand_per_env, w_per_env;
};

// Resident 256-thread blocks per SM the register allocation of the specialised kernels is held to (measured on
// B200, shared candidate table = the FP64-bound case): 3wrobot at 2 blocks (128 registers, <= 96 B of spills)
// instead of the 1 block its 162 registers allowed: 3.10 -> 2.64 ms for 262,144 x 256 RQL 'quadratic' N=10;
// NI at 4 blocks (64 registers, spills) was slower than at 3 (0.298 vs 0.290 ms), so it stays at 3; 2tank fits 4.
__host__ __device__ constexpr int actor_min_blocks(int sys, int na, bool rdiag)
{
    return (na > 0 && rdiag) ? (sys == RCG_SYS_3WROBOT_NI ? 3 : sys == RCG_SYS_3WROBOT ? 2 : 4) : 1;
}

template <typename T, int SYS, int MODE, int CS, bool RDIAG, int NA, bool LEAN = false>
__global__ void __launch_bounds__(kActorThreads, actor_min_blocks(SYS, NA, RDIAG))
actor_cost_kernel(const __grid_constant__ SysDev<T> S, const __grid_constant__ ObjDev<T> O,
                  const __grid_constant__ ActorArgs A, const T *__restrict__ state_sys_g, const T *__restrict__ obs_g,
                  const T *__restrict__ cand_g, const T *__restrict__ w_g, const int32_t *__restrict__ mask_g,
                  T *__restrict__ J_g, int32_t *__restrict__ argmin_g, T *__restrict__ Jmin_g, T *__restrict__ action_g,
                  T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    constexpr int kNone = 0x7fffffff;            // "no candidate": loses against any real index
    const int64_t E = A.E;
    const int C = A.C, seg = A.seg;
    const int lane = threadIdx.x & 31;
    const int slot = lane >> A.seg_shift, cl = lane & (seg - 1);
    const int epw = 32 >> A.seg_shift;           // environments per warp
    const int64_t ld = A.cand_per_env ? E * (int64_t)C : (int64_t)C;
    const int64_t warp0 = (int64_t)blockIdx.x * kActorWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kActorWarps;

    for (int64_t g = warp0; g < A.num_groups; g += nwarps) {
        const int64_t e = g * epw + slot;
        const bool active = e < E && (mask_g == nullptr || mask_g[e] != 0);
        T bestJ = T(0);
        int bestI = kNone;
        if (active) {
            T x0[N], ob[N], w[DIMC];
#pragma unroll
            for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob[i] = obs_g[i * E + e]; }
            if constexpr (MODE != RCG_MODE_MPC) {
#pragma unroll
                for (int i = 0; i < DIMC; ++i) w[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
            }
            T s0 = T(0), c0 = T(1);
            if constexpr (SYS != RCG_SYS_2TANK) sincos_t(x0[2], &s0, &c0);
            const T *cbase = cand_g + (A.cand_per_env ? e * (int64_t)C : 0);
            for (int c = cl; c < C; c += seg) {
                const T J = actor_cost_lane<T, SYS, MODE, CS, RDIAG, NA, LEAN>(S, O, x0, ob, s0, c0, cbase + c, ld, w);
                if (J_g) J_g[e * (int64_t)C + c] = J;
                if (bestI == kNone || argmin_better(J, c, bestJ, bestI)) { bestJ = J; bestI = c; }
            }
        }
        // arg-min across the lanes of this environment's segment (all 32 lanes take part in the shuffles)
        for (int off = seg >> 1; off > 0; off >>= 1) {
            const T oJ = __shfl_xor_sync(0xffffffffu, bestJ, off);
            const int oI = __shfl_xor_sync(0xffffffffu, bestI, off);
            if (oI != kNone && (bestI == kNone || argmin_better(oJ, oI, bestJ, bestI))) { bestJ = oJ; bestI = oI; }
        }
        if (active && cl == 0 && bestI != kNone) {
            if (argmin_g) argmin_g[e] = bestI;
            if (Jmin_g) Jmin_g[e] = bestJ;
            if (action_g || accum_g) {
                // first action of the best sequence (_actor_optimizer returns action_sqn[:dim_input])
                T act[M], obs_e[N];
                const T *cb = cand_g + (A.cand_per_env ? e * (int64_t)C : 0) + bestI;
#pragma unroll
                for (int j = 0; j < M; ++j) act[j] = cb[j * ld];
                if (action_g) {
#pragma unroll
                    for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                }
                if (accum_g) {            // upd_accum_obj of the sampling step (controllers.py:1093)
#pragma unroll
                    for (int i = 0; i < N; ++i) obs_e[i] = obs_g[i * E + e];
                    accum_g[e] += stage_obj<T, N, M, RDIAG, RDIAG>(O, obs_e, act) * sampling_time;
                }
            }
        }
    }
}


// ---- TMA-staged variant for per-environment candidates (the HBM-bound configuration) ---------------
// Same mapping and arithmetic as actor_cost_kernel, but the candidate stream never touches the
// load/store pipe of the lanes: the per-environment candidate array [Nactor*m][E*C] is described by a
// 2-D tensor map and one elected lane per warp issues ONE cp.async.bulk.tensor (TMA) per 32
// candidates -- a box of 32 consecutive (environment, candidate) columns x Nactor*m component rows --
// into the warp's private shared-memory ring, completion signalled on an mbarrier.  The ring keeps
// STAGES boxes in flight per warp while the lanes run the FP64-heavy rollouts of the current box, so
// HBM stays busy regardless of where the warps are in their arithmetic; the lanes read their candidate
// as conflict-free LDS.64 ([component][lane] layout = the box as TMA writes it) and spend no
// instructions on 64-bit address arithmetic or global loads.  The pipeline runs across environment
// boundaries (items = this warp's (environment group, 32-candidate box) pairs in order).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_box(void *dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// Ring depth and residency per horizon length.  A lane copies its candidate (its column of the box)
// into registers as soon as the box has landed and the slot is refilled at once, so STAGES boxes per
// warp are in flight while the rollouts run.  Measured on B200 (profiles/): three resident CTAs with a
// 3-deep ring (80 registers) beat four CTAs with a 2-deep ring that read the box lazily (64 registers,
// more moves, exposed LDS latency) -- 0.311 ms vs 0.346 ms on the 65,536 x 256 headline launch.
template <typename T, int L>
__host__ __device__ constexpr int tma_stages() { return (L * (int)sizeof(T) <= 64) ? 4 : (L * (int)sizeof(T) <= 96) ? 3 : 2; }
template <typename T, int L>
__host__ __device__ constexpr int tma_smem_bytes() { return kActorWarps * tma_stages<T, L>() * (L * 32 * (int)sizeof(T) + 8); }
template <typename T, int L>
__host__ __device__ constexpr int tma_min_ctas() { return (225 * 1024 / (tma_smem_bytes<T, L>() + 1024)) >= 3 ? 3 : 2; }

template <typename T, int SYS, int MODE, int CS, int NA, bool LEAN = false>
__global__ void __launch_bounds__(kActorThreads, tma_min_ctas<T, NA * SysDim<SYS>::m>())
actor_cost_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SysDev<T> S,
                      const __grid_constant__ ObjDev<T> O, const __grid_constant__ ActorArgs A,
                      const T *__restrict__ state_sys_g, const T *__restrict__ obs_g, const T *__restrict__ cand_g,
                      const T *__restrict__ w_g, const int32_t *__restrict__ mask_g, T *__restrict__ J_g,
                      int32_t *__restrict__ argmin_g, T *__restrict__ Jmin_g, T *__restrict__ action_g,
                      T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m, L = NA * M;
    constexpr int STAGES = tma_stages<T, L>();
    constexpr int BOX = L * 32;                                         // elements per box
    constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    constexpr int kNone = 0x7fffffff;
    extern __shared__ __align__(128) unsigned char actor_smem[];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ring = reinterpret_cast<T *>(actor_smem) + (size_t)wi * STAGES * BOX;
    uint64_t *bars = reinterpret_cast<uint64_t *>(actor_smem + (size_t)kActorWarps * STAGES * BOX * sizeof(T)) + wi * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int64_t E = A.E;
    const int C = A.C, seg = A.seg;
    const int slot = lane >> A.seg_shift, cl = lane & (seg - 1);
    const int epw = 32 >> A.seg_shift;                                  // environments per warp (1 when C >= 32)
    const int cpl = (C + seg - 1) >> A.seg_shift;                       // 32-candidate boxes per environment group
    const int64_t ld = E * (int64_t)C;
    const int64_t warp0 = (int64_t)blockIdx.x * kActorWarps + wi;
    const int64_t nwarps = (int64_t)gridDim.x * kActorWarps;
    const int64_t ngroups = (A.num_groups > warp0) ? (A.num_groups - warp0 + nwarps - 1) / nwarps : 0;

    auto env_active = [&](int64_t gi, int sl) -> int {
        if (gi >= ngroups) return 0;
        const int64_t e = (warp0 + gi * nwarps) * epw + sl;
        return (e < E && (mask_g == nullptr || mask_g[e] != 0)) ? 1 : 0;
    };
    // producer (lane 0): next box to fetch.  A group whose only environment is masked out is skipped
    // (plain arrive, nothing copied); groups of several small environments are always fetched.
    int64_t p_gi = 0;
    int p_ci = 0, p_act = 0, p_act_next = 0;
    if (lane == 0) {
        p_act = (epw == 1) ? env_active(0, 0) : (ngroups > 0);
        p_act_next = (epw == 1) ? env_active(1, 0) : (ngroups > 1);
    }
    auto issue = [&](int stage) {
        if (p_gi < ngroups) {
            if (p_act) {
                const int64_t x = (warp0 + p_gi * nwarps) * (int64_t)epw * C + (int64_t)p_ci * 32;
                mbar_arrive_expect_tx(&bars[stage], (uint32_t)(BOX * sizeof(T)));
                tma_load_box(ring + (size_t)stage * BOX, &tmap, (int)x, 0, &bars[stage]);
            } else {
                mbar_arrive(&bars[stage]);
            }
            if (++p_ci == cpl) {
                p_ci = 0;
                ++p_gi;
                p_act = p_act_next;
                p_act_next = (epw == 1) ? env_active(p_gi + 1, 0) : (p_gi + 1 < ngroups);
            }
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) issue(s);
    }

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t gi = 0; gi < ngroups; ++gi) {
        const int64_t e = (warp0 + gi * nwarps) * epw + slot;
        const bool active = env_active(gi, slot) != 0;
        T bestJ = T(0);
        int bestI = kNone;
        T x0[N], ob[N], w[DIMC];
        T s0 = T(0), c0 = T(1);
        if (active) {
#pragma unroll
            for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob[i] = obs_g[i * E + e]; }
            if constexpr (MODE != RCG_MODE_MPC) {
#pragma unroll
                for (int i = 0; i < DIMC; ++i) w[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
            }
            if constexpr (SYS != RCG_SYS_2TANK) sincos_t(x0[2], &s0, &c0);
        }
        for (int ci = 0; ci < cpl; ++ci) {
            mbar_wait(&bars[stage], phase);                        // this box has landed
            const int c = cl + ci * seg;
            const bool valid = active && c < C;
            T a[NA][M];
            int bad = 0;
            {
                const T *src = ring + (size_t)stage * BOX + lane;
#pragma unroll
                for (int k = 0; k < NA; ++k)
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        a[k][j] = src[(k * M + j) * 32];
                        if constexpr (LEAN) bad |= nonfinite_bits(a[k][j]);
                    }
            }
            __syncwarp();                                          // every lane has copied its column out
            if (lane == 0) issue(stage);                           // refill the slot
            if (valid) {
                ActorEval<T, SYS, MODE, CS, true, LEAN> ev(S, O, x0, ob, s0, c0, w);
#pragma unroll
                for (int k = 0; k < NA; ++k) ev.stage(k, k + 1 == NA, a[k]);
                const T J = (LEAN && bad) ? (T)CUDART_NAN : ev.J;
                if (J_g) J_g[e * (int64_t)C + c] = J;
                if (bestI == kNone || argmin_better(J, c, bestJ, bestI)) { bestJ = J; bestI = c; }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        for (int off = seg >> 1; off > 0; off >>= 1) {
            const T oJ = __shfl_xor_sync(0xffffffffu, bestJ, off);
            const int oI = __shfl_xor_sync(0xffffffffu, bestI, off);
            if (oI != kNone && (bestI == kNone || argmin_better(oJ, oI, bestJ, bestI))) { bestJ = oJ; bestI = oI; }
        }
        if (active && cl == 0 && bestI != kNone) {
            if (argmin_g) argmin_g[e] = bestI;
            if (Jmin_g) Jmin_g[e] = bestJ;
            if (action_g || accum_g) {
                T act[M];
                const T *cb = cand_g + e * (int64_t)C + bestI;
#pragma unroll
                for (int j = 0; j < M; ++j) act[j] = cb[j * ld];
                if (action_g) {
#pragma unroll
                    for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                }
                if (accum_g) accum_g[e] += stage_obj<T, N, M, true, true>(O, ob, act) * sampling_time;
            }
        }
    }
}

// ---- TMA-staged variant for RUNTIME horizons (any Nactor without a compile-time specialisation) ---------------
// Same staging idea, but a box holds kRtChunk stages of 32 candidates (kRtChunk * m component rows x 32 columns,
// 2 KB in fp64): the lanes evaluate their candidate chunk by chunk while the next chunks of the same candidates -- and
// then the next 32 candidates -- are in flight.  Rows beyond Nactor * m of the last chunk are out of bounds for the
// tensor map and arrive as zeros (never evaluated).  Lean objective only when LEAN (see ActorEval).
constexpr int kRtChunk = 4;
constexpr int kRtStages = 6;
template <typename T, int M>
__host__ __device__ constexpr int tma_rt_smem_bytes() { return kActorWarps * kRtStages * (kRtChunk * M * 32 * (int)sizeof(T) + 8); }

template <typename T, int SYS, int MODE, int CS, bool LEAN>
__global__ void __launch_bounds__(kActorThreads, 2)
actor_cost_tma_rt_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SysDev<T> S,
                         const __grid_constant__ ObjDev<T> O, const __grid_constant__ ActorArgs A,
                         const T *__restrict__ state_sys_g, const T *__restrict__ obs_g, const T *__restrict__ cand_g,
                         const T *__restrict__ w_g, const int32_t *__restrict__ mask_g, T *__restrict__ J_g,
                         int32_t *__restrict__ argmin_g, T *__restrict__ Jmin_g, T *__restrict__ action_g,
                         T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    constexpr int ROWS = kRtChunk * M, BOX = ROWS * 32, STAGES = kRtStages;
    constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    constexpr int kNone = 0x7fffffff;
    extern __shared__ __align__(128) unsigned char actor_smem[];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *ring = reinterpret_cast<T *>(actor_smem) + (size_t)wi * STAGES * BOX;
    uint64_t *bars = reinterpret_cast<uint64_t *>(actor_smem + (size_t)kActorWarps * STAGES * BOX * sizeof(T)) + wi * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int64_t E = A.E;
    const int C = A.C, seg = A.seg, na = O.Nactor;
    const int nchunks = (na + kRtChunk - 1) / kRtChunk;
    const int slot = lane >> A.seg_shift, cl = lane & (seg - 1);
    const int epw = 32 >> A.seg_shift;
    const int cpl = (C + seg - 1) >> A.seg_shift;                       // 32-candidate boxes per environment group
    const int64_t ld = E * (int64_t)C;
    const int64_t warp0 = (int64_t)blockIdx.x * kActorWarps + wi;
    const int64_t nwarps = (int64_t)gridDim.x * kActorWarps;
    const int64_t ngroups = (A.num_groups > warp0) ? (A.num_groups - warp0 + nwarps - 1) / nwarps : 0;

    auto env_active = [&](int64_t gi, int sl) -> int {
        if (gi >= ngroups) return 0;
        const int64_t e = (warp0 + gi * nwarps) * epw + sl;
        return (e < E && (mask_g == nullptr || mask_g[e] != 0)) ? 1 : 0;
    };
    // producer (lane 0): boxes in the order (group, 32-candidate box, chunk)
    int64_t p_gi = 0;
    int p_ci = 0, p_k = 0, p_act = 0;
    if (lane == 0) p_act = (epw == 1) ? env_active(0, 0) : (ngroups > 0);
    auto issue = [&](int stage) {
        if (p_gi < ngroups) {
            if (p_act) {
                const int64_t x = (warp0 + p_gi * nwarps) * (int64_t)epw * C + (int64_t)p_ci * 32;
                mbar_arrive_expect_tx(&bars[stage], (uint32_t)(BOX * sizeof(T)));
                tma_load_box(ring + (size_t)stage * BOX, &tmap, (int)x, p_k * ROWS, &bars[stage]);
            } else {
                mbar_arrive(&bars[stage]);
            }
            if (++p_k == nchunks) {
                p_k = 0;
                if (++p_ci == cpl) {
                    p_ci = 0;
                    ++p_gi;
                    p_act = (epw == 1) ? env_active(p_gi, 0) : (p_gi < ngroups);
                }
            }
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) issue(s);
    }

    int stage = 0;
    uint32_t phase = 0;
    for (int64_t gi = 0; gi < ngroups; ++gi) {
        const int64_t e = (warp0 + gi * nwarps) * epw + slot;
        const bool active = env_active(gi, slot) != 0;
        T bestJ = T(0);
        int bestI = kNone;
        T x0[N], ob[N], w[DIMC];
        T s0 = T(0), c0 = T(1);
        if (active) {
#pragma unroll
            for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob[i] = obs_g[i * E + e]; }
            if constexpr (MODE != RCG_MODE_MPC) {
#pragma unroll
                for (int i = 0; i < DIMC; ++i) w[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
            }
            if constexpr (SYS != RCG_SYS_2TANK) sincos_t(x0[2], &s0, &c0);
        }
        for (int ci = 0; ci < cpl; ++ci) {
            const int c = cl + ci * seg;
            const bool valid = active && c < C;
            ActorEval<T, SYS, MODE, CS, true, LEAN> ev(S, O, x0, ob, s0, c0, w);
            int bad = 0;
            for (int kc = 0; kc < nchunks; ++kc) {
                mbar_wait(&bars[stage], phase);                    // this chunk has landed
                T a[kRtChunk][M];
                {
                    const T *src = ring + (size_t)stage * BOX + lane;
#pragma unroll
                    for (int i = 0; i < kRtChunk; ++i)
#pragma unroll
                        for (int j = 0; j < M; ++j) a[i][j] = src[(i * M + j) * 32];
                }
                __syncwarp();
                if (lane == 0) issue(stage);                       // refill the slot
                if (valid) {
#pragma unroll
                    for (int i = 0; i < kRtChunk; ++i) {
                        const int k = kc * kRtChunk + i;
                        if (k < na) {
                            if constexpr (LEAN) bad |= nonfinite_bits(a[i][0]) | nonfinite_bits(a[i][M - 1]);
                            ev.stage(k, k + 1 == na, a[i]);
                        }
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (valid) {
                const T J = (LEAN && bad) ? (T)CUDART_NAN : ev.J;
                if (J_g) J_g[e * (int64_t)C + c] = J;
                if (bestI == kNone || argmin_better(J, c, bestJ, bestI)) { bestJ = J; bestI = c; }
            }
        }
        for (int off = seg >> 1; off > 0; off >>= 1) {
            const T oJ = __shfl_xor_sync(0xffffffffu, bestJ, off);
            const int oI = __shfl_xor_sync(0xffffffffu, bestI, off);
            if (oI != kNone && (bestI == kNone || argmin_better(oJ, oI, bestJ, bestI))) { bestJ = oJ; bestI = oI; }
        }
        if (active && cl == 0 && bestI != kNone) {
            if (argmin_g) argmin_g[e] = bestI;
            if (Jmin_g) Jmin_g[e] = bestJ;
            if (action_g || accum_g) {
                T act[M];
                const T *cb = cand_g + e * (int64_t)C + bestI;
#pragma unroll
                for (int j = 0; j < M; ++j) act[j] = cb[j * ld];
                if (action_g) {
#pragma unroll
                    for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                }
                if (accum_g) accum_g[e] += stage_obj<T, N, M, true, true>(O, ob, act) * sampling_time;
            }
        }
    }
}

template <typename T>
struct ActorLaunch {
    SysDev<T> S;
    ObjDev<T> O;
    ActorArgs A;
    const T *state_sys, *obs, *cand, *w;
    const int32_t *mask;
    T *J;
    int32_t *argmin;
    T *Jmin, *action, *accum;
    T sampling_time;
    bool rdiag;
    bool lean;                 // diagonal R1 with zero action entries, no target, gamma = 1 (ActorEval LEAN)
    int mode, cs;
    unsigned grid;
    int sms;
    int64_t blocks_needed;
    bool use_tma;              // per-env candidates staged by TMA (tmap valid)
    bool use_tma_rt;           // ... runtime-horizon variant (tmap box = kRtChunk stages)
    CUtensorMap tmap;
    cudaStream_t stream;
};

template <typename T, int SYS, int MODE, int CS, bool RDIAG, int NA, bool LEAN = false>
static void launch_actor_one(const ActorLaunch<T> &L)
{
    if constexpr (RDIAG && NA > 0) {
        if (L.use_tma) {
            constexpr int LL = NA * SysDim<SYS>::m;
            auto kern = actor_cost_tma_kernel<T, SYS, MODE, CS, NA, LEAN>;
            const size_t smem = (size_t)tma_smem_bytes<T, LL>();
            static unsigned long long configured = 0;         // per instantiation, one bit per device
            ensure_dyn_smem(kern, smem, configured);
            // Grid: at least one resident wave (sms x CTAs per SM); beyond that one CTA per kEnvsPerWarp environments
            // per warp rather than a persistent grid: short-lived CTAs let the block scheduler slot the concurrently
            // running rk45_advance launch of another environment block (engine.PipelinedLoop) between them as they
            // retire -- measured on B200 (profiles/r02_overlap_grid_sweep.txt): 0.333 -> 0.300 ms per control interval.
            constexpr int kEnvsPerWarp = 4;
            int64_t pg = (int64_t)L.sms * tma_min_ctas<T, LL>();
            if (L.blocks_needed / kEnvsPerWarp > pg) pg = L.blocks_needed / kEnvsPerWarp;
            const unsigned grid = (unsigned)(L.blocks_needed < pg ? L.blocks_needed : pg);
            kern<<<grid, kActorThreads, smem, L.stream>>>(L.tmap, L.S, L.O, L.A, L.state_sys, L.obs, L.cand, L.w, L.mask, L.J,
                                                         L.argmin, L.Jmin, L.action, L.accum, L.sampling_time);
            return;
        }
    }
    if constexpr (RDIAG && NA == 0) {
        if (L.use_tma_rt) {
            auto kern = actor_cost_tma_rt_kernel<T, SYS, MODE, CS, LEAN>;
            const size_t smem = (size_t)tma_rt_smem_bytes<T, SysDim<SYS>::m>();
            static unsigned long long configured = 0;         // per instantiation, one bit per device
            ensure_dyn_smem(kern, smem, configured);
            const int64_t pg = (int64_t)L.sms * 2;
            const unsigned grid = (unsigned)(L.blocks_needed < pg ? L.blocks_needed : pg);
            kern<<<grid, kActorThreads, smem, L.stream>>>(L.tmap, L.S, L.O, L.A, L.state_sys, L.obs, L.cand, L.w, L.mask, L.J,
                                                         L.argmin, L.Jmin, L.action, L.accum, L.sampling_time);
            return;
        }
    }
    actor_cost_kernel<T, SYS, MODE, CS, RDIAG, NA, LEAN><<<L.grid, kActorThreads, 0, L.stream>>>(
        L.S, L.O, L.A, L.state_sys, L.obs, L.cand, L.w, L.mask, L.J, L.argmin, L.Jmin, L.action, L.accum, L.sampling_time);
}

// Horizon specialisations (diagonal R, the presets' case): every Nactor from 3 to 10 (BASELINE.json's configs use 6, 10, 8,
// the presets default to 3, 5, 10); everything else runs the runtime loop.
template <typename T, int SYS, int MODE, int CS>
static void launch_actor_mc(const ActorLaunch<T> &L)
{
    if (!L.rdiag) { launch_actor_one<T, SYS, MODE, CS, false, 0>(L); return; }
    if constexpr (MODE != RCG_MODE_SQL) {            // SQL has no stage_obj term: nothing to lean out
        if (L.lean) {
            switch (L.O.Nactor) {
            case 3:  launch_actor_one<T, SYS, MODE, CS, true, 3, true>(L); return;
            case 4:  launch_actor_one<T, SYS, MODE, CS, true, 4, true>(L); return;
            case 5:  launch_actor_one<T, SYS, MODE, CS, true, 5, true>(L); return;
            case 6:  launch_actor_one<T, SYS, MODE, CS, true, 6, true>(L); return;
            case 7:  launch_actor_one<T, SYS, MODE, CS, true, 7, true>(L); return;
            case 8:  launch_actor_one<T, SYS, MODE, CS, true, 8, true>(L); return;
            case 9:  launch_actor_one<T, SYS, MODE, CS, true, 9, true>(L); return;
            case 10: launch_actor_one<T, SYS, MODE, CS, true, 10, true>(L); return;
            default: launch_actor_one<T, SYS, MODE, CS, true, 0, true>(L); return;      // runtime horizon, lean
            }
        }
    }
    switch (L.O.Nactor) {
    case 3:  launch_actor_one<T, SYS, MODE, CS, true, 3>(L); return;
    case 4:  launch_actor_one<T, SYS, MODE, CS, true, 4>(L); return;
    case 5:  launch_actor_one<T, SYS, MODE, CS, true, 5>(L); return;
    case 6:  launch_actor_one<T, SYS, MODE, CS, true, 6>(L); return;
    case 7:  launch_actor_one<T, SYS, MODE, CS, true, 7>(L); return;
    case 8:  launch_actor_one<T, SYS, MODE, CS, true, 8>(L); return;
    case 9:  launch_actor_one<T, SYS, MODE, CS, true, 9>(L); return;
    case 10: launch_actor_one<T, SYS, MODE, CS, true, 10>(L); return;
    default: break;
    }
    launch_actor_one<T, SYS, MODE, CS, true, 0>(L);
}

template <typename T, int SYS>
static int launch_actor_sys(const ActorLaunch<T> &L)
{
#define RCG_CASE_CS(MODE)                                                                      \
    switch (L.cs) {                                                                            \
    case RCG_CRITIC_QUAD_LIN:   launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_LIN>(L); break;   \
    case RCG_CRITIC_QUADRATIC:  launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUADRATIC>(L); break;  \
    case RCG_CRITIC_QUAD_NOMIX: launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_NOMIX>(L); break; \
    case RCG_CRITIC_QUAD_MIX:   launch_actor_mc<T, SYS, MODE, RCG_CRITIC_QUAD_MIX>(L); break;   \
    default: return RCG_EINVAL;                                                                \
    }
    if (L.mode == RCG_MODE_MPC) {
        launch_actor_mc<T, SYS, RCG_MODE_MPC, RCG_CRITIC_QUAD_NOMIX>(L);
    } else if (L.mode == RCG_MODE_RQL) {
        RCG_CASE_CS(RCG_MODE_RQL)
    } else if (L.mode == RCG_MODE_SQL) {
        RCG_CASE_CS(RCG_MODE_SQL)
    } else {
        return RCG_EINVAL;
    }
#undef RCG_CASE_CS
    return 0;
}

// one translation unit per (system, dtype): actor_ni.cu, actor_3w.cu, actor_2t.cu
int launch_actor_ni(const ActorLaunch<double> &L);
int launch_actor_ni(const ActorLaunch<float> &L);
int launch_actor_3w(const ActorLaunch<double> &L);
int launch_actor_3w(const ActorLaunch<float> &L);
int launch_actor_2t(const ActorLaunch<double> &L);
int launch_actor_2t(const ActorLaunch<float> &L);

}  // namespace rcg
