// actor_opt_impl.cuh -- the batched stand-in for CtrlOptPred._actor_optimizer (rcognita/controllers.py:1330-1427):
// a bounded minimiser of _actor_cost (controllers.py:1273-1328) over action sequences, one THREAD per
// (environment, start point).  The reference runs scipy's SLSQP with finite-difference gradients (13-41 cost
// evaluations per gradient); here
//   * the gradient is the exact adjoint of the Euler rollout: one forward sweep that keeps the predictor states
//     and the heading trigonometry, one reverse sweep that carries dJ/dstate back through (I + h df/dx)^T and
//     emits dJ/da_k = d(stage term)/da_k + h (df/da)^T lambda_{k+1}  (~3 cost evaluations per gradient);
//   * the minimiser is a projected limited-memory quasi-Newton method: L-BFGS two-loop recursion over the free
//     variables (every inner product masked by the binding set of the box), projected Armijo backtracking,
//     monotone -- see solve() below; the CPU checker under tests/ restates the same algorithm.
// Storage: with a compile-time horizon (NA > 0) the five working vectors and the rollout live in registers /
// thread-local memory; the (s, y) pairs of the quasi-Newton memory always live in a caller-provided global
// workspace laid out [slot][component][thread] (coalesced: consecutive threads = consecutive (env, start)
// columns); with a runtime horizon (NA = 0) everything lives there.
// fp64 only.  Costs are evaluated by the same device functions as rcg_actor_cost (stage_obj, critic,
// euler_step with the rotation-based heading trigonometry), so J here equals rcg_actor_cost's J bit for bit.
#pragma once

#include <math_constants.h>

#include <type_traits>

#include "actor_impl.cuh"

namespace rcg {

constexpr int kOptMem = 6;          // quasi-Newton pairs kept
constexpr int kOptThreads = 128;
constexpr int kOptMaxBacktracks = 40;

// d stage_obj / d (obs, act) scaled by gk (controllers.py:1063-1084).
template <typename T, int N, int M, bool RDIAG>
__device__ __forceinline__ void stage_obj_grad(const ObjDev<T> &O, const T *obs, const T *act, T gk, T *gobs, T *gact)
{
    constexpr int P = N + M;
    T chi[P], g[P];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = obs[i] - O.target[i];
#pragma unroll
    for (int j = 0; j < M; ++j) chi[N + j] = act[j];
    if constexpr (RDIAG) {
#pragma unroll
        for (int i = 0; i < P; ++i) g[i] = (O.R1[i * P + i] + O.R1[i * P + i]) * chi[i];
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < P; ++j) acc += (O.R1[i * P + j] + O.R1[j * P + i]) * chi[j];
            g[i] = acc;
        }
        if (O.stage_struct == RCG_STAGE_BIQUADRATIC) {
#pragma unroll
            for (int i = 0; i < P; ++i) {
                T acc = T(0);
#pragma unroll
                for (int j = 0; j < P; ++j) acc += (O.R2[i * P + j] + O.R2[j * P + i]) * (chi[j] * chi[j]);
                g[i] += T(2) * chi[i] * acc;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) gobs[i] = gk * g[i];
#pragma unroll
    for (int j = 0; j < M; ++j) gact[j] = gk * g[N + j];
}

// d _critic / d (obs, act) (controllers.py:1192-1214; feature order of critic<>() in rcg_device.cuh).
template <typename T, int N, int M, int CS, class WAcc>
__device__ __forceinline__ void critic_grad(const ObjDev<T> &O, const T *obs, const T *act, WAcc w, T *gobs, T *gact)
{
    constexpr int P = N + M;
    T chi[P], g[P];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = obs[i] - O.target[i];
#pragma unroll
    for (int j = 0; j < M; ++j) chi[N + j] = act[j];
#pragma unroll
    for (int i = 0; i < P; ++i) g[i] = T(0);
    int k = 0;
    if constexpr (CS == RCG_CRITIC_QUAD_LIN || CS == RCG_CRITIC_QUADRATIC) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = i; j < P; ++j) {
                const T wk = w(k++);
                g[i] += wk * chi[j];
                g[j] += wk * chi[i];
            }
        if constexpr (CS == RCG_CRITIC_QUAD_LIN) {
#pragma unroll
            for (int i = 0; i < P; ++i) g[i] += w(k++);
        }
    } else if constexpr (CS == RCG_CRITIC_QUAD_NOMIX) {
#pragma unroll
        for (int i = 0; i < P; ++i) g[i] = T(2) * w(k++) * chi[i];
    } else {                                                   // quad-mix: raw observation (:1212)
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = T(2) * w(k++) * obs[i];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const T wk = w(k++);
                g[i] += wk * act[j];
                g[N + j] += wk * obs[i];
            }
#pragma unroll
        for (int j = 0; j < M; ++j) g[N + j] += T(2) * w(k++) * act[j];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) gobs[i] = g[i];
#pragma unroll
    for (int j = 0; j < M; ++j) gact[j] = g[N + j];
}

// lam <- (I + h df/dx)^T lam and ga += h (df/da)^T lam at (x, a), lam = lambda_{k+1} on entry; (sn, cs) = sin/cos
// of the heading x[2] as the forward sweep used them (systems.py:308-323, :370-382, :412-419).
template <typename T, int SYS>
__device__ __forceinline__ void dyn_adjoint(const SysDev<T> &S, T h, const T *x, const T *a, T sn, T cs, T *lam, T *ga)
{
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {
        ga[0] += h * (cs * lam[0] + sn * lam[1]);
        ga[1] += h * lam[2];
        lam[2] += h * (a[0] * (cs * lam[1] - sn * lam[0]));
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {
        ga[0] += h * ((T(1) / S.pars[0]) * lam[3]);
        ga[1] += h * ((T(1) / S.pars[1]) * lam[4]);
        const T l2 = lam[2] + h * (x[3] * (cs * lam[1] - sn * lam[0]));
        const T l3 = lam[3] + h * (cs * lam[0] + sn * lam[1]);
        const T l4 = lam[4] + h * lam[2];
        lam[2] = l2; lam[3] = l3; lam[4] = l4;
    } else {
        const T tau1 = S.pars[0], tau2 = S.pars[1], K1 = S.pars[2], K2 = S.pars[3], K3 = S.pars[4];
        ga[0] += h * ((T(1) / tau1) * K1 * lam[0]);
        const T l0 = lam[0] + h * (-(T(1) / tau1) * lam[0] + (T(1) / tau2) * K2 * lam[1]);
        const T l1 = lam[1] + h * ((T(1) / tau2) * (T(-1) + T(2) * K3 * x[1]) * lam[1]);
        lam[0] = l0; lam[1] = l1;
    }
}

// ---- storage policies ------------------------------------------------------------------------------------
// Working vectors: 0 = x, 1 = g, 2 = d, 3 = x_trial, 4 = g_trial.
template <typename T, int LC, int NK, int N>
struct OptLocalMem {
    T vecs[5][LC];
    T Xs[NK][N], sn[NK], cs[NK];
    __device__ __forceinline__ T &v(int which, int i) { return vecs[which][i]; }
    __device__ __forceinline__ T &X(int k, int i) { return Xs[k][i]; }
    __device__ __forceinline__ T &S(int k) { return sn[k]; }
    __device__ __forceinline__ T &C(int k) { return cs[k]; }
};
template <typename T, int N>
struct OptGlobalMem {
    T *base;                 // this thread's column of the workspace, after the pair storage
    int64_t stride;          // threads in the launch
    int L, NA;
    __device__ __forceinline__ T &v(int which, int i) { return base[((int64_t)which * L + i) * stride]; }
    __device__ __forceinline__ T &X(int k, int i) { return base[((int64_t)5 * L + (int64_t)k * N + i) * stride]; }
    __device__ __forceinline__ T &S(int k) { return base[((int64_t)5 * L + (int64_t)NA * N + k) * stride]; }
    __device__ __forceinline__ T &C(int k) { return base[((int64_t)5 * L + (int64_t)NA * N + NA + k) * stride]; }
};

// Workspace layout (doubles): [0, 2) work-queue counter, [2, 2 + E*S) final cost per start (read by the select
// kernel when S > 1), then the quasi-Newton pairs [2][kOptMem][L][TR] and -- generic kernel only -- the working
// vectors and the rollout [5 L + Nactor (n + 2)][TR]; TR = launched threads <= E*S rounded up to a block.
__host__ __device__ inline int64_t opt_ws_per_thread(int na_runtime, int n, int m, bool generic)
{
    const int64_t L = (int64_t)na_runtime * m;
    return 2 * kOptMem * L + (generic ? 5 * L + (int64_t)na_runtime * (n + 2) : 0);
}
constexpr int64_t kOptWsHeader = 2;

struct OptArgs {
    int64_t E;
    int S, S_shift;                  // start points per environment (power of two <= 32)
    int w_per_env;
    int max_iter;
    double pg_tol, f_tol;
    int grad_only;                   // rcg_actor_grad: one value-and-gradient evaluation, no iteration
    int dynamic;                     // lanes pull problems from the work queue (else: problem = thread index)
};

// One LANE per problem (environment, start point), persistent: a lane that finishes its problem pulls the next one
// from a global counter, so the lanes of a warp stay busy although iteration counts differ by an order of magnitude
// between problems (measured before this change: mean 17, p99 27, max 95 iterations near the goal -- every warp waited
// for its slowest lane).  The body is a state machine -- one pass = one trial point: forward sweep, then either a
// backtracking step or (accepted) the reverse sweep, the quasi-Newton update and the next trial point.
// No residency target: holding the kernel to 3 or 4 blocks per SM (168 / 128 registers, 0.3-0.8 KB of spills) was
// measured 2x SLOWER for 3wrobot_NI N=6 (1.56 -> 3.4 ms per 65,536 solves) and only 8 % faster for 3wrobot N=10.
template <typename T, int SYS, int MODE, int CS, bool RDIAG, int NA>
__global__ void __launch_bounds__(kOptThreads)
actor_opt_kernel(const __grid_constant__ SysDev<T> Sd, const __grid_constant__ ObjDev<T> O,
                 const __grid_constant__ OptArgs A, const T *__restrict__ state_sys_g, const T *__restrict__ obs_g,
                 T *__restrict__ sqn_g, const T *__restrict__ w_g, const int32_t *__restrict__ mask_g,
                 T *__restrict__ ws_g, T *__restrict__ J_g, T *__restrict__ grad_g, int32_t *__restrict__ iters_g,
                 int32_t *__restrict__ nfev_g, int32_t *__restrict__ best_g, T *__restrict__ Jmin_g,
                 T *__restrict__ action_g, T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    constexpr int LC = (NA > 0) ? NA * M : 1, NK = (NA > 0) ? NA : 1;
    constexpr int UK = (NA > 0) ? NA : 1, UL = (NA > 0) ? NA * M : 1;      // unroll factors (1 = keep the runtime loop)
    const int na = (NA > 0) ? NA : O.Nactor;
    const int L = na * M;
    const int64_t E = A.E;
    const int64_t nprob = E << A.S_shift;
    const int64_t col = (int64_t)blockIdx.x * kOptThreads + threadIdx.x;    // this thread's workspace column
    const int64_t TR = (int64_t)gridDim.x * kOptThreads;
    unsigned long long *queue = reinterpret_cast<unsigned long long *>(ws_g);
    T *Jscr = ws_g + kOptWsHeader;

    // ---- box ----
    const T h = O.pred_step_size;
    T lo[M], hi[M], step0 = T(1);
#pragma unroll
    for (int j = 0; j < M; ++j) {
        lo[j] = Sd.has_bnds ? Sd.lo[j] : -CUDART_INF;
        hi[j] = Sd.has_bnds ? Sd.hi[j] : CUDART_INF;
    }
    if (Sd.has_bnds) {
        step0 = T(0);
#pragma unroll
        for (int j = 0; j < M; ++j) step0 = fmax(step0, hi[j] - lo[j]);
    }
    auto clip = [&](T v, int i) {
        const T l = lo[i % M], u = hi[i % M];
        v = (v < l) ? l : v;
        return (v > u) ? u : v;
    };

    // ---- storage ----
    using LMem = OptLocalMem<T, LC, NK, N>;
    using GMem = OptGlobalMem<T, N>;
    typename std::conditional<(NA > 0), LMem, GMem>::type mem;
    T *pairs = ws_g + kOptWsHeader + nprob + col;                      // [2][kOptMem][L][TR]
    if constexpr (NA == 0) {
        mem.base = ws_g + kOptWsHeader + nprob + (int64_t)2 * kOptMem * L * TR + col;
        mem.stride = TR;
        mem.L = L;
        mem.NA = na;
    }
    auto Sp = [&](int slot, int i) -> T & { return pairs[((int64_t)slot * L + i) * TR]; };
    auto Yp = [&](int slot, int i) -> T & { return pairs[((int64_t)(kOptMem + slot) * L + i) * TR]; };

    // ---- per-problem state ----
    T x0[N], ob0[N], w[DIMC];
    T s0 = T(0), c0 = T(1);
    const RegW<T, DIMC> wacc{w};
    int64_t p = -1, e = 0;
    bool need = true, idle = false, first = true, taken = false;
    int npairs = 0, head = 0, stall = 0, bt = 0, iters = 0, nfev = 0;
    T lam_ls = T(1), gs = T(0), J = T(0);

    // ---- forward sweep: cost of the sequence in vector `which`, keeps the rollout ----
    auto forward = [&](int which) -> T {
        T st[N], ob[N], sn = s0, cs = c0, Jc = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) { st[i] = x0[i]; ob[i] = ob0[i]; }
#pragma unroll UK
        for (int k = 0; k < na; ++k) {
            T a[M];
#pragma unroll
            for (int j = 0; j < M; ++j) a[j] = mem.v(which, k * M + j);
#pragma unroll
            for (int i = 0; i < N; ++i) mem.X(k, i) = st[i];
            mem.S(k) = sn;
            mem.C(k) = cs;
            const bool last = (k + 1 == na);
            if constexpr (MODE == RCG_MODE_MPC) {
                Jc += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, ob, a);
            } else if constexpr (MODE == RCG_MODE_RQL) {
                if (!last) Jc += O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, ob, a);
                else Jc += critic<T, N, M, CS>(O, ob, a, wacc);
            } else {
                Jc += critic<T, N, M, CS>(O, ob, a, wacc);
            }
            if (!last) {
                euler_step<T, SYS>(Sd, h, st, a, sn, cs);
#pragma unroll
                for (int i = 0; i < N; ++i) ob[i] = st[i];
            }
        }
        return Jc;
    };
    // ---- reverse sweep: gradient of the last forward(which) into vector `gout` ----
    auto backward = [&](int which, int gout) {
        T lam[N];
#pragma unroll
        for (int i = 0; i < N; ++i) lam[i] = T(0);
#pragma unroll UK
        for (int kk = 0; kk < na; ++kk) {
            const int k = na - 1 - kk;
            T a[M], xk[N], ob[N], ga[M], gobs[N], gact[M];
#pragma unroll
            for (int j = 0; j < M; ++j) { a[j] = mem.v(which, k * M + j); ga[j] = T(0); }
#pragma unroll
            for (int i = 0; i < N; ++i) { xk[i] = mem.X(k, i); ob[i] = (k == 0) ? ob0[i] : xk[i]; }
            if (kk > 0) dyn_adjoint<T, SYS>(Sd, h, xk, a, mem.S(k), mem.C(k), lam, ga);
            const bool use_critic = (MODE == RCG_MODE_SQL) || (MODE == RCG_MODE_RQL && kk == 0);
            if (use_critic) {
                if constexpr (MODE != RCG_MODE_MPC) critic_grad<T, N, M, CS>(O, ob, a, wacc, gobs, gact);
            } else {
                stage_obj_grad<T, N, M, RDIAG>(O, ob, a, O.gamma_pow[k], gobs, gact);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) mem.v(gout, k * M + j) = ga[j] + gact[j];
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) lam[i] += gobs[i];
            }
        }
    };

#pragma unroll 1
    for (;;) {
        if (need) {
            // ---- next problem (environments with mask == 0 are skipped) ----
#pragma unroll 1
            for (;;) {
                if (A.dynamic) p = (int64_t)atomicAdd(queue, 1ull);
                else { p = taken ? nprob : col; taken = true; }
                if (p >= nprob) { idle = true; break; }
                e = p >> A.S_shift;
                if (mask_g == nullptr || mask_g[e] != 0) break;
            }
            need = false;
            if (!idle) {
#pragma unroll
                for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob0[i] = obs_g[i * E + e]; }
                if constexpr (MODE != RCG_MODE_MPC) {
#pragma unroll
                    for (int i = 0; i < DIMC; ++i) w[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
                }
                s0 = T(0);
                c0 = T(1);
                if constexpr (SYS != RCG_SYS_2TANK) sincos_t(x0[2], &s0, &c0);
#pragma unroll UL
                for (int i = 0; i < L; ++i) mem.v(3, i) = clip(sqn_g[(int64_t)i * nprob + p], i);
                first = true;
                npairs = 0; head = 0; stall = 0; bt = 0; iters = 0; nfev = 0;
                lam_ls = T(1); gs = T(0); J = T(0);
            }
        }
        if (__all_sync(0xffffffffu, idle)) break;
        if (idle) continue;

        // ---- one pass of the projected L-BFGS state machine: evaluate the trial point (vector 3) ----
        bool finished = false;
        const T Jt = forward(3);
        bool accepted = true;
        if (!first) {
            ++nfev;
            // projected Armijo test; after the projection g.(x+ - x) can be >= 0 (the descent components clipped away):
            // a trial point is then only accepted if it does not raise the cost, which keeps the iteration monotone
            if (!(Jt <= J + T(1e-4) * fmin(gs, T(0)))) {                // failed: halve, or give up
                accepted = false;
                if (++bt >= kOptMaxBacktracks) {
                    finished = true;
                } else {
                    lam_ls *= T(0.5);
                    gs = T(0);
#pragma unroll UL
                    for (int i = 0; i < L; ++i) {
                        const T xi = mem.v(0, i);
                        const T xt = clip(xi + lam_ls * mem.v(2, i), i);
                        mem.v(3, i) = xt;
                        gs += mem.v(1, i) * (xt - xi);
                    }
                }
            }
        }
        if (accepted) {
            backward(3, 4);
            bool stop = false;
            if (!first) {
                ++iters;
#pragma unroll UL
                for (int i = 0; i < L; ++i) {
                    Sp(head, i) = mem.v(3, i) - mem.v(0, i);
                    Yp(head, i) = mem.v(4, i) - mem.v(1, i);
                }
                head = (head + 1 == kOptMem) ? 0 : head + 1;
                if (npairs < kOptMem) ++npairs;
                if (fabs(J - Jt) <= A.f_tol * fmax(fmax(fabs(J), fabs(Jt)), T(1))) {
                    if (++stall >= 2) stop = true;
                } else {
                    stall = 0;
                }
            }
            J = Jt;
            T pg = T(0);
            uint64_t fr = 0, fr_hi = 0;   // free-set bit mask (L <= 128: two words)
#pragma unroll UL
            for (int i = 0; i < L; ++i) {
                const T xi = mem.v(3, i), gi = mem.v(4, i);
                mem.v(0, i) = xi;
                mem.v(1, i) = gi;
                const T l = lo[i % M], u = hi[i % M];
                const bool binding = (xi <= l && gi > T(0)) || (xi >= u && gi < T(0));
                if (!binding) { if (i < 64) fr |= (1ull << i); else fr_hi |= (1ull << (i - 64)); }
                pg = fmax(pg, fabs(clip(xi - gi, i) - xi));
            }
            first = false;
            if (A.grad_only || stop || !(pg > A.pg_tol) || iters >= A.max_iter) {
                finished = true;
            } else {
                auto is_free = [&](int i) { return ((i < 64 ? (fr >> i) : (fr_hi >> (i - 64))) & 1ull) != 0; };
                // two-loop recursion on the free set: d <- H g_F
#pragma unroll UL
                for (int i = 0; i < L; ++i) mem.v(2, i) = is_free(i) ? mem.v(1, i) : T(0);
                T scale = step0 / pg;
                bool have_scale = false;
                T al[kOptMem], sy[kOptMem];
                // the pair loops stay rolled (al/sy are indexed dynamically): unrolled, the kernel outgrows the
                // instruction cache (ncu: stall_no_instruction was the top stall reason)
#pragma unroll 1
                for (int j = 0; j < npairs; ++j) {
                    int slot = head - 1 - j;
                    slot += (slot < 0) ? kOptMem : 0;
                    const T *sp = &Sp(slot, 0), *yp = &Yp(slot, 0);
                    T a = T(0), ss = T(0), yy = T(0), sq = T(0);
#pragma unroll UL
                    for (int i = 0; i < L; ++i) {
                        const T si = sp[(int64_t)i * TR], yi = yp[(int64_t)i * TR];
                        if (is_free(i)) { a += si * yi; ss += si * si; yy += yi * yi; sq += si * mem.v(2, i); }
                    }
                    al[j] = T(0);
                    sy[j] = T(0);
                    if (a > T(1e-10) * sqrt(ss * yy)) {
                        sy[j] = a;
                        al[j] = sq / a;
                        const T alj = al[j];
#pragma unroll UL
                        for (int i = 0; i < L; ++i)
                            if (is_free(i)) mem.v(2, i) -= alj * yp[(int64_t)i * TR];
                        if (!have_scale) { scale = a / yy; have_scale = true; }
                    }
                }
#pragma unroll UL
                for (int i = 0; i < L; ++i) mem.v(2, i) *= scale;
#pragma unroll 1
                for (int j = npairs - 1; j >= 0; --j) {
                    if (!(sy[j] > T(0))) continue;
                    int slot = head - 1 - j;
                    slot += (slot < 0) ? kOptMem : 0;
                    const T *sp = &Sp(slot, 0), *yp = &Yp(slot, 0);
                    T yr = T(0);
#pragma unroll UL
                    for (int i = 0; i < L; ++i)
                        if (is_free(i)) yr += yp[(int64_t)i * TR] * mem.v(2, i);
                    const T c = al[j] - yr / sy[j];
#pragma unroll UL
                    for (int i = 0; i < L; ++i)
                        if (is_free(i)) mem.v(2, i) += c * sp[(int64_t)i * TR];
                }
                T gd = T(0);
#pragma unroll UL
                for (int i = 0; i < L; ++i) {
                    const T di = -mem.v(2, i);
                    mem.v(2, i) = di;
                    gd += mem.v(1, i) * di;
                }
                if (!(gd < T(0)) || !isfinite(gd)) {                      // not a descent direction: restart
                    npairs = 0;
#pragma unroll UL
                    for (int i = 0; i < L; ++i) mem.v(2, i) = is_free(i) ? -mem.v(1, i) * (step0 / pg) : T(0);
                }
                lam_ls = T(1);
                bt = 0;
                gs = T(0);
#pragma unroll UL
                for (int i = 0; i < L; ++i) {
                    const T xi = mem.v(0, i);
                    const T xt = clip(xi + mem.v(2, i), i);
                    mem.v(3, i) = xt;
                    gs += mem.v(1, i) * (xt - xi);
                }
            }
        }
        if (finished) {
            // vector 0 = the minimiser (monotone: last accepted iterate), vector 1 its gradient, J its cost
            if (!A.grad_only) {
#pragma unroll UL
                for (int i = 0; i < L; ++i) sqn_g[(int64_t)i * nprob + p] = mem.v(0, i);
                Jscr[p] = J;
            } else if (grad_g) {
#pragma unroll UL
                for (int i = 0; i < L; ++i) grad_g[(int64_t)i * nprob + p] = mem.v(1, i);
            }
            if (J_g) J_g[p] = J;
            if (iters_g) iters_g[p] = iters;
            if (nfev_g) nfev_g[p] = nfev;
            if (!A.grad_only && A.S == 1) {
                // _actor_optimizer returns action_sqn[:dim_input] (:1427); upd_accum_obj of the sampling step (:1093)
                if (best_g) best_g[e] = 0;
                if (Jmin_g) Jmin_g[e] = J;
                T act[M];
#pragma unroll
                for (int j = 0; j < M; ++j) act[j] = mem.v(0, j);
                if (action_g) {
#pragma unroll
                    for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                }
                if (accum_g) accum_g[e] += stage_obj<T, N, M, RDIAG, RDIAG>(O, ob0, act) * sampling_time;
            }
            need = true;
        }
    }
}

// S > 1: best start per environment (np.argmin order over the final costs), action hand-over, upd_accum_obj.
template <typename T, int SYS, bool RDIAG>
__global__ void __launch_bounds__(256)
actor_opt_select_kernel(const __grid_constant__ ObjDev<T> O, int64_t E, int S, const T *__restrict__ obs_g,
                        const T *__restrict__ sqn_g, const T *__restrict__ Jscr, const int32_t *__restrict__ mask_g,
                        int32_t *__restrict__ best_g, T *__restrict__ Jmin_g, T *__restrict__ action_g,
                        T *__restrict__ accum_g, T sampling_time)
{
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || (mask_g && mask_g[e] == 0)) return;
    const int64_t nprob = E * S;
    T bestJ = Jscr[e * S];
    int bestI = 0;
    for (int s = 1; s < S; ++s) {
        const T Js = Jscr[e * S + s];
        if (argmin_better(Js, s, bestJ, bestI)) { bestJ = Js; bestI = s; }
    }
    if (best_g) best_g[e] = bestI;
    if (Jmin_g) Jmin_g[e] = bestJ;
    if (action_g || accum_g) {
        T act[M], ob[N];
#pragma unroll
        for (int j = 0; j < M; ++j) act[j] = sqn_g[(int64_t)j * nprob + e * S + bestI];
        if (action_g) {
#pragma unroll
            for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
        }
        if (accum_g) {
#pragma unroll
            for (int i = 0; i < N; ++i) ob[i] = obs_g[i * E + e];
            accum_g[e] += stage_obj<T, N, M, RDIAG, RDIAG>(O, ob, act) * sampling_time;
        }
    }
}

template <typename T>
struct OptLaunch {
    SysDev<T> S;
    ObjDev<T> O;
    OptArgs A;
    const T *state_sys, *obs, *w;
    T *sqn;
    const int32_t *mask;
    T *ws, *J, *grad;
    int32_t *iters, *nfev, *best;
    T *Jmin, *action, *accum;
    T sampling_time;
    bool rdiag;
    int mode, cs;
    unsigned grid;
    int sms;
    cudaStream_t stream;
};

// horizons with register-resident working vectors; anything else (and dense R) runs the generic kernel
// (every horizon from 3 to 10: a horizon without a specialisation costs ~5x -- 7.8 ms against 1.6 ms per 65,536 Sys3WRobotNI
// solves at N = 7 against N = 6 in round 1)
__host__ inline bool opt_horizon_specialised(int na) { return na >= 3 && na <= 10; }

template <typename T, int SYS, int MODE, int CS, bool RDIAG, int NA>
static void launch_opt_one(const OptLaunch<T> &L)
{
    auto kern = actor_opt_kernel<T, SYS, MODE, CS, RDIAG, NA>;
    unsigned grid = L.grid;                                  // one thread per problem ...
    if (L.A.dynamic) {                                       // ... or a persistent grid pulling from the work queue
        static int occ = 0;                                  // per instantiation
        if (occ == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kOptThreads, 0) != cudaSuccess || occ < 1) occ = 1;
        }
        const unsigned resident = (unsigned)(L.sms * occ);
        if (grid > resident) grid = resident;
    }
    kern<<<grid, kOptThreads, 0, L.stream>>>(L.S, L.O, L.A, L.state_sys, L.obs, L.sqn, L.w, L.mask, L.ws, L.J, L.grad, L.iters,
                                             L.nfev, L.best, L.Jmin, L.action, L.accum, L.sampling_time);
    if (L.A.dynamic && L.A.S > 1 && (L.best || L.Jmin || L.action || L.accum)) {
        actor_opt_select_kernel<T, SYS, RDIAG><<<(unsigned)((L.A.E + 255) / 256), 256, 0, L.stream>>>(
            L.O, L.A.E, L.A.S, L.obs, L.sqn, L.ws + kOptWsHeader, L.mask, L.best, L.Jmin, L.action, L.accum, L.sampling_time);
    }
}

template <typename T, int SYS, int MODE, int CS>
static void launch_opt_mc(const OptLaunch<T> &L)
{
    if (!L.rdiag) { launch_opt_one<T, SYS, MODE, CS, false, 0>(L); return; }
    switch (L.O.Nactor) {
    case 3:  launch_opt_one<T, SYS, MODE, CS, true, 3>(L); return;
    case 4:  launch_opt_one<T, SYS, MODE, CS, true, 4>(L); return;
    case 5:  launch_opt_one<T, SYS, MODE, CS, true, 5>(L); return;
    case 6:  launch_opt_one<T, SYS, MODE, CS, true, 6>(L); return;
    case 7:  launch_opt_one<T, SYS, MODE, CS, true, 7>(L); return;
    case 8:  launch_opt_one<T, SYS, MODE, CS, true, 8>(L); return;
    case 9:  launch_opt_one<T, SYS, MODE, CS, true, 9>(L); return;
    case 10: launch_opt_one<T, SYS, MODE, CS, true, 10>(L); return;
    default: break;
    }
    launch_opt_one<T, SYS, MODE, CS, true, 0>(L);
}

template <typename T, int SYS>
static int launch_opt_sys(const OptLaunch<T> &L)
{
#define RCG_CASE_CS(MODE)                                                                    \
    switch (L.cs) {                                                                          \
    case RCG_CRITIC_QUAD_LIN:   launch_opt_mc<T, SYS, MODE, RCG_CRITIC_QUAD_LIN>(L); break;   \
    case RCG_CRITIC_QUADRATIC:  launch_opt_mc<T, SYS, MODE, RCG_CRITIC_QUADRATIC>(L); break;  \
    case RCG_CRITIC_QUAD_NOMIX: launch_opt_mc<T, SYS, MODE, RCG_CRITIC_QUAD_NOMIX>(L); break; \
    case RCG_CRITIC_QUAD_MIX:   launch_opt_mc<T, SYS, MODE, RCG_CRITIC_QUAD_MIX>(L); break;   \
    default: return RCG_EINVAL;                                                              \
    }
    if (L.mode == RCG_MODE_MPC) {
        launch_opt_mc<T, SYS, RCG_MODE_MPC, RCG_CRITIC_QUAD_NOMIX>(L);
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

// one translation unit per system: actor_opt_ni.cu, actor_opt_3w.cu, actor_opt_2t.cu
int launch_opt_ni(const OptLaunch<double> &L);
int launch_opt_3w(const OptLaunch<double> &L);
int launch_opt_2t(const OptLaunch<double> &L);

}  // namespace rcg
