// actor_opt_quad.cuh -- rcg_actor_opt with G LANES PER PROBLEM (round 2): the same projected L-BFGS iteration as
// actor_opt_kernel (actor_opt_impl.cuh; stand-in for CtrlOptPred._actor_optimizer, rcognita/controllers.py:1330-1427),
// re-laid for the hardware after the round-2 source-level profile of the one-lane kernel (10.6 of 32 threads active per
// issued instruction, 43 % of the stalls on the quasi-Newton pair loads from the global workspace, FP64 pipe 3.5 %):
//   * everything a problem owns -- the kOptMem (s, y) pairs, x, g, d and the rollout of the accepted point -- lives in SHARED
//     memory, [component][problem] with the problem stride padded so that a quad's strided accesses are bank-conflict free;
//     nothing is spilled and no pair ever travels to L2 / HBM;
//   * the vector work (two-loop recursion, projected-gradient norm, free set) is DISTRIBUTED: lane r of a quad owns the
//     components r, r + G, ...; inner products are completed by log2(G) xor-shuffles inside the quad;
//   * the line search is PARALLEL: lane r evaluates the trial point of step length lam0 * 2^-r, the first lane (longest step)
//     that passes the projected Armijo test wins -- exactly the point sequential halving would have accepted, so the
//     iterates are those of the one-lane kernel up to the summation order of the distributed inner products; the quads of a
//     warp therefore stay in the same phase (one line-search pass, one gradient pass, one two-loop recursion per
//     iteration) instead of waiting for each other's rejected trial points;
//   * lane 0 of the quad keeps the rollout of its trial point; when a shorter step wins, one more pass re-evaluates it on
//     lane 0.  The reverse sweep at the accepted point is a serial recurrence over the horizon: lane 0 runs it alone and
//     publishes x+, g+ and the new pair.
// A persistent grid pulls problems from the work-queue counter in the workspace header like the one-lane kernel.
// Costs come from the same device functions as rcg_actor_cost (stage_obj, critic, euler_step), in the same order.
#pragma once

#include "actor_opt_impl.cuh"

namespace rcg {

constexpr int kQThreads = 128;

template <int SYS, int MODE, int CS, int NA, int G>
struct QuadCfg {
    static constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m;
    static constexpr int L = NA * M, LPL = (L + G - 1) / G;
    static constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    static constexpr bool W_REGS = (MODE == RCG_MODE_MPC) || DIMC <= 8;       // critic weights in registers, else read through L1
    static constexpr int P = kQThreads / G;                  // problems per block
    static constexpr int PS = P + 16 / G;                    // padded problem stride (doubles): (r * PS + q) distinct mod 16
    static constexpr int RS = N + 2;                         // rollout record per stage: state, sin, cos
    static constexpr int oS = 0, oY = kOptMem * L, oX = 2 * kOptMem * L, oG = oX + L, oD = oG + L, oR = oD + L;
    static constexpr int oA = oR + NA * RS;                  // two-loop coefficients al[kOptMem], 1 / (s.y)[kOptMem]
    static constexpr int TOT = oA + 2 * kOptMem;
    static constexpr int smem_bytes = TOT * PS * (int)sizeof(double);
    // Partial unrolling of the sweeps (measured, Sys3WRobot N=10, 262,144 solves): rolled 5.88 ms; forward sweep x2 5.50, x5
    // 5.23, FULLY unrolled 6.63 ms (the instruction-fetch cliff again); reverse sweep x2 on top of forward x2: 5.35; pair loops
    // x2: slower (5.71).  Hence: the largest divisor of the horizon within a per-system stage budget for the forward sweep
    // (a Sys3WRobot stage is ~1.6x a Sys3WRobotNI stage), two stages of the reverse sweep, pair loops rolled.
    static constexpr int largest_divisor(int n, int cap) { int d = 1; for (int c = 2; c <= cap && c <= n; ++c) if (n % c == 0) d = c; return d; }
    static constexpr int unroll_fwd = largest_divisor(NA, N >= 5 ? 5 : 7);
    static constexpr int unroll_bwd = (NA % 2 == 0) ? 2 : 1;
    static constexpr int fit = (228 * 1024) / (smem_bytes + 1024);
    static constexpr int min_blocks = fit < 1 ? 1 : (fit > 4 ? 4 : fit);
    static_assert(G == 2 || G == 4 || G == 8, "lanes per problem");
    static_assert(G % M == 0, "a lane's components must share one box row");
    static_assert(kOptMaxBacktracks % G == 0, "give-up count is a whole number of line-search passes");
};

template <typename T>
struct GlobW {                                               // critic weights read through the read-only path
    const T *p;
    int64_t stride;
    __device__ __forceinline__ T operator()(int i) const { return __ldg(p + (int64_t)i * stride); }
};

template <int SYS, int MODE, int CS, int NA, int G>
__global__ void __launch_bounds__(kQThreads, QuadCfg<SYS, MODE, CS, NA, G>::min_blocks)
actor_opt_quad_kernel(const __grid_constant__ SysDev<double> Sd, const __grid_constant__ ObjDev<double> O,
                      const __grid_constant__ OptArgs A, const double *__restrict__ state_sys_g,
                      const double *__restrict__ obs_g, double *__restrict__ sqn_g, const double *__restrict__ w_g,
                      const int32_t *__restrict__ mask_g, double *__restrict__ ws_g, double *__restrict__ J_g,
                      int32_t *__restrict__ iters_g, int32_t *__restrict__ nfev_g, int32_t *__restrict__ best_g,
                      double *__restrict__ Jmin_g, double *__restrict__ action_g, double *__restrict__ accum_g,
                      double sampling_time)
{
    using T = double;
    using Q = QuadCfg<SYS, MODE, CS, NA, G>;
    constexpr int N = Q::N, M = Q::M, L = Q::L, LPL = Q::LPL, DIMC = Q::DIMC, PS = Q::PS, RS = Q::RS;
    constexpr int oS = Q::oS, oY = Q::oY, oX = Q::oX, oG = Q::oG, oD = Q::oD, oR = Q::oR, oA = Q::oA;
    constexpr bool RDIAG = true;
    extern __shared__ double smq[];

    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x % G;                           // lane of the quad
    const int qbase = lane & ~(G - 1);                       // first warp lane of this quad
    const unsigned qmask = ((1u << G) - 1u) << qbase;
    T *col = smq + threadIdx.x / G;                          // this problem's column: element `off` at col[off * PS]
    auto at = [&](int off) -> T & { return col[off * PS]; };

    const int64_t E = A.E;
    const int64_t nprob = E << A.S_shift;
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
    auto clipj = [&](T v, int j) {                           // j = component % M, compile-time at every call site
        v = (v < lo[j]) ? lo[j] : v;
        return (v > hi[j]) ? hi[j] : v;
    };
    // the components a lane owns (r, r + G, ...) all sit in box row r % M because G % M == 0
    T mylo = lo[0], myhi = hi[0];
#pragma unroll
    for (int j = 1; j < M; ++j)
        if (r % M == j) { mylo = lo[j]; myhi = hi[j]; }
    auto clipmine = [&](T v) {
        v = (v < mylo) ? mylo : v;
        return (v > myhi) ? myhi : v;
    };
    const T lam_r = __longlong_as_double((long long)(1023 - r) << 52);       // 2^-r
    const T lam_pass = __longlong_as_double((long long)(1023 - G) << 52);    // 2^-G

    // quad reductions
    auto qsum = [&](T v) {
#pragma unroll
        for (int s = 1; s < G; s <<= 1) v += __shfl_xor_sync(qmask, v, s);
        return v;
    };
    auto qmax = [&](T v) {
#pragma unroll
        for (int s = 1; s < G; s <<= 1) v = fmax(v, __shfl_xor_sync(qmask, v, s));
        return v;
    };

    // ---- per-problem state (identical in the lanes of a quad unless noted) ----
    T x0[N], ob0[N], wreg[Q::W_REGS ? DIMC : 1];
    GlobW<T> wglob{w_g, 1};
    T s0 = T(0), c0 = T(1);
    int64_t p = -1, e = 0;
    bool need = true, idle = false, first = true;
    int npairs = 0, head = 0, stall = 0, iters = 0, nfev = 0;
    T J = T(0), pg = T(0);
    unsigned fr = 0;                                          // per lane: bit j = owned component r + G j is free

    // The sweeps are ROLLED loops over the horizon (one copy of a stage in the instruction stream): the fully unrolled
    // first version of this kernel was 7.6 k instructions (122 KB) for Sys3WRobot N = 10 and ran at the speed of the one-lane
    // kernel -- both were bound by instruction fetch, not by arithmetic or memory.
    auto stage_cost = [&](int k, bool last, const T *ob, const T *a) -> T {
        if constexpr (MODE == RCG_MODE_MPC) {
            return O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, ob, a);
        } else {
            if (MODE == RCG_MODE_RQL && !last) return O.gamma_pow[k] * stage_obj<T, N, M, RDIAG, RDIAG>(O, ob, a);
            if constexpr (Q::W_REGS) return critic<T, N, M, CS>(O, ob, a, RegW<T, DIMC>{wreg});
            else return critic<T, N, M, CS>(O, ob, a, wglob);
        }
    };

    // ---- forward sweep at clip(x + lam d): cost, and g.(x+ - x) of that trial point; lane 0 of the quad keeps its rollout ----
    auto forward = [&](T lam, T &gs_out) -> T {
        T st[N], ob[N], sn = s0, cs = c0, Jc = T(0), gs = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) { st[i] = x0[i]; ob[i] = ob0[i]; }
        const T *xp = col + oX * PS, *dp = col + oD * PS, *gp = col + oG * PS;
        T *rp = col + oR * PS;
#pragma unroll Q::unroll_fwd
        for (int k = 0; k < NA; ++k) {
            T a[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const T xi = xp[j * PS];
                a[j] = clipj(fma(lam, dp[j * PS], xi), j);
                gs += gp[j * PS] * (a[j] - xi);
            }
            if (r == 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) rp[i * PS] = st[i];
                rp[N * PS] = sn;
                rp[(N + 1) * PS] = cs;
            }
            const bool last = (k + 1 == NA);
            Jc += stage_cost(k, last, ob, a);
            if (!last) {
                euler_step<T, SYS>(Sd, h, st, a, sn, cs);
#pragma unroll
                for (int i = 0; i < N; ++i) ob[i] = st[i];
            }
            xp += M * PS; dp += M * PS; gp += M * PS; rp += RS * PS;
        }
        gs_out = gs;
        return Jc;
    };

    // ---- reverse sweep at clip(x + lam d) over the stored rollout (called by lane 0 of the quad only); publishes the new
    //      (s, y) pair into `slot` (when have_prev), x <- x+, g <- gradient ----
    auto backward = [&](T lam, bool have_prev, int slot) {
        T lamv[N];
#pragma unroll
        for (int i = 0; i < N; ++i) lamv[i] = T(0);
        T *xp = col + (oX + (NA - 1) * M) * PS, *gp = col + (oG + (NA - 1) * M) * PS;
        const T *dp = col + (oD + (NA - 1) * M) * PS, *rp = col + (oR + (NA - 1) * RS) * PS;
        T *sp = col + (oS + slot * L + (NA - 1) * M) * PS, *yp = col + (oY + slot * L + (NA - 1) * M) * PS;
#pragma unroll Q::unroll_bwd
        for (int k = NA - 1; k >= 0; --k) {
            T a[M], xo[M], ga[M], xk[N], ob[N], gobs[N], gact[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                xo[j] = xp[j * PS];
                a[j] = clipj(fma(lam, dp[j * PS], xo[j]), j);
                ga[j] = T(0);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) { xk[i] = rp[i * PS]; ob[i] = (k == 0) ? ob0[i] : xk[i]; }
            if (k < NA - 1) dyn_adjoint<T, SYS>(Sd, h, xk, a, rp[N * PS], rp[(N + 1) * PS], lamv, ga);
            bool done = false;
            if constexpr (MODE != RCG_MODE_MPC) {
                if (MODE == RCG_MODE_SQL || k == NA - 1) {
                    if constexpr (Q::W_REGS) critic_grad<T, N, M, CS>(O, ob, a, RegW<T, DIMC>{wreg}, gobs, gact);
                    else critic_grad<T, N, M, CS>(O, ob, a, wglob, gobs, gact);
                    done = true;
                }
            }
            if constexpr (MODE != RCG_MODE_SQL) {
                if (!done) stage_obj_grad<T, N, M, RDIAG>(O, ob, a, O.gamma_pow[k], gobs, gact);
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const T gt = ga[j] + gact[j];
                if (have_prev) {
                    sp[j * PS] = a[j] - xo[j];
                    yp[j * PS] = gt - gp[j * PS];
                }
                xp[j * PS] = a[j];
                gp[j * PS] = gt;
            }
            if (k > 0) {
#pragma unroll
                for (int i = 0; i < N; ++i) lamv[i] += gobs[i];
            }
            xp -= M * PS; gp -= M * PS; dp -= M * PS; rp -= RS * PS; sp -= M * PS; yp -= M * PS;
        }
    };

#pragma unroll 1
    for (;;) {
        if (need) {
            // ---- next problem (environments with mask == 0 are skipped); lane 0 of the quad pulls ----
#pragma unroll 1
            for (;;) {
                unsigned long long pp = 0;
                if (r == 0) pp = atomicAdd(queue, 1ull);
                p = (int64_t)__shfl_sync(qmask, pp, qbase);
                if (p >= nprob) { idle = true; break; }
                e = p >> A.S_shift;
                if (mask_g == nullptr || mask_g[e] != 0) break;
            }
            need = false;
            if (!idle) {
#pragma unroll
                for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob0[i] = obs_g[i * E + e]; }
                if constexpr (MODE != RCG_MODE_MPC) {
                    if constexpr (Q::W_REGS) {
#pragma unroll
                        for (int i = 0; i < DIMC; ++i) wreg[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
                    } else {
                        wglob.p = A.w_per_env ? w_g + e : w_g;
                        wglob.stride = A.w_per_env ? E : 1;
                    }
                }
                s0 = T(0);
                c0 = T(1);
                if constexpr (SYS != RCG_SYS_2TANK) sincos_t(x0[2], &s0, &c0);
#pragma unroll
                for (int j = 0; j < LPL; ++j) {
                    const int i = r + G * j;
                    if (i < L) {
                        at(oX + i) = clipmine(sqn_g[(int64_t)i * nprob + p]);
                        at(oG + i) = T(0);
                        at(oD + i) = T(0);
                    }
                }
                first = true;
                npairs = 0; head = 0; stall = 0; iters = 0; nfev = 0;
                J = T(0);
                __syncwarp(qmask);
            }
        }
        if (__all_sync(0xffffffffu, idle)) break;
        if (idle) continue;

        // ---- line search: G trial points per pass (lane r tries lam0 * 2^-r), the longest passing step wins.  The reverse
        //      sweep needs the rollout of the accepted point and only lane 0 keeps one: when a shorter step wins, one more
        //      pass re-evaluates it on lane 0.  First pass of a problem: lam0 = 0, i.e. the start point itself. ----
        bool finished = false;
        T lam_acc = T(0), Jt = T(0);
        {
            T lam0 = first ? T(0) : T(1);
            int bt = 0;
            bool confirm = first;
#pragma unroll 1
            for (;;) {
                const T lam = lam0 * lam_r;
                T gs;
                const T Jr = forward(lam, gs);
                // projected Armijo test; after the projection g.(x+ - x) can be >= 0 (the descent components clipped
                // away): a trial point is then only accepted if it does not raise the cost (monotone iteration)
                const bool ok = confirm || (Jr <= J + T(1e-4) * fmin(gs, T(0)));
                const unsigned b = __ballot_sync(qmask, ok) & qmask;
                if (b) {
                    const int src = __ffs(b) - 1;
                    if (!confirm) nfev += src - qbase + 1;
                    if (src == qbase) {
                        lam_acc = __shfl_sync(qmask, lam, qbase);
                        Jt = __shfl_sync(qmask, Jr, qbase);
                        break;
                    }
                    lam0 = __shfl_sync(qmask, lam, src);
                    confirm = true;
                    continue;
                }
                nfev += G;
                bt += G;
                if (bt >= kOptMaxBacktracks) { finished = true; break; }
                lam0 *= lam_pass;
            }
            __syncwarp(qmask);
        }

        if (!finished) {
            // ---- gradient at the accepted point: a serial recurrence over the horizon, run by lane 0 alone (it reads and
            //      rewrites x and g stage by stage; the other lanes wait at the barrier) ----
            if (r == 0) backward(lam_acc, !first, head);
            __syncwarp(qmask);
            bool stop = false;
            if (!first) {
                ++iters;
                head = (head + 1 == kOptMem) ? 0 : head + 1;
                if (npairs < kOptMem) ++npairs;
                if (fabs(J - Jt) <= A.f_tol * fmax(fmax(fabs(J), fabs(Jt)), T(1))) {
                    if (++stall >= 2) stop = true;
                } else {
                    stall = 0;
                }
            }
            J = Jt;
            first = false;
            // projected-gradient norm and free set over the owned components
            T gv[LPL];
            T pgl = T(0);
            fr = 0;
#pragma unroll
            for (int j = 0; j < LPL; ++j) {
                const int i = r + G * j;
                gv[j] = T(0);
                if (i < L) {
                    const T xi = at(oX + i);
                    gv[j] = at(oG + i);
                    const bool binding = (xi <= mylo && gv[j] > T(0)) || (xi >= myhi && gv[j] < T(0));
                    if (!binding) fr |= (1u << j);
                    pgl = fmax(pgl, fabs(clipmine(xi - gv[j]) - xi));
                }
            }
            pg = qmax(pgl);
            if (stop || !(pg > A.pg_tol) || iters >= A.max_iter) {
                finished = true;
            } else {
                // ---- two-loop recursion on the free set, distributed over the quad: d <- -H g_F ----
                T dv[LPL];
#pragma unroll
                for (int j = 0; j < LPL; ++j) dv[j] = ((fr >> j) & 1u) ? gv[j] : T(0);
                // (al_j = (s_j.q) / (s_j.y_j) and the second loop's (y_j.r) / (s_j.y_j) are formed with one reciprocal per pair,
                // the initial scaling (s.y) / (y.y) of the newest usable pair with one division after the loop)
                T a_first = T(0), yy_first = T(1);
                bool have_scale = false;
#pragma unroll 1
                for (int jj = 0; jj < npairs; ++jj) {
                    int slot = head - 1 - jj;
                    slot += (slot < 0) ? kOptMem : 0;
                    const T *sp = col + (oS + slot * L + r) * PS, *yp = col + (oY + slot * L + r) * PS;
                    T yv[LPL];
                    T a = T(0), ss = T(0), yy = T(0), sq = T(0);
#pragma unroll
                    for (int j = 0; j < LPL; ++j) {
                        yv[j] = T(0);
                        if (r + G * j < L) {
                            const T si = sp[j * G * PS];
                            yv[j] = yp[j * G * PS];
                            if ((fr >> j) & 1u) { a += si * yv[j]; ss += si * si; yy += yv[j] * yv[j]; sq += si * dv[j]; }
                        }
                    }
                    a = qsum(a); ss = qsum(ss); yy = qsum(yy); sq = qsum(sq);
                    T alv = T(0), rho = T(0);
                    if (a > T(1e-10) * sqrt(ss * yy)) {
                        rho = T(1) / a;
                        alv = sq * rho;
#pragma unroll
                        for (int j = 0; j < LPL; ++j)
                            if ((fr >> j) & 1u) dv[j] -= alv * yv[j];
                        if (!have_scale) { a_first = a; yy_first = yy; have_scale = true; }
                    }
                    // every lane of the quad holds the same (al, rho) bit for bit (xor-butterfly sums); lane 0 keeps them for
                    // the second loop
                    if (r == 0) {
                        at(oA + jj) = alv;
                        at(oA + kOptMem + jj) = rho;
                    }
                }
                __syncwarp(qmask);
                const T scale = have_scale ? a_first / yy_first : step0 / pg;
#pragma unroll
                for (int j = 0; j < LPL; ++j) dv[j] *= scale;
#pragma unroll 1
                for (int jj = npairs - 1; jj >= 0; --jj) {
                    const T rho = at(oA + kOptMem + jj);
                    if (!(rho > T(0))) continue;
                    int slot = head - 1 - jj;
                    slot += (slot < 0) ? kOptMem : 0;
                    const T *sp = col + (oS + slot * L + r) * PS, *yp = col + (oY + slot * L + r) * PS;
                    T sv[LPL];
                    T yr = T(0);
#pragma unroll
                    for (int j = 0; j < LPL; ++j) {
                        sv[j] = T(0);
                        if (r + G * j < L) {
                            sv[j] = sp[j * G * PS];
                            if ((fr >> j) & 1u) yr += yp[j * G * PS] * dv[j];
                        }
                    }
                    yr = qsum(yr);
                    const T c = at(oA + jj) - yr * rho;
#pragma unroll
                    for (int j = 0; j < LPL; ++j)
                        if ((fr >> j) & 1u) dv[j] += c * sv[j];
                }
                T gd = T(0);
#pragma unroll
                for (int j = 0; j < LPL; ++j) {
                    dv[j] = -dv[j];
                    gd += gv[j] * dv[j];
                }
                gd = qsum(gd);
                if (!(gd < T(0)) || !isfinite(gd)) {                      // not a descent direction: restart
                    npairs = 0;
                    const T sc = step0 / pg;
#pragma unroll
                    for (int j = 0; j < LPL; ++j) dv[j] = ((fr >> j) & 1u) ? -gv[j] * sc : T(0);
                }
#pragma unroll
                for (int j = 0; j < LPL; ++j)
                    if (r + G * j < L) at(oD + r + G * j) = dv[j];
                __syncwarp(qmask);
            }
        }
        if (finished) {
            // x = the minimiser (monotone: last accepted iterate), J its cost
#pragma unroll
            for (int j = 0; j < LPL; ++j) {
                const int i = r + G * j;
                if (i < L) sqn_g[(int64_t)i * nprob + p] = at(oX + i);
            }
            if (r == 0) {
                Jscr[p] = J;
                if (J_g) J_g[p] = J;
                if (iters_g) iters_g[p] = iters;
                if (nfev_g) nfev_g[p] = nfev;
                if (A.S == 1) {
                    // _actor_optimizer returns action_sqn[:dim_input] (:1427); upd_accum_obj of the sampling step (:1093)
                    if (best_g) best_g[e] = 0;
                    if (Jmin_g) Jmin_g[e] = J;
                    T act[M];
#pragma unroll
                    for (int j = 0; j < M; ++j) act[j] = at(oX + j);
                    if (action_g) {
#pragma unroll
                        for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                    }
                    if (accum_g) accum_g[e] += stage_obj<T, N, M, RDIAG, RDIAG>(O, ob0, act) * sampling_time;
                }
            }
            __syncwarp(qmask);                                            // the column is reused by the next problem
            need = true;
        }
    }
}

// Launch: a persistent grid of resident blocks (capped by the number of problems).  Returns false when the device cannot
// hold one block of this instantiation (the caller then runs the one-lane kernel).
template <int SYS, int MODE, int CS, int NA, int G>
static bool launch_opt_quad(const OptLaunch<double> &L)
{
    using Q = QuadCfg<SYS, MODE, CS, NA, G>;
    auto kern = actor_opt_quad_kernel<SYS, MODE, CS, NA, G>;
    static int occ[64] = {0};                    // per instantiation and device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return false;
    if (occ[dev] == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Q::smem_bytes) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[dev], kern, kQThreads, Q::smem_bytes) != cudaSuccess ||
            occ[dev] < 1) {
            (void)cudaGetLastError();
            occ[dev] = -1;
        }
    }
    if (occ[dev] < 0) return false;
    const int64_t nprob = L.A.E * (int64_t)L.A.S;
    int64_t grid = (nprob + Q::P - 1) / Q::P;
    const int64_t resident = (int64_t)L.sms * occ[dev];
    if (grid > resident) grid = resident;
    kern<<<(unsigned)grid, kQThreads, Q::smem_bytes, L.stream>>>(L.S, L.O, L.A, L.state_sys, L.obs, L.sqn, L.w, L.mask, L.ws, L.J,
                                                                 L.iters, L.nfev, L.best, L.Jmin, L.action, L.accum,
                                                                 L.sampling_time);
    return true;
}

// Dispatch over mode / critic structure / horizon for one system.  Returns 0 when a quad kernel (and, for S > 1, the
// select kernel) was launched, 1 when this problem shape has no quad instantiation (the caller runs the one-lane kernel).
template <int SYS, int MODE, int CS, int G>
static int launch_optq_g(const OptLaunch<double> &L)
{
    bool ok = false;
    switch (L.O.Nactor) {
    case 3:  ok = launch_opt_quad<SYS, MODE, CS, 3, G>(L); break;
    case 4:  ok = launch_opt_quad<SYS, MODE, CS, 4, G>(L); break;
    case 5:  ok = launch_opt_quad<SYS, MODE, CS, 5, G>(L); break;
    case 6:  ok = launch_opt_quad<SYS, MODE, CS, 6, G>(L); break;
    case 7:  ok = launch_opt_quad<SYS, MODE, CS, 7, G>(L); break;
    case 8:  ok = launch_opt_quad<SYS, MODE, CS, 8, G>(L); break;
    case 9:  ok = launch_opt_quad<SYS, MODE, CS, 9, G>(L); break;
    case 10: ok = launch_opt_quad<SYS, MODE, CS, 10, G>(L); break;
    default: return 1;
    }
    if (!ok) return 1;
    if (L.A.S > 1 && (L.best || L.Jmin || L.action || L.accum)) {
        actor_opt_select_kernel<double, SYS, true><<<(unsigned)((L.A.E + 255) / 256), 256, 0, L.stream>>>(
            L.O, L.A.E, L.A.S, L.obs, L.sqn, L.ws + kOptWsHeader, L.mask, L.best, L.Jmin, L.action, L.accum, L.sampling_time);
    }
    return 0;
}

template <int SYS, int MODE, int CS>
static int launch_optq_mc(const OptLaunch<double> &L)
{
    // eight lanes per problem (twice the resident warps at 128 registers, three weights per lane) were measured slower on
    // every bench point: Sys3WRobot N=10 7.35 against 5.90 ms per 262,144 solves, Sys3WRobotNI N=6 1.45 against 0.95 ms
    return launch_optq_g<SYS, MODE, CS, 4>(L);
}

template <int SYS>
static int launch_optq_sys(const OptLaunch<double> &L)
{
    if (!L.rdiag || L.A.grad_only || !L.A.dynamic) return 1;
#define RCG_CASE_CS(MODE)                                                                        \
    switch (L.cs) {                                                                              \
    case RCG_CRITIC_QUAD_LIN:   return launch_optq_mc<SYS, MODE, RCG_CRITIC_QUAD_LIN>(L);         \
    case RCG_CRITIC_QUADRATIC:  return launch_optq_mc<SYS, MODE, RCG_CRITIC_QUADRATIC>(L);        \
    case RCG_CRITIC_QUAD_NOMIX: return launch_optq_mc<SYS, MODE, RCG_CRITIC_QUAD_NOMIX>(L);       \
    case RCG_CRITIC_QUAD_MIX:   return launch_optq_mc<SYS, MODE, RCG_CRITIC_QUAD_MIX>(L);         \
    default: return RCG_EINVAL;                                                                  \
    }
    if (L.mode == RCG_MODE_MPC) return launch_optq_mc<SYS, RCG_MODE_MPC, RCG_CRITIC_QUAD_NOMIX>(L);
    if (L.mode == RCG_MODE_RQL) { RCG_CASE_CS(RCG_MODE_RQL) }
    if (L.mode == RCG_MODE_SQL) { RCG_CASE_CS(RCG_MODE_SQL) }
#undef RCG_CASE_CS
    return RCG_EINVAL;
}

// one translation unit per system: actor_optq_ni.cu, actor_optq_3w.cu
int launch_optq_ni(const OptLaunch<double> &L);
int launch_optq_3w(const OptLaunch<double> &L);

}  // namespace rcg
