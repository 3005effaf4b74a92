// critic_fit.cu -- batched CtrlOptPred._critic_optimizer (rcognita/controllers.py:1248-1271): the minimiser
// of _critic_cost (:1216-1245) over w in [Wmin, Wmax], started from w_critic_init, one environment per thread.
//
// _critic_cost is linear least squares in w:  J_c(w) = 1/2 * sum_r (phi_r . w - b_r)^2  with, for the
// buffer rows k = Ncritic-1 .. 1,  phi_r = phi(o[k-1], a[k-1])  and  b_r = gamma * Q(o[k], a[k]; w_prev) +
// stage_obj(o[k-1], a[k-1]).  There are K = Ncritic-1 (3 in every preset) rows against 3..35 unknowns, so the
// minimiser is not unique; the reference takes whatever SLSQP (maxiter 200, tol 1e-7) returns from
// w_critic_init.  Here: proximal-point iterations  w <- argmin J_c(w) + mu/2 |w - w_prev_iter|^2  within the
// box with a decreasing mu (continuation), each solved in the K-dimensional dual by a semismooth Newton method -- the prox solution is
// w = clip(w0 + Phi^T lam) with  mu*lam + Phi*clip(w0 + Phi^T lam) - b = 0  (a piecewise-linear monotone
// equation; Newton with the Gram matrix of the un-clipped columns terminates when the clip pattern stops
// changing).  The iterate with the smallest J_c is returned, so the result is never worse than w_critic_init.
// Parity is on the fitted COST against the reference's SLSQP result (tests/golden/critic_fit.json), never on
// the weights (hard part 5).
#include <cstdlib>
#include <type_traits>

#include <math_constants.h>

#include "rcg_host.h"

namespace rcg {

constexpr int kFitMaxK = 15;            // Ncritic - 1 <= 15

template <typename T>
struct GlobalWF {
    const T *w;
    int64_t stride, idx;
    __device__ __forceinline__ T operator()(int i) const { return w[i * stride + idx]; }
};

// Lane of this thread: the thread index itself (masked), or -- second phase of a two-phase fit -- entry
// `thread index` of the compacted list of environments whose first-phase budget ran out.  -1: nothing to do.
__device__ __forceinline__ int64_t fit_lane(int64_t E, const int32_t *mask, const int32_t *lane_list, const int32_t *lane_count)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane_list) return (idx < *lane_count) ? (int64_t)lane_list[idx] : -1;
    if (idx >= E || (mask && mask[idx] == 0)) return -1;
    return idx;
}
// First phase: an environment that used up its budget is queued for the second phase and writes nothing.
__device__ __forceinline__ bool fit_defer(bool exhausted, int64_t e, int32_t *todo_list, int32_t *todo_count)
{
    if (!todo_list || !exhausted) return false;
    todo_list[atomicAdd(todo_count, 1)] = (int32_t)e;
    return true;
}

template <int N, int M, int CS, bool RDIAG>
__global__ void __launch_bounds__(128)
critic_fit_kernel(const __grid_constant__ ObjDev<double> O, int64_t E, const double *__restrict__ obs_buf,
                  const double *__restrict__ act_buf, double *__restrict__ wprev_g, double lo, double hi,
                  const double *__restrict__ winit_g, double *__restrict__ w_g, const int32_t *__restrict__ mask,
                  double mu_rel, int max_outer, int max_newton, int max_evals, int update_prev, double *__restrict__ Jc_out,
                  int max_ls, const int32_t *__restrict__ lane_list, const int32_t *__restrict__ lane_count,
                  int32_t *__restrict__ todo_list, int32_t *__restrict__ todo_count)
{
    constexpr int D = dim_critic_c(CS, N, M);
    const int64_t e = fit_lane(E, mask, lane_list, lane_count);
    if (e < 0) return;
    const int K = O.Ncritic - 1;

    double Phi[kFitMaxK * D], b[kFitMaxK], lam[kFitMaxK], lt[kFitMaxK], F[kFitMaxK], dl[kFitMaxK];
    double H[kFitMaxK * kFitMaxK];
    double w0[D], wb[D];
    signed char st[D];

    // rows of the least-squares problem (controllers.py:1230-1242)
    const GlobalWF<double> wp{wprev_g, E, e};          // read in full before the optional update below
    double trace = 0, bb = 0;
    for (int r = 0; r < K; ++r) {
        const int k = K - r;
        double op[N], on[N], ap[M], an[M];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            op[i] = obs_buf[((int64_t)(k - 1) * N + i) * E + e];
            on[i] = obs_buf[((int64_t)k * N + i) * E + e];
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            ap[j] = act_buf[((int64_t)(k - 1) * M + j) * E + e];
            an[j] = act_buf[((int64_t)k * M + j) * E + e];
        }
        double ph[D];
        critic_phi<double, N, M, CS>(O, op, ap, ph);
#pragma unroll
        for (int j = 0; j < D; ++j) { Phi[r * D + j] = ph[j]; trace = fma(ph[j], ph[j], trace); }
        b[r] = O.gamma * critic<double, N, M, CS>(O, on, an, wp) + stage_obj<double, N, M, RDIAG>(O, op, ap);
        bb = fma(b[r], b[r], bb);
    }
    auto clipw = [&](double z) { return z < lo ? lo : (z > hi ? hi : z); };
    auto cost = [&](const double *w) {
        double J = 0;
        for (int r = 0; r < K; ++r) {
            double s = -b[r];
            for (int j = 0; j < D; ++j) s = fma(Phi[r * D + j], w[j], s);
            J = fma(0.5 * s, s, J);
        }
        return J;
    };
    for (int j = 0; j < D; ++j) { w0[j] = clipw(winit_g ? winit_g[j] : w_g[j * E + e]); wb[j] = w0[j]; }
    double Jbest = cost(w0);

    const double J0 = Jbest;
    int evals = 0;                                         // dual / Newton passes spent (budget: max_evals, 0 = none)
    if (K >= 1 && trace > 0 && isfinite(trace) && isfinite(bb)) {
        double mu = 0;
        auto zj = [&](const double *l, int j) {
            double z = w0[j];
            for (int r = 0; r < K; ++r) z = fma(Phi[r * D + j], l[r], z);
            return z;
        };
        auto dual = [&](const double *l) {             // the convex dual objective whose gradient is F
            double q = 0, s = 0;
            for (int r = 0; r < K; ++r) { q = fma(l[r], l[r], q); s = fma(-b[r], l[r], s); }
            s = fma(0.5 * mu, q, s);
            for (int j = 0; j < D; ++j) {
                const double z = zj(l, j);
                s += z < lo ? lo * z - 0.5 * lo * lo : (z > hi ? hi * z - 0.5 * hi * hi : 0.5 * z * z);
            }
            return s;
        };
        for (int outer = 0; outer < max_outer && evals < max_evals; ++outer) {
            mu = mu_rel * trace / K;                       // continuation: the proximal weight shrinks 100x per stage
            mu_rel *= 1e-2;
            for (int r = 0; r < K; ++r) lam[r] = 0;
            for (int it = 0; it < max_newton && evals < max_evals; ++it) {
                ++evals;
                for (int r = 0; r < K; ++r) {
                    F[r] = mu * lam[r] - b[r];
                    for (int q = 0; q <= r; ++q) H[r * kFitMaxK + q] = (q == r) ? mu : 0.0;
                }
                for (int j = 0; j < D; ++j) {
                    const double z = zj(lam, j), w = clipw(z);
                    const int s = (z > lo ? 1 : (z < lo ? -1 : 0)) + (z > hi ? 1 : (z < hi ? -1 : 0));   // -2 | 0 | 2 (+-1 on a bound)
                    st[j] = (signed char)s;
                    for (int r = 0; r < K; ++r) F[r] = fma(Phi[r * D + j], w, F[r]);
                    if (z > lo && z < hi)
                        for (int r = 0; r < K; ++r)
                            for (int q = 0; q <= r; ++q)
                                H[r * kFitMaxK + q] = fma(Phi[r * D + j], Phi[q * D + j], H[r * kFitMaxK + q]);
                }
                double fmax = 0;
                for (int r = 0; r < K; ++r) fmax = fmax > fabs(F[r]) ? fmax : fabs(F[r]);
                if (fmax <= 1e-13 * sqrt(bb)) break;               // stationary to rounding
                // Cholesky H = L L^T (lower, in place); H >= mu*I is positive definite
                bool spd = true;
                for (int r = 0; r < K && spd; ++r) {
                    for (int q = 0; q <= r; ++q) {
                        double s = H[r * kFitMaxK + q];
                        for (int c = 0; c < q; ++c) s = fma(-H[r * kFitMaxK + c], H[q * kFitMaxK + c], s);
                        if (q == r) {
                            if (!(s > 0)) { spd = false; break; }
                            H[r * kFitMaxK + r] = sqrt(s);
                        } else {
                            H[r * kFitMaxK + q] = s / H[q * kFitMaxK + q];
                        }
                    }
                }
                if (!spd) break;
                for (int r = 0; r < K; ++r) {                      // L y = -F
                    double s = -F[r];
                    for (int c = 0; c < r; ++c) s = fma(-H[r * kFitMaxK + c], dl[c], s);
                    dl[r] = s / H[r * kFitMaxK + r];
                }
                for (int r = K - 1; r >= 0; --r) {                 // L^T dl = y
                    double s = dl[r];
                    for (int c = r + 1; c < K; ++c) s = fma(-H[c * kFitMaxK + r], dl[c], s);
                    dl[r] = s / H[r * kFitMaxK + r];
                }
                double slope = 0;
                for (int r = 0; r < K; ++r) slope = fma(F[r], dl[r], slope);
                if (!(slope < 0)) break;
                const double D0 = dual(lam);
                double a = 1.0;
                bool ok = false;
                for (int ls = 0; ls < max_ls && evals < max_evals; ++ls) {     // Armijo backtracking on the dual
                    ++evals;
                    for (int r = 0; r < K; ++r) lt[r] = fma(a, dl[r], lam[r]);
                    if (dual(lt) <= D0 + 1e-4 * a * slope + 1e-14 * fabs(D0)) { ok = true; break; }
                    a *= 0.5;
                }
                if (!ok) break;
                for (int r = 0; r < K; ++r) lam[r] = lt[r];
                if (a == 1.0) {                                    // full step inside one linear piece: exact
                    bool same = true;
                    for (int j = 0; j < D; ++j) {
                        const double z = zj(lam, j);
                        const int s = (z > lo ? 1 : (z < lo ? -1 : 0)) + (z > hi ? 1 : (z < hi ? -1 : 0));
                        same = same && (s == st[j]);
                    }
                    if (same) break;
                }
            }
            double wn[D];
            for (int j = 0; j < D; ++j) wn[j] = clipw(zj(lam, j));
            const double Jn = cost(wn);
            if (Jn < Jbest) {                              // commit: new prox centre = best iterate
                Jbest = Jn;
                for (int j = 0; j < D; ++j) { w0[j] = wn[j]; wb[j] = wn[j]; }
                if (Jn <= 1e-12 * J0 || Jn <= 1e-20 * bb) break;
            }                                              // else: keep the centre, try the next (smaller) mu
        }
    }
    if (fit_defer(evals >= max_evals, e, todo_list, todo_count)) return;
    for (int j = 0; j < D; ++j) w_g[j * E + e] = wb[j];
    if (update_prev)                                   // controllers.py:1471: w_critic_prev = w_critic
        for (int j = 0; j < D; ++j) wprev_g[j * E + e] = wb[j];
    if (Jc_out) Jc_out[e] = Jbest;
}


// ---- fast path: K = Ncritic - 1 <= 3 rows (every preset) ------------------------------------------------
// Same algorithm, laid out for the register file: every critic feature is a product u[a] * u[b] of two entries
// of the row vector u = [chi, 1] (quad-mix: [observation, action, 1]) with compile-time (a, b), so the three
// rows are kept as 3 x (p+1) registers and the features are recomputed on the fly (one DMUL each) instead of
// being stored -- the generic kernel's K x D matrix in local memory thrashed L1 (528 ms for 1 M fits of the
// 28-weight 'quadratic' critic).  The two D-vectors (prox centre, trial point) live in shared memory,
// [weight][thread], conflict-free.  One Newton iteration = two unrolled passes over the D weights.
template <int CS, int N, int M>
struct FeatIdx {
    static constexpr int P = N + M;
    // entry j of phi = u[a(j)] * u[b(j)]; u[P] == 1
    __host__ __device__ static constexpr int a(int j)
    {
        if (CS == RCG_CRITIC_QUAD_LIN || CS == RCG_CRITIC_QUADRATIC) {
            int k = 0;
            for (int i = 0; i < P; ++i)
                for (int l = i; l < P; ++l) { if (k == j) return i; ++k; }
            return j - k;                                   // linear tail of quad-lin: chi[j - k] * 1
        } else if (CS == RCG_CRITIC_QUAD_NOMIX) {
            return j;
        } else {
            if (j < N) return j;
            if (j < N + N * M) return (j - N) / M;
            return N + (j - N - N * M);
        }
    }
    __host__ __device__ static constexpr int b(int j)
    {
        if (CS == RCG_CRITIC_QUAD_LIN || CS == RCG_CRITIC_QUADRATIC) {
            int k = 0;
            for (int i = 0; i < P; ++i)
                for (int l = i; l < P; ++l) { if (k == j) return l; ++k; }
            return P;
        } else if (CS == RCG_CRITIC_QUAD_NOMIX) {
            return j;
        } else {
            if (j < N) return j;
            if (j < N + N * M) return N + (j - N) % M;
            return N + (j - N - N * M);
        }
    }
};

template <int J, int END, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (J < END) {
        f(std::integral_constant<int, J>{});
        static_for<J + 1, END>(f);
    }
}

template <int N, int M, int CS, bool RDIAG>
__global__ void __launch_bounds__(128)
critic_fit3_kernel(const __grid_constant__ ObjDev<double> O, int64_t E, const double *__restrict__ obs_buf,
                   const double *__restrict__ act_buf, double *__restrict__ wprev_g, double lo, double hi,
                   const double *__restrict__ winit_g, double *__restrict__ w_g, const int32_t *__restrict__ mask,
                   double mu_rel, int max_outer, int max_newton, int max_evals, int update_prev, double *__restrict__ Jc_out,
                  int max_ls, const int32_t *__restrict__ lane_list, const int32_t *__restrict__ lane_count,
                  int32_t *__restrict__ todo_list, int32_t *__restrict__ todo_count)
{
    constexpr int D = dim_critic_c(CS, N, M), P = N + M;
    using FI = FeatIdx<CS, N, M>;
    extern __shared__ double fit_smem[];                   // [2][D][blockDim.x]
    const int64_t e = fit_lane(E, mask, lane_list, lane_count);
    if (e < 0) return;
    const int K = O.Ncritic - 1;
    double *wa = fit_smem + threadIdx.x;                   // prox centre / best iterate
    double *wbuf = fit_smem + (size_t)D * blockDim.x + threadIdx.x;
    const int ws = blockDim.x;

    double u[3][P + 1], b[3];
    const GlobalWF<double> wp{wprev_g, E, e};
    double trace = 0, bb = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int i = 0; i <= P; ++i) u[r][i] = 0.0;
        b[r] = 0.0;
        if (r < K) {
            const int k = K - r;
            double op[N], on[N], ap[M], an[M];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                op[i] = obs_buf[((int64_t)(k - 1) * N + i) * E + e];
                on[i] = obs_buf[((int64_t)k * N + i) * E + e];
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                ap[j] = act_buf[((int64_t)(k - 1) * M + j) * E + e];
                an[j] = act_buf[((int64_t)k * M + j) * E + e];
            }
#pragma unroll
            for (int i = 0; i < N; ++i) u[r][i] = (CS == RCG_CRITIC_QUAD_MIX) ? op[i] : op[i] - O.target[i];
#pragma unroll
            for (int j = 0; j < M; ++j) u[r][N + j] = ap[j];
            u[r][P] = 1.0;
            b[r] = O.gamma * critic<double, N, M, CS>(O, on, an, wp) + stage_obj<double, N, M, RDIAG>(O, op, ap);
        }
        bb = fma(b[r], b[r], bb);
    }
    static_for<0, D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double ph = u[r][FI::a(j)] * u[r][FI::b(j)];
            trace = fma(ph, ph, trace);
        }
    });
    auto clipw = [&](double z) { return z < lo ? lo : (z > hi ? hi : z); };
    // Each pass recomputes the feature products from u.  Without this fence the compiler keeps all 3*D products
    // alive across the passes (common-subexpression elimination): 255 registers and kilobytes of spills for the
    // 28- and 35-weight critics.  The empty asm makes the rows opaque per pass at no run-time cost.
    auto fence_rows = [&]() {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i <= P; ++i) asm volatile("" : "+d"(u[r][i]));
    };
    // start point and its cost
    fence_rows();
    double r0 = -b[0], r1 = -b[1], r2 = -b[2];
    static_for<0, D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const double w = clipw(winit_g ? winit_g[j] : w_g[j * E + e]);
        wa[j * ws] = w;
        r0 = fma(u[0][FI::a(j)] * u[0][FI::b(j)], w, r0);
        r1 = fma(u[1][FI::a(j)] * u[1][FI::b(j)], w, r1);
        r2 = fma(u[2][FI::a(j)] * u[2][FI::b(j)], w, r2);
    });
    double Jbest = 0.5 * (r0 * r0 + r1 * r1 + r2 * r2);
    const double J0 = Jbest;

    int evals = 0;                                         // dual / Newton passes spent (budget: max_evals, 0 = none)
    if (K >= 1 && trace > 0 && isfinite(trace) && isfinite(bb)) {
        for (int outer = 0; outer < max_outer && evals < max_evals; ++outer) {
            const double mu = mu_rel * trace / K;          // continuation: the proximal weight shrinks 100x per stage
            mu_rel *= 1e-2;
            double l0 = 0, l1 = 0, l2 = 0;
            for (int it = 0; it < max_newton && evals < max_evals; ++it) {
                ++evals;
                // pass A: residual F, Gram matrix H of the free columns, dual value and clip pattern at lam
                double F0 = mu * l0 - b[0], F1 = mu * l1 - b[1], F2 = mu * l2 - b[2];
                double H00 = mu, H10 = 0, H11 = mu, H20 = 0, H21 = 0, H22 = mu;
                double D0 = fma(0.5 * mu, l0 * l0 + l1 * l1 + l2 * l2, -(b[0] * l0 + b[1] * l1 + b[2] * l2));
                uint64_t lowA = 0, highA = 0;
                fence_rows();
                static_for<0, D>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    const double p0 = u[0][FI::a(j)] * u[0][FI::b(j)], p1 = u[1][FI::a(j)] * u[1][FI::b(j)],
                                 p2 = u[2][FI::a(j)] * u[2][FI::b(j)];
                    const double z = fma(p2, l2, fma(p1, l1, fma(p0, l0, wa[j * ws])));
                    const bool below = !(z > lo), above = !(z < hi);
                    const double w = below ? lo : (above ? hi : z);
                    D0 += below ? lo * z - 0.5 * lo * lo : (above ? hi * z - 0.5 * hi * hi : 0.5 * z * z);
                    lowA |= (uint64_t)below << j;
                    highA |= (uint64_t)above << j;
                    F0 = fma(p0, w, F0); F1 = fma(p1, w, F1); F2 = fma(p2, w, F2);
                    if (!below && !above) {
                        H00 = fma(p0, p0, H00); H10 = fma(p1, p0, H10); H11 = fma(p1, p1, H11);
                        H20 = fma(p2, p0, H20); H21 = fma(p2, p1, H21); H22 = fma(p2, p2, H22);
                    }
                });
                if (fmax(fmax(fabs(F0), fabs(F1)), fabs(F2)) <= 1e-13 * sqrt(bb)) break;    // stationary to rounding
                // Cholesky of the 3 x 3 system, solve H d = -F
                if (!(H00 > 0)) break;
                const double c00 = sqrt(H00), c10 = H10 / c00, c20 = H20 / c00;
                const double t11 = H11 - c10 * c10;
                if (!(t11 > 0)) break;
                const double c11 = sqrt(t11), c21 = (H21 - c20 * c10) / c11;
                const double t22 = H22 - c20 * c20 - c21 * c21;
                if (!(t22 > 0)) break;
                const double c22 = sqrt(t22);
                const double y0 = -F0 / c00, y1 = (-F1 - c10 * y0) / c11, y2 = (-F2 - c20 * y0 - c21 * y1) / c22;
                const double d2 = y2 / c22, d1 = (y1 - c21 * d2) / c11, d0 = (y0 - c10 * d1 - c20 * d2) / c00;
                const double slope = F0 * d0 + F1 * d1 + F2 * d2;
                if (!(slope < 0)) break;
                // pass B: Armijo backtracking on the dual; the clip pattern of the trial point comes for free
                double a = 1.0;
                bool ok = false, same = false;
                for (int ls = 0; ls < max_ls && evals < max_evals; ++ls) {
                    ++evals;
                    const double t0 = fma(a, d0, l0), t1 = fma(a, d1, l1), t2 = fma(a, d2, l2);
                    double Dt = fma(0.5 * mu, t0 * t0 + t1 * t1 + t2 * t2, -(b[0] * t0 + b[1] * t1 + b[2] * t2));
                    uint64_t lowB = 0, highB = 0;
                    fence_rows();
                    static_for<0, D>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        const double p0 = u[0][FI::a(j)] * u[0][FI::b(j)], p1 = u[1][FI::a(j)] * u[1][FI::b(j)],
                                     p2 = u[2][FI::a(j)] * u[2][FI::b(j)];
                        const double z = fma(p2, t2, fma(p1, t1, fma(p0, t0, wa[j * ws])));
                        const bool below = !(z > lo), above = !(z < hi);
                        Dt += below ? lo * z - 0.5 * lo * lo : (above ? hi * z - 0.5 * hi * hi : 0.5 * z * z);
                        lowB |= (uint64_t)below << j;
                        highB |= (uint64_t)above << j;
                    });
                    if (Dt <= D0 + 1e-4 * a * slope + 1e-14 * fabs(D0)) {
                        ok = true;
                        l0 = t0; l1 = t1; l2 = t2;
                        same = (a == 1.0) && lowB == lowA && highB == highA;
                        break;
                    }
                    if (todo_list) {                       // first phase of a two-phase fit: the unit step failed, so this
                        evals = max_evals;                 // problem needs a line search -- the second phase (one warp per
                        break;                             // problem) restarts it and searches the direction EXACTLY
                    }
                    a *= 0.5;
                }
                if (!ok || same) break;                    // full step inside one linear piece: exact
            }
            // prox step result, its cost
            double q0 = -b[0], q1 = -b[1], q2 = -b[2];
            fence_rows();
            static_for<0, D>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                const double p0 = u[0][FI::a(j)] * u[0][FI::b(j)], p1 = u[1][FI::a(j)] * u[1][FI::b(j)],
                             p2 = u[2][FI::a(j)] * u[2][FI::b(j)];
                const double w = clipw(fma(p2, l2, fma(p1, l1, fma(p0, l0, wa[j * ws]))));
                wbuf[j * ws] = w;
                q0 = fma(p0, w, q0); q1 = fma(p1, w, q1); q2 = fma(p2, w, q2);
            });
            const double Jn = 0.5 * (q0 * q0 + q1 * q1 + q2 * q2);
            if (Jn < Jbest) {                              // commit: new prox centre = best iterate
                Jbest = Jn;
                static_for<0, D>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    wa[j * ws] = wbuf[j * ws];
                });
                if (Jn <= 1e-12 * J0 || Jn <= 1e-20 * bb) break;
            }                                              // else: keep the centre, try the next (smaller) mu
        }
    }
    if (fit_defer(evals >= max_evals, e, todo_list, todo_count)) return;
    static_for<0, D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const double w = wa[j * ws];
        w_g[j * E + e] = w;
        if (update_prev) wprev_g[j * E + e] = w;           // controllers.py:1471
    });
    if (Jc_out) Jc_out[e] = Jbest;
}

// ---- second phase, K <= 3: ONE WARP PER PROBLEM ------------------------------------------------------------
// The environments that exhaust the first-phase budget need hundreds to thousands of dual evaluations; one
// environment per lane, each evaluation is a serial pass over the D weights (a ~2 us dependency chain for the
// 28-weight critic) and the launch lasts as long as its slowest lane.  Here lane j owns weight j (and j + 32 when
// D > 32): its three feature products stay in registers, a pass is a handful of FP64 operations per lane followed by a
// butterfly all-reduce (every lane ends up with the same sums, so control flow stays warp-uniform), the clip pattern
// is a pair of ballots.  Same algorithm and constants as critic_fit3_kernel; sums are formed in a different order,
// so results agree with the one-lane version to rounding, not bit for bit.  Warps pull problems from a queue.
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Residency: held to 6 blocks of 128 threads per SM (80 registers, ~50 bytes of spills): the kernel is issue-bound with
// long shuffle / FP64 dependency chains per warp, 24 resident warps instead of 12 (152 registers unconstrained) buy 16 %:
// 15.8 -> 13.3 ms per 1 M in-loop fits (5 blocks: 96 registers, 8 blocks: 64 registers and 200 bytes of spills, both slower).
template <int N, int M, int CS, bool RDIAG>
__global__ void __launch_bounds__(128, 6)
critic_fit3_warp_kernel(const __grid_constant__ ObjDev<double> O, int64_t E, const double *__restrict__ obs_buf,
                        const double *__restrict__ act_buf, double *__restrict__ wprev_g, double lo, double hi,
                        const double *__restrict__ winit_g, double *__restrict__ w_g, double mu_rel0, int max_outer,
                        int max_newton, int max_evals, int update_prev, double *__restrict__ Jc_out, int max_ls,
                        const int32_t *__restrict__ lane_list, const int32_t *__restrict__ lane_count,
                        int32_t *__restrict__ queue, const int32_t *__restrict__ mask = nullptr)
{
    constexpr int D = dim_critic_c(CS, N, M), P = N + M;
    constexpr int SLOTS = (D + 31) / 32;                    // weights per lane (1, or 2 for the 35-weight critic)
    using FI = FeatIdx<CS, N, M>;
    const int lane = threadIdx.x & 31;
    const int K = O.Ncritic - 1;
    // work items: the compacted list of a first phase, or (lane_list == NULL) every environment with mask != 0
    const int count = lane_list ? *lane_count : (int)E;
    auto clipw = [&](double z) { return z < lo ? lo : (z > hi ? hi : z); };

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(queue, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= count) return;
        const int64_t e = lane_list ? (int64_t)lane_list[item] : (int64_t)item;
        if (!lane_list && mask && mask[e] == 0) continue;

        // rows u[r] = [chi(o[k-1], a[k-1]), 1] and right-hand sides b[r] (every lane keeps a copy: 3 x 8 doubles)
        double u[3][P + 1], b[3];
        double bb = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int i = 0; i <= P; ++i) u[r][i] = 0.0;
            b[r] = 0.0;
            if (r < K) {
                const int k = K - r;
                double op[N], on[N], ap[M], an[M];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    op[i] = obs_buf[((int64_t)(k - 1) * N + i) * E + e];
                    on[i] = obs_buf[((int64_t)k * N + i) * E + e];
                }
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    ap[j] = act_buf[((int64_t)(k - 1) * M + j) * E + e];
                    an[j] = act_buf[((int64_t)k * M + j) * E + e];
                }
#pragma unroll
                for (int i = 0; i < N; ++i) u[r][i] = (CS == RCG_CRITIC_QUAD_MIX) ? op[i] : op[i] - O.target[i];
#pragma unroll
                for (int j = 0; j < M; ++j) u[r][N + j] = ap[j];
                u[r][P] = 1.0;
                const GlobalWF<double> wp{wprev_g, E, e};
                b[r] = O.gamma * critic<double, N, M, CS>(O, on, an, wp) + stage_obj<double, N, M, RDIAG>(O, op, ap);
            }
            bb = fma(b[r], b[r], bb);
        }
        // this lane's weights: feature products, prox centre
        double p0[SLOTS], p1[SLOTS], p2[SLOTS], wc[SLOTS], wn[SLOTS];
        bool own[SLOTS];
        double tr = 0;
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
            const int j = lane + 32 * sl;
            own[sl] = j < D;
            const int ja = own[sl] ? FI::a(j) : 0, jb = own[sl] ? FI::b(j) : 0;
            p0[sl] = own[sl] ? u[0][ja] * u[0][jb] : 0.0;
            p1[sl] = own[sl] ? u[1][ja] * u[1][jb] : 0.0;
            p2[sl] = own[sl] ? u[2][ja] * u[2][jb] : 0.0;
            wc[sl] = own[sl] ? clipw(winit_g ? winit_g[j] : w_g[j * E + e]) : 0.0;
            wn[sl] = wc[sl];
            tr += p0[sl] * p0[sl] + p1[sl] * p1[sl] + p2[sl] * p2[sl];
        }
        const double trace = warp_sum(tr);
        double r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) { r0 = fma(p0[sl], wc[sl], r0); r1 = fma(p1[sl], wc[sl], r1); r2 = fma(p2[sl], wc[sl], r2); }
        r0 = warp_sum(r0) - b[0]; r1 = warp_sum(r1) - b[1]; r2 = warp_sum(r2) - b[2];
        double Jbest = 0.5 * (r0 * r0 + r1 * r1 + r2 * r2);
        const double J0 = Jbest;
        double mu_rel = mu_rel0;

        if (K >= 1 && trace > 0 && isfinite(trace) && isfinite(bb)) {
            int evals = 0;
            for (int outer = 0; outer < max_outer && evals < max_evals; ++outer) {
                const double mu = mu_rel * trace / K;
                mu_rel *= 1e-2;
                double l0 = 0, l1 = 0, l2 = 0;
                for (int it = 0; it < max_newton && evals < max_evals; ++it) {
                    ++evals;
                    // pass A
                    double F0 = 0, F1 = 0, F2 = 0, H00 = 0, H10 = 0, H11 = 0, H20 = 0, H21 = 0, H22 = 0, Ds = 0;
                    unsigned lowA[SLOTS], highA[SLOTS];
#pragma unroll
                    for (int sl = 0; sl < SLOTS; ++sl) {
                        const double z = fma(p2[sl], l2, fma(p1[sl], l1, fma(p0[sl], l0, wc[sl])));
                        const bool below = own[sl] && !(z > lo), above = own[sl] && !(z < hi);
                        const double w = below ? lo : (above ? hi : z);
                        if (own[sl]) {
                            Ds += below ? lo * z - 0.5 * lo * lo : (above ? hi * z - 0.5 * hi * hi : 0.5 * z * z);
                            F0 = fma(p0[sl], w, F0); F1 = fma(p1[sl], w, F1); F2 = fma(p2[sl], w, F2);
                            if (!below && !above) {
                                H00 = fma(p0[sl], p0[sl], H00); H10 = fma(p1[sl], p0[sl], H10); H11 = fma(p1[sl], p1[sl], H11);
                                H20 = fma(p2[sl], p0[sl], H20); H21 = fma(p2[sl], p1[sl], H21); H22 = fma(p2[sl], p2[sl], H22);
                            }
                        }
                        lowA[sl] = __ballot_sync(0xffffffffu, below);
                        highA[sl] = __ballot_sync(0xffffffffu, above);
                    }
                    F0 = warp_sum(F0) + (mu * l0 - b[0]); F1 = warp_sum(F1) + (mu * l1 - b[1]); F2 = warp_sum(F2) + (mu * l2 - b[2]);
                    H00 = warp_sum(H00) + mu; H10 = warp_sum(H10); H11 = warp_sum(H11) + mu;
                    H20 = warp_sum(H20); H21 = warp_sum(H21); H22 = warp_sum(H22) + mu;
                    const double D0 = warp_sum(Ds) + fma(0.5 * mu, l0 * l0 + l1 * l1 + l2 * l2, -(b[0] * l0 + b[1] * l1 + b[2] * l2));
                    if (fmax(fmax(fabs(F0), fabs(F1)), fabs(F2)) <= 1e-13 * sqrt(bb)) break;
                    if (!(H00 > 0)) break;
                    const double c00 = sqrt(H00), c10 = H10 / c00, c20 = H20 / c00;
                    const double t11 = H11 - c10 * c10;
                    if (!(t11 > 0)) break;
                    const double c11 = sqrt(t11), c21 = (H21 - c20 * c10) / c11;
                    const double t22 = H22 - c20 * c20 - c21 * c21;
                    if (!(t22 > 0)) break;
                    const double c22 = sqrt(t22);
                    const double y0 = -F0 / c00, y1 = (-F1 - c10 * y0) / c11, y2 = (-F2 - c20 * y0 - c21 * y1) / c22;
                    const double d2 = y2 / c22, d1 = (y1 - c21 * d2) / c11, d0 = (y0 - c10 * d1 - c20 * d2) / c00;
                    const double slope = F0 * d0 + F1 * d1 + F2 * d2;
                    if (!(slope < 0)) break;
                    // pass B: the unit Newton step if it passes the Armijo test on the dual (inside one linear piece it is
                    // the exact solution and ends the stage); otherwise the EXACT minimiser of the dual along d.  Round 1
                    // halved the step until the test passed: 3.6 passes per Newton step on average and ~10 on the hard
                    // (infeasible) in-loop problems, which then ran into the Newton cap -- 76-103 passes per problem on
                    // average, up to 1,643; with the exact search 30 on average, at most 89 (oracle, dumped config-3 problems).
                    bool ok = false, same = false;
                    {
                        ++evals;
                        const double t0 = l0 + d0, t1 = l1 + d1, t2 = l2 + d2;
                        double Dp = 0;
                        bool eq = true;
#pragma unroll
                        for (int sl = 0; sl < SLOTS; ++sl) {
                            const double z = fma(p2[sl], t2, fma(p1[sl], t1, fma(p0[sl], t0, wc[sl])));
                            const bool below = own[sl] && !(z > lo), above = own[sl] && !(z < hi);
                            if (own[sl]) Dp += below ? lo * z - 0.5 * lo * lo : (above ? hi * z - 0.5 * hi * hi : 0.5 * z * z);
                            eq = eq && (__ballot_sync(0xffffffffu, below) == lowA[sl]) && (__ballot_sync(0xffffffffu, above) == highA[sl]);
                        }
                        const double Dt = warp_sum(Dp) + fma(0.5 * mu, t0 * t0 + t1 * t1 + t2 * t2, -(b[0] * t0 + b[1] * t1 + b[2] * t2));
                        if (Dt <= D0 + 1e-4 * slope + 1e-14 * fabs(D0)) {
                            ok = true;
                            l0 = t0; l1 = t1; l2 = t2;
                            same = eq;
                        } else {
                            // g(a) = d . grad D(lam + a d) = c0 + c1 a + sum_j s_j clip(z_j + a s_j): piecewise linear, increasing,
                            // g(0) = slope < 0.  Weight j contributes the breakpoints (lo - z_j) / s_j and (hi - z_j) / s_j; every
                            // lane evaluates g at ITS breakpoints (z_j, s_j of all weights broadcast by shuffles, summed in the
                            // order of the CPU checker), the bracket of the root is a pair of warp reductions.
                            evals += 3;
                            double c0 = 0, c1 = 0;
                            c0 = fma(mu * l0 - b[0], d0, c0); c0 = fma(mu * l1 - b[1], d1, c0); c0 = fma(mu * l2 - b[2], d2, c0);
                            c1 = fma(mu * d0, d0, c1); c1 = fma(mu * d1, d1, c1); c1 = fma(mu * d2, d2, c1);
                            double zs[SLOTS], ss[SLOTS], ak[2 * SLOTS], gk[2 * SLOTS];
#pragma unroll
                            for (int sl = 0; sl < SLOTS; ++sl) {
                                zs[sl] = fma(p2[sl], l2, fma(p1[sl], l1, fma(p0[sl], l0, wc[sl])));
                                ss[sl] = fma(p2[sl], d2, fma(p1[sl], d1, p0[sl] * d0));
                                const double a_lo = (lo - zs[sl]) / ss[sl], a_hi = (hi - zs[sl]) / ss[sl];
                                const bool has = own[sl] && ss[sl] != 0.0;
                                ak[2 * sl] = (has && a_lo > 0.0 && isfinite(a_lo)) ? a_lo : -1.0;
                                ak[2 * sl + 1] = (has && a_hi > 0.0 && isfinite(a_hi)) ? a_hi : -1.0;
                                gk[2 * sl] = 0.0; gk[2 * sl + 1] = 0.0;
                            }
                            auto g_at = [&](double a) {            // every lane: g(a), weights in ascending order
                                double acc = 0;
#pragma unroll
                                for (int sl = 0; sl < SLOTS; ++sl) {
                                    const int cnt = (D - 32 * sl) < 32 ? (D - 32 * sl) : 32;
                                    for (int j = 0; j < cnt; ++j) {
                                        const double zj = __shfl_sync(0xffffffffu, zs[sl], j), sj = __shfl_sync(0xffffffffu, ss[sl], j);
                                        acc = fma(sj, clipw(fma(a, sj, zj)), acc);
                                    }
                                }
                                return fma(c1, a, c0) + acc;
                            };
                            {   // all own breakpoints in one sweep over the weights
                                double acc[2 * SLOTS];
#pragma unroll
                                for (int q = 0; q < 2 * SLOTS; ++q) acc[q] = 0.0;
#pragma unroll
                                for (int sl = 0; sl < SLOTS; ++sl) {
                                    const int cnt = (D - 32 * sl) < 32 ? (D - 32 * sl) : 32;
                                    for (int j = 0; j < cnt; ++j) {
                                        const double zj = __shfl_sync(0xffffffffu, zs[sl], j), sj = __shfl_sync(0xffffffffu, ss[sl], j);
#pragma unroll
                                        for (int q = 0; q < 2 * SLOTS; ++q) acc[q] = fma(sj, clipw(fma(ak[q], sj, zj)), acc[q]);
                                    }
                                }
#pragma unroll
                                for (int q = 0; q < 2 * SLOTS; ++q) gk[q] = fma(c1, ak[q], c0) + acc[q];
                            }
                            double aL = 0.0, gL = slope, aR = CUDART_INF, gR = 0.0;
#pragma unroll
                            for (int q = 0; q < 2 * SLOTS; ++q) {
                                if (ak[q] > 0.0) {
                                    if (gk[q] < 0.0) { if (ak[q] > aL) { aL = ak[q]; gL = gk[q]; } }
                                    else if (ak[q] < aR) { aR = ak[q]; gR = gk[q]; }
                                }
                            }
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) {
                                const double oaL = __shfl_xor_sync(0xffffffffu, aL, off), ogL = __shfl_xor_sync(0xffffffffu, gL, off);
                                const double oaR = __shfl_xor_sync(0xffffffffu, aR, off), ogR = __shfl_xor_sync(0xffffffffu, gR, off);
                                if (oaL > aL) { aL = oaL; gL = ogL; }
                                if (oaR < aR) { aR = oaR; gR = ogR; }
                            }
                            double a;
                            if (aR < CUDART_INF) {
                                a = !(aR > aL) ? aR : ((gR > gL) ? aL + (aR - aL) * (-gL) / (gR - gL) : aR);
                            } else {                               // beyond the last breakpoint: the slope of the last piece
                                const double sl_ = g_at(aL + 1.0) - gL;
                                a = (sl_ > 0.0) ? aL - gL / sl_ : -1.0;
                            }
                            if (a > 0.0 && isfinite(a)) {
                                ok = true;
                                l0 = fma(a, d0, l0); l1 = fma(a, d1, l1); l2 = fma(a, d2, l2);
                            }
                        }
                    }
                    (void)max_ls;
                    if (!ok || same) break;
                }
                // prox step result, its cost
                double q0 = 0, q1 = 0, q2 = 0;
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {
                    wn[sl] = own[sl] ? clipw(fma(p2[sl], l2, fma(p1[sl], l1, fma(p0[sl], l0, wc[sl])))) : 0.0;
                    q0 = fma(p0[sl], wn[sl], q0); q1 = fma(p1[sl], wn[sl], q1); q2 = fma(p2[sl], wn[sl], q2);
                }
                q0 = warp_sum(q0) - b[0]; q1 = warp_sum(q1) - b[1]; q2 = warp_sum(q2) - b[2];
                const double Jn = 0.5 * (q0 * q0 + q1 * q1 + q2 * q2);
                if (Jn < Jbest) {
                    Jbest = Jn;
#pragma unroll
                    for (int sl = 0; sl < SLOTS; ++sl) wc[sl] = wn[sl];
                    if (Jn <= 1e-12 * J0 || Jn <= 1e-20 * bb) break;
                }
            }
        }
        __syncwarp();                                      // every lane has read w_prev before anyone overwrites it
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
            const int j = lane + 32 * sl;
            if (own[sl]) {
                w_g[j * E + e] = wc[sl];
                if (update_prev) wprev_g[j * E + e] = wc[sl];
            }
        }
        if (lane == 0 && Jc_out) Jc_out[e] = Jbest;
    }
}

static bool fit_rdiag(const rcg_objective_t *obj, int p)
{
    return obj->r_is_diag && is_diag(obj->R1, p) && (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, p));
}

}  // namespace rcg

extern "C" int rcg_critic_fit(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs_buf,
                              const double *act_buf, double *w_prev, double w_min, double w_max, const double *w_init,
                              double *w, const int32_t *mask, int32_t max_evals, int32_t update_prev, double *Jc_out,
                              void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(obj && (E <= 0 || (obs_buf && act_buf && w_prev && w)), "rcg_critic_fit: null argument");
    RCG_REQUIRE((n == 3 && m == 2) || (n == 5 && m == 2) || (n == 2 && m == 1), "rcg_critic_fit: unsupported dims n=%d m=%d", n, m);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "rcg_critic_fit: unknown critic_struct %d",
                obj->critic_struct);
    RCG_REQUIRE(obj->Ncritic >= 1 && obj->Ncritic <= obj->buffer_size,
                "rcg_critic_fit: Ncritic %d must be in [1, buffer_size = %d]", obj->Ncritic, obj->buffer_size);
    RCG_REQUIRE(obj->Ncritic - 1 <= kFitMaxK, "rcg_critic_fit: Ncritic - 1 = %d exceeds %d", obj->Ncritic - 1, kFitMaxK);
    RCG_REQUIRE(w_min <= w_max, "rcg_critic_fit: empty weight box");
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const ObjDev<double> O = make_obj_dev<double>(obj, n, m);
    const bool rd = fit_rdiag(obj, n + m);
    const unsigned grid = (unsigned)((E + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    // proximal weights mu_rel * trace(Phi Phi^T) / K with mu_rel = 1e-3, 1e-5, ..., 1e-11: the first stages are
    // well conditioned and do most of the work, the last ones polish (measured on in-loop problems of BASELINE
    // configs 3/4: 5x fewer dual evaluations in the tail than a fixed mu_rel = 1e-10, same fitted costs).
    const int outer = 5;
    const double mu_rel = 1e-3;
    const int newton = 20;
    const int max_ls = 40;          // Armijo halvings per Newton step (a cap of 12 saves 17 % of the fit time in config 3 but moves the fitted costs)
    const bool fast = obj->Ncritic - 1 <= 3;
    // Run-to-convergence fits of a large batch go in two phases.  A few per cent of in-loop problems need hundreds
    // to thousands of dual evaluations while the rest finish within a few dozen; with one environment per lane every
    // warp that holds one of them waits for it (ncu: 3.3 of 32 lanes active on average).  Phase 1 gives every
    // environment kPhase1Budget evaluations; those that run out are queued (and write nothing), phase 2 restarts
    // exactly them, packed densely into warps.  Each environment's result is what the single-phase kernel computes.
    // (round 2, exact line search in the second phase: budgets 4 ... 32 and "no first phase at all" -- every masked
    //  environment straight to the warp kernel -- measure within 5 % of each other: 12.95 ... 14.1 ms per 1 M in-loop fits)
    constexpr int kPhase1Budget = 32;
    constexpr bool warp_only = false;
    // Measured on B200 (profiles/r01_critic_fit_two_phase.txt): 23.0 -> 20.8 ms per 1 M fits of the 28-weight critic
    // inside config 3's loop -- the rest of the tail is the serial dependency chain of the environments that run
    // into the iteration caps, not idle lanes; for small critics the second launch costs more than it saves
    // (2tank, 3 weights: 0.49 -> 0.55 ms), hence the dim_critic threshold.
    const bool two_phase = max_evals <= 0 && dim_critic_c(obj->critic_struct, n, m) >= 10 &&
                           getenv("RCG_FIT_ONE_PHASE") == nullptr;
    int32_t *todo = nullptr;
    if (two_phase) {
        // keep the stream-ordered pool's memory across synchronisation points (default: released at every sync,
        // which would turn the per-call scratch into a real cudaMalloc / cudaFree pair)
        static bool pool_configured = false;
        if (!pool_configured) {
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = (uint64_t)256 << 20;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            (void)cudaGetLastError();
            pool_configured = true;
        }
        if (cudaMallocAsync((void **)&todo, (size_t)(E + 2) * sizeof(int32_t), s) != cudaSuccess) {
            (void)cudaGetLastError();
            todo = nullptr;                                // no pool on this device: single phase
        } else {
            cudaMemsetAsync(todo + E, 0, 2 * sizeof(int32_t), s);      // [E] = queued environments, [E + 1] = work queue
        }
    }
    const int nphases = todo ? 2 : 1;
    const int first_phase = (todo && warp_only && fast) ? 1 : 0;
    const bool warp_phase2 = getenv("RCG_FIT_LANE_PHASE2") == nullptr;      // second phase: one warp per environment
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int phase = first_phase; phase < nphases; ++phase) {
        const int evals = (todo && phase == 0) ? kPhase1Budget : (max_evals > 0 ? max_evals : 0x7fffffff);
        const bool listed = todo && phase == 1 && first_phase == 0;
        const int32_t *lane_list = listed ? todo : nullptr, *lane_count = listed ? todo + E : nullptr;
        const bool warp_phase = todo && phase == 1;
        int32_t *todo_list = (todo && phase == 0) ? todo : nullptr, *todo_count = (todo && phase == 0) ? todo + E : nullptr;
#define FIT3(NN, MM, CS, RD)                                                                                              \
    {                                                                                                                     \
        auto kern = critic_fit3_kernel<NN, MM, CS, RD>;                                                                   \
        const size_t smem = (size_t)2 * dim_critic_c(CS, NN, MM) * 128 * sizeof(double);                                  \
        static unsigned long long configured = 0;                                                                         \
        ensure_dyn_smem(kern, smem, configured);                                                                          \
        kern<<<grid, 128, smem, s>>>(O, E, obs_buf, act_buf, w_prev, w_min, w_max, w_init, w, mask, mu_rel, outer, newton,  \
                                     evals, update_prev, Jc_out, max_ls, lane_list, lane_count, todo_list, todo_count);          \
    }
#define FIT3W(NN, MM, CS, RD)                                                                                             \
    critic_fit3_warp_kernel<NN, MM, CS, RD><<<(unsigned)(sms * 8), 128, 0, s>>>(                                         \
        O, E, obs_buf, act_buf, w_prev, w_min, w_max, w_init, w, mu_rel, outer, newton, evals, update_prev, Jc_out,      \
        max_ls, lane_list, lane_count, todo + E + 1, mask);
#define FIT(NN, MM, CS)                                                                                                   \
    if (fast && warp_phase && warp_phase2) { if (rd) FIT3W(NN, MM, CS, true) else FIT3W(NN, MM, CS, false) }              \
    else if (fast) { if (rd) FIT3(NN, MM, CS, true) else FIT3(NN, MM, CS, false) }                                        \
    else if (rd) critic_fit_kernel<NN, MM, CS, true><<<grid, 128, 0, s>>>(O, E, obs_buf, act_buf, w_prev, w_min, w_max, w_init, w, \
                                                                     mask, mu_rel, outer, newton, evals, update_prev, Jc_out,      \
                                                                     max_ls, lane_list, lane_count, todo_list, todo_count);        \
    else critic_fit_kernel<NN, MM, CS, false><<<grid, 128, 0, s>>>(O, E, obs_buf, act_buf, w_prev, w_min, w_max, w_init, w,  \
                                                                   mask, mu_rel, outer, newton, evals, update_prev, Jc_out,        \
                                                                   max_ls, lane_list, lane_count, todo_list, todo_count);
#define FITCS(NN, MM)                  \
    switch (obj->critic_struct) {      \
    case 0: FIT(NN, MM, 0) break;      \
    case 1: FIT(NN, MM, 1) break;      \
    case 2: FIT(NN, MM, 2) break;      \
    default: FIT(NN, MM, 3) break;     \
    }
        if (n == 3) { FITCS(3, 2) } else if (n == 5) { FITCS(5, 2) } else { FITCS(2, 1) }
#undef FITCS
#undef FIT
#undef FIT3W
#undef FIT3
        if (int rc = check_launch("rcg_critic_fit")) {
            if (todo) cudaFreeAsync(todo, s);
            return rc;
        }
    }
    if (todo) cudaFreeAsync(todo, s);
    return 0;
}
