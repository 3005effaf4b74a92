// critic_fit.cu -- batched CtrlOptPred._critic_optimizer (rcognita/controllers.py:1248-1271): the minimiser
// of _critic_cost (:1216-1245) over w in [Wmin, Wmax], started from w_critic_init, one environment per thread.
//
// _critic_cost is linear least squares in w:  J_c(w) = 1/2 * sum_r (phi_r . w - b_r)^2  with, for the
// buffer rows k = Ncritic-1 .. 1,  phi_r = phi(o[k-1], a[k-1])  and  b_r = gamma * Q(o[k], a[k]; w_prev) +
// stage_obj(o[k-1], a[k-1]).  There are K = Ncritic-1 (3 in every preset) rows against 3..35 unknowns, so the
// minimiser is not unique; the reference takes whatever SLSQP (maxiter 200, tol 1e-7) returns from
// w_critic_init.  Here: proximal-point iterations  w <- argmin J_c(w) + mu/2 |w - w_prev_iter|^2  within the
// box, each solved exactly in the K-dimensional dual by a semismooth Newton method -- the prox solution is
// w = clip(w0 + Phi^T lam) with  mu*lam + Phi*clip(w0 + Phi^T lam) - b = 0  (a piecewise-linear monotone
// equation; Newton with the Gram matrix of the un-clipped columns terminates when the clip pattern stops
// changing).  The iterate with the smallest J_c is returned, so the result is never worse than w_critic_init.
// Parity is on the fitted COST against the reference's SLSQP result (tests/golden/critic_fit.json), never on
// the weights (hard part 5).
#include "rcg_host.h"

namespace rcg {

constexpr int kFitMaxK = 15;            // Ncritic - 1 <= 15

template <typename T>
struct GlobalWF {
    const T *w;
    int64_t stride, idx;
    __device__ __forceinline__ T operator()(int i) const { return w[i * stride + idx]; }
};

template <int N, int M, int CS, bool RDIAG>
__global__ void __launch_bounds__(128)
critic_fit_kernel(const __grid_constant__ ObjDev<double> O, int64_t E, const double *__restrict__ obs_buf,
                  const double *__restrict__ act_buf, const double *__restrict__ wprev_g, double lo, double hi,
                  double *__restrict__ w_g, const int32_t *__restrict__ mask, double mu_rel, int max_outer,
                  int max_newton, double *__restrict__ Jc_out)
{
    constexpr int D = dim_critic_c(CS, N, M);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (mask && mask[e] == 0) return;
    const int K = O.Ncritic - 1;

    double Phi[kFitMaxK * D], b[kFitMaxK], lam[kFitMaxK], lt[kFitMaxK], F[kFitMaxK], dl[kFitMaxK];
    double H[kFitMaxK * kFitMaxK];
    double w0[D], wb[D];
    signed char st[D];

    // rows of the least-squares problem (controllers.py:1230-1242)
    const GlobalWF<double> wp{wprev_g, E, e};
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
    for (int j = 0; j < D; ++j) { w0[j] = clipw(w_g[j * E + e]); wb[j] = w0[j]; }
    double Jbest = cost(w0);

    if (K >= 1 && trace > 0 && isfinite(trace) && isfinite(bb)) {
        const double mu = mu_rel * trace / K;
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
        for (int outer = 0; outer < max_outer; ++outer) {
            for (int r = 0; r < K; ++r) lam[r] = 0;
            for (int it = 0; it < max_newton; ++it) {
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
                for (int ls = 0; ls < 40; ++ls) {                  // Armijo backtracking on the dual
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
            for (int j = 0; j < D; ++j) w0[j] = wn[j];
            if (!(Jn < Jbest)) break;
            const bool improved = Jn < Jbest * (1 - 1e-3);
            Jbest = Jn;
            for (int j = 0; j < D; ++j) wb[j] = wn[j];
            if (!improved || Jn <= 1e-26 * bb) break;
        }
    }
    for (int j = 0; j < D; ++j) w_g[j * E + e] = wb[j];
    if (Jc_out) Jc_out[e] = Jbest;
}

static bool fit_rdiag(const rcg_objective_t *obj, int p)
{
    return obj->r_is_diag && is_diag(obj->R1, p) && (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, p));
}

}  // namespace rcg

extern "C" int rcg_critic_fit(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs_buf,
                              const double *act_buf, const double *w_prev, double w_min, double w_max, double *w,
                              const int32_t *mask, int32_t max_outer, double *Jc_out, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(obj && obs_buf && act_buf && w_prev && w, "rcg_critic_fit: null argument");
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
    const int outer = max_outer > 0 ? max_outer : 8;
    const double mu_rel = 1e-10;
    const int newton = 20;
#define FIT(NN, MM, CS)                                                                                                   \
    if (rd) critic_fit_kernel<NN, MM, CS, true><<<grid, 128, 0, s>>>(O, E, obs_buf, act_buf, w_prev, w_min, w_max, w, mask, \
                                                                     mu_rel, outer, newton, Jc_out);                       \
    else critic_fit_kernel<NN, MM, CS, false><<<grid, 128, 0, s>>>(O, E, obs_buf, act_buf, w_prev, w_min, w_max, w, mask,  \
                                                                   mu_rel, outer, newton, Jc_out);
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
    return check_launch("rcg_critic_fit");
}
