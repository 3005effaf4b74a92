// rcg_device.cuh -- device-side building blocks shared by the sm_100a kernels:
// the three environments' dynamics, the stage objective and the critic features.
// Every function cites the reference line it reproduces (paths relative to the rcognita
// v0.1.2 tree).  T is the arithmetic type of the path (double, or float for the _f32 twins).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rcg.h"

namespace rcg {

template <int SYS> struct SysDim;
template <> struct SysDim<RCG_SYS_3WROBOT_NI> { static constexpr int n = 3, m = 2; };
template <> struct SysDim<RCG_SYS_3WROBOT>    { static constexpr int n = 5, m = 2; };
template <> struct SysDim<RCG_SYS_2TANK>      { static constexpr int n = 2, m = 1; };

// Kernel-parameter copies of the host descriptors (passed by value -> constant bank).
template <typename T>
struct SysDev {
    T pars[8];
    T lo[RCG_MAX_M], hi[RCG_MAX_M];
    int has_bnds;
};

template <typename T>
struct ObjDev {
    int stage_struct, has_target, Nactor, Ncritic, buffer_size;
    T gamma, pred_step_size;
    T gamma_pow[RCG_MAX_NACTOR];
    T R1[RCG_MAX_P * RCG_MAX_P];
    T R2[RCG_MAX_P * RCG_MAX_P];
    T target[RCG_MAX_N];      // zeros when has_target == 0 (x - 0 == x exactly)
};

// ---- deterministic elementary functions (fp64 path) -------------------------------------------
// sin/cos and err ** -0.2 decide, through the adaptive step size, the last bits of the solver time
// t and with them the step on which the controller samples.  libm results differ
// in the last bit between implementations (CUDA's sincos: <= 2 ulp; numpy's: CPU dependent), so the
// fp64 path uses fully specified functions built from IEEE +,-,*,fma only -- the same operation
// sequence the CPU checker uses, so that both agree bit for bit on every lane:
//  - det_sincos: q = rint(x * 2/pi); r = x - q*pi/2 in double-double with pi/2 split into three
//    doubles; then the fdlibm / FreeBSD msun k_sin / k_cos kernels on (rh, rl)  (< 1 ulp).
//  - det_pow_m02: x ** -0.2 = libm pow() + one correction from the double-double residual
//    y^5 * x - 1 and the term for 0.2 != 1/5: correctly rounded (up to astronomically rare cases).
// Translation units that need bit-reproducibility (rk45.cu) are compiled with -fmad=false, so the
// plain * and + below are never contracted; the fma() calls are explicit.
// Coefficients in constant memory: as literals every use costs two UMOVs to materialise the 64-bit immediate
// (the RK45 kernels inline det_sincos seven times per step attempt).
struct DetSincosC {
    double two_over_pi, P1, P2, P3;
    double S1, S2, S3, S4, S5, S6;
    double C1, C2, C3, C4, C5, C6;
};
static __constant__ DetSincosC kDSC = {
    6.36619772367581382433e-01, 1.5707963267948966e+00, 6.123233995736766e-17, -1.4973849048591698e-33,
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,
    -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,
    2.08757232129817482790e-09, -1.13596475577881948265e-11};

static __device__ __noinline__ double2 sincos_libm(double x)      // cold path, kept out of line
{
    double2 r;
    sincos(x, &r.x, &r.y);
    return r;
}

__device__ __forceinline__ void det_sincos(double x, double *sn, double *cs)
{
    if (!(fabs(x) <= 1.0e5)) {                  // huge, inf, NaN: outside the regime of the path
        const double2 r = sincos_libm(x);
        *sn = r.x;
        *cs = r.y;
        return;
    }
    const double q = rint(__dmul_rn(x, kDSC.two_over_pi));
    const double P1 = kDSC.P1, P2 = kDSC.P2, P3 = kDSC.P3;
    const double ph = __dmul_rn(q, P1), pl = fma(q, P1, -ph);
    const double r = __dadd_rn(x, -ph);
    double t = fma(-q, P2, -pl);
    t = fma(-q, P3, t);
    const double rh = __dadd_rn(r, t);
    const double bb = __dadd_rn(rh, -r);
    const double rl = __dadd_rn(__dadd_rn(r, -__dadd_rn(rh, -bb)), __dadd_rn(t, -bb));     // two-sum
    const double S1 = kDSC.S1, S2 = kDSC.S2, S3 = kDSC.S3, S4 = kDSC.S4, S5 = kDSC.S5, S6 = kDSC.S6;
    const double C1 = kDSC.C1, C2 = kDSC.C2, C3 = kDSC.C3, C4 = kDSC.C4, C5 = kDSC.C5, C6 = kDSC.C6;
    const double z = __dmul_rn(rh, rh), w = __dmul_rn(z, z), v = __dmul_rn(z, rh);
    // every product and sum rounded separately (intrinsics: immune to -fmad contraction)
    const double ps = __dadd_rn(__dadd_rn(S2, __dmul_rn(z, __dadd_rn(S3, __dmul_rn(z, S4)))),
                                __dmul_rn(__dmul_rn(z, w), __dadd_rn(S5, __dmul_rn(z, S6))));
    const double ks = __dadd_rn(rh, -__dadd_rn(__dadd_rn(__dmul_rn(z, __dadd_rn(__dmul_rn(0.5, rl), -__dmul_rn(v, ps))), -rl),
                                               -__dmul_rn(v, S1)));
    const double pc = __dadd_rn(__dmul_rn(z, __dadd_rn(C1, __dmul_rn(z, __dadd_rn(C2, __dmul_rn(z, C3))))),
                                __dmul_rn(__dmul_rn(w, w), __dadd_rn(C4, __dmul_rn(z, __dadd_rn(C5, __dmul_rn(z, C6))))));
    const double hz = __dmul_rn(0.5, z), w1 = __dadd_rn(1.0, -hz);
    const double kc = __dadd_rn(w1, __dadd_rn(__dadd_rn(__dadd_rn(1.0, -w1), -hz),
                                              __dadd_rn(__dmul_rn(z, pc), -__dmul_rn(rh, rl))));
    const int n = (int)((long long)q & 3);
    *sn = (n == 0) ? ks : (n == 1) ? kc : (n == 2) ? -ks : -kc;
    *cs = (n == 0) ? kc : (n == 1) ? -ks : (n == 2) ? -kc : ks;
}

__device__ __forceinline__ double det_pow_m02(double x)
{
    const double y = pow(x, -0.2);
    if (!(x > 0.0 && x < (double)INFINITY)) return y;
    const double ah = __dmul_rn(y, y), al = fma(y, y, -ah);                                                  // y^2
    const double bh = __dmul_rn(ah, ah), bl = __dadd_rn(fma(ah, ah, -bh), __dmul_rn(2.0, __dmul_rn(ah, al)));   // y^4
    const double ch = __dmul_rn(bh, y), cl = __dadd_rn(fma(bh, y, -ch), __dmul_rn(bl, y));                   // y^5
    const double dh = __dmul_rn(ch, x), dl = __dadd_rn(fma(ch, x, -dh), __dmul_rn(cl, x));                   // y^5 x ~ 1
    const double g = __dadd_rn(__dadd_rn(dh, -1.0), dl);
    // the exponent is the double 0.2 = 1/5 + 1.1102230246251565e-17
    return __dadd_rn(y, __dmul_rn(-y, __dadd_rn(__dmul_rn(0.2, g), __dmul_rn(1.1102230246251565e-17, log(x)))));
}
__device__ __forceinline__ float det_pow_m02(float x) { return powf(x, -0.2f); }

__device__ __forceinline__ void sincos_t(double x, double *s, double *c) { det_sincos(x, s, c); }
__device__ __forceinline__ void sincos_t(float x, float *s, float *c) { sincosf(x, s, c); }

// System._state_dyn, is_disturb = 0:
//   Sys3WRobotNI rcognita/systems.py:370-382, Sys3WRobot :308-323, Sys2Tank :412-419.
// Operation order follows the Python expressions literally.
template <typename T, int SYS>
__device__ __forceinline__ void state_dyn(const SysDev<T> &S, const T *x, const T *a, T *d)
{
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {
        T s, c;
        sincos_t(x[2], &s, &c);
        d[0] = a[0] * c;
        d[1] = a[0] * s;
        d[2] = a[1];
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {
        T s, c;
        sincos_t(x[2], &s, &c);
        d[0] = x[3] * c;
        d[1] = x[3] * s;
        d[2] = x[4];
        d[3] = (T(1) / S.pars[0]) * a[0];          // 1/m * F
        d[4] = (T(1) / S.pars[1]) * a[1];          // 1/I * M
    } else {
        const T tau1 = S.pars[0], tau2 = S.pars[1], K1 = S.pars[2], K2 = S.pars[3], K3 = S.pars[4];
        d[0] = (T(1) / tau1) * (-x[0] + K1 * a[0]);
        d[1] = (T(1) / tau2) * ((-x[1] + K2 * x[0]) + K3 * (x[1] * x[1]));
    }
}

// ---- disturbance lanes (is_disturb = 1): rcognita/systems.py:228-231, :247-248, :316-318, :373-376, :325-345, :384-394
// dim_disturb is 2 for the two robots and 1 for the two-tank system (the presets' values).
template <int SYS> struct DistDim { static constexpr int nd = (SYS == RCG_SYS_2TANK) ? 1 : 2; };

struct DistDev {
    double sigma[2], mu[2], tau[2];          // pars_disturb = [sigma_disturb, mu_disturb, tau_disturb]
    unsigned long long seed;                 // Philox key
    long long env_offset;                    // global index of lane 0 (sharding does not change an environment's stream)
};

// _state_dyn(t, state, action, disturb) with a disturbance: Sys3WRobotNI adds disturb[0] to BOTH position rates and
// disturb[1] to the heading rate (systems.py:373-376, literally), Sys3WRobot adds the disturbance to force and moment
// before the division (:316-318), Sys2Tank ignores it (:412-419).
template <int SYS>
__device__ __forceinline__ void state_dyn_disturbed(const SysDev<double> &S, const double *x, const double *a, const double *q, double *d)
{
    if constexpr (SYS == RCG_SYS_3WROBOT_NI) {
        double s, c;
        det_sincos(x[2], &s, &c);
        d[0] = a[0] * c + q[0];
        d[1] = a[0] * s + q[0];
        d[2] = a[1] + q[1];
    } else if constexpr (SYS == RCG_SYS_3WROBOT) {
        double s, c;
        det_sincos(x[2], &s, &c);
        d[0] = x[3] * c;
        d[1] = x[3] * s;
        d[2] = x[4];
        d[3] = (1.0 / S.pars[0]) * (a[0] + q[0]);
        d[4] = (1.0 / S.pars[1]) * (a[1] + q[1]);
    } else {
        state_dyn<double, SYS>(S, x, a, d);
    }
}

// _disturb_dyn GIVEN the draws z[k] = randn(): Ddisturb[k] = -tau[k] * (disturb[k] + sigma[k] * (z[k] + mu[k]))
// (systems.py:341-343, :390-392; note the multiplication by tau).  Sys2Tank: zeros (:421-424).
template <int SYS>
__device__ __forceinline__ void disturb_dyn(const DistDev &D, const double *q, const double *z, double *dq)
{
    if constexpr (SYS == RCG_SYS_2TANK) {
        dq[0] = 0.0;
    } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) dq[k] = -D.tau[k] * (q[k] + D.sigma[k] * (z[k] + D.mu[k]));
    }
}

// Natural logarithm from IEEE operations only (the fdlibm / musl algorithm: x = 2^k (1 + f), s = f / (2 + f), a degree-14
// even polynomial in s), for normal positive x: the same operation sequence as the CPU checker, so the normal draws below
// agree bit for bit between the two.
__device__ __forceinline__ double det_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    unsigned int hx = (unsigned int)(bits >> 32);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int k = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    bits = ((unsigned long long)hx << 32) | (bits & 0xffffffffull);
    const double f = __dadd_rn(__longlong_as_double((long long)bits), -1.0);
    const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s), w = __dmul_rn(z, z);
    const double t1 = __dmul_rn(w, __dadd_rn(Lg2, __dmul_rn(w, __dadd_rn(Lg4, __dmul_rn(w, Lg6)))));
    const double t2 = __dmul_rn(z, __dadd_rn(Lg1, __dmul_rn(w, __dadd_rn(Lg3, __dmul_rn(w, __dadd_rn(Lg5, __dmul_rn(w, Lg7)))))));
    const double R = __dadd_rn(t2, t1);
    const double dk = (double)k;
    // s*(hfsq+R) + dk*ln2_lo - hfsq + f + dk*ln2_hi, left to right
    double r = __dadd_rn(__dmul_rn(s, __dadd_rn(hfsq, R)), __dmul_rn(dk, ln2_lo));
    r = __dadd_rn(r, -hfsq);
    r = __dadd_rn(r, f);
    return __dadd_rn(r, __dmul_rn(dk, ln2_hi));
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> four 32-bit words.
__device__ __forceinline__ void philox4x32_10(unsigned int c[4], unsigned int k0, unsigned int k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
        const unsigned int lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// The two standard-normal draws of RHS call number `call` of global environment `env` (what the reference takes from
// numpy's global randn(), systems.py:343 / :392 -- irreproducible there by construction; here a counter-based stream):
// Philox block (call, env) under key `seed` -> two uniforms in (0, 1) with 64 random bits each -> Box-Muller with the
// specified log / sincos above.
__device__ __forceinline__ void det_normal2(unsigned long long seed, unsigned long long env, unsigned int call, double *z)
{
    unsigned int c[4] = {call, 0u, (unsigned int)env, (unsigned int)(env >> 32)};
    philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
    // (x + 0.5) / 2^64 with x the 64-bit word: top 53 bits kept, the half keeps u away from 0
    const double u1 = __dmul_rn(__dadd_rn((double)((((unsigned long long)c[0] << 32) | c[1]) >> 11), 0.5), 1.1102230246251565e-16);
    const double u2 = __dmul_rn(__dadd_rn((double)((((unsigned long long)c[2] << 32) | c[3]) >> 11), 0.5), 1.1102230246251565e-16);
    const double r = sqrt(__dmul_rn(-2.0, det_log(u1)));
    double sn, cs;
    det_sincos(__dmul_rn(6.283185307179586, u2), &sn, &cs);
    z[0] = __dmul_rn(r, cs);
    z[1] = __dmul_rn(r, sn);
}

// The clipping of System.closed_loop_rhs (rcognita/systems.py:241-243): np.clip(a, lo, hi).
template <typename T, int M>
__device__ __forceinline__ void clip_action(const SysDev<T> &S, T *a)
{
    if (S.has_bnds) {
#pragma unroll
        for (int k = 0; k < M; ++k) {
            T v = a[k];
            v = (v < S.lo[k]) ? S.lo[k] : v;        // NaN stays NaN, like np.clip
            v = (v > S.hi[k]) ? S.hi[k] : v;
            a[k] = v;
        }
    }
}

// x @ R @ x evaluated left to right like numpy: v = x @ R, then v @ x.
template <typename T, int P, bool RDIAG>
__device__ __forceinline__ T quad_form(const T *x, const T *R)
{
    T out = T(0);
    if constexpr (RDIAG) {
#pragma unroll
        for (int j = 0; j < P; ++j) out += (x[j] * R[j * P + j]) * x[j];
    } else {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            T acc = T(0);
#pragma unroll
            for (int i = 0; i < P; ++i) acc += x[i] * R[i * P + j];
            out += acc * x[j];
        }
    }
    return out;
}

// CtrlOptPred.stage_obj (rcognita/controllers.py:1063-1084), a.k.a. rcost.
template <typename T, int N, int M, bool RDIAG, bool QUAD_ONLY = false>
__device__ __forceinline__ T stage_obj(const ObjDev<T> &O, const T *obs, const T *act)
{
    constexpr int P = N + M;
    T chi[P];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = obs[i] - O.target[i];
#pragma unroll
    for (int j = 0; j < M; ++j) chi[N + j] = act[j];
    if (QUAD_ONLY || O.stage_struct == RCG_STAGE_QUADRATIC) {
        return quad_form<T, P, RDIAG>(chi, O.R1);
    } else {
        T chi2[P];
#pragma unroll
        for (int i = 0; i < P; ++i) chi2[i] = chi[i] * chi[i];
        return quad_form<T, P, RDIAG>(chi2, O.R2) + quad_form<T, P, RDIAG>(chi, O.R1);
    }
}

// CtrlOptPred._critic (rcognita/controllers.py:1192-1214): w . phi(obs, act).  Feature order:
// uptria2vec is row-major i <= j (rcognita/utilities.py:81-96); np.kron(obs, act)[i*M + j] =
// obs_i * act_j; quad-mix uses the RAW observation (controllers.py:1212).  `w(i)` is any
// accessor of the i-th weight (shared memory, registers or a strided global load).
template <typename T, int N, int M, int CS, class WAcc>
__device__ __forceinline__ T critic(const ObjDev<T> &O, const T *obs, const T *act, WAcc w)
{
    constexpr int P = N + M;
    T chi[P];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = obs[i] - O.target[i];
#pragma unroll
    for (int j = 0; j < M; ++j) chi[N + j] = act[j];
    T q = T(0);
    int k = 0;
    if constexpr (CS == RCG_CRITIC_QUAD_LIN || CS == RCG_CRITIC_QUADRATIC) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = i; j < P; ++j) q += w(k++) * (chi[i] * chi[j]);
        if constexpr (CS == RCG_CRITIC_QUAD_LIN) {
#pragma unroll
            for (int i = 0; i < P; ++i) q += w(k++) * chi[i];
        }
    } else if constexpr (CS == RCG_CRITIC_QUAD_NOMIX) {
#pragma unroll
        for (int i = 0; i < P; ++i) q += w(k++) * (chi[i] * chi[i]);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) q += w(k++) * (obs[i] * obs[i]);
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) q += w(k++) * (obs[i] * act[j]);
#pragma unroll
        for (int j = 0; j < M; ++j) q += w(k++) * (act[j] * act[j]);
    }
    return q;
}

// The feature vector phi(obs, act) of CtrlOptPred._critic (same order as critic() above), written out
// explicitly: _critic_cost is linear least squares in w with rows phi (used by the critic fit).
template <typename T, int N, int M, int CS>
__device__ __forceinline__ void critic_phi(const ObjDev<T> &O, const T *obs, const T *act, T *phi)
{
    constexpr int P = N + M;
    T chi[P];
#pragma unroll
    for (int i = 0; i < N; ++i) chi[i] = obs[i] - O.target[i];
#pragma unroll
    for (int j = 0; j < M; ++j) chi[N + j] = act[j];
    int k = 0;
    if constexpr (CS == RCG_CRITIC_QUAD_LIN || CS == RCG_CRITIC_QUADRATIC) {
#pragma unroll
        for (int i = 0; i < P; ++i)
#pragma unroll
            for (int j = i; j < P; ++j) phi[k++] = chi[i] * chi[j];
        if constexpr (CS == RCG_CRITIC_QUAD_LIN) {
#pragma unroll
            for (int i = 0; i < P; ++i) phi[k++] = chi[i];
        }
    } else if constexpr (CS == RCG_CRITIC_QUAD_NOMIX) {
#pragma unroll
        for (int i = 0; i < P; ++i) phi[k++] = chi[i] * chi[i];
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) phi[k++] = obs[i] * obs[i];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) phi[k++] = obs[i] * act[j];
#pragma unroll
        for (int j = 0; j < M; ++j) phi[k++] = act[j] * act[j];
    }
}

__host__ __device__ constexpr int dim_critic_c(int cs, int n, int m)
{
    const int p = n + m;
    return cs == RCG_CRITIC_QUAD_LIN ? (p + 1) * p / 2 + p
         : cs == RCG_CRITIC_QUADRATIC ? (p + 1) * p / 2
         : cs == RCG_CRITIC_QUAD_NOMIX ? p
         : n + n * m + m;
}

// (J, index) ordering of np.argmin: NaN is minimal, ties go to the lower index.
template <typename T>
__device__ __forceinline__ bool argmin_better(T Ja, int ia, T Jb, int ib)
{
    const bool na = Ja != Ja, nb = Jb != Jb;
    if (na || nb) return (na && nb) ? (ia < ib) : na;
    if (Ja < Jb) return true;
    if (Ja > Jb) return false;
    return ia < ib;
}

}  // namespace rcg
