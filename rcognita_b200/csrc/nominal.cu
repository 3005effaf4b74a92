// nominal.cu -- CtrlNominal3WRobotNI (rcognita/controllers.py:1758-1956), the nominal parking controller of the
// non-holonomic integrator and the default ctrl_mode of presets/main_3wrobot_NI.py, for E environments: one
// thread per environment, closed form (no optimisation), fused with the controller's ZOH hand-over and
// upd_accum_obj.  Compiled with -fmad=false: every product and sum is rounded like the Python expression.
#include "rcg_host.h"

namespace rcg {

__device__ __forceinline__ double sgn(double v) { return (v > 0.0) ? 1.0 : (v < 0.0) ? -1.0 : v; }   // np.sign

// compute_action_vanila + the clipping of compute_action (:1915-1921): _Cart2NH (:1880-1893), _zeta (:1786-1834),
// _kappa (:1836-1853), uNI = ctrl_gain * kappa, _NH2ctrl_Cart (:1895-1905).  Expressions in Python's order;
// x**3 and |x|**(1/3) through pow() like numpy's scalar power.
__device__ __forceinline__ void nominal_ni(double gain, const SysDev<double> &S, const double *obs, double *action)
{
    const double xc = obs[0], yc = obs[1], alpha = obs[2];
    double sa, ca;
    sincos(alpha, &sa, &ca);
    double x[3], zeta[3];
    x[0] = alpha;
    x[1] = xc * ca + yc * sa;
    x[2] = -2 * (yc * ca - xc * sa) - alpha * (xc * ca + yc * sa);
    const double a2 = fabs(x[2]);
    if (x[0] == 0 && x[1] == 0) {                                   // nablaF with theta = 0
        const double st = x[0] * 1.0 + x[1] * 0.0 + sqrt(a2);
        zeta[0] = 4 * pow(x[0], 3.0) - 2 * pow(a2, 3.0) * 1.0 / pow(st, 3.0);
        zeta[1] = 4 * pow(x[1], 3.0) - 2 * pow(a2, 3.0) * 0.0 / pow(st, 3.0);
        zeta[2] = (3 * x[0] * 1.0 + 3 * x[1] * 0.0 + 2 * sqrt(a2)) * (x[2] * x[2]) * sgn(x[2]) / pow(st, 3.0);
    } else {                                                        // nablaL
        const double r = sqrt(x[0] * x[0] + x[1] * x[1]);
        const double sigma = r + sqrt(a2);
        const double q = pow(a2, 3.0) / pow(sigma, 3.0);
        zeta[0] = 4 * pow(x[0], 3.0) + q * 1 / pow(r, 3.0) * 2 * x[0];
        zeta[1] = 4 * pow(x[1], 3.0) + q * 1 / pow(r, 3.0) * 2 * x[1];
        zeta[2] = 3 * (a2 * a2) * sgn(x[2]) + q * 1 / sqrt(a2) * sgn(x[2]);
    }
    const double d0 = zeta[0] * 1.0 + zeta[1] * 0.0 + zeta[2] * x[1];
    const double d1 = zeta[0] * 0.0 + zeta[1] * 1.0 + zeta[2] * (-x[0]);
    const double k0 = -pow(fabs(d0), 1.0 / 3) * sgn(d0);
    const double k1 = -pow(fabs(d1), 1.0 / 3) * sgn(d1);
    const double u0 = gain * k0, u1 = gain * k1;
    action[0] = u1 + 1.0 / 2 * u0 * (x[2] + x[0] * x[1]);
    action[1] = u0;
    clip_action<double, 2>(S, action);
}

template <bool RDIAG>
__global__ void __launch_bounds__(256)
nominal_ni_kernel(const __grid_constant__ SysDev<double> S, const __grid_constant__ ObjDev<double> O, int64_t E,
                  double gain, const double *__restrict__ obs_g, const int32_t *__restrict__ mask_g,
                  double *__restrict__ action_g, double *__restrict__ accum_g, double sampling_time)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || (mask_g && mask_g[e] == 0)) return;
    double obs[3], act[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) obs[i] = obs_g[i * E + e];
    nominal_ni(gain, S, obs, act);
    action_g[e] = act[0];
    action_g[E + e] = act[1];
    if (accum_g) accum_g[e] += stage_obj<double, 3, 2, RDIAG, false>(O, obs, act) * sampling_time;
}

}  // namespace rcg

extern "C" int rcg_nominal_ni(const rcg_system_t *sys, int64_t E, const double *obs, double ctrl_gain,
                              const int32_t *mask, double *action, const rcg_objective_t *obj, double *accum,
                              double sampling_time, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(sys && (E <= 0 || (obs && action)), "rcg_nominal_ni: null argument");
    RCG_REQUIRE(sys->sys_id == RCG_SYS_3WROBOT_NI, "rcg_nominal_ni: the nominal controller is defined for Sys3WRobotNI "
                "(sys_id %d given)", sys->sys_id);
    RCG_REQUIRE(!accum || obj, "rcg_nominal_ni: accum needs the objective descriptor (stage_obj)");
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const SysDev<double> S = make_sys_dev<double>(sys);
    ObjDev<double> O{};
    bool rdiag = true;
    if (obj) {
        O = make_obj_dev<double>(obj, 3, 2);
        rdiag = obj->r_is_diag && is_diag(obj->R1, 5) && obj->stage_struct == RCG_STAGE_QUADRATIC;
    }
    const unsigned grid = (unsigned)((E + 255) / 256);
    if (rdiag)
        nominal_ni_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(S, O, E, ctrl_gain, obs, mask, action, accum, sampling_time);
    else
        nominal_ni_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(S, O, E, ctrl_gain, obs, mask, action, accum, sampling_time);
    return check_launch("rcg_nominal_ni");
}
