// actor_ilqr.cu -- rcg_actor_ilqr: the Gauss-Newton (iLQR) pre-pass of the actor optimiser, one thread per problem
// (environment, start point); the per-problem algorithm is actor_ilqr_core.cuh.  Storage: the action sequences stay in
// the caller's [Nactor*m][E*S] array (component-major: consecutive threads touch consecutive addresses); rollout,
// gains, feed-forward steps and the trial sequence live in the workspace with the same [double][problem] layout.
#include <cstdint>

#include "actor_ilqr_core.cuh"
#include "rcg_host.h"

namespace rcg {

constexpr int kIlqrThreads = 128;
constexpr int kIlqrMaxW = RCG_MAX_P * (RCG_MAX_P + 1) / 2 + RCG_MAX_P;

struct IlqrArgs {
    int64_t E;
    int S_shift, w_per_env, mode, cs, dimc, max_sweeps;
    double pg_tol;
};

template <int SYS>
__global__ void __launch_bounds__(kIlqrThreads)
actor_ilqr_kernel(const __grid_constant__ SysDev<double> Sd, const __grid_constant__ ObjDev<double> O,
                  const __grid_constant__ IlqrArgs A, const double *__restrict__ state_sys_g,
                  const double *__restrict__ obs_g, double *__restrict__ sqn_g, const double *__restrict__ w_g,
                  const int32_t *__restrict__ mask_g, double *__restrict__ ws_g, int32_t *__restrict__ sweeps_g)
{
    constexpr int N = SysDim<SYS>::n;
    const int64_t nprob = A.E << A.S_shift;
    const int64_t p = (int64_t)blockIdx.x * kIlqrThreads + threadIdx.x;
    if (p >= nprob) return;
    const int64_t e = p >> A.S_shift;
    if (mask_g && mask_g[e] == 0) return;
    double x0[N], ob0[N], w[kIlqrMaxW];
#pragma unroll
    for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * A.E + e]; ob0[i] = obs_g[i * A.E + e]; }
    if (A.mode != RCG_MODE_MPC)
        for (int i = 0; i < A.dimc; ++i) w[i] = A.w_per_env ? w_g[i * A.E + e] : w_g[i];
    const int sweeps = ilqr_presweeps<SYS>(Sd, O, A.mode, A.cs, x0, ob0, w, sqn_g + p, nprob, ws_g + p, nprob,
                                           A.max_sweeps, A.pg_tol);
    if (sweeps_g) sweeps_g[p] = sweeps;
}

static int ilqr_dimc(int cs, int n, int m)
{
    const int p = n + m;
    switch (cs) {
    case RCG_CRITIC_QUAD_LIN:   return p * (p + 1) / 2 + p;
    case RCG_CRITIC_QUADRATIC:  return p * (p + 1) / 2;
    case RCG_CRITIC_QUAD_NOMIX: return p;
    default:                    return n + n * m + m;
    }
}

}  // namespace rcg

extern "C" {

int64_t rcg_actor_ilqr_workspace_bytes(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S)
{
    if (!sys || !obj) return RCG_EINVAL;
    const int n = rcg::sys_n(sys->sys_id), m = rcg::sys_m(sys->sys_id);
    if (n <= 0 || obj->Nactor < 1 || obj->Nactor > RCG_MAX_NACTOR || E < 0 || S < 1) return RCG_EINVAL;
    return rcg::ilqr_ws_per_problem(obj->Nactor, n, m) * E * S * (int64_t)sizeof(double);
}

int rcg_actor_ilqr(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                   const double *obs, double *sqn, const double *w_critic, int32_t w_per_env, const int32_t *mask,
                   int32_t max_sweeps, double pg_tol, double *workspace, int64_t workspace_bytes, int32_t *sweeps_out,
                   void *stream)
{
    RCG_REQUIRE(sys && obj && (E <= 0 || (state_sys && obs && sqn)), "rcg_actor_ilqr: null argument");
    const int n = rcg::sys_n(sys->sys_id), m = rcg::sys_m(sys->sys_id);
    RCG_REQUIRE(n > 0, "rcg_actor_ilqr: unknown sys_id %d", sys->sys_id);
    RCG_REQUIRE(obj->mode >= RCG_MODE_MPC && obj->mode <= RCG_MODE_SQL, "rcg_actor_ilqr: unknown mode %d", obj->mode);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "rcg_actor_ilqr: unknown critic_struct %d",
                obj->critic_struct);
    RCG_REQUIRE(obj->Nactor >= 1 && obj->Nactor <= RCG_MAX_NACTOR, "rcg_actor_ilqr: Nactor %d out of range [1, %d]",
                obj->Nactor, RCG_MAX_NACTOR);
    RCG_REQUIRE(S >= 1 && S <= 32 && (S & (S - 1)) == 0, "rcg_actor_ilqr: S must be a power of two in [1, 32], got %d", S);
    RCG_REQUIRE(max_sweeps >= 0, "rcg_actor_ilqr: max_sweeps must be >= 0");
    RCG_REQUIRE(obj->mode == RCG_MODE_MPC || w_critic || E <= 0, "rcg_actor_ilqr: w_critic is required in RQL/SQL mode");
    const int64_t need = rcg_actor_ilqr_workspace_bytes(sys, obj, E < 0 ? 0 : E, S);
    RCG_REQUIRE(E <= 0 || (workspace && workspace_bytes >= need),
                "rcg_actor_ilqr: workspace too small (%lld bytes given, %lld needed)", (long long)workspace_bytes,
                (long long)need);
    if (int rc = rcg::require_device()) return rc;
    if (E <= 0) return 0;
    rcg::IlqrArgs A;
    A.E = E;
    A.S_shift = 0;
    while ((1 << A.S_shift) < S) ++A.S_shift;
    A.w_per_env = (int)w_per_env;
    A.mode = obj->mode;
    A.cs = obj->critic_struct;
    A.dimc = rcg::ilqr_dimc(obj->critic_struct, n, m);
    A.max_sweeps = max_sweeps;
    A.pg_tol = pg_tol;
    const rcg::SysDev<double> Sd = rcg::make_sys_dev<double>(sys);
    const rcg::ObjDev<double> O = rcg::make_obj_dev<double>(obj, n, m);
    const unsigned grid = (unsigned)((E * S + rcg::kIlqrThreads - 1) / rcg::kIlqrThreads);
    cudaStream_t st = (cudaStream_t)stream;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI:
        rcg::actor_ilqr_kernel<RCG_SYS_3WROBOT_NI><<<grid, rcg::kIlqrThreads, 0, st>>>(Sd, O, A, state_sys, obs, sqn, w_critic,
                                                                                      mask, workspace, sweeps_out);
        break;
    case RCG_SYS_3WROBOT:
        rcg::actor_ilqr_kernel<RCG_SYS_3WROBOT><<<grid, rcg::kIlqrThreads, 0, st>>>(Sd, O, A, state_sys, obs, sqn, w_critic, mask,
                                                                                   workspace, sweeps_out);
        break;
    default:
        rcg::actor_ilqr_kernel<RCG_SYS_2TANK><<<grid, rcg::kIlqrThreads, 0, st>>>(Sd, O, A, state_sys, obs, sqn, w_critic, mask,
                                                                                 workspace, sweeps_out);
        break;
    }
    return rcg::check_launch("rcg_actor_ilqr");
}

}  // extern "C"
