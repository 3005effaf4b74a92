// actor_ilqr.cu -- rcg_actor_ilqr: the Gauss-Newton (iLQR) pre-pass of the actor optimiser: a persistent grid whose lanes
// pull problems (environment, start point) from a work queue; the per-problem algorithm is actor_ilqr_core.cuh.  Storage: the action sequences stay in
// the caller's [Nactor*m][E*S] array (component-major: consecutive threads touch consecutive addresses); rollout,
// gains, feed-forward steps and the trial sequence live in the workspace with the same [double][problem] layout.
#include <cstdint>

#include "actor_ilqr_core.cuh"
#include "rcg_host.h"

namespace rcg {

constexpr int kIlqrThreads = 128;

constexpr int64_t kIlqrWsHeader = 2;          // doubles in front of the per-lane storage: the work-queue counter

struct IlqrArgs {
    int64_t E;
    int S_shift, w_per_env, mode, cs, dimc, max_sweeps;
    double pg_tol;
};

// The problems of one lane: pulled from a global counter (environments with mask == 0 are skipped).
template <int N>
struct IlqrQueueFeeder {
    const IlqrArgs &A;
    const double *state_sys_g, *obs_g, *w_g;
    double *sqn_g;
    const int32_t *mask_g;
    int32_t *sweeps_g;
    unsigned long long *queue;
    int64_t nprob, p;
    __device__ bool next(double *x0, double *ob0, double *w, double *&U, int64_t &us)
    {
        int64_t e;
        for (;;) {
            p = (int64_t)atomicAdd(queue, 1ull);
            if (p >= nprob) return false;
            e = p >> A.S_shift;
            if (mask_g == nullptr || mask_g[e] != 0) break;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * A.E + e]; ob0[i] = obs_g[i * A.E + e]; }
        if (A.mode != RCG_MODE_MPC)
            for (int i = 0; i < A.dimc; ++i) w[i] = A.w_per_env ? w_g[i * A.E + e] : w_g[i];
        U = sqn_g + p;
        us = nprob;
        return true;
    }
    __device__ bool all_idle(bool idle) const { return __all_sync(0xffffffffu, idle); }
    __device__ void done(int sweeps)
    {
        if (sweeps_g) sweeps_g[p] = sweeps;
    }
};

// Persistent grid: one LANE per problem at a time (see ilqr_run); ws_g = [header][ilqr_ws_per_problem][launched threads].
template <int SYS>
__global__ void __launch_bounds__(kIlqrThreads)
actor_ilqr_kernel(const __grid_constant__ SysDev<double> Sd, const __grid_constant__ ObjDev<double> O,
                  const __grid_constant__ IlqrArgs A, const double *__restrict__ state_sys_g,
                  const double *__restrict__ obs_g, double *__restrict__ sqn_g, const double *__restrict__ w_g,
                  const int32_t *__restrict__ mask_g, double *__restrict__ ws_g, int32_t *__restrict__ sweeps_g)
{
    constexpr int N = SysDim<SYS>::n;
    const int64_t col = (int64_t)blockIdx.x * kIlqrThreads + threadIdx.x;
    const int64_t TR = (int64_t)gridDim.x * kIlqrThreads;
    IlqrQueueFeeder<N> feed{A, state_sys_g, obs_g, w_g, sqn_g, mask_g, sweeps_g,
                            reinterpret_cast<unsigned long long *>(ws_g), A.E << A.S_shift, 0};
    ilqr_run<SYS>(Sd, O, A.mode, A.cs, feed, ws_g + kIlqrWsHeader + col, TR, A.max_sweeps, A.pg_tol);
}

template <int SYS>
static void launch_ilqr(const SysDev<double> &Sd, const ObjDev<double> &O, const IlqrArgs &A, const double *state_sys,
                        const double *obs, double *sqn, const double *w, const int32_t *mask, double *ws, int32_t *sweeps,
                        cudaStream_t st)
{
    auto kern = actor_ilqr_kernel<SYS>;
    static int occ = 0, sms = 0;                               // per instantiation
    if (occ == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kIlqrThreads, 0) != cudaSuccess || occ < 1) occ = 1;
    }
    const int64_t nprob = A.E << A.S_shift;
    unsigned grid = (unsigned)((nprob + kIlqrThreads - 1) / kIlqrThreads);
    const unsigned resident = (unsigned)(sms * occ);
    if (grid > resident) grid = resident;
    cudaMemsetAsync(ws, 0, kIlqrWsHeader * sizeof(double), st);                 // work-queue counter
    kern<<<grid, kIlqrThreads, 0, st>>>(Sd, O, A, state_sys, obs, sqn, w, mask, ws, sweeps);
}

}  // namespace rcg

extern "C" {

int64_t rcg_actor_ilqr_workspace_bytes(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S)
{
    if (!sys || !obj) return RCG_EINVAL;
    const int n = rcg::sys_n(sys->sys_id), m = rcg::sys_m(sys->sys_id);
    if (n <= 0 || obj->Nactor < 1 || obj->Nactor > RCG_MAX_NACTOR || E < 0 || S < 1) return RCG_EINVAL;
    const int64_t threads = (E * S + rcg::kIlqrThreads - 1) / rcg::kIlqrThreads * rcg::kIlqrThreads;      // upper bound of the launch
    return (rcg::kIlqrWsHeader + rcg::ilqr_ws_per_problem(obj->Nactor, n, m) * threads) * (int64_t)sizeof(double);
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
    A.dimc = rcg::ilqr_dim_critic(obj->mode, obj->critic_struct, n, m);
    A.max_sweeps = max_sweeps;
    A.pg_tol = pg_tol;
    cudaStream_t st = (cudaStream_t)stream;
    if (obj->stage_struct != RCG_STAGE_QUADRATIC || max_sweeps == 0) {          // nothing to do: sequences stay as they are
        if (sweeps_out) cudaMemsetAsync(sweeps_out, 0, sizeof(int32_t) * (size_t)(E * S), st);
        return 0;
    }
    const rcg::SysDev<double> Sd = rcg::make_sys_dev<double>(sys);
    const rcg::ObjDev<double> O = rcg::make_obj_dev<double>(obj, n, m);
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI:
        rcg::launch_ilqr<RCG_SYS_3WROBOT_NI>(Sd, O, A, state_sys, obs, sqn, w_critic, mask, workspace, sweeps_out, st);
        break;
    case RCG_SYS_3WROBOT:
        rcg::launch_ilqr<RCG_SYS_3WROBOT>(Sd, O, A, state_sys, obs, sqn, w_critic, mask, workspace, sweeps_out, st);
        break;
    default:
        rcg::launch_ilqr<RCG_SYS_2TANK>(Sd, O, A, state_sys, obs, sqn, w_critic, mask, workspace, sweeps_out, st);
        break;
    }
    return rcg::check_launch("rcg_actor_ilqr");
}

}  // extern "C"
