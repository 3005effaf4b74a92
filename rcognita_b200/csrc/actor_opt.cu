// actor_opt.cu -- host side of rcg_actor_opt / rcg_actor_grad / rcg_gather_sqn: argument checks, workspace
// sizing, launch geometry and dispatch to the per-system instantiations (actor_opt_impl.cuh, actor_opt_{ni,3w,2t}.cu).
#include <cstdint>

#include "actor_opt_quad.cuh"

namespace rcg {

static int g_opt_lanes = 0;                                   // rcg_actor_opt_lanes: 0 = default (4 where instantiated), 1 = one lane
static thread_local const char *g_last_opt_kernel = "";      // variant dispatched by the last rcg_actor_opt of this thread

// sqn_out[i][e] = cand[i][(e*C if per-env) + idx[e]]: the start point of the optimiser = the arg-min candidate.
__global__ void gather_sqn_kernel(int L, int64_t E, int C, const double *__restrict__ cand, int cand_per_env,
                                  const int32_t *__restrict__ idx, const int32_t *__restrict__ mask,
                                  double *__restrict__ out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || (mask && mask[e] == 0)) return;
    const int c = idx[e];
    if (c < 0 || c >= C) return;
    const int64_t ld = cand_per_env ? E * (int64_t)C : (int64_t)C;
    const double *src = cand + (cand_per_env ? e * (int64_t)C : 0) + c;
    for (int i = 0; i < L; ++i) out[(int64_t)i * E + e] = src[(int64_t)i * ld];
}

static int opt_common(const char *what, const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S,
                      const double *state_sys, const double *obs, double *sqn, const double *w_critic, int &n, int &m,
                      int &shift)
{
    RCG_REQUIRE(sys && obj && (E <= 0 || (state_sys && obs && sqn)), "%s: null argument", what);
    n = sys_n(sys->sys_id);
    m = sys_m(sys->sys_id);
    RCG_REQUIRE(n > 0, "%s: unknown sys_id %d", what, sys->sys_id);
    RCG_REQUIRE(obj->mode >= RCG_MODE_MPC && obj->mode <= RCG_MODE_SQL, "%s: unknown mode %d", what, obj->mode);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "%s: unknown critic_struct %d", what,
                obj->critic_struct);
    RCG_REQUIRE(obj->Nactor >= 1 && obj->Nactor <= RCG_MAX_NACTOR, "%s: Nactor %d out of range [1, %d]", what,
                obj->Nactor, RCG_MAX_NACTOR);
    RCG_REQUIRE(S >= 1 && S <= 32 && (S & (S - 1)) == 0, "%s: S must be a power of two in [1, 32], got %d", what, S);
    RCG_REQUIRE(obj->mode == RCG_MODE_MPC || w_critic || E <= 0, "%s: w_critic is required in RQL/SQL mode", what);
    shift = 0;
    while ((1 << shift) < S) ++shift;
    return 0;
}

// header (work queue + final costs) + per-thread storage for every launched thread (<= E*S rounded up to a block)
static int64_t opt_ws_doubles(int na, int n, int m, bool generic, int64_t E, int S)
{
    const int64_t threads = (E * S + kOptThreads - 1) / kOptThreads * kOptThreads;
    return kOptWsHeader + E * S + opt_ws_per_thread(na, n, m, generic) * threads;
}

static bool opt_is_generic(const rcg_objective_t *obj, int p)
{
    const bool rdiag = obj->r_is_diag && is_diag(obj->R1, p) && obj->stage_struct == RCG_STAGE_QUADRATIC;
    return !(rdiag && opt_horizon_specialised(obj->Nactor));
}

static int launch_opt(const char *what, const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S,
                      const double *state_sys, const double *obs, double *sqn, const double *w_critic,
                      int32_t w_per_env, const int32_t *mask, int grad_only, int32_t max_iter, double pg_tol,
                      double f_tol, double *ws, int64_t ws_bytes, double *J_out, double *grad_out, int32_t *iters_out,
                      int32_t *nfev_out, int32_t *best_out, double *Jmin_out, double *action_out, double *accum,
                      double sampling_time, void *stream)
{
    int n, m, shift;
    if (int rc = opt_common(what, sys, obj, E, S, state_sys, obs, sqn, w_critic, n, m, shift)) return rc;
    const bool generic = opt_is_generic(obj, n + m);
    const int64_t need = grad_only && !generic ? 0 : opt_ws_doubles(obj->Nactor, n, m, generic, E, S) * (int64_t)sizeof(double);
    RCG_REQUIRE(need == 0 || E <= 0 || (ws && ws_bytes >= need), "%s: workspace too small (%lld bytes given, %lld needed)", what,
                (long long)ws_bytes, (long long)need);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    OptLaunch<double> L;
    L.S = make_sys_dev<double>(sys);
    L.O = make_obj_dev<double>(obj, n, m);
    L.rdiag = obj->r_is_diag && is_diag(obj->R1, n + m) && obj->stage_struct == RCG_STAGE_QUADRATIC;
    L.A.E = E;
    L.A.S = S;
    L.A.S_shift = shift;
    L.A.w_per_env = (int)w_per_env;
    L.A.max_iter = max_iter;
    L.A.pg_tol = pg_tol;
    L.A.f_tol = f_tol;
    L.A.grad_only = grad_only;
    L.A.dynamic = grad_only ? 0 : 1;
    int dev = 0;
    L.sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&L.sms, cudaDevAttrMultiProcessorCount, dev);
    if (L.A.dynamic) cudaMemsetAsync(ws, 0, kOptWsHeader * sizeof(double), (cudaStream_t)stream);      // work-queue counter
    L.state_sys = state_sys; L.obs = obs; L.w = w_critic; L.sqn = sqn; L.mask = mask; L.ws = ws;
    L.J = J_out; L.grad = grad_out; L.iters = iters_out; L.nfev = nfev_out; L.best = best_out; L.Jmin = Jmin_out;
    L.action = action_out; L.accum = accum;
    L.sampling_time = sampling_time;
    L.mode = obj->mode; L.cs = obj->critic_struct;
    const int64_t threads = E * S;
    L.grid = (unsigned)((threads + kOptThreads - 1) / kOptThreads);
    L.stream = (cudaStream_t)stream;
    int rc = 1;
    // G lanes per problem, state in shared memory (actor_opt_quad.cuh): the two robots.  Sys2Tank (n = 2, m = 1, at most 10
    // variables) fits the one-lane kernel's registers and its problems take 1-7 iterations: measured 0.19 ms (one lane) against
    // 0.29 ms (four lanes) per 262,144 SQL N=8 solves, so it stays on the one-lane kernel.
    if (!grad_only && !generic && g_opt_lanes != 1) {
        switch (sys->sys_id) {
        case RCG_SYS_3WROBOT_NI: rc = launch_optq_ni(L); break;
        case RCG_SYS_3WROBOT:    rc = launch_optq_3w(L); break;
        default:                 break;
        }
        if (rc == 0) g_last_opt_kernel = "actor_opt_quad_kernel";
    }
    if (rc == 1) {                                             // one lane per problem (generic shapes, rcg_actor_grad)
        switch (sys->sys_id) {
        case RCG_SYS_3WROBOT_NI: rc = launch_opt_ni(L); break;
        case RCG_SYS_3WROBOT:    rc = launch_opt_3w(L); break;
        default:                 rc = launch_opt_2t(L); break;
        }
        if (!grad_only) g_last_opt_kernel = "actor_opt_kernel";
    }
    if (rc) { set_error("%s: bad mode/critic_struct", what); return rc; }
    return check_launch(what);
}

}  // namespace rcg

extern "C" {

int64_t rcg_actor_opt_workspace_bytes(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S)
{
    if (!sys || !obj) return RCG_EINVAL;
    const int n = rcg::sys_n(sys->sys_id), m = rcg::sys_m(sys->sys_id);
    if (n <= 0 || obj->Nactor < 1 || obj->Nactor > RCG_MAX_NACTOR || E < 0 || S < 1) return RCG_EINVAL;
    return rcg::opt_ws_doubles(obj->Nactor, n, m, rcg::opt_is_generic(obj, n + m), E, S) * (int64_t)sizeof(double);
}

int rcg_actor_grad(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                   const double *obs, const double *sqn, const double *w_critic, int32_t w_per_env, double *workspace,
                   int64_t workspace_bytes, double *J_out, double *grad_out, void *stream)
{
    return rcg::launch_opt("rcg_actor_grad", sys, obj, E, S, state_sys, obs, const_cast<double *>(sqn), w_critic,
                           w_per_env, nullptr, 1, 0, 0.0, 0.0, workspace, workspace_bytes, J_out, grad_out, nullptr,
                           nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, stream);
}

int rcg_actor_opt(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t S, const double *state_sys,
                  const double *obs, double *sqn, const double *w_critic, int32_t w_per_env, const int32_t *mask,
                  int32_t max_iter, double pg_tol, double f_tol, double *workspace, int64_t workspace_bytes,
                  double *J_out, int32_t *iters_out, int32_t *nfev_out, int32_t *best_out, double *Jmin_out,
                  double *action_out, double *accum, double sampling_time, void *stream)
{
    RCG_REQUIRE(max_iter >= 0, "rcg_actor_opt: max_iter must be >= 0");
    return rcg::launch_opt("rcg_actor_opt", sys, obj, E, S, state_sys, obs, sqn, w_critic, w_per_env, mask, 0, max_iter,
                           pg_tol, f_tol, workspace, workspace_bytes, J_out, nullptr, iters_out, nfev_out, best_out,
                           Jmin_out, action_out, accum, sampling_time, stream);
}

int rcg_actor_opt_lanes(int32_t lanes)
{
    const int prev = rcg::g_opt_lanes;
    if (lanes == 0 || lanes == 1 || lanes == 4) rcg::g_opt_lanes = lanes;
    return prev;
}

const char *rcg_last_actor_opt_kernel(void) { return rcg::g_last_opt_kernel; }

int rcg_gather_sqn(int32_t L, int64_t E, int32_t C, const double *cand, int32_t cand_per_env, const int32_t *idx,
                   const int32_t *mask, double *sqn_out, void *stream)
{
    RCG_REQUIRE(E <= 0 || (cand && idx && sqn_out), "rcg_gather_sqn: null argument");
    RCG_REQUIRE(L >= 1 && C >= 1, "rcg_gather_sqn: L and C must be >= 1");
    if (int rc = rcg::require_device()) return rc;
    if (E <= 0) return 0;
    rcg::gather_sqn_kernel<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(L, E, C, cand, cand_per_env, idx,
                                                                                         mask, sqn_out);
    return rcg::check_launch("rcg_gather_sqn");
}

}  // extern "C"
