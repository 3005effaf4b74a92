// actor_tab.cuh -- _actor_cost (rcognita/controllers.py:1273-1328) for a SHARED candidate table on the two robots, with
// everything that depends on the candidate alone taken out of the E x C inner loop.
//
// With one table for all environments the heading of the Euler predictor (systems.py:308-323, :370-382) splits into an
// environment part and a candidate part:
//   Sys3WRobotNI:  theta_k = theta_0 + phi_k,                    phi_k = h * sum_{j<k} omega_j           (the action itself)
//   Sys3WRobot:    theta_k = theta_0 + k h omega_0 + phi_k,      phi_k = h * sum_{j<k} cw_j,  cw_j = (h / I) * sum_{i<j} M_i,
//                  v_k = v_0 + cv_k,  omega_k = omega_0 + cw_k,  cv_k = (h / m) * sum_{i<k} F_i
// so a small kernel tabulates (sin phi_k, cos phi_k, phi_k and v_k resp. cv_k, cw_k) once per candidate and stage -- C x Nactor
// entries instead of E x C x Nactor polynomial sin/cos evaluations -- and the rollout gets sin / cos of the heading from ONE
// angle addition (4 FP64 instructions) instead of the 16-22 of a minimax sincos plus rotation; Sys3WRobot's environment part
// (theta_0 + k h omega_0 and its sin / cos) is computed once per environment by lanes 0 .. Nactor-1.  The table lives in
// shared memory (every warp of a block reads all of it for every environment).  The shared-table launches are FP64-bound
// (config 3's actor launch: 78 % of the FP64 pipe), so instructions are time.  Measured on a B200: Sys3WRobotNI MPC N=6,
// 65,536 x 256: 0.260 -> 0.159 ms (1.05e11 evaluations/s); Sys3WRobot RQL 'quadratic' N=10, 262,144 x 256: 2.41 -> 1.54 ms
// (4.4e10); a first version that read the table through L1 spent 24 % of its instructions on 64-bit addresses: 0.189 / 2.16 ms.
// The sums of the heading are formed in a different order than the reference's step-by-step update (theta_0 + (d_0 + d_1 +
// ...) instead of ((theta_0 + d_0) + d_1) + ...): costs agree with the other kernels and the oracle to ~1e-15 relative, not
// bit for bit.  Lean objective only (diagonal R1 with zero action weights, no target, gamma = 1: every preset), MPC and RQL,
// fp64, horizons 3..10, C a multiple of 32 (Sys3WRobotNI: or a power of two below 32), batches of >= 1,024 environments (one more launch), table within the block's
// shared memory; everything else runs actor_cost_kernel.  A candidate with a non-finite action costs NaN like in the
// reference (0 * a * a = NaN): one flag per candidate.
#pragma once

#include "actor_impl.cuh"

namespace rcg {

template <int SYS> struct TabFields { static constexpr int F = (SYS == RCG_SYS_3WROBOT) ? 5 : 4; };

// tab[(k * F + f) * C + c], k = 0 .. Nactor-1: the candidate part of the predictor state after k Euler steps.
template <int SYS>
__global__ void __launch_bounds__(128)
cand_table_kernel(const __grid_constant__ SysDev<double> S, double h, int C, int na, const double *__restrict__ cand_g,
                  double *__restrict__ tab_g, int32_t *__restrict__ bad_g)
{
    constexpr int M = SysDim<SYS>::m, F = TabFields<SYS>::F;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double phi = 0.0, cv = 0.0, cw = 0.0;
    int bad = 0;
    for (int k = 0; k < na; ++k) {
        double sn, cs;
        sincos(phi, &sn, &cs);
        double *t = tab_g + (int64_t)k * F * C + c;
        t[0] = sn;
        t[(int64_t)C] = cs;
        t[(int64_t)2 * C] = phi;
        const double a0 = cand_g[(int64_t)(k * M) * C + c], a1 = cand_g[(int64_t)(k * M + 1) * C + c];
        bad |= nonfinite_bits(a0) | nonfinite_bits(a1);
        if constexpr (SYS == RCG_SYS_3WROBOT) {
            t[(int64_t)3 * C] = cv;
            t[(int64_t)4 * C] = cw;
            phi = phi + h * cw;                                   // theta' = omega
            cv = cv + h * ((1.0 / S.pars[0]) * a0);               // v' = F / m
            cw = cw + h * ((1.0 / S.pars[1]) * a1);               // omega' = M / I
        } else {
            t[(int64_t)3 * C] = a0;                               // x' = v cos(theta): the action v_k itself
            phi = phi + h * a1;                                   // theta' = omega (= action 1)
        }
    }
    bad_g[c] = bad;
}

// Shared-memory footprint: the whole table [Nactor][F][C] (every warp of the block reads it for every environment; LDS with
// immediate offsets from one per-lane base: no address arithmetic, no global-load latency -- the first version read the table
// through L1 and spent 24 % of its instructions on 64-bit addresses) + per warp the environment part of the heading
// [Nactor][3] (Sys3WRobot: theta_0 + k h omega_0 and its sin / cos, computed once per environment by lanes 0 .. Nactor-1).
// In shared memory a candidate's entries are one ROW ([c][k * F + f], row stride odd: consecutive candidates fall into distinct
// 8-byte banks), so every read is an immediate offset from the lane's row pointer (the [k][f][C] layout of the global table
// cost an integer multiply-add per read: 13 % of the instructions).
template <int SYS, int NA>
__host__ __device__ constexpr int actor_tab_row() { return (NA * TabFields<SYS>::F) | 1; }
template <int SYS, int NA>
__host__ __device__ constexpr size_t actor_tab_smem_bytes(int C)
{
    return ((size_t)actor_tab_row<SYS, NA>() * C + (size_t)kActorWarps * NA * 3) * sizeof(double);
}

template <int SYS, int MODE, int CS, int NA>
__global__ void __launch_bounds__(kActorThreads, actor_min_blocks(SYS, NA, true))
actor_cost_tab_kernel(const __grid_constant__ SysDev<double> S, const __grid_constant__ ObjDev<double> O,
                      const __grid_constant__ ActorArgs A, const double *__restrict__ state_sys_g,
                      const double *__restrict__ obs_g, const double *__restrict__ cand_g, const double *__restrict__ tab_g,
                      const int32_t *__restrict__ bad_g, const double *__restrict__ w_g, const int32_t *__restrict__ mask_g,
                      double *__restrict__ J_g, int32_t *__restrict__ argmin_g, double *__restrict__ Jmin_g,
                      double *__restrict__ action_g, double *__restrict__ accum_g, double sampling_time)
{
    using T = double;
    constexpr int N = SysDim<SYS>::n, M = SysDim<SYS>::m, P = N + M, F = TabFields<SYS>::F;
    constexpr int DIMC = (MODE == RCG_MODE_MPC) ? 1 : dim_critic_c(CS, N, M);
    constexpr int kNone = 0x7fffffff;
    static_assert(SYS == RCG_SYS_3WROBOT || SYS == RCG_SYS_3WROBOT_NI, "robots only");
    static_assert(MODE == RCG_MODE_MPC || MODE == RCG_MODE_RQL, "lean objective: MPC and RQL");
    extern __shared__ double tab_s[];
    const int64_t E = A.E;
    const int C = A.C, seg = A.seg;                           // Sys3WRobot: seg == 32 (one environment per warp, host-checked)
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int slot = lane >> A.seg_shift, cl = lane & (seg - 1);
    const int epw = 32 >> A.seg_shift;                        // environments per warp (C < 32: one power-of-two lane segment each)
    const int64_t warp0 = (int64_t)blockIdx.x * kActorWarps + wi;
    const int64_t nwarps = (int64_t)gridDim.x * kActorWarps;
    const T h = O.pred_step_size;
    T r[N];
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = O.R1[i * P + i];
    constexpr int ROW = actor_tab_row<SYS, NA>();
    for (int i = threadIdx.x; i < NA * F * C; i += kActorThreads) {           // coalesced read of [k * F + f][c], transposed write
        const int j = i / C, c = i - j * C;
        tab_s[c * ROW + j] = tab_g[i];
    }
    T *envp = tab_s + (size_t)ROW * C + (size_t)wi * NA * 3;
    __syncthreads();

    for (int64_t g = warp0; g < A.num_groups; g += nwarps) {
        const int64_t e = g * epw + slot;
        const bool active = e < E && (mask_g == nullptr || mask_g[e] != 0);     // Sys3WRobot: warp-uniform
        T bestJ = T(0);
        int bestI = kNone;
        if (active) {
            T x0[N], ob0[N], w[DIMC];
#pragma unroll
            for (int i = 0; i < N; ++i) { x0[i] = state_sys_g[i * E + e]; ob0[i] = obs_g[i * E + e]; }
            if constexpr (MODE != RCG_MODE_MPC) {
#pragma unroll
                for (int i = 0; i < DIMC; ++i) w[i] = A.w_per_env ? w_g[i * E + e] : w_g[i];
            }
            T s0 = T(0), c0 = T(1);
            if constexpr (SYS == RCG_SYS_3WROBOT) {
                // the environment's own heading after k steps: theta_0 + k h omega_0 (lane k)
                __syncwarp();
                if (lane < NA) {
                    const T th = fma((T)lane, h * x0[4], x0[2]);
                    T sn, cs;
                    sincos_t(th, &sn, &cs);
                    envp[lane * 3 + 0] = sn;
                    envp[lane * 3 + 1] = cs;
                    envp[lane * 3 + 2] = th;
                }
                __syncwarp();
            } else {
                sincos_t(x0[2], &s0, &c0);
            }
            for (int c = cl; c < C; c += seg) {
                const T *t = tab_s + c * ROW;
                T x = x0[0], y = x0[1], J = T(0);
#pragma unroll
                for (int k = 0; k < NA; ++k) {
                    const T sp = t[k * F + 0], cp = t[k * F + 1], phi = t[k * F + 2];
                    T sE = s0, cE = c0, thE = x0[2];
                    if constexpr (SYS == RCG_SYS_3WROBOT) { sE = envp[k * 3 + 0]; cE = envp[k * 3 + 1]; thE = envp[k * 3 + 2]; }
                    const T sn = fma(sE, cp, cE * sp), cs = fma(cE, cp, -(sE * sp));          // sin / cos (theta_k)
                    T ob[N];
                    ob[0] = x; ob[1] = y; ob[2] = thE + phi;
                    T vk;
                    if constexpr (SYS == RCG_SYS_3WROBOT) {
                        vk = x0[3] + t[k * F + 3];
                        ob[3] = vk;
                        ob[4] = x0[4] + t[k * F + 4];
                    } else {
                        vk = t[k * F + 3];                                                      // the action v_k itself
                    }
                    if (k == 0) {
#pragma unroll
                        for (int i = 0; i < N; ++i) ob[i] = ob0[i];                            // stage 0 sees the observation
                    }
                    const bool last = (k + 1 == NA);
                    if (MODE == RCG_MODE_RQL && last) {
                        T a[M];
#pragma unroll
                        for (int j = 0; j < M; ++j) a[j] = __ldg(cand_g + (int64_t)(k * M + j) * C + c);
                        J += critic<T, N, M, CS>(O, ob, a, RegW<T, DIMC>{w});
                    } else {
                        T out = T(0);
#pragma unroll
                        for (int i = 0; i < N; ++i) out += (ob[i] * r[i]) * ob[i];
                        J += out;
                    }
                    if (!last) {
                        x = x + h * (vk * cs);
                        y = y + h * (vk * sn);
                    }
                }
                if (__ldg(bad_g + c)) J = (T)CUDART_NAN;
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
                T act[M], obs_e[N];
#pragma unroll
                for (int j = 0; j < M; ++j) act[j] = cand_g[(int64_t)j * C + bestI];
                if (action_g) {
#pragma unroll
                    for (int j = 0; j < M; ++j) action_g[j * E + e] = act[j];
                }
                if (accum_g) {
#pragma unroll
                    for (int i = 0; i < N; ++i) obs_e[i] = obs_g[i * E + e];
                    accum_g[e] += stage_obj<T, N, M, true, true>(O, obs_e, act) * sampling_time;
                }
            }
        }
    }
}

// scratch of the table path: Nactor * F * C doubles + C flags
inline size_t actor_tab_bytes(int sys, int na, int C)
{
    const int F = (sys == RCG_SYS_3WROBOT) ? 5 : 4;
    return ((size_t)na * F * C * sizeof(double) + 255) / 256 * 256 + (size_t)C * sizeof(int32_t);
}

template <int SYS, int MODE, int CS, int NA>
static int launch_actor_tab_one(const ActorLaunch<double> &L, void *scratch)
{
    constexpr int F = TabFields<SYS>::F;
    const int C = L.A.C, na = NA;
    const size_t smem = actor_tab_smem_bytes<SYS, NA>(C);
    // the table must fit the block's shared memory at the residency the register allocation is held to; Sys3WRobot keeps one
    // environment part per warp, i.e. one environment per warp
    if (smem > (size_t)(225 * 1024) / actor_min_blocks(SYS, NA, true) - 1024) return 1;
    if (SYS == RCG_SYS_3WROBOT && L.A.seg != 32) return 1;
    auto kern = actor_cost_tab_kernel<SYS, MODE, CS, NA>;
    static unsigned long long configured = 0;                 // per instantiation, one bit per device
    if (smem > 48 * 1024) {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        static size_t set_for[64] = {0};
        if (!(configured & bit) || set_for[dev & 63] < smem) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(225 * 1024) / actor_min_blocks(SYS, NA, true) - 1024));
            configured |= bit;
            set_for[dev & 63] = (size_t)(225 * 1024) / actor_min_blocks(SYS, NA, true) - 1024;
        }
    }
    double *tab = static_cast<double *>(scratch);
    int32_t *bad = reinterpret_cast<int32_t *>(static_cast<char *>(scratch) + ((size_t)na * F * C * sizeof(double) + 255) / 256 * 256);
    cand_table_kernel<SYS><<<(unsigned)((C + 127) / 128), 128, 0, L.stream>>>(L.S, L.O.pred_step_size, C, na, L.cand, tab, bad);
    // persistent grid: the blocks copy the table once and stride over the environments
    int64_t grid = (int64_t)L.sms * actor_min_blocks(SYS, NA, true);
    if (L.blocks_needed < grid) grid = L.blocks_needed;
    kern<<<(unsigned)grid, kActorThreads, smem, L.stream>>>(L.S, L.O, L.A, L.state_sys, L.obs, L.cand, tab, bad, L.w, L.mask, L.J,
                                                            L.argmin, L.Jmin, L.action, L.accum, L.sampling_time);
    return 0;
}

template <int SYS, int MODE, int CS>
static int launch_actor_tab_mc(const ActorLaunch<double> &L, void *scratch)
{
    switch (L.O.Nactor) {
    case 3:  return launch_actor_tab_one<SYS, MODE, CS, 3>(L, scratch);
    case 4:  return launch_actor_tab_one<SYS, MODE, CS, 4>(L, scratch);
    case 5:  return launch_actor_tab_one<SYS, MODE, CS, 5>(L, scratch);
    case 6:  return launch_actor_tab_one<SYS, MODE, CS, 6>(L, scratch);
    case 7:  return launch_actor_tab_one<SYS, MODE, CS, 7>(L, scratch);
    case 8:  return launch_actor_tab_one<SYS, MODE, CS, 8>(L, scratch);
    case 9:  return launch_actor_tab_one<SYS, MODE, CS, 9>(L, scratch);
    case 10: return launch_actor_tab_one<SYS, MODE, CS, 10>(L, scratch);
    default: return 1;
    }
}

// 0 = launched; 1 = this shape has no table kernel (the caller runs actor_cost_kernel)
template <int SYS>
static int launch_actor_tab_sys(const ActorLaunch<double> &L, void *scratch)
{
    if (L.mode == RCG_MODE_MPC) return launch_actor_tab_mc<SYS, RCG_MODE_MPC, RCG_CRITIC_QUAD_NOMIX>(L, scratch);
    if (L.mode != RCG_MODE_RQL) return 1;
    switch (L.cs) {
    case RCG_CRITIC_QUAD_LIN:   return launch_actor_tab_mc<SYS, RCG_MODE_RQL, RCG_CRITIC_QUAD_LIN>(L, scratch);
    case RCG_CRITIC_QUADRATIC:  return launch_actor_tab_mc<SYS, RCG_MODE_RQL, RCG_CRITIC_QUADRATIC>(L, scratch);
    case RCG_CRITIC_QUAD_NOMIX: return launch_actor_tab_mc<SYS, RCG_MODE_RQL, RCG_CRITIC_QUAD_NOMIX>(L, scratch);
    case RCG_CRITIC_QUAD_MIX:   return launch_actor_tab_mc<SYS, RCG_MODE_RQL, RCG_CRITIC_QUAD_MIX>(L, scratch);
    default: return 1;
    }
}

// one translation unit per system: actor_tab_ni.cu, actor_tab_3w.cu
int launch_actor_tab_ni(const ActorLaunch<double> &L, void *scratch);
int launch_actor_tab_3w(const ActorLaunch<double> &L, void *scratch);

}  // namespace rcg
