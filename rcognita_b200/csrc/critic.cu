// critic.cu -- batched CtrlOptPred.stage_obj / upd_accum_obj, _critic and _critic_cost
// (rcognita/controllers.py:1063-1093, :1192-1245) and the FIFO push of the controller
// buffers (rcognita/utilities.py:78-79).  One thread per environment (x weight vector).
#include "rcg_host.h"

namespace rcg {

template <typename T>
struct GlobalW {                 // weight i of lane `idx` in a [dimc][stride] array (shared vector: stride 1, idx 0)
    const T *w;
    int64_t stride, idx;
    __device__ __forceinline__ T operator()(int i) const { return w[i * stride + idx]; }
};

template <typename T, int N, int M, bool RDIAG>
__global__ void __launch_bounds__(256)
stage_obj_kernel(const __grid_constant__ ObjDev<T> O, int64_t E, const T *__restrict__ obs_g, const T *__restrict__ act_g,
                 T *__restrict__ out_g, T *__restrict__ accum_g, T scale)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    T obs[N], act[M];
#pragma unroll
    for (int i = 0; i < N; ++i) obs[i] = obs_g[i * E + e];
#pragma unroll
    for (int j = 0; j < M; ++j) act[j] = act_g[j * E + e];
    const T r = stage_obj<T, N, M, RDIAG>(O, obs, act);
    if (out_g) out_g[e] = r;
    if (accum_g) accum_g[e] += r * scale;            // upd_accum_obj, controllers.py:1093
}

template <typename T, int N, int M, int CS>
__global__ void __launch_bounds__(256)
critic_kernel(const __grid_constant__ ObjDev<T> O, int64_t E, const T *__restrict__ obs_g, const T *__restrict__ act_g,
              const T *__restrict__ w_g, int w_per_env, T *__restrict__ out_g)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    T obs[N], act[M];
#pragma unroll
    for (int i = 0; i < N; ++i) obs[i] = obs_g[i * E + e];
#pragma unroll
    for (int j = 0; j < M; ++j) act[j] = act_g[j * E + e];
    const GlobalW<T> w{w_g, w_per_env ? E : 1, w_per_env ? e : 0};
    out_g[e] = critic<T, N, M, CS>(O, obs, act, w);
}

// _critic_cost: thread (e, k) evaluates weight vector k of environment e against the
// Ncritic OLDEST rows of that environment's buffers (rows 0..Ncritic-1, controllers.py:1230-1234).
template <typename T, int N, int M, int CS, bool RDIAG>
__global__ void __launch_bounds__(256)
critic_cost_kernel(const __grid_constant__ ObjDev<T> O, int64_t E, int W, const T *__restrict__ obs_buf,
                   const T *__restrict__ act_buf, const T *__restrict__ w_g, const T *__restrict__ wprev_g,
                   T *__restrict__ out_g)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= E * (int64_t)W) return;
    const int64_t e = idx / W;
    const GlobalW<T> w{w_g, E * (int64_t)W, idx};
    const GlobalW<T> wp{wprev_g, E, e};
    T Jc = T(0);
    for (int k = O.Ncritic - 1; k > 0; --k) {
        T op[N], on[N], ap[M], an[M];
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
        const T q_prev = critic<T, N, M, CS>(O, op, ap, w);
        const T q_next = critic<T, N, M, CS>(O, on, an, wp);
        const T err = q_prev - O.gamma * q_next - stage_obj<T, N, M, RDIAG>(O, op, ap);   // :1240
        Jc += T(0.5) * (err * err);                                                       // :1242
    }
    out_g[idx] = Jc;
}

// push_vec on both buffers, for the lanes with mask != 0: row r <- row r+1, last row <- new.
template <typename T>
__global__ void __launch_bounds__(256)
push_buffers_kernel(int n, int m, int L, int64_t E, T *__restrict__ obs_buf, T *__restrict__ act_buf,
                    const T *__restrict__ obs, const T *__restrict__ act, const int32_t *__restrict__ mask)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (mask && mask[e] == 0) return;
    for (int r = 0; r + 1 < L; ++r) {
        for (int i = 0; i < n; ++i) obs_buf[((int64_t)r * n + i) * E + e] = obs_buf[((int64_t)(r + 1) * n + i) * E + e];
        for (int j = 0; j < m; ++j) act_buf[((int64_t)r * m + j) * E + e] = act_buf[((int64_t)(r + 1) * m + j) * E + e];
    }
    for (int i = 0; i < n; ++i) obs_buf[((int64_t)(L - 1) * n + i) * E + e] = obs[i * E + e];
    for (int j = 0; j < m; ++j) act_buf[((int64_t)(L - 1) * m + j) * E + e] = act[j * E + e];
}

// The clock test of CtrlOptPred.compute_action (controllers.py:1440-1442, and :1459-1468 for
// the critic clock): mask = (t - clock >= period) [& in_mask]; clock = t where mask.
__global__ void __launch_bounds__(256)
ctrl_sample_kernel(int64_t E, const double *__restrict__ t, double *__restrict__ clock, double period,
                   const int32_t *__restrict__ in_mask, int32_t *__restrict__ mask_out)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const double te = t[e];
    const bool fire = (in_mask == nullptr || in_mask[e] != 0) && (te - clock[e] >= period);
    if (fire) clock[e] = te;
    mask_out[e] = fire ? 1 : 0;
}

static bool obj_rdiag(const rcg_objective_t *obj, int p)
{
    return obj->r_is_diag && is_diag(obj->R1, p) && (obj->stage_struct == RCG_STAGE_QUADRATIC || is_diag(obj->R2, p));
}

static int dims_ok(int n, int m) { return (n == 3 && m == 2) || (n == 5 && m == 2) || (n == 2 && m == 1); }

#define RCG_DISPATCH_NM(n, m, CALL)          \
    if (n == 3) { CALL(3, 2) }               \
    else if (n == 5) { CALL(5, 2) }          \
    else { CALL(2, 1) }

template <typename T>
static int launch_critic_cost(const char *what, const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, int32_t W,
                              const T *obs_buf, const T *act_buf, const T *w, const T *w_prev, T *Jc_out, void *stream)
{
    RCG_REQUIRE(obj && (E <= 0 || (obs_buf && act_buf && w && w_prev && Jc_out)), "%s: null argument", what);
    RCG_REQUIRE(dims_ok(n, m), "%s: unsupported dims n=%d m=%d", what, n, m);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "%s: unknown critic_struct %d", what, obj->critic_struct);
    RCG_REQUIRE(W >= 1, "%s: W must be >= 1", what);
    RCG_REQUIRE(obj->Ncritic >= 1 && obj->Ncritic <= obj->buffer_size, "%s: Ncritic %d must be in [1, buffer_size = %d]", what,
                obj->Ncritic, obj->buffer_size);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const ObjDev<T> O = make_obj_dev<T>(obj, n, m);
    const bool rd = obj_rdiag(obj, n + m);
    const unsigned grid = (unsigned)((E * (int64_t)W + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
#define CC(NN, MM, CS)                                                                                                  \
    if (rd) critic_cost_kernel<T, NN, MM, CS, true><<<grid, 256, 0, s>>>(O, E, W, obs_buf, act_buf, w, w_prev, Jc_out); \
    else critic_cost_kernel<T, NN, MM, CS, false><<<grid, 256, 0, s>>>(O, E, W, obs_buf, act_buf, w, w_prev, Jc_out);
#define CALL(NN, MM)                     \
    switch (obj->critic_struct) {        \
    case 0: CC(NN, MM, 0) break;         \
    case 1: CC(NN, MM, 1) break;         \
    case 2: CC(NN, MM, 2) break;         \
    default: CC(NN, MM, 3) break;        \
    }
    RCG_DISPATCH_NM(n, m, CALL)
#undef CALL
#undef CC
    return check_launch(what);
}

}  // namespace rcg

extern "C" {

int rcg_stage_obj(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs, const double *act,
                  double *out, double *accum, double scale, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(obj && (E <= 0 || (obs && act && (out || accum))), "rcg_stage_obj: null argument");
    RCG_REQUIRE(dims_ok(n, m), "rcg_stage_obj: unsupported dims n=%d m=%d", n, m);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const ObjDev<double> O = make_obj_dev<double>(obj, n, m);
    const bool rd = obj_rdiag(obj, n + m);
    const unsigned grid = (unsigned)((E + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(NN, MM)                                                                                          \
    if (rd) stage_obj_kernel<double, NN, MM, true><<<grid, 256, 0, s>>>(O, E, obs, act, out, accum, scale);  \
    else stage_obj_kernel<double, NN, MM, false><<<grid, 256, 0, s>>>(O, E, obs, act, out, accum, scale);
    RCG_DISPATCH_NM(n, m, CALL)
#undef CALL
    return check_launch("rcg_stage_obj");
}

int rcg_critic(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, const double *obs, const double *act,
               const double *w, int32_t w_per_env, double *out, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(obj && (E <= 0 || (obs && act && w && out)), "rcg_critic: null argument");
    RCG_REQUIRE(dims_ok(n, m), "rcg_critic: unsupported dims n=%d m=%d", n, m);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "rcg_critic: unknown critic_struct %d",
                obj->critic_struct);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const ObjDev<double> O = make_obj_dev<double>(obj, n, m);
    const unsigned grid = (unsigned)((E + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(NN, MM)                                                                                                   \
    switch (obj->critic_struct) {                                                                                      \
    case 0: critic_kernel<double, NN, MM, 0><<<grid, 256, 0, s>>>(O, E, obs, act, w, w_per_env, out); break;           \
    case 1: critic_kernel<double, NN, MM, 1><<<grid, 256, 0, s>>>(O, E, obs, act, w, w_per_env, out); break;           \
    case 2: critic_kernel<double, NN, MM, 2><<<grid, 256, 0, s>>>(O, E, obs, act, w, w_per_env, out); break;           \
    default: critic_kernel<double, NN, MM, 3><<<grid, 256, 0, s>>>(O, E, obs, act, w, w_per_env, out); break;          \
    }
    RCG_DISPATCH_NM(n, m, CALL)
#undef CALL
    return check_launch("rcg_critic");
}

int rcg_critic_cost(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, int32_t W, const double *obs_buf,
                    const double *act_buf, const double *w, const double *w_prev, double *Jc_out, void *stream)
{
    return rcg::launch_critic_cost<double>("rcg_critic_cost", obj, n, m, E, W, obs_buf, act_buf, w, w_prev, Jc_out, stream);
}

int rcg_critic_cost_f32(const rcg_objective_t *obj, int32_t n, int32_t m, int64_t E, int32_t W, const float *obs_buf,
                        const float *act_buf, const float *w, const float *w_prev, float *Jc_out, void *stream)
{
    return rcg::launch_critic_cost<float>("rcg_critic_cost_f32", obj, n, m, E, W, obs_buf, act_buf, w, w_prev, Jc_out, stream);
}

int rcg_ctrl_sample(int64_t E, const double *t, double *clock, double period, const int32_t *in_mask, int32_t *mask_out,
                    void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(E <= 0 || (t && clock && mask_out), "rcg_ctrl_sample: null argument");
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    ctrl_sample_kernel<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(E, t, clock, period, in_mask, mask_out);
    return check_launch("rcg_ctrl_sample");
}

int rcg_push_buffers(int32_t n, int32_t m, int32_t buffer_size, int64_t E, double *obs_buf, double *act_buf,
                     const double *obs, const double *act, const int32_t *mask, void *stream)
{
    using namespace rcg;
    RCG_REQUIRE(E <= 0 || (obs_buf && act_buf && obs && act), "rcg_push_buffers: null argument");
    RCG_REQUIRE(n >= 1 && m >= 1 && buffer_size >= 1, "rcg_push_buffers: bad dims");
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;
    const unsigned grid = (unsigned)((E + 255) / 256);
    push_buffers_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, buffer_size, E, obs_buf, act_buf, obs, act,
                                                                        mask);
    return check_launch("rcg_push_buffers");
}

}  // extern "C"
