// actor.cu -- host side of rcg_actor_cost / rcg_actor_cost_f32: argument checks, launch geometry
// and dispatch to the per-system kernel instantiations (actor_impl.cuh, actor_{ni,3w,2t}_{f64,f32}.cu).
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "actor_tab.cuh"

namespace rcg {

static thread_local const char *g_last_actor_kernel = "";       // variant dispatched by the last rcg_actor_cost of this thread

// 2-D tensor map of a per-environment candidate array [rows = Nactor*m][cols = E*C] (row-major, cols
// contiguous), box = 32 columns x all rows.  cuTensorMapEncodeTiled is resolved through the runtime so
// that the library does not link against libcuda directly.
static int make_cand_tensor_map(CUtensorMap *tm, const void *base, size_t elem, int64_t cols, int rows, int box_rows)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            (void)cudaGetLastError();
            set_error("rcg_actor_cost: cuTensorMapEncodeTiled is not available from this driver");
            return RCG_ENODEV;
        }
        encode = (EncodeFn)fn;
    }
    // The descriptor depends only on (base, element size, shape, box): a controller calls with the same candidate array
    // at every sample, so the last few descriptors are kept per thread instead of being re-encoded on every launch.
    struct Key { const void *base; size_t elem; int64_t cols; int rows, box_rows; };
    constexpr int kCache = 8;
    static thread_local Key keys[kCache];
    static thread_local CUtensorMap maps[kCache];
    static thread_local int used = 0, next = 0;
    for (int i = 0; i < used; ++i)
        if (keys[i].base == base && keys[i].elem == elem && keys[i].cols == cols && keys[i].rows == rows && keys[i].box_rows == box_rows) {
            *tm = maps[i];
            return 0;
        }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * elem};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult r = encode(tm, elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                              const_cast<void *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("rcg_actor_cost: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return RCG_EINVAL;
    }
    keys[next] = Key{base, elem, cols, rows, box_rows};
    maps[next] = *tm;
    next = (next + 1) % kCache;
    if (used < kCache) ++used;
    return 0;
}

// Scratch of the shared-table path (actor_tab.cuh: Nactor * F * C doubles + C flags), one buffer per (thread, stream), grown
// on demand and kept.  Returns nullptr when none can be had without side effects: the stream is being captured (a captured
// graph must not depend on a buffer a later, larger call would replace) or the allocation fails.
static void *actor_tab_scratch(cudaStream_t stream, size_t bytes)
{
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return nullptr;
    }
    struct Slot { cudaStream_t stream; int dev; void *ptr; size_t bytes; };
    constexpr int kSlots = 16;
    static thread_local Slot slots[kSlots];
    static thread_local int used = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    Slot *s = nullptr;
    for (int i = 0; i < used; ++i)
        if (slots[i].stream == stream && slots[i].dev == dev) { s = &slots[i]; break; }
    if (!s) {
        if (used == kSlots) return nullptr;
        s = &slots[used++];
        *s = Slot{stream, dev, nullptr, 0};
    }
    if (s->bytes < bytes) {
        if (s->ptr) {
            cudaStreamSynchronize(stream);                   // the previous launch on this stream may still read it
            cudaFree(s->ptr);
        }
        s->ptr = nullptr;
        s->bytes = 0;
        if (cudaMalloc(&s->ptr, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            s->ptr = nullptr;
            return nullptr;
        }
        s->bytes = bytes;
    }
    return s->ptr;
}

// Smallest batch that takes the table path: RCG_ACTOR_TABLE_MIN_E (default 1024; 1 forces it for tests, a huge value or
// RCG_ACTOR_NO_TABLE=1 turns it off).
static int64_t actor_tab_min_envs()
{
    if (getenv("RCG_ACTOR_NO_TABLE")) return INT64_MAX;
    const char *v = getenv("RCG_ACTOR_TABLE_MIN_E");
    return v ? atoll(v) : 1024;
}

template <typename T>
static int launch_actor(const char *what, const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                        const T *state_sys, const T *obs, const T *cand, int32_t cand_per_env, const T *w_critic,
                        int32_t w_per_env, const int32_t *mask, T *J_out, int32_t *argmin_out, T *Jmin_out,
                        T *action_out, T *accum, double sampling_time, void *stream)
{
    RCG_REQUIRE(sys && obj && (E <= 0 || (state_sys && obs && cand)), "%s: null argument", what);
    const int n = sys_n(sys->sys_id), m = sys_m(sys->sys_id);
    RCG_REQUIRE(n > 0, "%s: unknown sys_id %d", what, sys->sys_id);
    RCG_REQUIRE(obj->mode >= RCG_MODE_MPC && obj->mode <= RCG_MODE_SQL, "%s: unknown mode %d", what, obj->mode);
    RCG_REQUIRE(obj->critic_struct >= 0 && obj->critic_struct <= 3, "%s: unknown critic_struct %d", what,
                obj->critic_struct);
    RCG_REQUIRE(obj->Nactor >= 1 && obj->Nactor <= RCG_MAX_NACTOR, "%s: Nactor %d out of range [1, %d]", what,
                obj->Nactor, RCG_MAX_NACTOR);
    RCG_REQUIRE(C >= 1, "%s: C must be >= 1", what);
    RCG_REQUIRE(obj->mode == RCG_MODE_MPC || w_critic || E <= 0, "%s: w_critic is required in RQL/SQL mode", what);
    if (int rc = require_device()) return rc;
    if (E <= 0) return 0;

    ActorLaunch<T> L;
    L.S = make_sys_dev<T>(sys);
    L.O = make_obj_dev<T>(obj, n, m);
    // fast kernels: diagonal R1 and the 'quadratic' structure (every preset); otherwise the general ones
    L.rdiag = obj->r_is_diag && is_diag(obj->R1, n + m) && obj->stage_struct == RCG_STAGE_QUADRATIC;
    // lean objective (every robot preset): no target, gamma = 1, zero weight on the action entries of diagonal R1
    L.lean = L.rdiag && !obj->has_target && getenv("RCG_ACTOR_NO_LEAN") == nullptr;
    for (int k = 0; k < obj->Nactor && L.lean; ++k) L.lean = (obj->gamma_pow[k] == 1.0);
    for (int j = n; j < n + m && L.lean; ++j) L.lean = (obj->R1[j * (n + m) + j] == 0.0);
    int seg = 32, shift = 5;
    while (shift > 0 && (seg >> 1) >= C) { seg >>= 1; --shift; }      // smallest power of two >= C, capped at 32
    const int epw = 32 / seg;
    L.A.E = E;
    L.A.C = (int)C;
    L.A.seg = seg;
    L.A.seg_shift = shift;
    L.A.num_groups = (E + epw - 1) / epw;
    L.A.cand_per_env = (int)cand_per_env;
    L.A.w_per_env = (int)w_per_env;
    L.state_sys = state_sys; L.obs = obs; L.cand = cand; L.w = w_critic; L.mask = mask;
    L.J = J_out; L.argmin = argmin_out; L.Jmin = Jmin_out; L.action = action_out; L.accum = accum;
    L.sampling_time = (T)sampling_time;
    L.mode = obj->mode; L.cs = obj->critic_struct;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent-style grid: a whole number of resident blocks per SM, warps stride over the environments
    const int64_t blocks_needed = (L.A.num_groups + kActorWarps - 1) / kActorWarps;
    const int64_t max_grid = (int64_t)sms * 8;
    L.grid = (unsigned)(blocks_needed < max_grid ? blocks_needed : max_grid);
    L.sms = sms;
    L.blocks_needed = blocks_needed;
    // TMA-staged kernel: per-environment candidates, diagonal R, a specialised horizon, and a candidate
    // count whose lane mapping is a contiguous box (C a multiple of 32, or a power of two below 32)
    L.use_tma = false;
    L.use_tma_rt = false;
    {
        const int na = obj->Nactor;
        const bool na_ok = na >= 3 && na <= 10;                   // horizons with a compile-time instantiation
        const bool c_ok = (C % 32 == 0) || (C < 32 && (C & (C - 1)) == 0);
        const int64_t cols = E * (int64_t)C;
        if (cand_per_env && L.rdiag && na_ok && c_ok && cols % (16 / (int)sizeof(T)) == 0 && cols < (int64_t)1 << 31 &&
            ((uintptr_t)cand % 16) == 0 && getenv("RCG_ACTOR_NO_TMA") == nullptr) {
            if (int rc = make_cand_tensor_map(&L.tmap, cand, sizeof(T), cols, na * m, na * m)) return rc;
            L.use_tma = true;
        } else if (cand_per_env && L.rdiag && !na_ok && c_ok && cols % (16 / (int)sizeof(T)) == 0 && cols < (int64_t)1 << 31 &&
                   ((uintptr_t)cand % 16) == 0 && getenv("RCG_ACTOR_NO_TMA") == nullptr) {
            // runtime horizon: boxes of kRtChunk stages; rows past Nactor*m of the last box are zero-filled
            if (int rc = make_cand_tensor_map(&L.tmap, cand, sizeof(T), cols, na * m, kRtChunk * m)) return rc;
            L.use_tma_rt = true;
        }
    }
    L.stream = (cudaStream_t)stream;
    g_last_actor_kernel = L.use_tma ? "actor_cost_tma_kernel" : L.use_tma_rt ? "actor_cost_tma_rt_kernel" : "actor_cost_kernel";
    // Shared candidate table on a robot with the presets' (lean) objective, MPC / RQL, horizons 3..10, fp64: the candidate
    // part of the heading is tabulated once per launch (actor_tab.cuh).  Not for small batches (one more launch), not while
    // the stream is being captured.
    if constexpr (std::is_same<T, double>::value) {
        const int na = obj->Nactor;
        if (!cand_per_env && ((C % 32 == 0) || (C < 32 && (C & (C - 1)) == 0)) && L.lean && L.rdiag && na >= 3 && na <= 10 && sys->sys_id != RCG_SYS_2TANK &&
            (obj->mode == RCG_MODE_MPC || obj->mode == RCG_MODE_RQL) && E >= actor_tab_min_envs()) {
            if (void *scratch = actor_tab_scratch(L.stream, actor_tab_bytes(sys->sys_id, na, C))) {
                const int rt = (sys->sys_id == RCG_SYS_3WROBOT_NI) ? launch_actor_tab_ni(L, scratch) : launch_actor_tab_3w(L, scratch);
                if (rt == 0) {
                    g_last_actor_kernel = "actor_cost_tab_kernel";
                    return check_launch(what);
                }
            }
        }
    }
    int rc;
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI: rc = launch_actor_ni(L); break;
    case RCG_SYS_3WROBOT:    rc = launch_actor_3w(L); break;
    default:                 rc = launch_actor_2t(L); break;
    }
    if (rc) { set_error("%s: bad mode/critic_struct", what); return rc; }
    return check_launch(what);
}

}  // namespace rcg

extern "C" {

const char *rcg_last_actor_kernel(void) { return rcg::g_last_actor_kernel; }

int rcg_actor_cost(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C, const double *state_sys,
                   const double *obs, const double *cand, int32_t cand_per_env, const double *w_critic,
                   int32_t w_per_env, const int32_t *mask, double *J_out, int32_t *argmin_out, double *Jmin_out,
                   double *action_out, double *accum, double sampling_time, void *stream)
{
    return rcg::launch_actor<double>("rcg_actor_cost", sys, obj, E, C, state_sys, obs, cand, cand_per_env, w_critic,
                                     w_per_env, mask, J_out, argmin_out, Jmin_out, action_out, accum, sampling_time,
                                     stream);
}

int rcg_actor_cost_f32(const rcg_system_t *sys, const rcg_objective_t *obj, int64_t E, int32_t C,
                       const float *state_sys, const float *obs, const float *cand, int32_t cand_per_env,
                       const float *w_critic, int32_t w_per_env, const int32_t *mask, float *J_out,
                       int32_t *argmin_out, float *Jmin_out, float *action_out, float *accum, double sampling_time,
                       void *stream)
{
    return rcg::launch_actor<float>("rcg_actor_cost_f32", sys, obj, E, C, state_sys, obs, cand, cand_per_env, w_critic,
                                    w_per_env, mask, J_out, argmin_out, Jmin_out, action_out, accum, sampling_time,
                                    stream);
}

}  // extern "C"
