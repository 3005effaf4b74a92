// rcg_host.h -- host-side helpers of the C ABI: error reporting, launch accounting and the
// conversion of the public descriptors (include/rcg.h) into kernel-parameter structs.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "rcg_device.cuh"

namespace rcg {

void set_error(const char *fmt, ...);
int  check_launch(const char *what);          // cudaGetLastError() after a launch, counts it
int  require_device();                        // 0, or RCG_ENODEV with the error string set

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: `done` (a static of the calling
// instantiation) keeps one bit per device so that a second GPU in the same process gets it too.
template <class K>
inline void ensure_dyn_smem(K kern, size_t smem, unsigned long long &done)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(done & bit)) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        done |= bit;
    }
}

inline bool is_diag(const double *R, int p)
{
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j)
            if (i != j && R[i * p + j] != 0.0) return false;
    return true;
}

template <typename T>
inline SysDev<T> make_sys_dev(const rcg_system_t *s)
{
    SysDev<T> d;
    for (int i = 0; i < 8; ++i) d.pars[i] = (T)s->pars[i];
    for (int i = 0; i < RCG_MAX_M; ++i) { d.lo[i] = (T)s->lo[i]; d.hi[i] = (T)s->hi[i]; }
    d.has_bnds = s->has_bnds;
    return d;
}

template <typename T>
inline ObjDev<T> make_obj_dev(const rcg_objective_t *o, int n, int m)
{
    ObjDev<T> d;
    const int p = n + m;
    d.stage_struct = o->stage_struct;
    d.has_target = o->has_target;
    d.Nactor = o->Nactor;
    d.Ncritic = o->Ncritic;
    d.buffer_size = o->buffer_size;
    d.gamma = (T)o->gamma;
    d.pred_step_size = (T)o->pred_step_size;
    for (int k = 0; k < RCG_MAX_NACTOR; ++k) d.gamma_pow[k] = (T)o->gamma_pow[k];
    // the public struct stores R row-major with leading dimension p; the device struct too
    for (int i = 0; i < RCG_MAX_P * RCG_MAX_P; ++i) {
        d.R1[i] = (i < p * p) ? (T)o->R1[i] : T(0);
        d.R2[i] = (i < p * p) ? (T)o->R2[i] : T(0);
    }
    for (int i = 0; i < RCG_MAX_N; ++i) d.target[i] = (o->has_target && i < n) ? (T)o->target[i] : T(0);
    return d;
}

inline int sys_n(int sys_id) { return sys_id == RCG_SYS_3WROBOT_NI ? 3 : sys_id == RCG_SYS_3WROBOT ? 5 : sys_id == RCG_SYS_2TANK ? 2 : -1; }
inline int sys_m(int sys_id) { return sys_id == RCG_SYS_2TANK ? 1 : (sys_id == RCG_SYS_3WROBOT_NI || sys_id == RCG_SYS_3WROBOT) ? 2 : -1; }

}  // namespace rcg

#define RCG_REQUIRE(cond, ...)                    \
    do {                                          \
        if (!(cond)) {                            \
            rcg::set_error(__VA_ARGS__);          \
            return RCG_EINVAL;                    \
        }                                         \
    } while (0)
