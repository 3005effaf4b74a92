// Test harness (NOT part of the product): compiles the __host__ __device__ core of the iLQR pre-pass
// (rcognita_b200/csrc/actor_ilqr_core.cuh) for the host, so that tests/test_ilqr_core_host.py can check its decisions
// against the CPU checker without a GPU.  Built by the test with `nvcc -shared`; nothing in rcognita_b200/ loads it.
#include "../../rcognita_b200/csrc/actor_ilqr_core.cuh"
#include "../../rcognita_b200/csrc/rcg_host.h"

#include <vector>

extern "C" int ilqr_host(const rcg_system_t *sys, const rcg_objective_t *obj, const double *x0, const double *ob0,
                         const double *w, double *U, int max_sweeps, double pg_tol)
{
    const int n = rcg::sys_n(sys->sys_id), m = rcg::sys_m(sys->sys_id);
    if (n <= 0) return -1;
    const rcg::SysDev<double> S = rcg::make_sys_dev<double>(sys);
    const rcg::ObjDev<double> O = rcg::make_obj_dev<double>(obj, n, m);
    std::vector<double> ws((size_t)rcg::ilqr_ws_per_problem(obj->Nactor, n, m));
    switch (sys->sys_id) {
    case RCG_SYS_3WROBOT_NI:
        return rcg::ilqr_presweeps<RCG_SYS_3WROBOT_NI>(S, O, obj->mode, obj->critic_struct, x0, ob0, w, U, 1, ws.data(), 1, max_sweeps, pg_tol);
    case RCG_SYS_3WROBOT:
        return rcg::ilqr_presweeps<RCG_SYS_3WROBOT>(S, O, obj->mode, obj->critic_struct, x0, ob0, w, U, 1, ws.data(), 1, max_sweeps, pg_tol);
    default:
        return rcg::ilqr_presweeps<RCG_SYS_2TANK>(S, O, obj->mode, obj->critic_struct, x0, ob0, w, U, 1, ws.data(), 1, max_sweeps, pg_tol);
    }
}
