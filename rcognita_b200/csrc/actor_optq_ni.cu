// actor_optq_ni.cu -- instantiates the lanes-per-problem actor-optimiser kernels of actor_opt_quad.cuh for one system.
#include "actor_opt_quad.cuh"

namespace rcg {
int launch_optq_ni(const OptLaunch<double> &L) { return launch_optq_sys<RCG_SYS_3WROBOT_NI>(L); }
}  // namespace rcg
