// actor_opt_2t.cu -- instantiates the actor-optimiser kernels of actor_opt_impl.cuh for one system.
#include "actor_opt_impl.cuh"

namespace rcg {
int launch_opt_2t(const OptLaunch<double> &L) { return launch_opt_sys<double, RCG_SYS_2TANK>(L); }
}  // namespace rcg
