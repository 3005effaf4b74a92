// actor_2t_f64.cu -- instantiates the actor-cost kernels of actor_impl.cuh for one (system, dtype).
#include "actor_impl.cuh"

namespace rcg {
int launch_actor_2t(const ActorLaunch<double> &L) { return launch_actor_sys<double, RCG_SYS_2TANK>(L); }
}  // namespace rcg
