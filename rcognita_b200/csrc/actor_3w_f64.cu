// actor_3w_f64.cu -- instantiates the actor-cost kernels of actor_impl.cuh for one (system, dtype).
#include "actor_impl.cuh"

namespace rcg {
int launch_actor_3w(const ActorLaunch<double> &L) { return launch_actor_sys<double, RCG_SYS_3WROBOT>(L); }
}  // namespace rcg
