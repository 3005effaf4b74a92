// actor_2t_f32.cu -- instantiates the actor-cost kernels of actor_impl.cuh for one (system, dtype).
#include "actor_impl.cuh"

namespace rcg {
int launch_actor_2t(const ActorLaunch<float> &L) { return launch_actor_sys<float, RCG_SYS_2TANK>(L); }
}  // namespace rcg
