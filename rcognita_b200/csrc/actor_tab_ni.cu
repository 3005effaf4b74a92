// actor_tab_ni.cu -- instantiates the shared-table actor kernels of actor_tab.cuh (candidate part of the heading tabulated once) for one system.
#include "actor_tab.cuh"

namespace rcg {
int launch_actor_tab_ni(const ActorLaunch<double> &L, void *scratch) { return launch_actor_tab_sys<RCG_SYS_3WROBOT_NI>(L, scratch); }
}  // namespace rcg
