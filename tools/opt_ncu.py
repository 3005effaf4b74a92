#!/usr/bin/env python3
"""One rcg_actor_opt launch per configuration for an ncu capture (profiles/): 3wrobot_NI MPC Nactor=6 near the
goal (interior minimisers, ~17 iterations) and 3wrobot RQL 'quadratic' Nactor=10."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from kernel_bench import actor_opt_point  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ni"
if which == "ni":
    print(actor_opt_point("3wrobotNI", "MPC", "quad-nomix", 6, 65536, 0.05, iters=1))
else:
    print(actor_opt_point("3wrobot", "RQL", "quadratic", 10, 65536, 1.0, iters=1))
torch.cuda.synchronize()
