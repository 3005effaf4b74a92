#!/usr/bin/env python3
"""Preset: 3wrobotNI on the B200 engine -- same flags as the reference's presets/main_3wrobot_NI.py (plus --num_envs,
--num_candidates, --seed, --state_spread).  Example:
    python presets/main_3wrobot_NI.py --ctrl_mode MPC --Nactor 6 --is_visualization '' --is_print_sim_step '' --is_log_data 1
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcognita_b200 import presets  # noqa: E402

if __name__ == "__main__":
    presets.main("3wrobotNI")
