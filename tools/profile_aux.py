"""Runs the SURVEY §8(f) kernels once each at the sizes bench.py's aux_kernels record uses (under ncu: see
profiles/README).  usage: python tools/profile_aux.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import closed_chain_motion_planner_b200 as pkg  # noqa: E402
from bench import aux_kernels  # noqa: E402

c = pkg.KinematicChainConstraint.from_config("dumbbell")
peak, _ = c.fp64PeakProbe(3)
print(json.dumps(aux_kernels(pkg, 0, peak), indent=1))
