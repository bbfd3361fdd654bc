"""The reference's constraint surface, batched, on one B200 (the snippet of README.md).  Run on a machine with the GPU:
python examples/quickstart.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import closed_chain_motion_planner_b200 as ccp

c = ccp.KinematicChainConstraint.from_config("dumbbell")  # two Panda arms, bases and start_joint from configs/dumbbell.yaml
x = c.config.start + 0.05 * np.random.default_rng(0).standard_normal(14)
ok = c.project(x)  # ConstraintFunction.h:57 — in place, one state
print("project():", ok, "residual", c.function(x))
r = c.projectBatch(np.random.default_rng(1).uniform(-2, 2, (100_000, 14)))  # r.x, r.ok, r.converged, r.iters, r.resid
print("projectBatch: ok fraction", r.ok.mean(), "mean iterations", r.iters.mean())
space = ccp.jy_ProjectedStateSpace(ccp.KinematicChainSpace(14), c)
V = space.allocStateSampler(pool_size=1 << 16).sampleUniformBatch(50_000)  # projected, wrapped, ok states (device tensor)
g = space.discreteGeodesicBatch(V[:1000].contiguous(), V[1000:2000].contiguous(), max_states=40)  # g.reached, g.n_states, g.states
print("sampler:", tuple(V.shape), "geodesics reached", float(g.reached.float().mean()), "mean states", float(g.n_states.float().mean()))
