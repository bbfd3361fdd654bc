"""Diagnostic: latency of small host-buffer calls (what a planner that projects one state at a time sees)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA

c = pkg.KinematicChainConstraint.from_config("dumbbell")
A = OracleA(c.config.arm_indices)
A.set_initial_position(c.config.start)
seeds = A.seeds_uniform(0, 0, 4096)
near = c.config.start[None, :] + 0.05 * np.random.default_rng(0).standard_normal((4096, 14))
for name, S in (("uniform seeds", seeds), ("seeds 0.05 rad from the manifold", near)):
    for cnt in (1, 16, 256, 4096):
        reps = 200 if cnt <= 256 else 50
        c.projectBatch(S[:cnt])
        t0 = time.perf_counter()
        its = 0
        for r in range(reps):
            res = c.projectBatch(S[(r * cnt) % (4096 - cnt + 1):][:cnt], want_resid=False) if cnt < 4096 else c.projectBatch(S, want_resid=False)
            its += int(res.iters.max())
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:34s} batch {cnt:5d}: {dt*1e6:8.1f} us per call  (max iterations per call, mean {its/reps:.0f})")
t0 = time.perf_counter()
r = A.project(near[:64], nthreads=1)
print(f"oracle A, one thread: {1e6*(time.perf_counter()-t0)/64:.0f} us per projection (near-manifold seeds)")
x1 = near[:1].copy()
for name, fn in (("function()", lambda: c.function(x1[0])), ("jacobian()", lambda: c.jacobian(x1[0])), ("isSatisfied()", lambda: c.isSatisfied(x1[0]))):
    fn()
    t0 = time.perf_counter()
    for _ in range(300):
        fn()
    print(f"single-state {name:14s}: {(time.perf_counter() - t0) / 300 * 1e6:6.1f} us per call")
