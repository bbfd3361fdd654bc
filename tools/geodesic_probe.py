"""Diagnostic: batched discreteGeodesic latency against batch size (edges between projected samples)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg

c = pkg.KinematicChainConstraint.from_config("dumbbell")
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
smp = space.allocStateSampler(pool_size=1_200_000, rng_seed=1)
pts = smp.sampleUniformBatch(1_200_000)
# the pool's row order is whatever the compaction atomics made it: sort it, then shuffle with a fixed seed, so that two
# runs (two builds) walk the SAME edges
import numpy as np

pn = pts.cpu().numpy()
pn = pn[np.lexsort(pn.T[::-1])]
pn = pn[np.random.default_rng(0).permutation(len(pn))]
pts = torch.from_numpy(pn).cuda()
E = min(100_000, pts.shape[0] // 2)
frm, to = pts[:E].contiguous(), pts[E:2 * E].contiguous()
import itertools

for edges, thr in itertools.product((5, 100, 1000, 5000, 20000, E), (0, 1 << 30)):
    c._lib.ccp_set_coop_threshold(c._h, thr)
    print("two lanes per edge " if thr else "one thread per edge", end=" ")
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = space.discreteGeodesicBatch(frm[:edges], to[:edges], max_states=40)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{edges:7d} edges: {best*1e3:8.3f} ms  {edges/best/1e3:9.1f} k edges/s  reached {float(r.reached.float().mean()):.3f}  "
          f"mean states {float(r.n_states.float().mean()):.1f}  mean Newton iterations per edge {float(r.iters.float().mean()):.0f}")
