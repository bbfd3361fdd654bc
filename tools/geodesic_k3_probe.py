"""Diagnostic: the three-arm discreteGeodesic walk, one thread per edge against four lanes per edge
(ccp_geodesic_coop3_kernel), on the same deterministic edge set.  usage: python tools/geodesic_k3_probe.py [config]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg

name = sys.argv[1] if len(sys.argv) > 1 else "stefan_three_arm"
c = pkg.KinematicChainConstraint.from_config(name)
n = c.getAmbientDimension()
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(n), c)
smp = space.allocStateSampler(pool_size=1 << 20, rng_seed=1)
pts = smp.sampleUniformBatch(10_000_000 if n == 21 else 2_400_000).cpu().numpy()
pts = pts[np.lexsort(pts.T[::-1])]  # the pool's row order depends on the compaction atomics: make it deterministic
pts = torch.from_numpy(pts[np.random.default_rng(0).permutation(len(pts))]).cuda()
E = min(200_000, pts.shape[0] // 2)
frm, to = pts[:E].contiguous(), pts[E:2 * E].contiguous()
for edges in [int(v) for v in os.environ.get("GEO_PROBE_EDGES", "5,100,1000,5000,10000,20000,30000").split(",")] + [E]:
    row = {}
    for kname, thr in (("thread_per_edge", 0), ("cooperative", 1 << 30)):
        c._lib.ccp_set_coop_threshold(c._h, thr)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = space.discreteGeodesicBatch(frm[:edges], to[:edges], max_states=40)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        row[kname] = round(best, 3)
    print(edges, "edges:", row, "x%.2f" % (row["thread_per_edge"] / row["cooperative"]), "reached", round(float(r.reached.float().mean()), 4),
          "longest edge", int(r.iters.max()), "trips")
