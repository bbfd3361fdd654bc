import os, sys
sys.path.insert(0, '.')
import torch
import closed_chain_motion_planner_b200 as pkg
c = pkg.KinematicChainConstraint.from_config("stefan_three_arm")
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(21), c)
smp = space.allocStateSampler(pool_size=1 << 20, rng_seed=1)
pts = smp.sampleUniformBatch(3_000_000)
E = min(50_000, pts.shape[0] // 2)
frm, to = pts[:E].contiguous(), pts[E:2 * E].contiguous()
for edges in (1000, E):
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = space.discreteGeodesicBatch(frm[:edges], to[:edges], max_states=40); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(os.environ.get("CCP_LIB", "")[-14:], edges, "edges", round(best, 3), "ms reached", float(r.reached.float().mean()))
