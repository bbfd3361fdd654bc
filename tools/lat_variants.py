import os, sys, json
sys.path.insert(0, '.')
import torch
import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA
c = pkg.KinematicChainConstraint.from_config("dumbbell")
A = OracleA(c.config.arm_indices)
c._lib.ccp_set_coop_threshold(c._h, 0)
for count in (16, 1000, 4000):
    seeds = torch.from_numpy(A.seeds_uniform(0, 0, count)).cuda(); out = torch.empty_like(seeds)
    best = 1e9
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.projectBatch(seeds, out=out, want_resid=False); e1.record(); torch.cuda.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    print(os.environ.get("CCP_PROJ_VARIANT"), count, round(best, 4))
