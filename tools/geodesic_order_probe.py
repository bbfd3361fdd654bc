"""Does longest-edge-first ordering shorten the geodesic kernel's tail?  (experiment; host-side reordering)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import closed_chain_motion_planner_b200 as pkg

c = pkg.KinematicChainConstraint.from_config("dumbbell")
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
smp = space.allocStateSampler(pool_size=1 << 17, rng_seed=3)
nv, knn = 20_000, 5
V = smp.sampleUniformBatch(120_000)[:nv].contiguous()
dm = torch.cdist(V, V); dm.fill_diagonal_(float("inf"))
nbr = dm.topk(knn, largest=False).indices
frm, to = V.repeat_interleave(knn, dim=0).contiguous(), V[nbr.reshape(-1)].contiguous()
def run(f, t):
    best = 1e9
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = space.discreteGeodesicBatch(f, t, max_states=40); e1.record(); torch.cuda.synchronize()
        if i: best = min(best, e0.elapsed_time(e1))
    return best, r
ms, r = run(frm, to)
trips = r.iters + r.n_states  # Newton trips + one bookkeeping trip per state
print("as given        %.3f ms; trips per edge mean %.1f max %d; states max %d" % (ms, trips.float().mean(), trips.max(), r.n_states.max()))
d = (frm - to).norm(dim=1)
o = torch.argsort(d, descending=True)
ms2, _ = run(frm[o].contiguous(), to[o].contiguous())
print("by distance desc %.3f ms" % ms2)
o = torch.argsort(trips, descending=True)
ms3, _ = run(frm[o].contiguous(), to[o].contiguous())
print("by true trips desc (oracle ordering) %.3f ms" % ms3)
o = torch.randperm(len(d), device="cuda")
ms4, _ = run(frm[o].contiguous(), to[o].contiguous())
print("random order %.3f ms" % ms4)
