"""Small pass over every kernel, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA

for cfgname in ("dumbbell", "stefan_three_arm"):
    c = pkg.KinematicChainConstraint.from_config(cfgname)
    n = c.getAmbientDimension()
    A = OracleA(c.config.arm_indices)
    s = A.seeds_uniform(0, 0, 9001)
    xa = torch.from_numpy(s).cuda()
    xs = torch.from_numpy(np.ascontiguousarray(s.T)).cuda()
    c._lib.ccp_set_coop_threshold(c._h, 0)  # the thread-per-sample kernel ...
    r = c.projectBatch(xa)
    r2 = c.projectBatch(xs, layout=pkg.CCP_LAYOUT_SOA)
    c._lib.ccp_set_coop_threshold(c._h, 1 << 30)  # ... and the cooperative one (two arms only; three arms fall through)
    rc = c.projectBatch(xa)
    rc2 = c.projectBatch(xs, layout=pkg.CCP_LAYOUT_SOA)
    torch.cuda.synchronize()
    assert torch.equal(rc.x, r.x) and torch.equal(rc2.x, r2.x) and torch.equal(rc.iters, r.iters)
    c._lib.ccp_set_coop_threshold(c._h, -1)
    compact = torch.zeros((3 * 9001, n), dtype=torch.float64, device="cuda")
    n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
    ps = [c.projectBatch(xa[i * 3000:(i + 1) * 3000].contiguous(), compact=compact, n_ok=n_ok, pipelined=True) for i in range(3)]
    c.flush(compact=compact, n_ok=n_ok)
    f = c.functionBatch(xa[:100].contiguous())
    J = c.jacobianBatch(xa[:100].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([p.iters for p in ps]), r.iters[:9000])
    print(cfgname, "ok", int(r.ok.sum()), int(n_ok))
c = pkg.KinematicChainConstraint.from_config("dumbbell")
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
smp = space.allocStateSampler(pool_size=4096, rng_seed=1)
pts = smp.sampleUniformBatch(4000)
pts = pts if isinstance(pts, torch.Tensor) else torch.from_numpy(np.asarray(pts)).cuda()
E = pts.shape[0] // 2
res = space.discreteGeodesicBatch(pts[:E].contiguous(), pts[E:2 * E].contiguous(), max_states=32)
pm = pkg.PandaModel()
lb, ub = pm.getJointLimit().T
q = lb + (ub - lb) * np.random.default_rng(0).uniform(0.1, 0.9, (300, 7))
T = pm.getTransform(q)[:, :3, :]
ik = pm.ikSampleBatch(T, restarts=15, rng_seed=2)
ik1 = pm.ikBatch(T, q)
torch.cuda.synchronize()
print("geodesic/ik ok", float(ik["ok"].mean()))
# streaming host batches: full outputs, compact outputs, device-generated seeds
from closed_chain_motion_planner_b200 import _capi

A = OracleA(c.config.arm_indices)
hb = [A.seeds_uniform(3, b * 250_000, 250_000) for b in range(3)]
pend = [c.submitHostBatch(x, pinned=True) for x in hb[:2]]
c.waitHostBatch(pend[0][0])
pend.append(c.submitCompactBatch(hb[2], want_flags=True))
c.waitHostBatch(pend[1][0])
nk = c.waitCompactBatch(*pend[2])
sa = _capi.SamplerArgs(rng_seed=5, first_index=0, mode=0, wrap_bounds=1, distance=0.0, near_host=None)
t, r = c.submitCompactBatch(sampler=sa, count=120_000)
nk2 = c.waitCompactBatch(t, r)
print("host batches ok", nk, nk2)
