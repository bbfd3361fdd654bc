"""Diagnostic: throughput of the batched IK kernels on one GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg

pm = pkg.PandaModel()
lb, ub = pm.getJointLimit().T
rng = np.random.default_rng(0)
n = 200_000
q_true = lb + (ub - lb) * rng.uniform(0.08, 0.92, (n, 7))
T = torch.from_numpy(pm.getTransform(q_true)[:, :3, :].copy()).cuda()
seeds = torch.from_numpy(np.clip(q_true + 0.4 * rng.standard_normal(q_true.shape), lb, ub)).cuda()
for name, fn in (("ikBatch (1 solve per target, seed 0.4 rad away)", lambda: pm.ikBatch(T, seeds)),
                 ("ikSampleBatch (15 restarts per target, N(mid, 0.3) seeds)", lambda: pm.ikSampleBatch(T, restarts=15, rng_seed=1))):
    r = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    extra = f", mean iterations {float(r['iters'].float().mean()):.1f}" if "iters" in r else f", mean successes {float(r['n_success'].float().mean()):.1f}/15"
    print(f"{name}: {n} targets in {dt*1e3:.2f} ms = {n/dt/1e6:.2f} M targets/s, success {float(r['ok'].float().mean()):.4f}{extra}")
