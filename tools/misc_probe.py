"""Diagnostic: (1) calibrated-DH (generic link code) against stock-Panda (structured) projection time,
(2) batched discreteGeodesic throughput."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

N = 2_000_000
cfg = pkg.grasping_point().loadConfig("dumbbell")


def build(dh):
    c = pkg.KinematicChainConstraint(14)
    c.setArmModels(*[pkg.ArmModel(name=nm, index=ix, t_wb=cfg.t_wb[ix], dh_offsets=dh) for nm, ix in zip(cfg.arm_names, cfg.arm_indices)])
    c.setInitialPosition(cfg.start)
    c.setTolerance(1e-3, 5e-3)
    return c


rng = np.random.default_rng(0)
dh = 1e-3 * rng.standard_normal((7, 4))
dh_noalpha = dh.copy()
dh_noalpha[:, 3] = 0.0
for name, d in (("stock Panda (structured alpha)", None), ("calibrated a, d, theta (alpha stock: structured)", dh_noalpha),
                ("calibrated a, d, theta, alpha (generic link code)", dh)):
    c = build(d)
    seeds = torch.empty((N, 14), dtype=torch.float64, device="cuda")
    a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    st = torch.cuda.current_stream().cuda_stream
    assert c._lib.ccp_generate_seeds(c._h, C.byref(a), N, 0, seeds.data_ptr(), st) == 0
    out = torch.empty_like(seeds)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = c.projectBatch(seeds, out=out, want_resid=False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {best / (N / 1e6):.3f} ms per 1M seeds, ok {float(r.ok.float().mean()):.4f}, mean iterations {float(r.iters.float().mean()):.2f}")

# geodesic: edges between projected samples 1.0 apart
c = pkg.KinematicChainConstraint.from_config("dumbbell")
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
smp = space.allocStateSampler(pool_size=400_000, rng_seed=1)
pts = smp.sampleUniformBatch(200_000)
pts = pts if isinstance(pts, torch.Tensor) else torch.from_numpy(np.asarray(pts)).cuda()
E = pts.shape[0] // 2
frm, to = pts[:E].contiguous(), pts[E:2 * E].contiguous()
# shorten: target 1.5 rad away along the chord
d = to - frm
to = frm + d * (1.5 / d.norm(dim=1, keepdim=True)).clamp(max=1.0)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter()
    res = space.discreteGeodesicBatch(frm, to, max_states=16)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
ns = res.n_states if isinstance(res.n_states, torch.Tensor) else torch.as_tensor(res.n_states)
rc = res.reached if isinstance(res.reached, torch.Tensor) else torch.as_tensor(res.reached)
print(f"discreteGeodesicBatch: {E} edges (1.5 rad chords, delta 0.25) in {dt*1e3:.2f} ms = {E/dt/1e6:.2f} M edges/s, "
      f"reached {float(rc.float().mean()):.3f}, mean states {float(ns.float().mean()):.2f}")
