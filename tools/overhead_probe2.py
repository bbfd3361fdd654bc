"""Diagnostic: cost per iteration when every sample runs exactly `cap` iterations (tolerance unreachable: all lanes of
a warp finish together) against the real mix (reference tolerances: lanes finish at different trips)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

c = pkg.KinematicChainConstraint.from_config("dumbbell")
lib, h = c._lib, c._h
n = 14
N = 2_000_000
d = torch.empty((N, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
st0 = torch.cuda.current_stream().cuda_stream
assert lib.ccp_generate_seeds(h, C.byref(a), N, 0, d.data_ptr(), st0) == 0
x_out = torch.empty_like(d)
ok = torch.empty(N, dtype=torch.uint8, device="cuda")
it = torch.empty(N, dtype=torch.int32, device="cuda")
for tol, label in ((1e-300, "lockstep (tolerance unreachable)"), (None, "real mix")):
    if tol is None:
        c.setTolerance(1e-3, 5e-3)
    else:
        c.setTolerance(tol, tol)
    for cap in (10, 20, 40, 60, 80):
        c.setOptions(0.30, cap, 1e-3)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert lib.ccp_project_batch(h, d.data_ptr(), N, 0, x_out.data_ptr(), ok.data_ptr(), None, it.data_ptr(), None, None, None, st0) == 0
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        mi = float(it.sum(dtype=torch.int64)) / N
        per = best / (N / 1e6)
        print(f"{label:34s} cap {cap:3d}: mean iterations {mi:7.3f}  {per:7.4f} ms per 1M  ->  {(per - 0.095) / mi * 1e3:6.2f} us per 1M iterations")
