"""Diagnostic (not part of the product): time of ONE Newton trip against the number of resident warps.
Tolerances are set so small that every sample runs exactly max_iter = 250 iterations; count selects how many
warps carry samples (the launcher spreads a small batch one warp per block)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

c = pkg.KinematicChainConstraint.from_config("dumbbell")
c.setTolerance(1e-300, 1e-300)
lib, h = c._lib, c._h
n = 14
NMAX = 148 * 384
d = torch.empty((NMAX, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
st0 = torch.cuda.current_stream().cuda_stream
assert lib.ccp_generate_seeds(h, C.byref(a), NMAX, 0, d.data_ptr(), st0) == 0
x_out = torch.empty_like(d)
it = torch.empty(NMAX, dtype=torch.int32, device="cuda")
for cnt in (1, 32, 64, 148 * 32, 148 * 64, 148 * 128, 148 * 256, 148 * 384, 148 * 384 * 2):
    if cnt > NMAX:
        continue
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert lib.ccp_project_batch(h, d.data_ptr(), cnt, 0, x_out.data_ptr(), None, None, it.data_ptr(), None, None, None, st0) == 0
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    trips = 251
    print(f"count {cnt:6d} ({cnt / 148 / 32:5.2f} warps/SM): {best:7.3f} ms = {best * 1e3 / trips:6.3f} us/trip = "
          f"{best * 1e-3 / trips * 1.965e9:7.0f} clk/trip   [iters min {int(it[:cnt].min())} max {int(it[:cnt].max())}]")
