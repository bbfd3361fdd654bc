"""Diagnostic (not part of the product): per-iteration and per-sample cost of the projection kernel, from runs
of the same 4M seeds with the iteration cap set to different values (time = a * iterations + b * samples)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

cfg = sys.argv[1] if len(sys.argv) > 1 else "dumbbell"
c = pkg.KinematicChainConstraint.from_config(cfg)
lib, h = c._lib, c._h
n = c.getAmbientDimension()
N = 4_000_000
d = torch.empty((N, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
st0 = torch.cuda.current_stream().cuda_stream
assert lib.ccp_generate_seeds(h, C.byref(a), N, 0, d.data_ptr(), st0) == 0
x_out = torch.empty_like(d)
ok = torch.empty(N, dtype=torch.uint8, device="cuda")
it = torch.empty(N, dtype=torch.int32, device="cuda")
rows = []
for cap in (0, 2, 5, 10, 20, 40, 80, 250):
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
    rows.append((cap, mi, best / (N / 1e6)))
    print(f"cap {cap:3d}: mean iterations {mi:7.3f}  {best / (N / 1e6):7.4f} ms per 1M samples")
A = np.array([[r[1], 1.0] for r in rows[:-1]])
y = np.array([r[2] for r in rows[:-1]])
(a_, b_), *_ = np.linalg.lstsq(A, y, rcond=None)
print(f"fit over caps 0..80: {a_*1e3:.3f} us per 1k iterations... a = {a_:.5f} ms per (1M samples x iteration), b = {b_:.4f} ms per 1M samples = {b_/a_:.2f} iterations' worth")
