"""Diagnostic (not part of the product): wall time of ccp_project_batch_host (pinned buffers) for the chunk count
in CCP_HOST_CHUNKS (read once per process; unset = the library's own choice)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA  # seeds only

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
c = pkg.KinematicChainConstraint.from_config("dumbbell")
A = OracleA(c.config.arm_indices)
hs = torch.from_numpy(A.seeds_uniform(0, 0, n)).pin_memory()
hx = torch.empty((n, 14), dtype=torch.float64).pin_memory()
hok = torch.empty(n, dtype=torch.uint8).pin_memory()
hit = torch.empty(n, dtype=torch.int32).pin_memory()
lib, h = c._lib, c._h
best = 1e9
for _ in range(6):
    t0 = time.perf_counter()
    assert lib.ccp_project_batch_host(h, hs.data_ptr(), n, hx.data_ptr(), hok.data_ptr(), None, hit.data_ptr(), None) == 0
    best = min(best, time.perf_counter() - t0)
print(f"chunks {os.environ.get('CCP_HOST_CHUNKS', 'auto'):>4}: {best * 1e3:.3f} ms  {n / best / 1e6:.1f} M projections/s  "
      f"[iters sum {int(hit.sum())}, ok {int(hok.sum())}]")
