"""Diagnostic (not part of the product): time the CCP_PROJ_VARIANT launch configurations (make TUNE=1).
usage: CCP_PROJ_VARIANT=v python tools/variant_probe.py [config] [counts...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

cfgname = sys.argv[1] if len(sys.argv) > 1 else "dumbbell"
counts = [int(v) for v in sys.argv[2:]] or [56832, 1_000_000, 4_000_000]
c = pkg.KinematicChainConstraint.from_config(cfgname)
lib, h = c._lib, c._h
n = c.getAmbientDimension()
NMAX = max(counts)
d = torch.empty((NMAX, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
assert lib.ccp_generate_seeds(h, C.byref(a), NMAX, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
x_out = torch.empty_like(d)
ok = torch.empty(NMAX, dtype=torch.uint8, device="cuda")
it = torch.empty(NMAX, dtype=torch.int32, device="cuda")
st0 = torch.cuda.current_stream().cuda_stream


def launch(cnt):
    rc = lib.ccp_project_batch(h, d.data_ptr(), cnt, 0, x_out.data_ptr(), ok.data_ptr(), None, it.data_ptr(), None, None, None, st0)
    assert rc == 0, lib.ccp_last_error(h)


launch(counts[0])
torch.cuda.synchronize()
out = []
for cnt in counts:
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch(cnt)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out.append(f"{cnt}: {best:.3f} ms")
chk = int(it[: counts[0]].sum().item())
print(f"variant {os.environ.get('CCP_PROJ_VARIANT', '0')}: " + "  ".join(out) + f"  [iters checksum {chk}]")
