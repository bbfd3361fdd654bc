"""Diagnostic: host-buffer projection throughput, synchronous call against the streaming submit/wait form.

usage: python tools/e2e_stream.py [count] [batches] [--pageable]     (CCP_HOST_LAG / CCP_HOST_CHUNKS are read by the library)
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

pageable = "--pageable" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
count = int(argv[0]) if len(argv) > 0 else 1_000_000
nb = int(argv[1]) if len(argv) > 1 else 12
c = pkg.KinematicChainConstraint.from_config("dumbbell")
lib, h = c._lib, c._h
n = c.getAmbientDimension()
d = torch.empty((count, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
assert lib.ccp_generate_seeds(h, C.byref(a), count, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
pin = (lambda t: t) if pageable else (lambda t: t.pin_memory())
seeds = pin(d.cpu())
R = 3
xo = [pin(torch.zeros((count, n), dtype=torch.float64)) for _ in range(R)]
ok = [pin(torch.zeros(count, dtype=torch.uint8)) for _ in range(R)]
it = [pin(torch.zeros(count, dtype=torch.int32)) for _ in range(R)]
print("host buffers:", "pageable" if pageable else "page-locked")


def sync_call(b):
    r = b % R
    assert lib.ccp_project_batch_host(h, seeds.data_ptr(), count, xo[r].data_ptr(), ok[r].data_ptr(), None,
                                      it[r].data_ptr(), None) == 0


for _ in range(3):
    sync_call(0)
t0 = time.perf_counter()
for b in range(nb):
    sync_call(b)
t1 = time.perf_counter()
ms = (t1 - t0) / nb * 1e3
print(f"lag={os.environ.get('CCP_HOST_LAG', 'default')} synchronous: {ms:.3f} ms per {count} = {count / ms / 1e3:.1f} M projections/s")
nok_sync = int(ok[(nb - 1) % R].numpy().sum())

tick = C.c_int64(0)
pend = []
t0 = time.perf_counter()
for b in range(nb):
    r = b % R
    assert lib.ccp_project_batch_host_submit(h, seeds.data_ptr(), count, xo[r].data_ptr(), ok[r].data_ptr(), None,
                                             it[r].data_ptr(), None, C.byref(tick)) == 0
    pend.append(tick.value)
    if len(pend) == 2:
        assert lib.ccp_project_batch_host_wait(h, pend.pop(0)) == 0
while pend:
    assert lib.ccp_project_batch_host_wait(h, pend.pop(0)) == 0
t1 = time.perf_counter()
ms = (t1 - t0) / nb * 1e3
print(f"lag={os.environ.get('CCP_HOST_LAG', 'default')} streaming (2 in flight): {ms:.3f} ms per {count} = {count / ms / 1e3:.1f} M projections/s")
assert int(ok[(nb - 1) % R].numpy().sum()) == nok_sync
