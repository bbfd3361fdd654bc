"""Diagnostic (not part of the product): how much of a projection launch is its tail?

  (a) kernel time against batch size on one stream: the slope is the steady-state rate, the intercept the
      launch tail (the last samples of a batch run up to 250 iterations while most lanes have nothing left);
  (b) the same batches issued round-robin on 1..4 streams: the next batch's blocks fill the SMs the previous
      batch's tail leaves idle.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

cfgname = sys.argv[1] if len(sys.argv) > 1 else "dumbbell"
c = pkg.KinematicChainConstraint.from_config(cfgname)
lib, h = c._lib, c._h
n = c.getAmbientDimension()
NMAX = 8_000_000
d = torch.empty((NMAX, n), dtype=torch.float64, device="cuda")
a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
assert lib.ccp_generate_seeds(h, C.byref(a), NMAX, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
x_out = torch.empty_like(d)
ok = torch.empty(NMAX, dtype=torch.uint8, device="cuda")
it = torch.empty(NMAX, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()


def launch(off, cnt, st):
    rc = lib.ccp_project_batch(h, d.data_ptr() + off * n * 8, cnt, 0, x_out.data_ptr() + off * n * 8, ok.data_ptr() + off, None,
                               it.data_ptr() + off * 4, None, None, None, st)
    assert rc == 0


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


st0 = torch.cuda.current_stream().cuda_stream
launch(0, 1_000_000, st0)
torch.cuda.synchronize()
print("(a) one launch, one stream")
for cnt in (56832, 125_000, 250_000, 500_000, 1_000_000, 2_000_000, 4_000_000, 8_000_000):
    ms = timed(lambda: launch(0, cnt, st0))
    print(f"  count {cnt:>9}: {ms:8.3f} ms   {cnt / ms / 1e3:7.1f} M projections/s   {ms / cnt * 1e6:6.3f} ms per 1M")

print("(b) 8 batches of 1M, round-robin over S streams (events on the legacy stream bracket all of them)")
for ns in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]

    def run():
        cur = torch.cuda.current_stream()
        e = torch.cuda.Event()
        e.record(cur)
        for s in streams:
            s.wait_event(e)
        for b in range(8):
            launch(b * 1_000_000, 1_000_000, streams[b % ns].cuda_stream)
        for s in streams:
            e2 = torch.cuda.Event()
            e2.record(s)
            cur.wait_event(e2)

    ms = timed(run)
    print(f"  {ns} stream(s): {ms:8.3f} ms total   {ms / 8:6.3f} ms per 1M batch   {8e6 / ms / 1e3:7.1f} M projections/s")
