"""Batch-size sweep of the projection path on one GPU (BASELINE configs[2], [3]): device-generated Seeds-U
(ccp_sample_project_batch, no seed buffer), outputs ok + iters only, CUDA-event time of the launch.
usage: python tools/sweep.py CONFIG COUNT [COUNT ...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi

cfg = sys.argv[1]
counts = [int(float(v)) for v in sys.argv[2:]]
c = pkg.KinematicChainConstraint.from_config(cfg)
lib, h = c._lib, c._h
peak, _ = c.fp64PeakProbe(5)
fl_iter, fl_tail = c.algorithmicFlops()
nmax = max(counts)
ok = torch.empty(nmax, dtype=torch.uint8, device="cuda")
it = torch.empty(nmax, dtype=torch.int32, device="cuda")
n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rows = []
for cnt in counts:
    best = 1e30
    reps = 3 if cnt <= 10_000_000 else 2
    for r in range(reps + 1):
        a = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
        n_ok.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.ccp_sample_project_batch(h, C.byref(a), cnt, 0, None, ok.data_ptr(), it.data_ptr(), None, n_ok.data_ptr(), st)
        assert rc == 0, lib.ccp_last_error(h)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    iters = int(it[:cnt].sum(dtype=torch.int64))
    flops = iters * fl_iter + cnt * fl_tail
    row = {"config": cfg, "arms": c.k_, "count": cnt, "ms": best, "projections_per_s": cnt / best * 1e3,
           "converged_per_s": int(n_ok) / best * 1e3, "ok_fraction": int(n_ok) / cnt, "mean_iters": iters / cnt,
           "tflops": flops / best / 1e9, "frac_of_measured_fp64_peak": flops / (best * 1e-3) / peak}
    rows.append(row)
    print(json.dumps(row), flush=True)
