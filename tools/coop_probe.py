"""Latency regime: complete launches of small batches on the thread-per-sample kernel and on the cooperative kernel
(two lanes per sample).  usage: python tools/coop_probe.py [config]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA  # seeds only

cfgname = sys.argv[1] if len(sys.argv) > 1 else "dumbbell"
c = pkg.KinematicChainConstraint.from_config(cfgname)
A = OracleA(c.config.arm_indices)
rows = []
for count in [int(v) for v in os.environ.get("COOP_PROBE_COUNTS", "1,16,256,1000,4000,10000,20000,40000,80000,160000,320000").split(",")]:
    seeds = torch.from_numpy(A.seeds_uniform(0, 0, count)).cuda()
    out = torch.empty_like(seeds)
    row = {"count": count}
    for name, thr in (("thread_per_sample", 0), ("cooperative", 1 << 30)):
        c._lib.ccp_set_coop_threshold(c._h, thr)
        best = 1e30
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c.projectBatch(seeds, out=out, want_resid=False)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        row[name + "_ms"] = round(best, 4)
    row["speedup"] = round(row["thread_per_sample_ms"] / row["cooperative_ms"], 3)
    rows.append(row)
    print(json.dumps(row), flush=True)
# one state per call through the host entry point (what a planner calling project() state by state pays)
x = c.config.start.copy()
x[0] += 0.05
for name, thr in (("thread_per_sample", 0), ("cooperative", 1 << 30)):
    c._lib.ccp_set_coop_threshold(c._h, thr)
    for _ in range(50):
        c.project(x.copy())
    t0 = time.perf_counter()
    for _ in range(500):
        c.project(x.copy())
    print(json.dumps({"single_state_project_us": round((time.perf_counter() - t0) / 500 * 1e6, 2), "kernel": name}))
