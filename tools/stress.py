"""Randomised consistency stress of the projection entry points (not part of the product): complete launches against
pipelined sequences of random sizes, both layouts, all configs; strided checks against the host build of the engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import make_model_desc
from oracle.oracle import OracleA, OracleB

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
bits = lambda t: t.cpu().numpy().view(np.uint64)
for cfgname in ("dumbbell", "Wine_Bottle", "stefan", "stefan_three_arm"):
    c = pkg.KinematicChainConstraint.from_config(cfgname)
    n = c.getAmbientDimension()
    A = OracleA(c.config.arm_indices)
    cfg = c.config
    arms = [pkg.ArmModel(name=nm, index=ix, t_wb=cfg.t_wb[ix]) for nm, ix in zip(cfg.arm_names, cfg.arm_indices)]
    B = OracleB(make_model_desc(arms))
    B.set_initial_position(cfg.start)
    for rnd in range(rounds):
        total = int(rng.integers(1, 400_000))
        seeds = A.seeds_uniform(int(rng.integers(0, 1000)), int(rng.integers(0, 10**6)), total)
        lay = int(rng.integers(0, 2))
        X = torch.from_numpy(seeds if lay == 0 else np.ascontiguousarray(seeds.T)).cuda()
        ref = c.projectBatch(X, layout=lay)
        # random split into pipelined launches (+ sometimes a complete launch in the middle), then flush
        cuts = np.sort(rng.integers(0, total + 1, size=int(rng.integers(1, 7))))
        cuts = np.concatenate([[0], cuts, [total]])
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = (X[a:b] if lay == 0 else X[:, a:b]).contiguous()
            parts.append(c.projectBatch(part, layout=lay, pipelined=bool(rng.integers(0, 4))))
        c.flush()
        torch.cuda.synchronize()
        cat = (lambda ts: torch.cat(ts, dim=0)) if lay == 0 else (lambda ts: torch.cat(ts, dim=1))
        x = cat([p.x for p in parts])
        it = torch.cat([p.iters for p in parts])
        ok = torch.cat([p.ok for p in parts])
        rs = cat([p.resid for p in parts])
        assert np.array_equal(bits(x), bits(ref.x)), (cfgname, rnd, "x")
        assert torch.equal(it, ref.iters) and torch.equal(ok, ref.ok), (cfgname, rnd, "flags")
        assert np.array_equal(bits(rs), bits(ref.resid)), (cfgname, rnd, "resid")
        sel = np.arange(0, total, max(1, total // 3000))
        rb = B.project(seeds[sel], nthreads=8)
        xg = (ref.x if lay == 0 else ref.x.T)[torch.from_numpy(sel).cuda()].cpu().numpy()
        assert np.array_equal(xg.view(np.uint64), rb["x"].view(np.uint64)), (cfgname, rnd, "host twin")
        assert np.array_equal(ref.iters.cpu().numpy()[sel], rb["iters"])
    print(cfgname, "ok", rounds, "rounds")
print("stress passed")
