"""The planner's inner loop with the batched entry points (not part of the product; BASELINE configs[4] evidence).

stefanBiPRM grows its roadmaps one vertex at a time: sample a projected state (jy_ProjectedStateSpace.cpp:10-15),
find the k = 5 nearest vertices, walk a discreteGeodesic to each (stefanBiPRM.cpp:315,397,463).  Here the same work
for a whole roadmap at once: N projected vertices from the pool-backed sampler, k nearest neighbours on the GPU,
N x k geodesics in one launch; the reference-faithful CPU oracle walks a bounded sample of the same edges for scale.
usage: python tools/roadmap_probe.py [config] [N] [k]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import closed_chain_motion_planner_b200 as pkg
from oracle.oracle import OracleA  # CPU reference for the bounded comparison only

cfgname = sys.argv[1] if len(sys.argv) > 1 else "stefan"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = pkg.KinematicChainConstraint.from_config(cfgname)
space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c)
smp = space.allocStateSampler(pool_size=max(6 * N, 4096), rng_seed=3)
smp.sampleUniformBatch(4096)  # warm-up
torch.cuda.synchronize()

t0 = time.perf_counter()
V = smp.sampleUniformBatch(6 * N)[:N].contiguous()
torch.cuda.synchronize()
t_sample = time.perf_counter() - t0
n = V.shape[0]
t0 = time.perf_counter()
d = torch.cdist(V, V)
d.fill_diagonal_(float("inf"))
nbr = d.topk(k, largest=False).indices  # (n, k)
torch.cuda.synchronize()
t_knn = time.perf_counter() - t0
frm = V.repeat_interleave(k, dim=0).contiguous()
to = V[nbr.reshape(-1)].contiguous()
t0 = time.perf_counter()
g = space.discreteGeodesicBatch(frm, to, max_states=40)
torch.cuda.synchronize()
t_geo = time.perf_counter() - t0
reached = g.reached.bool()
# connectivity of the roadmap
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components

src = torch.arange(n, device=V.device).repeat_interleave(k)[reached].cpu().numpy()
dst = nbr.reshape(-1)[reached].cpu().numpy()
ncomp, lab = connected_components(sp.coo_matrix((np.ones(len(src)), (src, dst)), shape=(n, n)), directed=False)
big = np.bincount(lab).max()
print(f"{cfgname}: {n} vertices in {t_sample*1e3:.1f} ms ({6*N} seeds projected), kNN {t_knn*1e3:.1f} ms, "
      f"{n*k} geodesics in {t_geo*1e3:.1f} ms ({n*k/t_geo/1e6:.2f} M edges/s), {float(reached.float().mean()):.3f} reached, "
      f"largest component {big}/{n}; total {1e3*(t_sample+t_knn+t_geo):.1f} ms")
# the same edges on the CPU, reference-faithful arithmetic, bounded sample
A = OracleA(c.config.arm_indices)
A.set_initial_position(c.config.start)
m = 64
t0 = time.perf_counter()
rc, ns, st = A.discrete_geodesic(frm[:m].cpu().numpy(), to[:m].cpu().numpy(), delta=0.25, lam=2.0, max_states=40,
                                     nthreads=A.max_threads)
t_cpu = time.perf_counter() - t0
agree = float(np.mean(rc == g.reached[:m].cpu().numpy()))
print(f"oracle A (FD Jacobian + SVD, {A.max_threads} threads): {m} of those edges in {t_cpu:.2f} s = {m/t_cpu:.1f} edges/s "
      f"-> the {n*k} edges would take {n*k/(m/t_cpu)/60:.1f} min; reached-flag agreement on the sample {agree:.3f}")
