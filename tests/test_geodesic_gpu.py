"""Batched discreteGeodesic and the pool-backed projected-state sampler on the GPU."""
import numpy as np
import pytest

from conftest import load_path, make_oracles, near_manifold_seeds
from test_geodesic_golden import CASES, segments

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _space(name, delta=0.25):
    import closed_chain_motion_planner_b200 as pkg

    c = pkg.KinematicChainConstraint.from_config(name, device=0)
    return pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), c, delta=delta, lam=2.0), c


@pytest.mark.parametrize("name", list(CASES))
def test_gpu_geodesics_reproduce_reference_path_rows(name):
    delta, tol = CASES[name]
    cfg, A, B = make_oracles(name)
    sp, c = _space(name, delta)
    P = load_path(name)
    segs = segments(P)
    frm = np.stack([cfg.start if s == 1 else P[s] for s, e in segs])
    to = np.stack([P[e] for s, e in segs])
    r = sp.discreteGeodesicBatch(frm, to, max_states=32)
    for k, (s, e) in enumerate(segs):
        gold = P[s + 1:e]
        got = r.states[k, 1:r.n_states[k]]
        assert len(got) == len(gold) and np.max(np.abs(got - gold)) < tol
    # single-edge reference-style call
    ok, geo = sp.discreteGeodesic(frm[0], to[0])
    assert len(geo) == r.n_states[0] and ok == bool(r.reached[0]) and np.array_equal(geo[0], frm[0])


@pytest.mark.parametrize("name", ["stefan", "Wine_Bottle"])
def test_geodesic_batch_bit_exact_and_flags(name):
    cfg, A, B = make_oracles(name)
    sp, c = _space(name)
    # edges between projected near-manifold states (what the planner connects)
    pts = B.project(near_manifold_seeds(cfg, 600, seed=7, radius=0.6), nthreads=8)
    good = pts["x"][pts["ok"] == 1]
    rng = np.random.default_rng(1)
    e = 256
    i, j = rng.integers(0, len(good), e), rng.integers(0, len(good), e)
    frm, to = good[i], good[j]
    frm[0] = to[0]  # zero-length edge
    r = sp.discreteGeodesicBatch(frm, to, max_states=48)
    rc, ns, st, it = B.discrete_geodesic(frm, to, max_states=48)
    assert np.array_equal(r.reached, rc) and np.array_equal(r.n_states, ns) and np.array_equal(r.iters, it)
    for k in range(e):
        assert np.array_equal(_bits(r.states[k, :ns[k]]), _bits(st[k, :ns[k]]))
    assert r.reached[0] == 1 and r.n_states[0] == 1
    assert 0.2 < r.reached.mean() <= 1.0
    # against the reference-faithful oracle (FD Jacobian): same verdicts and the same number of states on
    # (almost) every edge; states within 1e-4
    ra, nsa, sta = A.discrete_geodesic(frm[:64], to[:64], max_states=48, nthreads=A.max_threads)
    assert np.mean(ra == r.reached[:64]) >= 0.95 and np.mean(nsa == r.n_states[:64]) >= 0.9
    # every returned state is on the manifold and inside the limits; consecutive states are <= lambda*delta apart
    for k in range(0, e, 9):
        s_k = r.states[k, :ns[k]]
        if ns[k] > 1:
            assert np.all(c.isSatisfiedBatch(s_k[1:]) == 1) and np.all(c.jointValidBatch(s_k[1:]) == 1)
            assert np.all(np.linalg.norm(np.diff(s_k, axis=0), axis=1) <= 0.5 + 1e-12)
    # device tensors in -> device tensors out
    rd = sp.discreteGeodesicBatch(torch.from_numpy(frm).cuda(), torch.from_numpy(to).cuda(), max_states=48)
    assert torch.equal(rd.n_states.cpu(), torch.from_numpy(ns)) and torch.equal(rd.reached.cpu(), torch.from_numpy(rc))
    # running out of room is reported as not reached, never as an overflow
    r2 = sp.discreteGeodesicBatch(frm, to, max_states=2)
    assert r2.n_states.max() <= 2 and np.all(r2.reached[ns > 2] == 0)


def test_pool_backed_sampler():
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles("stefan")
    sp, c = _space("stefan")
    smp = sp.allocStateSampler(pool_size=4096, rng_seed=5)
    xs = np.stack([smp.sampleUniform() for _ in range(1500)])
    assert smp.launches == 2  # ~22 % of 4096 per refill
    assert np.all(c.isSatisfiedBatch(xs) == 1)
    assert np.all(xs >= -np.pi) and np.all(xs < np.pi)  # enforceBounds applied
    # the pool is exactly the ok subset of the projected counter stream (same seeds as the oracle's generator)
    seeds = A.seeds_uniform(5, 0, 4096)
    rb = B.project(seeds, nthreads=8)
    want = A.enforce_bounds(rb["x"][rb["ok"] == 1])
    key = lambda m: m[np.lexsort(m.T[::-1])]
    assert np.array_equal(key(xs[: len(want)]), key(want))
    # near / gaussian draws land on the manifold close to the centre
    y = smp.sampleUniformNear(cfg.start, 0.1)
    z = smp.sampleGaussian(cfg.start, 0.05)
    assert c.isSatisfied(y) and c.isSatisfied(z)
    assert np.linalg.norm(y - cfg.start) < 1.0 and np.linalg.norm(z - cfg.start) < 1.0
    out = np.zeros(14)
    assert smp.sampleUniform(out) is out and c.isSatisfied(out)
    b = smp.sampleUniformNearBatch(cfg.start, 0.25, 2000)
    assert b.shape[0] > 1800  # near-manifold seeds almost always succeed
