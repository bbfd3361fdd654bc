"""Batched projected-state sampler (jy_ProjectedStateSampler, jy_ProjectedStateSpace.cpp:10-29) on the GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_oracles

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


def _args(**kw):
    from closed_chain_motion_planner_b200 import _capi

    d = dict(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    d.update(kw)
    return _capi.SamplerArgs(**d)


@pytest.fixture(scope="module")
def con():
    import closed_chain_motion_planner_b200 as pkg

    return pkg.KinematicChainConstraint.from_config("stefan", device=0)


def test_seed_stream_matches_oracle(con):
    cfg, A, B = make_oracles("stefan")
    n = 5000
    out = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    a = _args(rng_seed=11, first_index=777)
    assert con._lib.ccp_generate_seeds(con._h, C.byref(a), n, 0, out.data_ptr(), st) == 0
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(A.seeds_uniform(11, 777, n)))
    outs = torch.empty((14, n), dtype=torch.float64, device="cuda")
    assert con._lib.ccp_generate_seeds(con._h, C.byref(a), n, 1, outs.data_ptr(), st) == 0
    assert torch.equal(outs.T.contiguous(), out)


def test_sample_project_equals_generate_then_project(con):
    cfg, A, B = make_oracles("stefan")
    n = 20000
    st = torch.cuda.current_stream().cuda_stream
    x = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    it = torch.empty(n, dtype=torch.int32, device="cuda")
    comp = torch.zeros((n, 14), dtype=torch.float64, device="cuda")
    nok = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = _args(rng_seed=2, first_index=10)
    rc = con._lib.ccp_sample_project_batch(con._h, C.byref(a), n, 0, x.data_ptr(), ok.data_ptr(), it.data_ptr(),
                                           comp.data_ptr(), nok.data_ptr(), st)
    assert rc == 0
    seeds = torch.from_numpy(A.seeds_uniform(2, 10, n)).cuda()
    r = con.projectBatch(seeds)
    assert torch.equal(r.x, x) and torch.equal(r.ok, ok) and torch.equal(r.iters, it)
    k = int(nok.item())
    assert k == int(ok.sum().item()) and 0.15 * n < k < 0.35 * n
    got = comp[:k].cpu().numpy()
    want = x[ok.bool()].cpu().numpy()
    key = lambda a_: a_[np.lexsort(a_.T[::-1])]
    assert np.array_equal(key(got), key(want))  # compaction order is unspecified, content is exact
    assert bool((comp[k:] == 0).all())


def test_wrap_bounds_epilogue(con):
    """enforceBounds after project (jy_ProjectedStateSpace.cpp:14): fmod wrap into [-pi, pi)."""
    cfg, A, B = make_oracles("stefan")
    n = 4000
    st = torch.cuda.current_stream().cuda_stream
    x = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    a = _args(rng_seed=5, wrap_bounds=1)
    assert con._lib.ccp_sample_project_batch(con._h, C.byref(a), n, 0, x.data_ptr(), None, None, None, None, st) == 0
    r = con.projectBatch(torch.from_numpy(A.seeds_uniform(5, 0, n)).cuda())
    want = A.enforce_bounds(r.x.cpu().numpy())
    assert np.array_equal(_bits(x.cpu().numpy()), _bits(want))
    assert bool((x >= -np.pi).all()) and bool((x < np.pi).all())
    y = r.x.clone()
    assert con._lib.ccp_enforce_bounds_batch(con._h, y.data_ptr(), n, 0, st) == 0
    assert torch.equal(y, x)


def test_sample_near_and_gaussian(con):
    cfg, A, B = make_oracles("stefan")
    n = 8000
    st = torch.cuda.current_stream().cuda_stream
    near = np.ascontiguousarray(cfg.start)
    nearp = near.ctypes.data_as(C.POINTER(C.c_double))
    lb = np.tile([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973], 2)
    ub = np.tile([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973], 2)
    s = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    a = _args(rng_seed=1, mode=1, distance=0.25, near_host=nearp)
    assert con._lib.ccp_generate_seeds(con._h, C.byref(a), n, 0, s.data_ptr(), st) == 0
    sn = s.cpu().numpy()
    assert np.all(sn >= np.maximum(lb, near - 0.25) - 1e-15) and np.all(sn <= np.minimum(ub, near + 0.25) + 1e-15)
    assert np.all(np.abs(sn.mean(0) - 0.5 * (np.maximum(lb, near - 0.25) + np.minimum(ub, near + 0.25))) < 0.02)
    a = _args(rng_seed=1, mode=2, distance=0.1, near_host=nearp)
    assert con._lib.ccp_generate_seeds(con._h, C.byref(a), n, 0, s.data_ptr(), st) == 0
    sg = s.cpu().numpy()
    assert np.all(sg >= lb) and np.all(sg <= ub)
    assert np.all(np.abs(sg.mean(0) - near) < 0.01) and np.all(np.abs(sg.std(0) - 0.1) < 0.01)
    # projecting near-manifold samples: (almost) all succeed, and the sampler kernel equals generate+project
    x = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    a = _args(rng_seed=1, mode=1, distance=0.25, near_host=nearp)
    assert con._lib.ccp_sample_project_batch(con._h, C.byref(a), n, 0, x.data_ptr(), ok.data_ptr(), None, None, None, st) == 0
    r = con.projectBatch(torch.from_numpy(sn).cuda())
    assert torch.equal(r.x, x) and torch.equal(r.ok, ok) and ok.float().mean().item() > 0.95
    # mode without a near state is rejected
    a = _args(mode=1, distance=0.25)
    assert con._lib.ccp_generate_seeds(con._h, C.byref(a), n, 0, s.data_ptr(), st) == -1


def test_python_sampler_return_failed_switch(con):
    """return_failed=True reproduces the reference's sampler (jy_ProjectedStateSpace.cpp:13 ignores project()'s return
    value): every seed's wrapped last iterate in stream order; the default pool holds only states with project()==true."""
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles("stefan")
    space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), con)
    s_all = pkg.jy_ProjectedStateSampler(space, pool_size=256, rng_seed=9, return_failed=True)
    got = np.stack([s_all.sampleUniform() for _ in range(300)])  # spans two refills
    rb = B.project(A.seeds_uniform(9, 0, 512), nthreads=4)
    want = B.enforce_bounds(rb["x"]).reshape(-1, 14)
    assert np.array_equal(_bits(got), _bits(want[:300])) and 0 < rb["ok"][:300].mean() < 1
    s_ok = pkg.jy_ProjectedStateSampler(space, pool_size=256, rng_seed=9)
    first = s_ok.sampleUniform()
    assert first.tobytes() in {r.tobytes() for r in want[:256][rb["ok"][:256].astype(bool)]}
