"""The cooperative projection kernel (csrc/ccp_coop.cu: two lanes per sample, one arm each) against the
thread-per-sample kernel: BIT-IDENTICAL states, flags, iteration counts, residuals, compaction — so the launcher may
pick either by batch size (ccp_set_coop_threshold).  Also against the host build of the engine arithmetic."""
import ctypes as C

import numpy as np
import pytest

from conftest import CONFIGS, make_oracles, near_manifold_seeds

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FORCE, NEVER = 1 << 30, 0


def _bits(t):
    return t.cpu().numpy().view(np.uint64)


def _both(c, fn):
    """fn() with the thread-per-sample kernel, then with the cooperative one"""
    out = []
    for thr in (NEVER, FORCE):
        assert c._lib.ccp_set_coop_threshold(c._h, thr) == 0
        out.append(fn())
        torch.cuda.synchronize()
    assert c._lib.ccp_set_coop_threshold(c._h, -1) == 0
    return out


@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("name", CONFIGS)
def test_coop_equals_thread_per_sample(name, layout):
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles(name)
    c = pkg.KinematicChainConstraint.from_config(name, device=0)
    lay = pkg.CCP_LAYOUT_AOS if layout == "aos" else pkg.CCP_LAYOUT_SOA
    for count in (1, 2, 15, 16, 17, 37, 1000, 20_001):
        seeds = np.concatenate([A.seeds_uniform(1, 0, count), near_manifold_seeds(cfg, count, seed=3)])[:max(count, 1)]
        xs = torch.from_numpy(seeds if layout == "aos" else np.ascontiguousarray(seeds.T)).cuda()

        def run():
            compact = torch.zeros((len(seeds), 14), dtype=torch.float64, device="cuda")
            n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
            r = c.projectBatch(xs, layout=lay, compact=compact, n_ok=n_ok)
            return r, compact, n_ok

        (r0, c0, n0), (r1, c1, n1) = _both(c, run)
        assert np.array_equal(_bits(r0.x), _bits(r1.x)), (name, layout, count)
        assert torch.equal(r0.ok, r1.ok) and torch.equal(r0.converged, r1.converged) and torch.equal(r0.iters, r1.iters)
        assert np.array_equal(_bits(r0.resid), _bits(r1.resid))
        k = int(n0.item())
        assert k == int(n1.item()) == int(r0.ok.sum())
        srt = lambda m: m[np.lexsort(m.T[::-1])]
        assert np.array_equal(srt(c0[:k].cpu().numpy()).view(np.uint64), srt(c1[:k].cpu().numpy()).view(np.uint64))
        if count == 1000:  # and against the host twin
            rb = B.project(seeds, nthreads=8)
            x1 = r1.x.cpu().numpy() if layout == "aos" else r1.x.cpu().numpy().T
            assert np.array_equal(x1.view(np.uint64), rb["x"].view(np.uint64)) and np.array_equal(r1.iters.cpu().numpy(), rb["iters"])


def test_coop_edge_cases_options_and_calibrated_model():
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles("Wine_Bottle")
    c = pkg.KinematicChainConstraint.from_config("Wine_Bottle", device=0)
    seeds = A.seeds_uniform(9, 0, 300)
    seeds[5, 3] = np.nan
    seeds[40, 0] = np.inf
    seeds[41, :] = 1e300
    xs = torch.from_numpy(seeds).cuda()
    for opts in (dict(), dict(max_iter=0), dict(max_iter=3), dict(damping=1e-4, clamp=True), dict(step=0.5, joint_margin=0.0)):
        c.setOptions(**opts)
        r0, r1 = _both(c, lambda: c.projectBatch(xs))
        assert np.array_equal(_bits(r0.x), _bits(r1.x)), opts
        assert torch.equal(r0.ok, r1.ok) and torch.equal(r0.iters, r1.iters) and np.array_equal(_bits(r0.resid), _bits(r1.resid))
    c.setOptions()
    c.setTolerance(1e-2, 5e-2)
    r0, r1 = _both(c, lambda: c.projectBatch(xs))
    assert np.array_equal(_bits(r0.x), _bits(r1.x)) and torch.equal(r0.iters, r1.iters)
    # host entry points (in place in page-locked memory, one launch) take the same kernel choice
    assert c._lib.ccp_set_coop_threshold(c._h, FORCE) == 0
    rh = c.projectBatch(seeds)
    assert np.array_equal(rh.x.view(np.uint64), r1.x.cpu().numpy().view(np.uint64)) and np.array_equal(rh.ok, r1.ok.cpu().numpy())
    # calibrated (alpha-offset) arms: the generic link code
    gp = pkg.grasping_point()
    dh = 1e-2 * np.random.default_rng(4).standard_normal((2, 7, 4))
    arms = [pkg.ArmModel("l", 0, gp.t_wb[0], dh_offsets=dh[0]), pkg.ArmModel("t", 2, gp.t_wb[2], dh_offsets=dh[1])]
    cc = pkg.KinematicChainConstraint(14)
    cc.setArmModels(*arms)
    cc.setInitialPosition(cfg.start)
    s2 = torch.from_numpy(A.seeds_uniform(2, 0, 700)).cuda()
    r0, r1 = _both(cc, lambda: cc.projectBatch(s2))
    assert np.array_equal(_bits(r0.x), _bits(r1.x)) and torch.equal(r0.ok, r1.ok) and torch.equal(r0.iters, r1.iters)


@pytest.mark.parametrize("layout", ["aos", "soa"])
@pytest.mark.parametrize("calibrated", [False, True])
def test_coop_three_arms_equals_thread_per_sample(layout, calibrated):
    """21 DoF: four lanes per sample (ccp_project_coop3_kernel; lane 3 repeats arm 0 for the second residual pair) against the
    thread-per-sample kernel and the host build of the engine arithmetic — bit for bit, stock and calibrated link code,
    every output, the sampler's wrap + compaction, non-default options."""
    import closed_chain_motion_planner_b200 as pkg
    from oracle.oracle import OracleA, OracleB
    from test_parity_full_gpu import Q3

    gp = pkg.grasping_point()
    dh = 1e-2 * np.random.default_rng(11).standard_normal((3, 7, 4)) if calibrated else [None] * 3
    arms = [pkg.ArmModel(f"arm{i}", i, gp.t_wb[i], dh_offsets=dh[i]) for i in range(3)]
    c = pkg.KinematicChainConstraint(21)
    c.setArmModels(*arms)
    c.setInitialPosition(Q3)
    A = OracleA([0, 1, 2])
    B = OracleB(pkg.make_model_desc(arms))
    B.set_initial_position(Q3)
    lay = pkg.CCP_LAYOUT_AOS if layout == "aos" else pkg.CCP_LAYOUT_SOA
    rng = np.random.default_rng(5)
    for count in (1, 7, 8, 9, 33, 1000, 6001):
        d = rng.standard_normal((count, 21))
        d *= 0.25 / np.linalg.norm(d, axis=1, keepdims=True)
        seeds = np.concatenate([A.seeds_uniform(2, 0, count), Q3[None, :] + d])[:max(count, 1)]
        if count == 33:
            seeds[3, 5] = np.nan
            seeds[4, :] = 1e300
        xs = torch.from_numpy(seeds if layout == "aos" else np.ascontiguousarray(seeds.T)).cuda()

        def run():
            compact = torch.zeros((len(seeds), 21), dtype=torch.float64, device="cuda")
            n_ok = torch.zeros(1, dtype=torch.int64, device="cuda")
            r = c.projectBatch(xs, layout=lay, compact=compact, n_ok=n_ok)
            return r, compact, n_ok

        (r0, c0, n0), (r1, c1, n1) = _both(c, run)
        assert np.array_equal(_bits(r0.x), _bits(r1.x)), (layout, calibrated, count)
        assert torch.equal(r0.ok, r1.ok) and torch.equal(r0.converged, r1.converged) and torch.equal(r0.iters, r1.iters)
        assert np.array_equal(_bits(r0.resid), _bits(r1.resid))
        k = int(n0.item())
        assert k == int(n1.item()) == int(r0.ok.sum())
        srt = lambda m: m[np.lexsort(m.T[::-1])]
        assert np.array_equal(srt(c0[:k].cpu().numpy()).view(np.uint64), srt(c1[:k].cpu().numpy()).view(np.uint64))
        if count == 1000:
            rb = B.project(seeds, nthreads=8)
            x1 = r1.x.cpu().numpy() if layout == "aos" else r1.x.cpu().numpy().T
            assert np.array_equal(x1.view(np.uint64), rb["x"].view(np.uint64)) and np.array_equal(r1.iters.cpu().numpy(), rb["iters"])
            assert int(r1.ok.sum()) > 0
    if layout == "aos":
        xs = torch.from_numpy(A.seeds_uniform(3, 0, 400)).cuda()
        for opts in (dict(max_iter=0), dict(max_iter=3), dict(damping=1e-4, clamp=True), dict(step=0.5, joint_margin=0.0)):
            c.setOptions(**opts)
            r0, r1 = _both(c, lambda: c.projectBatch(xs))
            assert np.array_equal(_bits(r0.x), _bits(r1.x)), opts
            assert torch.equal(r0.ok, r1.ok) and torch.equal(r0.iters, r1.iters) and np.array_equal(_bits(r0.resid), _bits(r1.resid))
        c.setOptions()
        # the sampler path: seed kernel -> projection with the enforceBounds wrap
        from closed_chain_motion_planner_b200 import _capi

        def sample():
            a = _capi.SamplerArgs(rng_seed=3, first_index=5, mode=0, wrap_bounds=1, distance=0.0, near_host=None)
            x = torch.empty((3000, 21), dtype=torch.float64, device="cuda")
            ok = torch.empty(3000, dtype=torch.uint8, device="cuda")
            it = torch.empty(3000, dtype=torch.int32, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            assert c._lib.ccp_sample_project_batch(c._h, C.byref(a), 3000, 0, x.data_ptr(), ok.data_ptr(), it.data_ptr(), None, None, st) == 0
            return x, ok, it

        (x0, ok0, it0), (x1, ok1, it1) = _both(c, sample)
        assert np.array_equal(_bits(x0), _bits(x1)) and torch.equal(ok0, ok1) and torch.equal(it0, it1)
        # host entry point (one launch, in place) takes the same kernel choice
        assert c._lib.ccp_set_coop_threshold(c._h, FORCE) == 0
        hs = A.seeds_uniform(4, 0, 50)
        rh = c.projectBatch(hs)
        assert c._lib.ccp_set_coop_threshold(c._h, NEVER) == 0
        rt = c.projectBatch(hs)
        assert c._lib.ccp_set_coop_threshold(c._h, -1) == 0
        assert np.array_equal(rh.x.view(np.uint64), rt.x.view(np.uint64)) and np.array_equal(rh.ok, rt.ok)


def test_coop_sampler_wrap_and_golden_rows():
    """the sampler path (seed kernel -> projection with the enforceBounds wrap and compaction) and the reference's dumped
    path rows through the cooperative kernel"""
    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi
    from conftest import load_path

    c = pkg.KinematicChainConstraint.from_config("dumbbell", device=0)
    n = 5000

    def run():
        a = _capi.SamplerArgs(rng_seed=3, first_index=77, mode=0, wrap_bounds=1, distance=0.0, near_host=None)
        x = torch.empty((n, 14), dtype=torch.float64, device="cuda")
        ok = torch.empty(n, dtype=torch.uint8, device="cuda")
        it = torch.empty(n, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        assert c._lib.ccp_sample_project_batch(c._h, C.byref(a), n, 0, x.data_ptr(), ok.data_ptr(), it.data_ptr(), None, None, st) == 0
        return x, ok, it

    (x0, ok0, it0), (x1, ok1, it1) = _both(c, run)
    assert np.array_equal(_bits(x0), _bits(x1)) and torch.equal(ok0, ok1) and torch.equal(it0, it1)
    assert float(x1.abs().max()) <= np.pi
    P = load_path("dumbbell")
    keep = np.ones(len(P), bool)
    keep[-1] = False
    for i in range(len(P) - 1):
        if np.array_equal(P[i], P[i + 1]):
            keep[i] = keep[i + 1] = False
    rows = torch.from_numpy(P[keep]).cuda()
    assert c._lib.ccp_set_coop_threshold(c._h, FORCE) == 0
    r = c.projectBatch(rows)
    assert bool((r.ok == 1).all()) and int(r.iters.max()) <= 1 and float((r.x - rows).abs().max()) < 5e-4


@pytest.mark.parametrize("name", ["stefan", "Wine_Bottle", "stefan_three_arm"])
def test_coop_geodesic_equals_thread_per_edge(name):
    """discreteGeodesic with two lanes per edge (ccp_geodesic_coop_kernel; three arms: four lanes, ccp_geodesic_coop3_kernel)
    against the one-thread walk and the host twin: the same states, state counts, reached flags and iteration totals, bit
    for bit."""
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles(name)
    c = pkg.KinematicChainConstraint.from_config(name, device=0)
    space = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(c.getAmbientDimension()), c)
    # (three arms: 4 % of uniform seeds end inside the limits)
    two = c.getAmbientDimension() == 14
    smp = space.allocStateSampler(pool_size=8192 if two else 1 << 16, rng_seed=5)
    V = smp.sampleUniformBatch(8000 if two else 60_000)
    assert V.shape[0] >= 1400
    for E in (1, 5, 17, 700):
        frm = torch.cat([torch.from_numpy(cfg.start[None, :]).cuda().repeat(E // 2 + 1, 1), V[:E]])[:E].contiguous()
        to = V[E:2 * E].contiguous()
        if E == 5:
            to[0] = frm[0]  # an edge that is "already there"
        r0, r1 = _both(c, lambda: space.discreteGeodesicBatch(frm, to, max_states=24))
        assert torch.equal(r0.reached, r1.reached) and torch.equal(r0.n_states, r1.n_states) and torch.equal(r0.iters, r1.iters)
        for e in range(E):
            k = int(r0.n_states[e])
            assert np.array_equal(_bits(r0.states[e, :k]), _bits(r1.states[e, :k])), (name, E, e)
        rc_b, ns_b, st_b, it_b = B.discrete_geodesic(frm.cpu().numpy(), to.cpu().numpy(), delta=0.25, lam=2.0, max_states=24)
        assert np.array_equal(r1.reached.cpu().numpy(), rc_b) and np.array_equal(r1.n_states.cpu().numpy(), ns_b)
        assert np.array_equal(r1.iters.cpu().numpy(), it_b)
        for e in range(E):
            assert np.array_equal(_bits(r1.states[e, :ns_b[e]]), st_b[e, :ns_b[e]].view(np.uint64))


@pytest.mark.parametrize("name", ["dumbbell", "stefan_three_arm"])
def test_stock_specialisation_changes_no_bit(name):
    """The kernels specialised to the reference's stock Panda table and base frames (PANDA = 2: exact-zero link terms and the
    diagonal base-to-base rotation are skipped, their constants not loaded) against the structured-alpha kernels
    (CCP_MODEL_NO_STOCK) on the same model: states, flags, iteration counts, residuals, function and Jacobian values —
    identical bits.  Both projection kernels and the geodesic walk."""
    import closed_chain_motion_planner_b200 as pkg
    from closed_chain_motion_planner_b200 import _capi

    cfg = pkg.grasping_point().loadConfig(name)
    arms = [pkg.ArmModel(name=nm, index=ix, t_wb=cfg.t_wb[ix]) for nm, ix in zip(cfg.arm_names, cfg.arm_indices)]
    cs = pkg.KinematicChainConstraint(7 * len(arms))
    cs.setArmModels(*arms)
    cn = pkg.KinematicChainConstraint(7 * len(arms))
    cn._arms = list(arms)
    d = pkg.make_model_desc(arms)
    d.flags = _capi.CCP_MODEL_NO_STOCK
    import ctypes as C

    h = C.c_void_p()
    assert cn._lib.ccp_create(C.byref(d), 0, C.byref(h)) == 0
    cn._h, cn._desc = h, d
    for c in (cs, cn):
        c.setInitialPosition(cfg.start)
    n = cs.getAmbientDimension()
    rng = np.random.default_rng(1)
    lb = np.tile([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973], n // 7)
    ub = np.tile([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973], n // 7)
    seeds = np.concatenate([rng.uniform(lb, ub, (20_000, n)), cfg.start[None, :] + 0.1 * rng.standard_normal((10_000, n))])
    xs = torch.from_numpy(seeds).cuda()
    for count in (300, len(seeds)):  # cooperative kernel (two arms) / thread-per-sample kernel
        a, b = cs.projectBatch(xs[:count].contiguous()), cn.projectBatch(xs[:count].contiguous())
        torch.cuda.synchronize()
        assert np.array_equal(_bits(a.x), _bits(b.x)) and torch.equal(a.ok, b.ok) and torch.equal(a.iters, b.iters)
        assert np.array_equal(_bits(a.resid), _bits(b.resid))
    assert np.array_equal(_bits(cs.functionBatch(xs[:2000].contiguous())), _bits(cn.functionBatch(xs[:2000].contiguous())))
    assert np.array_equal(_bits(cs.jacobianBatch(xs[:500].contiguous())), _bits(cn.jacobianBatch(xs[:500].contiguous())))
    if n == 14:
        ss = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), cs)
        sn = pkg.jy_ProjectedStateSpace(pkg.KinematicChainSpace(14), cn)
        V = a.x[a.ok.bool()][:400].contiguous()
        frm = torch.from_numpy(cfg.start[None, :]).cuda().repeat(200, 1).contiguous()
        ga, gb = ss.discreteGeodesicBatch(frm, V[:200].contiguous(), max_states=24), sn.discreteGeodesicBatch(frm, V[:200].contiguous(), max_states=24)
        assert torch.equal(ga.reached, gb.reached) and torch.equal(ga.n_states, gb.n_states) and torch.equal(ga.iters, gb.iters)
        for e in range(200):
            k = int(ga.n_states[e])
            assert np.array_equal(_bits(ga.states[e, :k]), _bits(gb.states[e, :k]))
