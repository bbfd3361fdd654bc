"""GPU parity: the CUDA path (through the C ABI) against the CPU oracles on identical seeds.

  vs ORACLE-B (host build of the engine arithmetic)  : BIT-EXACT x, flags, iteration counts, residuals
  vs ORACLE-A (reference-faithful FD/SVD/libm)       : converged/ok flags >= 99.9 %, residual under the
                                                       reference tolerance, joint-vector distance distribution
  golden path rows (reference dumps)                 : fixed points of the GPU project
  full-size properties (1M / 10M)                    : tolerance + limits on every ok sample, idempotence,
                                                       determinism across launches
"""
import numpy as np
import pytest

from conftest import CONFIGS, load_path, make_oracles, near_manifold_seeds

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint64)


@pytest.fixture(scope="module")
def constraints():
    import closed_chain_motion_planner_b200 as pkg

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return {name: pkg.KinematicChainConstraint.from_config(name, device=0) for name in CONFIGS}


@pytest.mark.parametrize("name", CONFIGS)
def test_reference_chain_bit_exact(constraints, name):
    cfg, A, B = make_oracles(name)
    t, q = constraints[name].getInitChain()
    tb, qb = B.get_reference()
    assert np.array_equal(_bits(t), _bits(tb)) and np.array_equal(_bits(q), _bits(qb))
    _, ta = A.init_chain()
    assert np.allclose(t, ta, atol=1e-14)


@pytest.mark.parametrize("name", CONFIGS)
@pytest.mark.parametrize("layout", ["aos", "soa"])
def test_project_bit_exact_vs_engine_arithmetic(constraints, name, layout):
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles(name)
    c = constraints[name]
    seeds = np.concatenate([A.seeds_uniform(0, 0, 4096), near_manifold_seeds(cfg, 1024, seed=1)])
    rb = B.project(seeds, nthreads=8)
    if layout == "aos":
        r = c.projectBatch(torch.from_numpy(seeds).cuda())
        x, rs = r.x.cpu().numpy(), r.resid.cpu().numpy()
    else:
        r = c.projectBatch(torch.from_numpy(np.ascontiguousarray(seeds.T)).cuda(), layout=pkg.CCP_LAYOUT_SOA)
        x, rs = r.x.cpu().numpy().T, r.resid.cpu().numpy().T
    assert np.array_equal(_bits(x), _bits(rb["x"]))
    assert np.array_equal(r.ok.cpu().numpy(), rb["ok"])
    assert np.array_equal(r.converged.cpu().numpy(), rb["converged"])
    assert np.array_equal(r.iters.cpu().numpy(), rb["iters"])
    assert np.array_equal(_bits(rs), _bits(rb["resid"]))


@pytest.mark.parametrize("count", [1, 37, 512, 513, 3001])  # <= 512: in place in page-locked host memory; odd counts
@pytest.mark.parametrize("name", CONFIGS)
def test_project_host_path_matches_device_path(constraints, name, count):
    cfg, A, B = make_oracles(name)
    c = constraints[name]
    seeds = A.seeds_uniform(3, 100, count)
    rh = c.projectBatch(seeds)
    rd = c.projectBatch(torch.from_numpy(seeds).cuda())
    assert np.array_equal(_bits(rh.x), _bits(rd.x.cpu().numpy()))
    assert np.array_equal(rh.ok, rd.ok.cpu().numpy()) and np.array_equal(rh.iters, rd.iters.cpu().numpy())
    assert np.array_equal(_bits(rh.resid), _bits(rd.resid.cpu().numpy()))
    assert np.array_equal(rh.converged, rd.converged.cpu().numpy())


@pytest.mark.parametrize("name", CONFIGS)
def test_project_flags_vs_reference_faithful_oracle(constraints, name):
    """north_star: converged/failed flags agree on >= 99.9 % of samples; every ok sample's residual is
    under the reference tolerance when re-evaluated by the reference-faithful function()."""
    cfg, A, B = make_oracles(name)
    c = constraints[name]
    seeds = np.concatenate([A.seeds_uniform(0, 0, 1500), near_manifold_seeds(cfg, 500, seed=2)])
    ra = A.project(seeds, nthreads=A.max_threads)
    r = c.projectBatch(seeds)
    agree_ok = np.mean(r.ok == ra["ok"])
    agree_cv = np.mean(r.converged == ra["converged"])
    assert agree_ok >= 0.999 and agree_cv >= 0.999, (agree_ok, agree_cv)
    okm = r.ok == 1
    f = A.function(r.x[okm], nthreads=4)
    assert np.all(f[:, 0] <= 1e-3 * (1 + 1e-9)) and np.all(f[:, 1] < 5e-3)
    assert np.all(A.joint_valid(r.x[okm]) == 1)
    # joint-vector distance to the FD-Jacobian reference: reported honestly (chaotic on uniform seeds,
    # SURVEY §0); the near-manifold half must meet the 1e-6 gate on >= 80 % (dumbbell: see DESIGN.md)
    both = okm & (ra["ok"] == 1)
    d = np.max(np.abs(r.x - ra["x"]), axis=1)
    near = np.zeros(len(seeds), bool)
    near[1500:] = True
    frac_near = np.mean(d[both & near] <= 1e-6)
    frac_uni = np.mean(d[both & ~near] <= 1e-6)
    print(f"\n[{name}] ok-flag agreement {agree_ok:.4f}; |x_gpu - x_refA| <= 1e-6: near-manifold {frac_near:.3f}, "
          f"uniform {frac_uni:.3f}; median {np.median(d[both]):.2e}, max {d[both].max():.2e}")
    # The gate is BASELINE.md's 0.8 unless the reference-faithful oracle itself cannot reach it: its agreement with
    # ITSELF on the same seeds moved by one ulp is the ceiling (dumbbell's near-planar start: ~0.67), and the engine
    # must sit at that ceiling.  tests/test_parity_full_gpu.py repeats this at 10 000 seeds per regime.
    r2 = A.project(np.nextafter(seeds[1500:], np.inf), nthreads=A.max_threads)
    both2 = (ra["ok"][1500:] == 1) & (r2["ok"] == 1)
    ceiling = np.mean(np.max(np.abs(ra["x"][1500:] - r2["x"]), axis=1)[both2] <= 1e-6)
    print(f"[{name}] near-manifold ceiling (oracle A vs itself, 1 ulp apart): {ceiling:.3f}")
    assert frac_near >= min(0.8, ceiling - 0.05), (frac_near, ceiling)
    assert d[both].max() < 5e-3


@pytest.mark.parametrize("name", CONFIGS)
def test_function_and_jacobian_batch(constraints, name):
    cfg, A, B = make_oracles(name)
    c = constraints[name]
    x = A.seeds_uniform(5, 0, 1000)
    f = c.functionBatch(x)
    assert np.array_equal(_bits(f), _bits(B.function(x)))
    assert np.max(np.abs(f - A.function(x, nthreads=4))) < 5e-14
    J = c.jacobianBatch(x[:200])
    assert np.array_equal(_bits(J), _bits(B.jacobian(x[:200])))
    assert np.max(np.abs(J - A.jacobian(x[:200], fd=True, nthreads=8))) < 1e-6
    # single-state reference-style calls
    assert np.array_equal(c.function(x[0]), f[0]) and np.array_equal(c.jacobian(x[0]), J[0])
    # SOA device layout
    import closed_chain_motion_planner_b200 as pkg

    xs = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    fs = c.functionBatch(xs, layout=pkg.CCP_LAYOUT_SOA).cpu().numpy().T
    assert np.array_equal(_bits(fs), _bits(f))
    Js = c.jacobianBatch(xs[:, :200].contiguous(), layout=pkg.CCP_LAYOUT_SOA).cpu().numpy()
    assert np.array_equal(_bits(np.ascontiguousarray(np.moveaxis(Js, 2, 0))), _bits(J))


def test_single_state_project_in_place_semantics(constraints):
    """project() mutates x in place, also on failure, and returns converged AND jointValid."""
    cfg, A, B = make_oracles("stefan")
    c = constraints["stefan"]
    x = cfg.start.copy()
    x[0] += 0.1
    x[9] -= 0.2
    x0 = x.copy()
    assert c.project(x) is True
    assert not np.array_equal(x, x0)
    assert c.isSatisfied(x) and c.jointValid(x)
    rb = B.project(x0)
    assert np.array_equal(_bits(x), _bits(rb["x"][0]))
    # a seed that converges outside the joint limits: returns False, x still moved to the last iterate
    seeds = A.seeds_uniform(0, 0, 200)
    rb = B.project(seeds)
    bad = np.where((rb["converged"] == 1) & (rb["ok"] == 0))[0][0]
    y = seeds[bad].copy()
    assert c.project(y) is False
    assert np.array_equal(_bits(y), _bits(rb["x"][bad])) and c.isSatisfied(y) and not c.jointValid(y)
    with pytest.raises(ValueError):
        c.project(np.zeros(13))
    with pytest.raises(ValueError):
        c.setTolerance(0.0, 1e-3)  # the reference throws ompl::Exception


@pytest.mark.parametrize("name", ["dumbbell", "Wine_Bottle"])
def test_golden_path_rows_are_fixed_points_on_gpu(constraints, name):
    c = constraints[name]
    P = load_path(name)
    keep = np.ones(len(P), bool)
    keep[-1] = False
    for i in range(len(P) - 1):
        if np.array_equal(P[i], P[i + 1]):
            keep[i] = keep[i + 1] = False
    rows = P[keep]
    r = c.projectBatch(rows)
    assert np.all(r.ok == 1) and np.all(r.iters <= 1)
    assert np.max(np.abs(r.x - rows)) < 5e-4
    f = c.functionBatch(rows)
    assert np.all(f[:, 0] < 1e-3 * 1.002) and np.all(f[:, 1] < 5e-3 * 1.002)


def test_edge_cases(constraints):
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles("Wine_Bottle")
    c = constraints["Wine_Bottle"]
    # empty batch
    r = c.projectBatch(np.zeros((0, 14)))
    assert r.x.shape == (0, 14) and r.ok.shape == (0,)
    r = c.projectBatch(torch.zeros((0, 14), dtype=torch.float64, device="cuda"))
    assert r.x.shape == (0, 14)
    # one state; NaN / inf seeds fail cleanly and do not disturb their neighbours
    seeds = A.seeds_uniform(9, 0, 67)
    seeds[5, 3] = np.nan
    seeds[40, 0] = np.inf
    seeds[41, :] = 1e300
    r = c.projectBatch(seeds)
    rb = B.project(seeds)
    assert r.ok[5] == 0 and r.ok[40] == 0 and r.ok[41] == 0
    assert np.array_equal(r.ok, rb["ok"]) and np.array_equal(r.iters, rb["iters"])
    good = np.ones(67, bool)
    good[[5, 40, 41]] = False
    assert np.array_equal(_bits(r.x[good]), _bits(rb["x"][good]))
    # iteration cap: max_iter = 0 returns the seed untouched; cap = 3 stops after 3 steps
    c.setOptions(max_iter=0)
    r0 = c.projectBatch(seeds[:8])
    assert np.all(r0.iters == 0) and np.array_equal(_bits(r0.x[:5]), _bits(seeds[:5]))
    c.setOptions(max_iter=3)
    B.set_options(max_iter=3)
    r3 = c.projectBatch(seeds[:32])
    assert r3.iters.max() == 3 and np.array_equal(_bits(r3.x[:5]), _bits(B.project(seeds[:32])["x"][:5]))
    c.setOptions()
    B.set_options()
    # tolerance change is honoured
    c.setTolerance(1e-2, 5e-2)
    B.set_tolerance(1e-2, 5e-2)
    rl = c.projectBatch(seeds[:32][good[:32]])
    assert np.array_equal(rl.iters, B.project(seeds[:32][good[:32]])["iters"])
    c.setTolerance(1e-3, 5e-3)
    # call-order error: project before setInitialPosition
    c2 = pkg.KinematicChainConstraint(14)
    c2.setArmModels(pkg.ArmModel("a", 0), pkg.ArmModel("b", 1))
    with pytest.raises(pkg.CcpError, match="setInitialPosition"):
        c2.projectBatch(seeds[:4])
    # wrong shapes / dtypes on the device path
    with pytest.raises(ValueError):
        c.projectBatch(torch.zeros((4, 13), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        c.projectBatch(torch.zeros((4, 14), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("name,n", [("dumbbell", 1_000_000), ("Wine_Bottle", 10_000_000)])
def test_full_size_properties(constraints, name, n):
    """BASELINE configs[1] (dumbbell, 1M uniform seeds) and configs[2] (Wine_Bottle, 10M) at full size, checked through
    size-independent properties."""
    import ctypes as C

    from closed_chain_motion_planner_b200 import _capi

    c = constraints[name]
    lib = c._lib
    seeds = torch.empty((n, 14), dtype=torch.float64, device="cuda")
    args = _capi.SamplerArgs(rng_seed=0, first_index=0, mode=0, wrap_bounds=0, distance=0.0, near_host=None)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ccp_generate_seeds(c._h, C.byref(args), n, 0, seeds.data_ptr(), st) == 0
    r1 = c.projectBatch(seeds)
    r2 = c.projectBatch(seeds)
    torch.cuda.synchronize()
    # determinism: lane-refill order must not leak into results
    assert torch.equal(r1.x, r2.x) and torch.equal(r1.ok, r2.ok) and torch.equal(r1.iters, r2.iters)
    ok = r1.ok.bool()
    cv = r1.converged.bool()
    frac_ok, frac_cv = ok.float().mean().item(), cv.float().mean().item()
    assert 0.15 < frac_ok < 0.30 and frac_cv > 0.98, (frac_ok, frac_cv)  # SURVEY §6 probe: ~20.7 % / ~100 %
    del r2
    # every converged sample is under tolerance; every ok sample is inside the limits by the margin
    f = c.functionBatch(r1.x)
    assert bool((f[cv, 0] <= 1e-3).all()) and bool((f[cv, 1] < 5e-3).all())
    lb = torch.tensor([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973] * 2, device="cuda", dtype=torch.float64)
    ub = torch.tensor([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973] * 2, device="cuda", dtype=torch.float64)
    xin = r1.x[ok]
    assert bool((xin >= lb + 1e-3).all()) and bool((xin <= ub - 1e-3).all())
    assert torch.equal(c.jointValidBatch(r1.x).bool(), ((r1.x >= lb + 1e-3) & (r1.x <= ub - 1e-3)).all(dim=1))
    assert torch.equal(c.isSatisfiedBatch(r1.x).bool() & cv, cv)
    # not-converged samples ran into the cap
    assert bool((r1.iters[~cv] == 250).all()) and int(r1.iters.max()) <= 250
    # idempotence: projecting converged outputs again takes 0 iterations and changes nothing
    xc = r1.x[cv].contiguous()
    r3 = c.projectBatch(xc)
    assert int(r3.iters.max()) == 0 and torch.equal(r3.x, xc)
    # checksum of checksums against the host build of the engine arithmetic on a strided sample
    cfg, A, B = make_oracles(name)
    sel = torch.arange(0, n, 997 * (n // 1_000_000), device="cuda")
    rb = B.project(seeds[sel].cpu().numpy(), nthreads=8)
    assert np.array_equal(_bits(r1.x[sel].cpu().numpy()), _bits(rb["x"]))
    assert np.array_equal(r1.iters[sel].cpu().numpy(), rb["iters"])


def test_three_arm_projection(constraints):
    """21-DoF extension (BASELINE configs[3]): bit-exact vs the host build, flags vs the 3-arm oracle-A."""
    import closed_chain_motion_planner_b200 as pkg
    from oracle.oracle import OracleA, OracleB

    gp = pkg.grasping_point()
    arms = [pkg.ArmModel("panda_left", 0, gp.t_wb[0]), pkg.ArmModel("panda_right", 1, gp.t_wb[1]),
            pkg.ArmModel("panda_top", 2, gp.t_wb[2])]
    c = pkg.KinematicChainConstraint(21)
    c.setArmModels(*arms)
    q = np.array([-0.16661368, -0.7661184, -0.03369873, -2.37254935, -0.09888003, 1.6927669, 0.17440837] * 3)
    q[7:14] += 0.05
    q[14:] -= 0.07
    c.setInitialPosition(q)
    A = OracleA([0, 1, 2])
    A.set_initial_position(q)
    B = OracleB(pkg.make_model_desc(arms))
    B.set_initial_position(q)
    rng = np.random.default_rng(0)
    xs = np.concatenate([q[None, :] + 0.08 * rng.standard_normal((600, 21)), A.seeds_uniform(0, 0, 400)])
    r = c.projectBatch(xs)
    rb = B.project(xs, nthreads=8)
    assert np.array_equal(_bits(r.x), _bits(rb["x"])) and np.array_equal(r.ok, rb["ok"])
    assert np.array_equal(r.iters, rb["iters"]) and np.array_equal(_bits(r.resid), _bits(rb["resid"]))
    ra = A.project(xs[:300], nthreads=A.max_threads)
    assert np.mean(ra["converged"] == r.converged[:300]) >= 0.99
    f = c.functionBatch(xs[:50])
    assert f.shape == (50, 4) and np.max(np.abs(f - A.function(xs[:50]))) < 5e-14
    assert c.jacobianBatch(xs[:5]).shape == (5, 4, 21)


def test_three_arm_host_paths():
    """21-DoF states through every host path (in place in page-locked memory, one launch, chunked, streaming):
    bit-identical to the device launch."""
    import closed_chain_motion_planner_b200 as pkg
    from oracle.oracle import OracleA

    gp = pkg.grasping_point()
    arms = [pkg.ArmModel("panda_left", 0, gp.t_wb[0]), pkg.ArmModel("panda_right", 1, gp.t_wb[1]),
            pkg.ArmModel("panda_top", 2, gp.t_wb[2])]
    c = pkg.KinematicChainConstraint(21)
    c.setArmModels(*arms)
    q = np.array([-0.16661368, -0.7661184, -0.03369873, -2.37254935, -0.09888003, 1.6927669, 0.17440837] * 3)
    q[7:14] += 0.05
    q[14:] -= 0.07
    c.setInitialPosition(q)
    A = OracleA([0, 1, 2])
    rng = np.random.default_rng(3)
    big = np.concatenate([q[None, :] + 0.08 * rng.standard_normal((200_000, 21)), A.seeds_uniform(0, 0, 50_000)])
    ref = c.projectBatch(torch.from_numpy(big).cuda())
    torch.cuda.synchronize()

    def same(r, lo, hi):
        assert np.array_equal(_bits(ref.x[lo:hi].cpu().numpy()), _bits(r.x))
        assert np.array_equal(ref.ok[lo:hi].cpu().numpy(), r.ok) and np.array_equal(ref.iters[lo:hi].cpu().numpy(), r.iters)
        assert np.array_equal(_bits(ref.resid[lo:hi].cpu().numpy()), _bits(r.resid))

    for count in (1, 300, 5_000):
        same(c.projectBatch(big[:count]), 0, count)
    same(c.projectBatch(big, pinned=True), 0, len(big))
    half = len(big) // 2
    t0, r0 = c.submitHostBatch(big[:half], want_resid=True, pinned=True)
    t1, r1 = c.submitHostBatch(big[half:], want_resid=True, pinned=True)
    c.waitHostBatch(t0)
    c.waitHostBatch(t1)
    same(r0, 0, half)
    same(r1, half, len(big))


def test_panda_model_api(constraints):
    """RobotModel virtuals (panda_rbdl.h:13-23) through the GPU batch kernels."""
    import closed_chain_motion_planner_b200 as pkg

    cfg, A, B = make_oracles("stefan")
    pm = pkg.PandaModel()
    assert pm.getDof() == 7 and pm.getJointLimit().shape == (7, 2)
    T0 = pm.getTransform(np.zeros(7))
    assert T0.shape == (4, 4) and np.allclose(T0[:3, 3], [0.088, 0, 0.926], atol=1e-12)
    q = np.random.default_rng(0).uniform(-2.5, 2.5, (500, 7))
    T = pm.getTransform(q)
    Tb, Jb = B.arm_fk(0, q)
    assert np.array_equal(_bits(T[:, :3, :]), _bits(Tb))
    assert np.max(np.abs(T[:, :3, :] - A.arm_transform(0, q))) < 5e-15
    J = pm.getJacobianMatrix(q)
    assert np.array_equal(_bits(J), _bits(Jb)) and np.max(np.abs(J - A.arm_jacobian(0, q))) < 5e-15
    assert np.allclose(pm.getRotation(q[0]), T[0, :3, :3]) and np.allclose(pm.getTranslation(q[0]), T[0, :3, 3])
    # calibrated model: offsets change the kinematics consistently with the reference-style model build
    dh = 1e-2 * np.random.default_rng(1).standard_normal((7, 4))
    from oracle.oracle import OracleA

    Ac = OracleA([0, 1], dh_offsets=np.stack([dh, dh]))
    Ac.set_arm_base(0, np.eye(4)[:3].reshape(12))
    pmc = pkg.PandaModel(dh)
    assert np.max(np.abs(pmc.getTransform(q[:50])[:, :3, :] - Ac.arm_transform(0, q[:50]))) < 5e-14
