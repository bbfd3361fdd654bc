"""Pins ORACLE-A (reference-faithful CPU restatement) against every artefact the reference ships for the
projection path: the dumped solution paths, the EE poses quoted in the config
comments, Franka's flange pose, and the survey's known-answer values."""
import numpy as np
import pytest

from conftest import CONFIGS, load_cfg, load_path, make_oracles

TOL1, TOL2 = 1e-3, 5e-3


def _vertex_mask(P):
    """Rows that appear twice in a row are roadmap vertices (TRAC-IK output, not project output); the
    last row is the goal vertex (the path ends there, so it is printed once)."""
    dup = np.zeros(len(P), bool)
    dup[-1] = True
    for i in range(len(P) - 1):
        if np.array_equal(P[i], P[i + 1]):
            dup[i] = dup[i + 1] = True
    return dup


@pytest.mark.parametrize("name", ["dumbbell", "Wine_Bottle"])
def test_path_first_row_is_start_joint(name):
    cfg = load_cfg(name)
    P = load_path(name)
    # printAsMatrix prints 6 significant digits (ConstrainedPlanningCommon.cpp:217-221)
    assert np.allclose(P[0], cfg.start, rtol=1e-5, atol=1e-6)
    assert np.array_equal(P[0], P[1])


@pytest.mark.parametrize("name", ["dumbbell", "Wine_Bottle"])
def test_path_interior_rows_are_project_outputs(name):
    """Every non-vertex row came out of KinematicChainConstraint::project inside discreteGeodesic
    (jy_ProjectedStateSpace.cpp:65): it must satisfy the tolerances w.r.t. init_chain(start_joint), and
    — because the loop stops at the FIRST iterate under tolerance with a 0.7 contraction per step — sit
    just under tolerance1."""
    cfg, A, _ = make_oracles(name)
    P = load_path(name)
    interior = P[~_vertex_mask(P)]
    assert len(interior) >= 4
    f = A.function(interior)
    slack = 1 + 2e-3  # 6-digit print precision of the dump
    assert np.all(f[:, 0] < TOL1 * slack), f[:, 0]
    assert np.all(f[:, 1] < TOL2 * slack), f[:, 1]
    assert np.all(f[:, 0] > 0.6 * TOL1), "rows should sit just under tolerance1 (first iterate below it)"
    r = A.project(interior)
    assert np.all(r["iters"] <= 1), r["iters"]  # fixed points of project (0 its; 1 if print rounding crossed tol)
    assert np.all(r["ok"] == 1)


@pytest.mark.parametrize("name", ["dumbbell", "Wine_Bottle"])
def test_path_vertex_rows_are_not_project_outputs(name):
    """Roadmap vertices come from IK and only satisfy the constraint loosely: this is what
    tells project outputs apart in the dumps, and it guards the init_chain / arm-order restatement."""
    cfg, A, _ = make_oracles(name)
    P = load_path(name)
    dup = _vertex_mask(P)
    dup[:2] = False  # the start rows are exact
    f = A.function(P[dup])
    assert np.all(f[:, 0] < 0.05) and np.all(f[:, 1] < 0.15)
    assert np.any(f[:, 0] > TOL1)
    f0 = A.function(P[:1])
    assert f0[0, 0] < 2e-5 and f0[0, 1] < 2e-5


def test_config_comment_poses():
    """config/Wine_Bottle.yaml:21-22 and config/dumbbell.yaml:22-23 quote the EE targets used to make the
    start configurations; FK(start) must land there (checks base frames, DH table, flange, Rz(-pi/4))."""
    cfg, A, _ = make_oracles("Wine_Bottle")
    T = A.arm_transform(0, cfg.start[:7])[0]
    p_world = T[:, 3] + np.array([0.0, 0.3, 1.006])
    assert np.allclose(p_world, [0.45, 0.11, 1.40], atol=1e-2)
    cfg, A, _ = make_oracles("dumbbell")
    T = A.arm_transform(0, cfg.start[:7])[0]
    p_left = T[:, 3] + np.array([0.0, 0.3, 1.006])
    assert np.allclose(p_left[[0, 2]], [0.35, 1.38], atol=1e-2)
    T2 = A.arm_transform(1, cfg.start[7:])[0]
    p_top = np.diag([-1.0, -1.0, 1.0]) @ T2[:, 3] + np.array([1.35, 0.3, 1.006])
    assert np.allclose(p_top[[0, 2]], [0.95, 1.38], atol=1e-2)
    assert abs(p_left[1] - 0.3) < 1e-2 and abs(p_top[1] - 0.3) < 1e-2  # y = object start y (stale comment)


def test_franka_flange_pose():
    """FK(0) of the Panda flange + 0.107 m, turned -45 deg: p = (0.088, 0, 0.926)."""
    _, A, B = make_oracles("stefan")
    for T in (A.arm_transform(0, np.zeros(7))[0], B.arm_fk(0, np.zeros(7))[0][0]):
        assert np.allclose(T[:, 3], [0.088, 0.0, 0.926], atol=1e-12)
        s = np.sqrt(0.5)
        assert np.allclose(T[:, :3], [[s, s, 0], [s, -s, 0], [0, 0, -1]], atol=1e-12)
    q = np.array([0, -0.785, 0, -1.571, 0, 1.571, 0.785])
    assert np.allclose(A.arm_transform(0, q)[0][:, 3], [0.186274417967, 0.0, 0.931094788159], atol=1e-11)


KAT = {
    # init chain and one function/project triple per config
    "stefan": dict(t0=(0.292056451288, 0.381692128649, 0.007945445027),
                   q0=(-0.113562544313, -0.004794166599, -0.008319704794, 0.993484447290),
                   f=(0.161742843168, 0.298603488234), iters=23, fend=(8.389e-4, 7.943e-4)),
    "dumbbell": dict(t0=(0.593535975762, -0.013928001854, 0.004071953333),
                     q0=(0.027274510846, -0.008470577377, -0.007378754425, 0.999564857506),
                     f=(0.153872342356, 0.295473010483), iters=19, fend=(8.250e-4, 1.0476e-3)),
    "Wine_Bottle": dict(t0=(-0.088290011624, -0.005935844107, 0.215053361662),
                        q0=(0.015043740652, -0.017380477445, 0.999652064350, -0.012936580375),
                        f=(0.154064535956, 0.299582153540), iters=18, fend=(9.659e-4, 2.2492e-3)),
}


@pytest.mark.parametrize("name", CONFIGS)
def test_known_answers(name):
    cfg, A, B = make_oracles(name)
    k = KAT[name]
    _, t0 = A.init_chain()
    assert np.allclose(t0, k["t0"], atol=2e-12)
    tb, qb = B.get_reference()
    assert np.allclose(tb, k["t0"], atol=2e-12)
    assert np.allclose(qb, k["q0"], atol=2e-12) or np.allclose(-qb, k["q0"], atol=2e-12)
    x = cfg.start.copy()
    x[0] += 0.1
    x[9] -= 0.2
    for orc in (A, B):
        assert np.allclose(orc.function(x)[0], k["f"], atol=2e-12)
        r = orc.project(x)
        assert r["ok"][0] == 1 and r["iters"][0] == k["iters"]
        assert np.allclose(r["resid"][0], k["fend"], rtol=2e-3)


@pytest.mark.parametrize("name", CONFIGS)
def test_start_is_on_manifold_and_fixed(name):
    cfg, A, B = make_oracles(name)
    for orc in (A, B):
        f = orc.function(cfg.start)[0]
        assert np.all(f < 1e-12)
        r = orc.project(cfg.start)
        assert r["iters"][0] == 0 and np.array_equal(r["x"][0], cfg.start)
    # Wine_Bottle's start has q7 = 2.8898, 7.5e-3 from the limit: still jointValid (margin 1e-3)
    assert A.joint_valid(cfg.start)[0] == 1
