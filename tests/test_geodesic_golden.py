"""The reference's dumped solution paths are `path.interpolate()` outputs (ConstrainedPlanningCommon.cpp:217-221):
between two consecutive roadmap vertices every printed row is a state of
jy_ProjectedStateSpace::discreteGeodesic(v_i, v_{i+1}, interpolate=true) (jy_ProjectedStateSpace.cpp:32-96), i.e. an
actual OUTPUT of the reference's project().  Re-walking the same edges with the CPU oracles must reproduce those rows
to the 6 significant digits they were printed with.  This pins the oracles (FK, residual, FD Jacobian, step, loop exit,
interpolate, the traversal checks) against real reference results — 21 rows for Wine_Bottle, 4 for dumbbell.

dumbbell: the dump was produced with delta = 0.5 (row spacing 0.50; the shipped source says 0.25 at
ConstrainedPlanningCommon.cpp:118 — the constant was edited between runs); Wine_Bottle matches the shipped 0.25."""
import numpy as np
import pytest

from conftest import load_path, make_oracles

CASES = {"Wine_Bottle": (0.25, 2e-5), "dumbbell": (0.5, 3e-4)}


def segments(P):
    """(first, last) row indices of every vertex-to-vertex segment; vertices are the duplicated rows + the goal."""
    idx = [i for i in range(len(P) - 1) if np.array_equal(P[i], P[i + 1])]
    out = []
    for k, i in enumerate(idx):
        out.append((i + 1, idx[k + 1] if k + 1 < len(idx) else len(P) - 1))
    return out


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("which", ["A", "B"])
def test_oracle_geodesics_reproduce_reference_path_rows(name, which):
    delta, tol = CASES[name]
    cfg, A, B = make_oracles(name)
    P = load_path(name)
    rows = 0
    for s, e in segments(P):
        v0 = cfg.start if s == 1 else P[s]
        gold = P[s + 1:e]
        if which == "A":
            rc, ns, st = A.discrete_geodesic(v0, P[e], delta=delta)
        else:
            rc, ns, st, _ = B.discrete_geodesic(v0, P[e], delta=delta)
        got = st[0, 1:ns[0]]
        assert len(got) == len(gold), (s, e, len(got), len(gold))
        assert np.max(np.abs(got - gold)) < tol, (s, e, np.max(np.abs(got - gold)))
        rows += len(gold)
    assert rows == (21 if name == "Wine_Bottle" else 4)


def test_interpolate_restatements_agree():
    from closed_chain_motion_planner_b200 import KinematicChainSpace

    _, A, B = make_oracles("stefan")
    sp = KinematicChainSpace(14)
    rng = np.random.default_rng(0)
    a = rng.uniform(-3.5, 3.5, (200, 14))
    b = rng.uniform(-3.5, 3.5, (200, 14))
    for t in (0.0, 0.1, 0.5, 1.0):
        for i in range(0, 200, 7):
            want = A.interpolate(a[i], b[i], t)
            assert np.allclose(sp.interpolate(a[i], b[i], t), want, atol=1e-15)
    assert abs(sp.distance(a[0], b[0]) - np.linalg.norm(a[0] - b[0])) < 1e-14
    x = np.array([0.5, -3.0, 3.5, -3.5, np.pi, -np.pi, 7.0])
    y = x.copy()
    sp.enforceBounds(y)
    assert np.array_equal(y, A.enforce_bounds(x))
    assert sp.equalStates(x, x + 5e-11) and not sp.equalStates(x, x + 2e-10)
