"""The engine's own sin/cos/atan2 (csrc/ccp_core.h) against libm, and the enforceBounds wrap."""
import numpy as np

from conftest import make_oracles


def _ulp_err(a, ref):
    return np.abs(a - ref) / np.spacing(np.abs(ref) + 1e-300)


def test_sincos_accuracy():
    _, A, B = make_oracles("dumbbell")
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-8, 8, 200000), rng.uniform(-1e3, 1e3, 50000),
                        np.linspace(-np.pi, np.pi, 10001), [0.0, np.pi / 2, np.pi / 4, -np.pi / 4, 1e-300, -0.0]])
    s, c = B.sincos(x)
    assert np.max(np.abs(s - np.sin(x))) < 2.3e-16 and np.max(np.abs(c - np.cos(x))) < 2.3e-16
    big = np.abs(np.sin(x)) > 1e-3
    assert np.max(_ulp_err(s[big], np.sin(x)[big])) <= 2.0
    big = np.abs(np.cos(x)) > 1e-3
    assert np.max(_ulp_err(c[big], np.cos(x)[big])) <= 2.0
    # large arguments degrade gracefully, never blow up
    xl = rng.uniform(-1e8, 1e8, 10000)
    s, c = B.sincos(xl)
    assert np.max(np.abs(s - np.sin(xl))) < 1e-7 and np.all(np.abs(s * s + c * c - 1) < 1e-12)
    s, c = B.sincos(np.array([np.nan, np.inf]))
    assert np.all(np.isnan(s)) and np.all(np.isnan(c))


def test_atan2_pos_accuracy():
    _, A, B = make_oracles("dumbbell")
    rng = np.random.default_rng(1)
    y = np.abs(rng.standard_normal(200000))
    x = np.abs(rng.standard_normal(200000))
    y[:1000] = 0.0
    x[1000:2000] = 0.0
    y[2000:3000] *= 1e-9
    x[3000:4000] *= 1e-9
    r = B.atan2_pos(y, x)
    ref = np.arctan2(y, x)
    assert np.max(np.abs(r - ref)) < 4.5e-16
    assert B.atan2_pos(np.array([0.0]), np.array([0.0]))[0] == 0.0
    assert np.isnan(B.atan2_pos(np.array([np.nan]), np.array([1.0]))[0])
    # region boundaries
    t = np.array([0.198912367379658, 0.6681786379192989, 1.0, 0.41421356237309503])
    assert np.max(np.abs(B.atan2_pos(t, np.ones(4)) - np.arctan(t))) < 2.3e-16


def test_enforce_bounds_wrap_quirk():
    """KinematicChain.h:118-130 wraps with fmod into [-pi, pi): joint 6 in (pi, 3.7525] becomes NEGATIVE
    (invalid) — kept, as the reference does it."""
    _, A, B = make_oracles("dumbbell")
    x = np.array([0.5, -3.0, 3.5, -3.5, np.pi, -np.pi, 7.0, -7.0, 3.7525, 100.0])
    wa, wb = A.enforce_bounds(x), B.enforce_bounds(x)
    assert np.array_equal(wa, wb)
    assert np.all(wa >= -np.pi) and np.all(wa < np.pi)
    assert wa[2] < 0 and abs(wa[2] - (3.5 - 2 * np.pi)) < 1e-15
    assert wa[4] == -np.pi and wa[5] == -np.pi
