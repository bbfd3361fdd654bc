"""ctypes front-ends of the two CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.

  OracleA — oracle_a.c: reference-faithful restatement (libm, OMPL FD stencil, SVD min-norm solve)
            of ConstraintFunction.h:31-120 + panda_rbdl.cpp:9-161.
  OracleB — oracle_b.cpp: the engine's own arithmetic header compiled for the host; exists only to
            check that device results are bit-reproducible on the host.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "build")
_ROOT = os.path.dirname(_HERE)


def build(force: bool = False) -> None:
    a = os.path.join(_BUILD, "liboracle_a.so")
    b = os.path.join(_BUILD, "liboracle_b.so")
    if force or not (os.path.exists(a) and os.path.exists(b)):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _as_states(x, n):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if x.ndim == 1:
        x = x[None, :]
    assert x.ndim == 2 and x.shape[1] == n, f"expected (count,{n}) states, got {x.shape}"
    return x


class OracleA:
    """Reference-faithful CPU oracle.  arm_index: grasping_point::t_wb order (0 left, 1 right, 2 top)."""

    def __init__(self, arm_index, dh_offsets=None):
        build()
        self.lib = C.CDLL(os.path.join(_BUILD, "liboracle_a.so"))
        self.lib.oa_model_size.restype = C.c_int
        self.lib.oa_project.restype = C.c_int
        self.lib.oa_max_threads.restype = C.c_int
        self.n_arms = len(arm_index)
        self.n = 7 * self.n_arms
        self.m = 2 * (self.n_arms - 1)
        self._buf = C.create_string_buffer(self.lib.oa_model_size())
        idx = (C.c_int * self.n_arms)(*arm_index)
        if dh_offsets is not None:
            dh = np.ascontiguousarray(dh_offsets, dtype=np.float64).reshape(self.n_arms, 7, 4)
            self.lib.oa_default_model(self._buf, C.c_int(self.n_arms), idx, _dp(dh))
        else:
            self.lib.oa_default_model(self._buf, C.c_int(self.n_arms), idx, None)

    @property
    def max_threads(self):
        return int(self.lib.oa_max_threads())

    def set_arm_base(self, arm, twb12):
        t = np.ascontiguousarray(twb12, dtype=np.float64).reshape(12)
        self.lib.oa_set_arm_base(self._buf, C.c_int(arm), _dp(t))

    def set_initial_position(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == (self.n,)
        self.lib.oa_set_initial_position(self._buf, _dp(q))

    def set_tolerance(self, t1, t2):
        self.lib.oa_set_tolerance(self._buf, C.c_double(t1), C.c_double(t2))

    def set_options(self, step=0.30, max_iter=250, margin=1e-3):
        self.lib.oa_set_options(self._buf, C.c_double(step), C.c_int(max_iter), C.c_double(margin))

    def init_chain(self, pair=0):
        R = np.zeros(9)
        t = np.zeros(3)
        self.lib.oa_get_init_chain(self._buf, C.c_int(pair), _dp(R), _dp(t))
        return R.reshape(3, 3), t

    def function(self, x, nthreads=1):
        x = _as_states(x, self.n)
        f = np.zeros((x.shape[0], self.m))
        self.lib.oa_function_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(f), C.c_int(nthreads))
        return f

    def jacobian(self, x, fd=True, nthreads=1):
        x = _as_states(x, self.n)
        J = np.zeros((x.shape[0], self.m, self.n))
        self.lib.oa_jacobian_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(J), C.c_int(1 if fd else 0),
                                   C.c_int(nthreads))
        return J

    def project(self, x, fd=True, nthreads=1):
        """Returns dict(x, ok, converged, iters, resid); input is not modified."""
        x = _as_states(x, self.n).copy()
        cnt = x.shape[0]
        ok = np.zeros(cnt, np.uint8)
        cv = np.zeros(cnt, np.uint8)
        it = np.zeros(cnt, np.int32)
        rs = np.zeros((cnt, self.m))
        self.lib.oa_project_batch(self._buf, _dp(x), C.c_int64(cnt), C.c_int(1 if fd else 0), _dp(ok), _dp(cv),
                                  _dp(it), _dp(rs), C.c_int(nthreads))
        return dict(x=x, ok=ok, converged=cv, iters=it, resid=rs)

    def joint_valid(self, x):
        x = _as_states(x, self.n)
        self.lib.oa_joint_valid.restype = C.c_int
        return np.array([self.lib.oa_joint_valid(self._buf, _dp(x[i])) for i in range(x.shape[0])], np.uint8)

    def is_satisfied(self, x):
        x = _as_states(x, self.n)
        self.lib.oa_is_satisfied.restype = C.c_int
        return np.array([self.lib.oa_is_satisfied(self._buf, _dp(x[i])) for i in range(x.shape[0])], np.uint8)

    def arm_transform(self, arm, q):
        q = _as_states(q, 7)
        T = np.zeros((q.shape[0], 12))
        for i in range(q.shape[0]):
            self.lib.oa_arm_transform(self._buf, C.c_int(arm), _dp(q[i]), _dp(T[i]))
        return T.reshape(-1, 3, 4)

    def arm_jacobian(self, arm, q):
        q = _as_states(q, 7)
        J = np.zeros((q.shape[0], 42))
        for i in range(q.shape[0]):
            self.lib.oa_arm_jacobian(self._buf, C.c_int(arm), _dp(q[i]), _dp(J[i]))
        return J.reshape(-1, 6, 7)

    def discrete_geodesic(self, frm, to, delta=0.25, lam=2.0, fd=True, max_states=64, nthreads=1):
        """jy_ProjectedStateSpace::discreteGeodesic(interpolate=true) per edge -> (reached, n_states, states)."""
        frm, to = _as_states(frm, self.n), _as_states(to, self.n)
        e = frm.shape[0]
        states = np.zeros((e, max_states, self.n))
        ns = np.zeros(e, np.int32)
        rc = np.zeros(e, np.uint8)
        self.lib.oa_discrete_geodesic_batch(self._buf, _dp(frm), _dp(to), C.c_int64(e), C.c_double(delta), C.c_double(lam),
                                            C.c_int(1 if fd else 0), C.c_int(max_states), _dp(states), _dp(ns), _dp(rc),
                                            C.c_int(nthreads))
        return rc, ns, states

    def interpolate(self, frm, to, t):
        frm = np.ascontiguousarray(frm, dtype=np.float64)
        to = np.ascontiguousarray(to, dtype=np.float64)
        out = np.zeros_like(frm)
        self.lib.oa_interpolate(_dp(frm), _dp(to), C.c_double(t), _dp(out), C.c_int(frm.size))
        return out

    def seeds_uniform(self, seed, first, count):
        x = np.zeros((count, self.n))
        self.lib.oa_seeds_uniform(self._buf, C.c_uint64(seed), C.c_int64(first), C.c_int64(count), _dp(x))
        return x

    @staticmethod
    def enforce_bounds(x):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        lib = C.CDLL(os.path.join(_BUILD, "liboracle_a.so"))
        lib.oa_enforce_bounds(_dp(x), C.c_int(x.size))
        return x


class OracleB:
    """Host build of the engine arithmetic (bitwise host/device reproducibility check only)."""

    def __init__(self, desc):
        build()
        self.lib = C.CDLL(os.path.join(_BUILD, "liboracle_b.so"))
        self.lib.ob_model_size.restype = C.c_int
        self._buf = C.create_string_buffer(self.lib.ob_model_size())
        rc = self.lib.ob_create(C.byref(desc), self._buf)
        if rc != 0:
            raise ValueError("invalid model description")
        self.n_arms = int(desc.n_arms)
        self.n = 7 * self.n_arms
        self.m = 2 * (self.n_arms - 1)

    def set_initial_position(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        assert q.shape == (self.n,)
        self.lib.ob_set_reference(self._buf, _dp(q))

    def get_reference(self, pair=0):
        t = np.zeros(3)
        q = np.zeros(4)
        self.lib.ob_get_reference(self._buf, C.c_int(pair), _dp(t), _dp(q))
        return t, q

    def set_tolerance(self, t1, t2):
        self.lib.ob_set_tolerance(self._buf, C.c_double(t1), C.c_double(t2))

    def set_options(self, step=0.30, max_iter=250, margin=1e-3, damping=0.0, clamp=False):
        self.lib.ob_set_options(self._buf, C.c_double(step), C.c_int(max_iter), C.c_double(margin))
        self.lib.ob_set_modes(self._buf, C.c_double(damping), C.c_int(1 if clamp else 0))

    def function(self, x):
        x = _as_states(x, self.n)
        f = np.zeros((x.shape[0], self.m))
        self.lib.ob_function_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(f))
        return f

    def jacobian(self, x):
        x = _as_states(x, self.n)
        J = np.zeros((x.shape[0], self.m, self.n))
        self.lib.ob_jacobian_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(J))
        return J

    def project(self, x, nthreads=1):
        x = _as_states(x, self.n).copy()
        cnt = x.shape[0]
        ok = np.zeros(cnt, np.uint8)
        cv = np.zeros(cnt, np.uint8)
        it = np.zeros(cnt, np.int32)
        rs = np.zeros((cnt, self.m))
        self.lib.ob_project_batch(self._buf, _dp(x), C.c_int64(cnt), _dp(ok), _dp(cv), _dp(it), _dp(rs),
                                  C.c_int(nthreads))
        return dict(x=x, ok=ok, converged=cv, iters=it, resid=rs)

    def joint_valid(self, x):
        x = _as_states(x, self.n)
        out = np.zeros(x.shape[0], np.uint8)
        self.lib.ob_joint_valid_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(out))
        return out

    def is_satisfied(self, x):
        x = _as_states(x, self.n)
        out = np.zeros(x.shape[0], np.uint8)
        self.lib.ob_is_satisfied_batch(self._buf, _dp(x), C.c_int64(x.shape[0]), _dp(out))
        return out

    def arm_fk(self, arm, q):
        q = _as_states(q, 7)
        T = np.zeros((q.shape[0], 12))
        J = np.zeros((q.shape[0], 42))
        self.lib.ob_arm_fk_batch(self._buf, C.c_int(arm), _dp(q), C.c_int64(q.shape[0]), _dp(T), _dp(J))
        return T.reshape(-1, 3, 4), J.reshape(-1, 6, 7)

    def ik(self, arm, targets, seeds, max_iter=200, eps_pos=1e-5, eps_rot=1e-5, damping=1e-4, margin=1e-3):
        """Host twin of ccp_ik_batch: targets (count, 12) row-major 3x4, seeds (count, 7)."""
        T = np.ascontiguousarray(targets, dtype=np.float64).reshape(-1, 12)
        q0 = _as_states(seeds, 7)
        cnt = T.shape[0]
        q = np.zeros((cnt, 7))
        ok = np.zeros(cnt, np.uint8)
        it = np.zeros(cnt, np.int32)
        err = np.zeros((cnt, 2))
        self.lib.ob_ik_batch(self._buf, C.c_int(arm), _dp(T), _dp(q0), C.c_int64(cnt), C.c_int(max_iter), C.c_double(eps_pos),
                             C.c_double(eps_rot), C.c_double(damping), C.c_double(margin), _dp(q), _dp(ok), _dp(it), _dp(err))
        return dict(q=q, ok=ok, iters=it, err=err)

    def discrete_geodesic(self, frm, to, delta=0.25, lam=2.0, max_states=64):
        frm, to = _as_states(frm, self.n), _as_states(to, self.n)
        e = frm.shape[0]
        states = np.zeros((e, max_states, self.n))
        ns = np.zeros(e, np.int32)
        rc = np.zeros(e, np.uint8)
        it = np.zeros(e, np.int32)
        self.lib.ob_geodesic_batch(self._buf, _dp(frm), _dp(to), C.c_int64(e), C.c_double(delta), C.c_double(lam),
                                   C.c_int(max_states), _dp(states), _dp(ns), _dp(rc), _dp(it))
        return rc, ns, states, it

    def seeds_uniform(self, seed, first, count):
        x = np.zeros((count, self.n))
        self.lib.ob_seeds_uniform(self._buf, C.c_uint64(seed), C.c_int64(first), C.c_int64(count), _dp(x))
        return x

    def enforce_bounds(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        self.lib.ob_enforce_bounds(_dp(x), C.c_int64(x.size))
        return x

    def sincos(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.zeros_like(x)
        c = np.zeros_like(x)
        self.lib.ob_sincos(_dp(x), C.c_int64(x.size), _dp(s), _dp(c))
        return s, c

    def atan2_pos(self, y, x):
        y = np.ascontiguousarray(y, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros_like(x)
        self.lib.ob_atan2_pos(_dp(y), _dp(x), C.c_int64(x.size), _dp(o))
        return o
