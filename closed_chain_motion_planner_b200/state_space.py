"""Host-side mirror of the reference's constrained state space seam, backed by the batched GPU kernels.

  KinematicChainSpace         include/closed_chain_motion_planner/kinematics/KinematicChain.h:69-174
  jy_ProjectedStateSpace      include/.../base/jy_ProjectedStateSpace.h:31-54, src/base/jy_ProjectedStateSpace.cpp:32-96
  jy_ProjectedStateSampler    src/base/jy_ProjectedStateSpace.cpp:5-29

This is the "next" row of the scope table: the callers either side of project().  The
sequential walk of one edge stays sequential; the batch is across edges (`discreteGeodesicBatch`) and across
samples (`jy_ProjectedStateSampler`, a pool refilled by ONE sample->project->wrap->compact kernel launch).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _capi
from .constraint import KinematicChainConstraint, _check

PI = math.pi


class KinematicChainSpace:
    """RealVectorStateSpace with the Panda bounds for every arm (KinematicChain.h:72-111)."""

    LOW = (-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973)
    HIGH = (2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973)

    def __init__(self, numLinks: int = 14):
        if numLinks % 7:
            raise ValueError("numLinks must be a multiple of 7")
        self.dimension_ = numLinks
        k = numLinks // 7
        self.low = np.tile(np.array(self.LOW), k)
        self.high = np.tile(np.array(self.HIGH), k)

    def getDimension(self) -> int:
        return self.dimension_

    def enforceBounds(self, state: np.ndarray) -> None:
        """KinematicChain.h:118-130: fmod wrap into [-pi, pi) IN PLACE (not a clamp)."""
        v = np.fmod(state, 2.0 * PI)
        v = np.where(v < -PI, v + 2.0 * PI, np.where(v >= PI, v - 2.0 * PI, v))
        state[...] = v

    def equalStates(self, s1, s2) -> bool:
        """KinematicChain.h:132-144."""
        return bool(np.all(np.abs(np.asarray(s1) - np.asarray(s2)) <= 1e-10))

    def distance(self, s1, s2) -> float:
        """RealVectorStateSpace::distance (inherited): Euclidean."""
        d = np.asarray(s1, dtype=np.float64) - np.asarray(s2, dtype=np.float64)
        return float(math.sqrt(float(np.dot(d, d))))

    def interpolate(self, frm, to, t: float, state: Optional[np.ndarray] = None) -> np.ndarray:
        """KinematicChain.h:145-171."""
        frm = np.asarray(frm, dtype=np.float64)
        to = np.asarray(to, dtype=np.float64)
        diff = to - frm
        near = np.abs(diff) <= PI
        d2 = np.where(diff > 0.0, 2.0 * PI - diff, -2.0 * PI - diff)
        far = frm - d2 * t
        far = np.where(far > PI, far - 2.0 * PI, np.where(far < -PI, far + 2.0 * PI, far))
        out = np.where(near, frm + diff * t, far)
        if state is not None:
            state[...] = out
            return state
        return out


@dataclass
class GeodesicResult:
    reached: object   # uint8[edges]: discreteGeodesic's return value
    n_states: object  # int32[edges]
    states: object    # float64[edges][max_states][n]; states[e][:n_states[e]] valid, states[e][0] = from
    iters: object     # int32[edges]: Newton iterations spent on the edge


class jy_ProjectedStateSpace:
    """ConstrainedStateSpace with the reference's traversal (delta, lambda from ConstrainedPlanningCommon.cpp:118-119)."""

    def __init__(self, ambientSpace: KinematicChainSpace, constraint: KinematicChainConstraint, delta: float = 0.25,
                 lam: float = 2.0):
        self.space_ = ambientSpace
        self.constraint_ = constraint
        self.delta_ = float(delta)
        self.lambda_ = float(lam)

    def setDelta(self, delta: float):
        if delta <= 0:
            raise ValueError("delta must be positive")
        self.delta_ = float(delta)

    def setLambda(self, lam: float):
        if lam <= 1:
            raise ValueError("lambda must be > 1")
        self.lambda_ = float(lam)

    def getConstraint(self) -> KinematicChainConstraint:
        return self.constraint_

    def distance(self, a, b) -> float:
        return self.space_.distance(a, b)

    def allocStateSampler(self, pool_size: int = 65536, rng_seed: int = 0) -> "jy_ProjectedStateSampler":
        """jy_ProjectedStateSpace.h:41-50."""
        return jy_ProjectedStateSampler(self, pool_size=pool_size, rng_seed=rng_seed)

    allocDefaultStateSampler = allocStateSampler

    # -- traversal ------------------------------------------------------------------------------------
    def discreteGeodesicBatch(self, frm, to, max_states: int = 64) -> GeodesicResult:
        """discreteGeodesic(from, to, interpolate=True, &geodesic) for many edges in ONE kernel launch.
        numpy (edges, n) in -> numpy out; torch CUDA tensors in -> torch out (async on the current stream)."""
        import torch

        c = self.constraint_
        c._need()
        n = c.getAmbientDimension()
        host = not type(frm).__module__.startswith("torch")
        dev = torch.device("cuda", c.device)
        if host:
            f = torch.from_numpy(np.ascontiguousarray(frm, dtype=np.float64).reshape(-1, n)).to(dev)
            t = torch.from_numpy(np.ascontiguousarray(to, dtype=np.float64).reshape(-1, n)).to(dev)
        else:
            f, t = frm, to
        if f.shape != t.shape or f.dim() != 2 or f.shape[1] != n or f.dtype != torch.float64 or not f.is_contiguous() \
                or not t.is_contiguous():
            raise ValueError(f"from/to must be contiguous float64 (edges, {n}) of equal shape")
        e = f.shape[0]
        states = torch.empty((e, max_states, n), dtype=torch.float64, device=dev)
        ns = torch.empty(e, dtype=torch.int32, device=dev)
        rc = torch.empty(e, dtype=torch.uint8, device=dev)
        it = torch.empty(e, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(c._lib, c._h, c._lib.ccp_geodesic_batch(c._h, f.data_ptr(), t.data_ptr(), e, self.delta_, self.lambda_,
                                                      max_states, states.data_ptr(), ns.data_ptr(), rc.data_ptr(),
                                                      it.data_ptr(), stream))
        if host:
            return GeodesicResult(rc.cpu().numpy(), ns.cpu().numpy(), states.cpu().numpy(), it.cpu().numpy())
        return GeodesicResult(rc, ns, states, it)

    def discreteGeodesic(self, frm, to, interpolate: bool = True, max_states: int = 256) -> Tuple[bool, List[np.ndarray]]:
        """jy_ProjectedStateSpace.cpp:32-96 for one edge: (reached, geodesic states incl. `from`).
        Only interpolate=True is offered on the device: the reference's interpolate=False additionally asks the
        MoveIt validity checker at every step (:66), which stays on the host — validate the returned states."""
        if not interpolate:
            raise NotImplementedError("interpolate=False needs the host collision checker: validate the returned states")
        r = self.discreteGeodesicBatch(np.asarray(frm, dtype=np.float64)[None, :], np.asarray(to, dtype=np.float64)[None, :],
                                       max_states=max_states)
        k = int(r.n_states[0])
        return bool(r.reached[0]), [r.states[0, i].copy() for i in range(k)]


class jy_ProjectedStateSampler:
    """WrapperStateSampler + project + enforceBounds (jy_ProjectedStateSpace.cpp:10-29), batched.

    sampleUniform() pops one state from a pool of PROJECTED states; an empty pool is refilled by one call of
    ccp_sample_project_batch (counter-based seed kernel, then the projection kernel with the optional [-pi,pi) wrap
    and the compaction in its epilogue).  Unlike the reference — which ignores project()'s return value (:13) and hands failed
    projections to the planner — the pool only holds states with project()==true.  sampleUniformNear / sampleGaussian
    draw a small batch around the given state and return the first success (or the wrapped last iterate of the first
    draw if none succeeded, which is what the reference would have returned).
    """

    def __init__(self, space: jy_ProjectedStateSpace, pool_size: int = 65536, rng_seed: int = 0, wrap_bounds: bool = True,
                 group=None, return_failed: bool = False):
        self.space_ = space
        self.constraint_ = space.getConstraint()
        self.pool_size = int(pool_size)
        self.rng_seed = int(rng_seed)
        self.wrap_bounds = bool(wrap_bounds)
        self.group = group
        # True reproduces the reference's distribution: jy_ProjectedStateSpace.cpp:13 ignores project()'s return value, so
        # failed projections (their wrapped last iterate) reach the planner too, in stream order
        self.return_failed = bool(return_failed)
        self._sharded = None
        self._next_index = 0  # position in the counter-based stream
        self._pool = np.zeros((0, self.constraint_.getAmbientDimension()))
        self._pos = 0
        self.launches = 0

    # -- batched primitives -----------------------------------------------------------------------------
    def _sample_project(self, count: int, mode: int, near=None, distance: float = 0.0, want_all: bool = False):
        import torch

        c = self.constraint_
        c._need()
        n = c.getAmbientDimension()
        dev = torch.device("cuda", c.device)
        near_arr = None if near is None else np.ascontiguousarray(near, dtype=np.float64)
        args = _capi.SamplerArgs(rng_seed=self.rng_seed, first_index=self._next_index, mode=mode,
                                 wrap_bounds=1 if self.wrap_bounds else 0, distance=float(distance),
                                 near_host=None if near_arr is None else near_arr.ctypes.data_as(C.POINTER(C.c_double)))
        self._next_index += count
        compact = torch.empty((count, n), dtype=torch.float64, device=dev)
        n_ok = torch.zeros(1, dtype=torch.int64, device=dev)
        x_all = torch.empty((count, n), dtype=torch.float64, device=dev) if want_all else None
        ok_all = torch.empty(count, dtype=torch.uint8, device=dev) if want_all else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _check(c._lib, c._h, c._lib.ccp_sample_project_batch(
            c._h, C.byref(args), count, _capi.CCP_LAYOUT_AOS, x_all.data_ptr() if want_all else None,
            ok_all.data_ptr() if want_all else None, None, compact.data_ptr(), n_ok.data_ptr(), stream))
        self.launches += 1
        k = int(n_ok.item())
        return compact[:k], x_all, ok_all

    def sampleUniformBatch(self, count: int):
        """`count` uniform seeds -> the projected states that succeeded (torch (k, n), k <= count)."""
        if self.group is not None:
            if self._sharded is None:  # created once: its symmetric-memory pool is kept across refills
                from .dist import ShardedSampleProjector

                self._sharded = ShardedSampleProjector(self.constraint_, self.group)
            states, _ = self._sharded.sample_project(
                self.rng_seed, self._next_index, count, wrap_bounds=self.wrap_bounds)
            self._next_index += count
            self.launches += 1
            return states
        return self._sample_project(count, 0)[0]

    def sampleUniformNearBatch(self, near, distance: float, count: int):
        return self._sample_project(count, 1, near, distance)[0]

    def sampleGaussianBatch(self, mean, stdDev: float, count: int):
        return self._sample_project(count, 2, mean, stdDev)[0]

    # -- OMPL StateSampler surface ------------------------------------------------------------------------
    def sampleUniform(self, state: Optional[np.ndarray] = None) -> np.ndarray:
        """jy_ProjectedStateSpace.cpp:10-15."""
        while self._pos >= len(self._pool):
            if self.return_failed:
                self._pool = self._sample_project(self.pool_size, 0, want_all=True)[1].cpu().numpy()
            else:
                self._pool = self.sampleUniformBatch(self.pool_size).cpu().numpy()
            self._pos = 0
        s = self._pool[self._pos]
        self._pos += 1
        if state is not None:
            state[...] = s
            return state
        return s.copy()

    def _first_success(self, mode, center, spread, state, tries):
        _, x_all, ok_all = self._sample_project(tries, mode, center, spread, want_all=True)
        okh = ok_all.cpu().numpy()
        first = int(np.argmax(okh)) if okh.any() else 0  # lowest stream index that succeeded: deterministic
        s = x_all[first].cpu().numpy()
        if state is not None:
            state[...] = s
            return state
        return s

    def sampleUniformNear(self, near, distance: float, state: Optional[np.ndarray] = None, tries: int = 32) -> np.ndarray:
        """jy_ProjectedStateSpace.cpp:17-22."""
        return self._first_success(1, near, distance, state, tries)

    def sampleGaussian(self, mean, stdDev: float, state: Optional[np.ndarray] = None, tries: int = 32) -> np.ndarray:
        """jy_ProjectedStateSpace.cpp:24-29."""
        return self._first_success(2, mean, stdDev, state, tries)
