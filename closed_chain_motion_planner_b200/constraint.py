"""Host-side mirror of the reference's constraint / kinematics interface over the C ABI.

Same names, argument meaning and error behaviour as the reference so parity tests read like the
reference's call sites:

  KinematicChainConstraint   include/closed_chain_motion_planner/base/constraints/ConstraintFunction.h:21-137
  PandaModel (RobotModel)    include/closed_chain_motion_planner/kinematics/panda_rbdl.h:8-77
  ArmModel                   include/closed_chain_motion_planner/kinematics/panda_model.h:7-23
  grasping_point             src/kinematics/grasping_point.cpp:5-65

plus the batched entry points the north star adds (projectBatch, functionBatch, jacobianBatch ...).
All arithmetic runs in the CUDA library (csrc/libccp.so); there is no CPU fallback here — if the
library or a GPU is missing, construction raises.  numpy arrays go through the *_host C entry
points (copies included), torch CUDA tensors through the device-pointer entry points.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import CCP_LAYOUT_AOS, CCP_LAYOUT_SOA, CcpError

_CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


# ------------------------------------------------------------------------------------------------
# configuration (grasping_point)
# ------------------------------------------------------------------------------------------------
class grasping_point:
    """Base frames of the three arms + YAML problem loader (grasping_point.cpp:5-65)."""

    def __init__(self):
        left = np.eye(4)
        left[:3, 3] = (0.0, 0.3, 1.006)
        right = np.eye(4)
        right[:3, 3] = (0.0, -0.3, 1.006)
        top = np.eye(4)
        top[:3, 3] = (1.35, 0.3, 1.006)
        top[:3, :3] = np.diag([-1.0, -1.0, 1.0])
        self.t_wb = [left, right, top]
        self.start: Optional[np.ndarray] = None
        self.obj_name = ""
        self.arm_name1 = self.arm_name2 = ""
        self.arm_index1 = self.arm_index2 = -1
        self.t_wo_start = np.eye(4)
        self.t_wo_goal = np.eye(4)
        self.arm_names: list[str] = []
        self.arm_indices: list[int] = []

    @staticmethod
    def _quat_xyzw_to_R(q):
        x, y, z, w = (float(v) for v in q)
        return np.array(
            [
                [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
            ]
        )

    def loadConfig(self, file_name: str = "stefan"):
        """file_name: a config name under configs/ (stefan, dumbbell, Wine_Bottle, ...) or a YAML path."""
        import yaml

        path = file_name if os.path.exists(file_name) else os.path.join(_CONFIG_DIR, file_name + ".yaml")
        with open(path) as f:
            y = yaml.safe_load(f)
        self.obj_name = y["obj_name"]
        self.start = np.asarray(y["start_joint"], dtype=np.float64)
        for key, dst in (("start", "t_wo_start"), ("goal", "t_wo_goal")):
            T = np.eye(4)
            T[:3, 3] = y[f"t_wo_{key}_pos"]
            T[:3, :3] = self._quat_xyzw_to_R(y[f"t_wo_{key}_quat"])  # x, y, z, w (grasping_point.cpp:40)
            setattr(self, dst, T)
        arms = []
        k = 1
        while f"arm{k}" in y:
            arms.append((y[f"arm{k}"]["name"], int(y[f"arm{k}"]["index"])))
            k += 1
        self.arm_name1, self.arm_index1 = arms[0]
        self.arm_name2, self.arm_index2 = arms[1]
        # the constraint's arm order is the alphabetical std::map order of the names
        # (ConstrainedPlanningCommon.cpp:89-91,126), not the YAML order
        arms_sorted = sorted(arms, key=lambda a: a[0])
        self.arm_names = [a[0] for a in arms_sorted]
        self.arm_indices = [a[1] for a in arms_sorted]
        return self


@dataclass
class ArmModel:
    """panda_model.h:7-23 (only the members the constraint reads)."""

    name: str = ""
    index: int = 0
    t_wb: np.ndarray = field(default_factory=lambda: np.eye(4))
    dh_offsets: Optional[np.ndarray] = None  # 7x4 (a, d, theta, alpha) calibration, panda_rbdl.cpp:92-95
    t_7e: np.ndarray = field(default_factory=lambda: np.eye(4))  # never read by the constraint
    t_o7: np.ndarray = field(default_factory=lambda: np.eye(4))


def _fill_arm_desc(A: _capi.ArmDesc, arm: ArmModel):
    al = [0.0, -math.pi / 2, math.pi / 2, math.pi / 2, -math.pi / 2, math.pi / 2, math.pi / 2]
    aa = [0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088]
    dd = [0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0]
    off = np.zeros((7, 4)) if arm.dh_offsets is None else np.asarray(arm.dh_offsets, dtype=np.float64).reshape(7, 4)
    for i in range(7):
        A.dh_a[i] = aa[i] + off[i, 0]
        A.dh_d[i] = dd[i] + off[i, 1]
        A.dh_theta_offset[i] = off[i, 2]
        A.dh_alpha[i] = al[i] + off[i, 3]
    T = np.asarray(arm.t_wb, dtype=np.float64)
    for r in range(3):
        for c in range(4):
            A.t_wb[4 * r + c] = T[r, c]
    A.flange = 0.107
    A.ee_yaw = -math.pi / 4.0


def make_model_desc(arms: Sequence[ArmModel]) -> _capi.ModelDesc:
    d = _capi.default_model_desc([0] * len(arms))
    for a, arm in enumerate(arms):
        _fill_arm_desc(d.arm[a], arm)
    return d


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _check(lib, h, rc):
    if rc != 0:
        msg = lib.ccp_last_error(h)
        raise CcpError(f"ccp error {rc}: {msg.decode() if msg else ''}")


def _host_array(shape, dtype, pinned: bool) -> np.ndarray:
    """Zeroed host array; page-locked (through torch) when `pinned`."""
    if not pinned:
        return np.zeros(shape, dtype)
    import torch

    return torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory().numpy()


@dataclass
class ProjectResult:
    x: object  # projected states (same container type / layout as the input)
    ok: object  # project()'s return value per state: converged AND jointValid
    converged: object
    iters: object
    resid: object


@dataclass
class CompactResult:
    states: np.ndarray  # (capacity, n): the first min(n_ok, capacity) rows are the batch's ok states, order unspecified
    index: Optional[np.ndarray]  # (capacity,) int32: seed index of each packed row
    ok: Optional[np.ndarray]
    iters: Optional[np.ndarray]
    n_ok: int = -1  # set by waitCompactBatch


# ------------------------------------------------------------------------------------------------
# the constraint
# ------------------------------------------------------------------------------------------------
class KinematicChainConstraint:
    """ompl::base::Constraint-shaped closed-chain constraint running on the GPU.

    KinematicChainConstraint(links) -> ambient dimension `links` (14 or 21), co-dimension 2*(links/7 - 1).
    """

    def __init__(self, links: int = 14, device: int = 0):
        if links not in (14, 21):
            raise ValueError("links must be 14 (two arms) or 21 (three arms)")
        self.n_ = int(links)
        self.k_ = self.n_ // 7
        self.device = int(device)
        self._lib = _capi.load_library()
        self._h = C.c_void_p()
        self._arms: list[ArmModel] = []
        self.lb_ = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
        self.ub_ = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])

    # -- construction ---------------------------------------------------------------------------
    def setArmModels(self, *arms: ArmModel):
        """ConstraintFunction.h:122-126 (two arms in the reference; three for the 21-DoF extension)."""
        if len(arms) != self.k_:
            raise ValueError(f"expected {self.k_} arm models")
        self._arms = list(arms)
        self._destroy()
        desc = make_model_desc(self._arms)
        h = C.c_void_p()
        rc = self._lib.ccp_create(C.byref(desc), self.device, C.byref(h))
        if rc != 0:
            raise CcpError(f"ccp_create failed ({rc}): {self._lib.ccp_last_error(None).decode()}")
        self._h = h
        self._desc = desc

    @classmethod
    def from_config(cls, name: str, device: int = 0) -> "KinematicChainConstraint":
        """Everything ConstrainedProblem::_setEnvironment + setConstrainedOptions do for the constraint
        (ConstrainedPlanningCommon.cpp:85-132): arm models in map order, init chain from start_joint,
        tolerances 1e-3 / 5e-3."""
        cfg = grasping_point().loadConfig(name)
        c = cls(7 * len(cfg.arm_indices), device)
        c.setArmModels(*[ArmModel(name=nm, index=ix, t_wb=cfg.t_wb[ix]) for nm, ix in zip(cfg.arm_names, cfg.arm_indices)])
        c.setInitialPosition(cfg.start)
        c.setTolerance(0.001, 0.005)
        c.config = cfg
        return c

    def _need(self):
        if not self._h:
            raise CcpError("setArmModels() must be called first")

    def _destroy(self):
        if getattr(self, "_h", None):
            self._lib.ccp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    # -- OMPL Constraint surface ----------------------------------------------------------------
    def getAmbientDimension(self) -> int:
        return self.n_

    def getCoDimension(self) -> int:
        return 2 * (self.k_ - 1)

    def setInitialPosition(self, init_joint):
        """ConstraintFunction.h:31-40."""
        self._need()
        q = np.ascontiguousarray(init_joint, dtype=np.float64)
        if q.shape != (self.n_,):
            raise ValueError(f"init_joint must have {self.n_} entries")
        _check(self._lib, self._h, self._lib.ccp_set_reference(self._h, q.ctypes.data))

    def getInitChain(self, pair: int = 0):
        t = np.zeros(3)
        q = np.zeros(4)
        _check(self._lib, self._h, self._lib.ccp_get_reference(self._h, pair, t.ctypes.data, q.ctypes.data))
        return t, q

    def setTolerance(self, tolerance1: float, tolerance2: float):
        """ConstraintFunction.h:104-112; the reference throws ompl::Exception for non-positive values."""
        self._need()
        if tolerance1 <= 0 or tolerance2 <= 0:
            raise ValueError("ompl::base::Constraint::setProjectionTolerance(): tolerance must be positive.")
        _check(self._lib, self._h, self._lib.ccp_set_tolerance(self._h, tolerance1, tolerance2))

    def setMaxIterations(self, n: int):
        """The reference's setMaxIterations(1000) (ConstrainedPlanningCommon.cpp:129) sets an OMPL base
        member the loop never reads; the loop uses the private 250 (ConstraintFunction.h:26).  Kept as a
        no-op for drop-in compatibility; use setOptions(max_iter=...) to really change the cap."""
        self._ompl_max_iterations = int(n)

    def setOptions(self, step: float = 0.30, max_iter: int = 250, joint_margin: float = 1e-3, damping: float = 0.0,
                   clamp: bool = False):
        """step / cap / margin of the reference (ConstraintFunction.h:71,26,45) plus two opt-in modes that the
        reference does not have and parity runs keep off: `damping` = lambda^2 of a damped-least-squares step,
        `clamp` = clamp every iterate to the joint limits."""
        self._need()
        o = _capi.Options(step=step, max_iter=max_iter, clamp=1 if clamp else 0, joint_margin=joint_margin,
                          damping=damping)
        _check(self._lib, self._h, self._lib.ccp_set_options(self._h, C.byref(o)))

    def function(self, x, out=None):
        """ConstraintFunction.h:84-102: out = (err_p, err_r) [per chain]."""
        f = self.functionBatch(np.asarray(x, dtype=np.float64).reshape(1, self.n_))[0]
        if out is not None:
            out[...] = f
        return f

    def jacobian(self, x, out=None):
        """ompl::base::Constraint::jacobian (called at ConstraintFunction.h:70), analytic: (m, n)."""
        J = self.jacobianBatch(np.asarray(x, dtype=np.float64).reshape(1, self.n_))[0]
        if out is not None:
            out[...] = J
        return J

    def project(self, x) -> bool:
        """ConstraintFunction.h:57-82: x (writable float64 array of n) is updated IN PLACE, also on failure."""
        xa = np.asarray(x)
        if xa.dtype != np.float64 or xa.shape != (self.n_,) or not xa.flags.writeable:
            raise ValueError("project() needs a writable float64 vector of the ambient dimension")
        r = self.projectBatch(xa.reshape(1, self.n_))
        xa[...] = r.x[0]
        return bool(r.ok[0])

    def isSatisfied(self, x) -> bool:
        """ConstraintFunction.h:114-120."""
        return bool(self.isSatisfiedBatch(np.asarray(x, dtype=np.float64).reshape(1, self.n_))[0])

    def jointValid(self, q) -> bool:
        """ConstraintFunction.h:43-55."""
        return bool(self.jointValidBatch(np.asarray(q, dtype=np.float64).reshape(1, self.n_))[0])

    # -- batched entry points -------------------------------------------------------------------
    def _dev_args(self, X, layout):
        import torch

        if X.dtype != torch.float64 or not X.is_cuda or not X.is_contiguous():
            raise ValueError("device states must be contiguous float64 CUDA tensors")
        if X.device.index != self.device:
            raise ValueError("tensor is on another device than the constraint")
        if layout == CCP_LAYOUT_AOS:
            if X.dim() != 2 or X.shape[1] != self.n_:
                raise ValueError(f"AOS states must have shape (count, {self.n_})")
            count = X.shape[0]
        else:
            if X.dim() != 2 or X.shape[0] != self.n_:
                raise ValueError(f"SOA states must have shape ({self.n_}, count)")
            count = X.shape[1]
        return count, torch.cuda.current_stream(X.device).cuda_stream

    def projectBatch(self, X, layout: int = CCP_LAYOUT_AOS, out=None, want_resid: bool = True,
                     compact=None, n_ok=None, pipelined: bool = False, pinned: bool = False) -> ProjectResult:
        """Batched project().  numpy (count, n) -> host path; torch CUDA tensor -> device path (async on the
        current stream).  `out` (device path) may alias X for in-place projection.  Device path only:
        `compact` ((>=count, n) float64) receives the ok states densely packed by the kernel epilogue and
        `n_ok` (int64[1], zeroed by the caller) their count.  `pipelined=True` (device path) parks the samples
        still iterating when the batch runs dry instead of idling the GPU on them: the result tensors are
        complete only after the next non-pipelined projectBatch or flush() (ccp_project_batch_pipelined).
        `pinned=True` (host path) returns the results in page-locked arrays: the library then copies each chunk out the
        moment its last sample has finished instead of a fixed number of launches later."""
        self._need()
        m = self.getCoDimension()
        if _is_torch(X):
            import torch

            count, stream = self._dev_args(X, layout)
            xo = torch.empty_like(X) if out is None else out
            ok = torch.empty(count, dtype=torch.uint8, device=X.device)
            cv = torch.empty(count, dtype=torch.uint8, device=X.device)
            it = torch.empty(count, dtype=torch.int32, device=X.device)
            rs = None
            if want_resid:
                shape = (count, m) if layout == CCP_LAYOUT_AOS else (m, count)
                rs = torch.empty(shape, dtype=torch.float64, device=X.device)
            fn = self._lib.ccp_project_batch_pipelined if pipelined else self._lib.ccp_project_batch
            _check(self._lib, self._h, fn(
                self._h, X.data_ptr(), count, layout, xo.data_ptr(), ok.data_ptr(), cv.data_ptr(), it.data_ptr(),
                rs.data_ptr() if rs is not None else None,
                compact.data_ptr() if compact is not None else None,
                n_ok.data_ptr() if n_ok is not None else None, stream))
            return ProjectResult(xo, ok, cv, it, rs)
        X = np.ascontiguousarray(X, dtype=np.float64)
        if layout != CCP_LAYOUT_AOS:
            raise ValueError("host path is AOS only")
        if X.ndim != 2 or X.shape[1] != self.n_:
            raise ValueError(f"states must have shape (count, {self.n_})")
        count = X.shape[0]
        xo = _host_array((count, self.n_), np.float64, pinned)
        ok = _host_array((count,), np.uint8, pinned)
        cv = _host_array((count,), np.uint8, pinned)
        it = _host_array((count,), np.int32, pinned)
        rs = _host_array((count, m), np.float64, pinned)
        _check(self._lib, self._h, self._lib.ccp_project_batch_host(
            self._h, X.ctypes.data, count, xo.ctypes.data, ok.ctypes.data, cv.ctypes.data, it.ctypes.data,
            rs.ctypes.data))
        return ProjectResult(xo, ok, cv, it, rs)

    def submitHostBatch(self, X: np.ndarray, want_resid: bool = False, pinned: bool = False):
        """Streaming host path (ccp_project_batch_host_submit): enqueue a host batch and return (ticket, result); the
        result arrays are complete after waitHostBatch(ticket).  Submit the next batch before waiting for this one and
        the GPU never idles on a batch's stragglers nor on its copies.  X must stay alive until the wait."""
        self._need()
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.n_:
            raise ValueError(f"states must have shape (count, {self.n_})")
        count = X.shape[0]
        xo = _host_array((count, self.n_), np.float64, pinned)
        ok = _host_array((count,), np.uint8, pinned)
        cv = _host_array((count,), np.uint8, pinned)
        it = _host_array((count,), np.int32, pinned)
        rs = _host_array((count, self.getCoDimension()), np.float64, pinned) if want_resid else None
        t = C.c_int64(0)
        _check(self._lib, self._h, self._lib.ccp_project_batch_host_submit(
            self._h, X.ctypes.data, count, xo.ctypes.data, ok.ctypes.data, cv.ctypes.data, it.ctypes.data,
            rs.ctypes.data if rs is not None else None, C.byref(t)))
        res = ProjectResult(xo, ok, cv, it, rs)
        res._keepalive = X
        return int(t.value), res

    def waitHostBatch(self, ticket: int):
        self._need()
        _check(self._lib, self._h, self._lib.ccp_project_batch_host_wait(self._h, int(ticket)))

    def submitCompactBatch(self, X: Optional[np.ndarray] = None, *, sampler=None, count: Optional[int] = None,
                           capacity: Optional[int] = None, want_index: bool = True, want_flags: bool = False,
                           pinned: bool = True):
        """General streaming host batch (ccp_host_batch_submit) with COMPACT outputs: only the states with
        project() == true come back, densely packed, plus (optionally) the seed index of each packed row and the per-seed
        ok / iteration arrays.  Either X (host states, (count, n)) or `sampler` (_capi.SamplerArgs: the seeds are generated
        on the device, nothing is copied in) with `count`.  Returns (ticket, CompactResult); call waitCompactBatch(ticket,
        result) to complete it: result.n_ok rows of result.states / result.index are then valid."""
        self._need()
        b = _capi.HostBatch()
        keep = []
        if (X is None) == (sampler is None):
            raise ValueError("give either X or sampler")
        if X is not None:
            X = np.ascontiguousarray(X, dtype=np.float64)
            if X.ndim != 2 or X.shape[1] != self.n_:
                raise ValueError(f"states must have shape (count, {self.n_})")
            count = X.shape[0]
            b.seeds_host = X.ctypes.data
            keep.append(X)
        else:
            if count is None:
                raise ValueError("count is required with sampler arguments")
            b.sampler = C.pointer(sampler)
            keep.append(sampler)
        cap = int(count if capacity is None else capacity)
        res = CompactResult(states=_host_array((cap, self.n_), np.float64, pinned),
                            index=_host_array((cap,), np.int32, pinned) if want_index else None,
                            ok=_host_array((count,), np.uint8, pinned) if want_flags else None,
                            iters=_host_array((count,), np.int32, pinned) if want_flags else None)
        b.count = int(count)
        b.compact_host = res.states.ctypes.data
        b.compact_index_host = res.index.ctypes.data if want_index else None
        b.compact_capacity = cap
        b.ok_host = res.ok.ctypes.data if want_flags else None
        b.iters_host = res.iters.ctypes.data if want_flags else None
        t = C.c_int64(0)
        _check(self._lib, self._h, self._lib.ccp_host_batch_submit(self._h, C.byref(b), C.byref(t)))
        res._keepalive = keep
        return int(t.value), res

    def waitCompactBatch(self, ticket: int, result: "CompactResult") -> int:
        self._need()
        nk = C.c_int64(-1)
        _check(self._lib, self._h, self._lib.ccp_host_batch_wait(self._h, int(ticket), C.byref(nk)))
        result.n_ok = int(nk.value)
        return result.n_ok

    def flush(self, compact=None, n_ok=None, stream=None):
        """Completes the samples parked by pipelined projections (ccp_project_flush); async on the current stream."""
        self._need()
        if stream is None:
            import torch

            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._lib, self._h, self._lib.ccp_project_flush(
            self._h, compact.data_ptr() if compact is not None else None,
            n_ok.data_ptr() if n_ok is not None else None, stream))

    def pipelineOpen(self) -> bool:
        self._need()
        return self._lib.ccp_project_pipeline_open(self._h) == 1

    def functionBatch(self, X, layout: int = CCP_LAYOUT_AOS):
        self._need()
        m = self.getCoDimension()
        if _is_torch(X):
            import torch

            count, stream = self._dev_args(X, layout)
            f = torch.empty((count, m) if layout == CCP_LAYOUT_AOS else (m, count), dtype=torch.float64, device=X.device)
            _check(self._lib, self._h, self._lib.ccp_function_batch(self._h, X.data_ptr(), count, layout, f.data_ptr(), stream))
            return f
        X = np.ascontiguousarray(X, dtype=np.float64)
        f = np.zeros((X.shape[0], m))
        _check(self._lib, self._h, self._lib.ccp_function_batch_host(self._h, X.ctypes.data, X.shape[0], f.ctypes.data))
        return f

    def jacobianBatch(self, X, layout: int = CCP_LAYOUT_AOS):
        self._need()
        m = self.getCoDimension()
        if _is_torch(X):
            import torch

            count, stream = self._dev_args(X, layout)
            shape = (count, m, self.n_) if layout == CCP_LAYOUT_AOS else (m, self.n_, count)
            J = torch.empty(shape, dtype=torch.float64, device=X.device)
            _check(self._lib, self._h, self._lib.ccp_jacobian_batch(self._h, X.data_ptr(), count, layout, J.data_ptr(), stream))
            return J
        X = np.ascontiguousarray(X, dtype=np.float64)
        J = np.zeros((X.shape[0], m, self.n_))
        _check(self._lib, self._h, self._lib.ccp_jacobian_batch_host(self._h, X.ctypes.data, X.shape[0], J.ctypes.data))
        return J

    def _flags(self, fn, X, layout):
        import torch

        host = not _is_torch(X)
        if host:
            X = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(f"cuda:{self.device}")
        count, stream = self._dev_args(X, layout)
        out = torch.empty(count, dtype=torch.uint8, device=X.device)
        _check(self._lib, self._h, fn(self._h, X.data_ptr(), count, layout, out.data_ptr(), stream))
        return out.cpu().numpy() if host else out

    def isSatisfiedBatch(self, X, layout: int = CCP_LAYOUT_AOS):
        self._need()
        return self._flags(self._lib.ccp_is_satisfied_batch, X, layout)

    def jointValidBatch(self, X, layout: int = CCP_LAYOUT_AOS):
        self._need()
        return self._flags(self._lib.ccp_joint_valid_batch, X, layout)

    # -- goal sampling (jy_ValidStateSampler::sampleCalibGoal / sampleRandomGoal, batched) ---------
    def graspFrames(self, t_wo_start, start_joint) -> np.ndarray:
        """t_o7 of every arm as the reference derives it (ConstrainedPlanningCommon.cpp:105-111):
        t_o7_a = t_wo_start^-1 * t_wb_a * getTransform(start segment of arm a).  Returns (K, 3, 4)."""
        self._need()
        q = np.ascontiguousarray(start_joint, dtype=np.float64).reshape(self.k_, 7)
        Two_inv = np.linalg.inv(np.asarray(t_wo_start, dtype=np.float64).reshape(4, 4))
        out = np.zeros((self.k_, 3, 4))
        for a in range(self.k_):
            T = np.zeros(12)
            _check(self._lib, self._h, self._lib.ccp_arm_fk_batch_host(self._h, a, q[a].ctypes.data, 1, T.ctypes.data, None))
            Tb = np.vstack([T.reshape(3, 4), [0, 0, 0, 1]])
            out[a] = (Two_inv @ np.asarray(self._arms[a].t_wb, dtype=np.float64) @ Tb)[:3]
        return out

    def sampleGoalBatch(self, T_obj, t_o7, q_ref=None, restarts: int = 15, rng_seed: int = 0, sigma: float = 0.3, **ik_opts):
        """Closed-chain goal configurations for a batch of object poses (ccp_goal_sample_batch): per arm the IK target is
        t_wb^-1 * T_obj * t_o7 (ik_task.cpp:16-27), `restarts` solves side by side, the seeded one (q_ref's segment) wins,
        else the nearest success.  T_obj: (n, 3, 4) or (n, 4, 4) world poses; t_o7: (K, 3, 4) (graspFrames); q_ref: (n, 7K)
        or a single 7K-vector or None (sampleRandomGoal).  Returns dict(q=(n, 7K), ok=(n,)) as numpy arrays."""
        self._need()
        T = np.ascontiguousarray(np.asarray(T_obj, dtype=np.float64)[..., :3, :].reshape(-1, 12))
        n = T.shape[0]
        to7 = np.ascontiguousarray(np.asarray(t_o7, dtype=np.float64)[..., :3, :].reshape(self.k_, 12))
        ref = None
        if q_ref is not None:
            ref = np.asarray(q_ref, dtype=np.float64)
            ref = np.ascontiguousarray(np.broadcast_to(ref.reshape(-1, self.n_), (n, self.n_)))
        o = _capi.IkOptions()
        self._lib.ccp_ik_default_options(C.byref(o))
        for k, v in ik_opts.items():
            setattr(o, k, v)
        q = np.zeros((n, self.n_))
        ok = np.zeros(n, np.uint8)
        _check(self._lib, self._h, self._lib.ccp_goal_sample_batch_host(
            self._h, T.ctypes.data, n, to7.ctypes.data, ref.ctypes.data if ref is not None else None, restarts, rng_seed, sigma,
            C.byref(o), q.ctypes.data, ok.ctypes.data))
        return dict(q=q, ok=ok)

    # -- measurement helpers --------------------------------------------------------------------
    def fp64PeakProbe(self, repeats: int = 5):
        self._need()
        fl = C.c_double()
        ms = C.c_double()
        _check(self._lib, self._h, self._lib.ccp_fp64_peak_probe(self._h, repeats, C.byref(fl), C.byref(ms)))
        return fl.value, ms.value

    def launchCount(self) -> int:
        return int(self._lib.ccp_launch_count(self._h)) if self._h else 0

    def algorithmicFlops(self):
        self._need()
        a = C.c_double()
        b = C.c_double()
        _check(self._lib, self._h, self._lib.ccp_algorithmic_flops(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


# ------------------------------------------------------------------------------------------------
# kinematics model (RobotModel / PandaModel)
# ------------------------------------------------------------------------------------------------
class PandaModel:
    """panda_rbdl.h:43-77.  q are 7-vectors in the arm's base frame; batched variants take (count, 7)."""

    kDof = 7

    def __init__(self, dh: Optional[np.ndarray] = None, device: int = 0):
        self._c = None
        self.device = device
        self.initModel(dh)

    def initModel(self, dh: Optional[np.ndarray] = None):
        """panda_rbdl.cpp:66-148: dh = 7x4 calibration offsets (a, d, theta, alpha) or None."""
        c = KinematicChainConstraint(14, self.device)
        arm = ArmModel(name="panda", index=0, t_wb=np.eye(4), dh_offsets=dh)
        c.setArmModels(arm, ArmModel(name="panda_b", index=1, t_wb=np.eye(4), dh_offsets=dh))
        self._c = c

    def getDof(self) -> int:
        return self.kDof

    def getJointLimit(self) -> np.ndarray:
        """panda_rbdl.cpp:44-55."""
        return np.stack([self._c.lb_, self._c.ub_], axis=1)

    def _fk(self, q, want_T, want_J):
        import torch

        q = np.ascontiguousarray(q, dtype=np.float64)
        single = q.ndim == 1
        q2 = q.reshape(-1, 7)
        dev = f"cuda:{self.device}"
        qd = torch.from_numpy(q2).to(dev)
        c = self._c
        stream = torch.cuda.current_stream(qd.device).cuda_stream
        T = J = None
        if want_T:
            Td = torch.empty((q2.shape[0], 12), dtype=torch.float64, device=dev)
            _check(c._lib, c._h, c._lib.ccp_fk_batch(c._h, 0, qd.data_ptr(), q2.shape[0], CCP_LAYOUT_AOS, Td.data_ptr(), stream))
            T = Td.cpu().numpy().reshape(-1, 3, 4)
            T = T[0] if single else T
        if want_J:
            Jd = torch.empty((q2.shape[0], 42), dtype=torch.float64, device=dev)
            _check(c._lib, c._h, c._lib.ccp_arm_jacobian_batch(c._h, 0, qd.data_ptr(), q2.shape[0], CCP_LAYOUT_AOS, Jd.data_ptr(), stream))
            J = Jd.cpu().numpy().reshape(-1, 6, 7)
            J = J[0] if single else J
        return T, J

    def getTransform(self, q) -> np.ndarray:
        """panda_rbdl.cpp:35-42 -> 4x4 homogeneous (or (count,4,4))."""
        T, _ = self._fk(q, True, False)
        if T.ndim == 2:
            return np.vstack([T, [0, 0, 0, 1]])
        bottom = np.tile(np.array([[[0.0, 0, 0, 1]]]), (T.shape[0], 1, 1))
        return np.concatenate([T, bottom], axis=1)

    def getRotation(self, q) -> np.ndarray:
        T, _ = self._fk(q, True, False)
        return T[..., :3, :3]

    def getTranslation(self, q) -> np.ndarray:
        T, _ = self._fk(q, True, False)
        return T[..., :3, 3]

    # ---- batched pose IK (goal sampling; ik_task.cpp:16-49, panda_tracik.cpp:62-88,140-158) ----
    def _ik_opts(self, **kw):
        o = _capi.IkOptions()
        self._c._lib.ccp_ik_default_options(C.byref(o))
        for k, v in kw.items():
            if v is not None:
                setattr(o, k, v)
        return o

    def ikBatch(self, targets, seeds, max_iter=None, eps_pos=None, eps_rot=None, damping=None, joint_margin=None):
        """One damped-Newton IK solve per (target, seed) pair on the GPU.  targets: (count, 3, 4) or (count, 4, 4) EE
        poses in the arm's base frame (what getTransform returns); seeds: (count, 7).  Returns dict(q, ok, iters, err)
        as numpy arrays (torch CUDA tensors in -> torch CUDA tensors out)."""
        import torch

        c = self._c
        is_t = _is_torch(targets)
        dev = torch.device("cuda", c.device)
        T = targets if is_t else torch.from_numpy(np.ascontiguousarray(targets, dtype=np.float64))
        T = T.to(dev)[..., :3, :].reshape(-1, 12).contiguous()
        q0 = seeds if _is_torch(seeds) else torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.float64))
        q0 = q0.to(dev).reshape(-1, 7).contiguous()
        cnt = T.shape[0]
        if q0.shape[0] != cnt:
            raise ValueError("one seed per target")
        q = torch.empty((cnt, 7), dtype=torch.float64, device=dev)
        ok = torch.empty(cnt, dtype=torch.uint8, device=dev)
        it = torch.empty(cnt, dtype=torch.int32, device=dev)
        err = torch.empty((cnt, 2), dtype=torch.float64, device=dev)
        o = self._ik_opts(max_iter=max_iter, eps_pos=eps_pos, eps_rot=eps_rot, damping=damping, joint_margin=joint_margin)
        _check(c._lib, c._h, c._lib.ccp_ik_batch(c._h, 0, T.data_ptr(), q0.data_ptr(), cnt, C.byref(o), q.data_ptr(),
                                                 ok.data_ptr(), it.data_ptr(), err.data_ptr(),
                                                 torch.cuda.current_stream(dev).cuda_stream))
        res = dict(q=q, ok=ok, iters=it, err=err)
        return res if is_t else {k: v.cpu().numpy() for k, v in res.items()}

    def ikSampleBatch(self, targets, restarts: int = 15, rng_seed: int = 0, sigma: float = 0.3, q_ref=None, **opts):
        """The goal sampler's per-arm loop for a batch of targets (jy_ConstrainedValidStateSampler.h:63-189): `restarts`
        solves per target side by side (restart 0 from q_ref if given, the rest from N(mid-range, sigma) clipped to the
        limits); the seeded solution wins, else the successful one nearest to q_ref (without q_ref: the lowest-numbered
        success).  Restarts that can no longer win are abandoned, so n_success counts the successful restarts that got to
        finish.  Returns dict(q, ok, n_success)."""
        import torch

        c = self._c
        is_t = _is_torch(targets)
        dev = torch.device("cuda", c.device)
        T = targets if is_t else torch.from_numpy(np.ascontiguousarray(targets, dtype=np.float64))
        T = T.to(dev)[..., :3, :].reshape(-1, 12).contiguous()
        cnt = T.shape[0]
        ref = None
        if q_ref is not None:
            ref = q_ref if _is_torch(q_ref) else torch.from_numpy(np.ascontiguousarray(q_ref, dtype=np.float64))
            ref = ref.to(dev).reshape(-1, 7).contiguous()
            if ref.shape[0] != cnt:
                raise ValueError("one reference configuration per target")
        q = torch.zeros((cnt, 7), dtype=torch.float64, device=dev)
        ok = torch.empty(cnt, dtype=torch.uint8, device=dev)
        ns = torch.empty(cnt, dtype=torch.int32, device=dev)
        o = self._ik_opts(**opts)
        _check(c._lib, c._h, c._lib.ccp_ik_sample_batch(c._h, 0, T.data_ptr(), cnt, restarts, rng_seed, sigma,
                                                        ref.data_ptr() if ref is not None else None, C.byref(o),
                                                        q.data_ptr(), ok.data_ptr(), ns.data_ptr(),
                                                        torch.cuda.current_stream(dev).cuda_stream))
        res = dict(q=q, ok=ok, n_success=ns)
        return res if is_t else {k: v.cpu().numpy() for k, v in res.items()}

    def getJacobianMatrix(self, q) -> np.ndarray:
        """panda_rbdl.cpp:9-22: 6x7, rows [linear; angular]."""
        _, J = self._fk(q, False, True)
        return J
