"""Host logic and the C-ABI boundary, without a GPU: the library loads, exports every symbol
include/ccp.h declares, and refuses loudly to compute without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_cfg

import closed_chain_motion_planner_b200 as pkg
from closed_chain_motion_planner_b200 import _capi


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ccp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ccp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_ctypes_table_agree():
    assert _declared_symbols() == sorted(_capi.SYMBOLS.keys())


def test_library_exports_every_declared_symbol():
    lib = _capi.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ccp_version()


def test_struct_sizes_match_header():
    # ccp_arm_desc: 4*7 + 12 + 2 doubles; ccp_model_desc: 2 int32 + 3 arms + 14 doubles
    assert C.sizeof(_capi.ArmDesc) == 8 * (28 + 12 + 2)
    assert C.sizeof(_capi.ModelDesc) == 8 + 3 * C.sizeof(_capi.ArmDesc) + 8 * 14
    assert C.sizeof(_capi.Options) == 32


def test_default_model_matches_python_twin():
    lib = _capi.load_library()
    d = _capi.ModelDesc()
    idx = (C.c_int32 * 2)(0, 2)
    assert lib.ccp_default_model(2, idx, C.byref(d)) == 0
    p = _capi.default_model_desc([0, 2])
    assert bytes(d) == bytes(p)
    assert lib.ccp_default_model(4, idx, C.byref(d)) != 0
    assert lib.ccp_default_model(2, (C.c_int32 * 2)(0, 5), C.byref(d)) != 0


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    lib = _capi.load_library()
    d = _capi.default_model_desc([0, 1])
    h = C.c_void_p()
    rc = lib.ccp_create(C.byref(d), 0, C.byref(h))
    assert rc == -2 and not h  # CCP_ERR_CUDA
    assert b"no CPU fallback" in lib.ccp_last_error(None)
    c = pkg.KinematicChainConstraint(14)
    with pytest.raises(pkg.CcpError):
        c.setArmModels(pkg.ArmModel("a", 0), pkg.ArmModel("b", 1))
    with pytest.raises(pkg.CcpError):
        pkg.KinematicChainConstraint(14).project(np.zeros(14))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(pkg.CcpError, match="no CPU fallback"):
        _capi.load_library(str(tmp_path / "libccp.so"))


def test_config_loader_and_arm_order():
    for name, idx in (("stefan", [0, 2]), ("dumbbell", [0, 2]), ("Wine_Bottle", [0, 1])):
        cfg = load_cfg(name)
        assert cfg.obj_name == name and cfg.start.shape == (14,)
        assert cfg.arm_indices == idx
        assert cfg.arm_names == sorted(cfg.arm_names)  # std::map order, ConstrainedPlanningCommon.cpp:89-91
    cfg = load_cfg("stefan")
    # quaternions are x,y,z,w (grasping_point.cpp:40): w = 0.4418 -> rotation of ~127.6 deg about ~z
    R = cfg.t_wo_start[:3, :3]
    assert abs(np.trace(R) - (1 + 2 * np.cos(2 * np.arccos(0.4418)))) < 2e-3
    gp = pkg.grasping_point()
    assert np.allclose(gp.t_wb[2][:3, :3], np.diag([-1, -1, 1])) and np.allclose(gp.t_wb[2][:3, 3], [1.35, 0.3, 1.006])
    assert np.allclose(gp.t_wb[1][:3, 3], [0, -0.3, 1.006])


def test_config_loader_reversed_yaml_order(tmp_path):
    """The constraint's arm order is the map (alphabetical) order even if the YAML lists them reversed."""
    src = open(os.path.join(ROOT, "configs", "dumbbell.yaml")).read()
    src = src.replace("arm1:\n  name: panda_left\n  index: 0", "arm1:\n  name: panda_top\n  index: 2", 1)
    src = src.replace("arm2:\n  name: panda_top\n  index: 2", "arm2:\n  name: panda_left\n  index: 0", 1)
    p = tmp_path / "rev.yaml"
    p.write_text(src)
    cfg = pkg.grasping_point().loadConfig(str(p))
    assert cfg.arm_name1 == "panda_top" and cfg.arm_names == ["panda_left", "panda_top"] and cfg.arm_indices == [0, 2]


def test_model_desc_with_calibration_offsets():
    dh = np.arange(28, dtype=float).reshape(7, 4) * 1e-3
    gp = pkg.grasping_point()
    d = pkg.make_model_desc([pkg.ArmModel("l", 0, gp.t_wb[0], dh), pkg.ArmModel("t", 2, gp.t_wb[2])])
    assert d.n_arms == 2
    assert abs(d.arm[0].dh_a[3] - (0.0825 + dh[3, 0])) < 1e-18
    assert abs(d.arm[0].dh_d[0] - (0.333 + dh[0, 1])) < 1e-18
    assert d.arm[0].dh_theta_offset[5] == dh[5, 2]
    assert abs(d.arm[0].dh_alpha[1] - (-np.pi / 2 + dh[1, 3])) < 1e-18
    assert d.arm[1].dh_theta_offset[5] == 0.0
    assert list(d.arm[1].t_wb)[:4] == [-1.0, 0.0, 0.0, 1.35]


def test_constraint_argument_checks_without_gpu():
    with pytest.raises(ValueError):
        pkg.KinematicChainConstraint(15)
    c = pkg.KinematicChainConstraint(14)
    assert c.getAmbientDimension() == 14 and c.getCoDimension() == 2
    assert pkg.KinematicChainConstraint(21).getCoDimension() == 4
    with pytest.raises(ValueError):
        c.setArmModels(pkg.ArmModel("a", 0))


def test_null_arguments_are_refused_not_dereferenced():
    """Entry points that can be reached without a handle return CCP_ERR_INVALID instead of crashing."""
    lib = _capi.load_library()
    t = C.c_int64(0)
    assert lib.ccp_project_batch_host_wait(None, 1) == -1
    assert lib.ccp_project_batch_host_submit(None, None, 1, None, None, None, None, None, C.byref(t)) == -1
    assert lib.ccp_project_batch_host(None, None, 1, None, None, None, None, None) == -1
    assert lib.ccp_host_alloc(None, 64) == -1
    assert lib.ccp_host_register(None, 64) == -1 and lib.ccp_host_unregister(None) == -1
    lib.ccp_host_free(None)  # a no-op
    p = C.c_void_p()
    assert lib.ccp_host_alloc(C.byref(p), 0) == 0 and not p.value
