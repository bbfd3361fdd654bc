"""The CMake target INTEGRATION.md tells a maintainer to add (csrc/CMakeLists.txt) configures, builds for sm_100a and
exports every symbol include/ccp.h declares.  Skipped where cmake or nvcc is absent."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("cmake") is None or not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs cmake and nvcc")
def test_cmake_target_builds_and_exports_the_abi(tmp_path):
    src = os.path.join(ROOT, "closed_chain_motion_planner_b200", "csrc")
    cfg = subprocess.run(["cmake", "-S", src, "-B", str(tmp_path), "-DCMAKE_BUILD_TYPE=Release",
                          "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc"], capture_output=True, text=True)
    assert cfg.returncode == 0, cfg.stdout[-2000:] + cfg.stderr[-2000:]
    bld = subprocess.run(["cmake", "--build", str(tmp_path), "-j8"], capture_output=True, text=True)
    assert bld.returncode == 0, bld.stdout[-2000:] + bld.stderr[-2000:]
    lib = os.path.join(str(tmp_path), "libccp.so")
    assert os.path.exists(lib)
    exported = set(re.findall(r" T (ccp_\w+)", subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True,
                                                              text=True).stdout))
    header = open(os.path.join(ROOT, "include", "ccp.h")).read()
    declared = set(re.findall(r"^(?:int|void|int64_t|const char\*)\s+(ccp_\w+)\(", header, flags=re.M))
    assert declared and declared <= exported, sorted(declared - exported)
