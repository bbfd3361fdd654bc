"""The OMPL adaptor a maintainer drops into the reference (include/closed_chain_motion_planner_b200/ompl_adaptor/
ConstraintFunction.h).  OMPL and Eigen are not installed here, so it is compiled against minimal
stand-ins of the two interfaces (tests/stubs/): every signature cited from ConstraintFunction.h:24,31,57,84,104,114,122
meets a compiler (CPU, -std=c++14 like the reference's CMakeLists.txt:4), and on a GPU box the compiled program is driven
through ompl::base::Constraint's virtuals and checked against the host build of the engine arithmetic and oracle A."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_oracles

LIB_DIR = os.path.join(ROOT, "closed_chain_motion_planner_b200", "csrc")
SRC = os.path.join(ROOT, "tests", "cpp", "test_ompl_adaptor.cpp")
INC = ["-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include")]


def test_adaptor_compiles_against_stub_headers():
    r = subprocess.run(["/usr/bin/g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only"] + INC + [SRC],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
def test_adaptor_through_ompl_virtuals(tmp_path):
    cfg, A, B = make_oracles("stefan")
    exe = tmp_path / "test_ompl_adaptor"
    subprocess.check_call(["/usr/bin/g++", "-std=c++14", "-O1"] + INC + [SRC, "-o", str(exe), "-L", LIB_DIR, "-lccp",
                                                                         f"-Wl,-rpath,{LIB_DIR}"])
    count = 5000
    seeds = A.seeds_uniform(6, 0, count)
    (tmp_path / "start.bin").write_bytes(cfg.start.tobytes())
    (tmp_path / "seeds.bin").write_bytes(np.int64(count).tobytes() + seeds.tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "start.bin"), str(tmp_path / "seeds.bin"), str(tmp_path / "out.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(tmp_path / "out.bin", "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    fx, J, x0, flags = take(np.float64, 2), take(np.float64, 28).reshape(2, 14), take(np.float64, 14), take(np.uint8, 3)
    single, single_ok = take(np.float64, 64 * 14).reshape(64, 14), take(np.uint8, 64)
    X, ok = take(np.float64, count * 14).reshape(count, 14), take(np.uint8, count)
    bits = lambda a: np.ascontiguousarray(a).view(np.uint64)
    rb = B.project(seeds, nthreads=8)
    assert np.array_equal(bits(fx), bits(B.function(seeds[0])[0])) and np.array_equal(bits(J), bits(B.jacobian(seeds[0])[0]))
    assert np.array_equal(bits(x0), bits(rb["x"][0])) and flags[0] == rb["ok"][0]
    assert flags[1] == A.is_satisfied(x0)[0] and flags[2] == A.joint_valid(x0)[0]
    # project(State *) one at a time == projectBatch == the host twin, failures written back too
    assert np.array_equal(bits(single), bits(rb["x"][:64])) and np.array_equal(single_ok, rb["ok"][:64])
    assert np.array_equal(bits(X), bits(rb["x"])) and np.array_equal(ok, rb["ok"])
    ra = A.project(seeds[:1000], nthreads=A.max_threads)
    assert np.mean(ra["ok"] == ok[:1000]) >= 0.999
