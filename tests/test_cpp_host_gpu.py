"""The C++ host mirror of the reference interface (ConstraintFunction.hpp) driven from a compiled C++ program,
checked bit for bit against the host build of the engine arithmetic and semantically against oracle A."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_oracles

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("count", [777, 300_000])  # 300 000: the chunked host path with page-locked results
def test_cpp_constraint_program(tmp_path, count):
    cfg, A, B = make_oracles("dumbbell")
    exe = tmp_path / "test_constraint"
    lib_dir = os.path.join(ROOT, "closed_chain_motion_planner_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_constraint.cpp"), "-o", str(exe),
                           "-L", lib_dir, "-lccp", f"-Wl,-rpath,{lib_dir}"])
    seeds = A.seeds_uniform(4, 0, count)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.int64(count).tobytes())
        f.write(cfg.start.tobytes())
        f.write(seeds.tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(tmp_path / "out.bin", "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    fx, J, x0, flags = take(np.float64, 2), take(np.float64, 28).reshape(2, 14), take(np.float64, 14), take(np.uint8, 3)
    X, ok, iters = take(np.float64, count * 14).reshape(count, 14), take(np.uint8, count), take(np.int32, count)
    T, Jg = take(np.float64, 12).reshape(3, 4), take(np.float64, 42).reshape(6, 7)
    rb = B.project(seeds, nthreads=8)
    bits = lambda a: np.ascontiguousarray(a).view(np.uint64)
    assert np.array_equal(bits(fx), bits(B.function(seeds[0])[0]))
    assert np.array_equal(bits(J), bits(B.jacobian(seeds[0])[0]))
    assert np.array_equal(bits(x0), bits(rb["x"][0])) and flags[0] == rb["ok"][0]
    assert flags[1] == A.is_satisfied(x0)[0] and flags[2] == A.joint_valid(x0)[0]
    assert np.array_equal(bits(X), bits(rb["x"])) and np.array_equal(ok, rb["ok"]) and np.array_equal(iters, rb["iters"])
    assert np.allclose(T[:, 3], [0.088, 0, 0.926], atol=1e-12)
    assert np.max(np.abs(Jg - A.arm_jacobian(0, cfg.start[:7])[0])) < 1e-14
