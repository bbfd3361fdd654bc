"""The C++ mirror of the reference's state-space seam (ProjectedStateSpace.hpp: sampler pool, discreteGeodesic, goal IK)
driven from a compiled C++ program and checked against the CPU oracles."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_oracles

pytestmark = pytest.mark.gpu


def test_cpp_state_space_program(tmp_path):
    from oracle.oracle import OracleA

    cfg, A, B = make_oracles("stefan")
    exe = tmp_path / "test_state_space"
    lib_dir = os.path.join(ROOT, "closed_chain_motion_planner_b200", "csrc")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_state_space.cpp"), "-o", str(exe),
                           "-L", lib_dir, "-lccp", f"-Wl,-rpath,{lib_dir}"])
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(cfg.start.tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = open(tmp_path / "out.bin", "rb").read()
    off = 0

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off)
        off += a.nbytes
        return a

    S, E, MS, NT = 300, 64, 48, 32
    goal_q, goal_ok, goal_sat = take(np.float64, 28).reshape(2, 14), take(np.uint8, 2), take(np.uint8, 2)
    # goal sampling: "object at the world origin" sends both arms back to the start; 1 cm along x is another closed pose
    assert goal_ok.all() and goal_sat.all()
    assert np.max(np.abs(goal_q[0] - cfg.start)) < 1e-3
    f_goal = A.function(goal_q)
    assert f_goal[:, 0].max() < 1e-4 and f_goal[:, 1].max() < 1e-4
    ahead = take(np.float64, 700 * 14).reshape(700, 14)
    ahead_refills = int(take(np.int64, 1)[0])
    failed_too = take(np.float64, 120 * 14).reshape(120, 14)
    samples = take(np.float64, S * 14).reshape(S, 14)
    sat, jv = take(np.uint8, S), take(np.uint8, S)
    refills = int(take(np.int64, 1)[0])
    near_ok, gauss_ok, reached0 = take(np.uint8, 3)
    near_state, gauss_state = take(np.float64, 14), take(np.float64, 14)
    reached, n_states = take(np.uint8, E), take(np.int32, E)
    states = take(np.float64, E * MS * 14).reshape(E, MS, 14)
    n_one = int(take(np.int32, 1)[0])
    targets = take(np.float64, NT * 12).reshape(NT, 3, 4)
    qbest, ikok = take(np.float64, NT * 7).reshape(NT, 7), take(np.uint8, NT)

    # sampler: every popped state is on the manifold, judged by the reference-faithful oracle; it is inside the joint
    # limits too, except for the reference's own quirk: enforceBounds (KinematicChain.h:118-130) wraps a joint-6 value
    # in (pi, 3.7525] to a negative one, which jointValid then rejects
    assert sat.all() and refills >= 1
    f = A.function(samples)
    assert np.all(f[:, 0] <= 1e-3 * (1 + 1e-9)) and np.all(f[:, 1] < 5e-3 * (1 + 1e-9))
    assert np.all(np.abs(samples) <= np.pi)  # the wrap was applied
    unwrapped = samples.copy()
    for j in (5, 12):
        quirk = samples[:, j] < -0.0175
        unwrapped[quirk, j] += 2 * np.pi
        assert np.all((unwrapped[quirk, j] > np.pi) & (unwrapped[quirk, j] <= 3.7525))
    assert np.all(A.joint_valid(unwrapped) == 1)
    assert np.array_equal(jv, A.joint_valid(samples))
    # the pool is the compacted ok-states of the counter-based seed stream: the same SET as projecting the stream's
    # first 2000 seeds with the host build of the engine and wrapping (order within a refill is unspecified)
    seeds = A.seeds_uniform(11, 0, 2000)
    rb = B.project(seeds, nthreads=4)
    want = rb["x"][rb["ok"].astype(bool)]
    want = B.enforce_bounds(want).reshape(-1, 14)
    first = samples[: min(S, len(want))]
    keyset = {row.tobytes() for row in want}
    assert all(row.tobytes() in keyset for row in first)
    # the prefetching sampler walks the same stream pool by pool: its first pool is the same set, the second the ok
    # states of seeds 2000 .. 3999
    k1 = len(want)
    assert ahead_refills >= 2 and k1 < 700
    assert {r_.tobytes() for r_ in ahead[:k1]} == keyset
    rb2 = B.project(A.seeds_uniform(11, 2000, 2000), nthreads=4)
    want2 = B.enforce_bounds(rb2["x"][rb2["ok"].astype(bool)]).reshape(-1, 14)
    keyset2 = {row.tobytes() for row in want2}
    assert all(r_.tobytes() in keyset2 for r_ in ahead[k1:min(700, k1 + len(want2))])
    # returnFailed: every seed's wrapped last iterate, in stream order (what the reference's sampler hands out)
    rb3 = B.project(A.seeds_uniform(11, 0, 120), nthreads=4)
    assert np.array_equal(failed_too.view(np.uint64), B.enforce_bounds(rb3["x"]).reshape(-1, 14).view(np.uint64))
    assert 0 < rb3["ok"].mean() < 1
    # near / gaussian draws around the start project back onto the manifold
    assert near_ok and gauss_ok
    for s_ in (near_state, gauss_state):
        fs = A.function(s_[None, :])[0]
        assert fs[0] <= 1e-3 * (1 + 1e-9) and fs[1] < 5e-3 * (1 + 1e-9)
    # geodesics: against the host twin of the kernel, bit for bit
    frm = np.repeat(cfg.start[None, :], E, axis=0)
    rc_b, ns_b, st_b, it_b = B.discrete_geodesic(frm, samples[:E], delta=0.25, lam=2.0, max_states=MS)
    assert np.array_equal(reached, rc_b) and np.array_equal(n_states, ns_b)
    for e in range(E):
        assert np.array_equal(states[e, : n_states[e]].view(np.uint64), st_b[e, : ns_b[e]].view(np.uint64))
    assert n_one == n_states[0] and bool(reached0) == bool(reached[0])
    # goal IK: every accepted answer hits its target by the reference-faithful FK
    A0 = OracleA([0, 1])
    A0.set_arm_base(0, np.eye(4)[:3].reshape(12))
    assert ikok.mean() > 0.9
    okm = ikok.astype(bool)
    Tq = A0.arm_transform(0, qbest[okm])
    assert np.abs(Tq[:, :, 3] - targets[okm][:, :, 3]).max() <= 1e-5 * (1 + 1e-6)
