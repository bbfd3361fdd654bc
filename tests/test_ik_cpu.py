"""The IK arithmetic (csrc/ccp_ik.h, host build) against the reference-faithful FK of oracle A — no GPU needed."""
import numpy as np


def test_host_ik_solutions_hit_the_target():
    from closed_chain_motion_planner_b200._capi import default_model_desc
    from oracle.oracle import OracleA, OracleB

    lb = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
    ub = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
    A = OracleA([0, 2])  # left arm at (0, 0.3, 1.006): targets in the arm's base frame = oracle A with identity base
    A.set_arm_base(0, np.eye(4)[:3].reshape(12))
    B = OracleB(default_model_desc([0, 2]))
    rng = np.random.default_rng(3)
    q_true = lb + (ub - lb) * rng.uniform(0.1, 0.9, (300, 7))
    T = A.arm_transform(0, q_true)
    seeds = np.clip(q_true + 0.3 * rng.standard_normal(q_true.shape), lb, ub)
    r = B.ik(0, T.reshape(-1, 12), seeds)
    ok = r["ok"].astype(bool)
    assert ok.mean() > 0.8
    Tq = A.arm_transform(0, r["q"][ok])
    assert np.abs(Tq[:, :, 3] - T[ok][:, :, 3]).max() <= 1e-5 * (1 + 1e-6)
    R = np.einsum("nij,nkj->nik", T[ok][:, :, :3], Tq[:, :, :3])
    ang = np.arccos(np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2, -1, 1))
    assert ang.max() <= 1.8e-5
    assert np.all(r["q"] >= lb) and np.all(r["q"] <= ub)
    # exact seeds need no iteration
    r0 = B.ik(0, T.reshape(-1, 12), q_true)
    assert np.all(r0["ok"] == 1) and np.all(r0["iters"] == 0)
