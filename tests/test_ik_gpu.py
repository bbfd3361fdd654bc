"""Batched pose IK (ccp_ik_batch / ccp_ik_sample_batch; replaces the TRAC-IK calls of ik_task.cpp:16-49 and the goal
sampler's per-arm loop, jy_ConstrainedValidStateSampler.h:63-189).  TRAC-IK's arithmetic is third party; the parity
criterion is the one an IK answer has: FK(q) — evaluated by the REFERENCE-FAITHFUL FK of oracle A — hits the target
within TRAC-IK's tolerance inside the joint limits.  The device solver is also bit-exact against its host build."""
import numpy as np
import pytest

from conftest import make_oracles

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

LB = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
UB = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])


@pytest.fixture(scope="module")
def setup():
    import closed_chain_motion_planner_b200 as pkg
    from oracle.oracle import OracleA

    assert torch.cuda.is_available()
    pm = pkg.PandaModel()
    cfg, A, B = make_oracles("dumbbell")
    A0 = OracleA([0, 1])
    A0.set_arm_base(0, np.eye(4)[:3].reshape(12))  # FK in the arm's own base frame
    rng = np.random.default_rng(0)
    q_true = LB + (UB - LB) * rng.uniform(0.08, 0.92, (4000, 7))  # reachable targets, away from the limits
    T = A0.arm_transform(0, q_true)  # (count, 3, 4) by the reference-faithful FK
    return pm, A0, B, rng, q_true, T


def _pose_error(A0, q, T):
    Tq = A0.arm_transform(0, q)
    ep = np.abs(Tq[:, :, 3] - T[:, :, 3]).max(axis=1)
    R = np.einsum("nij,nkj->nik", T[:, :, :3], Tq[:, :, :3])  # R_t R^T
    ang = np.arccos(np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2, -1, 1))
    return ep, ang


def test_ik_solutions_hit_the_target(setup):
    pm, A0, B, rng, q_true, T = setup
    seeds = np.clip(q_true + 0.4 * rng.standard_normal(q_true.shape), LB, UB)
    r = pm.ikBatch(T, seeds)
    ok = r["ok"].astype(bool)
    assert ok.mean() > 0.8, ok.mean()  # a single Newton run from a nearby seed (measured 0.85)
    ep, ang = _pose_error(A0, r["q"][ok], T[ok])
    assert ep.max() <= 1.0e-5 * (1 + 1e-6) and ang.max() <= 1.8e-5  # eps on every component
    assert np.all(r["q"][ok] >= LB + 1e-3 - 1e-15) and np.all(r["q"][ok] <= UB - 1e-3 + 1e-15)
    assert np.all(r["q"] >= LB) and np.all(r["q"] <= UB)  # clamped every step, failures included
    assert np.all(r["err"][ok, 0] <= 1e-5) and np.all(r["err"][ok, 1] <= 1e-5)
    # a seed that already solves the problem costs no iteration
    r0 = pm.ikBatch(T[:64], q_true[:64])
    assert np.all(r0["ok"] == 1) and np.all(r0["iters"] == 0)


def test_ik_bit_exact_vs_host_build(setup):
    pm, A0, B, rng, q_true, T = setup
    seeds = np.clip(q_true + 0.6 * rng.standard_normal(q_true.shape), LB, UB)[:1500]
    # the handle's arm 0 sits in the identity base frame; the host twin needs the same model: use the engine's
    # own FK targets so both sides see identical inputs
    from closed_chain_motion_planner_b200 import make_model_desc, ArmModel
    from oracle.oracle import OracleB

    Bm = OracleB(make_model_desc([ArmModel(name="panda", index=0, t_wb=np.eye(4)), ArmModel(name="panda_b", index=1, t_wb=np.eye(4))]))
    r = pm.ikBatch(T[:1500], seeds)
    rb = Bm.ik(0, T[:1500].reshape(-1, 12), seeds)
    assert np.array_equal(r["q"].view(np.uint64), rb["q"].view(np.uint64))
    assert np.array_equal(r["ok"], rb["ok"]) and np.array_equal(r["iters"], rb["iters"])
    assert np.array_equal(r["err"].view(np.uint64), rb["err"].view(np.uint64))


def test_ik_sample_restarts(setup):
    pm, A0, B, rng, q_true, T = setup
    n = 2000
    # no reference: 15 restarts from N(mid-range, 0.3) as the reference's sampleRandomGoal does
    r = pm.ikSampleBatch(T[:n], restarts=15, rng_seed=7, sigma=0.3)
    ok = r["ok"].astype(bool)
    assert ok.mean() > 0.97, ok.mean()
    ep, ang = _pose_error(A0, r["q"][ok], T[:n][ok])
    assert ep.max() <= 1.0e-5 * (1 + 1e-6) and ang.max() <= 1.8e-5
    assert np.all(r["n_success"] <= 15) and np.all((r["n_success"] > 0) == ok)
    # deterministic in (rng_seed, target index)
    r2 = pm.ikSampleBatch(T[:n], restarts=15, rng_seed=7, sigma=0.3)
    assert np.array_equal(r["q"].view(np.uint64), r2["q"].view(np.uint64))
    # with the true configuration as reference the seeded solve wins with the reference itself
    r3 = pm.ikSampleBatch(T[:n], restarts=15, rng_seed=7, sigma=0.3, q_ref=q_true[:n])
    assert np.all(r3["ok"] == 1) and np.array_equal(r3["q"], q_true[:n])
    # (restarts still running when the seeded one succeeds are abandoned, as the reference only draws them after a failed
    # seeded solve: at least the seeded restart is counted, the answer does not depend on how many others got to finish)
    assert np.all(r3["n_success"] >= 1) and np.all(r3["n_success"] <= 15)
    # with a perturbed reference the answer is a valid solution at least as near to it as the no-reference answer
    ref = np.clip(q_true[:n] + 0.3 * rng.standard_normal((n, 7)), LB, UB)
    r4 = pm.ikSampleBatch(T[:n], restarts=16, rng_seed=7, sigma=0.3, q_ref=ref)
    ok4 = r4["ok"].astype(bool)
    assert ok4.mean() > 0.97
    r5 = pm.ikSampleBatch(T[:n], restarts=16, rng_seed=7, sigma=0.3, q_ref=ref)  # the same answer whichever lanes raced
    assert np.array_equal(r4["q"].view(np.uint64), r5["q"].view(np.uint64)) and np.array_equal(r4["ok"], r5["ok"])
    ep, ang = _pose_error(A0, r4["q"][ok4], T[:n][ok4])
    assert ep.max() <= 1.0e-5 * (1 + 1e-6) and ang.max() <= 1.8e-5


def test_ik_unreachable_target_fails_cleanly(setup):
    pm, A0, B, rng, q_true, T = setup
    Tfar = T[:32].copy()
    Tfar[:, :, 3] += np.array([3.0, 0.0, 0.0])  # 3 m away: outside the workspace
    r = pm.ikSampleBatch(Tfar, restarts=8, rng_seed=1)
    assert np.all(r["ok"] == 0) and np.all(r["n_success"] == 0)
    r1 = pm.ikBatch(Tfar, q_true[:32])
    assert np.all(r1["ok"] == 0) and np.all(r1["iters"] == 200) and np.all(np.isfinite(r1["q"]))


def test_goal_sample_batch_closes_the_chain():
    """ccp_goal_sample_batch (≙ sampleCalibGoal / sampleRandomGoal for a batch of object poses): per-arm IK on
    t_wb^-1 T_obj t_o7; every ok row is a closed-chain configuration — both arms hold the object at the pose asked for, by the
    reference-faithful FK, so the constraint residual is at IK tolerance and project() accepts the row unchanged."""
    import closed_chain_motion_planner_b200 as pkg
    from oracle.oracle import OracleA

    c = pkg.KinematicChainConstraint.from_config("stefan", device=0)
    cfg = c.config
    A = OracleA(cfg.arm_indices)
    A.set_initial_position(cfg.start)
    gp = pkg.grasping_point()
    t_o7 = c.graspFrames(cfg.t_wo_start, cfg.start)
    # the grasp frames reproduce the start: t_wb_a FK(start_a) t_o7_a^-1 == t_wo_start
    for a, ix in enumerate(cfg.arm_indices):
        Tb = np.vstack([A.arm_transform(a, cfg.start[7 * a:7 * a + 7])[0], [0, 0, 0, 1]])
        Two = gp.t_wb[ix] @ Tb @ np.linalg.inv(np.vstack([t_o7[a], [0, 0, 0, 1]]))
        assert np.max(np.abs(Two - cfg.t_wo_start)) < 1e-12
    # object poses around the start pose: small translations and rotations
    rng = np.random.default_rng(3)
    n = 3000
    from scipy.spatial.transform import Rotation

    T_obj = np.tile(cfg.t_wo_start[None, :, :], (n, 1, 1))
    T_obj[:, :3, 3] += rng.uniform(-0.06, 0.06, (n, 3))
    dR = Rotation.from_rotvec(rng.uniform(-0.15, 0.15, (n, 3))).as_matrix()
    T_obj[:, :3, :3] = np.einsum("nij,jk->nik", dR, cfg.t_wo_start[:3, :3])
    for q_ref in (cfg.start, None):
        r = c.sampleGoalBatch(T_obj, t_o7, q_ref=q_ref, restarts=15, rng_seed=11)
        ok = r["ok"].astype(bool)
        assert ok.mean() > (0.9 if q_ref is not None else 0.6), ok.mean()
        q = r["q"][ok]
        # each arm holds the object where it was asked to (reference-faithful FK)
        for a, ix in enumerate(cfg.arm_indices):
            Tq = A.arm_transform(a, q[:, 7 * a:7 * a + 7])
            Tb = np.concatenate([Tq, np.tile(np.array([[[0.0, 0, 0, 1]]]), (len(q), 1, 1))], axis=1)
            Two = gp.t_wb[ix][None] @ Tb @ np.linalg.inv(np.vstack([t_o7[a], [0, 0, 0, 1]]))[None]
            assert np.abs(Two[:, :3, 3] - T_obj[ok][:, :3, 3]).max() < 3e-5
            assert np.abs(Two[:, :3, :3] - T_obj[ok][:, :3, :3]).max() < 3e-5
        # the closed-chain constraint is satisfied at IK tolerance; project() takes 0 iterations and keeps the row
        f = A.function(q)
        assert f[:, 0].max() < 1e-4 and f[:, 1].max() < 1e-4
        assert np.all(q >= np.tile(LB, 2) + 1e-3 - 1e-12) and np.all(q <= np.tile(UB, 2) - 1e-3 + 1e-12)
        p = c.projectBatch(q)
        assert np.all(p.ok == 1) and np.all(p.iters == 0) and np.array_equal(p.x, q)
        if q_ref is not None:  # seeded: the answers stay near the reference configuration
            assert np.median(np.linalg.norm(q - cfg.start[None, :], axis=1)) < 1.0
    # an unreachable object pose fails cleanly
    far = T_obj[:16].copy()
    far[:, :3, 3] += np.array([0.0, 0.0, 3.0])
    r = c.sampleGoalBatch(far, t_o7, q_ref=cfg.start, restarts=8)
    assert np.all(r["ok"] == 0)
    with pytest.raises(pkg.CcpError):
        c.sampleGoalBatch(T_obj[:4], t_o7, restarts=0)
