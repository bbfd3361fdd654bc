import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CONFIGS = ("stefan", "dumbbell", "Wine_Bottle")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """CPU checkers are built on demand; the CUDA library must already be built (build())."""
    from oracle import oracle

    oracle.build()


def load_cfg(name):
    from closed_chain_motion_planner_b200 import grasping_point

    return grasping_point().loadConfig(name)


def make_oracles(name):
    from closed_chain_motion_planner_b200._capi import default_model_desc
    from oracle.oracle import OracleA, OracleB

    cfg = load_cfg(name)
    A = OracleA(cfg.arm_indices)
    A.set_initial_position(cfg.start)
    B = OracleB(default_model_desc(cfg.arm_indices))
    B.set_initial_position(cfg.start)
    return cfg, A, B


def near_manifold_seeds(cfg, count, seed=0, radius=0.25):
    """start_joint + a random step of length `radius` (= the planner's delta, one discreteGeodesic step)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((count, cfg.start.size))
    d *= radius / np.linalg.norm(d, axis=1, keepdims=True)
    return cfg.start[None, :] + d


def load_path(name):
    rows = []
    with open(os.path.join(GOLDEN, f"{name}_path.txt")) as f:
        for line in f:
            t = line.split()
            if len(t) == 14:
                rows.append([float(v) for v in t])
    return np.array(rows)
