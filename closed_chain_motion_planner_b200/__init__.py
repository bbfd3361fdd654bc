"""B200-native batched closed-chain constraint projection (hot path of
jkw0701/closed_chain_motion_planner behind the reference's own constraint interface).

Public surface:
  KinematicChainConstraint, PandaModel, ArmModel, grasping_point   (reference names)
  ProjectResult, CCP_LAYOUT_AOS, CCP_LAYOUT_SOA
The CUDA library is csrc/libccp.so (build: `python -c "import __graft_entry__ as g; g.build()"`).
"""
from ._capi import CCP_LAYOUT_AOS, CCP_LAYOUT_SOA, CcpError, load_library  # noqa: F401
from .constraint import (  # noqa: F401
    ArmModel,
    KinematicChainConstraint,
    PandaModel,
    ProjectResult,
    grasping_point,
    make_model_desc,
)

from .state_space import (  # noqa: F401,E402
    GeodesicResult,
    KinematicChainSpace,
    jy_ProjectedStateSampler,
    jy_ProjectedStateSpace,
)

__all__ = [
    "GeodesicResult",
    "KinematicChainSpace",
    "jy_ProjectedStateSampler",
    "jy_ProjectedStateSpace",
    "ArmModel",
    "KinematicChainConstraint",
    "PandaModel",
    "ProjectResult",
    "grasping_point",
    "make_model_desc",
    "CCP_LAYOUT_AOS",
    "CCP_LAYOUT_SOA",
    "CcpError",
    "load_library",
]
