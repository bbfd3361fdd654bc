"""Copies the reference's only golden artefacts for the projection path into tests/golden/.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed copies.

Sources:
  debug/dumbbell_path.txt, debug/Wine_Bottle_path.txt — path.printAsMatrix dumps written at
  src/base/constraints/ConstrainedPlanningCommon.cpp:217-221 after path.interpolate(): every
  non-duplicated row is an output of KinematicChainConstraint::project inside discreteGeodesic.
"""
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("dumbbell_path.txt", "Wine_Bottle_path.txt"):
    shutil.copyfile(os.path.join(REF, "debug", name), os.path.join(HERE, name))
    print("copied", name)
