"""Diagnostic: runs the joint zoo one constraint type at a time in dataflow mode against the oracle and reports which types fail / stall."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bepuphysics2_b200 import scenes
from bepuphysics2_b200.native import EXEC_DATAFLOW
from tests import util

for kin in (0.0, 0.05):
    for t in sorted(scenes.JOINT_ZOO_TYPES):
        scene = scenes.joint_zoo(300, 150, seed=20 + t, types=[t], kinematic_fraction=kin)
        kw = dict(substeps=2, velocity_iterations=2)
        try:
            ref = util.run_oracle(util.make_sim(scene, **kw), 1 / 60.0, frames=2)
            got = util.run_gpu(util.make_sim(scene, **kw), 1 / 60.0, frames=2, strict=True, mode=EXEC_DATAFLOW)
            util.compare(ref, got, exact=True)
            print("type %d kinematic %.2f: ok" % (t, kin), flush=True)
        except Exception as e:  # noqa: BLE001
            print("type %d kinematic %.2f: FAILED %s" % (t, kin, str(e)[:300]), flush=True)
