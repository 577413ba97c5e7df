"""Development diagnostic: error statistics of the FMA/approx-div build against the oracle after one frame."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from bepuphysics2_b200 import scenes
from bepuphysics2_b200.native import EXEC_GRAPH
from tests import util

DT = 1 / 60
for name, scene, kw in (("pile", scenes.shape_pile(3000, seed=5), dict(substeps=8, velocity_iterations=2)),
                        ("ragdolls 1x4", scenes.ragdolls(60, seed=5), dict(substeps=1, velocity_iterations=4)),
                        ("ragdolls 8x2", scenes.ragdolls(60, seed=5), dict(substeps=8, velocity_iterations=2)),
                        ("ragdolls rest pose", scenes.ragdolls(60, seed=5, pose_noise=0.0), dict(substeps=1, velocity_iterations=4))):
    a = util.make_sim(scene, **kw)
    b = util.make_sim(scene, **kw)
    ref = util.run_oracle(a, DT)
    got = util.run_gpu(b, DT, strict=False, mode=EXEC_GRAPH)
    for label, cols in (("pose", np.r_[0:7]), ("linear v", np.r_[8:11]), ("angular v", np.r_[12:15])):
        x, y = ref["bodies"][:, cols].astype(np.float64), got["bodies"][:, cols].astype(np.float64)
        d = np.abs(x - y)
        print("%-20s %-10s rel rms %.2e  max abs %.2e  p99.9 %.2e  scale(max|x|) %.2f" % (name, label, np.sqrt((d ** 2).sum() / (x ** 2).sum()), d.max(), np.quantile(d, 0.999), np.abs(x).max()))
    imp_ref = np.concatenate([t["impulses"].ravel() for t in ref["type_batches"]]).astype(np.float64)
    imp_got = np.concatenate([t["impulses"].ravel() for t in got["type_batches"]]).astype(np.float64)
    d = np.abs(imp_ref - imp_got)
    print("%-20s %-10s rel rms %.2e  max abs %.2e  p99.9 %.2e  scale %.2f" % (name, "impulses", np.sqrt((d ** 2).sum() / (imp_ref ** 2).sum()), d.max(), np.quantile(d, 0.999), np.abs(imp_ref).max()))
