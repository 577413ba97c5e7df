"""Diagnostic: dataflow mode against the oracle at growing pile sizes (warps own one bundle per pass below ~20 k bodies, several above)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bepuphysics2_b200 import scenes
from bepuphysics2_b200.native import EXEC_DATAFLOW
from tests import util

sizes = [int(x) for x in sys.argv[1:]] or [5000, 20000, 30000, 60000, 100000]
for n in sizes:
    scene = scenes.shape_pile(n, seed=5)
    kw = dict(substeps=2, velocity_iterations=1)
    try:
        ref = util.run_oracle(util.make_sim(scene, **kw), 1 / 60.0, frames=2, threads=8, simd=True)
        got = util.run_gpu(util.make_sim(scene, **kw), 1 / 60.0, frames=2, strict=True, mode=EXEC_DATAFLOW)
        util.compare(ref, got, exact=True)
        print("bodies %d: ok (solve %.3f ms)" % (n, got["timings"]["solve_ms"]), flush=True)
    except Exception as e:  # noqa: BLE001
        print("bodies %d: FAILED %s" % (n, str(e)[:400]), flush=True)
