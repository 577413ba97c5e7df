"""One-off check at benchmark scale: the strict build against the oracle, bit for bit, on the 100 k (and optionally 1 M) body pile, several frames.
The unit tests use small scenes; this exercises the stage kernels with hundreds of CTAs per launch and several grids in flight (PDL)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np

import util
from bepuphysics2_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=100_000)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--mode", type=int, default=0)
args = ap.parse_args()
scene = scenes.shape_pile(args.bodies, seed=5)
kw = dict(substeps=8, velocity_iterations=2)
a, b = util.make_sim(scene, **kw), util.make_sim(scene, **kw)
t = time.time()
ref = util.run_oracle(a, 1 / 60.0, frames=args.frames, threads=16, simd=True)
t_cpu = time.time() - t
for rep in range(3):  # repeat the GPU run: a race would not show every time
    c = util.make_sim(scene, **kw)
    got = util.run_gpu(c, 1 / 60.0, frames=args.frames, strict=True, mode=args.mode)
    util.compare(ref, got, exact=True)
print("bit-exact: %d bodies, %d constraints, %d frames x 3 repetitions (oracle %.1f s)" % (args.bodies, a.constraint_count, args.frames, t_cpu))
