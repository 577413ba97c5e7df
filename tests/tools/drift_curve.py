"""Development diagnostic (CPU only): how fast does a run with FMA contraction drift away from the non-contracting evaluation?

The reference computes without FMA contraction; the library's default `bepu_fast` kernels contract. One frame of that build is
held to a tolerance by the GPU tests; over many frames a rigid-body simulation is chaotic and the two trajectories separate. This tool shows the
size of that effect with a CPU proxy: the oracle compiled with `-ffp-contract=fast -mfma` against the regular `-ffp-contract=off` build, same
scenes, same seeds (the GPU's contraction choices differ in detail, the growth rate does not).

    python tests/tools/drift_curve.py
"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

from bepuphysics2_b200 import scenes
from oracle import binding as ob
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
DT = 1 / 60


def build_contracting_oracle():
    out = os.path.join(HERE, "bin", "libbepu_oracle_fma.so")
    src = os.path.join(HERE, "..", "..", "oracle", "bepu_oracle.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=fast", "-mfma", "-fno-fast-math", "-march=x86-64-v3", "-Wno-psabi",
                           "-shared", "-o", out, src])
    lib = C.CDLL(out)
    lib.oracle_solve.argtypes = [C.POINTER(ob.OracleScene), C.c_float]
    return lib


def main():
    exact = ob.load()
    fma = build_contracting_oracle()
    marks = (1, 2, 4, 8, 16, 32, 64, 128)
    for name, scene, kw in (("pile 3000 bodies, 8 x 2", scenes.shape_pile(3000, seed=5), dict(substeps=8, velocity_iterations=2)),
                            ("60 ragdolls, 1 x 4", scenes.ragdolls(60, seed=5), dict(substeps=1, velocity_iterations=4))):
        a, b = util.make_sim(scene, **kw), util.make_sim(scene, **kw)
        print(name)
        print("  frame | rel. RMS difference: position  linear velocity  angular velocity | max |dp|")
        for frame in range(1, marks[-1] + 1):
            ob._LIB = exact
            ob.solve(a, DT, simd=True)
            ob._LIB = fma
            ob.solve(b, DT, simd=True)
            if frame in marks:
                x, y = a.bodies.astype(np.float64), b.bodies.astype(np.float64)
                rel = lambda cols: np.sqrt(((x[:, cols] - y[:, cols]) ** 2).sum() / max((x[:, cols] ** 2).sum(), 1e-300))
                print("  %5d | %.2e  %.2e  %.2e | %.2e" % (frame, rel(np.r_[4:7]), rel(np.r_[8:11]), rel(np.r_[12:15]), np.abs(x[:, 4:7] - y[:, 4:7]).max()))
        ob._LIB = exact


if __name__ == "__main__":
    main()
