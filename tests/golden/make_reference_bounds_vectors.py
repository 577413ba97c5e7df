"""Generates tests/golden/reference_bounds_vectors.npz: known-answer vectors for the convex-primitive bounds of PredictBoundingBoxes,
produced by the REFERENCE'S OWN C# text: CapsuleWide / BoxWide / CylinderWide.GetBounds and BoundingBoxHelpers.GetAngularBoundsExpansion /
GetBoundsExpansion, transpiled mechanically (oracle/ref_transpile/cs2cpp.py) and glued as BoundingBoxBatcher.ExecuteConvexBatch glues them
(ref_convex_bounds in the generated harness). Run where /root/reference exists; the file is committed so the check runs anywhere.

    python tests/golden/make_reference_bounds_vectors.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_transpile"))
import build_ref  # noqa: E402

FP = C.POINTER(C.c_float)
SAMPLES = 48


def make_inputs(rng, n):
    """Random shapes (types 0, 1, 2, 4), poses and velocities: slow and fast, so that the pi/3 clamp of the angular expansion, both margin bounds and
    the expansion clamp are all hit."""
    types = rng.choice([0, 1, 2, 4], size=n).astype(np.int32)
    dims = rng.uniform(0.05, 3.0, size=(n, 3)).astype(np.float32)
    margins = np.stack([rng.choice([0.0, 0.01, 0.2], size=n), rng.choice([0.05, 1.0, 3.40282347e+38], size=n)], axis=1).astype(np.float32)
    allow = rng.integers(0, 2, size=n).astype(np.int32)
    q = rng.normal(size=(n, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    pos = rng.uniform(-50, 50, size=(n, 3)).astype(np.float32)
    lin = (rng.normal(size=(n, 3)) * rng.choice([0.0, 0.1, 5.0, 80.0], size=(n, 1))).astype(np.float32)
    ang = (rng.normal(size=(n, 3)) * rng.choice([0.0, 0.3, 10.0, 200.0], size=(n, 1))).astype(np.float32)
    return types, dims, margins, allow, q, pos, lin, ang


def evaluate(lib, types, dims, margins, allow, q, pos, lin, ang, dt):
    out = np.zeros((types.shape[0], 7), dtype=np.float32)
    for i in range(types.shape[0]):
        rc = lib.ref_convex_bounds(int(types[i]), dims[i].ctypes.data_as(FP), margins[i].ctypes.data_as(FP), int(allow[i]), q[i].ctypes.data_as(FP), pos[i].ctypes.data_as(FP),
                                   lin[i].ctypes.data_as(FP), ang[i].ctypes.data_as(FP), dt, out[i].ctypes.data_as(FP))
        assert rc == 0
    return out


def load_ref():
    lib = C.CDLL(build_ref.build())
    lib.ref_convex_bounds.argtypes = [C.c_int32, FP, FP, C.c_int32, FP, FP, FP, FP, C.c_float, FP]
    return lib


def main():
    lib = load_ref()
    rng = np.random.default_rng(20260923)
    out = {}
    for k, dt in enumerate((1.0 / 60.0, 1.0 / 240.0)):
        inputs = make_inputs(rng, SAMPLES)
        for name, a in zip(("types", "dims", "margins", "allow", "q", "pos", "lin", "ang"), inputs):
            out["set%d_%s" % (k, name)] = a
        out["set%d_dt" % k] = np.float32(dt)
        out["set%d_out" % k] = evaluate(lib, *inputs, dt)
    path = os.path.join(ROOT, "tests", "golden", "reference_bounds_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
