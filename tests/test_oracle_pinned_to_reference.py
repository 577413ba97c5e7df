"""Pins the hand-written oracle (oracle/*.h) to the REFERENCE'S OWN C# TEXT.

The reference cannot be built here (no .NET). Its constraint functions, wide math and pose integration are straight-line code, so
oracle/ref_transpile/cs2cpp.py transpiles those sources mechanically (syntax only; every operator and call is the C# text's, in its order) into
oracle/_ref/libbepu_ref.so, compiled without FMA contraction like RyuJIT's Vector<float> code. Two layers of checks, all BIT FOR BIT:

  * committed known-answer vectors (tests/golden/reference_vectors.npz, generated from the transpiled reference by
    tests/golden/make_reference_vectors.py): every one of the 44 constraint types x {WarmStart, Solve, IncrementallyUpdateForSubstep}, the four
    PoseIntegration functions, and 38 chains of 1000 x (WarmStart; Solve) on the reference's constraint micro-benchmark inputs
    (DemoBenchmarks/*ConstraintBenchmarks*.cs) -- these run anywhere, with or without /root/reference;
  * live: when the reference tree (or a prebuilt oracle/_ref) is present, fresh random inputs through both libraries.

What stays outside the pin: the solver driver (substep loop, batch order, integration responsibilities, gather/scatter, the TypeProcessor bundle
loops), which is generic / unsafe C# the transpiler does not cover; tests/test_oracle.py holds those to closed-form answers.
TEST INFRASTRUCTURE: nothing in the product loads either library."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_transpile"))
FP = C.POINTER(C.c_float)
DT = 1.0 / 240.0


def _ptr(a):
    return a.ctypes.data_as(FP)


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def oracle(libs):
    lib = ob.load()
    lib.oracle_eval_lane.argtypes = [C.c_int32, C.c_int32, FP, C.c_float, FP, FP, FP, C.c_int32]
    lib.oracle_eval_integration.argtypes = [C.c_int32, FP, FP]
    return lib


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


def test_golden_vectors_cover_every_registered_type(oracle, golden):
    registered = sorted(t for t in range(64) if ob.type_info(t) is not None)
    in_file = sorted(int(k[5:7]) for k in golden.files if k.startswith("lane_") and k.endswith("_states"))
    assert registered == in_file and len(registered) == 44


def test_oracle_reproduces_the_reference_constraint_functions_bit_for_bit(oracle, golden):
    checked = 0
    for type_id in sorted(t for t in range(64) if ob.type_info(t) is not None):
        key = "lane_%02d_" % type_id
        states, vel, imp, pre = (golden[key + n] for n in ("states", "velocities", "impulses", "prestep"))
        bodies, prestep_rows, impulse_rows = ob.type_info(type_id)
        assert states.shape[1:] == (bodies, 14) and imp.shape[1] == impulse_rows and pre.shape[1] == prestep_rows
        for i in range(states.shape[0]):
            for stage in (0, 1, 2):
                v, a, p = vel[i].copy(), imp[i].copy(), pre[i].copy()
                st = np.ascontiguousarray(states[i])
                assert oracle.oracle_eval_lane(type_id, stage, _ptr(st), DT, _ptr(p), _ptr(a), _ptr(v), 1) == 0
                what = "type %d stage %d sample %d" % (type_id, stage, i)
                assert np.array_equal(_bits(v), _bits(golden[key + "out%d_velocities" % stage][i])), what + ": velocities"
                assert np.array_equal(_bits(a), _bits(golden[key + "out%d_impulses" % stage][i])), what + ": accumulated impulses"
                assert np.array_equal(_bits(p), _bits(golden[key + "out%d_prestep" % stage][i])), what + ": prestep"
                checked += 1
        # the vectors exercise the functions: WarmStart and Solve moved the velocities
        assert not np.array_equal(golden[key + "out1_velocities"], vel)
    assert checked == 44 * 3 * 6


def test_oracle_reproduces_the_reference_pose_integration_bit_for_bit(oracle, golden):
    """PoseIntegration.Integrate (custom Sin/Cos, normalisation, |w| <= 1e-15 fallback), RotateInverseInertia, IntegrateAngularVelocityConserveMomentum,
    ...WithGyroscopicTorque (BepuPhysics/PoseIntegrator.cs:L146-253)."""
    for op in range(4):
        ins, outs = golden["integration_%d_in" % op], golden["integration_%d_out" % op]
        for i in range(ins.shape[0]):
            got = np.zeros(outs.shape[1], dtype=np.float32)
            assert oracle.oracle_eval_integration(op, _ptr(np.ascontiguousarray(ins[i])), _ptr(got)) == 0
            assert np.array_equal(_bits(got), _bits(outs[i])), "integration function %d sample %d" % (op, i)


def test_oracle_reproduces_the_reference_benchmark_chains_bit_for_bit(oracle, golden):
    """DemoBenchmarks/{One,Two,Three,Four}BodyConstraintBenchmarks[Deep].cs: 1000 x (WarmStart; Solve) at dt = 1/60 from rest on the benchmark's own
    prestep data; a chain amplifies any difference in a single evaluation, and exercises warm starting with the solver's own accumulated impulses."""
    names = list(golden["bench_names"])
    assert len(names) >= 30
    for n, name in enumerate(names):
        k = "bench_%02d_" % n
        tid = int(golden[k + "type"])
        st, p = np.ascontiguousarray(golden[k + "states"]), golden[k + "prestep"].copy()
        v, a = np.zeros_like(golden[k + "out_velocities"]), np.zeros_like(golden[k + "out_impulses"])
        for _ in range(1000):
            oracle.oracle_eval_lane(tid, 0, _ptr(st), 1.0 / 60.0, _ptr(p), _ptr(a), _ptr(v), 1)
            oracle.oracle_eval_lane(tid, 1, _ptr(st), 1.0 / 60.0, _ptr(p), _ptr(a), _ptr(v), 1)
        assert np.isfinite(v).all(), name
        assert np.array_equal(_bits(v), _bits(golden[k + "out_velocities"])), name + ": velocities after 1000 iterations"
        assert np.array_equal(_bits(a), _bits(golden[k + "out_impulses"])), name + ": accumulated impulses after 1000 iterations"


@pytest.fixture(scope="module")
def reference_library():
    import build_ref

    path = build_ref.build()
    if path is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libbepu_ref.so is present")
    lib = C.CDLL(path)
    lib.ref_eval_lane.argtypes = [C.c_int32, C.c_int32, FP, C.c_float, FP, FP, FP, C.c_int32]
    lib.ref_eval_integration.argtypes = [C.c_int32, FP, FP]
    lib.ref_covered_types.argtypes = [C.POINTER(C.c_int32), C.c_int32]
    return lib


def test_transpiled_reference_covers_all_types_and_reproduces_the_committed_vectors(reference_library, golden):
    ids = (C.c_int32 * 64)()
    n = reference_library.ref_covered_types(ids, 64)
    assert sorted(ids[:n]) == sorted(t for t in range(64) if ob.type_info(t) is not None)
    key = "lane_07_"  # the committed file is what this library produces (regenerate with tests/golden/make_reference_vectors.py)
    v, a, p = golden[key + "velocities"][0].copy(), golden[key + "impulses"][0].copy(), golden[key + "prestep"][0].copy()
    assert reference_library.ref_eval_lane(7, 1, _ptr(np.ascontiguousarray(golden[key + "states"][0])), DT, _ptr(p), _ptr(a), _ptr(v), 1) == 0
    assert np.array_equal(_bits(v), _bits(golden[key + "out1_velocities"][0]))


def test_oracle_matches_the_transpiled_reference_on_fresh_inputs(oracle, reference_library):
    """Every type, every stage, 32 fresh samples per type (other seeds than the committed vectors), plus degenerate lanes: zero velocities and
    impulses, identical poses, a kinematic partner (zero inverse mass and inertia)."""
    from tests.test_device_source_on_host import _prestep_samples, _random_states

    samples = _prestep_samples()
    rng = np.random.default_rng(991)
    for type_id in sorted(samples):
        bodies, prestep_rows, impulse_rows = ob.type_info(type_id)
        for n, prestep in enumerate(samples[type_id]):
            states = _random_states(rng, bodies)
            vel = rng.normal(0, 1.5, (bodies, 6)).astype(np.float32)
            imp = np.abs(rng.normal(0, 0.2, impulse_rows)).astype(np.float32)
            if n % 8 == 5:
                vel[:] = 0
                imp[:] = 0
            if n % 8 == 6 and bodies > 1:
                states[1, 7:14] = 0  # kinematic partner
            if n % 8 == 7:
                states[:, 3:7] = (0, 0, 0, 1)
            for stage in (0, 1, 2):
                p1, a1, v1 = prestep.copy(), imp.copy(), vel.copy()
                p2, a2, v2 = prestep.copy(), imp.copy(), vel.copy()
                assert oracle.oracle_eval_lane(type_id, stage, _ptr(states), DT, _ptr(p1), _ptr(a1), _ptr(v1), 1) == 0
                assert reference_library.ref_eval_lane(type_id, stage, _ptr(states), DT, _ptr(p2), _ptr(a2), _ptr(v2), 1) == 0
                what = "type %d stage %d sample %d" % (type_id, stage, n)
                assert np.array_equal(_bits(v1), _bits(v2)), what + ": velocities"
                assert np.array_equal(_bits(a1), _bits(a2)), what + ": accumulated impulses"
                assert np.array_equal(_bits(p1), _bits(p2)), what + ": prestep"
