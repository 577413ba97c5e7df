"""Device-side batch colouring (bepucuda_color_constraints).

CPU: the oracle's sequential first fit (Solver.cs:L1182-1199 restated) reproduces the batches the host mirror's Solver.Add sequence builds, its
hashed / by-priority orders give valid batches, and feeding the current batches back as priorities is a fixed point (what BatchCompressor.cs:L233
works towards). GPU: the device result is IDENTICAL to the oracle's for every order (integer work: bit-exact), on small scenes, through the fallback
batch, and at the benchmark's size; a simulation built from device-computed batches solves bit-exactly against the oracle on the same batches."""
import numpy as np
import pytest

import bepuphysics2_b200 as bp
from bepuphysics2_b200 import coloring, scenes
from oracle import binding as ob
from tests import util

DT = 1.0 / 60.0


def _scenes():
    return [("pile", scenes.shape_pile(3000, seed=5), 64), ("ragdolls", scenes.ragdolls(40, seed=5), 64), ("zoo", scenes.joint_zoo(600, per_type=40, seed=3), 64),
            ("fallback", scenes.fallback_stress(600, hubs=3, seed=5), 8)]


def test_oracle_first_fit_is_the_host_mirrors_solver_add(libs):
    for name, scene, threshold in _scenes():
        sim = util.make_sim(scene, fallback_batch_threshold=threshold)
        refs, host_batches = sim.constraint_references()
        assert np.array_equal(refs, coloring.scene_references(scene)), name
        got, count = ob.first_fit_batches(refs, sim.body_count, threshold, order=0)
        assert np.array_equal(got, host_batches), name
        assert count == sim.batch_count, name


def test_oracle_orders_give_valid_batches_and_compression_is_a_fixed_point(libs):
    for name, scene, threshold in _scenes():
        refs = coloring.scene_references(scene)
        n_bodies = len(scene["bodies"])
        first, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=0)
        hashed, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=1)
        assert coloring.check_batches(refs, first, threshold) and coloring.check_batches(refs, hashed, threshold), name
        # a first-fit result is already compressed: recolouring in (batch, index) order moves nothing
        again, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=2, priorities=first)
        assert np.array_equal(again, first), name
        # a fragmented layout (every constraint pushed up by a random amount, still valid) only ever moves down, never up
        rng = np.random.default_rng(1)
        spread = (first.astype(np.int64) * 3 + rng.integers(0, 3, size=first.shape)).astype(np.int32)
        if threshold == 64 and spread.max() < 64:
            assert coloring.check_batches(refs, spread, threshold)
            packed, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=2, priorities=spread)
            assert (packed <= spread).all() and coloring.check_batches(refs, packed, threshold), name
            assert packed.max() <= first.max() * 3 and (np.bincount(packed) > 0).all()


def test_hash_matches_the_documented_formula(libs):
    cuda, _ = bp.load_libraries()
    for i in (0, 1, 2, 12345, 2 ** 31, 2 ** 32 - 1):
        h = (i * 0x9E3779B1) & 0xFFFFFFFF
        h ^= h >> 15
        h = (h * 0x85EBCA77) & 0xFFFFFFFF
        h ^= h >> 13
        h = (h * 0xC2B2AE3D) & 0xFFFFFFFF
        h ^= h >> 16
        assert cuda.bepucuda_color_hash(i) == h


@pytest.mark.gpu
def test_device_colouring_is_identical_to_sequential_first_fit(libs):
    ts = bp.CudaTimestepper(bp.Simulation())
    try:
        for name, scene, threshold in _scenes():
            refs = coloring.scene_references(scene)
            n_bodies = len(scene["bodies"])
            for order in (0, 1):
                want, want_count = ob.first_fit_batches(refs, n_bodies, threshold, order=order)
                got, count, rounds = ts.color_constraints(refs, n_bodies, threshold, order=order)
                assert np.array_equal(got, want), (name, order)
                assert count == want_count and rounds >= 1
            first, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=0)
            rng = np.random.default_rng(2)
            priorities = rng.integers(0, 1 << 20, size=first.shape).astype(np.uint32)
            want, _ = ob.first_fit_batches(refs, n_bodies, threshold, order=2, priorities=priorities)
            got, _, _ = ts.color_constraints(refs, n_bodies, threshold, order=2, priorities=priorities)
            assert np.array_equal(got, want), (name, "priorities")
            again, _, _ = ts.color_constraints(refs, n_bodies, threshold, order=2, priorities=first)
            assert np.array_equal(again, first), (name, "fixed point")
        # narrower reference rows, an empty list, bad arguments
        one = np.array([[0], [0], [1], [0 | coloring.KINEMATIC_BIT]], dtype=np.int32)
        got, count, _ = ts.color_constraints(one, 2, 64, order=0)
        assert got.tolist() == [0, 1, 0, 0] and count == 2
        got, count, _ = ts.color_constraints(np.zeros((0, 2), dtype=np.int32), 5, 64)
        assert got.size == 0 and count == 0
        with pytest.raises(bp.BepuCudaError):
            ts.color_constraints(np.array([[7, -1]], dtype=np.int32), 3, 64)
    finally:
        ts.close()


@pytest.mark.gpu
def test_device_colouring_at_benchmark_size(libs):
    """C2's 333 k manifolds: hashed order (few dependent rounds) and insertion order (the reference's own sequence) both match the oracle."""
    scene = scenes.shape_pile(100_000, seed=5)
    refs = coloring.scene_references(scene)
    ts = bp.CudaTimestepper(bp.Simulation())
    try:
        for order in (1, 0):
            want, want_count = ob.first_fit_batches(refs, 100_000, 64, order=order)
            got, count, rounds = ts.color_constraints(refs, 100_000, 64, order=order)
            assert np.array_equal(got, want) and count == want_count
            assert coloring.check_batches(refs, got)
            if order == 1:
                assert rounds <= 64  # a pseudo-random order keeps the dependency chains short
    finally:
        ts.close()


@pytest.mark.gpu
def test_solve_on_device_coloured_batches_is_bit_exact(libs):
    """The solver consumes device-computed batches like host-computed ones: GPU (strict) == oracle on a simulation built from them."""
    scene = scenes.shape_pile(3000, seed=9)
    refs = coloring.scene_references(scene)
    ts = bp.CudaTimestepper(bp.Simulation())
    try:
        batches, count, _ = ts.color_constraints(refs, len(scene["bodies"]), 64, order=1)
    finally:
        ts.close()
    sims = []
    for _ in range(2):
        sim = bp.Simulation(substeps=4, velocity_iterations=2)
        coloring.build_with_batches(scene, sim, batches)
        assert sim.batch_count == count
        sims.append(sim)
    ref = util.run_oracle(sims[0], DT, frames=2)
    got = util.run_gpu(sims[1], DT, frames=2, strict=True)
    util.compare(ref, got, exact=True)
