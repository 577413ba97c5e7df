"""Device-side contact constraint update, first slice: accumulated impulses stay resident on the device and are redistributed there
from the old to the new contact feature ids (NarrowPhase.RedistributeImpulses, CollisionDetection/NarrowPhaseConstraintUpdate.cs:L81-135), and
only the motion half of BodyDynamics crosses the bus. CPU part: the oracle's restatement of RedistributeImpulses against hand-worked cases.
GPU part: several frames of (narrow-phase-like update -> solve) through the resident path against the host path of the oracle, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from bepuphysics2_b200 import scenes
from oracle import binding as ob
from tests import util

DT = 1.0 / 60.0


def _redistribute(old_ids, old_imp, new_ids):
    lib = ob.load()
    ip, fp = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.oracle_redistribute_impulses.argtypes = [C.c_int32, ip, fp, C.c_int32, ip, fp]
    oi = np.array(old_ids, dtype=np.int32)
    om = np.array(old_imp, dtype=np.float32)
    ni = np.array(new_ids, dtype=np.int32)
    out = np.zeros(len(new_ids), dtype=np.float32)
    lib.oracle_redistribute_impulses(len(old_ids), oi.ctypes.data_as(ip), om.ctypes.data_as(fp), len(new_ids), ni.ctypes.data_as(ip), out.ctypes.data_as(fp))
    return out


def test_redistribute_impulses_known_answers(libs):
    """NarrowPhaseConstraintUpdate.cs:L81-135 on cases worked by hand."""
    # all matched, permuted: every new contact takes the impulse of the old contact with its id
    assert np.array_equal(_redistribute([10, 11, 12, 13], [1, 2, 3, 4], [13, 10, 12, 11]), np.array([4, 1, 3, 2], dtype=np.float32))
    # one new contact unmatched: it receives what the vanished old contact had
    assert np.array_equal(_redistribute([10, 11, 12, 13], [1, 2, 3, 4], [10, 99, 12, 13]), np.array([1, 2, 3, 4], dtype=np.float32))
    # two unmatched share the remainder (2 + 4) / 2 equally
    assert np.array_equal(_redistribute([10, 11, 12, 13], [1, 2, 3, 4], [10, 98, 12, 99]), np.array([1, 3, 3, 3], dtype=np.float32))
    # nothing matches: the total is spread evenly
    assert np.array_equal(_redistribute([1, 2], [3, 5], [7, 8]), np.array([4, 4], dtype=np.float32))
    # fewer new contacts than old, one match: matched keeps its own, nothing unmatched -> the rest is dropped
    assert np.array_equal(_redistribute([1, 2, 3], [3, 5, 7], [2]), np.array([5], dtype=np.float32))
    # duplicate ids: the inner loop always stops at the FIRST old contact with the id (only its impulse is zeroed, L96-101), so the second new contact matches it again and gets 0
    assert np.array_equal(_redistribute([5, 5], [2, 6], [5, 5]), np.array([2, 0], dtype=np.float32))


def _contact_counts(type_id):
    return (type_id & 3) + 1 if type_id <= 7 else (type_id - 6 if type_id <= 10 else type_id - 13)


def _initial_features(sim, rng):
    return {(tb.batch_index, tb.type_batch_index): rng.integers(0, 1 << 20, size=(tb.constraint_count, _contact_counts(tb.type_id)), dtype=np.int32)
            for tb in sim.type_batches() if tb.type_id <= 17}


def _narrow_phase_like_update(sim, features, rng):
    """What a frame of collision detection does to an unchanged topology: new depths, and some contacts replaced by new features."""
    new = {}
    for tb in sim.type_batches():
        key = (tb.batch_index, tb.type_batch_index)
        if key not in features:
            continue
        ids = features[key].copy()
        n = ids.shape[1]
        change = rng.random(ids.shape) < 0.2
        ids[change] = rng.integers(1 << 20, 1 << 21, size=int(change.sum()), dtype=np.int32)
        swap = rng.random(ids.shape[0]) < 0.3  # the narrow phase often reports the same features in another order
        if n > 1:
            ids[swap] = ids[swap][:, ::-1]
        new[key] = ids
        # depth rows: convex prestep rows 4 i + 3; nonconvex: after MaterialProperties (4 rows) [+ OffsetB (3 rows) for two bodies], contact i = 7 rows
        # (Offset xyz, Depth, Normal xyz) with the depth at + 3 (tests/golden/type_layouts.json)
        base = 4 if tb.type_id <= 10 else 7
        rows = [4 * i + 3 for i in range(n)] if tb.type_id <= 7 else [base + 7 * i + 3 for i in range(n)]
        for r in rows:  # the narrow phase writes fresh depths every frame (whatever the solver's incremental updates left in the row)
            tb.prestep[:, r, :] = rng.uniform(-0.02, 0.05, size=tb.prestep[:, r, :].shape).astype(np.float32)
    return new


@pytest.mark.gpu
@pytest.mark.parametrize("nonconvex", [0.0, 0.5])
def test_resident_impulses_with_device_side_redistribution_bit_exact(libs, nonconvex):
    import bepuphysics2_b200 as bp

    scene = scenes.shape_pile(3000, seed=31, nonconvex_fraction=nonconvex)
    kw = dict(substeps=3, velocity_iterations=2)
    host, dev = util.make_sim(scene, **kw), util.make_sim(scene, **kw)
    rng_h, rng_d = np.random.default_rng(77), np.random.default_rng(77)
    feat_h, feat_d = _initial_features(host, rng_h), _initial_features(dev, rng_d)
    ts = bp.CudaTimestepper(dev, strict_fp=True)
    ts.register_host_buffers()
    ts.describe()
    ts.set_contact_features(feat_d)
    for frame in range(4):
        if frame > 0:
            # host path: the narrow phase gathers the old impulses, redistributes, scatters, writes the new description (prestep)
            new_h = _narrow_phase_like_update(host, feat_h, rng_h)
            for tb in host.type_batches():
                key = (tb.batch_index, tb.type_batch_index)
                if key in new_h:
                    ob.update_contact_impulses(tb, feat_h[key], new_h[key])
            feat_h = new_h
            # device path: prestep + new ids go up, the impulses never leave the device; bodies: motion half only
            new_d = _narrow_phase_like_update(dev, feat_d, rng_d)
            ts.upload_body_motion()
            ts.update_contacts(new_d)
            feat_d = new_d
        ob.solve(host, DT)
        ts.solve_device_only(DT)
        ts.download_body_motion()
        t = ts.timings()
        assert t.d2h_bytes == dev.body_count * 64
        for cols in (np.r_[0:7], np.r_[8:11], np.r_[12:15]):
            assert np.array_equal(host.bodies[:, cols].view(np.uint32), dev.bodies[:, cols].view(np.uint32)), "frame %d" % frame
    ts.download_bodies()  # the motion-only download leaves the host's world-inertia half untouched: fetch it once for the full comparison
    ts.download_impulses()
    ts.download_prestep()
    ts.close()
    util.compare(util.snapshot(host), util.snapshot(dev), exact=True)
    moved = sum(int((feat_d[k] >= (1 << 20)).sum()) for k in feat_d)
    assert moved > 1000  # the test really replaced contacts
