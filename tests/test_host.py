"""CPU tests of the host-side mirror: Solver.Add batch assignment invariants (Solver.cs:L984-1140,L1182-1199) and buffer layouts."""
import numpy as np

from bepuphysics2_b200 import scenes
from tests import util

DYN_LIMIT = 1 << 30


def _batch_bodies(sim):
    per_batch = {}
    for tb in sim.type_batches():
        refs = tb.body_references
        per_batch.setdefault(tb.batch_index, []).append(refs[refs >= 0])
    return {b: np.concatenate(v) for b, v in per_batch.items()}


def test_no_dynamic_body_twice_in_a_synchronized_batch(libs):
    sim = util.make_sim(scenes.shape_pile(4000, seed=2))
    for b, refs in _batch_bodies(sim).items():
        dyn = refs[refs < DYN_LIMIT]
        assert np.unique(dyn).size == dyn.size, "batch %d repeats a dynamic body" % b


def test_greedy_first_fit_uses_the_lowest_batch(libs):
    """Every constraint in batch b > 0 must be blocked from each lower batch by one of its dynamic bodies."""
    sim = util.make_sim(scenes.box_stacks(3, 6))
    sets = {b: set(r[r < DYN_LIMIT].tolist()) for b, r in _batch_bodies(sim).items()}
    for tb in sim.type_batches():
        if tb.batch_index == 0:
            continue
        W = sim.bundle_width
        for c in range(tb.constraint_count):
            refs = tb.body_references[c // W, :, c % W]
            dyn = [int(r) for r in refs if 0 <= r < DYN_LIMIT]
            # adds are sequential, so being blocked by the *final* contents of a lower batch is necessary (not sufficient); check necessity
            for lower in range(tb.batch_index):
                assert any(d in sets[lower] for d in dyn)


def test_kinematics_do_not_block_and_are_flagged(libs):
    sim = util.make_sim(scenes.box_stacks(16, 2))
    # 16 bottom boxes all touch the single kinematic ground: they must share batch 0 with it flagged kinematic
    tb0 = [t for t in sim.type_batches() if t.batch_index == 0][0]
    refs_b = tb0.body_references[:, 1, :].ravel()
    refs_b = refs_b[refs_b >= 0]
    assert (refs_b >= DYN_LIMIT).sum() == 16
    assert set((refs_b[refs_b >= DYN_LIMIT] & (DYN_LIMIT - 1)).tolist()) == {0}
    assert sim.constrained_kinematics.tolist() == [0]
    assert sim.batch_count == 2


def test_fallback_batch_bundles_never_share_a_dynamic_body(libs):
    sim = util.make_sim(scenes.fallback_stress(400, hubs=2, seed=3), fallback_batch_threshold=6)
    assert sim.batch_count == 7
    fb = [t for t in sim.type_batches() if t.batch_index == 6]
    assert fb, "fallback batch must exist"
    for tb in fb:
        for k in range(tb.bundle_count):
            refs = tb.body_references[k].ravel()
            dyn = refs[(refs >= 0) & (refs < DYN_LIMIT)]
            assert np.unique(dyn).size == dyn.size


def test_trailing_lanes_are_empty_and_impulses_start_at_zero(libs):
    sim = util.make_sim(scenes.box_stacks(1, 5))
    for tb in sim.type_batches():
        flat = tb.body_references[:, 0, :].ravel()
        assert (flat[tb.constraint_count:] == -1).all()
        assert not tb.accumulated_impulses.any()
