"""One constraint graph over several GPUs: the host partitioner (CPU: invariants, and a world-size-2 gloo run showing both ranks derive
the same global tables from the same scene) and, on the GPU, several ranks as contexts of one process on one device against the oracle, bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

from bepuphysics2_b200 import scenes, sharding
from tests import util

DT = 1.0 / 60.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_rows(sim):
    rows = {}
    for tb in sim.type_batches():
        nb = tb.body_references.shape[1]
        rows[(tb.batch_index, tb.type_batch_index)] = tb.body_references.transpose(0, 2, 1).reshape(-1, nb)[:tb.constraint_count]
    return rows


@pytest.mark.parametrize("rank_count", [2, 3, 8])
def test_partition_covers_every_constraint_once_and_masks_name_the_referencing_ranks(rank_count):
    sim = util.make_sim(scenes.shape_pile(3000, seed=3), substeps=2, velocity_iterations=2)
    shards, first_batch, constrained, masks = sharding.partition(sim, rank_count)
    rows = _global_rows(sim)
    seen = {k: np.zeros(v.shape[0], dtype=np.int32) for k, v in rows.items()}
    expect_masks = np.zeros(sim.body_count, dtype=np.uint8)
    expect_first = np.full(sim.body_count, sharding.INT32_MAX, dtype=np.int64)
    for r, shard in enumerate(shards):
        for tb in shard:
            key = (tb["batch_index"], tb["type_batch_index"])
            seen[key][tb["source"]] += 1
            # the compacted rows are the global rows of the source constraints, in the order the shard stores them
            nb = tb["refs"].shape[1]
            packed = tb["refs"].transpose(0, 2, 1).reshape(-1, nb)[:tb["count"]]
            assert np.array_equal(packed, rows[key][tb["source"]])
            assert (tb["refs"].transpose(0, 2, 1).reshape(-1, nb)[tb["count"]:] == -1).all()
            dyn = (packed >= 0) & ((packed & sharding.KINEMATIC_BIT) == 0)
            idx = (packed & sharding.INDEX_MASK)[dyn]
            expect_masks[idx] |= np.uint8(1 << r)
            np.minimum.at(expect_first, idx, tb["batch_index"])
            # boundary constraints (a dynamic body some other rank references too) come first
            boundary = (((masks[packed & sharding.INDEX_MASK] & ~np.uint8(1 << r)) != 0) & dyn).any(axis=1)
            assert not (np.diff(boundary.astype(np.int8)) > 0).any()
    for key, count in seen.items():
        assert (count == 1).all(), key
    assert np.array_equal(expect_masks, masks)
    assert np.array_equal(expect_first, first_batch.astype(np.int64))
    assert np.array_equal(constrained != 0, expect_first != sharding.INT32_MAX)  # no constrained kinematics in this scene
    # within a batch no dynamic body is written by two ranks: what one rank pushes never collides with another rank's write
    for batch in range(sim.batch_count):
        writers = np.zeros(sim.body_count, dtype=np.int32)
        for shard in shards:
            touched = np.zeros(sim.body_count, dtype=bool)
            for tb in shard:
                if tb["batch_index"] == batch:
                    touched[tb["idx"][tb["dynamic"]]] = True
            writers += touched
        assert writers.max() <= 1


def test_push_lists_name_exactly_the_shared_bodies_a_rank_writes():
    sim = util.make_sim(scenes.shape_pile(2000, seed=4))
    shards, first_batch, constrained, masks = sharding.partition(sim, 2)
    for rank in range(2):
        pushes = sharding.pushes_for_rank(shards[rank], rank, 2, first_batch, masks)
        for batch, (bodies, dst, owner) in pushes.items():
            assert (dst == 1 - rank).all()
            assert (masks[bodies] == 3).all()
            assert np.array_equal(owner != 0, first_batch[bodies] == batch)
            written = np.concatenate([tb["idx"][tb["dynamic"]] for tb in shards[rank] if tb["batch_index"] == batch])
            assert np.array_equal(np.sort(bodies), np.sort(written[masks[written] == 3]))


_GLOO_SCRIPT = r"""
import hashlib, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
from bepuphysics2_b200 import scenes, sharding
from tests import util
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sim = util.make_sim(scenes.shape_pile(1500, seed=9), substeps=2, velocity_iterations=1)
shards, first_batch, constrained, masks = sharding.partition(sim, world)
digest = hashlib.sha256(first_batch.tobytes() + constrained.tobytes() + masks.tobytes()).hexdigest()
mine = sum(tb["count"] for tb in shards[rank])
got = [None] * world
dist.all_gather_object(got, (digest, mine))
assert len({d for d, _ in got}) == 1, "ranks disagree on the global tables"
assert sum(m for _, m in got) == sim.constraint_count
if rank == 0:
    print("OK", sim.constraint_count, [m for _, m in got])
dist.destroy_process_group()
"""


def test_two_gloo_ranks_derive_the_same_global_tables(tmp_path):
    script = tmp_path / "ranks.py"
    script.write_text(_GLOO_SCRIPT % ROOT)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-3000:]


def _one_process_ranks(sim_scene, rank_count, frames, fused, libs, **kw):
    """`rank_count` contexts on device 0, wired to one another with bepucuda_shard_import_contexts; every rank's referenced bodies and its own
    impulses against the oracle run on the whole graph."""
    from oracle import binding as ob

    sim = util.make_sim(sim_scene, **kw)
    solvers = [sharding.ShardedSolver(sim, r, rank_count, 0, strict_fp=True, fused_pushes=fused) for r in range(rank_count)]
    try:
        for s in solvers:
            s.export_handles()
        for s in solvers:
            s.import_contexts(solvers)
        for s in solvers:
            s.describe()
        for s in solvers:
            s.synchronize()
        for _ in range(frames):
            ob.solve(sim, DT)
            for s in solvers:  # asynchronous launches: the ranks' graphs run side by side on the device and meet at their exchange points
                s.solve(DT)
        by_key = {(tb.batch_index, tb.type_batch_index): tb for tb in sim.type_batches()}
        for s in solvers:
            got = s.download()
            mine = s.referenced_bodies()
            assert mine.size > 0
            assert np.array_equal(sim.bodies[mine][:, util.MOTION].view(np.uint32), got[mine][:, util.MOTION].view(np.uint32)), "rank %d bodies" % s.rank
            for tb in s.shard:
                g = by_key[(tb["batch_index"], tb["type_batch_index"])]
                ref = g.accumulated_impulses.transpose(0, 2, 1).reshape(-1, g.accumulated_impulses.shape[1])[tb["source"]]
                have = tb["impulses"].transpose(0, 2, 1).reshape(-1, tb["impulses"].shape[1])[:tb["count"]]
                assert np.array_equal(ref.view(np.uint32), have.view(np.uint32)), "rank %d impulses of batch %d" % (s.rank, tb["batch_index"])
        shared = int(((solvers[0].masks & (solvers[0].masks - 1)) != 0).sum())
        assert shared > 0
    finally:
        for s in solvers:
            s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("rank_count,fused", [(2, True), (3, True), (2, False)])
def test_one_graph_over_several_ranks_bit_exact(libs, rank_count, fused):
    _one_process_ranks(scenes.shape_pile(2500, seed=6), rank_count, frames=3, fused=fused, libs=libs, substeps=4, velocity_iterations=2)


@pytest.mark.gpu
def test_one_graph_over_two_ranks_ragdolls_bit_exact(libs):
    _one_process_ranks(scenes.ragdolls(60, seed=2), 2, frames=2, fused=True, libs=libs, substeps=2, velocity_iterations=2)


@pytest.mark.gpu
def test_one_graph_over_two_ranks_joint_zoo_bit_exact(libs):
    """Three- and four-body constraints, kinematic bodies, every joint type: the pushes of body slots 2 and 3 and the kinematic stages."""
    _one_process_ranks(scenes.joint_zoo(1200, per_type=60, seed=8), 2, frames=2, fused=True, libs=libs, substeps=3, velocity_iterations=2)
