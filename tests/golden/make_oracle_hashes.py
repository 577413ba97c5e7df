"""Regenerates tests/golden/oracle_hashes.json: SHA-256 of the oracle's output (body states, accumulated impulses, prestep) on a few seeded
scenes. This does NOT pin the oracle to the reference (nothing can here: see DESIGN.md §5); it pins the oracle to ITSELF across rounds, so that an
accidental change to the checker -- which every GPU parity test leans on -- cannot go unnoticed. Run after an intended oracle change:

    python tests/golden/make_oracle_hashes.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import util
from bepuphysics2_b200 import scenes

CASES = {
    "box_stacks_4x8": (lambda: scenes.box_stacks(4, 8), dict(substeps=2, velocity_iterations=2)),
    "shape_pile_500": (lambda: scenes.shape_pile(500, seed=1, nonconvex_fraction=0.3), dict(substeps=4, velocity_iterations=2)),
    "ragdolls_10": (lambda: scenes.ragdolls(10, seed=2), dict(substeps=2, velocity_iterations=3)),
    "joint_zoo_200": (lambda: scenes.joint_zoo(200, 16, seed=3), dict(substeps=3, velocity_iterations=1)),
    "fallback_300": (lambda: scenes.fallback_stress(300, hubs=2, seed=4), dict(substeps=2, velocity_iterations=2, fallback_batch_threshold=6)),
}


def digest(name):
    make, kw = CASES[name]
    snap = util.run_oracle(util.make_sim(make(), **kw), 1.0 / 60.0, frames=3)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(snap["bodies"][:, util.MEANINGFUL]).tobytes())
    for tb in snap["type_batches"]:
        v = np.broadcast_to(tb["valid"][:, None, :], tb["impulses"].shape)
        h.update(np.ascontiguousarray(np.where(v, tb["impulses"], 0)).tobytes())
        v = np.broadcast_to(tb["valid"][:, None, :], tb["prestep"].shape)
        h.update(np.ascontiguousarray(np.where(v, tb["prestep"], 0)).tobytes())
    return h.hexdigest()


if __name__ == "__main__":
    out = {name: digest(name) for name in CASES}
    with open(os.path.join(ROOT, "tests", "golden", "oracle_hashes.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
