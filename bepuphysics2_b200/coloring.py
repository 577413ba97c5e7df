"""Host-side helpers around bepucuda_color_constraints: a scene's constraints as the flat list of encoded body references the
colouring consumes, and a Simulation built from device-computed batches."""
import numpy as np

KINEMATIC_BIT = 1 << 30


def scene_references(scene):
    """references[n, 4] (int32, -1 = unused slot, bit 30 = kinematic body) of a scene's constraints in add order (the order scenes.build uses)."""
    bodies = np.asarray(scene["bodies"], dtype=np.float32).reshape(-1, 32)
    kinematic = np.all(bodies[:, 16:23] == 0.0, axis=1)  # Bodies.cs:L326-331: inverse mass and inverse inertia all zero
    rows = []
    for _type_id, handles, _prestep in scene["constraints"]:
        h = np.asarray(handles, dtype=np.int32)
        h = h.reshape(h.shape[0], -1)
        enc = np.where(kinematic[h], h | KINEMATIC_BIT, h).astype(np.int32)
        rows.append(np.concatenate([enc, np.full((h.shape[0], 4 - h.shape[1]), -1, dtype=np.int32)], axis=1))
    return np.concatenate(rows, axis=0) if rows else np.zeros((0, 4), dtype=np.int32)


def build_with_batches(scene, simulation, batches):
    """scenes.build with every constraint placed in the batch `batches` names (in add order)."""
    simulation.add_bodies(scene["bodies"])
    begin = 0
    for type_id, handles, prestep in scene["constraints"]:
        n = len(handles)
        simulation.add_constraints_in_batches(type_id, handles, prestep, batches[begin:begin + n])
        begin += n
    return simulation


def check_batches(references, batches, fallback_batch_threshold=64):
    """True when no synchronized batch references a dynamic body twice (the invariant every stage kernel relies on)."""
    refs = np.asarray(references)
    dynamic = (refs >= 0) & ((refs & KINEMATIC_BIT) == 0)
    b = np.repeat(np.asarray(batches)[:, None], refs.shape[1], axis=1)
    keep = dynamic & (b < fallback_batch_threshold)
    pairs = b[keep].astype(np.int64) * (1 << 32) + refs[keep].astype(np.int64)
    return np.unique(pairs).size == pairs.size
