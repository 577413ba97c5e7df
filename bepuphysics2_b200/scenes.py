"""Seeded synthetic scenes for the BASELINE.json configs.

The reference's collision detection cannot run here, so contact manifolds are synthesised directly in the prestep layouts the
narrow phase would write (Constraints/Contact/ContactConvexTypes.cs prestep structs), with the material the reference benchmarks use
(DemoBenchmarks/ShapePileBenchmark.cs:L54-66: SpringSettings(30 Hz, zeta 1), MaximumRecoveryVelocity 2, friction 1).
Every generator is a pure function of its arguments (numpy PCG64 with a fixed seed, default 5 like the reference's `new Random(5)`).

A scene is a dict: {"bodies": float32[n, 32] BodyDynamics records, "constraints": [(type_id, int32[m, bodies], float32[m, prestep])...],
"description": str}. `build(scene, ...)` adds it to a host `Simulation` in order, which assigns batches like Solver.Add does.
"""
import math

import numpy as np

TWO_PI = np.float32(6.283185307179586477)


def make_bodies(position, orientation=None, linear=None, angular=None, inverse_mass=None, inverse_inertia=None):
    """Packs BodyDynamics records (BepuPhysics/BodyProperties.cs:L11-46,L318-338): floats 0-3 orientation xyzw, 4-6 position, 8-10 linear,
    12-14 angular, 16-21 local inverse inertia (XX,YX,YY,ZX,ZY,ZZ), 22 inverse mass, 24-30 world inverse inertia + inverse mass (left zero)."""
    position = np.asarray(position, dtype=np.float32).reshape(-1, 3)
    n = position.shape[0]
    b = np.zeros((n, 32), dtype=np.float32)
    b[:, 3] = 1.0
    if orientation is not None:
        b[:, 0:4] = np.asarray(orientation, dtype=np.float32).reshape(n, 4)
    b[:, 4:7] = position
    if linear is not None:
        b[:, 8:11] = np.asarray(linear, dtype=np.float32).reshape(n, 3)
    if angular is not None:
        b[:, 12:15] = np.asarray(angular, dtype=np.float32).reshape(n, 3)
    if inverse_inertia is not None:
        b[:, 16:22] = np.asarray(inverse_inertia, dtype=np.float32).reshape(n, 6)
    if inverse_mass is not None:
        b[:, 22] = np.asarray(inverse_mass, dtype=np.float32).reshape(n)
    return b


def spring(frequency_hz, damping_ratio):
    """SpringSettings(frequency, dampingRatio) -> (AngularFrequency, TwiceDampingRatio) (Constraints/SpringSettings.cs:L75-80)."""
    return np.float32(frequency_hz) * TWO_PI, np.float32(damping_ratio) * np.float32(2)


def convex_prestep(offsets_a, depths, normal, offset_b=None, friction=1.0, spring_settings=None, max_recovery=2.0):
    """Rows of Contact{N}[OneBody]PrestepData: [OffsetA xyz, Depth] x N, (OffsetB xyz), Normal xyz, FrictionCoefficient, AngularFrequency,
    TwiceDampingRatio, MaximumRecoveryVelocity. offsets_a: [m, N, 3], depths: [m, N], normal: [m, 3], offset_b: [m, 3] or None (one body)."""
    offsets_a = np.asarray(offsets_a, dtype=np.float32)
    m, N, _ = offsets_a.shape
    af, tdr = spring_settings if spring_settings is not None else spring(30, 1)
    cols = []
    for i in range(N):
        cols.append(offsets_a[:, i, :])
        cols.append(np.asarray(depths, dtype=np.float32)[:, i : i + 1])
    if offset_b is not None:
        cols.append(np.asarray(offset_b, dtype=np.float32).reshape(m, 3))
    cols.append(np.asarray(normal, dtype=np.float32).reshape(m, 3))
    mat = np.empty((m, 4), dtype=np.float32)
    mat[:, 0], mat[:, 1], mat[:, 2], mat[:, 3] = friction, af, tdr, max_recovery
    cols.append(mat)
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float32)


def nonconvex_prestep(offsets, depths, normals, offset_b=None, friction=1.0, spring_settings=None, max_recovery=2.0):
    """Rows of Contact{N}Nonconvex[OneBody]PrestepData: material (4), (OffsetB xyz), [Offset xyz, Depth, Normal xyz] x N."""
    offsets = np.asarray(offsets, dtype=np.float32)
    m, N, _ = offsets.shape
    af, tdr = spring_settings if spring_settings is not None else spring(30, 1)
    mat = np.empty((m, 4), dtype=np.float32)
    mat[:, 0], mat[:, 1], mat[:, 2], mat[:, 3] = friction, af, tdr, max_recovery
    cols = [mat]
    if offset_b is not None:
        cols.append(np.asarray(offset_b, dtype=np.float32).reshape(m, 3))
    for i in range(N):
        cols.append(offsets[:, i, :])
        cols.append(np.asarray(depths, dtype=np.float32)[:, i : i + 1])
        cols.append(np.asarray(normals, dtype=np.float32)[:, i, :])
    return np.ascontiguousarray(np.concatenate(cols, axis=1), dtype=np.float32)


CONVEX_ONE_BODY = {1: 0, 2: 1, 3: 2, 4: 3}
CONVEX_TWO_BODY = {1: 4, 2: 5, 3: 6, 4: 7}
NONCONVEX_ONE_BODY = {2: 8, 3: 9, 4: 10}
NONCONVEX_TWO_BODY = {2: 15, 3: 16, 4: 17}


def box_stacks(columns=16, height=16, penetration=0.01):
    """Config 1 (plumbing): `columns` stacks of `height` unit boxes (mass 1, inverse inertia diag 6) on one kinematic ground body.
    One Contact4 (type 7) per box: 4 corner contacts against the body below, normal (0,1,0), depth `penetration`."""
    n = columns * height
    pos = np.zeros((n + 1, 3), dtype=np.float32)
    pos[0] = (0, -0.5, 0)  # ground slab, top face at y = 0
    spacing = 1.0 - penetration
    idx = 1
    handles_a, handles_b = [], []
    for c in range(columns):
        for i in range(height):
            pos[idx] = (3.0 * c, 0.5 - penetration + i * spacing, 0.0)
            handles_a.append(idx)
            handles_b.append(idx - 1 if i > 0 else 0)
            idx += 1
    inv_mass = np.ones(n + 1, dtype=np.float32)
    inv_inertia = np.zeros((n + 1, 6), dtype=np.float32)
    inv_inertia[:, 0] = inv_inertia[:, 2] = inv_inertia[:, 5] = 6.0
    inv_mass[0] = 0
    inv_inertia[0] = 0
    bodies = make_bodies(pos, inverse_mass=inv_mass, inverse_inertia=inv_inertia)
    a = np.asarray(handles_a, dtype=np.int32)
    b = np.asarray(handles_b, dtype=np.int32)
    corners = np.array([[-0.5, -0.5, -0.5], [0.5, -0.5, -0.5], [-0.5, -0.5, 0.5], [0.5, -0.5, 0.5]], dtype=np.float32)
    offsets = np.broadcast_to(corners, (n, 4, 3)).copy()
    depths = np.full((n, 4), penetration, dtype=np.float32)
    normal = np.broadcast_to(np.array([0, 1, 0], dtype=np.float32), (n, 3))
    offset_b = pos[b] - pos[a]
    pre = convex_prestep(offsets, depths, normal, offset_b)
    return {"bodies": bodies, "constraints": [(7, np.stack([a, b], axis=1), pre)], "description": "%d columns x %d unit boxes, Contact4 only" % (columns, height)}


# ShapePileBenchmark shapes (DemoBenchmarks/ShapePileBenchmark.cs:L109-164), mass 1: local inverse inertia diagonals.
def _shape_inverse_inertias():
    def inv(ixx, iyy, izz):
        return np.array([1.0 / ixx, 0, 1.0 / iyy, 0, 0, 1.0 / izz], dtype=np.float32)

    r = 1.5
    sphere = inv(0.4 * r * r, 0.4 * r * r, 0.4 * r * r)
    # capsule radius 1, length 1 (axis y): approximate as cylinder + two hemispheres by volume-weighted mass split
    cr, cl = 1.0, 1.0
    vc, vs = math.pi * cr * cr * cl, 4.0 / 3.0 * math.pi * cr ** 3
    mc, ms = vc / (vc + vs), vs / (vc + vs)
    iyy = mc * 0.5 * cr * cr + ms * 0.4 * cr * cr
    ixx = mc * (cl * cl / 12.0 + cr * cr / 4.0) + ms * (0.4 * cr * cr + cl * cl / 4.0 + 3.0 * cl * cr / 8.0)
    capsule = inv(ixx, iyy, ixx)
    w, h, d = 1.0, 3.0, 2.0
    box = inv((h * h + d * d) / 12.0, (w * w + d * d) / 12.0, (w * w + h * h) / 12.0)
    cyr, cyl = 1.5, 0.3
    cylinder = inv(cyl * cyl / 12.0 + cyr * cyr / 4.0, 0.5 * cyr * cyr, cyl * cyl / 12.0 + cyr * cyr / 4.0)
    hull = inv(0.39, 0.39, 0.39)  # dodecahedron-ish hull, a little lighter in rotation than the unit sphere
    return np.stack([sphere, capsule, box, cylinder, hull])


_NEIGHBOR_OFFSETS = np.array(
    [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, -1, 0), (1, 0, 1), (1, 0, -1), (0, 1, 1), (0, 1, -1), (1, 1, 1), (1, 1, -1), (1, -1, 1), (1, -1, -1)],
    dtype=np.int64,
)


def _random_unit_quaternions(rng, n):
    q = rng.standard_normal((n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    return q.astype(np.float32)


def _tangent_frame(normal):
    ref = np.where(np.abs(normal[:, 0:1]) < 0.7, np.array([[1, 0, 0]], dtype=np.float32), np.array([[0, 1, 0]], dtype=np.float32))
    t1 = np.cross(normal, ref)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(normal, t1)
    return t1.astype(np.float32), t2.astype(np.float32)


def shape_pile(body_count=100_000, manifolds_per_body=3.5, one_body_fraction=0.10, seed=5, dims=None, nonconvex_fraction=0.0):
    """Config 2/4: a settled-pile contact graph on a jittered 3-D lattice. Bodies are the five ShapePileBenchmark shapes (mass 1) at random
    orientations; each body touches lattice neighbours (13 forward directions) with the probability that yields `manifolds_per_body`
    constraints per body; manifold contact counts are drawn {1: 15%, 2: 25%, 3: 15%, 4: 45%}; `one_body_fraction` of the manifolds are against
    the static world (one-body types 0-3). Optionally a fraction of the multi-contact manifolds use the nonconvex types (8-10, 15-17)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if dims is None:
        ny = max(4, int(round((body_count / 100.0) ** (1.0 / 3.0) * 2.15)))  # 100k -> 100 x 10 x 100
        nx = int(math.ceil(math.sqrt(body_count / ny)))
        nz = int(math.ceil(body_count / (nx * ny)))
    else:
        nx, ny, nz = dims
    n = body_count
    ids = np.arange(n, dtype=np.int64)
    gx, gy, gz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    spacing = 2.0
    pos = np.stack([gx, gy, gz], axis=1).astype(np.float32) * np.float32(spacing)
    pos += rng.uniform(-0.25, 0.25, size=(n, 3)).astype(np.float32)
    shape = rng.integers(0, 5, size=n)
    inv_inertia = _shape_inverse_inertias()[shape]
    bodies = make_bodies(
        pos,
        orientation=_random_unit_quaternions(rng, n),
        linear=rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32),
        angular=rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32),
        inverse_mass=np.ones(n, dtype=np.float32),
        inverse_inertia=inv_inertia,
    )
    two_body_target = manifolds_per_body * (1.0 - one_body_fraction)
    p_pair = min(1.0, two_body_target / len(_NEIGHBOR_OFFSETS))
    pairs_a, pairs_b = [], []
    for off in _NEIGHBOR_OFFSETS:
        ox, oy, oz = gx + off[0], gy + off[1], gz + off[2]
        valid = (ox >= 0) & (ox < nx) & (oy >= 0) & (oy < ny) & (oz >= 0) & (oz < nz)
        other = ox + oy * nx + oz * nx * ny
        valid &= other < n
        valid &= rng.random(n) < p_pair
        pairs_a.append(ids[valid])
        pairs_b.append(other[valid])
    a = np.concatenate(pairs_a)
    b = np.concatenate(pairs_b)
    order = np.argsort(a, kind="stable")  # narrow-phase-like order: by first body
    a, b = a[order].astype(np.int32), b[order].astype(np.int32)
    m2 = a.shape[0]
    m1 = int(round(n * manifolds_per_body * one_body_fraction))
    one_a = np.sort(rng.choice(n, size=m1, replace=m1 > n)).astype(np.int32)

    def manifold_geometry(count, pa, pb):
        counts = rng.choice(np.array([1, 2, 3, 4]), size=count, p=[0.15, 0.25, 0.15, 0.45])
        if pb is not None:
            d = pa - pb
            normal = d / np.linalg.norm(d, axis=1, keepdims=True)
            mid = 0.5 * (pa + pb)
        else:
            normal = np.tile(np.array([[0, 1, 0]], dtype=np.float32), (count, 1))
            mid = pa - normal * np.float32(0.9)
        normal = normal + rng.normal(0, 0.05, size=(count, 3)).astype(np.float32)
        normal = (normal / np.linalg.norm(normal, axis=1, keepdims=True)).astype(np.float32)
        t1, t2 = _tangent_frame(normal)
        ang = rng.uniform(0, 2 * math.pi, size=(count, 1)).astype(np.float32) + np.arange(4, dtype=np.float32)[None, :] * np.float32(math.pi / 2)
        rad = rng.uniform(0.15, 0.45, size=(count, 4)).astype(np.float32)
        pts = mid[:, None, :] + (np.cos(ang) * rad)[:, :, None] * t1[:, None, :] + (np.sin(ang) * rad)[:, :, None] * t2[:, None, :]
        offsets = (pts - pa[:, None, :]).astype(np.float32)
        depths = rng.uniform(-0.02, 0.05, size=(count, 4)).astype(np.float32)
        return counts, normal, offsets, depths

    constraints = []
    pa, pb = pos[a], pos[b]
    counts, normal, offsets, depths = manifold_geometry(m2, pa, pb)
    noncv = rng.random(m2) < nonconvex_fraction
    for N in (1, 2, 3, 4):
        sel = (counts == N) & ~(noncv & (N > 1))
        if sel.any():
            pre = convex_prestep(offsets[sel][:, :N], depths[sel][:, :N], normal[sel], (pb - pa)[sel])
            constraints.append((CONVEX_TWO_BODY[N], np.stack([a[sel], b[sel]], axis=1), pre))
        sel = (counts == N) & noncv & (N > 1)
        if N > 1 and sel.any():
            k = int(sel.sum())
            normals = normal[sel][:, None, :] + rng.normal(0, 0.03, size=(k, N, 3)).astype(np.float32)
            normals = (normals / np.linalg.norm(normals, axis=2, keepdims=True)).astype(np.float32)
            pre = nonconvex_prestep(offsets[sel][:, :N], depths[sel][:, :N], normals, (pb - pa)[sel])
            constraints.append((NONCONVEX_TWO_BODY[N], np.stack([a[sel], b[sel]], axis=1), pre))
    if m1 > 0:
        pa1 = pos[one_a]
        counts, normal, offsets, depths = manifold_geometry(m1, pa1, None)
        noncv = rng.random(m1) < nonconvex_fraction
        for N in (1, 2, 3, 4):
            sel = (counts == N) & ~(noncv & (N > 1))
            if sel.any():
                pre = convex_prestep(offsets[sel][:, :N], depths[sel][:, :N], normal[sel], None)
                constraints.append((CONVEX_ONE_BODY[N], one_a[sel].reshape(-1, 1), pre))
            sel = (counts == N) & noncv & (N > 1)
            if N > 1 and sel.any():
                k = int(sel.sum())
                normals = normal[sel][:, None, :] + rng.normal(0, 0.03, size=(k, N, 3)).astype(np.float32)
                normals = (normals / np.linalg.norm(normals, axis=2, keepdims=True)).astype(np.float32)
                pre = nonconvex_prestep(offsets[sel][:, :N], depths[sel][:, :N], normals, None)
                constraints.append((NONCONVEX_ONE_BODY[N], one_a[sel].reshape(-1, 1), pre))
    total = sum(c[1].shape[0] for c in constraints)
    return {
        "bodies": bodies,
        "constraints": constraints,
        "description": "shape pile: %d bodies on a %dx%dx%d jittered lattice, %d contact manifolds (types 0-7%s)" % (n, nx, ny, nz, total, ", 8-10, 15-17" if nonconvex_fraction > 0 else ""),
    }


def fallback_stress(body_count=50_000, hubs=50, seed=5, neighbour_manifolds_per_body=1.0):
    """Config 5: every body touches one of a few dynamic hub bodies, so hubs exceed FallbackBatchThreshold constraints and most constraints
    land in the sequential fallback batch; plus ordinary neighbour contacts."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = body_count
    pos = rng.uniform(-50, 50, size=(n, 3)).astype(np.float32)
    inv_inertia = _shape_inverse_inertias()[rng.integers(0, 5, size=n)]
    inv_mass = np.ones(n, dtype=np.float32)
    inv_mass[:hubs] = 0.01  # heavy hubs
    inv_inertia[:hubs] *= 0.01
    bodies = make_bodies(pos, orientation=_random_unit_quaternions(rng, n), linear=rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32),
                         angular=rng.uniform(-0.3, 0.3, size=(n, 3)).astype(np.float32), inverse_mass=inv_mass, inverse_inertia=inv_inertia)
    others = np.arange(hubs, n, dtype=np.int32)
    hub_of = (rng.integers(0, hubs, size=others.shape[0])).astype(np.int32)
    constraints = []

    def add_pairs(a, b, label_rng):
        m = a.shape[0]
        counts = label_rng.choice(np.array([1, 2, 3, 4]), size=m, p=[0.15, 0.25, 0.15, 0.45])
        d = pos[a] - pos[b]
        normal = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-3)).astype(np.float32)
        t1, t2 = _tangent_frame(normal)
        ang = label_rng.uniform(0, 2 * math.pi, size=(m, 1)).astype(np.float32) + np.arange(4, dtype=np.float32)[None, :] * np.float32(math.pi / 2)
        rad = label_rng.uniform(0.15, 0.45, size=(m, 4)).astype(np.float32)
        offs = (-normal * np.float32(0.9))[:, None, :] + (np.cos(ang) * rad)[:, :, None] * t1[:, None, :] + (np.sin(ang) * rad)[:, :, None] * t2[:, None, :]
        depths = label_rng.uniform(-0.02, 0.05, size=(m, 4)).astype(np.float32)
        offset_b = (-normal * np.float32(1.8)).astype(np.float32)
        for N in (1, 2, 3, 4):
            sel = counts == N
            if sel.any():
                constraints.append((CONVEX_TWO_BODY[N], np.stack([a[sel], b[sel]], axis=1), convex_prestep(offs[sel][:, :N].astype(np.float32), depths[sel][:, :N], normal[sel], offset_b[sel])))

    add_pairs(others, hub_of, rng)
    k = int(round(n * neighbour_manifolds_per_body))
    na = rng.integers(hubs, n, size=k).astype(np.int32)
    nb_ = rng.integers(hubs, n, size=k).astype(np.int32)
    keep = na != nb_
    add_pairs(na[keep], nb_[keep], rng)
    total = sum(c[1].shape[0] for c in constraints)
    return {"bodies": bodies, "constraints": constraints, "description": "fallback stress: %d bodies, %d hubs, %d manifolds" % (n, hubs, total)}


def build(scene, simulation):
    """Adds a scene to a host Simulation (Bodies.Add, then Solver.Add per constraint in list order)."""
    simulation.add_bodies(scene["bodies"])
    for type_id, handles, prestep in scene["constraints"]:
        simulation.add_constraints(type_id, handles, prestep)
    return simulation
