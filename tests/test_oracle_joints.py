"""CPU tests pinning the oracle's restatement of every joint / motor / servo / limit type beyond the ragdoll set on analytic known answers:
with gravity and damping off, each constraint must drive its own error measure (computed here independently, in float64, from the body
states) to its target. A wrong sign, swapped body or mis-indexed prestep row makes the measure diverge instead."""
import math

import numpy as np
import pytest

import bepuphysics2_b200 as bp
from bepuphysics2_b200 import scenes
from tests import util

DT = 1.0 / 60.0
FMAX = float(np.finfo(np.float32).max)
SPRING = [float(x) for x in scenes.spring(30, 1)]
SERVO = [FMAX, 0.0, FMAX]  # ServoSettings.Default: unlimited speed and force
MOTOR = [FMAX, 1e5]  # MotorSettingsWide.Damping = 1 / softness (MotorSettings.cs): a nearly rigid motor


def _q(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    return np.r_[a * math.sin(angle / 2), math.cos(angle / 2)].astype(np.float32)


def _rot(v, q):
    return scenes.qrot(np.asarray(v, dtype=np.float64)[None, :], np.asarray(q, dtype=np.float64)[None, :])[0]


def _still():
    d = bp.IntegratorDesc.default()
    d.gravity[1] = 0.0
    d.linear_damping = 0.0
    d.angular_damping = 0.0
    return d


def _bodies(count):
    """`count` dynamic bodies on a small triangle/tetrahedron, rotated and moving a little."""
    pos = np.array([[0, 0, 0], [1.2, 0.3, -0.2], [0.1, 1.1, 0.4], [-0.3, 0.2, 1.3]], dtype=np.float32)[:count]
    orient = np.stack([_q([1, 2, 3], 0.4), _q([-1, 0.5, 0.2], -0.7), _q([0, 1, 0], 0.3), _q([1, 0, 1], 1.0)])[:count]
    lin = np.array([[0.1, 0, -0.1], [-0.2, 0.1, 0], [0, 0.05, 0.1], [0.05, -0.1, 0]], dtype=np.float32)[:count]
    ang = np.array([[0.2, -0.1, 0.3], [0, 0.3, -0.2], [0.1, 0.1, 0], [0, 0, 0.2]], dtype=np.float32)[:count]
    inertia = np.tile(np.array([[2.0, 0, 2.5, 0, 0, 3.0]], dtype=np.float32), (count, 1))
    return scenes.make_bodies(pos, orientation=orient, linear=lin, angular=ang, inverse_mass=np.array([1.0, 0.7, 1.3, 0.9], dtype=np.float32)[:count], inverse_inertia=inertia)


def _run(type_id, prestep, body_count, frames=240, substeps=4, angular_damping=0.0):
    handles = np.arange(body_count, dtype=np.int32)[None, :]
    scene = {"bodies": _bodies(body_count), "constraints": [(type_id, handles, np.asarray([prestep], dtype=np.float32))]}
    integ = _still()
    integ.angular_damping = angular_damping
    sim = util.make_sim(scene, substeps=substeps, velocity_iterations=2, integrator=integ)
    for _ in range(frames):
        util.ob.solve(sim, DT)
    b = sim.bodies.astype(np.float64)
    assert np.isfinite(b).all()
    return [{"q": b[i, 0:4], "p": b[i, 4:7], "v": b[i, 8:11], "w": b[i, 12:15]} for i in range(body_count)]


def _anchor(body, local):
    return body["p"] + _rot(local, body["q"])


def _anchor_velocity(body, local):
    return body["v"] + np.cross(body["w"], _rot(local, body["q"]))


def test_weld_holds_relative_pose(libs):
    offset, target = [0.8, 0.2, -0.1], _q([0, 0, 1], 0.5)
    a, b = _run(31, np.r_[offset, target, SPRING], 2)
    assert np.linalg.norm(b["p"] - _anchor(a, offset)) < 2e-3
    want = scenes.qcat(target[None, :].astype(np.float64), a["q"][None, :])[0]  # LocalOrientation * orientationA
    assert abs(abs(np.dot(want, b["q"])) - 1) < 1e-5


def test_ball_socket_servo_joins_anchors(libs):
    oa, ob_ = [0.5, 0.1, 0], [-0.4, 0.2, 0.1]
    a, b = _run(53, np.r_[oa, ob_, SPRING, SERVO], 2)
    assert np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_)) < 2e-3


def test_ball_socket_motor_drives_anchor_velocity(libs):
    ob_, target = [0.3, 0.1, -0.2], [0.5, -0.3, 0.2]
    a, b = _run(52, np.r_[ob_, target, MOTOR], 2, frames=30)
    # offsetA = (pB - pA) + offsetB: both anchors are the same world point; B's velocity there minus A's equals the target in A's frame
    world = _anchor(b, ob_)
    va = a["v"] + np.cross(a["w"], world - a["p"])
    vb = b["v"] + np.cross(b["w"], world - b["p"])
    assert np.allclose(vb - va, _rot(target, a["q"]), atol=2e-2)


def test_distance_servo_reaches_target_distance(libs):
    oa, ob_ = [0.2, 0, 0.1], [0, -0.2, 0.1]
    a, b = _run(33, np.r_[oa, ob_, [1.7], SERVO, SPRING], 2)
    assert np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_)) == pytest.approx(1.7, abs=3e-3)


@pytest.mark.parametrize("lo,hi", [(2.0, 2.5), (0.2, 0.6)])
def test_distance_limit_keeps_distance_in_range(libs, lo, hi):
    oa, ob_ = [0.2, 0, 0.1], [0, -0.2, 0.1]
    a, b = _run(34, np.r_[oa, ob_, [lo, hi], SPRING], 2)
    d = np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_))
    assert lo - 5e-3 <= d <= hi + 5e-3


def test_center_distance_constraint_and_limit(libs):
    a, b = _run(35, np.r_[[2.0], SPRING], 2)
    assert np.linalg.norm(b["p"] - a["p"]) == pytest.approx(2.0, abs=3e-3)
    a, b = _run(55, np.r_[[2.0, 2.4], SPRING], 2)
    assert 2.0 - 5e-3 <= np.linalg.norm(b["p"] - a["p"]) <= 2.4 + 5e-3
    a, b = _run(55, np.r_[[0.3, 0.7], SPRING], 2)
    assert 0.3 - 5e-3 <= np.linalg.norm(b["p"] - a["p"]) <= 0.7 + 5e-3


def test_point_on_line_servo_puts_anchor_on_the_line(libs):
    oa, ob_, direction = [0.1, 0.2, 0], [0, 0.1, -0.1], np.array([1, 1, 0]) / math.sqrt(2)
    a, b = _run(37, np.r_[oa, ob_, direction, SERVO, SPRING], 2)
    rel = _anchor(b, ob_) - _anchor(a, oa)
    d = _rot(direction, a["q"])
    assert np.linalg.norm(rel - d * np.dot(rel, d)) < 3e-3


def test_linear_axis_servo_limit_and_motor(libs):
    oa, ob_, normal = [0.1, 0.2, 0], [0, 0.1, -0.1], np.array([0, 1, 0.0])
    a, b = _run(38, np.r_[oa, ob_, normal, [0.6], SERVO, SPRING], 2)
    assert np.dot(_anchor(b, ob_) - _anchor(a, oa), _rot(normal, a["q"])) == pytest.approx(0.6, abs=3e-3)
    a, b = _run(40, np.r_[oa, ob_, normal, [1.0, 1.5], SPRING], 2)
    assert 1.0 - 5e-3 <= np.dot(_anchor(b, ob_) - _anchor(a, oa), _rot(normal, a["q"])) <= 1.5 + 5e-3
    a, b = _run(39, np.r_[oa, ob_, normal, [0.8], MOTOR], 2, frames=30)
    n = _rot(normal, a["q"])
    anchor_b = _anchor(b, ob_)  # jacobians act at B's anchor: A's velocity is measured at the closest point on the plane to it
    closest = anchor_b - n * np.dot(anchor_b - _anchor(a, oa), n)
    va = a["v"] + np.cross(a["w"], closest - a["p"])
    assert np.dot(_anchor_velocity(b, ob_) - va, n) == pytest.approx(0.8, abs=2e-2)


def test_angular_hinge_aligns_axes_and_swivel_hinge_keeps_them_perpendicular(libs):
    ha, hb = np.array([0, 1, 0.0]), np.array([1, 0, 0.0])
    a, b = _run(23, np.r_[ha, hb, SPRING], 2)
    assert np.dot(_rot(ha, a["q"]), _rot(hb, b["q"])) > 1 - 1e-5
    a, b = _run(24, np.r_[ha, hb, SPRING], 2)
    assert abs(np.dot(_rot(ha, a["q"]), _rot(hb, b["q"]))) < 2e-3


def test_twist_and_axis_motors_reach_target_relative_spin(libs):
    axis = np.array([0, 0, 1.0])
    a, b = _run(28, np.r_[axis, axis, [1.5], MOTOR], 2, frames=20)
    j = _rot(axis, a["q"]) + _rot(axis, b["q"])
    j /= np.linalg.norm(j)
    assert np.dot(a["w"] - b["w"], j) == pytest.approx(1.5, abs=2e-2)
    a, b = _run(41, np.r_[axis, [1.5], MOTOR], 2, frames=20)
    assert np.dot(a["w"] - b["w"], _rot(axis, a["q"])) == pytest.approx(1.5, abs=2e-2)


def test_one_body_servos_and_motors(libs):
    target_q = _q([1, 1, 0], 0.9)
    (a,) = _run(42, np.r_[target_q, SPRING, SERVO], 1)
    assert abs(abs(np.dot(a["q"], target_q.astype(np.float64))) - 1) < 1e-5
    (a,) = _run(43, np.r_[[0.5, -1.0, 0.25], MOTOR], 1, frames=20)
    assert np.allclose(a["w"], [0.5, -1.0, 0.25], atol=1e-2)
    offset, target = [0.3, 0.1, -0.2], [1.0, 2.0, -0.5]
    (a,) = _run(44, np.r_[offset, target, SPRING, SERVO], 1)
    # a point servo leaves the rotation free: the spring energy ends up as spin about the grab point (no damping here), so the anchor orbits it slightly
    assert np.linalg.norm(_anchor(a, offset) - target) < 2e-2
    (a,) = _run(45, np.r_[offset, [0.4, 0.0, -0.3], MOTOR], 1, frames=20)
    assert np.allclose(_anchor_velocity(a, offset), [0.4, 0.0, -0.3], atol=1e-2)


def test_area_and_volume_constraints_reach_target(libs):
    a, b, c = _run(36, np.r_[[2.0], SPRING], 3)
    assert np.linalg.norm(np.cross(b["p"] - a["p"], c["p"] - a["p"])) == pytest.approx(2.0, rel=5e-3)
    a, b, c, d = _run(32, np.r_[[2.5], SPRING], 4)
    assert np.dot(np.cross(b["p"] - a["p"], c["p"] - a["p"]), d["p"] - a["p"]) == pytest.approx(2.5, rel=5e-3)


def test_angular_axis_gear_motor_reproduces_reference_apply_of_accumulated_impulse(libs):
    """AngularAxisGearMotor.cs:L112 applies the accumulated impulse (not the corrective one) in Solve. One frame, one substep, one iteration from rest
    impulse 0 must therefore equal the plain formula: csi = (wB.axis - wA.jA) * effectiveMass, velocities changed by exactly that impulse."""
    axis, scale = np.array([0, 0, 1.0]), 2.0
    scene = {"bodies": _bodies(2), "constraints": [(54, np.array([[0, 1]], dtype=np.int32), np.asarray([np.r_[axis, [scale], [FMAX, 0.5]]], dtype=np.float32))]}
    sim = util.make_sim(scene, substeps=1, velocity_iterations=1, integrator=_still())
    before = sim.bodies.astype(np.float64).copy()
    util.ob.solve(sim, DT)
    after = sim.bodies.astype(np.float64)
    qa = before[0, 0:4]
    ax = _rot(axis, qa)
    ja = ax * scale

    def world_inverse_inertia(q):
        r = np.stack([_rot([1, 0, 0], q), _rot([0, 1, 0], q), _rot([0, 0, 1], q)], axis=1)  # columns = rotated basis vectors
        return r @ np.diag([2.0, 2.5, 3.0]) @ r.T

    ia, ib = world_inverse_inertia(qa), world_inverse_inertia(before[1, 0:4])
    wa, wb = before[0, 12:15], before[1, 12:15]
    dtd = DT * 0.5
    soft = 1 / (dtd + 1)
    cfm = dtd * soft
    csi = (np.dot(wb, ax) - np.dot(wa, ja)) * cfm / (ja @ ia @ ja + ax @ ib @ ax)
    assert np.allclose(after[0, 12:15], wa + ia @ ja * csi, atol=1e-5)
    assert np.allclose(after[1, 12:15], wb - ib @ ax * csi, atol=1e-5)
    tb = sim.type_batches()[0]
    assert tb.accumulated_impulses[0, 0, 0] == pytest.approx(csi, rel=1e-4)


def test_joint_zoo_scalar_and_simd_are_bit_identical(libs):
    scene = scenes.joint_zoo(400, 24, seed=4)
    a = util.run_oracle(util.make_sim(scene, substeps=2, velocity_iterations=2), DT, frames=2, simd=False)
    b = util.run_oracle(util.make_sim(scene, substeps=2, velocity_iterations=2), DT, frames=2, simd=True)
    util.compare(a, b, exact=True)


# ---- the ragdoll joint set (types 22, 25-27, 29, 30, 46, 47): the same kind of independent known answers ---------------------------------------
def _qcat64(a, b):
    return scenes.qcat(np.asarray(a, dtype=np.float64)[None, :], np.asarray(b, dtype=np.float64)[None, :])[0].astype(np.float64)


def _twist_angle(a, b, basis_a, basis_b):
    """Signed twist of B's basis about the shared Z axis relative to A's, measured like TwistServoFunctions.ComputeCurrentAngle but in float64:
    rotate B's basis so its Z meets A's Z by the shortest arc, then the angle of B's X in A's XY plane."""
    qa, qb = _qcat64(basis_a, a["q"]), _qcat64(basis_b, b["q"])
    ax, ay, az = _rot([1, 0, 0], qa), _rot([0, 1, 0], qa), _rot([0, 0, 1], qa)
    bx, bz = _rot([1, 0, 0], qb), _rot([0, 0, 1], qb)
    axis = np.cross(bz, az)
    s, c = np.linalg.norm(axis), np.dot(bz, az)
    if s > 1e-12:
        k = axis / s
        ang = math.atan2(s, c)
        bx = bx * math.cos(ang) + np.cross(k, bx) * math.sin(ang) + k * np.dot(k, bx) * (1 - math.cos(ang))  # Rodrigues
    return math.atan2(np.dot(bx, ay), np.dot(bx, ax))


def test_ball_socket_hinge_and_swivel_hinge_geometry(libs):
    oa, ob_ = [0.5, 0.1, 0], [-0.4, 0.2, 0.1]
    a, b = _run(22, np.r_[oa, ob_, SPRING], 2)
    assert np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_)) < 2e-3
    ha, hb = np.array([0, 1, 0.0]), np.array([1, 0, 0.0])
    a, b = _run(47, np.r_[oa, ha, ob_, hb, SPRING], 2)
    assert np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_)) < 3e-3
    assert np.dot(_rot(ha, a["q"]), _rot(hb, b["q"])) > 1 - 1e-4
    a, b = _run(46, np.r_[oa, ha, ob_, hb, SPRING], 2)
    assert np.linalg.norm(_anchor(a, oa) - _anchor(b, ob_)) < 3e-3
    assert abs(np.dot(_rot(ha, a["q"]), _rot(hb, b["q"]))) < 3e-3


def test_swing_limit_keeps_axes_within_the_cone(libs):
    axis_a, axis_b = np.array([0, 1, 0.0]), np.array([1, 0, 0.0])  # the bodies of _bodies() start ~90 degrees apart in these axes
    min_dot = math.cos(0.5)
    # an inequality with undamped bodies just bounces off the cone; with some angular damping the pair settles inside it
    a, b = _run(25, np.r_[axis_a, axis_b, [min_dot], SPRING], 2, angular_damping=0.5)
    assert np.dot(_rot(axis_a, a["q"]), _rot(axis_b, b["q"])) > min_dot - 5e-3


def test_twist_servo_and_limit_control_the_twist_angle(libs):
    basis = scenes.basis_quaternion([0, 0, 1], [1, 0, 0])
    # Only the twist is constrained, so the two Z axes are free to swing apart (in a ragdoll a SwingLimit holds them together) and the twist
    # measure degenerates once they oppose each other: check while they are still roughly aligned.
    a, b = _run(26, np.r_[basis, basis, [0.7], SPRING, SERVO], 2, frames=12)
    assert np.dot(_rot([0, 0, 1], a["q"]), _rot([0, 0, 1], b["q"])) > 0.5
    assert _twist_angle(a, b, basis, basis) == pytest.approx(0.7, abs=5e-3)
    a, b = _run(27, np.r_[basis, basis, [-0.2, 0.1], SPRING], 2, frames=12)
    assert np.dot(_rot([0, 0, 1], a["q"]), _rot([0, 0, 1], b["q"])) > 0.5
    assert -0.2 - 5e-3 <= _twist_angle(a, b, basis, basis) <= 0.1 + 5e-3


def test_angular_servo_and_motor(libs):
    target = _q([0, 0, 1], 0.8)
    a, b = _run(29, np.r_[target, SPRING, SERVO], 2)
    want = _qcat64(target, a["q"])  # TargetRelativeRotationLocalA * orientationA
    assert abs(abs(np.dot(want, b["q"])) - 1) < 1e-5
    w = [0.4, -0.6, 0.2]
    a, b = _run(30, np.r_[w, MOTOR], 2, frames=20)
    assert np.allclose(a["w"] - b["w"], _rot(w, a["q"]), atol=2e-2)
