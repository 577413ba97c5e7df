// ORACLE — test infrastructure only (see bepu_math.h header). Joint / motor / servo / limit constraint functions, each restating the
// reference file:line cited next to it. Pinned bit for bit to the reference's C# text through oracle/ref_transpile (tests/test_oracle_pinned_to_reference.py).
#pragma once
#include "bepu_contacts.h"
#include "bepu_math.h"

namespace bepu_oracle {

template <class F> struct V4 { F x, y, z, w; };

// ---- shared settings helpers -------------------------------------------------------------------------------------------------------
// Constraints/MotorSettings.cs:L70-98 ComputeSoftness. Rows: MaximumForce, Damping.
template <class F> inline void motor_softness(const F& maximumForce, const F& damping, float dt, F& effectiveMassCFMScale, F& softnessImpulseScale, F& maximumImpulse) {
    F dtWide = bc<F>(dt);
    F dtd = dtWide * damping;
    maximumImpulse = maximumForce * dtWide;
    softnessImpulseScale = bc<F>(1.0f) / (dtd + bc<F>(1.0f));
    effectiveMassCFMScale = dtd * softnessImpulseScale;
}
// Constraints/ServoSettings.cs:L75-85 (1-DOF). Servo rows: MaximumSpeed, BaseSpeed, MaximumForce.
template <class F>
inline void servo_clamped_bias_velocity(const F& error, const F& positionErrorToVelocity, const F& maximumSpeed, const F& baseSpeedSetting, const F& maximumForce, float dt, float inverseDt,
                                        F& clampedBiasVelocity, F& maximumImpulse) {
    F baseSpeed = vmin(baseSpeedSetting, vabs(error) * bc<F>(inverseDt));
    F biasVelocity = error * positionErrorToVelocity;
    clampedBiasVelocity = sel(lt(biasVelocity, bc<F>(0.0f)), vmax(-maximumSpeed, vmin(-baseSpeed, biasVelocity)), vmin(maximumSpeed, vmax(baseSpeed, biasVelocity)));
    maximumImpulse = maximumForce * bc<F>(dt);
}
// ServoSettings.cs:L116-128 (3-DOF, axis + length form)
template <class F>
inline void servo_clamped_bias_velocity3(const V3<F>& errorAxis, const F& errorLength, const F& positionErrorToBiasVelocity, const F& maximumSpeed, const F& baseSpeedSetting,
                                         const F& maximumForce, float dt, float inverseDt, V3<F>& clampedBiasVelocity, F& maximumImpulse) {
    F baseSpeed = vmin(baseSpeedSetting, errorLength * bc<F>(inverseDt));
    F unclampedBiasSpeed = errorLength * positionErrorToBiasVelocity;
    F targetSpeed = vmax(baseSpeed, unclampedBiasSpeed);
    F scl = vmin(bc<F>(1.0f), maximumSpeed / targetSpeed);
    MaskOf<F> useFallback = lt(targetSpeed, bc<F>(1e-10f));
    scl = sel(useFallback, bc<F>(1.0f), scl);
    clampedBiasVelocity = scale(errorAxis, F(scl * unclampedBiasSpeed));
    maximumImpulse = maximumForce * bc<F>(dt);
}
// ServoSettings.cs:L145-151 ClampImpulse (1-DOF)
template <class F> inline void servo_clamp_impulse(const F& maximumImpulse, F& accumulated, F& csi) {
    F previous = accumulated;
    accumulated = vmax(-maximumImpulse, vmin(maximumImpulse, accumulated + csi));
    csi = accumulated - previous;
}
// ServoSettings.cs:L167-178 ClampImpulse (3-DOF)
template <class F> inline void servo_clamp_impulse3(const F& maximumImpulse, V3<F>& accumulated, V3<F>& csi) {
    V3<F> previous = accumulated;
    accumulated = add(accumulated, csi);
    F magnitude = length(accumulated);
    F impulseScale = sel(lt(vabs(magnitude), bc<F>(1e-10f)), bc<F>(1.0f), vmin(maximumImpulse / magnitude, bc<F>(1.0f)));
    accumulated = scale(accumulated, impulseScale);
    csi = sub(accumulated, previous);
}
// Constraints/InequalityHelpers.cs:L16-21 ClampPositive
template <class F> inline void clamp_positive(F& accumulated, F& impulse) {
    F previous = accumulated;
    accumulated = vmax(bc<F>(0.0f), accumulated + impulse);
    impulse = accumulated - previous;
}

// ---- BallSocket (22): BallSocket.cs:L57-90, BallSocketShared.cs ----------------------------------------------------------------------
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, AngularFrequency, TwiceDampingRatio. Impulses: xyz.
template <class F>
inline void ball_socket_apply_impulse(Velocity<F>& vA, Velocity<F>& vB, const V3<F>& offsetA, const V3<F>& offsetB, const Inertia<F>& iA, const Inertia<F>& iB, const V3<F>& csi) {  // BallSocketShared.cs:L29-45
    V3<F> wsi = cross(offsetA, csi);
    vA.ang = add(vA.ang, transform(wsi, iA.t));
    vA.lin = add(vA.lin, scale(csi, iA.inv_mass));
    wsi = cross(csi, offsetB);
    vB.ang = add(vB.ang, transform(wsi, iB.t));
    vB.lin = sub(vB.lin, scale(csi, iB.inv_mass));
}
template <class F>
inline Sym3<F> ball_socket_effective_mass(const Inertia<F>& iA, const Inertia<F>& iB, const V3<F>& offsetA, const V3<F>& offsetB, const F& effectiveMassCFMScale) {  // L10-26
    Sym3<F> inverseEffectiveMass = add(skew_sandwich(offsetA, iA.t), skew_sandwich(offsetB, iB.t));
    F linearContribution = iA.inv_mass + iB.inv_mass;
    inverseEffectiveMass.xx = inverseEffectiveMass.xx + linearContribution;
    inverseEffectiveMass.yy = inverseEffectiveMass.yy + linearContribution;
    inverseEffectiveMass.zz = inverseEffectiveMass.zz + linearContribution;
    return scale(invert(inverseEffectiveMass), effectiveMassCFMScale);
}
template <class F>
inline V3<F> ball_socket_corrective_impulse(const Velocity<F>& vA, const Velocity<F>& vB, const V3<F>& offsetA, const V3<F>& offsetB, const V3<F>& biasVelocity,
                                            const Sym3<F>& effectiveMass, const F& softnessImpulseScale, const V3<F>& accumulated) {  // L47-62
    V3<F> csv = sub(vA.lin, vB.lin);
    csv = add(csv, cross(vA.ang, offsetA));
    csv = add(csv, cross(offsetB, vB.ang));
    csv = sub(biasVelocity, csv);
    V3<F> corrective = transform(csv, effectiveMass);
    return sub(corrective, scale(accumulated, softnessImpulseScale));
}
template <class F> struct BallSocket {
    static constexpr int kPrestepRows = 8, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offsetA = transform(p.get3(0), qA), offsetB = transform(p.get3(3), qB);
        ball_socket_apply_impulse(vA, vB, offsetA, offsetB, iA, iB, a.get3(0));
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offsetA = transform(p.get3(0), qA), offsetB = transform(p.get3(3), qB);
        F pe2v, cfm, soft;
        compute_springiness(p.get(6), p.get(7), dt, pe2v, cfm, soft);
        Sym3<F> effectiveMass = ball_socket_effective_mass(iA, iB, offsetA, offsetB, cfm);
        V3<F> ab = sub(pB, pA);
        V3<F> anchorB = add(ab, offsetB);
        V3<F> error = sub(anchorB, offsetA);
        V3<F> biasVelocity = scale(error, pe2v);
        V3<F> acc = a.get3(0);
        V3<F> corrective = ball_socket_corrective_impulse(vA, vB, offsetA, offsetB, biasVelocity, effectiveMass, soft, acc);
        acc = add(acc, corrective);
        ball_socket_apply_impulse(vA, vB, offsetA, offsetB, iA, iB, corrective);
        a.set3(0, acc);
    }
};

// ---- shared 1-DOF angular apply (SwingLimit.cs:L77-84, TwistServo.cs ApplyImpulse) -------------------------------------------------
template <class F> inline void angular1_apply_impulse(const V3<F>& impulseToVelocityA, const V3<F>& negatedImpulseToVelocityB, const F& csi, V3<F>& wA, V3<F>& wB) {
    wA = add(wA, scale(impulseToVelocityA, csi));
    wB = sub(wB, scale(negatedImpulseToVelocityB, csi));
}

// ---- SwingLimit (25): SwingLimit.cs:L86-149 --------------------------------------------------------------------------------------
// Prestep rows: AxisLocalA xyz, AxisLocalB xyz, MinimumDot, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct SwingLimit {
    static constexpr int kPrestepRows = 9, kImpulseRows = 1;
    static void jacobian(const V3<F>& axisLocalA, const V3<F>& axisLocalB, const Q4<F>& qA, const Q4<F>& qB, V3<F>& axisA, V3<F>& axisB, V3<F>& jacobianA) {
        axisA = transform(axisLocalA, qA);
        axisB = transform(axisLocalB, qB);
        jacobianA = cross(axisA, axisB);
        V3<F> fallback = find_perpendicular(axisA);
        F lengthSquared = dot(jacobianA, jacobianA);
        jacobianA = sel3<F>(lt(lengthSquared, bc<F>(1e-7f)), fallback, jacobianA);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> axisA, axisB, j;
        jacobian(p.get3(0), p.get3(3), qA, qB, axisA, axisB, j);
        angular1_apply_impulse(transform(j, iA.t), transform(j, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> axisA, axisB, j;
        jacobian(p.get3(0), p.get3(3), qA, qB, axisA, axisB, j);
        V3<F> impulseToVelocityA = transform(j, iA.t), negatedImpulseToVelocityB = transform(j, iB.t);
        F angularContributionA = dot(impulseToVelocityA, j), angularContributionB = dot(negatedImpulseToVelocityB, j);
        F pe2v, cfm, soft;
        compute_springiness(p.get(7), p.get(8), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / (angularContributionA + angularContributionB);
        F axisDot = dot(axisA, axisB);
        F error = axisDot - p.get(6);
        F biasVelocity = -vmin(error * bc<F>(inverseDt), error * pe2v);
        F csv = dot(sub(vA.ang, vB.ang), j);
        F acc = a.get(0);
        F csi = effectiveMass * (biasVelocity - csv) - acc * soft;
        clamp_positive(acc, csi);
        angular1_apply_impulse(impulseToVelocityA, negatedImpulseToVelocityB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- Twist family helpers: TwistServo.cs:L69-126 -------------------------------------------------------------------------------------
template <class F> inline void transform_unit_xz(const Q4<F>& r, V3<F>& x, V3<F>& z) {  // QuaternionWide.cs:L467-492
    F qX2 = r.x + r.x, qY2 = r.y + r.y, qZ2 = r.z + r.z;
    F YY = qY2 * r.y, ZZ = qZ2 * r.z;
    x.x = bc<F>(1.0f) - YY - ZZ;
    F XY = qX2 * r.y, ZW = qZ2 * r.w;
    x.y = XY + ZW;
    F XZ = qX2 * r.z, YW = qY2 * r.w;
    x.z = XZ - YW;
    F XX = qX2 * r.x, XW = qX2 * r.w, YZ = qY2 * r.z;
    z.x = XZ + YW;
    z.y = YZ - XW;
    z.z = bc<F>(1.0f) - XX - YY;
}
template <class F>
inline void twist_jacobian_full(const Q4<F>& qA, const Q4<F>& qB, const Q4<F>& localBasisA, const Q4<F>& localBasisB, V3<F>& basisBX, V3<F>& basisBZ, M33<F>& basisA, V3<F>& jacobianA) {  // L69-84
    Q4<F> basisQuaternionA = concatenate(localBasisA, qA);
    Q4<F> basisQuaternionB = concatenate(localBasisB, qB);
    transform_unit_xz(basisQuaternionB, basisBX, basisBZ);
    basisA = matrix_from_quaternion(basisQuaternionA);
    jacobianA = add(basisA.z, basisBZ);
    F len = length(jacobianA);
    jacobianA = scale(jacobianA, F(bc<F>(1.0f) / len));
    jacobianA = sel3<F>(lt(len, bc<F>(1e-10f)), basisA.z, jacobianA);
}
template <class F> inline V3<F> twist_jacobian_only(const Q4<F>& qA, const Q4<F>& qB, const Q4<F>& localBasisA, const Q4<F>& localBasisB) {  // L139-151
    Q4<F> basisQuaternionA = concatenate(localBasisA, qA);
    Q4<F> basisQuaternionB = concatenate(localBasisB, qB);
    V3<F> basisAZ = transform_unit_z(basisQuaternionA), basisBZ = transform_unit_z(basisQuaternionB);
    V3<F> j = add(basisAZ, basisBZ);
    F len = length(j);
    j = scale(j, F(bc<F>(1.0f) / len));
    return sel3<F>(lt(len, bc<F>(1e-10f)), basisAZ, j);
}
template <class F> inline F twist_current_angle(const V3<F>& basisBX, const V3<F>& basisBZ, const M33<F>& basisA) {  // L86-96
    Q4<F> aligningRotation = quaternion_between_normalized(basisBZ, basisA.z);
    V3<F> alignedBasisBX = transform(basisBX, aligningRotation);
    F x = dot(alignedBasisBX, basisA.x), y = dot(alignedBasisBX, basisA.y);
    F absAngle = acos_approx(x);
    return sel(lt(y, bc<F>(0.0f)), -absAngle, absAngle);
}
template <class F>
inline void twist_effective_mass(float dt, const F& angularFrequency, const F& twiceDampingRatio, const Sym3<F>& iA, const Sym3<F>& iB, const V3<F>& jacobianA, V3<F>& impulseToVelocityA,
                                 V3<F>& negatedImpulseToVelocityB, F& positionErrorToVelocity, F& softnessImpulseScale, F& effectiveMass, V3<F>& velocityToImpulseA) {  // L98-126
    impulseToVelocityA = transform(jacobianA, iA);
    negatedImpulseToVelocityB = transform(jacobianA, iB);
    F unsoftenedInverseEffectiveMass = dot(impulseToVelocityA, jacobianA) + dot(negatedImpulseToVelocityB, jacobianA);
    F cfm;
    compute_springiness(angularFrequency, twiceDampingRatio, dt, positionErrorToVelocity, cfm, softnessImpulseScale);
    effectiveMass = cfm / unsoftenedInverseEffectiveMass;
    velocityToImpulseA = scale(jacobianA, effectiveMass);
}

// ---- TwistLimit (27): TwistLimit.cs:L68-123 -------------------------------------------------------------------------------------------
// Prestep rows: LocalBasisA xyzw, LocalBasisB xyzw, MinimumAngle, MaximumAngle, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct TwistLimit {
    static constexpr int kPrestepRows = 12, kImpulseRows = 1;
    static void jacobian(const Q4<F>& qA, const Q4<F>& qB, const Rows<F>& p, F& error, V3<F>& jacobianA) {
        V3<F> basisBX, basisBZ;
        M33<F> basisA;
        twist_jacobian_full(qA, qB, p.get4(0), p.get4(4), basisBX, basisBZ, basisA, jacobianA);
        F angle = twist_current_angle(basisBX, basisBZ, basisA);
        F minError = signed_angle_difference(p.get(8), angle);
        F maxError = signed_angle_difference(p.get(9), angle);
        MaskOf<F> useMin = lt(vabs(minError), vabs(maxError));
        error = sel(useMin, -minError, maxError);
        jacobianA = sel3<F>(useMin, neg(jacobianA), jacobianA);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F error;
        V3<F> j;
        jacobian(qA, qB, p, error, j);
        angular1_apply_impulse(transform(j, iA.t), transform(j, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F error;
        V3<F> j;
        jacobian(qA, qB, p, error, j);
        V3<F> i2vA, ni2vB, v2iA;
        F pe2v, soft, effectiveMass;
        twist_effective_mass(dt, p.get(10), p.get(11), iA.t, iB.t, j, i2vA, ni2vB, pe2v, soft, effectiveMass, v2iA);
        F biasVelocity = sel(lt(error, bc<F>(0.0f)), error * bc<F>(inverseDt), error * pe2v);
        F biasImpulse = biasVelocity * effectiveMass;
        F csiVelocityComponent = dot(sub(vA.ang, vB.ang), v2iA);
        F acc = a.get(0);
        F csi = biasImpulse - acc * soft - csiVelocityComponent;
        clamp_positive(acc, csi);
        angular1_apply_impulse(i2vA, ni2vB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- TwistServo (26): TwistServo.cs:L153-192 ------------------------------------------------------------------------------------------
// Prestep rows: LocalBasisA xyzw, LocalBasisB xyzw, TargetAngle, AngularFrequency, TwiceDampingRatio, MaximumSpeed, BaseSpeed, MaximumForce. Impulse: 1.
template <class F> struct TwistServo {
    static constexpr int kPrestepRows = 14, kImpulseRows = 1;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> j = twist_jacobian_only(qA, qB, p.get4(0), p.get4(4));
        angular1_apply_impulse(transform(j, iA.t), transform(j, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> basisBX, basisBZ, j;
        M33<F> basisA;
        twist_jacobian_full(qA, qB, p.get4(0), p.get4(4), basisBX, basisBZ, basisA, j);
        V3<F> i2vA, ni2vB, v2iA;
        F pe2v, soft, effectiveMass;
        twist_effective_mass(dt, p.get(9), p.get(10), iA.t, iB.t, j, i2vA, ni2vB, pe2v, soft, effectiveMass, v2iA);
        F angle = twist_current_angle(basisBX, basisBZ, basisA);
        F error = signed_angle_difference(p.get(8), angle);
        F clampedBiasVelocity, maximumImpulse;
        servo_clamped_bias_velocity(error, pe2v, p.get(11), p.get(12), p.get(13), dt, inverseDt, clampedBiasVelocity, maximumImpulse);
        F biasImpulse = clampedBiasVelocity * effectiveMass;
        F csiVelocityComponent = dot(sub(vA.ang, vB.ang), v2iA);
        F acc = a.get(0);
        F csi = biasImpulse - acc * soft - csiVelocityComponent;
        F previous = acc;
        acc = vmin(vmax(acc + csi, -maximumImpulse), maximumImpulse);
        csi = acc - previous;
        angular1_apply_impulse(i2vA, ni2vB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- AngularServo (29) / AngularMotor (30): AngularServo.cs:L58-109, AngularMotor.cs:L52-79 -------------------------------------------
template <class F> inline void angular3_apply_impulse(V3<F>& wA, V3<F>& wB, const Sym3<F>& i2vA, const Sym3<F>& ni2vB, const V3<F>& csi) {  // AngularServo.cs:L60-66
    wA = add(wA, transform(csi, i2vA));
    wB = sub(wB, transform(csi, ni2vB));
}
// Prestep rows: TargetVelocityLocalA xyz, MaximumForce, Damping. Impulses: xyz.
template <class F> struct AngularMotor {
    static constexpr int kPrestepRows = 5, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>&, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>&, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        angular3_apply_impulse(vA.ang, vB.ang, iA.t, iB.t, a.get3(0));
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(3), p.get(4), dt, cfm, soft, maximumImpulse);
        Sym3<F> unsoftenedEffectiveMass = invert(add(iA.t, iB.t));
        V3<F> biasVelocity = transform(p.get3(0), qA);
        V3<F> csv = sub(vA.ang, vB.ang);
        csv = sub(biasVelocity, csv);
        V3<F> csi = transform(csv, unsoftenedEffectiveMass);
        csi = scale(csi, cfm);
        V3<F> acc = a.get3(0);
        csi = sub(csi, scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        angular3_apply_impulse(vA.ang, vB.ang, iA.t, iB.t, csi);
        a.set3(0, acc);
    }
};
// Prestep rows: TargetRelativeRotationLocalA xyzw, AngularFrequency, TwiceDampingRatio, MaximumSpeed, BaseSpeed, MaximumForce. Impulses: xyz.
template <class F> struct AngularServo {
    static constexpr int kPrestepRows = 9, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>&, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>&, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        angular3_apply_impulse(vA.ang, vB.ang, iA.t, iB.t, a.get3(0));
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        Q4<F> targetOrientationB = concatenate(p.get4(0), qA);
        Q4<F> inverseTarget = conjugate(targetOrientationB);
        Q4<F> errorRotation = concatenate(inverseTarget, qB);
        V3<F> errorAxis;
        F errorLength;
        axis_angle_from_quaternion(errorRotation, errorAxis, errorLength);
        F pe2v, cfm, soft;
        compute_springiness(p.get(4), p.get(5), dt, pe2v, cfm, soft);
        Sym3<F> unsoftenedEffectiveMass = invert(add(iA.t, iB.t));
        V3<F> clampedBiasVelocity;
        F maximumImpulse;
        servo_clamped_bias_velocity3(errorAxis, errorLength, pe2v, p.get(6), p.get(7), p.get(8), dt, inverseDt, clampedBiasVelocity, maximumImpulse);
        V3<F> csv = sub(vA.ang, vB.ang);
        csv = sub(clampedBiasVelocity, csv);
        V3<F> csi = transform(csv, unsoftenedEffectiveMass);
        csi = scale(csi, cfm);
        V3<F> acc = a.get3(0);
        csi = sub(csi, scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        angular3_apply_impulse(vA.ang, vB.ang, iA.t, iB.t, csi);
        a.set3(0, acc);
    }
};

// ---- Hinge (47): Hinge.cs:L89-222; AngularHinge.cs:L74-112 GetErrorAngles; Symmetric5x5Wide.cs:L21-51 ----------------------------------
template <class F> inline V2<F> hinge_error_angles(const V3<F>& hingeAxisA, const V3<F>& hingeAxisB, const M23<F>& jacobianA) {
    F hingeAxisBDotX = dot(hingeAxisB, jacobianA.x), hingeAxisBDotY = dot(hingeAxisB, jacobianA.y);
    V3<F> onPlaneX = sub(hingeAxisB, scale(jacobianA.x, hingeAxisBDotX));
    V3<F> onPlaneY = sub(hingeAxisB, scale(jacobianA.y, hingeAxisBDotY));
    F xLength = length(onPlaneX), yLength = length(onPlaneY);
    F scaleX = bc<F>(1.0f) / xLength, scaleY = bc<F>(1.0f) / yLength;
    onPlaneX = scale(onPlaneX, scaleX);
    onPlaneY = scale(onPlaneY, scaleY);
    F epsilon = bc<F>(1e-7f);
    onPlaneX = sel3<F>(lt(xLength, epsilon), hingeAxisA, onPlaneX);
    onPlaneY = sel3<F>(lt(yLength, epsilon), hingeAxisA, onPlaneY);
    F hbxha = dot(onPlaneX, hingeAxisA), hbyha = dot(onPlaneY, hingeAxisA);
    V2<F> e{acos_approx(hbxha), acos_approx(hbyha)};
    F hbxay = dot(onPlaneX, jacobianA.y), hbyax = dot(onPlaneY, jacobianA.x);
    e.x = sel(lt(hbxay, bc<F>(0.0f)), e.x, -e.x);
    e.y = sel(lt(hbyax, bc<F>(0.0f)), -e.y, e.y);
    return e;
}
template <class F> struct Sym5 { Sym3<F> A; M23<F> B; Sym2<F> D; };
template <class F> inline Sym5<F> invert5(const Sym3<F>& a, const M23<F>& b, const Sym2<F>& d) {
    Sym2<F> invD = invert(d);
    // Symmetric2x2Wide.MultiplyTransposed(b, invD): bT * invD, stored as 2x3
    M23<F> bTInvD;
    bTInvD.x = {b.x.x * invD.xx + b.y.x * invD.yx, b.x.y * invD.xx + b.y.y * invD.yx, b.x.z * invD.xx + b.y.z * invD.yx};
    bTInvD.y = {b.x.x * invD.yx + b.y.x * invD.yy, b.x.y * invD.yx + b.y.y * invD.yy, b.x.z * invD.yx + b.y.z * invD.yy};
    Sym3<F> bTInvDB = complete_matrix_sandwich_t(bTInvD, b);
    Sym3<F> resultAInverse{a.xx - bTInvDB.xx, a.yx - bTInvDB.yx, a.yy - bTInvDB.yy, a.zx - bTInvDB.zx, a.zy - bTInvDB.zy, a.zz - bTInvDB.zz};
    Sym5<F> r;
    r.A = invert(resultAInverse);
    // Symmetric3x3Wide.MultiplyByTransposed(result.A, bTInvD)
    M23<F> n;
    n.x.x = r.A.xx * bTInvD.x.x + r.A.yx * bTInvD.x.y + r.A.zx * bTInvD.x.z;
    n.y.x = r.A.xx * bTInvD.y.x + r.A.yx * bTInvD.y.y + r.A.zx * bTInvD.y.z;
    n.x.y = r.A.yx * bTInvD.x.x + r.A.yy * bTInvD.x.y + r.A.zy * bTInvD.x.z;
    n.y.y = r.A.yx * bTInvD.y.x + r.A.yy * bTInvD.y.y + r.A.zy * bTInvD.y.z;
    n.x.z = r.A.zx * bTInvD.x.x + r.A.zy * bTInvD.x.y + r.A.zz * bTInvD.x.z;
    n.y.z = r.A.zx * bTInvD.y.x + r.A.zy * bTInvD.y.y + r.A.zz * bTInvD.y.z;
    r.B.x = neg(n.x);
    r.B.y = neg(n.y);
    r.D = add(complete_matrix_sandwich(bTInvD, n), invD);
    return r;
}
// Prestep rows: LocalOffsetA xyz, LocalHingeAxisA xyz, LocalOffsetB xyz, LocalHingeAxisB xyz, AngularFrequency, TwiceDampingRatio. Impulses: BallSocket xyz, Hinge xy.
template <class F> struct Hinge {
    static constexpr int kPrestepRows = 14, kImpulseRows = 5;
    static void apply_impulse(const V3<F>& offsetA, const V3<F>& offsetB, const M23<F>& hingeJacobian, const Inertia<F>& iA, const Inertia<F>& iB, const V3<F>& ballSocket, const V2<F>& hinge,
                              Velocity<F>& vA, Velocity<F>& vB) {  // L70-92
        vA.lin = add(vA.lin, scale(ballSocket, iA.inv_mass));
        V3<F> ballSocketAngularImpulseA = cross(offsetA, ballSocket);
        V3<F> hingeAngularImpulseA = transform(hinge, hingeJacobian);
        V3<F> angularImpulseA = add(ballSocketAngularImpulseA, hingeAngularImpulseA);
        vA.ang = add(vA.ang, transform(angularImpulseA, iA.t));
        vB.lin = sub(vB.lin, scale(ballSocket, iB.inv_mass));
        V3<F> ballSocketAngularImpulseB = cross(ballSocket, offsetB);
        V3<F> angularImpulseB = sub(ballSocketAngularImpulseB, hingeAngularImpulseA);
        vB.ang = add(vB.ang, transform(angularImpulseB, iB.t));
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        M33<F> mA = matrix_from_quaternion(qA);
        V3<F> offsetA = transform(p.get3(0), mA);
        V3<F> offsetB = transform(p.get3(6), qB);
        V3<F> localAX, localAY;
        build_orthonormal_basis(p.get3(3), localAX, localAY);
        M23<F> hingeJacobian{transform(localAX, mA), transform(localAY, mA)};
        apply_impulse(offsetA, offsetB, hingeJacobian, iA, iB, a.get3(0), a.get2(3), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        M33<F> mA = matrix_from_quaternion(qA), mB = matrix_from_quaternion(qB);
        V3<F> offsetA = transform(p.get3(0), mA), hingeAxisA = transform(p.get3(3), mA);
        V3<F> offsetB = transform(p.get3(6), mB), hingeAxisB = transform(p.get3(9), mB);
        V3<F> localAX, localAY;
        build_orthonormal_basis(p.get3(3), localAX, localAY);
        M23<F> hingeJacobian{transform(localAX, mA), transform(localAY, mA)};
        Sym3<F> A = add(skew_sandwich(offsetA, iA.t), skew_sandwich(offsetB, iB.t));
        F linearContribution = iA.inv_mass + iB.inv_mass;
        A.xx = A.xx + linearContribution;
        A.yy = A.yy + linearContribution;
        A.zz = A.zz + linearContribution;
        M23<F> hingeInertiaA = multiply(hingeJacobian, iA.t), hingeInertiaB = multiply(hingeJacobian, iB.t);
        Sym2<F> D = add(complete_matrix_sandwich(hingeInertiaA, hingeJacobian), complete_matrix_sandwich(hingeInertiaB, hingeJacobian));
        M23<F> B;
        B.x = add(cross(hingeInertiaA.x, offsetA), cross(hingeInertiaB.x, offsetB));
        B.y = add(cross(hingeInertiaA.y, offsetA), cross(hingeInertiaB.y, offsetB));
        Sym5<F> effectiveMass = invert5(A, B, D);
        F pe2v, cfm, soft;
        compute_springiness(p.get(12), p.get(13), dt, pe2v, cfm, soft);
        V3<F> anchorB = add(sub(pB, pA), offsetB);
        V3<F> ballSocketError = sub(anchorB, offsetA);
        V3<F> ballSocketBiasVelocity = scale(ballSocketError, pe2v);
        V2<F> errorAngles = hinge_error_angles(hingeAxisA, hingeAxisB, hingeJacobian);
        V2<F> hingeBiasVelocity = scale(errorAngles, F(-pe2v));
        V3<F> ballSocketAngularCSVA = cross(vA.ang, offsetA);
        V2<F> hingeCSVA = transform_by_transpose(vA.ang, hingeJacobian);
        V3<F> ballSocketAngularCSVB = cross(offsetB, vB.ang);
        V2<F> negatedHingeCSVB = transform_by_transpose(vB.ang, hingeJacobian);
        V3<F> ballSocketAngularCSV = add(ballSocketAngularCSVA, ballSocketAngularCSVB);
        V3<F> ballSocketLinearCSV = sub(vA.lin, vB.lin);
        V3<F> ballSocketCSV = add(ballSocketAngularCSV, ballSocketLinearCSV);
        ballSocketCSV = sub(ballSocketBiasVelocity, ballSocketCSV);
        V2<F> hingeCSV = sub(hingeCSVA, negatedHingeCSVB);
        hingeCSV = sub(hingeBiasVelocity, hingeCSV);
        // Symmetric5x5Wide.TransformWithoutOverlap
        const Sym5<F>& m = effectiveMass;
        const V3<F>& v0 = ballSocketCSV;
        const V2<F>& v1 = hingeCSV;
        V3<F> csiBall;
        V2<F> csiHinge;
        csiBall.x = v0.x * m.A.xx + v0.y * m.A.yx + v0.z * m.A.zx + v1.x * m.B.x.x + v1.y * m.B.y.x;
        csiBall.y = v0.x * m.A.yx + v0.y * m.A.yy + v0.z * m.A.zy + v1.x * m.B.x.y + v1.y * m.B.y.y;
        csiBall.z = v0.x * m.A.zx + v0.y * m.A.zy + v0.z * m.A.zz + v1.x * m.B.x.z + v1.y * m.B.y.z;
        csiHinge.x = v0.x * m.B.x.x + v0.y * m.B.x.y + v0.z * m.B.x.z + v1.x * m.D.xx + v1.y * m.D.yx;
        csiHinge.y = v0.x * m.B.y.x + v0.y * m.B.y.y + v0.z * m.B.y.z + v1.x * m.D.yx + v1.y * m.D.yy;
        csiBall = scale(csiBall, cfm);
        csiHinge = scale(csiHinge, cfm);
        V3<F> accBall = a.get3(0);
        V2<F> accHinge = a.get2(3);
        csiBall = sub(csiBall, scale(accBall, soft));
        csiHinge = sub(csiHinge, scale(accHinge, soft));
        accBall = add(accBall, csiBall);
        accHinge = add(accHinge, csiHinge);
        apply_impulse(offsetA, offsetB, hingeJacobian, iA, iB, csiBall, csiHinge, vA, vB);
        a.set3(0, accBall);
        a.set2(3, accHinge);
    }
};

// ---- SwivelHinge (46): SwivelHinge.cs:L86-215; Symmetric4x4Wide.cs:L46-80 ----------------------------------------------------------------
template <class F> struct Sym4 { F xx, yx, yy, zx, zy, zz, wx, wy, wz, ww; };
template <class F> inline Sym4<F> invert4(const Sym4<F>& m) {
    F s0 = m.xx * m.yy - m.yx * m.yx;
    F s1 = m.xx * m.zy - m.yx * m.zx;
    F s2 = m.xx * m.wy - m.yx * m.wx;
    F s3 = m.yx * m.zy - m.yy * m.zx;
    F s4 = m.yx * m.wy - m.yy * m.wx;
    F s5 = m.zx * m.wy - m.zy * m.wx;
    F c5 = m.zz * m.ww - m.wz * m.wz;
    F c4 = m.zy * m.ww - m.wy * m.wz;
    F c3 = m.zy * m.wz - m.wy * m.zz;
    F c2 = m.zx * m.ww - m.wx * m.wz;
    F c1 = m.zx * m.wz - m.wx * m.zz;
    F inverseDeterminant = bc<F>(1.0f) / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * s5);
    Sym4<F> r;
    r.xx = (m.yy * c5 - m.zy * c4 + m.wy * c3) * inverseDeterminant;
    r.yx = (-m.yx * c5 + m.zy * c2 - m.wy * c1) * inverseDeterminant;
    r.yy = (m.xx * c5 - m.zx * c2 + m.wx * c1) * inverseDeterminant;
    r.zx = (m.yx * c4 - m.yy * c2 + m.wy * s5) * inverseDeterminant;
    r.zy = (-m.xx * c4 + m.yx * c2 - m.wx * s5) * inverseDeterminant;
    r.zz = (m.wx * s4 - m.wy * s2 + m.ww * s0) * inverseDeterminant;
    r.wx = (-m.yx * c3 + m.yy * c1 - m.zy * s5) * inverseDeterminant;
    r.wy = (m.xx * c3 - m.yx * c1 + m.zx * s5) * inverseDeterminant;
    r.wz = (-m.wx * s3 + m.wy * s1 - m.wz * s0) * inverseDeterminant;
    r.ww = (m.zx * s3 - m.zy * s1 + m.zz * s0) * inverseDeterminant;
    return r;
}
// Prestep rows: LocalOffsetA xyz, LocalSwivelAxisA xyz, LocalOffsetB xyz, LocalHingeAxisB xyz, AngularFrequency, TwiceDampingRatio. Impulses: xyzw.
template <class F> struct SwivelHinge {
    static constexpr int kPrestepRows = 14, kImpulseRows = 4;
    static void apply_impulse(const V3<F>& offsetA, const V3<F>& offsetB, const V3<F>& swivelHingeJacobian, const Inertia<F>& iA, const Inertia<F>& iB, const V4<F>& csi, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> ballSocketCSI{csi.x, csi.y, csi.z};
        vA.lin = add(vA.lin, scale(ballSocketCSI, iA.inv_mass));
        V3<F> ballSocketAngularImpulseA = cross(offsetA, ballSocketCSI);
        V3<F> swivelHingeAngularImpulseA = scale(swivelHingeJacobian, csi.w);
        vA.ang = add(vA.ang, transform(add(ballSocketAngularImpulseA, swivelHingeAngularImpulseA), iA.t));
        vB.lin = sub(vB.lin, scale(ballSocketCSI, iB.inv_mass));
        V3<F> ballSocketAngularImpulseB = cross(ballSocketCSI, offsetB);
        vB.ang = add(vB.ang, transform(sub(ballSocketAngularImpulseB, swivelHingeAngularImpulseA), iB.t));
    }
    static void jacobian(const Rows<F>& p, const Q4<F>& qA, const Q4<F>& qB, V3<F>& swivelAxis, V3<F>& hingeAxis, V3<F>& offsetA, V3<F>& offsetB, V3<F>& j) {
        M33<F> mA = matrix_from_quaternion(qA), mB = matrix_from_quaternion(qB);
        offsetA = transform(p.get3(0), mA);
        swivelAxis = transform(p.get3(3), mA);
        offsetB = transform(p.get3(6), mB);
        hingeAxis = transform(p.get3(9), mB);
        j = cross(swivelAxis, hingeAxis);
        F lengthSquared = length_squared(j);
        j = sel3<F>(lt(lengthSquared, bc<F>(1e-3f)), hingeAxis, j);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> swivelAxis, hingeAxis, offsetA, offsetB, j;
        jacobian(p, qA, qB, swivelAxis, hingeAxis, offsetA, offsetB, j);
        apply_impulse(offsetA, offsetB, j, iA, iB, V4<F>{a.get(0), a.get(1), a.get(2), a.get(3)}, vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> swivelAxis, hingeAxis, offsetA, offsetB, j;
        jacobian(p, qA, qB, swivelAxis, hingeAxis, offsetA, offsetB, j);
        Sym3<F> upperLeft = add(skew_sandwich(offsetA, iA.t), skew_sandwich(offsetB, iB.t));
        F linearContribution = iA.inv_mass + iB.inv_mass;
        Sym4<F> m;
        m.xx = upperLeft.xx + linearContribution; m.yx = upperLeft.yx; m.yy = upperLeft.yy + linearContribution;
        m.zx = upperLeft.zx; m.zy = upperLeft.zy; m.zz = upperLeft.zz + linearContribution;
        V3<F> swivelHingeInertiaA = transform(j, iA.t), swivelHingeInertiaB = transform(j, iB.t);
        m.ww = dot(swivelHingeInertiaA, j) + dot(swivelHingeInertiaB, j);
        V3<F> upperRight = add(cross(swivelHingeInertiaA, offsetA), cross(swivelHingeInertiaB, offsetB));
        m.wx = upperRight.x; m.wy = upperRight.y; m.wz = upperRight.z;
        Sym4<F> e = invert4(m);
        F pe2v, cfm, soft;
        compute_springiness(p.get(12), p.get(13), dt, pe2v, cfm, soft);
        V3<F> anchorB = add(sub(pB, pA), offsetB);
        V3<F> ballSocketError = sub(anchorB, offsetA);
        V4<F> biasVelocity{ballSocketError.x * pe2v, ballSocketError.y * pe2v, ballSocketError.z * pe2v, bc<F>(0.0f)};
        F error = dot(hingeAxis, swivelAxis);
        biasVelocity.w = pe2v * -error;
        V3<F> ballSocketAngularCSVA = cross(vA.ang, offsetA);
        F swivelHingeCSVA = dot(j, vA.ang);
        V3<F> ballSocketAngularCSVB = cross(offsetB, vB.ang);
        F negatedSwivelHingeCSVB = dot(j, vB.ang);
        V3<F> ballSocketAngularCSV = add(ballSocketAngularCSVA, ballSocketAngularCSVB);
        V3<F> ballSocketLinearCSV = sub(vA.lin, vB.lin);
        V4<F> csv{ballSocketAngularCSV.x + ballSocketLinearCSV.x, ballSocketAngularCSV.y + ballSocketLinearCSV.y, ballSocketAngularCSV.z + ballSocketLinearCSV.z, swivelHingeCSVA - negatedSwivelHingeCSVB};
        csv = {biasVelocity.x - csv.x, biasVelocity.y - csv.y, biasVelocity.z - csv.z, biasVelocity.w - csv.w};
        V4<F> csi;
        csi.x = csv.x * e.xx + csv.y * e.yx + csv.z * e.zx + csv.w * e.wx;
        csi.y = csv.x * e.yx + csv.y * e.yy + csv.z * e.zy + csv.w * e.wy;
        csi.z = csv.x * e.zx + csv.y * e.zy + csv.z * e.zz + csv.w * e.wz;
        csi.w = csv.x * e.wx + csv.y * e.wy + csv.z * e.wz + csv.w * e.ww;
        csi = {csi.x * cfm, csi.y * cfm, csi.z * cfm, csi.w * cfm};
        V4<F> acc{a.get(0), a.get(1), a.get(2), a.get(3)};
        csi = {csi.x - acc.x * soft, csi.y - acc.y * soft, csi.z - acc.z * soft, csi.w - acc.w * soft};
        acc = {acc.x + csi.x, acc.y + csi.y, acc.z + csi.z, acc.w + csi.w};
        apply_impulse(offsetA, offsetB, j, iA, iB, csi, vA, vB);
        a.set(0, acc.x); a.set(1, acc.y); a.set(2, acc.z); a.set(3, acc.w);
    }
};

template <class R> inline void register_joints(R& r) {
    typedef typename R::Lane F;
    r.template joint2<BallSocket<F>>(22);
    r.template joint2<SwingLimit<F>>(25);
    r.template joint2<TwistServo<F>>(26);
    r.template joint2<TwistLimit<F>>(27);
    r.template joint2<AngularServo<F>>(29);
    r.template joint2<AngularMotor<F>>(30);
    r.template joint2<SwivelHinge<F>>(46);
    r.template joint2<Hinge<F>>(47);
}

}  // namespace bepu_oracle
