// ORACLE — test infrastructure only (see bepu_math.h header). The remaining joint / motor / servo / limit constraint functions of the
// reference's default type set (DefaultTypes.cs), each restating the reference file:line cited next to it. Pinned bit for bit to the reference's C# text
// through oracle/ref_transpile (tests/test_oracle_pinned_to_reference.py).
//
// MathHelper.FastReciprocal / FastReciprocalSquareRoot (MathHelper.cs:L380-413) are hardware approximations (rcpps / rsqrtps) on x86 and
// exact 1/v, 1/sqrt(v) elsewhere; their approximation error differs between CPU vendors, so this restatement (and the CUDA path it checks)
// uses the exact fallback definition the reference itself carries for non-AVX targets.
#pragma once
#include "bepu_joints.h"

namespace bepu_oracle {

// ---- more settings helpers ---------------------------------------------------------------------------------------------------------
// ServoSettings.cs:L132-142 (3-DOF, error-vector form)
template <class F>
inline void servo_clamped_bias_velocity3e(const V3<F>& error, const F& positionErrorToBiasVelocity, const F& maximumSpeed, const F& baseSpeedSetting, const F& maximumForce, float dt,
                                          float inverseDt, V3<F>& clampedBiasVelocity, F& maximumImpulse) {
    F errorLength = length(error);
    V3<F> errorAxis = scale(error, F(bc<F>(1.0f) / errorLength));
    errorAxis = sel3<F>(lt(errorLength, bc<F>(1e-10f)), v3bc<F>(0.0f, 0.0f, 0.0f), errorAxis);
    servo_clamped_bias_velocity3(errorAxis, errorLength, positionErrorToBiasVelocity, maximumSpeed, baseSpeedSetting, maximumForce, dt, inverseDt, clampedBiasVelocity, maximumImpulse);
}
// ServoSettings.cs:L88-113 (2-DOF, error-vector form)
template <class F>
inline void servo_clamped_bias_velocity2e(const V2<F>& error, const F& positionErrorToBiasVelocity, const F& maximumSpeed, const F& baseSpeedSetting, const F& maximumForce, float dt,
                                          float inverseDt, V2<F>& clampedBiasVelocity, F& maximumImpulse) {
    F errorLength = length(error);
    V2<F> errorAxis = scale(error, F(bc<F>(1.0f) / errorLength));
    MaskOf<F> useFallbackAxis = lt(errorLength, bc<F>(1e-10f));
    errorAxis.x = sel(useFallbackAxis, bc<F>(0.0f), errorAxis.x);
    errorAxis.y = sel(useFallbackAxis, bc<F>(0.0f), errorAxis.y);
    F baseSpeed = vmin(baseSpeedSetting, errorLength * bc<F>(inverseDt));
    F unclampedBiasSpeed = errorLength * positionErrorToBiasVelocity;
    F targetSpeed = vmax(baseSpeed, unclampedBiasSpeed);
    F scl = vmin(bc<F>(1.0f), maximumSpeed / targetSpeed);
    scl = sel(lt(targetSpeed, bc<F>(1e-10f)), bc<F>(1.0f), scl);
    clampedBiasVelocity = scale(errorAxis, F(scl * unclampedBiasSpeed));
    maximumImpulse = maximumForce * bc<F>(dt);
}
// ServoSettings.cs:L153-164 ClampImpulse (2-DOF)
template <class F> inline void servo_clamp_impulse2(const F& maximumImpulse, V2<F>& accumulated, V2<F>& csi) {
    V2<F> previous = accumulated;
    V2<F> unclamped = add(accumulated, csi);
    F magnitude = length(unclamped);
    F impulseScale = sel(lt(vabs(magnitude), bc<F>(1e-10f)), bc<F>(1.0f), vmin(maximumImpulse / magnitude, bc<F>(1.0f)));
    accumulated = scale(unclamped, impulseScale);
    csi = sub(accumulated, previous);
}
// InequalityHelpers.cs:L9-12
template <class F> inline F inequality_bias_velocity(const F& error, const F& positionErrorToVelocity, float inverseDt) { return vmin(error * bc<F>(inverseDt), error * positionErrorToVelocity); }

// ---- Weld (31): Weld.cs:L83-221; Symmetric6x6Wide.cs:L84-129 LDLTSolve ----------------------------------------------------------------
// Prestep rows: LocalOffset xyz, LocalOrientation xyzw, AngularFrequency, TwiceDampingRatio. Impulses: Orientation xyz, Offset xyz.
template <class F> inline M33<F> cross_product_matrix(const V3<F>& v) {  // Matrix3x3Wide.cs:L169-180
    F zero = bc<F>(0.0f);
    return M33<F>{V3<F>{zero, -v.z, v.y}, V3<F>{v.z, zero, -v.x}, V3<F>{-v.y, v.x, zero}};
}
template <class F> inline M33<F> multiply_sym_matrix(const Sym3<F>& a, const M33<F>& b) {  // Symmetric3x3Wide.cs:L343-356
    M33<F> r;
    r.x.x = a.xx * b.x.x + a.yx * b.y.x + a.zx * b.z.x;
    r.x.y = a.xx * b.x.y + a.yx * b.y.y + a.zx * b.z.y;
    r.x.z = a.xx * b.x.z + a.yx * b.y.z + a.zx * b.z.z;
    r.y.x = a.yx * b.x.x + a.yy * b.y.x + a.zy * b.z.x;
    r.y.y = a.yx * b.x.y + a.yy * b.y.y + a.zy * b.z.y;
    r.y.z = a.yx * b.x.z + a.yy * b.y.z + a.zy * b.z.z;
    r.z.x = a.zx * b.x.x + a.zy * b.y.x + a.zz * b.z.x;
    r.z.y = a.zx * b.x.y + a.zy * b.y.y + a.zz * b.z.y;
    r.z.z = a.zx * b.x.z + a.zy * b.y.z + a.zz * b.z.z;
    return r;
}
template <class F> inline Sym3<F> complete_matrix_sandwich_transpose(const M33<F>& a, const M33<F>& b) {  // Symmetric3x3Wide.cs:L508-518
    Sym3<F> r;
    r.xx = a.x.x * b.x.x + a.y.x * b.y.x + a.z.x * b.z.x;
    r.yx = a.x.y * b.x.x + a.y.y * b.y.x + a.z.y * b.z.x;
    r.yy = a.x.y * b.x.y + a.y.y * b.y.y + a.z.y * b.z.y;
    r.zx = a.x.z * b.x.x + a.y.z * b.y.x + a.z.z * b.z.x;
    r.zy = a.x.z * b.x.y + a.y.z * b.y.y + a.z.z * b.z.y;
    r.zz = a.x.z * b.x.z + a.y.z * b.y.z + a.z.z * b.z.z;
    return r;
}
template <class F> inline void ldlt_solve6(const V3<F>& v0, const V3<F>& v1, const Sym3<F>& a, const M33<F>& b, const Sym3<F>& d, V3<F>& result0, V3<F>& result1) {
    F one = bc<F>(1.0f);
    F d1 = a.xx;
    F inverseD1 = one / d1;
    F l21 = inverseD1 * a.yx, l31 = inverseD1 * a.zx, l41 = inverseD1 * b.x.x, l51 = inverseD1 * b.x.y, l61 = inverseD1 * b.x.z;
    F d2 = a.yy - l21 * l21 * d1;
    F inverseD2 = one / d2;
    F l32 = inverseD2 * (a.zy - l31 * l21 * d1);
    F l42 = inverseD2 * (b.y.x - l41 * l21 * d1);
    F l52 = inverseD2 * (b.y.y - l51 * l21 * d1);
    F l62 = inverseD2 * (b.y.z - l61 * l21 * d1);
    F d3 = a.zz - l31 * l31 * d1 - l32 * l32 * d2;
    F inverseD3 = one / d3;
    F l43 = inverseD3 * (b.z.x - l41 * l31 * d1 - l42 * l32 * d2);
    F l53 = inverseD3 * (b.z.y - l51 * l31 * d1 - l52 * l32 * d2);
    F l63 = inverseD3 * (b.z.z - l61 * l31 * d1 - l62 * l32 * d2);
    F d4 = d.xx - l41 * l41 * d1 - l42 * l42 * d2 - l43 * l43 * d3;
    F inverseD4 = one / d4;
    F l54 = inverseD4 * (d.yx - l51 * l41 * d1 - l52 * l42 * d2 - l53 * l43 * d3);
    F l64 = inverseD4 * (d.zx - l61 * l41 * d1 - l62 * l42 * d2 - l63 * l43 * d3);
    F d5 = d.yy - l51 * l51 * d1 - l52 * l52 * d2 - l53 * l53 * d3 - l54 * l54 * d4;
    F inverseD5 = one / d5;
    F l65 = inverseD5 * (d.zy - l61 * l51 * d1 - l62 * l52 * d2 - l63 * l53 * d3 - l64 * l54 * d4);
    F d6 = d.zz - l61 * l61 * d1 - l62 * l62 * d2 - l63 * l63 * d3 - l64 * l64 * d4 - l65 * l65 * d5;
    F inverseD6 = one / d6;
    result0.x = v0.x;
    result0.y = v0.y - l21 * result0.x;
    result0.z = v0.z - l31 * result0.x - l32 * result0.y;
    result1.x = v1.x - l41 * result0.x - l42 * result0.y - l43 * result0.z;
    result1.y = v1.y - l51 * result0.x - l52 * result0.y - l53 * result0.z - l54 * result1.x;
    result1.z = v1.z - l61 * result0.x - l62 * result0.y - l63 * result0.z - l64 * result1.x - l65 * result1.y;
    result1.z = result1.z * inverseD6;
    result1.y = result1.y * inverseD5 - l65 * result1.z;
    result1.x = result1.x * inverseD4 - l64 * result1.z - l54 * result1.y;
    result0.z = result0.z * inverseD3 - l63 * result1.z - l53 * result1.y - l43 * result1.x;
    result0.y = result0.y * inverseD2 - l62 * result1.z - l52 * result1.y - l42 * result1.x - l32 * result0.z;
    result0.x = result0.x * inverseD1 - l61 * result1.z - l51 * result1.y - l41 * result1.x - l31 * result0.z - l21 * result0.y;
}
template <class F> struct Weld {
    static constexpr int kPrestepRows = 9, kImpulseRows = 6;
    static void apply_impulse(const Inertia<F>& iA, const Inertia<F>& iB, const V3<F>& offset, const V3<F>& orientationCSI, const V3<F>& offsetCSI, Velocity<F>& vA, Velocity<F>& vB) {  // L85-114
        vA.lin = add(vA.lin, scale(offsetCSI, iA.inv_mass));
        V3<F> offsetWorldImpulse = cross(offset, offsetCSI);
        V3<F> angularImpulseA = add(offsetWorldImpulse, orientationCSI);
        vA.ang = add(vA.ang, transform(angularImpulseA, iA.t));
        vB.lin = sub(vB.lin, scale(offsetCSI, iB.inv_mass));
        vB.ang = sub(vB.ang, transform(orientationCSI, iB.t));
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offset = transform(p.get3(0), qA);
        apply_impulse(iA, iB, offset, a.get3(0), a.get3(3), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offset = transform(p.get3(0), qA);
        Sym3<F> jmjtA = add(iA.t, iB.t);
        M33<F> xAB = cross_product_matrix(offset);
        M33<F> jmjtB = multiply_sym_matrix(iA.t, xAB);
        Sym3<F> jmjtD = complete_matrix_sandwich_transpose(xAB, jmjtB);
        F diagonalAdd = iA.inv_mass + iB.inv_mass;
        jmjtD.xx = jmjtD.xx + diagonalAdd;
        jmjtD.yy = jmjtD.yy + diagonalAdd;
        jmjtD.zz = jmjtD.zz + diagonalAdd;
        V3<F> positionError = sub(sub(pB, pA), offset);
        Q4<F> targetOrientationB = concatenate(p.get4(3), qA);
        Q4<F> rotationError = concatenate(conjugate(targetOrientationB), qB);
        V3<F> rotationErrorAxis;
        F rotationErrorLength;
        axis_angle_from_quaternion(rotationError, rotationErrorAxis, rotationErrorLength);
        F pe2v, cfm, soft;
        compute_springiness(p.get(7), p.get(8), dt, pe2v, cfm, soft);
        V3<F> orientationBiasVelocity = scale(rotationErrorAxis, F(rotationErrorLength * pe2v));
        V3<F> offsetBiasVelocity = scale(positionError, pe2v);
        V3<F> orientationCSV, offsetCSV;
        orientationCSV.x = orientationBiasVelocity.x - vA.ang.x + vB.ang.x;
        orientationCSV.y = orientationBiasVelocity.y - vA.ang.y + vB.ang.y;
        orientationCSV.z = orientationBiasVelocity.z - vA.ang.z + vB.ang.z;
        offsetCSV.x = offsetBiasVelocity.x - vA.lin.x + vB.lin.x - (vA.ang.y * offset.z - vA.ang.z * offset.y);
        offsetCSV.y = offsetBiasVelocity.y - vA.lin.y + vB.lin.y - (vA.ang.z * offset.x - vA.ang.x * offset.z);
        offsetCSV.z = offsetBiasVelocity.z - vA.lin.z + vB.lin.z - (vA.ang.x * offset.y - vA.ang.y * offset.x);
        V3<F> orientationCSI, offsetCSI;
        ldlt_solve6(orientationCSV, offsetCSV, jmjtA, jmjtB, jmjtD, orientationCSI, offsetCSI);
        V3<F> accOrientation = a.get3(0), accOffset = a.get3(3);
        orientationCSI.x = orientationCSI.x * cfm - accOrientation.x * soft;
        orientationCSI.y = orientationCSI.y * cfm - accOrientation.y * soft;
        orientationCSI.z = orientationCSI.z * cfm - accOrientation.z * soft;
        accOrientation = add(accOrientation, orientationCSI);
        offsetCSI.x = offsetCSI.x * cfm - accOffset.x * soft;
        offsetCSI.y = offsetCSI.y * cfm - accOffset.y * soft;
        offsetCSI.z = offsetCSI.z * cfm - accOffset.z * soft;
        accOffset = add(accOffset, offsetCSI);
        apply_impulse(iA, iB, offset, orientationCSI, offsetCSI, vA, vB);
        a.set3(0, accOrientation);
        a.set3(3, accOffset);
    }
};

// ---- AngularHinge (23): AngularHinge.cs:L71-223 -------------------------------------------------------------------------------------------
// Prestep rows: LocalHingeAxisA xyz, LocalHingeAxisB xyz, AngularFrequency, TwiceDampingRatio. Impulses: xy.
template <class F> struct AngularHinge {
    static constexpr int kPrestepRows = 8, kImpulseRows = 2;
    static void apply_impulse(const M23<F>& i2vA, const M23<F>& ni2vB, const V2<F>& csi, V3<F>& wA, V3<F>& wB) {  // L114-120
        wA = add(wA, transform(csi, i2vA));
        wB = sub(wB, transform(csi, ni2vB));
    }
    static void jacobians(const V3<F>& localHingeAxisA, const Q4<F>& qA, V3<F>& hingeAxisA, M23<F>& jacobianA) {  // L122-131
        V3<F> localAX, localAY;
        build_orthonormal_basis(localHingeAxisA, localAX, localAY);
        M33<F> mA = matrix_from_quaternion(qA);
        hingeAxisA = transform(localHingeAxisA, mA);
        jacobianA.x = transform(localAX, mA);
        jacobianA.y = transform(localAY, mA);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> hingeAxisA;
        M23<F> jacobianA;
        jacobians(p.get3(0), qA, hingeAxisA, jacobianA);
        apply_impulse(multiply(jacobianA, iA.t), multiply(jacobianA, iB.t), a.get2(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> hingeAxisA;
        M23<F> jacobianA;
        jacobians(p.get3(0), qA, hingeAxisA, jacobianA);
        V3<F> hingeAxisB = transform(p.get3(3), qB);
        M23<F> i2vA = multiply(jacobianA, iA.t), ni2vB = multiply(jacobianA, iB.t);
        Sym2<F> inverseEffectiveMass = add(complete_matrix_sandwich(i2vA, jacobianA), complete_matrix_sandwich(ni2vB, jacobianA));
        Sym2<F> effectiveMass = invert(inverseEffectiveMass);
        F pe2v, cfm, soft;
        compute_springiness(p.get(6), p.get(7), dt, pe2v, cfm, soft);
        V2<F> errorAngle = hinge_error_angles(hingeAxisA, hingeAxisB, jacobianA);
        V2<F> biasVelocity = scale(errorAngle, F(-pe2v));
        V2<F> biasImpulse = transform(biasVelocity, effectiveMass);
        V3<F> difference = sub(vA.ang, vB.ang);
        V2<F> csv = transform_by_transpose(difference, jacobianA);
        V2<F> csi = transform(csv, effectiveMass);
        csi = scale(csi, cfm);
        V2<F> acc = a.get2(0);
        V2<F> softnessContribution = scale(acc, soft);
        csi = add(softnessContribution, csi);
        csi = sub(biasImpulse, csi);
        acc = add(acc, csi);
        apply_impulse(i2vA, ni2vB, csi, vA.ang, vB.ang);
        a.set2(0, acc);
    }
};

// ---- AngularSwivelHinge (24): AngularSwivelHinge.cs:L71-148 ---------------------------------------------------------------------------------
// Prestep rows: LocalSwivelAxisA xyz, LocalHingeAxisB xyz, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct AngularSwivelHinge {
    static constexpr int kPrestepRows = 8, kImpulseRows = 1;
    static void jacobian(const Rows<F>& p, const Q4<F>& qA, const Q4<F>& qB, V3<F>& swivelAxis, V3<F>& hingeAxis, V3<F>& jacobianA) {  // L82-94
        swivelAxis = transform(p.get3(0), qA);
        hingeAxis = transform(p.get3(3), qB);
        jacobianA = cross(swivelAxis, hingeAxis);
        V3<F> fallbackJacobian = find_perpendicular(swivelAxis);
        F jacobianLengthSquared = dot(jacobianA, jacobianA);
        jacobianA = sel3<F>(lt(jacobianLengthSquared, bc<F>(1e-3f)), fallbackJacobian, jacobianA);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> swivelAxis, hingeAxis, j;
        jacobian(p, qA, qB, swivelAxis, hingeAxis, j);
        angular1_apply_impulse(transform(j, iA.t), transform(j, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> swivelAxis, hingeAxis, j;
        jacobian(p, qA, qB, swivelAxis, hingeAxis, j);
        V3<F> i2vA = transform(j, iA.t), ni2vB = transform(j, iB.t);
        F angularA = dot(i2vA, j), angularB = dot(ni2vB, j);
        F pe2v, cfm, soft;
        compute_springiness(p.get(6), p.get(7), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / (angularA + angularB);
        F error = dot(hingeAxis, swivelAxis);
        F biasVelocity = -(pe2v * error);
        F csv = dot(sub(vA.ang, vB.ang), j);
        F acc = a.get(0);
        F csi = effectiveMass * (biasVelocity - csv) - acc * soft;
        acc = acc + csi;
        angular1_apply_impulse(i2vA, ni2vB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- TwistMotor (28): TwistMotor.cs:L77-126 ---------------------------------------------------------------------------------------------------
// Prestep rows: LocalAxisA xyz, LocalAxisB xyz, TargetVelocity, MaximumForce, Damping. Impulse: 1.
template <class F> struct TwistMotor {
    static constexpr int kPrestepRows = 9, kImpulseRows = 1;
    static V3<F> jacobian(const Q4<F>& qA, const Q4<F>& qB, const V3<F>& localAxisA, const V3<F>& localAxisB) {  // L79-89
        V3<F> axisA = transform(localAxisA, qA), axisB = transform(localAxisB, qB);
        V3<F> j = add(axisA, axisB);
        F len = length(j);
        j = scale(j, F(bc<F>(1.0f) / len));
        return sel3<F>(lt(len, bc<F>(1e-10f)), axisA, j);
    }
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> j = jacobian(qA, qB, p.get3(0), p.get3(3));
        angular1_apply_impulse(transform(j, iA.t), transform(j, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> j = jacobian(qA, qB, p.get3(0), p.get3(3));
        V3<F> i2vA = transform(j, iA.t), ni2vB = transform(j, iB.t);  // TwistServo.cs:L133-144
        F unsoftenedInverseEffectiveMass = dot(i2vA, j) + dot(ni2vB, j);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(7), p.get(8), dt, cfm, soft, maximumImpulse);
        F effectiveMass = cfm / unsoftenedInverseEffectiveMass;
        V3<F> velocityToImpulseA = scale(j, effectiveMass);
        F biasImpulse = p.get(6) * effectiveMass;
        F csiVelocityComponent = dot(sub(vA.ang, vB.ang), velocityToImpulseA);
        F acc = a.get(0);
        F csi = biasImpulse - acc * soft - csiVelocityComponent;
        F previous = acc;
        acc = vmax(vmin(acc + csi, maximumImpulse), -maximumImpulse);
        csi = acc - previous;
        angular1_apply_impulse(i2vA, ni2vB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- AngularAxisMotor (41): AngularAxisMotor.cs:L69-106 -----------------------------------------------------------------------------------------
// Prestep rows: LocalAxisA xyz, TargetVelocity, MaximumForce, Damping. Impulse: 1.
template <class F> struct AngularAxisMotor {
    static constexpr int kPrestepRows = 6, kImpulseRows = 1;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> axis = transform(p.get3(0), qA);
        angular1_apply_impulse(transform(axis, iA.t), transform(axis, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> jA = transform(p.get3(0), qA);
        V3<F> jIA = transform(jA, iA.t);
        F contributionA = dot(jA, jIA);
        V3<F> jIB = transform(jA, iB.t);
        F contributionB = dot(jA, jIB);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(4), p.get(5), dt, cfm, soft, maximumImpulse);
        F acc = a.get(0);
        F csi = (p.get(3) + dot(vB.ang, jA) - dot(vA.ang, jA)) * cfm / (contributionA + contributionB) - acc * soft;
        servo_clamp_impulse(maximumImpulse, acc, csi);
        angular1_apply_impulse(jIA, jIB, csi, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- AngularAxisGearMotor (54): AngularAxisGearMotor.cs:L70-114 ----------------------------------------------------------------------------------
// Prestep rows: LocalAxisA xyz, VelocityScale, MaximumForce, Damping. Impulse: 1.
// Reference behaviour reproduced as written: Solve's final ApplyImpulse (L112) is given the clamped ACCUMULATED impulse, not the corrective impulse.
template <class F> struct AngularAxisGearMotor {
    static constexpr int kPrestepRows = 6, kImpulseRows = 1;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> axis = transform(p.get3(0), qA);
        V3<F> jA = scale(axis, p.get(3));
        angular1_apply_impulse(transform(jA, iA.t), transform(axis, iB.t), a.get(0), vA.ang, vB.ang);
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>&, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> axis = transform(p.get3(0), qA);
        V3<F> jA = scale(axis, p.get(3));
        V3<F> i2vA = transform(jA, iA.t);
        F contributionA = dot(jA, i2vA);
        V3<F> ni2vB = transform(axis, iB.t);
        F contributionB = dot(axis, ni2vB);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(4), p.get(5), dt, cfm, soft, maximumImpulse);
        F effectiveMass = cfm / (contributionA + contributionB);
        F unscaledCSVA = dot(vA.ang, jA);
        F negatedCSVB = dot(vB.ang, axis);
        F acc = a.get(0);
        F csi = (negatedCSVB - unscaledCSVA) * effectiveMass - acc * soft;
        servo_clamp_impulse(maximumImpulse, acc, csi);
        angular1_apply_impulse(i2vA, ni2vB, acc, vA.ang, vB.ang);
        a.set(0, acc);
    }
};

// ---- BallSocketMotor (52) / BallSocketServo (53): BallSocketMotor.cs:L68-97, BallSocketServo.cs:L75-107, BallSocketShared.cs:L126-134 -------------
template <class F>
inline void ball_socket_solve_clamped(Velocity<F>& vA, Velocity<F>& vB, const V3<F>& offsetA, const V3<F>& offsetB, const V3<F>& biasVelocity, const Sym3<F>& effectiveMass,
                                      const F& soft, const F& maximumImpulse, V3<F>& acc, const Inertia<F>& iA, const Inertia<F>& iB) {
    V3<F> corrective = ball_socket_corrective_impulse(vA, vB, offsetA, offsetB, biasVelocity, effectiveMass, soft, acc);
    servo_clamp_impulse3(maximumImpulse, acc, corrective);
    ball_socket_apply_impulse(vA, vB, offsetA, offsetB, iA, iB, corrective);
}
// Prestep rows: LocalOffsetB xyz, TargetVelocityLocalA xyz, MaximumForce, Damping. Impulses: xyz.
template <class F> struct BallSocketMotor {
    static constexpr int kPrestepRows = 8, kImpulseRows = 3;
    static void warm_start(const V3<F>& pA, const Q4<F>&, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> targetOffsetB = transform(p.get3(0), qB);
        ball_socket_apply_impulse(vA, vB, add(sub(pB, pA), targetOffsetB), targetOffsetB, iA, iB, a.get3(0));
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> targetOffsetB = transform(p.get3(0), qB);
        V3<F> offsetA = add(sub(pB, pA), targetOffsetB);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(6), p.get(7), dt, cfm, soft, maximumImpulse);
        Sym3<F> effectiveMass = ball_socket_effective_mass(iA, iB, offsetA, targetOffsetB, cfm);
        V3<F> biasVelocity = neg(transform(p.get3(3), qA));
        V3<F> acc = a.get3(0);
        ball_socket_solve_clamped(vA, vB, offsetA, targetOffsetB, biasVelocity, effectiveMass, soft, maximumImpulse, acc, iA, iB);
        a.set3(0, acc);
    }
};
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, AngularFrequency, TwiceDampingRatio, MaximumSpeed, BaseSpeed, MaximumForce. Impulses: xyz.
template <class F> struct BallSocketServo {
    static constexpr int kPrestepRows = 11, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>&, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offsetA = transform(p.get3(0), qA), offsetB = transform(p.get3(3), qB);
        ball_socket_apply_impulse(vA, vB, offsetA, offsetB, iA, iB, a.get3(0));
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> offsetA = transform(p.get3(0), qA), offsetB = transform(p.get3(3), qB);
        F pe2v, cfm, soft;
        compute_springiness(p.get(6), p.get(7), dt, pe2v, cfm, soft);
        Sym3<F> effectiveMass = ball_socket_effective_mass(iA, iB, offsetA, offsetB, cfm);
        V3<F> ab = sub(pB, pA);
        V3<F> anchorB = add(ab, offsetB);
        V3<F> error = sub(anchorB, offsetA);
        V3<F> biasVelocity;
        F maximumImpulse;
        servo_clamped_bias_velocity3e(error, pe2v, p.get(8), p.get(9), p.get(10), dt, inverseDt, biasVelocity, maximumImpulse);
        V3<F> acc = a.get3(0);
        ball_socket_solve_clamped(vA, vB, offsetA, offsetB, biasVelocity, effectiveMass, soft, maximumImpulse, acc, iA, iB);
        a.set3(0, acc);
    }
};

// ---- DistanceServo (33): DistanceServo.cs:L107-226 ------------------------------------------------------------------------------------------------
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, TargetDistance, MaximumSpeed, BaseSpeed, MaximumForce, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct DistanceServo {
    static constexpr int kPrestepRows = 12, kImpulseRows = 1;
    static void get_distance(const Q4<F>& qA, const V3<F>& ab, const Q4<F>& qB, const V3<F>& localOffsetA, const V3<F>& localOffsetB, V3<F>& anchorOffsetA, V3<F>& anchorOffsetB,
                             V3<F>& anchorOffset, F& distance) {  // L109-117
        anchorOffsetA = transform(localOffsetA, qA);
        anchorOffsetB = transform(localOffsetB, qB);
        V3<F> anchorB = add(anchorOffsetB, ab);
        anchorOffset = sub(anchorB, anchorOffsetA);
        distance = length(anchorOffset);
    }
    static void jacobian(const F& distance, const V3<F>& anchorOffsetA, const V3<F>& anchorOffsetB, V3<F>& direction, V3<F>& angularJA, V3<F>& angularJB) {  // L119-130
        direction = sel3<F>(lt(distance, bc<F>(1e-9f)), v3bc<F>(1.0f, 0.0f, 0.0f), direction);
        angularJA = cross(anchorOffsetA, direction);
        angularJB = cross(direction, anchorOffsetB);
    }
    static void apply_impulse(const F& imA, const F& imB, const V3<F>& direction, const V3<F>& ai2vA, const V3<F>& ai2vB, const F& csi, Velocity<F>& vA, Velocity<F>& vB) {  // L160-183
        vA.lin = add(scale(direction, F(csi * imA)), vA.lin);
        vA.ang = add(scale(ai2vA, csi), vA.ang);
        vB.lin = sub(vB.lin, scale(direction, F(csi * imB)));
        vB.ang = add(scale(ai2vB, csi), vB.ang);
    }
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> anchorOffsetA, anchorOffsetB, anchorOffset, angularJA, angularJB;
        F distance;
        get_distance(qA, sub(pB, pA), qB, p.get3(0), p.get3(3), anchorOffsetA, anchorOffsetB, anchorOffset, distance);
        V3<F> direction = scale(anchorOffset, F(bc<F>(1.0f) / distance));
        jacobian(distance, anchorOffsetA, anchorOffsetB, direction, angularJA, angularJB);
        apply_impulse(iA.inv_mass, iB.inv_mass, direction, transform(angularJA, iA.t), transform(angularJB, iB.t), a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> anchorOffsetA, anchorOffsetB, anchorOffset, angularJA, angularJB;
        F distance;
        get_distance(qA, sub(pB, pA), qB, p.get3(0), p.get3(3), anchorOffsetA, anchorOffsetB, anchorOffset, distance);
        V3<F> direction = scale(anchorOffset, F(bc<F>(1.0f) / distance));
        jacobian(distance, anchorOffsetA, anchorOffsetB, direction, angularJA, angularJB);  // ComputeTransforms L132-158
        V3<F> ai2vA = transform(angularJA, iA.t), ai2vB = transform(angularJB, iB.t);
        F angularContributionA = dot(angularJA, ai2vA), angularContributionB = dot(angularJB, ai2vB);
        F inverseEffectiveMass = iA.inv_mass + iB.inv_mass + angularContributionA + angularContributionB;
        F pe2v, cfm, soft;
        compute_springiness(p.get(10), p.get(11), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / inverseEffectiveMass;
        F error = distance - p.get(6);
        F clampedBiasVelocity, maximumImpulse;
        servo_clamped_bias_velocity(error, pe2v, p.get(7), p.get(8), p.get(9), dt, inverseDt, clampedBiasVelocity, maximumImpulse);
        F linearCSVA = dot(vA.lin, direction), negatedLinearCSVB = dot(vB.lin, direction);
        F angularCSVA = dot(vA.ang, angularJA), angularCSVB = dot(vB.ang, angularJB);
        F acc = a.get(0);
        F csi = (clampedBiasVelocity - linearCSVA - angularCSVA + negatedLinearCSVB - angularCSVB) * effectiveMass - acc * soft;
        servo_clamp_impulse(maximumImpulse, acc, csi);
        apply_impulse(iA.inv_mass, iB.inv_mass, direction, ai2vA, ai2vB, csi, vA, vB);
        a.set(0, acc);
    }
};

// ---- DistanceLimit (34): DistanceLimit.cs:L103-181 ----------------------------------------------------------------------------------------------------
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, MinimumDistance, MaximumDistance, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct DistanceLimit {
    static constexpr int kPrestepRows = 10, kImpulseRows = 1;
    static void apply_impulse(const V3<F>& linearJA, const V3<F>& angularJA, const V3<F>& angularJB, const Inertia<F>& iA, const Inertia<F>& iB, const F& csi, Velocity<F>& vA, Velocity<F>& vB) {  // L105-114
        V3<F> impulseScaledLinearJacobian = scale(linearJA, csi);
        vA.lin = add(vA.lin, scale(impulseScaledLinearJacobian, iA.inv_mass));
        vB.lin = sub(vB.lin, scale(impulseScaledLinearJacobian, iB.inv_mass));
        vA.ang = add(vA.ang, transform(scale(angularJA, csi), iA.t));
        vB.ang = add(vB.ang, transform(scale(angularJB, csi), iB.t));
    }
    static void jacobians(const Rows<F>& p, const V3<F>& pA, const Q4<F>& qA, const V3<F>& pB, const Q4<F>& qB, MaskOf<F>& useMinimum, F& distance, V3<F>& direction, V3<F>& angularJA,
                          V3<F>& angularJB) {  // L116-139
        V3<F> offsetA = transform(p.get3(0), qA), offsetB = transform(p.get3(3), qB);
        V3<F> anchorOffset = add(sub(offsetB, offsetA), sub(pB, pA));
        distance = length(anchorOffset);
        useMinimum = lt(vabs(distance - p.get(6)), vabs(distance - p.get(7)));
        F sign = sel(useMinimum, bc<F>(-1.0f), bc<F>(1.0f));
        direction = scale(anchorOffset, F(sign / distance));
        direction = sel3<F>(lt(distance, bc<F>(1e-9f)), v3bc<F>(1.0f, 0.0f, 0.0f), direction);
        angularJA = cross(offsetA, direction);
        angularJB = cross(direction, offsetB);
    }
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        MaskOf<F> useMinimum;
        F distance;
        V3<F> direction, angularJA, angularJB;
        jacobians(p, pA, qA, pB, qB, useMinimum, distance, direction, angularJA, angularJB);
        apply_impulse(direction, angularJA, angularJB, iA, iB, a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        MaskOf<F> useMinimum;
        F distance;
        V3<F> direction, angularJA, angularJB;
        jacobians(p, pA, qA, pB, qB, useMinimum, distance, direction, angularJA, angularJB);
        F linearCSVA = dot(vA.lin, direction), negatedLinearCSVB = dot(vB.lin, direction);
        F angularCSVA = dot(vA.ang, angularJA), angularCSVB = dot(vB.ang, angularJB);
        F csv = linearCSVA - negatedLinearCSVB + angularCSVA + angularCSVB;
        F angularContributionA = vector_sandwich(angularJA, iA.t), angularContributionB = vector_sandwich(angularJB, iB.t);
        F inverseEffectiveMass = iA.inv_mass + iB.inv_mass + angularContributionA + angularContributionB;
        F pe2v, cfm, soft;
        compute_springiness(p.get(8), p.get(9), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / inverseEffectiveMass;
        F error = sel(useMinimum, F(p.get(6) - distance), F(distance - p.get(7)));
        F biasVelocity = inequality_bias_velocity(error, pe2v, inverseDt);
        F acc = a.get(0);
        F csi = -acc * soft - effectiveMass * (csv - biasVelocity);
        clamp_positive(acc, csi);
        apply_impulse(direction, angularJA, angularJB, iA, iB, csi, vA, vB);
        a.set(0, acc);
    }
};

// ---- PointOnLineServo (37): PointOnLineServo.cs:L82-193 -------------------------------------------------------------------------------------------------
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, LocalDirection xyz, MaximumSpeed, BaseSpeed, MaximumForce, AngularFrequency, TwiceDampingRatio. Impulses: xy.
template <class F> struct PointOnLineServo {
    static constexpr int kPrestepRows = 14, kImpulseRows = 2;
    static void apply_impulse(Velocity<F>& vA, Velocity<F>& vB, const M23<F>& linearJacobian, const M23<F>& angularJA, const M23<F>& angularJB, const Inertia<F>& iA, const Inertia<F>& iB,
                              const V2<F>& csi) {  // L84-101
        V3<F> linearImpulseA = transform(csi, linearJacobian);
        V3<F> angularImpulseA = transform(csi, angularJA);
        V3<F> angularImpulseB = transform(csi, angularJB);
        V3<F> angularChangeA = transform(angularImpulseA, iA.t), angularChangeB = transform(angularImpulseB, iB.t);
        V3<F> linearChangeA = scale(linearImpulseA, iA.inv_mass), negatedLinearChangeB = scale(linearImpulseA, iB.inv_mass);
        vA.lin = add(linearChangeA, vA.lin);
        vA.ang = add(angularChangeA, vA.ang);
        vB.lin = sub(vB.lin, negatedLinearChangeB);
        vB.ang = add(angularChangeB, vB.ang);
    }
    static void jacobians(const V3<F>& ab, const Q4<F>& qA, const Q4<F>& qB, const V3<F>& localDirection, const V3<F>& localOffsetA, const V3<F>& localOffsetB, V3<F>& anchorOffset,
                          M23<F>& linearJacobian, M23<F>& angularJA, M23<F>& angularJB) {  // L103-126
        V3<F> localTangentX, localTangentY;
        build_orthonormal_basis(localDirection, localTangentX, localTangentY);
        M33<F> mA = matrix_from_quaternion(qA);
        V3<F> anchorA = transform(localOffsetA, mA);
        V3<F> offsetB = transform(localOffsetB, qB);
        V3<F> direction = transform(localDirection, mA);
        V3<F> anchorB = add(offsetB, ab);
        anchorOffset = sub(anchorB, anchorA);
        F d = dot(anchorOffset, direction);
        V3<F> lineStartToClosestPointOnLine = scale(direction, d);
        V3<F> offsetA = add(lineStartToClosestPointOnLine, anchorA);
        linearJacobian.x = transform(localTangentX, mA);
        linearJacobian.y = transform(localTangentY, mA);
        angularJA.x = cross(offsetA, linearJacobian.x);
        angularJA.y = cross(offsetA, linearJacobian.y);
        angularJB.x = cross(linearJacobian.x, offsetB);
        angularJB.y = cross(linearJacobian.y, offsetB);
    }
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> anchorOffset;
        M23<F> linearJacobian, angularJA, angularJB;
        jacobians(sub(pB, pA), qA, qB, p.get3(6), p.get3(0), p.get3(3), anchorOffset, linearJacobian, angularJA, angularJB);
        apply_impulse(vA, vB, linearJacobian, angularJA, angularJB, iA, iB, a.get2(0));
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> anchorOffset;
        M23<F> linearJacobian, angularJA, angularJB;
        jacobians(sub(pB, pA), qA, qB, p.get3(6), p.get3(0), p.get3(3), anchorOffset, linearJacobian, angularJA, angularJB);
        Sym2<F> linearContribution = sandwich_scale(linearJacobian, F(iA.inv_mass + iB.inv_mass));
        Sym2<F> angularContributionA = matrix_sandwich(angularJA, iA.t), angularContributionB = matrix_sandwich(angularJB, iB.t);
        Sym2<F> inverseEffectiveMass = add(angularContributionA, angularContributionB);
        inverseEffectiveMass = add(inverseEffectiveMass, linearContribution);
        Sym2<F> effectiveMass = invert(inverseEffectiveMass);
        F pe2v, cfm, soft;
        compute_springiness(p.get(12), p.get(13), dt, pe2v, cfm, soft);
        effectiveMass = Sym2<F>{effectiveMass.xx * cfm, effectiveMass.yx * cfm, effectiveMass.yy * cfm};
        V2<F> linearCSVA = transform_by_transpose(vA.lin, linearJacobian), negatedLinearCSVB = transform_by_transpose(vB.lin, linearJacobian);
        V2<F> angularCSVA = transform_by_transpose(vA.ang, angularJA), angularCSVB = transform_by_transpose(vB.ang, angularJB);
        V2<F> linearCSV = sub(linearCSVA, negatedLinearCSVB);
        V2<F> angularCSV = add(angularCSVA, angularCSVB);
        V2<F> csv = add(linearCSV, angularCSV);
        V2<F> error{dot(anchorOffset, linearJacobian.x), dot(anchorOffset, linearJacobian.y)};
        V2<F> biasVelocity;
        F maximumImpulse;
        servo_clamped_bias_velocity2e(error, pe2v, p.get(9), p.get(10), p.get(11), dt, inverseDt, biasVelocity, maximumImpulse);
        csv = sub(biasVelocity, csv);
        V2<F> csi = transform(csv, effectiveMass);
        V2<F> acc = a.get2(0);
        V2<F> softnessContribution = scale(acc, soft);
        csi = sub(csi, softnessContribution);
        servo_clamp_impulse2(maximumImpulse, acc, csi);
        apply_impulse(vA, vB, linearJacobian, angularJA, angularJB, iA, iB, csi);
        a.set2(0, acc);
    }
};

// ---- LinearAxis family: LinearAxisServo.cs:L182-248, LinearAxisMotor.cs:L82-110, LinearAxisLimit.cs:L90-151 --------------------------------------------------
template <class F> inline void linear_axis_apply_impulse(const V3<F>& linearJA, const V3<F>& ai2vA, const V3<F>& ai2vB, const Inertia<F>& iA, const Inertia<F>& iB, const F& csi, Velocity<F>& vA,
                                                        Velocity<F>& vB) {  // LinearAxisServo.cs:L173-180
    vA.lin = add(vA.lin, scale(linearJA, F(csi * iA.inv_mass)));
    vB.lin = sub(vB.lin, scale(linearJA, F(csi * iB.inv_mass)));
    vA.ang = add(vA.ang, scale(ai2vA, csi));
    vB.ang = add(vB.ang, scale(ai2vB, csi));
}
template <class F>
inline void linear_axis_jacobians(const V3<F>& ab, const Q4<F>& qA, const Q4<F>& qB, const V3<F>& localPlaneNormalA, const V3<F>& localOffsetA, const V3<F>& localOffsetB, F& planeNormalDot,
                                  V3<F>& normal, V3<F>& angularJA, V3<F>& angularJB) {  // LinearAxisServo.cs:L182-197
    M33<F> mA = matrix_from_quaternion(qA);
    normal = transform(localPlaneNormalA, mA);
    V3<F> anchorA = transform(localOffsetA, mA);
    V3<F> offsetB = transform(localOffsetB, qB);
    V3<F> anchorB = add(ab, offsetB);
    planeNormalDot = dot(sub(anchorB, anchorA), normal);
    V3<F> offsetFromAToClosestPointOnPlaneToB = sub(anchorB, scale(normal, planeNormalDot));
    angularJA = cross(offsetFromAToClosestPointOnPlaneToB, normal);
    angularJB = cross(normal, offsetB);
}
template <class F>
inline void linear_axis_effective_mass(const V3<F>& angularJA, const V3<F>& angularJB, const Inertia<F>& iA, const Inertia<F>& iB, const F& cfm, V3<F>& ai2vA, V3<F>& ai2vB, F& effectiveMass) {  // L199-209
    ai2vA = transform(angularJA, iA.t);
    ai2vB = transform(angularJB, iB.t);
    F angularContributionA = dot(angularJA, ai2vA), angularContributionB = dot(angularJB, ai2vB);
    effectiveMass = cfm / (iA.inv_mass + iB.inv_mass + angularContributionA + angularContributionB);
}
template <class F> inline F linear_axis_csv(const Velocity<F>& vA, const Velocity<F>& vB, const V3<F>& normal, const V3<F>& angularJA, const V3<F>& angularJB) {
    return dot(sub(vA.lin, vB.lin), normal) + dot(vA.ang, angularJA) + dot(vB.ang, angularJB);
}
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, LocalPlaneNormal xyz, TargetOffset, MaximumSpeed, BaseSpeed, MaximumForce, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct LinearAxisServo {
    static constexpr int kPrestepRows = 15, kImpulseRows = 1;
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F planeNormalDot;
        V3<F> normal, angularJA, angularJB;
        linear_axis_jacobians(sub(pB, pA), qA, qB, p.get3(6), p.get3(0), p.get3(3), planeNormalDot, normal, angularJA, angularJB);
        linear_axis_apply_impulse(normal, transform(angularJA, iA.t), transform(angularJB, iB.t), iA, iB, a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F planeNormalDot;
        V3<F> normal, angularJA, angularJB;
        linear_axis_jacobians(sub(pB, pA), qA, qB, p.get3(6), p.get3(0), p.get3(3), planeNormalDot, normal, angularJA, angularJB);
        F pe2v, cfm, soft;
        compute_springiness(p.get(13), p.get(14), dt, pe2v, cfm, soft);
        V3<F> ai2vA, ai2vB;
        F effectiveMass;
        linear_axis_effective_mass(angularJA, angularJB, iA, iB, cfm, ai2vA, ai2vB, effectiveMass);
        F biasVelocity, maximumImpulse;
        servo_clamped_bias_velocity(F(planeNormalDot - p.get(9)), pe2v, p.get(10), p.get(11), p.get(12), dt, inverseDt, biasVelocity, maximumImpulse);
        F csv = linear_axis_csv(vA, vB, normal, angularJA, angularJB);
        F acc = a.get(0);
        F csi = effectiveMass * (biasVelocity - csv) - acc * soft;
        servo_clamp_impulse(maximumImpulse, acc, csi);
        linear_axis_apply_impulse(normal, ai2vA, ai2vB, iA, iB, csi, vA, vB);
        a.set(0, acc);
    }
};
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, LocalPlaneNormal xyz, TargetVelocity, MaximumForce, Damping. Impulse: 1.
template <class F> struct LinearAxisMotor {
    static constexpr int kPrestepRows = 12, kImpulseRows = 1;
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        LinearAxisServo<F>::warm_start(pA, qA, iA, pB, qB, iB, p, a, vA, vB);  // identical body: LinearAxisMotor.cs:L84-90
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F planeNormalDot;
        V3<F> normal, angularJA, angularJB;
        linear_axis_jacobians(sub(pB, pA), qA, qB, p.get3(6), p.get3(0), p.get3(3), planeNormalDot, normal, angularJA, angularJB);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(10), p.get(11), dt, cfm, soft, maximumImpulse);
        V3<F> ai2vA, ai2vB;
        F effectiveMass;
        linear_axis_effective_mass(angularJA, angularJB, iA, iB, cfm, ai2vA, ai2vB, effectiveMass);
        F csv = linear_axis_csv(vA, vB, normal, angularJA, angularJB);
        F acc = a.get(0);
        F csi = effectiveMass * (-p.get(9) - csv) - acc * soft;
        servo_clamp_impulse(maximumImpulse, acc, csi);
        linear_axis_apply_impulse(normal, ai2vA, ai2vB, iA, iB, csi, vA, vB);
        a.set(0, acc);
    }
};
// Prestep rows: LocalOffsetA xyz, LocalOffsetB xyz, LocalPlaneNormal xyz, MinimumOffset, MaximumOffset, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct LinearAxisLimit {
    static constexpr int kPrestepRows = 13, kImpulseRows = 1;
    static void jacobians(const V3<F>& ab, const Q4<F>& qA, const Q4<F>& qB, const Rows<F>& p, F& error, V3<F>& normal, V3<F>& angularJA, V3<F>& angularJB) {  // LinearAxisLimit.cs:L92-119
        M33<F> mA = matrix_from_quaternion(qA);
        normal = transform(p.get3(6), mA);
        V3<F> anchorA = transform(p.get3(0), mA);
        V3<F> offsetB = transform(p.get3(3), qB);
        V3<F> anchorB = add(ab, offsetB);
        F planeNormalDot = dot(sub(anchorB, anchorA), normal);
        F minimumError = p.get(9) - planeNormalDot;
        F maximumError = planeNormalDot - p.get(10);
        MaskOf<F> useMin = lt(vabs(minimumError), vabs(maximumError));
        error = sel(useMin, minimumError, maximumError);
        normal = sel3<F>(useMin, neg(normal), normal);
        V3<F> offsetFromAToClosestPointOnPlaneToB = sub(anchorB, scale(normal, planeNormalDot));
        angularJA = cross(offsetFromAToClosestPointOnPlaneToB, normal);
        angularJB = cross(normal, offsetB);
    }
    static void warm_start(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F error;
        V3<F> normal, angularJA, angularJB;
        jacobians(sub(pB, pA), qA, qB, p, error, normal, angularJA, angularJB);
        linear_axis_apply_impulse(normal, transform(angularJA, iA.t), transform(angularJB, iB.t), iA, iB, a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>& qB, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        F error;
        V3<F> normal, angularJA, angularJB;
        jacobians(sub(pB, pA), qA, qB, p, error, normal, angularJA, angularJB);
        F pe2v, cfm, soft;
        compute_springiness(p.get(11), p.get(12), dt, pe2v, cfm, soft);
        V3<F> ai2vA, ai2vB;
        F effectiveMass;
        linear_axis_effective_mass(angularJA, angularJB, iA, iB, cfm, ai2vA, ai2vB, effectiveMass);
        F biasVelocity = inequality_bias_velocity(error, pe2v, inverseDt);
        F csv = linear_axis_csv(vA, vB, normal, angularJA, angularJB);
        F acc = a.get(0);
        F csi = effectiveMass * (biasVelocity - csv) - acc * soft;
        clamp_positive(acc, csi);
        linear_axis_apply_impulse(normal, ai2vA, ai2vB, iA, iB, csi, vA, vB);
        a.set(0, acc);
    }
};

// ---- CenterDistanceConstraint (35) / CenterDistanceLimit (55): CenterDistanceConstraint.cs:L69-133, CenterDistanceLimit.cs:L78-132 --------------------------
template <class F> inline void center_distance_apply_impulse(const V3<F>& jacobianA, const F& imA, const F& imB, const F& impulse, Velocity<F>& a, Velocity<F>& b) {  // L71-80
    a.lin = add(a.lin, scale(jacobianA, F(impulse * imA)));
    b.lin = sub(b.lin, scale(jacobianA, F(impulse * imB)));
}
// Prestep rows: TargetDistance, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct CenterDistanceConstraint {
    static constexpr int kPrestepRows = 3, kImpulseRows = 1;
    static void warm_start(const V3<F>& pA, const Q4<F>&, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>&, const Inertia<F>& iB, const Rows<F>&, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> ab = sub(pB, pA);
        F lengthSquared = length_squared(ab);
        F inverseDistance = bc<F>(1.0f) / vsqrt(lengthSquared);  // FastReciprocalSquareRoot, exact form (see header)
        V3<F> jacobianA = scale(ab, inverseDistance);
        jacobianA = sel3<F>(lt(lengthSquared, bc<F>(1e-10f)), v3bc<F>(1.0f, 0.0f, 0.0f), jacobianA);
        center_distance_apply_impulse(jacobianA, iA.inv_mass, iB.inv_mass, a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>&, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>&, const Inertia<F>& iB, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> ab = sub(pB, pA);
        F distance = length(ab);
        F inverseDistance = bc<F>(1.0f) / distance;  // FastReciprocal, exact form
        V3<F> jacobianA = scale(ab, inverseDistance);
        jacobianA = sel3<F>(lt(distance, bc<F>(1e-5f)), v3bc<F>(1.0f, 0.0f, 0.0f), jacobianA);
        F pe2v, cfm, soft;
        compute_springiness(p.get(1), p.get(2), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / (iA.inv_mass + iB.inv_mass);
        F biasVelocity = (distance - p.get(0)) * pe2v;
        F linearCSVA = dot(vA.lin, jacobianA), negatedCSVB = dot(vB.lin, jacobianA);
        F acc = a.get(0);
        F csi = (biasVelocity - (linearCSVA - negatedCSVB)) * effectiveMass - acc * soft;
        acc = acc + csi;
        center_distance_apply_impulse(jacobianA, iA.inv_mass, iB.inv_mass, csi, vA, vB);
        a.set(0, acc);
    }
};
// Prestep rows: MinimumDistance, MaximumDistance, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct CenterDistanceLimit {
    static constexpr int kPrestepRows = 4, kImpulseRows = 1;
    static void jacobian(const F& minimumDistance, const F& maximumDistance, const V3<F>& pA, const V3<F>& pB, V3<F>& jacobianA, F& distance, MaskOf<F>& useMinimum) {  // L80-96
        V3<F> ab = sub(pB, pA);
        distance = length(ab);
        F inverseDistance = bc<F>(1.0f) / distance;  // FastReciprocal, exact form
        jacobianA = scale(ab, inverseDistance);
        jacobianA = sel3<F>(lt(distance, bc<F>(1e-5f)), v3bc<F>(1.0f, 0.0f, 0.0f), jacobianA);
        useMinimum = lt(vabs(distance - minimumDistance), vabs(distance - maximumDistance));
        jacobianA = sel3<F>(useMinimum, neg(jacobianA), jacobianA);
    }
    static void warm_start(const V3<F>& pA, const Q4<F>&, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>&, const Inertia<F>& iB, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> jacobianA;
        F distance;
        MaskOf<F> useMinimum;
        jacobian(p.get(0), p.get(1), pA, pB, jacobianA, distance, useMinimum);
        center_distance_apply_impulse(jacobianA, iA.inv_mass, iB.inv_mass, a.get(0), vA, vB);
    }
    static void solve(const V3<F>& pA, const Q4<F>&, const Inertia<F>& iA, const V3<F>& pB, const Q4<F>&, const Inertia<F>& iB, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA, Velocity<F>& vB) {
        V3<F> jacobianA;
        F distance;
        MaskOf<F> useMinimum;
        jacobian(p.get(0), p.get(1), pA, pB, jacobianA, distance, useMinimum);
        F pe2v, cfm, soft;
        compute_springiness(p.get(2), p.get(3), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / (iA.inv_mass + iB.inv_mass);
        F error = sel(useMinimum, F(p.get(0) - distance), F(distance - p.get(1)));
        F biasVelocity = inequality_bias_velocity(error, pe2v, inverseDt);
        F csv = dot(vA.lin, jacobianA) - dot(vB.lin, jacobianA);
        F acc = a.get(0);
        F csi = -acc * soft - effectiveMass * (csv - biasVelocity);
        clamp_positive(acc, csi);
        center_distance_apply_impulse(jacobianA, iA.inv_mass, iB.inv_mass, csi, vA, vB);
        a.set(0, acc);
    }
};

// ---- one-body joints: signature (position, orientation, inertia, ..., velocity) ---------------------------------------------------------------------------
// OneBodyAngularServo (42): OneBodyAngularServo.cs:L69-109. Prestep rows: TargetOrientation xyzw, AngularFrequency, TwiceDampingRatio, MaximumSpeed, BaseSpeed, MaximumForce. Impulses: xyz.
template <class F> struct OneBodyAngularServo {
    static constexpr int kPrestepRows = 9, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>&, const Inertia<F>& iA, const Rows<F>&, const Rows<F>& a, Velocity<F>& vA) { vA.ang = add(vA.ang, transform(a.get3(0), iA.t)); }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        Q4<F> inverseOrientation = conjugate(qA);
        Q4<F> errorRotation = concatenate(inverseOrientation, p.get4(0));
        V3<F> errorAxis;
        F errorLength;
        axis_angle_from_quaternion(errorRotation, errorAxis, errorLength);
        F pe2v, cfm, soft;
        compute_springiness(p.get(4), p.get(5), dt, pe2v, cfm, soft);
        Sym3<F> effectiveMass = invert(iA.t);
        V3<F> clampedBiasVelocity;
        F maximumImpulse;
        servo_clamped_bias_velocity3(errorAxis, errorLength, pe2v, p.get(6), p.get(7), p.get(8), dt, inverseDt, clampedBiasVelocity, maximumImpulse);
        V3<F> csv = sub(clampedBiasVelocity, vA.ang);
        V3<F> csi = transform(csv, effectiveMass);
        V3<F> acc = a.get3(0);
        csi = sub(scale(csi, cfm), scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        vA.ang = add(vA.ang, transform(csi, iA.t));
        a.set3(0, acc);
    }
};
// OneBodyAngularMotor (43): OneBodyAngularMotor.cs:L61-93. Prestep rows: TargetVelocity xyz, MaximumForce, Damping. Impulses: xyz.
template <class F> struct OneBodyAngularMotor {
    static constexpr int kPrestepRows = 5, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>&, const Inertia<F>& iA, const Rows<F>&, const Rows<F>& a, Velocity<F>& vA) { vA.ang = add(vA.ang, transform(a.get3(0), iA.t)); }
    static void solve(const V3<F>&, const Q4<F>&, const Inertia<F>& iA, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(3), p.get(4), dt, cfm, soft, maximumImpulse);
        Sym3<F> unsoftenedEffectiveMass = invert(iA.t);
        V3<F> csi = transform(sub(p.get3(0), vA.ang), unsoftenedEffectiveMass);
        V3<F> acc = a.get3(0);
        csi = sub(scale(csi, cfm), scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        vA.ang = add(vA.ang, transform(csi, iA.t));
        a.set3(0, acc);
    }
};
template <class F> inline void one_body_linear_apply_impulse(const V3<F>& offset, const Inertia<F>& inertia, Velocity<F>& vA, const V3<F>& csi) {  // OneBodyLinearServo.cs:L93-105
    V3<F> wsi = cross(offset, csi);
    vA.ang = add(vA.ang, transform(wsi, inertia.t));
    vA.lin = add(vA.lin, scale(csi, inertia.inv_mass));
}
template <class F> inline Sym3<F> one_body_linear_effective_mass(const V3<F>& offset, const Inertia<F>& inertia) {
    Sym3<F> inverseEffectiveMass = skew_sandwich(offset, inertia.t);
    inverseEffectiveMass.xx = inverseEffectiveMass.xx + inertia.inv_mass;
    inverseEffectiveMass.yy = inverseEffectiveMass.yy + inertia.inv_mass;
    inverseEffectiveMass.zz = inverseEffectiveMass.zz + inertia.inv_mass;
    return invert(inverseEffectiveMass);
}
// OneBodyLinearServo (44): OneBodyLinearServo.cs:L77-146. Prestep rows: LocalOffset xyz, Target xyz, AngularFrequency, TwiceDampingRatio, MaximumSpeed, BaseSpeed, MaximumForce. Impulses: xyz.
template <class F> struct OneBodyLinearServo {
    static constexpr int kPrestepRows = 11, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        one_body_linear_apply_impulse(transform(p.get3(0), qA), iA, vA, a.get3(0));
    }
    static void solve(const V3<F>& pA, const Q4<F>& qA, const Inertia<F>& iA, float dt, float inverseDt, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        V3<F> offset = transform(p.get3(0), qA);
        F pe2v, cfm, soft;
        compute_springiness(p.get(6), p.get(7), dt, pe2v, cfm, soft);
        V3<F> worldGrabPoint = add(offset, pA);
        V3<F> error = sub(p.get3(3), worldGrabPoint);
        V3<F> biasVelocity;
        F maximumImpulse;
        servo_clamped_bias_velocity3e(error, pe2v, p.get(8), p.get(9), p.get(10), dt, inverseDt, biasVelocity, maximumImpulse);
        V3<F> csv = sub(sub(biasVelocity, cross(vA.ang, offset)), vA.lin);
        Sym3<F> effectiveMass = one_body_linear_effective_mass(offset, iA);
        V3<F> csi = transform(csv, effectiveMass);
        V3<F> acc = a.get3(0);
        csi = sub(scale(csi, cfm), scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        one_body_linear_apply_impulse(offset, iA, vA, csi);
        a.set3(0, acc);
    }
};
// OneBodyLinearMotor (45): OneBodyLinearMotor.cs:L67-100. Prestep rows: LocalOffset xyz, TargetVelocity xyz, MaximumForce, Damping. Impulses: xyz.
template <class F> struct OneBodyLinearMotor {
    static constexpr int kPrestepRows = 8, kImpulseRows = 3;
    static void warm_start(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        one_body_linear_apply_impulse(transform(p.get3(0), qA), iA, vA, a.get3(0));
    }
    static void solve(const V3<F>&, const Q4<F>& qA, const Inertia<F>& iA, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>& vA) {
        V3<F> offset = transform(p.get3(0), qA);
        F cfm, soft, maximumImpulse;
        motor_softness(p.get(6), p.get(7), dt, cfm, soft, maximumImpulse);
        V3<F> csv = sub(sub(p.get3(3), cross(vA.ang, offset)), vA.lin);
        Sym3<F> effectiveMass = one_body_linear_effective_mass(offset, iA);
        V3<F> csi = transform(csv, effectiveMass);
        V3<F> acc = a.get3(0);
        csi = sub(scale(csi, cfm), scale(acc, soft));
        servo_clamp_impulse3(maximumImpulse, acc, csi);
        one_body_linear_apply_impulse(offset, iA, vA, csi);
        a.set3(0, acc);
    }
};

// ---- three / four body constraints: signature (positions[], inverse masses[], ..., velocities[]) -----------------------------------------------------------
// AreaConstraint (36): AreaConstraint.cs:L76-197. Prestep rows: TargetScaledArea, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct AreaConstraint {
    static constexpr int kBodies = 3, kPrestepRows = 3, kImpulseRows = 1;
    struct Jacobian { F normalLength, contributionA, contributionB, contributionC, inverseJacobianLength; V3<F> negatedJacobianA, jacobianB, jacobianC; };
    static void apply_impulse(const F* im, const Jacobian& j, const F& impulse, Velocity<F>* v) {  // L78-91
        V3<F> negativeVelocityChangeA = scale(j.negatedJacobianA, F(im[0] * impulse));
        V3<F> velocityChangeB = scale(j.jacobianB, F(im[1] * impulse));
        V3<F> velocityChangeC = scale(j.jacobianC, F(im[2] * impulse));
        v[0].lin = sub(v[0].lin, negativeVelocityChangeA);
        v[1].lin = add(v[1].lin, velocityChangeB);
        v[2].lin = add(v[2].lin, velocityChangeC);
    }
    static Jacobian jacobian(const V3<F>* pos) {  // L93-138
        Jacobian j;
        V3<F> ab = sub(pos[1], pos[0]), ac = sub(pos[2], pos[0]);
        V3<F> abxac = cross(ab, ac);
        j.normalLength = length(abxac);
        V3<F> normal = scale(abxac, F(sel(gt(j.normalLength, bc<F>(1e-10f)), F(bc<F>(1.0f) / j.normalLength), bc<F>(0.0f))));
        j.jacobianB = cross(ac, normal);
        j.jacobianC = cross(normal, ab);
        j.negatedJacobianA = add(j.jacobianB, j.jacobianC);
        j.contributionA = dot(j.negatedJacobianA, j.negatedJacobianA);
        j.contributionB = dot(j.jacobianB, j.jacobianB);
        j.contributionC = dot(j.jacobianC, j.jacobianC);
        F jacobianLengthSquared = j.contributionA + j.contributionB + j.contributionC;
        jacobianLengthSquared = vmax(bc<F>(1e-14f), jacobianLengthSquared);
        j.inverseJacobianLength = bc<F>(1.0f) / vsqrt(jacobianLengthSquared);  // FastReciprocalSquareRoot, exact form
        return j;
    }
    static void warm_start(const V3<F>* pos, const F* im, const Rows<F>&, const Rows<F>& a, Velocity<F>* v) {
        Jacobian j = jacobian(pos);
        apply_impulse(im, j, F(j.inverseJacobianLength * a.get(0)), v);
    }
    static void solve(const V3<F>* pos, const F* im, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>* v) {
        Jacobian j = jacobian(pos);
        F inverseJacobianLengthSquared = j.inverseJacobianLength * j.inverseJacobianLength;
        F inverseEffectiveMass = vmax(bc<F>(1e-14f), F(inverseJacobianLengthSquared * (j.contributionA * im[0] + j.contributionB * im[1] + j.contributionC * im[2])));
        F pe2v, cfm, soft;
        compute_springiness(p.get(1), p.get(2), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / inverseEffectiveMass;
        F biasVelocity = (p.get(0) - j.normalLength) * j.inverseJacobianLength * pe2v;
        F negatedVelocityContributionA = dot(j.negatedJacobianA, v[0].lin);
        F velocityContributionB = dot(j.jacobianB, v[1].lin);
        F velocityContributionC = dot(j.jacobianC, v[2].lin);
        F csv = j.inverseJacobianLength * (velocityContributionB + velocityContributionC - negatedVelocityContributionA);
        F acc = a.get(0);
        F csi = (biasVelocity - csv) * effectiveMass - acc * soft;
        acc = acc + csi;
        apply_impulse(im, j, F(j.inverseJacobianLength * csi), v);
        a.set(0, acc);
    }
};
// VolumeConstraint (32): VolumeConstraint.cs:L76-186. Prestep rows: TargetScaledVolume, AngularFrequency, TwiceDampingRatio. Impulse: 1.
template <class F> struct VolumeConstraint {
    static constexpr int kBodies = 4, kPrestepRows = 3, kImpulseRows = 1;
    struct Jacobian { F contributionA, contributionB, contributionC, contributionD, inverseJacobianLength; V3<F> ad, negatedJA, jacobianB, jacobianC, jacobianD; };
    static void apply_impulse(const F* im, const Jacobian& j, const F& impulse, Velocity<F>* v) {  // L78-94
        V3<F> negativeVelocityChangeA = scale(j.negatedJA, F(im[0] * impulse));
        V3<F> velocityChangeB = scale(j.jacobianB, F(im[1] * impulse));
        V3<F> velocityChangeC = scale(j.jacobianC, F(im[2] * impulse));
        V3<F> velocityChangeD = scale(j.jacobianD, F(im[3] * impulse));
        v[0].lin = sub(v[0].lin, negativeVelocityChangeA);
        v[1].lin = add(v[1].lin, velocityChangeB);
        v[2].lin = add(v[2].lin, velocityChangeC);
        v[3].lin = add(v[3].lin, velocityChangeD);
    }
    static Jacobian jacobian(const V3<F>* pos) {  // L96-123
        Jacobian j;
        V3<F> ab = sub(pos[1], pos[0]), ac = sub(pos[2], pos[0]);
        j.ad = sub(pos[3], pos[0]);
        j.jacobianB = cross(ac, j.ad);
        j.jacobianC = cross(j.ad, ab);
        j.jacobianD = cross(ab, ac);
        j.negatedJA = add(j.jacobianB, j.jacobianC);
        j.negatedJA = add(j.jacobianD, j.negatedJA);
        j.contributionA = dot(j.negatedJA, j.negatedJA);
        j.contributionB = dot(j.jacobianB, j.jacobianB);
        j.contributionC = dot(j.jacobianC, j.jacobianC);
        j.contributionD = dot(j.jacobianD, j.jacobianD);
        F jacobianLengthSquared = j.contributionA + j.contributionB + j.contributionC + j.contributionD;
        jacobianLengthSquared = vmax(bc<F>(1e-14f), jacobianLengthSquared);
        j.inverseJacobianLength = bc<F>(1.0f) / vsqrt(jacobianLengthSquared);  // FastReciprocalSquareRoot, exact form
        return j;
    }
    static void warm_start(const V3<F>* pos, const F* im, const Rows<F>&, const Rows<F>& a, Velocity<F>* v) {
        Jacobian j = jacobian(pos);
        apply_impulse(im, j, F(j.inverseJacobianLength * a.get(0)), v);
    }
    static void solve(const V3<F>* pos, const F* im, float dt, float, const Rows<F>& p, const Rows<F>& a, Velocity<F>* v) {
        Jacobian j = jacobian(pos);
        F inverseJacobianLengthSquared = j.inverseJacobianLength * j.inverseJacobianLength;
        F inverseEffectiveMass =
            vmax(bc<F>(1e-14f), F(inverseJacobianLengthSquared * (j.contributionA * im[0] + j.contributionB * im[1] + j.contributionC * im[2] + j.contributionD * im[3])));
        F pe2v, cfm, soft;
        compute_springiness(p.get(1), p.get(2), dt, pe2v, cfm, soft);
        F effectiveMass = cfm / inverseEffectiveMass;
        F volume = dot(j.jacobianD, j.ad);
        F biasVelocity = (p.get(0) - volume) * j.inverseJacobianLength * pe2v;
        F negatedVelocityContributionA = dot(j.negatedJA, v[0].lin);
        F velocityContributionB = dot(j.jacobianB, v[1].lin);
        F velocityContributionC = dot(j.jacobianC, v[2].lin);
        F velocityContributionD = dot(j.jacobianD, v[3].lin);
        F csv = j.inverseJacobianLength * (velocityContributionB + velocityContributionC + velocityContributionD - negatedVelocityContributionA);
        F acc = a.get(0);
        F csi = (biasVelocity - csv) * effectiveMass - acc * soft;
        acc = acc + csi;
        apply_impulse(im, j, F(j.inverseJacobianLength * csi), v);
        a.set(0, acc);
    }
};

template <class R> inline void register_joints_more(R& r) {
    typedef typename R::Lane F;
    r.template joint2<AngularHinge<F>>(23);
    r.template joint2<AngularSwivelHinge<F>>(24);
    r.template joint2<TwistMotor<F>>(28);
    r.template joint2<Weld<F>>(31);
    r.template jointN<VolumeConstraint<F>>(32);
    r.template joint2<DistanceServo<F>>(33);
    r.template joint2<DistanceLimit<F>>(34);
    r.template joint2<CenterDistanceConstraint<F>>(35);
    r.template jointN<AreaConstraint<F>>(36);
    r.template joint2<PointOnLineServo<F>>(37);
    r.template joint2<LinearAxisServo<F>>(38);
    r.template joint2<LinearAxisMotor<F>>(39);
    r.template joint2<LinearAxisLimit<F>>(40);
    r.template joint2<AngularAxisMotor<F>>(41);
    r.template joint1<OneBodyAngularServo<F>>(42);
    r.template joint1<OneBodyAngularMotor<F>>(43);
    r.template joint1<OneBodyLinearServo<F>>(44);
    r.template joint1<OneBodyLinearMotor<F>>(45);
    r.template joint2<BallSocketMotor<F>>(52);
    r.template joint2<BallSocketServo<F>>(53);
    r.template joint2<AngularAxisGearMotor<F>>(54);
    r.template joint2<CenterDistanceLimit<F>>(55);
}

}  // namespace bepu_oracle
