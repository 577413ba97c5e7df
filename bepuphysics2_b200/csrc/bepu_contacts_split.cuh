// EXPERIMENT ("two lanes per two-body constraint"): only compiled into the kernels with -DBEPU_SPLIT_CONTACTS (off in the shipped
// library); the arithmetic is checked on the CPU by tests/test_device_source_on_host.py.
//
// A two-body contact manifold evaluated by a PAIR of lanes: lane `side` 0 owns body A, lane 1 owns body B. Each lane gathers, updates and
// scatters only its own body and computes only its own body's jacobian terms; the few scalars the other side needs cross through
// X::swap(v) (device: __shfl_xor_sync(0xffffffff, v, 1); the host test uses a two-thread rendezvous). Every floating-point expression is
// assembled in the operand order of the one-lane functions in csrc/bepu_contacts.cuh (which follow ContactConvexTypes.cs, PenetrationLimit.cs,
// TangentFriction.cs, TwistFriction.cs), so both lanes hold bit-identical impulses and the result is bit-identical to the one-lane evaluation
// (tests/test_device_source_on_host.py, split-lane test).
// The dependent chain per lane roughly halves (one cross product, one inertia sandwich, two dot products and one body update per row instead
// of two each), for 3 exchanges per penetration row, 10 for the tangent friction, 2 for the twist friction and none in WarmStart.
#pragma once
#include "bepu_contacts.cuh"

namespace BEPU_NS {

template <class X> BEPU_DI Sym2 swap_sym2(Sym2 m) { return {X::swap(m.xx), X::swap(m.yx), X::swap(m.yy)}; }
template <class X> BEPU_DI V2 swap_v2(V2 v) { return {X::swap(v.x), X::swap(v.y)}; }

// PenetrationLimit.Solve (PenetrationLimit.cs:L78-131), own side only. offsetMine is the contact offset from the own body's centre.
template <class X>
BEPU_DI void split_penetration_solve(bool isA, const Inertia& mine, float linear, V3 normal, V3 offsetA, V3 offsetB, float depth, const Springiness& sp, float maxRecovery,
                                     float inverseDt, float& accumulated, Velocity& v) {
    const V3 angular = isA ? cross(offsetA, normal) : cross(normal, offsetB);
    const float sandwich = vector_sandwich(angular, mine.t);
    const float csvLinear = dot(v.lin, normal);  // body B: this is negatedCSVBLinear
    const float csvAngular = dot(v.ang, angular);
    const float sandwichOther = X::swap(sandwich), csvLinearOther = X::swap(csvLinear), csvAngularOther = X::swap(csvAngular);
    const float angularA0 = isA ? sandwich : sandwichOther, angularB0 = isA ? sandwichOther : sandwich;
    const float csvaLinear = isA ? csvLinear : csvLinearOther, negatedCSVBLinear = isA ? csvLinearOther : csvLinear;
    const float csvaAngular = isA ? csvAngular : csvAngularOther, csvbAngular = isA ? csvAngularOther : csvAngular;
    float effectiveMass = sp.effective_mass_cfm_scale / (linear + angularA0 + angularB0);
    float biasVelocity = fmin_ps(depth * inverseDt, fmin_ps(depth * sp.position_error_to_velocity, maxRecovery));
    float negatedCSI = accumulated * sp.softness_impulse_scale + (csvaLinear - negatedCSVBLinear + csvaAngular + csvbAngular - biasVelocity) * effectiveMass;
    float previous = accumulated;
    accumulated = fmax_ps(0.0f, accumulated - negatedCSI);
    const float impulse = accumulated - previous;
    // PenetrationLimit.ApplyImpulse (L45-65), own body
    const V3 dLin = normal * (impulse * mine.inv_mass);
    v.lin = isA ? v.lin + dLin : v.lin - dLin;
    v.ang = v.ang + transform(angular * impulse, mine.t);
}
BEPU_DI void split_penetration_warm_start(bool isA, const Inertia& mine, V3 normal, V3 offsetA, V3 offsetB, float accumulated, Velocity& v) {
    const V3 angular = isA ? cross(offsetA, normal) : cross(normal, offsetB);
    const V3 dLin = normal * (accumulated * mine.inv_mass);
    v.lin = isA ? v.lin + dLin : v.lin - dLin;
    v.ang = v.ang + transform(angular * accumulated, mine.t);
}

// TangentFriction (TangentFriction.cs:L17-100), own side only.
BEPU_DI M23 split_tangent_angular(bool isA, V3 tX, V3 tY, V3 offsetA, V3 offsetB) {
    return isA ? M23{cross(offsetA, tX), cross(offsetA, tY)} : M23{cross(tX, offsetB), cross(tY, offsetB)};
}
BEPU_DI void split_tangent_apply(bool isA, const M23& linearA, const M23& angularMine, const Inertia& mine, V2 impulse, Velocity& v) {
    const V3 linearImpulseA = transform(impulse, linearA);
    const V3 dLin = linearImpulseA * mine.inv_mass;
    v.lin = isA ? v.lin + dLin : v.lin - dLin;
    v.ang = v.ang + transform(transform(impulse, angularMine), mine.t);
}
template <class X>
BEPU_DI void split_tangent_solve(bool isA, V3 tX, V3 tY, V3 offsetA, V3 offsetB, const Inertia& mine, float maximumImpulse, V2& accumulated, Velocity& v) {
    const M23 linearA{tX, tY};
    const M23 angularMine = split_tangent_angular(isA, tX, tY, offsetA, offsetB);
    const Sym2 linearMine = sandwich_scale(linearA, mine.inv_mass);
    const Sym2 angularSandwichMine = matrix_sandwich(angularMine, mine.t);
    const V2 csvLinearMine = transform_by_transpose(v.lin, linearA);
    const V2 csvAngularMine = transform_by_transpose(v.ang, angularMine);
    const Sym2 linearOther = swap_sym2<X>(linearMine), angularSandwichOther = swap_sym2<X>(angularSandwichMine);
    const V2 csvLinearOther = swap_v2<X>(csvLinearMine), csvAngularOther = swap_v2<X>(csvAngularMine);
    Sym2 linear = (isA ? linearMine : linearOther) + (isA ? linearOther : linearMine);
    Sym2 angular = (isA ? angularSandwichMine : angularSandwichOther) + (isA ? angularSandwichOther : angularSandwichMine);
    Sym2 effectiveMass = invert(linear + angular);
    const V2 csvaLinear = isA ? csvLinearMine : csvLinearOther, csvbLinear = isA ? csvLinearOther : csvLinearMine;
    const V2 csvaAngular = isA ? csvAngularMine : csvAngularOther, csvbAngular = isA ? csvAngularOther : csvAngularMine;
    V2 csv = (csvbLinear - csvaLinear) - (csvaAngular + csvbAngular);
    V2 csi = transform(csv, effectiveMass);
    V2 previous = accumulated;
    accumulated = accumulated + csi;
    float magnitude = length(accumulated);
    float scl = fmin_ps(1.0f, maximumImpulse / fmax_ps(1e-16f, magnitude));
    accumulated = accumulated * scl;
    split_tangent_apply(isA, linearA, angularMine, mine, accumulated - previous, v);
}

// TwistFriction (TwistFriction.cs:L322-343), own side only.
BEPU_DI void split_twist_apply(bool isA, V3 axis, const Inertia& mine, float impulse, Velocity& v) {
    const V3 d = transform(axis * impulse, mine.t);
    v.ang = isA ? v.ang + d : v.ang - d;
}
template <class X> BEPU_DI void split_twist_solve(bool isA, V3 axis, const Inertia& mine, float maximumImpulse, float& accumulated, Velocity& v) {
    const float sandwich = vector_sandwich(axis, mine.t), csv = dot(v.ang, axis);
    const float sandwichOther = X::swap(sandwich), csvOther = X::swap(csv);
    float inverseEffectiveMass = (isA ? sandwich : sandwichOther) + (isA ? sandwichOther : sandwich);
    float effectiveMass = (0.0f == inverseEffectiveMass) ? 0.0f : 1.0f / inverseEffectiveMass;
    float negatedCSI = ((isA ? csv : csvOther) - (isA ? csvOther : csv)) * effectiveMass;
    float previous = accumulated;
    accumulated = fmin_ps(maximumImpulse, fmax_ps(-maximumImpulse, accumulated - negatedCSI));
    split_twist_apply(isA, axis, mine, accumulated - previous, v);
}

// Contact{N} (two bodies) by a lane pair: same rows, same order of sub-solves as ConvexTwoBody<N> (ContactConvexTypes.cs, e.g. L1473-1513).
template <int N> struct ConvexTwoBodySplit {
    typedef ConvexLayout<N, true> L;
    template <class PR, class AR> BEPU_DI static void warm_start(int side, const Inertia& mine, PR p, AR a, Velocity& v) {
        const bool isA = side == 0;
        V3 normal = ldrow3(p, L::kNormal), offsetB = ldrow3(p, L::kOffsetB);
        V3 x, z;
        build_orthonormal_basis(normal, x, z);
        V3 offs[N];
        float depths[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { offs[i] = ldrow3(p, 4 * i); depths[i] = ldrow(p, 4 * i + 3); }
        V3 centerA;
        if constexpr (N == 1) centerA = offs[0]; else centerA = friction_center<N>(offs, depths);
        V3 centerB = centerA - offsetB;
        split_tangent_apply(isA, M23{x, z}, split_tangent_angular(isA, x, z, centerA, centerB), mine, V2{ldacc(a, 0), ldacc(a, 1)}, v);
#pragma unroll
        for (int i = 0; i < N; ++i) split_penetration_warm_start(isA, mine, normal, offs[i], offs[i] - offsetB, ldacc(a, 2 + i), v);
        split_twist_apply(isA, normal, mine, ldacc(a, N + 2), v);
    }
    template <class X, class PR, class AR> BEPU_DI static void solve(int side, const Inertia& mine, float dt, float inverseDt, PR p, AR a, Velocity& v) {
        const bool isA = side == 0;
        V3 normal = ldrow3(p, L::kNormal), offsetB = ldrow3(p, L::kOffsetB);
        float friction = ldrow(p, L::kFriction), maxRecovery = ldrow(p, L::kMaxRecovery);
        Springiness sp = compute_springiness(ldrow(p, L::kAngularFrequency), ldrow(p, L::kTwiceDampingRatio), dt);
        const float inverseMassOther = X::swap(mine.inv_mass);
        const float linear = (isA ? mine.inv_mass : inverseMassOther) + (isA ? inverseMassOther : mine.inv_mass);  // iA.inv_mass + iB.inv_mass
        V3 offs[N];
        float depths[N], pen[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { offs[i] = ldrow3(p, 4 * i); depths[i] = ldrow(p, 4 * i + 3); pen[i] = ldacc(a, 2 + i); }
#pragma unroll
        for (int i = 0; i < N; ++i) split_penetration_solve<X>(isA, mine, linear, normal, offs[i], offs[i] - offsetB, depths[i], sp, maxRecovery, inverseDt, pen[i], v);
        V3 x, z;
        build_orthonormal_basis(normal, x, z);
        V2 tangent{ldacc(a, 0), ldacc(a, 1)};
        float twist = ldacc(a, N + 2);
        if constexpr (N == 1) {
            split_tangent_solve<X>(isA, x, z, offs[0], offs[0] - offsetB, mine, friction * pen[0], tangent, v);
            split_twist_solve<X>(isA, normal, mine, friction * pen[0] * fmax_ps(0.0f, depths[0]), twist, v);
        } else {
            float premultiplied = (1.0f / N) * friction;
            float penSum = pen[0];
#pragma unroll
            for (int i = 1; i < N; ++i) penSum = penSum + pen[i];
            V3 centerA = friction_center<N>(offs, depths);
            split_tangent_solve<X>(isA, x, z, centerA, centerA - offsetB, mine, premultiplied * penSum, tangent, v);
            float twistSum = pen[0] * distance(centerA, offs[0]);
#pragma unroll
            for (int i = 1; i < N; ++i) twistSum = twistSum + pen[i] * distance(centerA, offs[i]);
            split_twist_solve<X>(isA, normal, mine, premultiplied * twistSum, twist, v);
        }
        if (isA) {  // both lanes hold the same impulses; one writes them
            stacc(a, 0, tangent.x);
            stacc(a, 1, tangent.y);
#pragma unroll
            for (int i = 0; i < N; ++i) stacc(a, 2 + i, pen[i]);
            stacc(a, N + 2, twist);
        }
    }
};

}  // namespace BEPU_NS
