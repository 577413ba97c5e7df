// Solver stage kernels (sm_100a): WarmStart (with embedded pose/velocity integration), Solve, IncrementallyUpdateForSubstep,
// the kinematic prepasses and the final pose pass. Included once per numerics flavour (BEPU_NS = bepu_fast with FMA
// contraction, bepu_strict with -fmad=false).
//
// Mapping: one warp = one 32-lane bundle, lane = constraint (TwoBodyTypeProcessor.cs:L176-224 processes one
// Vector<float>.Count-wide bundle per loop trip; here the trip is a warp). Constraint rows are 128-B coalesced lines;
// body state is gathered as 32-B records (two LDG.128 each) straight from L2 (ld.global.cg) — body data is the random
// part of the access pattern and is written by other warps of the same launch sequence, so it never goes through L1.
#pragma once
#include "bepu_contacts.cuh"
#include "bepu_joints_more.cuh"
#include "bepu_device_types.h"

namespace BEPU_NS {

using namespace bepucuda;

// ---- body records: one 32-byte record = one DRAM sector = ONE 256-bit load/store (LDG.E.256 / STG.E.256, new on sm_100) ----------------
struct F8 { float a, b, c, d, e, f, g, h; };
BEPU_DI F8 ld256(const float4* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a), "=f"(r.b), "=f"(r.c), "=f"(r.d), "=f"(r.e), "=f"(r.f), "=f"(r.g), "=f"(r.h)
                 : "l"(p)
                 : "memory");
    return r;
}
BEPU_DI void st256(float4* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
BEPU_DI void load_velocity(const float4* vel, uint32_t i, Velocity& v) {
    const F8 r = ld256(vel + 2 * (size_t)i);
    v.lin = {r.a, r.b, r.c};
    v.ang = {r.e, r.f, r.g};
}
BEPU_DI void store_velocity(float4* vel, uint32_t i, const Velocity& v) { st256(vel + 2 * (size_t)i, v.lin.x, v.lin.y, v.lin.z, 0.0f, v.ang.x, v.ang.y, v.ang.z, 0.0f); }
BEPU_DI void load_inertia(const float4* in, uint32_t i, Inertia& r) {
    const F8 x = ld256(in + 2 * (size_t)i);
    r.t = {x.a, x.b, x.c, x.d, x.e, x.f};
    r.inv_mass = x.g;
}
BEPU_DI void store_inertia(float4* in, uint32_t i, const Inertia& r) { st256(in + 2 * (size_t)i, r.t.xx, r.t.yx, r.t.yy, r.t.zx, r.t.zy, r.t.zz, r.inv_mass, 0.0f); }
BEPU_DI void load_pose(const float4* pose, uint32_t i, V3& pos, Q4& q) {
    const F8 x = ld256(pose + 2 * (size_t)i);
    q = {x.a, x.b, x.c, x.d};
    pos = {x.e, x.f, x.g};
}
BEPU_DI void store_pose(float4* pose, uint32_t i, V3 pos, Q4 q) { st256(pose + 2 * (size_t)i, q.x, q.y, q.z, q.w, pos.x, pos.y, pos.z, 0.0f); }

// ---- integration (PoseIntegrator.cs:L99-261, TypeProcessor.cs:L1204-1283, Demos/DemoCallbacks.cs:L99-102) -----------
BEPU_DI Q4 integrate_orientation(Q4 start, V3 w, float halfDt) {  // PoseIntegrator.cs:L146-164
    float speed = length(w);
    float halfAngle = speed * halfDt;
    float s = sin_approx(halfAngle);
    float scl = s / speed;
    Q4 q{w.x * scl, w.y * scl, w.z * scl, cos_approx(halfAngle)};
    Q4 end = normalize(concatenate(start, q));
    return speed > 1e-15f ? end : start;
}
BEPU_DI Sym3 rotate_inverse_inertia(Sym3 local, Q4 q) { return rotation_sandwich(matrix_from_quaternion(q), local); }  // L166-175
BEPU_DI void callback_integrate_velocity(Velocity& v, float gx, float gy, float gz, float linearDampingDt, float angularDampingDt) {
    v.lin = (v.lin + V3{gx, gy, gz}) * linearDampingDt;
    v.ang = v.ang * angularDampingDt;
}
BEPU_DI void fallback_if_inertia_incompatible(V3 previous, V3& w) {  // L180-190
    const float inf = __int_as_float(0x7f800000);
    bool useNew = fabsf(w.x) < inf && fabsf(w.y) < inf && fabsf(w.z) < inf;
    w = useNew ? w : previous;
}
// The two momentum-conserving modes are rare (AngularIntegrationMode.Nonconserving is the default everywhere in the
// reference's demos/benchmarks); keeping them out of line keeps their registers out of the hot WarmStart path.
__device__ __noinline__ void integrate_angular_conserve_momentum(Q4 previousOrientation, Sym3 localInverseInertia, Sym3 worldInverseInertia, V3& w) {  // L192-206
    M33 prevR = matrix_from_quaternion(previousOrientation);
    V3 localPrevW = transform_by_transposed(w, prevR);
    Sym3 localInertiaTensor = invert(localInverseInertia);
    V3 angularMomentum = transform(transform(localPrevW, localInertiaTensor), prevR);
    V3 previous = w;
    w = transform(angularMomentum, worldInverseInertia);
    fallback_if_inertia_incompatible(previous, w);
}
__device__ __noinline__ void integrate_angular_gyroscopic(Q4 orientation, Sym3 localInverseInertia, V3& w, float dt) {  // L208-253
    M33 R = matrix_from_quaternion(orientation);
    V3 localW = transform_by_transposed(w, R);
    Sym3 I = invert(localInverseInertia);
    V3 localMomentum = transform(localW, I);
    V3 residual = cross(localMomentum, localW) * dt;
    M33 skewMomentum{{0.0f, -localMomentum.z, localMomentum.y}, {localMomentum.z, 0.0f, -localMomentum.x}, {-localMomentum.y, localMomentum.x, 0.0f}};
    M33 skewVelocity{{0.0f, -localW.z, localW.y}, {localW.z, 0.0f, -localW.x}, {-localW.y, localW.x, 0.0f}};
    M33 tsv = multiply(skewVelocity, I);
    M33 J;
    V3 cx = (tsv.x - skewMomentum.x) * dt, cy = (tsv.y - skewMomentum.y) * dt, cz = (tsv.z - skewMomentum.z) * dt;
    J.x = {I.xx + cx.x, I.yx + cx.y, I.zx + cx.z};
    J.y = {I.yx + cy.x, I.yy + cy.y, I.zy + cy.z};
    J.z = {I.zx + cz.x, I.zy + cz.y, I.zz + cz.z};
    V3 newtonStep = transform(residual, invert(J));
    localW = localW - newtonStep;
    V3 previous = w;
    w = transform(localW, R);
    fallback_if_inertia_incompatible(previous, w);
}

// GatherAndIntegrate for one body slot of one lane (TypeProcessor.cs:L1298-1397), after the velocity has been loaded. The lane integrates iff the
// device body reference carries kRefIntegrateBit; all other lanes read the world inertia their owner constraint stored earlier in this substep,
// which is bit-identical to what the reference's bundle-wide recompute would give them.
template <int STAGE, bool NeedsPose>
BEPU_DI void warm_start_body(uint32_t enc, const BodyBuffers& B, const FrameParams& fp, BodyState& b, Velocity& v) {
    const uint32_t idx = enc & kRefIndexMask;
    if (enc & kRefIntegrateBit) {
        Inertia local;
        load_inertia(B.inertia_local, idx, local);
        load_pose(B.pose, idx, b.pos, b.q);
        b.inertia.inv_mass = local.inv_mass;
        if (STAGE == kStageWarmStart) {
            // IntegratePoseAndVelocity, TypeProcessor.cs:L1204-1248
            b.pos = b.pos + v.lin * fp.dt;
            Q4 previousOrientation = b.q;
            b.q = integrate_orientation(b.q, v.ang, fp.dt * 0.5f);
            b.inertia.t = rotate_inverse_inertia(local.t, b.q);
            if (fp.angular_mode == 1) integrate_angular_conserve_momentum(previousOrientation, local.t, b.inertia.t, v.ang);
            else if (fp.angular_mode == 2) integrate_angular_gyroscopic(b.q, local.t, v.ang, fp.dt);
            store_pose(B.pose, idx, b.pos, b.q);
        } else {
            // IntegrateVelocity, TypeProcessor.cs:L1251-1283
            b.inertia.t = rotate_inverse_inertia(local.t, b.q);
            if (fp.angular_mode == 1) {
                Q4 previousOrientation = integrate_orientation(b.q, v.ang, fp.dt * -0.5f);
                integrate_angular_conserve_momentum(previousOrientation, local.t, b.inertia.t, v.ang);
            } else if (fp.angular_mode == 2) {
                integrate_angular_gyroscopic(b.q, local.t, v.ang, fp.dt);
            }
        }
        callback_integrate_velocity(v, fp.gravity_dt[0], fp.gravity_dt[1], fp.gravity_dt[2], fp.linear_damping_dt, fp.angular_damping_dt);
        store_inertia(B.inertia_world, idx, b.inertia);
    } else {
        load_inertia(B.inertia_world, idx, b.inertia);
        if (NeedsPose) load_pose(B.pose, idx, b.pos, b.q);
        if (STAGE == kStageWarmStartFirst && fp.angular_mode != 0 && (enc & kRefBundleIntegratesBit) && !(enc & kRefKinematicBit)) {
            // Reference quirk, reproduced for identical results: in the first substep IntegrateVelocity runs the momentum-conserving angular
            // update on EVERY lane of a bundle that contains an integrating lane and only masks the callback afterwards
            // (TypeProcessor.cs:L1259-1281), so non-owning dynamic lanes of such a (host-width) bundle get the update too.
            Inertia local;
            load_inertia(B.inertia_local, idx, local);
            V3 pos;
            Q4 q;
            load_pose(B.pose, idx, pos, q);
            if (fp.angular_mode == 1) integrate_angular_conserve_momentum(integrate_orientation(q, v.ang, fp.dt * -0.5f), local.t, b.inertia.t, v.ang);
            else integrate_angular_gyroscopic(q, local.t, v.ang, fp.dt);
        }
    }
}
template <int STAGE, bool NeedsPose>
BEPU_DI void gather_for_warm_start(uint32_t enc, const BodyBuffers& B, const FrameParams& fp, BodyState& b, Velocity& v) {
    load_velocity(B.velocity, enc & kRefIndexMask, v);
    warm_start_body<STAGE, NeedsPose>(enc, B, fp, b, v);
}

// ---- uniform call shapes over contact and joint types ------------------------------------------------------------------
template <class T> BEPU_DI void call_warm_start(const BodyState* b, const float* p, const float* a, Velocity* v) {
    if constexpr (T::kNeedsPose) T::warm_start(b, p, a, v);
    else if constexpr (T::kBodies == 2) T::warm_start(b[0].inertia, b[1].inertia, p, a, v[0], v[1]);
    else T::warm_start(b[0].inertia, p, a, v[0]);
}
template <class T> BEPU_DI void call_solve(const BodyState* b, float dt, float inverseDt, const float* p, float* a, Velocity* v) {
    if constexpr (T::kNeedsPose) T::solve(b, dt, inverseDt, p, a, v);
    else if constexpr (T::kBodies == 2) T::solve(b[0].inertia, b[1].inertia, dt, inverseDt, p, a, v[0], v[1]);
    else T::solve(b[0].inertia, dt, inverseDt, p, a, v[0]);
}
template <class T> BEPU_DI void call_incremental(float dt, const Velocity* v, float* p) {
    if constexpr (T::kIncremental) {
        if constexpr (T::kBodies == 2) T::incremental_update(dt, v[0], v[1], p);
        else T::incremental_update(dt, v[0], p);
    }
}

// One constraint lane of one stage. refs/p/a address this lane in row 0 of the bundle; enc0/enc1 are the (possibly prefetched) first two body references.
template <class T, int STAGE>
BEPU_DI void run_lane(const int32_t* refs, float* p, float* a, uint32_t enc0, uint32_t enc1, const BodyBuffers& B, const FrameParams& fp) {
    constexpr int NB = T::kBodies;
    uint32_t enc[NB];
    enc[0] = enc0;
    if constexpr (NB > 1) enc[1] = enc1;
#pragma unroll
    for (int s = 2; s < NB; ++s) enc[s] = (uint32_t)__ldg(refs + s * kLanes);
    if ((int32_t)enc[0] == kRefEmpty) return;  // trailing lane of the last bundle, or a hole in a fallback bundle
    BodyState b[NB];
    Velocity v[NB];
    if constexpr (STAGE == kStageIncremental) {
#pragma unroll
        for (int s = 0; s < NB; ++s) load_velocity(B.velocity, enc[s] & kRefIndexMask, v[s]);
        call_incremental<T>(fp.dt, v, p);
    } else if constexpr (STAGE == kStageSolve) {
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            const uint32_t idx = enc[s] & kRefIndexMask;
            load_velocity(B.velocity, idx, v[s]);
            load_inertia(B.inertia_world, idx, b[s].inertia);
            if (T::kNeedsPose) load_pose(B.pose, idx, b[s].pos, b[s].q);
        }
        call_solve<T>(b, fp.dt, fp.inverse_dt, p, a, v);
#pragma unroll
        for (int s = 0; s < NB; ++s)
            if (!(enc[s] & kRefKinematicBit)) store_velocity(B.velocity, enc[s] & kRefIndexMask, v[s]);
    } else {
#pragma unroll
        for (int s = 0; s < NB; ++s) gather_for_warm_start<STAGE, T::kNeedsPose>(enc[s], B, fp, b[s], v[s]);
        call_warm_start<T>(b, p, a, v);
#pragma unroll
        for (int s = 0; s < NB; ++s)
            if (!(enc[s] & kRefKinematicBit)) store_velocity(B.velocity, enc[s] & kRefIndexMask, v[s]);
    }
}

BEPU_DI WorkRecord load_record(const WorkRecord* r) {
    const int4 lo = __ldg(reinterpret_cast<const int4*>(r)), hi = __ldg(reinterpret_cast<const int4*>(r) + 1);
    WorkRecord w;
    w.refs = reinterpret_cast<int32_t*>((unsigned long long)(unsigned int)lo.x | ((unsigned long long)(unsigned int)lo.y << 32));
    w.prestep = reinterpret_cast<float*>((unsigned long long)(unsigned int)lo.z | ((unsigned long long)(unsigned int)lo.w << 32));
    w.impulses = reinterpret_cast<float*>((unsigned long long)(unsigned int)hi.x | ((unsigned long long)(unsigned int)hi.y << 32));
    w.type_id = hi.z;
    w.live_lanes = hi.w;
    return w;
}

// Type registry: BatchTypeId constants of the reference (Contact/ContactConvexTypes.cs, ContactNonconvexTypes.cs, joint files).
#define BEPU_CONTACT_TYPES(X)                                                                                                     \
    X(0, ConvexOneBody<1>) X(1, ConvexOneBody<2>) X(2, ConvexOneBody<3>) X(3, ConvexOneBody<4>)                                   \
    X(4, ConvexTwoBody<1>) X(5, ConvexTwoBody<2>) X(6, ConvexTwoBody<3>) X(7, ConvexTwoBody<4>)                                   \
    X(8, NonconvexOneBody<2>) X(9, NonconvexOneBody<3>) X(10, NonconvexOneBody<4>)                                                \
    X(15, NonconvexTwoBody<2>) X(16, NonconvexTwoBody<3>) X(17, NonconvexTwoBody<4>)

template <int STAGE>
BEPU_DI void run_bundle(const WorkRecord& rec, int lane, uint32_t enc0, uint32_t enc1, const BodyBuffers& B, const FrameParams& fp) {
    const int32_t* refs = rec.refs + lane;
    float* p = rec.prestep + lane;
    float* a = rec.impulses + lane;
    switch (rec.type_id) {
#define BEPU_CASE(ID, T) \
    case ID: run_lane<T, STAGE>(refs, p, a, enc0, enc1, B, fp); break;
        BEPU_CONTACT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES_MORE(BEPU_CASE)
#undef BEPU_CASE
        default: break;
    }
}
// The reference arena is padded, so reading a second body-reference row is always in bounds (one-body types ignore it).
template <int STAGE> BEPU_DI void run_bundle(const WorkRecord& rec, int lane, const BodyBuffers& B, const FrameParams& fp) {
    const uint32_t enc0 = (uint32_t)__ldg(rec.refs + lane), enc1 = (uint32_t)__ldg(rec.refs + kLanes + lane);
    run_bundle<STAGE>(rec, lane, enc0, enc1, B, fp);
}

// ---- kernels ------------------------------------------------------------------------------------------------------------
constexpr int kStageBlockThreads = 64;
#ifndef BEPU_STAGE_MIN_BLOCKS
#define BEPU_STAGE_MIN_BLOCKS 1
#endif

template <int STAGE>
__global__ void __launch_bounds__(kStageBlockThreads, BEPU_STAGE_MIN_BLOCKS) constraint_stage_kernel(const WorkRecord* __restrict__ records, int work_count, BodyBuffers B, const FrameParams* __restrict__ fpp) {
    // Programmatic dependent launch: let the NEXT stage's grid become resident right away, and do everything that does not depend on
    // the previous stage (work record, body references, frame scalars: all immutable during a solve) before waiting for it.
    asm volatile("griddepcontrol.launch_dependents;");
    const int warp = (blockIdx.x * kStageBlockThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    WorkRecord rec{};
    uint32_t enc0 = (uint32_t)kRefEmpty, enc1 = 0;
    const bool active = warp < work_count;
    if (active) {
        rec = load_record(records + warp);
        enc0 = (uint32_t)__ldg(rec.refs + lane);
        enc1 = (uint32_t)__ldg(rec.refs + kLanes + lane);
    }
    const FrameParams fp = *fpp;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (active) run_bundle<STAGE>(rec, lane, enc0, enc1, B, fp);
}

// IntegrateKinematicVelocities / IntegrateKinematicPosesAndVelocities (PoseIntegrator.cs:L451-487, L493-535)
template <int STAGE> BEPU_DI void run_kinematic(int i, const int32_t* kinematics, const BodyBuffers& B, const FrameParams& fp) {
    const uint32_t idx = (uint32_t)kinematics[i];
    Velocity v;
    load_velocity(B.velocity, idx, v);
    if (STAGE == kStageKinematic) {
        V3 pos;
        Q4 q;
        load_pose(B.pose, idx, pos, q);
        pos = pos + v.lin * fp.dt;
        q = integrate_orientation(q, v.ang, fp.dt * 0.5f);
        store_pose(B.pose, idx, pos, q);
    }
    if (fp.integrate_velocity_for_kinematics) {
        callback_integrate_velocity(v, fp.gravity_dt[0], fp.gravity_dt[1], fp.gravity_dt[2], fp.linear_damping_dt, fp.angular_damping_dt);
        store_velocity(B.velocity, idx, v);
    }
}
template <int STAGE>
__global__ void kinematic_stage_kernel(const int32_t* __restrict__ kinematics, int count, BodyBuffers B, const FrameParams* __restrict__ fpp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const FrameParams fp = *fpp;
    run_kinematic<STAGE>(i, kinematics, B, fp);
}

// IntegrateBundlesAfterSubstepping (PoseIntegrator.cs:L537-693), per body.
BEPU_DI void run_final_pose(int i, const BodyBuffers& B, const FrameParams& fp) {
    V3 pos;
    Q4 q;
    Velocity v;
    load_pose(B.pose, i, pos, q);
    load_velocity(B.velocity, i, v);
    if (B.constrained[i]) {
        // constrained bodies: the one trailing pose integration of (velocity -> solve) -> (pose -> velocity -> solve) ... -> pose
        q = integrate_orientation(q, v.ang, fp.dt * 0.5f);
        pos = pos + v.lin * fp.dt;
        store_pose(B.pose, i, pos, q);
        return;
    }
    Inertia local;
    load_inertia(B.inertia_local, i, local);
    const bool kinematic = local.inv_mass == 0.0f && local.t.xx == 0.0f && local.t.yx == 0.0f && local.t.yy == 0.0f && local.t.zx == 0.0f && local.t.zy == 0.0f && local.t.zz == 0.0f;
    const bool integrateVelocity = fp.integrate_velocity_for_kinematics || !kinematic;
    const float dt = fp.final_dt, halfDt = fp.final_dt * 0.5f;
    for (int step = 0; step < fp.final_steps; ++step) {
        if (integrateVelocity)
            callback_integrate_velocity(v, fp.final_gravity_dt[0], fp.final_gravity_dt[1], fp.final_gravity_dt[2], fp.final_linear_damping_dt, fp.final_angular_damping_dt);
        pos = pos + v.lin * dt;
        if (fp.angular_mode == 1) {
            Q4 previousOrientation = q;
            q = integrate_orientation(q, v.ang, halfDt);
            integrate_angular_conserve_momentum(previousOrientation, local.t, rotate_inverse_inertia(local.t, q), v.ang);
        } else if (fp.angular_mode == 2) {
            q = integrate_orientation(q, v.ang, halfDt);
            integrate_angular_gyroscopic(q, local.t, v.ang, dt);
        } else {
            q = integrate_orientation(q, v.ang, halfDt);
        }
    }
    store_pose(B.pose, i, pos, q);
    if (integrateVelocity) store_velocity(B.velocity, i, v);
}
__global__ void final_pose_kernel(BodyBuffers B, const FrameParams* __restrict__ fpp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.count) return;
    const FrameParams fp = *fpp;
    run_final_pose(i, B, fp);
}

}  // namespace BEPU_NS
