// PredictBoundingBoxes kernel (sm_100a), compiled -fmad=false with IEEE sqrt: the results are bit-identical to a non-contracting CPU evaluation of
// the reference's expressions (tests/test_bounds.py). One thread per body, fully coalesced: reads the 32-B pose / velocity / local inertia records
// the solver keeps resident (96 B), the 32-B shape record and the 8-B activity; writes 32 B of bounds + margin and the activity. HBM-bound
// (168 B per body); the bounding boxes of every body are independent, so there is nothing to order.
#define BEPU_NS bepu_bounds_math
#include "bepu_device_math.cuh"
#include "bepu_bounds.h"

namespace bepucuda {

namespace {

using namespace bepu_bounds_math;

struct LocalBounds { V3 max; float maximumRadius, maximumAngularExpansion; };

// IConvexShape wide GetBounds, one lane: Sphere.cs:L149-160, Capsule.cs:L226-239, Box.cs:L211-222, Cylinder.cs:L222-235. All four are symmetric (min = -max).
__device__ __forceinline__ LocalBounds shape_bounds(const BodyShape& s, Q4 q) {
    LocalBounds r;
    if (s.type == 0) {
        r.max = {s.a, s.a, s.a};
        r.maximumRadius = 0.0f;
        r.maximumAngularExpansion = 0.0f;
    } else if (s.type == 1) {
        const float radius = s.a, halfLength = s.b;
        V3 segmentOffset = transform_unit_y(q) * halfLength;
        segmentOffset = {fabsf(segmentOffset.x), fabsf(segmentOffset.y), fabsf(segmentOffset.z)};
        r.max = {segmentOffset.x + radius, segmentOffset.y + radius, segmentOffset.z + radius};
        r.maximumRadius = halfLength + radius;
        r.maximumAngularExpansion = halfLength;
    } else if (s.type == 2) {
        const float halfWidth = s.a, halfHeight = s.b, halfLength = s.c;
        const M33 basis = matrix_from_quaternion(q);
        r.max.x = fabsf(halfWidth * basis.x.x) + fabsf(halfHeight * basis.y.x) + fabsf(halfLength * basis.z.x);
        r.max.y = fabsf(halfWidth * basis.x.y) + fabsf(halfHeight * basis.y.y) + fabsf(halfLength * basis.z.y);
        r.max.z = fabsf(halfWidth * basis.x.z) + fabsf(halfHeight * basis.y.z) + fabsf(halfLength * basis.z.z);
        r.maximumRadius = sqrtf(halfWidth * halfWidth + halfHeight * halfHeight + halfLength * halfLength);
        // as written in the reference (Box.cs:L221): HalfLength appears twice, HalfWidth not at all
        r.maximumAngularExpansion = r.maximumRadius - fmin_ps(halfLength, fmin_ps(halfHeight, halfLength));
    } else {
        const float radius = s.a, halfLength = s.b;
        const V3 y = transform_unit_y(q);
        const V3 squared = {1.0f - y.x * y.x, 1.0f - y.y * y.y, 1.0f - y.z * y.z};
        r.max.x = fabsf(halfLength * y.x) + sqrtf(fmax_ps(0.0f, squared.x)) * radius;
        r.max.y = fabsf(halfLength * y.y) + sqrtf(fmax_ps(0.0f, squared.y)) * radius;
        r.max.z = fabsf(halfLength * y.z) + sqrtf(fmax_ps(0.0f, squared.z)) * radius;
        r.maximumRadius = sqrtf(halfLength * halfLength + radius * radius);
        r.maximumAngularExpansion = r.maximumRadius - fmin_ps(halfLength, radius);
    }
    return r;
}

// BoundingBoxHelpers.GetAngularBoundsExpansion (BoundingBoxHelpers.cs:L12-45)
__device__ __forceinline__ float angular_bounds_expansion(float angularSpeed, float dt, float maximumRadius, float maximumAngularExpansion) {
    const float a = fmin_ps(angularSpeed * dt, 3.14159274f / 3.0f);
    const float a2 = a * a;
    const float a4 = a2 * a2;
    const float a6 = a4 * a2;
    const float cosAngleMinusOne = a2 * (-1.0f / 2.0f) + a4 * (1.0f / 24.0f) - a6 * (1.0f / 720.0f);
    return fmin_ps(maximumAngularExpansion, sqrtf(-2.0f * maximumRadius * maximumRadius * cosAngleMinusOne));
}

__global__ void predict_bounding_boxes_kernel(BodyBuffers B, const BodyShape* __restrict__ shapes, BodyActivityRecord* __restrict__ activities, float4* __restrict__ bounds,
                                              const __grid_constant__ PredictParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.count) return;
    const float4 q4 = B.pose[2 * (size_t)i], p4 = B.pose[2 * (size_t)i + 1];
    const float4 l4 = B.velocity[2 * (size_t)i], w4 = B.velocity[2 * (size_t)i + 1];
    const float4 i0 = B.inertia_local[2 * (size_t)i], i1 = B.inertia_local[2 * (size_t)i + 1];
    const Q4 orientation = {q4.x, q4.y, q4.z, q4.w};
    const V3 position = {p4.x, p4.y, p4.z};
    Velocity velocity = {{l4.x, l4.y, l4.z}, {w4.x, w4.y, w4.z}};
    // Bodies.IsKinematic (Bodies.cs:L326-331): every bit of inverse mass and inverse inertia is zero
    const bool kinematic = (__float_as_uint(i1.z) | __float_as_uint(i0.x) | __float_as_uint(i0.y) | __float_as_uint(i0.z) | __float_as_uint(i0.w) | __float_as_uint(i1.x) | __float_as_uint(i1.y)) == 0u;
    const bool integrate = p.integrate_velocity_for_kinematics != 0 || !kinematic;
    const float sleepEnergy = length_squared(velocity.lin) + length_squared(velocity.ang);
    // the integrated velocity is only used for the prediction; it is not stored (PoseIntegrator.cs:L339)
    if (integrate) {
        velocity.lin = (velocity.lin + V3{p.gravity_dt[0], p.gravity_dt[1], p.gravity_dt[2]}) * p.linear_damping_dt;
        velocity.ang = velocity.ang * p.angular_damping_dt;
    }
    // UpdateSleepCandidacy (PoseIntegrator.cs:L286-304)
    BodyActivityRecord activity = activities[i];
    if (sleepEnergy > activity.sleep_threshold) {
        activity.timesteps_under_threshold_count = 0;
        activity.sleep_candidate = 0;
    } else if (activity.timesteps_under_threshold_count < 255) {
        ++activity.timesteps_under_threshold_count;
        if (activity.timesteps_under_threshold_count >= activity.minimum_timesteps_under_threshold) activity.sleep_candidate = 1;
    }
    activities[i] = activity;

    const BodyShape shape = shapes[i];
    if (!(shape.type == 0 || shape.type == 1 || shape.type == 2 || shape.type == 4)) {
        bounds[2 * (size_t)i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        bounds[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    // BoundingBoxBatcher.ExecuteConvexBatch (BoundingBoxBatcher.cs:L176-197)
    const LocalBounds local = shape_bounds(shape, orientation);
    const float angularBoundsExpansion = angular_bounds_expansion(length(velocity.ang), p.dt, local.maximumRadius, local.maximumAngularExpansion);
    float speculativeMargin = length(velocity.lin) * p.dt + angularBoundsExpansion;
    speculativeMargin = fmax_ps(shape.minimum_speculative_margin, fmin_ps(shape.maximum_speculative_margin, speculativeMargin));
    const float maximumBoundsExpansion = shape.allow_expansion_beyond_speculative_margin ? 3.40282347e+38f : speculativeMargin;
    // BoundingBoxHelpers.GetBoundsExpansion (BoundingBoxHelpers.cs:L49-58)
    const V3 linearDisplacement = velocity.lin * p.dt;
    V3 minExpansion = {fmin_ps(0.0f, linearDisplacement.x) - angularBoundsExpansion, fmin_ps(0.0f, linearDisplacement.y) - angularBoundsExpansion, fmin_ps(0.0f, linearDisplacement.z) - angularBoundsExpansion};
    V3 maxExpansion = {fmax_ps(0.0f, linearDisplacement.x) + angularBoundsExpansion, fmax_ps(0.0f, linearDisplacement.y) + angularBoundsExpansion, fmax_ps(0.0f, linearDisplacement.z) + angularBoundsExpansion};
    minExpansion = {fmax_ps(-maximumBoundsExpansion, minExpansion.x), fmax_ps(-maximumBoundsExpansion, minExpansion.y), fmax_ps(-maximumBoundsExpansion, minExpansion.z)};
    maxExpansion = {fmin_ps(maximumBoundsExpansion, maxExpansion.x), fmin_ps(maximumBoundsExpansion, maxExpansion.y), fmin_ps(maximumBoundsExpansion, maxExpansion.z)};
    const V3 bundleMin = position + ((-local.max) + minExpansion);
    const V3 bundleMax = position + (local.max + maxExpansion);
    bounds[2 * (size_t)i] = make_float4(bundleMin.x, bundleMin.y, bundleMin.z, speculativeMargin);
    bounds[2 * (size_t)i + 1] = make_float4(bundleMax.x, bundleMax.y, bundleMax.z, 1.0f);
}

}  // namespace

void launch_predict_bounding_boxes(const BodyBuffers& B, const BodyShape* shapes, BodyActivityRecord* activities, float4* bounds, const PredictParams& params, cudaStream_t s) {
    if (B.count <= 0) return;
    predict_bounding_boxes_kernel<<<(unsigned)((B.count + 255) / 256), 256, 0, s>>>(B, shapes, activities, bounds, params);
}

}  // namespace bepucuda
