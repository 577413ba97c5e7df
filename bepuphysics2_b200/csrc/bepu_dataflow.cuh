// Dataflow persistent solver kernel (BEPUCUDA_EXEC_DATAFLOW).
//
// A (batch, stage) of a 100 k-body scene is half a wave of warps, so with one grid-wide barrier (or kernel boundary) per (batch, stage) a frame is a
// chain of ~390 latency-bound steps of launch + gather + math + scatter. But a constraint in batch k only depends on the (at most one per body)
// constraints of earlier batches that touch ITS bodies. This kernel keeps exactly those dependencies and drops the rest:
//
//   * every dynamic body carries a version counter in the two padding words of its 32-byte velocity record (one in each 16-byte half, so a torn
//     read is detected); a version counts the constraint evaluations that have written the body since the last reset;
//   * a lane that is the r-th of the K constraints on body X (device batch order: "rank", "degree") expects version P*K + r at WarmStart/Solve pass
//     P (passes counted from the reset) and publishes P*K + r + 1 together with the new velocity in ONE 256-bit store. Per body this replays
//     exactly the reference's Gauss-Seidel order (Solver_Solve.cs:L1447-1476), so results are bit-identical to the barrier schedule;
//   * waiting costs no memory bandwidth: every bundle has ONE notification counter. A lane that has written body X adds 1 to the counter of the
//     bundle holding the NEXT constraint on X (static: the "successor" table built at bepucuda_end_constraints); a warp polls only its bundle's
//     counter (one 4-byte load per poll for the whole warp) until all of its (lane, body) dependencies of this pass have reported, and only then
//     gathers the velocity records. A producer releases (fence.acq_rel.gpu) between its record stores and its notifications; the gather still
//     checks every record's version, so a consumer can never compute on a stale record;
//   * warps own bundles statically (bundle g -> warp g mod T) and walk their bundles in program order; while a warp waits it already holds its
//     bundle's prestep + impulse block in its shared-memory slab (one cp.async.bulk pair), its body references and -- in Solve passes -- the world
//     inertias, so a dependency link costs (notification visible) + (velocity gather) + math + (store);
//   * progress: the earliest unfinished evaluation in program order never waits (all its producers are earlier), and every warp reaches its items in
//     program order, so with all CTAs co-resident (cooperative launch) a pass cannot deadlock. A spin limit raises an error flag instead of hanging.
//
// One launch = one WarmStart or Solve pass over ALL device batches (dataflow_pass_kernel<STAGE>): a step of 8 substeps x (1 + 2) passes is 24 of these
// plus the whole-set stages the reference has anyway (IncrementallyUpdateForSubstep, kinematic prepass, final pose pass) instead of 392 stage kernels.
// Everything is inlined into the type switch: a call per bundle would put the callee-saved registers and by-reference arguments in local memory.
#pragma once
#include "bepu_persistent.cuh"

namespace BEPU_NS {

constexpr unsigned int kDataflowSpinLimit = 20000000u;  // ~1 s of polling: a dependency that never arrives is a bug, not a reason to hang the GPU
#ifndef BEPU_DATAFLOW_THREADS
#define BEPU_DATAFLOW_THREADS 256
#endif
constexpr int kDataflowThreads = BEPU_DATAFLOW_THREADS;
constexpr int kDataflowWarps = kDataflowThreads / 32;
constexpr int kDataflowSmemBytes = kDataflowWarps * kStageSlabBytes + kDataflowWarps * 8;
#ifndef BEPU_DATAFLOW_MINB
#define BEPU_DATAFLOW_MINB 2
#endif

// Body records inside a pass are read and written with STRONG gpu-scope accesses (LDG/STG.E.256.STRONG.GPU): they are performed at the L2, the
// coherence point of all SMs. Weak (.cg) stores were observed to stay invisible to other SMs for as long as a consumer spun on the record.
BEPU_DI F8 ld256_strong(const float4* p) {
    F8 r;
    asm volatile("ld.relaxed.gpu.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a), "=f"(r.b), "=f"(r.c), "=f"(r.d), "=f"(r.e), "=f"(r.f), "=f"(r.g), "=f"(r.h)
                 : "l"(p)
                 : "memory");
    return r;
}
BEPU_DI void st256_strong(float4* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.relaxed.gpu.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
BEPU_DI void store_velocity_versioned(float4* vel, uint32_t i, const Velocity& v, uint32_t version) {
    const float ver = __uint_as_float(version);
    st256_strong(vel + 2 * (size_t)i, v.lin.x, v.lin.y, v.lin.z, ver, v.ang.x, v.ang.y, v.ang.z, ver);
}
BEPU_DI unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
BEPU_DI void red_add_u32(unsigned int* p, unsigned int v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// A dependency that never arrives is a bug in the tables, not a reason to hang the GPU: the first lane to give up records what it was waiting for
// (error_flag[0] = 4, [1] = what: 1 counter, 2 velocity version, 3 inertia stamp, 4 pose stamp, [2] = expected, [3] = observed, [4] = body or bundle).
BEPU_DI void report_stall(int32_t* error_flag, int what, uint32_t expected, uint32_t observed, uint32_t where) {
    if (atomicCAS(error_flag, 0, 4) == 0) {
        error_flag[1] = what;
        error_flag[2] = (int32_t)expected;
        error_flag[3] = (int32_t)observed;
        error_flag[4] = (int32_t)where;
    }
}

// World inertia / pose records written by the integrating (first) constraint of a body carry a stamp in their padding word: the number of the
// WarmStart pass that wrote them + 1. Readers of the same substep check that stamp instead of relying on a fence in the writer.
BEPU_DI void store_inertia_stamped(float4* in, uint32_t i, const Inertia& r, uint32_t stamp) {
    st256_strong(in + 2 * (size_t)i, r.t.xx, r.t.yx, r.t.yy, r.t.zx, r.t.zy, r.t.zz, r.inv_mass, __uint_as_float(stamp));
}
BEPU_DI void store_pose_stamped(float4* pose, uint32_t i, V3 pos, Q4 q, uint32_t stamp) {
    st256_strong(pose + 2 * (size_t)i, q.x, q.y, q.z, q.w, pos.x, pos.y, pos.z, __uint_as_float(stamp));
}
// Re-reads a record until its tag word(s) carry `tag` (`both`: the two version words of a velocity record). Only entered when the first read,
// issued together with the bundle's other gathers, was early.
BEPU_DI bool reload_until(const float4* p, uint32_t tag, bool both, F8& x, int what, uint32_t where, int32_t* error_flag) {
    unsigned int spins = 0;
    bool good;
    do {
        x = ld256_strong(p);
        good = __float_as_uint(x.h) == tag && (!both || __float_as_uint(x.d) == tag);
    } while (!good && ++spins < kDataflowSpinLimit / 16);
    if (!good) report_stall(error_flag, what, tag, __float_as_uint(x.h), where);
    return good;
}
BEPU_DI bool load_inertia_stamped(const float4* in, uint32_t i, Inertia& r, uint32_t stamp, int32_t* error_flag) {
    F8 x = ld256_strong(in + 2 * (size_t)i);
    bool good = __float_as_uint(x.h) == stamp;
    if (!good) good = reload_until(in + 2 * (size_t)i, stamp, false, x, 3, i, error_flag);
    r.t = {x.a, x.b, x.c, x.d, x.e, x.f};
    r.inv_mass = x.g;
    return good;
}
BEPU_DI bool load_pose_stamped(const float4* pose, uint32_t i, V3& pos, Q4& q, uint32_t stamp, int32_t* error_flag) {
    F8 x = ld256_strong(pose + 2 * (size_t)i);
    bool good = __float_as_uint(x.h) == stamp;
    if (!good) good = reload_until(pose + 2 * (size_t)i, stamp, false, x, 4, i, error_flag);
    q = {x.a, x.b, x.c, x.d};
    pos = {x.e, x.f, x.g};
    return good;
}

// GatherAndIntegrate of warm_start_body (bepu_solver_kernels.cuh) with stamped records. `stamp` = this WarmStart pass + 1; pose_stamp = 0 when the
// pose was not rewritten in this substep (first substep: IntegrateVelocity only, TypeProcessor.cs:L1251-1283).
template <int STAGE, bool NeedsPose>
BEPU_DI void warm_start_body_dataflow(uint32_t enc, const BodyBuffers& B, const FrameParams& fp, BodyState& b, Velocity& v, uint32_t stamp, int32_t* error_flag) {
    const uint32_t idx = enc & kRefIndexMask;
    if (enc & kRefIntegrateBit) {
        Inertia local;
        load_inertia(B.inertia_local, idx, local);
        load_pose(B.pose, idx, b.pos, b.q);
        b.inertia.inv_mass = local.inv_mass;
        if (STAGE == kStageWarmStart) {
            b.pos = b.pos + v.lin * fp.dt;
            Q4 previousOrientation = b.q;
            b.q = integrate_orientation(b.q, v.ang, fp.dt * 0.5f);
            b.inertia.t = rotate_inverse_inertia(local.t, b.q);
            if (fp.angular_mode == 1) integrate_angular_conserve_momentum(previousOrientation, local.t, b.inertia.t, v.ang);
            else if (fp.angular_mode == 2) integrate_angular_gyroscopic(b.q, local.t, v.ang, fp.dt);
            store_pose_stamped(B.pose, idx, b.pos, b.q, stamp);
        } else {
            b.inertia.t = rotate_inverse_inertia(local.t, b.q);
            if (fp.angular_mode == 1) {
                Q4 previousOrientation = integrate_orientation(b.q, v.ang, fp.dt * -0.5f);
                integrate_angular_conserve_momentum(previousOrientation, local.t, b.inertia.t, v.ang);
            } else if (fp.angular_mode == 2) {
                integrate_angular_gyroscopic(b.q, local.t, v.ang, fp.dt);
            }
        }
        callback_integrate_velocity(v, fp.gravity_dt[0], fp.gravity_dt[1], fp.gravity_dt[2], fp.linear_damping_dt, fp.angular_damping_dt);
        store_inertia_stamped(B.inertia_world, idx, b.inertia, stamp);
    } else if (enc & kRefKinematicBit) {
        load_inertia(B.inertia_world, idx, b.inertia);  // kinematic: never written inside a region
        if (NeedsPose) load_pose(B.pose, idx, b.pos, b.q);
    } else {
        load_inertia_stamped(B.inertia_world, idx, b.inertia, stamp, error_flag);
        if (NeedsPose) {
            if (STAGE == kStageWarmStart) load_pose_stamped(B.pose, idx, b.pos, b.q, stamp, error_flag);
            else load_pose(B.pose, idx, b.pos, b.q);
        }
        if (STAGE == kStageWarmStartFirst && fp.angular_mode != 0 && (enc & kRefBundleIntegratesBit)) {
            // reference quirk of the momentum-conserving modes (see warm_start_body)
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

// Everything a lane does for its constraint in one pass; shaped like run_lane (bepu_solver_kernels.cuh) so that the register allocation of the math
// is the stage kernels'. Inlined into the type switch of the kernel (no calls: the ABI's register save/restore and by-reference arguments would live
// in local memory). Order: static words (references, chain, successor) -> wait for this pass's notifications -> ALL body records gathered in one
// round trip -> tag checks (re-read only what was early) -> math -> versioned stores -> notifications.
template <class T, int STAGE>
BEPU_DI void run_lane_dataflow(const WorkRecord& rec, int lane, long long chain_delta, long long succ_delta, unsigned int* counters, unsigned int* my_counter, unsigned int target,
                               unsigned int first, uint32_t slab_addr, uint32_t bar, uint32_t parity, uint32_t prestep_bytes, const BodyBuffers& B, const FrameParams& fp,
                               uint32_t pass_index, uint32_t ws_stamp, bool pose_stamped, int32_t* error_flag) {
    constexpr int NB = T::kBodies;
    const int32_t* refs = rec.refs + lane;
    const StagedRows p{slab_addr + lane * 4, bar, parity};
    const StagedAcc a{slab_addr + prestep_bytes + lane * 4, rec.impulses + lane};
    uint32_t enc[NB], version[NB];
    int32_t succ[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        enc[s] = ldg_nc_u32(refs + s * kLanes);
        const uint32_t chain = ldg_nc_u32(refs + chain_delta + s * kLanes);
        succ[s] = (int32_t)ldg_nc_u32(refs + succ_delta + s * kLanes);
        version[s] = pass_index * (chain >> kChainDegreeShift) + (chain & kChainRankMask);  // expected now; + 1 is what this lane publishes
    }
    // wait for the notifications of this pass: one 4-byte poll for the whole warp
    if (target != first) {
        unsigned int spins = 0;
        while ((int)(ld_relaxed_u32(my_counter) - target) < 0) {
            if (++spins > kDataflowSpinLimit || ((spins & 1023u) == 0u && *reinterpret_cast<volatile int32_t*>(error_flag) == 4)) {
                if (lane == 0) report_stall(error_flag, 1, target, ld_relaxed_u32(my_counter), (uint32_t)(my_counter - counters));  // results are void; drain quickly
                return;
            }
            if (spins > 2) __nanosleep(32);
        }
        if (lane == 0) *my_counter = first;  // all of this pass's notifications are in: ready for the next pass (which starts after a kernel boundary)
    }
    if ((int32_t)enc[0] == kRefEmpty) return;  // trailing lane of the last bundle, or a hole in a fallback bundle
    // gather: every record this lane needs, issued back to back
    F8 rv[NB], ri[NB], rp[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        const uint32_t idx = enc[s] & kRefIndexMask;
        rv[s] = ld256_strong(B.velocity + 2 * (size_t)idx);
        if constexpr (STAGE == kStageSolve) {
            ri[s] = ld256_strong(B.inertia_world + 2 * (size_t)idx);
            if (T::kNeedsPose) rp[s] = ld256_strong(B.pose + 2 * (size_t)idx);
        }
    }
    BodyState b[NB];
    Velocity v[NB];
    bool ok = true;
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        const uint32_t idx = enc[s] & kRefIndexMask;
        const bool dynamic = !(enc[s] & kRefKinematicBit);  // kinematic records are read-only inside a pass: nothing to check
        if (dynamic && (__float_as_uint(rv[s].d) != version[s] || __float_as_uint(rv[s].h) != version[s]))
            ok &= reload_until(B.velocity + 2 * (size_t)idx, version[s], true, rv[s], 2, idx, error_flag);
        v[s].lin = {rv[s].a, rv[s].b, rv[s].c};
        v[s].ang = {rv[s].e, rv[s].f, rv[s].g};
        if constexpr (STAGE == kStageSolve) {
            if (dynamic && __float_as_uint(ri[s].h) != ws_stamp) ok &= reload_until(B.inertia_world + 2 * (size_t)idx, ws_stamp, false, ri[s], 3, idx, error_flag);
            b[s].inertia.t = {ri[s].a, ri[s].b, ri[s].c, ri[s].d, ri[s].e, ri[s].f};
            b[s].inertia.inv_mass = ri[s].g;
            if (T::kNeedsPose) {
                if (dynamic && pose_stamped && __float_as_uint(rp[s].h) != ws_stamp) ok &= reload_until(B.pose + 2 * (size_t)idx, ws_stamp, false, rp[s], 4, idx, error_flag);
                b[s].q = {rp[s].a, rp[s].b, rp[s].c, rp[s].d};
                b[s].pos = {rp[s].e, rp[s].f, rp[s].g};
            }
        }
    }
    if (!ok) return;  // a dependency never arrived: the error flag is set, results are void
    rows_ready(p);
    if constexpr (STAGE == kStageSolve) {
        call_solve<T>(b, fp.dt, fp.inverse_dt, p, a, v);
    } else {
#pragma unroll
        for (int s = 0; s < NB; ++s) warm_start_body_dataflow<STAGE, T::kNeedsPose>(enc[s], B, fp, b[s], v[s], ws_stamp, error_flag);
        call_warm_start<T>(b, p, a, v);
    }
#pragma unroll
    for (int s = 0; s < NB; ++s)
        if (!(enc[s] & kRefKinematicBit)) store_velocity_versioned(B.velocity, enc[s] & kRefIndexMask, v[s], version[s] + 1u);
    if (fp.tune[2]) asm volatile("fence.acq_rel.gpu;" ::: "memory");  // development knob (A/B): release between the records and the notifications
#pragma unroll
    for (int s = 0; s < NB; ++s)
        if (!(enc[s] & kRefKinematicBit) && succ[s] >= 0) red_add_u32(counters + succ[s], 1u);  // the body's last constraint of the pass has nobody to wake
}

// kContactsOnly: the type switch holds the 14 contact types only (scenes without joints: shorter code, lower register pressure).
#define BEPU_DATAFLOW_ARGS rec, lane, df.chain_delta, df.succ_delta, df.counters, my_counter, target, first, slab_addr, bar, parity, prestep_bytes, B, fp, pass_index, ws_stamp, pose_stamped, error_flag
template <int STAGE, bool kContactsOnly>
BEPU_DI void run_bundle_dataflow(const WorkRecord& rec, int lane, const DataflowTables& df, unsigned int* my_counter, unsigned int target, unsigned int first, uint32_t slab_addr, uint32_t bar,
                                 uint32_t parity, uint32_t prestep_bytes, const BodyBuffers& B, const FrameParams& fp, uint32_t pass_index, uint32_t ws_stamp, bool pose_stamped,
                                 int32_t* error_flag) {
    switch (rec.type_id) {
#define BEPU_CASE(ID, T) \
    case ID: run_lane_dataflow<T, STAGE>(BEPU_DATAFLOW_ARGS); break;
        BEPU_CONTACT_TYPES(BEPU_CASE)
#undef BEPU_CASE
        default:
            if constexpr (!kContactsOnly) {
                switch (rec.type_id) {
#define BEPU_CASE(ID, T) \
    case ID: run_lane_dataflow<T, STAGE>(BEPU_DATAFLOW_ARGS); break;
                    BEPU_JOINT_TYPES(BEPU_CASE)
                    BEPU_JOINT_TYPES_MORE(BEPU_CASE)
#undef BEPU_CASE
                    default: break;
                }
            }
            break;
    }
}

// One WarmStart or Solve pass over the whole active set (all device batches) in ONE cooperative launch: inside the pass the Gauss-Seidel order per
// body is kept by the version / notification protocol above; passes are separated by kernel boundaries, so at the start of a pass every body is at
// version pass * degree and the first constraint on each body needs no notification. A bundle's counter therefore starts each pass at `first`
// (dep_counts.y) and must reach dep_counts.x; the warp that consumed it resets it for the next pass.
template <int STAGE, bool kContactsOnly>
__global__ void __launch_bounds__(kDataflowThreads, BEPU_DATAFLOW_MINB)
dataflow_pass_kernel(const WorkRecord* __restrict__ records, int work_count, DataflowTables df, BodyBuffers B, const FrameParams* __restrict__ fpp, uint32_t pass_offset,
                     uint32_t ws_pass_offset, int pose_stamped, int32_t* error_flag) {
    extern __shared__ __align__(128) unsigned char dataflow_smem[];
    const FrameParams fp = *fpp;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int total_warps = gridDim.x * kDataflowWarps;
    const uint32_t slab_addr = smem_u32(dataflow_smem) + warp_in_block * kStageSlabBytes;
    const uint32_t bar = smem_u32(dataflow_smem + kDataflowWarps * kStageSlabBytes + warp_in_block * 8);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    uint32_t parity = 0;
    const uint32_t pass_index = fp.pass_base + pass_offset;       // version pass since the body versions were reset
    const uint32_t ws_stamp = fp.pass_base + ws_pass_offset + 1u;  // stamp of this substep's WarmStart pass
    const uint64_t policy = l2_evict_first_policy();
    for (int g = warp_in_block * gridDim.x + blockIdx.x; g < work_count; g += total_warps) {
        // The slab address is loop-invariant; laundering it keeps the compiler from hoisting every row address of every type out of this loop
        // (hundreds of live registers, spilled), which the single-bundle stage kernels never suffer from.
        uint32_t slab_it = slab_addr, bar_it = bar;
        asm volatile("" : "+r"(slab_it), "+r"(bar_it));
        const WorkRecord rec = load_record(records + g);
        const int2 deps = __ldg(df.dep_counts + g);
        unsigned int* my_counter = df.counters + g;
        const uint32_t prestep_bytes = kStageRowCounts.prestep[rec.type_id] * (kLanes * 4);
        const uint32_t impulse_bytes = kStageRowCounts.impulses[rec.type_id] * (kLanes * 4);
        __syncwarp();  // every lane is done with the slab's previous contents
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar_it, prestep_bytes + impulse_bytes);
            bulk_copy_g2s(slab_it, rec.prestep, prestep_bytes, bar_it, policy);
            bulk_copy_g2s(slab_it + prestep_bytes, rec.impulses, impulse_bytes, bar_it, policy);
        }
        run_bundle_dataflow<STAGE, kContactsOnly>(rec, lane, df, my_counter, (unsigned int)deps.x, (unsigned int)deps.y, slab_it, bar_it, parity, prestep_bytes, B, fp, pass_index, ws_stamp,
                                                  pose_stamped != 0, error_flag);
        parity ^= 1u;
    }
}

template <int STAGE, bool kContactsOnly>
static int launch_dataflow_pass_t(const WorkRecord* records, int work_count, const DataflowTables& df, const BodyBuffers& B, const FrameParams* fp, uint32_t pass_offset,
                                  uint32_t ws_pass_offset, int pose_stamped, int32_t* error_flag, int blocks_per_sm, cudaStream_t s) {
    static int grid_limit[64] = {};
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return (int)e;
    auto kernel = dataflow_pass_kernel<STAGE, kContactsOnly>;
    if (grid_limit[device & 63] == 0) {
        int sms = 0, max_per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDataflowSmemBytes);
        // Maximum shared-memory carveout: co-residency of the whole grid is what makes a pass deadlock-free, and a smaller preferred carveout that
        // "just fits" two CTAs per SM was observed to leave the second CTA of every SM unscheduled (the cooperative-launch check uses the
        // occupancy calculator, which assumes the largest carveout).
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, kernel, kDataflowThreads, kDataflowSmemBytes);
        if (e != cudaSuccess) return (int)e;
        if (max_per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
        grid_limit[device & 63] = sms * 1024 + max_per_sm;
    }
    const int sms = grid_limit[device & 63] / 1024, max_per_sm = grid_limit[device & 63] % 1024;
    int per_sm = blocks_per_sm <= 0 ? max_per_sm : (blocks_per_sm < max_per_sm ? blocks_per_sm : max_per_sm);
    int grid = sms * per_sm;  // every CTA must be resident: the pass deadlocks otherwise (cooperative launch enforces it)
    const int needed = (work_count + kDataflowWarps - 1) / kDataflowWarps;
    if (grid > needed) grid = needed;
    if (grid < 1) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kDataflowThreads);
    cfg.dynamicSmemBytes = kDataflowSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag);
}
template <bool kContactsOnly>
static int launch_dataflow_pass_c(int stage, const WorkRecord* records, int work_count, const DataflowTables& df, const BodyBuffers& B, const FrameParams* fp, uint32_t pass_offset,
                                  uint32_t ws_pass_offset, int pose_stamped, int32_t* error_flag, int blocks_per_sm, cudaStream_t s) {
    switch (stage) {
        case kStageWarmStartFirst: return launch_dataflow_pass_t<kStageWarmStartFirst, kContactsOnly>(records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag, blocks_per_sm, s);
        case kStageWarmStart: return launch_dataflow_pass_t<kStageWarmStart, kContactsOnly>(records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag, blocks_per_sm, s);
        default: return launch_dataflow_pass_t<kStageSolve, kContactsOnly>(records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag, blocks_per_sm, s);
    }
}
static int launch_dataflow_pass(int stage, const WorkRecord* records, int work_count, const DataflowTables& df, const BodyBuffers& B, const FrameParams* fp, uint32_t pass_offset,
                                uint32_t ws_pass_offset, int pose_stamped, int32_t* error_flag, int blocks_per_sm, int contacts_only, cudaStream_t s) {
    if (contacts_only) return launch_dataflow_pass_c<true>(stage, records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag, blocks_per_sm, s);
#ifdef BEPU_DATAFLOW_DEV_CONTACTS_ONLY
    return (int)cudaErrorNotSupported;  // development builds compile the contacts-only instantiations alone
#else
    return launch_dataflow_pass_c<false>(stage, records, work_count, df, B, fp, pass_offset, ws_pass_offset, pose_stamped, error_flag, blocks_per_sm, s);
#endif
}

}  // namespace BEPU_NS
