// Dataflow persistent solver kernel (BEPUCUDA_EXEC_DATAFLOW).
//
// A (batch, stage) of a 100 k-body scene is half a wave of warps, so with one grid-wide barrier (or kernel boundary) per (batch, stage) a frame is a
// chain of ~390 latency-bound steps. But a constraint in batch k only depends on the (at most one per body) constraints of earlier batches that
// touch ITS bodies. This kernel keeps exactly those dependencies and drops the rest:
//
//   * every dynamic body carries a version counter in the two padding words of its 32-byte velocity record (one in each 16-byte half, so a torn
//     read is detected); a version counts the constraint evaluations that have written the body since the last reset;
//   * a lane that is the r-th of the K constraints on body X (device batch order: "rank", "degree") expects version P*K + r at WarmStart/Solve pass
//     P (passes counted from the reset) and publishes P*K + r + 1 together with the new velocity in ONE 256-bit store. Per body this replays
//     exactly the reference's Gauss-Seidel order (Solver_Solve.cs:L1447-1476), so results are bit-identical to the barrier schedule;
//   * warps own bundles statically (bundle g -> warp g mod T) and walk their bundles in program order pass by pass, so a warp is usually already
//     resident, has its work record, references and prestep in flight, and is polling L2 when its producers publish: a dependency link costs
//     (poll hit = the velocity gather itself) + compute + (store visible in L2), with no kernel boundary and no grid barrier;
//   * progress: the earliest unfinished evaluation in program order never waits (all its producers are earlier), and every warp reaches its items in
//     program order, so with all CTAs co-resident (cooperative launch) the schedule cannot deadlock. A spin limit raises an error flag instead of hanging.
//
// Grid barriers remain only where the reference has whole-set passes: IncrementallyUpdateForSubstep, the kinematic prepass, the final pose pass
// (3 per substep instead of 49 at 16 batches x (1 + 2 iterations)).
#pragma once
#include "bepu_persistent.cuh"

namespace BEPU_NS {

constexpr unsigned int kDataflowSpinLimit = 4000000u;  // ~1 s of polling: a dependency that never arrives is a bug, not a reason to hang the GPU

BEPU_DI void store_velocity_versioned(float4* vel, uint32_t i, const Velocity& v, uint32_t version) {
    const float ver = __uint_as_float(version);
    st256(vel + 2 * (size_t)i, v.lin.x, v.lin.y, v.lin.z, ver, v.ang.x, v.ang.y, v.ang.z, ver);
}

// Out of line per (type, stage): each gets its own register allocation (92-113 registers, no spills). Inlined into one switch, ptxas spilled ~7 KB.
template <class T, int STAGE>
__device__ __noinline__ void run_lane_dataflow(const WorkRecord& rec, int lane, long long chain_delta, const BodyBuffers& B, const FrameParams& fp, uint32_t pass_index, int32_t* error_flag) {
    constexpr int NB = T::kBodies;
    const int32_t* refs = rec.refs + lane;
    const GlobalRows p{rec.prestep + lane};
    const GlobalAcc a{rec.impulses + lane};
    uint32_t enc[NB], expect[NB];
    bool dynamic[NB], ready[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) enc[s] = (uint32_t)__ldg(refs + s * kLanes);
    const bool empty = (int32_t)enc[0] == kRefEmpty;
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        const uint32_t chain = __ldg(reinterpret_cast<const uint32_t*>(refs + chain_delta) + s * kLanes);
        dynamic[s] = !empty && !(enc[s] & kRefKinematicBit);
        expect[s] = pass_index * (chain >> kChainDegreeShift) + (chain & kChainRankMask);
        ready[s] = !dynamic[s];
    }
    BodyState b[NB];
    Velocity v[NB];
    if constexpr (STAGE == kStageSolve) {
        // World inertia and pose were written by this substep's WarmStart pass, which this warp already executed for this bundle: fetch them while waiting.
        if (!empty) {
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                const uint32_t idx = enc[s] & kRefIndexMask;
                load_inertia(B.inertia_world, idx, b[s].inertia);
                if (T::kNeedsPose) load_pose(B.pose, idx, b[s].pos, b[s].q);
            }
        }
    }
    // Wait for both bodies: the poll IS the velocity gather.
    unsigned int spins = 0;
    while (true) {
        bool all = true;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            if (!ready[s]) {
                const F8 r = ld256(B.velocity + 2 * (size_t)(enc[s] & kRefIndexMask));
                if (__float_as_uint(r.d) == expect[s] && __float_as_uint(r.h) == expect[s]) {
                    v[s].lin = {r.a, r.b, r.c};
                    v[s].ang = {r.e, r.f, r.g};
                    ready[s] = true;
                } else {
                    all = false;
                }
            }
        }
        if (__all_sync(0xffffffffu, all)) break;
        if (++spins > kDataflowSpinLimit) {
            if (!all) atomicExch(error_flag, 4);
            break;
        }
        // once any dependency has timed out the results are void anyway: stop waiting everywhere so the kernel drains quickly
        if ((spins & 255u) == 0u && __any_sync(0xffffffffu, *reinterpret_cast<volatile int32_t*>(error_flag) == 4)) break;
        if (spins > 4) __nanosleep(spins > 64 ? (fp.tune[1] > 0 ? fp.tune[1] : 200) : (fp.tune[0] > 0 ? fp.tune[0] : 40));
    }
    if (empty) return;
#pragma unroll
    for (int s = 0; s < NB; ++s)
        if (!dynamic[s]) load_velocity(B.velocity, enc[s] & kRefIndexMask, v[s]);  // kinematic: read-only inside a region
    if constexpr (STAGE == kStageSolve) {
        call_solve<T>(b, fp.dt, fp.inverse_dt, p, a, v);
    } else {
        bool owner = false;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
            warm_start_body<STAGE, T::kNeedsPose>(enc[s], B, fp, b[s], v[s]);
            owner = owner || (enc[s] & kRefIntegrateBit);
        }
        call_warm_start<T>(b, p, a, v);
        // The owner's pose / world inertia stores must be visible before the version that lets the next constraint on the body read them.
        if (owner) __threadfence();
    }
#pragma unroll
    for (int s = 0; s < NB; ++s)
        if (dynamic[s]) store_velocity_versioned(B.velocity, enc[s] & kRefIndexMask, v[s], expect[s] + 1u);
}

template <int STAGE>
BEPU_DI void run_bundle_dataflow(const WorkRecord* __restrict__ record, int lane, long long chain_delta, BodyBuffers B, const FrameParams* __restrict__ fpp, uint32_t pass_index,
                                                 int32_t* error_flag) {
    const WorkRecord rec = load_record(record);
    const FrameParams fp = *fpp;
    switch (rec.type_id) {
#define BEPU_CASE(ID, T) \
    case ID: run_lane_dataflow<T, STAGE>(rec, lane, chain_delta, B, fp, pass_index, error_flag); break;
        BEPU_CONTACT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES_MORE(BEPU_CASE)
#undef BEPU_CASE
        default: break;
    }
}

static __global__ void __launch_bounds__(kPersistentThreads, 2)
dataflow_solve_kernel(const StageOp* __restrict__ program, int op_count, const WorkRecord* __restrict__ records, long long chain_delta, const int32_t* __restrict__ kinematics,
                      BodyBuffers B, const FrameParams* __restrict__ fpp, unsigned int* barrier_counter, int32_t* error_flag) {
    const FrameParams fp = *fpp;
    constexpr int kWarpsPerBlock = kPersistentThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int total_warps = gridDim.x * kWarpsPerBlock;
    const int first_warp_item = warp_in_block * gridDim.x + blockIdx.x;
    const int total_threads = gridDim.x * kPersistentThreads;
    const int first_thread_item = threadIdx.x * gridDim.x + blockIdx.x;
    unsigned int barrier_target = 0;
    uint32_t pass_counter = fp.pass_base;
    for (int op_index = 0; op_index < op_count; ++op_index) {
        const StageOp op = program[op_index];
        switch (op.stage) {
            case kStageRegion: {
                const int solve_passes = op.pad >> 1;
                const bool first_substep = (op.pad & 1) != 0;
                for (int pass = 0; pass <= solve_passes; ++pass) {
                    const uint32_t pass_index = pass_counter + (uint32_t)pass;
                    for (int g = first_warp_item; g < op.work_count; g += total_warps) {
                        const WorkRecord* rec = records + op.work_begin + g;
                        if (pass > 0) run_bundle_dataflow<kStageSolve>(rec, lane, chain_delta, B, fpp, pass_index, error_flag);
                        else if (first_substep) run_bundle_dataflow<kStageWarmStartFirst>(rec, lane, chain_delta, B, fpp, pass_index, error_flag);
                        else run_bundle_dataflow<kStageWarmStart>(rec, lane, chain_delta, B, fpp, pass_index, error_flag);
                    }
                }
                pass_counter += (uint32_t)solve_passes + 1u;
                break;
            }
            case kStageIncremental:
                for (int i = first_warp_item; i < op.work_count; i += total_warps) run_bundle<kStageIncremental>(load_record(records + op.work_begin + i), lane, B, fp);
                break;
            case kStageKinematicFirst:
                for (int i = first_thread_item; i < op.work_count; i += total_threads) run_kinematic<kStageKinematicFirst>(i, kinematics, B, fp);
                break;
            case kStageKinematic:
                for (int i = first_thread_item; i < op.work_count; i += total_threads) run_kinematic<kStageKinematic>(i, kinematics, B, fp);
                break;
            case kStageFinalPose:
                for (int i = blockIdx.x * kPersistentThreads + threadIdx.x; i < B.count; i += total_threads) run_final_pose(i, B, fp);
                break;
            default: break;
        }
        if (op_index + 1 < op_count) {
            barrier_target += gridDim.x;
            grid_barrier(barrier_counter, barrier_target);
        }
    }
}

static int launch_dataflow(const StageOp* program, int op_count, const WorkRecord* records, long long chain_delta, const int32_t* kinematics, const BodyBuffers& B, const FrameParams* fp,
                           unsigned int* barrier_counter, int32_t* error_flag, int blocks_per_sm, cudaStream_t s) {
    static int sms = 0, max_per_sm = 0;
    if (sms == 0) {
        int device = 0;
        cudaError_t e = cudaGetDevice(&device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, dataflow_solve_kernel, kPersistentThreads, 0);
        if (e != cudaSuccess) return (int)e;
    }
    if (max_per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    int per_sm = blocks_per_sm <= 0 ? 1 : blocks_per_sm;
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    const int grid = sms * per_sm;
    BodyBuffers Bc = B;
    void* args[] = {(void*)&program, (void*)&op_count, (void*)&records, (void*)&chain_delta, (void*)&kinematics, (void*)&Bc, (void*)&fp, (void*)&barrier_counter, (void*)&error_flag};
    return (int)cudaLaunchCooperativeKernel((const void*)dataflow_solve_kernel, dim3(grid), dim3(kPersistentThreads), args, 0, s);
}

}  // namespace BEPU_NS
