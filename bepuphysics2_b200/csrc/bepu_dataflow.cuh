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
//     gathers the velocity records. The counter is a wake-up hint, not the synchronisation: the gather still checks every record's version and
//     re-reads a record whose store has not landed yet, so no fence is needed between a producer's record store and its notification;
//   * warps own bundles statically (bundle g -> warp g mod T) and walk their bundles in program order pass by pass; while a warp waits it already
//     holds its bundle's prestep + impulse block in its shared-memory slab (one cp.async.bulk pair), its body references and -- in Solve passes --
//     the world inertias, so a dependency link costs (notification visible) + (velocity gather) + math + (store);
//   * progress: the earliest unfinished evaluation in program order never waits (all its producers are earlier), and every warp reaches its items in
//     program order, so with all CTAs co-resident (cooperative launch) the schedule cannot deadlock. A spin limit raises an error flag instead of hanging.
//
// Grid barriers remain only where the reference has whole-set passes: IncrementallyUpdateForSubstep, the kinematic prepass, the final pose pass.
#pragma once
#include "bepu_persistent.cuh"

namespace BEPU_NS {

constexpr unsigned int kDataflowSpinLimit = 20000000u;  // ~1 s of polling: a dependency that never arrives is a bug, not a reason to hang the GPU
constexpr int kDataflowWarps = kPersistentThreads / 32;
constexpr int kDataflowSmemBytes = kDataflowWarps * kStageSlabBytes + kDataflowWarps * 8;

BEPU_DI void store_velocity_versioned(float4* vel, uint32_t i, const Velocity& v, uint32_t version) {
    const float ver = __uint_as_float(version);
    st256(vel + 2 * (size_t)i, v.lin.x, v.lin.y, v.lin.z, ver, v.ang.x, v.ang.y, v.ang.z, ver);
}
BEPU_DI unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
BEPU_DI void red_add_u32(unsigned int* p, unsigned int v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// What a warp knows about its bundle before it waits.
struct DataflowBundle {
    WorkRecord rec;
    unsigned int* counters;      // base of the per-bundle notification counters
    unsigned int* my_counter;    // this bundle's
    unsigned int target;         // notifications that must have arrived before this pass may gather
    long long chain_delta, succ_delta;
    uint32_t slab_addr, bar, parity, prestep_bytes;
};

// Out of line per (type, stage): each gets its own register allocation. Inlined into one switch, ptxas spills.
template <class T, int STAGE>
__device__ __noinline__ void run_lane_dataflow(const DataflowBundle& w, int lane, const BodyBuffers& B, const FrameParams& fp, uint32_t pass_index, int32_t* error_flag) {
    constexpr int NB = T::kBodies;
    const int32_t* refs = w.rec.refs + lane;
    const StagedRows p{w.slab_addr + lane * 4, w.bar, w.parity};
    const StagedAcc a{w.slab_addr + w.prestep_bytes + lane * 4, w.rec.impulses + lane};
    uint32_t enc[NB], expect[NB];
    int32_t succ[NB];
    bool dynamic[NB], ready[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) enc[s] = (uint32_t)__ldg(refs + s * kLanes);
    const bool empty = (int32_t)enc[0] == kRefEmpty;
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        const uint32_t chain = __ldg(reinterpret_cast<const uint32_t*>(refs + w.chain_delta) + s * kLanes);
        succ[s] = __ldg(refs + w.succ_delta + s * kLanes);
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
    rows_ready(p);
    // 1. wait for the notifications of this pass (one 4-byte poll per warp)
    unsigned int spins = 0;
    bool failed = false;
    if (w.target != 0u) {
        while ((int)(ld_relaxed_u32(w.my_counter) - w.target) < 0) {
            if (++spins > kDataflowSpinLimit || ((spins & 1023u) == 0u && *reinterpret_cast<volatile int32_t*>(error_flag) == 4)) { failed = true; break; }
            if (spins > 2) __nanosleep(fp.tune[0] > 0 ? fp.tune[0] : 64);
        }
    }
    // 2. gather; a record whose store has not landed yet (the notification overtook it) is simply read again
    spins = 0;
    while (!failed) {
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
        if (++spins > kDataflowSpinLimit / 16) failed = true;
    }
    if (__any_sync(0xffffffffu, failed)) {
        atomicExch(error_flag, 4);  // results are void; drain quickly
        return;
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
        if (dynamic[s]) {
            store_velocity_versioned(B.velocity, enc[s] & kRefIndexMask, v[s], expect[s] + 1u);
            red_add_u32(w.counters + succ[s], 1u);
        }
}

template <int STAGE>
BEPU_DI void run_bundle_dataflow(const DataflowBundle& w, int lane, const BodyBuffers& B, const FrameParams& fp, uint32_t pass_index, int32_t* error_flag) {
    switch (w.rec.type_id) {
#define BEPU_CASE(ID, T) \
    case ID: run_lane_dataflow<T, STAGE>(w, lane, B, fp, pass_index, error_flag); break;
        BEPU_CONTACT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES(BEPU_CASE)
        BEPU_JOINT_TYPES_MORE(BEPU_CASE)
#undef BEPU_CASE
        default: break;
    }
}

static __global__ void __launch_bounds__(kPersistentThreads, 2)
dataflow_solve_kernel(const StageOp* __restrict__ program, int op_count, const WorkRecord* __restrict__ records, DataflowTables df, const int32_t* __restrict__ kinematics,
                      BodyBuffers B, const FrameParams* __restrict__ fpp, unsigned int* barrier_counter, int32_t* error_flag) {
    extern __shared__ __align__(128) unsigned char dataflow_smem[];
    const FrameParams fp = *fpp;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int total_warps = gridDim.x * kDataflowWarps;
    const int first_warp_item = warp_in_block * gridDim.x + blockIdx.x;
    const int total_threads = gridDim.x * kPersistentThreads;
    const int first_thread_item = threadIdx.x * gridDim.x + blockIdx.x;
    const uint32_t slab_addr = smem_u32(dataflow_smem) + warp_in_block * kStageSlabBytes;
    const uint32_t bar = smem_u32(dataflow_smem + kDataflowWarps * kStageSlabBytes + warp_in_block * 8);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    uint32_t parity = 0;
    unsigned int barrier_target = 0;
    uint32_t pass_counter = fp.pass_base;   // version passes since the body versions were reset
    uint32_t solve_pass = 0;                // notification passes since this solve started (the counters are reset per solve)
    const uint64_t policy = l2_evict_first_policy();
    for (int op_index = 0; op_index < op_count; ++op_index) {
        const StageOp op = program[op_index];
        switch (op.stage) {
            case kStageRegion: {
                const int solve_passes = op.pad >> 1;
                const bool first_substep = (op.pad & 1) != 0;
                for (int pass = 0; pass <= solve_passes; ++pass) {
                    const uint32_t pass_index = pass_counter + (uint32_t)pass;
                    for (int g = first_warp_item; g < op.work_count; g += total_warps) {
                        DataflowBundle w;
                        w.rec = load_record(records + op.work_begin + g);
                        const int2 deps = __ldg(df.dep_counts + op.work_begin + g);
                        w.counters = df.counters;
                        w.my_counter = df.counters + op.work_begin + g;
                        w.target = (solve_pass + (uint32_t)pass + 1u) * (unsigned int)deps.x;
                        w.chain_delta = df.chain_delta;
                        w.succ_delta = df.succ_delta;
                        w.slab_addr = slab_addr;
                        w.bar = bar;
                        w.parity = parity;
                        w.prestep_bytes = kStageRowCounts.prestep[w.rec.type_id] * (kLanes * 4);
                        const uint32_t impulse_bytes = kStageRowCounts.impulses[w.rec.type_id] * (kLanes * 4);
                        __syncwarp();  // every lane is done with the slab's previous contents
                        if (lane == 0) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            mbar_expect_tx(bar, w.prestep_bytes + impulse_bytes);
                            bulk_copy_g2s(slab_addr, w.rec.prestep, w.prestep_bytes, bar, policy);
                            bulk_copy_g2s(slab_addr + w.prestep_bytes, w.rec.impulses, impulse_bytes, bar, policy);
                        }
                        if (pass > 0) run_bundle_dataflow<kStageSolve>(w, lane, B, fp, pass_index, error_flag);
                        else if (first_substep) run_bundle_dataflow<kStageWarmStartFirst>(w, lane, B, fp, pass_index, error_flag);
                        else run_bundle_dataflow<kStageWarmStart>(w, lane, B, fp, pass_index, error_flag);
                        parity ^= 1u;
                    }
                }
                pass_counter += (uint32_t)solve_passes + 1u;
                solve_pass += (uint32_t)solve_passes + 1u;
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

static int launch_dataflow(const StageOp* program, int op_count, const WorkRecord* records, const DataflowTables& df, const int32_t* kinematics, const BodyBuffers& B, const FrameParams* fp,
                           unsigned int* barrier_counter, int32_t* error_flag, int blocks_per_sm, cudaStream_t s) {
    int device = 0, sms = 0, max_per_sm = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dataflow_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDataflowSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dataflow_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_per_sm, dataflow_solve_kernel, kPersistentThreads, kDataflowSmemBytes);
    if (e != cudaSuccess) return (int)e;
    if (max_per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    int per_sm = blocks_per_sm <= 0 ? 2 : blocks_per_sm;
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    const int grid = sms * per_sm;
    BodyBuffers Bc = B;
    DataflowTables dfc = df;
    void* args[] = {(void*)&program, (void*)&op_count, (void*)&records, (void*)&dfc, (void*)&kinematics, (void*)&Bc, (void*)&fp, (void*)&barrier_counter, (void*)&error_flag};
    return (int)cudaLaunchCooperativeKernel((const void*)dataflow_solve_kernel, dim3(grid), dim3(kPersistentThreads), args, kDataflowSmemBytes, s);
}

}  // namespace BEPU_NS
