// Persistent cooperative solver kernel: the whole (substep, stage, batch) sequence of Solver.Solve
// (Solver_Solve.cs:L1419-1479) + the final pose pass in ONE launch, with a grid-wide barrier where the reference's
// multithreaded path has a sync point (Solver_Solve.cs:L395-401).
//
// Latency engineering (a stage of a 100k-body pile is ~650 warps: one thin wave, so a stage costs one warp's critical path):
//   - work items are dealt round-robin across CTAs so a small batch still spreads over all 148 SMs;
//   - each warp fetches its NEXT stage's 32-B work record and body references BEFORE arriving at the barrier (they are
//     immutable during a solve), so after the barrier the first thing issued is the body gather from L2;
//   - the barrier is a single monotonically increasing counter: one red.add per CTA, ld.acquire polling, no reset phase.
#pragma once
#include "bepu_solver_kernels.cuh"

namespace BEPU_NS {

constexpr int kPersistentThreads = 256;

BEPU_DI void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // make this CTA's scatters visible at gpu scope before signalling
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

BEPU_DI bool is_constraint_stage(int stage) { return stage <= kStageIncremental; }

// The kernel itself is only compiled in its own unit (bepu_dataflow.cuh includes this header for the barrier helpers).
#if !defined(BEPU_UNIT) || BEPU_UNIT == 4
template <int STAGE>
BEPU_DI void run_constraint_op(const StageOp& op, const WorkRecord* __restrict__ records, bool prefetched, const WorkRecord& rec0, uint32_t enc0, uint32_t enc1, int first_item, int stride,
                               int lane, const BodyBuffers& B, const FrameParams& fp) {
    if (first_item < op.work_count) {
        if (prefetched) run_bundle<STAGE>(rec0, lane, enc0, enc1, B, fp);
        else run_bundle<STAGE>(load_record(records + op.work_begin + first_item), lane, B, fp);
        for (int i = first_item + stride; i < op.work_count; i += stride) run_bundle<STAGE>(load_record(records + op.work_begin + i), lane, B, fp);
    }
}

static __global__ void __launch_bounds__(kPersistentThreads, 2)
persistent_solve_kernel(const StageOp* __restrict__ program, int op_count, const WorkRecord* __restrict__ records, const int32_t* __restrict__ kinematics, BodyBuffers B,
                        const FrameParams* __restrict__ fpp, unsigned int* barrier_counter) {
    const FrameParams fp = *fpp;
    constexpr int kWarpsPerBlock = kPersistentThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int total_warps = gridDim.x * kWarpsPerBlock;
    const int first_warp_item = warp_in_block * gridDim.x + blockIdx.x;  // item i -> CTA (i % grid), warp (i / grid) % warpsPerBlock
    const int total_threads = gridDim.x * kPersistentThreads;
    const int first_thread_item = threadIdx.x * gridDim.x + blockIdx.x;
    unsigned int barrier_target = 0;

    StageOp op = program[0];
    WorkRecord rec{};
    uint32_t enc0 = 0, enc1 = 0;
    bool prefetched = false;
    for (int op_index = 0; op_index < op_count; ++op_index) {
        StageOp next{};
        const bool has_next = op_index + 1 < op_count;
        if (has_next) next = program[op_index + 1];
        switch (op.stage) {
            case kStageWarmStartFirst: run_constraint_op<kStageWarmStartFirst>(op, records, prefetched, rec, enc0, enc1, first_warp_item, total_warps, lane, B, fp); break;
            case kStageWarmStart: run_constraint_op<kStageWarmStart>(op, records, prefetched, rec, enc0, enc1, first_warp_item, total_warps, lane, B, fp); break;
            case kStageSolve: run_constraint_op<kStageSolve>(op, records, prefetched, rec, enc0, enc1, first_warp_item, total_warps, lane, B, fp); break;
            case kStageIncremental: run_constraint_op<kStageIncremental>(op, records, prefetched, rec, enc0, enc1, first_warp_item, total_warps, lane, B, fp); break;
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
        if (!has_next) break;
        // Prefetch the next stage's record + body references while the rest of the grid is still finishing this stage.
        prefetched = false;
        if (is_constraint_stage(next.stage) && first_warp_item < next.work_count) {
            rec = load_record(records + next.work_begin + first_warp_item);
            enc0 = (uint32_t)__ldg(rec.refs + lane);
            enc1 = (uint32_t)__ldg(rec.refs + kLanes + lane);
            prefetched = true;
        }
        barrier_target += gridDim.x;
        grid_barrier(barrier_counter, barrier_target);
        op = next;
    }
}

static int launch_persistent(const StageOp* program, int op_count, const WorkRecord* records, const int32_t* kinematics, const BodyBuffers& B, const FrameParams* fp,
                             unsigned int* barrier_counter, int blocks_per_sm, cudaStream_t s) {
    // Queried once per process (thread-safe static initialisation); every device of a node is the same part, so one answer serves all contexts.
    struct Occupancy { int sms = 0, max_per_sm = 0, error = 0; };
    static const Occupancy occ = [] {
        Occupancy o;
        int device = 0;
        cudaError_t e = cudaGetDevice(&device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&o.sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o.max_per_sm, persistent_solve_kernel, kPersistentThreads, 0);
        o.error = (int)e;
        return o;
    }();
    if (occ.error != 0) return occ.error;
    const int sms = occ.sms, max_per_sm = occ.max_per_sm;
    if (max_per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
    int per_sm = blocks_per_sm <= 0 ? 1 : blocks_per_sm;
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    const int grid = sms * per_sm;
    BodyBuffers Bc = B;
    void* args[] = {(void*)&program, (void*)&op_count, (void*)&records, (void*)&kinematics, (void*)&Bc, (void*)&fp, (void*)&barrier_counter};
    return (int)cudaLaunchCooperativeKernel((const void*)persistent_solve_kernel, dim3(grid), dim3(kPersistentThreads), args, 0, s);
}

#endif

}  // namespace BEPU_NS
