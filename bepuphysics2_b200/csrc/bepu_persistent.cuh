// Persistent cooperative solver kernel: the whole (substep, stage, batch) sequence of Solver.Solve
// (Solver_Solve.cs:L1419-1479) + the final pose pass in ONE launch, with a grid-wide barrier where the reference's
// multithreaded path has a sync point (Solver_Solve.cs:L395-401). One CTA set stays resident on all 148 SMs; work items
// of a stage are dealt round-robin across CTAs so a small batch still spreads over the whole chip.
#pragma once
#include "bepu_solver_kernels.cuh"

namespace BEPU_NS {

constexpr int kPersistentThreads = 256;

// Sense-free generation barrier: state[0] = arrival count, state[1] = generation. Same fence pattern as cooperative
// groups' grid.sync(): block barrier, one thread publishes with a gpu-scope fence + atomic, spins on the generation,
// fences again, block barrier. Requires all CTAs co-resident (cooperative launch).
BEPU_DI void grid_barrier(unsigned int* state) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* gen_ptr = state + 1;
        const unsigned int gen = *gen_ptr;
        __threadfence();
        const unsigned int prev = atomicAdd(state, 1u);
        if (prev == gridDim.x - 1) {
            state[0] = 0;
            __threadfence();
            atomicAdd(state + 1, 1u);
        } else {
            while (*gen_ptr == gen) { __nanosleep(20); }
        }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kPersistentThreads, 1)
persistent_solve_kernel(const StageOp* __restrict__ program, int op_count, const DeviceTypeBatch* __restrict__ tbs, const WorkItem* __restrict__ work,
                        const int32_t* __restrict__ kinematics, BodyBuffers B, const FrameParams* __restrict__ fpp, unsigned int* barrier_state) {
    const FrameParams fp = *fpp;
    constexpr int kWarpsPerBlock = kPersistentThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int total_warps = gridDim.x * kWarpsPerBlock;
    const int first_warp_item = warp_in_block * gridDim.x + blockIdx.x;  // item i -> CTA (i % grid), warp (i / grid) % warpsPerBlock
    const int total_threads = gridDim.x * kPersistentThreads;
    const int first_thread_item = threadIdx.x * gridDim.x + blockIdx.x;
    for (int op_index = 0; op_index < op_count; ++op_index) {
        const StageOp op = program[op_index];
        switch (op.stage) {
            case kStageWarmStartFirst:
                for (int i = first_warp_item; i < op.work_count; i += total_warps) {
                    const WorkItem w = work[op.work_begin + i];
                    run_bundle<kStageWarmStartFirst>(tbs[w.type_batch], w.bundle, lane, B, fp);
                }
                break;
            case kStageWarmStart:
                for (int i = first_warp_item; i < op.work_count; i += total_warps) {
                    const WorkItem w = work[op.work_begin + i];
                    run_bundle<kStageWarmStart>(tbs[w.type_batch], w.bundle, lane, B, fp);
                }
                break;
            case kStageSolve:
                for (int i = first_warp_item; i < op.work_count; i += total_warps) {
                    const WorkItem w = work[op.work_begin + i];
                    run_bundle<kStageSolve>(tbs[w.type_batch], w.bundle, lane, B, fp);
                }
                break;
            case kStageIncremental:
                for (int i = first_warp_item; i < op.work_count; i += total_warps) {
                    const WorkItem w = work[op.work_begin + i];
                    run_bundle<kStageIncremental>(tbs[w.type_batch], w.bundle, lane, B, fp);
                }
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
        if (op_index + 1 < op_count) grid_barrier(barrier_state);
    }
}

static int persistent_grid_size() {
    int device = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_solve_kernel, kPersistentThreads, 0) != cudaSuccess) return 0;
    return sms * (per_sm < 1 ? 1 : per_sm);
}

static int launch_persistent(const StageOp* program, int op_count, const DeviceTypeBatch* tbs, const WorkItem* work, const int32_t* kinematics, const BodyBuffers& B,
                             const FrameParams* fp, unsigned int* barrier_state, cudaStream_t s) {
    static int grid = 0;
    if (grid == 0) grid = persistent_grid_size();
    if (grid <= 0) return (int)cudaErrorLaunchFailure;
    BodyBuffers Bc = B;
    void* args[] = {(void*)&program, (void*)&op_count, (void*)&tbs, (void*)&work, (void*)&kinematics, (void*)&Bc, (void*)&fp, (void*)&barrier_state};
    return (int)cudaLaunchCooperativeKernel((const void*)persistent_solve_kernel, dim3(grid), dim3(kPersistentThreads), args, 0, s);
}

}  // namespace BEPU_NS
