// Compiled twice: -DBEPU_NS=bepu_fast (default FMA contraction) and -DBEPU_NS=bepu_strict -fmad=false.
#include "bepu_solver_kernels.cuh"
#include "bepu_persistent.cuh"
#include "bepu_dataflow.cuh"
#include "bepu_layout_kernels.h"

namespace BEPU_NS {

template <int STAGE>
static void launch_stage_t(const WorkRecord* records, int work_count, const BodyBuffers& B, const FrameParams* fp, bool pdl, cudaStream_t s) {
    const unsigned blocks = (unsigned)(((size_t)work_count * 32 + kStageBlockThreads - 1) / kStageBlockThreads);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kStageBlockThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, constraint_stage_kernel<STAGE>, records, work_count, B, fp);
}
static void launch_constraint_stage(int stage, const WorkRecord* records, int work_count, const BodyBuffers& B, const FrameParams* fp, bool pdl, cudaStream_t s) {
    if (work_count <= 0) return;
    switch (stage) {
        case kStageWarmStartFirst: launch_stage_t<kStageWarmStartFirst>(records, work_count, B, fp, pdl, s); break;
        case kStageWarmStart: launch_stage_t<kStageWarmStart>(records, work_count, B, fp, pdl, s); break;
        case kStageSolve: launch_stage_t<kStageSolve>(records, work_count, B, fp, pdl, s); break;
        case kStageIncremental: launch_stage_t<kStageIncremental>(records, work_count, B, fp, pdl, s); break;
        default: break;
    }
}
static void launch_kinematic_stage(int stage, const int32_t* kinematics, int count, const BodyBuffers& B, const FrameParams* fp, cudaStream_t s) {
    if (count <= 0) return;
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (stage == kStageKinematicFirst) kinematic_stage_kernel<kStageKinematicFirst><<<blocks, 128, 0, s>>>(kinematics, count, B, fp);
    else kinematic_stage_kernel<kStageKinematic><<<blocks, 128, 0, s>>>(kinematics, count, B, fp);
}
static void launch_final_pose(const BodyBuffers& B, const FrameParams* fp, cudaStream_t s) {
    if (B.count <= 0) return;
    final_pose_kernel<<<(unsigned)((B.count + 255) / 256), 256, 0, s>>>(B, fp);
}

static const bepucuda::SolverLaunchers kLaunchers = {&launch_constraint_stage, &launch_kinematic_stage, &launch_final_pose, &launch_persistent, &launch_dataflow};

}  // namespace BEPU_NS

namespace bepucuda {
#define BEPU_CAT2(a, b) a##b
#define BEPU_CAT(a, b) BEPU_CAT2(a, b)
const SolverLaunchers* BEPU_CAT(get_launchers_, BEPU_NS)() { return &BEPU_NS::kLaunchers; }
}  // namespace bepucuda
