// Micro-benchmark: cost of one grid-wide barrier inside a persistent cooperative kernel on B200, for the barrier designs considered for the
// persistent solver (tools only; not part of the library). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// V0: monotonic counter, every CTA polls the counter line (the library's current barrier).
__device__ __forceinline__ void barrier_v0(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire(counter) < target) {}
    }
    __syncthreads();
}
// V1: red.release arrive, same polling.
__device__ __forceinline__ void barrier_v1(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release(counter, 1u);
        while (ld_acquire(counter) < target) {}
    }
    __syncthreads();
}
// V2: two-level arrive (groups of `group` CTAs), last arriver publishes the epoch to one flag line per group; CTAs poll their group's flag.
// layout: words [0]: top counter; [32 * (1 + g)]: group counter g; [32 * (1 + G + g)]: flag g   (one 128-B line each)
__device__ __forceinline__ void barrier_v2(unsigned* mem, unsigned epoch, int group, int groups) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const int g = blockIdx.x / group;
        const int members = min(group, (int)gridDim.x - g * group);
        unsigned* group_counter = mem + 32 * (1 + g);
        unsigned* flag = mem + 32 * (1 + groups + g);
        const unsigned old = atom_add_acq_rel(group_counter, 1u);
        if (old == epoch * members - 1) {  // last of the group this epoch
            const unsigned top_old = atom_add_acq_rel(mem, 1u);
            if (top_old == epoch * groups - 1) {  // last group: release everyone
                for (int i = 0; i < groups; ++i) st_release(mem + 32 * (1 + groups + i), epoch);
            }
        }
        while (ld_acquire(flag) < epoch) {}
    }
    __syncthreads();
}
// V3: like V2 but the flags are written by a whole warp of the last CTA (parallel stores) and pollers use relaxed loads + one fence.
__device__ __forceinline__ void barrier_v3(unsigned* mem, unsigned epoch, int group, int groups) {
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = blockIdx.x / group;
        const int members = min(group, (int)gridDim.x - g * group);
        unsigned last = 0;
        if (threadIdx.x == 0) {
            const unsigned old = atom_add_acq_rel(mem + 32 * (1 + g), 1u);
            if (old == epoch * members - 1) {
                const unsigned top_old = atom_add_acq_rel(mem, 1u);
                last = top_old == epoch * groups - 1;
            }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last)
            for (int i = threadIdx.x; i < groups; i += 32) st_release(mem + 32 * (1 + groups + i), epoch);
        if (threadIdx.x == 0) {
            const unsigned* flag = mem + 32 * (1 + groups + g);
            while (ld_relaxed(flag) < epoch) {}
            __threadfence();
        }
    }
    __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(256) bench_kernel(unsigned* mem, float* data, int iters, int group, int groups, int work) {
    cg::grid_group grid = cg::this_grid();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = gridDim.x * blockDim.x;
    float acc = 0.f;
    for (int i = 1; i <= iters; ++i) {
        if (work) {  // one dependent gather + scatter to another thread's slot, like a solver stage
            const int src = (gid * 7 + i * 13) % total;
            acc = __ldcg(data + src) * 0.5f + 1.0f;
            __stcg(data + gid, acc);
        }
        if (V == 0) barrier_v0(mem, (unsigned)i * gridDim.x);
        else if (V == 1) barrier_v1(mem, (unsigned)i * gridDim.x);
        else if (V == 2) barrier_v2(mem, (unsigned)i, group, groups);
        else if (V == 3) barrier_v3(mem, (unsigned)i, group, groups);
        else grid.sync();
    }
    if (acc == 12345.f) data[0] = acc;
}

template <int V> void run(const char* name, int blocks_per_sm, int group, int work, unsigned* mem, float* data, int sms) {
    const int grid = sms * blocks_per_sm;
    const int groups = (grid + group - 1) / group;
    int iters = 2000;
    void* args[] = {(void*)&mem, (void*)&data, (void*)&iters, (void*)&group, (void*)&groups, (void*)&work};
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(mem, 0, 32 * 4 * (1 + 2 * 1024)));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)bench_kernel<V>, dim3(grid), dim3(256), args, 0, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    printf("%-28s ctas/sm=%d group=%3d work=%d : %.3f us per barrier\n", name, blocks_per_sm, group, work, best * 1000.f / iters);
}

__global__ void __launch_bounds__(64) chain_kernel(float* data, int work, int use_pdl) {
    if (use_pdl) asm volatile("griddepcontrol.launch_dependents;");
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (use_pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (work) {
        const int total = gridDim.x * blockDim.x;
        float v = __ldcg(data + (gid * 7 + 13) % total) * 0.5f + 1.0f;
        __stcg(data + gid, v);
    }
}
// A CUDA graph of `n` dependent kernels (the library's GRAPH mode shape): time per node, with and without programmatic dependent launch.
void run_chain(int blocks, int work, int use_pdl, float* data) {
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const int n = 400;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; ++i) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(64);
        cfg.stream = s;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = use_pdl ? 1 : 0;
        CK(cudaLaunchKernelEx(&cfg, chain_kernel, data, work, use_pdl));
    }
    CK(cudaStreamEndCapture(s, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, s));
        CK(cudaGraphLaunch(exec, s));
        CK(cudaEventRecord(b, s));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    printf("graph chain blocks=%5d work=%d pdl=%d : %.3f us per node\n", blocks, work, use_pdl, best * 1000.f / n);
    CK(cudaGraphExecDestroy(exec));
    CK(cudaGraphDestroy(graph));
    CK(cudaStreamDestroy(s));
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned* mem;
    float* data;
    CK(cudaMalloc(&mem, 32 * 4 * (1 + 2 * 1024)));
    CK(cudaMalloc(&data, sizeof(float) * 148 * 4 * 256 * 2));
    CK(cudaMemset(data, 0, sizeof(float) * 148 * 4 * 256 * 2));
    printf("SMs: %d\n", sms);
    for (int work = 0; work <= 1; ++work)
        for (int bps = 1; bps <= 2; ++bps) {
            run<0>("v0 counter+poll (current)", bps, 1, work, mem, data, sms);
            run<1>("v1 red.release+poll", bps, 1, work, mem, data, sms);
            for (int group : {4, 8, 16, 32}) run<2>("v2 two-level + flags", bps, group, work, mem, data, sms);
            for (int group : {8, 16}) run<3>("v3 two-level, warp flags", bps, group, work, mem, data, sms);
            run<4>("cg grid.sync()", bps, 1, work, mem, data, sms);
        }
    for (int blocks : {1, 148, 800})
        for (int work = 0; work <= 1; ++work)
            for (int pdl = 0; pdl <= 1; ++pdl) run_chain(blocks, work, pdl, data);
    return 0;
}
