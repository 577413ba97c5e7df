// Host stand-in for <cuda_runtime.h>, used ONLY by tests/device_on_host: lets g++ compile the constraint headers of
// bepuphysics2_b200/csrc (plain scalar C++ behind __device__) so that their arithmetic can be checked against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))

// cache-hinted loads / stores are plain memory accesses on the host
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float __int_as_float(int i) {
    float f;
    std::memcpy(&f, &i, sizeof f);
    return f;
}
