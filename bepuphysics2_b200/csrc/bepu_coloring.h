// Device-side batch colouring: the batch search of Solver.Add / FindCandidateBatch (Solver.cs:L984-1014, L1182-1199) for a whole
// constraint set at once, and the fixed point BatchCompressor works towards (BatchCompressor.cs:L233) when fed the current batch indices as priorities.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace bepucuda {

// Priority key of constraint c: constraints are assigned in ascending key order; ties cannot occur (the low word is the constraint index).
//   order 0 (insertion order): key = c                       -- exactly the sequence of Solver.Add calls
//   order 1 (hashed order)   : key = color_hash(c) << 32 | c -- a fixed pseudo-random permutation: O(log n) dependent rounds on bounded-degree graphs
//   order 2 (by priority)    : key = priorities[c] << 32 | c -- caller-defined (current batch index: compression)
__host__ __device__ inline uint32_t color_hash(uint32_t c) {
    uint32_t h = c * 0x9E3779B1u;
    h ^= h >> 15;
    h *= 0x85EBCA77u;
    h ^= h >> 13;
    h *= 0xC2B2AE3Du;
    h ^= h >> 16;
    return h;
}

struct ColoringBuffers {
    const int32_t* refs;             // [constraint][slot], slot stride = bodies_per_constraint; -1 = unused slot, bit 30 = kinematic (never blocks)
    const uint32_t* priorities;      // order 2 only
    unsigned long long* body_min;    // per body: lowest key among its uncoloured constraints (all ones = none)
    unsigned long long* body_mask;   // per body: synchronized batches already holding one of its constraints (batchReferencedHandles, transposed)
    int32_t* batch_out;              // per constraint: assigned batch index, -1 while uncoloured
    int32_t* list[2];                // uncoloured constraints, ping-pong
    unsigned int* counts;            // counts[r] = length of the list round r reads
    int32_t constraint_count, bodies_per_constraint, body_count, fallback_threshold, order;
};
// One round = launch_color_round: every uncoloured constraint that holds the lowest key on each of its dynamic bodies takes the lowest batch
// none of those bodies is in yet (first fit; the fallback batch when all synchronized batches are blocked). Reads list[round & 1] (counts[round]
// entries), appends the rest to list[(round + 1) & 1] / counts[round + 1], which must be zero beforehand.
void launch_color_init(const ColoringBuffers& cb, cudaStream_t s);
void launch_color_round(const ColoringBuffers& cb, int round, cudaStream_t s);

}  // namespace bepucuda
