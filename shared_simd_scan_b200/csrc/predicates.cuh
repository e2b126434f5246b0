// predicates.cuh -- the K query bounds of a shared scan, held in shared memory in left-aligned
// form (see sss::aligned_range), plus per-predicate match counters.
//
// Counterpart of the reference keeping the predicate keys in a std::vector<int> and broadcasting
// them with _mm_set1_epi32 inside the inner loop (src/simd_scan_shared.cpp:73-76) or in 2K
// pre-shifted registers (src/simd_scan.hpp:177-191).
#pragma once

#include "common.cuh"

namespace sss {

constexpr int SSS_MAX_K = 1024;  // == SSS_MAX_PREDICATES of the C ABI

struct PredicateTable {
    uint32_t* lo_al;
    uint32_t* span_al;
    uint32_t* valid;   // 0xFFFFFFFF, or 0 for an empty range (lo > hi or lo >= 2^w)
    uint32_t* counts;  // per-CTA match counters, flushed once with 64-bit global atomics

    __device__ __forceinline__ explicit PredicateTable(uint8_t* smem16)
        : lo_al(reinterpret_cast<uint32_t*>(smem16)), span_al(lo_al + SSS_MAX_K), valid(span_al + SSS_MAX_K),
          counts(valid + SSS_MAX_K)
    {
    }

    // all threads; visible after the next __syncthreads
    __device__ __forceinline__ void load(int w, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                         int K)
    {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const AlignedRange r = aligned_range(w, lo[k], hi[k]);
            lo_al[k] = r.lo_al;
            span_al[k] = r.span_al;
            valid[k] = r.empty ? 0u : 0xFFFFFFFFu;
            counts[k] = 0;
        }
    }

    // all threads, once, after the last tile
    __device__ __forceinline__ void flush_counts(unsigned long long* __restrict__ global, int K)
    {
        __syncthreads();
        if (global == nullptr) return;
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            if (counts[k]) atomicAdd(&global[k], (unsigned long long)counts[k]);
    }
};

}  // namespace sss
