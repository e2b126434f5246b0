// consumers.cu -- the step after the scan in a column store.  The reference
// stops at the bitvector; its header only records the intent of combining range predicates
// (src/simd_scan.hpp:75-84).  Two small, bandwidth-bound operators:
//
//   bitvec_combine : out = a OP b over 32-bit words (+ popcount), for conjunctions / disjunctions of
//                    predicates on several columns.  Bytes per code: 3/8.
//   unpack_at      : late materialisation -- decode the codes of another packed column at a position
//                    list (random access: a word pair + funnel shift per position).
#include "common.cuh"
#include "kernels.h"

namespace sss {

__device__ __forceinline__ uint32_t combine(int op, uint32_t a, uint32_t b)
{
    switch (op) {
        case 0: return a & b;
        case 1: return a | b;
        case 2: return a & ~b;
        default: return a ^ b;
    }
}

__global__ void __launch_bounds__(kThreads)
bitvec_combine_kernel(int op, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n,
                      uint32_t* __restrict__ out, unsigned long long* __restrict__ count,
                      unsigned long long* __restrict__ count_slot)
{
    const uint64_t n_words = (n + 31) >> 5;
    const uint64_t n_vec = n_words >> 2;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    uint32_t hits = 0;
    // a, b and out may alias: every element is read before it is written, by the same thread
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_vec; i += stride) {
        const uint4 x = reinterpret_cast<const uint4*>(a)[i], y = reinterpret_cast<const uint4*>(b)[i];
        uint4 r;
        r.x = combine(op, x.x, y.x);
        r.y = combine(op, x.y, y.y);
        r.z = combine(op, x.z, y.z);
        r.w = combine(op, x.w, y.w);
        hits += __popc(r.x) + __popc(r.y) + __popc(r.z) + __popc(r.w);
        reinterpret_cast<uint4*>(out)[i] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) {  // up to 3 trailing words, the last possibly ragged
        const uint64_t w = (n_vec << 2) + threadIdx.x;
        if (w < n_words) {
            uint32_t r = combine(op, a[w], b[w]);
            const uint32_t tail_bits = (uint32_t)(n & 31);
            if (tail_bits != 0 && w == n_words - 1) r &= (1u << tail_bits) - 1u;
            hits += __popc(r);
            out[w] = r;
        }
    }
    if (count != nullptr) cta_finish_count(hits, count_slot, count, false);
}

__global__ void __launch_bounds__(kThreads)
unpack_at_kernel(const uint32_t* __restrict__ packed, int w, const unsigned long long* __restrict__ positions,
                 uint64_t num, int32_t* __restrict__ out)
{
    const uint32_t mask = code_mask(w);
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < num; i += stride) {
        const unsigned long long bit = positions[i] * (unsigned long long)w;
        const uint64_t wi = bit >> 5;
        const uint32_t sh = (uint32_t)(bit & 31);
        const uint32_t lo = __ldg(packed + wi);
        // the second word only matters when the code straddles; the buffer's 256-byte padding
        // (sss_packed_bytes) keeps the read legal for the last code
        const uint32_t hi = (sh + (uint32_t)w > 32) ? __ldg(packed + wi + 1) : 0u;
        out[i] = (int32_t)(__funnelshift_r(lo, hi, sh) & mask);
    }
}

static unsigned grid_for(const Device& dev, uint64_t items, int per_sm)
{
    const uint64_t want = (items + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)dev.sm_count * per_sm;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

cudaError_t launch_bitvec_combine(const Device& dev, int op, const uint32_t* a, const uint32_t* b, uint64_t n,
                                  uint32_t* out, const CountTarget& count, cudaStream_t stream)
{
    const uint64_t n_vec = ((n + 31) / 32) / 4;
    bitvec_combine_kernel<<<grid_for(dev, n_vec, 8), kThreads, 0, stream>>>(op, a, b, n, out, count.result, count.slot);
    return cudaGetLastError();
}

cudaError_t launch_unpack_at(const Device& dev, const void* packed, int w, const unsigned long long* positions,
                             uint64_t num, int32_t* out, cudaStream_t stream)
{
    unpack_at_kernel<<<grid_for(dev, num, 16), kThreads, 0, stream>>>(static_cast<const uint32_t*>(packed), w,
                                                                      positions, num, out);
    return cudaGetLastError();
}

}  // namespace sss
