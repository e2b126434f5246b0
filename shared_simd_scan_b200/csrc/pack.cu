// pack.cu -- the compressor (codes -> packed bit stream) and the synthetic-column generator.
//
// sss_pack replaces compress_9bit_input (reference src/simd_scan_compression.cpp:53-104) for any
// width 1..32 and uint32 codes: code i occupies stream bits [i*w, (i+1)*w), little-endian.  The
// reference scatters each code into one or two uint64 words with a serial cursor; here every thread
// owns one 32-bit OUTPUT word and gathers the (at most 32/w + 2) codes that overlap it, so there
// are no write conflicts and stores are fully coalesced.
//
// sss_generate_packed produces the benchmark column (uniform codes from a counter-based
// SplitMix64) directly in packed form, with no unpacked copy in HBM -- the 2^30..2^34-code columns
// of BASELINE.json never exist unpacked.
#include "common.cuh"
#include "kernels.h"

namespace sss {

// must stay identical to orc_mix64 / orc_generate_codes (oracle/oracle.c)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct FromArray {
    const uint32_t* codes;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return __ldg(codes + i); }
};

struct FromGenerator {
    uint64_t seed, first;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const
    {
        return (uint32_t)mix64(seed + (first + i) * 0xD1342543DE82EF95ull);
    }
};

template <class Source>
__global__ void __launch_bounds__(kThreads)
pack_kernel(Source src, uint64_t n, int w, uint32_t* __restrict__ packed, uint64_t packed_words)
{
    const uint32_t mask = code_mask(w);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t word = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; word < packed_words; word += stride) {
        const uint64_t bit0 = word * 32;
        uint64_t c = bit0 / (uint64_t)w;  // first code overlapping this word
        uint32_t acc = 0;
        for (; c < n; c++) {
            const int64_t rel = (int64_t)(c * (uint64_t)w) - (int64_t)bit0;  // code's offset inside the word
            if (rel >= 32) break;
            const uint32_t v = src(c) & mask;
            acc |= rel >= 0 ? (v << rel) : (v >> (-rel));
        }
        packed[word] = acc;  // words past the payload are written as zero padding
    }
}

static unsigned grid_for(const Device& dev, uint64_t items)
{
    const uint64_t want = (items + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)dev.sm_count * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_pack(const Device& dev, const uint32_t* codes, uint64_t n, int w, uint32_t* packed,
                        uint64_t packed_words, cudaStream_t stream)
{
    pack_kernel<<<grid_for(dev, packed_words), kThreads, 0, stream>>>(FromArray{codes}, n, w, packed, packed_words);
    return cudaGetLastError();
}

cudaError_t launch_generate_packed(const Device& dev, uint32_t* packed, uint64_t packed_words, uint64_t n, int w,
                                   uint64_t seed, uint64_t first_index, cudaStream_t stream)
{
    pack_kernel<<<grid_for(dev, packed_words), kThreads, 0, stream>>>(FromGenerator{seed, first_index}, n, w, packed,
                                                                     packed_words);
    return cudaGetLastError();
}

}  // namespace sss
