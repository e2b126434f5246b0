// positions.cu -- bitvector -> ascending position list.
//
// The reference produces no position list; the north star asks for one.  It is
// defined as the ascending indices of the set bits in get_bit order (reference src/util.cpp:51-58).
// A warp-aggregated atomic append would emit the right SET but in nondeterministic order, so this
// is a deterministic three-step expansion instead: per-tile popcounts, an exclusive scan of the
// tile counts, then an ordered expansion in which every thread knows its exact output offset.
// Input traffic is n/8 bytes read twice; output is 8 bytes per match.
#include "common.cuh"
#include "kernels.h"

namespace sss {

constexpr int kPosWordsPerThread = 4;
constexpr int kPosTileWords = kThreads * kPosWordsPerThread;  // 1024 words = 32768 codes per tile

static inline uint64_t pos_tiles(uint64_t n)
{
    const uint64_t words = (n + 31) / 32;
    return (words + kPosTileWords - 1) / kPosTileWords;
}

size_t positions_workspace_bytes(uint64_t n)
{
    // [tiles] u64 offsets, then [tiles] u32 counts
    return (size_t)(pos_tiles(n) * 12 + 16);
}

__device__ __forceinline__ uint32_t load_word(const uint32_t* __restrict__ bitvec, uint64_t word, uint64_t n_words,
                                              uint32_t tail_bits)
{
    if (word >= n_words) return 0u;
    uint32_t v = __ldg(bitvec + word);
    if (tail_bits != 0 && word == n_words - 1) v &= (1u << tail_bits) - 1u;
    return v;
}

__global__ void __launch_bounds__(kThreads)
pos_count_kernel(const uint32_t* __restrict__ bitvec, uint64_t n, uint64_t tiles, uint32_t* __restrict__ tile_counts)
{
    __shared__ uint32_t warp_sums[kWarps];
    const uint64_t n_words = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t w0 = tile * kPosTileWords + threadIdx.x * kPosWordsPerThread;
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < kPosWordsPerThread; i++) c += __popc(load_word(bitvec, w0 + i, n_words, tail_bits));
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t s = 0;
#pragma unroll
            for (int i = 0; i < kWarps; i++) s += warp_sums[i];
            tile_counts[tile] = s;
        }
        __syncthreads();
    }
}

// single CTA: exclusive scan of the tile counts into 64-bit offsets, total -> *num
__global__ void __launch_bounds__(1024)
pos_scan_kernel(const uint32_t* __restrict__ tile_counts, uint64_t tiles, unsigned long long* __restrict__ tile_offsets,
                unsigned long long* __restrict__ num)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < tiles ? tile_counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = warp_tot[lane];
            unsigned long long ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, ti, d);
                if (lane >= d) ti += o;
            }
            warp_tot[lane] = ti - t;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        const unsigned long long excl = carry + warp_tot[warp] + (incl - v);
        if (i < tiles) tile_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && num != nullptr) *num = carry_s;
}

// Ordered expansion.  A tile is 1024 bitvector words; warp w owns words [128w, 128w+128) in four
// rounds of 32 (lane l loads word l of the round, coalesced).  Within a round the words are walked
// one at a time by the whole warp: lane l owns bit l, its rank among the set bits below it is a
// popcount, so the set bits of a word are written by one store instruction to consecutive slots --
// ascending order and coalesced stores at any selectivity.  All-zero words are skipped via a ballot.
__global__ void __launch_bounds__(kThreads)
pos_expand_kernel(const uint32_t* __restrict__ bitvec, uint64_t n, uint64_t tiles,
                  const unsigned long long* __restrict__ tile_offsets, uint64_t base,
                  unsigned long long* __restrict__ positions, uint64_t capacity)
{
    __shared__ uint32_t warp_tot[kWarps];
    const uint64_t n_words = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t below = (1u << lane) - 1u;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t w0 = tile * kPosTileWords + warp * 128;
        uint32_t v[4];
        uint32_t c = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            v[r] = load_word(bitvec, w0 + r * 32 + lane, n_words, tail_bits);
            c += __popc(v[r]);
        }
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) warp_tot[warp] = c;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int i = 0; i < kWarps; i++) before += (i < (int)warp) ? warp_tot[i] : 0u;
        uint64_t o = tile_offsets[tile] + before;  // first output slot of this warp's words
#pragma unroll
        for (int r = 0; r < 4; r++) {
            // exclusive prefix of the words' popcounts across the lanes of this round
            const uint32_t pc = __popc(v[r]);
            uint32_t incl = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            const uint32_t excl = incl - pc;
            const uint32_t round_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (round_total > 96) {
                // dense round: walk the words, one coalesced store per non-empty word
                uint32_t nonzero = __ballot_sync(0xFFFFFFFFu, v[r] != 0);
                while (nonzero) {
                    const int i = __ffs(nonzero) - 1;
                    nonzero &= nonzero - 1;
                    const uint32_t word = __shfl_sync(0xFFFFFFFFu, v[r], i);
                    const uint32_t start = __shfl_sync(0xFFFFFFFFu, excl, i);
                    if ((word >> lane) & 1u) {
                        const uint64_t slot = o + start + __popc(word & below);
                        if (slot < capacity) positions[slot] = base + (w0 + (uint64_t)r * 32 + i) * 32 + lane;
                    }
                }
            } else {
                // sparse round (<= 3 set bits per word on average): every lane expands its own word
                uint32_t bits = v[r];
                uint64_t slot = o + excl;
                const uint64_t code0 = base + (w0 + (uint64_t)r * 32 + lane) * 32;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (slot < capacity) positions[slot] = code0 + b;
                    slot++;
                }
            }
            o += round_total;
        }
        __syncthreads();
    }
}

cudaError_t launch_positions(const Device& dev, const uint32_t* bitvec, uint64_t n, uint64_t base,
                             unsigned long long* positions, uint64_t capacity, unsigned long long* num,
                             void* workspace, cudaStream_t stream)
{
    const uint64_t tiles = pos_tiles(n);
    unsigned long long* tile_offsets = static_cast<unsigned long long*>(workspace);
    uint32_t* tile_counts = reinterpret_cast<uint32_t*>(tile_offsets + tiles);
    if (tiles == 0) return num ? cudaMemsetAsync(num, 0, sizeof(unsigned long long), stream) : cudaSuccess;
    const uint64_t cap = (uint64_t)dev.sm_count * 8;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    pos_count_kernel<<<grid, kThreads, 0, stream>>>(bitvec, n, tiles, tile_counts);
    pos_scan_kernel<<<1, 1024, 0, stream>>>(tile_counts, tiles, tile_offsets, num);
    pos_expand_kernel<<<grid, kThreads, 0, stream>>>(bitvec, n, tiles, tile_offsets, base, positions, capacity);
    return cudaGetLastError();
}

}  // namespace sss
