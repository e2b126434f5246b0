// positions.cu -- bitvector -> ascending position list, in ONE pass over the bitvector.
//
// The reference produces no position list; the north star asks for one.  It is
// defined as the ascending indices of the set bits in get_bit order (reference src/util.cpp:51-58).
// A warp-aggregated atomic append would emit the right SET but in nondeterministic order, so the
// output offset of every tile is computed exactly, by a single-pass prefix sum with decoupled
// look-back: a CTA takes the next tile from a ticket counter, popcounts it, publishes the tile's
// total, sums the totals / inclusive prefixes its predecessors have published (they are running or
// finished: tickets are handed out in order), publishes its own inclusive prefix and expands its
// tile at that offset.  Round 1 read the bitvector twice and scanned the tile counts with one CTA
// (three launches, 0.24-0.28 of the HBM peak).
//
// Expansion is ordered AND coalesced: a warp owns 256 consecutive words in 8 rounds of 32 (lane l
// loads word l of the round); every lane expands the set bits of its word into a warp-private
// shared-memory stage at its exclusive offset within the round, then the warp writes the stage out
// with consecutive 8-byte stores.
//
// Algorithmic bytes: n/8 (bitvector, read once) + 8 per match.
#include "common.cuh"
#include "kernels.h"

namespace sss {

constexpr int kPosRounds = 8;                                // words per lane per tile (32 measured slower: 0.18 -> 0.22 ms at 0.1 % selectivity)
constexpr int kPosTileWords = kThreads * kPosRounds;         // 2048 words = 65536 codes per tile
constexpr int kPosWarpWords = 32 * kPosRounds;               // 256 words per warp
constexpr int kPosStage = 1024;                              // positions per round and warp (32 words x 32 bits)

// tile state word: [value : 42][flag : 2][epoch : 20].  A state whose epoch is not this launch's is
// "not yet written", so the array needs no clearing between launches (capi.cu owns the epoch).
constexpr unsigned long long kPosFlagAggregate = 1, kPosFlagPrefix = 2;
constexpr int kPosEpochBits = 20;

static inline uint64_t pos_tiles(uint64_t n)
{
    const uint64_t words = (n + 31) / 32;
    return (words + kPosTileWords - 1) / kPosTileWords;
}

size_t positions_workspace_bytes(uint64_t n)
{
    // {ticket, done} counters (fixed place: they must be zero between launches whatever n is), then
    // [tiles] u64 states
    return (size_t)(32 + pos_tiles(n) * 8);
}

uint32_t positions_epoch_limit() { return (1u << kPosEpochBits) - 1; }

__device__ __forceinline__ unsigned long long pos_pack(unsigned long long value, unsigned long long flag, uint32_t epoch)
{
    return (value << (kPosEpochBits + 2)) | (flag << kPosEpochBits) | epoch;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads)
positions_kernel(const uint32_t* __restrict__ bitvec, uint64_t n, uint64_t tiles, unsigned long long* __restrict__ state,
                 unsigned long long* __restrict__ counters, uint32_t epoch, uint64_t base,
                 unsigned long long* __restrict__ positions, uint64_t capacity, unsigned long long* __restrict__ num)
{
    __shared__ uint32_t warp_tot[kWarps];
    __shared__ unsigned long long s_tile, s_prefix;
    __shared__ uint32_t stage[kWarps][kPosStage];
    const uint64_t n_words = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[0], 1ull);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= tiles) break;

        // ---- load + count ------------------------------------------------------------------------
        const uint64_t w0 = tile * kPosTileWords + warp * kPosWarpWords;
        uint32_t v[kPosRounds];
        uint32_t c = 0;
#pragma unroll
        for (int r = 0; r < kPosRounds; r++) {
            const uint64_t word = w0 + r * 32 + lane;
            uint32_t x = word < n_words ? __ldcs(bitvec + word) : 0u;
            if (tail_bits != 0 && word == n_words - 1) x &= (1u << tail_bits) - 1u;
            v[r] = x;
            c += __popc(x);
        }
        c = __reduce_add_sync(0xFFFFFFFFu, c);
        if (lane == 0) warp_tot[warp] = c;
        __syncthreads();

        // ---- decoupled look-back (warp 0) ----------------------------------------------------------
        if (warp == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int i = 0; i < kWarps; i++) total += warp_tot[i];
            unsigned long long exclusive = 0;
            if (tile != 0) {
                if (lane == 0) st_relaxed_u64(&state[tile], pos_pack(total, kPosFlagAggregate, epoch));
                int64_t look = (int64_t)tile - 1 - lane;  // lane l inspects tile - 1 - l of the window
                for (;;) {
                    unsigned long long s = 0;
                    bool ready = true, prefix = false;
                    if (look >= 0) {
                        for (;;) {
                            s = ld_relaxed_u64(&state[look]);
                            if ((uint32_t)(s & ((1u << kPosEpochBits) - 1)) == epoch && ((s >> kPosEpochBits) & 3) != 0) break;
                            __nanosleep(20);
                        }
                        prefix = ((s >> kPosEpochBits) & 3) == kPosFlagPrefix;
                    } else {
                        ready = false;  // before the first tile: contributes nothing, ends the walk
                        prefix = true;
                    }
                    // the nearest predecessor that carries an inclusive prefix closes the sum
                    const uint32_t pmask = __ballot_sync(0xFFFFFFFFu, prefix);
                    const uint32_t upto = pmask ? (uint32_t)(__ffs(pmask) - 1) : 31u;
                    unsigned long long add = (ready && lane <= upto) ? (s >> (kPosEpochBits + 2)) : 0ull;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) add += __shfl_xor_sync(0xFFFFFFFFu, add, d);
                    exclusive += add;
                    if (pmask) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                st_relaxed_u64(&state[tile], pos_pack(exclusive + total, kPosFlagPrefix, epoch));
                s_prefix = exclusive;
                if (tile == tiles - 1 && num != nullptr) *num = exclusive + total;
            }
        }
        __syncthreads();

        // ---- ordered expansion ---------------------------------------------------------------------
        uint32_t before = 0;
#pragma unroll
        for (int i = 0; i < kWarps; i++) before += (i < (int)warp) ? warp_tot[i] : 0u;
        uint64_t o = s_prefix + before;  // first output slot of this warp's words
        if (c != 0) {  // warp-uniform: nothing set in these 8192 codes
            uint32_t* st = stage[warp];
#pragma unroll
            for (int r2 = 0; r2 < kPosRounds; r2 += 2) {
                // exclusive prefix of the words' popcounts across the lanes, two rounds per scan
                const uint32_t pc = __popc(v[r2]) | (__popc(v[r2 + 1]) << 16);
                uint32_t incl = pc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += t;
                }
                const uint32_t excl2 = incl - pc;
                const uint32_t tot2 = __shfl_sync(0xFFFFFFFFu, incl, 31);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r = r2 + h;
                    const uint32_t round_total = (tot2 >> (16 * h)) & 0xFFFFu;
                    if (round_total == 0) continue;  // warp-uniform
                    uint32_t slot = (excl2 >> (16 * h)) & 0xFFFFu;
                    uint32_t bits = v[r];
                    const uint32_t code0 = (uint32_t)(warp * kPosWarpWords + r * 32 + lane) * 32u;  // within the tile
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        st[slot++] = code0 + b;
                    }
                    __syncwarp();
                    const uint64_t tile_code0 = base + tile * (uint64_t)(kPosTileWords * 32);
                    for (uint32_t i = lane; i < round_total; i += 32) {
                        const uint64_t dst = o + i;
                        if (dst < capacity) __stcs(positions + dst, tile_code0 + st[i]);
                    }
                    __syncwarp();
                    o += round_total;
                }
            }
        }
        // (the __syncthreads at the top of the next iteration protects warp_tot / s_tile / s_prefix)
    }
    // the last CTA out re-arms the ticket counter for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&counters[1], 1ull) == gridDim.x - 1) {
            counters[0] = 0;
            counters[1] = 0;
        }
    }
}

cudaError_t launch_positions(const Device& dev, const uint32_t* bitvec, uint64_t n, uint64_t base,
                             unsigned long long* positions, uint64_t capacity, unsigned long long* num,
                             void* workspace, uint32_t epoch, cudaStream_t stream)
{
    const uint64_t tiles = pos_tiles(n);
    if (tiles == 0) return num ? cudaMemsetAsync(num, 0, sizeof(unsigned long long), stream) : cudaSuccess;
    unsigned long long* counters = static_cast<unsigned long long*>(workspace);
    unsigned long long* state = counters + 4;
    // as many CTAs as stay resident (the look-back only ever waits for tiles whose ticket was drawn,
    // so a smaller or larger grid is still correct)
    const uint64_t cap = (uint64_t)dev.sm_count * 6;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    positions_kernel<<<grid, kThreads, 0, stream>>>(bitvec, n, tiles, state, counters, epoch, base, positions, capacity,
                                                    num);
    return cudaGetLastError();
}

}  // namespace sss
