// positions.cu -- bitvector -> ascending position list.
//
// The reference produces no position list; the north star asks for one.  It is
// defined as the ascending indices of the set bits in get_bit order (reference src/util.cpp:51-58).
// A warp-aggregated atomic append would emit the right SET but in nondeterministic order, so every
// 4 KB SEGMENT of the bitvector (1024 words = 32768 codes, one warp) gets its exact output offset:
//
//   pos_count_kernel   popcount per segment, walking the bitvector from its END to its start
//   pos_offsets_kernel exclusive prefix sum over groups of 8 segments (one CTA; 4096 groups at 2^30 bits)
//   pos_expand_kernel  every warp expands its segment at its offset, walking from the START
//
// History (profiles/r02_pos_*.txt).  Round 1: the same three steps, but per-lane serial expansion and
// scattered stores, 0.24-0.28 of the HBM peak.  Round 2 first tried ONE pass with a decoupled
// look-back (a CTA draws a tile ticket, publishes its popcount, sums what its predecessors
// published): 0.12-0.28 of the peak when few bits are set, and a trace of the tile iterations showed
// why -- the bitvector streams at ~200 tiles of 32 KB per microsecond, a tile's words wait in
// REGISTERS while its look-back polls the predecessors (6-10 us: the slowest of the aggregates it
// needs, several L2 round trips), and the register file holds only 19 MB of tiles: 1.8 TB/s whatever
// the window width or ticket timing.  Nothing waits here: the bitvector is read twice instead, and the
// second read mostly comes out of the 126 MB L2 -- the count pass ends at the low addresses, the
// expansion starts there and reads with evict-first loads so that it does not displace what it has
// not reached yet.  (After a scan the bitvector's freshly written end is still in L2 as well, which
// is why the count pass starts there.)
//
// Also measured and dropped: a sparse scratch -- the count pass leaving every segment's non-zero pieces
// compacted in the workspace so that the expansion reads 18 MB instead of the 134 MB bitvector at 0.1 %
// selectivity.  The expansion took the same 30 us with 20 MB of DRAM reads as with 134 MB: at low
// selectivity it is bound by a warp's chain of dependent L2 round trips (counts -> offset -> words ->
// stores) at 32 warps per SM and by its instruction count, not by the bytes.
//
// Expansion is ordered AND coalesced: a warp owns 1024 consecutive words in 8 rounds of 128 (lane l
// loads words 4l..4l+3 of the round with one 128-bit load, all eight in flight); two rounds share a
// warp scan; every lane expands the set bits of its words into a warp-private shared-memory stage at
// its exclusive offset, rounds are collected there until the stage is full, then the warp writes the
// stage out with consecutive 8-byte stores.  A round with more than half of its bits set skips the
// stage: the warp takes its words one at a time, lane b owns bit b and stores at the rank of its bit.
//
// sss_scan_range_positions skips the first kernel: the scan that writes the bitvector adds every segment's
// matches to seg_count itself (scan.cu, kSeg; launch_positions(..., counted = true)).  The three kernels are
// launched with programmatic stream serialization (pdl_enter, common.cuh): 5-60 us launches whose gaps showed
// (sss_positions at 0.1 % selectivity 53.9 -> 47.7 us, profiles/r02_pos_pdl.txt).
//
// Algorithmic bytes: n/8 (bitvector, counted once) + 8 per match.
#include "common.cuh"
#include "kernels.h"

namespace sss {

constexpr int kPosRounds = 8;                                // 128-bit loads per lane and segment
constexpr int kPosRoundWords = 32 * 4;                       // words per warp and round
constexpr int kPosSegWords = kPosRoundWords * kPosRounds;    // 1024 words = 32768 codes per segment (warp)
#ifndef SSS_POS_STAGE
#define SSS_POS_STAGE 2048
#endif
constexpr int kPosStage = SSS_POS_STAGE;                              // positions a warp collects before it writes them out
#ifndef SSS_POS_SPARSE
#define SSS_POS_SPARSE 128
#endif
constexpr int kPosSparse = SSS_POS_SPARSE;                   // segments with at most this many set bits take the item path
static_assert(kPosRounds * 32 * 16 <= kPosStage * 2, "the items (a lane's 128-bit piece of a round) live in the stage");
static_assert(kPosSegWords * 32 <= 65536, "stage entries are 16-bit offsets within the segment");

static inline uint64_t pos_segments(uint64_t n)
{
    const uint64_t words = (n + 31) / 32;
    return (words + kPosSegWords - 1) / kPosSegWords;
}

static inline uint64_t pos_groups(uint64_t n) { return (pos_segments(n) + kWarps - 1) / kWarps; }

size_t positions_workspace_bytes(uint64_t n)
{
    // [groups of 8 segments] u64 offsets (padded to 16 bytes), then [segments] u32 counts
    return (size_t)(pos_groups(n) * 16 + pos_segments(n) * 4 + 16);
}

uint32_t positions_epoch_limit() { return 0xFFFFFFFFu; }  // the workspace carries no state between launches

// The warp's eight rounds of segment `seg`: lane l holds words seg*1024 + r*128 + 4l .. +3 in v[r];
// words past the end read as 0 and the ragged last word is masked.
template <bool kStreaming>
__device__ __forceinline__ void pos_load_segment(const uint32_t* __restrict__ bitvec, uint64_t seg, uint64_t n_words,
                                                 uint32_t tail_bits, bool aligned, uint32_t lane, uint4 (&v)[kPosRounds])
{
    const uint64_t w0 = seg * kPosSegWords + lane * 4;
#pragma unroll
    for (int r = 0; r < kPosRounds; r++) {
        const uint64_t word = w0 + r * kPosRoundWords;
        if (aligned && word + 4 < n_words) {
            const uint4* p = reinterpret_cast<const uint4*>(bitvec + word);
            v[r] = kStreaming ? __ldcs(p) : __ldg(p);
        } else {
            uint32_t x[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                x[j] = word + j < n_words ? __ldg(bitvec + word + j) : 0u;
                if (tail_bits != 0 && word + j == n_words - 1) x[j] &= (1u << tail_bits) - 1u;
            }
            v[r] = make_uint4(x[0], x[1], x[2], x[3]);
        }
    }
}

__device__ __forceinline__ uint32_t popc4(const uint4& a) { return __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w); }

// ---- 1: popcount per segment, last segment first ------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
pos_count_kernel(const uint32_t* __restrict__ bitvec, uint64_t n, uint64_t segments, uint32_t* __restrict__ seg_count)
{
    pdl_enter();  // (launched with programmatic stream serialization, like the two kernels below: common.cuh)
    const uint64_t n_words = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool aligned = (reinterpret_cast<uintptr_t>(bitvec) & 15) == 0;
    const uint64_t i = (uint64_t)blockIdx.x * kWarps + warp;
    if (i >= segments) return;
    const uint64_t seg = segments - 1 - i;
    uint4 v[kPosRounds];
    pos_load_segment<false>(bitvec, seg, n_words, tail_bits, aligned, lane, v);
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < kPosRounds; r++) c += popc4(v[r]);
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    if (lane == 0) seg_count[seg] = c;
}

// ---- 2: exclusive prefix sum over GROUPS of kWarps segments (one CTA of 1024 threads) -------------------
// A group is what one CTA of the expansion handles; that CTA adds up its own eight counts.  Thread t
// of iteration i owns group i * 1024 + t: its counts are one 32-byte run and a warp reads 1 KB in
// one piece (a first version gave every thread 32 consecutive counts and took 25 us for 32768
// segments: 128-byte strides between the lanes, one sector per lane and store).
constexpr int kPosScanThreads = 1024;
static_assert(kWarps == 8, "a group is two 128-bit loads of segment counts");
__global__ void __launch_bounds__(kPosScanThreads)
pos_offsets_kernel(const uint32_t* __restrict__ seg_count, uint64_t segments, unsigned long long* __restrict__ group_off,
                   unsigned long long* __restrict__ num)
{
    pdl_enter();
    __shared__ unsigned long long warp_sum[kPosScanThreads / 32];
    __shared__ unsigned long long carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t groups = (segments + kWarps - 1) / kWarps;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    // kPosScanUnroll iterations' counts are requested together: one memory latency per 4096 groups, not four
    constexpr int kU = 4;
    for (uint64_t g0 = 0; g0 < groups; g0 += (uint64_t)kPosScanThreads * kU) {
        unsigned long long mine[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const uint64_t g = g0 + (uint64_t)u * kPosScanThreads + threadIdx.x, first = g * kWarps;
            mine[u] = 0;
            if (first + kWarps <= segments) {
                const uint4 a = *reinterpret_cast<const uint4*>(seg_count + first);
                const uint4 b = *reinterpret_cast<const uint4*>(seg_count + first + 4);
                mine[u] = (unsigned long long)a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
            } else {
                for (uint64_t s = first; s < segments; s++) mine[u] += seg_count[s];
            }
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const uint64_t g = g0 + (uint64_t)u * kPosScanThreads + threadIdx.x;
            unsigned long long incl = mine[u];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            if (lane == 31) warp_sum[warp] = incl;
            __syncthreads();
            unsigned long long before = carry;
            for (uint32_t i = 0; i < warp; i++) before += warp_sum[i];
            if (g < groups) group_off[g] = before + incl - mine[u];
            __syncthreads();
            if (threadIdx.x == kPosScanThreads - 1) carry = before + incl;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0 && num != nullptr) *num = carry;
}

// ---- 3: ordered expansion, first segment first ---------------------------------------------------------
#ifndef SSS_POS_CTAS
#define SSS_POS_CTAS 4
#endif
__global__ void __launch_bounds__(kThreads, SSS_POS_CTAS)
pos_expand_kernel(const uint32_t* __restrict__ bitvec, uint64_t n, uint64_t segments,
                  const uint32_t* __restrict__ seg_count, const unsigned long long* __restrict__ group_off, uint64_t base,
                  unsigned long long* __restrict__ positions, uint64_t capacity)
{
    pdl_enter();
    __shared__ __align__(16) uint16_t stage[kWarps][kPosStage];
    __shared__ uint8_t piece_id[kWarps][kPosRounds * 32];
    const uint64_t n_words = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool aligned = (reinterpret_cast<uintptr_t>(bitvec) & 15) == 0;
    const uint64_t seg = (uint64_t)blockIdx.x * kWarps + warp;
    if (seg >= segments) return;
    // the segment's words are requested before anything depends on its count or offset
    uint4 v[kPosRounds];
    pos_load_segment<true>(bitvec, seg, n_words, tail_bits, aligned, lane, v);
    // the group's offset plus the counts of the group's segments before this one (lane i holds count i)
    const uint64_t seg0 = (uint64_t)blockIdx.x * kWarps;
    const uint32_t cnt = (lane < (uint32_t)kWarps && seg0 + lane < segments) ? __ldg(seg_count + seg0 + lane) : 0u;
    uint64_t o = __ldg(group_off + blockIdx.x);
    const uint32_t my_cnt = __shfl_sync(0xFFFFFFFFu, cnt, warp);
    if (my_cnt == 0) return;  // nothing set in these 32768 codes
    o += __reduce_add_sync(0xFFFFFFFFu, lane < warp ? cnt : 0u);
    if (o >= capacity) return;  // the caller's list is full before this segment

    uint16_t* st = stage[warp];
    const uint64_t seg_code0 = base + seg * (uint64_t)(kPosSegWords * 32);

    // Few bits set (at most kPosSparse in the segment, known from its count): the general path below
    // costs ~850 warp instructions per segment whatever it finds (32 per-word loops, four scans), which
    // made the expansion issue-bound at 0.1-1 % selectivity (profiles/r02_ncu_pos_expand_sel001_*.txt).
    // Here the lanes' non-zero 128-bit pieces are compacted first (one ballot per round places them in
    // word order; a segment has 256 pieces, so they always fit the stage), then lane i expands item i:
    // work follows the matches, not the words.
    if (my_cnt <= (uint32_t)kPosSparse) {
        uint4* item_words = reinterpret_cast<uint4*>(st);  // [<= 256]
        uint8_t* item_piece = piece_id[warp];              // [<= 256] round * 32 + lane
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t items = 0;
#pragma unroll
        for (int r = 0; r < kPosRounds; r++) {
            const bool any = (v[r].x | v[r].y | v[r].z | v[r].w) != 0;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, any);
            if (any) {
                const uint32_t slot = items + __popc(m & lt);
                item_words[slot] = v[r];
                item_piece[slot] = (uint8_t)(r * 32 + lane);
            }
            items += __popc(m);
        }
        __syncwarp();
        for (uint32_t i0 = 0; i0 < items; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint4 q = make_uint4(0, 0, 0, 0);
            uint32_t piece = 0;
            if (i < items) {
                q = item_words[i];
                piece = item_piece[i];
            }
            const uint32_t pc = popc4(q);
            uint32_t incl = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            uint64_t dst = o + (incl - pc);
            const uint64_t code = seg_code0 + piece * 128u;  // piece = (round, lane): words 128 r + 4 lane .. + 3
            const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t bits = x[j];
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (dst < capacity) __stcs(positions + dst, code + j * 32 + b);
                    dst++;
                }
            }
            o += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        return;
    }

    // two rounds share a scan: popcounts of a lane and round are <= 128, of a round <= 4096
    uint32_t pc2[kPosRounds / 2], incl2[kPosRounds / 2];
#pragma unroll
    for (int r2 = 0; r2 < kPosRounds / 2; r2++) {
        pc2[r2] = popc4(v[2 * r2]) | (popc4(v[2 * r2 + 1]) << 16);
        uint32_t incl = pc2[r2];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        incl2[r2] = incl;
    }

    uint32_t fill = 0;  // positions in the stage (warp-uniform)

    auto flush = [&]() {
        if (fill == 0) return;
        __syncwarp();
        for (uint32_t i = lane; i < fill; i += 32) {
            const uint64_t dst = o + i;
            if (dst < capacity) __stcs(positions + dst, seg_code0 + st[i]);
        }
        o += fill;
        fill = 0;
        __syncwarp();
    };
    auto expand = [&](const uint4& q, uint32_t slot, uint32_t code0) {
        const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t bits = x[j];
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                st[slot++] = (uint16_t)(code0 + j * 32 + b);
            }
        }
    };

#pragma unroll
    for (int r = 0; r < kPosRounds; r++) {
        const int sh = 16 * (r & 1);
        const uint32_t mine = (pc2[r >> 1] >> sh) & 0xFFFFu;
        const uint32_t excl = ((incl2[r >> 1] >> sh) & 0xFFFFu) - mine;
        const uint32_t round_total = (__shfl_sync(0xFFFFFFFFu, incl2[r >> 1], 31) >> sh) & 0xFFFFu;
        if (round_total == 0) continue;  // warp-uniform
        const uint32_t code0 = (uint32_t)(r * kPosRoundWords + lane * 4) * 32u;  // within the segment
        if (round_total <= (uint32_t)kPosStage) {
            if (fill + round_total > (uint32_t)kPosStage) flush();
            expand(v[r], fill + excl, code0);
            fill += round_total;
        } else {
            // more than half of the round's bits are set: the warp takes the round's words one at a
            // time, lane b owns bit b and stores its position at the rank of its bit
            flush();
            const uint32_t x[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll 1
            for (int src = 0; src < 32; src++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t word = __shfl_sync(0xFFFFFFFFu, x[j], src);
                    const uint64_t dst = o + __popc(word & ((1u << lane) - 1u));
                    if (((word >> lane) & 1u) && dst < capacity)
                        __stcs(positions + dst, seg_code0 + (uint32_t)(r * kPosRoundWords + src * 4 + j) * 32u + lane);
                    o += __popc(word);
                }
            }
        }
    }
    flush();
}

uint32_t* positions_seg_counts(void* workspace, uint64_t n)
{
    return reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + pos_groups(n) * 16);
}
size_t positions_seg_count_bytes(uint64_t n) { return (size_t)pos_segments(n) * 4; }

cudaError_t launch_positions(const Device& dev, const uint32_t* bitvec, uint64_t n, uint64_t base,
                             unsigned long long* positions, uint64_t capacity, unsigned long long* num,
                             void* workspace, uint32_t /*epoch*/, cudaStream_t stream, bool counted)
{
    const uint64_t segments = pos_segments(n);
    if (segments == 0) return num ? cudaMemsetAsync(num, 0, sizeof(unsigned long long), stream) : cudaSuccess;
    unsigned long long* group_off = static_cast<unsigned long long*>(workspace);
    uint32_t* seg_count = positions_seg_counts(workspace, n);
    const unsigned grid = (unsigned)((segments + kWarps - 1) / kWarps);
    // the three launches of a list are 5-60 us each: their CTAs are scheduled while the previous kernel drains
    // (programmatic stream serialization; every kernel waits in pdl_enter() before it touches memory)
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = dev.pdl ? 1 : 0;
    cudaError_t e = cudaSuccess;
    // counted: the scan that wrote the bitvector has filled seg_count itself (scan.cu, kSeg)
    if (!counted) {
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kThreads);
        e = cudaLaunchKernelEx(&cfg, pos_count_kernel, bitvec, n, segments, seg_count);
        if (e != cudaSuccess) return e;
    }
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kPosScanThreads);
    e = cudaLaunchKernelEx(&cfg, pos_offsets_kernel, (const uint32_t*)seg_count, segments, group_off, num);
    if (e != cudaSuccess) return e;
    if (capacity != 0) {
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kThreads);
        e = cudaLaunchKernelEx(&cfg, pos_expand_kernel, bitvec, n, segments, (const uint32_t*)seg_count,
                               (const unsigned long long*)group_off, base, positions, capacity);
    }
    return e;
}

}  // namespace sss
