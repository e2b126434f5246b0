// shared.cu -- shared scan: K range predicates evaluated in ONE pass over the packed column,
// producing K bitvectors and K match counts.
//
// Replaces shared_scan_128_standard / _standard_unrolled / _parallel / _sequential* / _threaded and
// the 256-bit twins (reference src/simd_scan_shared.cpp:9-466).  Like shared_scan_128_standard
// (:49-86) the column is decoded once for all predicates.  Unlike it, the predicates are not looped
// over per code: the K bounds are compiled into a predicate index (predindex.cuh) that maps a code
// to a precomputed K-bit row "which predicates match" with two shared-memory loads, and the
// 32 codes x K bits a thread collects for its 32-code block are turned into K result words with
// 32x32 bit-matrix transposes in registers.  Cost per code at K = 64: ~30 integer ops instead of
// ~130 for the compare loop, which is what moves this kernel from issue-bound towards HBM-bound.
// Each predicate's result is a whole 32-bit word per block and lanes write consecutive words of
// the same bitvector (128-byte lines) instead of the reference's K scattered byte stores (:82).
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 + K/8.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "predindex.cuh"

namespace sss {

// Threads per CTA of the index-driven kernels: ONE 512-thread CTA per SM.  The kernels are bound by
// instruction issue and shared-memory wavefronts, not by occupancy, and a single CTA per SM keeps
// one copy of the locator (32 KB) and leaves room for 32 lane-private replicas of the row table.
#ifndef SSS_SHARED_THREADS
#define SSS_SHARED_THREADS 512
#endif
constexpr int kSThreads = SSS_SHARED_THREADS;

template <int W>
struct SharedCfg {
    static constexpr int kBlocksPerTile = kSThreads;  // one 32-code block per thread per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    // compute-bound kernels: a shallow ring is enough and leaves shared memory for the row table
    static constexpr int kStages = kTileBytes >= 32 * 1024 ? 2 : 3;
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
};

// Up to kDirectCountMaxK predicates the kernel counts matches by popcounting its result words
// (K/32 instructions per code) and adds them straight to the caller's counters; above, counting goes
// through the interval histogram (one shared-memory atomic per code, independent of K).
static_assert(kDirectCountMaxK == 8, "the K <= 8 path of shared_scan_kernel is the direct-count path");

// kSmem: merged row table in shared memory (otherwise the masks are read through L1 from global
// memory and the interval histogram is a separate shared-memory array)
template <int W, bool kSmem>
__global__ void __launch_bounds__(kSThreads, 1)
shared_scan_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                   const uint8_t* __restrict__ g_index, int K, uint32_t* __restrict__ bitvecs, uint64_t stride_words,
                   unsigned long long* __restrict__ counts, bool count_matches, const IndexSmemPlan plan)
{
    using Cfg = SharedCfg<W>;
    using Rot = BlockRot<W, 1>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_index = smem + Cfg::kRingBytes + plan.hist_bytes;
    const IntervalHist hist = hist_init(smem + Cfg::kRingBytes, plan);  // kSmem: empty, the rows carry the counters
    const PredIndexView ix = load_pred_index(g_index, s_index, plan);
    typename Cfg::Ring ring(smem, packed, readable_bytes16);
    ring.start();  // contains the __syncthreads that publishes the index and the zeroed counters

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const int cols = (K + 31) >> 5;
    const uint32_t lane_rot = Rot::lane_rot();
    uint32_t direct[kDirectCountMaxK] = {};  // K <= 8: this thread's match counts

    uint32_t iter = 0;
    for (uint64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++iter) {
        ring.wait(iter);
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + threadIdx.x;
        const bool active = gb < n_blocks;
        const bool last = tail_bits != 0 && gb == n_blocks - 1;  // the one ragged block of the column
        // register j holds code (j + rot * kPieceCodes) mod 32; the ragged block is read unrotated
        const uint32_t rot = last ? 0u : lane_rot;
        uint32_t x[W];
        load_block<W, 1>(buf, threadIdx.x, x, rot);
        uint32_t loaded = 0;
#pragma unroll
        for (int k = 0; k < W; k++) loaded |= x[k];
        ring.release_warp(iter, tile, loaded);  // the stage is free once every warp has copied its words
        if (!active) continue;

        const uint32_t unrot = rot * Rot::kPieceCodes;
        const uint32_t keep = last ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;
        uint32_t id[32];
        block_ids<W>(x, ix, id);
        uint32_t* p = bitvecs + gb;  // this block's word of predicate 0; += stride_words per predicate
        if (K <= kDirectCountMaxK) {  // one byte of predicates per code: the cheap 32 x 8 transpose
            rows_of_ids<kSmem>(ix, id);
            uint32_t m[32], r[8];
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], 0);  // (no counter column: K <= 8)
            transpose32x8(m, r);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < K) {
                    uint32_t v = r[k];
                    if constexpr (Rot::kRot) v = __funnelshift_l(v, v, unrot);
                    v &= keep;
                    direct[k] += __popc(v);
                    st_stream_u32(p, v);
                }
                p += stride_words;
            }
            continue;
        }
        if (!kSmem && count_matches) {  // separate histogram, indexed by the interval id
            if (!last) {
#pragma unroll
                for (int j = 0; j < 32; j++) hist.add(id[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if ((uint32_t)j < tail_bits) hist.add(id[j]);
            }
        }
        rows_of_ids<kSmem>(ix, id);
        if (kSmem && count_matches) {  // the rows carry their interval's counter
            if (!last) {
#pragma unroll
                for (int j = 0; j < 32; j++) row_count(id[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if ((uint32_t)j < tail_bits) row_count(id[j]);
            }
        }
        for (int c = 0; c < cols; c++) {
            uint32_t m[32];
            // derived from the kernel parameter so that it stays in a uniform register: LDS [R + UR]
            const uint32_t col_bytes = kSmem ? (((uint32_t)c + plan.has_counter) << (plan.rep_log2 + 2)) : (uint32_t)c * 4u;
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], col_bytes);
            transpose32(m);  // m[k'] bit j = predicate 32c+k' on code j
            const int kk = K - 32 * c;  // predicates in this column
            if (kk >= 32 && !last && !Rot::kRot) {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    st_stream_u32(p, m[k]);
                    p += stride_words;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    uint32_t v = m[k];
                    if constexpr (Rot::kRot) v = __funnelshift_l(v, v, unrot);
                    if (k < kk) st_stream_u32(p, v & keep);
                    p += stride_words;
                }
            }
        }
    }
    __syncthreads();
    if (count_matches) {
        if (K <= kDirectCountMaxK) {
#pragma unroll
            for (int k = 0; k < kDirectCountMaxK; k++) {
                const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, direct[k]);
                if ((threadIdx.x & 31) == 0 && k < K && s != 0) atomicAdd(&counts[k], (unsigned long long)s);
            }
        } else if (kSmem) {
            merged_hist_flush(ix, plan, s_index, g_index);
        } else {
            hist_flush(smem + Cfg::kRingBytes, plan, ix.num_bp, g_index);
        }
    }
}

template <int W, bool kSmem>
static cudaError_t launch_ws(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                             uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                             const IndexSmemPlan& plan, cudaStream_t stream)
{
    using Cfg = SharedCfg<W>;
    const size_t smem = Cfg::kRingBytes + plan.total_bytes;
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_kernel<W, kSmem>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_kernel<W, kSmem><<<grid, kSThreads, smem, stream>>>(packed, readable, n,
                                                                   static_cast<const uint8_t*>(index), K, bitvecs,
                                                                   stride_words, counts, count_matches, plan);
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                            uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                            cudaStream_t stream)
{
    const IndexSmemPlan plan =
        plan_index_smem(K, W, SharedCfg<W>::kRingBytes, /*use_hist=*/K > kDirectCountMaxK, /*merged_ok=*/true);
    if (plan.masks_in_smem)
        return launch_ws<W, true>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
    return launch_ws<W, false>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const void*, int, uint32_t*, uint64_t,
                                 unsigned long long*, bool, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_shared_scan(const Device& dev, const void* packed, uint64_t n, int w, const void* index, int K,
                               uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                               cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, stream);
}

}  // namespace sss
