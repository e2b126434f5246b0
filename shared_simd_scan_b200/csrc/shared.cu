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
#include <cstdlib>
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
#ifndef SSS_STORE_FMA
#define SSS_STORE_FMA 0
#endif
// knock-out switches for bottleneck hunting (results are WRONG with any of them set; never shipped):
// SSS_KO_STORES: only predicate 0 of a column is stored; SSS_KO_ATOMS: no interval counting;
// SSS_KO_TRANSPOSE: mask words stored untransposed; SSS_KO_PLAINST: plain instead of streaming stores
#ifndef SSS_KO_STORES
#define SSS_KO_STORES 0
#endif
#ifndef SSS_KO_ATOMS
#define SSS_KO_ATOMS 0
#endif
#ifndef SSS_KO_TRANSPOSE
#define SSS_KO_TRANSPOSE 0
#endif
#ifndef SSS_KO_PLAINST
#define SSS_KO_PLAINST 0
#endif
// SSS_KO_LOCATE: row handles from the code's low bits, no locator lookup; SSS_KO_MASKLD: no mask loads
#ifndef SSS_KO_LOCATE
#define SSS_KO_LOCATE 0
#endif
#ifndef SSS_KO_MASKLD
#define SSS_KO_MASKLD 0
#endif
#if SSS_KO_PLAINST
#define SSS_ST(p, v) (*(p) = (v))
#else
#define SSS_ST(p, v) st_stream_u32((p), (v))
#endif
// the wide stores are PLAIN stores: with 64 and more output streams per CTA the evict-first hint of a
// streaming store measured 1.5-3.5 % slower (profiles/r02_ab_plainst_wide.txt, 20-24 bit)
#ifndef SSS_WIDE_STREAMING
#define SSS_WIDE_STREAMING 0
#endif
#if SSS_WIDE_STREAMING
#define SSS_ST4(p, v) __stcs((p), (v))
#else
#define SSS_ST4(p, v) (*(p) = (v))
#endif

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

// Wide stores (kWide).  A thread's 32 result words of a column belong to 32 different bitvectors, so
// the natural store is one 32-bit word per thread and predicate: 128 bytes per warp instruction.  The
// SM sustains far fewer bytes of such narrow stores next to other traffic than of 128-bit stores
// (tools/probes/stream_pattern.py: 64 output streams + the column's reads: 4.3 TB/s with 32-bit stores,
// 5.5 TB/s with 128-bit stores; knocking the stores out of this kernel made it 1.65x faster while
// knocking out the transposes, the lookups or the counting changed nothing, profiles/r02_knockouts.txt).
// So a warp parks the 32 x 32 words of a column in a warp-private shared-memory stage [predicate][block]
// (conflict-free 32-bit stores), and reads them back as 128-bit pieces along the BLOCK axis: lane
// (k mod 4, q) gets words 4q..4q+3 of predicate k and writes them with ONE 128-bit store -- 8 store
// instructions per column and thread instead of 32, four full 128-byte lines each.
// kStage = predicates a warp parks per round (32 / kStage rounds per column): 32 -> 4 KB per warp (64 KB per
// CTA), 16 -> 2 KB (32 KB); 0 = narrow stores.  The stage competes with the lane-private table replicas for
// shared memory, so the launcher picks the largest stage that costs no (32) or at most one (16) halving
// of the row-table replicas (profiles/r02_ab_wide_v2.txt, K = 64: 17-19 bit 0.73 -> 0.80, 20-31 bit 0.75 ->
// 0.78-0.82, 10-13 bit 0.71 -> 0.79; K = 256 / 512: 0.61 -> 0.70).
SSS_HD constexpr int stage_bytes(int k_stage) { return (kSThreads / 32) * k_stage * 32 * 4; }

// kSmem: merged row table in shared memory (otherwise the masks are read through L1 from global
// memory and the interval histogram is a separate shared-memory array)
template <int W, bool kSmem, int kStage>
__global__ void __launch_bounds__(kSThreads, 1)
shared_scan_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                   const uint8_t* __restrict__ g_index, int K, uint32_t* __restrict__ bitvecs, uint64_t stride_words,
                   unsigned long long* __restrict__ counts, bool count_matches, const IndexSmemPlan plan,
                   unsigned long long* __restrict__ tickets)
{
    using Cfg = SharedCfg<W>;
    using Rot = BlockRot<W, 1>;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool kWide = kStage > 0;
    constexpr int kStagePreds = kStage;
    constexpr int kFront = Cfg::kRingBytes + stage_bytes(kStage);  // ring, then the wide-store stage
    uint32_t* const stage = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes) + (threadIdx.x >> 5) * (kStage * 32);
    uint8_t* s_index = smem + kFront + plan.hist_bytes;
    const IntervalHist hist = hist_init(smem + kFront, plan);  // kSmem: empty, the rows carry the counters
    const PredIndexView ix = load_pred_index(g_index, s_index, plan);
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();  // contains the __syncthreads that publishes the index and the zeroed counters

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const int cols = (K + 31) >> 5;
    const uint32_t lane_rot = Rot::lane_rot();
    const bool narrow = stride_words < (1ull << 30);  // bitvector stride below 4 GiB
    const uint32_t stride_bytes32 = (uint32_t)(stride_words * 4);
    uint32_t direct[kDirectCountMaxK] = {};  // K <= 8: this thread's match counts

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
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
        ring.release_warp(iter, loaded);  // the stage is free once every warp has copied its words
        // kWide: the warp exchanges its result words through shared memory, so lanes past the column's end
        // stay with their warp (their lookups run on whatever the tile holds -- every index stays inside the
        // tables --, they count nothing and store nothing)
        const uint64_t gb_warp0 = gb - (threadIdx.x & 31);
        if (kWide ? gb_warp0 >= n_blocks : !active) continue;

        const uint32_t unrot = rot * Rot::kPieceCodes;
        const uint32_t keep = last ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;
        uint32_t id[32];
#if SSS_KO_LOCATE
#pragma unroll
        for (int j = 0; j < 32; j++) id[j] = ((x[j % W] >> (j & 15)) & 63u) * ix.row_bytes + ix.rows_saddr;
#else
        if constexpr (kSmem) block_rows<W>(x, ix, id);  // row handles directly
        else block_ids<W>(x, ix, id);
#endif
        uint32_t* p = bitvecs + gb;  // this block's word of predicate 0; += stride_words per predicate
        if (K <= kDirectCountMaxK) {  // one byte of predicates per code: the cheap 32 x 8 transpose
            if constexpr (!kSmem) rows_of_ids<kSmem>(ix, id);
            uint32_t m[32], r[8];
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], 0);  // (no counter column: K <= 8)
            transpose32x8(m, r);
            if (kWide && !active) continue;
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
        if (!kSmem && count_matches && active) {  // separate histogram, indexed by the interval id
            if (!last) {
#pragma unroll
                for (int j = 0; j < 32; j++) hist.add(id[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if ((uint32_t)j < tail_bits) hist.add(id[j]);
            }
        }
        if constexpr (!kSmem) rows_of_ids<kSmem>(ix, id);
        if (kSmem && count_matches && active && !SSS_KO_ATOMS) {  // the rows carry their interval's counter
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
#if SSS_KO_MASKLD
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = id[j] * 2654435761u + col_bytes;
#else
#pragma unroll
            for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], col_bytes);
#endif
#if !SSS_KO_TRANSPOSE
            transpose32(m);  // m[k'] bit j = predicate 32c+k' on code j
#endif
            const int kk = K - 32 * c;  // predicates in this column
#if SSS_KO_STORES
            {
                uint32_t acc = 0;
#pragma unroll
                for (int k = 0; k < 32; k++) acc ^= m[k];
                SSS_ST(p, acc);
                p += 32 * stride_words;
                continue;
            }
#endif
            if constexpr (kWide) {
                const uint32_t lane = threadIdx.x & 31;
                const uint32_t live = active ? keep : 0u;
                const uint32_t kq = lane >> 3, q4 = (lane & 7) * 4;
                const uint64_t gbq = gb_warp0 + q4;  // first of the 4 consecutive blocks this lane stores
                uint32_t* pw = bitvecs + (uint64_t)(32 * c + (int)kq) * stride_words + gbq;
                const uint32_t* src = stage + kq * 32 + q4;
                const bool whole = kk >= 32 && gb_warp0 + 32 <= n_blocks;  // warp-uniform
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += (kWide ? kStagePreds : 32)) {  // rounds of kStagePreds predicates
                    __syncwarp();  // the previous round's wide reads of the stage are done
#pragma unroll
                    for (int k = 0; k < kStagePreds; k++) {
                        uint32_t v = m[r0 + k];
                        if constexpr (Rot::kRot) v = __funnelshift_l(v, v, unrot);
                        stage[k * 32 + lane] = v & live;
                    }
                    __syncwarp();
                    if (whole) {
#pragma unroll
                        for (int i = 0; i < kStagePreds / 4; i++) {
                            SSS_ST4(reinterpret_cast<uint4*>(pw), *reinterpret_cast<const uint4*>(src + i * 128));
                            pw += 4 * stride_words;
                        }
                    } else {  // the last predicates of a ragged K, or the last warp of the column
#pragma unroll
                        for (int i = 0; i < kStagePreds / 4; i++) {
                            const uint4 v = *reinterpret_cast<const uint4*>(src + i * 128);
                            if (r0 + (int)kq + 4 * i < kk) {
                                if (gbq + 4 <= n_blocks) {
                                    SSS_ST4(reinterpret_cast<uint4*>(pw), v);
                                } else {
                                    if (gbq + 0 < n_blocks) st_stream_u32(pw + 0, v.x);
                                    if (gbq + 1 < n_blocks) st_stream_u32(pw + 1, v.y);
                                    if (gbq + 2 < n_blocks) st_stream_u32(pw + 2, v.z);
                                }
                            }
                            pw += 4 * stride_words;
                        }
                    }
                }
                continue;
            }
            if (kk >= 32 && !last && !Rot::kRot && narrow) {
                // address of predicate k's word = p + k * stride: one IMAD.WIDE each (FMA pipe) instead of
                // a 64-bit add chain (two ALU instructions per store)
#if SSS_STORE_FMA
#pragma unroll
                for (int k = 0; k < 32; k++) st_stream_u32(ptr_mad_fma(p, stride_bytes32, k), m[k]);
#else
#pragma unroll
                for (int k = 0; k < 32; k++)
                    SSS_ST(reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(p) + (uint64_t)stride_bytes32 * (uint32_t)k),
                           m[k]);
#endif
                p += 32 * stride_words;
            } else if (kk >= 32 && !last && !Rot::kRot) {
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
    ring.finish();  // (contains the __syncthreads that closes the counting)
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
            hist_flush(smem + kFront, plan, ix.num_bp, g_index);
        }
    }
}

// ---- many predicates (K > kWideMinK): one warp per 32-code block, one lane per 32-predicate column ----
//
// With more than ~16 columns the K-bit rows no longer fit in shared memory at useful widths, and the
// thread-per-block mapping would read them from global memory one word per lane from 32 different
// rows.  Here, as in the wide linear kernel, lane j first locates code j of the warp's block, the 32
// row handles are broadcast with shuffles, and lane c reads column c of the 32 rows -- consecutive
// words of one row across the lanes, a coalesced 128-byte read per code.  After the 32x32 transpose
// lane c holds the block's result words of predicates 32c..32c+31: words of 32 different bitvectors.
// Stored directly that would be K scattered 4-byte writes per block, so the 16 warps of a CTA work on
// 16 CONSECUTIVE blocks, park their words in a shared-memory stage [predicate][block], and the CTA
// then writes every predicate's 16 words as one 64-byte run (a half-warp per predicate).
constexpr int kWideMinK = 512;
constexpr int kWideThreads = 512;
constexpr int kWideWarps = kWideThreads / 32;  // blocks per step
constexpr int kWideRowPad = kWideWarps + 1;  // odd pitch: conflict-free for the lane = column accesses
static_assert(kWideWarps == 16, "the drain below maps a half-warp to a predicate's 16-word run");

template <int W>
struct SharedWideCfg {
    static constexpr int kBlocksPerTile = 256;
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    static constexpr int kStages = kTileBytes >= 16 * 1024 ? 2 : 3;
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
    static constexpr int kSteps = kBlocksPerTile / kWideWarps;
};
SSS_HD inline size_t wide_stage_bytes(int K) { return (size_t)32 * ((K + 31) / 32) * kWideRowPad * 4; }

// interval id of code `j` (runtime) of a block whose W words start at `blk` in shared memory
template <int W>
__device__ __forceinline__ uint32_t wide_code_id(const uint32_t* blk, uint32_t j, const PredIndexView& ix)
{
    const uint32_t bit = j * W, wi = bit >> 5;
    const uint32_t v = __funnelshift_r(blk[wi], blk[wi + 1 < (uint32_t)W ? wi + 1 : wi], bit & 31) & code_mask(W);
    if constexpr (W <= kDirectMaxW) {
        return static_cast<const uint32_t*>(ix.locator)[v];
    } else {
        const uint32_t t = v << (32 - W);
        const uint32_t e = static_cast<const uint32_t*>(ix.locator)[t >> (32 - kBucketBits)];
        uint32_t id = (e & kBucketIdMask) + ((t << kBucketBits) > e ? 1u : 0u);
        if (e & kSeveralBit) {
            uint32_t i = e & kBucketIdMask;
            while (t > ix.thr[i]) i++;
            id = i;
        }
        return id;
    }
}

template <int W, bool kSmem>
__global__ void __launch_bounds__(kWideThreads, 1)
shared_scan_wide_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                        const uint8_t* __restrict__ g_index, int K, uint32_t* __restrict__ bitvecs,
                        uint64_t stride_words, bool count_matches, const IndexSmemPlan plan,
                        unsigned long long* __restrict__ tickets)
{
    using Cfg = SharedWideCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes);
    uint8_t* s_hist = smem + Cfg::kRingBytes + wide_stage_bytes(K);
    const IntervalHist hist = hist_init(s_hist, plan);
    const PredIndexView ix = load_pred_index(g_index, s_hist + plan.hist_bytes, plan);
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t cols = ((uint32_t)K + 31) >> 5;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_base = kSmem ? ix.rows_saddr : 0u;
    const uint32_t my_col = lane < cols ? lane : 0;  // lanes past the last column idle along

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        uint32_t dep = 0;
        for (uint32_t step = 0; step < (uint32_t)Cfg::kSteps; step++) {
            const uint64_t gb0 = tile * (uint64_t)Cfg::kBlocksPerTile + step * kWideWarps;  // first block of the step
            if (gb0 >= n_blocks) break;  // CTA-uniform
            const uint32_t blk = step * kWideWarps + warp;
            const uint64_t gb = gb0 + warp;
            if (gb < n_blocks) {  // warp-uniform
                const uint32_t valid = (tail_bits != 0 && gb == n_blocks - 1) ? tail_bits : 32u;
                const uint32_t id = wide_code_id<W>(buf + blk * W, lane, ix);
                dep |= id;
                if (count_matches && lane < valid) hist.add(id);
                const uint32_t row = id * ix.row_bytes + rows_base;
                uint32_t m[32];
#pragma unroll
                for (int j = 0; j < 32; j++)
                    m[j] = row_mask_word<kSmem>(ix, __shfl_sync(0xFFFFFFFFu, row, j), 4 * my_col);
                if (valid != 32) {  // only the column's last block: codes past n contribute zero bits
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if ((uint32_t)j >= valid) m[j] = 0;
                }
                transpose32(m);  // m[k] bit j = predicate 32*lane+k on code j
                if (lane < cols) {
#pragma unroll
                    for (int k = 0; k < 32; k++) stage[(k * cols + lane) * kWideRowPad + warp] = m[k];
                }
            }
            __syncthreads();
            // drain: warp w owns columns 2w and 2w+1; its two half-warps walk the 32 predicates of one
            // column each, lane i of a half writing word i of the predicate's 16-word run.  A store
            // instruction so touches two bitvectors (two pages) -- with one bitvector per LANE (32
            // pages per instruction, 256-bit stores) the kernel was bound by address translation.
            const uint64_t left = n_blocks - gb0;
            const uint32_t words = left < (uint64_t)kWideWarps ? (uint32_t)left : (uint32_t)kWideWarps;
            {
                const uint32_t c = 2 * warp + (lane >> 4), i = lane & 15;
                const int kk = K - 32 * (int)c;  // predicates in my column (<= 0: none)
                if (c < cols && i < words) {
                    const uint32_t* src = stage + c * kWideRowPad + i;
                    uint32_t* dst = bitvecs + (uint64_t)(32 * c) * stride_words + gb0 + i;
#pragma unroll 8
                    for (int k = 0; k < 32; k++) {
                        if (k < kk) st_stream_u32(dst, src[0]);
                        src += cols * kWideRowPad;
                        dst += stride_words;
                    }
                }
            }
            __syncthreads();
        }
        ring.release_warp(iter, dep);
    }
    ring.finish();
    if (count_matches) hist_flush(s_hist, plan, ix.num_bp, g_index);
}

template <int W, bool kSmem>
static cudaError_t launch_wide(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                               uint32_t* bitvecs, uint64_t stride_words, bool count_matches, const IndexSmemPlan& plan,
                               cudaStream_t stream)
{
    using Cfg = SharedWideCfg<W>;
    const size_t smem = Cfg::kRingBytes + wide_stage_bytes(K) + plan.total_bytes;
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_wide_kernel<W, kSmem>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_wide_kernel<W, kSmem><<<grid, kWideThreads, smem, stream>>>(
        packed, readable, n, static_cast<const uint8_t*>(index), K, bitvecs, stride_words, count_matches, plan,
        dev.next_tickets());
    return cudaGetLastError();
}

template <int W, bool kSmem, int kStage>
static cudaError_t launch_ws(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                             uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                             const IndexSmemPlan& plan, cudaStream_t stream)
{
    using Cfg = SharedCfg<W>;
    const size_t smem = Cfg::kRingBytes + stage_bytes(kStage) + plan.total_bytes;
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_kernel<W, kSmem, kStage>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_kernel<W, kSmem, kStage><<<grid, kSThreads, smem, stream>>>(
        packed, readable, n, static_cast<const uint8_t*>(index), K, bitvecs, stride_words, counts, count_matches, plan,
        dev.next_tickets());
    return cudaGetLastError();
}

// Wide stores need every bitvector 16-byte aligned (base and stride) and at least one full column of
// predicates; below 10 bits the kernel is closer to its instruction budget than to its store budget and
// the narrow stores measured 1-2 % faster.
//
// OFF since the persistent CTAs draw their tiles from a ticket counter.  The wide stores were measured
// and tuned under the static tile order (K = 64 at 17 bit: 0.72 -> 0.80).  With the dynamic order the
// plain 32-bit stores -- one word per thread and predicate, no stage, no second trip through shared
// memory -- are faster at EVERY width (profiles/r02_ab_wide_after_tickets.txt, K = 64: 10-13 bit 0.79 ->
// 0.85, 17 bit 0.80 -> 0.86, 23 bit 0.85 -> 0.88, 28-32 bit 0.76-0.81 -> 0.88-0.91; K = 1024 +0-5 points):
// CTAs no longer walk the 64 output streams in lockstep, which is what the narrow stores suffered from.
// -DSSS_SHARED_WIDE=1 rebuilds the wide path (kept, parity-tested by tools/ab_variants.sh CHECK=1).
#ifndef SSS_SHARED_WIDE
#define SSS_SHARED_WIDE 0
#endif
#ifndef SSS_WIDE_MIN_K
#define SSS_WIDE_MIN_K 32
#endif
#ifndef SSS_WIDE_MIN_W
#define SSS_WIDE_MIN_W 10
#endif
static bool wide_ok(int W, int K, const uint32_t* bitvecs, uint64_t stride_words)
{
    return SSS_SHARED_WIDE && W >= SSS_WIDE_MIN_W && K >= SSS_WIDE_MIN_K && (reinterpret_cast<uintptr_t>(bitvecs) & 15) == 0 &&
           (stride_words & 3) == 0;
}

template <int W>
static IndexSmemPlan thread_plan(int K, int k_stage)
{
    return plan_index_smem(K, W, SharedCfg<W>::kRingBytes + stage_bytes(k_stage), /*use_hist=*/K > kDirectCountMaxK,
                           /*merged_ok=*/true);
}

// the stage to use for K predicates at W bits (0: narrow stores) and the plan that goes with it
template <int W>
static int pick_stage(int K, bool wide_possible, IndexSmemPlan& plan)
{
    plan = thread_plan<W>(K, 0);
    if (!wide_possible || !plan.merged) return 0;
    const IndexSmemPlan p32 = thread_plan<W>(K, 32);
    if (p32.merged && p32.rep_log2 == plan.rep_log2) {
        plan = p32;
        return 32;
    }
    const IndexSmemPlan p16 = thread_plan<W>(K, 16);
    if (p16.merged && p16.rep_log2 + 1 >= plan.rep_log2) {
        plan = p16;
        return 16;
    }
    return 0;
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                            uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                            cudaStream_t stream)
{
    IndexSmemPlan plan;
    const int k_stage = pick_stage<W>(K, wide_ok(W, K, bitvecs, stride_words), plan);
    // many predicates whose row table does not fit in shared memory: the warp-per-block kernel
    // (measured at K = 1024: 18 G codes/s at every width, against 26 G for the thread-per-block kernel
    // while the table fits -- W <= 9 -- and 5-12 G once it reads its rows from global memory)
    if (K > kWideMinK && !plan.merged) {
        const IndexSmemPlan wp = plan_index_smem(K, W, SharedWideCfg<W>::kRingBytes + wide_stage_bytes(K),
                                                 /*use_hist=*/true, /*merged_ok=*/false);
        if (wp.masks_in_smem)
            return launch_wide<W, true>(dev, packed, n, index, K, bitvecs, stride_words, count_matches, wp, stream);
        return launch_wide<W, false>(dev, packed, n, index, K, bitvecs, stride_words, count_matches, wp, stream);
    }
    if (plan.masks_in_smem) {
        if (k_stage == 32)
            return launch_ws<W, true, 32>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
        if (k_stage == 16)
            return launch_ws<W, true, 16>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
        return launch_ws<W, true, 0>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
    }
    return launch_ws<W, false, 0>(dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, plan, stream);
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const void*, int, uint32_t*, uint64_t,
                                 unsigned long long*, bool, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

// Predicate groups.  Once the merged row table of all K predicates no longer fits in shared memory
// the kernels above fall to 0.31-0.36 of the HBM peak (rows read through L1, or the warp-per-block
// kernel's 64-byte runs), while a table that fits runs at 0.60 whatever the width.  The bitvectors of
// different predicates are independent, so the device-resident entry point evaluates such a K as
// ceil(K / G) scans of G predicates each (own index, own bitvector slice, own counts): the column is
// read ceil(K / G) times, which at K / 8 output bytes per code is noise (K = 1024, 32 bit, G = 256:
// 16 B read against 128 B written per code).  Returns G (a multiple of 32), or K when the table fits
// or no smaller group would fit either.
template <int W>
static int group_w(int K)
{
    auto merged = [](int k) { return thread_plan<W>(k, 0).merged != 0; };
    if (K <= 128 || merged(K)) return K;
    for (int g = 512; g >= 128; g >>= 1)
        if (g < K && merged(g)) return g;
    return K;
}

template <int... Ws>
static constexpr auto make_group_table(std::integer_sequence<int, Ws...>)
{
    return std::array<int (*)(int), sizeof...(Ws)>{&group_w<Ws + 1>...};
}

int shared_scan_predicate_group(int w, int K)
{
    static constexpr auto table = make_group_table(std::make_integer_sequence<int, 32>{});
    // developer switch: SSS_PRED_GROUP=<multiple of 32> forces the group size (A/B runs)
    static const int forced = [] {
        const char* e = std::getenv("SSS_PRED_GROUP");
        const int g = e ? std::atoi(e) : 0;
        return g >= 32 && g % 32 == 0 ? g : 0;
    }();
    if (forced && forced < K) return forced;
    return table[w - 1](K);
}

cudaError_t launch_shared_scan(const Device& dev, const void* packed, uint64_t n, int w, const void* index, int K,
                               uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts, bool count_matches,
                               cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, index, K, bitvecs, stride_words, counts, count_matches, stream);
}

}  // namespace sss
