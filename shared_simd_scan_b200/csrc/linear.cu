// linear.cu -- shared scan with ONE linear output: the byte for (group g of 8 codes, predicate k)
// lives at out[g*K + k].
//
// Replaces shared_scan_128_linear_standard / shared_scan_128_linear_simple /
// shared_scan_128_linear_static<NUM> (reference src/simd_scan_shared_linear.cpp:9-82,
// src/simd_scan.hpp:122-236).  The layout is pinned by the reference's only shared-scan unit test
// (test/simd_scan_tests.cpp:108-150).
//
// Same predicate index as shared.cu (predindex.cuh): a code is mapped to its precomputed K-bit match
// row.  A thread owns one 32-code block = 4 output rows; for every column of 32 predicates it loads
// the 32 rows' words and runs a 5-stage exchange network in registers -- three bit stages (8x8 bit
// transposes inside every byte lane of every 8-register group) and two byte stages (PRMT) -- whose
// result registers ARE the output: 8 registers per row, each holding 4 consecutive predicate bytes.
// That is the same instruction count as the 32x32 transpose of shared.cu; the linear layout costs
// nothing extra.  Every row piece a thread writes is 32 contiguous bytes (one full sector).
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 + K/8.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "predindex.cuh"

namespace sss {

// developer switch: 0 = no row spreading (A/B runs; results stay correct)
#ifndef SSS_LINEAR_SPREAD
#define SSS_LINEAR_SPREAD 1
#endif

// Threads per CTA of the index-driven kernels: one 512-thread CTA per SM (see shared.cu).
#ifndef SSS_SHARED_THREADS
#define SSS_SHARED_THREADS 512
#endif
constexpr int kSThreads = SSS_SHARED_THREADS;

template <int W>
struct LinearCfg {
    static constexpr int kBlocksPerTile = kSThreads;
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    // compute-bound kernels: a shallow ring is enough and leaves shared memory for the row table
    static constexpr int kStages = kTileBytes >= 32 * 1024 ? 2 : 3;
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
};

// more than this many predicates: one warp per block (shared_scan_linear_wide_kernel)
#ifndef SSS_LINEAR_WIDE_MIN_K
#define SSS_LINEAR_WIDE_MIN_K (16 * 32)
#endif
constexpr int kLinearWideMinK = SSS_LINEAR_WIDE_MIN_K;
constexpr int kLinearWideFallbackK = 12 * 32;  // ... and from here on when the merged row table does not fit

// in : a[8i + e] byte p bit f  = predicate 8p+f on code 8i+e           (i = row, e = code in group)
// out: a[8i + 4*f2 + p] byte (f & 3) bit e, f2 = f >> 2                 = row i, bytes 8p+4*f2 .. +3
// i.e. output word q (predicates 4q..4q+3) of row i is a[8i + 4*(q & 1) + (q >> 1)].
__device__ __forceinline__ void exchange_linear(uint32_t (&a)[32])
{
#pragma unroll
    for (int g = 0; g < 32; g += 8)
#pragma unroll
        for (int k = g; k < g + 4; k++) exchange_bits<4>(a[k], a[k + 4]);  // e2 <-> f2
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) exchange_bits<2>(a[k], a[k + 2]);  // e1 <-> f1
#pragma unroll
    for (int k = 0; k < 32; k += 2) exchange_bits<1>(a[k], a[k + 1]);      // e0 <-> f0
    // now a[8i + f] byte p bit e.  Byte stages: register bit f1 <-> byte-index bit p1, then f0 <-> p0
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) {
            const uint32_t lo = a[k], hi = a[k + 2];
            a[k] = __byte_perm(lo, hi, 0x5410);
            a[k + 2] = __byte_perm(lo, hi, 0x7632);
        }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const uint32_t lo = a[k], hi = a[k + 1];
        a[k] = __byte_perm(lo, hi, 0x6240);
        a[k + 1] = __byte_perm(lo, hi, 0x7351);
    }
}

// 32 consecutive predicate bytes of one output row (8 words, already in row order); kk = predicates
// that exist in this column (>= 32: all)
__device__ __forceinline__ void store_row_piece(uint8_t* dst, const uint32_t (&ws)[8], int kk, bool vec32, bool vec16,
                                                bool vec4)
{
    if (vec32 && kk >= 32) {
        st_v8_u32(dst, ws);
        return;
    }
    int q0 = 0;
    if (vec16 && kk >= 16) {
        __stcs(reinterpret_cast<uint4*>(dst), make_uint4(ws[0], ws[1], ws[2], ws[3]));
        q0 = 4;
        if (kk >= 32) {
            __stcs(reinterpret_cast<uint4*>(dst) + 1, make_uint4(ws[4], ws[5], ws[6], ws[7]));
            return;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (q < q0) continue;
        if (4 * q + 4 <= kk && vec4) {
            __stcs(reinterpret_cast<uint32_t*>(dst) + q, ws[q]);
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (4 * q + b < kk) dst[4 * q + b] = (uint8_t)(ws[q] >> (8 * b));
        }
    }
}

// interval id of code `j` (runtime) of a block whose W words start at `blk` in shared memory
template <int W>
__device__ __forceinline__ uint32_t code_id_runtime(const uint32_t* blk, uint32_t j, const PredIndexView& ix)
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

// ---- many predicates (more than 16 columns of 32): one warp per 32-code block, one lane per column ----
//
// With a row pitch of K >= 544 bytes the thread-per-block mapping above makes every warp store hit
// 32 different rows.  Here lane j first locates code j of the block (one lookup per code, lane
// parallel), the 32 interval ids are broadcast with shuffles, lane c then reads column c of the 32
// mask rows (consecutive words of the same row across lanes: conflict-free), runs the exchange
// network and owns bytes 32c..32c+31 of the block's four output rows -- so a warp writes each row as
// one contiguous run of K bytes.
template <int W, bool kSmem>
__global__ void __launch_bounds__(kSThreads, 1)
shared_scan_linear_wide_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                               const uint8_t* __restrict__ g_index, int K, uint8_t* __restrict__ out,
                               bool count_matches, const IndexSmemPlan plan, unsigned long long* __restrict__ tickets)
{
    using Cfg = LinearCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    const IntervalHist hist = hist_init(smem + Cfg::kRingBytes, plan);
    const PredIndexView ix = load_pred_index(g_index, smem + Cfg::kRingBytes + plan.hist_bytes, plan);
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const int cols = (K + 31) >> 5;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rows_base = kSmem ? ix.rows_saddr : 0u;
    const bool vec32 = (K & 31) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0;
    const bool vec16 = (K & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const bool vec4 = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    const int my_col = (int)lane < cols ? (int)lane : 0;   // lanes past the last column idle along
    const int kk = K - 32 * (int)lane;                      // predicates in my column (<= 0: none)
    // rows read from global memory (through L1): my column of row 0
    const uint8_t* g_col = reinterpret_cast<const uint8_t*>(ix.masks) + 4 * my_col;

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        uint32_t dep = 0;
        for (uint32_t b = 0; b < 32; b++) {
            const uint32_t blk = warp * 32 + b;
            const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + blk;
            if (gb >= n_blocks) break;  // warp-uniform
            const uint32_t valid = (tail_bits != 0 && gb == n_blocks - 1) ? tail_bits : 32u;
            const uint32_t id = code_id_runtime<W>(buf + blk * W, lane, ix);
            dep |= id;
            if (count_matches && lane < valid) hist.add(id);
            uint32_t m[32];
            const uint32_t row_off = id * ix.row_bytes + rows_base;  // my code's mask row
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t r = __shfl_sync(0xFFFFFFFFu, row_off, j);
                // (measured: the IMAD.WIDE address pays from 16 bit on, 0.66 -> 0.69 of the HBM peak at
                //  K = 1024, and costs two points below, where the ring is three stages deep)
                if constexpr (kSmem || W < 16) m[j] = row_mask_word<kSmem>(ix, r, 4 * my_col);
                else m[j] = __ldg(reinterpret_cast<const uint32_t*>(ptr_add_fma(g_col, r)));
            }
            if (valid != 32) {  // only the column's last block: codes past n contribute zero bits
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if ((uint32_t)j >= valid) m[j] = 0;
            }
            exchange_linear(m);
            if (kk > 0) {
                uint8_t* row0 = out + gb * 4 * (uint64_t)K + 32 * lane;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t ws[8] = {m[8 * i + 0], m[8 * i + 4], m[8 * i + 1], m[8 * i + 5],
                                            m[8 * i + 2], m[8 * i + 6], m[8 * i + 3], m[8 * i + 7]};
                    store_row_piece(row0 + (uint64_t)i * K, ws, kk, vec32, vec16, vec4);
                }
            }
        }
        ring.release_warp(iter, dep);
    }
    ring.finish();
    if (count_matches) hist_flush(smem + Cfg::kRingBytes, plan, ix.num_bp, g_index);
}

template <int W, bool kSmem>
__global__ void __launch_bounds__(kSThreads, 1)
shared_scan_linear_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                          const uint8_t* __restrict__ g_index, int K, uint8_t* __restrict__ out,
                          unsigned long long* __restrict__ counts, bool count_matches, const IndexSmemPlan plan,
                          unsigned long long* __restrict__ tickets)
{
    using Cfg = LinearCfg<W>;
    using Rot = BlockRot<W, 8>;  // rotations by whole rows (8 codes) only
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_index = smem + Cfg::kRingBytes + plan.hist_bytes;
    const IntervalHist hist = hist_init(smem + Cfg::kRingBytes, plan);  // kSmem: empty, the rows carry the counters
    const PredIndexView ix = load_pred_index(g_index, s_index, plan);
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const int cols = (K + 31) >> 5;
    const uint32_t lane_rot = Rot::lane_rot();
    // widest store the row pitch K and the base alignment allow
    const bool vec32 = (K & 31) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0;
    const bool vec16 = (K & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const bool vec4 = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
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
        load_block<W, 8>(buf, threadIdx.x, x, rot);
        uint32_t loaded = 0;
#pragma unroll
        for (int k = 0; k < W; k++) loaded |= x[k];
        ring.release_warp(iter, loaded);  // the stage is free once every warp has copied its words
        if (active) {
            uint32_t id[32];
            if constexpr (kSmem) block_rows<W>(x, ix, id);  // row handles directly
        else block_ids<W>(x, ix, id);
            uint8_t* row0 = out + gb * 4 * (uint64_t)K;  // rows 4*gb .. 4*gb+3
            if (K <= kDirectCountMaxK) {
                // per-predicate result words as in shared.cu, then bytes -> rows
                if constexpr (!kSmem) rows_of_ids<kSmem>(ix, id);
                uint32_t m[32], r[8];
#pragma unroll
                for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], 0);  // (no counter column)
                transpose32x8(m, r);
                const uint32_t keep = last ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if constexpr (Rot::kRot) r[k] = __funnelshift_l(r[k], r[k], rot * Rot::kPieceCodes);
                    r[k] &= keep;
                    if (k < K) direct[k] += __popc(r[k]);
                }
                store_block_small(row0, r, K);
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
            if constexpr (!kSmem) rows_of_ids<kSmem>(ix, id);
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
            // Row spreading.  A thread's four rows are 32-byte pieces at a 4K-byte block stride; when all
            // lanes of a store instruction write the SAME row of their blocks, its 32 sectors land on one
            // offset modulo the stride and the launch loses a quarter of its speed (measured: the widths
            // whose rotated block loads happen to permute the rows over all four positions, 16 and 32
            // bit, ran 27 % faster than with the permutation switched off).  So every lane rotates its
            // 32 row handles by `extra` rows (8 codes each; the order of a block's codes is free until
            // the store) such that, together with the load rotation, lanes l, l+1, l+2, l+3 sit on four
            // different rows: two conditional-move stages, 2 instructions per code.
            constexpr int kLoadWays = Rot::kRot ? (Rot::kPieces >= 4 ? 4 : Rot::kPieces) : 1;  // rows the load rotation spreads over
            uint32_t extra = 0;
            if constexpr (kLoadWays < 4) {
                extra = (last || !SSS_LINEAR_SPREAD) ? 0u : ((threadIdx.x & 31u) & (uint32_t)(4 / kLoadWays - 1));
                if (extra & 1u) {  // rotate by one row: position j takes the handle of position j + 8
                    uint32_t t[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) t[j] = id[j];
#pragma unroll
                    for (int j = 0; j < 24; j++) id[j] = id[j + 8];
#pragma unroll
                    for (int j = 0; j < 8; j++) id[24 + j] = t[j];
                }
                if constexpr (kLoadWays == 1) {
                    if (extra & 2u) {  // rotate by two rows
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const uint32_t t = id[j];
                            id[j] = id[j + 16];
                            id[j + 16] = t;
                        }
                    }
                }
            }
            // register row i is output row (i + row_rot) & 3
            const uint32_t row_rot = rot * (Rot::kPieceCodes / 8) + extra;
            for (int c = 0; c < cols; c++) {
                uint32_t m[32];
                // derived from the kernel parameter so that it stays in a uniform register: LDS [R + UR]
                const uint32_t col_bytes =
                    kSmem ? (((uint32_t)c + plan.has_counter) << (plan.rep_log2 + 2)) : (uint32_t)c * 4u;
#pragma unroll
                for (int j = 0; j < 32; j++) m[j] = row_mask_word<kSmem>(ix, id[j], col_bytes);
                if (last) {  // codes past n contribute zero bits
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if ((uint32_t)j >= tail_bits) m[j] = 0;
                }
                exchange_linear(m);
                const int kk = K - 32 * c;  // predicates in this column
#pragma unroll
                for (int i = 0; i < 4; i++) {
#ifdef SSS_EXPERIMENT_NO_ROW_PERMUTE  // timing experiment only (wrong results): rows leave in register order
                    const uint32_t orow = (uint32_t)i;
#else
                    const uint32_t orow = (i + row_rot) & 3u;
#endif
                    uint8_t* dst = row0 + (uint64_t)orow * K + 32 * c;
                    // word q of the row = m[8i + 4*(q&1) + (q>>1)]
                    const uint32_t ws[8] = {m[8 * i + 0], m[8 * i + 4], m[8 * i + 1], m[8 * i + 5],
                                            m[8 * i + 2], m[8 * i + 6], m[8 * i + 3], m[8 * i + 7]};
                    store_row_piece(dst, ws, kk, vec32, vec16, vec4);
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
            hist_flush(smem + Cfg::kRingBytes, plan, ix.num_bp, g_index);
        }
    }
}

template <int W, bool kSmem>
static cudaError_t launch_ws(const Device& dev, const void* packed, uint64_t n, const void* index, int K, uint8_t* out,
                             unsigned long long* counts, bool count_matches, const IndexSmemPlan& plan, bool wide,
                             cudaStream_t stream)
{
    using Cfg = LinearCfg<W>;
    const size_t smem = Cfg::kRingBytes + plan.total_bytes;
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    if (wide) {
        static size_t configured_wide[64] = {};
        if (configured_wide[dev.ordinal & 63] < smem) {
            cudaError_t e = cudaFuncSetAttribute(shared_scan_linear_wide_kernel<W, kSmem>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured_wide[dev.ordinal & 63] = smem;
        }
        shared_scan_linear_wide_kernel<W, kSmem><<<grid, kSThreads, smem, stream>>>(
            packed, readable, n, static_cast<const uint8_t*>(index), K, out, count_matches, plan, dev.next_tickets());
        return cudaGetLastError();
    }
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_linear_kernel<W, kSmem>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    shared_scan_linear_kernel<W, kSmem><<<grid, kSThreads, smem, stream>>>(
        packed, readable, n, static_cast<const uint8_t*>(index), K, out, counts, count_matches, plan, dev.next_tickets());
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K, uint8_t* out,
                            unsigned long long* counts, bool count_matches, cudaStream_t stream)
{
    // the wide kernel reads mask rows lane-contiguously (conflict-free already): no mask replicas.  It
    // also takes over from kLinearWideFallbackK predicates on when the thread-per-block kernel's merged
    // row table does not fit (K = 512 at W >= 30: 0.46 against 0.21 of the HBM peak).
    bool wide = K > kLinearWideMinK;
    IndexSmemPlan plan = plan_index_smem(K, W, LinearCfg<W>::kRingBytes, /*use_hist=*/K > kDirectCountMaxK,
                                         /*merged_ok=*/!wide);
    if (!wide && !plan.merged && K >= kLinearWideFallbackK) {
        wide = true;
        plan = plan_index_smem(K, W, LinearCfg<W>::kRingBytes, /*use_hist=*/true, /*merged_ok=*/false);
    }
    if (plan.masks_in_smem)
        return launch_ws<W, true>(dev, packed, n, index, K, out, counts, count_matches, plan, wide, stream);
    return launch_ws<W, false>(dev, packed, n, index, K, out, counts, count_matches, plan, wide, stream);
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const void*, int, uint8_t*,
                                 unsigned long long*, bool, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_shared_scan_linear(const Device& dev, const void* packed, uint64_t n, int w, const void* index,
                                      int K, uint8_t* out, unsigned long long* counts, bool count_matches,
                                      cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, index, K, out, counts, count_matches, stream);
}

}  // namespace sss
