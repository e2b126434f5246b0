// predindex.cuh -- the predicate index of a shared scan: K range predicates compiled, once per call,
// into a structure that answers "which of the K predicates does this code satisfy?" with two
// shared-memory loads instead of K compares.
//
// The reference evaluates every predicate against every decoded code (the inner `for key_id` loop of
// shared_scan_128_standard, src/simd_scan_shared.cpp:73-83): K compares + K movemasks per 8 codes.
// On the GPU that is integer-issue bound long before it is HBM bound (K = 64: ~130 ops per code).
// The structure of the problem is much smaller than K x n: the 2K bounds cut the code domain into at
// most 2K+1 elementary intervals, and every code of one interval satisfies the same set of
// predicates.  So:
//
//   breakpoints  bp[0..M)   sorted, distinct: every lo_k (> 0) and hi_k + 1 (<= 2^W - 1)
//   interval id  id(v) = #{ i : bp[i] <= v }                      in [0, M]
//   mask rows    mask[id][k] = 1  iff  lo_k <= start(id) <= hi_k   ceil(K/32) words per row
//   locator      W <= 13 : lut[v] = id(v)                        (direct; 16-bit entries in global memory)
//                W  > 13 : 8192 buckets on the top 13 bits of the code, ONE 32-bit word each:
//                          [threshold : 19][several breakpoints inside : 1][id at bucket start : 12];
//                          a bucket with at most one inner breakpoint resolves with one compare of the
//                          code's remaining bits against the threshold, the rare others re-search
//
// Built on the device by one CTA (bitonic sort of <= 2048 bounds, compaction, row/bucket fill) on
// the caller's stream -- no host round trip.  The scan kernels copy it to shared memory at start.
#pragma once

#include "common.cuh"

namespace sss {

constexpr int kMaxK = 1024;            // == SSS_MAX_PREDICATES
constexpr int kMaxBp = 2 * kMaxK;      // breakpoints
constexpr int kDirectMaxW = 13;        // direct locator up to 2^13 code values
constexpr int kBucketBits = 13;
constexpr int kBuckets = 1 << kBucketBits;
constexpr int kSharedSmemBudget = 224 * 1024;  // per CTA, ONE 512-thread CTA per SM: the mask table lives in shared
                                               // memory whenever ring + counters + index fit in this

struct PredIndexHeader {
    uint32_t num_bp;      // M
    uint32_t row_words;   // row stride in words: ceil(K/32) made odd
    uint32_t direct;      // 1: lut16 locator, 0: bucket locator
    uint32_t K;
    uint32_t w;
    uint32_t reserved;
    uint32_t pad[2];
};

// Match counts.  Popcounting and warp-reducing every result word costs 3 instructions per predicate
// per 32 codes -- at K = 1024 more than the scan itself.  Instead the kernels count how many codes
// fall into each elementary INTERVAL (one shared-memory atomic per code, independent of K) and a
// tiny epilogue kernel turns the interval histogram into per-predicate counts:
//     count_k = sum over intervals i of hist[i] * mask[i][k].
//
// Mask rows are stored with an ODD word stride: the 32 lanes of a warp read the same column of 32
// different rows at once, and with a power-of-two stride (K = 1024: 32 words) every lane would hit
// the same shared-memory bank.
inline uint32_t pred_row_stride_words(int K) { return (((uint32_t)K + 31) / 32) | 1u; }

// global-memory layout of the index (all offsets 16-byte aligned)
struct PredIndexLayout {
    static constexpr size_t kHeader = 0;                                   // PredIndexHeader (32 B)
    static constexpr size_t kThr = 32;                                     // uint32 thr[kMaxBp]: (bp << (32-W)) - 1
    static constexpr size_t kLocator = kThr + (size_t)kMaxBp * 4 + 16;          // lut16[2^W] or uint32 bucket[4096]
    static constexpr size_t kLocatorBytes = (size_t)kBuckets * 4;          // 16 KB == 2^13 * 2
    static constexpr size_t kMasks = kLocator + kLocatorBytes;             // uint32 mask[(M+1)][row_words]
    static constexpr size_t kMasksBytes = ((size_t)(kMaxBp + 1) * (kMaxK / 32 + 1) * 4 + 255) / 256 * 256;
    static constexpr size_t kHist = kMasks + kMasksBytes + 256;            // unsigned long long hist[kMaxBp + 1]
    static constexpr size_t kTotal = kHist + (size_t)(kMaxBp + 1) * 8 + 256;
};

cudaError_t launch_build_pred_index(const uint32_t* d_lo, const uint32_t* d_hi, int K, int w, void* d_index,
                                    cudaStream_t stream);
// counts[k] += sum_i hist[i] * mask[i][k], then clears the histogram
cudaError_t launch_counts_from_histogram(void* d_index, int K, unsigned long long* d_counts, cudaStream_t stream);

// ---- device-side view used by the scan kernels ----------------------------------------------------

struct PredIndexView {
    const uint32_t* thr;      // shared
    const void* locator;      // shared (replica 0; plain indexing is valid only without replicas)
    uint32_t loc_saddr;       // shared-window address of this lane's replica of locator entry 0
    uint32_t loc_stride;      // bytes between consecutive locator entries (entry size x replicas)
    const uint32_t* masks;    // global table (kSmem = false)
    uint32_t rows_saddr;      // shared-window address of this lane's replica of row 0 (kSmem = true)
    uint32_t row_bytes;       // bytes between the rows of consecutive intervals
    uint32_t col_bytes;       // bytes between consecutive 32-predicate columns of one row
    uint32_t mask0_bytes;     // offset of mask column 0 inside a row (merged rows with a counter: col_bytes)
    uint32_t num_bp;
    bool direct;
};

// ---- shared-memory plan of an index-driven launch ----------------------------------------------------
//
// Per CTA, after the tile ring:  [separate histogram][locator][thresholds][row table].
//
// The lookups of a warp go to 32 unrelated addresses, so on average ~3 of them fall into the same
// bank (ncu, K = 64: 2.9 wavefronts per LDS, 2.8 per ATOMS).  Shared memory that is left over is
// therefore spent on LANE-PRIVATE REPLICAS of the small per-interval tables: with R = 2^r copies
// interleaved word by word, lane l uses copy l mod R; at R = 32 every lane stays in its own bank
// and every access is a single wavefront.
//
//   locator (merged plans): entry e of replica r at (e * R_l + r) entries; sized after the rows.
//   merged rows (thread-per-block kernels, table fits):
//       row[id] = [match counter][mask column 0][mask column 1]..., each R words; one multiply-add
//       turns an interval id into the address of its row = of its counter, the mask columns are at
//       warp-uniform offsets from it.  Row stride in R-word groups is made odd while R < 32.
//   separate (wide linear kernel; tables too large): histogram hist[id][R_h] in front, masks as the
//       flat global layout (row stride = ceil(K/32) made odd), in shared memory if they fit.
struct IndexSmemPlan {
    uint32_t merged;          // 1: merged rows, 0: separate histogram
    uint32_t rep_log2;        // merged: replicas of the row table
    uint32_t row_groups;      // merged: row stride in R-word groups
    uint32_t has_counter;     // merged: group 0 of every row is the interval's match counter
    uint32_t masks_in_smem;
    uint32_t loc_rep_log2;    // replicas of the locator (merged plans only; from what the rows leave over)
    uint32_t hist_rep_log2;   // separate: replicas of the histogram
    uint32_t hist_bytes;      // separate: bytes of the histogram in front of the index (0: none)
    uint32_t total_bytes;     // everything after the ring
};

inline IndexSmemPlan plan_index_smem(int K, int w, size_t ring_bytes, bool use_hist, bool merged_ok)
{
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    IndexSmemPlan p{};
    const size_t cols = ((size_t)K + 31) / 32;
    const size_t max_rows = (w >= 31) ? (size_t)(2 * K + 1) : std::min<size_t>((size_t)2 * K + 1, (size_t)1 << w);
    // (in shared memory the direct table is widened to 32-bit entries: one entry per bank and lane)
    const size_t locator = (w <= kDirectMaxW) ? ((size_t)4 << w) : (size_t)kBuckets * 4;
    const size_t thr = (w <= kDirectMaxW) ? 0 : (std::min<size_t>(2 * (size_t)K, kMaxBp) + 1) * 4;
    const size_t index = up16(locator) + up16(thr) + 64;
    const size_t budget = (size_t)kSharedSmemBudget;
    if (merged_ok) {
        const size_t groups0 = cols + (use_hist ? 1 : 0);
        for (int r = 5; r >= 0; r--) {
            const size_t groups = (r == 5) ? groups0 : (groups0 | 1);
            const size_t bytes = up16((max_rows * groups * 4) << r);
            if (ring_bytes + index + bytes <= budget) {
                p.merged = 1;
                p.masks_in_smem = 1;
                p.rep_log2 = (uint32_t)r;
                p.row_groups = (uint32_t)groups;
                p.has_counter = use_hist ? 1u : 0u;
                while (p.loc_rep_log2 < 5 &&
                       ring_bytes + up16(locator << (p.loc_rep_log2 + 1)) + up16(thr) + 64 + bytes <= budget)
                    p.loc_rep_log2++;
                p.total_bytes = (uint32_t)(up16(locator << p.loc_rep_log2) + up16(thr) + 64 + bytes);
                return p;
            }
        }
    }
    auto hist_bytes = [&](uint32_t r) { return use_hist ? up16((max_rows * 4) << r) : (size_t)0; };
    const size_t flat = up16(max_rows * (cols | 1) * 4);
    // (a kernel that asked for merged rows and did not get them reads its masks from global memory)
    p.masks_in_smem = (!merged_ok && ring_bytes + index + hist_bytes(0) + flat <= budget) ? 1u : 0u;
    const size_t masks = p.masks_in_smem ? flat : 0;
    // masks read from global memory go through L1, which shares the SM's 256 KB with shared memory:
    // keep the footprint small then
    const size_t hist_cap = p.masks_in_smem ? 32 * 1024 : 8 * 1024 + 64;
    while (use_hist && p.hist_rep_log2 < 5 && hist_bytes(p.hist_rep_log2 + 1) <= hist_cap &&
           ring_bytes + index + hist_bytes(p.hist_rep_log2 + 1) + masks <= budget)
        p.hist_rep_log2++;
    p.hist_bytes = (uint32_t)hist_bytes(p.hist_rep_log2);
    p.total_bytes = (uint32_t)(p.hist_bytes + index + masks);
    return p;
}

// All threads: copy the index into shared memory at `smem` (16-byte aligned; behind the separate
// histogram if there is one), expanding the mask table into the plan's layout and zeroing the
// merged counters.  Visible after the next __syncthreads.
__device__ __forceinline__ PredIndexView load_pred_index(const uint8_t* __restrict__ g_index, uint8_t* smem,
                                                         const IndexSmemPlan& plan)
{
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
    PredIndexView v;
    v.num_bp = h->num_bp;
    v.direct = h->direct != 0;
    const uint32_t g_row_words = h->row_words;
    const uint32_t lr = plan.loc_rep_log2, LR = 1u << lr;
    const uint32_t locator_bytes = (v.direct ? (4u << h->w) : (uint32_t)kBuckets * 4) << lr;
    const uint32_t thr_bytes = v.direct ? 0u : (v.num_bp + 1) * 4;  // + the 0xFFFFFFFF sentinel
    uint8_t* s_loc = smem;
    uint8_t* s_thr = s_loc + ((locator_bytes + 15) & ~15u);
    uint8_t* s_rows = s_thr + ((thr_bytes + 15) & ~15u);
    auto copy = [&](uint8_t* dst, const uint8_t* src, uint32_t bytes) {
        const uint32_t n16 = (bytes + 15) / 16;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x)
            reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    };
    if (v.direct) {  // 16-bit entries in global memory, widened here; LR replicas each
        const uint16_t* src = reinterpret_cast<const uint16_t*>(g_index + PredIndexLayout::kLocator);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_loc);
        for (uint32_t i = threadIdx.x; i < ((1u << h->w) << lr); i += blockDim.x) dst[i] = src[i >> lr];
    } else if (lr == 0) {
        copy(s_loc, g_index + PredIndexLayout::kLocator, locator_bytes);
    } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kLocator);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_loc);
        for (uint32_t i = threadIdx.x; i < ((uint32_t)kBuckets << lr); i += blockDim.x) dst[i] = __ldg(src + (i >> lr));
    }
    copy(s_thr, g_index + PredIndexLayout::kThr, thr_bytes);
    v.locator = s_loc;
    v.loc_saddr = smem_u32(s_loc) + (threadIdx.x & (LR - 1)) * 4u;
    v.loc_stride = 4u << lr;
    v.thr = reinterpret_cast<const uint32_t*>(s_thr);
    const uint32_t* g_masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
    v.masks = g_masks;
    v.mask0_bytes = 0;
    if (plan.merged) {
        // thread -> (row, group, replica); consecutive threads write consecutive words
        const uint32_t r = plan.rep_log2, R = 1u << r;
        const uint32_t cols = (h->K + 31) / 32, groups = plan.row_groups, first = plan.has_counter;
        const uint32_t total = ((v.num_bp + 1) * groups) << r;
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_rows);
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const uint32_t cell = i >> r, row = cell / groups, c = cell - row * groups - first;  // c: mask column
            dst[i] = c < cols ? __ldg(g_masks + row * g_row_words + c) : 0u;  // (counter and padding: zero)
        }
        v.rows_saddr = smem_u32(s_rows) + ((threadIdx.x & (R - 1)) << 2);
        v.row_bytes = (groups << r) * 4;
        v.col_bytes = 4u << r;
        v.mask0_bytes = plan.has_counter * v.col_bytes;
    } else if (plan.masks_in_smem) {
        copy(s_rows, g_index + PredIndexLayout::kMasks, (v.num_bp + 1) * g_row_words * 4);
        v.rows_saddr = smem_u32(s_rows);
        v.row_bytes = g_row_words * 4;
        v.col_bytes = 4;
    } else {
        v.rows_saddr = 0;
        v.row_bytes = g_row_words * 4;
        v.col_bytes = 4;
    }
    return v;
}

// merged rows: add the counters of this CTA (summed over the replicas) to the index's global histogram.
// All threads, after the __syncthreads that follows the last count.
__device__ __forceinline__ void merged_hist_flush(const PredIndexView& ix, const IndexSmemPlan& plan, uint8_t* s_index,
                                                  const uint8_t* g_index)
{
    const uint32_t R = 1u << plan.rep_log2;
    // recompute the table base the way load_pred_index laid it out
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
    const uint32_t locator_bytes = (ix.direct ? (4u << h->w) : (uint32_t)kBuckets * 4) << plan.loc_rep_log2;
    const uint32_t thr_bytes = ix.direct ? 0u : (ix.num_bp + 1) * 4;
    const uint32_t* rows =
        reinterpret_cast<const uint32_t*>(s_index + ((locator_bytes + 15) & ~15u) + ((thr_bytes + 15) & ~15u));
    unsigned long long* g_hist =
        reinterpret_cast<unsigned long long*>(const_cast<uint8_t*>(g_index) + PredIndexLayout::kHist);
    for (uint32_t i = threadIdx.x; i <= ix.num_bp; i += blockDim.x) {
        const uint32_t* c = rows + (((size_t)i * plan.row_groups) << plan.rep_log2);
        uint32_t s = 0;
        for (uint32_t r = 0; r < R; r++) s += c[(r + i) & (R - 1)];  // rotated start: no bank conflicts
        if (s) atomicAdd(&g_hist[i], (unsigned long long)s);
    }
}

// The separate interval histogram of a CTA: `rep_log2` lane-private replicas per interval.
struct IntervalHist {
    uint32_t* mine;   // this lane's replica of interval 0
    uint32_t shift;   // log2(replicas)
    __device__ __forceinline__ void add(uint32_t id) const { atomicAdd(mine + (id << shift), 1u); }
};

// All threads, before the __syncthreads that precedes the first add()
__device__ __forceinline__ IntervalHist hist_init(uint8_t* smem, const IndexSmemPlan& plan)
{
    uint32_t* h = reinterpret_cast<uint32_t*>(smem);
    const uint32_t words = plan.hist_bytes / 4;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) h[i] = 0;
    const uint32_t R = 1u << plan.hist_rep_log2;
    return IntervalHist{h + (threadIdx.x & (R - 1)), plan.hist_rep_log2};
}

// All threads, after the __syncthreads that follows the last add(): fold the replicas and add the
// CTA's interval counts to the global histogram of the index
__device__ __forceinline__ void hist_flush(const uint8_t* smem, const IndexSmemPlan& plan, uint32_t num_bp,
                                           const uint8_t* g_index)
{
    const uint32_t* h = reinterpret_cast<const uint32_t*>(smem);
    const uint32_t R = 1u << plan.hist_rep_log2;
    unsigned long long* g_hist =
        reinterpret_cast<unsigned long long*>(const_cast<uint8_t*>(g_index) + PredIndexLayout::kHist);
    for (uint32_t i = threadIdx.x; i <= num_bp; i += blockDim.x) {
        uint32_t s = 0;
        for (uint32_t r = 0; r < R; r++) s += h[(i << plan.hist_rep_log2) + ((r + i) & (R - 1))];
        if (s) atomicAdd(&g_hist[i], (unsigned long long)s);
    }
}

// ---- one 32-code block: shared memory -> registers ------------------------------------------------
//
// A thread owns a whole block (W consecutive words of the tile), so the 32 lanes of a warp read with
// a stride of W words.  That is conflict-free only for odd W.  Widest conflict-free access per width:
//   odd W          32-bit loads, stride W (odd)
//   W = 2 mod 4    64-bit loads, stride W/2 (odd) -- a half-warp per wavefront
//   W = 4 mod 8    128-bit loads, stride W/4 (odd) -- a quarter-warp per wavefront
//   W = 0 mod 8    128-bit loads with a lane-dependent ROTATION: the block is cut into kPieces equal
//                  pieces of whole 16-byte chunks that hold whole codes (8/16/32 bit: one chunk,
//                  24 bit: three chunks = 16 codes) and lane l starts at piece rot(l), so that the
//                  eight lanes of a quarter-warp touch eight different bank groups.  Register j then
//                  holds code (j + rot * kPieceCodes) mod 32 of the block; callers undo that on the
//                  RESULT (one funnel shift per result word, or a permutation of the output rows).
// kGranule: the rotation has to be a multiple of this many codes (1: bitvector words, 8: the linear
// layout's rows of 8 codes; 32-bit codes then rotate by two chunks and keep a 2-way conflict).
template <int W, int kGranule>
struct BlockRot {
    static constexpr int kChunks = W / 4;
    static constexpr int kBaseChunks = (W == 24) ? 3 : 1;
    static constexpr int kBaseCodes = (W % 8 == 0) ? kBaseChunks * 128 / W : 32;
    static constexpr int kMul = kBaseCodes >= kGranule ? 1 : kGranule / kBaseCodes;
    static constexpr int kPieceChunks = kBaseChunks * kMul;
    static constexpr int kPieces = (W % 8 == 0) ? kChunks / kPieceChunks : 1;
    static constexpr int kPieceCodes = 32 / kPieces;
    static constexpr bool kRot = kPieces > 1;
    // lanes l and l + 8/g start in the same bank group (g = gcd(kChunks, 8)): spread them over the pieces
    static constexpr int kGcd = (kChunks % 8 == 0) ? 8 : (kChunks % 4 == 0 ? 4 : (kChunks % 2 == 0 ? 2 : 1));
    static __device__ __forceinline__ uint32_t lane_rot()
    {
        if constexpr (!kRot) return 0u;
        return (((threadIdx.x & 31u) * kGcd) >> 3) & (uint32_t)(kPieces - 1);
    }
};

template <int W, int kGranule>
__device__ __forceinline__ void load_block(const uint32_t* buf, uint32_t blk, uint32_t (&x)[W], uint32_t rot)
{
    using R = BlockRot<W, kGranule>;
    if constexpr (W % 4 == 0) {
        const uint4* src = reinterpret_cast<const uint4*>(buf) + (size_t)blk * R::kChunks;
#pragma unroll
        for (int q = 0; q < R::kChunks; q++) {
            uint32_t c = q;
            if constexpr (R::kRot) {
                c = q + rot * R::kPieceChunks;
                c = c >= (uint32_t)R::kChunks ? c - R::kChunks : c;
            }
            const uint4 v = src[c];
            x[4 * q + 0] = v.x;
            x[4 * q + 1] = v.y;
            x[4 * q + 2] = v.z;
            x[4 * q + 3] = v.w;
        }
    } else if constexpr (W % 2 == 0) {
        const uint2* src = reinterpret_cast<const uint2*>(buf) + (size_t)blk * (W / 2);
#pragma unroll
        for (int q = 0; q < W / 2; q++) {
            const uint2 v = src[q];
            x[2 * q + 0] = v.x;
            x[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) x[k] = buf[(size_t)blk * W + k];
    }
}

// ---- interval ids of the 32 codes of a block ---------------------------------------------------
//
// id(code) = number of breakpoints <= code, from the LEFT-aligned code t (garbage allowed in the
// low 32-W bits).  Branch-free for all 32 codes so that the 32 locator loads overlap.
//
// Bucket word e = thr19 << 13 | several << 12 | id_at_bucket_start.  The bucket fixes the top 13 bits
// of t, so only the remaining 19 decide:  t >= breakpoint  <=>  (t << 13) > e  with
// thr19 = (left-aligned breakpoint & 0x7FFFF) - 1 (the low 13 bits of t << 13 are zero, so whatever
// sits below thr19 in e cannot turn a tie into "greater"), and thr19 = 0x7FFFF (never exceeded) for a
// bucket without an inner breakpoint.  Buckets with two or more inner breakpoints (K = 64: ~0.01 % of
// the codes of a uniform column; every equality predicate, whose breakpoints key and key+1 are
// adjacent) set `several`; a block that met one patches the ids of exactly those codes (fix()).
constexpr int kThrBits = 32 - kBucketBits;           // 19
constexpr uint32_t kThrMask = (1u << kThrBits) - 1;  // 0x7FFFF
constexpr uint32_t kBucketIdMask = (1u << (kBucketBits - 1)) - 1;  // 12 bits: ids 0..2048
constexpr uint32_t kSeveralBit = 1u << (kBucketBits - 1);
static_assert(kMaxBp <= (int)kBucketIdMask, "interval ids must fit below the several flag");

SSS_HD inline uint32_t make_bucket_word(uint32_t id_at_start, bool several, uint32_t thr19)
{
    return (thr19 << kBucketBits) | (several ? kSeveralBit : 0u) | id_at_start;
}

// entry `idx` of this lane's replica of the locator: one multiply-add for the address
__device__ __forceinline__ uint32_t locator_entry(const PredIndexView& ix, uint32_t idx)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(idx * ix.loc_stride + ix.loc_saddr));
    return v;
}

// Left-aligned code with the shift of a non-straddling code issued as a MULTIPLY by a power of two that
// ptxas cannot see (constant bank): written as `x << a`, ptxas merges it with the `>> 19` of the bucket
// index into `(x >> b) & 0x1fff` -- two ALU-pipe instructions (SHF + LOP3) where IMAD (FMA pipe) + SHF do,
// and the ALU pipe is the one these kernels are short of (ncu: ALU 67 %, FMA 16 % at K = 64).
#ifndef SSS_LOCATE_FMA
#define SSS_LOCATE_FMA 1
#endif
static __constant__ uint32_t c_pow2[32] = {
    1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,  1u << 8,  1u << 9,  1u << 10,
    1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21,
    1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

template <int W, int J>
__device__ __forceinline__ uint32_t located_code(const uint32_t (&x)[W])
{
#if SSS_LOCATE_FMA
    constexpr CodeSlot s = code_slot(W, J);
    if constexpr (s.end_shift != 32 && !s.straddles) return x[s.last_word] * c_pow2[32 - s.end_shift];
#endif
    return left_aligned_code<W, J>(x);
}

template <int W, int J>
struct BlockIds {
    static __device__ __forceinline__ void fast(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&id)[32],
                                                uint32_t& flags)
    {
        const uint32_t t = left_aligned_code<W, J>(x);
        if constexpr (W <= kDirectMaxW) {
            id[J] = locator_entry(ix, t >> (32 - W));
        } else {
            const uint32_t e = locator_entry(ix, t >> (32 - kBucketBits));
            id[J] = (e & kBucketIdMask) + ((t << kBucketBits) > e ? 1u : 0u);
            flags |= e;
        }
        BlockIds<W, J + 1>::fast(x, ix, id, flags);
    }
    // executed only by blocks that met a `several` bucket: the codes that fell into one walk the
    // (sorted) thresholds from the bucket's first breakpoint; thr[num_bp] is a 0xFFFFFFFF sentinel
    static __device__ __forceinline__ void fix(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&id)[32])
    {
        if constexpr (W > kDirectMaxW) {
            const uint32_t t = left_aligned_code<W, J>(x);
            const uint32_t e = locator_entry(ix, t >> (32 - kBucketBits));
            if (e & kSeveralBit) {
                uint32_t i = e & kBucketIdMask;
                while (t > ix.thr[i]) i++;
                id[J] = i;
            }
        }
        BlockIds<W, J + 1>::fix(x, ix, id);
    }
};
template <int W>
struct BlockIds<W, 32> {
    static __device__ __forceinline__ void fast(const uint32_t (&)[W], const PredIndexView&, uint32_t (&)[32], uint32_t&) {}
    static __device__ __forceinline__ void fix(const uint32_t (&)[W], const PredIndexView&, uint32_t (&)[32]) {}
};

template <int W>
__device__ __forceinline__ void block_ids(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&id)[32])
{
    uint32_t flags = 0;
    BlockIds<W, 0>::fast(x, ix, id, flags);
    if (flags & kSeveralBit) BlockIds<W, 0>::fix(x, ix, id);
}

// Row handles straight from the locator (merged rows in shared memory): the id never materialises --
// row = (bucket's id at start) * row_bytes + base, plus one row when the code passed the threshold
// (a multiply-add and a predicated add instead of mask, add, select, multiply-add).
template <int W, int J>
struct BlockRows {
    static __device__ __forceinline__ void fast(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&row)[32],
                                                uint32_t& flags)
    {
        const uint32_t t = located_code<W, J>(x);
        if constexpr (W <= kDirectMaxW) {
            row[J] = locator_entry(ix, t >> (32 - W)) * ix.row_bytes + ix.rows_saddr;
        } else {
            const uint32_t e = locator_entry(ix, t >> (32 - kBucketBits));
            uint32_t r = (e & kBucketIdMask) * ix.row_bytes + ix.rows_saddr;
            // one predicated add (ptxas otherwise selects 0 / row_bytes first and adds after)
            asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}"
                : "+r"(r)
                : "r"(t << kBucketBits), "r"(e), "r"(ix.row_bytes));
            row[J] = r;
            flags |= e;
        }
        BlockRows<W, J + 1>::fast(x, ix, row, flags);
    }
    static __device__ __forceinline__ void fix(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&row)[32])
    {
        if constexpr (W > kDirectMaxW) {
            const uint32_t t = left_aligned_code<W, J>(x);
            const uint32_t e = locator_entry(ix, t >> (32 - kBucketBits));
            if (e & kSeveralBit) {
                uint32_t i = e & kBucketIdMask;
                while (t > ix.thr[i]) i++;
                row[J] = i * ix.row_bytes + ix.rows_saddr;
            }
        }
        BlockRows<W, J + 1>::fix(x, ix, row);
    }
};
template <int W>
struct BlockRows<W, 32> {
    static __device__ __forceinline__ void fast(const uint32_t (&)[W], const PredIndexView&, uint32_t (&)[32], uint32_t&) {}
    static __device__ __forceinline__ void fix(const uint32_t (&)[W], const PredIndexView&, uint32_t (&)[32]) {}
};

template <int W>
__device__ __forceinline__ void block_rows(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&row)[32])
{
    uint32_t flags = 0;
    BlockRows<W, 0>::fast(x, ix, row, flags);
    if (flags & kSeveralBit) BlockRows<W, 0>::fix(x, ix, row);
}

// interval ids -> row handles: the shared-window address of the lane's replica of the row (kSmem),
// or the row's byte offset in the global table.  One multiply-add per code; the mask columns and the
// match counter are then at warp-uniform offsets.
template <bool kSmem>
__device__ __forceinline__ void rows_of_ids(const PredIndexView& ix, uint32_t (&id)[32])
{
#pragma unroll
    for (int j = 0; j < 32; j++) id[j] = id[j] * ix.row_bytes + (kSmem ? ix.rows_saddr : 0u);
}

// mask word of a row handle at column byte offset `col_bytes` (warp-uniform)
template <bool kSmem>
__device__ __forceinline__ uint32_t row_mask_word(const PredIndexView& ix, uint32_t row, uint32_t col_bytes)
{
    if constexpr (kSmem) {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(row + col_bytes));
        return v;
    } else {
        return __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ix.masks) + row + col_bytes));
    }
}

// merged rows: count one code of this row's interval (lane-private counter: no conflicts at R = 32)
__device__ __forceinline__ void row_count(uint32_t row)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(row) : "memory");
}

// ---- 32 x 32 bit-matrix transpose in registers ------------------------------------------------------

// (a & m) | (b & ~m) in ONE LOP3 (ptxas does not fuse the two-immediate form on its own)
__device__ __forceinline__ uint32_t bit_select(uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// Pipe balance (measured, not adopted here).  Integer instructions issue on two pipes: LOP3, PRMT,
// shifts and compares on the ALU pipe, IMAD on the FMA pipe.  An exchange stage is 3 ALU + 1 FMA
// instructions per register pair (ptxas already turns `b << S` into IMAD.SHL).  Issuing `a >> S` as
// umulhi(a, 2^(32-S)) = IMAD.HI (the power of two loaded from constant memory so that ptxas cannot fold
// it back into a shift) makes it 2 + 2 -- and was SLOWER in every kernel that transposes (linear K = 64:
// 0.85 -> 0.80 of the HBM peak, linear K = 1024: 0.66 -> 0.62, shared K = 64 unchanged): IMAD.HI does not
// run at the rate of a shift.  The K <= 8 lookup in narrow.cu, which was at 86 % ALU pipe utilisation,
// is the one place where the trade pays.  -DSSS_SHR_FMA=1 rebuilds the variant.
#ifndef SSS_SHR_FMA
#define SSS_SHR_FMA 0
#endif
static __constant__ uint32_t c_shr_mul[5] = {0u, 1u << 31, 1u << 30, 0u, 1u << 28};  // [S] = 2^(32 - S), S = 1, 2, 4

// base + off as ONE IMAD.WIDE on the FMA pipe.  ptxas forms a 64-bit address with two instructions
// (IADD3 + carry add; `p + k * stride` with a constant k even with three), most of them on the ALU pipe
// these kernels are short of; the multiplier 1 is loaded from constant memory so that it cannot be
// folded back into the add pair.
static __constant__ uint32_t c_one = 1u;
template <class T>
__device__ __forceinline__ T* ptr_add_fma(T* base, uint32_t off_bytes)
{
    uint64_t a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(off_bytes), "r"(c_one), "l"(base));
    return reinterpret_cast<T*>(a);
}

// base + k * stride as ONE IMAD.WIDE (FMA pipe), k = 0..31 read from the constant bank so that ptxas
// can neither fold it nor strength-reduce the 32 addresses of a column into a 64-bit add chain
static __constant__ uint32_t c_iota[32] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15,
                                           16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31};
template <class T>
__device__ __forceinline__ T* ptr_mad_fma(T* base, uint32_t stride_bytes, int k)
{
    uint64_t a;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(stride_bytes), "r"(c_iota[k]), "l"(base));
    return reinterpret_cast<T*>(a);
}

// one exchange stage on a register pair: swaps the high S-bit groups of `lo` with the low ones of `hi`
template <int S>
__device__ __forceinline__ void exchange_bits(uint32_t& lo, uint32_t& hi)
{
    constexpr uint32_t M = S == 4 ? 0x0F0F0F0Fu : (S == 2 ? 0x33333333u : 0x55555555u);
    const uint32_t a = lo, b = hi;
    lo = bit_select(a, b << S, M);
#if SSS_SHR_FMA
    hi = bit_select(__umulhi(a, c_shr_mul[S]), b, M);
#else
    hi = bit_select(a >> S, b, M);
#endif
}

// in : a[j] bit k  = element (j, k)        (j = code index, k = predicate index)
// out: a[k] bit j  = element (j, k)
// Two byte-granular stages with PRMT, three bit-granular stages with shift + bit-select LOP3.
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
#pragma unroll
    for (int k = 0; k < 16; k++) {  // 16 x 16 blocks
        const uint32_t lo = a[k], hi = a[k + 16];
        a[k] = __byte_perm(lo, hi, 0x5410);
        a[k + 16] = __byte_perm(lo, hi, 0x7632);
    }
#pragma unroll
    for (int g = 0; g < 32; g += 16)
#pragma unroll
        for (int k = g; k < g + 8; k++) {  // 8 x 8 blocks
            const uint32_t lo = a[k], hi = a[k + 8];
            a[k] = __byte_perm(lo, hi, 0x6240);
            a[k + 8] = __byte_perm(lo, hi, 0x7351);
        }
#pragma unroll
    for (int g = 0; g < 32; g += 8)
#pragma unroll
        for (int k = g; k < g + 4; k++) exchange_bits<4>(a[k], a[k + 4]);
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) exchange_bits<2>(a[k], a[k + 2]);
#pragma unroll
    for (int k = 0; k < 32; k += 2) exchange_bits<1>(a[k], a[k + 1]);
}

// ---- up to 8 predicates: 32 x 8 transpose --------------------------------------------------------
// in : m[j] bits 0..7 = the predicates on code j.   out: r[k] bit j = predicate k on code j.
// Four codes are packed per word so that byte b of word q is code 8b + q; the three bit stages then
// leave predicate k's 32 results in word k already in code order -- 72 ops instead of 256.

// bit stages of the 32 x 8 transpose.  in: r[q] byte b = the 8 predicate bits of code 8b + q;
// out: r[k] bit j = predicate k on code j
__device__ __forceinline__ void transpose_bits_32x8(uint32_t (&r)[8])
{
#pragma unroll
    for (int k = 0; k < 4; k++) exchange_bits<4>(r[k], r[k + 4]);
#pragma unroll
    for (int g = 0; g < 8; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) exchange_bits<2>(r[k], r[k + 2]);
#pragma unroll
    for (int k = 0; k < 8; k += 2) exchange_bits<1>(r[k], r[k + 1]);
}

__device__ __forceinline__ void transpose32x8(const uint32_t (&m)[32], uint32_t (&r)[8])
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t lo = __byte_perm(m[q], m[q + 8], 0x0040);        // m[q].b0, m[q+8].b0
        const uint32_t hi = __byte_perm(m[q + 16], m[q + 24], 0x0040);  // m[q+16].b0, m[q+24].b0
        r[q] = __byte_perm(lo, hi, 0x5410);
    }
    transpose_bits_32x8(r);
}

// ---- bitvector layout, up to 8 predicates: 128-bit stores through a warp-private stage --------------------
// r[k] = the block's result word of predicate k.  Stored per thread that is one 32-bit word per
// predicate (128 bytes per warp instruction); the SM moves far more bytes of 128-bit stores next to
// the kernel's other traffic (shared.cu, "Wide stores").  The warp parks its 8 x 32 words in `stage`
// (1 KB, conflict-free) and lane (k mod 4, q) writes words 4q..4q+3 of predicate k with one 128-bit
// store: 2 store instructions per thread instead of 8.  All 32 lanes of the warp call this together;
// `warp0` = bitvector 0's word of the warp's FIRST block; base and stride 16-byte aligned.
constexpr int kSmallStageBytesPerWarp = 8 * 32 * 4;
__device__ __forceinline__ void store_words8_wide(uint32_t* stage, const uint32_t (&r)[8], int K, uint32_t* warp0,
                                                  uint64_t stride_words)
{
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();  // the previous block's wide reads of the stage are done
#pragma unroll
    for (int k = 0; k < 8; k++) stage[k * 32 + lane] = r[k];
    __syncwarp();
    const uint32_t kq = lane >> 3, q4 = (lane & 7) * 4;
    uint32_t* pw = warp0 + (uint64_t)kq * stride_words + q4;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if ((int)kq + 4 * i < K)
            __stcs(reinterpret_cast<uint4*>(pw), *reinterpret_cast<const uint4*>(stage + (kq + 4 * i) * 32 + q4));
        pw += 4 * stride_words;
    }
}

// ---- linear layout, up to 8 predicates ----------------------------------------------------------------
// 4 x 4 byte transpose: in a,b,c,d = four predicates' result words (byte g = codes 8g..8g+7);
// out o[g] = bytes (a.g, b.g, c.g, d.g) = four consecutive predicate bytes of row g
__device__ __forceinline__ void byte_transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t (&o)[4])
{
    const uint32_t t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
    const uint32_t t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
    o[0] = __byte_perm(t0, t1, 0x5410);
    o[1] = __byte_perm(t0, t1, 0x7632);
    o[2] = __byte_perm(t2, t3, 0x5410);
    o[3] = __byte_perm(t2, t3, 0x7632);
}

// The 4 rows x K bytes of a block for K <= 8, from the per-predicate result words r[k] (bit j = code j):
// the rows of a block are contiguous (4K bytes), so K = 1, 2, 4, 8 are ONE 4/8/16/32-byte store.
__device__ __forceinline__ void store_block_small(uint8_t* dst, const uint32_t (&r)[8], int K)
{
    uint32_t lo[4], hi[4];
    byte_transpose4(r[0], r[1], r[2], r[3], lo);
    if (K == 1 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        __stcs(reinterpret_cast<uint32_t*>(dst), r[0]);
    } else if (K == 2 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
        __stcs(reinterpret_cast<uint2*>(dst),
               make_uint2(__byte_perm(lo[0], lo[1], 0x5410), __byte_perm(lo[2], lo[3], 0x5410)));
    } else if (K == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        __stcs(reinterpret_cast<uint4*>(dst), make_uint4(lo[0], lo[1], lo[2], lo[3]));
    } else {
        byte_transpose4(r[4], r[5], r[6], r[7], hi);
        if (K == 8 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
            const uint32_t ws[8] = {lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], lo[3], hi[3]};
            st_v8_u32(dst, ws);
        } else {
#pragma unroll
            for (int g = 0; g < 4; g++)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < K) dst[g * K + k] = (uint8_t)((k < 4 ? lo[g] : hi[g]) >> (8 * (k & 3)));
        }
    }
}

}  // namespace sss
