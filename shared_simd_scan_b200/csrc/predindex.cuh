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
//   locator      W <= 13 : lut[v] = id(v) * row_bytes             (direct, 2 bytes per code value)
//                W  > 13 : 8192 buckets on the top 13 bits of the code, ONE 32-bit word each:
//                          [id at bucket start : 12][several breakpoints inside : 1][threshold : 19];
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
constexpr int kSharedSmemBudget = 112 * 1024;  // per CTA, two CTAs per SM: the mask table lives in shared
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
    const void* locator;      // shared
    const uint32_t* masks;    // shared or global
    uint32_t masks_saddr;     // shared-window address of the mask table (when it lives in shared memory)
    uint32_t row_words;
    uint32_t num_bp;
    bool direct;
};

// shared-memory bytes the kernels need for the index of this launch; `other_bytes` is what the
// kernel already uses (tile ring + interval counters)
inline size_t pred_index_smem_bytes(int K, int w, size_t other_bytes, bool* masks_in_smem)
{
    const size_t stride = pred_row_stride_words(K);
    const size_t max_rows = (w >= 31) ? (size_t)(2 * K + 1) : std::min<size_t>((size_t)2 * K + 1, (size_t)1 << w);
    const size_t mask_bytes = (max_rows * stride * 4 + 15) & ~(size_t)15;
    const size_t locator = (w <= kDirectMaxW) ? ((size_t)2 << w) : (size_t)kBuckets * 4;
    const size_t thr = (w <= kDirectMaxW) ? 0 : (std::min<size_t>(2 * (size_t)K, kMaxBp) + 1) * 4;
    const size_t base = ((locator + 15) & ~(size_t)15) + ((thr + 15) & ~(size_t)15);
    *masks_in_smem = other_bytes + base + mask_bytes + 64 <= (size_t)kSharedSmemBudget;
    return base + (*masks_in_smem ? mask_bytes : 0);
}

// All threads: copy the index into shared memory at `smem` (16-byte aligned).  Visible after the
// next __syncthreads.
__device__ __forceinline__ PredIndexView load_pred_index(const uint8_t* __restrict__ g_index, uint8_t* smem,
                                                         bool masks_in_smem)
{
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
    PredIndexView v;
    v.num_bp = h->num_bp;
    v.row_words = h->row_words;
    v.direct = h->direct != 0;
    const uint32_t locator_bytes = v.direct ? (2u << h->w) : (uint32_t)kBuckets * 4;
    const uint32_t thr_bytes = v.direct ? 0u : (v.num_bp + 1) * 4;  // + the 0xFFFFFFFF sentinel
    const uint32_t mask_bytes = (v.num_bp + 1) * v.row_words * 4;
    uint8_t* s_loc = smem;
    uint8_t* s_thr = s_loc + ((locator_bytes + 15) & ~15u);
    uint8_t* s_mask = s_thr + ((thr_bytes + 15) & ~15u);
    auto copy = [&](uint8_t* dst, const uint8_t* src, uint32_t bytes) {
        const uint32_t n16 = (bytes + 15) / 16;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x)
            reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    };
    copy(s_loc, g_index + PredIndexLayout::kLocator, locator_bytes);
    copy(s_thr, g_index + PredIndexLayout::kThr, thr_bytes);
    v.locator = s_loc;
    v.thr = reinterpret_cast<const uint32_t*>(s_thr);
    if (masks_in_smem) {
        copy(s_mask, g_index + PredIndexLayout::kMasks, mask_bytes);
        v.masks = reinterpret_cast<const uint32_t*>(s_mask);
        v.masks_saddr = smem_u32(s_mask);
    } else {
        v.masks_saddr = 0;
        v.masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
    }
    return v;
}

// ---- interval ids of the 32 codes of a block ---------------------------------------------------
//
// id(code) = number of breakpoints <= code, from the LEFT-aligned code t (garbage allowed in the
// low 32-W bits).  Branch-free for all 32 codes so that the 32 locator loads overlap.
//
// Bucket word e = id_at_bucket_start << 20 | several << 19 | thr19.  The bucket fixes the top 13
// bits of t, so only the remaining 19 decide:  t >= breakpoint  <=>  (t << 13) > (e << 13)  with
// thr19 = (left-aligned breakpoint & 0x7FFFF) - 1, and thr19 = 0x7FFFF (never exceeded) for a bucket
// without an inner breakpoint.  Buckets with two or more inner breakpoints (K = 64: ~0.01 % of the
// codes of a uniform column; every equality predicate, whose breakpoints key and key+1 are
// adjacent) set `several`; a block that met one patches the ids of exactly those codes (fix()).
constexpr int kThrBits = 32 - kBucketBits;           // 19
constexpr uint32_t kThrMask = (1u << kThrBits) - 1;  // 0x7FFFF
constexpr uint32_t kSeveralBit = 1u << kThrBits;

template <int W, int J>
struct BlockIds {
    static __device__ __forceinline__ void fast(const uint32_t (&x)[W], const PredIndexView& ix, uint32_t (&id)[32],
                                                uint32_t& flags)
    {
        const uint32_t t = left_aligned_code<W, J>(x);
        if constexpr (W <= kDirectMaxW) {
            id[J] = static_cast<const uint16_t*>(ix.locator)[t >> (32 - W)];
        } else {
            const uint32_t e = static_cast<const uint32_t*>(ix.locator)[t >> (32 - kBucketBits)];
            id[J] = (e >> (kThrBits + 1)) + ((t << kBucketBits) > (e << kBucketBits) ? 1u : 0u);
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
            const uint32_t e = static_cast<const uint32_t*>(ix.locator)[t >> (32 - kBucketBits)];
            if (e & kSeveralBit) {
                uint32_t i = e >> (kThrBits + 1);
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

// word c of the mask row of interval `id`; kSmem selects shared-memory addressing (LDS) over the
// generic path used when the table is too large for shared memory (K > 128 or so)
template <bool kSmem>
__device__ __forceinline__ uint32_t mask_word(const PredIndexView& ix, uint32_t id, int c)
{
    const uint32_t word = id * ix.row_words + (uint32_t)c;
    if constexpr (kSmem) {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ix.masks_saddr + word * 4));
        return v;
    } else {
        return __ldg(ix.masks + word);
    }
}

// Same with the row already resolved to a byte offset (id * row_words * 4) and the column to bytes:
// lets a caller hoist the multiply out of its inner loop.
template <bool kSmem>
__device__ __forceinline__ uint32_t mask_word_at(const PredIndexView& ix, uint32_t row_byte_offset, uint32_t col_bytes)
{
    if constexpr (kSmem) {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ix.masks_saddr + row_byte_offset + col_bytes));
        return v;
    } else {
        return __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ix.masks) + row_byte_offset +
                                                       col_bytes));
    }
}

// ---- 32 x 32 bit-matrix transpose in registers ------------------------------------------------------

// (a & m) | (b & ~m) in ONE LOP3 (ptxas does not fuse the two-immediate form on its own)
__device__ __forceinline__ uint32_t bit_select(uint32_t a, uint32_t b, uint32_t m)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// one exchange stage on a register pair: swaps the high S-bit groups of `lo` with the low ones of `hi`
template <int S>
__device__ __forceinline__ void exchange_bits(uint32_t& lo, uint32_t& hi)
{
    constexpr uint32_t M = S == 4 ? 0x0F0F0F0Fu : (S == 2 ? 0x33333333u : 0x55555555u);
    const uint32_t a = lo, b = hi;
    lo = bit_select(a, b << S, M);
    hi = bit_select(a >> S, b, M);
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
__device__ __forceinline__ void transpose32x8(const uint32_t (&m)[32], uint32_t (&r)[8])
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint32_t lo = __byte_perm(m[q], m[q + 8], 0x0040);        // m[q].b0, m[q+8].b0
        const uint32_t hi = __byte_perm(m[q + 16], m[q + 24], 0x0040);  // m[q+16].b0, m[q+24].b0
        r[q] = __byte_perm(lo, hi, 0x5410);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) exchange_bits<4>(r[k], r[k + 4]);
#pragma unroll
    for (int g = 0; g < 8; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) exchange_bits<2>(r[k], r[k + 2]);
#pragma unroll
    for (int k = 0; k < 8; k += 2) exchange_bits<1>(r[k], r[k + 1]);
}

}  // namespace sss
