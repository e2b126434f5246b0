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
//                W  > 13 : 4096 buckets on the top 12 bits of the code: (first row offset, count of
//                          breakpoints inside, left-aligned first breakpoint - 1); a bucket with at
//                          most one breakpoint resolves with one compare, more fall to a short loop
//
// Built on the device by one CTA (bitonic sort of <= 2048 bounds, compaction, row/bucket fill) on
// the caller's stream -- no host round trip.  The scan kernels copy it to shared memory at start.
#pragma once

#include "common.cuh"

namespace sss {

constexpr int kMaxK = 1024;            // == SSS_MAX_PREDICATES
constexpr int kMaxBp = 2 * kMaxK;      // breakpoints
constexpr int kDirectMaxW = 13;        // direct locator up to 2^13 code values
constexpr int kBucketBits = 12;
constexpr int kBuckets = 1 << kBucketBits;
constexpr int kMaskSmemBytes = 16 * 1024;  // mask tables up to this size live in shared memory

struct PredIndexHeader {
    uint32_t num_bp;      // M
    uint32_t row_words;   // ceil(K/32) rounded up to a power of two
    uint32_t direct;      // 1: lut16 locator, 0: bucket locator
    uint32_t K;
    uint32_t w;
    uint32_t row_shift;   // log2(row bytes)
    uint32_t pad[2];
};

// Match counts.  Popcounting and warp-reducing every result word costs 3 instructions per predicate
// per 32 codes -- at K = 1024 more than the scan itself.  Instead the kernels count how many codes
// fall into each elementary INTERVAL (one shared-memory atomic per code, independent of K) and a
// tiny epilogue kernel turns the interval histogram into per-predicate counts:
//     count_k = sum over intervals i of hist[i] * mask[i][k].
// Row sizes are rounded up to a power of two so that a row's byte offset shifts down to its id.
inline uint32_t pred_row_words(int K)
{
    const uint32_t rw = ((uint32_t)K + 31) / 32;
    uint32_t p = 1;
    while (p < rw) p <<= 1;
    return p;
}

// global-memory layout of the index (all offsets 16-byte aligned)
struct PredIndexLayout {
    static constexpr size_t kHeader = 0;                                   // PredIndexHeader (32 B)
    static constexpr size_t kThr = 32;                                     // uint32 thr[kMaxBp]: (bp << (32-W)) - 1
    static constexpr size_t kLocator = kThr + (size_t)kMaxBp * 4;          // lut16[2^W] or uint2 bucket[4096]
    static constexpr size_t kLocatorBytes = (size_t)kBuckets * 8;          // 32 KB >= 2^13 * 2
    static constexpr size_t kMasks = kLocator + kLocatorBytes;             // uint32 mask[(M+1)][row_words]
    static constexpr size_t kMasksBytes = (size_t)(kMaxBp + 1) * (kMaxK / 32) * 4;
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
    uint32_t row_bytes;
    uint32_t row_shift;
    uint32_t num_bp;
    bool direct;
};

// shared-memory bytes the kernels need for the index of this launch
inline size_t pred_index_smem_bytes(int K, int w, bool* masks_in_smem)
{
    const size_t row_words = pred_row_words(K);
    const size_t max_rows = (w >= 31) ? (size_t)(2 * K + 1) : std::min<size_t>((size_t)2 * K + 1, (size_t)1 << w);
    const size_t mask_bytes = max_rows * row_words * 4;
    *masks_in_smem = mask_bytes <= (size_t)kMaskSmemBytes;
    const size_t locator = (w <= kDirectMaxW) ? ((size_t)2 << w) : (size_t)kBuckets * 8;
    const size_t thr = (w <= kDirectMaxW) ? 0 : std::min<size_t>(2 * (size_t)K, kMaxBp) * 4;
    return ((locator + 15) & ~(size_t)15) + ((thr + 15) & ~(size_t)15) + (*masks_in_smem ? ((mask_bytes + 15) & ~(size_t)15) : 0);
}

// All threads: copy the index into shared memory at `smem` (16-byte aligned).  Visible after the
// next __syncthreads.
__device__ __forceinline__ PredIndexView load_pred_index(const uint8_t* __restrict__ g_index, uint8_t* smem,
                                                         bool masks_in_smem)
{
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
    PredIndexView v;
    v.num_bp = h->num_bp;
    v.row_bytes = h->row_words * 4;
    v.row_shift = h->row_shift;
    v.direct = h->direct != 0;
    const uint32_t locator_bytes = v.direct ? (2u << h->w) : (uint32_t)kBuckets * 8;
    const uint32_t thr_bytes = v.direct ? 0u : v.num_bp * 4;
    const uint32_t mask_bytes = (v.num_bp + 1) * v.row_bytes;
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
    } else {
        v.masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
    }
    return v;
}

// byte offset of the mask row of a LEFT-aligned code t (garbage allowed in the low 32-W bits)
template <int W>
__device__ __forceinline__ uint32_t mask_row_offset(const PredIndexView& ix, uint32_t t)
{
    if constexpr (W <= kDirectMaxW) {
        return static_cast<const uint16_t*>(ix.locator)[t >> (32 - W)] * ix.row_bytes;
    } else {
        const uint2 e = static_cast<const uint2*>(ix.locator)[t >> (32 - kBucketBits)];
        // e.x = threshold of the bucket's first breakpoint (0xFFFFFFFF: none), e.y = first id | count << 16
        uint32_t id = (e.y & 0xFFFFu) + (t > e.x ? 1u : 0u);
        if (e.y >= (2u << 16)) {  // rare: several breakpoints inside one bucket
            const uint32_t first = e.y & 0xFFFFu, cnt = e.y >> 16;
            for (uint32_t i = 1; i < cnt; i++) id += (t > ix.thr[first + i]) ? 1u : 0u;
        }
        return id * ix.row_bytes;
    }
}

// ---- 32 x 32 bit-matrix transpose in registers ------------------------------------------------------
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
        for (int k = g; k < g + 4; k++) {  // 4 x 4 blocks
            const uint32_t lo = a[k], hi = a[k + 4];
            a[k] = (lo & 0x0F0F0F0Fu) | ((hi << 4) & 0xF0F0F0F0u);
            a[k + 4] = ((lo >> 4) & 0x0F0F0F0Fu) | (hi & 0xF0F0F0F0u);
        }
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; k++) {  // 2 x 2 blocks
            const uint32_t lo = a[k], hi = a[k + 2];
            a[k] = (lo & 0x33333333u) | ((hi << 2) & 0xCCCCCCCCu);
            a[k + 2] = ((lo >> 2) & 0x33333333u) | (hi & 0xCCCCCCCCu);
        }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {  // 1 x 1
        const uint32_t lo = a[k], hi = a[k + 1];
        a[k] = (lo & 0x55555555u) | ((hi << 1) & 0xAAAAAAAAu);
        a[k + 1] = ((lo >> 1) & 0x55555555u) | (hi & 0xAAAAAAAAu);
    }
}

}  // namespace sss
