// predindex.cu -- builds the predicate index (predindex.cuh) on the device: one CTA of 1024 threads.
//
//   1. every valid predicate contributes the breakpoints lo (if > 0) and hi+1 (if <= 2^W - 1)
//   2. bitonic sort of the (<= 2048) breakpoints in shared memory, 64-bit keys (2^32 = "none")
//   3. compaction of the distinct values (block prefix sum)
//   4. mask rows: row i, predicate k  <-  lo_k <= start_i <= hi_k   (fill_masks_kernel, many CTAs)
//   5. locator: direct lut (W <= 13) or 8192 buckets on the top 13 bits
#include <algorithm>

#include "predindex.cuh"

namespace sss {

constexpr int kBuildThreads = 1024;
constexpr unsigned long long kNone = 1ull << 32;

// number of breakpoints <= v  (bp sorted ascending, M entries)
__device__ __forceinline__ uint32_t count_le(const uint32_t* bp, uint32_t M, unsigned long long v)
{
    uint32_t lo = 0, hi = M;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((unsigned long long)bp[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kBuildThreads, 1)
build_pred_index_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int K, int w,
                        uint8_t* __restrict__ index)
{
    __shared__ unsigned long long keys[kMaxBp];
    __shared__ uint32_t bp[kMaxBp];
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_M;

    const uint32_t tid = threadIdx.x;
    const uint32_t maxv = code_mask(w);

    // 1. bounds -> candidate breakpoints
    for (uint32_t k = tid; k < (uint32_t)kMaxK; k += kBuildThreads) {
        unsigned long long b0 = kNone, b1 = kNone;
        if (k < (uint32_t)K) {
            const uint32_t l = lo[k];
            uint32_t h = hi[k];
            if (l <= h && l <= maxv) {  // valid predicate (hi clamps to the domain)
                if (h > maxv) h = maxv;
                if (l > 0) b0 = l;
                if (h < maxv) b1 = (unsigned long long)h + 1;
            }
        }
        keys[2 * k] = b0;
        keys[2 * k + 1] = b1;
    }
    __syncthreads();

    // 2. bitonic sort, 2048 keys, 1024 threads (one compare-exchange each per step)
    for (uint32_t size = 2; size <= (uint32_t)kMaxBp; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            const uint32_t pos = 2 * tid - (tid & (stride - 1));
            const bool up = ((pos & size) == 0);
            const unsigned long long a = keys[pos], b = keys[pos + stride];
            if ((a > b) == up) {
                keys[pos] = b;
                keys[pos + stride] = a;
            }
            __syncthreads();
        }
    }

    // 3. distinct values -> bp[0..M)
    uint32_t flags[2], incl;
    {
        uint32_t c = 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const uint32_t i = 2 * tid + e;
            const bool first = keys[i] != kNone && (i == 0 || keys[i] != keys[i - 1]);
            flags[e] = first ? 1u : 0u;
            c += flags[e];
        }
        incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((tid & 31) >= (uint32_t)d) incl += o;
        }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            uint32_t t = warp_tot[tid], ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, ti, d);
                if (tid >= (uint32_t)d) ti += o;
            }
            warp_tot[tid] = ti - t;
            if (tid == 31) s_M = ti;
        }
        __syncthreads();
        uint32_t at = warp_tot[tid >> 5] + incl - c;
#pragma unroll
        for (int e = 0; e < 2; e++)
            if (flags[e]) bp[at++] = (uint32_t)keys[2 * tid + e];
    }
    __syncthreads();
    const uint32_t M = s_M;
    const uint32_t row_words = (((uint32_t)K + 31) / 32) | 1u;  // odd stride, see pred_row_stride_words

    // header + thresholds of the left-aligned compare:  t >= bp << (32-W)  <=>  t > thr
    if (tid == 0) {
        PredIndexHeader* h = reinterpret_cast<PredIndexHeader*>(index + PredIndexLayout::kHeader);
        h->num_bp = M;
        h->row_words = row_words;
        h->direct = (w <= kDirectMaxW) ? 1u : 0u;
        h->K = (uint32_t)K;
        h->w = (uint32_t)w;
        h->reserved = 0;
    }
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(index + PredIndexLayout::kHist);
    for (uint32_t i = tid; i <= M; i += kBuildThreads) hist[i] = 0;
    uint32_t* thr = reinterpret_cast<uint32_t*>(index + PredIndexLayout::kThr);
    for (uint32_t i = tid; i < M; i += kBuildThreads) thr[i] = (bp[i] << (32 - w)) - 1u;
    if (tid == 0) thr[M] = 0xFFFFFFFFu;  // sentinel: ends every forward walk

    // 4. mask rows: filled by fill_masks_kernel (many CTAs) right after this kernel

    // 5. locator
    if (w <= kDirectMaxW) {
        uint16_t* lut = reinterpret_cast<uint16_t*>(index + PredIndexLayout::kLocator);
        for (uint32_t v = tid; v < (1u << w); v += kBuildThreads) lut[v] = (uint16_t)count_le(bp, M, v);
    } else {
        uint32_t* bucket = reinterpret_cast<uint32_t*>(index + PredIndexLayout::kLocator);
        const int shift = w - kBucketBits;
        for (uint32_t b = tid; b < (uint32_t)kBuckets; b += kBuildThreads) {
            const unsigned long long start = (unsigned long long)b << shift;
            const unsigned long long end = start + (1ull << shift) - 1;
            const uint32_t base = count_le(bp, M, start);        // breakpoints <= bucket start: in the id already
            const uint32_t inner = count_le(bp, M, end) - base;  // start < breakpoint <= end
            // threshold on the bits below the bucket bits of the LEFT-aligned code
            const uint32_t thr = inner ? (((bp[base] << (32 - w)) & kThrMask) - 1u) : kThrMask;
            bucket[b] = make_bucket_word(base, inner >= 2, thr);
        }
    }
}

// mask rows, one thread per 32-bit word: row i, word c, bit b  <-  lo_k <= start_i <= hi_k, k = 32c + b.
// Reads the breakpoints back from the thresholds the build kernel wrote (bp = (thr + 1) >> (32 - w)).
__global__ void __launch_bounds__(256)
fill_masks_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int K, int w,
                  uint8_t* __restrict__ index)
{
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(index + PredIndexLayout::kHeader);
    const uint32_t M = h->num_bp, row_words = h->row_words;
    const uint32_t* thr = reinterpret_cast<const uint32_t*>(index + PredIndexLayout::kThr);
    uint32_t* masks = reinterpret_cast<uint32_t*>(index + PredIndexLayout::kMasks);
    const uint32_t maxv = code_mask(w);
    const uint32_t total_words = (M + 1) * row_words;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_words; idx += gridDim.x * blockDim.x) {
        const uint32_t row = idx / row_words, c = idx - row * row_words;
        const uint32_t start = row == 0 ? 0u : ((thr[row - 1] + 1u) >> (32 - w));
        uint32_t word = 0;
        for (uint32_t b = 0; b < 32; b++) {
            const uint32_t k = c * 32 + b;
            if (k < (uint32_t)K) {
                const uint32_t l = __ldg(lo + k);
                uint32_t hh = __ldg(hi + k);
                if (l <= hh && l <= maxv) {  // valid predicate (hi clamps to the domain)
                    if (hh > maxv) hh = maxv;
                    word |= ((l <= start && start <= hh) ? 1u : 0u) << b;
                }
            }
        }
        masks[idx] = word;
    }
}

// counts[k] += sum over intervals i of hist[i] * mask[i][k].  One CTA per 32 predicates; thread
// (slice, p) sums the rows i = slice, slice + 32, ... for predicate p.  The histogram is cleared by
// the next build_pred_index_kernel.
__global__ void __launch_bounds__(1024)
counts_from_histogram_kernel(const uint8_t* __restrict__ index, int K, unsigned long long* __restrict__ counts)
{
    __shared__ unsigned long long acc[32];
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(index + PredIndexLayout::kHeader);
    const uint32_t M = h->num_bp, row_words = h->row_words;
    const uint32_t* masks = reinterpret_cast<const uint32_t*>(index + PredIndexLayout::kMasks);
    const unsigned long long* hist = reinterpret_cast<const unsigned long long*>(index + PredIndexLayout::kHist);
    const uint32_t p = threadIdx.x & 31, slice = threadIdx.x >> 5, c = blockIdx.x;
    if (threadIdx.x < 32) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long sum = 0;
    for (uint32_t i = slice; i <= M; i += 32)
        if ((masks[i * row_words + c] >> p) & 1u) sum += hist[i];
    if (sum) atomicAdd(&acc[p], sum);
    __syncthreads();
    const uint32_t k = c * 32 + threadIdx.x;
    if (threadIdx.x < 32 && k < (uint32_t)K) counts[k] += acc[threadIdx.x];
}

size_t pred_index_bytes() { return PredIndexLayout::kTotal; }

cudaError_t launch_counts_from_histogram(void* d_index, int K, unsigned long long* d_counts, cudaStream_t stream)
{
    counts_from_histogram_kernel<<<(K + 31) / 32, 1024, 0, stream>>>(static_cast<const uint8_t*>(d_index), K, d_counts);
    return cudaGetLastError();
}

cudaError_t launch_build_pred_index(const uint32_t* d_lo, const uint32_t* d_hi, int K, int w, void* d_index,
                                    cudaStream_t stream)
{
    build_pred_index_kernel<<<1, kBuildThreads, 0, stream>>>(d_lo, d_hi, K, w, static_cast<uint8_t*>(d_index));
    const unsigned words = (unsigned)((2 * K + 1) * pred_row_stride_words(K));
    fill_masks_kernel<<<(words + 255) / 256, 256, 0, stream>>>(d_lo, d_hi, K, w, static_cast<uint8_t*>(d_index));
    return cudaGetLastError();
}

}  // namespace sss
