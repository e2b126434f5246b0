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

// Threads per CTA of the index-driven kernels.  They are latency-bound (dependent shared-memory
// lookups), so more resident warps help: 2 CTAs x 320 threads = 20 warps/SM within the register file
// at <= 102 registers per thread.
#ifndef SSS_SHARED_THREADS
#define SSS_SHARED_THREADS 256
#endif
constexpr int kSThreads = SSS_SHARED_THREADS;

template <int W>
struct SharedCfg {
    static constexpr int kBlocksPerTile = kSThreads;  // one 32-code block per thread per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    // compute-bound kernels: a shallow ring is enough and leaves shared memory for the mask table
    static constexpr int kStages = kTileBytes >= 16 * 1024 ? 2 : 3;
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
};

constexpr int kCounterBytes = (kMaxBp + 1 + 3) / 4 * 16;  // up to 2K+1 uint32 interval counters, 16-byte multiple

// kSmem: the mask table fits in shared memory (otherwise it is read through L1 from global memory)
template <int W, bool kSmem>
__global__ void __launch_bounds__(kSThreads, 2)
shared_scan_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                   const uint8_t* __restrict__ g_index, int K, uint32_t* __restrict__ bitvecs, uint64_t stride_words,
                   bool count_matches)
{
    using Cfg = SharedCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    // per-CTA interval histogram (match counts are derived from it, see predindex.cuh)
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes);
    for (int i = threadIdx.x; i < 2 * K + 1; i += kSThreads) s_hist[i] = 0;
    const PredIndexView ix = load_pred_index(g_index, smem + Cfg::kRingBytes + kCounterBytes, kSmem);
    typename Cfg::Ring ring(smem, packed, readable_bytes16);
    ring.start();  // contains the __syncthreads that publishes the index and the zeroed counters

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const int cols = (K + 31) >> 5;

    uint32_t iter = 0;
    for (uint64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++iter) {
        ring.wait(iter);
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + threadIdx.x;
        const bool active = gb < n_blocks;
        uint32_t x[W];
#pragma unroll
        for (int k = 0; k < W; k++) x[k] = active ? buf[threadIdx.x * W + k] : 0u;
        uint32_t loaded = 0;
#pragma unroll
        for (int k = 0; k < W; k++) loaded |= x[k];
        ring.release_warp(iter, tile, loaded);  // the stage is free once every warp has copied its words
        if (active) {
            const bool last = tail_bits != 0 && gb == n_blocks - 1;  // the one ragged block of the column
            uint32_t id[32];
            block_ids<W>(x, ix, id);
            if (count_matches) {
                if (!last) {
#pragma unroll
                    for (int j = 0; j < 32; j++) atomicAdd(&s_hist[id[j]], 1u);
                } else {
                    for (uint32_t j = 0; j < tail_bits; j++) {
                        uint32_t v = 0;
#pragma unroll
                        for (int q = 0; q < 32; q++) v = (q == (int)j) ? id[q] : v;
                        atomicAdd(&s_hist[v], 1u);
                    }
                }
            }
            uint32_t* p = bitvecs + gb;  // this block's word of predicate 0; += stride_words per predicate
            if (K <= 8) {  // one byte of predicates per code: the cheap 32 x 8 transpose
                uint32_t m[32], r[8];
#pragma unroll
                for (int j = 0; j < 32; j++) m[j] = mask_word<kSmem>(ix, id[j], 0);
                transpose32x8(m, r);
                const uint32_t keep = last ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k < K) st_stream_u32(p, r[k] & keep);
                    p += stride_words;
                }
            } else {
                for (int c = 0; c < cols; c++) {
                    uint32_t m[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) m[j] = mask_word<kSmem>(ix, id[j], c);
                    transpose32(m);  // m[k'] bit j = predicate 32c+k' on code j
                    if (last) {
                        const uint32_t keep = (1u << tail_bits) - 1u;
#pragma unroll
                        for (int k = 0; k < 32; k++) m[k] &= keep;
                    }
                    const int kk = K - 32 * c;  // predicates in this column
                    if (kk >= 32) {
#pragma unroll
                        for (int k = 0; k < 32; k++) {
                            st_stream_u32(p, m[k]);
                            p += stride_words;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; k++) {
                            if (k < kk) st_stream_u32(p, m[k]);
                            p += stride_words;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (count_matches) {
        unsigned long long* g_hist =
            reinterpret_cast<unsigned long long*>(const_cast<uint8_t*>(g_index) + PredIndexLayout::kHist);
        for (uint32_t i = threadIdx.x; i <= ix.num_bp; i += kSThreads)
            if (s_hist[i]) atomicAdd(&g_hist[i], (unsigned long long)s_hist[i]);
    }
}

template <int W, bool kSmem>
static cudaError_t launch_ws(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                             uint32_t* bitvecs, uint64_t stride_words, bool count_matches, size_t smem,
                             cudaStream_t stream)
{
    using Cfg = SharedCfg<W>;
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
    const uint64_t max_ctas = (uint64_t)dev.sm_count * 2;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_kernel<W, kSmem><<<grid, kSThreads, smem, stream>>>(
        packed, readable, n, static_cast<const uint8_t*>(index), K, bitvecs, stride_words, count_matches);
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K,
                            uint32_t* bitvecs, uint64_t stride_words, bool count_matches, cudaStream_t stream)
{
    bool masks_in_smem = false;
    const size_t other = SharedCfg<W>::kRingBytes + kCounterBytes;
    const size_t smem = other + pred_index_smem_bytes(K, W, other, &masks_in_smem) + 64;
    if (masks_in_smem)
        return launch_ws<W, true>(dev, packed, n, index, K, bitvecs, stride_words, count_matches, smem, stream);
    return launch_ws<W, false>(dev, packed, n, index, K, bitvecs, stride_words, count_matches, smem, stream);
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const void*, int, uint32_t*, uint64_t, bool,
                                 cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_shared_scan(const Device& dev, const void* packed, uint64_t n, int w, const void* index, int K,
                               uint32_t* bitvecs, uint64_t stride_words, bool count_matches, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, index, K, bitvecs, stride_words, count_matches, stream);
}

}  // namespace sss
