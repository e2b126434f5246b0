// shared.cu -- shared scan: K range predicates evaluated in ONE pass over the packed column,
// producing K bitvectors and K match counts.
//
// Replaces shared_scan_128_standard / _standard_unrolled / _parallel / _sequential* / _threaded and
// the 256-bit twins (reference src/simd_scan_shared.cpp:9-466).  Like shared_scan_128_standard
// (:49-86) the column is decoded once and every predicate is evaluated against the decoded codes;
// unlike it the unit of work is a 32-code block per thread, so each predicate's result is a whole
// 32-bit word and lanes write consecutive words of the same bitvector (128-byte lines) instead of
// the reference's K scattered single-byte stores (:82).
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 + K/8.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "predicates.cuh"

namespace sss {

template <int W>
struct SharedCfg {
    static constexpr int kBlocksPerTile = kThreads;  // one 32-code block per thread per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    static constexpr int kStages = ring_stages(kTileBytes);
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kPredSmem = 3 * SSS_MAX_K * 4 + SSS_MAX_K * 4;  // lo_al, span_al, valid, counts
    static constexpr int kSmemBytes = Ring::kSmemBytes + 8 /*align*/ + kPredSmem;
};

template <int W>
__global__ void __launch_bounds__(kThreads, 2)
shared_scan_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                   const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int K,
                   uint32_t* __restrict__ bitvecs, uint64_t stride_words, unsigned long long* __restrict__ counts)
{
    using Cfg = SharedCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    PredicateTable preds(smem + ((Cfg::Ring::kSmemBytes + 15) & ~15));
    preds.load(W, lo, hi, K);
    typename Cfg::Ring ring(smem, packed, readable_bytes16);
    ring.start();  // contains the __syncthreads that publishes the predicate table

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31;

    uint32_t iter = 0;
    for (uint64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++iter) {
        ring.wait(iter);
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + threadIdx.x;
        const bool active = gb < n_blocks;
        uint32_t t[32];
        if (active) {
            uint32_t x[W];
#pragma unroll
            for (int k = 0; k < W; k++) x[k] = buf[threadIdx.x * W + k];
            LeftAlignAll<W, 0>::run(x, t);
        }
        const uint32_t keep = (active && tail_bits != 0 && gb == n_blocks - 1) ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;
        for (int k = 0; k < K; k++) {
            uint32_t res = 0;
            if (active) {
                const uint32_t lo_al = preds.lo_al[k], span_al = preds.span_al[k];
#pragma unroll
                for (int j = 0; j < 32; j++) res |= ((t[j] - lo_al) <= span_al ? 1u : 0u) << j;
                res &= preds.valid[k] & keep;
                st_stream_u32(bitvecs + (uint64_t)k * stride_words + gb, res);
            }
            const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, __popc(res));
            if (lane == 0 && c) atomicAdd(&preds.counts[k], c);
        }
        ring.release(iter, tile);
    }
    preds.flush_counts(counts, K);
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const uint32_t* lo, const uint32_t* hi,
                            int K, uint32_t* bitvecs, uint64_t stride_words, unsigned long long* counts,
                            cudaStream_t stream)
{
    using Cfg = SharedCfg<W>;
    constexpr int smem = Cfg::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * 2;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_kernel<W><<<grid, kThreads, smem, stream>>>(packed, readable, n, lo, hi, K, bitvecs, stride_words,
                                                           counts);
    return cudaGetLastError();
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const uint32_t*, const uint32_t*, int,
                                 uint32_t*, uint64_t, unsigned long long*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_shared_scan(const Device& dev, const void* packed, uint64_t n, int w, const uint32_t* lo,
                               const uint32_t* hi, int K, uint32_t* bitvecs, uint64_t stride_words,
                               unsigned long long* counts, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, lo, hi, K, bitvecs, stride_words, counts, stream);
}

}  // namespace sss
