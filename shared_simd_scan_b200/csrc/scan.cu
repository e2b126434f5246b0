// scan.cu -- fused unpack + range predicate -> bitvector + match count.
//
// Replaces scan_unvectorized / scan_128 / scan_128_unrolled / scan_256 / scan_256_unrolled
// (reference src/simd_scan.cpp:20-306).  Those walk the column 8 or 32 codes at a time with
// pshufb + pand + pcmpeqd + movmskps; here each thread owns one 32-code block (= W consecutive
// 32-bit words = one 32-bit word of the result bitvector, the same unit scan_*_unrolled writes,
// src/simd_scan.cpp:183,211), extracts every code LEFT-aligned with one funnel shift whose
// distance is a compile-time constant, and evaluates lo <= code <= hi as one subtract + one
// unsigned compare (see sss::aligned_range).  Decoded codes never leave registers.
//
// Data movement: persistent CTAs stream contiguous tiles of the packed column into a shared-memory
// ring with TMA bulk copies (common.cuh TileRing); lanes write consecutive result words, so every
// warp store is one full 128-byte line.
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 (packed input) + 1/8 (bitvector).
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace sss {

template <int W>
struct ScanCfg {
    // 32-code blocks per thread per tile: keeps tiles at 16-32 KB for every width
    static constexpr int kR = (W <= 1) ? 16 : (W <= 2) ? 8 : (W <= 5) ? 4 : (W <= 10) ? 2 : 1;
    static constexpr int kBlocksPerTile = kThreads * kR;  // = result words per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    static constexpr int kStages = (kTileBytes > 24 * 1024) ? 3 : 4;
    using Ring = TileRing<kTileBytes, kStages>;
};

template <int W, int J>
struct EvalBlock {
    static __device__ __forceinline__ uint32_t run(const uint32_t (&x)[W], uint32_t lo_al, uint32_t span_al)
    {
        const uint32_t hit = (left_aligned_code<W, J>(x) - lo_al) <= span_al ? (1u << J) : 0u;
        return hit | EvalBlock<W, J + 1>::run(x, lo_al, span_al);
    }
};
template <int W>
struct EvalBlock<W, 32> {
    static __device__ __forceinline__ uint32_t run(const uint32_t (&)[W], uint32_t, uint32_t) { return 0u; }
};

template <int W>
__global__ void __launch_bounds__(kThreads, 2)
scan_range_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n, uint32_t lo_al,
                  uint32_t span_al, uint32_t* __restrict__ bitvec, unsigned long long* __restrict__ count)
{
    using Cfg = ScanCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    typename Cfg::Ring ring(smem, packed, readable_bytes16);
    ring.start();

    const uint64_t n_blocks = (n + 31) >> 5;  // result words
    const uint32_t tail_bits = (uint32_t)(n & 31);
    uint32_t hits = 0;

    uint32_t iter = 0;
    for (uint64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++iter) {
        ring.wait(iter);
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        const uint64_t tile_block0 = tile * (uint64_t)Cfg::kBlocksPerTile;
#pragma unroll
        for (int r = 0; r < Cfg::kR; r++) {
            const uint32_t b = r * kThreads + threadIdx.x;
            const uint64_t gb = tile_block0 + b;
            if (gb < n_blocks) {
                uint32_t x[W];
#pragma unroll
                for (int k = 0; k < W; k++) x[k] = buf[b * W + k];
                uint32_t res = EvalBlock<W, 0>::run(x, lo_al, span_al);
                if (tail_bits != 0 && gb == n_blocks - 1) res &= (1u << tail_bits) - 1u;
                hits += __popc(res);
                st_stream_u32(bitvec + gb, res);
            }
        }
        ring.release(iter, tile);
    }
    if (count != nullptr) cta_add_u64(hits, count);
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, AlignedRange r, uint32_t* bitvec,
                            unsigned long long* count, cudaStream_t stream)
{
    using Cfg = ScanCfg<W>;
    constexpr int smem = Cfg::Ring::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e = cudaFuncSetAttribute(scan_range_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * 2;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    scan_range_kernel<W><<<grid, kThreads, smem, stream>>>(packed, readable, n, r.lo_al, r.span_al, bitvec, count);
    return cudaGetLastError();
}

using ScanLauncher = cudaError_t (*)(const Device&, const void*, uint64_t, AlignedRange, uint32_t*,
                                     unsigned long long*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<ScanLauncher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_scan_range(const Device& dev, const void* packed, uint64_t n, int w, AlignedRange r,
                              uint32_t* bitvec, unsigned long long* count, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, r, bitvec, count, stream);
}

}  // namespace sss
