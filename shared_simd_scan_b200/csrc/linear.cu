// linear.cu -- shared scan with ONE linear output: the byte for (group g of 8 codes, predicate k)
// lives at out[g*K + k].
//
// Replaces shared_scan_128_linear_standard / shared_scan_128_linear_simple /
// shared_scan_128_linear_static<NUM> (reference src/simd_scan_shared_linear.cpp:9-82,
// src/simd_scan.hpp:122-236).  The layout is pinned by the reference's only shared-scan unit test
// (test/simd_scan_tests.cpp:108-150).
//
// A thread owns one 32-code block = 4 output rows.  For four predicates at a time it builds the four
// 32-bit result words, regroups their bytes with PRMT into one 4-byte piece per row and stores it,
// so every store covers 4 consecutive bytes of a K-byte row.
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 + K/8.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "predicates.cuh"

namespace sss {

template <int W>
struct LinearCfg {
    static constexpr int kBlocksPerTile = kThreads;
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    static constexpr int kStages = ring_stages(kTileBytes);
    using Ring = TileRing<kTileBytes, kStages>;
    static constexpr int kPredSmem = 4 * SSS_MAX_K * 4;
    static constexpr int kSmemBytes = Ring::kSmemBytes + 8 + kPredSmem;
};

__device__ __forceinline__ uint32_t eval_word(const uint32_t (&t)[32], uint32_t lo_al, uint32_t span_al)
{
    uint32_t res = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) res |= ((t[j] - lo_al) <= span_al ? 1u : 0u) << j;
    return res;
}

template <int W>
__global__ void __launch_bounds__(kThreads, 2)
shared_scan_linear_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                          const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int K,
                          uint8_t* __restrict__ out, unsigned long long* __restrict__ counts)
{
    using Cfg = LinearCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    PredicateTable preds(smem + ((Cfg::Ring::kSmemBytes + 15) & ~15));
    preds.load(W, lo, hi, K);
    typename Cfg::Ring ring(smem, packed, readable_bytes16);
    ring.start();

    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t lane = threadIdx.x & 31;
    const bool vec4 = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;

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
        uint8_t* row0 = out + gb * 4 * (uint64_t)K;  // rows 4*gb .. 4*gb+3
        if (vec4) {
            for (int k = 0; k < K; k += 4) {
                uint32_t r[4] = {0, 0, 0, 0};
                if (active) {
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        r[q] = eval_word(t, preds.lo_al[k + q], preds.span_al[k + q]) & preds.valid[k + q] & keep;
                    // byte i of r[q] is row i, predicate k+q  ->  one word per row
                    const uint32_t a01 = __byte_perm(r[0], r[1], 0x5140);  // r0.b0 r1.b0 r0.b1 r1.b1
                    const uint32_t a23 = __byte_perm(r[2], r[3], 0x5140);
                    const uint32_t b01 = __byte_perm(r[0], r[1], 0x7362);  // r0.b2 r1.b2 r0.b3 r1.b3
                    const uint32_t b23 = __byte_perm(r[2], r[3], 0x7362);
                    uint32_t* p = reinterpret_cast<uint32_t*>(row0 + k);
                    const uint64_t rs = (uint64_t)K / 4;
                    __stcs(p, __byte_perm(a01, a23, 0x5410));
                    __stcs(p + rs, __byte_perm(a01, a23, 0x7632));
                    __stcs(p + 2 * rs, __byte_perm(b01, b23, 0x5410));
                    __stcs(p + 3 * rs, __byte_perm(b01, b23, 0x7632));
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, __popc(r[q]));
                    if (lane == 0 && c) atomicAdd(&preds.counts[k + q], c);
                }
            }
        } else {
            for (int k = 0; k < K; k++) {
                uint32_t r = 0;
                if (active) {
                    r = eval_word(t, preds.lo_al[k], preds.span_al[k]) & preds.valid[k] & keep;
#pragma unroll
                    for (int i = 0; i < 4; i++) row0[(uint64_t)i * K + k] = (uint8_t)(r >> (8 * i));
                }
                const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, __popc(r));
                if (lane == 0 && c) atomicAdd(&preds.counts[k], c);
            }
        }
        ring.release(iter, tile);
    }
    preds.flush_counts(counts, K);
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const uint32_t* lo, const uint32_t* hi,
                            int K, uint8_t* out, unsigned long long* counts, cudaStream_t stream)
{
    using Cfg = LinearCfg<W>;
    constexpr int smem = Cfg::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e =
            cudaFuncSetAttribute(shared_scan_linear_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * 2;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    shared_scan_linear_kernel<W><<<grid, kThreads, smem, stream>>>(packed, readable, n, lo, hi, K, out, counts);
    return cudaGetLastError();
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const uint32_t*, const uint32_t*, int,
                                 uint8_t*, unsigned long long*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_shared_scan_linear(const Device& dev, const void* packed, uint64_t n, int w, const uint32_t* lo,
                                      const uint32_t* hi, int K, uint8_t* out, unsigned long long* counts,
                                      cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, lo, hi, K, out, counts, stream);
}

}  // namespace sss
