// unpack.cu -- packed W-bit codes -> one zero-extended int32 per code.
//
// Replaces the nine decompress_* variants (reference src/simd_scan_decompression.cpp:6-470:
// movdqu + pshufb + pmulld/psrld or vpsrlvd + store).  The kernel is write-dominated (4 bytes out
// per W/8 bytes in), so the thread mapping is chosen for the stores: each lane owns 8 consecutive
// codes and writes them with one 256-bit store (sm_100 STG.E.256: a full 32-byte sector per lane), a
// warp therefore writes 1 KB of contiguous output per instruction.  Codes are pulled out of the shared-memory tile with a word pair + funnel
// shift (SHF.R.W), which is exact for every width 1..32 (the reference's 4-byte gather is not,
// SURVEY.md Appendix B).
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 + 4.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace sss {

template <int W>
struct UnpackCfg {
    static constexpr int kCodesPerThread = 8;                                // one 32-byte (256-bit) store per lane
#ifndef SSS_UNPACK_MIN_TILE_KB
#define SSS_UNPACK_MIN_TILE_KB 1
#endif
    // 8-code groups per thread per tile: 4, more at narrow widths so that a tile stays >= MIN_TILE_KB
    static constexpr int kR = (W * 4 >= SSS_UNPACK_MIN_TILE_KB * 4) ? 4 : (SSS_UNPACK_MIN_TILE_KB * 4 + W - 1) / W;
    static constexpr int kCodesPerTile = kThreads * kCodesPerThread * kR;    // 8192
    static constexpr int kTileBytes = kCodesPerTile / 32 * W * 4;            // 1 KB * W
    static constexpr int kStages = ring_stages(kTileBytes);
    using Ring = TileRing<kTileBytes, kStages>;
};

#ifndef SSS_UNPACK_CTAS
#define SSS_UNPACK_CTAS 2
#endif
template <int W>
__global__ void __launch_bounds__(kThreads, SSS_UNPACK_CTAS)
unpack_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n, int32_t* __restrict__ out,
              unsigned long long* __restrict__ tickets)
{
    using Cfg = UnpackCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();

    constexpr uint32_t mask = code_mask(W);
    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        const uint64_t tile_code0 = tile * (uint64_t)Cfg::kCodesPerTile;
#pragma unroll
        for (int r = 0; r < Cfg::kR; r++) {
            const uint32_t c0 = (r * kThreads + threadIdx.x) * Cfg::kCodesPerThread;  // code index in tile
            const uint64_t g0 = tile_code0 + c0;
            if (g0 < n) {
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const uint32_t bit = (c0 + e) * W;
                    const uint32_t wi = bit >> 5;
                    // buf[wi + 1] is only significant when the code straddles; the stage carries a
                    // 16-byte pad so the read is always inside shared memory
                    v[e] = __funnelshift_r(buf[wi], buf[wi + 1], bit & 31) & mask;
                }
                if (g0 + 8 <= n) {
#ifdef SSS_UNPACK_ST128
                    __stcs(reinterpret_cast<uint4*>(out + g0), make_uint4(v[0], v[1], v[2], v[3]));
                    __stcs(reinterpret_cast<uint4*>(out + g0) + 1, make_uint4(v[4], v[5], v[6], v[7]));
#else
                    st_v8_u32(out + g0, v);
#endif
                } else {
                    for (int e = 0; e < 8 && g0 + e < n; e++) out[g0 + e] = (int32_t)v[e];
                }
            }
        }
        ring.release(iter);
    }
    ring.finish();
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, int32_t* out, cudaStream_t stream)
{
    using Cfg = UnpackCfg<W>;
    constexpr int smem = Cfg::Ring::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e = cudaFuncSetAttribute(unpack_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t readable = round_up_u64(packed_payload_bytes(n, W), 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * SSS_UNPACK_CTAS;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    unpack_kernel<W><<<grid, kThreads, smem, stream>>>(packed, readable, n, out, dev.next_tickets());
    return cudaGetLastError();
}

using UnpackLauncher = cudaError_t (*)(const Device&, const void*, uint64_t, int32_t*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<UnpackLauncher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}

cudaError_t launch_unpack(const Device& dev, const void* packed, uint64_t n, int w, int32_t* out,
                          cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, out, stream);
}

}  // namespace sss
