// pack.cu -- the compressor (codes -> packed bit stream) and the synthetic-column generator.
//
// sss_pack replaces compress_9bit_input (reference src/simd_scan_compression.cpp:53-104) for any
// width 1..32 and uint32 codes: code i occupies stream bits [i*w, (i+1)*w), little-endian.  The
// reference scatters each code into one or two uint64 words with a serial cursor.
//
// pack_block_kernel<W> (full 32-code blocks): a thread packs one block = 32 codes -> W words with
// compile-time shifts.  Both sides of that are strided per thread (128 bytes in, 4W bytes out), so
// both go through a per-warp shared-memory transposition: the warp loads its 1024 codes with
// coalesced 16-byte loads, parks them in a [32 blocks][8+1 chunks] padded array (conflict-free both
// ways), each thread reads back its own block, packs, parks its W words at an odd word pitch, and the
// warp streams the 32*W words out as whole 128-byte lines.
//
// pack_kernel<Source> (generic, runtime w): one thread per OUTPUT word gathering the codes that
// overlap it.  Used for the ragged tail of sss_pack and for sss_generate_packed, which produces the
// benchmark column (uniform codes from a counter-based SplitMix64) directly in packed form -- the
// 2^30..2^34-code columns of BASELINE.json never exist unpacked.
//
// Roofline: HBM.  Algorithmic bytes per code = 4 + W/8.
#include <algorithm>
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace sss {

// must stay identical to orc_mix64 / orc_generate_codes (oracle/oracle.c)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct FromArray {
    const uint32_t* codes;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return __ldg(codes + i); }
};

struct FromGenerator {
    uint64_t seed, first;
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const
    {
        return (uint32_t)mix64(seed + (first + i) * 0xD1342543DE82EF95ull);
    }
};

// ---- generic: one thread per output word ---------------------------------------------------------

template <class Source>
__global__ void __launch_bounds__(kThreads)
pack_kernel(Source src, uint64_t n, int w, uint32_t* __restrict__ packed, uint64_t first_word, uint64_t end_word)
{
    const uint32_t mask = code_mask(w);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t word = first_word + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; word < end_word; word += stride) {
        const uint64_t bit0 = word * 32;
        uint64_t c = bit0 / (uint64_t)w;  // first code overlapping this word
        uint32_t acc = 0;
        for (; c < n; c++) {
            const int64_t rel = (int64_t)(c * (uint64_t)w) - (int64_t)bit0;  // code's offset inside the word
            if (rel >= 32) break;
            const uint32_t v = src(c) & mask;
            acc |= rel >= 0 ? (v << rel) : (v >> (-rel));
        }
        packed[word] = acc;  // words past the payload are written as zero padding
    }
}

// ---- full blocks: one thread per 32-code block, warp-level transposition through shared memory ----

template <int W, int J>
struct PackBlock {
    static __device__ __forceinline__ void run(const uint32_t (&c)[32], uint32_t (&out)[W])
    {
        constexpr CodeSlot s = code_slot(W, J);
        const uint32_t v = c[J] & code_mask(W);
        if constexpr (s.shift == 0) {
            out[s.first_word] = v;  // the first code to touch this word
        } else {
            out[s.first_word] |= v << s.shift;
        }
        if constexpr (s.straddles) out[s.last_word] = v >> (32 - s.shift);
        PackBlock<W, J + 1>::run(c, out);
    }
};
template <int W>
struct PackBlock<W, 32> {
    static __device__ __forceinline__ void run(const uint32_t (&)[32], uint32_t (&)[W]) {}
};

#ifndef SSS_PACK_WARPS
#define SSS_PACK_WARPS 4
#endif
// Measured (tools/ab_pack.py): short-lived CTAs beat a grid capped at 12 CTAs per SM by 10-30 %, and plain
// read-only loads beat evict-first streaming loads by 5-8 %.
#ifndef SSS_PACK_LOAD
#define SSS_PACK_LOAD __ldg
#endif
// Rounds (of 1024 codes) a warp packs back to back.  The codes of round r+1 are requested BEFORE round r is
// packed and written: with one round per warp the loads were outstanding for only ~40 % of a warp's life
// (ncu: issue 30 %, 5.6-5.9 TB/s for a read-dominated kernel whose traffic mix streams at 6.8 TB/s).
// Measured (profiles/r02_ab_pack_rounds.txt, fraction of the copy peak at 1 / 9 / 17 / 23 / 32 bit): one round
// 1.03 / 0.88 / 0.88 / 0.89 / 0.91, two rounds 1.02 / 1.06 / 1.04 / 1.02 / 0.87, four 1.02 / 1.03 / 1.01 / 1.00 / 0.84.
// At 32 bit the kernel is a copy (as many bytes written as read: the 1:1 mix tops out at 0.91) and stays
// with one round.
#ifndef SSS_PACK_ROUNDS
#define SSS_PACK_ROUNDS 2
#endif
constexpr int kPackWarps = SSS_PACK_WARPS;  // 128-thread CTAs: 4 warps x (codes 4608 B + words <= 4224 B) of shared memory
template <int W>
struct PackRounds {
    static constexpr int value = W == 32 ? 1 : SSS_PACK_ROUNDS;
};

template <int W>
__global__ void __launch_bounds__(kPackWarps * 32)
pack_block_kernel(const uint32_t* __restrict__ codes, uint64_t n_full_blocks, uint32_t* __restrict__ packed)
{
    constexpr int WP = W | 1;                                   // odd word pitch: conflict-free block rows
    __shared__ __align__(16) uint4 s_codes[kPackWarps][32 * 9];  // [block][8 chunks + 1 pad]
    __shared__ uint32_t s_words[kPackWarps][32 * WP];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t groups = (n_full_blocks + 31) / 32;  // a warp packs 32 blocks = 1024 codes per round
    constexpr int kPackRounds = PackRounds<W>::value;
    const uint64_t g0 = ((uint64_t)blockIdx.x * kPackWarps + warp) * kPackRounds;
    if (g0 >= groups) return;
    const uint32_t rounds = (uint32_t)((groups - g0) < (uint64_t)kPackRounds ? (groups - g0) : kPackRounds);

    // chunk c = i*32 + lane of a round -> block c/8, slot c%8; nb = blocks of the round (32 but for the last)
    auto blocks_of = [&](uint64_t g) {
        const uint64_t block0 = g * 32;
        return (uint32_t)((n_full_blocks - block0) < 32 ? (n_full_blocks - block0) : 32);
    };
    auto request = [&](uint64_t g, uint4 (&v)[8]) {
        const uint4* src = reinterpret_cast<const uint4*>(codes + g * 1024);
        const uint32_t nb = blocks_of(g);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t c = i * 32 + lane;
            if (c < nb * 8) v[i] = SSS_PACK_LOAD(src + c);
        }
    };
    uint4 v[8];
    request(g0, v);
    for (uint32_t r = 0; r < rounds; r++) {
        const uint64_t g = g0 + r;
        const uint32_t nb = blocks_of(g);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t c = i * 32 + lane;
            if (c < nb * 8) s_codes[warp][(c >> 3) * 9 + (c & 7)] = v[i];
        }
        __syncwarp();
        if (r + 1 < rounds) request(g + 1, v);  // in flight while this round is packed and written
        uint32_t out[W];
        if (lane < nb) {
            uint32_t cds[32];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 t = s_codes[warp][lane * 9 + q];
                cds[4 * q + 0] = t.x;
                cds[4 * q + 1] = t.y;
                cds[4 * q + 2] = t.z;
                cds[4 * q + 3] = t.w;
            }
            PackBlock<W, 0>::run(cds, out);
#pragma unroll
            for (int k = 0; k < W; k++) s_words[warp][lane * WP + k] = out[k];
        }
        __syncwarp();
        // stream the nb*W words out, 128 bytes per warp store
        uint32_t* dst = packed + g * 32 * W;
        const uint32_t total = nb * W;
        uint32_t b = lane / W, k = lane - b * W;  // word idx = b * W + k; idx += 32 per step
        for (uint32_t idx = lane; idx < total; idx += 32) {
            __stcs(dst + idx, s_words[warp][b * WP + k]);
            k += 32 % W;
            b += 32 / W;
            if (k >= (uint32_t)W) {
                k -= W;
                b++;
            }
        }
        __syncwarp();
    }
}

template <int W>
static cudaError_t launch_block_w(const Device& dev, const uint32_t* codes, uint64_t n_full_blocks, uint32_t* packed,
                                  cudaStream_t stream)
{
    (void)dev;
    const uint64_t groups = (n_full_blocks + 31) / 32;
    constexpr int kPackRounds = PackRounds<W>::value;
    const uint64_t want = (groups + (uint64_t)kPackWarps * kPackRounds - 1) / ((uint64_t)kPackWarps * kPackRounds);
    if (want > 0x7FFFFFFFull) return cudaErrorInvalidConfiguration;
    pack_block_kernel<W><<<(unsigned)want, kPackWarps * 32, 0, stream>>>(codes, n_full_blocks, packed);
    return cudaGetLastError();
}

using PackLauncher = cudaError_t (*)(const Device&, const uint32_t*, uint64_t, uint32_t*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<PackLauncher, sizeof...(Ws)>{&launch_block_w<Ws + 1>...};
}

static unsigned grid_for(const Device& dev, uint64_t items)
{
    const uint64_t want = (items + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)dev.sm_count * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_pack(const Device& dev, const uint32_t* codes, uint64_t n, int w, uint32_t* packed,
                        uint64_t packed_words, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    // full 32-code blocks through the block kernel (needs 16-byte aligned codes), the ragged tail
    // and the zero padding through the generic one
    uint64_t full_blocks = (reinterpret_cast<uintptr_t>(codes) & 15) == 0 ? n / 32 : 0;
    if (full_blocks) {
        cudaError_t e = table[w - 1](dev, codes, full_blocks, packed, stream);
        if (e != cudaSuccess) return e;
    }
    const uint64_t first_word = full_blocks * (uint64_t)w;
    if (first_word < packed_words)
        pack_kernel<<<grid_for(dev, packed_words - first_word), kThreads, 0, stream>>>(FromArray{codes}, n, w, packed,
                                                                                      first_word, packed_words);
    return cudaGetLastError();
}

cudaError_t launch_generate_packed(const Device& dev, uint32_t* packed, uint64_t packed_words, uint64_t n, int w,
                                   uint64_t seed, uint64_t first_index, cudaStream_t stream)
{
    pack_kernel<<<grid_for(dev, packed_words), kThreads, 0, stream>>>(FromGenerator{seed, first_index}, n, w, packed,
                                                                     0, packed_words);
    return cudaGetLastError();
}

}  // namespace sss
