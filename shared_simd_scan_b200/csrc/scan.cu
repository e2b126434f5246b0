// scan.cu -- fused unpack + range predicate -> bitvector + match count.
//
// Replaces scan_unvectorized / scan_128 / scan_128_unrolled / scan_256 / scan_256_unrolled
// (reference src/simd_scan.cpp:20-306).  Those walk the column 8 or 32 codes at a time with
// pshufb + pand + pcmpeqd + movmskps; here each thread owns 32-code blocks (= W consecutive 32-bit
// words = one 32-bit word of the result bitvector, the same unit scan_*_unrolled writes,
// src/simd_scan.cpp:183,211) and evaluates lo <= code <= hi on them without ever materialising the
// decoded codes in memory.  Three evaluators, chosen per width at compile time:
//
//   W = 1        the predicate is a boolean function of the bit itself: one LOP3 per 32 codes
//   W = 2..12    word-parallel compare on the packed words (swar.cuh): ~8W/32 instructions per code;
//                these widths are integer-issue bound, not HBM bound, with a per-code pipeline
//   W = 13..32   per-code pipeline: one funnel shift with a compile-time distance moves the code
//                LEFT-aligned into a register, then one subtract + one unsigned compare
//                (sss::aligned_range); already at the HBM roofline for these widths
//
// Data movement: persistent CTAs stream contiguous tiles of the packed column into a shared-memory
// ring with TMA bulk copies (common.cuh TileRing).  A thread reads its blocks from shared memory
// with the widest conflict-free access for its width (128-bit loads over 1-4 blocks whose chunk
// stride is odd; for W = 8/16/32, where the stride is a power of two, each lane starts at a
// different 16-byte chunk of its block and the result word is rotated back with one funnel shift).
// Lanes write consecutive result words, so every warp store covers whole 128-byte lines.
//
// Roofline: HBM.  Algorithmic bytes per code = W/8 (packed input) + 1/8 (bitvector).
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "swar.cuh"

namespace sss {

// developer switch: plain instead of streaming (evict-first) stores of the result words
#ifndef SSS_SCAN_PLAINST
#define SSS_SCAN_PLAINST 0
#endif
#if SSS_SCAN_PLAINST
#define SSS_SCAN_ST(p, v) (*(p) = (v))
#else
#define SSS_SCAN_ST(p, v) __stcs((p), (v))
#endif

// widths whose scans always take the ring-less kernel (see scan_direct_kernel)
#ifndef SSS_SCAN_DIRECT_MAXW
#define SSS_SCAN_DIRECT_MAXW 2
#endif

enum ScanMode : int { kScalar = 0, kSwarLoTop0 = 1, kSwarLoTop1 = 2, kOneBit = 3, kBool2 = 4 };

template <int W>
struct ScanPlan {
    static constexpr bool kSwar = (W >= 2 && W <= 12);
    static constexpr bool kRot = (W == 8 || W == 16 || W == 32);
    // 128-bit shared-memory loads (group of kG blocks per thread, kG*W a multiple of 4 words)
    static constexpr bool kVec = (W == 1) || kRot || (kSwar && W != 11);
    static constexpr int kG = !kVec ? 1 : (W == 1 ? 4 : (W % 4 == 0 ? 1 : (W % 2 == 0 ? 2 : 4)));
    static constexpr int kGroupWords = kG * W;
#ifndef SSS_SCAN_TILE_WORDS
#define SSS_SCAN_TILE_WORDS 32  // target words per thread per tile (x 1 KiB = tile bytes).  16 in round 1; with
#endif                          // tiles drawn from a ticket counter, tiles below ~20 KB (3, 6, 11-13 bit) lost 5-15 %
#ifndef SSS_SCAN_MAX_CTAS
#define SSS_SCAN_MAX_CTAS 2
#endif
    // (W = 1 runs ring-less, one CTA per tile -- scan_direct_kernel -- and does best with 32 words per thread)
    static constexpr int kTileWords = (W == 1) ? 32 : SSS_SCAN_TILE_WORDS;
    static constexpr int kR = (kTileWords / kGroupWords) > 1 ? (kTileWords / kGroupWords) : 1;  // groups per thread per tile
    static constexpr int kBlocksPerTile = kThreads * kR * kG;                                // = result words per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
    static constexpr int kStages = ring_stages(kTileBytes);
    using Ring = TileRing<kTileBytes, kStages>;
    // resident CTAs per SM: as many as shared memory allows (227 KB usable, 1 KB reserved per CTA)
    static constexpr int kFit = (227 * 1024) / (Ring::kSmemBytes + 1024 + 64);
    static constexpr int kCtasPerSm = kFit < 1 ? 1 : (kFit > SSS_SCAN_MAX_CTAS ? SSS_SCAN_MAX_CTAS : kFit);
};

struct ScalarArgs {
    uint32_t lo_al, span_al;
};
struct OneBitArgs {
    uint32_t flip, all;  // result = (x ^ flip) | all
};
// W = 2: the predicate is a boolean function of the code's two bits; m[v] = 0x55555555 if code value v
// matches, else 0 (bit 2i of the result word = the predicate on field i)
struct Bool2Args {
    uint32_t m[4];
};
inline __host__ __device__ Bool2Args make_bool2(uint32_t lo, uint32_t hi)
{
    Bool2Args a;
    for (uint32_t v = 0; v < 4; v++) a.m[v] = (lo <= v && v <= hi) ? 0x55555555u : 0u;
    return a;
}
template <int W>
struct ScanArgs {
    ScalarArgs scalar;
    OneBitArgs onebit;
    Bool2Args bool2;
    std::conditional_t<(W >= 2 && W <= 16), SwarConsts<(W >= 2 && W <= 16) ? W : 2>, ScalarArgs> swar;
};

// ---- per-code evaluator (W >= 13) -----------------------------------------------------------------

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

template <int W, int kMode>
__device__ __forceinline__ uint32_t eval_block(const uint32_t (&x)[W], const ScanArgs<W>& a)
{
    if constexpr (kMode == kOneBit) {
        return (x[0] ^ a.onebit.flip) | a.onebit.all;
    } else if constexpr (kMode == kBool2) {
        // Two-bit codes.  The word-parallel subtract chains of swar.cuh cost ~36 ALU-pipe instructions per
        // 32 codes here -- at 2 bits the HBM roofline is 17 T codes/s and the kernel was ALU-bound at 0.85 of
        // it.  A predicate on a 2-bit code is one of 16 boolean functions of (b1, b0): a two-level
        // multiplexer over four launch-time masks evaluates it for the 16 fields of a word with 1 shift +
        // 3 LOP3.  Word 0's results sit at the even bits; word 1's are added at the odd bits (one IMAD) and a
        // three-stage inverse perfect shuffle + one byte permute bring code i to bit i.
        static_assert(W == 2 || kMode != kBool2, "two-bit evaluator");
        uint32_t r[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint32_t b0 = x[i % W], b1 = x[i % W] >> 1;
            const uint32_t hi_half = (a.bool2.m[3] & b0) | (a.bool2.m[2] & ~b0);  // b1 = 1: values 3 / 2
            const uint32_t lo_half = (a.bool2.m[1] & b0) | (a.bool2.m[0] & ~b0);  // b1 = 0: values 1 / 0
            r[i] = (hi_half & b1) | (lo_half & ~b1);
        }
        uint32_t z = r[1] * 2u + r[0];  // (disjoint bits: the add is an OR, issued on the FMA pipe)
        uint32_t u;
        u = (z ^ (z >> 1)) & 0x22222222u; z = z ^ u ^ (u << 1);
        u = (z ^ (z >> 2)) & 0x0C0C0C0Cu; z = z ^ u ^ (u << 2);
        u = (z ^ (z >> 4)) & 0x00F000F0u; z = z ^ u ^ (u << 4);
        return __byte_perm(z, 0u, 0x3120);  // the 8-bit stage of the shuffle swaps bytes 1 and 2
    } else if constexpr (kMode == kScalar) {
        return EvalBlock<W, 0>::run(x, a.scalar.lo_al, a.scalar.span_al);
    } else {
        return swar_eval_block<W, kMode == kSwarLoTop1>(x, a.swar);
    }
}

// ---- shared memory -> registers ---------------------------------------------------------------------

// lane-dependent starting chunk for the power-of-two strides (see file header)
template <int W>
__device__ __forceinline__ uint32_t rot_chunks()
{
    const uint32_t lane = threadIdx.x & 31;
    if constexpr (W == 32) return lane & 7;
    if constexpr (W == 16) return (lane >> 1) & 3;
    if constexpr (W == 8) return (lane >> 2) & 1;
    return 0;
}

// words of thread group `tg` (kG consecutive blocks) -> x[kG][W]
template <int W>
__device__ __forceinline__ void load_group(const uint32_t* buf, uint32_t tg, uint32_t (&x)[ScanPlan<W>::kG][W],
                                           uint32_t rot)
{
    using P = ScanPlan<W>;
    if constexpr (P::kVec) {
        constexpr int chunks = P::kGroupWords / 4;
        const uint4* src = reinterpret_cast<const uint4*>(buf) + (size_t)tg * chunks;
        uint32_t* flat = &x[0][0];
#pragma unroll
        for (int q = 0; q < chunks; q++) {
            int c = q;
            if constexpr (P::kRot) c = (q + rot) & (chunks - 1);
            const uint4 v = src[c];
            flat[4 * q + 0] = v.x;
            flat[4 * q + 1] = v.y;
            flat[4 * q + 2] = v.z;
            flat[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) x[0][k] = buf[(size_t)tg * W + k];
    }
}

// ---- the kernel ---------------------------------------------------------------------------------

// One tile.  kChecked = false is the hot path (every block of the tile holds 32 valid codes: no
// bounds or tail handling, 32-bit indexing); kChecked = true handles the last, ragged tile.
template <int W, int kMode, bool kChecked, bool kSeg = false, int kR = ScanPlan<W>::kR>
__device__ __forceinline__ uint32_t scan_tile(const uint32_t* buf, const ScanArgs<W>& args, uint32_t rot,
                                              uint32_t* __restrict__ tile_out, uint64_t blocks_left,
                                              uint32_t tail_bits, uint32_t* __restrict__ tile_seg = nullptr)
{
    using P = ScanPlan<W>;
    // kSeg: the matches of every position-list segment (2^kPosSegShift result words) are added to
    // tile_seg[segment - the tile's first segment], CTA-shared counters that the caller adds to the global
    // segment counts after the tile's barrier.  A round's 256 * kG blocks and a tile's kBlocksPerTile are
    // powers of two, so a warp's blocks of a round lie in ONE segment, the same for every warp of the CTA,
    // known at compile time: one REDUX + one shared-memory atomic per warp and segment (or tile, if shorter),
    // none when the warp found nothing.  (Global atomics straight from the warps -- 2^20 of them per 2^30
    // codes, 32 per counter, all concurrent ones within a few cache lines -- cost 60 us per scan.)
    constexpr int kRoundBlocks = kThreads * P::kG;
    static_assert(!kSeg || ((kRoundBlocks & (kRoundBlocks - 1)) == 0 && kRoundBlocks <= (1 << kPosSegShift)),
                  "a round must not straddle position-list segments");
    uint32_t hits = 0, seg_hits = 0;
#pragma unroll
    for (int r = 0; r < kR; r++) {
        const uint32_t tg = r * kThreads + threadIdx.x;
        const uint32_t b0 = tg * P::kG;  // first block of this group, relative to the tile
        if (!kChecked || b0 < blocks_left) {
            uint32_t x[P::kG][W];
            load_group<W>(buf, tg, x, rot);
            uint32_t res[P::kG];
#pragma unroll
            for (int b = 0; b < P::kG; b++) {
                uint32_t v = eval_block<W, kMode>(x[b], args);
                if constexpr (P::kRot) v = __funnelshift_l(v, v, rot * (128 / W));
                if constexpr (kChecked) {
                    if (tail_bits != 0 && b0 + b == blocks_left - 1) v &= (1u << tail_bits) - 1u;
                    if (b0 + b >= blocks_left) v = 0;
                }
                if constexpr (kSeg) seg_hits += __popc(v);
                else hits += __popc(v);
                res[b] = v;
            }
            uint32_t* dst = tile_out + b0;
            if (P::kG == 4 && (!kChecked || b0 + 4 <= blocks_left)) {
                SSS_SCAN_ST(reinterpret_cast<uint4*>(dst), make_uint4(res[0], res[1 % P::kG], res[2 % P::kG], res[3 % P::kG]));
            } else if (P::kG == 2 && (!kChecked || b0 + 2 <= blocks_left)) {
                SSS_SCAN_ST(reinterpret_cast<uint2*>(dst), make_uint2(res[0], res[1 % P::kG]));
            } else {
#pragma unroll
                for (int b = 0; b < P::kG; b++)
                    if (!kChecked || b0 + b < blocks_left) SSS_SCAN_ST(dst + b, res[b]);
            }
        }
        if constexpr (kSeg) {
            constexpr int sh = kPosSegShift;
            const int seg_rel = (r * kRoundBlocks) >> sh;
            if (r == kR - 1 || (((r + 1) * kRoundBlocks) >> sh) != seg_rel) {  // (compile-time per unrolled r)
                const uint32_t c = __reduce_add_sync(0xFFFFFFFFu, seg_hits);
                // (explicit red.shared from lane 0: atomicAdd() on shared memory brings a warp-aggregation
                // preamble of a dozen instructions and a divergent branch)
                if ((threadIdx.x & 31) == 0)
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(tile_seg + seg_rel)), "r"(c) : "memory");
                hits += seg_hits;
                seg_hits = 0;
            }
        }
    }
    return hits;
}

// every tile this CTA owns (static round-robin); returns the thread's match count
// position-list segments a tile covers (tiles shorter than a segment add to the one they lie in)
template <int kBlocksPerTile>
constexpr int kSegsPerTile = kBlocksPerTile > (1 << kPosSegShift) ? (kBlocksPerTile >> kPosSegShift) : 1;
constexpr int kMaxSegsPerTile = 8;

// after the tile's barrier: the CTA's counters of the tile go to the global segment counts and are cleared
template <int kBlocksPerTile>
__device__ __forceinline__ void flush_seg_counts(uint32_t* acc, uint32_t* __restrict__ seg_count, uint64_t tile_block0)
{
    static_assert(kSegsPerTile<kBlocksPerTile> <= kMaxSegsPerTile, "seg_acc rows");
    // (the last warp's first lanes: thread 0 is busy refilling the ring)
    const uint32_t t = threadIdx.x - (kThreads - 32);
    if (t < (uint32_t)kSegsPerTile<kBlocksPerTile>) {
        const uint32_t c = acc[t];
        if (c != 0) {
            acc[t] = 0;
            // an explicit RED: atomicAdd() with its result unused came out as ATOMG ..., RZ, and the flushing
            // thread -- the one that refills the ring -- then sat out an L2 round trip per tile (+30 % per scan)
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(seg_count + (tile_block0 >> kPosSegShift) + t),
                         "r"(c)
                         : "memory");
        }
    }
}

template <int W, int kMode, bool kSeg = false>
__device__ __forceinline__ uint32_t scan_tiles(typename ScanPlan<W>::Ring& ring, const ScanArgs<W>& args, uint64_t n,
                                               uint32_t* __restrict__ bitvec, uint32_t* __restrict__ seg_count = nullptr,
                                               uint32_t (*seg_acc)[kMaxSegsPerTile] = nullptr)
{
    using P = ScanPlan<W>;
    const uint64_t n_blocks = (n + 31) >> 5;  // result words
    const uint64_t full_blocks = n >> 5;      // blocks whose 32 codes are all valid
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint32_t rot = rot_chunks<W>();
    uint32_t hits = 0;

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % P::kStages);
        const uint64_t tile_block0 = tile * (uint64_t)P::kBlocksPerTile;
        uint32_t* tile_out = bitvec + tile_block0;
        // (two rows of counters: the warps of iteration i + 1 add while row i & 1 is being flushed; row i & 1
        // is next used by iteration i + 2, behind the barrier of i + 1)
        uint32_t* tile_seg = kSeg ? seg_acc[iter & 1u] : nullptr;
        if (tile_block0 + P::kBlocksPerTile <= full_blocks) {
            hits += scan_tile<W, kMode, false, kSeg>(buf, args, rot, tile_out, 0, 0, tile_seg);
        } else {
            hits += scan_tile<W, kMode, true, kSeg>(buf, args, rot, tile_out, n_blocks - tile_block0, tail_bits, tile_seg);
        }
        ring.release(iter);  // (CTA barrier inside)
        if constexpr (kSeg) flush_seg_counts<P::kBlocksPerTile>(tile_seg, seg_count, tile_block0);
    }
    ring.finish();
    return hits;
}

template <int W, int kMode, bool kSeg>
__global__ void __launch_bounds__(kThreads, ScanPlan<W>::kCtasPerSm)
scan_range_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                  const __grid_constant__ ScanArgs<W> args, uint32_t* __restrict__ bitvec,
                  unsigned long long* __restrict__ count, unsigned long long* __restrict__ count_slot,
                  bool accumulate, unsigned long long* __restrict__ tickets, uint32_t* __restrict__ seg_count)
{
    using P = ScanPlan<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    typename P::Ring ring(smem, packed, readable_bytes16, tickets);
    uint32_t (*acc)[kMaxSegsPerTile] = nullptr;
    if constexpr (kSeg) {
        __shared__ uint32_t seg_acc[2][kMaxSegsPerTile];
        if (threadIdx.x < 2 * kMaxSegsPerTile) (&seg_acc[0][0])[threadIdx.x] = 0;
        __syncthreads();
        acc = seg_acc;
    }
    ring.start();
    const uint32_t hits = scan_tiles<W, kMode, kSeg>(ring, args, n, bitvec, seg_count, acc);
    if (count != nullptr) cta_finish_count(hits, count_slot, count, accumulate);
}

// One-bit codes without the ring: one CTA per tile, the tile read straight from global memory with
// 128-bit loads.  A 2^30-code scan at W = 1 is a ~50 us kernel whose output is as large as its input;
// there the ring's prologue and the persistent CTAs' drain are a visible share of the launch, and a
// plain grid of short-lived CTAs keeps more of the machine busy at both ends.
// Measured (profiles/r01_ab_scan_narrow.txt): +5 % at W = 1, nothing at W = 2, 3.
// End of round 2: the kernel is launched with programmatic stream serialization (pdl_enter, common.cuh), so
// back-to-back scans lose no launch gap: W = 1 0.886 -> 0.93 of the copy peak (46.3 -> 44.0 us), W = 2 -- now
// direct as well -- 0.932 -> 0.991 (66.0 -> 62.0 us).  The persistent ring kernel is launched the plain way: with
// the attribute it gained 1 % at some widths and lost 2-8 % at 3, 5, 11-13 and 17 bit, wherever its CTAs let
// the next kernel in (profiles/r02_pdl_ab.txt, r02_pdl_ab2.txt).

// The ring-less kernels' tile.  A thread owns W (or kG * W) consecutive words.  Read in place, the lanes of a
// warp load touch up to 32 different 128-byte lines (128-byte lane stride at 32 bit), and from about 9 bit on
// the kernel is bound by L1 wavefronts: 2^24 codes inside a CUDA graph took 9.3 / 12.8 / 19.1 us per scan at
// 17 / 23 / 32 bit.  So from 9 bit on the CTA fetches its tile with ONE bulk copy into shared memory and
// evaluates it there, conflict-free like the ring's stages (dynamic shared memory: kTileBytes + 16): 7.8 / 8.5 /
// 8.9 us, 2^22 codes at 32 bit 11.3 -> 4.8 us.  Below 9 bit the lanes' pieces share lines (1-2, 4, 8 bit: the
// loads are contiguous) and the copy's extra round trip costs more than it saves
// (profiles/r02_direct_staged_short.txt).
#ifndef SSS_DIRECT_STAGED_MINW
#define SSS_DIRECT_STAGED_MINW 9
#endif
// (11 bit, the one word-parallel width with scalar loads -- a 44-byte lane stride whose 11 loads re-use the same
// lines -- stays in place: BASELINE configs[0], 2^24 x 11 bit, measured 6.2 us per scan in place, 6.9 us staged)
template <int W>
constexpr bool kDirectStaged = (W >= SSS_DIRECT_STAGED_MINW) && (W != 11);
// The ring-less kernels' tile: the plan's at 1-2 bit (long columns: 32 words per thread do best); above -- short
// columns only -- SSS_DIRECT_ROUNDS rounds per thread (0 = the plan's): a column of a few hundred tiles is ONE
// wave of CTAs that all load, then all compute, then all store, and smaller tiles might give the waves something
// to overlap with.  Measured (profiles/r02_direct_rounds.txt): one round per thread helps only where the plan's
// tile leaves fewer than ~300 CTAs (2^24 codes at 3 / 4 bit: 6.6 -> 4.9 us) and costs 10-90 % elsewhere (2^26 codes
// at 4 / 8 bit: 7.3 -> 14.1 / 9.7 -> 15.1 us), so the plan's tile stays.
#ifndef SSS_DIRECT_ROUNDS
#define SSS_DIRECT_ROUNDS 0
#endif
template <int W>
struct DirectPlan {
    using P = ScanPlan<W>;
    static constexpr int kR = (W <= SSS_SCAN_DIRECT_MAXW || SSS_DIRECT_ROUNDS <= 0 || SSS_DIRECT_ROUNDS > P::kR) ? P::kR : SSS_DIRECT_ROUNDS;
    static constexpr int kBlocksPerTile = kThreads * kR * P::kG;
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
};
template <int W>
constexpr int kDirectSmemBytes = kDirectStaged<W> ? DirectPlan<W>::kTileBytes + 16 : 0;

template <int W>
__device__ __forceinline__ const uint32_t* direct_tile_words(const void* __restrict__ packed, uint64_t n, uint32_t& rot)
{
    using P = DirectPlan<W>;
    const uint64_t tile_off = blockIdx.x * (uint64_t)P::kTileBytes;
    const uint8_t* src = static_cast<const uint8_t*>(packed) + tile_off;
    if constexpr (!kDirectStaged<W>) {
        rot = 0;
        return reinterpret_cast<const uint32_t*>(src);
    } else {
        extern __shared__ __align__(128) uint8_t smem[];
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P::kTileBytes);
        const uint64_t readable = round_up_u64(((n + 31) >> 5) * (uint64_t)W * 4, 16);
        const uint64_t left = readable - tile_off;
        const uint32_t bytes = left < (uint64_t)P::kTileBytes ? (uint32_t)left : (uint32_t)P::kTileBytes;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s_plain(smem, src, bytes, bar);
        }
        mbar_wait(bar, 0);
        rot = rot_chunks<W>();
        return reinterpret_cast<const uint32_t*>(smem);
    }
}

template <int W, int kMode, bool kSeg>
__global__ void __launch_bounds__(kThreads)
scan_direct_kernel(const void* __restrict__ packed, uint64_t n, const __grid_constant__ ScanArgs<W> args,
                   uint32_t* __restrict__ bitvec, unsigned long long* __restrict__ count,
                   unsigned long long* __restrict__ count_slot, bool accumulate, uint32_t* __restrict__ seg_count)
{
    using P = DirectPlan<W>;
    pdl_enter();
    const uint64_t n_blocks = (n + 31) >> 5;
    const uint64_t full_blocks = n >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    const uint64_t tile_block0 = blockIdx.x * (uint64_t)P::kBlocksPerTile;
    uint32_t rot;
    const uint32_t* buf = direct_tile_words<W>(packed, n, rot);
    uint32_t* tile_out = bitvec + tile_block0;
    uint32_t* tile_seg = nullptr;
    if constexpr (kSeg) {
        __shared__ uint32_t seg_acc[kMaxSegsPerTile];
        if (threadIdx.x < kMaxSegsPerTile) seg_acc[threadIdx.x] = 0;
        __syncthreads();
        tile_seg = seg_acc;
    }
    uint32_t hits;
    if (tile_block0 + P::kBlocksPerTile <= full_blocks)
        hits = scan_tile<W, kMode, false, kSeg, P::kR>(buf, args, rot, tile_out, 0, 0, tile_seg);
    else
        hits = scan_tile<W, kMode, true, kSeg, P::kR>(buf, args, rot, tile_out, n_blocks - tile_block0, tail_bits, tile_seg);
    if constexpr (kSeg) {
        __syncthreads();
        flush_seg_counts<P::kBlocksPerTile>(tile_seg, seg_count, tile_block0);
    }
    if (count != nullptr) cta_finish_count(hits, count_slot, count, accumulate);
}

// The same scan with its bounds in DEVICE memory: a shared scan with a single predicate is this
// kernel (one bitvector, one count -- for K = 1 the linear layout is the bitvector too), and the
// shared-scan entry points take their bounds as device arrays.  The per-launch constants that
// launch_w() prepares on the host are built here by every thread (a few hundred instructions once
// per thread, against thousands of blocks scanned); an invalid or empty predicate writes zeros.
template <int W>
__global__ void __launch_bounds__(kThreads, ScanPlan<W>::kCtasPerSm)
scan_range_dev_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                      const uint32_t* __restrict__ d_lo, const uint32_t* __restrict__ d_hi,
                      uint32_t* __restrict__ bitvec, unsigned long long* __restrict__ count,
                      unsigned long long* __restrict__ count_slot, bool accumulate,
                      unsigned long long* __restrict__ tickets)
{
    using P = ScanPlan<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lo = __ldg(d_lo);
    uint32_t hi = __ldg(d_hi);
    uint32_t hits = 0;
    if (lo > hi || lo > code_mask(W)) {  // matches nothing (kernel-uniform)
        const uint64_t n_blocks = (n + 31) >> 5;
        for (uint64_t i = (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < n_blocks; i += (uint64_t)gridDim.x * kThreads)
            bitvec[i] = 0u;
    } else {
        if (hi > code_mask(W)) hi = code_mask(W);
        typename P::Ring ring(smem, packed, readable_bytes16, tickets);
        ring.start();
        ScanArgs<W> args{};
        if constexpr (W == 1) {
            args.onebit.flip = (lo == 0 && hi == 0) ? 0xFFFFFFFFu : 0u;
            args.onebit.all = (lo == 0 && hi == 1) ? 0xFFFFFFFFu : 0u;
            hits = scan_tiles<W, kOneBit>(ring, args, n, bitvec);
        } else if constexpr (W == 2) {
            args.bool2 = make_bool2(lo, hi);
            hits = scan_tiles<W, kBool2>(ring, args, n, bitvec);
        } else if constexpr (P::kSwar) {
            args.swar = make_swar_consts<W>(lo, hi);
            hits = args.swar.lo_top ? scan_tiles<W, kSwarLoTop1>(ring, args, n, bitvec)
                                    : scan_tiles<W, kSwarLoTop0>(ring, args, n, bitvec);
        } else {
            const AlignedRange r = aligned_range(W, lo, hi);
            args.scalar.lo_al = r.lo_al;
            args.scalar.span_al = r.span_al;
            hits = scan_tiles<W, kScalar>(ring, args, n, bitvec);
        }
    }
    if (count != nullptr) cta_finish_count(hits, count_slot, count, accumulate);
}

// ... and its ring-less form (one CTA per tile, programmatic launch) for one- and two-bit codes and for short
// columns, like scan_direct_kernel: a K = 1 shared scan over a short column otherwise pays the ring's prologue.
template <int W, int kMode>
__device__ __forceinline__ uint32_t scan_direct_tile(const void* __restrict__ packed, uint64_t n, const ScanArgs<W>& args,
                                                     uint32_t* __restrict__ bitvec)
{
    using P = DirectPlan<W>;
    const uint64_t n_blocks = (n + 31) >> 5;
    const uint64_t full_blocks = n >> 5;
    const uint64_t tile_block0 = blockIdx.x * (uint64_t)P::kBlocksPerTile;
    uint32_t rot;
    const uint32_t* buf = direct_tile_words<W>(packed, n, rot);
    if (tile_block0 + P::kBlocksPerTile <= full_blocks)
        return scan_tile<W, kMode, false, false, P::kR>(buf, args, rot, bitvec + tile_block0, 0, 0);
    return scan_tile<W, kMode, true, false, P::kR>(buf, args, rot, bitvec + tile_block0, n_blocks - tile_block0, (uint32_t)(n & 31));
}

template <int W>
__global__ void __launch_bounds__(kThreads)
scan_direct_dev_kernel(const void* __restrict__ packed, uint64_t n, const uint32_t* __restrict__ d_lo,
                       const uint32_t* __restrict__ d_hi, uint32_t* __restrict__ bitvec,
                       unsigned long long* __restrict__ count, unsigned long long* __restrict__ count_slot, bool accumulate)
{
    using P = ScanPlan<W>;
    pdl_enter();
    const uint32_t lo = __ldg(d_lo);
    uint32_t hi = __ldg(d_hi);
    uint32_t hits = 0;
    if (lo > hi || lo > code_mask(W)) {  // matches nothing (kernel-uniform): this tile's result words are zero
        const uint64_t n_blocks = (n + 31) >> 5;
        const uint64_t b0 = blockIdx.x * (uint64_t)DirectPlan<W>::kBlocksPerTile;
        for (uint32_t i = threadIdx.x; i < (uint32_t)DirectPlan<W>::kBlocksPerTile && b0 + i < n_blocks; i += kThreads) bitvec[b0 + i] = 0u;
    } else {
        if (hi > code_mask(W)) hi = code_mask(W);
        ScanArgs<W> args{};
        if constexpr (W == 1) {
            args.onebit.flip = (lo == 0 && hi == 0) ? 0xFFFFFFFFu : 0u;
            args.onebit.all = (lo == 0 && hi == 1) ? 0xFFFFFFFFu : 0u;
            hits = scan_direct_tile<W, kOneBit>(packed, n, args, bitvec);
        } else if constexpr (W == 2) {
            args.bool2 = make_bool2(lo, hi);
            hits = scan_direct_tile<W, kBool2>(packed, n, args, bitvec);
        } else if constexpr (P::kSwar) {
            args.swar = make_swar_consts<W>(lo, hi);
            hits = args.swar.lo_top ? scan_direct_tile<W, kSwarLoTop1>(packed, n, args, bitvec)
                                    : scan_direct_tile<W, kSwarLoTop0>(packed, n, args, bitvec);
        } else {
            const AlignedRange r = aligned_range(W, lo, hi);
            args.scalar.lo_al = r.lo_al;
            args.scalar.span_al = r.span_al;
            hits = scan_direct_tile<W, kScalar>(packed, n, args, bitvec);
        }
    }
    if (count != nullptr) cta_finish_count(hits, count_slot, count, accumulate);
}

// ---- launch ---------------------------------------------------------------------------------------

template <int W, int kMode, bool kSeg>
static cudaError_t launch_mode_seg(const Device& dev, const void* packed, uint64_t n, const ScanArgs<W>& args,
                                   uint32_t* bitvec, const CountTarget& ct, cudaStream_t stream)
{
    using P = ScanPlan<W>;
    constexpr int smem = P::Ring::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e =
            cudaFuncSetAttribute(scan_range_kernel<W, kMode, kSeg>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + P::kTileBytes - 1) / P::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * P::kCtasPerSm;
    // Short columns (BASELINE configs[0]: 2^24 codes = 1024 tiles, a 4 us kernel at the HBM roofline): the
    // persistent ring's prologue -- tickets, first bulk copies, barrier round trips -- is then most of the
    // launch, and a plain grid of one short-lived CTA per tile reading straight from global memory / L2
    // starts working at once (cfg 1 in a CUDA graph: 9.1 -> see profiles/r02_bench extras `cfg1`).
    const bool small = tiles <= (uint64_t)dev.scan_direct_tiles * max_ctas;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cfg.attrs = attr;
    if ((W <= SSS_SCAN_DIRECT_MAXW || small) && tiles < (1ull << 31)) {
        cfg.numAttrs = dev.pdl ? 1 : 0;  // (the ring kernel below: no attribute, see scan_direct_kernel)
        cfg.gridDim = dim3((unsigned)((readable + DirectPlan<W>::kTileBytes - 1) / DirectPlan<W>::kTileBytes));
        cfg.dynamicSmemBytes = kDirectSmemBytes<W>;
        return cudaLaunchKernelEx(&cfg, scan_direct_kernel<W, kMode, kSeg>, packed, n, args, bitvec, ct.result, ct.slot,
                                  ct.accumulate, ct.seg_count);
    }
    cfg.gridDim = dim3((unsigned)(tiles < max_ctas ? tiles : max_ctas));
    cfg.dynamicSmemBytes = smem;
    return cudaLaunchKernelEx(&cfg, scan_range_kernel<W, kMode, kSeg>, packed, readable, n, args, bitvec, ct.result, ct.slot,
                              ct.accumulate, dev.next_tickets(), ct.seg_count);
}

// ct.seg_count set: the variant that also fills the position list's segment counts (sss_scan_range_positions)
template <int W, int kMode>
static cudaError_t launch_mode(const Device& dev, const void* packed, uint64_t n, const ScanArgs<W>& args,
                               uint32_t* bitvec, const CountTarget& ct, cudaStream_t stream)
{
    if (ct.seg_count != nullptr) return launch_mode_seg<W, kMode, true>(dev, packed, n, args, bitvec, ct, stream);
    return launch_mode_seg<W, kMode, false>(dev, packed, n, args, bitvec, ct, stream);
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, uint32_t lo, uint32_t hi,
                            uint32_t* bitvec, const CountTarget& count, cudaStream_t stream)
{
    // caller guarantees lo <= hi <= 2^W - 1 (non-empty range, already clamped)
    ScanArgs<W> args{};
    if constexpr (W == 1) {
        args.onebit.flip = (lo == 0 && hi == 0) ? 0xFFFFFFFFu : 0u;
        args.onebit.all = (lo == 0 && hi == 1) ? 0xFFFFFFFFu : 0u;
        return launch_mode<W, kOneBit>(dev, packed, n, args, bitvec, count, stream);
    } else if constexpr (W == 2) {
        args.bool2 = make_bool2(lo, hi);
        return launch_mode<W, kBool2>(dev, packed, n, args, bitvec, count, stream);
    } else if constexpr (ScanPlan<W>::kSwar) {
        args.swar = make_swar_consts<W>(lo, hi);
        if (args.swar.lo_top) return launch_mode<W, kSwarLoTop1>(dev, packed, n, args, bitvec, count, stream);
        return launch_mode<W, kSwarLoTop0>(dev, packed, n, args, bitvec, count, stream);
    } else {
        const AlignedRange r = aligned_range(W, lo, hi);
        args.scalar.lo_al = r.lo_al;
        args.scalar.span_al = r.span_al;
        return launch_mode<W, kScalar>(dev, packed, n, args, bitvec, count, stream);
    }
}

template <int W>
static cudaError_t launch_w_dev(const Device& dev, const void* packed, uint64_t n, const uint32_t* d_lo,
                                const uint32_t* d_hi, uint32_t* bitvec, const CountTarget& ct, cudaStream_t stream)
{
    using P = ScanPlan<W>;
    constexpr int smem = P::Ring::kSmemBytes;
    static bool configured[64] = {};
    if (!configured[dev.ordinal & 63]) {
        cudaError_t e = cudaFuncSetAttribute(scan_range_dev_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = true;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + P::kTileBytes - 1) / P::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * P::kCtasPerSm;
    const bool small = tiles <= (uint64_t)dev.scan_direct_tiles * max_ctas;
    if ((W <= SSS_SCAN_DIRECT_MAXW || small) && tiles < (1ull << 31)) {
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((readable + DirectPlan<W>::kTileBytes - 1) / DirectPlan<W>::kTileBytes));
        cfg.blockDim = dim3(kThreads);
        cfg.stream = stream;
        cfg.attrs = attr;
        cfg.numAttrs = dev.pdl ? 1 : 0;
        cfg.dynamicSmemBytes = kDirectSmemBytes<W>;
        return cudaLaunchKernelEx(&cfg, scan_direct_dev_kernel<W>, packed, n, d_lo, d_hi, bitvec, ct.result, ct.slot,
                                  ct.accumulate);
    }
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    scan_range_dev_kernel<W><<<grid, kThreads, smem, stream>>>(packed, readable, n, d_lo, d_hi, bitvec, ct.result,
                                                               ct.slot, ct.accumulate, dev.next_tickets());
    return cudaGetLastError();
}

using ScanLauncher = cudaError_t (*)(const Device&, const void*, uint64_t, uint32_t, uint32_t, uint32_t*,
                                     const CountTarget&, cudaStream_t);
using ScanDevLauncher = cudaError_t (*)(const Device&, const void*, uint64_t, const uint32_t*, const uint32_t*,
                                        uint32_t*, const CountTarget&, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<ScanLauncher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}
template <int... Ws>
static constexpr auto make_dev_table(std::integer_sequence<int, Ws...>)
{
    return std::array<ScanDevLauncher, sizeof...(Ws)>{&launch_w_dev<Ws + 1>...};
}

cudaError_t launch_scan_range(const Device& dev, const void* packed, uint64_t n, int w, AlignedRange r,
                              uint32_t* bitvec, const CountTarget& count, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, 32>{});
    // recover the clamped code-domain bounds from the aligned form
    const int s = 32 - w;
    const uint32_t lo = r.lo_al >> s;
    const uint32_t hi = (r.lo_al + r.span_al) >> s;
    return table[w - 1](dev, packed, n, lo, hi, bitvec, count, stream);
}

cudaError_t launch_scan_range_dev(const Device& dev, const void* packed, uint64_t n, int w, const uint32_t* d_lo,
                                  const uint32_t* d_hi, uint32_t* bitvec, const CountTarget& count, cudaStream_t stream)
{
    static constexpr auto table = make_dev_table(std::make_integer_sequence<int, 32>{});
    return table[w - 1](dev, packed, n, d_lo, d_hi, bitvec, count, stream);
}

}  // namespace sss
