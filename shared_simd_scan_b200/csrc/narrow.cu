// narrow.cu -- shared scans with at most 8 predicates, both output layouts: NARROW codes (W <= 13) through
// a table indexed by groups of codes (below), wider codes (14..28 bit) through a bucket table whose
// entries carry the match masks (second half of the file).
//
// The index-driven kernels (shared.cu, linear.cu) spend ~12 instructions per code at K <= 8: extract
// the code, look up its interval, look up the interval's mask row.  That is right for wide codes, but
// a column of 1..6-bit codes has a roofline of 1.1-1.8 bytes per code, i.e. 3.6-5.8 T codes/s, which
// 12 instructions per code cannot reach (issue-bound at ~2 T codes/s; profiles/r01_ncu_v3_shared_w5_k8.txt).
// For narrow codes the answer does not need an interval at all: a group of G adjacent codes (G * W <=
// 12 bits; G = 4, 2, 1 for W <= 3, <= 6, <= 13) indexes a table that holds the G codes' 8-bit match
// masks in one 32-bit entry -- one lookup per G codes instead of two per code.  Every CTA derives that table in its prologue from the
// predicate index of the call (direct locator + mask rows, predindex.cuh) and keeps up to 32
// lane-private replicas of it in shared memory (no bank conflicts).  The entries' bytes are brought
// into the 32 x 8 transpose's input order with 16 PRMTs; from there on it is the K <= 8 path of the
// other kernels: three bit stages, popcount for the counts, per-predicate result words (bitvectors) or
// the block's 4K bytes in one store (linear layout).
//
// Replaces, for this corner, the same reference functions as shared.cu / linear.cu
// (src/simd_scan_shared.cpp:34-151, src/simd_scan_shared_linear.cpp:9-82).
// Roofline: HBM.  Algorithmic bytes per code = W/8 + K/8.
#include <array>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "predindex.cuh"

namespace sss {

constexpr int kNThreads = 512;
constexpr int kNarrowTableMaxW = 15;                 // widest one-code table (32 KB of mask bytes per replica)
constexpr uint64_t kNarrowTableMinN = 1ull << 24;    // ... used from this many codes on at 14 / 15 bit
#ifndef SSS_NARROW_CTAS
#define SSS_NARROW_CTAS 1  // CTAs per SM (2 measured no better: the kernel is bound by its 8 write streams)
#endif
constexpr int kNCtasPerSm = SSS_NARROW_CTAS;
// (measured: with at most 8 output streams the stores are not what limits these kernels, and the detour
// through the stage costs 2-8 %: profiles/r02_ab_k8_wide.txt.  Off; -DSSS_NARROW_WIDE=1 rebuilds it.)
#ifndef SSS_NARROW_WIDE
#define SSS_NARROW_WIDE 0
#endif
// the warps' wide-store stages (16 KB) -- shared memory the table replicas get while the stage is off
constexpr int kNStageBytes = SSS_NARROW_WIDE ? (kNThreads / 32) * kSmallStageBytesPerWarp : 0;
static bool narrow_wide_ok(const void* out, uint64_t stride_words, bool linear)
{
    return SSS_NARROW_WIDE && !linear && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (stride_words & 3) == 0;
}
static_assert(kNarrowMaxK == kDirectCountMaxK, "the narrow path counts by popcount");

template <int W>
struct NarrowCfg {
    static constexpr int kG = W <= 3 ? 4 : (W <= 6 ? 2 : 1);  // codes per lookup
    static constexpr int kBits = kG * W;       // index bits
    static constexpr int kEntries = 1 << kBits;
    static constexpr int kLookups = 32 / kG;
    // Byte tables (W >= kByteMinW, one code per lookup).  An entry of a one-code table is a single mask
    // byte; kept as a 32-bit word it wastes three quarters of the shared memory that decides how many
    // lane-private replicas fit -- and the replicas decide the bank conflicts of 32 random lookups
    // (12 bit: 4 replicas, 2.6 wavefronts per lookup, 0.69 of the HBM peak; with three ring stages 8
    // replicas and 0.73).  Packed four codes to the word the same memory holds four times the replicas
    // (12 bit: 32, conflict-free; 13 bit: 16).  Replicas stay interleaved word by word, so a code's
    // byte sits at ((code >> 2) * R + replica) * 4 + (code & 3): two more integer instructions per code.
#ifndef SSS_NARROW_BYTE_MIN_W
#define SSS_NARROW_BYTE_MIN_W 12
#endif
    static constexpr bool kByte = kG == 1 && W >= SSS_NARROW_BYTE_MIN_W;
    static constexpr int kTableBytes = kByte ? kEntries : kEntries * 4;  // per replica
    // blocks per thread per tile: a block of narrow codes is only W words, and a tile has to be large
    // enough (>= 16 KB) for the ring's per-tile bookkeeping to disappear behind the work
    static constexpr int kBT = 8 / W > 1 ? 8 / W : 1;
    static constexpr int kBlocksPerTile = kNThreads * kBT;
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
#ifndef SSS_NARROW_STAGES_WIDE
#define SSS_NARROW_STAGES_WIDE 3  // 11-13 bit: a shallower ring buys a doubling of the table replicas (11 bit 0.69 -> 0.82)
#endif
    static constexpr int kStages = W >= 15 ? 2 : (W >= 11 ? SSS_NARROW_STAGES_WIDE : 4);  // (15 bit: 4 replicas of 32 KB)
    using Ring = TileRing<kTileBytes, kStages, 16, /*kDynamic=*/false>;  // static tile order: see common.cuh
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
    static_assert(kBits <= kNarrowTableMaxW, "group index must stay small enough to replicate");
};

inline int narrow_rep_log2(size_t ring_bytes, int table_bytes)
{
    int r = 5;
    while (r > 0 && ring_bytes + ((size_t)table_bytes << r) + kNStageBytes + 64 > (size_t)(224 / kNCtasPerSm - 4) * 1024) r--;
    return r;
}

// group g of the block: G adjacent codes = kBits bits at bit offset g * kBits of the W words
template <int W, int G>
__device__ __forceinline__ uint32_t group_index(const uint32_t (&x)[W])
{
    using Cfg = NarrowCfg<W>;
    constexpr int bit = G * Cfg::kBits, wi = bit >> 5, sh = bit & 31;
    constexpr uint32_t mask = (1u << Cfg::kBits) - 1u;
    if constexpr (sh + Cfg::kBits <= 32) {
        return (x[wi] >> sh) & mask;
    } else {
        return __funnelshift_r(x[wi], x[wi + 1 < W ? wi + 1 : wi], sh) & mask;
    }
}

template <int W, int G>
struct GroupLookups {
    static __device__ __forceinline__ void run(const uint32_t (&x)[W], uint32_t base, uint32_t stride,
                                               uint32_t (&e)[NarrowCfg<W>::kLookups])
    {
        if constexpr (NarrowCfg<W>::kByte) {
            // byte of code v: word (v >> 2) of the lane's replica, byte v & 3.  v * stride / 4 puts the word
            // offset right and drags the low two bits along, shifted: take those out, add them back unshifted
            // (`unshift` = stride / 4 - 1).
            const uint32_t v = group_index<W, G>(x);
            asm("ld.shared.u8 %0, [%1];" : "=r"(e[G]) : "r"(v * (stride >> 2) + base - (v & 3u) * ((stride >> 2) - 1u)));
        } else {
            asm("ld.shared.u32 %0, [%1];" : "=r"(e[G]) : "r"(group_index<W, G>(x) * stride + base));
        }
        GroupLookups<W, G + 1>::run(x, base, stride, e);
    }
};
template <int W>
struct GroupLookups<W, NarrowCfg<W>::kLookups> {
    static __device__ __forceinline__ void run(const uint32_t (&)[W], uint32_t, uint32_t,
                                               uint32_t (&)[NarrowCfg<W>::kLookups])
    {
    }
};

template <int W, bool kAll8>
__global__ void __launch_bounds__(kNThreads, kNCtasPerSm)
shared_scan_narrow_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                          const uint8_t* __restrict__ g_index, int K_arg, void* __restrict__ out, uint64_t stride_words,
                          bool linear, unsigned long long* __restrict__ counts, int rep_log2, bool wide,
                          unsigned long long* __restrict__ tickets)
{
    // kAll8: exactly 8 predicates (no per-predicate guards in the epilogue)
    const int K = kAll8 ? 8 : K_arg;
    using Cfg = NarrowCfg<W>;
    extern __shared__ __align__(128) uint8_t smem[];
    // ---- prologue: the group table, `R` replicas interleaved entry by entry --------------------------
    if constexpr (W > kDirectMaxW) {
        // 14, 15 bit: the call's index locates codes through buckets, not directly.  With at most 8 predicates
        // there are at most 16 breakpoints, so the byte table is filled from the index's thresholds: the mask
        // of code c is the row of interval #{breakpoints <= c}.  One replica is computed (16 compares per
        // code, 8-16 words per thread), the others are copies of it.
        static_assert(Cfg::kByte, "wide one-code tables are byte tables");
        __shared__ uint32_t s_thr[2 * kNarrowMaxK], s_mask[2 * kNarrowMaxK + 1];
        const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
        const uint32_t* thr = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kThr);
        const uint32_t* masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
        const uint32_t num_bp = h->num_bp, row_words = h->row_words;
        if (threadIdx.x < 2 * kNarrowMaxK) s_thr[threadIdx.x] = threadIdx.x < num_bp ? thr[threadIdx.x] : 0xFFFFFFFFu;
        if (threadIdx.x <= 2 * kNarrowMaxK) s_mask[threadIdx.x] = threadIdx.x <= num_bp ? masks[threadIdx.x * row_words] & 0xFFu : 0u;
        __syncthreads();
        uint32_t* table = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes);
        for (uint32_t v = threadIdx.x; v < (uint32_t)(Cfg::kTableBytes / 4); v += kNThreads) {
            uint32_t entry = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t t = (4 * v + c) << (32 - W);  // left-aligned, like the thresholds
                uint32_t id = 0;
#pragma unroll
                for (int k = 0; k < 2 * kNarrowMaxK; k++) id += t > s_thr[k] ? 1u : 0u;
                entry |= s_mask[id] << (8 * c);
            }
            table[v << rep_log2] = entry;  // replica 0
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < ((uint32_t)(Cfg::kTableBytes / 4) << rep_log2); i += kNThreads)
            if (i & ((1u << rep_log2) - 1u)) table[i] = table[i & ~((1u << rep_log2) - 1u)];
    } else {
        const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
        const uint16_t* lut = reinterpret_cast<const uint16_t*>(g_index + PredIndexLayout::kLocator);  // W <= 13: direct
        const uint32_t* masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
        const uint32_t row_words = h->row_words;
        uint32_t* table = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes);
        for (uint32_t i = threadIdx.x; i < ((uint32_t)(Cfg::kTableBytes / 4) << rep_log2); i += kNThreads) {
            const uint32_t v = i >> rep_log2;
            uint32_t entry = 0;
            if constexpr (Cfg::kByte) {  // word v of a replica = the mask bytes of codes 4v .. 4v+3
#pragma unroll
                for (int c = 0; c < 4; c++) entry |= (masks[(uint32_t)lut[4 * v + c] * row_words] & 0xFFu) << (8 * c);
            } else {
#pragma unroll
                for (int c = 0; c < Cfg::kG; c++) {
                    const uint32_t code = (v >> (c * W)) & code_mask(W);
                    entry |= (masks[(uint32_t)lut[code] * row_words] & 0xFFu) << (8 * c);
                }
            }
            table[i] = entry;
        }
    }
    const uint32_t R = 1u << rep_log2;
    // the warp's wide-store stage sits behind the table
    uint32_t* const wstage = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes + ((size_t)Cfg::kTableBytes << rep_log2)) +
                             (threadIdx.x >> 5) * (kSmallStageBytesPerWarp / 4);
    const uint32_t table_base = smem_u32(smem + Cfg::kRingBytes) + ((threadIdx.x & (R - 1)) << 2);
    const uint32_t table_stride = 4u << rep_log2;
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();  // contains the __syncthreads that publishes the table

    using Rot = BlockRot<W, 1>;
    const uint32_t lane_rot = Rot::lane_rot();
    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    uint32_t direct[kNarrowMaxK] = {};

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        // thread t owns blocks t, t + 512, ... of the tile (lanes stay adjacent: conflict-free loads,
        // coalesced stores)
        uint32_t x[Cfg::kBT][W];
        uint32_t loaded = 0;
#pragma unroll
        for (int u = 0; u < Cfg::kBT; u++) {
            // (W = 8 reads its block rotated by lane_rot pieces, see BlockRot; undone on the result words.
            //  The column's ragged last block may come out rotated too: its `keep` mask is applied after.)
            load_block<W, 1>(buf, u * kNThreads + threadIdx.x, x[u], lane_rot);
#pragma unroll
            for (int k = 0; k < W; k++) loaded |= x[u][k];
        }
        ring.release_warp(iter, loaded);
#pragma unroll
        for (int u = 0; u < Cfg::kBT; u++) {
            const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + u * kNThreads + threadIdx.x;
            if (gb >= n_blocks) break;
            const uint32_t keep = (tail_bits != 0 && gb == n_blocks - 1) ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;

            uint32_t e[Cfg::kLookups];
            GroupLookups<W, 0>::run(x[u], table_base, table_stride, e);
            // entries -> r[q] byte b = the predicate bits of code 8b + q (the 32 x 8 transpose's input order)
            uint32_t r[8];
            if constexpr (Cfg::kG == 4) {  // e[g] bytes = codes 4g .. 4g+3
                uint32_t o[4];
                byte_transpose4(e[0], e[2], e[4], e[6], o);
                r[0] = o[0]; r[1] = o[1]; r[2] = o[2]; r[3] = o[3];
                byte_transpose4(e[1], e[3], e[5], e[7], o);
                r[4] = o[0]; r[5] = o[1]; r[6] = o[2]; r[7] = o[3];
            } else if constexpr (Cfg::kG == 1) {  // e[j] = the mask of code j: the plain 32 x 8 transpose
                transpose32x8(e, r);
            } else {  // e[g] bytes 0, 1 = codes 2g, 2g+1
#pragma unroll
                for (int hq = 0; hq < 4; hq++) {
                    const uint32_t t0 = __byte_perm(e[hq], e[4 + hq], 0x5140);
                    const uint32_t t1 = __byte_perm(e[8 + hq], e[12 + hq], 0x5140);
                    r[2 * hq] = __byte_perm(t0, t1, 0x5410);
                    r[2 * hq + 1] = __byte_perm(t0, t1, 0x7632);
                }
            }
            if constexpr (Cfg::kG != 1) transpose_bits_32x8(r);
            if constexpr (Rot::kRot) {
#pragma unroll
                for (int k = 0; k < 8; k++) r[k] = __funnelshift_l(r[k], r[k], lane_rot * Rot::kPieceCodes);
            }
            if (keep != 0xFFFFFFFFu) {  // the column's ragged last block only
#pragma unroll
                for (int k = 0; k < 8; k++) r[k] &= keep;
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (kAll8 || k < K) direct[k] += __popc(r[k]);
            const uint64_t gb_warp0 = gb - (threadIdx.x & 31);
            if (linear) {
                store_block_small(static_cast<uint8_t*>(out) + gb * 4 * (uint64_t)K, r, K);
            } else if (SSS_NARROW_WIDE && wide && gb_warp0 + 32 <= n_blocks) {  // warp-uniform: every lane has a block
                store_words8_wide(wstage, r, K, static_cast<uint32_t*>(out) + gb_warp0, stride_words);
            } else {
                uint32_t* p = static_cast<uint32_t*>(out) + gb;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (kAll8 || k < K) st_stream_u32(p, r[k]);
                    p += stride_words;
                }
            }
        }
    }
    ring.finish();
    if (counts != nullptr) {
#pragma unroll
        for (int k = 0; k < kNarrowMaxK; k++) {
            const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, direct[k]);
            if ((threadIdx.x & 31) == 0 && k < K && s != 0) atomicAdd(&counts[k], (unsigned long long)s);
        }
    }
}

// ---- up to 8 predicates over WIDER codes (14 <= W <= kFewMaxW): the bucket word carries the masks ----
//
// From 14 bits on a code is located through the 8192-bucket locator, and the index-driven kernels then
// read the interval's mask row: two dependent, randomly addressed shared-memory loads and ~18
// instructions per code (profiles/r01_ncu_v3_shared_w17_k8.txt: shared-memory pipe 79 %, issue 71 %,
// 0.64 of the HBM peak).  With at most 8 predicates the two mask bytes a bucket can resolve to fit into
// the bucket word itself:
//
//     entry = thr15 << 17 | several << 16 | mask_after << 8 | mask_before
//
// thr15 = the top 15 bits of the standard word's 19-bit threshold (predindex.cuh; exact while the
// code's bits below the bucket bits number at most 15, i.e. W <= 28).  The code is extracted RIGHT-
// aligned (v: code in the low W bits, later codes above), so `v << (45 - W)` is the threshold field
// with zeros below it -- `(v << (45 - W)) > entry` is the compare of the standard lookup, and the
// entry's low bits cannot turn a tie into "greater".  The table address is one rotate + one LOP3
// (`(rot(v) & index bits) | lane replica`).  Every CTA derives the table in its prologue from the
// call's standard index.  Codes in a `several` bucket (two or more inner breakpoints: every equality
// predicate) are patched from the global index by the blocks that met one.
constexpr int kFewMaxW = kFewPredicatesMaxW;
constexpr uint32_t kFewSeveral = 1u << 16;

template <int W>
struct FewCfg {
    static constexpr int kBlocksPerTile = kNThreads;  // one block per thread per tile
    static constexpr int kTileBytes = kBlocksPerTile * W * 4;
#ifndef SSS_FEW_TWO_STAGE_KB
#define SSS_FEW_TWO_STAGE_KB 32
#endif
    static constexpr int kStages = kTileBytes >= SSS_FEW_TWO_STAGE_KB * 1024 ? 2 : 3;
    using Ring = TileRing<kTileBytes, kStages, 16, /*kDynamic=*/false>;  // static tile order: see common.cuh
    static constexpr int kRingBytes = (Ring::kSmemBytes + 15) & ~15;
    // lane-private replicas of the table from what the ring leaves over (compile-time: the index shift
    // below then folds into the code's extraction shift)
    static constexpr int rep_log2()
    {
        int r = 5;
        while (r > 0 && kRingBytes + ((kBuckets * 4) << r) + kNStageBytes + 64 > 220 * 1024) r--;
        return r;
    }
    static constexpr int kRepLog2 = rep_log2();
    static constexpr int kTableBytes = (kBuckets * 4) << kRepLog2;
    static constexpr uint32_t kIdxMask = (uint32_t)(kBuckets - 1) << (2 + kRepLog2);
    // right shift that moves a RIGHT-aligned code's top 13 bits onto kIdxMask (negative: left shift)
    static constexpr int kIdxShift = W - kBucketBits - 2 - kRepLog2;
};

// x >> S for S >= 0, x << -S otherwise
template <int S>
__device__ __forceinline__ uint32_t shift_right_signed(uint32_t x)
{
    if constexpr (S >= 0) return x >> S;
    else return x << (-S);
}

// Pipe balance.  Integer work issues on two pipes, each good for one warp instruction every two
// cycles per scheduler: shifts, LOP3, PRMT and compares on the ALU pipe, IMAD on the FMA pipe.  A
// lookup written with shifts alone is bound by the ALU pipe (ncu, 17 bit: ALU 86 %, FMA 8 %), so half
// of the shifts are issued as multiplies: `x << s` is IMAD.SHL anyway, and `x >> s` = umulhi(x, 2^(32-s))
// becomes IMAD.HI when the power of two comes from a kernel parameter (an immediate would be folded
// back into a shift by ptxas).
struct Pow2Table {
    uint32_t p[32];  // p[k] = 1 << k
};

template <int S>
__device__ __forceinline__ uint32_t shr_fma(uint32_t x, const Pow2Table& P)
{
    if constexpr (S == 0) return x;
    else return __umulhi(x, P.p[32 - S]);
}

// Code J of the block: shared-memory address of its table entry and the threshold field
// `v << (45 - W)` (v = the right-aligned code, later codes above it).
template <int W, int J>
__device__ __forceinline__ void few_code(const uint32_t (&x)[W], const Pow2Table& P, uint32_t table_base,
                                         uint32_t lane_off, uint32_t& addr, uint32_t& field)
{
    using Cfg = FewCfg<W>;
    constexpr int bit = J * W, wi = bit >> 5, sh = bit & 31;
    constexpr bool kFmaRoute = (J % 3) != 0;  // address arithmetic on the FMA pipe for two codes out of three
    uint32_t v;
    if constexpr (sh + W <= 32) v = kFmaRoute ? x[wi] >> sh : shr_fma<sh>(x[wi], P);
    else v = __funnelshift_r(x[wi], x[wi + 1 < W ? wi + 1 : wi], sh);
    field = v << (45 - W);
    if constexpr (kFmaRoute) {
        const uint32_t idx = __umulhi(v << (32 - W), P.p[kBucketBits]);  // top 13 bits of the code
        addr = idx * (4u << Cfg::kRepLog2) + (table_base + lane_off);
    } else {
        addr = bit_select(shift_right_signed<Cfg::kIdxShift>(v), lane_off, Cfg::kIdxMask) + table_base;
    }
}

template <int W, int J>
struct FewLookups {
    // two codes per step: their entries join the block's `several` flags in ONE three-input LOP3
    static __device__ __forceinline__ void run(const uint32_t (&x)[W], const Pow2Table& P, uint32_t table_base,
                                               uint32_t lane_off, uint32_t (&m)[32], uint32_t& flags)
    {
        uint32_t addr0, field0, addr1, field1, e0, e1;
        few_code<W, J>(x, P, table_base, lane_off, addr0, field0);
        few_code<W, J + 1>(x, P, table_base, lane_off, addr1, field1);
        asm("ld.shared.u32 %0, [%1];" : "=r"(e0) : "r"(addr0));
        asm("ld.shared.u32 %0, [%1];" : "=r"(e1) : "r"(addr1));
        asm("lop3.b32 %0, %0, %1, %2, 0xFE;" : "+r"(flags) : "r"(e0), "r"(e1));
        m[J] = field0 > e0 ? __umulhi(e0, P.p[24]) : e0;  // (e >> 8 as IMAD.HI); byte 0 = the code's match mask
        m[J + 1] = field1 > e1 ? __umulhi(e1, P.p[24]) : e1;
        FewLookups<W, J + 2>::run(x, P, table_base, lane_off, m, flags);
    }
    static __device__ __forceinline__ void fix(const uint32_t (&x)[W], const uint8_t* __restrict__ g_index,
                                               uint32_t row_words, uint32_t (&m)[32])
    {
        constexpr int bit = J * W, wi = bit >> 5, sh = bit & 31;
        uint32_t v = x[wi] >> sh;
        if constexpr (sh + W > 32) v = __funnelshift_r(x[wi], x[wi + 1 < W ? wi + 1 : wi], sh);
        const uint32_t t = v << (32 - W);  // left-aligned, zeros below
        const uint32_t e = __ldg(reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kLocator) +
                                 (t >> (32 - kBucketBits)));
        if (e & kSeveralBit) {
            const uint32_t* thr = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kThr);
            uint32_t i = e & kBucketIdMask;
            while (t > __ldg(thr + i)) i++;  // thr[num_bp] = 0xFFFFFFFF ends the walk
            m[J] = __ldg(reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks) + i * row_words);
        }
        FewLookups<W, J + 1>::fix(x, g_index, row_words, m);
    }
};
template <int W>
struct FewLookups<W, 32> {
    static __device__ __forceinline__ void run(const uint32_t (&)[W], const Pow2Table&, uint32_t, uint32_t,
                                               uint32_t (&)[32], uint32_t&)
    {
    }
    static __device__ __forceinline__ void fix(const uint32_t (&)[W], const uint8_t*, uint32_t, uint32_t (&)[32]) {}
};

template <int W, bool kAll8>
__global__ void __launch_bounds__(kNThreads, 1)
shared_scan_few_kernel(const void* __restrict__ packed, uint64_t readable_bytes16, uint64_t n,
                       const uint8_t* __restrict__ g_index, int K_arg, void* __restrict__ out, uint64_t stride_words,
                       bool linear, unsigned long long* __restrict__ counts, const Pow2Table P, bool wide,
                       unsigned long long* __restrict__ tickets)
{
    // kAll8: exactly 8 predicates (no per-predicate guards in the epilogue)
    const int K = kAll8 ? 8 : K_arg;
    using Cfg = FewCfg<W>;
    constexpr int rep_log2 = Cfg::kRepLog2;
    static_assert(W > kDirectMaxW && W <= kFewMaxW && W - kBucketBits <= 15, "bucket locator widths, 15-bit thresholds");
    extern __shared__ __align__(128) uint8_t smem[];
    const PredIndexHeader* h = reinterpret_cast<const PredIndexHeader*>(g_index + PredIndexLayout::kHeader);
    const uint32_t row_words = h->row_words;
    // ---- prologue: the mask-carrying bucket table, `R` replicas interleaved entry by entry ------------
    {
        const uint32_t num_bp = h->num_bp;
        const uint32_t* bucket = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kLocator);
        const uint32_t* masks = reinterpret_cast<const uint32_t*>(g_index + PredIndexLayout::kMasks);
        uint32_t* table = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes);
        for (uint32_t i = threadIdx.x; i < ((uint32_t)kBuckets << rep_log2); i += kNThreads) {
            const uint32_t e = bucket[i >> rep_log2];
            const uint32_t id = e & kBucketIdMask;
            const uint32_t before = masks[id * row_words] & 0xFFu;
            const uint32_t after = id < num_bp ? masks[(id + 1) * row_words] & 0xFFu : 0u;
            table[i] = (e & 0xFFFE0000u) | ((e & kSeveralBit) ? kFewSeveral : 0u) | (after << 8) | before;
        }
    }
    const uint32_t table_base = smem_u32(smem + Cfg::kRingBytes);
    uint32_t* const wstage = reinterpret_cast<uint32_t*>(smem + Cfg::kRingBytes + Cfg::kTableBytes) +
                             (threadIdx.x >> 5) * (kSmallStageBytesPerWarp / 4);
    const uint32_t lane_off = (threadIdx.x & ((1u << rep_log2) - 1u)) << 2;
    typename Cfg::Ring ring(smem, packed, readable_bytes16, tickets);
    ring.start();  // contains the __syncthreads that publishes the table

    using Rot = BlockRot<W, 1>;
    const uint32_t lane_rot = Rot::lane_rot();
    const uint64_t n_blocks = (n + 31) >> 5;
    const uint32_t tail_bits = (uint32_t)(n & 31);
    uint32_t direct[kNarrowMaxK] = {};

    for (uint32_t iter = 0;; ++iter) {
        const uint64_t tile = ring.acquire(iter);
        if (tile == kNoTile) break;
        const uint32_t* buf = ring.stage_words(iter % Cfg::kStages);
        uint32_t x[W];
        load_block<W, 1>(buf, threadIdx.x, x, lane_rot);  // (W = 16, 24: rotated by lane_rot pieces, see BlockRot)
        uint32_t loaded = 0;
#pragma unroll
        for (int k = 0; k < W; k++) loaded |= x[k];
        ring.release_warp(iter, loaded);
        const uint64_t gb = tile * (uint64_t)Cfg::kBlocksPerTile + threadIdx.x;
        if (gb >= n_blocks) continue;
        const uint32_t keep = (tail_bits != 0 && gb == n_blocks - 1) ? ((1u << tail_bits) - 1u) : 0xFFFFFFFFu;

        uint32_t m[32], flags = 0;
        FewLookups<W, 0>::run(x, P, table_base, lane_off, m, flags);
        if (flags & kFewSeveral) FewLookups<W, 0>::fix(x, g_index, row_words, m);
        uint32_t r[8];
        transpose32x8(m, r);
        if constexpr (Rot::kRot) {
#pragma unroll
            for (int k = 0; k < 8; k++) r[k] = __funnelshift_l(r[k], r[k], lane_rot * Rot::kPieceCodes);
        }
        if (keep != 0xFFFFFFFFu) {  // the column's ragged last block only
#pragma unroll
            for (int k = 0; k < 8; k++) r[k] &= keep;
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (kAll8 || k < K) direct[k] += __popc(r[k]);
        const uint64_t gb_warp0 = gb - (threadIdx.x & 31);
        if (linear) {
            store_block_small(static_cast<uint8_t*>(out) + gb * 4 * (uint64_t)K, r, K);
        } else if (SSS_NARROW_WIDE && wide && gb_warp0 + 32 <= n_blocks) {  // warp-uniform: every lane has a block
            store_words8_wide(wstage, r, K, static_cast<uint32_t*>(out) + gb_warp0, stride_words);
        } else {
            uint32_t* p = static_cast<uint32_t*>(out) + gb;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (kAll8 || k < K) st_stream_u32(p, r[k]);
                p += stride_words;
            }
        }
    }
    ring.finish();
    if (counts != nullptr) {
#pragma unroll
        for (int k = 0; k < kNarrowMaxK; k++) {
            const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, direct[k]);
            if ((threadIdx.x & 31) == 0 && k < K && s != 0) atomicAdd(&counts[k], (unsigned long long)s);
        }
    }
}

template <int W>
static cudaError_t launch_few_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K, void* out,
                                uint64_t stride_words, bool linear, unsigned long long* counts, cudaStream_t stream)
{
    using Cfg = FewCfg<W>;
    const size_t smem = Cfg::kRingBytes + Cfg::kTableBytes + kNStageBytes + 64;
    const bool wide = narrow_wide_ok(out, stride_words, linear);
    Pow2Table P;
    for (int k = 0; k < 32; k++) P.p[k] = 1u << k;
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_few_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(shared_scan_few_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    if (K == 8)
        shared_scan_few_kernel<W, true><<<grid, kNThreads, smem, stream>>>(
            packed, readable, n, static_cast<const uint8_t*>(index), K, out, stride_words, linear, counts, P, wide,
            dev.next_tickets());
    else
        shared_scan_few_kernel<W, false><<<grid, kNThreads, smem, stream>>>(
            packed, readable, n, static_cast<const uint8_t*>(index), K, out, stride_words, linear, counts, P, wide,
            dev.next_tickets());
    return cudaGetLastError();
}

template <int W>
static cudaError_t launch_w(const Device& dev, const void* packed, uint64_t n, const void* index, int K, void* out,
                            uint64_t stride_words, bool linear, unsigned long long* counts, cudaStream_t stream)
{
    using Cfg = NarrowCfg<W>;
    const int r = narrow_rep_log2(Cfg::kRingBytes, Cfg::kTableBytes);
    const size_t smem = Cfg::kRingBytes + ((size_t)Cfg::kTableBytes << r) + kNStageBytes + 64;
    const bool wide = narrow_wide_ok(out, stride_words, linear);
    static size_t configured[64] = {};
    if (configured[dev.ordinal & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(shared_scan_narrow_kernel<W, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(shared_scan_narrow_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev.ordinal & 63] = smem;
    }
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t readable = round_up_u64(n_blocks * (uint64_t)W * 4, 16);
    const uint64_t tiles = (readable + Cfg::kTileBytes - 1) / Cfg::kTileBytes;
    const uint64_t max_ctas = (uint64_t)dev.sm_count * kNCtasPerSm;
    const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
    if (K == 8)
        shared_scan_narrow_kernel<W, true><<<grid, kNThreads, smem, stream>>>(
            packed, readable, n, static_cast<const uint8_t*>(index), K, out, stride_words, linear, counts, r, wide,
            dev.next_tickets());
    else
        shared_scan_narrow_kernel<W, false><<<grid, kNThreads, smem, stream>>>(
            packed, readable, n, static_cast<const uint8_t*>(index), K, out, stride_words, linear, counts, r, wide,
            dev.next_tickets());
    return cudaGetLastError();
}

using Launcher = cudaError_t (*)(const Device&, const void*, uint64_t, const void*, int, void*, uint64_t, bool,
                                 unsigned long long*, cudaStream_t);

template <int... Ws>
static constexpr auto make_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_w<Ws + 1>...};
}
template <int... Ws>
static constexpr auto make_few_table(std::integer_sequence<int, Ws...>)
{
    return std::array<Launcher, sizeof...(Ws)>{&launch_few_w<Ws + kNarrowMaxW + 1>...};
}

cudaError_t launch_shared_scan_narrow(const Device& dev, const void* packed, uint64_t n, int w, const void* index,
                                      int K, void* out, uint64_t stride_words, bool linear,
                                      unsigned long long* counts, cudaStream_t stream)
{
    static constexpr auto table = make_table(std::make_integer_sequence<int, kNarrowMaxW>{});
    static constexpr auto few = make_few_table(std::make_integer_sequence<int, kFewMaxW - kNarrowMaxW>{});
    // 14 and 15 bit: the one-code byte table (ten instructions per code instead of the bucket table's 16.5)
    // once the column is long enough for its per-CTA table fill (~3 us) not to show
    if (w == 14 && n >= kNarrowTableMinN)
        return launch_w<14>(dev, packed, n, index, K, out, stride_words, linear, counts, stream);
    if (w == 15 && n >= kNarrowTableMinN)
        return launch_w<15>(dev, packed, n, index, K, out, stride_words, linear, counts, stream);
    if (w > kNarrowMaxW)
        return few[w - kNarrowMaxW - 1](dev, packed, n, index, K, out, stride_words, linear, counts, stream);
    return table[w - 1](dev, packed, n, index, K, out, stride_words, linear, counts, stream);
}

}  // namespace sss
