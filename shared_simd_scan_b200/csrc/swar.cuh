// swar.cuh -- word-parallel (SWAR) range predicate on packed codes, for narrow bit widths.
//
// At narrow widths the fused scan is bound by integer issue, not by HBM: a per-code pipeline
// (extract, subtract, compare, insert bit) costs ~4 ALU instructions per code, and a B200 SM issues
// 64 integer lane-ops per clock, so one code per 4 ops tops out near 4 T codes/s while the HBM
// roofline at w = 4 is > 10 T codes/s.  The answer is the one the reference gives for its own
// machine -- "compare in place, do not unpack" (scan_128 compares masked lanes against a
// pre-shifted key, reference src/simd_scan.cpp:126-130) -- taken to word granularity: all the codes
// of a 32-bit word are compared at once, no matter how they straddle word boundaries.
//
// For a 32-code block X (W words, fields of W bits, NO spare guard bits) and lo <= hi:
//   H = the top bit of every field, L = ~H, d = hi - lo
//   T = (X | H) - rep(lo & ~msb)          field-wise x_low - lo_low; forcing the top bit makes every
//                                         field non-negative, so borrows never cross fields and one
//                                         multi-word subtract-with-borrow chain does all 32 fields
//   u = x - lo (mod 2^W):  u_low = T_low,  u_top = T_top ^ x_top ^ ~lo_top
//   G = rep(d | msb) - (T & L)            field-wise d_low - u_low, same trick
//   in range  <=>  d >= u  <=>  (d_top & ~u_top) | (~(d_top ^ u_top) & G_top)
// which is two LOP3 per word once lo_top is a compile-time flag.  The 32 result bits sit at the
// top bit of each field; they are collected with a multiply (bits of one word at stride W are
// gathered to the top of the product by a constant whose partial products never collide, valid for
// groups of at most W bits) and shifted into the result word with one funnel shift per group.
//
// Cost per word: 6 ALU ops + (1 IMAD + 1 SHF) per gather group, i.e. about 8W/32 instructions per
// code instead of ~4.  tools/swar_model.py is the executable model of exactly
// these operations and checks them against brute force for every width.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace sss {

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---- compile-time tables ----------------------------------------------------------------------

// top bit of every field, restricted to word i of the block
SSS_HD constexpr uint32_t swar_hmask(int w, int i)
{
    uint32_t m = 0;
    for (int j = 0; j < 32; j++) {
        const int p = j * w + w - 1;
        if ((p >> 5) == i) m |= 1u << (p & 31);
    }
    return m;
}

SSS_HD constexpr int swar_popc(uint32_t v)
{
    int c = 0;
    for (; v; v &= v - 1) c++;
    return c;
}

// gather groups of word i: consecutive runs of at most w result bits (ascending)
SSS_HD constexpr int swar_num_groups(int w, int i) { return (swar_popc(swar_hmask(w, i)) + w - 1) / w; }

struct SwarGroup {
    uint32_t mask;   // the result bits of this group
    uint32_t magic;  // multiplier that moves them, in order, to the top `count` bits of the product
    int count;
};

SSS_HD constexpr SwarGroup swar_group(int w, int i, int g)
{
    const uint32_t h = swar_hmask(w, i);
    // first pass: how many bits fall in this group
    int count = 0, seen = 0;
    for (int p = 0; p < 32; p++) {
        if (!((h >> p) & 1)) continue;
        if (seen / w == g) count++;
        seen++;
    }
    uint32_t mask = 0, magic = 0;
    int k = 0;
    seen = 0;
    for (int p = 0; p < 32; p++) {
        if (!((h >> p) & 1)) continue;
        if (seen / w == g) {
            mask |= 1u << p;
            magic |= 1u << ((32 - count + k) - p);
            k++;
        }
        seen++;
    }
    return SwarGroup{mask, magic, count};
}

// does a field straddle the boundary between word i-1 and word i?
SSS_HD constexpr bool swar_borrow_in(int w, int i) { return i > 0 && ((32 * i) % w) != 0; }

// ---- per-launch constants (replicated bounds), passed by value as a kernel parameter -----------

template <int W>
struct SwarConsts {
    uint32_t lol[W];  // rep(lo & ~msb)
    uint32_t dh[W];   // rep(d | msb)
    uint32_t drh[W];  // rep(d & msb)
    uint32_t lo_top;  // top bit of lo (selects the kernel instantiation; kept for reference)
};

#if defined(__CUDA_ARCH__)
#define SSS_UNROLL _Pragma("unroll")
#else
#define SSS_UNROLL
#endif

template <int W>
SSS_HD inline SwarConsts<W> make_swar_consts(uint32_t lo, uint32_t hi)
{
    SwarConsts<W> c{};
    const uint32_t msb = 1u << (W - 1);
    const uint32_t d = hi - lo;
    // every index below is a compile-time constant once the loops are unrolled, so that in device
    // code (scan_range_dev_kernel builds its constants itself) the arrays stay in registers
    auto put = [&](uint32_t (&dst)[W], uint32_t v) {
        SSS_UNROLL
        for (int i = 0; i < W; i++) dst[i] = 0;
        SSS_UNROLL
        for (int j = 0; j < 32; j++) {
            const int p = j * W;
            const unsigned long long vv = (unsigned long long)v << (p & 31);
            dst[p >> 5] |= (uint32_t)vv;
            if ((p >> 5) + 1 < W) dst[(p >> 5) + 1 < W ? (p >> 5) + 1 : 0] |= (uint32_t)(vv >> 32);
        }
    };
    put(c.lol, lo & ~msb);
    put(c.dh, d | msb);
    put(c.drh, d & msb);
    c.lo_top = (lo >> (W - 1)) & 1u;
    return c;
}

// ---- the block evaluator -------------------------------------------------------------------------

// dst = a - b over W words; a borrow is carried from word i-1 into word i only where a field
// straddles that boundary (elsewhere it is provably zero, so the chain is cut there)
template <int W, int I>
__device__ __forceinline__ void swar_sub_word(uint32_t& dst, uint32_t a, uint32_t b)
{
    constexpr bool in = swar_borrow_in(W, I);
    constexpr bool out = (I + 1 < W) && swar_borrow_in(W, I + 1);
    if constexpr (!in && !out) {
        dst = a - b;
    } else if constexpr (!in && out) {
        asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(dst) : "r"(a), "r"(b));
    } else if constexpr (in && out) {
        asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(dst) : "r"(a), "r"(b));
    } else {
        asm volatile("subc.u32 %0, %1, %2;" : "=r"(dst) : "r"(a), "r"(b));
    }
}

// 32 codes (W words) -> 32 result bits.  kLoTop = top bit of lo.
template <int W, bool kLoTop>
__device__ __forceinline__ uint32_t swar_eval_block(const uint32_t (&x)[W], const SwarConsts<W>& c)
{
    static_assert(W >= 2 && W <= 16, "SWAR path is instantiated for 2..16 bits");
    uint32_t t[W], g[W];
    static_for<0, W>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr uint32_t H = swar_hmask(W, i);
        swar_sub_word<W, i>(t[i], x[i] | H, c.lol[i]);
    });
    static_for<0, W>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr uint32_t H = swar_hmask(W, i);
        swar_sub_word<W, i>(g[i], c.dh[i], t[i] & ~H);
    });
    if constexpr (W == 2) {
        // 16 result bits per word at stride 2: the multiply-gather moves only 2 bits at a time here,
        // so interleave the two words' results (word 0 -> even bits, word 1 -> odd bits) and undo the
        // interleave with the 4-step inverse perfect shuffle (even bits -> low half, odd -> high half)
        uint32_t r[2];
        static_for<0, 2>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr uint32_t H = 0xAAAAAAAAu;
            const uint32_t v = kLoTop ? ((t[i] ^ x[i]) | ~H) : ((t[i] ^ x[i] ^ H) | ~H);
            const uint32_t d = c.drh[i];
            r[i] = (d & ~v) | (~(d ^ v) & g[i]);
        });
        uint32_t z = (r[0] >> 1) | r[1];
        uint32_t u;
        u = (z ^ (z >> 1)) & 0x22222222u; z = z ^ u ^ (u << 1);
        u = (z ^ (z >> 2)) & 0x0C0C0C0Cu; z = z ^ u ^ (u << 2);
        u = (z ^ (z >> 4)) & 0x00F000F0u; z = z ^ u ^ (u << 4);
        u = (z ^ (z >> 8)) & 0x0000FF00u; z = z ^ u ^ (u << 8);
        return z;
    }
    uint32_t res = 0;
    static_for<0, W>([&](auto rc) {
        constexpr int i = W - 1 - decltype(rc)::value;  // highest word first: code 0 must end at bit 0
        constexpr uint32_t H = swar_hmask(W, i);
        // v: u_top at the field tops, all ones elsewhere (so that r is zero off the field tops)
        const uint32_t v = kLoTop ? ((t[i] ^ x[i]) | ~H) : ((t[i] ^ x[i] ^ H) | ~H);
        const uint32_t d = c.drh[i];
        const uint32_t r = (d & ~v) | (~(d ^ v) & g[i]);
        constexpr int ng = swar_num_groups(W, i);
        static_for<0, ng>([&](auto gc) {
            constexpr int gi = ng - 1 - decltype(gc)::value;
            constexpr SwarGroup grp = swar_group(W, i, gi);
            // no pre-mask: result bit i times magic term k lands at 32-count+k+W*(i-first-k); for
            // i-first-k > 0 that is past bit 31 (dropped), for < 0 it is below the window and all
            // such positions are distinct (|k-k'| < W), so nothing collides with or carries into
            // the top `count` bits
            res = __funnelshift_l(r * grp.magic, res, grp.count);
        });
    });
    return res;
}

}  // namespace sss
