// simd_scan_commons.hpp -- per-width layout constants of the packed column format.
//
// Counterpart of the reference's src/simd_scan_commons.hpp.  There the per-width constants are
// SSE/AVX shuffle, shift, clean and predicate masks keyed on byte offset (w*i)/8 and bit padding
// (w*i)%8 with a period of 8 codes (:5-170).  On the GPU everything lives in 32-bit word space
// (32 codes = exactly w words), so the constants are word index / funnel
// shift pairs with a period of 32 codes, all evaluated at compile time for template<int W> kernels.
// Working in word space is also what makes every width 1..32 exact: the reference's 4-byte gather
// breaks at w = 27, 29, 30, 31, a word pair + funnel shift never does.
//
// Usable from host and device code (constexpr only).
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define SSS_HD __host__ __device__
#else
#define SSS_HD
#endif

namespace sss {

// all-ones mask of a w-bit code; counterpart of `(1 << compression) - 1` (commons :62) without the
// undefined shift at w = 32
SSS_HD constexpr uint32_t code_mask(int w) { return w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u); }

// Position of code j (0..31) inside its 32-code block of w words.
struct CodeSlot {
    int first_word;   // word holding the code's lowest bit
    int last_word;    // word holding the code's highest bit (== first_word unless it straddles)
    int shift;        // bit offset of the code inside first_word           (reference: "padding")
    int end_shift;    // 1..32: bits of last_word up to and including the code's top bit
    bool straddles;
};

SSS_HD constexpr CodeSlot code_slot(int w, int j)
{
    const int start = j * w;
    const int end = start + w - 1;
    return CodeSlot{start >> 5, end >> 5, start & 31, (end & 31) + 1, (start >> 5) != (end >> 5)};
}

// Range predicate on left-aligned codes.  A code moved to the top w bits of a 32-bit lane (whatever
// sits in the low 32-w bits) lies in [lo,hi] iff  (lane - lo_al) <= span_al  as unsigned numbers:
// the low bits of lo_al are 0 and those of lo_al + span_al are all 1, so they never change the
// outcome.  One subtract + one compare per code, no mask, no right shift.  Counterpart of the
// reference's "compare in place against a pre-shifted predicate" (commons :72-90).
struct AlignedRange {
    uint32_t lo_al;
    uint32_t span_al;
    bool empty;
};

SSS_HD constexpr AlignedRange aligned_range(int w, uint32_t lo, uint32_t hi)
{
    const uint32_t m = code_mask(w);
    if (lo > hi || lo > m) return AlignedRange{0u, 0u, true};
    if (hi > m) hi = m;
    const int s = 32 - w;
    const uint32_t fill = (s == 0) ? 0u : ((1u << s) - 1u);
    const uint32_t lo_al = lo << s;
    const uint32_t hi_al = (hi << s) | fill;
    return AlignedRange{lo_al, hi_al - lo_al, false};
}

// buffer-size helpers shared by the C ABI and the C++ layer (reference: src/simd_scan.hpp:20-40)
SSS_HD constexpr uint64_t packed_payload_bytes(uint64_t n, int w) { return (n * (uint64_t)w + 7) / 8; }
SSS_HD constexpr uint64_t round_up_u64(uint64_t x, uint64_t g) { return (x + g - 1) / g * g; }

}  // namespace sss
