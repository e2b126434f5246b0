// simd_scan.hpp -- the reference's public surface (RRr89/Shared_SIMD_Scan src/simd_scan.hpp:20-120),
// re-created on top of the B200 C ABI (include/sss_b200.h).
//
// Every entry point keeps the reference's name, argument list, argument meaning (`input_size` is a
// number of CODES, `input` is the packed buffer cast to __m128i*, outputs are caller-sized vectors
// that are written from index 0 and never resized) and tail behaviour, so a
// translation unit written against the reference compiles and links against libsimd_scan_host.so
// unchanged.  All variants of one family map to the same CUDA kernel; what distinguishes them here
// is only the tail mode they ask for.
//
// Differences a caller can observe:
//   * BITS_NEEDED may be predefined to any width 1..32 (the reference hard-codes 9, :12, and its
//     SIMD paths are exact only for {1..26, 28, 32}); it is forwarded to the library at run time.
//   * A CUDA failure throws std::runtime_error (the reference cannot fail).  There is no CPU path.
//   * Additive API for the north star (range predicates, 64-bit counts, position lists, shared-scan
//     counts) lives in namespace sss_host below.
#pragma once

#include <immintrin.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#ifndef BITS_NEEDED
#define BITS_NEEDED 9  // reference default, src/simd_scan.hpp:12
#endif

// ---- buffer sizes in bytes, padding included (reference :20-40) -----------------------------------

constexpr size_t compressed_buffer_size(uint8_t compression, size_t input_array_size)
{
    return (static_cast<size_t>(compression) * input_array_size + 7) / 8 + 256;
}

constexpr size_t decompression_output_buffer_size(size_t input_array_size) { return input_array_size * 4 + 32; }

constexpr size_t scan_output_buffer_size(size_t input_array_size) { return (input_array_size + 7) / 8 + 32; }

// ---- implemented in libsimd_scan_host.so ----------------------------------------------------------

namespace sss_host {

// tail modes, == SSS_TAIL_* of sss_b200.h
enum Tail : int { kExact = 1, kGroup8 = 8, kWord32 = 32, kCursor32 = 132, kCursor64 = 164 };

std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint16_t>& input);
std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint32_t>& input);
void decompress(int w, const void* input, size_t input_size, int* output);
int scan(int w, Tail tail, int predicate_key, const void* input, size_t input_size, std::vector<uint8_t>& output);
void shared_scan(int w, Tail tail, const std::vector<int>& predicate_keys, const void* input, size_t input_size,
                 std::vector<std::vector<uint8_t>>& outputs);
void shared_scan_linear(int w, const std::vector<int>& predicate_keys, const void* input, size_t input_size,
                        std::vector<uint8_t>& output);

// additive API (BASELINE.json north star; the commented-out intent at reference :75-84)
struct Range {
    uint32_t lo, hi;  // inclusive, unsigned
};
uint64_t scan_range(int w, Range r, const void* input, size_t input_size, std::vector<uint8_t>& output,
                    std::vector<uint64_t>* positions = nullptr);
void shared_scan_range(int w, const std::vector<Range>& ranges, const void* input, size_t input_size,
                       std::vector<std::vector<uint8_t>>& outputs, std::vector<uint64_t>* counts = nullptr);
void shared_scan_linear_range(int w, const std::vector<Range>& ranges, const void* input, size_t input_size,
                              std::vector<uint8_t>& output, std::vector<uint64_t>* counts = nullptr);

}  // namespace sss_host

// ---- compression (reference :46) -------------------------------------------------------------------

inline std::unique_ptr<uint64_t[]> compress_9bit_input(std::vector<uint16_t>& input)
{
    return sss_host::compress(BITS_NEEDED, input);
}

// ---- decompression (reference :51-73): nine variants, one kernel -----------------------------------

#define SSS_DECOMPRESS_VARIANT(name)                                           \
    inline void name(__m128i* input, size_t input_size, int* output)           \
    {                                                                          \
        sss_host::decompress(BITS_NEEDED, input, input_size, output);          \
    }
SSS_DECOMPRESS_VARIANT(decompress_unvectorized)
SSS_DECOMPRESS_VARIANT(decompress_128_sweep)
SSS_DECOMPRESS_VARIANT(decompress_128_nosweep)
SSS_DECOMPRESS_VARIANT(decompress_128_9bit)
SSS_DECOMPRESS_VARIANT(decompress_128)
SSS_DECOMPRESS_VARIANT(decompress_128_unrolled)
SSS_DECOMPRESS_VARIANT(decompress_128_aligned)
SSS_DECOMPRESS_VARIANT(decompress_256)
SSS_DECOMPRESS_VARIANT(decompress_256_avx2)
#undef SSS_DECOMPRESS_VARIANT

// ---- single-predicate scan (reference :89-96); returns the hit count -----------------------------

#define SSS_SCAN_VARIANT(name, tail)                                                                          \
    inline int name(int predicate_key, __m128i* input, size_t input_size, std::vector<uint8_t>& output)       \
    {                                                                                                         \
        return sss_host::scan(BITS_NEEDED, sss_host::tail, predicate_key, input, input_size, output);         \
    }
SSS_SCAN_VARIANT(scan_unvectorized, kCursor64)
SSS_SCAN_VARIANT(scan_128, kGroup8)
SSS_SCAN_VARIANT(scan_128_unrolled, kWord32)
SSS_SCAN_VARIANT(scan_256, kGroup8)
SSS_SCAN_VARIANT(scan_256_unrolled, kWord32)
#undef SSS_SCAN_VARIANT

// ---- shared scan, K bitvectors (reference :102-113) -------------------------------------------------

#define SSS_SHARED_VARIANT(name, tail)                                                                        \
    inline void name(std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,               \
                     std::vector<std::vector<uint8_t>>& outputs)                                              \
    {                                                                                                         \
        sss_host::shared_scan(BITS_NEEDED, sss_host::tail, predicate_keys, input, input_size, outputs);       \
    }
SSS_SHARED_VARIANT(shared_scan_128_sequential, kGroup8)
SSS_SHARED_VARIANT(shared_scan_128_sequential_unrolled, kWord32)
SSS_SHARED_VARIANT(shared_scan_128_threaded, kGroup8)
SSS_SHARED_VARIANT(shared_scan_128_standard, kGroup8)
SSS_SHARED_VARIANT(shared_scan_128_standard_unrolled, kWord32)
SSS_SHARED_VARIANT(shared_scan_128_parallel, kCursor32)
SSS_SHARED_VARIANT(shared_scan_256_sequential, kGroup8)
SSS_SHARED_VARIANT(shared_scan_256_standard, kGroup8)
SSS_SHARED_VARIANT(shared_scan_256_parallel, kCursor32)
#undef SSS_SHARED_VARIANT

// ---- shared scan, one linear output vector (reference :119-122) -------------------------------------

inline void shared_scan_128_linear_standard(std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,
                                            std::vector<uint8_t>& outputs)
{
    sss_host::shared_scan_linear(BITS_NEEDED, predicate_keys, input, input_size, outputs);
}

// K must be one of 1, 2, 4, ..., 1024; anything else prints the reference's message to std::cerr
// and leaves `output` untouched (reference src/simd_scan_shared_linear.cpp:64-82)
void shared_scan_128_linear_simple_w(int w, std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,
                                     std::vector<uint8_t>& output);

inline void shared_scan_128_linear_simple(std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,
                                          std::vector<uint8_t>& output)
{
    shared_scan_128_linear_simple_w(BITS_NEEDED, predicate_keys, input, input_size, output);
}

template <size_t NUM>
void shared_scan_128_linear_static(std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,
                                   std::vector<uint8_t>& output)
{
    // the reference reads exactly NUM keys (src/simd_scan.hpp:178-191)
    std::vector<int> keys(predicate_keys.begin(), predicate_keys.begin() + NUM);
    sss_host::shared_scan_linear(BITS_NEEDED, keys, input, input_size, output);
}
