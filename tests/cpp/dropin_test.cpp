// dropin_test.cpp -- the reference's own unit tests (test/simd_scan_tests.cpp:6-150,
// test/util_tests.cpp:15-36), restated against include/simd_scan.hpp + libsimd_scan_host.so.
// A translation unit written for the reference header must compile and pass unchanged; this file
// uses only reference names.  Built twice (BITS_NEEDED = 9 default, and -DBITS_NEEDED=17) by
// __graft_entry__.build(); run on the GPU box by tests/test_gpu_dropin.py.
#include "simd_scan.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <iostream>
#include <thread>

static int failures = 0;
#define REQUIRE(cond)                                                          \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            failures++;                                                        \
        }                                                                      \
    } while (0)

// bit order of the reference's get_bit (src/util.cpp:51-58)
static bool get_bit(std::vector<uint8_t> const& v, size_t i) { return (v[i / 8] >> (i % 8)) & 1; }

typedef void (*decompress_t)(__m128i*, size_t, int*);
typedef int (*scan_t)(int, __m128i*, size_t, std::vector<uint8_t>&);
typedef void (*shared_t)(std::vector<int> const&, __m128i*, size_t, std::vector<std::vector<uint8_t>>&);

static void test_compress_decompress()
{
    // test/simd_scan_tests.cpp:6-43
    size_t input_size = (BITS_NEEDED >= 16 ? (1u << 16) : (1u << BITS_NEEDED)) - 3;
    std::vector<uint16_t> input_numbers(input_size);
    for (size_t i = 0; i < input_size; i++) input_numbers[i] = (uint16_t)i;
    auto compressed = compress_9bit_input(input_numbers);
    __m128i* compressed_ptr = (__m128i*)compressed.get();

    decompress_t variants[] = {decompress_unvectorized, decompress_128_sweep, decompress_128_nosweep,
                               decompress_128_9bit, decompress_128, decompress_128_unrolled,
                               decompress_128_aligned, decompress_256, decompress_256_avx2};
    for (decompress_t f : variants) {
        size_t output_buffer_size = decompression_output_buffer_size(input_size) / sizeof(int);
        auto result_buffer = std::make_unique<int[]>(output_buffer_size);
        f(compressed_ptr, input_numbers.size(), result_buffer.get());
        bool same = true;
        for (size_t i = 0; i < input_numbers.size(); i++) same &= (input_numbers[i] == result_buffer[i]);
        REQUIRE(same);
        // nothing written past input_size ints
        for (size_t i = input_size; i < output_buffer_size; i++) same &= (result_buffer[i] == 0);
        REQUIRE(same);
    }
#if BITS_NEEDED == 9
    // golden bytes of the unmodified reference
    const uint8_t expect[18] = {0x00, 0x02, 0x08, 0x18, 0x40, 0xa0, 0x80, 0x81, 0x03,
                                0x08, 0x12, 0x28, 0x58, 0xc0, 0xa0, 0x81, 0x83, 0x07};
    REQUIRE(std::memcmp(compressed.get(), expect, 18) == 0);
    REQUIRE(compressed_buffer_size(9, 509) == 829);
    REQUIRE(decompression_output_buffer_size(509) == 2068);
    REQUIRE(scan_output_buffer_size(12) == 34);
#endif
}

static std::vector<uint16_t> twelve() { return {1, 2, 3, 3, 2, 1, 1, 2, 3, 1, 2, 3}; }

static void test_scan()
{
    // test/simd_scan_tests.cpp:45-82, for every scan variant
    auto input_numbers = twelve();
    auto compressed = compress_9bit_input(input_numbers);
    __m128i* compressed_ptr = (__m128i*)compressed.get();
    scan_t variants[] = {scan_unvectorized, scan_128, scan_128_unrolled, scan_256, scan_256_unrolled};
    for (scan_t f : variants) {
        std::vector<uint8_t> output(scan_output_buffer_size(input_numbers.size()));
        int predicate_key = 3;
        int hits = f(predicate_key, compressed_ptr, input_numbers.size(), output);
        REQUIRE(hits == 4);
        for (size_t i = 0; i < input_numbers.size(); i++)
            REQUIRE(get_bit(output, i) == (input_numbers[i] == predicate_key));
    }
#if BITS_NEEDED == 9
    const uint8_t packed[14] = {0x01, 0x04, 0x0c, 0x18, 0x20, 0x20, 0x40, 0x00, 0x01, 0x03, 0x02, 0x08, 0x18, 0x00};
    REQUIRE(std::memcmp(compressed.get(), packed, 14) == 0);
    std::vector<uint8_t> output(scan_output_buffer_size(12));
    REQUIRE(scan_128(1, compressed_ptr, 12, output) == 4);
    REQUIRE(output[0] == 0x61 && output[1] == 0x02);
    REQUIRE(scan_128(2, compressed_ptr, 12, output) == 4);
    REQUIRE(output[0] == 0x92 && output[1] == 0x04);
    // tail semantics: key 0 only matches the zero padding
    std::vector<uint8_t> o8(scan_output_buffer_size(12)), o32(scan_output_buffer_size(12)), o64(scan_output_buffer_size(12));
    REQUIRE(scan_128(0, compressed_ptr, 12, o8) == 4 && o8[1] == 0xf0 && o8[2] == 0);
    REQUIRE(scan_128_unrolled(0, compressed_ptr, 12, o32) == 20 && o32[1] == 0xf0 && o32[2] == 0xff && o32[3] == 0xff && o32[4] == 0);
    REQUIRE(scan_unvectorized(0, compressed_ptr, 12, o64) == 3 && o64[1] == 0x70 && o64[2] == 0);
#endif
}

static void test_shared_scan()
{
    // test/simd_scan_tests.cpp:84-106, for every shared variant
    auto input_numbers = twelve();
    auto compressed = compress_9bit_input(input_numbers);
    __m128i* compressed_ptr = (__m128i*)compressed.get();
    std::vector<int> predicate_keys{1, 2, 3};
    shared_t variants[] = {shared_scan_128_sequential, shared_scan_128_sequential_unrolled, shared_scan_128_threaded,
                           shared_scan_128_standard, shared_scan_128_standard_unrolled, shared_scan_128_parallel,
                           shared_scan_256_sequential, shared_scan_256_standard, shared_scan_256_parallel};
    for (shared_t f : variants) {
        auto output_buffer_size = scan_output_buffer_size(input_numbers.size());
        std::vector<std::vector<uint8_t>> outputs(predicate_keys.size(), std::vector<uint8_t>(output_buffer_size));
        f(predicate_keys, compressed_ptr, input_numbers.size(), outputs);
        for (size_t key_id = 0; key_id < predicate_keys.size(); key_id++)
            for (size_t i = 0; i < input_numbers.size(); i++)
                REQUIRE(get_bit(outputs[key_id], i) == (input_numbers[i] == predicate_keys[key_id]));
    }
}

static void test_simple_shared_scan()
{
    // test/simd_scan_tests.cpp:108-150: the linear layout, whole-vector equality incl. padding
    auto input_numbers = twelve();
    auto compressed = compress_9bit_input(input_numbers);
    __m128i* compressed_ptr = (__m128i*)compressed.get();
    std::vector<int> predicate_keys1{1};
    std::vector<int> predicate_keys2{2, 3};
    size_t output_buffer_size = scan_output_buffer_size(input_numbers.size());
    std::vector<uint8_t> outputs1(predicate_keys1.size() * output_buffer_size);
    std::vector<uint8_t> outputs2(predicate_keys2.size() * output_buffer_size);
    std::vector<uint8_t> compare_output(output_buffer_size);

    shared_scan_128_linear_simple(predicate_keys1, compressed_ptr, input_numbers.size(), outputs1);
    int hits = scan_128(predicate_keys1[0], compressed_ptr, input_numbers.size(), compare_output);
    REQUIRE(hits == 4);
    REQUIRE(outputs1 == compare_output);

    shared_scan_128_linear_simple(predicate_keys2, compressed_ptr, input_numbers.size(), outputs2);
    hits = scan_128(predicate_keys2[0], compressed_ptr, input_numbers.size(), compare_output);
    REQUIRE(hits == 4);
    for (size_t i = 0; i < compare_output.size(); ++i) REQUIRE(outputs2[i * 2] == compare_output[i]);
    hits = scan_128(predicate_keys2[1], compressed_ptr, input_numbers.size(), compare_output);
    REQUIRE(hits == 4);
    for (size_t i = 0; i < compare_output.size(); ++i) REQUIRE(outputs2[i * 2 + 1] == compare_output[i]);

    std::vector<uint8_t> std_out(predicate_keys2.size() * output_buffer_size);
    shared_scan_128_linear_standard(predicate_keys2, compressed_ptr, input_numbers.size(), std_out);
    REQUIRE(std_out == outputs2);
    std::vector<uint8_t> static_out(predicate_keys2.size() * output_buffer_size);
    shared_scan_128_linear_static<2>(predicate_keys2, compressed_ptr, input_numbers.size(), static_out);
    REQUIRE(static_out == outputs2);
#if BITS_NEEDED == 9
    REQUIRE(outputs2[0] == 0x92 && outputs2[1] == 0x0c && outputs2[2] == 0x04 && outputs2[3] == 0x09);
#endif

    // unsupported K: message on std::cerr, output untouched (src/simd_scan_shared_linear.cpp:79-80)
    std::vector<int> three{1, 2, 3};
    std::vector<uint8_t> untouched(3 * output_buffer_size, 0xAB);
    std::stringstream captured;
    auto* old = std::cerr.rdbuf(captured.rdbuf());
    shared_scan_128_linear_simple(three, compressed_ptr, input_numbers.size(), untouched);
    std::cerr.rdbuf(old);
    REQUIRE(captured.str() == "not supported for 3 predicate keys!\n");
    bool same = true;
    for (uint8_t b : untouched) same &= (b == 0xAB);
    REQUIRE(same);
}

static void test_additive_api()
{
    // range predicate + 64-bit count + ascending position list (no reference counterpart)
    std::vector<uint32_t> codes(100003);
    const uint32_t mask = BITS_NEEDED >= 32 ? 0xFFFFFFFFu : ((1u << BITS_NEEDED) - 1u);
    for (size_t i = 0; i < codes.size(); i++) codes[i] = (uint32_t)(i * 2654435761u) & mask;
    auto packed = sss_host::compress(BITS_NEEDED, codes);
    std::vector<uint8_t> out(scan_output_buffer_size(codes.size()));
    std::vector<uint64_t> pos;
    const sss_host::Range r{mask / 4, mask / 2};
    uint64_t hits = sss_host::scan_range(BITS_NEEDED, r, packed.get(), codes.size(), out, &pos);
    std::vector<uint64_t> expect;
    for (size_t i = 0; i < codes.size(); i++)
        if (codes[i] >= r.lo && codes[i] <= r.hi) expect.push_back(i);
    REQUIRE(hits == expect.size());
    REQUIRE(pos == expect);
    std::vector<std::vector<uint8_t>> outs(2, std::vector<uint8_t>(scan_output_buffer_size(codes.size())));
    std::vector<uint64_t> counts;
    sss_host::shared_scan_range(BITS_NEEDED, {r, {1, 0}}, packed.get(), codes.size(), outs, &counts);
    REQUIRE(counts.size() == 2 && counts[0] == hits && counts[1] == 0);
    REQUIRE(outs[0] == out);
}

static void test_concurrent_callers()
{
    // The reference's functions are stateless (src/simd_scan.cpp:103-158) and its own threaded variant
    // calls scan_128 from an OpenMP team (src/simd_scan_shared.cpp:25-32): 8 threads call scan_128 and
    // shared_scan_128_standard at the same time, each on its OWN column, and every result must equal
    // the single-threaded one.
    constexpr int kThreads = 8;
    const uint32_t mask = BITS_NEEDED >= 16 ? 0xFFFFu : ((1u << BITS_NEEDED) - 1u);
    const size_t n = 200003;
    std::vector<std::vector<uint16_t>> columns(kThreads, std::vector<uint16_t>(n));
    for (int t = 0; t < kThreads; t++)
        for (size_t i = 0; i < n; i++) columns[t][i] = (uint16_t)(((i + 1) * 2654435761u * (t + 3)) >> 7) & mask;
    std::vector<std::unique_ptr<uint64_t[]>> packed(kThreads);
    for (int t = 0; t < kThreads; t++) packed[t] = compress_9bit_input(columns[t]);
    const std::vector<int> keys = {1, 2, 3, (int)mask, 0, 77 & (int)mask};
    struct Result {
        int hits = 0;
        std::vector<uint8_t> scan;
        std::vector<std::vector<uint8_t>> shared;
    };
    auto run = [&](int t, Result& r) {
        r.scan.assign(scan_output_buffer_size(n), 0);
        r.hits = scan_128(columns[t][5], (__m128i*)packed[t].get(), n, r.scan);
        r.shared.assign(keys.size(), std::vector<uint8_t>(scan_output_buffer_size(n), 0));
        shared_scan_128_standard(keys, (__m128i*)packed[t].get(), n, r.shared);
    };
    std::vector<Result> serial(kThreads), parallel(kThreads);
    for (int t = 0; t < kThreads; t++) run(t, serial[t]);
    for (int round = 0; round < 3; round++) {
        std::vector<std::thread> team;
        for (int t = 0; t < kThreads; t++) team.emplace_back([&, t] { run(t, parallel[t]); });
        for (auto& th : team) th.join();
        for (int t = 0; t < kThreads; t++) {
            REQUIRE(parallel[t].hits == serial[t].hits);
            REQUIRE(parallel[t].scan == serial[t].scan);
            REQUIRE(parallel[t].shared == serial[t].shared);
        }
    }
    // and the serial results are right
    for (int t = 0; t < kThreads; t++) {
        int expect = 0;
        bool bits = true;
        for (size_t i = 0; i < n; i++) {
            const bool hit = columns[t][i] == columns[t][5];
            expect += hit;
            bits &= get_bit(serial[t].scan, i) == hit;
        }
        REQUIRE(serial[t].hits == expect);
        REQUIRE(bits);
    }
}

int main()
{
    try {
        test_compress_decompress();
        test_scan();
        test_shared_scan();
        test_simple_shared_scan();
        test_additive_api();
        test_concurrent_callers();
    } catch (std::exception const& e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    if (failures == 0) std::printf("All tests passed (BITS_NEEDED=%d)\n", BITS_NEEDED);
    return failures == 0 ? 0 : 1;
}
