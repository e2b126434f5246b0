// ref_shim.cpp -- C-linkage doorway into the UNMODIFIED reference, one shared object per bit width.
//
// TEST INFRASTRUCTURE ONLY.  oracle/build_ref.sh compiles this file together with the reference's
// own translation units (read in place from /root/reference/src, never copied into the repo) into
// oracle/_ref/libref_wNN.so.  The only thing that differs between widths is the value of the
// reference's BITS_NEEDED macro (src/simd_scan.hpp:12), substituted on the fly by the build
// recipe.  Nothing here re-implements reference logic: each entry point sizes the std::vector
// outputs the way the reference's tests do (test/simd_scan_tests.cpp:55-56,94-95,118-120), calls
// the reference function, and copies the vector out.
//
// Used by: tests/ (pin the oracle restatement), tests/golden/make_golden.py (golden vectors),
// bench.py cpu_baseline / --impl reference (timing).  Never by the product path.

#include "simd_scan.hpp"  // the reference header (src/simd_scan.hpp), BITS_NEEDED substituted

#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace {

typedef void (*decompress_fn)(__m128i*, size_t, int*);
typedef int (*scan_fn)(int, __m128i*, size_t, std::vector<uint8_t>&);
typedef void (*shared_fn)(std::vector<int> const&, __m128i*, size_t, std::vector<std::vector<uint8_t>>&);
typedef void (*linear_fn)(std::vector<int> const&, __m128i*, size_t, std::vector<uint8_t>&);

// variant ids are part of this shim's ABI (mirrored in tests/refshim.py)
decompress_fn decompress_variant(int v)
{
    switch (v) {
        case 0: return decompress_unvectorized;   // src/simd_scan_decompression.cpp:6
        case 1: return decompress_128_sweep;      // :57
        case 2: return decompress_128_nosweep;    // :120
        case 3: return decompress_128_9bit;       // :177 (w == 9 only)
        case 4: return decompress_128;            // :237
        case 5: return decompress_128_unrolled;   // :273
        case 6: return decompress_128_aligned;    // :346
#ifdef __AVX__
        case 7: return decompress_256;            // :392
#endif
#ifdef __AVX2__
        case 8: return decompress_256_avx2;       // :427
#endif
    }
    return nullptr;
}

scan_fn scan_variant(int v)
{
    switch (v) {
        case 0: return scan_unvectorized;   // src/simd_scan.cpp:20
        case 1: return scan_128;            // :103
        case 2: return scan_128_unrolled;   // :175
#ifdef __AVX__
        case 3: return scan_256;            // :220
        case 4: return scan_256_unrolled;   // :273
#endif
    }
    return nullptr;
}

shared_fn shared_variant(int v)
{
    switch (v) {
        case 0: return shared_scan_128_sequential;           // src/simd_scan_shared.cpp:9
        case 1: return shared_scan_128_sequential_unrolled;  // :17
        case 2: return shared_scan_128_threaded;             // :25
        case 3: return shared_scan_128_standard;             // :34
        case 4: return shared_scan_128_standard_unrolled;    // :111
        case 5: return shared_scan_128_parallel;             // :154
#ifdef __AVX__
        case 6: return shared_scan_256_sequential;           // :287
        case 7: return shared_scan_256_standard;             // :295
        case 8: return shared_scan_256_parallel;             // :333
#endif
    }
    return nullptr;
}

linear_fn linear_variant(int v)
{
    switch (v) {
        case 0: return shared_scan_128_linear_standard;  // src/simd_scan_shared_linear.cpp:9
        case 1: return shared_scan_128_linear_simple;    // :64
    }
    return nullptr;
}

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Range partition at multiples of 64 codes (=> byte-aligned shard starts for every width).
struct Shard { size_t first, count; };
std::vector<Shard> make_shards(size_t n, int threads)
{
    if (threads < 1) threads = 1;
    size_t per = ((n + threads - 1) / threads + 63) / 64 * 64;
    std::vector<Shard> s;
    for (int t = 0; t < threads; t++) {
        size_t first = (size_t)t * per;
        if (first >= n) break;
        s.push_back({first, std::min(per, n - first)});
    }
    return s;
}

__m128i* shard_ptr(const void* packed, size_t first)
{
    return (__m128i*)((uint8_t*)packed + first * (size_t)BITS_NEEDED / 8);
}

}  // namespace

extern "C" {

int ref_bits(void) { return BITS_NEEDED; }

size_t ref_compressed_buffer_size(size_t n) { return compressed_buffer_size(BITS_NEEDED, n); }
size_t ref_decompression_output_buffer_size(size_t n) { return decompression_output_buffer_size(n); }
size_t ref_scan_output_buffer_size(size_t n) { return scan_output_buffer_size(n); }

// compress_9bit_input (src/simd_scan_compression.cpp:53); out gets (buffer_size/8)*8 bytes.
size_t ref_compress(const uint16_t* in, size_t n, uint8_t* out, size_t out_bytes)
{
    std::vector<uint16_t> v(in, in + n);
    auto buf = compress_9bit_input(v);
    size_t produced = compressed_buffer_size(BITS_NEEDED, n) / sizeof(uint64_t) * sizeof(uint64_t);
    std::memcpy(out, buf.get(), std::min(produced, out_bytes));
    return produced;
}

int ref_decompress(int variant, const void* packed, size_t n, int* out)
{
    decompress_fn f = decompress_variant(variant);
    if (!f) return -1;
    f((__m128i*)packed, n, out);
    return 0;
}

// returns hits; out gets scan_output_buffer_size(n) bytes (zero-initialised vector, as the tests do)
long ref_scan(int variant, int key, const void* packed, size_t n, uint8_t* out)
{
    scan_fn f = scan_variant(variant);
    if (!f) return -1;
    std::vector<uint8_t> o(scan_output_buffer_size(n));
    int hits = f(key, (__m128i*)packed, n, o);
    std::memcpy(out, o.data(), o.size());
    return hits;
}

// outs: K consecutive regions of scan_output_buffer_size(n) bytes
int ref_shared_scan(int variant, const int* keys, size_t K, const void* packed, size_t n, uint8_t* outs)
{
    shared_fn f = shared_variant(variant);
    if (!f) return -1;
    size_t sz = scan_output_buffer_size(n);
    std::vector<int> k(keys, keys + K);
    std::vector<std::vector<uint8_t>> o(K, std::vector<uint8_t>(sz));
    f(k, (__m128i*)packed, n, o);
    for (size_t i = 0; i < K; i++) std::memcpy(outs + i * sz, o[i].data(), sz);
    return 0;
}

// out: K * scan_output_buffer_size(n) bytes (test/simd_scan_tests.cpp:119)
int ref_shared_scan_linear(int variant, const int* keys, size_t K, const void* packed, size_t n, uint8_t* out)
{
    linear_fn f = linear_variant(variant);
    if (!f) return -1;
    std::vector<int> k(keys, keys + K);
    std::vector<uint8_t> o(K * scan_output_buffer_size(n));
    f(k, (__m128i*)packed, n, o);
    std::memcpy(out, o.data(), o.size());
    return 0;
}

// ---- timing (CPU baseline).  Outputs are allocated once outside the timed region; every thread
// runs the unmodified reference function on its own code range with a private output vector.
// Returns the best wall time in seconds over `reps`; *hits gets the summed return values.

double ref_time_scan(int variant, int key, const void* packed, size_t n, int threads, int reps, long long* hits)
{
    scan_fn f = scan_variant(variant);
    if (!f) return -1.0;
    auto shards = make_shards(n, threads);
    std::vector<std::vector<uint8_t>> outs;
    for (auto& s : shards) outs.emplace_back(scan_output_buffer_size(s.count));
    std::vector<long long> h(shards.size());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        std::vector<std::thread> th;
        for (size_t i = 1; i < shards.size(); i++)
            th.emplace_back([&, i] { h[i] = f(key, shard_ptr(packed, shards[i].first), shards[i].count, outs[i]); });
        h[0] = f(key, shard_ptr(packed, shards[0].first), shards[0].count, outs[0]);
        for (auto& t : th) t.join();
        best = std::min(best, now_s() - t0);
    }
    long long total = 0;
    for (auto x : h) total += x;
    if (hits) *hits = total;
    return best;
}

double ref_time_decompress(int variant, const void* packed, size_t n, int threads, int reps)
{
    decompress_fn f = decompress_variant(variant);
    if (!f) return -1.0;
    auto shards = make_shards(n, threads);
    std::vector<std::vector<int>> outs;
    for (auto& s : shards) outs.emplace_back(decompression_output_buffer_size(s.count) / sizeof(int));
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        std::vector<std::thread> th;
        for (size_t i = 1; i < shards.size(); i++)
            th.emplace_back([&, i] { f(shard_ptr(packed, shards[i].first), shards[i].count, outs[i].data()); });
        f(shard_ptr(packed, shards[0].first), shards[0].count, outs[0].data());
        for (auto& t : th) t.join();
        best = std::min(best, now_s() - t0);
    }
    return best;
}

double ref_time_shared_scan(int variant, const int* keys, size_t K, const void* packed, size_t n, int threads, int reps)
{
    shared_fn f = shared_variant(variant);
    if (!f) return -1.0;
    auto shards = make_shards(n, threads);
    std::vector<int> k(keys, keys + K);
    std::vector<std::vector<std::vector<uint8_t>>> outs;
    for (auto& s : shards) outs.emplace_back(K, std::vector<uint8_t>(scan_output_buffer_size(s.count)));
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        std::vector<std::thread> th;
        for (size_t i = 1; i < shards.size(); i++)
            th.emplace_back([&, i] { f(k, shard_ptr(packed, shards[i].first), shards[i].count, outs[i]); });
        f(k, shard_ptr(packed, shards[0].first), shards[0].count, outs[0]);
        for (auto& t : th) t.join();
        best = std::min(best, now_s() - t0);
    }
    return best;
}

double ref_time_shared_scan_linear(int variant, const int* keys, size_t K, const void* packed, size_t n, int threads, int reps)
{
    linear_fn f = linear_variant(variant);
    if (!f) return -1.0;
    auto shards = make_shards(n, threads);
    std::vector<int> k(keys, keys + K);
    std::vector<std::vector<uint8_t>> outs;
    for (auto& s : shards) outs.emplace_back(K * scan_output_buffer_size(s.count));
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        double t0 = now_s();
        std::vector<std::thread> th;
        for (size_t i = 1; i < shards.size(); i++)
            th.emplace_back([&, i] { f(k, shard_ptr(packed, shards[i].first), shards[i].count, outs[i]); });
        f(k, shard_ptr(packed, shards[0].first), shards[0].count, outs[0]);
        for (auto& t : th) t.join();
        best = std::min(best, now_s() - t0);
    }
    return best;
}

}  // extern "C"
