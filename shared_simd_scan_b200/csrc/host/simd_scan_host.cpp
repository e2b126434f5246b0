// simd_scan_host.cpp -- the C++ drop-in layer behind include/simd_scan.hpp (libsimd_scan_host.so).
//
// Thin by design: translate the reference's argument conventions (std::vector outputs, `int` keys,
// per-variant tail behaviour) into one call of the C ABI (include/sss_b200.h), which stages the
// host buffers through the GPU.  No computation happens here; there is no CPU path.
#include "simd_scan.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <string>

#include "sss_b200.h"

namespace {

sss_ctx* context()
{
    static sss_ctx* ctx = nullptr;
    static std::once_flag once;
    static int status = SSS_OK;
    std::call_once(once, [] {
        const char* dev = std::getenv("SSS_DEVICE");
        status = sss_ctx_create(dev ? std::atoi(dev) : 0, &ctx);
    });
    if (status != SSS_OK)
        throw std::runtime_error(std::string("shared_simd_scan_b200: cannot create a GPU context: ") +
                                 sss_error_string(status));
    return ctx;
}

void must(int status, const char* what)
{
    if (status != SSS_OK)
        throw std::runtime_error(std::string("shared_simd_scan_b200: ") + what + ": " + sss_error_string(status));
}

// `int predicate_key` -> inclusive range.  The reference compares 32-bit lanes for equality
// (src/simd_scan.cpp:130), so at w == 32 a negative key is the code with the same bit pattern; at
// smaller widths a key outside [0, 2^w) matches nothing.
sss_host::Range key_range(int w, int key)
{
    if (w == 32) return {static_cast<uint32_t>(key), static_cast<uint32_t>(key)};
    if (key < 0 || static_cast<uint32_t>(key) > ((1u << w) - 1u)) return {1u, 0u};
    return {static_cast<uint32_t>(key), static_cast<uint32_t>(key)};
}

void split(const std::vector<sss_host::Range>& r, std::vector<uint32_t>& lo, std::vector<uint32_t>& hi)
{
    lo.resize(r.size());
    hi.resize(r.size());
    for (size_t i = 0; i < r.size(); i++) {
        lo[i] = r[i].lo;
        hi[i] = r[i].hi;
    }
}

std::vector<sss_host::Range> key_ranges(int w, const std::vector<int>& keys)
{
    std::vector<sss_host::Range> r(keys.size());
    for (size_t i = 0; i < keys.size(); i++) r[i] = key_range(w, keys[i]);
    return r;
}

void require_bytes(size_t have, uint64_t need, const char* what)
{
    if (have < need)
        throw std::runtime_error(std::string("shared_simd_scan_b200: ") + what +
                                 ": output vector smaller than the bytes this variant writes (size it with "
                                 "scan_output_buffer_size)");
}

void shared_impl(int w, int tail, const std::vector<sss_host::Range>& ranges, const void* input, size_t n,
                 std::vector<std::vector<uint8_t>>& outputs, std::vector<uint64_t>* counts)
{
    if (ranges.empty()) return;
    if (outputs.size() < ranges.size())
        throw std::runtime_error("shared_simd_scan_b200: shared scan: fewer output vectors than predicates");
    std::vector<uint32_t> lo, hi;
    split(ranges, lo, hi);
    const uint64_t need = sss_tail_out_bytes(n, w, tail);
    std::vector<void*> ptrs(ranges.size());
    for (size_t k = 0; k < ranges.size(); k++) {
        require_bytes(outputs[k].size(), need, "shared scan");
        ptrs[k] = outputs[k].data();
    }
    std::vector<uint64_t> local;
    if (counts) counts->assign(ranges.size(), 0);
    must(sss_host_shared_scan_ptrs(context(), input, n, w, lo.data(), hi.data(), (int)ranges.size(), tail, ptrs.data(),
                                   counts ? counts->data() : nullptr),
         "shared scan");
}

void linear_impl(int w, const std::vector<sss_host::Range>& ranges, const void* input, size_t n,
                 std::vector<uint8_t>& output, std::vector<uint64_t>* counts)
{
    if (ranges.empty()) return;
    std::vector<uint32_t> lo, hi;
    split(ranges, lo, hi);
    require_bytes(output.size(), sss_tail_out_bytes(n, w, SSS_TAIL_8) * ranges.size(), "linear shared scan");
    if (counts) counts->assign(ranges.size(), 0);
    must(sss_host_shared_scan_linear(context(), input, n, w, lo.data(), hi.data(), (int)ranges.size(), SSS_TAIL_8,
                                     output.data(), counts ? counts->data() : nullptr),
         "linear shared scan");
}

}  // namespace

namespace sss_host {

std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint32_t>& input)
{
    const size_t bytes = sss_packed_bytes(input.size(), w);
    auto buffer = std::make_unique<uint64_t[]>((bytes + 7) / 8);  // zero-initialised, like the reference's
    must(sss_host_pack(context(), input.data(), input.size(), w, buffer.get()), "compress");
    return buffer;
}

std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint16_t>& input)
{
    return compress(w, std::vector<uint32_t>(input.begin(), input.end()));
}

void decompress(int w, const void* input, size_t input_size, int* output)
{
    must(sss_host_unpack(context(), input, input_size, w, output), "decompress");
}

int scan(int w, Tail tail, int predicate_key, const void* input, size_t input_size, std::vector<uint8_t>& output)
{
    const Range r = key_range(w, predicate_key);
    require_bytes(output.size(), sss_tail_out_bytes(input_size, w, tail), "scan");
    uint64_t hits = 0;
    must(sss_host_scan_range(context(), input, input_size, w, r.lo, r.hi, tail, output.data(), &hits), "scan");
    return static_cast<int>(hits);  // the reference's return type (overflows past 2^31 matches)
}

void shared_scan(int w, Tail tail, const std::vector<int>& predicate_keys, const void* input, size_t input_size,
                 std::vector<std::vector<uint8_t>>& outputs)
{
    shared_impl(w, tail, key_ranges(w, predicate_keys), input, input_size, outputs, nullptr);
}

void shared_scan_linear(int w, const std::vector<int>& predicate_keys, const void* input, size_t input_size,
                        std::vector<uint8_t>& output)
{
    linear_impl(w, key_ranges(w, predicate_keys), input, input_size, output, nullptr);
}

uint64_t scan_range(int w, Range r, const void* input, size_t input_size, std::vector<uint8_t>& output,
                    std::vector<uint64_t>* positions)
{
    require_bytes(output.size(), sss_tail_out_bytes(input_size, w, SSS_TAIL_EXACT), "scan_range");
    uint64_t hits = 0;
    must(sss_host_scan_range(context(), input, input_size, w, r.lo, r.hi, SSS_TAIL_EXACT, output.data(), &hits),
         "scan_range");
    if (positions) {
        // ascending positions of the set bits (get_bit order, reference src/util.cpp:51-58); the
        // bitvector is already on the host, so this is plain bit iteration, not a second scan
        positions->clear();
        positions->reserve(hits);
        for (size_t byte = 0; byte < (input_size + 7) / 8; byte++)
            for (uint8_t b = output[byte]; b; b &= b - 1)
                positions->push_back(byte * 8 + static_cast<unsigned>(__builtin_ctz(b)));
    }
    return hits;
}

void shared_scan_range(int w, const std::vector<Range>& ranges, const void* input, size_t input_size,
                       std::vector<std::vector<uint8_t>>& outputs, std::vector<uint64_t>* counts)
{
    shared_impl(w, SSS_TAIL_EXACT, ranges, input, input_size, outputs, counts);
}

void shared_scan_linear_range(int w, const std::vector<Range>& ranges, const void* input, size_t input_size,
                              std::vector<uint8_t>& output, std::vector<uint64_t>* counts)
{
    linear_impl(w, ranges, input, input_size, output, counts);
}

}  // namespace sss_host

void shared_scan_128_linear_simple_w(int w, std::vector<int> const& predicate_keys, __m128i* input, size_t input_size,
                                     std::vector<uint8_t>& output)
{
    const size_t K = predicate_keys.size();
    const bool supported = K >= 1 && K <= 1024 && (K & (K - 1)) == 0;
    if (!supported) {
        std::cerr << "not supported for " << K << " predicate keys!" << std::endl;
        return;
    }
    sss_host::shared_scan_linear(w, predicate_keys, input, input_size, output);
}
