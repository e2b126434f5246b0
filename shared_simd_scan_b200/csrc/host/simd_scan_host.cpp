// simd_scan_host.cpp -- the C++ drop-in layer behind include/simd_scan.hpp (libsimd_scan_host.so).
//
// Thin by design: translate the reference's argument conventions (std::vector outputs, `int` keys,
// per-variant tail behaviour) into one call of the C ABI (include/sss_b200.h), which stages the
// host buffers through the GPU.  No computation happens here; there is no CPU path.
#include "simd_scan.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "sss_b200.h"

namespace {

// The reference's functions are stateless and may be called from any number of threads at once
// (src/simd_scan.cpp:103-158; its own shared_scan_128_threaded calls scan_128 from an OpenMP team,
// src/simd_scan_shared.cpp:25-32).  A GPU context is not stateless -- streams, staging buffers, the
// predicate index of the shared scan in flight -- so the drop-in keeps a POOL of contexts: every call
// leases one for its duration, concurrent callers get different contexts (created on demand, up to
// SSS_HOST_CONTEXTS, default 8; further callers wait for a free one).  SSS_DEVICE picks the GPU;
// SSS_DEVICES=0,1,... spreads the pool's contexts round-robin over several GPUs.
class ContextPool {
public:
    sss_ctx* acquire()
    {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            if (!free_.empty()) {
                sss_ctx* c = free_.back();
                free_.pop_back();
                return c;
            }
            if (created_ < limit()) {
                const int ordinal = created_++;
                lk.unlock();
                sss_ctx* c = nullptr;
                const int status = sss_ctx_create(device_for(ordinal), &c);
                if (status != SSS_OK) {
                    lk.lock();
                    created_--;
                    cv_.notify_one();
                    throw std::runtime_error(std::string("shared_simd_scan_b200: cannot create a GPU context: ") +
                                             sss_error_string(status));
                }
                return c;
            }
            cv_.wait(lk);
        }
    }
    void release(sss_ctx* c)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            free_.push_back(c);
        }
        cv_.notify_one();
    }

private:
    static int limit()
    {
        static const int v = [] {
            const char* e = std::getenv("SSS_HOST_CONTEXTS");
            const int n = e ? std::atoi(e) : 8;
            return n < 1 ? 1 : (n > 64 ? 64 : n);
        }();
        return v;
    }
    static int device_for(int ordinal)
    {
        static const std::vector<int> devices = [] {
            std::vector<int> d;
            if (const char* list = std::getenv("SSS_DEVICES")) {
                for (const char* p = list; *p;) {
                    char* end = nullptr;
                    const long v = std::strtol(p, &end, 10);
                    if (end == p) break;
                    d.push_back((int)v);
                    p = (*end == ',') ? end + 1 : end;
                }
            }
            if (d.empty()) {
                const char* one = std::getenv("SSS_DEVICE");
                d.push_back(one ? std::atoi(one) : 0);
            }
            return d;
        }();
        return devices[(size_t)ordinal % devices.size()];
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<sss_ctx*> free_;  // (contexts live until process exit, like the CUDA primary context)
    int created_ = 0;
};

ContextPool& pool()
{
    static ContextPool* p = new ContextPool();  // never destroyed: no CUDA calls during static destruction
    return *p;
}

// a context for the duration of one call
class Lease {
public:
    Lease() : ctx_(pool().acquire()) {}
    ~Lease() { pool().release(ctx_); }
    Lease(const Lease&) = delete;
    Lease& operator=(const Lease&) = delete;
    operator sss_ctx*() const { return ctx_; }

private:
    sss_ctx* ctx_;
};

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
    Lease ctx;
    must(sss_host_shared_scan_ptrs(ctx, input, n, w, lo.data(), hi.data(), (int)ranges.size(), tail, ptrs.data(),
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
    Lease ctx;
    must(sss_host_shared_scan_linear(ctx, input, n, w, lo.data(), hi.data(), (int)ranges.size(), SSS_TAIL_8,
                                     output.data(), counts ? counts->data() : nullptr),
         "linear shared scan");
}

}  // namespace

namespace sss_host {

std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint32_t>& input)
{
    const size_t bytes = sss_packed_bytes(input.size(), w);
    auto buffer = std::make_unique<uint64_t[]>((bytes + 7) / 8);  // zero-initialised, like the reference's
    Lease ctx;
    must(sss_host_pack(ctx, input.data(), input.size(), w, buffer.get()), "compress");
    return buffer;
}

std::unique_ptr<uint64_t[]> compress(int w, const std::vector<uint16_t>& input)
{
    return compress(w, std::vector<uint32_t>(input.begin(), input.end()));
}

void decompress(int w, const void* input, size_t input_size, int* output)
{
    Lease ctx;
    must(sss_host_unpack(ctx, input, input_size, w, output), "decompress");
}

int scan(int w, Tail tail, int predicate_key, const void* input, size_t input_size, std::vector<uint8_t>& output)
{
    const Range r = key_range(w, predicate_key);
    require_bytes(output.size(), sss_tail_out_bytes(input_size, w, tail), "scan");
    uint64_t hits = 0;
    Lease ctx;
    must(sss_host_scan_range(ctx, input, input_size, w, r.lo, r.hi, tail, output.data(), &hits), "scan");
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
    Lease ctx;
    must(sss_host_scan_range(ctx, input, input_size, w, r.lo, r.hi, SSS_TAIL_EXACT, output.data(), &hits),
         "scan_range");
    if (positions) {
        // ascending positions of the set bits (get_bit order, reference src/util.cpp:51-58), expanded on
        // the GPU from the bitvector (n/8 bytes up, 8 bytes per match down)
        positions->assign(hits, 0);
        uint64_t num = 0;
        must(sss_host_positions(ctx, output.data(), input_size, 0, positions->data(), hits, &num), "scan_range positions");
        if (num != hits) throw std::runtime_error("shared_simd_scan_b200: scan_range: position count != match count");
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
