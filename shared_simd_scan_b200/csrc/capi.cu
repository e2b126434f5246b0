// capi.cu -- the C ABI (include/sss_b200.h): context, argument checking, device-resident entry
// points, and the host-buffer entry points that the C++ drop-in layer (include/simd_scan.hpp)
// calls.  Every CUDA return code is checked and surfaced as an int status; there is no CPU path.
#include <cuda_runtime.h>
#include <sched.h>
#include <cctype>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/sss_b200.h"
#include "kernels.h"

using namespace sss;

namespace {

constexpr int kSlots = 3;                       // host<->device staging ring depth
constexpr uint32_t kCountSlots = 1024;          // scans in flight that may count concurrently
constexpr uint64_t kChunkBytes = 64ull << 20;   // target bytes per staged chunk (larger of in/out)

struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr;
    size_t in_cap = 0;
    void* d_out = nullptr;
    size_t out_cap = 0;
    void* h_bounce = nullptr;  // pinned; results on their way to pageable or scattered host destinations
    size_t bounce_cap = 0;
    void* h_in_bounce = nullptr;  // pinned; pageable input on its way to the device
    size_t in_bounce_cap = 0;
};

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? SSS_OK : SSS_ERR_CUDA_BASE + (int)e; }

#define SSS_CK(expr)                                       \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return cuda_status(_e);     \
    } while (0)

#define SSS_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != SSS_OK) return _s;  \
    } while (0)

#define SSS_LOCK(ctx) std::lock_guard<std::recursive_mutex> _guard((ctx)->mu)

}  // namespace

struct sss_ctx {
    Device dev;
    Slot slots[kSlots];
    uint32_t* d_lo = nullptr;  // SSS_MAX_PREDICATES each, for the host entry points
    uint32_t* d_hi = nullptr;
    unsigned long long* d_counts = nullptr;  // SSS_MAX_PREDICATES + 1
    void* d_ws = nullptr;                    // positions workspace (positions.cu: counters + tile states)
    size_t ws_cap = 0;
    uint32_t ws_epoch = 0;                   // last epoch the workspace has seen since it was zeroed
    // Serialises the calls made with this context: the host entry points for their whole (synchronous)
    // duration, the device entry points while they enqueue.  Calls from several threads are therefore
    // safe; to run them CONCURRENTLY give every thread its own context (the C++ drop-in layer keeps a pool).
    std::recursive_mutex mu;
    void* d_index = nullptr;                 // predicate index of the shared scan in flight (predindex.cuh)
    unsigned long long* d_count_slots = nullptr;  // kCountSlots x {accumulator, ticket}, zero between launches
    uint32_t next_count_slot = 0;
};

namespace {

int grow(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return SSS_OK;
    if (*p) SSS_CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    SSS_CK(cudaMalloc(p, need));
    *cap = need;
    return SSS_OK;
}

int grow_pinned(void** p, size_t* cap, size_t need)
{
    if (*cap >= need) return SSS_OK;
    if (*p) SSS_CK(cudaFreeHost(*p));
    *p = nullptr;
    *cap = 0;
    SSS_CK(cudaMallocHost(p, need));
    *cap = need;
    return SSS_OK;
}

int check_common(const sss_ctx* ctx, const void* packed, int w)
{
    if (ctx == nullptr || packed == nullptr) return SSS_ERR_INVALID_ARG;
    if (w < 1 || w > 32) return SSS_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(packed) & 15) return SSS_ERR_INVALID_ARG;
    return SSS_OK;
}

int check_k(int K) { return (K < 1 || K > SSS_MAX_PREDICATES) ? SSS_ERR_INVALID_ARG : SSS_OK; }

// Tail modes (SSS_TAIL_*).  The cursor modes restate how far the reference's word-cursor loops run:
// scan_unvectorized walks ceil(w*n/64) 64-bit words and also evaluates a code that merely STARTS in
// the last word (reference src/simd_scan.cpp:25,39-91), leaving early once `oi == input_size`
// compares its output BYTE index with the code count (:58); shared_scan_*_parallel does the same
// over 32-bit words and 4-byte flushes (reference src/simd_scan_shared.cpp:160,184-266,219).
bool tail_mode_ok(int g)
{
    return g == SSS_TAIL_EXACT || g == SSS_TAIL_8 || g == SSS_TAIL_32 || g == SSS_TAIL_CURSOR32 ||
           g == SSS_TAIL_CURSOR64;
}

uint64_t eval_codes_for(uint64_t n, int w, int g)
{
    if (n == 0) return 0;
    const uint64_t bits = n * (uint64_t)w;
    switch (g) {
        case SSS_TAIL_EXACT: return n;
        case SSS_TAIL_8: return round_up_u64(n, 8);
        case SSS_TAIL_32: return round_up_u64(n, 32);
        case SSS_TAIL_CURSOR64: {
            const uint64_t span = round_up_u64(bits, 64);
            return std::min((span + w - 1) / w, 8 * n);
        }
        case SSS_TAIL_CURSOR32: {
            const uint64_t span = round_up_u64(bits, 32);
            return std::min((span + w - 1) / w, 32 * ((n + 3) / 4));
        }
    }
    return 0;
}

uint64_t out_bytes_for(uint64_t n, int w, int g)
{
    switch (g) {
        case SSS_TAIL_EXACT:
        case SSS_TAIL_8: return (n + 7) / 8;
        case SSS_TAIL_32: return 4 * ((n + 31) / 32);
        case SSS_TAIL_CURSOR64: return (eval_codes_for(n, w, g) + 7) / 8;
        case SSS_TAIL_CURSOR32: return 4 * ((eval_codes_for(n, w, g) + 31) / 32);
    }
    return 0;
}

// codes per staged chunk: a multiple of 1024 codes (=> whole words in, whole words out, 16-byte
// aligned packed offsets for every width)
uint64_t chunk_codes(double in_bytes_per_code, double out_bytes_per_code)
{
    const double per = std::max(in_bytes_per_code, out_bytes_per_code);
    uint64_t c = (uint64_t)((double)kChunkBytes / per);
    c = c / 1024 * 1024;
    return c < 1024 ? 1024 : c;
}


// ---- pageable host memory -----------------------------------------------------------------------
//
// The reference's callers hand in make_unique / std::vector memory (src/simd_scan.hpp:46-120), i.e.
// PAGEABLE buffers.  Measured on the B200 box (profiles/r02_h2d_probe.txt): cudaMemcpy from pageable
// memory 10 GB/s (to: 18 GB/s), pinned 55 GB/s, cudaHostRegister on the fly 7 GB/s to pin plus a slower
// unregister -- it never pays.  What pays is staging through the slots' pinned bounce buffers with a few
// host threads doing the memcpy (8 threads: 34 GB/s end to end, overlapped with the DMA of the
// neighbouring chunks).
int copy_threads()
{
    static const int v = [] {
        const char* e = std::getenv("SSS_COPY_THREADS");
        int t = e ? std::atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
        return t < 1 ? 1 : (t > 32 ? 32 : t);
    }();
    return v;
}

void par_memcpy(void* dst, const void* src, size_t bytes)
{
    const int T = copy_threads();
    if (T <= 1 || bytes < (1u << 20)) {
        std::memcpy(dst, src, bytes);
        return;
    }
    const size_t per = ((bytes + T - 1) / T + 4095) & ~(size_t)4095;
    std::vector<std::thread> team;
    for (int t = 1; t < T; t++) {
        const size_t off = per * t;
        if (off >= bytes) break;
        team.emplace_back([=] {
            std::memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(per, bytes - off));
        });
    }
    std::memcpy(dst, src, std::min(per, bytes));
    for (auto& t : team) t.join();
}

bool is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// The staging pipeline of one host-buffer call: chunk c uses slot c % kSlots (own stream, own device
// buffers, own bounce buffers).  acquire() makes a slot reusable: it waits for the slot's stream and
// then hands the slot's finished results to their pageable / scattered destinations, so the host
// copies of chunk c overlap the DMA and kernels of chunks c+1 and c+2.
class HostPipe {
public:
    struct Piece {
        void* dst;
        size_t src_off, bytes;
    };
    HostPipe(sss_ctx* ctx, bool in_pageable, bool out_staged) : ctx_(ctx), in_pageable_(in_pageable), out_staged_(out_staged) {}
    bool out_staged() const { return out_staged_; }

    int acquire(int sl)
    {
        Slot& s = ctx_->slots[sl];
        if (!busy_[sl]) return SSS_OK;
        // stream order alone protects the device buffers; the host has to wait only before it touches a
        // bounce buffer again (pageable input) or owes the slot's results to somebody
        if (!in_pageable_ && owed_[sl].empty()) return SSS_OK;
        SSS_CK(cudaStreamSynchronize(s.stream));
        for (const Piece& p : owed_[sl])
            par_memcpy(p.dst, static_cast<const uint8_t*>(s.h_bounce) + p.src_off, p.bytes);
        owed_[sl].clear();
        busy_[sl] = false;
        return SSS_OK;
    }
    // host -> slot.d_in (after acquire)
    int copy_in(int sl, const void* src, size_t bytes, size_t cap_hint)
    {
        Slot& s = ctx_->slots[sl];
        busy_[sl] = true;
        if (!in_pageable_) return cuda_status(cudaMemcpyAsync(s.d_in, src, bytes, cudaMemcpyHostToDevice, s.stream));
        SSS_TRY(grow_pinned(&s.h_in_bounce, &s.in_bounce_cap, std::max(bytes, cap_hint)));
        par_memcpy(s.h_in_bounce, src, bytes);
        return cuda_status(cudaMemcpyAsync(s.d_in, s.h_in_bounce, bytes, cudaMemcpyHostToDevice, s.stream));
    }
    // staged results: slot.d_out[0, extent) -> the slot's bounce buffer now, `pieces` of it -> their
    // destinations when the slot is next acquired (or at finish)
    int copy_out_staged(int sl, size_t extent, size_t cap_hint, std::vector<Piece>&& pieces)
    {
        Slot& s = ctx_->slots[sl];
        busy_[sl] = true;
        SSS_TRY(grow_pinned(&s.h_bounce, &s.bounce_cap, std::max(extent, cap_hint)));
        SSS_CK(cudaMemcpyAsync(s.h_bounce, s.d_out, extent, cudaMemcpyDeviceToHost, s.stream));
        owed_[sl] = std::move(pieces);
        return SSS_OK;
    }
    // oldest chunk first
    int finish(int next_slot)
    {
        for (int i = 0; i < kSlots; i++) SSS_TRY(acquire((next_slot + i) % kSlots));
        for (auto& s : ctx_->slots) SSS_CK(cudaStreamSynchronize(s.stream));
        return SSS_OK;
    }

private:
    sss_ctx* ctx_;
    bool in_pageable_, out_staged_;
    bool busy_[kSlots] = {};
    std::vector<Piece> owed_[kSlots];
};

}  // namespace

extern "C" {

int sss_version(void) { return 100; }

const char* sss_error_string(int status)
{
    switch (status) {
        case SSS_OK: return "ok";
        case SSS_ERR_INVALID_ARG: return "invalid argument";
        case SSS_ERR_UNSUPPORTED: return "unsupported";
        case SSS_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    }
    if (status >= SSS_ERR_NCCL_BASE) return "NCCL error (status - 2000 = ncclResult_t)";
    if (status >= SSS_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - SSS_ERR_CUDA_BASE));
    return "unknown status";
}

int sss_ctx_create(int device, sss_ctx** out)
{
    if (out == nullptr) return SSS_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return SSS_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return SSS_ERR_INVALID_ARG;
    SSS_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SSS_CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return SSS_ERR_UNSUPPORTED;  // kernels are sm_100a only
    sss_ctx* ctx = new (std::nothrow) sss_ctx();
    if (ctx == nullptr) return SSS_ERR_CUDA_BASE + (int)cudaErrorMemoryAllocation;
    ctx->dev.ordinal = device;
    ctx->dev.sm_count = prop.multiProcessorCount;
    ctx->dev.smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* e = std::getenv("SSS_SCAN_DIRECT_TILES")) ctx->dev.scan_direct_tiles = (uint32_t)std::atoi(e);
    if (const char* e = std::getenv("SSS_PDL")) ctx->dev.pdl = std::atoi(e) != 0;
    for (auto& s : ctx->slots) {
        cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            sss_ctx_destroy(ctx);
            return cuda_status(e);
        }
    }
    cudaError_t e = cudaMalloc(&ctx->d_lo, SSS_MAX_PREDICATES * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_hi, SSS_MAX_PREDICATES * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counts, (SSS_MAX_PREDICATES + 1) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_index, pred_index_bytes());
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dev.ticket_pool, kTicketSlots * 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->dev.ticket_pool, 0, kTicketSlots * 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_count_slots, kCountSlots * 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_count_slots, 0, kCountSlots * 2 * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        sss_ctx_destroy(ctx);
        return cuda_status(e);
    }
    *out = ctx;
    return SSS_OK;
}

int sss_ctx_destroy(sss_ctx* ctx)
{
    if (ctx == nullptr) return SSS_OK;
    cudaSetDevice(ctx->dev.ordinal);
    for (auto& s : ctx->slots) {
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_bounce) cudaFreeHost(s.h_bounce);
        if (s.h_in_bounce) cudaFreeHost(s.h_in_bounce);
    }
    if (ctx->d_lo) cudaFree(ctx->d_lo);
    if (ctx->d_hi) cudaFree(ctx->d_hi);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->d_ws) cudaFree(ctx->d_ws);
    if (ctx->d_index) cudaFree(ctx->d_index);
    if (ctx->d_count_slots) cudaFree(ctx->d_count_slots);
    if (ctx->dev.ticket_pool) cudaFree(ctx->dev.ticket_pool);
    delete ctx;
    return SSS_OK;
}

int sss_ctx_sm_count(const sss_ctx* ctx) { return ctx ? ctx->dev.sm_count : 0; }

size_t sss_packed_bytes(uint64_t n, int w) { return (size_t)(packed_payload_bytes(n, w) + 256); }
size_t sss_unpack_bytes(uint64_t n) { return (size_t)(n * 4 + 32); }
size_t sss_bitvec_bytes(uint64_t n) { return (size_t)((n + 7) / 8 + 32); }

uint64_t sss_tail_eval_codes(uint64_t n, int w, int tail_granule)
{
    return (w < 1 || w > 32 || !tail_mode_ok(tail_granule)) ? 0 : eval_codes_for(n, w, tail_granule);
}

uint64_t sss_tail_out_bytes(uint64_t n, int w, int tail_granule)
{
    return (w < 1 || w > 32 || !tail_mode_ok(tail_granule)) ? 0 : out_bytes_for(n, w, tail_granule);
}

// ---- device-resident entry points ----------------------------------------------------------

int sss_pack(sss_ctx* ctx, const uint32_t* d_codes, uint64_t n, int w, void* d_packed, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    if (d_codes == nullptr && n != 0) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const size_t bytes = sss_packed_bytes(n, w);
    const uint64_t words = bytes / 4;
    auto st = static_cast<cudaStream_t>(stream);
    if (bytes % 4) SSS_CK(cudaMemsetAsync(static_cast<uint8_t*>(d_packed) + words * 4, 0, bytes % 4, st));
    return cuda_status(launch_pack(ctx->dev, d_codes, n, w, static_cast<uint32_t*>(d_packed), words, st));
}

int sss_generate_packed(sss_ctx* ctx, void* d_packed, uint64_t n, int w, uint64_t seed, uint64_t first_index,
                        void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const size_t bytes = sss_packed_bytes(n, w);
    const uint64_t words = bytes / 4;
    auto st = static_cast<cudaStream_t>(stream);
    if (bytes % 4) SSS_CK(cudaMemsetAsync(static_cast<uint8_t*>(d_packed) + words * 4, 0, bytes % 4, st));
    return cuda_status(
        launch_generate_packed(ctx->dev, static_cast<uint32_t*>(d_packed), words, n, w, seed, first_index, st));
}

int sss_unpack(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, int32_t* d_out, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    if (d_out == nullptr || (reinterpret_cast<uintptr_t>(d_out) & 31)) return SSS_ERR_INVALID_ARG;  // 256-bit stores
    if (n == 0) return SSS_OK;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return cuda_status(launch_unpack(ctx->dev, d_packed, n, w, d_out, static_cast<cudaStream_t>(stream)));
}

static int scan_range_impl(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                           void* d_bitvec, unsigned long long* d_count, bool accumulate, cudaStream_t st,
                           uint32_t* d_seg_count = nullptr)
{
    const AlignedRange r = aligned_range(w, lo, hi);
    if (n == 0 || r.empty) {  // nothing can match: clear what a scan would have written
        if (d_count && !accumulate) SSS_CK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
        if (n != 0) SSS_CK(cudaMemsetAsync(d_bitvec, 0, 4 * ((n + 31) / 32), st));
        return SSS_OK;
    }
    CountTarget ct{d_count, ctx->d_count_slots + 2 * (ctx->next_count_slot++ % kCountSlots), accumulate, d_seg_count};
    return cuda_status(launch_scan_range(ctx->dev, d_packed, n, w, r, static_cast<uint32_t*>(d_bitvec), ct, st));
}

int sss_scan_range(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, uint32_t lo, uint32_t hi, void* d_bitvec,
                   uint64_t* d_count, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    if (d_bitvec == nullptr || (reinterpret_cast<uintptr_t>(d_bitvec) & 15)) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return scan_range_impl(ctx, d_packed, n, w, lo, hi, d_bitvec, reinterpret_cast<unsigned long long*>(d_count), false,
                           static_cast<cudaStream_t>(stream));
}

static int positions_workspace(sss_ctx* ctx, uint64_t n, cudaStream_t st)
{
    const size_t need = positions_workspace_bytes(n);
    if (ctx->ws_cap < need || ctx->ws_epoch >= positions_epoch_limit()) {
        // (re)allocating or re-zeroing the workspace synchronises; callers that time this path warm it up first
        SSS_CK(cudaStreamSynchronize(st));
        SSS_TRY(grow(&ctx->d_ws, &ctx->ws_cap, need + need / 4));
        SSS_CK(cudaMemsetAsync(ctx->d_ws, 0, ctx->ws_cap, st));
        ctx->ws_epoch = 0;
    }
    return SSS_OK;
}

static int positions_impl(sss_ctx* ctx, const void* d_bitvec, uint64_t n, uint64_t base, uint64_t* d_positions,
                          uint64_t capacity, uint64_t* d_num, cudaStream_t st, bool counted = false)
{
    SSS_TRY(positions_workspace(ctx, n, st));
    return cuda_status(launch_positions(ctx->dev, static_cast<const uint32_t*>(d_bitvec), n, base,
                                        reinterpret_cast<unsigned long long*>(d_positions), capacity,
                                        reinterpret_cast<unsigned long long*>(d_num), ctx->d_ws, ++ctx->ws_epoch, st,
                                        counted));
}

int sss_positions(sss_ctx* ctx, const void* d_bitvec, uint64_t n, uint64_t base, uint64_t* d_positions,
                  uint64_t capacity, uint64_t* d_num, void* stream)
{
    if (ctx == nullptr || d_bitvec == nullptr || (reinterpret_cast<uintptr_t>(d_bitvec) & 3)) return SSS_ERR_INVALID_ARG;
    if (d_positions == nullptr && capacity != 0) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return positions_impl(ctx, d_bitvec, n, base, d_positions, capacity, d_num, static_cast<cudaStream_t>(stream));
}

int sss_scan_range_positions(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                             void* d_bitvec, uint64_t* d_count, uint64_t base, uint64_t* d_positions,
                             uint64_t capacity, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    if (d_bitvec == nullptr || (reinterpret_cast<uintptr_t>(d_bitvec) & 15)) return SSS_ERR_INVALID_ARG;
    if (d_positions == nullptr && capacity != 0) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    auto st = static_cast<cudaStream_t>(stream);
    // The scan writes the bitvector, the count AND the matches of every 4 KB segment of the bitvector (one
    // warp-aggregated atomic per warp and segment into the position workspace), so the position list starts
    // at its second step: group offsets, then the ordered expansion, which reads the bitvector while it is
    // still (mostly) L2-resident.  An empty predicate writes zeros without a scan: the count pass then runs.
    const bool fused = n != 0 && !aligned_range(w, lo, hi).empty && capacity != 0;
    uint32_t* seg = nullptr;
    if (fused) {
        SSS_TRY(positions_workspace(ctx, n, st));
        seg = positions_seg_counts(ctx->d_ws, n);
        SSS_CK(cudaMemsetAsync(seg, 0, positions_seg_count_bytes(n), st));
    }
    SSS_TRY(scan_range_impl(ctx, d_packed, n, w, lo, hi, d_bitvec, reinterpret_cast<unsigned long long*>(d_count), false, st,
                            seg));
    return positions_impl(ctx, d_bitvec, n, base, d_positions, capacity, nullptr, st, fused);
}

// One shared scan = build the predicate index, run the scan (which fills the index's interval
// histogram), turn the histogram into per-predicate counts.  The index lives in the context, so a
// context serves one shared scan at a time (calls on one stream serialise naturally).
int sss_bitvec_combine(sss_ctx* ctx, int op, const void* d_a, const void* d_b, uint64_t n, void* d_out,
                       uint64_t* d_count, void* stream)
{
    if (ctx == nullptr || d_a == nullptr || d_b == nullptr || d_out == nullptr) return SSS_ERR_INVALID_ARG;
    if (op < SSS_OP_AND || op > SSS_OP_XOR) return SSS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(d_a) | reinterpret_cast<uintptr_t>(d_b) | reinterpret_cast<uintptr_t>(d_out)) & 15)
        return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    auto st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (d_count) SSS_CK(cudaMemsetAsync(d_count, 0, sizeof(uint64_t), st));
        return SSS_OK;
    }
    CountTarget ct{reinterpret_cast<unsigned long long*>(d_count),
                   ctx->d_count_slots + 2 * (ctx->next_count_slot++ % kCountSlots), false};
    return cuda_status(launch_bitvec_combine(ctx->dev, op, static_cast<const uint32_t*>(d_a),
                                             static_cast<const uint32_t*>(d_b), n, static_cast<uint32_t*>(d_out), ct, st));
}

int sss_unpack_at(sss_ctx* ctx, const void* d_packed, int w, const uint64_t* d_positions, uint64_t num, int32_t* d_out,
                  void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    if (num != 0 && (d_positions == nullptr || d_out == nullptr)) return SSS_ERR_INVALID_ARG;
    if (num == 0) return SSS_OK;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return cuda_status(launch_unpack_at(ctx->dev, d_packed, w, reinterpret_cast<const unsigned long long*>(d_positions),
                                        num, d_out, static_cast<cudaStream_t>(stream)));
}

int sss_calibrate(sss_ctx* ctx, int mode, const void* d_a, void* d_b, uint64_t bytes, uint32_t ratio, void* stream)
{
    if (ctx == nullptr || mode < SSS_CALIB_READ || mode > SSS_CALIB_STREAMS128) return SSS_ERR_INVALID_ARG;
    if (mode != SSS_CALIB_WRITE && mode < SSS_CALIB_STREAMS32 && (d_a == nullptr || (reinterpret_cast<uintptr_t>(d_a) & 15))) return SSS_ERR_INVALID_ARG;
    if (mode != SSS_CALIB_READ && (d_b == nullptr || (reinterpret_cast<uintptr_t>(d_b) & 15))) return SSS_ERR_INVALID_ARG;
    if (mode == SSS_CALIB_MIX && (ratio == 0 || ratio > 64)) return SSS_ERR_INVALID_ARG;
    if (mode >= SSS_CALIB_STREAMS32 && (ratio == 0 || ratio > SSS_MAX_PREDICATES)) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return cuda_status(launch_calibrate(ctx->dev, mode, d_a, d_b, bytes, ratio, ctx->d_counts + SSS_MAX_PREDICATES,
                                        static_cast<cudaStream_t>(stream)));
}

// true when the scan kernel itself adds the match counts to `d_counts` (no histogram epilogue)
static bool counts_directly(bool, int K) { return K <= kDirectCountMaxK; }

// One predicate: the shared scan IS the single-predicate scan (one bitvector, one count; the linear
// layout of K = 1 is the bitvector too), run with its bounds read from device memory.  No index.
static int single_predicate_body(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, const uint32_t* d_lo,
                                 const uint32_t* d_hi, void* d_out, unsigned long long* d_count, bool accumulate,
                                 cudaStream_t st)
{
    if (n == 0) {
        if (d_count && !accumulate) SSS_CK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
        return SSS_OK;
    }
    CountTarget ct{d_count, ctx->d_count_slots + 2 * (ctx->next_count_slot++ % kCountSlots), accumulate};
    return cuda_status(launch_scan_range_dev(ctx->dev, d_packed, n, w, d_lo, d_hi, static_cast<uint32_t*>(d_out), ct, st));
}

static int shared_scan_body(sss_ctx* ctx, bool linear, const void* d_packed, uint64_t n, int w, int K, void* d_out,
                            uint64_t stride_bytes, unsigned long long* d_counts, cudaStream_t st)
{
    if (n == 0) return SSS_OK;
    if (K <= kNarrowMaxK && w <= kFewPredicatesMaxW)  // few predicates: group-table / mask-bucket kernels
        return cuda_status(launch_shared_scan_narrow(ctx->dev, d_packed, n, w, ctx->d_index, K, d_out, stride_bytes / 4,
                                                     linear, d_counts, st));
    if (linear)
        return cuda_status(launch_shared_scan_linear(ctx->dev, d_packed, n, w, ctx->d_index, K,
                                                     static_cast<uint8_t*>(d_out), d_counts, d_counts != nullptr, st));
    return cuda_status(launch_shared_scan(ctx->dev, d_packed, n, w, ctx->d_index, K, static_cast<uint32_t*>(d_out),
                                          stride_bytes / 4, d_counts, d_counts != nullptr, st));
}

static int shared_impl(sss_ctx* ctx, bool linear, const void* d_packed, uint64_t n, int w, const uint32_t* d_lo,
                       const uint32_t* d_hi, int K, void* d_out, uint64_t stride_bytes, unsigned long long* d_counts,
                       cudaStream_t st)
{
    if (K == 1 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0)  // (the scan kernel stores 128-bit pieces)
        return single_predicate_body(ctx, d_packed, n, w, d_lo, d_hi, d_out, d_counts, false, st);
    if (d_counts) SSS_CK(cudaMemsetAsync(d_counts, 0, (size_t)K * sizeof(unsigned long long), st));
    // bitvector layout: predicates in groups whose row table fits in shared memory (shared.cu); the
    // index buffer is reused group after group in stream order
    const int G = linear ? K : shared_scan_predicate_group(w, K);
    for (int k0 = 0; k0 < K; k0 += G) {
        const int Kg = std::min(G, K - k0);
        void* out_g = linear ? d_out : static_cast<void*>(static_cast<uint8_t*>(d_out) + (uint64_t)k0 * stride_bytes);
        unsigned long long* counts_g = d_counts ? d_counts + k0 : nullptr;
        SSS_CK(launch_build_pred_index(d_lo + k0, d_hi + k0, Kg, w, ctx->d_index, st));
        SSS_TRY(shared_scan_body(ctx, linear, d_packed, n, w, Kg, out_g, stride_bytes, counts_g, st));
        if (counts_g && !counts_directly(linear, Kg))
            SSS_CK(launch_counts_from_histogram(ctx->d_index, Kg, counts_g, st));
    }
    return SSS_OK;
}

int sss_shared_scan(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, const uint32_t* d_lo, const uint32_t* d_hi,
                    int K, void* d_bitvecs, uint64_t stride_bytes, uint64_t* d_counts, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    SSS_TRY(check_k(K));
    if (d_lo == nullptr || d_hi == nullptr || d_bitvecs == nullptr) return SSS_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(d_bitvecs) & 3) || (stride_bytes & 3) || stride_bytes < 4 * ((n + 31) / 32))
        return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return shared_impl(ctx, false, d_packed, n, w, d_lo, d_hi, K, d_bitvecs, stride_bytes,
                       reinterpret_cast<unsigned long long*>(d_counts), static_cast<cudaStream_t>(stream));
}

int sss_shared_scan_linear(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, const uint32_t* d_lo,
                           const uint32_t* d_hi, int K, void* d_out, uint64_t* d_counts, void* stream)
{
    SSS_TRY(check_common(ctx, d_packed, w));
    SSS_TRY(check_k(K));
    if (d_lo == nullptr || d_hi == nullptr || d_out == nullptr) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    return shared_impl(ctx, true, d_packed, n, w, d_lo, d_hi, K, d_out, 0,
                       reinterpret_cast<unsigned long long*>(d_counts), static_cast<cudaStream_t>(stream));
}

// ---- host-buffer entry points ----------------------------------------------------------------
//
// The column is cut into chunks of a multiple of 1024 codes.  Chunk c uses staging slot c % 3
// (own stream, own device buffers): H2D copy, kernel and D2H copy are enqueued on that stream, so
// the copies of neighbouring chunks overlap the kernel and a slot is never reused before its
// previous D2H has finished (same-stream ordering).

int sss_host_unpack(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, int32_t* h_out)
{
    if (ctx == nullptr || h_packed == nullptr || h_out == nullptr || w < 1 || w > 32) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const uint64_t C = chunk_codes(w / 8.0, 4.0);
    const uint64_t Ca = std::min<uint64_t>(C, round_up_u64(n, 1024));
    const size_t in_cap = round_up_u64(packed_payload_bytes(Ca, w), 16) + 16;
    HostPipe pipe(ctx, is_pageable(h_packed), is_pageable(h_out));
    int slot = 0;
    for (uint64_t first = 0; first < n; first += C, slot = (slot + 1) % kSlots) {
        Slot& s = ctx->slots[slot];
        SSS_TRY(pipe.acquire(slot));
        const uint64_t cnt = std::min(C, n - first);
        const uint64_t in_bytes = round_up_u64(packed_payload_bytes(cnt, w), 16);
        SSS_TRY(grow(&s.d_in, &s.in_cap, in_cap));
        SSS_TRY(grow(&s.d_out, &s.out_cap, Ca * 4));
        const uint8_t* src = static_cast<const uint8_t*>(h_packed) + first * (uint64_t)w / 8;
        SSS_TRY(pipe.copy_in(slot, src, in_bytes, in_cap));
        SSS_CK(launch_unpack(ctx->dev, s.d_in, cnt, w, static_cast<int32_t*>(s.d_out), s.stream));
        if (pipe.out_staged())
            SSS_TRY(pipe.copy_out_staged(slot, cnt * 4, Ca * 4, {{h_out + first, 0, cnt * 4}}));
        else
            SSS_CK(cudaMemcpyAsync(h_out + first, s.d_out, cnt * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    return pipe.finish(slot);
}

int sss_host_scan_range(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                        int tail_granule, void* h_bitvec, uint64_t* h_count)
{
    if (ctx == nullptr || h_packed == nullptr || h_bitvec == nullptr || w < 1 || w > 32) return SSS_ERR_INVALID_ARG;
    if (!tail_mode_ok(tail_granule)) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const uint64_t n_eval = eval_codes_for(n, w, tail_granule);
    const uint64_t out_total = out_bytes_for(n, w, tail_granule);
    const uint64_t C = chunk_codes(w / 8.0, 1 / 8.0);
    const uint64_t Ca = std::min<uint64_t>(C, round_up_u64(n_eval, 1024));
    const size_t in_cap = round_up_u64(Ca / 32 * (uint64_t)w * 4, 16) + 16;
    SSS_CK(cudaMemsetAsync(ctx->d_counts, 0, sizeof(unsigned long long), ctx->slots[0].stream));
    SSS_CK(cudaStreamSynchronize(ctx->slots[0].stream));
    // A pageable bitvector goes through the slot's bounce buffer too, small as it is: cudaMemcpyAsync into
    // pageable memory returns only when the copy is done -- i.e. after the chunk's H2D copy and kernel --, which
    // serialised the host's staging memcpy of chunk c+1 with the DMA of chunk c (16 GB/s whatever the
    // number of copy threads, tools/pageable_probe.py).
    HostPipe pipe(ctx, is_pageable(h_packed), is_pageable(h_bitvec));
    int slot = 0;
    for (uint64_t first = 0; first < n_eval; first += C, slot = (slot + 1) % kSlots) {
        Slot& s = ctx->slots[slot];
        SSS_TRY(pipe.acquire(slot));
        const uint64_t cnt = std::min(C, n_eval - first);
        const uint64_t in_bytes = round_up_u64(((cnt + 31) / 32) * (uint64_t)w * 4, 16);
        SSS_TRY(grow(&s.d_in, &s.in_cap, in_cap));
        SSS_TRY(grow(&s.d_out, &s.out_cap, Ca / 8 + 16));
        const uint8_t* src = static_cast<const uint8_t*>(h_packed) + first * (uint64_t)w / 8;
        SSS_TRY(pipe.copy_in(slot, src, in_bytes, in_cap));
        SSS_TRY(scan_range_impl(ctx, s.d_in, cnt, w, lo, hi, s.d_out, ctx->d_counts, true, s.stream));
        const uint64_t out_off = first / 8;
        const uint64_t out_bytes = std::min<uint64_t>(4 * ((cnt + 31) / 32), out_total - out_off);
        uint8_t* dst = static_cast<uint8_t*>(h_bitvec) + out_off;
        if (pipe.out_staged())
            SSS_TRY(pipe.copy_out_staged(slot, out_bytes, Ca / 8 + 16, {{dst, 0, (size_t)out_bytes}}));
        else
            SSS_CK(cudaMemcpyAsync(dst, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    SSS_TRY(pipe.finish(slot));
    if (h_count) SSS_CK(cudaMemcpy(h_count, ctx->d_counts, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return SSS_OK;
}

static int host_shared_impl(sss_ctx* ctx, bool linear, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                            const uint32_t* h_hi, int K, int tail_granule, void* h_out, void* const* h_out_ptrs,
                            uint64_t stride_bytes, uint64_t* h_counts)
{
    if (ctx == nullptr || h_packed == nullptr || (h_out == nullptr && h_out_ptrs == nullptr) || h_lo == nullptr ||
        h_hi == nullptr)
        return SSS_ERR_INVALID_ARG;
    if (w < 1 || w > 32) return SSS_ERR_INVALID_ARG;
    SSS_TRY(check_k(K));
    if (!tail_mode_ok(tail_granule)) return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const uint64_t n_eval = eval_codes_for(n, w, tail_granule);
    const uint64_t out_total = out_bytes_for(n, w, tail_granule);  // per predicate (bitvectors) / rows (linear)
    if (!linear && h_out_ptrs == nullptr && stride_bytes < out_total) return SSS_ERR_INVALID_ARG;
    const uint64_t C = chunk_codes(w / 8.0, K / 8.0);
    cudaStream_t s0 = ctx->slots[0].stream;
    SSS_CK(cudaMemcpyAsync(ctx->d_lo, h_lo, (size_t)K * 4, cudaMemcpyHostToDevice, s0));
    SSS_CK(cudaMemcpyAsync(ctx->d_hi, h_hi, (size_t)K * 4, cudaMemcpyHostToDevice, s0));
    SSS_CK(cudaMemsetAsync(ctx->d_counts, 0, (size_t)K * sizeof(unsigned long long), s0));
    const bool single = K == 1;  // see single_predicate_body
    if (!single) SSS_CK(launch_build_pred_index(ctx->d_lo, ctx->d_hi, K, w, ctx->d_index, s0));
    SSS_CK(cudaStreamSynchronize(s0));
    // Results go through the slot's pinned bounce buffer -- one device->host copy per chunk, then a
    // host-side scatter once the chunk has landed -- when the destination is pageable, and always for
    // per-predicate destination vectors (h_out_ptrs): K small copies per chunk (1024 x 8 KB at K = 1024)
    // into pageable memory are what the reference's signature would otherwise cost.
    const bool staged = h_out_ptrs != nullptr || is_pageable(h_out);
    HostPipe pipe(ctx, is_pageable(h_packed), staged);
    const uint64_t chunk_stride = C / 8;  // device-side bitvector stride inside the slot (full chunks)
    const uint64_t Ca = std::min<uint64_t>(C, round_up_u64(n_eval, 1024));  // codes the staging buffers are sized for
    const size_t in_cap = round_up_u64(Ca / 32 * (uint64_t)w * 4, 16) + 16;
    const size_t out_cap = (Ca / 8 + 128) * (uint64_t)K;
    int slot = 0;
    for (uint64_t first = 0; first < n_eval; first += C, slot = (slot + 1) % kSlots) {
        Slot& s = ctx->slots[slot];
        SSS_TRY(pipe.acquire(slot));  // the slot's previous chunk is delivered before its buffers are reused
        const uint64_t cnt = std::min(C, n_eval - first);
        const uint64_t in_bytes = round_up_u64(((cnt + 31) / 32) * (uint64_t)w * 4, 16);
        SSS_TRY(grow(&s.d_in, &s.in_cap, in_cap));
        SSS_TRY(grow(&s.d_out, &s.out_cap, out_cap));
        const uint8_t* src = static_cast<const uint8_t*>(h_packed) + first * (uint64_t)w / 8;
        SSS_TRY(pipe.copy_in(slot, src, in_bytes, in_cap));
        // a short (last) chunk packs its bitvectors tightly so that small columns do not move whole-chunk strides
        const uint64_t dev_stride = std::min(chunk_stride, round_up_u64(4 * ((cnt + 31) / 32), 128));
        if (single)
            SSS_TRY(single_predicate_body(ctx, s.d_in, cnt, w, ctx->d_lo, ctx->d_hi, s.d_out,
                                          h_counts != nullptr ? ctx->d_counts : nullptr, true, s.stream));
        else
            SSS_TRY(shared_scan_body(ctx, linear, s.d_in, cnt, w, K, s.d_out, dev_stride,
                                     h_counts != nullptr ? ctx->d_counts : nullptr, s.stream));
        const uint64_t out_off = first / 8;  // bytes per predicate / rows already produced
        const uint64_t piece = std::min<uint64_t>(4 * ((cnt + 31) / 32), out_total - out_off);
        if (linear) {
            uint8_t* dst = static_cast<uint8_t*>(h_out) + out_off * (uint64_t)K;
            if (staged)
                SSS_TRY(pipe.copy_out_staged(slot, piece * (uint64_t)K, out_cap, {{dst, 0, piece * (uint64_t)K}}));
            else
                SSS_CK(cudaMemcpyAsync(dst, s.d_out, piece * (uint64_t)K, cudaMemcpyDeviceToHost, s.stream));
        } else if (staged) {
            std::vector<HostPipe::Piece> pieces((size_t)K);
            for (int k = 0; k < K; k++) {
                uint8_t* base = h_out_ptrs != nullptr ? static_cast<uint8_t*>(h_out_ptrs[k])
                                                      : static_cast<uint8_t*>(h_out) + (uint64_t)k * stride_bytes;
                pieces[k] = HostPipe::Piece{base + out_off, (size_t)((uint64_t)k * dev_stride), (size_t)piece};
            }
            SSS_TRY(pipe.copy_out_staged(slot, (uint64_t)(K - 1) * dev_stride + piece, out_cap, std::move(pieces)));
        } else {
            SSS_CK(cudaMemcpy2DAsync(static_cast<uint8_t*>(h_out) + out_off, stride_bytes, s.d_out, dev_stride, piece,
                                     (size_t)K, cudaMemcpyDeviceToHost, s.stream));
        }
    }
    SSS_TRY(pipe.finish(slot));
    if (h_counts) {
        if (!single && !counts_directly(linear, K))
            SSS_CK(launch_counts_from_histogram(ctx->d_index, K, ctx->d_counts, s0));
        SSS_CK(cudaMemcpyAsync(h_counts, ctx->d_counts, (size_t)K * sizeof(uint64_t), cudaMemcpyDeviceToHost, s0));
        SSS_CK(cudaStreamSynchronize(s0));
    }
    return SSS_OK;
}

int sss_host_shared_scan(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                         const uint32_t* h_hi, int K, int tail_granule, void* h_bitvecs, uint64_t stride_bytes,
                         uint64_t* h_counts)
{
    return host_shared_impl(ctx, false, h_packed, n, w, h_lo, h_hi, K, tail_granule, h_bitvecs, nullptr, stride_bytes,
                            h_counts);
}

int sss_host_shared_scan_ptrs(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                              const uint32_t* h_hi, int K, int tail_granule, void* const* h_bitvec_ptrs,
                              uint64_t* h_counts)
{
    if (h_bitvec_ptrs == nullptr) return SSS_ERR_INVALID_ARG;
    for (int k = 0; k < K && k < SSS_MAX_PREDICATES; k++)
        if (h_bitvec_ptrs[k] == nullptr) return SSS_ERR_INVALID_ARG;
    return host_shared_impl(ctx, false, h_packed, n, w, h_lo, h_hi, K, tail_granule, nullptr, h_bitvec_ptrs, 0,
                            h_counts);
}

int sss_host_shared_scan_linear(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                                const uint32_t* h_hi, int K, int tail_granule, void* h_out, uint64_t* h_counts)
{
    return host_shared_impl(ctx, true, h_packed, n, w, h_lo, h_hi, K, tail_granule, h_out, nullptr, 0, h_counts);
}

// Position list of a HOST bitvector (the result of sss_host_scan_range): the bitvector goes to the
// device in pieces of 2^25 bits, every piece is expanded there (positions.cu) and its positions come
// back appended to h_positions -- n/8 bytes up, 8 bytes per match down.  One stream: the kernels share
// the context's look-back workspace and the running offset needs every piece's count anyway.
int sss_host_positions(sss_ctx* ctx, const void* h_bitvec, uint64_t n, uint64_t base, uint64_t* h_positions,
                       uint64_t capacity, uint64_t* h_num)
{
    if (ctx == nullptr || (h_bitvec == nullptr && n != 0) || (h_positions == nullptr && capacity != 0))
        return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    constexpr uint64_t kPiece = 1ull << 25;  // codes per piece: 4 MiB of bitvector, at most 256 MiB of positions
    Slot& s = ctx->slots[0];
    uint64_t total = 0;
    for (uint64_t first = 0; first < n; first += kPiece) {
        const uint64_t cnt = std::min(kPiece, n - first);
        const uint64_t room = total < capacity ? capacity - total : 0;
        const uint64_t cap_piece = std::min(room, cnt);
        SSS_TRY(grow(&s.d_in, &s.in_cap, kPiece / 8 + 16));
        SSS_TRY(grow(&s.d_out, &s.out_cap, std::max<uint64_t>(cap_piece, 1) * 8));
        const uint64_t in_bytes = (cnt + 7) / 8;
        SSS_CK(cudaMemsetAsync(static_cast<uint8_t*>(s.d_in) + in_bytes / 4 * 4, 0, 8, s.stream));  // ragged last word
        SSS_CK(cudaMemcpyAsync(s.d_in, static_cast<const uint8_t*>(h_bitvec) + first / 8, in_bytes, cudaMemcpyHostToDevice,
                               s.stream));
        SSS_TRY(positions_impl(ctx, s.d_in, cnt, base + first, static_cast<uint64_t*>(s.d_out), cap_piece,
                               reinterpret_cast<uint64_t*>(ctx->d_counts), s.stream));
        unsigned long long got = 0;
        SSS_CK(cudaMemcpyAsync(&got, ctx->d_counts, sizeof(got), cudaMemcpyDeviceToHost, s.stream));
        SSS_CK(cudaStreamSynchronize(s.stream));
        const uint64_t take = std::min<uint64_t>(got, cap_piece);
        if (take)
            SSS_CK(cudaMemcpyAsync(h_positions + total, s.d_out, take * 8, cudaMemcpyDeviceToHost, s.stream));
        total += got;
    }
    SSS_CK(cudaStreamSynchronize(s.stream));
    if (h_num) *h_num = total;
    return SSS_OK;
}

int sss_host_pack(sss_ctx* ctx, const uint32_t* h_codes, uint64_t n, int w, void* h_packed)
{
    if (ctx == nullptr || h_packed == nullptr || (h_codes == nullptr && n != 0) || w < 1 || w > 32)
        return SSS_ERR_INVALID_ARG;
    SSS_LOCK(ctx);
    SSS_CK(cudaSetDevice(ctx->dev.ordinal));
    const size_t total = sss_packed_bytes(n, w);
    std::memset(static_cast<uint8_t*>(h_packed) + packed_payload_bytes(n, w), 0, total - packed_payload_bytes(n, w));
    const uint64_t C = chunk_codes(4.0, w / 8.0);
    const uint64_t Ca = std::min<uint64_t>(C, round_up_u64(n, 1024));
    const size_t out_cap = Ca / 32 * (uint64_t)w * 4 + 16;
    HostPipe pipe(ctx, n != 0 && is_pageable(h_codes), is_pageable(h_packed));
    int slot = 0;
    for (uint64_t first = 0; first < n; first += C, slot = (slot + 1) % kSlots) {
        Slot& s = ctx->slots[slot];
        SSS_TRY(pipe.acquire(slot));
        const uint64_t cnt = std::min(C, n - first);
        const uint64_t words = (cnt * (uint64_t)w + 31) / 32;
        SSS_TRY(grow(&s.d_in, &s.in_cap, Ca * 4));
        SSS_TRY(grow(&s.d_out, &s.out_cap, out_cap));
        SSS_TRY(pipe.copy_in(slot, h_codes + first, cnt * 4, Ca * 4));
        SSS_CK(launch_pack(ctx->dev, static_cast<const uint32_t*>(s.d_in), cnt, w, static_cast<uint32_t*>(s.d_out), words,
                           s.stream));
        const uint64_t off = first * (uint64_t)w / 8;
        const uint64_t bytes = std::min<uint64_t>(words * 4, total - off);
        uint8_t* dst = static_cast<uint8_t*>(h_packed) + off;
        if (pipe.out_staged())
            SSS_TRY(pipe.copy_out_staged(slot, bytes, out_cap, {{dst, 0, (size_t)bytes}}));
        else
            SSS_CK(cudaMemcpyAsync(dst, s.d_out, bytes, cudaMemcpyDeviceToHost, s.stream));
    }
    return pipe.finish(slot);
}

void* sss_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}

void sss_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int sss_bind_host_thread(int device, int* n_cpus)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return SSS_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return SSS_ERR_INVALID_ARG;
    char bdf[32] = {0};
    SSS_CK(cudaDeviceGetPCIBusId(bdf, sizeof bdf, device));
    for (char* c = bdf; *c; c++) *c = (char)tolower(*c);
    cpu_set_t now;
    CPU_ZERO(&now);
    sched_getaffinity(0, sizeof now, &now);
    char path[128], list[4096] = {0};
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf);
    if (FILE* f = fopen(path, "r")) {
        const size_t k = fread(list, 1, sizeof list - 1, f);
        list[k] = 0;
        fclose(f);
    }
    // "0-15,32-47" -> the CPUs of that list this thread is allowed on already (a container's cpuset stays in force)
    cpu_set_t want;
    CPU_ZERO(&want);
    int n = 0;
    for (const char* p = list; *p && *p != '\n';) {
        char* end = nullptr;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) break;
        p = end;
        if (*p == '-') {
            b = strtol(p + 1, &end, 10);
            p = end;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &now)) {
                CPU_SET(c, &want);
                n++;
            }
        if (*p == ',') p++;
    }
    if (n > 0 && sched_setaffinity(0, sizeof want, &want) == 0) {
        if (n_cpus) *n_cpus = n;
        return SSS_OK;
    }
    if (n_cpus) *n_cpus = CPU_COUNT(&now);
    return SSS_OK;
}

int sss_host_register(void* p, size_t bytes) { return cuda_status(cudaHostRegister(p, bytes, cudaHostRegisterDefault)); }
int sss_host_unregister(void* p) { return cuda_status(cudaHostUnregister(p)); }

}  // extern "C"
