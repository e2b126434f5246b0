// mgpu.cu -- the multi-GPU entry points of the C ABI (include/sss_b200.h, sss_mgpu_*).
//
// The reference is single-process, single-socket code; its only parallel variant partitions the
// PREDICATES over OpenMP threads and re-reads the column once per predicate
// (src/simd_scan_shared.cpp:25-32).  Codes are independent, so here the COLUMN is range-partitioned
// (SURVEY.md section 8e): one process drives N GPUs, shard g = codes [first_g, first_g + count_g) with
// first_g a multiple of 128 codes (16-byte aligned in the packed stream for any width and in every
// bitvector), every GPU scans its shard with the single-GPU kernels on its own stream, the per-query
// match counts are combined with ONE ncclAllReduce(sum, uint64 x K) over NVLink, and position lists
// are rebased by first_g -- concatenated in device order they are globally ascending.  No column
// data ever crosses GPUs.
//
// Built on the public single-GPU ABI only (one sss_ctx per device).  NCCL is loaded at run time
// (dlopen), preferring a copy that is already in the process (torch brings its own), so that
// libsss_b200.so neither pins an NCCL version nor needs one when a single GPU is used.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/sss_b200.h"

namespace {

struct NcclApi {
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool ok = false;
};

const NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process?
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return a;
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(h, "ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? SSS_OK : SSS_ERR_CUDA_BASE + (int)e; }
inline int nccl_status(ncclResult_t r) { return r == ncclSuccess ? SSS_OK : SSS_ERR_NCCL_BASE + (int)r; }

#define MG_CK(expr)                                    \
    do {                                               \
        cudaError_t _e = (expr);                       \
        if (_e != cudaSuccess) return cuda_status(_e); \
    } while (0)
#define MG_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != SSS_OK) return _s; \
    } while (0)

constexpr uint64_t kShardAlign = 128;  // codes: 128 codes = w 16-byte vectors

}  // namespace

struct sss_mgpu {
    int n = 0;
    std::vector<int> device;
    std::vector<sss_ctx*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<ncclComm_t> comm;  // empty when n == 1
    std::vector<uint32_t*> d_lo, d_hi;
    std::mutex mu;
};

extern "C" {

int sss_mgpu_create(int n_devices, const int* devices, sss_mgpu** out)
{
    if (out == nullptr) return SSS_ERR_INVALID_ARG;
    *out = nullptr;
    if (n_devices < 1 || n_devices > SSS_MGPU_MAX_DEVICES) return SSS_ERR_INVALID_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return SSS_ERR_NO_DEVICE;
    }
    std::vector<int> dev(n_devices);
    for (int i = 0; i < n_devices; i++) {
        dev[i] = devices ? devices[i] : i;
        if (dev[i] < 0 || dev[i] >= count) return SSS_ERR_INVALID_ARG;
        for (int j = 0; j < i; j++)
            if (dev[j] == dev[i]) return SSS_ERR_INVALID_ARG;  // a device carries one shard
    }
    if (n_devices > 1 && !nccl().ok) return SSS_ERR_UNSUPPORTED;  // no libnccl.so.2 to be found
    sss_mgpu* m = new (std::nothrow) sss_mgpu();
    if (m == nullptr) return SSS_ERR_CUDA_BASE + (int)cudaErrorMemoryAllocation;
    m->n = n_devices;
    m->device = dev;
    m->ctx.assign(n_devices, nullptr);
    m->stream.assign(n_devices, nullptr);
    m->d_lo.assign(n_devices, nullptr);
    m->d_hi.assign(n_devices, nullptr);
    int status = SSS_OK;
    for (int i = 0; i < n_devices && status == SSS_OK; i++) {
        status = sss_ctx_create(dev[i], &m->ctx[i]);
        if (status != SSS_OK) break;
        cudaError_t e = cudaSetDevice(dev[i]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&m->d_lo[i], SSS_MAX_PREDICATES * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&m->d_hi[i], SSS_MAX_PREDICATES * sizeof(uint32_t));
        status = cuda_status(e);
    }
    if (status == SSS_OK && n_devices > 1) {
        m->comm.assign(n_devices, nullptr);
        status = nccl_status(nccl().CommInitAll(m->comm.data(), n_devices, dev.data()));
        if (status != SSS_OK) m->comm.clear();
    }
    if (status != SSS_OK) {
        sss_mgpu_destroy(m);
        return status;
    }
    *out = m;
    return SSS_OK;
}

int sss_mgpu_destroy(sss_mgpu* m)
{
    if (m == nullptr) return SSS_OK;
    for (int i = 0; i < m->n; i++) {
        cudaSetDevice(m->device[i]);
        if (m->stream[i]) cudaStreamSynchronize(m->stream[i]);
        if (i < (int)m->comm.size() && m->comm[i]) nccl().CommDestroy(m->comm[i]);
        if (m->stream[i]) cudaStreamDestroy(m->stream[i]);
        if (m->d_lo[i]) cudaFree(m->d_lo[i]);
        if (m->d_hi[i]) cudaFree(m->d_hi[i]);
        if (m->ctx[i]) sss_ctx_destroy(m->ctx[i]);
    }
    delete m;
    return SSS_OK;
}

int sss_mgpu_device_count(const sss_mgpu* m) { return m ? m->n : 0; }
int sss_mgpu_device(const sss_mgpu* m, int shard) { return (m && shard >= 0 && shard < m->n) ? m->device[shard] : -1; }
sss_ctx* sss_mgpu_ctx(sss_mgpu* m, int shard) { return (m && shard >= 0 && shard < m->n) ? m->ctx[shard] : nullptr; }
void* sss_mgpu_stream(sss_mgpu* m, int shard) { return (m && shard >= 0 && shard < m->n) ? m->stream[shard] : nullptr; }

int sss_shard_range(uint64_t n_global, int n_shards, int shard, uint64_t* first, uint64_t* count)
{
    if (n_shards < 1 || shard < 0 || shard >= n_shards || first == nullptr || count == nullptr) return SSS_ERR_INVALID_ARG;
    uint64_t per = (n_global + (uint64_t)n_shards - 1) / (uint64_t)n_shards;
    per = (per + kShardAlign - 1) / kShardAlign * kShardAlign;
    const uint64_t f = std::min(per * (uint64_t)shard, n_global);
    *first = f;
    *count = std::min(per, n_global - f);
    return SSS_OK;
}

int sss_mgpu_sync(sss_mgpu* m)
{
    if (m == nullptr) return SSS_ERR_INVALID_ARG;
    for (int i = 0; i < m->n; i++) {
        MG_CK(cudaSetDevice(m->device[i]));
        MG_CK(cudaStreamSynchronize(m->stream[i]));
    }
    return SSS_OK;
}

// one ncclAllReduce(sum, uint64 x count) over the shards' count arrays, in place, each on its shard's stream
static int allreduce_counts(sss_mgpu* m, uint64_t* const* d_counts, size_t count)
{
    if (m->n == 1) return SSS_OK;
    const NcclApi& api = nccl();
    ncclResult_t r = api.GroupStart();
    for (int i = 0; i < m->n && r == ncclSuccess; i++)
        r = api.AllReduce(d_counts[i], d_counts[i], count, ncclUint64, ncclSum, m->comm[i], m->stream[i]);
    const ncclResult_t r2 = api.GroupEnd();
    return nccl_status(r != ncclSuccess ? r : r2);
}

int sss_mgpu_scan_range(sss_mgpu* m, const void* const* d_packed, uint64_t n_global, int w, uint32_t lo, uint32_t hi,
                        void* const* d_bitvec, uint64_t* const* d_count, uint64_t* const* d_positions,
                        const uint64_t* capacity, uint64_t* const* d_num_local)
{
    if (m == nullptr || d_packed == nullptr || d_bitvec == nullptr || d_count == nullptr) return SSS_ERR_INVALID_ARG;
    if (d_positions != nullptr && capacity == nullptr) return SSS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> guard(m->mu);
    for (int i = 0; i < m->n; i++) {
        uint64_t first, cnt;
        MG_TRY(sss_shard_range(n_global, m->n, i, &first, &cnt));
        if (d_count[i] == nullptr) return SSS_ERR_INVALID_ARG;
        MG_CK(cudaSetDevice(m->device[i]));
        if (d_positions != nullptr)
            MG_TRY(sss_scan_range_positions(m->ctx[i], d_packed[i], cnt, w, lo, hi, d_bitvec[i], d_count[i], first,
                                            d_positions[i], capacity[i], m->stream[i]));
        else
            MG_TRY(sss_scan_range(m->ctx[i], d_packed[i], cnt, w, lo, hi, d_bitvec[i], d_count[i], m->stream[i]));
        if (d_num_local != nullptr && d_num_local[i] != nullptr)
            MG_CK(cudaMemcpyAsync(d_num_local[i], d_count[i], sizeof(uint64_t), cudaMemcpyDeviceToDevice, m->stream[i]));
    }
    return allreduce_counts(m, d_count, 1);
}

int sss_mgpu_shared_scan(sss_mgpu* m, const void* const* d_packed, uint64_t n_global, int w, const uint32_t* h_lo,
                         const uint32_t* h_hi, int K, int linear, void* const* d_out, uint64_t stride_bytes,
                         uint64_t* const* d_counts)
{
    if (m == nullptr || d_packed == nullptr || d_out == nullptr || d_counts == nullptr || h_lo == nullptr || h_hi == nullptr)
        return SSS_ERR_INVALID_ARG;
    if (K < 1 || K > SSS_MAX_PREDICATES) return SSS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> guard(m->mu);
    for (int i = 0; i < m->n; i++) {
        uint64_t first, cnt;
        MG_TRY(sss_shard_range(n_global, m->n, i, &first, &cnt));
        if (d_counts[i] == nullptr) return SSS_ERR_INVALID_ARG;
        MG_CK(cudaSetDevice(m->device[i]));
        // query bounds are replicated: K x 8 bytes per device
        MG_CK(cudaMemcpyAsync(m->d_lo[i], h_lo, (size_t)K * 4, cudaMemcpyHostToDevice, m->stream[i]));
        MG_CK(cudaMemcpyAsync(m->d_hi[i], h_hi, (size_t)K * 4, cudaMemcpyHostToDevice, m->stream[i]));
        if (linear)
            MG_TRY(sss_shared_scan_linear(m->ctx[i], d_packed[i], cnt, w, m->d_lo[i], m->d_hi[i], K, d_out[i], d_counts[i],
                                          m->stream[i]));
        else
            MG_TRY(sss_shared_scan(m->ctx[i], d_packed[i], cnt, w, m->d_lo[i], m->d_hi[i], K, d_out[i], stride_bytes,
                                   d_counts[i], m->stream[i]));
    }
    return allreduce_counts(m, d_counts, (size_t)K);
}

// ---- host buffers: one column in host memory, every GPU pulls its own shard over its own PCIe link ----

int sss_mgpu_host_scan_range(sss_mgpu* m, const void* h_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                             int tail_granule, void* h_bitvec, uint64_t* h_count)
{
    if (m == nullptr || h_packed == nullptr || h_bitvec == nullptr || w < 1 || w > 32) return SSS_ERR_INVALID_ARG;
    if (n != 0 && sss_tail_eval_codes(n, w, tail_granule) == 0) return SSS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> guard(m->mu);
    std::vector<int> status(m->n, SSS_OK);
    std::vector<uint64_t> hits(m->n, 0);
    std::vector<std::thread> team;
    int last = 0;  // the last non-empty shard carries the variant's tail behaviour
    for (int i = 0; i < m->n; i++) {
        uint64_t first, cnt;
        MG_TRY(sss_shard_range(n, m->n, i, &first, &cnt));
        if (cnt != 0) last = i;
    }
    for (int i = 0; i < m->n; i++) {
        team.emplace_back([&, i] {
            sss_bind_host_thread(m->device[i], nullptr);  // this thread's staging buffers and copies stay next to its GPU
            uint64_t first, cnt;
            sss_shard_range(n, m->n, i, &first, &cnt);
            if (cnt == 0 && !(i == 0 && n == 0)) return;
            const uint8_t* in = static_cast<const uint8_t*>(h_packed) + first * (uint64_t)w / 8;
            uint8_t* out = static_cast<uint8_t*>(h_bitvec) + first / 8;
            status[i] = sss_host_scan_range(m->ctx[i], in, cnt, w, lo, hi, i == last ? tail_granule : SSS_TAIL_EXACT, out,
                                            &hits[i]);
        });
    }
    for (auto& t : team) t.join();
    uint64_t total = 0;
    for (int i = 0; i < m->n; i++) {
        if (status[i] != SSS_OK) return status[i];
        total += hits[i];
    }
    if (h_count) *h_count = total;
    return SSS_OK;
}

int sss_mgpu_host_shared_scan(sss_mgpu* m, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                              const uint32_t* h_hi, int K, int tail_granule, void* h_bitvecs, uint64_t stride_bytes,
                              uint64_t* h_counts)
{
    if (m == nullptr || h_packed == nullptr || h_bitvecs == nullptr || h_lo == nullptr || h_hi == nullptr || w < 1 || w > 32)
        return SSS_ERR_INVALID_ARG;
    if (K < 1 || K > SSS_MAX_PREDICATES) return SSS_ERR_INVALID_ARG;
    if (n != 0 && sss_tail_eval_codes(n, w, tail_granule) == 0) return SSS_ERR_INVALID_ARG;
    if (stride_bytes < sss_tail_out_bytes(n, w, tail_granule)) return SSS_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> guard(m->mu);
    std::vector<int> status(m->n, SSS_OK);
    std::vector<std::vector<uint64_t>> counts(m->n, std::vector<uint64_t>(K, 0));
    int last = 0;
    for (int i = 0; i < m->n; i++) {
        uint64_t first, cnt;
        MG_TRY(sss_shard_range(n, m->n, i, &first, &cnt));
        if (cnt != 0) last = i;
    }
    std::vector<std::thread> team;
    for (int i = 0; i < m->n; i++) {
        team.emplace_back([&, i] {
            sss_bind_host_thread(m->device[i], nullptr);  // this thread's staging buffers and copies stay next to its GPU
            uint64_t first, cnt;
            sss_shard_range(n, m->n, i, &first, &cnt);
            if (cnt == 0) return;
            const uint8_t* in = static_cast<const uint8_t*>(h_packed) + first * (uint64_t)w / 8;
            uint8_t* out = static_cast<uint8_t*>(h_bitvecs) + first / 8;
            // a shard's slice of bitvector k starts at out + k * stride_bytes: the caller's stride serves every shard
            status[i] = sss_host_shared_scan(m->ctx[i], in, cnt, w, h_lo, h_hi, K, i == last ? tail_granule : SSS_TAIL_EXACT,
                                             out, stride_bytes, counts[i].data());
        });
    }
    for (auto& t : team) t.join();
    for (int i = 0; i < m->n; i++)
        if (status[i] != SSS_OK) return status[i];
    if (h_counts)
        for (int k = 0; k < K; k++) {
            uint64_t s = 0;
            for (int i = 0; i < m->n; i++) s += counts[i][k];
            h_counts[k] = s;
        }
    return SSS_OK;
}

}  // extern "C"
