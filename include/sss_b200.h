/*
 * sss_b200.h -- C ABI of the B200-native Shared_SIMD_Scan hot path (libsss_b200.so).
 *
 * This is the drop-in boundary: plain C, raw pointers and sizes, int status codes, no C++ or
 * torch types.  The reference (RRr89/Shared_SIMD_Scan) has no FFI of its own -- its public surface
 * is the set of C++ free functions in src/simd_scan.hpp:20-120 -- so each entry point below names
 * the reference function(s) it replaces.  include/simd_scan.hpp re-creates those C++ signatures on
 * top of this ABI.
 *
 * Conventions
 *  - Packed column format = the reference compressor's (src/simd_scan_compression.cpp:53-104):
 *    code i occupies bits [i*w, (i+1)*w) of a little-endian bit stream, w in [1,32].
 *    The buffer must be readable up to sss_packed_bytes(n, w) (= compressed_buffer_size,
 *    src/simd_scan.hpp:20-26) and its base must be 16-byte aligned.
 *  - Bitvector: bit i%8 of byte i/8 is the predicate on code i (src/simd_scan.cpp:48-52,
 *    src/util.cpp:51-58).  Device entry points write whole 32-bit words, 4*ceil(n/32) bytes,
 *    with bits >= n cleared; buffers are sized with sss_bitvec_bytes(n) (= scan_output_buffer_size,
 *    src/simd_scan.hpp:35-40) and must be 16-byte aligned (any cudaMalloc pointer, and any shard
 *    slice that starts at a multiple of 128 codes).
 *  - Predicates are inclusive unsigned ranges lo <= code <= hi.  The reference's equality scan
 *    (`decompressed_element == predicate_key`, src/simd_scan.cpp:47) is lo == hi == key.
 *    lo > hi, or lo >= 2^w, matches nothing.
 *  - Counts and positions are 64-bit (the reference's `int hits` overflows past 2^31 codes).
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device entry points
 *    only enqueue work; host entry points (sss_host_*) are synchronous.
 *  - Threads: every entry point takes the context's lock (host entry points for their whole
 *    synchronous duration, device entry points while they enqueue), so a context may be shared by
 *    threads -- its calls serialise.  For CONCURRENT calls give every thread its own context (the C++
 *    drop-in layer, csrc/host/simd_scan_host.cpp, keeps a pool; the reference's functions are stateless,
 *    src/simd_scan.cpp:103-158).  A context serves one shared scan / one position pass at a time per
 *    stream order: its predicate index and look-back workspace live in the context, so device entry
 *    points of ONE context must be enqueued on one stream (or on streams the caller orders).
 *  - Return value: SSS_OK, a negative SSS_ERR_* argument error, 1000 + cudaError_t, or (multi-GPU entry
 *    points) 2000 + ncclResult_t.
 *    There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SSS_B200_H
#define SSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSS_OK 0
#define SSS_ERR_INVALID_ARG (-1)
#define SSS_ERR_UNSUPPORTED (-2)
#define SSS_ERR_NO_DEVICE (-3)
#define SSS_ERR_CUDA_BASE 1000 /* 1000 + cudaError_t */
#define SSS_ERR_NCCL_BASE 2000 /* 2000 + ncclResult_t (multi-GPU entry points) */

/* tail_granule values of the host entry points: how far past code n-1 the reference variant being
 * replaced keeps evaluating the predicate on the buffer padding. */
#define SSS_TAIL_EXACT 1       /* canonical: nothing past n */
#define SSS_TAIL_8 8           /* scan_128 / scan_256 / shared_*_standard / linear_*: next multiple of 8 */
#define SSS_TAIL_32 32         /* *_unrolled: next multiple of 32, whole 32-bit words written */
#define SSS_TAIL_CURSOR32 132  /* shared_scan_*_parallel: every code that starts inside ceil(w*n/32) words */
#define SSS_TAIL_CURSOR64 164  /* scan_unvectorized: every code that starts inside ceil(w*n/64) 64-bit words */

#define SSS_MAX_PREDICATES 1024 /* largest K of the reference's shared scans (src/simd_scan_shared_linear.cpp:76) */

typedef struct sss_ctx sss_ctx;

/* ---- lifecycle --------------------------------------------------------------------------- */
int sss_version(void);
const char* sss_error_string(int status);
int sss_ctx_create(int device, sss_ctx** out);
int sss_ctx_destroy(sss_ctx* ctx);
int sss_ctx_sm_count(const sss_ctx* ctx);

/* ---- buffer-size helpers (bytes, including the reference's SIMD padding) ---------------------- */
size_t sss_packed_bytes(uint64_t n, int w);  /* compressed_buffer_size, src/simd_scan.hpp:20-26 */
size_t sss_unpack_bytes(uint64_t n);         /* decompression_output_buffer_size, :28-33 */
size_t sss_bitvec_bytes(uint64_t n);         /* scan_output_buffer_size, :35-40 */

/* Pure arithmetic (no device needed): number of codes a tail mode evaluates and the number of
 * output bytes (per bitvector, or rows for the linear layout) it writes.  0 for a bad mode. */
uint64_t sss_tail_eval_codes(uint64_t n, int w, int tail_granule);
uint64_t sss_tail_out_bytes(uint64_t n, int w, int tail_granule);

/* ---- device-resident entry points (pointers are device pointers) --------------------------- */

/* compress_9bit_input (src/simd_scan_compression.cpp:53-104), any w, uint32 codes masked to w bits.
 * Writes the whole sss_packed_bytes(n,w) buffer (payload + zero padding). */
int sss_pack(sss_ctx* ctx, const uint32_t* d_codes, uint64_t n, int w, void* d_packed, void* stream);

/* Synthetic uniform column packed in place: code_i = mix64(seed + (first_index+i)*C) & (2^w-1),
 * identical to the oracle's generator (stands in for the reference benchmark's data generation,
 * src/benchmark.cpp:81,173,277). */
int sss_generate_packed(sss_ctx* ctx, void* d_packed, uint64_t n, int w, uint64_t seed,
                        uint64_t first_index, void* stream);

/* decompress_* (src/simd_scan_decompression.cpp:6-470): one zero-extended int32 per code.
 * d_out must be 32-byte aligned (the kernel writes 256-bit vectors). */
int sss_unpack(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, int32_t* d_out, void* stream);

/* scan_* (src/simd_scan.cpp:20-306) generalised to a range.  *d_count is overwritten with the
 * number of matches in [0,n); d_count may be NULL. */
int sss_scan_range(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                   void* d_bitvec, uint64_t* d_count, void* stream);

/* Ascending position list of a bitvector (no reference function): positions of the
 * set bits among [0,n), each plus `base`, written in ascending order; *d_num = how many exist
 * (may exceed `capacity`, only the first `capacity` are written). */
int sss_positions(sss_ctx* ctx, const void* d_bitvec, uint64_t n, uint64_t base,
                  uint64_t* d_positions, uint64_t capacity, uint64_t* d_num, void* stream);

/* The north star's scan output in one call: bitvector + match count (sss_scan_range) and the ascending
 * position list of the matches (sss_positions; the position pass reads the bitvector the scan has just
 * written).  *d_count = number of matches; positions beyond `capacity` are not written. */
int sss_scan_range_positions(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                             void* d_bitvec, uint64_t* d_count, uint64_t base, uint64_t* d_positions,
                             uint64_t capacity, void* stream);

/* Downstream consumers of a scan result (the reference stops at the bitvector).
 * Combine two bitvectors over the first n bits: op = SSS_OP_AND / OR / ANDNOT (a & ~b) / XOR; writes
 * 4*ceil(n/32) bytes with bits >= n cleared, *d_count (optional) = popcount of the result.  d_out may
 * alias d_a or d_b. */
#define SSS_OP_AND 0
#define SSS_OP_OR 1
#define SSS_OP_ANDNOT 2
#define SSS_OP_XOR 3
int sss_bitvec_combine(sss_ctx* ctx, int op, const void* d_a, const void* d_b, uint64_t n, void* d_out,
                       uint64_t* d_count, void* stream);

/* Late materialisation: d_out[i] = the code at d_positions[i] of a packed column (any order, any
 * repeats; positions must be < the column's code count). */
int sss_unpack_at(sss_ctx* ctx, const void* d_packed, int w, const uint64_t* d_positions, uint64_t num,
                  int32_t* d_out, void* stream);

/* Memory-bandwidth calibration (replaces bench_memory<T> / bench_memcpy, src/benchmark_misc.cpp:9-52):
 * one streaming pass with the access mix of this path, to be timed by the caller.
 *   SSS_CALIB_READ  reads `bytes` from d_a          SSS_CALIB_WRITE writes `bytes` to d_b
 *   SSS_CALIB_COPY  d_a -> d_b, `bytes` each way    SSS_CALIB_MIX   reads `bytes`, writes bytes/ratio
 * (MIX with ratio = W is the traffic shape of the fused scan at W bits).  Buffers 16-byte aligned. */
#define SSS_CALIB_READ 0
#define SSS_CALIB_WRITE 1
#define SSS_CALIB_COPY 2
#define SSS_CALIB_MIX 3
#define SSS_CALIB_STREAMS32 4   /* writes `bytes` to d_b as `ratio` = K interleaved output streams, 32-bit stores: */
#define SSS_CALIB_STREAMS128 5  /* the shared scan's bitvector write pattern without its computation (128-bit stores) */
int sss_calibrate(sss_ctx* ctx, int mode, const void* d_a, void* d_b, uint64_t bytes, uint32_t ratio, void* stream);

/* shared_scan_* (src/simd_scan_shared.cpp:9-466): K bitvectors, predicate k's at
 * d_bitvecs + k*stride_bytes (stride a multiple of 4; a 128-byte aligned base and stride keep every
 * warp store on whole lines, worth ~3 %); d_counts[K] overwritten (may be NULL). */
int sss_shared_scan(sss_ctx* ctx, const void* d_packed, uint64_t n, int w, const uint32_t* d_lo,
                    const uint32_t* d_hi, int K, void* d_bitvecs, uint64_t stride_bytes,
                    uint64_t* d_counts, void* stream);

/* shared_scan_128_linear_* (src/simd_scan_shared_linear.cpp:9-82): byte for (group g of 8 codes,
 * predicate k) at d_out[g*K + k]; writes 4*ceil(n/32) rows of K bytes. */
int sss_shared_scan_linear(sss_ctx* ctx, const void* d_packed, uint64_t n, int w,
                           const uint32_t* d_lo, const uint32_t* d_hi, int K, void* d_out,
                           uint64_t* d_counts, void* stream);

/* ---- host-buffer entry points (what the C++ drop-in layer calls) --------------------------------
 * Synchronous: packed input is copied host->device in chunks overlapped with the kernels and the
 * device->host copies of the results.  tail_granule (SSS_TAIL_*) reproduces the tail behaviour of
 * the reference variant being replaced: codes in
 * [n, sss_tail_eval_codes(n, w, mode)) are decoded from the buffer padding and evaluated, exactly
 * like the reference's loops; sss_tail_out_bytes(n, w, mode) bytes are written, nothing beyond.
 * Pinned buffers (sss_host_alloc / sss_host_register) transfer at full PCIe speed. */
int sss_host_unpack(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, int32_t* h_out);
int sss_host_scan_range(sss_ctx* ctx, const void* h_packed, uint64_t n, int w, uint32_t lo,
                        uint32_t hi, int tail_granule, void* h_bitvec, uint64_t* h_count);
int sss_host_shared_scan(sss_ctx* ctx, const void* h_packed, uint64_t n, int w,
                         const uint32_t* h_lo, const uint32_t* h_hi, int K, int tail_granule,
                         void* h_bitvecs, uint64_t stride_bytes, uint64_t* h_counts);
/* same, one separately allocated bitvector per predicate (std::vector<std::vector<uint8_t>>& outputs of
 * the reference, src/simd_scan.hpp:102-113) */
int sss_host_shared_scan_ptrs(sss_ctx* ctx, const void* h_packed, uint64_t n, int w,
                              const uint32_t* h_lo, const uint32_t* h_hi, int K, int tail_granule,
                              void* const* h_bitvec_ptrs, uint64_t* h_counts);
int sss_host_shared_scan_linear(sss_ctx* ctx, const void* h_packed, uint64_t n, int w,
                                const uint32_t* h_lo, const uint32_t* h_hi, int K, int tail_granule,
                                void* h_out, uint64_t* h_counts);
int sss_host_pack(sss_ctx* ctx, const uint32_t* h_codes, uint64_t n, int w, void* h_packed);
/* Ascending positions (+ base) of the set bits among the first n bits of a HOST bitvector, expanded on
 * the device; *h_num = how many exist, at most `capacity` are written. */
int sss_host_positions(sss_ctx* ctx, const void* h_bitvec, uint64_t n, uint64_t base, uint64_t* h_positions,
                       uint64_t capacity, uint64_t* h_num);

void* sss_host_alloc(size_t bytes); /* pinned host memory, NULL on failure */
void sss_host_free(void* p);
int sss_host_register(void* p, size_t bytes);
int sss_host_unregister(void* p);
/* Bind the CALLING host thread to the CPUs next to a GPU (its PCIe root's NUMA node, from
 * /sys/bus/pci/devices/<bdf>/local_cpulist), so that the pinned buffers this thread allocates and
 * first touches afterwards are node-local and its copies do not cross the socket interconnect.  On a
 * one-node box, or where sysfs does not say, it changes nothing and returns SSS_OK; *n_cpus (may be
 * NULL) = CPUs the thread may run on afterwards.  The multi-GPU host entry points call it for their
 * per-GPU threads; a one-process-per-GPU launcher calls it once per rank (bench.py does). */
int sss_bind_host_thread(int device, int* n_cpus);

/* ---- multi-GPU entry points (the reference is single-device) --------
 * One process drives N GPUs.  The column is RANGE-partitioned: shard g holds codes
 * [first_g, first_g + count_g) of the global column (sss_shard_range: every shard but the last is a
 * multiple of 128 codes, so each starts 16-byte aligned in the packed stream for any width and on a
 * byte boundary of every bitvector).  Every GPU scans its own shard with the single-GPU kernels on its
 * own stream; the per-query match counts are combined with ONE ncclAllReduce(sum, uint64 x K) over
 * NVLink; position lists are rebased by first_g, so the shards' lists in device order are globally
 * ascending.  No column data crosses GPUs.  Contrast: the reference's only parallel variant partitions
 * the predicates over OpenMP threads and re-reads the column per predicate (src/simd_scan_shared.cpp:25-32).
 * NCCL (libnccl.so.2) is loaded at run time and only when N > 1.
 * Device entry points enqueue on the shards' streams (sss_mgpu_stream) and return; sss_mgpu_sync waits.
 * Array arguments have one entry per shard, in shard order. */
#define SSS_MGPU_MAX_DEVICES 16
typedef struct sss_mgpu sss_mgpu;

int sss_mgpu_create(int n_devices, const int* devices /* NULL: 0..n_devices-1 */, sss_mgpu** out);
int sss_mgpu_destroy(sss_mgpu* m);
int sss_mgpu_device_count(const sss_mgpu* m);
int sss_mgpu_device(const sss_mgpu* m, int shard);      /* CUDA ordinal of the shard's GPU */
sss_ctx* sss_mgpu_ctx(sss_mgpu* m, int shard);          /* the shard's single-GPU context (pack, generate, ...) */
void* sss_mgpu_stream(sss_mgpu* m, int shard);          /* cudaStream_t the shard's work is enqueued on */
int sss_mgpu_sync(sss_mgpu* m);

/* Pure arithmetic: the contiguous range partition of n_global codes into n_shards shards. */
int sss_shard_range(uint64_t n_global, int n_shards, int shard, uint64_t* first, uint64_t* count);

/* scan_* over the global column.  d_packed[g] / d_bitvec[g]: shard g's packed codes / its slice of the
 * bitvector, on GPU g.  d_count[g] (one uint64 on GPU g, required) receives the GLOBAL match count on
 * every GPU.  Optional position lists: d_positions[g] (capacity[g] entries, on GPU g) receives shard g's
 * matches as GLOBAL code indices, ascending; d_num_local[g] (optional) = how many shard g has. */
int sss_mgpu_scan_range(sss_mgpu* m, const void* const* d_packed, uint64_t n_global, int w, uint32_t lo, uint32_t hi,
                        void* const* d_bitvec, uint64_t* const* d_count, uint64_t* const* d_positions,
                        const uint64_t* capacity, uint64_t* const* d_num_local);

/* shared_scan_* / shared_scan_128_linear_* over the global column: K bounds in HOST memory (replicated to
 * every GPU), d_out[g] = shard g's K bitvector slices (stride_bytes apart) or, linear != 0, its rows of
 * the [n/8][K] byte matrix; d_counts[g] (K uint64 on GPU g) receives the GLOBAL per-query counts. */
int sss_mgpu_shared_scan(sss_mgpu* m, const void* const* d_packed, uint64_t n_global, int w, const uint32_t* h_lo,
                         const uint32_t* h_hi, int K, int linear, void* const* d_out, uint64_t stride_bytes,
                         uint64_t* const* d_counts);

/* Host buffers (synchronous): ONE column in host memory; every GPU pulls its own shard over its own
 * PCIe link (one host thread per GPU), results land in the caller's buffers, counts are summed on the
 * host.  Same arguments and tail modes as sss_host_scan_range / sss_host_shared_scan. */
int sss_mgpu_host_scan_range(sss_mgpu* m, const void* h_packed, uint64_t n, int w, uint32_t lo, uint32_t hi,
                             int tail_granule, void* h_bitvec, uint64_t* h_count);
int sss_mgpu_host_shared_scan(sss_mgpu* m, const void* h_packed, uint64_t n, int w, const uint32_t* h_lo,
                              const uint32_t* h_hi, int K, int tail_granule, void* h_bitvecs, uint64_t stride_bytes,
                              uint64_t* h_counts);

#ifdef __cplusplus
}
#endif
#endif /* SSS_B200_H */
