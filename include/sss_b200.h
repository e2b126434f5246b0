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
 *  - A context is not thread-safe: use it from one host thread at a time (one context per thread
 *    is fine), and it serves one shared scan at a time (its predicate index lives in the context).
 *  - Return value: SSS_OK, a negative SSS_ERR_* argument error, or 1000 + cudaError_t.
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

void* sss_host_alloc(size_t bytes); /* pinned host memory, NULL on failure */
void sss_host_free(void* p);
int sss_host_register(void* p, size_t bytes);
int sss_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SSS_B200_H */
