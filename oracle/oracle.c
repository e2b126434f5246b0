/*
 * oracle.c -- plain-C restatement of the Shared_SIMD_Scan hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for who may load it and how its parity is pinned).
 *
 * Written for clarity, not speed: one bit cursor, one code at a time.  Every function cites the
 * reference lines whose behaviour it restates (paths relative to the reference root).
 */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static inline uint32_t code_mask(int w) { return w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u); }

/* src/simd_scan.hpp:20-26 */
size_t orc_compressed_buffer_size(int w, size_t n)
{
    size_t bits = (size_t)w * n;
    return bits / 8 + (bits % 8 != 0) + 256;
}

/* src/simd_scan.hpp:28-33 */
size_t orc_decompression_output_buffer_size(size_t n) { return n * 4 + 32; }

/* src/simd_scan.hpp:35-40 */
size_t orc_scan_output_buffer_size(size_t n) { return n / 8 + (n % 8 != 0) + 32; }

/*
 * src/simd_scan_compression.cpp:66-101.  The reference ORs `input[i] << (i*w)` into a zeroed
 * uint64 array and splits straddling codes over two words; the net effect is the little-endian
 * bit stream written here byte by byte (the reference's shift by i*w relies on x86 masking the
 * count to 6 bits, :72 -- the stream position below is the defined equivalent).
 */
void orc_pack(const uint32_t* codes, size_t n, int w, uint8_t* out)
{
    memset(out, 0, orc_compressed_buffer_size(w, n));
    const uint32_t m = code_mask(w);
    for (size_t i = 0; i < n; i++) {
        uint64_t v = codes[i] & m;
        size_t bit = i * (size_t)w;
        size_t byte = bit >> 3;
        unsigned sh = (unsigned)(bit & 7);
        /* w + sh <= 39 bits: spread over at most 5 bytes */
        uint64_t chunk = v << sh;
        for (int b = 0; b < 5 && chunk; b++, chunk >>= 8) out[byte + b] |= (uint8_t)(chunk & 0xFF);
    }
}

/* SURVEY.md Appendix A; equivalent to the cursor walk of src/simd_scan_decompression.cpp:20-54. */
uint32_t orc_code_at(const uint8_t* packed, size_t i, int w)
{
    size_t bit = i * (size_t)w;
    size_t byte = bit >> 3;
    unsigned sh = (unsigned)(bit & 7);
    uint64_t chunk = 0;
    int nbytes = (int)((sh + (unsigned)w + 7) >> 3);
    for (int b = 0; b < nbytes; b++) chunk |= (uint64_t)packed[byte + b] << (8 * b);
    return (uint32_t)(chunk >> sh) & code_mask(w);
}

/* src/simd_scan_decompression.cpp:6-55: output[oi++] = decompressed_element, zero-extended. */
void orc_unpack(const uint8_t* packed, size_t n, int w, int32_t* out)
{
    for (size_t i = 0; i < n; i++) out[i] = (int32_t)orc_code_at(packed, i, w);
}

static inline size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

/*
 * src/simd_scan.cpp:46-56 (predicate, bit order, per-byte popcount) and :122-155 / :195-214 for
 * the 8- and 32-code granules.  Codes in [n, round_up(n, granule)) are decoded from the buffer
 * (zero padding in every buffer the reference produces) exactly as the SIMD loops do.
 */
size_t orc_scan_range(const uint8_t* packed, size_t n, int w, uint32_t lo, uint32_t hi,
                      int granule, uint8_t* out, uint64_t* count)
{
    size_t evaluated = granule <= 1 ? n : round_up(n, (size_t)granule);
    size_t bytes = granule == 32 ? 4 * ((n + 31) / 32) : (n + 7) / 8;
    memset(out, 0, bytes);
    uint64_t hits = 0;
    for (size_t i = 0; i < evaluated; i++) {
        uint32_t v = orc_code_at(packed, i, w);
        if (v >= lo && v <= hi) {
            out[i >> 3] |= (uint8_t)(1u << (i & 7));
            hits++;
        }
    }
    if (count) *count = hits;
    return bytes;
}

/* Bit order: src/util.cpp:51-58 (get_bit). */
size_t orc_positions(const uint8_t* bitvec, size_t n, uint64_t base, uint64_t* positions,
                     size_t capacity)
{
    size_t k = 0;
    for (size_t i = 0; i < n && k < capacity; i++)
        if ((bitvec[i >> 3] >> (i & 7)) & 1) positions[k++] = base + i;
    return k;
}

/* src/simd_scan_shared.cpp:49-86: unpack once, loop over the K predicates, outputs[key][g]. */
void orc_shared_scan(const uint8_t* packed, size_t n, int w, const uint32_t* lo,
                     const uint32_t* hi, size_t K, int granule, uint8_t* outs, size_t stride,
                     uint64_t* counts)
{
    size_t evaluated = granule <= 1 ? n : round_up(n, (size_t)granule);
    size_t bytes = granule == 32 ? 4 * ((n + 31) / 32) : (n + 7) / 8;
    for (size_t k = 0; k < K; k++) {
        memset(outs + k * stride, 0, bytes);
        if (counts) counts[k] = 0;
    }
    for (size_t i = 0; i < evaluated; i++) {
        uint32_t v = orc_code_at(packed, i, w);
        for (size_t k = 0; k < K; k++) {
            if (v >= lo[k] && v <= hi[k]) {
                outs[k * stride + (i >> 3)] |= (uint8_t)(1u << (i & 7));
                if (counts) counts[k]++;
            }
        }
    }
}

/* src/simd_scan_shared_linear.cpp:24-61: outputs[output_index*K + key_id]. */
void orc_shared_scan_linear(const uint8_t* packed, size_t n, int w, const uint32_t* lo,
                            const uint32_t* hi, size_t K, int granule, uint8_t* out,
                            uint64_t* counts)
{
    size_t evaluated = granule <= 1 ? n : round_up(n, (size_t)granule);
    size_t rows = (granule == 32 ? 4 * ((n + 31) / 32) : (n + 7) / 8);
    memset(out, 0, rows * K);
    if (counts) memset(counts, 0, K * sizeof(uint64_t));
    for (size_t i = 0; i < evaluated; i++) {
        uint32_t v = orc_code_at(packed, i, w);
        for (size_t k = 0; k < K; k++) {
            if (v >= lo[k] && v <= hi[k]) {
                out[(i >> 3) * K + k] |= (uint8_t)(1u << (i & 7));
                if (counts) counts[k]++;
            }
        }
    }
}

uint64_t orc_bitvec_combine(int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out)
{
    uint64_t hits = 0;
    size_t words = (n + 31) / 32;
    memset(out, 0, words * 4);
    for (size_t i = 0; i < n; i++) {
        int x = (a[i >> 3] >> (i & 7)) & 1, y = (b[i >> 3] >> (i & 7)) & 1, r;
        switch (op) {
            case 0: r = x & y; break;
            case 1: r = x | y; break;
            case 2: r = x & !y; break;
            default: r = x ^ y; break;
        }
        if (r) {
            out[i >> 3] |= (uint8_t)(1u << (i & 7));
            hits++;
        }
    }
    return hits;
}

void orc_unpack_at(const uint8_t* packed, int w, const uint64_t* positions, size_t num, int32_t* out)
{
    for (size_t i = 0; i < num; i++) out[i] = (int32_t)orc_code_at(packed, (size_t)positions[i], w);
}

/* SplitMix64 finaliser (public-domain constant set); SURVEY.md section 8d synthetic inputs. */
uint64_t orc_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void orc_generate_codes(uint32_t* codes, size_t n, int w, uint64_t seed, uint64_t first_index)
{
    const uint32_t m = code_mask(w);
    for (size_t i = 0; i < n; i++)
        codes[i] = (uint32_t)orc_mix64(seed + (first_index + i) * 0xD1342543DE82EF95ull) & m;
}

/* ---- timing helper (cpu_baseline "port") -------------------------------------------------- */

typedef struct {
    const uint8_t* packed;
    size_t first, count;
    int w;
    uint32_t lo, hi;
    uint8_t* out;
    uint64_t hits;
} shard_t;

static void* shard_main(void* p)
{
    shard_t* s = (shard_t*)p;
    /* shard starts at a multiple of 64 codes => byte aligned in the packed stream */
    const uint8_t* base = s->packed + (s->first * (size_t)s->w) / 8;
    orc_scan_range(base, s->count, s->w, s->lo, s->hi, 1, s->out, &s->hits);
    return NULL;
}

double orc_time_scan_range(const uint8_t* packed, size_t n, int w, uint32_t lo, uint32_t hi,
                           int threads, int reps, uint64_t* count)
{
    if (threads < 1) threads = 1;
    shard_t* sh = (shard_t*)calloc((size_t)threads, sizeof(shard_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    size_t per = round_up((n + (size_t)threads - 1) / (size_t)threads, 64);
    for (int t = 0; t < threads; t++) {
        size_t first = (size_t)t * per;
        size_t cnt = first >= n ? 0 : (n - first < per ? n - first : per);
        sh[t] = (shard_t){packed, first, cnt, w, lo, hi, NULL, 0};
        sh[t].out = (uint8_t*)malloc(orc_scan_output_buffer_size(cnt));
    }
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, shard_main, &sh[t]);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        clock_gettime(CLOCK_MONOTONIC, &b);
        double s = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
        if (s < best) best = s;
    }
    uint64_t total = 0;
    for (int t = 0; t < threads; t++) {
        total += sh[t].hits;
        free(sh[t].out);
    }
    if (count) *count = total;
    free(sh);
    free(th);
    return best;
}
