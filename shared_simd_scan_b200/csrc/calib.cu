// calib.cu -- memory-bandwidth calibration kernels.
//
// The reference ships two host micro-benchmarks next to its scans -- bench_memory<T> (an element
// loop copy) and bench_memcpy (src/benchmark_misc.cpp:9-52) -- to put its scan rates next to what the
// machine's memory can do.  The counterpart here measures what a B200's HBM delivers for the access
// MIXES of this path, because the roofline denominator (a copy: 50 % reads, 50 % writes) is not the
// mix of a scan: the fused scan at W bits reads W bytes for every byte it writes.
//
//   mode READ   : stream `bytes` from a, fold them into one word (so the loads cannot be elided)
//   mode WRITE  : stream `bytes` into b
//   mode COPY   : a -> b, `bytes` each way            (what MEASURED_PEAKS.json's hbm_gbs measures)
//   mode MIX    : read `bytes` from a, write bytes/ratio to b: the traffic shape of scan_range_kernel<ratio>
//
// One CTA per contiguous CHUNK, 128-bit accesses, up to 8 independent loads in flight per thread.  The chunks
// are handed out by the hardware CTA scheduler as SMs become free -- the same dynamic order the product
// kernels get from their tile tickets.  (Until the end of round 2 these were grid-stride loops, i.e. a static
// order, and undersold the hardware by 10-15 %: tools/probes/copy_patterns.cu measures 6.0-6.2 TB/s for a
// grid-stride copy and 6.85 TB/s for the chunked one, profiles/r02_copy_patterns.txt.)
#include "common.cuh"
#include "kernels.h"

namespace sss {

constexpr int kCalibUnroll = 8;   // READ / WRITE / COPY: vectors per thread and chunk (32 KB chunks)
constexpr int kCalibMixOut = 2;   // MIX: output vectors per thread and chunk (8 KB out, ratio x 8 KB in)

__global__ void __launch_bounds__(kThreads)
calib_kernel(int mode, const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n_vec, uint32_t ratio,
             unsigned long long* __restrict__ sink)
{
    uint32_t acc = 0;
    if (mode == 3) {  // MIX: read ratio * kCalibMixOut vectors per thread, write kCalibMixOut
        const uint64_t n_out = n_vec / ratio;
        const uint64_t out0 = (uint64_t)blockIdx.x * (kThreads * kCalibMixOut);
        const uint64_t in0 = out0 * ratio;
        const uint32_t per_thread = ratio * kCalibMixOut;
        for (uint32_t i0 = 0; i0 < per_thread; i0 += kCalibUnroll) {
            uint4 v[kCalibUnroll];
#pragma unroll
            for (int u = 0; u < kCalibUnroll; u++) {
                const uint64_t i = in0 + (uint64_t)(i0 + u) * kThreads + threadIdx.x;
                v[u] = (i0 + u < per_thread && i < n_vec) ? __ldcs(a + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kCalibUnroll; u++) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
#pragma unroll
        for (int u = 0; u < kCalibMixOut; u++) {
            const uint64_t o = out0 + (uint64_t)u * kThreads + threadIdx.x;
            if (o < n_out) __stcs(b + o, make_uint4(acc, acc, acc, acc));
        }
        return;
    }
    const uint64_t base = (uint64_t)blockIdx.x * (kThreads * kCalibUnroll) + threadIdx.x;
    if (mode == 1) {  // WRITE
        const uint4 v = make_uint4((uint32_t)base, 1u, 2u, 3u);
#pragma unroll
        for (int u = 0; u < kCalibUnroll; u++) {
            const uint64_t i = base + (uint64_t)u * kThreads;
            if (i < n_vec) __stcs(b + i, v);
        }
        return;
    }
    uint4 v[kCalibUnroll];
#pragma unroll
    for (int u = 0; u < kCalibUnroll; u++) {
        const uint64_t i = base + (uint64_t)u * kThreads;
        v[u] = i < n_vec ? __ldcs(a + i) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kCalibUnroll; u++) {
        const uint64_t i = base + (uint64_t)u * kThreads;
        if (mode == 2) {  // COPY
            if (i < n_vec) __stcs(b + i, v[u]);
        } else {
            acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        }
    }
    if (mode == 0 && acc == 0x9E3779B9u && sink != nullptr) atomicAdd(sink, 1ull);  // keeps the loads alive
}

// mode STREAMS32 / STREAMS128: the WRITE pattern of the shared scan's bitvector layout without any of
// its computation -- `ratio` = K independent output streams of bytes/K each; a 512-thread CTA owns tiles
// of 512 (x4) consecutive words and writes its piece of EVERY stream, one 32-bit (128-bit) store per
// thread and stream, streams innermost (exactly the order shared_scan_kernel stores in).  Tells apart
// "the memory system dislikes K x 148 interleaved 2 KB runs" from "the kernel does not overlap".
template <int kVec>
__global__ void __launch_bounds__(512)
calib_streams_kernel(const uint4* __restrict__ a, uint32_t read_vec_per_thread, uint32_t* __restrict__ b,
                     uint64_t stream_words, uint32_t K)
{
    const uint64_t tile_words = 512ull * kVec;
    const uint64_t tiles = stream_words / tile_words;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        uint32_t* p = b + tile * tile_words + (uint64_t)threadIdx.x * kVec;
        uint32_t v = (uint32_t)tile ^ threadIdx.x;
        if (a != nullptr) {  // the scan's input side: read_vec_per_thread x 16 bytes per thread and tile, contiguous per tile
            const uint4* src = a + tile * 512ull * read_vec_per_thread + threadIdx.x;
            for (uint32_t r = 0; r < read_vec_per_thread; r++) {
                const uint4 q = __ldcs(src + r * 512u);
                v ^= q.x ^ q.y ^ q.z ^ q.w;
            }
        }
#pragma unroll 8
        for (uint32_t k = 0; k < K; k++) {
            if constexpr (kVec == 4) __stcs(reinterpret_cast<uint4*>(p), make_uint4(v, k, v, k));
            else __stcs(p, v + k);
            p += stream_words;
        }
    }
}

cudaError_t launch_calibrate(const Device& dev, int mode, const void* a, void* b, uint64_t bytes, uint32_t ratio,
                             unsigned long long* sink, cudaStream_t stream)
{
    const uint64_t n_vec = bytes / 16;
    if (n_vec == 0) return cudaSuccess;
    if (mode == 4 || mode == 5) {
        const uint32_t K = ratio == 0 ? 1 : ratio;
        const uint64_t stream_words = bytes / 4 / K / 2048 * 2048;
        if (stream_words == 0) return cudaSuccess;
        // with d_a given, every thread also reads 4 x 16 bytes per tile (the packed input of ~16-bit codes)
        const uint32_t rv = a != nullptr ? (mode == 4 ? 4u : 16u) : 0u;
        if (mode == 4)
            calib_streams_kernel<1><<<dev.sm_count, 512, 0, stream>>>(static_cast<const uint4*>(a), rv, static_cast<uint32_t*>(b),
                                                                      stream_words, K);
        else
            calib_streams_kernel<4><<<dev.sm_count, 512, 0, stream>>>(static_cast<const uint4*>(a), rv, static_cast<uint32_t*>(b),
                                                                      stream_words, K);
        return cudaGetLastError();
    }
    const uint32_t r = ratio == 0 ? 1 : ratio;
    const uint64_t per_cta = mode == 3 ? (uint64_t)kThreads * kCalibMixOut * r : (uint64_t)kThreads * kCalibUnroll;
    const uint64_t ctas = (n_vec + per_cta - 1) / per_cta;
    if (ctas >= (1ull << 31)) return cudaErrorInvalidValue;
    calib_kernel<<<(unsigned)ctas, kThreads, 0, stream>>>(mode, static_cast<const uint4*>(a), static_cast<uint4*>(b), n_vec, r,
                                                         sink);
    return cudaGetLastError();
}

}  // namespace sss
