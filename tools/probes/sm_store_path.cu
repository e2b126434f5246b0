// sm_store_path.cu -- developer probe: what does ONE SM's store path sustain, and does it share cycles
// with shared-memory traffic?  Every CTA (one per SM, 512 threads) writes its own small L2-resident region
// over and over, so neither HBM nor the launch matter; reported: bytes per clock per SM.
//   st32   : 32-bit store per lane  (128 B per warp instruction)
//   st128  : 128-bit store per lane (512 B per warp instruction)
//   bulk   : STS.32 into a warp-private 4 KB stage + ONE cp.async.bulk shared->global of 4 KB per warp round
//   +lds   : the same while the odd warps hammer shared memory with conflict-free LDS (1 wavefront each)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sm_store_path sm_store_path.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, bool LDS_NOISE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, uint32_t region_words, int iters, unsigned long long* cycles, uint32_t* sink)
{
    extern __shared__ __align__(128) uint32_t sm[];  // [16 warps][1024 words] stage, then 16 KB noise table
    uint32_t* noise = sm + 16 * 1024;
    for (int i = threadIdx.x; i < 4096; i += 512) noise[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* mine = out + (size_t)blockIdx.x * region_words;
    const bool storer = !LDS_NOISE || (warp & 1) == 0;
    const uint32_t n_storers = LDS_NOISE ? 8 : 16;
    const uint32_t sw = LDS_NOISE ? warp / 2 : warp;
    uint32_t acc = 0;
    const long long t0 = clock64();
    if (storer) {
        for (int it = 0; it < iters; it++) {
            if (MODE == 0) {
                // 32 stores of 128 B per warp round (like one column of 32 predicates)
                uint32_t* p = mine + ((it * n_storers + sw) * 1024 + lane) % region_words;
#pragma unroll
                for (int k2 = 0; k2 < 32; k2++) __stcs(p + k2 * 32, acc + k2);
            } else if (MODE == 1) {
                uint32_t* p = mine + ((it * n_storers + sw) * 1024 + lane * 4) % region_words;
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) __stcs(reinterpret_cast<uint4*>(p + k2 * 128), make_uint4(acc, k2, it, lane));
            } else {
                uint32_t* st = sm + warp * 1024;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int k2 = 0; k2 < 32; k2++) st[k2 * 32 + lane] = acc + k2;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    uint32_t* p = mine + ((it * n_storers + sw) * 1024) % region_words;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p), "r"(smem_u32(st)), "r"(4096u) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            acc += it;
        }
        if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        uint32_t idx = lane;
        for (int it = 0; it < iters * 4; it++) {
#pragma unroll
            for (int k2 = 0; k2 < 32; k2++) acc += noise[(idx + k2 * 32) & 4095];  // consecutive lanes: conflict-free
            idx += acc & 32;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE, bool NOISE>
int run(const char* name, uint32_t* out, unsigned long long* cyc, uint32_t* sink, int sms)
{
    const uint32_t region_words = 64 * 1024 / 4 * 4;  // 256 KB per SM: 148 x 256 KB = 37 MB, L2-resident
    const int iters = 4000;
    const size_t smem = (16 * 1024 + 4096) * 4;
    CK(cudaFuncSetAttribute(k<MODE, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<MODE, NOISE><<<sms, 512, smem>>>(out, region_words, 10, cyc, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE, NOISE><<<sms, 512, smem>>>(out, region_words, iters, cyc, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    const double storers = NOISE ? 8 : 16;
    const double bytes_per_sm = (double)iters * storers * 4096;
    printf("%-14s %6.2f B/clk/SM   (%.0f GB/s chip, %.0f cycles, %.3f ms)%s\n", name, bytes_per_sm / avg, bytes_per_sm * sms / ms / 1e6, avg, ms,
           NOISE ? "  [8 warps store, 8 warps do LDS]" : "");
    return 0;
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint32_t *out, *sink; unsigned long long* cyc;
    CK(cudaMalloc(&out, (size_t)sms * 256 * 1024)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, 8 * 256));
    run<0, false>("st32", out, cyc, sink, sms);
    run<1, false>("st128", out, cyc, sink, sms);
    run<2, false>("bulk4k", out, cyc, sink, sms);
    run<0, true>("st32+lds", out, cyc, sink, sms);
    run<1, true>("st128+lds", out, cyc, sink, sms);
    run<2, true>("bulk4k+lds", out, cyc, sink, sms);
    return 0;
}
