// copy_patterns.cu -- developer probe: which copy-kernel shape reaches the driver's measured copy peak
// (MEASURED_PEAKS.json: 6.55 TB/s, torch's b.copy_(a))?  2 GiB -> 2 GiB, read + write bytes counted.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_patterns copy_patterns.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// A: grid-stride, streaming hints, U loads in flight
template <int U, bool CS> __global__ void __launch_bounds__(256) copy_gs(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + u * stride; if (i < n) v[u] = CS ? __ldcs(a + i) : a[i]; }
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + u * stride; if (i < n) { if (CS) __stcs(b + i, v[u]); else b[i] = v[u]; } }
    }
}
// B: one CTA per contiguous chunk of T*U vectors (non-persistent), like torch's elementwise kernels
template <int T, int U, bool CS> __global__ void __launch_bounds__(T) copy_chunk(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n)
{
    const size_t base = (size_t)blockIdx.x * T * U + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) v[u] = CS ? __ldcs(a + i) : a[i]; }
#pragma unroll
    for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) { if (CS) __stcs(b + i, v[u]); else b[i] = v[u]; } }
}
// C: persistent CTAs that draw contiguous chunks from an atomic ticket counter (dynamic balance, no relaunch)
template <int T, int U> __global__ void __launch_bounds__(T) copy_ticket(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n, unsigned long long* ctr)
{
    __shared__ unsigned long long s_chunk;
    const size_t chunks = (n + (size_t)T * U - 1) / ((size_t)T * U);
    for (;;) {
        if (threadIdx.x == 0) s_chunk = atomicAdd(ctr, 1ull);
        __syncthreads();
        const size_t c = s_chunk;
        __syncthreads();
        if (c >= chunks) break;
        const size_t base = c * T * U + threadIdx.x;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) v[u] = a[i]; }
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) b[i] = v[u]; }
    }
}
// D: persistent CTAs, STATIC round-robin over contiguous chunks (what TileRing does)
template <int T, int U> __global__ void __launch_bounds__(T) copy_static(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n)
{
    const size_t chunks = (n + (size_t)T * U - 1) / ((size_t)T * U);
    for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const size_t base = c * T * U + threadIdx.x;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) v[u] = a[i]; }
#pragma unroll
        for (int u = 0; u < U; u++) { size_t i = base + (size_t)u * T; if (i < n) b[i] = v[u]; }
    }
}
template <class F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) f();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}
int main()
{
    const size_t bytes = 2ull << 30, n = bytes / 16;
    uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto rep = [&](const char* name, float ms) { printf("%-44s %7.1f GB/s\n", name, 2.0 * bytes / ms / 1e6); };
    rep("cudaMemcpy D2D", timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }));
    rep("grid-stride 8 CTA/SM, U=8, .cs", timeit([&] { copy_gs<8, true><<<sms * 8, 256>>>(a, b, n); }));
    rep("grid-stride 8 CTA/SM, U=8, plain", timeit([&] { copy_gs<8, false><<<sms * 8, 256>>>(a, b, n); }));
    rep("grid-stride 4 CTA/SM, U=4, plain", timeit([&] { copy_gs<4, false><<<sms * 4, 256>>>(a, b, n); }));
    rep("grid-stride 16 CTA/SM (128 thr), U=4, plain", timeit([&] { copy_gs<4, false><<<sms * 16, 128>>>(a, b, n); }));
    rep("chunk T=128 U=4 plain (torch-like)", timeit([&] { copy_chunk<128, 4, false><<<(unsigned)((n + 511) / 512), 128>>>(a, b, n); }));
    rep("chunk T=128 U=4 .cs", timeit([&] { copy_chunk<128, 4, true><<<(unsigned)((n + 511) / 512), 128>>>(a, b, n); }));
    rep("chunk T=256 U=8 plain", timeit([&] { copy_chunk<256, 8, false><<<(unsigned)((n + 2047) / 2048), 256>>>(a, b, n); }));
    rep("chunk T=256 U=8 .cs", timeit([&] { copy_chunk<256, 8, true><<<(unsigned)((n + 2047) / 2048), 256>>>(a, b, n); }));
    rep("chunk T=512 U=4 plain", timeit([&] { copy_chunk<512, 4, false><<<(unsigned)((n + 2047) / 2048), 512>>>(a, b, n); }));
    rep("chunk T=256 U=16 plain", timeit([&] { copy_chunk<256, 16, false><<<(unsigned)((n + 4095) / 4096), 256>>>(a, b, n); }));
    unsigned long long* ctr; CK(cudaMalloc(&ctr, 8));
    rep("ticket T=256 U=8, 8 CTA/SM", timeit([&] { cudaMemsetAsync(ctr, 0, 8); copy_ticket<256, 8><<<sms * 8, 256>>>(a, b, n, ctr); }));
    rep("ticket T=256 U=8, 4 CTA/SM", timeit([&] { cudaMemsetAsync(ctr, 0, 8); copy_ticket<256, 8><<<sms * 4, 256>>>(a, b, n, ctr); }));
    rep("ticket T=512 U=8, 2 CTA/SM", timeit([&] { cudaMemsetAsync(ctr, 0, 8); copy_ticket<512, 8><<<sms * 2, 512>>>(a, b, n, ctr); }));
    rep("static round-robin T=256 U=8, 8 CTA/SM", timeit([&] { copy_static<256, 8><<<sms * 8, 256>>>(a, b, n); }));
    rep("static round-robin T=256 U=8, 4 CTA/SM", timeit([&] { copy_static<256, 8><<<sms * 4, 256>>>(a, b, n); }));
    rep("static round-robin T=512 U=8, 2 CTA/SM", timeit([&] { copy_static<512, 8><<<sms * 2, 512>>>(a, b, n); }));
    rep("static round-robin T=512 U=8, 1 CTA/SM", timeit([&] { copy_static<512, 8><<<sms * 1, 512>>>(a, b, n); }));
    rep("ticket T=512 U=8, 1 CTA/SM", timeit([&] { cudaMemsetAsync(ctr, 0, 8); copy_ticket<512, 8><<<sms * 1, 512>>>(a, b, n, ctr); }));
    return 0;
}
