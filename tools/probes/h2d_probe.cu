// h2d_probe.cu -- developer probe: how fast do host buffers reach the GPU on this box?
// pinned vs pageable cudaMemcpy, cudaHostRegister on the fly, and a hand-rolled staging pipeline
// (T threads memcpy pageable -> pinned bounce buffers, DMA overlapped).  nvcc -O2 -o h2d_probe h2d_probe.cu -lpthread
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void par_memcpy(char* dst, const char* src, size_t bytes, int T)
{
    if (T <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    size_t per = (bytes + T - 1) / T;
    for (int t = 0; t < T; t++) {
        size_t off = per * t;
        if (off >= bytes) break;
        size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy(dst + off, src + off, len); });
    }
    for (auto& t : th) t.join();
}
int main(int argc, char** argv)
{
    size_t bytes = (argc > 1 ? atoll(argv[1]) : 2048) << 20;
    int dev = argc > 2 ? atoi(argv[2]) : 0;
    CK(cudaSetDevice(dev));
    char *d, *hp, *hq;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMallocHost(&hp, bytes));
    hq = (char*)malloc(bytes);
    memset(hp, 1, bytes);
    memset(hq, 2, bytes);
    cudaStream_t s[4];
    for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    auto time_it = [&](const char* name, auto fn, int reps = 3) {
        fn();
        double best = 1e9;
        for (int i = 0; i < reps; i++) { double t0 = now(); fn(); best = std::min(best, now() - t0); }
        printf("%-58s %8.2f GB/s  (%.1f ms)\n", name, bytes / best / 1e9, best * 1e3);
    };
    time_it("H2D pinned, 1 stream", [&] { CK(cudaMemcpyAsync(d, hp, bytes, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); });
    time_it("H2D pinned, 4 streams x 64 MiB chunks", [&] {
        size_t c = 64 << 20; int k = 0;
        for (size_t off = 0; off < bytes; off += c, k++) CK(cudaMemcpyAsync(d + off, hp + off, std::min(c, bytes - off), cudaMemcpyHostToDevice, s[k & 3]));
        for (auto& x : s) CK(cudaStreamSynchronize(x)); });
    time_it("D2H pinned, 1 stream", [&] { CK(cudaMemcpyAsync(hp, d, bytes, cudaMemcpyDeviceToHost, s[0])); CK(cudaStreamSynchronize(s[0])); });
    time_it("H2D + D2H pinned concurrently (bytes each way)", [&] {
        CK(cudaMemcpyAsync(d, hp, bytes / 2, cudaMemcpyHostToDevice, s[0]));
        CK(cudaMemcpyAsync(hp + bytes / 2, d + bytes / 2, bytes / 2, cudaMemcpyDeviceToHost, s[1]));
        CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1])); });
    time_it("H2D pageable (cudaMemcpy)", [&] { CK(cudaMemcpy(d, hq, bytes, cudaMemcpyHostToDevice)); });
    time_it("D2H pageable (cudaMemcpy)", [&] { CK(cudaMemcpy(hq, d, bytes, cudaMemcpyDeviceToHost)); });
    {
        double t0 = now(); CK(cudaHostRegister(hq, bytes, cudaHostRegisterDefault)); double t1 = now();
        CK(cudaMemcpyAsync(d, hq, bytes, cudaMemcpyHostToDevice, s[0])); CK(cudaStreamSynchronize(s[0])); double t2 = now();
        CK(cudaHostUnregister(hq)); double t3 = now();
        printf("cudaHostRegister %.1f ms (%.2f GB/s), copy %.1f ms (%.2f GB/s), unregister %.1f ms; all-in %.2f GB/s\n", (t1 - t0) * 1e3,
               bytes / (t1 - t0) / 1e9, (t2 - t1) * 1e3, bytes / (t2 - t1) / 1e9, (t3 - t2) * 1e3, bytes / (t3 - t0) / 1e9);
        t0 = now(); CK(cudaHostRegister(hq, bytes, cudaHostRegisterDefault)); t1 = now(); CK(cudaHostUnregister(hq)); t2 = now();
        printf("  second time: register %.1f ms, unregister %.1f ms\n", (t1 - t0) * 1e3, (t2 - t1) * 1e3);
    }
    // staged pipeline: bounce buffers of 32 MiB, T copier threads
    const size_t B = 32 << 20;
    char* bounce[3];
    cudaEvent_t done[3];
    for (int i = 0; i < 3; i++) { CK(cudaMallocHost(&bounce[i], B)); CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); }
    for (int T : {1, 2, 4, 8}) {
        char name[96];
        snprintf(name, sizeof name, "H2D pageable via 3 x 32 MiB pinned bounce, %d memcpy threads", T);
        time_it(name, [&] {
            int k = 0;
            for (size_t off = 0; off < bytes; off += B, k = (k + 1) % 3) {
                CK(cudaEventSynchronize(done[k]));
                size_t len = std::min(B, bytes - off);
                par_memcpy(bounce[k], hq + off, len, T);
                CK(cudaMemcpyAsync(d + off, bounce[k], len, cudaMemcpyHostToDevice, s[0]));
                CK(cudaEventRecord(done[k], s[0]));
            }
            CK(cudaStreamSynchronize(s[0])); });
        snprintf(name, sizeof name, "D2H pageable via 3 x 32 MiB pinned bounce, %d memcpy threads", T);
        time_it(name, [&] {
            // chunk c: DMA into bounce[c % 3]; scatter chunk c-2 meanwhile
            std::vector<size_t> offs;
            for (size_t off = 0; off < bytes; off += B) offs.push_back(off);
            for (size_t c = 0; c < offs.size() + 2; c++) {
                if (c < offs.size()) {
                    size_t len = std::min(B, bytes - offs[c]);
                    CK(cudaMemcpyAsync(bounce[c % 3], d + offs[c], len, cudaMemcpyDeviceToHost, s[0]));
                    CK(cudaEventRecord(done[c % 3], s[0]));
                }
                if (c >= 2) {
                    size_t p = c - 2;
                    CK(cudaEventSynchronize(done[p % 3]));
                    par_memcpy(hq + offs[p], bounce[p % 3], std::min(B, bytes - offs[p]), T);
                }
            } });
    }
    time_it("host memcpy pageable->pageable 1 thread (reference point)", [&] { memcpy(hq, hp, bytes); });
    time_it("host memcpy 8 threads", [&] { par_memcpy(hq, hp, bytes, 8); });
    return 0;
}
