// h2d_multi.cu -- developer probe: concurrent pinned H2D bandwidth on every GPU of the box, with the
// copying thread (and therefore its pinned buffer, first touch) left where the OS put it vs bound to
// the GPU's NUMA-local CPUs (/sys/bus/pci/devices/<bdf>/local_cpulist).
// nvcc -O3 -std=c++17 -o h2d_multi h2d_multi.cu -lpthread
#include <cuda_runtime.h>
#include <sched.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static std::string read_file(const std::string& p)
{
    FILE* f = fopen(p.c_str(), "r");
    if (!f) return "";
    char buf[4096];
    size_t k = fread(buf, 1, sizeof buf - 1, f);
    fclose(f);
    buf[k] = 0;
    while (k && (buf[k - 1] == '\n' || buf[k - 1] == ' ')) buf[--k] = 0;
    return buf;
}
static std::string bdf_of(int dev)
{
    char id[32];
    cudaDeviceGetPCIBusId(id, sizeof id, dev);
    for (char* c = id; *c; c++) *c = (char)tolower(*c);
    return id;
}
static bool bind_to_cpulist(const std::string& list)
{
    cpu_set_t set;
    CPU_ZERO(&set);
    const char* p = list.c_str();
    int n = 0;
    while (*p) {
        int a = (int)strtol(p, (char**)&p, 10), b = a;
        if (*p == '-') b = (int)strtol(p + 1, (char**)&p, 10);
        for (int c = a; c <= b; c++) { CPU_SET(c, &set); n++; }
        if (*p == ',') p++;
    }
    return n && sched_setaffinity(0, sizeof set, &set) == 0;
}

int main()
{
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    printf("%d GPUs, %u host threads\n", ndev, std::thread::hardware_concurrency());
    for (int d = 0; d < ndev; d++) {
        const std::string bdf = bdf_of(d);
        printf("GPU %d %s numa_node %s local_cpulist %s\n", d, bdf.c_str(), read_file("/sys/bus/pci/devices/" + bdf + "/numa_node").c_str(),
               read_file("/sys/bus/pci/devices/" + bdf + "/local_cpulist").c_str());
    }
    const size_t bytes = 1ull << 30;
    for (int mode = 0; mode < 2; mode++) {
        for (int active : {1, 2, 4, 8}) {
            if (active > ndev) break;
            std::vector<double> gbs(active, 0.0);
            std::atomic<int> ready{0}, go{0};
            std::vector<std::thread> th;
            for (int d = 0; d < active; d++)
                th.emplace_back([&, d] {
                    if (mode == 1) bind_to_cpulist(read_file("/sys/bus/pci/devices/" + bdf_of(d) + "/local_cpulist"));
                    cudaSetDevice(d);
                    void *h, *g;
                    cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
                    memset(h, 1, bytes);
                    cudaMalloc(&g, bytes);
                    cudaStream_t s;
                    cudaStreamCreate(&s);
                    cudaMemcpyAsync(g, h, bytes, cudaMemcpyHostToDevice, s);
                    cudaStreamSynchronize(s);
                    ready++;
                    while (!go.load()) {}
                    auto t0 = std::chrono::steady_clock::now();
                    for (int r = 0; r < 4; r++) cudaMemcpyAsync(g, h, bytes, cudaMemcpyHostToDevice, s);
                    cudaStreamSynchronize(s);
                    gbs[d] = 4.0 * bytes / std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 1e9;
                    cudaFree(g);
                    cudaFreeHost(h);
                });
            while (ready.load() < active) {}
            go = 1;
            for (auto& t : th) t.join();
            double sum = 0;
            printf("%s, %d GPUs copying:", mode ? "bound to local_cpulist" : "unbound", active);
            for (double x : gbs) { printf(" %.1f", x); sum += x; }
            printf("  | total %.1f GB/s\n", sum);
        }
    }
    return 0;
}
