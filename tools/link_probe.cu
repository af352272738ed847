// Developer probe (not part of the product): what the host link of this box gives N GPUs at once.
// One host thread per GPU; every thread moves pinned H2D and D2H copies on two streams, all GPUs released together.
// Prints one JSON object per configuration: which GPUs, copy shape (1-D / 2-D strided), pinned-memory flavour
// (default / write-combined for the H2D source), NUMA placement (thread bound to the GPU's node before allocating, or not).
// The end-to-end numbers of bench.py (`e2e.link_frac`) are quoted against the duplex figure measured here.
//   nvcc -O2 -std=c++17 -o tools/link_probe tools/link_probe.cu && ./tools/link_probe [maxGpus] [GiB per buffer]
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

static int numa_node_of_gpu(int dev, char* busId /* [32] */) {
    busId[0] = 0;
    if (cudaDeviceGetPCIBusId(busId, 32, dev) != cudaSuccess) return -1;
    for (char* p = busId; *p; ++p) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", busId);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// binds the calling thread to the CPUs of a NUMA node (first-touch then places its pinned allocations there)
static bool bind_to_node(int node) {
    if (node < 0) return false;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char buf[4096];
    if (!fgets(buf, sizeof buf, f)) { fclose(f); return false; }
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a, b;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; ++c) { CPU_SET(c, &set); ++n; } }
        else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); ++n; }
    }
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

struct Cfg { int nGpus; bool twoD, wc, numa, h2d, d2h; };

struct Res { double h2dGBs = 0, d2hGBs = 0; int node = -1; bool bound = false; };

static void worker(int dev, const Cfg& c, size_t bytes, std::atomic<int>* ready, std::atomic<int>* go, Res* out) {
    cudaSetDevice(dev);
    char bus[32];
    out->node = numa_node_of_gpu(dev, bus);
    if (c.numa) out->bound = bind_to_node(out->node);
    void *hIn = nullptr, *hOut = nullptr, *dIn = nullptr, *dOut = nullptr;
    cudaHostAlloc(&hIn, bytes, c.wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    cudaHostAlloc(&hOut, bytes, cudaHostAllocDefault);
    cudaMalloc(&dIn, bytes);
    cudaMalloc(&dOut, bytes);
    memset(hIn, 1, bytes);
    memset(hOut, 0, bytes);
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t a0, a1, b0, b1;
    cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
    // 2-D: rows of 1 MiB at a 1 MiB + 4 KiB host pitch inside the same buffer (what a strided caller layout looks like)
    const size_t rowB = (size_t)1 << 20, pitch = rowB + 4096;
    const size_t rows = bytes / pitch;
    auto issue = [&]() {
        if (c.h2d) {
            if (c.twoD) cudaMemcpy2DAsync(dIn, rowB, hIn, pitch, rowB, rows, cudaMemcpyHostToDevice, a);
            else cudaMemcpyAsync(dIn, hIn, bytes, cudaMemcpyHostToDevice, a);
        }
        if (c.d2h) {
            if (c.twoD) cudaMemcpy2DAsync(hOut, pitch, dOut, rowB, rowB, rows, cudaMemcpyDeviceToHost, b);
            else cudaMemcpyAsync(hOut, dOut, bytes, cudaMemcpyDeviceToHost, b);
        }
    };
    issue();
    cudaDeviceSynchronize();
    ready->fetch_add(1);
    while (go->load() == 0) std::this_thread::yield();
    const int reps = 3;
    cudaEventRecord(a0, a); cudaEventRecord(b0, b);
    for (int r = 0; r < reps; ++r) issue();
    cudaEventRecord(a1, a); cudaEventRecord(b1, b);
    cudaDeviceSynchronize();
    float msA = 0, msB = 0;
    cudaEventElapsedTime(&msA, a0, a1);
    cudaEventElapsedTime(&msB, b0, b1);
    const double moved = (double)(c.twoD ? rows * rowB : bytes) * reps;
    if (c.h2d && msA > 0) out->h2dGBs = moved / (msA * 1e-3) / 1e9;
    if (c.d2h && msB > 0) out->d2hGBs = moved / (msB * 1e-3) / 1e9;
    cudaFreeHost(hIn); cudaFreeHost(hOut); cudaFree(dIn); cudaFree(dOut);
    cudaStreamDestroy(a); cudaStreamDestroy(b);
}

int main(int argc, char** argv) {
    int nDev = 0;
    cudaGetDeviceCount(&nDev);
    int maxG = argc > 1 ? atoi(argv[1]) : nDev;
    if (maxG > nDev) maxG = nDev;
    const double gib = argc > 2 ? atof(argv[2]) : 2.0;
    const size_t bytes = (size_t)(gib * 1073741824.0);
    printf("{\"probe\": \"link\", \"gpus_visible\": %d, \"buffer_gib\": %.2f, \"host_cpus\": %ld}\n", nDev, gib, sysconf(_SC_NPROCESSORS_ONLN));
    std::vector<Cfg> cfgs;
    for (int n = 1; n <= maxG; n *= 2) {
        cfgs.push_back({n, false, false, false, true, false});
        cfgs.push_back({n, false, false, false, false, true});
        cfgs.push_back({n, false, false, false, true, true});
        cfgs.push_back({n, true, false, false, true, true});
        cfgs.push_back({n, false, true, false, true, true});
        cfgs.push_back({n, false, false, true, true, true});
    }
    for (const Cfg& c : cfgs) {
        std::vector<Res> res(c.nGpus);
        std::atomic<int> ready{0}, go{0};
        std::vector<std::thread> th;
        for (int d = 0; d < c.nGpus; ++d) th.emplace_back(worker, d, std::cref(c), bytes, &ready, &go, &res[d]);
        while (ready.load() < c.nGpus) std::this_thread::yield();
        go.store(1);
        for (auto& t : th) t.join();
        double sumIn = 0, sumOut = 0;
        std::string nodes = "[", per = "[";
        for (int d = 0; d < c.nGpus; ++d) {
            sumIn += res[d].h2dGBs; sumOut += res[d].d2hGBs;
            char b[96];
            snprintf(b, sizeof b, "%s%d", d ? ", " : "", res[d].node); nodes += b;
            snprintf(b, sizeof b, "%s[%.1f, %.1f]", d ? ", " : "", res[d].h2dGBs, res[d].d2hGBs); per += b;
        }
        nodes += "]"; per += "]";
        printf("{\"gpus\": %d, \"h2d\": %s, \"d2h\": %s, \"copy\": \"%s\", \"pinned\": \"%s\", \"numa_bound\": %s, \"h2d_gbs_total\": %.1f, "
               "\"d2h_gbs_total\": %.1f, \"duplex_gbs_total\": %.1f, \"gpu_numa_nodes\": %s, \"per_gpu_gbs\": %s}\n",
               c.nGpus, c.h2d ? "true" : "false", c.d2h ? "true" : "false", c.twoD ? "2d" : "1d", c.wc ? "write-combined" : "default",
               (c.numa && res[0].bound) ? "true" : "false", sumIn, sumOut, sumIn + sumOut, nodes.c_str(), per.c_str());
        fflush(stdout);
    }
    return 0;
}
