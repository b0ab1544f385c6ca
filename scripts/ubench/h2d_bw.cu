// h2d_bw.cu — what the box gives pinned host -> device copies, per GPU alone and 2 / 4 / 8 GPUs at once.
//
// The e2e metric of bench.py moves 74.6 MB per 4K evaluation over PCIe; this micro-benchmark is the ceiling that
// number is reported against (profiles/r2_h2d_bw_*.txt).  One host thread per GPU, each with its own pinned buffer
// and stream, copying `--mb` megabytes back to back for `--ms` milliseconds; CUDA events time each GPU's stream.
//
//   --bind none    threads and their pinned buffers stay wherever the OS puts them
//   --bind local   each thread first binds to the GPU's local CPUs (sysfs local_cpulist of the PCI function),
//                  then allocates and first-touches its pinned buffer (node-local pages)
//   --bind spread  the visible CPUs are split evenly over the GPUs (for boxes whose sysfs shows one NUMA node)
//
// Build: nvcc -O2 -std=c++17 -o scripts/ubench/h2d_bw scripts/ubench/h2d_bw.cu -lpthread
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

static std::string slurp(const std::string &path)
{
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    return s;
}

static std::vector<int> parse_cpulist(const std::string &s)
{
    std::vector<int> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        const size_t dash = tok.find('-');
        const int a = atoi(tok.substr(0, dash).c_str());
        const int b = dash == std::string::npos ? a : atoi(tok.substr(dash + 1).c_str());
        for (int c = a; c <= b; ++c) out.push_back(c);
    }
    return out;
}

static std::string pci_dir(int dev)
{
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, sizeof id, dev) != cudaSuccess) return "";
    for (char *p = id; *p; ++p) *p = (char)tolower(*p);
    return std::string("/sys/bus/pci/devices/") + id;
}

static bool bind_cpus(const std::vector<int> &cpus)
{
    if (cpus.empty()) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    return sched_setaffinity(0, sizeof set, &set) == 0;
}

struct Result {
    double gbs = 0.0;
    std::string bound;
};

static void worker(int dev, int idx, int ngpu, const std::string &bind, size_t bytes, int ms, std::atomic<int> *ready,
                   std::atomic<bool> *go, Result *res)
{
    if (bind == "local") {
        const std::string l = slurp(pci_dir(dev) + "/local_cpulist");
        res->bound = bind_cpus(parse_cpulist(l)) ? l : "(unavailable)";
    } else if (bind == "spread") {
        cpu_set_t all;
        sched_getaffinity(0, sizeof all, &all);
        std::vector<int> cpus;
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &all)) cpus.push_back(c);
        const size_t per = cpus.size() / ngpu;
        std::vector<int> mine(cpus.begin() + idx * per, cpus.begin() + (idx + 1) * per);
        char buf[64];
        snprintf(buf, sizeof buf, "%d..%d", mine.empty() ? -1 : mine.front(), mine.empty() ? -1 : mine.back());
        res->bound = bind_cpus(mine) ? buf : "(unavailable)";
    }
    cudaSetDevice(dev);
    void *h = nullptr, *d = nullptr;
    cudaStream_t st;
    cudaEvent_t e0, e1;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&d, bytes) != cudaSuccess) {
        fprintf(stderr, "gpu %d: allocation failed\n", dev);
        ready->fetch_add(1);
        return;
    }
    memset(h, 1, bytes);   // first touch by the bound thread
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    ready->fetch_add(1);
    while (!go->load()) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    long long copies = 0;
    cudaEventRecord(e0, st);
    while (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() < ms) {
        for (int i = 0; i < 4; ++i) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
        copies += 4;
        cudaStreamSynchronize(st);
    }
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    res->gbs = (double)copies * bytes / 1e9 / (t / 1e3);
    cudaFreeHost(h);
    cudaFree(d);
}

static void run_set(const std::vector<int> &devs, const std::string &bind, size_t bytes, int ms)
{
    std::vector<Result> res(devs.size());
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (size_t i = 0; i < devs.size(); ++i)
        th.emplace_back(worker, devs[i], (int)i, (int)devs.size(), bind, bytes, ms, &ready, &go, &res[i]);
    while (ready.load() < (int)devs.size()) std::this_thread::yield();
    go.store(true);
    for (auto &t : th) t.join();
    double sum = 0.0;
    printf("bind=%-6s gpus=[", bind.c_str());
    for (size_t i = 0; i < devs.size(); ++i) printf("%s%d", i ? "," : "", devs[i]);
    printf("]  per-GPU GB/s:");
    for (size_t i = 0; i < devs.size(); ++i) {
        printf(" %.1f", res[i].gbs);
        sum += res[i].gbs;
    }
    printf("  | aggregate %.1f GB/s, mean %.1f", sum, sum / devs.size());
    if (bind != "none") printf("  (gpu %d bound to cpus %s)", devs[0], res[0].bound.c_str());
    printf("\n");
    fflush(stdout);
}

int main(int argc, char **argv)
{
    size_t mb = 128;
    int ms = 600;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--mb") && i + 1 < argc) mb = (size_t)atoi(argv[++i]);
        else if (!strcmp(argv[i], "--ms") && i + 1 < argc) ms = atoi(argv[++i]);
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    printf("host: %ld online CPUs; NUMA nodes:", sysconf(_SC_NPROCESSORS_ONLN));
    for (int node = 0; node < 16; ++node) {
        const std::string l = slurp("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
        if (!l.empty()) printf(" node%d={%s}", node, l.c_str());
    }
    printf("\n");
    for (int d = 0; d < n; ++d) {
        const std::string dir = pci_dir(d);
        printf("gpu %d: %s numa_node=%s local_cpulist=%s\n", d, dir.c_str(), slurp(dir + "/numa_node").c_str(),
               slurp(dir + "/local_cpulist").c_str());
    }
    const size_t bytes = mb << 20;
    for (const char *bind : {"none", "local", "spread"}) {
        for (int d = 0; d < n; ++d) run_set({d}, bind, bytes, ms);          // each GPU alone
        for (int k = 2; k <= n; k *= 2) {                                     // 2, 4, 8 at once
            std::vector<int> devs;
            for (int d = 0; d < k; ++d) devs.push_back(d);
            run_set(devs, bind, bytes, ms);
        }
    }
    return 0;
}
