// The N-way host<->device copy ceiling: what the end-to-end (host-array) entries can reach when N GPUs of one box
// copy at the same time.  One host thread per device, pinned buffers, 32 MiB chunks (the pipeline's order of
// magnitude); three traffic patterns per device count: H2D only, D2H only, and the 2:1 H2D:D2H mix of the headline
// workload (128 B in, 64 B out per item) on two streams.  All devices start together (a spin barrier) and the
// figure is total bytes / wall time from the common start to the last device's completion.
// Standalone: nvcc -O2 -std=c++17 tools/pcie_peak_multi.cu -o build/pcie_peak_multi ; prints one JSON line.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static const size_t TOTAL = 1ull << 30, CHUNK = 32ull << 20;

struct Dev { void *d0, *d1, *h0, *h1; cudaStream_t s0, s1; };

static double run(std::vector<Dev>& devs, int n, bool h2d, bool d2h, unsigned d2h_every) {
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<std::chrono::steady_clock::time_point> done((size_t)n);
        std::vector<std::thread> th;
        for (int g = 0; g < n; ++g)
            th.emplace_back([&, g] {
                Dev& v = devs[(size_t)g];
                CK(cudaSetDevice(g));
                CK(cudaDeviceSynchronize());
                ready++;
                while (!go.load(std::memory_order_acquire)) {}
                for (size_t off = 0; off < TOTAL; off += CHUNK) {
                    if (h2d) CK(cudaMemcpyAsync((char*)v.d0 + off, (char*)v.h0 + off, CHUNK, cudaMemcpyHostToDevice, v.s0));
                    if (d2h && (off / CHUNK) % d2h_every == 0) CK(cudaMemcpyAsync((char*)v.h1 + off, (char*)v.d1 + off, CHUNK, cudaMemcpyDeviceToHost, v.s1));
                }
                CK(cudaStreamSynchronize(v.s0));
                CK(cudaStreamSynchronize(v.s1));
                done[(size_t)g] = std::chrono::steady_clock::now();
            });
        while (ready.load() < n) {}
        const auto t0 = std::chrono::steady_clock::now();
        go.store(true, std::memory_order_release);
        for (auto& t : th) t.join();
        double sec = 0;
        for (auto& d : done) { double s = std::chrono::duration<double>(d - t0).count(); if (s > sec) sec = s; }
        const double bytes = (double)TOTAL * n * ((h2d ? 1.0 : 0.0) + (d2h ? 1.0 / d2h_every : 0.0));
        if (bytes / sec / 1e9 > best) best = bytes / sec / 1e9;
    }
    return best;
}

int main() {
    int visible = 0;
    CK(cudaGetDeviceCount(&visible));
    std::vector<Dev> devs((size_t)visible);
    for (int g = 0; g < visible; ++g) {
        Dev& v = devs[(size_t)g];
        CK(cudaSetDevice(g));
        CK(cudaMalloc(&v.d0, TOTAL)); CK(cudaMalloc(&v.d1, TOTAL));
        CK(cudaHostAlloc(&v.h0, TOTAL, cudaHostAllocPortable)); CK(cudaHostAlloc(&v.h1, TOTAL, cudaHostAllocPortable));
        memset(v.h0, 1, TOTAL); memset(v.h1, 2, TOTAL);
        CK(cudaMemset(v.d0, 0, TOTAL)); CK(cudaMemset(v.d1, 0, TOTAL));
        CK(cudaStreamCreateWithFlags(&v.s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&v.s1, cudaStreamNonBlocking));
    }
    printf("{\"visible_devices\": %d, \"chunk\": \"32MiB\", \"unit\": \"GB/s (aggregate over the devices)\", \"host_threads\": \"one per device\"", visible);
    for (int n = 1; n <= visible; n *= 2) {
        printf(", \"n%d\": {\"h2d\": %.2f, \"d2h\": %.2f, \"mix_2to1\": %.2f, \"bidir\": %.2f}", n, run(devs, n, true, false, 1),
               run(devs, n, false, true, 1), run(devs, n, true, true, 2), run(devs, n, true, true, 1));
        fflush(stdout);
    }
    printf("}\n");
    return 0;
}
