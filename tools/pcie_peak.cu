// Measures the host<->device copy ceilings the end-to-end (host-buffer) entries are bound by:
// pinned H2D alone, pinned D2H alone, and both directions at once on two streams, for the chunk
// size the pipeline uses (32 MiB) and for one large copy.
// Standalone: nvcc -O2 tools/pcie_peak.cu -o build/pcie_peak ; prints one JSON line.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static double run(void* d0, void* d1, void* h0, void* h1, size_t total, size_t chunk, bool h2d, bool d2h, unsigned d2h_every) {
    cudaStream_t s0, s1;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    cudaEvent_t b0, e0, b1, e1;
    CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&b1)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(b0, s0)); CK(cudaEventRecord(b1, s1));
        for (size_t off = 0; off < total; off += chunk) {
            if (h2d) CK(cudaMemcpyAsync((char*)d0 + off, (char*)h0 + off, chunk, cudaMemcpyHostToDevice, s0));
            if (d2h && (off / chunk) % d2h_every == 0) CK(cudaMemcpyAsync((char*)h1 + off, (char*)d1 + off, chunk, cudaMemcpyDeviceToHost, s1));
        }
        CK(cudaEventRecord(e0, s0)); CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
        float m0, m1, m01, m10;
        CK(cudaEventElapsedTime(&m0, b0, e0)); CK(cudaEventElapsedTime(&m1, b1, e1));
        CK(cudaEventElapsedTime(&m01, b0, e1)); CK(cudaEventElapsedTime(&m10, b1, e0));
        float ms = m0 > m1 ? m0 : m1; if (m01 > ms) ms = m01; if (m10 > ms) ms = m10;
        double gbs = (double)total * ((h2d ? 1.0 : 0.0) + (d2h ? 1.0 / d2h_every : 0.0)) / (ms * 1e6);
        if (gbs > best) best = gbs;
    }
    return best;
}

int main() {
    const size_t total = 1ull << 30;
    void *d0, *d1, *h0, *h1, *hw;
    CK(cudaMalloc(&d0, total)); CK(cudaMalloc(&d1, total));
    CK(cudaHostAlloc(&h0, total, cudaHostAllocDefault)); CK(cudaHostAlloc(&h1, total, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hw, total, cudaHostAllocWriteCombined));
    memset(h0, 1, total); memset(h1, 2, total); memset(hw, 3, total);
    CK(cudaMemset(d0, 0, total)); CK(cudaMemset(d1, 0, total));
    printf("{");
    const size_t chunks[3] = {32ull << 20, 256ull << 20, 1ull << 30};
    const char* names[3] = {"32MiB", "256MiB", "1GiB"};
    for (int i = 0; i < 3; ++i) {
        printf("\"h2d_%s\": %.2f, ", names[i], run(d0, d1, h0, h1, total, chunks[i], true, false, 1));
        printf("\"d2h_%s\": %.2f, ", names[i], run(d0, d1, h0, h1, total, chunks[i], false, true, 1));
        printf("\"bidir_sum_%s\": %.2f, ", names[i], run(d0, d1, h0, h1, total, chunks[i], true, true, 1));
    }
    printf("\"h2d_writecombined_32MiB\": %.2f, ", run(d0, d1, hw, h1, total, 32ull << 20, true, false, 1));
    printf("\"bidir_sum_writecombined_32MiB\": %.2f, ", run(d0, d1, hw, h1, total, 32ull << 20, true, true, 1));
    printf("\"bidir_sum_2to1_32MiB\": %.2f, ", run(d0, d1, h0, h1, total, 32ull << 20, true, true, 2));
    printf("\"unit\": \"GB/s\"}\n");
    return 0;
}
