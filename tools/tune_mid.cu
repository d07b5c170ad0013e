// Sweep of the row-per-thread contraction kernel's geometry (THREADS, ring depth) for the 8x8 and
// 16x16 products.  Standalone:
//   nvcc -std=c++17 -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a tools/tune_mid.cu -o build/tune_mid
// Output: CSV  shape,threads,stages,us,GB/s  (rotating operand sets larger than L2, CUDA events).
#include "../tensor_b200/csrc/t5_mid_ops.cu"
#include <cstdio>
#include <cstdlib>
using namespace t5;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill(double* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = ((double)(splitmix64(i + 12345) >> 11) * 0x1.0p-53) * 2.0 - 1.0;
}

static double *g_a, *g_b, *g_c;
static unsigned long long* g_sched;
static const int SETS = 2;

template <int P, int THREADS, int STAGES>
void run(size_t n) {
    if constexpr (mid::Cfg<double, P, P, P, THREADS, STAGES>::SMEM <= 227 * 1024) {
        Shard s{}; s.stream = 0; s.layout = T5L_AOS; s.count = n; s.sched = g_sched;
        const size_t d = (size_t)P * P;
        for (int i = 0; i < 3; ++i) if (mid::launch<double, P, P, P, THREADS, STAGES>(s, g_a, g_b, g_c) != T5I_OK) { printf("launch failed\n"); return; }
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int iters = 20;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) {
            const size_t off = (size_t)(i % SETS) * n * d;
            mid::launch<double, P, P, P, THREADS, STAGES>(s, g_a + off, g_b + off, g_c + off);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms / iters * 1e3;
        printf("%dx%d,%d,%d,%.2f,%.1f\n", P, P, THREADS, STAGES, us, 3.0 * d * 8 * n / us / 1e3);
        fflush(stdout);
    }
}

int main() {
    const size_t n = 1 << 20, d = 64;
    CK(cudaMalloc(&g_a, SETS * n * d * 8)); CK(cudaMalloc(&g_b, SETS * n * d * 8)); CK(cudaMalloc(&g_c, SETS * n * d * 8));
    CK(cudaMalloc(&g_sched, 16)); CK(cudaMemset(g_sched, 0, 16));
    fill<<<2048, 256>>>(g_a, SETS * n * d); fill<<<2048, 256>>>(g_b, SETS * n * d); CK(cudaDeviceSynchronize());
    printf("shape,threads,stages,us,GBps\n");
    run<8, 64, 3>(n); run<8, 64, 4>(n); run<8, 64, 6>(n);
    run<8, 128, 2>(n); run<8, 128, 3>(n); run<8, 128, 4>(n); run<8, 128, 6>(n);
    run<8, 256, 2>(n); run<8, 256, 3>(n); run<8, 256, 4>(n); run<8, 256, 6>(n);
    run<8, 512, 2>(n); run<8, 512, 3>(n); run<8, 512, 4>(n);
    run<8, 1024, 2>(n); run<8, 1024, 3>(n);
    const size_t n16 = 1 << 18;   // 16x16: 2 KiB records, same bytes
    run<16, 128, 2>(n16); run<16, 128, 3>(n16); run<16, 256, 2>(n16); run<16, 256, 3>(n16); run<16, 512, 2>(n16); run<16, 512, 3>(n16);
    return 0;
}
