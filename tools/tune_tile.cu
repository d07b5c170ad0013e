// Sweep of the AoS tile-pipeline geometry (THREADS, records/thread, ring depth) for the
// hot thread-per-item kernels.  Standalone: nvcc -fmad=false ... tools/tune_tile.cu -o build/tune_tile
// Output: one CSV line per (op, config): name,threads,ipt,stages,ctas_per_sm,smem,us,GB/s
#include "../tensor_b200/csrc/t5_small_ops.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace t5;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill(double* p, size_t n, int as_float) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = splitmix64(i + 12345);
        if (as_float) { float2 v; v.x = ((float)(x >> 40) * 0x1.0p-24f) * 2.0f - 1.0f; v.y = ((float)((x >> 8) & 0xffffff) * 0x1.0p-24f) * 2.0f - 1.0f; ((float2*)p)[i] = v; }
        else p[i] = ((double)(x >> 11) * 0x1.0p-53) * 2.0 - 1.0;
    }
}

static void *g_in0, *g_in1, *g_out0, *g_out1, *g_flush;
static unsigned long long* g_sched;
static const size_t FLUSH = 256ull << 20;

template <class Op, int THREADS, int IPT, int STAGES, int MINB>
void run(const char* name, size_t n, int alg_bytes, bool flush) {
    using C = TileCfg<Op, THREADS, IPT, STAGES>;
    if (C::SMEM > 227 * 1024) return;
    auto kern = tile_kernel<Op, THREADS, IPT, STAGES, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, C::SMEM));
    if (occ < 1) return;
    size_t tiles = (n + C::TILE - 1) / C::TILE;
    int grid = (int)std::min<size_t>(tiles, (size_t)occ * 148);
    size_t claim = tiles / ((size_t)grid * 16);
    claim = claim < 1 ? 1 : (claim > 8 ? 8 : claim);
    if (getenv("T5_CLAIM")) claim = atoi(getenv("T5_CLAIM"));
    TileArgs a{g_in0, g_in1, g_out0, g_out1, n, g_sched, (unsigned)claim};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // `flush` now means: rotate over operand sets (> L2 in total) instead of re-reading one set; launches are
    // back to back with the PDL attribute, exactly as the library issues them
    constexpr size_t REC = C::I0::bytes > C::I1::bytes ? (C::I0::bytes > C::O0::bytes ? C::I0::bytes : C::O0::bytes)
                                                        : (C::I1::bytes > C::O0::bytes ? C::I1::bytes : C::O0::bytes);
    size_t fit = ((size_t)1 << 31) / (n * REC);
    const int sets = flush ? (int)(fit < 1 ? 1 : (fit > 8 ? 8 : fit)) : 1;
    auto launch = [&](int i) {
        TileArgs b = a;
        const size_t k = (size_t)(i % sets) * n;
        b.in0 = (const char*)g_in0 + k * C::I0::bytes; b.in1 = (const char*)g_in1 + k * C::I1::bytes;
        b.out0 = (char*)g_out0 + k * C::O0::bytes; b.out1 = (char*)g_out1 + k * C::O1::bytes;
        CK(pdl_launch(kern, grid, THREADS, C::SMEM, 0, b));
    };
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    const int iters = 40;
    float total = 0;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch(i);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&total, e0, e1));
    CK(cudaGetLastError());
    double us = total / iters * 1e3;
    printf("%s,%zu,%d,%d,%d,%d,%d,%d,%.2f,%.1f\n", name, n, THREADS, IPT, STAGES, MINB, occ, C::SMEM, us, alg_bytes * (double)n / us / 1e3);
    fflush(stdout);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

template <class Op>
void sweep(const char* name, size_t n, int alg_bytes, bool flush) {
#define CFG(T, I, S) run<Op, T, I, S, 1>(name, n, alg_bytes, flush);
    CFG(64, 1, 2) CFG(64, 1, 3) CFG(64, 1, 4) CFG(64, 2, 3)
    CFG(128, 1, 2) CFG(128, 1, 3) CFG(128, 1, 4) CFG(128, 2, 2) CFG(128, 2, 3)
    CFG(256, 1, 2) CFG(256, 1, 3) CFG(256, 1, 4) CFG(256, 2, 2) CFG(256, 2, 3) CFG(256, 4, 2) CFG(256, 4, 3)
#undef CFG
}

// register-capped variants (__launch_bounds__ min blocks) for the latency-bound elimination ops
template <class Op>
void sweep_minb(const char* name, size_t n, int alg_bytes, bool flush) {
    run<Op, 64, 1, 2, 8>(name, n, alg_bytes, flush);
    run<Op, 64, 1, 2, 10>(name, n, alg_bytes, flush);
    run<Op, 64, 1, 3, 8>(name, n, alg_bytes, flush);
    run<Op, 128, 1, 2, 4>(name, n, alg_bytes, flush);
    run<Op, 128, 1, 2, 5>(name, n, alg_bytes, flush);
    run<Op, 128, 1, 2, 6>(name, n, alg_bytes, flush);
    run<Op, 256, 1, 2, 2>(name, n, alg_bytes, flush);
    run<Op, 256, 1, 2, 3>(name, n, alg_bytes, flush);
}

int main(int argc, char** argv) {
    const size_t NMAX = 1ull << 24;
    const size_t BYTES = NMAX * 128;  // biggest stream: 16Mi x 128 B
    CK(cudaMalloc(&g_in0, BYTES)); CK(cudaMalloc(&g_in1, BYTES)); CK(cudaMalloc(&g_out0, BYTES)); CK(cudaMalloc(&g_out1, NMAX * 8));
    CK(cudaMalloc(&g_flush, FLUSH)); CK(cudaMalloc(&g_sched, 16)); CK(cudaMemset(g_sched, 0, 16));
    const char* which = argc > 1 ? argv[1] : "all";
    auto want = [&](const char* s) { return !strcmp(which, "all") || strstr(s, which); };
    printf("op,n,threads,ipt,stages,minb,ctas_per_sm,smem,us,GBps\n");
    // float-typed inputs
    fill<<<2048, 256>>>((double*)g_in0, BYTES / 8, 1); fill<<<2048, 256>>>((double*)g_in1, BYTES / 8, 1); CK(cudaDeviceSynchronize());
    if (want("matmul4x4_f32")) sweep<MatmulOp<float, 4, 4, 4>>("matmul4x4_f32", 1 << 24, 192, false);
    if (want("transpose4x4_f32")) sweep<TransposeOp<unsigned, 4, 4>>("transpose4x4_f32", 1 << 24, 128, false);
    if (want("matvec4_f32")) sweep<MatmulOp<float, 4, 4, 1>>("matvec4_f32", 1 << 24, 96, false);
    if (want("dot4_f32")) sweep<MatmulOp<float, 1, 4, 1>>("dot4_f32", 1 << 24, 36, false);
    // double-typed inputs
    fill<<<2048, 256>>>((double*)g_in0, BYTES / 8, 0); fill<<<2048, 256>>>((double*)g_in1, BYTES / 8, 0); CK(cudaDeviceSynchronize());
    if (want("matmul3x3_f64_1M")) sweep<MatmulOp<double, 3, 3, 3>>("matmul3x3_f64_1M", 1 << 20, 216, true);
    if (want("matmul3x3_f64_16M")) sweep<MatmulOp<double, 3, 3, 3>>("matmul3x3_f64_16M", 1 << 24, 216, false);
    if (want("matmul3x3_i64")) sweep<MatmulOp<long long, 3, 3, 3>>("matmul3x3_i64_16M", 1 << 24, 216, false);
    if (want("det3_f64")) { sweep<DetOp<double, 3>>("det3_f64_4M", 1 << 22, 80, true); sweep_minb<DetOp<double, 3>>("det3_f64_4M", 1 << 22, 80, true); }
    if (want("det4_f64")) { sweep<DetOp<double, 4>>("det4_f64_4M", 1 << 22, 136, true); sweep_minb<DetOp<double, 4>>("det4_f64_4M", 1 << 22, 136, true); }
    if (want("inverse3_f64")) { sweep<InverseOp<double, 3>>("inverse3_f64_4M", 1 << 22, 145, true); sweep_minb<InverseOp<double, 3>>("inverse3_f64_4M", 1 << 22, 145, true); }
    if (want("inverse4_f64")) { sweep<InverseOp<double, 4>>("inverse4_f64_4M", 1 << 22, 257, false); sweep_minb<InverseOp<double, 4>>("inverse4_f64_4M", 1 << 22, 257, false); }
    if (want("solve3_f64")) { sweep<SolveOp<double, 3, 1>>("solve3_f64_4M", 1 << 22, 121, true); sweep_minb<SolveOp<double, 3, 1>>("solve3_f64_4M", 1 << 22, 121, true); }
    if (want("solve4_f64")) { sweep<SolveOp<double, 4, 1>>("solve4_f64_4M", 1 << 22, 193, false); sweep_minb<SolveOp<double, 4, 1>>("solve4_f64_4M", 1 << 22, 193, false); }
    return 0;
}
