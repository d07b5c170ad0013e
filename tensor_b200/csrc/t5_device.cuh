// Device-side helpers shared by every kernel translation unit (sm_100a only).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

#ifndef T5_SM_COUNT_FALLBACK
#define T5_SM_COUNT_FALLBACK 148
#endif

namespace t5 {

// ---- cp.async (LDGSTS) 16-byte global->shared copy, L2-only (.cg) ------------
// src_bytes < 16 zero-fills the remainder (ragged batch tails).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes = 16) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
template <int BYTES>   // 4- or 8-byte asynchronous copies (rows that are not whole 16-byte chunks)
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gsrc) {
    static_assert(BYTES == 4 || BYTES == 8, "cp.async.ca moves 4, 8 or 16 bytes");
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- programmatic dependent launch (PDL) ----------------------------------------
// The persistent kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization
// (pdl_launch, below): `pdl_trigger` lets the NEXT kernel of the stream be rasterised onto SMs as
// this grid's CTAs retire, `pdl_wait` blocks it until this grid has finished and its writes (results
// and the re-armed tile-scheduler words) are visible.  That hides the ~2-3 us launch/ramp gap that
// dominates the 1Mi-item configs (a 226 MB stream is 35 us of HBM time).  Both are no-ops for a
// plain <<<>>> launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

template <class Kern, class Args>
static inline cudaError_t pdl_launch(Kern kern, int grid, int threads, size_t smem, cudaStream_t stream, const Args& args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args);
}

// ---- streaming 16-byte global store / load (no L1 allocation) -----------------
__device__ __forceinline__ void st_global_stream16(void* gdst, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"l"(gdst), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_global_stream16(const void* gsrc) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(gsrc));
    return v;
}

// ---- element arithmetic: Int64 wraps (computed unsigned); FP never fuses
// (the TUs that include this are compiled with -fmad=false) ---------------------
template <class T> struct Arith {
    __device__ __forceinline__ static T zero() { return T(0); }
    __device__ __forceinline__ static T mul(T a, T b) { return a * b; }
    __device__ __forceinline__ static T add(T a, T b) { return a + b; }
    __device__ __forceinline__ static T sub(T a, T b) { return a - b; }
};
template <> struct Arith<long long> {
    using U = unsigned long long;
    __device__ __forceinline__ static long long zero() { return 0; }
    __device__ __forceinline__ static long long mul(long long a, long long b) { return (long long)((U)a * (U)b); }
    __device__ __forceinline__ static long long add(long long a, long long b) { return (long long)((U)a + (U)b); }
    __device__ __forceinline__ static long long sub(long long a, long long b) { return (long long)((U)a - (U)b); }
};

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- shared-memory item access: widest vector the item's byte size allows ----
// (items whose size is a multiple of 16 B are laid out with an odd 16-B stride,
// 8·odd and 4·odd sizes are naturally conflict-free; see t5_tile.cuh).
template <int BYTES>
__device__ __forceinline__ void lds_item(const unsigned char* p, void* regs) {
    if constexpr (BYTES % 16 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4*>(regs)[i] = reinterpret_cast<const uint4*>(p)[i];
    } else if constexpr (BYTES % 8 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2*>(regs)[i] = reinterpret_cast<const uint2*>(p)[i];
    } else if constexpr (BYTES % 4 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 4; ++i) reinterpret_cast<uint32_t*>(regs)[i] = reinterpret_cast<const uint32_t*>(p)[i];
    } else {
        #pragma unroll
        for (int i = 0; i < BYTES; ++i) reinterpret_cast<unsigned char*>(regs)[i] = p[i];
    }
}
template <int BYTES>
__device__ __forceinline__ void sts_item(unsigned char* p, const void* regs) {
    if constexpr (BYTES % 16 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 16; ++i) reinterpret_cast<uint4*>(p)[i] = reinterpret_cast<const uint4*>(regs)[i];
    } else if constexpr (BYTES % 8 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2*>(p)[i] = reinterpret_cast<const uint2*>(regs)[i];
    } else if constexpr (BYTES % 4 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 4; ++i) reinterpret_cast<uint32_t*>(p)[i] = reinterpret_cast<const uint32_t*>(regs)[i];
    } else {
        #pragma unroll
        for (int i = 0; i < BYTES; ++i) p[i] = reinterpret_cast<const unsigned char*>(regs)[i];
    }
}

}  // namespace t5
