// AoS tile pipeline: the skeleton every thread-per-item kernel runs in.
//
// A batch is `n_items` records per stream, records concatenated (the reference's
// wire order, src/Data/Tensor/Vector.hs:360-367).  Records are 4..128 bytes, so
// a thread reading "its" record straight from HBM would touch 32 scattered
// sectors per warp instruction.  Instead a persistent CTA walks tiles of
// TILE = THREADS*IPT records:
//
//   global --cp.async 16 B (LDGSTS, coalesced, STAGES-deep ring)--> shared slots
//   slot   --LDS (record-per-thread, conflict-free by construction)--> registers
//   registers: Op::run  (pure register arithmetic, shared with the SoA skeleton)
//   result <= 16 B : STG straight from registers (consecutive threads, consecutive
//                    addresses: already coalesced)
//   result  > 16 B : written IN PLACE over the thread's own input slot, then
//                    LDS.128 + STG.128 (coalesced, streaming) by the whole CTA
//
// Conflict-free slots.  A record of B bytes occupies a slot of `stride` bytes:
//   B in {32, 64, 128}      : stride B, 16-B chunks XOR-swizzled inside each 128-B row
//                             (chunk w of record i lives at w ^ ((i >> (3-log2 c)) & (c-1)),
//                             c = B/16) - the same pattern TMA's 32/64/128B swizzle uses;
//   B = 16*even (not 2^k)   : stride B+16 (odd number of chunks);
//   B = 16*odd, 8*odd, 4*odd: stride B (LDS.128 / LDS.64 / LDS.32 hit distinct banks).
//
// Tiles are handed out by a global atomic counter (dynamic scheduling): with a static
// round-robin the slowest SM (far-die L2, busier DRAM partition) finished ~17% after
// the average one (ncu, profiles/r01_*), which capped DRAM utilisation.
#pragma once
#include "t5_device.cuh"
#include <type_traits>

namespace t5 {

// Per-launch scalars of an Op (today: the coefficients of polyEval).  Travels inside the
// kernel parameter block, i.e. the constant bank: reading v[j] is an LDC, not a global load.
struct OpParams {
    int n;                       // number of valid entries of v
    unsigned long long v[16];    // bit patterns of the element type, low-order bytes first
};
template <class T> __device__ __forceinline__ T param_as(unsigned long long raw) {
    if constexpr (sizeof(T) == 8) { T x; memcpy(&x, &raw, 8); return x; }
    else { unsigned lo = (unsigned)raw; T x; memcpy(&x, &lo, 4); return x; }
}
template <class Op, class = void> struct UsesParams : std::false_type {};
template <class Op> struct UsesParams<Op, std::void_t<decltype(Op::uses_params)>> : std::true_type {};
template <class Op, class A0, class A1, class B0, class B1>
__device__ __forceinline__ void run_op(const A0* r0, const A1* r1, B0* w0, B1* w1, const OpParams& prm) {
    if constexpr (UsesParams<Op>::value) Op::run(r0, r1, w0, w1, prm);
    else Op::run(r0, r1, w0, w1);
}

// Ops whose result record does not fit in registers next to their working set (the 7x7 / 8x8 Double
// inverse: 64 + 64 doubles) declare `streams_out0` and provide
//   run_stream(const In0::type*, const In1::type*, Writer&, Out1::type*)
// which hands each result element to the writer as soon as it is final: into the thread's own
// (already consumed) input slot in the AoS pipeline, straight to its row in the SoA one.
template <class Op, class = void> struct StreamsOut : std::false_type {};
template <class Op> struct StreamsOut<Op, std::void_t<decltype(Op::streams_out0)>> : std::true_type {};

constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

template <class T, int N>
struct Rec {
    using type = T;
    static constexpr int n = N;
    static constexpr int bytes = N * (int)sizeof(T);
    static constexpr int cpi = bytes / 16;                                   // 16-B chunks per record (if bytes % 16 == 0)
    static constexpr bool vec16 = bytes > 0 && bytes % 16 == 0;
    static constexpr bool swz = vec16 && (cpi == 2 || cpi == 4 || cpi == 8); // power-of-two records: XOR swizzle
    static constexpr int pad = (vec16 && !swz && (cpi % 2 == 0)) ? 16 : 0;
    static constexpr int stride = bytes + pad;
    static constexpr int swz_shift = swz ? 3 - ilog2(cpi) : 0;

    // byte offset of 16-B chunk w of record `item` inside a stage
    __device__ __forceinline__ static int chunk_off(int item, int w) {
        if constexpr (swz) return item * bytes + ((w ^ ((item >> swz_shift) & (cpi - 1))) << 4);
        else return item * stride + (w << 4);
    }
    // byte offset of the ch-th 16-B chunk of the tile's contiguous global image
    __device__ __forceinline__ static int tile_chunk_off(int ch) {
        if constexpr (vec16 && (swz || pad)) return chunk_off(ch / cpi, ch % cpi);
        else return ch << 4;
    }
    // byte offset of element e of record `item` (for ops that write results element by element)
    __device__ __forceinline__ static int elem_off(int item, int e) {
        constexpr int sz = (int)sizeof(T);
        if constexpr (vec16) return chunk_off(item, (e * sz) >> 4) + ((e * sz) & 15);
        else return item * stride + e * sz;
    }
    __device__ __forceinline__ static void load(const unsigned char* stage, int item, void* regs) {
        if constexpr (vec16) {
            #pragma unroll
            for (int w = 0; w < cpi; ++w) reinterpret_cast<uint4*>(regs)[w] = *reinterpret_cast<const uint4*>(stage + chunk_off(item, w));
        } else {
            lds_item<bytes>(stage + item * stride, regs);
        }
    }
    __device__ __forceinline__ static void store(unsigned char* stage, int item, const void* regs) {
        if constexpr (vec16) {
            #pragma unroll
            for (int w = 0; w < cpi; ++w) *reinterpret_cast<uint4*>(stage + chunk_off(item, w)) = reinterpret_cast<const uint4*>(regs)[w];
        } else {
            sts_item<bytes>(stage + item * stride, regs);
        }
    }
};
using NoRec = Rec<unsigned char, 0>;

template <class R>
struct RegArr {
    alignas(16) typename R::type v[R::n > 0 ? R::n : 1];
};

template <class Op, class = void> struct LazyIn1 : std::false_type {};
template <class Op> struct LazyIn1<Op, std::void_t<decltype(Op::lazy_in1)>> : std::true_type {};

template <class R>
struct SlotReader {                 // second operand read on demand from the thread's slot (run-time element index)
    const unsigned char* stage;
    int item;
    __device__ __forceinline__ typename R::type get(int e) const {
        return *reinterpret_cast<const typename R::type*>(stage + R::elem_off(item, e));
    }
};
template <class T>
struct SoaReader {
    const T* base;                  // row 0 of this item inside the staged tile
    int pitch;                      // items per staged row
    __device__ __forceinline__ T get(int e) const { return base[e * pitch]; }
};

template <class R>
struct SlotWriter {
    static constexpr bool in_place = true;          // the results replace the thread's input record in shared memory
    unsigned char* stage;
    int item;
    template <int E>
    __device__ __forceinline__ void put(typename R::type v) const {
        *reinterpret_cast<typename R::type*>(stage + R::elem_off(item, E)) = v;
    }
};
template <class T>
struct SoaWriter {
    static constexpr bool in_place = false;
    T* base;                        // row 0 of this item
    unsigned long long stride;
    template <int E>
    __device__ __forceinline__ void put(T v) const { base[(size_t)E * stride] = v; }
};


// Ops that park part of their working set in the thread's own (already loaded) In0 staging area declare
// `scratch_in0` and take a fifth run_stream argument with compile-time-indexed get<E>() / put<E>(v): the thread's
// input slot in the AoS pipeline (the SAME bytes the streamed results go to when the result record replaces In0:
// the op must only overwrite what it no longer reads), its column of the staged tile in the SoA one.
template <class Op, class = void> struct ScratchIn0 : std::false_type {};
template <class Op> struct ScratchIn0<Op, std::void_t<decltype(Op::scratch_in0)>> : std::true_type {};
template <class R>
struct SlotScratch {
    unsigned char* stage;
    int item;
    template <int E> __device__ __forceinline__ typename R::type get() const {
        return *reinterpret_cast<const typename R::type*>(stage + R::elem_off(item, E));
    }
    template <int E> __device__ __forceinline__ void put(typename R::type v) const {
        *reinterpret_cast<typename R::type*>(stage + R::elem_off(item, E)) = v;
    }
};
template <class T, int PITCH>
struct SoaScratch {
    T* base;                        // this item's column of the staged [element][item] tile
    template <int E> __device__ __forceinline__ T get() const { return base[E * PITCH]; }
    template <int E> __device__ __forceinline__ void put(T v) const { base[E * PITCH] = v; }
};

struct TileArgs {
    const void* in0;
    const void* in1;
    void* out0;
    void* out1;
    unsigned long long n_items;
    unsigned long long* sched;  // [0] next tile, [1] finished CTAs; both zero between launches
    unsigned claim;             // tiles claimed per atomic (>= 1)
    OpParams prm;
};

// Tile claims: one atomicAdd on the scheduler word hands out `claim` consecutive tiles.  A
// single L2 address sustains only ~300 M atomics/s chip-wide (measured: the 16Mi-record dot
// kernel was capped at exactly that tile rate), so large batches claim several tiles at once.
struct TileClaimer {
    unsigned long long base = 0;
    unsigned left = 0;
    __device__ __forceinline__ unsigned long long next(unsigned long long* sched, unsigned claim) {
        if (left == 0) { base = atomicAdd(sched, (unsigned long long)claim); left = claim; }
        --left;
        return base++;
    }
};

// ---- cooperative tile movement ------------------------------------------------
template <class R, int TILE, int THREADS>
__device__ __forceinline__ void tile_load_async(unsigned char* stage, const unsigned char* gsrc, int n_valid, int tid) {
    if constexpr (R::bytes > 0) {
        constexpr int NCH = TILE * R::bytes / 16;  // TILE % 32 == 0 makes this exact
        if (n_valid == TILE) {
            #pragma unroll
            for (int i = 0; i < (NCH + THREADS - 1) / THREADS; ++i) {
                int ch = tid + i * THREADS;
                if ((NCH % THREADS == 0) || ch < NCH) cp_async16(stage + R::tile_chunk_off(ch), gsrc + (size_t)ch * 16);
            }
        } else {
            const int vbytes = n_valid * R::bytes;
            const int nch = (vbytes + 15) / 16;
            for (int ch = tid; ch < nch; ch += THREADS) {
                int rem = vbytes - ch * 16;
                cp_async16(stage + R::tile_chunk_off(ch), gsrc + (size_t)ch * 16, rem < 16 ? rem : 16);
            }
        }
    }
}

template <class R, int TILE, int THREADS>
__device__ __forceinline__ void tile_store(const unsigned char* stage, unsigned char* gdst, int n_valid, int tid) {
    if constexpr (R::bytes > 0) {
        constexpr int NCH = TILE * R::bytes / 16;
        if (n_valid == TILE) {
            #pragma unroll
            for (int i = 0; i < (NCH + THREADS - 1) / THREADS; ++i) {
                int ch = tid + i * THREADS;
                if ((NCH % THREADS == 0) || ch < NCH)
                    st_global_stream16(gdst + (size_t)ch * 16, *reinterpret_cast<const uint4*>(stage + R::tile_chunk_off(ch)));
            }
        } else {
            const int vbytes = n_valid * R::bytes;
            const int nfull = vbytes / 16;
            for (int ch = tid; ch < nfull; ch += THREADS)
                st_global_stream16(gdst + (size_t)ch * 16, *reinterpret_cast<const uint4*>(stage + R::tile_chunk_off(ch)));
            // ragged last chunk (only possible when the record size is not a multiple of 16: linear slots)
            for (int b = nfull * 16 + tid; b < vbytes; b += THREADS) gdst[b] = stage[b];
        }
    }
}

// result records of <= 16 B (and the status byte) go straight from registers to HBM
template <int BYTES>
__device__ __forceinline__ void stg_direct(unsigned char* g, const void* regs) {
    if constexpr (BYTES == 16) *reinterpret_cast<uint4*>(g) = *reinterpret_cast<const uint4*>(regs);
    else if constexpr (BYTES % 8 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 8; ++i) reinterpret_cast<uint2*>(g)[i] = reinterpret_cast<const uint2*>(regs)[i];
    } else if constexpr (BYTES % 4 == 0) {
        #pragma unroll
        for (int i = 0; i < BYTES / 4; ++i) reinterpret_cast<unsigned*>(g)[i] = reinterpret_cast<const unsigned*>(regs)[i];
    } else {
        #pragma unroll
        for (int i = 0; i < BYTES; ++i) g[i] = reinterpret_cast<const unsigned char*>(regs)[i];
    }
}

// ---- the kernel ------------------------------------------------------------------
// Op provides: In0, In1, Out0, Out1 (Rec types; NoRec when absent) and
//   __device__ static void run(const In0::type*, const In1::type*, Out0::type*, Out1::type*)
// operating on register arrays.
template <class Op, int THREADS, int IPT, int STAGES>
struct TileCfg {
    using I0 = typename Op::In0; using I1 = typename Op::In1;
    using O0 = typename Op::Out0; using O1 = typename Op::Out1;
    static constexpr int TILE = THREADS * IPT;
    // where Out0 goes: -1 direct from registers, 0 / 1 in place over In0 / In1, 2 its own stage
    static constexpr int OUT0 = O0::bytes <= 16 ? -1 : (O0::bytes == I0::bytes ? 0 : (O0::bytes == I1::bytes ? 1 : 2));
    static_assert(O1::bytes <= 16, "second output must be a small record (status byte)");
    static constexpr int IN_STAGE = TILE * (I0::stride + I1::stride);
    static constexpr int OUT_STAGE = OUT0 == 2 ? TILE * O0::stride : 0;
    static constexpr int SMEM = STAGES * IN_STAGE + OUT_STAGE;
};

template <class Op, int THREADS, int IPT, int STAGES, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) tile_kernel(TileArgs a) {
    using C = TileCfg<Op, THREADS, IPT, STAGES>;
    using I0 = typename C::I0; using I1 = typename C::I1;
    using O0 = typename C::O0; using O1 = typename C::O1;
    constexpr int TILE = C::TILE;
    static_assert(STAGES >= 2, "need a ring");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long tile_q[STAGES];

    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (a.n_items + TILE - 1) / TILE;
    const unsigned char* g_in0 = (const unsigned char*)a.in0;
    const unsigned char* g_in1 = (const unsigned char*)a.in1;
    unsigned char* g_out0 = (unsigned char*)a.out0;
    unsigned char* g_out1 = (unsigned char*)a.out1;

    auto issue = [&](unsigned long long tile, int stage) {
        if (tile < n_tiles) {
            unsigned long long first = tile * TILE;
            unsigned long long left = a.n_items - first;
            int nv = left < (unsigned long long)TILE ? (int)left : TILE;
            unsigned char* s0 = smem + stage * C::IN_STAGE;
            tile_load_async<I0, TILE, THREADS>(s0, g_in0 + first * I0::bytes, nv, tid);
            tile_load_async<I1, TILE, THREADS>(s0 + TILE * I0::stride, g_in1 + first * I1::bytes, nv, tid);
        }
        cp_async_commit();  // always commit: keeps the group count uniform
    };

    pdl_trigger();
    pdl_wait();   // nothing of the previous launch (operands it produced, scheduler words) is touched before this
    // prologue: claim and start the first STAGES-1 tiles; keep one more claim in flight
    unsigned long long next_claim = 0;
    TileClaimer claimer;
    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) tile_q[s] = claimer.next(a.sched, a.claim);
        next_claim = claimer.next(a.sched, a.claim);
    }
    __syncthreads();
    #pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(tile_q[s], s);

    int stage = 0;
    while (true) {
        const unsigned long long tile = tile_q[stage];
        if (tile >= n_tiles) break;  // claims are monotonic per CTA: nothing later is valid either
        int pf = stage + STAGES - 1; if (pf >= STAGES) pf -= STAGES;
        if (tid == 0) tile_q[pf] = next_claim;   // slot pf was consumed one iteration ago
        cp_async_wait<STAGES - 2>();
        __syncthreads();  // (A) this tile has landed for every thread; stage pf is free; tile_q[pf] visible
        issue(tile_q[pf], pf);
        if (tid == 0) next_claim = claimer.next(a.sched, a.claim);  // consumed next iteration: latency hidden

        const unsigned long long first = tile * TILE;
        const unsigned long long left = a.n_items - first;
        const int nv = left < (unsigned long long)TILE ? (int)left : TILE;
        unsigned char* s0 = smem + stage * C::IN_STAGE;
        unsigned char* s1 = s0 + TILE * I0::stride;
        unsigned char* so = C::OUT0 == 0 ? s0 : (C::OUT0 == 1 ? s1 : smem + STAGES * C::IN_STAGE);
        #pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * THREADS;
            if (it < nv) {
                RegArr<I0> r0; RegArr<I1> r1; RegArr<O0> w0; RegArr<O1> w1;
                if constexpr (I0::bytes > 0) I0::load(s0, it, r0.v);
                if constexpr (I1::bytes > 0 && !LazyIn1<Op>::value) I1::load(s1, it, r1.v);
                if constexpr (StreamsOut<Op>::value) {
                    static_assert(C::OUT0 == 0 || C::OUT0 == 1, "streamed results go in place over an input record");
                    SlotWriter<std::conditional_t<C::OUT0 == 0, I0, I1>> wr{so, it};
                    if constexpr (LazyIn1<Op>::value) {
                        SlotReader<I1> rd{s1, it};
                        Op::run_stream(r0.v, rd, wr, w1.v);
                    } else if constexpr (ScratchIn0<Op>::value) {
                        SlotScratch<I0> scr{s0, it};
                        Op::run_stream(r0.v, r1.v, wr, w1.v, scr);
                    } else {
                        Op::run_stream(r0.v, r1.v, wr, w1.v);
                    }
                } else {
                    run_op<Op>(r0.v, r1.v, w0.v, w1.v, a.prm);
                }
                if constexpr (O0::bytes > 0 && !StreamsOut<Op>::value) {
                    if constexpr (C::OUT0 == -1) stg_direct<O0::bytes>(g_out0 + (first + it) * O0::bytes, w0.v);
                    else if constexpr (C::OUT0 == 0) I0::store(so, it, w0.v);   // in place: same thread, same slot
                    else if constexpr (C::OUT0 == 1) I1::store(so, it, w0.v);
                    else O0::store(so, it, w0.v);
                }
                if constexpr (O1::bytes > 0) stg_direct<O1::bytes>(g_out1 + (first + it) * O1::bytes, w1.v);
            }
        }
        if constexpr (C::OUT0 >= 0) {
            __syncthreads();  // (B) every result record of the tile is in shared memory
            if constexpr (C::OUT0 == 0) tile_store<I0, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
            else if constexpr (C::OUT0 == 1) tile_store<I1, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
            else tile_store<O0, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
        }
        if (++stage == STAGES) stage = 0;
    }
    cp_async_wait<0>();
    // the last CTA to leave re-arms the scheduler for the next launch on this stream
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

// ---- one-stage variant for the largest records -------------------------------------------------
// An 8x8 Double record is 528 B of slot; a two-stage ring is 1 056 B per thread, three 64-thread CTAs per SM
// = six warps, too few for the FP64 dependency chains of det / solve / inverse (ncu: FP64 pipe ~21 % busy,
// `wait` the top stall).  With ONE stage per CTA the registers become the limit (four CTAs = eight warps)
// and the copy/compute overlap comes from the other resident CTAs instead of from a prefetched tile.
template <class Op, int THREADS, int IPT, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) tile_kernel_1s(TileArgs a) {
    using C = TileCfg<Op, THREADS, IPT, 1>;
    using I0 = typename C::I0; using I1 = typename C::I1;
    using O0 = typename C::O0; using O1 = typename C::O1;
    constexpr int TILE = C::TILE;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long tile_q[2];

    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (a.n_items + TILE - 1) / TILE;
    const unsigned char* g_in0 = (const unsigned char*)a.in0;
    const unsigned char* g_in1 = (const unsigned char*)a.in1;
    unsigned char* g_out0 = (unsigned char*)a.out0;
    unsigned char* g_out1 = (unsigned char*)a.out1;

    pdl_trigger();
    pdl_wait();
    TileClaimer claimer;
    if (tid == 0) tile_q[0] = claimer.next(a.sched, a.claim);
    __syncthreads();
    int q = 0;
    while (true) {
        const unsigned long long tile = tile_q[q];
        if (tile >= n_tiles) break;
        const unsigned long long first = tile * TILE;
        const unsigned long long left = a.n_items - first;
        const int nv = left < (unsigned long long)TILE ? (int)left : TILE;
        unsigned char* s0 = smem;
        unsigned char* s1 = s0 + TILE * I0::stride;
        unsigned char* so = C::OUT0 == 0 ? s0 : (C::OUT0 == 1 ? s1 : smem + C::IN_STAGE);
        tile_load_async<I0, TILE, THREADS>(s0, g_in0 + first * I0::bytes, nv, tid);
        tile_load_async<I1, TILE, THREADS>(s1, g_in1 + first * I1::bytes, nv, tid);
        cp_async_commit();
        if (tid == 0) tile_q[q ^ 1] = claimer.next(a.sched, a.claim);   // the next claim travels while this tile lands
        cp_async_wait<0>();
        __syncthreads();
        #pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * THREADS;
            if (it < nv) {
                RegArr<I0> r0; RegArr<I1> r1; RegArr<O0> w0; RegArr<O1> w1;
                if constexpr (I0::bytes > 0) I0::load(s0, it, r0.v);
                if constexpr (I1::bytes > 0 && !LazyIn1<Op>::value) I1::load(s1, it, r1.v);
                if constexpr (StreamsOut<Op>::value) {
                    static_assert(C::OUT0 == 0 || C::OUT0 == 1, "streamed results go in place over an input record");
                    SlotWriter<std::conditional_t<C::OUT0 == 0, I0, I1>> wr{so, it};
                    if constexpr (LazyIn1<Op>::value) {
                        SlotReader<I1> rd{s1, it};
                        Op::run_stream(r0.v, rd, wr, w1.v);
                    } else if constexpr (ScratchIn0<Op>::value) {
                        SlotScratch<I0> scr{s0, it};
                        Op::run_stream(r0.v, r1.v, wr, w1.v, scr);
                    } else {
                        Op::run_stream(r0.v, r1.v, wr, w1.v);
                    }
                } else {
                    run_op<Op>(r0.v, r1.v, w0.v, w1.v, a.prm);
                }
                if constexpr (O0::bytes > 0 && !StreamsOut<Op>::value) {
                    if constexpr (C::OUT0 == -1) stg_direct<O0::bytes>(g_out0 + (first + it) * O0::bytes, w0.v);
                    else if constexpr (C::OUT0 == 0) I0::store(so, it, w0.v);
                    else if constexpr (C::OUT0 == 1) I1::store(so, it, w0.v);
                    else O0::store(so, it, w0.v);
                }
                if constexpr (O1::bytes > 0) stg_direct<O1::bytes>(g_out1 + (first + it) * O1::bytes, w1.v);
            }
        }
        if constexpr (C::OUT0 >= 0) {
            __syncthreads();
            if constexpr (C::OUT0 == 0) tile_store<I0, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
            else if constexpr (C::OUT0 == 1) tile_store<I1, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
            else tile_store<O0, TILE, THREADS>(so, g_out0 + first * O0::bytes, nv, tid);
        }
        __syncthreads();      // the slot is free again; tile_q[q ^ 1] is visible
        q ^= 1;
    }
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

// ---- SoA skeleton: element e of item b at base[e * stride + b] ----------------------------
// Thread-per-item access is already coalesced in this layout (a warp reads 128/256 contiguous
// bytes per element), but a thread that loads its 9..32 elements, runs a division chain and
// stores leaves HBM idle while it computes.  So SoA batches go through the same machinery as AoS:
// a STAGES-deep cp.async ring fills shared [element][item-in-tile] rows (one contiguous
// TILE*sizeof(T) segment per element: conflict-free for stride-1 LDS with no swizzle at all), the
// dynamic tile scheduler hands out tiles, and results go straight from registers to their rows.
struct SoaArgs {
    const void* in0;
    const void* in1;
    void* out0;
    void* out1;
    unsigned long long n_items;
    unsigned long long stride;  // shard capacity in items (multiple of 32)
    unsigned long long* sched;  // tile scheduler words, as TileArgs
    unsigned claim;
    OpParams prm;
};

template <class R, int TILE, int THREADS>
__device__ __forceinline__ void soa_rows_load_async(unsigned char* sdst, const unsigned char* grow0, unsigned long long stride,
                                                    int n_valid, int tid) {
    if constexpr (R::n > 0 && R::bytes > 0) {
        using T = typename R::type;
        constexpr int ROWB = TILE * (int)sizeof(T), CPR = ROWB / 16, NCH = R::n * CPR, IPC = 16 / (int)sizeof(T);
        #pragma unroll
        for (int i = 0; i < (NCH + THREADS - 1) / THREADS; ++i) {
            const int ch = tid + i * THREADS;
            if ((NCH % THREADS == 0) || ch < NCH) {
                const int e = ch / CPR, c = ch - e * CPR;
                // rows are padded to 32 items, so a chunk that starts on a valid item is readable in full
                if (c * IPC < n_valid) cp_async16(sdst + e * ROWB + c * 16, grow0 + ((size_t)e * stride) * sizeof(T) + (size_t)c * 16);
            }
        }
    }
}

template <class Op, int THREADS, int IPT, int STAGES, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) soa_tile_kernel(SoaArgs a) {
    using I0 = typename Op::In0; using I1 = typename Op::In1;
    using O0 = typename Op::Out0; using O1 = typename Op::Out1;
    using T0 = typename I0::type; using T1 = typename I1::type;
    constexpr int TILE = THREADS * IPT;
    constexpr int IN_STAGE = TILE * (I0::bytes + I1::bytes);
    static_assert(STAGES >= 2, "need a ring");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long tile_q[STAGES];

    const int tid = threadIdx.x;
    const unsigned long long n_tiles = (a.n_items + TILE - 1) / TILE;
    const unsigned char* g_in0 = (const unsigned char*)a.in0;
    const unsigned char* g_in1 = (const unsigned char*)a.in1;
    typename O0::type* out0 = (typename O0::type*)a.out0;
    typename O1::type* out1 = (typename O1::type*)a.out1;

    auto issue = [&](unsigned long long tile, int stage) {
        if (tile < n_tiles) {
            const unsigned long long first = tile * TILE;
            const unsigned long long left = a.n_items - first;
            const int nv = left < (unsigned long long)TILE ? (int)left : TILE;
            unsigned char* s0 = smem + stage * IN_STAGE;
            soa_rows_load_async<I0, TILE, THREADS>(s0, g_in0 + first * sizeof(T0), a.stride, nv, tid);
            soa_rows_load_async<I1, TILE, THREADS>(s0 + TILE * I0::bytes, g_in1 + first * sizeof(T1), a.stride, nv, tid);
        }
        cp_async_commit();
    };

    pdl_trigger();
    pdl_wait();
    unsigned long long next_claim = 0;
    TileClaimer claimer;
    if (tid == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) tile_q[s] = claimer.next(a.sched, a.claim);
        next_claim = claimer.next(a.sched, a.claim);
    }
    __syncthreads();
    #pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(tile_q[s], s);

    int stage = 0;
    while (true) {
        const unsigned long long tile = tile_q[stage];
        if (tile >= n_tiles) break;
        int pf = stage + STAGES - 1; if (pf >= STAGES) pf -= STAGES;
        if (tid == 0) tile_q[pf] = next_claim;
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(tile_q[pf], pf);
        if (tid == 0) next_claim = claimer.next(a.sched, a.claim);

        const unsigned long long first = tile * TILE;
        const unsigned long long left = a.n_items - first;
        const int nv = left < (unsigned long long)TILE ? (int)left : TILE;
        const T0* s0 = reinterpret_cast<const T0*>(smem + stage * IN_STAGE);
        const T1* s1 = reinterpret_cast<const T1*>(smem + stage * IN_STAGE + TILE * I0::bytes);
        #pragma unroll
        for (int j = 0; j < IPT; ++j) {
            const int it = tid + j * THREADS;
            if (it < nv) {
                RegArr<I0> r0; RegArr<I1> r1; RegArr<O0> w0; RegArr<O1> w1;
                #pragma unroll
                for (int e = 0; e < I0::n; ++e) r0.v[e] = s0[e * TILE + it];
                if constexpr (!LazyIn1<Op>::value) {
                    #pragma unroll
                    for (int e = 0; e < I1::n; ++e) r1.v[e] = s1[e * TILE + it];
                }
                const unsigned long long b = first + it;
                if constexpr (StreamsOut<Op>::value) {
                    SoaWriter<typename O0::type> wr{out0 + b, a.stride};
                    if constexpr (LazyIn1<Op>::value) {
                        SoaReader<T1> rd{s1 + it, TILE};
                        Op::run_stream(r0.v, rd, wr, w1.v);
                    } else if constexpr (ScratchIn0<Op>::value) {
                        SoaScratch<T0, TILE> scr{const_cast<T0*>(s0) + it};
                        Op::run_stream(r0.v, r1.v, wr, w1.v, scr);
                    } else {
                        Op::run_stream(r0.v, r1.v, wr, w1.v);
                    }
                } else {
                    run_op<Op>(r0.v, r1.v, w0.v, w1.v, a.prm);
                    #pragma unroll
                    for (int e = 0; e < O0::n; ++e) out0[(size_t)e * a.stride + b] = w0.v[e];
                }
                #pragma unroll
                for (int e = 0; e < O1::n; ++e) out1[(size_t)e * a.stride + b] = w1.v[e];
            }
        }
        if (++stage == STAGES) stage = 0;
    }
    cp_async_wait<0>();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.sched + 1, 1ull) == (unsigned long long)gridDim.x - 1) {
            a.sched[0] = 0;
            a.sched[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace t5
