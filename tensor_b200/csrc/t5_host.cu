// C-ABI host layer (include/t5b200.h): contexts, sharded device buffers, the batch
// partitioner, per-operation validation + dispatch to the kernel TUs, the
// host-buffer (end-to-end) pipeline, the NCCL dot-sum exchange and the timing
// helpers.  No kernel lives here; no CPU arithmetic path exists: when no kernel
// covers a request the entry returns T5_ERR_UNSUPPORTED.
#include "../../include/t5b200.h"
#include "t5_internal.h"

#include <dlfcn.h>
#include <nccl.h>  // types/enums only; the functions are resolved with dlsym at run time

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

using namespace t5;

// -------------------------------------------------------------------------------------
// NCCL, loaded lazily so that the library has no link-time dependency on it.
// -------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
    }
};

struct HostSlot {  // staging of the host-buffer pipeline
    cudaStream_t stream = nullptr;
    void* d[4] = {nullptr, nullptr, nullptr, nullptr};   // <= 2 inputs + <= 2 outputs
    size_t cap[4] = {0, 0, 0, 0};
};

struct t5_ctx {
    std::mutex mu;
    int ndev = 0;
    int dev[T5_MAX_DEVICES];
    cudaStream_t stream[T5_MAX_DEVICES];
    cudaEvent_t ev0[T5_MAX_DEVICES], ev1[T5_MAX_DEVICES];
    // Stream-ordered allocator: every t5_buf comes from a per-device pool on the device's launch
    // stream (cudaMallocFromPoolAsync / cudaFreeAsync, release threshold = never).  "Outputs are
    // fresh buffers"  makes alloc + free part of EVERY operation: with
    // cudaMalloc / sync + cudaFree that was ~0.3 ms per call, several times the kernels it served.
    cudaMemPool_t pool[T5_MAX_DEVICES];
    void* scratch[T5_MAX_DEVICES];   // dot-sum block partials + the device partial (at the end)
    void* host_word[T5_MAX_DEVICES]; // 8 pinned bytes per device: the dot-sum partial comes back with an async copy on the
                                     // launch stream, so the call blocks once (stream sync) instead of twice (sync + cudaMemcpy)
    void* flush[T5_MAX_DEVICES];     // L2 flush target (lazy)
    void* coef[T5_MAX_DEVICES];      // polyEval coefficients for the any-shape kernel (lazy, grow-only)
    size_t coef_cap[T5_MAX_DEVICES];
    // Index tables of the table-driven gathers (transpose of rank >= 3, every structural op), one device copy per
    // device, keyed by (op, dims, params).  Bounded: beyond TABLE_CACHE_MAX entries the least recently used one is
    // released (a loop of `ta i` over the rows of a large tensor would otherwise keep one table per i for ever).
    struct Table {
        uint32_t* dev[T5_MAX_DEVICES];
        std::vector<uint32_t> host;        // kept until every device has its copy
        std::vector<uint64_t> out_dims;
        size_t d_out = 0;
        uint64_t last_use = 0;
    };
    std::map<std::vector<uint64_t>, Table> perm_cache;
    uint64_t table_clock = 0;
    std::mutex aux;                      // perm_cache + last_error: the host-array pipeline touches them from one thread per device
    HostSlot slots[T5_MAX_DEVICES][3];
    std::string last_error;
    std::atomic<uint64_t> launches{0};
    // Ownership: Haskell finalizers run t5_free at an arbitrary time, possibly AFTER t5_shutdown.
    // t5_shutdown releases every device resource, the memory of still-live buffers included, and marks the context
    // closed; the struct itself (mutex + registry) survives until the last t5_free of such an orphaned buffer.
    std::unordered_set<t5_buf*> bufs;
    int live_bufs = 0;
    bool closed = false;
    NcclApi nccl;
    ncclComm_t comms[T5_MAX_DEVICES];
    bool nccl_ready = false, nccl_failed = false, used_nccl = false;
};

struct t5_buf {
    t5_ctx* ctx;
    int dtype, layout;
    size_t elems, batch;
    void* ptr[T5_MAX_DEVICES];
    size_t first[T5_MAX_DEVICES], count[T5_MAX_DEVICES], cap[T5_MAX_DEVICES], bytes[T5_MAX_DEVICES];
};

static constexpr size_t PART_OFF = 148 * 8 * 8;          // dot-sum device partial
static constexpr size_t SCHED_OFF = PART_OFF + 64;        // 4 x {next tile, finished CTAs}: main stream + 3 pipeline slots
static constexpr size_t SCRATCH_BYTES = SCHED_OFF + 4 * 16;
static constexpr size_t FLUSH_BYTES = 256ull << 20;

static constexpr size_t TABLE_CACHE_MAX = 256;            // index tables kept per context (LRU beyond that)

static void set_error(t5_ctx* c, const std::string& msg) {
    if (!c) return;
    std::lock_guard<std::mutex> lk(c->aux);
    c->last_error = msg;
}

static int cuda_fail(t5_ctx* c, cudaError_t e, const char* where) {
    set_error(c, std::string(where) + ": " + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? T5_ERR_OOM : T5_ERR_CUDA;
}
#define CU(c, call)                                                     \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return cuda_fail((c), e__, #call);      \
    } while (0)

// Every entry that touches the device takes the context lock and refuses a context that t5_shutdown has closed
// (its struct is still alive only because orphaned buffers refer to it).
#define LOCKED(c)                                   \
    std::lock_guard<std::mutex> lk((c)->mu);        \
    if ((c)->closed) return T5_ERR_INVALID

// No C++ exception crosses the C boundary (include/t5b200.h): the std::vector / std::map / std::thread work of an
// entry runs inside a function-try-block that ends with this handler.
#define T5_NOEXCEPT_TAIL                                          \
    catch (const std::bad_alloc&) { return T5_ERR_OOM; }          \
    catch (...) { return T5_ERR_INVALID; }

static int map_internal(t5_ctx* c, int r, const char* what) {
    switch (r) {
        case T5I_OK: return T5_OK;
        case T5I_CUDA: {
            cudaError_t e = cudaGetLastError();
            set_error(c, std::string(what) + ": launch failed: " + cudaGetErrorString(e));
            return T5_ERR_CUDA;
        }
        default: return T5_ERR_UNSUPPORTED;
    }
}

#pragma GCC visibility push(default)
extern "C" {

const char* t5_strerror(int code) {
    switch (code) {
        case T5_OK: return "ok";
        case T5_ERR_SHAPE: return "shape or length mismatch (WrongListLength)";
        case T5_ERR_UNSUPPORTED: return "no device kernel for this dtype/shape/layout (there is no CPU fallback)";
        case T5_ERR_CUDA: return "CUDA error";
        case T5_ERR_NCCL: return "NCCL error";
        case T5_ERR_OOM: return "out of device memory";
        case T5_ERR_INVALID: return "invalid argument";
    }
    return "unknown error code";
}

const char* t5_version(void) { return "t5b200 sm_100a r2"; }

const char* t5_last_error(t5_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int t5_partition(size_t batch, int ndev, int g, size_t* first, size_t* count) {
    if (ndev < 1 || g < 0 || g >= ndev || !first || !count) return T5_ERR_INVALID;
    size_t per = (batch + (size_t)ndev - 1) / (size_t)ndev;
    per = (per + 255) / 256 * 256;
    size_t f = per * (size_t)g;
    if (f > batch) f = batch;
    size_t l = per * (size_t)(g + 1);
    if (l > batch) l = batch;
    *first = f;
    *count = l - f;
    return T5_OK;
}

static cudaError_t make_pool(int dev, cudaMemPool_t* pool) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaError_t e = cudaMemPoolCreate(pool, &props);
    if (e != cudaSuccess) return e;
    unsigned long long keep = ~0ull;                      // freed blocks stay in the pool until shutdown
    return cudaMemPoolSetAttribute(*pool, cudaMemPoolAttrReleaseThreshold, &keep);
}

int t5_init(const int* devices, int ndev, t5_ctx** out) try {
    if (!out) return T5_ERR_INVALID;
    *out = nullptr;
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible < 1) return T5_ERR_CUDA;
    std::vector<int> devs;
    if (!devices || ndev <= 0) {
        for (int i = 0; i < visible; ++i) devs.push_back(i);
    } else {
        for (int i = 0; i < ndev; ++i) {
            if (devices[i] < 0 || devices[i] >= visible) return T5_ERR_INVALID;
            devs.push_back(devices[i]);
        }
    }
    if ((int)devs.size() > T5_MAX_DEVICES) return T5_ERR_INVALID;
    t5_ctx* c = new (std::nothrow) t5_ctx();
    if (!c) return T5_ERR_OOM;
    c->ndev = (int)devs.size();
    for (int g = 0; g < c->ndev; ++g) {
        c->dev[g] = devs[g];
        c->pool[g] = nullptr;
        c->stream[g] = nullptr; c->ev0[g] = c->ev1[g] = nullptr; c->scratch[g] = c->flush[g] = c->coef[g] = nullptr; c->coef_cap[g] = 0; c->host_word[g] = nullptr;
    }
    for (int g = 0; g < c->ndev; ++g) {
        if ((e = cudaSetDevice(c->dev[g])) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&c->stream[g], cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreate(&c->ev0[g])) != cudaSuccess || (e = cudaEventCreate(&c->ev1[g])) != cudaSuccess ||
            (e = cudaMalloc(&c->scratch[g], SCRATCH_BYTES)) != cudaSuccess ||
            (e = cudaMallocHost(&c->host_word[g], 8)) != cudaSuccess ||
            (e = cudaMemset(c->scratch[g], 0, SCRATCH_BYTES)) != cudaSuccess ||
            (e = make_pool(c->dev[g], &c->pool[g])) != cudaSuccess) {
            int rc = cuda_fail(nullptr, e, "t5_init");
            t5_shutdown(c);
            return rc;
        }
    }
    *out = c;
    return T5_OK;
} T5_NOEXCEPT_TAIL

int t5_shutdown(t5_ctx* c) {
    if (!c) return T5_ERR_INVALID;
    bool destroy = false;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (c->closed) return T5_ERR_INVALID;      // already shut down; only orphaned buffers keep the struct alive
        c->closed = true;
        for (int g = 0; g < c->ndev; ++g) {
            cudaSetDevice(c->dev[g]);
            // buffers still owned by the caller (a ForeignPtr whose finalizer has not run yet): their device memory goes
            // back with the pool NOW, the handles become orphans that a later t5_free only deletes
            for (t5_buf* b : c->bufs)
                if (b->ptr[g] && c->stream[g]) { cudaFreeAsync(b->ptr[g], c->stream[g]); b->ptr[g] = nullptr; }
            if (c->stream[g]) cudaStreamSynchronize(c->stream[g]);
            for (auto& s : c->slots[g]) {
                if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); s.stream = nullptr; }
                for (auto& p : s.d) if (p) { cudaFree(p); p = nullptr; }
            }
            if (c->nccl_ready && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comms[g]);
            if (c->pool[g]) cudaMemPoolDestroy(c->pool[g]);     // returns every cached block to the device
            if (c->scratch[g]) cudaFree(c->scratch[g]);
            if (c->host_word[g]) cudaFreeHost(c->host_word[g]);
            if (c->flush[g]) cudaFree(c->flush[g]);
            if (c->coef[g]) cudaFree(c->coef[g]);
            if (c->ev0[g]) cudaEventDestroy(c->ev0[g]);
            if (c->ev1[g]) cudaEventDestroy(c->ev1[g]);
            if (c->stream[g]) cudaStreamDestroy(c->stream[g]);
            c->stream[g] = nullptr; c->pool[g] = nullptr; c->scratch[g] = c->flush[g] = c->coef[g] = c->host_word[g] = nullptr;
            c->ev0[g] = c->ev1[g] = nullptr;
        }
        c->nccl_ready = false;
        for (auto& kv : c->perm_cache)
            for (int g = 0; g < c->ndev; ++g)
                if (kv.second.dev[g]) { cudaSetDevice(c->dev[g]); cudaFree(kv.second.dev[g]); }
        c->perm_cache.clear();
        destroy = c->live_bufs == 0;
    }
    if (destroy) delete c;
    return T5_OK;
}

int t5_live_buffers(t5_ctx* c) {
    if (!c) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    return c->live_bufs;
}

int t5_ndev(t5_ctx* c) { return c ? c->ndev : 0; }

int t5_sync(t5_ctx* c) {
    if (!c) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        CU(c, cudaStreamSynchronize(c->stream[g]));
    }
    return T5_OK;
}

uint64_t t5_launch_count(t5_ctx* c) { return c ? c->launches.load() : 0; }

// ---- buffers --------------------------------------------------------------------------
int t5_alloc(t5_ctx* c, int dtype, size_t elems, size_t batch, int layout, t5_buf** out) try {
    if (!c || !out) return T5_ERR_INVALID;
    *out = nullptr;
    const int es = dtype_size(dtype);
    if (es == 0 || (layout != T5_AOS && layout != T5_SOA)) return T5_ERR_INVALID;
    if (elems == 0) return T5_ERR_SHAPE;
    t5_buf* b = new (std::nothrow) t5_buf();
    if (!b) return T5_ERR_OOM;
    b->ctx = c; b->dtype = dtype; b->layout = layout; b->elems = elems; b->batch = batch;
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->closed) { delete b; return T5_ERR_INVALID; }
    for (int g = 0; g < T5_MAX_DEVICES; ++g) { b->ptr[g] = nullptr; b->first[g] = b->count[g] = b->cap[g] = b->bytes[g] = 0; }
    try { c->bufs.insert(b); } catch (...) { delete b; return T5_ERR_OOM; }     // registered before anything is allocated
    for (int g = 0; g < c->ndev; ++g) {
        t5_partition(batch, c->ndev, g, &b->first[g], &b->count[g]);
        b->cap[g] = (b->count[g] + 31) / 32 * 32;
        size_t items = layout == T5_SOA ? b->cap[g] : b->count[g];
        size_t bytes = (items * elems * (size_t)es + 255) / 256 * 256;
        if (bytes == 0) bytes = 256;
        b->bytes[g] = bytes;
        cudaError_t e = cudaSetDevice(c->dev[g]);
        if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&b->ptr[g], bytes, c->pool[g], c->stream[g]);
        // SoA rows are padded to 32 items; the pad must read as zero (dot_sum folds whole rows, whole-row copies
        // propagate it).  Only the pad is cleared - clearing the whole buffer cost a full extra write pass on every
        // SoA result (4Mi 2x3x4 Double transpose: 0.39 -> 0.27 ms)
        if (e == cudaSuccess && layout == T5_SOA && b->cap[g] != b->count[g])
            e = cudaMemset2DAsync((char*)b->ptr[g] + b->count[g] * (size_t)es, b->cap[g] * (size_t)es, 0,
                                  (b->cap[g] - b->count[g]) * (size_t)es, elems, c->stream[g]);
        if (e != cudaSuccess) {
            int rc = cuda_fail(c, e, "t5_alloc");
            for (int h = 0; h <= g; ++h)
                if (b->ptr[h]) { cudaSetDevice(c->dev[h]); cudaFreeAsync(b->ptr[h], c->stream[h]); }
            c->bufs.erase(b);
            delete b;
            return rc;
        }
    }
    c->live_bufs++;
    *out = b;
    return T5_OK;
} T5_NOEXCEPT_TAIL

int t5_free(t5_buf* b) {
    if (!b) return T5_ERR_INVALID;
    t5_ctx* c = b->ctx;
    bool destroy = false;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        if (!c->closed)
            for (int g = 0; g < c->ndev; ++g) {
                if (!b->ptr[g]) continue;
                // work reading/writing the buffer may still be in flight: the release is ordered after it on
                // the launch stream and never blocks the caller (a GHC finalizer thread)
                cudaSetDevice(c->dev[g]);
                cudaFreeAsync(b->ptr[g], c->stream[g]);
            }
        c->bufs.erase(b);
        c->live_bufs--;
        // after t5_shutdown the buffer is an orphan (its memory went back with the pool): only the handle is
        // deleted, and the LAST orphan takes the context struct with it
        destroy = c->closed && c->live_bufs == 0;
    }
    delete b;
    if (destroy) delete c;
    return T5_OK;
}

size_t t5_buf_batch(const t5_buf* b) { return b ? b->batch : 0; }
size_t t5_buf_elems(const t5_buf* b) { return b ? b->elems : 0; }
int t5_buf_dtype(const t5_buf* b) { return b ? b->dtype : -1; }
int t5_buf_layout(const t5_buf* b) { return b ? b->layout : -1; }

void* t5_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void t5_pinned_free(void* p) { if (p) cudaFreeHost(p); }

static int xfer(t5_buf* b, void* host, size_t bytes, bool up) {
    if (!b) return T5_ERR_INVALID;
    t5_ctx* c = b->ctx;
    const size_t es = dtype_size(b->dtype);
    const size_t item = b->elems * es;
    if (bytes != b->batch * item) return T5_ERR_SHAPE;
    if (b->batch && !host) return T5_ERR_INVALID;
    LOCKED(c);
    std::vector<void*> tmp(c->ndev, nullptr);
    int rc = T5_OK;
    for (int g = 0; g < c->ndev && rc == T5_OK; ++g) {
        if (!b->count[g]) continue;
        cudaError_t e = cudaSetDevice(c->dev[g]);
        char* h = (char*)host + b->first[g] * item;
        const size_t n = b->count[g] * item;
        if (e == cudaSuccess && b->layout == T5_AOS) {
            e = up ? cudaMemcpyAsync(b->ptr[g], h, n, cudaMemcpyHostToDevice, c->stream[g])
                   : cudaMemcpyAsync(h, b->ptr[g], n, cudaMemcpyDeviceToHost, c->stream[g]);
        } else if (e == cudaSuccess) {
            e = cudaMallocFromPoolAsync(&tmp[g], n, c->pool[g], c->stream[g]);
            if (e == cudaSuccess && up) {
                e = cudaMemcpyAsync(tmp[g], h, n, cudaMemcpyHostToDevice, c->stream[g]);
                if (e == cudaSuccess && relayout(c->stream[g], (int)es, b->elems, b->count[g], b->cap[g], true, tmp[g], b->ptr[g]) != T5I_OK) e = cudaGetLastError();
                c->launches++;
            } else if (e == cudaSuccess) {
                if (relayout(c->stream[g], (int)es, b->elems, b->count[g], b->cap[g], false, b->ptr[g], tmp[g]) != T5I_OK) e = cudaGetLastError();
                c->launches++;
                if (e == cudaSuccess) e = cudaMemcpyAsync(h, tmp[g], n, cudaMemcpyDeviceToHost, c->stream[g]);
            }
        }
        if (e != cudaSuccess) rc = cuda_fail(c, e, up ? "t5_upload" : "t5_download");
    }
    for (int g = 0; g < c->ndev; ++g) {
        cudaSetDevice(c->dev[g]);
        cudaError_t e = cudaStreamSynchronize(c->stream[g]);
        if (e != cudaSuccess && rc == T5_OK) rc = cuda_fail(c, e, "sync after transfer");
        if (tmp[g]) cudaFreeAsync(tmp[g], c->stream[g]);
    }
    return rc;
}

int t5_upload(t5_buf* b, const void* host, size_t bytes) try { return xfer(b, const_cast<void*>(host), bytes, true); } T5_NOEXCEPT_TAIL
int t5_download(t5_buf* b, void* host, size_t bytes) try { return xfer(b, host, bytes, false); } T5_NOEXCEPT_TAIL

int t5_relayout(t5_buf* dst, const t5_buf* src) {
    if (!dst || !src || dst->ctx != src->ctx || dst == src) return T5_ERR_INVALID;
    if (dst->dtype != src->dtype || dst->elems != src->elems || dst->batch != src->batch) return T5_ERR_SHAPE;
    t5_ctx* c = dst->ctx;
    LOCKED(c);
    const int es = dtype_size(dst->dtype);
    for (int g = 0; g < c->ndev; ++g) {
        if (!dst->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        if (dst->layout == src->layout) {
            CU(c, cudaMemcpyAsync(dst->ptr[g], src->ptr[g], src->bytes[g], cudaMemcpyDeviceToDevice, c->stream[g]));
        } else {
            const bool to_soa = dst->layout == T5_SOA;
            int r = relayout(c->stream[g], es, dst->elems, dst->count[g], to_soa ? dst->cap[g] : src->cap[g], to_soa, src->ptr[g], dst->ptr[g]);
            c->launches++;
            if (r != T5I_OK) return map_internal(c, r, "t5_relayout");
        }
    }
    return T5_OK;
}

// ---- operand validation ------------------------------------------------------------------
static int check(const t5_ctx* c, const t5_buf* b, int dtype, size_t elems, const t5_buf* like) {
    if (!b || b->ctx != c) return T5_ERR_INVALID;
    if (b->dtype != dtype) return T5_ERR_INVALID;
    if (b->elems != elems) return T5_ERR_SHAPE;
    if (like && (b->batch != like->batch)) return T5_ERR_SHAPE;
    if (like && (b->layout != like->layout)) return T5_ERR_INVALID;
    return T5_OK;
}
#define CHECK(expr) do { int rc__ = (expr); if (rc__ != T5_OK) return rc__; } while (0)

static Shard shard_of(const t5_ctx* c, const t5_buf* b, int g) {
    Shard s;
    s.stream = c->stream[g]; s.layout = b->layout; s.count = b->count[g]; s.soa_stride = b->cap[g]; s.first = b->first[g];
    s.sched = (unsigned long long*)((char*)c->scratch[g] + SCHED_OFF);
    return s;
}

// one shard of a (P x K).(K x Q) product; shared by t5_matmul / t5_dot / t5_prod and
// the host-buffer pipeline
static int product_shard(const Shard& s, int dtype, int P, int K, int Q, int nc, const void* A, const void* B, void* C) {
    int r = T5I_NOT_SPECIALISED;
    // the tensor-core kernels first: their shape rules (dmma_takes / tf32x3_takes - exactly what t5_matmul_is_exact reports)
    // exclude everything the exact thread-per-item kernels are good at
    if (s.layout == T5L_AOS && dtype == T5D_F64 && nc != 0) r = dmma_matmul(s, P, K, Q, A, B, C);
    if (s.layout == T5L_AOS && dtype == T5D_F32 && nc != 0) r = tf32x3_matmul(s, P, K, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc != 0) r = small_matmul(s, dtype, P, K, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc == 0) r = small_outer(s, dtype, P, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc != 0) r = mid_matmul(s, dtype, P, K, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc != 0 && P == 1 && Q == 1 && K > 4) r = generic_longdot(s, dtype, K, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc != 0 && s.layout == T5L_AOS) r = generic_stagedprod(s, dtype, P, K, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED && nc != 0) r = exact_blocked_matmul(s, dtype, P, K, Q, A, B, C);
    if (r == T5I_NOT_SPECIALISED) r = generic_prod(s, dtype, P, K, Q, nc, A, B, C);
    return r;
}

// ---- AoS detour for SoA operands ------------------------------------------------------------------
// SoA is the tiny-record layout: the tile pipeline (records up to 8x8), outer products and the staged contractions
// read it directly.  Every other kernel - row-per-thread / staged / blocked / tensor-core products, eliminations
// beyond n = 8, the any-shape extras - takes AoS.  For those shapes a SoA shard is re-laid out into stream-ordered
// AoS temporaries, the AoS kernel runs, and the results are re-laid out back: one extra read + write pass per
// operand (t5_relayout's kernels), still several times faster than a thread-per-item contraction of a large record.
// fn(shard, in_ptrs, out_ptrs) -> T5I_* runs the op's kernel chain for one shard.
extern "C++" {
template <class F>
static int run_shard(t5_ctx* c, int g, const t5_buf* const* ins, int nin, t5_buf* const* outs, int nout, bool force_detour,
                     const char* what, F fn) {
    const void* ip[4] = {nullptr, nullptr, nullptr, nullptr};
    void* op[4] = {nullptr, nullptr, nullptr, nullptr};
    const t5_buf* like = ins[0];
    Shard s = shard_of(c, like, g);
    int r = T5I_NOT_SPECIALISED;
    if (!(force_detour && s.layout == T5L_SOA)) {
        for (int i = 0; i < nin; ++i) ip[i] = ins[i]->ptr[g];
        for (int i = 0; i < nout; ++i) op[i] = outs[i]->ptr[g];
        r = fn(s, ip, op);
    }
    if ((r == T5I_NOT_SPECIALISED || r == T5I_UNSUPPORTED) && s.layout == T5L_SOA) {
        void* tmp[8]; int nt = 0;
        cudaError_t e = cudaSuccess;
        auto stage = [&](const t5_buf* b) -> void* {
            if (b->elems == 1) return b->ptr[g];              // one element per item: SoA and AoS coincide
            void* p = nullptr;
            const size_t bytes = like->count[g] * b->elems * (size_t)dtype_size(b->dtype);
            if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&p, bytes ? bytes : 256, c->pool[g], s.stream);
            if (e == cudaSuccess) tmp[nt++] = p;
            return p;
        };
        int rr = T5I_OK;
        for (int i = 0; i < nin && e == cudaSuccess; ++i) {
            void* p = stage(ins[i]);
            ip[i] = p;
            if (e == cudaSuccess && p != ins[i]->ptr[g]) {
                rr = relayout(s.stream, dtype_size(ins[i]->dtype), ins[i]->elems, s.count, ins[i]->cap[g], false, ins[i]->ptr[g], p);
                c->launches++;
                if (rr != T5I_OK) break;
            }
        }
        for (int i = 0; i < nout && e == cudaSuccess; ++i) op[i] = stage(outs[i]);
        if (e == cudaSuccess && rr == T5I_OK) {
            Shard sa = s; sa.layout = T5L_AOS; sa.soa_stride = 0;
            r = fn(sa, ip, op);
            for (int i = 0; i < nout && r == T5I_OK; ++i)
                if (op[i] != outs[i]->ptr[g]) {
                    r = relayout(s.stream, dtype_size(outs[i]->dtype), outs[i]->elems, s.count, outs[i]->cap[g], true, op[i], outs[i]->ptr[g]);
                    c->launches++;
                }
        } else if (rr != T5I_OK) r = rr;
        for (int i = 0; i < nt; ++i) cudaFreeAsync(tmp[i], s.stream);
        if (e != cudaSuccess) return cuda_fail(c, e, what);
    }
    if (r != T5I_OK) return map_internal(c, r, what);
    c->launches++;
    return T5_OK;
}
}  // extern "C++"

// SoA products that would otherwise fall to the thread-per-item contraction with B too large to stage in shared
// memory (t5_generic.cu generic_soa_prod): run the AoS kernels through the detour instead
static bool soa_product_detours(int dtype, int P, int K, int Q, int nc) {
    if (nc == 0) return (long long)Q * 64 * 16 > 200 * 1024;
    // tensor-core shapes behave the same in both layouts (t5_matmul_is_exact depends on dtype and shape only)
    if ((dtype == T5_F64 && dmma_takes(P, K, Q)) || (dtype == T5_F32 && tf32x3_takes(P, K, Q))) return true;
    if (P == 1 && Q == 1) return false;                              // long dot: its own SoA-capable kernel
    const long long es = dtype == T5_F32 ? 4 : 8;
    return (long long)K * Q * 128 * es > 200 * 1024;
}

static int product_all(t5_ctx* c, int dtype, int P, int K, int Q, int nc, const t5_buf* A, const t5_buf* B, t5_buf* C, const char* what) {
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (P < 1 || K < 1 || Q < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)P * K, nullptr));
    CHECK(check(c, B, dtype, (size_t)K * Q, A));
    CHECK(check(c, C, dtype, (size_t)P * Q, A));
    if (C == A || C == B) return T5_ERR_INVALID;
    LOCKED(c);
    const t5_buf* ins[2] = {A, B};
    t5_buf* outs[1] = {C};
    const bool detour = A->layout == T5_SOA && soa_product_detours(dtype, P, K, Q, nc);
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        CHECK(run_shard(c, g, ins, 2, outs, 1, detour, what, [&](const Shard& s, const void* const* ip, void* const* op) {
            return product_shard(s, dtype, P, K, Q, nc, ip[0], ip[1], op[0]);
        }));
    }
    return T5_OK;
}

int t5_matmul(t5_ctx* c, int dtype, int m, int k, int n, const t5_buf* A, const t5_buf* B, t5_buf* C) {
    if (!c) return T5_ERR_INVALID;
    return product_all(c, dtype, m, k, n, 1, A, B, C, "t5_matmul");
}

int t5_dot(t5_ctx* c, int dtype, int n, const t5_buf* X, const t5_buf* Y, t5_buf* OUT) {
    if (!c) return T5_ERR_INVALID;
    return product_all(c, dtype, 1, n, 1, 1, X, Y, OUT, "t5_dot");
}

// (rankA, dimsA) x (rankB, dimsB) contracted over nc indices -> the P x K . K x Q view
static int prod_dims(int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB, int nc, int* Pp, int* Kp, int* Qp) {
    if (rankA < 0 || rankB < 0 || rankA > 16 || rankB > 16 || (rankA && !dimsA) || (rankB && !dimsB)) return T5_ERR_INVALID;
    if (nc < 0 || nc > rankA || nc > rankB) return T5_ERR_SHAPE;
    uint64_t P = 1, K = 1, Q = 1;
    for (int a = 0; a < rankA - nc; ++a) { P *= dimsA[a]; if (P > (1u << 28)) return T5_ERR_SHAPE; }
    for (int a = 0; a < nc; ++a) {
        if (dimsA[rankA - nc + a] != dimsB[a]) return T5_ERR_SHAPE;
        K *= dimsB[a];
        if (K > (1u << 28)) return T5_ERR_SHAPE;
    }
    for (int a = nc; a < rankB; ++a) { Q *= dimsB[a]; if (Q > (1u << 28)) return T5_ERR_SHAPE; }
    if (P == 0 || K == 0 || Q == 0) return T5_ERR_SHAPE;
    *Pp = (int)P; *Kp = (int)K; *Qp = (int)Q;
    return T5_OK;
}

int t5_prod(t5_ctx* c, int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB, int nc,
            const t5_buf* A, const t5_buf* B, t5_buf* C) {
    if (!c) return T5_ERR_INVALID;
    int P, K, Q;
    CHECK(prod_dims(rankA, dimsA, rankB, dimsB, nc, &P, &K, &Q));
    return product_all(c, dtype, P, K, Q, nc, A, B, C, "t5_prod");
}

// ---- dot_sum: the one entry with a cross-device exchange ---------------------------------
static bool ensure_nccl(t5_ctx* c) {
    if (c->nccl_ready) return true;
    if (c->nccl_failed) return false;
    if (!c->nccl.load() || c->nccl.CommInitAll(c->comms, c->ndev, c->dev) != ncclSuccess) {
        c->nccl_failed = true;
        return false;
    }
    c->nccl_ready = true;
    return true;
}

int t5_dot_sum_used_nccl(t5_ctx* c) { return c && c->used_nccl ? 1 : 0; }

int t5_matmul_is_exact(int dtype, int m, int k, int n) {
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return -T5_ERR_UNSUPPORTED;
    if (m < 1 || k < 1 || n < 1) return -T5_ERR_SHAPE;
    if (dtype == T5_F64) return dmma_takes(m, k, n) ? 0 : 1;
    if (dtype == T5_F32) return tf32x3_takes(m, k, n) ? 0 : 1;
    return 1;
}

int t5_dot_sum(t5_ctx* c, int dtype, int n, const t5_buf* X, const t5_buf* Y, void* host_scalar) {
    if (!c || !host_scalar) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, X, dtype, (size_t)n, nullptr));
    CHECK(check(c, Y, dtype, (size_t)n, X));
    LOCKED(c);
    const size_t part_off = PART_OFF;
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        Shard s = shard_of(c, X, g);
        const bool soa = X->layout == T5_SOA;
        const size_t elems = (soa ? X->cap[g] : X->count[g]) * (size_t)n;
        int r = dot_sum_partial(s, dtype, elems, soa ? X->cap[g] : 0, X->count[g], X->ptr[g], Y->ptr[g], c->scratch[g], (char*)c->scratch[g] + part_off);
        if (r != T5I_OK) return map_internal(c, r, "t5_dot_sum");
        c->launches += 2;
    }
    c->used_nccl = false;
    double dsum = 0.0;
    uint64_t isum = 0;
    if (c->ndev > 1 && ensure_nccl(c)) {
        // 1-element all-reduce over NVLink; every device ends with the total, device 0's copy is returned
        ncclDataType_t ty = dtype == T5_I64 ? ncclUint64 : ncclFloat64;
        if (c->nccl.GroupStart() != ncclSuccess) return T5_ERR_NCCL;
        for (int g = 0; g < c->ndev; ++g) {
            void* p = (char*)c->scratch[g] + part_off;
            ncclResult_t nr = c->nccl.AllReduce(p, p, 1, ty, ncclSum, c->comms[g], c->stream[g]);
            if (nr != ncclSuccess) {
                c->last_error = std::string("ncclAllReduce: ") + (c->nccl.GetErrorString ? c->nccl.GetErrorString(nr) : "?");
                c->nccl.GroupEnd();
                return T5_ERR_NCCL;
            }
        }
        if (c->nccl.GroupEnd() != ncclSuccess) return T5_ERR_NCCL;
        // every device ends with the total; device 0's copy returns through its pinned word on the launch stream:
        // ONE blocking point per call (it used to be a sync of every stream, then a synchronous 8-byte cudaMemcpy)
        CU(c, cudaSetDevice(c->dev[0]));
        CU(c, cudaMemcpyAsync(c->host_word[0], (char*)c->scratch[0] + part_off, 8, cudaMemcpyDeviceToHost, c->stream[0]));
        CU(c, cudaStreamSynchronize(c->stream[0]));
        if (dtype == T5_I64) isum = *(const volatile uint64_t*)c->host_word[0];
        else dsum = *(const volatile double*)c->host_word[0];
        c->used_nccl = true;
    } else {
        // single device, or NCCL not loadable: partials added on the host in device order
        for (int g = 0; g < c->ndev; ++g) {
            CU(c, cudaSetDevice(c->dev[g]));
            CU(c, cudaMemcpyAsync(c->host_word[g], (char*)c->scratch[g] + part_off, 8, cudaMemcpyDeviceToHost, c->stream[g]));
        }
        for (int g = 0; g < c->ndev; ++g) {
            CU(c, cudaSetDevice(c->dev[g]));
            CU(c, cudaStreamSynchronize(c->stream[g]));
            if (dtype == T5_I64) isum += *(const volatile uint64_t*)c->host_word[g];
            else dsum += *(const volatile double*)c->host_word[g];
        }
    }
    if (dtype == T5_F64) *(double*)host_scalar = dsum;
    else if (dtype == T5_F32) *(float*)host_scalar = (float)dsum;
    else *(int64_t*)host_scalar = (int64_t)isum;
    return T5_OK;
}

// ---- transpose -----------------------------------------------------------------------------
// transposeV ds = linearize ds . reverse . unlinearize (reverse ds)   (Vector.hs:306-307),
// written with 0-based indices: out linear index i over the REVERSED shape maps to the
// input offset of the reversed multi-index.
static void build_perm(const std::vector<uint64_t>& dims, std::vector<uint32_t>& perm) {
    const int r = (int)dims.size();
    size_t d = 1;
    for (auto x : dims) d *= x;
    perm.resize(d);
    std::vector<uint64_t> in_stride(r, 1);
    for (int a = r - 2; a >= 0; --a) in_stride[a] = in_stride[a + 1] * dims[a + 1];
    for (size_t i = 0; i < d; ++i) {
        // unlinearize i over reversed dims (dims[r-1], ..., dims[0]); last (fastest) out axis is in-axis 0
        size_t rem = i, off = 0;
        for (int a = 0; a < r; ++a) {           // a = input axis, which is output axis r-1-a (fastest first)
            size_t idx = rem % dims[a];
            rem /= dims[a];
            off += idx * in_stride[a];
        }
        perm[i] = (uint32_t)off;
    }
}

// ---- index-table cache -------------------------------------------------------------------------
// Looks the table of `key` up, building it with build(table, out_dims) -> T5_* on a miss, and makes sure device
// slot g (g < 0: none, sizes only) holds a copy.  Tables are built ONLY on a miss.  Serialised by c->aux: the
// host-array pipeline reaches it from one thread per device.
extern "C++" {
template <class Build>
static int table_get(t5_ctx* c, int g, const std::vector<uint64_t>& key, Build build, const uint32_t** dev, size_t* d_out,
                     std::vector<uint64_t>* out_dims) {
    std::lock_guard<std::mutex> lk(c->aux);
    auto it = c->perm_cache.find(key);
    if (it == c->perm_cache.end()) {
        t5_ctx::Table t;
        for (auto& p : t.dev) p = nullptr;
        int rc = build(t.host, t.out_dims);
        if (rc != T5_OK) return rc;
        t.d_out = t.host.size();
        if (c->perm_cache.size() >= TABLE_CACHE_MAX) {          // release the least recently used table
            auto victim = c->perm_cache.end();
            for (auto jt = c->perm_cache.begin(); jt != c->perm_cache.end(); ++jt)
                if (jt->second.last_use + 64 < c->table_clock && (victim == c->perm_cache.end() || jt->second.last_use < victim->second.last_use))
                    victim = jt;
            if (victim != c->perm_cache.end()) {
                for (int h = 0; h < c->ndev; ++h)
                    if (victim->second.dev[h]) {                  // launches that read it may still be in flight on any stream
                        cudaSetDevice(c->dev[h]);
                        cudaDeviceSynchronize();
                        cudaFree(victim->second.dev[h]);
                    }
                c->perm_cache.erase(victim);
            }
        }
        it = c->perm_cache.emplace(key, std::move(t)).first;
    }
    t5_ctx::Table& t = it->second;
    t.last_use = ++c->table_clock;
    if (g >= 0 && !t.dev[g]) {
        if (t.host.size() != t.d_out) {                            // host copy already dropped, a new device asks: rebuild
            std::vector<uint64_t> od;
            int rc = build(t.host, od);
            if (rc != T5_OK) return rc;
        }
        uint32_t* p = nullptr;
        cudaError_t e = cudaSetDevice(c->dev[g]);
        if (e == cudaSuccess) e = cudaMalloc(&p, (t.d_out ? t.d_out : 1) * 4);
        if (e == cudaSuccess && t.d_out) e = cudaMemcpy(p, t.host.data(), t.d_out * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (p) cudaFree(p);
            c->last_error = std::string("index table upload: ") + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? T5_ERR_OOM : T5_ERR_CUDA;
        }
        t.dev[g] = p;
        bool everywhere = true;
        for (int h = 0; h < c->ndev; ++h) everywhere = everywhere && t.dev[h];
        if (everywhere) std::vector<uint32_t>().swap(t.host);
    }
    if (g >= 0) cudaSetDevice(c->dev[g]);
    if (dev) *dev = g >= 0 ? t.dev[g] : nullptr;
    if (d_out) *d_out = t.d_out;
    if (out_dims) *out_dims = t.out_dims;
    return T5_OK;
}
}  // extern "C++"

// One shard of a transpose.  sq = dims with the extent-1 axes squeezed out, d = elements per item.
// Caller holds the context lock and has set the device.
static int transpose_shard(t5_ctx* c, int g, const Shard& s, int es, const std::vector<uint64_t>& sq, size_t d,
                           const void* in, void* out, size_t bytes) {
    if (sq.size() <= 1) {  // rank 0/1 (after squeezing): the identity
        CU(c, cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s.stream));
        return T5_OK;
    }
    int r = T5I_NOT_SPECIALISED;
    if (sq.size() == 2) r = small_transpose(s, es, (int)sq[0], (int)sq[1], in, out);
    if (r == T5I_NOT_SPECIALISED && sq.size() == 2) r = generic_transpose2d(s, es, (int)sq[0], (int)sq[1], in, out);
    if (r == T5I_NOT_SPECIALISED) {
        const uint32_t* tab = nullptr;
        CHECK(table_get(c, g, sq, [&](std::vector<uint32_t>& perm, std::vector<uint64_t>&) { build_perm(sq, perm); return (int)T5_OK; },
                        &tab, nullptr, nullptr));
        r = s.layout == T5L_SOA ? generic_soa_gather(s, es, d, d, tab, in, nullptr, out)   // SoA: a copy of whole rows
                                : generic_permute(s, es, d, tab, in, out);
    }
    if (r != T5I_OK) return map_internal(c, r, "t5_transpose");
    c->launches++;
    return T5_OK;
}

static int transpose_args(int dtype, int rank, const uint64_t* dims, std::vector<uint64_t>& sq, size_t& d) {
    if (rank < 0 || rank > 16 || (rank && !dims)) return T5_ERR_INVALID;
    const int es = dtype_size(dtype);
    if (es != 4 && es != 8) return T5_ERR_UNSUPPORTED;
    d = 1;
    sq.clear();
    for (int a = 0; a < rank; ++a) {
        const uint64_t x = dims[a];
        if (x == 0 || x > (1u << 28)) return T5_ERR_SHAPE;
        d *= x;
        if (d > (1u << 28)) return T5_ERR_SHAPE;
        if (x != 1) sq.push_back(x);   // axes of extent 1 do not move data
    }
    return T5_OK;
}

int t5_transpose(t5_ctx* c, int dtype, int rank, const uint64_t* dims, const t5_buf* A, t5_buf* OUT) try {
    if (!c) return T5_ERR_INVALID;
    std::vector<uint64_t> sq;
    size_t d = 1;
    CHECK(transpose_args(dtype, rank, dims, sq, d));
    const int es = dtype_size(dtype);
    CHECK(check(c, A, dtype, d, nullptr));
    CHECK(check(c, OUT, dtype, d, A));
    if (OUT == A) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        CHECK(transpose_shard(c, g, shard_of(c, A, g), es, sq, d, A->ptr[g], OUT->ptr[g], A->bytes[g]));
    }
    return T5_OK;
} T5_NOEXCEPT_TAIL

// ---- det / inverse / solve -----------------------------------------------------------------
// Kernel chains, fastest applicable first; shared by the device-buffer entries and the host-array pipeline.
static int det_chain(const Shard& s, int dtype, int n, const void* A, void* O) {
    int r = small_det(s, dtype, n, A, O);
    if (r == T5I_NOT_SPECIALISED) r = reg9_det(s, dtype, n, A, O);
    if (r == T5I_NOT_SPECIALISED) r = rowlane_det(s, dtype, n, A, O);
    if (r == T5I_NOT_SPECIALISED) r = warp_det(s, dtype, n, A, O);
    if (r == T5I_NOT_SPECIALISED) r = lu_det(s, dtype, n, A, O);
    if (r == T5I_NOT_SPECIALISED) r = generic_det(s, dtype, n, A, O);
    return r;
}
static int inverse_chain(const Shard& s, int dtype, int n, const void* A, void* O, void* st) {
    int r = small_inverse(s, dtype, n, A, O, st);
    if (r == T5I_NOT_SPECIALISED) r = reg9_inverse(s, dtype, n, A, O, st);
    if (r == T5I_NOT_SPECIALISED) r = rowlane_inverse(s, dtype, n, A, O, st);
    if (r == T5I_NOT_SPECIALISED) r = warp_inverse(s, dtype, n, A, O, st);
    if (r == T5I_NOT_SPECIALISED) r = lu_inverse(s, dtype, n, A, O, st);
    if (r == T5I_NOT_SPECIALISED) r = generic_solve(s, dtype, n, n, true, A, nullptr, O, st);
    return r;
}
static int solve_chain(const Shard& s, int dtype, int n, int nrhs, const void* A, const void* B, void* X, void* st) {
    int r = small_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = reg9_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = many_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = rowlane_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = warp_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = lu_solve(s, dtype, n, nrhs, A, B, X, st);
    if (r == T5I_NOT_SPECIALISED) r = generic_solve(s, dtype, n, nrhs, false, A, B, X, st);
    return r;
}

// one entry = validate, lock, then one run_shard per device
#define FOR_EACH_SHARD(c, like, ...)                                  \
    for (int g = 0; g < (c)->ndev; ++g) {                             \
        if (!(like)->count[g]) continue;                              \
        CU(c, cudaSetDevice((c)->dev[g]));                            \
        CHECK(run_shard(c, g, __VA_ARGS__));                          \
    }

int t5_det(t5_ctx* c, int dtype, int n, const t5_buf* A, t5_buf* OUT) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, 1, A));
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[1] = {OUT};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 1, false, "t5_det", [&](const Shard& s, const void* const* ip, void* const* op) {
        return det_chain(s, dtype, n, ip[0], op[0]);
    });
    return T5_OK;
}

int t5_inverse(t5_ctx* c, int dtype, int n, const t5_buf* A, t5_buf* OUT, t5_buf* STATUS) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, (size_t)n * n, A));
    CHECK(check(c, STATUS, T5_U8, 1, A));
    if (OUT == A) return T5_ERR_INVALID;
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[2] = {OUT, STATUS};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 2, false, "t5_inverse", [&](const Shard& s, const void* const* ip, void* const* op) {
        return inverse_chain(s, dtype, n, ip[0], op[0], op[1]);
    });
    return T5_OK;
}

int t5_solve(t5_ctx* c, int dtype, int n, int nrhs, const t5_buf* A, const t5_buf* B, t5_buf* X, t5_buf* STATUS) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1 || nrhs < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, B, dtype, (size_t)n * nrhs, A));
    CHECK(check(c, X, dtype, (size_t)n * nrhs, A));
    CHECK(check(c, STATUS, T5_U8, 1, A));
    if (X == A || X == B) return T5_ERR_INVALID;
    LOCKED(c);
    const t5_buf* ins[2] = {A, B};
    t5_buf* outs[2] = {X, STATUS};
    FOR_EACH_SHARD(c, A, ins, 2, outs, 2, false, "t5_solve", [&](const Shard& s, const void* const* ip, void* const* op) {
        return solve_chain(s, dtype, n, nrhs, ip[0], ip[1], op[0], op[1]);
    });
    return T5_OK;
}

// ---- tr / polyEval / charPoly / minPoly (the rest of SquareMatrix) -------------
int t5_trace(t5_ctx* c, int dtype, int n, const t5_buf* A, t5_buf* OUT) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, 1, A));
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[1] = {OUT};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 1, false, "t5_trace", [&](const Shard& s, const void* const* ip, void* const* op) {
        int r = small_trace(s, dtype, n, ip[0], op[0]);
        if (r == T5I_NOT_SPECIALISED) r = generic_trace(s, dtype, n, ip[0], op[0]);
        return r;
    });
    return T5_OK;
}

// polyEval coefficients / `pure` records ride in a grow-only per-device parameter buffer (stream-ordered copy, so a
// previous launch still reading the old contents is not disturbed)
static int param_upload(t5_ctx* c, int g, const void* host, size_t bytes) {
    if (c->coef_cap[g] < bytes) {
        if (c->coef[g]) { CU(c, cudaStreamSynchronize(c->stream[g])); CU(c, cudaFree(c->coef[g])); c->coef[g] = nullptr; c->coef_cap[g] = 0; }
        const size_t cap = bytes < 4096 ? 4096 : bytes;
        CU(c, cudaMalloc(&c->coef[g], cap));
        c->coef_cap[g] = cap;
    }
    if (bytes) CU(c, cudaMemcpyAsync(c->coef[g], host, bytes, cudaMemcpyHostToDevice, c->stream[g]));
    return T5_OK;
}

int t5_poly_eval(t5_ctx* c, int dtype, int n, const void* coeffs, int ncoef, const t5_buf* A, t5_buf* OUT) try {
    if (!c || ncoef < 0 || (ncoef > 0 && !coeffs)) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, (size_t)n * n, A));
    if (OUT == A) return T5_ERR_INVALID;
    const size_t cbytes = (size_t)ncoef * dtype_size(dtype);
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[1] = {OUT};
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        bool uploaded = false;
        int up_rc = T5_OK;
        CHECK(run_shard(c, g, ins, 1, outs, 1, false, "t5_poly_eval", [&](const Shard& s, const void* const* ip, void* const* op) {
            int r = small_poly_eval(s, dtype, n, coeffs, ncoef, ip[0], op[0]);
            if (r == T5I_NOT_SPECIALISED) {                     // any n <= 8 / any degree: the coefficients go to the device
                if (!uploaded) { up_rc = param_upload(c, g, coeffs, cbytes); uploaded = true; }
                if (up_rc != T5_OK) return (int)T5I_CUDA;
                r = generic_poly_eval(s, dtype, n, c->coef[g], ncoef, ip[0], op[0]);
            }
            return r;
        }));
    }
    return T5_OK;
} T5_NOEXCEPT_TAIL

int t5_char_poly(t5_ctx* c, int dtype, int n, const t5_buf* A, t5_buf* OUT) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, (size_t)n + 1, A));
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[1] = {OUT};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 1, false, "t5_char_poly", [&](const Shard& s, const void* const* ip, void* const* op) {
        int r = small_char_poly(s, dtype, n, ip[0], op[0]);
        if (r == T5I_NOT_SPECIALISED) r = generic_char_poly(s, dtype, n, ip[0], op[0]);
        return r;
    });
    return T5_OK;
}

int t5_min_poly(t5_ctx* c, int dtype, int n, const t5_buf* A, t5_buf* OUT, t5_buf* DEGREE) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)n * n, nullptr));
    CHECK(check(c, OUT, dtype, (size_t)n + 1, A));
    CHECK(check(c, DEGREE, T5_U8, 1, A));
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[2] = {OUT, DEGREE};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 2, false, "t5_min_poly", [&](const Shard& s, const void* const* ip, void* const* op) {
        int r = small_min_poly(s, dtype, n, ip[0], op[0], op[1]);
        if (r == T5I_NOT_SPECIALISED) r = generic_min_poly(s, dtype, n, ip[0], op[0], op[1]);
        return r;
    });
    return T5_OK;
}

int t5_replicate(t5_ctx* c, int dtype, const void* record, t5_buf* OUT) try {
    if (!c || !OUT || OUT->ctx != c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (OUT->dtype != dtype) return T5_ERR_INVALID;
    if (OUT->elems && !record) return T5_ERR_INVALID;
    const size_t rbytes = OUT->elems * dtype_size(dtype);
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!OUT->count[g] || !rbytes) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        CHECK(param_upload(c, g, record, rbytes));              // the record rides in the per-device parameter buffer
        int r = generic_replicate(shard_of(c, OUT, g), dtype_size(dtype), OUT->elems, c->coef[g], OUT->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_replicate");
        c->launches++;
    }
    return T5_OK;
} T5_NOEXCEPT_TAIL

int t5_row_echelon(t5_ctx* c, int dtype, int rows, int cols, int pivot_cols, const t5_buf* A, t5_buf* OUT, t5_buf* RANK) {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (rows < 1 || cols < 1 || pivot_cols < 0 || pivot_cols > cols) return T5_ERR_SHAPE;
    CHECK(check(c, A, dtype, (size_t)rows * cols, nullptr));
    CHECK(check(c, OUT, dtype, (size_t)rows * cols, A));
    CHECK(check(c, RANK, T5_U8, 1, A));
    LOCKED(c);
    const t5_buf* ins[1] = {A};
    t5_buf* outs[2] = {OUT, RANK};
    FOR_EACH_SHARD(c, A, ins, 1, outs, 2, false, "t5_row_echelon", [&](const Shard& s, const void* const* ip, void* const* op) {
        int r = small_row_echelon(s, dtype, rows, cols, pivot_cols, ip[0], op[0], op[1]);
        if (r == T5I_NOT_SPECIALISED) r = generic_row_echelon(s, dtype, rows, cols, pivot_cols, ip[0], op[0], op[1]);
        return r;
    });
    return T5_OK;
}

// ---- element-wise --------------------------------------------------------------------------
int t5_zip(t5_ctx* c, int dtype, int op, const t5_buf* A, const t5_buf* B, t5_buf* C) {
    if (!c || !A) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (op < 0 || op > 2) return T5_ERR_INVALID;
    CHECK(check(c, A, dtype, A->elems, nullptr));
    CHECK(check(c, B, dtype, A->elems, A));
    CHECK(check(c, C, dtype, A->elems, A));
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const size_t elems = (A->layout == T5_SOA ? A->cap[g] : A->count[g]) * A->elems;
        int r = generic_zip(shard_of(c, A, g), dtype, op, elems, A->ptr[g], B->ptr[g], C->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_zip");
        c->launches++;
    }
    return T5_OK;
}

int t5_scale(t5_ctx* c, int dtype, const void* scalar, const t5_buf* A, t5_buf* C) {
    if (!c || !A || !scalar) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    CHECK(check(c, A, dtype, A->elems, nullptr));
    CHECK(check(c, C, dtype, A->elems, A));
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const size_t elems = (A->layout == T5_SOA ? A->cap[g] : A->count[g]) * A->elems;
        int r = generic_scale(shard_of(c, A, g), dtype, scalar, elems, A->ptr[g], C->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_scale");
        c->launches++;
    }
    return T5_OK;
}

// ---- host-buffer pipeline --------------------------------------------------------------------
// The end-to-end form of every hot-path operation: operands and results are HOST AoS arrays
// (what a Haskell instance method on ordinary tensors holds).  The batch is cut into chunks;
// chunk j of device g runs on slot j % 3 (own stream, own staging buffers): H2D of the
// inputs, the kernel, D2H of the results.  With three slots the two PCIe directions and the
// SMs work on different chunks at the same time.  Pinned host memory (t5_pinned_alloc) gives
// full PCIe speed; pageable memory is correct but staged by the driver.
static int slot_reserve(t5_ctx* c, HostSlot& s, int i, size_t bytes) {
    if (!s.stream) CU(c, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (s.cap[i] >= bytes) return T5_OK;
    if (s.d[i]) { CU(c, cudaStreamSynchronize(s.stream)); CU(c, cudaFree(s.d[i])); s.d[i] = nullptr; s.cap[i] = 0; }
    CU(c, cudaMalloc(&s.d[i], bytes));
    s.cap[i] = bytes;
    return T5_OK;
}

struct HostOperand { const void* in; void* out; size_t item_bytes; };   // exactly one of in / out is set

// launch(g, shard, dev_ptrs[]) -> T5_* code; dev_ptrs follow the order of `ops`.
extern "C++" {
// One device's share of a host-array call: its chunks rotate over the device's three slots.  Whatever happens -
// a failed reservation, copy or launch - the function only returns after every slot of the device has drained:
// the caller's host arrays are never in flight when the entry returns (it may free them on error).
template <class Launch>
static int host_pipeline_device(t5_ctx* c, int g, size_t batch, size_t chunk, const HostOperand* ops, int nops, Launch& launch) {
    size_t first, count;
    t5_partition(batch, c->ndev, g, &first, &count);
    if (!count) return T5_OK;
    int rc = T5_OK;
    cudaError_t e = cudaSetDevice(c->dev[g]);
    if (e != cudaSuccess) return cuda_fail(c, e, "host pipeline: cudaSetDevice");
    const size_t slot_items = chunk < count ? chunk : count;      // staging is sized by what the call needs, not by the cap
    size_t j = 0;
    for (size_t off = 0; off < count && rc == T5_OK; off += chunk, ++j) {
        const size_t cnt = count - off < chunk ? count - off : chunk;
        HostSlot& s = c->slots[g][j % 3];
        void* dev[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int i = 0; i < nops && rc == T5_OK; ++i) {
            rc = slot_reserve(c, s, i, slot_items * ops[i].item_bytes);
            dev[i] = s.d[i];
        }
        if (rc != T5_OK) break;
        const size_t it = first + off;
        for (int i = 0; i < nops && rc == T5_OK; ++i)
            if (ops[i].in) {
                e = cudaMemcpyAsync(dev[i], (const char*)ops[i].in + it * ops[i].item_bytes, cnt * ops[i].item_bytes,
                                    cudaMemcpyHostToDevice, s.stream);
                if (e != cudaSuccess) rc = cuda_fail(c, e, "host pipeline: H2D copy");
            }
        if (rc != T5_OK) break;
        Shard sh; sh.stream = s.stream; sh.layout = T5L_AOS; sh.count = cnt; sh.soa_stride = 0; sh.first = it;
        sh.sched = (unsigned long long*)((char*)c->scratch[g] + SCHED_OFF) + 2 * (1 + j % 3);
        rc = launch(g, sh, dev);
        if (rc != T5_OK) break;
        for (int i = 0; i < nops && rc == T5_OK; ++i)
            if (ops[i].out) {
                e = cudaMemcpyAsync((char*)ops[i].out + it * ops[i].item_bytes, dev[i], cnt * ops[i].item_bytes,
                                    cudaMemcpyDeviceToHost, s.stream);
                if (e != cudaSuccess) rc = cuda_fail(c, e, "host pipeline: D2H copy");
            }
    }
    for (auto& s : c->slots[g])
        if (s.stream) {
            e = cudaStreamSynchronize(s.stream);
            if (e != cudaSuccess && rc == T5_OK) rc = cuda_fail(c, e, "host pipeline sync");
        }
    return rc;
}

template <class Launch>
static int host_pipeline(t5_ctx* c, size_t batch, const HostOperand* ops, int nops, Launch launch) {
    if (nops < 1 || nops > 4) return T5_ERR_INVALID;
    size_t per_item = 0;
    for (int i = 0; i < nops; ++i) {
        if (batch && !ops[i].in && !ops[i].out) return T5_ERR_INVALID;
        per_item += ops[i].item_bytes;
    }
    // Chunk size.  Every cudaMemcpyAsync costs ~15 us of setup on this platform (3 per chunk), and the
    // pipeline's fill and drain expose about 0.7 of one chunk's transfer time; with T = bytes / 70 GB/s the
    // sum 45 us * N + 0.7 T / N is smallest at chunk = sqrt(bytes * 4.5 MB): 120 MB for the 3.2 GB headline
    // call (measured 372 M items/s with fixed 32-MiB chunks, 392 M/s = 0.91 of the H2D ceiling with 128 MiB),
    // 32 MB for the 226 MB 3x3 call.  Clamped to [4, 256] MiB; a multiple of 256 items when that many fit,
    // else whatever fits (one item at least: a 128 MiB record is staged alone, not 256 at a time).
    const double total_bytes = (double)batch * (double)per_item / (double)(c->ndev > 0 ? c->ndev : 1);
    double cb = sqrt(total_bytes * 4.5e6);
    cb = cb < (double)(4ull << 20) ? (double)(4ull << 20) : (cb > (double)(256ull << 20) ? (double)(256ull << 20) : cb);
    size_t chunk = (size_t)cb / (per_item ? per_item : 1);
    chunk = chunk >= 256 ? chunk / 256 * 256 : (chunk ? chunk : 1);
    LOCKED(c);
    if (c->ndev == 1) return host_pipeline_device(c, 0, batch, chunk, ops, nops, launch);
    // One host thread per device: a single thread would enqueue device 0's whole chunk list before device 1 saw its
    // first copy (~45 us per chunk, ~3 ms before device 7 starts at 8 devices).  The context lock is held by the
    // caller's thread for the whole call; the workers only touch their own device's slots.
    std::vector<int> rcs((size_t)c->ndev, (int)T5_OK);
    std::vector<std::thread> workers;
    workers.reserve((size_t)c->ndev);
    for (int g = 1; g < c->ndev; ++g) {
        try {
            workers.emplace_back([&, g] { rcs[(size_t)g] = host_pipeline_device(c, g, batch, chunk, ops, nops, launch); });
        } catch (...) {                                       // no thread to be had: this device runs on the caller's thread
            rcs[(size_t)g] = host_pipeline_device(c, g, batch, chunk, ops, nops, launch);
        }
    }
    rcs[0] = host_pipeline_device(c, 0, batch, chunk, ops, nops, launch);
    for (auto& w : workers) w.join();
    for (int r : rcs) if (r != T5_OK) return r;
    return T5_OK;
}
}  // extern "C++"

static int product_host(t5_ctx* c, int dtype, int P, int K, int Q, int nc, size_t batch, const void* A, const void* B, void* C,
                        const char* what) {
    if (dtype != T5_F64 && dtype != T5_F32 && dtype != T5_I64) return T5_ERR_UNSUPPORTED;
    if (P < 1 || K < 1 || Q < 1) return T5_ERR_SHAPE;
    const size_t es = dtype_size(dtype);
    const HostOperand ops[3] = {{A, nullptr, (size_t)P * K * es}, {B, nullptr, (size_t)K * Q * es}, {nullptr, C, (size_t)P * Q * es}};
    return host_pipeline(c, batch, ops, 3, [&](int, const Shard& sh, void* const* d) {
        int r = product_shard(sh, dtype, P, K, Q, nc, d[0], d[1], d[2]);
        if (r != T5I_OK) return map_internal(c, r, what);
        c->launches++;
        return (int)T5_OK;
    });
}

int t5_matmul_host(t5_ctx* c, int dtype, int m, int k, int n, size_t batch, const void* A, const void* B, void* C) try {
    if (!c) return T5_ERR_INVALID;
    return product_host(c, dtype, m, k, n, 1, batch, A, B, C, "t5_matmul_host");
} T5_NOEXCEPT_TAIL

int t5_dot_host(t5_ctx* c, int dtype, int n, size_t batch, const void* X, const void* Y, void* OUT) try {
    if (!c) return T5_ERR_INVALID;
    return product_host(c, dtype, 1, n, 1, 1, batch, X, Y, OUT, "t5_dot_host");
} T5_NOEXCEPT_TAIL

int t5_prod_host(t5_ctx* c, int dtype, int rankA, const uint64_t* dimsA, int rankB, const uint64_t* dimsB, int nc,
                 size_t batch, const void* A, const void* B, void* C) try {
    if (!c) return T5_ERR_INVALID;
    int P, K, Q;
    CHECK(prod_dims(rankA, dimsA, rankB, dimsB, nc, &P, &K, &Q));
    return product_host(c, dtype, P, K, Q, nc, batch, A, B, C, "t5_prod_host");
} T5_NOEXCEPT_TAIL

int t5_transpose_host(t5_ctx* c, int dtype, int rank, const uint64_t* dims, size_t batch, const void* A, void* OUT) try {
    if (!c) return T5_ERR_INVALID;
    std::vector<uint64_t> sq;
    size_t d = 1;
    CHECK(transpose_args(dtype, rank, dims, sq, d));
    const int es = dtype_size(dtype);
    const HostOperand ops[2] = {{A, nullptr, d * es}, {nullptr, OUT, d * es}};
    return host_pipeline(c, batch, ops, 2, [&](int g, const Shard& sh, void* const* dv) {
        return transpose_shard(c, g, sh, es, sq, d, dv[0], dv[1], sh.count * d * es);
    });
} T5_NOEXCEPT_TAIL

int t5_det_host(t5_ctx* c, int dtype, int n, size_t batch, const void* A, void* OUT) try {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    const size_t es = dtype_size(dtype);
    const HostOperand ops[2] = {{A, nullptr, (size_t)n * n * es}, {nullptr, OUT, es}};
    return host_pipeline(c, batch, ops, 2, [&](int, const Shard& sh, void* const* d) {
        int r = det_chain(sh, dtype, n, d[0], d[1]);
        if (r != T5I_OK) return map_internal(c, r, "t5_det_host");
        c->launches++;
        return (int)T5_OK;
    });
} T5_NOEXCEPT_TAIL

int t5_inverse_host(t5_ctx* c, int dtype, int n, size_t batch, const void* A, void* OUT, uint8_t* status) try {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1) return T5_ERR_SHAPE;
    const size_t es = dtype_size(dtype);
    const HostOperand ops[3] = {{A, nullptr, (size_t)n * n * es}, {nullptr, OUT, (size_t)n * n * es}, {nullptr, status, 1}};
    return host_pipeline(c, batch, ops, 3, [&](int, const Shard& sh, void* const* d) {
        int r = inverse_chain(sh, dtype, n, d[0], d[1], d[2]);
        if (r != T5I_OK) return map_internal(c, r, "t5_inverse_host");
        c->launches++;
        return (int)T5_OK;
    });
} T5_NOEXCEPT_TAIL

int t5_solve_host(t5_ctx* c, int dtype, int n, int nrhs, size_t batch, const void* A, const void* B, void* X, uint8_t* status) try {
    if (!c) return T5_ERR_INVALID;
    if (dtype != T5_F64 && dtype != T5_F32) return T5_ERR_UNSUPPORTED;
    if (n < 1 || nrhs < 1) return T5_ERR_SHAPE;
    const size_t es = dtype_size(dtype);
    const HostOperand ops[4] = {{A, nullptr, (size_t)n * n * es}, {B, nullptr, (size_t)n * nrhs * es},
                                {nullptr, X, (size_t)n * nrhs * es}, {nullptr, status, 1}};
    return host_pipeline(c, batch, ops, 4, [&](int, const Shard& sh, void* const* d) {
        int r = solve_chain(sh, dtype, n, nrhs, d[0], d[1], d[2], d[3]);
        if (r != T5I_OK) return map_internal(c, r, "t5_solve_host");
        c->launches++;
        return (int)T5_OK;
    });
} T5_NOEXCEPT_TAIL

// ---- measurement helpers -----------------------------------------------------------------------
int t5_timer_begin(t5_ctx* c) {
    if (!c) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        CU(c, cudaEventRecord(c->ev0[g], c->stream[g]));
    }
    return T5_OK;
}

int t5_timer_end(t5_ctx* c, float* ms_max) {
    if (!c || !ms_max) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        CU(c, cudaEventRecord(c->ev1[g], c->stream[g]));
    }
    float mx = 0.f;
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        CU(c, cudaEventSynchronize(c->ev1[g]));
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, c->ev0[g], c->ev1[g]));
        if (ms > mx) mx = ms;
    }
    *ms_max = mx;
    return T5_OK;
}

int t5_timer_per_device(t5_ctx* c, float* ms, int cap) {
    if (!c || !ms || cap < 0) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev && g < cap; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        CU(c, cudaEventSynchronize(c->ev1[g]));
        CU(c, cudaEventElapsedTime(&ms[g], c->ev0[g], c->ev1[g]));
    }
    return T5_OK;
}

int t5_flush_l2(t5_ctx* c) {
    if (!c) return T5_ERR_INVALID;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        CU(c, cudaSetDevice(c->dev[g]));
        if (!c->flush[g]) CU(c, cudaMalloc(&c->flush[g], FLUSH_BYTES));
        CU(c, cudaMemsetAsync(c->flush[g], 0x5a, FLUSH_BYTES, c->stream[g]));
    }
    return T5_OK;
}

int t5_fill_synthetic(t5_buf* b, uint64_t seed, uint64_t op_id) {
    if (!b) return T5_ERR_INVALID;
    t5_ctx* c = b->ctx;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!b->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        int r = fill_synthetic(shard_of(c, b, g), b->dtype, b->elems, seed, op_id, b->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_fill_synthetic");
        c->launches++;
    }
    return T5_OK;
}

int t5_plant_specials(t5_buf* b, int n, uint64_t swap_every, uint64_t sing_every) {
    if (!b) return T5_ERR_INVALID;
    if (n < 1 || b->elems != (size_t)n * n) return T5_ERR_SHAPE;
    t5_ctx* c = b->ctx;
    LOCKED(c);
    for (int g = 0; g < c->ndev; ++g) {
        if (!b->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        int r = plant_specials(shard_of(c, b, g), b->dtype, n, swap_every, sing_every, b->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_plant_specials");
        c->launches++;
    }
    return T5_OK;
}

}  // extern "C"
#pragma GCC visibility pop

// ---- structural operations: affine gather tables + dispatch -----------------------------------
// Every structural op of the flat-vector backend (Vector.hs:260-354, :400-429) maps an output
// multi-index to an input offset affinely: in = base + sum_a oidx[a] * stride[a].
namespace {
struct View { std::vector<uint64_t> dims, stride; uint64_t base = 0; };

std::vector<uint64_t> row_major_strides(const std::vector<uint64_t>& d) {
    std::vector<uint64_t> s(d.size(), 1);
    for (int a = (int)d.size() - 2; a >= 0; --a) s[a] = s[a + 1] * d[a + 1];
    return s;
}
uint64_t prod_of(const std::vector<uint64_t>& d) { uint64_t p = 1; for (auto x : d) p *= x; return p; }

void view_table(const View& v, uint64_t add, std::vector<uint32_t>& out) {
    const int r = (int)v.dims.size();
    const uint64_t n = prod_of(v.dims);
    std::vector<uint64_t> idx(r, 0);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t off = v.base;
        for (int a = 0; a < r; ++a) off += idx[a] * v.stride[a];
        out.push_back((uint32_t)(off + add));
        for (int a = r - 1; a >= 0; --a) { if (++idx[a] < v.dims[a]) break; idx[a] = 0; }
    }
}

// returns T5_OK and fills table (+ out dims) or an error code
int build_structural(int op, const std::vector<uint64_t>& dims, const std::vector<uint64_t>& prm,
                     std::vector<uint32_t>& table, std::vector<uint64_t>& out_dims) {
    const int r = (int)dims.size();
    for (auto x : dims) if (x == 0 || x > (1u << 28)) return T5_ERR_SHAPE;
    if (prod_of(dims) > (1u << 28)) return T5_ERR_SHAPE;
    const auto st = row_major_strides(dims);
    table.clear();
    switch (op) {
        case 0: {  // append: params = axis, dimsB...
            if ((int)prm.size() != 1 + r || r < 1) return T5_ERR_SHAPE;
            const int ax = (int)prm[0] - 1;
            if (ax < 0 || ax >= r) return T5_ERR_SHAPE;
            std::vector<uint64_t> db(prm.begin() + 1, prm.end());
            for (int a = 0; a < r; ++a) if (a != ax && db[a] != dims[a]) return T5_ERR_SHAPE;
            if (db[ax] == 0 || prod_of(db) + prod_of(dims) > (1u << 28)) return T5_ERR_SHAPE;
            const auto sb = row_major_strides(db);
            out_dims = dims; out_dims[ax] = dims[ax] + db[ax];
            const uint64_t dA = prod_of(dims), n = prod_of(out_dims);
            std::vector<uint64_t> idx(r, 0);
            for (uint64_t i = 0; i < n; ++i) {
                uint64_t off = 0;
                if (idx[ax] < dims[ax]) { for (int a = 0; a < r; ++a) off += idx[a] * st[a]; }
                else { for (int a = 0; a < r; ++a) off += (a == ax ? idx[a] - dims[ax] : idx[a]) * sb[a]; off += dA; }
                table.push_back((uint32_t)off);
                for (int a = r - 1; a >= 0; --a) { if (++idx[a] < out_dims[a]) break; idx[a] = 0; }
            }
            return T5_OK;
        }
        case 1: case 2: {  // split halves: params = axis, first_extent
            if (prm.size() != 2 || r < 1) return T5_ERR_SHAPE;
            const int ax = (int)prm[0] - 1;
            if (ax < 0 || ax >= r || prm[1] == 0 || prm[1] >= dims[ax]) return T5_ERR_SHAPE;
            View v; v.dims = dims; v.stride = st;
            if (op == 1) v.dims[ax] = prm[1];
            else { v.dims[ax] = dims[ax] - prm[1]; v.base = prm[1] * st[ax]; }
            out_dims = v.dims;
            view_table(v, 0, table);
            return T5_OK;
        }
        case 3: {  // at: params = index over the tail axes
            if (r < 1 || (int)prm.size() != r - 1) return T5_ERR_SHAPE;
            View v; v.dims = {dims[0]}; v.stride = {st[0]};
            for (int a = 1; a < r; ++a) { if (prm[a - 1] < 1 || prm[a - 1] > dims[a]) return T5_ERR_SHAPE; v.base += (prm[a - 1] - 1) * st[a]; }
            out_dims = v.dims;
            view_table(v, 0, table);
            return T5_OK;
        }
        case 4: {  // ta: params = i
            if (r < 1 || prm.size() != 1 || prm[0] < 1 || prm[0] > dims[0]) return T5_ERR_SHAPE;
            View v; v.dims.assign(dims.begin() + 1, dims.end()); v.stride.assign(st.begin() + 1, st.end());
            v.base = (prm[0] - 1) * st[0];
            out_dims = v.dims;
            view_table(v, 0, table);
            return T5_OK;
        }
        case 5: {  // slice: params = slicer
            if ((int)prm.size() != r) return T5_ERR_SHAPE;
            View v;
            for (int a = 0; a < r; ++a) {
                if (prm[a] == 0) { v.dims.push_back(dims[a]); v.stride.push_back(st[a]); }
                else { if (prm[a] > dims[a]) return T5_ERR_SHAPE; v.base += (prm[a] - 1) * st[a]; }
            }
            out_dims = v.dims;
            view_table(v, 0, table);
            return T5_OK;
        }
        case 6: {  // permute_axes: params = perm
            if ((int)prm.size() != r) return T5_ERR_SHAPE;
            std::vector<int> seen(r, 0);
            View v;
            for (int a = 0; a < r; ++a) {
                if (prm[a] >= (uint64_t)r || seen[prm[a]]++) return T5_ERR_SHAPE;
                v.dims.push_back(dims[prm[a]]); v.stride.push_back(st[prm[a]]);
            }
            out_dims = v.dims;
            view_table(v, 0, table);
            return T5_OK;
        }
    }
    return T5_ERR_INVALID;
}
}  // namespace

// table cached per (op, dims, params) and per device; built (and thereby validated) only on a miss
static std::vector<uint64_t> structural_key(int op, const std::vector<uint64_t>& dims, const std::vector<uint64_t>& prm) {
    std::vector<uint64_t> key{0xFFFFFFFFull, (uint64_t)op, (uint64_t)dims.size()};
    key.insert(key.end(), dims.begin(), dims.end());
    key.insert(key.end(), prm.begin(), prm.end());
    return key;
}

static int structural_table_dev(t5_ctx* c, int g, int op, const std::vector<uint64_t>& dims, const std::vector<uint64_t>& prm,
                                const uint32_t** dev_table, size_t* d_out, std::vector<uint64_t>* out_dims) {
    return table_get(c, g, structural_key(op, dims, prm),
                     [&](std::vector<uint32_t>& table, std::vector<uint64_t>& od) { return build_structural(op, dims, prm, table, od); },
                     dev_table, d_out, out_dims);
}

static int structural_unary(t5_ctx* c, int dtype, int op, int rank, const uint64_t* dims, const std::vector<uint64_t>& prm,
                            const t5_buf* A, t5_buf* OUT, const char* what) {
    if (!c || rank < 0 || rank > 16 || (rank && !dims)) return T5_ERR_INVALID;
    const int es = dtype_size(dtype);
    if (es != 4 && es != 8) return T5_ERR_UNSUPPORTED;
    std::vector<uint64_t> dv(dims, dims + rank);
    LOCKED(c);
    size_t d_out = 0;
    CHECK(structural_table_dev(c, -1, op, dv, prm, nullptr, &d_out, nullptr));
    const size_t d_in = prod_of(dv);
    CHECK(check(c, A, dtype, d_in, nullptr));
    CHECK(check(c, OUT, dtype, d_out, A));
    if (OUT == A) return T5_ERR_INVALID;
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const uint32_t* tab; size_t n;
        CHECK(structural_table_dev(c, g, op, dv, prm, &tab, &n, nullptr));
        Shard s = shard_of(c, A, g);
        int r = s.layout == T5L_SOA ? generic_soa_gather(s, es, d_in, d_out, tab, A->ptr[g], nullptr, OUT->ptr[g])
                                    : generic_gather(s, es, d_in, d_out, tab, A->ptr[g], OUT->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, what);
        c->launches++;
    }
    return T5_OK;
}

#pragma GCC visibility push(default)
extern "C" {

int t5_structural_table(int op, int rank, const uint64_t* dims, const uint64_t* params, int nparams,
                        uint32_t* table_out, size_t table_cap, size_t* d_out) try {
    if (rank < 0 || rank > 16 || (rank && !dims) || nparams < 0 || (nparams && !params) || !d_out) return T5_ERR_INVALID;
    std::vector<uint64_t> dv(dims, dims + rank), pv(params, params + nparams);
    std::vector<uint32_t> table; std::vector<uint64_t> od;
    int rc = build_structural(op, dv, pv, table, od);
    if (rc != T5_OK) return rc;
    *d_out = table.size();
    if (table_out) for (size_t i = 0; i < table.size() && i < table_cap; ++i) table_out[i] = table[i];
    return T5_OK;
} T5_NOEXCEPT_TAIL

int t5_append(t5_ctx* c, int dtype, int axis, int rank, const uint64_t* dimsA, const uint64_t* dimsB,
              const t5_buf* A, const t5_buf* B, t5_buf* OUT) try {
    if (!c || rank < 1 || rank > 16 || !dimsA || !dimsB) return T5_ERR_INVALID;
    const int es = dtype_size(dtype);
    if (es != 4 && es != 8) return T5_ERR_UNSUPPORTED;
    std::vector<uint64_t> da(dimsA, dimsA + rank), prm{(uint64_t)axis};
    prm.insert(prm.end(), dimsB, dimsB + rank);
    LOCKED(c);
    size_t d_out = 0;
    CHECK(structural_table_dev(c, -1, 0, da, prm, nullptr, &d_out, nullptr));
    const size_t dA = prod_of(da), dB = d_out - dA;
    CHECK(check(c, A, dtype, dA, nullptr));
    CHECK(check(c, B, dtype, dB, A));
    CHECK(check(c, OUT, dtype, dA + dB, A));
    if (OUT == A || OUT == B) return T5_ERR_INVALID;
    for (int g = 0; g < c->ndev; ++g) {
        if (!A->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const uint32_t* tab; size_t n;
        CHECK(structural_table_dev(c, g, 0, da, prm, &tab, &n, nullptr));
        Shard s = shard_of(c, A, g);
        int r = s.layout == T5L_SOA ? generic_soa_gather(s, es, dA, dA + dB, tab, A->ptr[g], B->ptr[g], OUT->ptr[g])
                                    : generic_gather2(s, es, dA, dB, dA + dB, tab, A->ptr[g], B->ptr[g], OUT->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_append");
        c->launches++;
    }
    return T5_OK;
} T5_NOEXCEPT_TAIL

static int split_locked(t5_ctx* c, int dtype, int axis, int rank, const uint64_t* dims, uint64_t first_extent,
                        const t5_buf* IN, t5_buf* OUT1, t5_buf* OUT2) {
    const int es = dtype_size(dtype);
    std::vector<uint64_t> dv(dims, dims + rank), prm{(uint64_t)axis, first_extent};
    LOCKED(c);
    size_t d1 = 0, d2 = 0;
    CHECK(structural_table_dev(c, -1, 1, dv, prm, nullptr, &d1, nullptr));
    CHECK(structural_table_dev(c, -1, 2, dv, prm, nullptr, &d2, nullptr));
    const size_t d_in = prod_of(dv);
    CHECK(check(c, IN, dtype, d_in, nullptr));
    CHECK(check(c, OUT1, dtype, d1, IN));
    CHECK(check(c, OUT2, dtype, d2, IN));
    // the two gather tables partition the input: invert them into one scatter table, read the input once
    std::vector<uint64_t> key{0xFFFFFFFEull, (uint64_t)rank};
    key.insert(key.end(), dv.begin(), dv.end());
    key.insert(key.end(), prm.begin(), prm.end());
    auto build = [&](std::vector<uint32_t>& sc, std::vector<uint64_t>& od) {
        std::vector<uint32_t> t1, t2;
        int rc = build_structural(1, dv, prm, t1, od);
        if (rc == T5_OK) rc = build_structural(2, dv, prm, t2, od);
        if (rc != T5_OK) return rc;
        sc.assign(d_in, 0);
        for (size_t o = 0; o < t1.size(); ++o) sc[t1[o]] = (uint32_t)o;
        for (size_t o = 0; o < t2.size(); ++o) sc[t2[o]] = 0x80000000u | (uint32_t)o;
        return (int)T5_OK;
    };
    for (int g = 0; g < c->ndev; ++g) {
        if (!IN->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const uint32_t* tab = nullptr;
        CHECK(table_get(c, g, key, build, &tab, nullptr, nullptr));
        int r = generic_scatter2(shard_of(c, IN, g), es, d_in, d1, d2, tab, IN->ptr[g], OUT1->ptr[g], OUT2->ptr[g]);
        if (r != T5I_OK) return map_internal(c, r, "t5_split");
        c->launches++;
    }
    return T5_OK;
}

// AoS records small enough to stage: both halves from one read of the input, 16-byte accesses on every side
// (4Mi 4x4 -> 4x3 | 4x1: the two Float gathers ran at 0.55 of the HBM peak, the Double scatter at 0.78)
static int split_staged(t5_ctx* c, int dtype, int axis, int rank, const uint64_t* dims, uint64_t first_extent,
                        const t5_buf* IN, t5_buf* OUT1, t5_buf* OUT2, bool* done) {
    *done = false;
    const int es = dtype_size(dtype);
    std::vector<uint64_t> dv(dims, dims + rank), prm{(uint64_t)axis, first_extent};
    LOCKED(c);
    size_t d1 = 0, d2 = 0;
    CHECK(structural_table_dev(c, -1, 1, dv, prm, nullptr, &d1, nullptr));
    CHECK(structural_table_dev(c, -1, 2, dv, prm, nullptr, &d2, nullptr));
    const size_t d_in = prod_of(dv);
    if (!staged_gather_fits(es, d_in, d1 + d2, d1 && d2)) return T5_OK;
    CHECK(check(c, IN, dtype, d_in, nullptr));
    CHECK(check(c, OUT1, dtype, d1, IN));
    CHECK(check(c, OUT2, dtype, d2, IN));
    for (int g = 0; g < c->ndev; ++g) {
        if (!IN->count[g]) continue;
        CU(c, cudaSetDevice(c->dev[g]));
        const uint32_t *tab1, *tab2; size_t n;
        CHECK(structural_table_dev(c, g, 1, dv, prm, &tab1, &n, nullptr));
        CHECK(structural_table_dev(c, g, 2, dv, prm, &tab2, &n, nullptr));
        const Shard s = shard_of(c, IN, g);
        int r = staged_gather(s, es, d_in, 0, d1, d2, tab1, tab2, IN->ptr[g], nullptr, OUT1->ptr[g], OUT2->ptr[g]);
        if (r == T5I_NOT_SPECIALISED) {                       // (a misaligned shard: never with the library's own allocations)
            r = generic_gather(s, es, d_in, d1, tab1, IN->ptr[g], OUT1->ptr[g]);
            if (r == T5I_OK) r = generic_gather(s, es, d_in, d2, tab2, IN->ptr[g], OUT2->ptr[g]);
        }
        if (r != T5I_OK) return map_internal(c, r, "t5_split");
        c->launches++;
    }
    *done = true;
    return T5_OK;
}

int t5_split(t5_ctx* c, int dtype, int axis, int rank, const uint64_t* dims, uint64_t first_extent,
             const t5_buf* IN, t5_buf* OUT1, t5_buf* OUT2) try {
    if (!c || !IN || rank < 0 || rank > 16 || (rank && !dims)) return T5_ERR_INVALID;
    if (OUT1 == OUT2 || OUT1 == IN || OUT2 == IN) return T5_ERR_INVALID;
    const int es = dtype_size(dtype);
    if (es != 4 && es != 8) return T5_ERR_UNSUPPORTED;
    if (IN->layout != T5_SOA && OUT1 && OUT2 && OUT1->layout != T5_SOA && OUT2->layout != T5_SOA) {
        bool done = false;
        const int rc = split_staged(c, dtype, axis, rank, dims, first_extent, IN, OUT1, OUT2, &done);
        if (rc != T5_OK || done) return rc;
    }
    if (es == 4 || IN->layout == T5_SOA) {      // (SoA: each half is a copy of whole rows)
        // 4-byte elements: the scatter's short store runs cost more than the second read saves
        // (4Mi 4x4 -> 4x3 | 4x1: Float 0.145 ms as two gathers vs 0.160 ms; Double 0.278 vs 0.211 ms)
        std::vector<uint64_t> prm{(uint64_t)axis, first_extent};
        CHECK(structural_unary(c, dtype, 1, rank, dims, prm, IN, OUT1, "t5_split"));
        return structural_unary(c, dtype, 2, rank, dims, prm, IN, OUT2, "t5_split");
    }
    return split_locked(c, dtype, axis, rank, dims, first_extent, IN, OUT1, OUT2);
} T5_NOEXCEPT_TAIL

int t5_at(t5_ctx* c, int dtype, int rank, const uint64_t* dims, const uint64_t* index, const t5_buf* A, t5_buf* OUT) try {
    if (rank < 1 || (rank > 1 && !index)) return T5_ERR_INVALID;
    return structural_unary(c, dtype, 3, rank, dims, std::vector<uint64_t>(index, index + (rank - 1)), A, OUT, "t5_at");
} T5_NOEXCEPT_TAIL

int t5_ta(t5_ctx* c, int dtype, int rank, const uint64_t* dims, uint64_t index, const t5_buf* A, t5_buf* OUT) try {
    return structural_unary(c, dtype, 4, rank, dims, std::vector<uint64_t>{index}, A, OUT, "t5_ta");
} T5_NOEXCEPT_TAIL

int t5_slice(t5_ctx* c, int dtype, int rank, const uint64_t* dims, const uint64_t* slicer, const t5_buf* A, t5_buf* OUT) try {
    if (rank < 0 || (rank && !slicer)) return T5_ERR_INVALID;
    return structural_unary(c, dtype, 5, rank, dims, std::vector<uint64_t>(slicer, slicer + rank), A, OUT, "t5_slice");
} T5_NOEXCEPT_TAIL

int t5_permute_axes(t5_ctx* c, int dtype, int rank, const uint64_t* dims, const int* perm, const t5_buf* A, t5_buf* OUT) try {
    if (rank < 0 || (rank && !perm)) return T5_ERR_INVALID;
    std::vector<uint64_t> p;
    for (int a = 0; a < rank; ++a) { if (perm[a] < 0) return T5_ERR_SHAPE; p.push_back((uint64_t)perm[a]); }
    return structural_unary(c, dtype, 6, rank, dims, p, A, OUT, "t5_permute_axes");
} T5_NOEXCEPT_TAIL

}  // extern "C"
#pragma GCC visibility pop
