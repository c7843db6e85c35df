// Context, error reporting, launch accounting and per-class event timing.
#include "common.cuh"

#include <algorithm>
#include <mutex>
#include <unordered_map>

static thread_local char g_err[1024] = "";

// ---- device-buffer cache ------------------------------------------------------
namespace {
struct PoolRec {
    hg_ctx* ctx;
    size_t size;
};
std::mutex g_pool_mu;
std::unordered_map<void*, PoolRec> g_pool_live;  // every block handed out by hg_dmalloc_bytes
std::unordered_map<void*, PoolRec> g_hpool_live;  // pinned host blocks handed out by hg_hmalloc_bytes

size_t pool_cap_bytes() {
    static size_t cap = [] {
        const char* e = getenv("HG_POOL_GB");
        const double gb = e ? atof(e) : 64.0;
        return (size_t)(gb * 1073741824.0);
    }();
    return cap;
}
}  // namespace

cudaError_t hg_dmalloc_bytes(hg_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    const size_t want = (std::max<size_t>(bytes, 1) + 511) / 512 * 512;
    if (pool_cap_bytes() > 0) {  // every size is cached (rounded up to 512 B)
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = ctx->pool_free.lower_bound(want);
        if (it != ctx->pool_free.end() && it->first <= want + want / 8) {
            *p = it->second;
            ctx->pool_cached -= it->first;
            g_pool_live[*p] = PoolRec{ctx, it->first};
            ctx->pool_free.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {  // give the cache back to the driver and try once more
        cudaGetLastError();
        hg_pool_trim(ctx);
        e = cudaMalloc(p, want);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool_live[*p] = PoolRec{ctx, want};
    }
    return e;
}

void hg_dfree(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_pool_live.find(p);
        if (it != g_pool_live.end()) {
            const PoolRec rec = it->second;
            g_pool_live.erase(it);
            if (rec.ctx && rec.ctx->pool_cached + rec.size <= pool_cap_bytes()) {
                rec.ctx->pool_free.emplace(rec.size, p);
                rec.ctx->pool_cached += rec.size;
                return;
            }
        }
    }
    cudaFree(p);
}

cudaError_t hg_hmalloc_bytes(hg_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    const size_t want = (std::max<size_t>(bytes, 1) + 511) / 512 * 512;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = ctx->hpool_free.lower_bound(want);
        if (it != ctx->hpool_free.end() && it->first <= 2 * want) {
            *p = it->second;
            g_hpool_live[*p] = PoolRec{ctx, it->first};
            ctx->hpool_free.erase(it);
            return cudaSuccess;
        }
    }
    const cudaError_t e = cudaMallocHost(p, want);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_hpool_live[*p] = PoolRec{ctx, want};
    }
    return e;
}

void hg_hfree(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        auto it = g_hpool_live.find(p);
        if (it != g_hpool_live.end()) {
            const PoolRec rec = it->second;
            g_hpool_live.erase(it);
            if (rec.ctx) {
                rec.ctx->hpool_free.emplace(rec.size, p);
                return;
            }
        }
    }
    cudaFreeHost(p);
}

static thread_local hg_ctx* g_alloc_ctx = nullptr;
hg_alloc_scope::hg_alloc_scope(hg_ctx* c) : prev(g_alloc_ctx) { g_alloc_ctx = c; }
hg_alloc_scope::~hg_alloc_scope() { g_alloc_ctx = prev; }
cudaError_t hg_dmalloc_cur(void** p, size_t bytes) {
    return g_alloc_ctx ? hg_dmalloc_bytes(g_alloc_ctx, p, bytes) : cudaMalloc(p, std::max<size_t>(bytes, 1));
}
cudaError_t hg_hmalloc_cur(void** p, size_t bytes) {
    return g_alloc_ctx ? hg_hmalloc_bytes(g_alloc_ctx, p, bytes) : cudaMallocHost(p, std::max<size_t>(bytes, 1));
}

void hg_pool_trim(hg_ctx* ctx) {
    std::vector<void*> blocks, hblocks;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (auto& kv : ctx->pool_free) blocks.push_back(kv.second);
        ctx->pool_free.clear();
        ctx->pool_cached = 0;
        for (auto& kv : ctx->hpool_free) hblocks.push_back(kv.second);
        ctx->hpool_free.clear();
    }
    if (!blocks.empty() || !hblocks.empty()) cudaStreamSynchronize(ctx->stream);
    for (void* b : blocks) cudaFree(b);
    for (void* b : hblocks) cudaFreeHost(b);
}

extern "C" int hg_ctx_trim(hg_ctx* ctx) {
    HG_REQUIRE(ctx, "hg_ctx_trim: ctx is NULL");
    HG_CUDA(cudaSetDevice(ctx->device));
    hg_pool_trim(ctx);
    return HG_OK;
}

void hg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* hg_last_error(void) { return g_err; }
extern "C" int hg_version(void) { return 100; }

extern "C" int hg_ctx_create(int device, void* stream, hg_ctx** out) {
    HG_REQUIRE(out != nullptr, "hg_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        hg_set_error("hg_ctx_create: no CUDA device available (%s); libhgmres has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return HG_ERR_CUDA;
    }
    HG_REQUIRE(device >= 0 && device < ndev, "hg_ctx_create: device %d out of range [0,%d)", device,
               ndev);
    HG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        hg_set_error("hg_ctx_create: device %d is sm_%d%d; libhgmres is built for sm_100a only",
                     device, prop.major, prop.minor);
        return HG_ERR_CUDA;
    }
    hg_ctx* c = new (std::nothrow) hg_ctx();
    if (!c) {
        hg_set_error("hg_ctx_create: out of host memory");
        return HG_ERR_NOMEM;
    }
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t ce = cudaSuccess;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        c->own_stream = ce == cudaSuccess;
    }
    if (ce == cudaSuccess) ce = cudaMalloc(&c->d_scalars, 64 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemsetAsync(c->d_scalars, 0, 64 * sizeof(double), c->stream);
    if (ce == cudaSuccess) ce = cudaMallocHost(&c->h_scalars, 64 * sizeof(double));
    if (ce != cudaSuccess) {  // release what was created: no leak on the error path
        hg_set_error("hg_ctx_create: %s", cudaGetErrorString(ce));
        if (c->d_scalars) cudaFree(c->d_scalars);
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return HG_ERR_CUDA;
    }
    *out = c;
    return HG_OK;
}

extern "C" int hg_ctx_destroy(hg_ctx* ctx) {
    if (!ctx) return HG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& p : ctx->pending) {
        ctx->event_pool.push_back(p.e0);
        ctx->event_pool.push_back(p.e1);
    }
    for (auto ev : ctx->event_pool) cudaEventDestroy(ev);
    hg_pool_trim(ctx);
    {   // blocks still held by live objects outlive the context: they will be cudaFree'd directly
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (auto& kv : g_pool_live)
            if (kv.second.ctx == ctx) kv.second.ctx = nullptr;
        for (auto& kv : g_hpool_live)
            if (kv.second.ctx == ctx) kv.second.ctx = nullptr;
    }
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HG_OK;
}

extern "C" int hg_ctx_sync(hg_ctx* ctx) {
    HG_REQUIRE(ctx, "hg_ctx_sync: ctx is NULL");
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    return HG_OK;
}

extern "C" int hg_ctx_launch_count(hg_ctx* ctx, uint64_t* out) {
    HG_REQUIRE(ctx && out, "hg_ctx_launch_count: NULL argument");
    *out = ctx->launches;
    return HG_OK;
}

int hg_ensure_partials(hg_ctx* ctx, size_t ndoubles) {
    if (ndoubles <= ctx->partials_cap) return HG_OK;
    // growing frees the old buffer: wait for anything that may still read it
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_partials) HG_CUDA(cudaFree(ctx->d_partials));
    ctx->d_partials = nullptr;
    ctx->partials_cap = 0;
    size_t cap = ndoubles + ndoubles / 2 + 1024;
    HG_CUDA(cudaMalloc(&ctx->d_partials, cap * sizeof(double)));
    ctx->partials_cap = cap;
    return HG_OK;
}

// ---- launch bracket ---------------------------------------------------------
static cudaEvent_t take_event(hg_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

hg_launch_scope::hg_launch_scope(hg_ctx* c, int k, double bytes) : ctx(c), klass(k) {
    ctx->launches++;
    if (ctx->timing) {
        e0 = take_event(ctx);
        e1 = take_event(ctx);
        ctx->t_bytes[klass] += bytes;
        cudaEventRecord(e0, ctx->stream);
    }
}

hg_launch_scope::~hg_launch_scope() {
    if (e0) {
        cudaEventRecord(e1, ctx->stream);
        ctx->pending.push_back({klass, e0, e1});
    }
}

extern "C" int hg_ctx_timing_enable(hg_ctx* ctx, int on) {
    HG_REQUIRE(ctx, "hg_ctx_timing_enable: ctx is NULL");
    ctx->timing = on != 0;
    return HG_OK;
}

static int fold_pending(hg_ctx* ctx) {
    HG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto& p : ctx->pending) {
        float ms = 0.f;
        HG_CUDA(cudaEventElapsedTime(&ms, p.e0, p.e1));
        ctx->t_ms[p.klass] += (double)ms;
        ctx->t_count[p.klass] += 1;
        ctx->event_pool.push_back(p.e0);
        ctx->event_pool.push_back(p.e1);
    }
    ctx->pending.clear();
    return HG_OK;
}

extern "C" int hg_ctx_timing_get(hg_ctx* ctx, int kernel_class, double* ms, uint64_t* launches,
                                 double* bytes) {
    HG_REQUIRE(ctx, "hg_ctx_timing_get: ctx is NULL");
    HG_REQUIRE(kernel_class >= 0 && kernel_class < HG_K_NCLASSES, "hg_ctx_timing_get: bad class");
    HG_TRY(fold_pending(ctx));
    if (ms) *ms = ctx->t_ms[kernel_class];
    if (launches) *launches = ctx->t_count[kernel_class];
    if (bytes) *bytes = ctx->t_bytes[kernel_class];
    return HG_OK;
}

extern "C" int hg_ctx_timing_reset(hg_ctx* ctx) {
    HG_REQUIRE(ctx, "hg_ctx_timing_reset: ctx is NULL");
    HG_TRY(fold_pending(ctx));
    for (int i = 0; i < HG_K_NCLASSES; ++i) {
        ctx->t_ms[i] = 0;
        ctx->t_count[i] = 0;
        ctx->t_bytes[i] = 0;
    }
    return HG_OK;
}

extern "C" int hg_host_register(void* ptr, size_t bytes) {
    HG_REQUIRE(ptr && bytes, "hg_host_register: NULL/empty range");
    HG_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return HG_OK;
}

extern "C" int hg_host_unregister(void* ptr) {
    HG_REQUIRE(ptr, "hg_host_unregister: NULL");
    HG_CUDA(cudaHostUnregister(ptr));
    return HG_OK;
}
