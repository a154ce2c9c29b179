// context.cu -- context lifecycle, memory helpers, timers and error reporting of the C ABI.
#include "internal.h"
#include "mt_checkpoints.h"

#include <cstring>

namespace clumpy {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

// Order-preserving float <-> uint32 encoding so that atomicMin/atomicMax on unsigned ints reduce
// floats.  NaNs never reach the atomics (the kernels only submit values that won a `<` test).
__global__ void minmax_reset_kernel(uint32_t* mm) {
    mm[0] = 0xFFFFFFFFu; // +max
    mm[1] = 0u;          // lowest
}

static float decode_ordered(uint32_t u) {
    uint32_t bits = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

int minmax_reset(clumpy_cuda_ctx* ctx) {
    minmax_reset_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_minmax);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

int minmax_fetch(clumpy_cuda_ctx* ctx, float* minval, float* maxval) {
    CK(cudaMemcpyAsync(ctx->h_minmax, ctx->d_minmax, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // An untouched slot decodes to the reference's initial numeric_limits max / lowest.
    float lo = ctx->h_minmax[0] == 0xFFFFFFFFu ? 3.402823466e+38f : decode_ordered(ctx->h_minmax[0]);
    float hi = ctx->h_minmax[1] == 0u ? -3.402823466e+38f : decode_ordered(ctx->h_minmax[1]);
    if (minval) *minval = lo;
    if (maxval) *maxval = hi;
    return CLUMPY_OK;
}

// The cache remembers the size each buffer was allocated with; a request takes the smallest cached buffer that is
// large enough but not more than 1.5x larger (a 64 MB buffer handed out for a 32 MB request would be missing when
// the next 64 MB request comes, and pinning is what the cache exists to avoid).
void* ctx_pinned_get(clumpy_cuda_ctx* ctx, size_t bytes) {
    int best = -1;
    for (int i = 0; i < (int)ctx->pinned_cache.size(); ++i) {
        const size_t have = ctx->pinned_cache[i].second;
        if (have >= bytes && have <= bytes + bytes / 2 && (best < 0 || have < ctx->pinned_cache[best].second)) best = i;
    }
    void* p = nullptr;
    size_t have = bytes;
    if (best >= 0) {
        p = ctx->pinned_cache[best].first;
        have = ctx->pinned_cache[best].second;
        ctx->pinned_cache.erase(ctx->pinned_cache.begin() + best);
    } else {
        p = clumpy_cuda_host_alloc(bytes);
    }
    if (p) ctx->pinned_out.emplace_back(p, have);
    return p;
}

void ctx_pinned_put(clumpy_cuda_ctx* ctx, void* p, size_t /*bytes_requested*/) {
    if (!p) return;
    size_t have = 0;
    for (size_t i = 0; i < ctx->pinned_out.size(); ++i)
        if (ctx->pinned_out[i].first == p) { have = ctx->pinned_out[i].second; ctx->pinned_out.erase(ctx->pinned_out.begin() + i); break; }
    if (have == 0 || ctx->pinned_cache.size() >= 16) { cudaFreeHost(p); return; }
    ctx->pinned_cache.emplace_back(p, have);
}

} // namespace clumpy

using namespace clumpy;

CLUMPY_API const char* clumpy_cuda_last_error(void) { return g_error.c_str(); }

CLUMPY_API int clumpy_cuda_device_count(int* count) {
    REQUIRE(count, "count is NULL");
    *count = 0;
    CK(cudaGetDeviceCount(count));
    return CLUMPY_OK;
}

static int create_impl(clumpy_cuda_ctx* ctx, int device, void* stream, const cudaDeviceProp& prop);

CLUMPY_API int clumpy_cuda_create(int device, void* stream, clumpy_cuda_ctx** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) {
        set_error("no CUDA device is visible; libclumpy_cuda has no CPU fallback");
        return CLUMPY_ERR_CUDA;
    }
    REQUIRE(device >= 0 && device < ndev, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                  prop.major, prop.minor);
        return CLUMPY_ERR_CUDA;
    }
    clumpy_cuda_ctx* ctx = new (std::nothrow) clumpy_cuda_ctx();
    if (!ctx) { set_error("out of host memory"); return CLUMPY_ERR_NOMEM; }
    const int rc = create_impl(ctx, device, stream, prop);
    if (rc != CLUMPY_OK) { clumpy_cuda_destroy(ctx); return rc; } // releases whatever was created
    *out = ctx;
    return CLUMPY_OK;
}

static int create_impl(clumpy_cuda_ctx* ctx, int device, void* stream, const cudaDeviceProp& prop) {
    ctx->device = device;
    mt_checkpoints_prefetch(); // (background thread, once per process: the reset-offset generator's table)
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, greatest));
    }
    CK(cudaEventCreate(&ctx->t0));
    CK(cudaEventCreate(&ctx->t1));
    CK(cudaMalloc(&ctx->d_minmax, 8));
    CK(cudaMallocHost(&ctx->h_minmax, 8));
    {
        // keep what commands free in the device's default pool instead of returning it to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return CLUMPY_OK;
}

CLUMPY_API void clumpy_cuda_destroy(clumpy_cuda_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (cudaStream_t s : {ctx->copy_stream, ctx->aux_stream})
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->d_minmax) cudaFree(ctx->d_minmax);
    if (ctx->mt_checkpoints_dev) cudaFree(ctx->mt_checkpoints_dev);
    if (ctx->h_minmax) cudaFreeHost(ctx->h_minmax);
    cudaGetLastError();
    for (auto& b : ctx->pinned_cache) cudaFreeHost(b.first);
    delete ctx;
}

CLUMPY_API int clumpy_cuda_sync(clumpy_cuda_ctx* ctx) {
    REQUIRE(ctx, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->aux_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return CLUMPY_OK;
}

CLUMPY_API void* clumpy_cuda_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

CLUMPY_API void clumpy_cuda_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

CLUMPY_API int clumpy_cuda_dev_alloc(clumpy_cuda_ctx* ctx, size_t bytes, void** dptr) {
    REQUIRE(ctx && dptr, "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(dptr, bytes ? bytes : 1));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_dev_free(clumpy_cuda_ctx* ctx, void* dptr) {
    REQUIRE(ctx, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(dptr));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_memcpy_h2d(clumpy_cuda_ctx* ctx, void* dst_dev, const void* src_host,
                                      size_t bytes) {
    REQUIRE(ctx && (bytes == 0 || (dst_dev && src_host)), "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_memcpy_d2h(clumpy_cuda_ctx* ctx, void* dst_host, const void* src_dev,
                                      size_t bytes) {
    REQUIRE(ctx && (bytes == 0 || (dst_host && src_dev)), "null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_timer_start(clumpy_cuda_ctx* ctx) {
    REQUIRE(ctx, "ctx is NULL");
    CK(cudaEventRecord(ctx->t0, ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_timer_stop(clumpy_cuda_ctx* ctx, float* elapsed_ms) {
    REQUIRE(ctx && elapsed_ms, "null argument");
    CK(cudaEventRecord(ctx->t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(elapsed_ms, ctx->t0, ctx->t1));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_launch_count(clumpy_cuda_ctx* ctx, uint64_t* launches) {
    REQUIRE(ctx && launches, "null argument");
    *launches = ctx->launches;
    return CLUMPY_OK;
}
