// Context, device workspace pool, stage timers and small utility kernels of libsmt_b200.so.
#include <cstdarg>

#include "common.cuh"

static const char* kStageNames[ST_COUNT] = {"census", "cost", "arms", "aggregate", "scanline", "wta", "post", "copy"};

smt_status_t smt_fail(smt_ctx* ctx, smt_status_t st, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return st;
}

void* smt_pool_get(smt_ctx* ctx, size_t bytes) {
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 255) & ~size_t(255);
    int best = -1;
    for (size_t k = 0; k < ctx->pool.size(); ++k) {
        auto& b = ctx->pool[k];
        if (!b.used && b.bytes >= bytes && (best < 0 || b.bytes < ctx->pool[best].bytes)) best = (int)k;
    }
    // do not hand a multi-GB block to a tiny request
    if (best >= 0 && (ctx->pool[best].bytes <= 2 * bytes || ctx->pool[best].bytes <= (1u << 20))) {
        ctx->pool[best].used = true;
        return ctx->pool[best].p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        // free cached blocks and retry once
        for (auto it = ctx->pool.begin(); it != ctx->pool.end();) {
            if (!it->used) {
                cudaFree(it->p);
                it = ctx->pool.erase(it);
            } else {
                ++it;
            }
        }
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    ctx->pool.push_back({p, bytes, true});
    return p;
}

void smt_pool_put(smt_ctx* ctx, void* p) {
    for (auto& b : ctx->pool)
        if (b.p == p) {
            b.used = false;
            return;
        }
}

StageTimer::StageTimer(smt_ctx* c, int stage) : ctx(c) {
    if (!ctx->timing) return;
    smt_ctx::Pending pd;
    pd.stage = stage;
    if (cudaEventCreate(&pd.a) != cudaSuccess || cudaEventCreate(&pd.b) != cudaSuccess) return;
    cudaEventRecord(pd.a, ctx->stream);
    ctx->pending.push_back(pd);
    idx = (int)ctx->pending.size() - 1;
}
StageTimer::~StageTimer() {
    if (idx >= 0) cudaEventRecord(ctx->pending[idx].b, ctx->stream);
}

SideStream::SideStream(smt_ctx* c) : ctx(c), main(c->stream) {
    if (!ctx->side) {
        if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) { ctx->side = nullptr; return; }
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_side[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_side[1], cudaEventDisableTiming);
    }
    cudaEventRecord(ctx->ev_fork, main);
    cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
    ctx->stream = ctx->side;
    ok = true;
}

extern "C" {

int32_t smt_abi_version(void) { return 2; }

const char* smt_status_string(smt_status_t st) {
    switch (st) {
        case SMT_OK: return "ok";
        case SMT_ERR_INVALID_ARG: return "invalid argument";
        case SMT_ERR_UNSUPPORTED: return "unsupported shape or parameter";
        case SMT_ERR_OOM: return "out of device memory";
        case SMT_ERR_CUDA: return "CUDA error";
        case SMT_ERR_INPUT_RANGE: return "image is not integer-valued in 0..255";
    }
    return "unknown status";
}

smt_status_t smt_create(int32_t device, void* stream, smt_ctx** out) {
    if (!out) return SMT_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return SMT_ERR_CUDA;  // no CPU fallback by design
    }
    if (cudaSetDevice(device) != cudaSuccess) return SMT_ERR_CUDA;
    smt_ctx* c = new smt_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return SMT_ERR_CUDA;
        }
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        c->sm_count = prop.multiProcessorCount;
        c->l2_bytes = (size_t)prop.l2CacheSize;
    }
    if (cudaMalloc(&c->d_lut, 320 * sizeof(float)) != cudaSuccess || cudaMalloc(&c->d_flags, 16 * sizeof(int)) != cudaSuccess ||
        cudaMemset(c->d_flags, 0, 16 * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&c->h_pinned, 64 * sizeof(int32_t)) != cudaSuccess) {
        smt_destroy(c);
        return SMT_ERR_OOM;
    }
    *out = c;
    return SMT_OK;
}

void smt_destroy(smt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& pd : ctx->pending) {
        cudaEventDestroy(pd.a);
        cudaEventDestroy(pd.b);
    }
    for (auto& b : ctx->pool) cudaFree(b.p);
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->d_redo) cudaFree(ctx->d_redo);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->side) {
        cudaStreamSynchronize(ctx->side);
        cudaStreamDestroy(ctx->side);
        cudaEventDestroy(ctx->ev_fork);
        cudaEventDestroy(ctx->ev_side[0]);
        cudaEventDestroy(ctx->ev_side[1]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

}  // extern "C"

int* smt_redo_state(smt_ctx* ctx, size_t ntasks) {
    if (!ctx->d_redo || ctx->d_redo_cap < ntasks) {
        if (ctx->d_redo) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->d_redo);
            ctx->d_redo = nullptr;
        }
        const size_t cap = ntasks + ntasks / 2 + 1024;
        if (cudaMalloc(&ctx->d_redo, (FX_ST_LIST + cap) * sizeof(int)) != cudaSuccess) {
            ctx->d_redo_cap = 0;
            return nullptr;
        }
        cudaMemsetAsync(ctx->d_redo, 0, FX_ST_LIST * sizeof(int), ctx->stream);
        ctx->d_redo_cap = cap;
    }
    return ctx->d_redo;
}

extern "C" {

smt_status_t smt_sync(smt_ctx* ctx) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SMT_OK;
}

void* smt_get_stream(const smt_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

const char* smt_last_error_string(const smt_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

uint64_t smt_kernel_launches(const smt_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t smt_num_stages(void) { return ST_COUNT; }
const char* smt_stage_name(int32_t s) { return (s >= 0 && s < ST_COUNT) ? kStageNames[s] : ""; }

smt_status_t smt_stage_timing_enable(smt_ctx* ctx, int32_t enable) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    ctx->timing = enable != 0;
    return SMT_OK;
}

smt_status_t smt_stage_timing_read(smt_ctx* ctx, double* ms, int32_t reset) {
    if (!ctx || !ms) return SMT_ERR_INVALID_ARG;
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& pd : ctx->pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pd.a, pd.b) == cudaSuccess) ctx->stage_ms[pd.stage] += t;
        cudaEventDestroy(pd.a);
        cudaEventDestroy(pd.b);
    }
    ctx->pending.clear();
    for (int s = 0; s < ST_COUNT; ++s) ms[s] = ctx->stage_ms[s];
    if (reset)
        for (int s = 0; s < ST_COUNT; ++s) ctx->stage_ms[s] = 0;
    return SMT_OK;
}

smt_status_t smt_trim(smt_ctx* ctx) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto it = ctx->pool.begin(); it != ctx->pool.end();) {
        if (!it->used) {
            cudaFree(it->p);
            it = ctx->pool.erase(it);
        } else {
            ++it;
        }
    }
    return SMT_OK;
}

smt_status_t smt_device_alloc(smt_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return SMT_ERR_INVALID_ARG;
    DeviceGuard dg(ctx->device);
    SMT_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return SMT_OK;
}
smt_status_t smt_device_free(smt_ctx* ctx, void* p) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    DeviceGuard dg(ctx->device);
    SMT_CUDA(ctx, cudaFree(p));
    return SMT_OK;
}
smt_status_t smt_memcpy_h2d(smt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return SMT_ERR_INVALID_ARG;
    SMT_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SMT_OK;
}
smt_status_t smt_memcpy_d2h(smt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return SMT_ERR_INVALID_ARG;
    SMT_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SMT_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- utility kernels
__global__ void f32_to_u8_kernel(const float* __restrict__ in, size_t n, uint8_t* __restrict__ out, int* __restrict__ bad) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float v = in[k];
    int q = (int)v;
    if (!(v >= 0.f && v <= 255.f) || (float)q != v) {
        *bad = 1;
        q = 0;
    }
    out[k] = (uint8_t)q;
}
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ in, size_t n, float* __restrict__ out) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (float)in[k];
}
// OpenCV's 8-bit BGR2GRAY fixed-point formula (every reference main() calls cvtColor, e.g. AD-CensusV1/main.cpp:19-20)
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, size_t n, uint8_t* __restrict__ gray) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int b = bgr[3 * k], g = bgr[3 * k + 1], r = bgr[3 * k + 2];
    gray[k] = (uint8_t)((1868 * b + 9617 * g + 4899 * r + 8192) >> 14);
}

smt_status_t k_f32_to_u8(smt_ctx* ctx, const float* in, size_t n, uint8_t* out) {
    SMT_ALLOC(ctx, flag, sizeof(int));
    SMT_CUDA(ctx, cudaMemsetAsync(flag.p, 0, sizeof(int), ctx->stream));
    SMT_LAUNCH(ctx, f32_to_u8_kernel, cdiv((long long)n, 256), 256, 0, in, n, out, flag.as<int>());
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned[0])
        return smt_fail(ctx, SMT_ERR_INPUT_RANGE,
                        "float image must hold integer grey values 0..255 (AD-CensusV1/main.cpp:50-53)");
    return SMT_OK;
}
smt_status_t k_u8_to_f32(smt_ctx* ctx, const uint8_t* in, size_t n, float* out) {
    SMT_LAUNCH(ctx, u8_to_f32_kernel, cdiv((long long)n, 256), 256, 0, in, n, out);
    return SMT_OK;
}
smt_status_t k_bgr_to_gray(smt_ctx* ctx, const uint8_t* bgr, size_t n, uint8_t* gray) {
    SMT_LAUNCH(ctx, bgr_to_gray_kernel, cdiv((long long)n, 256), 256, 0, bgr, n, gray);
    return SMT_OK;
}
