// Image-sized pre-processing and display export around the hot path: what the reference's
// drivers do with OpenCV before and after calling the matchers.
//   smt_bgr_to_gray        cvtColor(BGR2GRAY), 8-bit fixed point                      (AD-CensusV1/main.cpp:19-20 and every main)
//   smt_replicate_pad      copyMakeBorder(..., BORDER_REPLICATE)                      (SADmain.cpp:47-48, ASWeight.cpp:54-57)
//   smt_u8_to_f32          the uchar -> float image copies                            (AD-CensusV1/main.cpp:46-55)
//   smt_transform_to_show  TransformToShow: float disparity -> 8-bit                  (AD-CensusV1/PostProcessing.h:137-154)
//   smt_normalize_minmax_u8  normalize(src, dst, 0, 255, NORM_MINMAX) + convertTo(CV_8U)  (main.cpp:106-111, SADmain.cpp:87-95)
#include "common.cuh"

__global__ void __launch_bounds__(256) replicate_pad_kernel(const uint8_t* __restrict__ in, int H, int W, int pad, uint8_t* __restrict__ out) {
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)Hp * Wp) return;
    const int i = (int)(p / Wp), j = (int)(p - (size_t)i * Wp);
    const int y = min(max(i - pad, 0), H - 1), x = min(max(j - pad, 0), W - 1);
    out[p] = in[(size_t)y * W + x];
}

// PostProcessing.h:143-151: `const int v = disp[..]`; 0 stays 0, otherwise static_cast<uchar>(v).  A float outside the
// int range (+inf from the LR check, 2147483648.f from RemoveSpeckles) converts to INT_MIN on x86, whose low byte is 0.
__global__ void __launch_bounds__(256) transform_to_show_kernel(const float* __restrict__ disp, size_t n, uint8_t* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float f = disp[p];
    int v = 0;
    if (f > -2147483648.f && f < 2147483648.f) v = (int)f;  // NaN fails both comparisons -> 0 (x86: INT_MIN -> 0)
    out[p] = (uint8_t)(v & 0xff);
}

__device__ __forceinline__ unsigned fkey2(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__global__ void __launch_bounds__(256) minmax_f32_kernel(const float* __restrict__ v, size_t n, unsigned* __restrict__ out) {
    unsigned lo = 0xffffffffu, hi = 0u;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const unsigned key = fkey2(v[k]);
        lo = min(lo, key);
        hi = max(hi, key);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}
// OpenCV: scale = 255/(max-min) (double), dst = saturate_cast<uchar>(cvRound(src*scale + shift)), shift = -min*scale;
// a constant image gives scale 0 -> all `alpha` (0).  cvRound is round-half-to-even (lrint).
__global__ void __launch_bounds__(256) normalize_u8_kernel(const float* __restrict__ v, size_t n, const unsigned* __restrict__ mm,
                                                           uint8_t* __restrict__ out) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    auto unkey = [](unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); };
    const double lo = (double)unkey(mm[0]), hi = (double)unkey(mm[1]);
    const double scale = (hi - lo) > 2.220446049250313e-16 ? 255.0 / (hi - lo) : 0.0;
    const double shift = 0.0 - lo * scale;
    // convertTo computes in float for CV_32F -> CV_8U: dst = saturate(round(src * (float)scale + (float)shift))
    const float r = __fmaf_rn(v[p], (float)scale, (float)shift);
    int q = __float2int_rn(r);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    out[p] = (uint8_t)q;
}

extern "C" {

smt_status_t smt_bgr_to_gray(smt_ctx* ctx, const uint8_t* bgr, int32_t rows, int32_t cols, uint8_t* gray, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, bgr && gray && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_bgr_to_gray: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* d = ms.in(bgr, n * 3);
    uint8_t* o = ms.out(gray, n);
    SMT_TRY(ms.check());
    SMT_TRY(k_bgr_to_gray(ctx, d, n, o));
    return ms.finish();
}

smt_status_t smt_replicate_pad(smt_ctx* ctx, const uint8_t* img, int32_t rows, int32_t cols, int32_t pad, uint8_t* out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, img && out && rows > 0 && cols > 0 && pad >= 0, SMT_ERR_INVALID_ARG, "smt_replicate_pad: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols, np = (size_t)(rows + 2 * pad) * (cols + 2 * pad);
    Marshal ms(ctx, space);
    const uint8_t* d = ms.in(img, n);
    uint8_t* o = ms.out(out, np);
    SMT_TRY(ms.check());
    SMT_LAUNCH(ctx, replicate_pad_kernel, cdiv((long long)np, 256), 256, 0, d, rows, cols, pad, o);
    return ms.finish();
}

smt_status_t smt_u8_to_f32(smt_ctx* ctx, const uint8_t* img, size_t count, float* out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, img && out && count > 0, SMT_ERR_INVALID_ARG, "smt_u8_to_f32: bad argument");
    DeviceGuard dev_guard(ctx->device);
    Marshal ms(ctx, space);
    const uint8_t* d = ms.in(img, count);
    float* o = ms.out(out, count);
    SMT_TRY(ms.check());
    SMT_TRY(k_u8_to_f32(ctx, d, count, o));
    return ms.finish();
}

smt_status_t smt_transform_to_show(smt_ctx* ctx, const float* disp, int32_t rows, int32_t cols, uint8_t* out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_transform_to_show: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* d = ms.in(disp, n);
    uint8_t* o = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_LAUNCH(ctx, transform_to_show_kernel, cdiv((long long)n, 256), 256, 0, d, n, o);
    return ms.finish();
}

smt_status_t smt_normalize_minmax_u8(smt_ctx* ctx, const float* src, int32_t rows, int32_t cols, uint8_t* out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, src && out && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_normalize_minmax_u8: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* d = ms.in(src, n);
    uint8_t* o = ms.out(out, n);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, mm, 2 * sizeof(unsigned));
    const unsigned init[2] = {0xffffffffu, 0u};
    SMT_CUDA(ctx, cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    SMT_LAUNCH(ctx, minmax_f32_kernel, ctx->sm_count * 4, 256, 0, d, n, mm.as<unsigned>());
    SMT_LAUNCH(ctx, normalize_u8_kernel, cdiv((long long)n, 256), 256, 0, d, n, mm.as<unsigned>(), o);
    return ms.finish();
}

}  // extern "C"
