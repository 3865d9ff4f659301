// Cross-based cost aggregation.
//
// Replaces CrossAggregator::{BuildArms, FindHorizontalArm, FindVerticalArm, ComputeSupPixelCount, Aggregate,
// AggregateInArms} (CBLSM/cross_aggregator.cpp:76-118, :135-394).
//
// Arms and support counts are integer work on the image only (one kernel each).  Aggregation is, per iteration and
// per disparity slice, two 1-D passes: sum along one arm pair, then along the orthogonal pair, divide by the support
// count; the order alternates per iteration (:102,:116).  The reference gathers one strided slice at a time; here a
// pass handles all D slices of a pixel together (D innermost => every tap is a coalesced row of the volume).
//
// Two summation modes:
//   SMT_SUM_REFERENCE_ORDER  each cell adds its taps sequentially in the reference's order (t ascending, starting
//                            from 0.0f, :358-383) -> bit-identical to the reference.
//   SMT_SUM_FAST             see agg_line_kernel below.
#include <cstdlib>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------- arms
__device__ __forceinline__ int color_dist(int b0, int g0, int r0, int b1, int g1, int r1) {
    return max(abs(r0 - r1), max(abs(g0 - g1), abs(b0 - b1)));  // cross_aggregator.h:78-80
}

// walks one direction from (x,y); step is the pixel stride of that direction, room the number of in-image pixels
__device__ __forceinline__ int walk_arm(const uint8_t* __restrict__ p0, long long step, int room, int Lmax, int L2, int t1,
                                        int t2) {
    const int b0 = p0[0], g0 = p0[1], r0 = p0[2];
    int bl = b0, gl = g0, rl = r0;
    int arm = 0;
    const uint8_t* p = p0 + step;
    for (int n = 0; n < Lmax; ++n) {
        if (n >= room) break;  // border (:154-163)
        const int b = p[0], g = p[1], r = p[2];
        const int d1 = color_dist(b, g, r, b0, g0, r0);
        if (d1 >= t1) break;                                  // :169-172
        if (n > 0 && color_dist(b, g, r, bl, gl, rl) >= t1) break;  // :175-180
        if (n + 1 > L2 && d1 >= t2) break;                    // :183-187
        ++arm;
        bl = b;
        gl = g;
        rl = r;
        p += step;
    }
    return arm;
}

__global__ void __launch_bounds__(128) cross_arms_kernel(const uint8_t* __restrict__ bgr, int H, int W, int L1, int L2, int t1,
                                                         int t2, uchar4* __restrict__ arms) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const int Lmax = min(L1, 255);  // MAX_ARM_LENGTH (cross_aggregator.h:24)
    const uint8_t* p0 = bgr + ((size_t)y * W + x) * 3;
    uchar4 a;
    a.x = (unsigned char)walk_arm(p0, -3, x, Lmax, L2, t1, t2);
    a.y = (unsigned char)walk_arm(p0, 3, W - 1 - x, Lmax, L2, t1, t2);
    a.z = (unsigned char)walk_arm(p0, -3ll * W, y, Lmax, L2, t1, t2);
    a.w = (unsigned char)walk_arm(p0, 3ll * W, H - 1 - y, Lmax, L2, t1, t2);
    arms[(size_t)y * W + x] = a;
}

// Packed variant for thresholds 1..128 (the reference's 20 and 6): one 32-bit word per pixel (B, G, R, 0), the three
// channel differences in one VABSDIFF4, "some byte >= t" as the carry-free SWAR test ((d + (128 - t)*0x01010101) | d) & 0x80808080
// (bytes of d + (128 - t) cannot carry: 255 + 127 < 512 only matters per byte when d >= 128, and then the `| d` term
// already has the top bit set -- so the sum is formed on the low 7 bits only).
__global__ void __launch_bounds__(256) pack_bgr_kernel(const uint8_t* __restrict__ bgr, size_t n, uint32_t* __restrict__ out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    out[k] = (uint32_t)bgr[3 * k] | ((uint32_t)bgr[3 * k + 1] << 8) | ((uint32_t)bgr[3 * k + 2] << 16);
}

__device__ __forceinline__ bool any_byte_ge(unsigned d, unsigned bias) {  // bias = (128 - t) * 0x01010101, 1 <= t <= 128
    return ((((d & 0x7f7f7f7fu) + bias) | d) & 0x80808080u) != 0u;
}

__device__ __forceinline__ int walk_arm_packed(const uint32_t* __restrict__ p0, long long step, int room, int Lmax, int L2,
                                               unsigned bias1, unsigned bias2) {
    const unsigned c0 = *p0;
    unsigned cl = c0;
    int arm = 0;
    const uint32_t* p = p0 + step;
    const int lim = min(Lmax, room);  // border (:154-163)
    for (int n = 0; n < lim; ++n) {
        const unsigned c = *p;
        const unsigned d0 = __vabsdiffu4(c, c0);
        if (any_byte_ge(d0, bias1)) break;                               // :169-172
        if (n > 0 && any_byte_ge(__vabsdiffu4(c, cl), bias1)) break;     // :175-180
        if (n + 1 > L2 && any_byte_ge(d0, bias2)) break;                 // :183-187
        ++arm;
        cl = c;
        p += step;
    }
    return arm;
}

__global__ void __launch_bounds__(128) cross_arms_packed_kernel(const uint32_t* __restrict__ pix, int H, int W, int L1, int L2,
                                                                unsigned bias1, unsigned bias2, uchar4* __restrict__ arms) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const int Lmax = min(L1, 255);
    const uint32_t* p0 = pix + (size_t)y * W + x;
    uchar4 a;
    a.x = (unsigned char)walk_arm_packed(p0, -1, x, Lmax, L2, bias1, bias2);
    a.y = (unsigned char)walk_arm_packed(p0, 1, W - 1 - x, Lmax, L2, bias1, bias2);
    a.z = (unsigned char)walk_arm_packed(p0, -(long long)W, y, Lmax, L2, bias1, bias2);
    a.w = (unsigned char)walk_arm_packed(p0, (long long)W, H - 1 - y, Lmax, L2, bias1, bias2);
    arms[(size_t)y * W + x] = a;
}

// sup_hfirst(y,x) = sum_{t=-top..bottom} (left+right+1)(y+t,x);  sup_vfirst(y,x) = sum_{t=-left..right} (top+bottom+1)(y,x+t)
// (cross_aggregator.cpp:271-325; uint16 storage as in the reference).
__global__ void __launch_bounds__(128) support_count_kernel(const uchar4* __restrict__ arms, int H, int W,
                                                            uint16_t* __restrict__ sup0, uint16_t* __restrict__ sup1) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uchar4 a = arms[(size_t)y * W + x];
    int c0 = 0, c1 = 0;
    for (int t = -(int)a.z; t <= (int)a.w; ++t) {
        const uchar4 b = arms[(size_t)(y + t) * W + x];
        c0 += (int)(uint16_t)(b.x + b.y + 1);
    }
    for (int t = -(int)a.x; t <= (int)a.y; ++t) {
        const uchar4 b = arms[(size_t)y * W + x + t];
        c1 += (int)(uint16_t)(b.z + b.w + 1);
    }
    if (sup0) sup0[(size_t)y * W + x] = (uint16_t)c0;
    if (sup1) sup1[(size_t)y * W + x] = (uint16_t)c1;
}

// scratch (optional): rows*cols uint32 owned by the caller -- needed when the call runs on a stream other than the one
// later pool allocations are used on (the pool recycles a block as soon as the host-side owner lets go of it)
smt_status_t k_cross_arms(smt_ctx* ctx, const uint8_t* bgr, int H, int W, int L1, int L2, int t1, int t2, uint8_t* arms,
                          uint16_t* sup0, uint16_t* sup1, uint32_t* scratch) {
    StageTimer t(ctx, ST_ARMS);
    dim3 grid(cdiv(W, 128), H);
    if (t1 >= 1 && t1 <= 128 && t2 >= 1 && t2 <= 128) {
        const size_t npix = (size_t)H * W;
        SMT_ALLOC(ctx, own, scratch ? 0 : npix * sizeof(uint32_t));
        uint32_t* packed = scratch ? scratch : own.as<uint32_t>();
        SMT_LAUNCH(ctx, pack_bgr_kernel, cdiv((long long)npix, 256), 256, 0, bgr, npix, packed);
        SMT_LAUNCH(ctx, cross_arms_packed_kernel, grid, 128, 0, packed, H, W, L1, L2, (unsigned)(128 - t1) * 0x01010101u,
                   (unsigned)(128 - t2) * 0x01010101u, reinterpret_cast<uchar4*>(arms));
    } else {
        SMT_LAUNCH(ctx, cross_arms_kernel, grid, 128, 0, bgr, H, W, L1, L2, t1, t2, reinterpret_cast<uchar4*>(arms));
    }
    if (sup0 || sup1)
        SMT_LAUNCH(ctx, support_count_kernel, grid, 128, 0, reinterpret_cast<const uchar4*>(arms), H, W, sup0, sup1);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- reference-order passes
// One thread per (pixel, d), d fastest.  HORIZ: taps (y, x+t), t=-left..right; else (y+t, x), t=-top..bottom.
// sup != nullptr => second pass of an iteration: divide by the support count (float / int, IEEE division, :389).
template <bool HORIZ>
__global__ void __launch_bounds__(256) agg_pass_exact_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                             const uchar4* __restrict__ arms, const uint16_t* __restrict__ sup,
                                                             int H, int W, int D, int st_lo, int st_hi) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t ncell = (size_t)H * W * D;
    if (cell >= ncell) return;
    const size_t pix = cell / D;
    const int d = (int)(cell - pix * D);
    const uchar4 a = arms[pix];
    // arms stop at the image border; the clamps only act on the halo rows of a row band, whose results are never stored
    const int pos = HORIZ ? (int)(pix % W) : (int)(pix / W), len = HORIZ ? W : H;
    if (pos < st_lo || pos >= st_hi) return;
    const int lo = -min(HORIZ ? (int)a.x : (int)a.z, pos);
    const int hi = min(HORIZ ? (int)a.y : (int)a.w, len - 1 - pos);
    const long long step = HORIZ ? (long long)D : (long long)W * D;
    const float* p = in + (long long)pix * D + d + lo * step;
    float acc = 0.0f;
    for (int t = lo; t <= hi; ++t) {
        acc = __fadd_rn(acc, *p);
        p += step;
    }
    if (sup) acc = __fdiv_rn(acc, (float)(int)sup[pix]);
    out[cell] = acc;
}

// Four disparities AND four consecutive positions of the line per thread: the windows of neighbouring positions overlap
// almost completely, so each tap row is loaded once for the four outputs and added (in ascending tap order, exactly the
// reference's order for every output) to the ones whose window contains it.  A quarter of the L1 traffic again.
#ifndef SMT_EXACT_NP
#define SMT_EXACT_NP 4
#endif
constexpr int EXACT_NP = SMT_EXACT_NP;  // consecutive positions per thread
template <bool HORIZ>
__global__ void __launch_bounds__(256) agg_pass_exact44_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                               const uchar4* __restrict__ arms, const uint16_t* __restrict__ sup,
                                                               int H, int W, int D4, int st_lo, int st_hi) {
    const int len = HORIZ ? W : H, nlines = HORIZ ? H : W;
    const int ngrp = (len + EXACT_NP - 1) / EXACT_NP;
    const size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (size_t)nlines * ngrp * D4) return;
    const int d4 = (int)(item % D4);
    const size_t lg = item / D4;
    const int grp = (int)(lg % ngrp), line = (int)(lg / ngrp);
    const int p0 = grp * EXACT_NP;
    const long long pstride = HORIZ ? 1 : (long long)W;                       // pixel stride along the line
    const long long pbase = HORIZ ? (long long)line * W : (long long)line;   // pixel of position 0
    int s[EXACT_NP], e[EXACT_NP];
    int tmin = len, tmax = -1;
#pragma unroll
    for (int i = 0; i < EXACT_NP; ++i) {
        const int pos = p0 + i;
        s[i] = 1;
        e[i] = 0;  // empty window for positions beyond the line or outside the stored range
        if (pos < len && pos >= st_lo && pos < st_hi) {
            const uchar4 a = arms[pbase + pstride * pos];
            s[i] = max(pos - (HORIZ ? (int)a.x : (int)a.z), 0);  // clamps: see agg_pass_exact_kernel
            e[i] = min(pos + (HORIZ ? (int)a.y : (int)a.w), len - 1);
            tmin = min(tmin, s[i]);
            tmax = max(tmax, e[i]);
        }
    }
    float4 acc[EXACT_NP];
#pragma unroll
    for (int i = 0; i < EXACT_NP; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* p = in + (pbase + pstride * tmin) * D4 + d4;
    const long long step = pstride * D4;
    for (int t = tmin; t <= tmax; ++t) {
        const float4 v = *p;
#pragma unroll
        for (int i = 0; i < EXACT_NP; ++i)
            if (t >= s[i] && t <= e[i]) {
                acc[i].x = __fadd_rn(acc[i].x, v.x);
                acc[i].y = __fadd_rn(acc[i].y, v.y);
                acc[i].z = __fadd_rn(acc[i].z, v.z);
                acc[i].w = __fadd_rn(acc[i].w, v.w);
            }
        p += step;
    }
#pragma unroll
    for (int i = 0; i < EXACT_NP; ++i) {
        const int pos = p0 + i;
        if (pos < len && pos >= st_lo && pos < st_hi) {
            const long long pix = pbase + pstride * pos;
            float4 r = acc[i];
            if (sup) {
                const float c = (float)(int)sup[pix];
                r.x = __fdiv_rn(r.x, c);
                r.y = __fdiv_rn(r.y, c);
                r.z = __fdiv_rn(r.z, c);
                r.w = __fdiv_rn(r.w, c);
            }
            out[pix * D4 + d4] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------- fast line passes
// (implemented in cross_agg_fast.cu)
smt_status_t k_aggregate_cross_fast(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms, const uint16_t* sup0,
                                    const uint16_t* sup1, int H, int W, int D, int iters, float vmax, float** result, float* wta_disp);

// one reference-order pass (taps added sequentially in ascending order, starting from 0.0f)
smt_status_t k_agg_exact_pass(smt_ctx* ctx, const float* src, float* dst, const uint8_t* arms8, const uint16_t* s, int H, int W, int D,
                              int horiz, int st_lo, int st_hi) {
    const uchar4* arms = reinterpret_cast<const uchar4*>(arms8);
    if (st_hi <= 0) {
        st_lo = 0;
        st_hi = horiz ? W : H;
    }
    const size_t ncell = (size_t)H * W * D;
    const unsigned grid = cdiv((long long)ncell, 256);
    const bool vec4 = D % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    if (vec4) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        {
            const long long items = (long long)(horiz ? H : W) * (((horiz ? W : H) + EXACT_NP - 1) / EXACT_NP) * (D / 4);
            if (horiz)
                SMT_LAUNCH(ctx, agg_pass_exact44_kernel<true>, cdiv(items, 256), 256, 0, s4, d4, arms, s, H, W, D / 4, st_lo, st_hi);
            else
                SMT_LAUNCH(ctx, agg_pass_exact44_kernel<false>, cdiv(items, 256), 256, 0, s4, d4, arms, s, H, W, D / 4, st_lo, st_hi);
        }
    } else {
        if (horiz)
            SMT_LAUNCH(ctx, agg_pass_exact_kernel<true>, grid, 256, 0, src, dst, arms, s, H, W, D, st_lo, st_hi);
        else
            SMT_LAUNCH(ctx, agg_pass_exact_kernel<false>, grid, 256, 0, src, dst, arms, s, H, W, D, st_lo, st_hi);
    }
    return SMT_OK;
}

smt_status_t k_aggregate_cross(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms8, const uint16_t* sup0,
                               const uint16_t* sup1, int H, int W, int D, int iters, smt_sum_mode_t mode, float vmax, float** result,
                               float* wta_disp) {
    if (mode == SMT_SUM_FAST) return k_aggregate_cross_fast(ctx, vol, tmp, arms8, sup0, sup1, H, W, D, iters, vmax, result, wta_disp);
    auto pass = [&](bool horiz, const float* src, float* dst, const uint16_t* s) -> smt_status_t {
        return k_agg_exact_pass(ctx, src, dst, arms8, s, H, W, D, horiz ? 1 : 0);
    };
    bool hfirst = true;
    float* cur = vol;
    float* oth = tmp;
    {
        StageTimer t(ctx, ST_AGGREGATE);
        for (int k = 0; k < iters; ++k) {
            SMT_TRY(pass(hfirst, cur, oth, nullptr));
            SMT_TRY(pass(!hfirst, oth, cur, hfirst ? sup0 : sup1));
            hfirst = !hfirst;
        }
    }
    *result = cur;  // every iteration ends back in `vol`
    if (wta_disp) SMT_TRY(k_wta(ctx, cur, H, W, D, wta_disp));  // its own stage (ST_WTA)
    return SMT_OK;
}

// min / max of a volume (order-preserving uint keys + atomics); used to pick the compensated-prefix variant
__device__ __forceinline__ unsigned fkey(float f) {
    unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ v, size_t n, unsigned* __restrict__ out) {
    unsigned lo = 0xffffffffu, hi = 0u;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const float x = v[k];
        const unsigned key = (x == x) ? fkey(x) : 0xffffffffu;  // NaN poisons the max
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
static smt_status_t volume_nonneg_max(smt_ctx* ctx, const float* vol, size_t n, float* vmax) {
    SMT_ALLOC(ctx, mm, 2 * sizeof(unsigned));
    const unsigned init[2] = {0xffffffffu, 0u};
    SMT_CUDA(ctx, cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    SMT_LAUNCH(ctx, minmax_kernel, ctx->sm_count * 8, 256, 0, vol, n, mm.as<unsigned>());
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, mm.p, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const unsigned klo = (unsigned)ctx->h_pinned[0], khi = (unsigned)ctx->h_pinned[1];
    auto unkey = [](unsigned k) {
        unsigned u = k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
        float f;
        memcpy(&f, &u, 4);
        return f;
    };
    const float lo = unkey(klo), hi = unkey(khi);
    *vmax = (khi != 0xffffffffu && lo >= 0.f && hi < 1e30f) ? (hi > 0.f ? hi : 1.f) : -1.f;
    return SMT_OK;
}

// longest arm in a {left,right,top,bottom} map (host value; synchronises the stream)
__global__ void __launch_bounds__(256) arms_max_kernel(const unsigned* __restrict__ arms, size_t npix, int* __restrict__ out) {
    unsigned m = 0u;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < npix; k += (size_t)gridDim.x * blockDim.x) m = __vmaxu4(m, arms[k]);
    m = max(max(m & 0xffu, (m >> 8) & 0xffu), max((m >> 16) & 0xffu, m >> 24));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, (int)m);
}
smt_status_t k_arms_max(smt_ctx* ctx, const uint8_t* arms8, size_t npix, int* longest) {
    int* slot = ctx->d_flags + 1;
    SMT_CUDA(ctx, cudaMemsetAsync(slot, 0, sizeof(int), ctx->stream));
    SMT_LAUNCH(ctx, arms_max_kernel, ctx->sm_count * 4, 256, 0, reinterpret_cast<const unsigned*>(arms8), npix, slot);
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *longest = ctx->h_pinned[0];
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

smt_status_t smt_cross_arms(smt_ctx* ctx, const uint8_t* bgr, int32_t rows, int32_t cols, int32_t L1, int32_t L2, int32_t t1,
                            int32_t t2, uint8_t* arms_out, uint16_t* sup_hfirst, uint16_t* sup_vfirst, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, bgr && rows > 0 && cols > 0, SMT_ERR_INVALID_ARG, "smt_cross_arms: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* dbgr = ms.in(bgr, npix * 3);
    uint8_t* darms = ms.out(arms_out, npix * 4);
    uint16_t* ds0 = ms.out(sup_hfirst, npix);
    uint16_t* ds1 = ms.out(sup_vfirst, npix);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, tmp_arms, darms ? 0 : npix * 4);
    if (!darms) darms = tmp_arms.as<uint8_t>();
    SMT_TRY(k_cross_arms(ctx, dbgr, rows, cols, L1, L2, t1, t2, darms, ds0, ds1));
    return ms.finish();
}

smt_status_t smt_aggregate_cross(smt_ctx* ctx, const uint8_t* bgr, const float* cost_init, int32_t rows, int32_t cols, int32_t D,
                                 int32_t L1, int32_t L2, int32_t t1, int32_t t2, int32_t iters, smt_sum_mode_t mode,
                                 float* cost_out, uint8_t* arms_out, uint16_t* sup_hfirst, uint16_t* sup_vfirst,
                                 smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, bgr && cost_init && cost_out && rows > 0 && cols > 0 && D > 0 && iters >= 0, SMT_ERR_INVALID_ARG,
                "smt_aggregate_cross: bad argument");
    SMT_REQUIRE(ctx, mode == SMT_SUM_REFERENCE_ORDER || mode == SMT_SUM_FAST, SMT_ERR_INVALID_ARG,
                "smt_aggregate_cross: unknown sum mode");
    SMT_REQUIRE(ctx, mode != SMT_SUM_FAST || L1 <= 34, SMT_ERR_UNSUPPORTED,
                "smt_aggregate_cross: SMT_SUM_FAST tiles assume arms <= 34 (cross_L1 <= 34, the reference default)");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols, nvol = npix * D;
    Marshal ms(ctx, space);
    const uint8_t* dbgr = ms.in(bgr, npix * 3);
    const float* dinit = ms.in(cost_init, nvol);
    float* dout = ms.out(cost_out, nvol);
    uint8_t* darms = ms.out(arms_out, npix * 4);
    uint16_t* ds0 = ms.out(sup_hfirst, npix);
    uint16_t* ds1 = ms.out(sup_vfirst, npix);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, aux, npix * 8);  // arms (4 B) + two uint16 maps when the caller does not want them
    if (!darms) darms = aux.as<uint8_t>();
    if (!ds0) ds0 = reinterpret_cast<uint16_t*>(aux.as<uint8_t>() + npix * 4);
    if (!ds1) ds1 = reinterpret_cast<uint16_t*>(aux.as<uint8_t>() + npix * 6);
    SMT_ALLOC(ctx, tmp, nvol * sizeof(float));
    SMT_TRY(k_cross_arms(ctx, dbgr, rows, cols, L1, L2, t1, t2, darms, ds0, ds1));
    // the reference memcpy's cost_init into its own volume first (:108)
    if (dout != dinit)
        SMT_CUDA(ctx, cudaMemcpyAsync(dout, dinit, nvol * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    float* res = nullptr;
    float vmax = -1.f;
    if (mode == SMT_SUM_FAST) SMT_TRY(volume_nonneg_max(ctx, dout, nvol, &vmax));
    SMT_TRY(k_aggregate_cross(ctx, dout, tmp.as<float>(), darms, ds0, ds1, rows, cols, D, iters, mode, vmax, &res));
    if (res != dout)
        SMT_CUDA(ctx, cudaMemcpyAsync(dout, res, nvol * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return ms.finish();
}

}  // extern "C"
