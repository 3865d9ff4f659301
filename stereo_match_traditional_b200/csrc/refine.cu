// Multi-step disparity refinement after the left-right check: iterative region voting, proper interpolation and
// depth-discontinuity adjustment (Mei et al. 2011, section 2.4).
//
// PARITY UNPINNED: the reference implements none of these steps -- only the option fields irv_ts, irv_th and
// do_discontinuity_adjustment exist (CBLSM/adcensus_types.h:58-59,65) and nothing reads them -- but BASELINE.json's
// north_star names them.  The specification is fixed in oracle/port/oracle_port.c (orc_refine_*), which is what these
// kernels are checked against, bit for bit: every phase reads the map as it was when the phase (or the voting iteration)
// began, so the result does not depend on a visiting order and one thread (or warp) per pixel is exact.
// Conventions: disparity float [H][W], +inf = invalid (LeftRightConsistency, PostProcessing.h:93,131); cls uint8 [H][W]:
// 0 valid, 1 occlusion, 2 mismatch (smt_invalid_class_map builds it from LeftRightConsistency's two lists); arms =
// CrossAggregator's {left,right,top,bottom} bytes (cross_aggregator.h:17-20).
#include <cmath>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) cls_scatter_kernel(const int* __restrict__ list, int n, int H, int W, unsigned char value,
                                                          unsigned char* __restrict__ cls, int* __restrict__ bad) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int y = list[2 * k], x = list[2 * k + 1];
    if (y < 0 || y >= H || x < 0 || x >= W) {
        *bad = 1;
        return;
    }
    cls[(size_t)y * W + x] = value;
}

// invalid pixels of the map, any order (the voting iteration is synchronous)
__global__ void __launch_bounds__(256) irv_collect_kernel(const unsigned char* __restrict__ cls, size_t n, int* __restrict__ list,
                                                          int* __restrict__ count) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inv = p < n && cls[p] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, inv);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (inv) list[base + __popc(m & ((1u << lane) - 1u))] = (int)p;
}

// one warp per invalid pixel: histogram of the valid disparities in its cross support region
constexpr int IRV_WARPS = 8;
__global__ void __launch_bounds__(IRV_WARPS * 32) irv_vote_kernel(const float* __restrict__ disp, const uchar4* __restrict__ arms,
                                                                  const int* __restrict__ list, const int* __restrict__ count, int H,
                                                                  int W, int D, int ts, float th, float* __restrict__ decided) {
    extern __shared__ int s_hist_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* hist = s_hist_all + warp * D;
    const int n = *count;
    for (int item = blockIdx.x * IRV_WARPS + warp; item < n; item += gridDim.x * IRV_WARPS) {
        for (int d = lane; d < D; d += 32) hist[d] = 0;
        __syncwarp();
        const int p = list[item];
        const int y = p / W, x = p - y * W;
        const uchar4 a = arms[p];
        int cnt = 0;
        for (int t = -(int)a.z; t <= (int)a.w; ++t) {
            const int yy = y + t;
            if (yy < 0 || yy >= H) continue;
            const uchar4 b = arms[(size_t)yy * W + x];
            for (int s = -(int)b.x + lane; s <= (int)b.y; s += 32) {
                const int xx = x + s;
                if (xx < 0 || xx >= W) continue;
                const float d = disp[(size_t)yy * W + xx];
                if (d == INFINITY || !(d >= 0.f)) continue;
                const int di = (int)lroundf(d);
                if (di < 0 || di >= D) continue;
                atomicAdd(&hist[di], 1);
                ++cnt;
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        // first most frequent disparity: maximise (count, -d)
        int bc = -1, bd = 0;
        for (int d = lane; d < D; d += 32) {
            const int h = hist[d];
            if (h > bc) {
                bc = h;
                bd = d;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o), od = __shfl_xor_sync(0xffffffffu, bd, o);
            if (oc > bc || (oc == bc && od < bd)) {
                bc = oc;
                bd = od;
            }
        }
        if (lane == 0) {
            float v = INFINITY;
            if (cnt > ts && __fdiv_rn((float)bc, (float)cnt) > th) v = (float)bd;
            decided[item] = v;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) irv_apply_kernel(const int* __restrict__ list, const int* __restrict__ count,
                                                        const float* __restrict__ decided, float* __restrict__ disp,
                                                        unsigned char* __restrict__ cls) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count) return;
    const float v = decided[k];
    if (v != INFINITY) {
        disp[list[k]] = v;
        cls[list[k]] = 0;
    }
}

struct Dirs16 {
    float sina[16], cosa[16];
};

__global__ void __launch_bounds__(128) interp_kernel(const float* __restrict__ disp, const unsigned char* __restrict__ cls,
                                                     const unsigned char* __restrict__ bgr, int H, int W, int max_len, Dirs16 dirs,
                                                     float* __restrict__ next) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x;
    const unsigned char c = cls[p];
    if (!c) return;
    float best_d = INFINITY;
    int best_dist = 1 << 30;
    const int b0 = bgr[p * 3], g0 = bgr[p * 3 + 1], r0 = bgr[p * 3 + 2];
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        const float sa = dirs.sina[k], ca = dirs.cosa[k];
        for (int m = 1; m < max_len; ++m) {
            const int yy = (int)lroundf(__fadd_rn((float)y, __fmul_rn((float)m, sa)));
            const int xx = (int)lroundf(__fadd_rn((float)x, __fmul_rn((float)m, ca)));
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) break;
            const size_t q = (size_t)yy * W + xx;
            const float d = disp[q];
            if (d == INFINITY) continue;
            if (c == 1) {
                if (d < best_d) best_d = d;
            } else {
                const int dist = max(max(abs(b0 - (int)bgr[q * 3]), abs(g0 - (int)bgr[q * 3 + 1])), abs(r0 - (int)bgr[q * 3 + 2]));
                if (dist < best_dist) {
                    best_dist = dist;
                    best_d = d;
                }
            }
            break;
        }
    }
    next[p] = best_d;
}

__global__ void __launch_bounds__(256) interp_apply_kernel(const float* __restrict__ next, size_t n, float* __restrict__ disp,
                                                           unsigned char* __restrict__ cls) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !cls[p]) return;
    const float v = next[p];
    if (v != INFINITY) {
        disp[p] = v;
        cls[p] = 0;
    }
}

__global__ void __launch_bounds__(128) discontinuity_kernel(const float* __restrict__ disp, const float* __restrict__ cost, int H, int W,
                                                            int D, float thres, float* __restrict__ next) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x;
    const float d = disp[p];
    float out = d;
    if (y >= 1 && y < H - 1 && x >= 1 && x < W - 1) {
        const float* r0 = disp + p - W;
        const float* r1 = disp + p;
        const float* r2 = disp + p + W;
        const float gx = __fadd_rn(__fadd_rn(__fsub_rn(r0[1], r0[-1]), __fsub_rn(__fmul_rn(2.f, r1[1]), __fmul_rn(2.f, r1[-1]))),
                                   __fsub_rn(r2[1], r2[-1]));
        const float gy = __fadd_rn(__fadd_rn(__fsub_rn(r2[-1], r0[-1]), __fsub_rn(__fmul_rn(2.f, r2[0]), __fmul_rn(2.f, r0[0]))),
                                   __fsub_rn(r2[1], r0[1]));
        if (__fadd_rn(fabsf(gx), fabsf(gy)) > thres && d != INFINITY && d >= 0.f && lroundf(d) < D) {
            const float* c = cost + p * D;
            float cbest = c[lroundf(d)];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float d2 = k == 0 ? r1[-1] : r1[1];
                if (d2 == INFINITY || !(d2 >= 0.f) || lroundf(d2) >= D) continue;
                const float c2 = c[lroundf(d2)];
                if (c2 < cbest) {
                    cbest = c2;
                    out = d2;
                }
            }
        }
    }
    next[p] = out;
}

}  // namespace

smt_status_t k_refine_region_voting(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* arms, int H, int W, int D, int ts, float th,
                                    int iters) {
    StageTimer t(ctx, ST_POST);
    const size_t n = (size_t)H * W;
    SMT_ALLOC(ctx, list, n * sizeof(int));
    SMT_ALLOC(ctx, decided, n * sizeof(float));
    SMT_ALLOC(ctx, count, sizeof(int));
    const size_t smem = (size_t)IRV_WARPS * D * sizeof(int);
    if (smem > 48 * 1024) SMT_CUDA(ctx, cudaFuncSetAttribute(irv_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int it = 0; it < iters; ++it) {
        SMT_CUDA(ctx, cudaMemsetAsync(count.p, 0, sizeof(int), ctx->stream));
        SMT_LAUNCH(ctx, irv_collect_kernel, cdiv((long long)n, 256), 256, 0, cls, n, list.as<int>(), count.as<int>());
        SMT_LAUNCH(ctx, irv_vote_kernel, ctx->sm_count * 8, IRV_WARPS * 32, smem, disp, reinterpret_cast<const uchar4*>(arms), list.as<int>(),
                   count.as<int>(), H, W, D, ts, th, decided.as<float>());
        SMT_LAUNCH(ctx, irv_apply_kernel, cdiv((long long)n, 256), 256, 0, list.as<int>(), count.as<int>(), decided.as<float>(), disp, cls);
    }
    return SMT_OK;
}

smt_status_t k_refine_interpolate(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* bgr, int H, int W, int max_len) {
    StageTimer t(ctx, ST_POST);
    const size_t n = (size_t)H * W;
    SMT_ALLOC(ctx, next, n * sizeof(float));
    Dirs16 dirs;
    for (int k = 0; k < 16; ++k) {  // float32 angle k * (pi / 8), float sin / cos: the tables of oracle.refine_directions()
        const float ang = (float)k * (float)(M_PI / 8);
        dirs.sina[k] = sinf(ang);
        dirs.cosa[k] = cosf(ang);
    }
    dim3 grid(cdiv(W, 128), H);
    SMT_LAUNCH(ctx, interp_kernel, grid, 128, 0, disp, cls, bgr, H, W, max_len, dirs, next.as<float>());
    SMT_LAUNCH(ctx, interp_apply_kernel, cdiv((long long)n, 256), 256, 0, next.as<float>(), n, disp, cls);
    return SMT_OK;
}

smt_status_t k_refine_discontinuity(smt_ctx* ctx, float* disp, const float* cost, int H, int W, int D, float thres) {
    StageTimer t(ctx, ST_POST);
    const size_t n = (size_t)H * W;
    SMT_ALLOC(ctx, next, n * sizeof(float));
    dim3 grid(cdiv(W, 128), H);
    SMT_LAUNCH(ctx, discontinuity_kernel, grid, 128, 0, disp, cost, H, W, D, thres, next.as<float>());
    SMT_CUDA(ctx, cudaMemcpyAsync(disp, next.p, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return SMT_OK;
}

smt_status_t k_invalid_class_map(smt_ctx* ctx, const int32_t* occ, int n_occ, const int32_t* mis, int n_mis, int H, int W, uint8_t* cls) {
    SMT_ALLOC(ctx, bad, sizeof(int));
    SMT_CUDA(ctx, cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
    SMT_CUDA(ctx, cudaMemsetAsync(cls, 0, (size_t)H * W, ctx->stream));
    if (n_occ > 0) SMT_LAUNCH(ctx, cls_scatter_kernel, cdiv(n_occ, 256), 256, 0, occ, n_occ, H, W, (unsigned char)1, cls, bad.as<int>());
    if (n_mis > 0) SMT_LAUNCH(ctx, cls_scatter_kernel, cdiv(n_mis, 256), 256, 0, mis, n_mis, H, W, (unsigned char)2, cls, bad.as<int>());
    SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned[0]) return smt_fail(ctx, SMT_ERR_INVALID_ARG, "smt_invalid_class_map: list entry outside the frame");
    return SMT_OK;
}

extern "C" {

smt_status_t smt_invalid_class_map(smt_ctx* ctx, const int32_t* occ, int32_t n_occ, const int32_t* mis, int32_t n_mis, int32_t rows,
                                   int32_t cols, uint8_t* cls, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, cls && rows > 0 && cols > 0 && n_occ >= 0 && n_mis >= 0 && (occ || !n_occ) && (mis || !n_mis), SMT_ERR_INVALID_ARG,
                "smt_invalid_class_map: bad argument");
    DeviceGuard dev_guard(ctx->device);
    Marshal ms(ctx, space);
    const int32_t* docc = n_occ ? ms.in(occ, 2 * (size_t)n_occ) : nullptr;
    const int32_t* dmis = n_mis ? ms.in(mis, 2 * (size_t)n_mis) : nullptr;
    uint8_t* dcls = ms.out(cls, (size_t)rows * cols);
    SMT_TRY(ms.check());
    SMT_TRY(k_invalid_class_map(ctx, docc, n_occ, dmis, n_mis, rows, cols, dcls));
    return ms.finish();
}

smt_status_t smt_refine_region_voting(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* arms, int32_t rows, int32_t cols, int32_t D,
                                      int32_t ts, float th, int32_t iters, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && cls && arms && rows > 0 && cols > 0 && D > 0 && iters >= 0, SMT_ERR_INVALID_ARG,
                "smt_refine_region_voting: bad argument");
    SMT_REQUIRE(ctx, (size_t)IRV_WARPS * D * sizeof(int) <= 200 * 1024, SMT_ERR_UNSUPPORTED, "smt_refine_region_voting: D too large");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    float* d = ms.out(disp, n, true);
    uint8_t* c = ms.out(cls, n, true);
    const uint8_t* a = ms.in(arms, n * 4);
    SMT_TRY(ms.check());
    SMT_TRY(k_refine_region_voting(ctx, d, c, a, rows, cols, D, ts, th, iters));
    return ms.finish();
}

smt_status_t smt_refine_interpolate(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* bgr, int32_t rows, int32_t cols,
                                    int32_t max_len, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && cls && bgr && rows > 0 && cols > 0 && max_len >= 1, SMT_ERR_INVALID_ARG, "smt_refine_interpolate: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    float* d = ms.out(disp, n, true);
    uint8_t* c = ms.out(cls, n, true);
    const uint8_t* b = ms.in(bgr, n * 3);
    SMT_TRY(ms.check());
    SMT_TRY(k_refine_interpolate(ctx, d, c, b, rows, cols, max_len));
    return ms.finish();
}

smt_status_t smt_refine_discontinuity(smt_ctx* ctx, float* disp, const float* cost, int32_t rows, int32_t cols, int32_t D, float thres,
                                      smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, disp && cost && rows > 0 && cols > 0 && D > 0, SMT_ERR_INVALID_ARG, "smt_refine_discontinuity: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t n = (size_t)rows * cols;
    Marshal ms(ctx, space);
    float* d = ms.out(disp, n, true);
    const float* c = ms.in(cost, n * D);
    SMT_TRY(ms.check());
    SMT_TRY(k_refine_discontinuity(ctx, d, c, rows, cols, D, thres));
    return ms.finish();
}

}  // extern "C"
