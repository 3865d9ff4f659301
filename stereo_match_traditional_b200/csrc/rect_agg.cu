// Grey / colour arm maps with the two-level threshold, rectangle-mean aggregation, and the plain AD volume.
//
// Replaces CrossArmAggregation::{Compute{Left,Right,Top,Button}ArmLength, AggregationVertical, Aggregation}
// (AD-CensusV1/CrossArm.cpp:147-598, :60-145) and CBLSM's ArmLength{L,R,Up,Down}, costAggregationV5, ComputeAD,
// ComputeADRight (CBLSM/CBLSM.h:536-966, :1179-1224, :327-381).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------- AD volume
// AD(i,j,d) = |L(i,j) - R(i,max(j-d,0))| (view 0) or |L(i,min(j+d,W-1)) - R(i,j)| (view 1): the reference copies
// the previous disparity's cost when the match leaves the image, which is the clamped match (CBLSM.h:340-349).
__global__ void __launch_bounds__(256) ad_cost_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, int H, int W,
                                                      int D, int view, float* __restrict__ out) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)H * W * D) return;
    const size_t pix = cell / D;
    const int d = (int)(cell - pix * D);
    const int i = (int)(pix / W), j = (int)(pix - (size_t)i * W);
    int a, b;
    if (view == 0) {
        a = L[pix];
        b = R[(size_t)i * W + max(j - d, 0)];
    } else {
        a = L[(size_t)i * W + min(j + d, W - 1)];
        b = R[pix];
    }
    out[cell] = (float)abs(a - b);
}

// ---------------------------------------------------------------------------------------------- arm maps
// One walk per (pixel, direction) -- see oracle/port/oracle_port.c walk_grey_arm for the sequential statement:
//   k = 1, 2, ...: once k exceeds `sec` the threshold becomes 6 and STAYS 6 (CrossArm.cpp:168-173);
//   stop at the border, at k > maxlen, or when the largest channel difference exceeds the threshold, keeping arm 1
//   if that happens on the first neighbour of a pixel at least two away from the border (:187).
// The sticky threshold is a sequential dependence over pixels in row-major order (and, for CrossArmAggregation, over
// the four passes Left, Right, Top, Button; main.cpp:69-72).  Parallel form: pass A finds the FIRST walk (in that
// order) that reaches k = sec+1 under the initial threshold (atomicMin over a global order index); pass B walks every
// pixel with T(k) = 6 if (order > trigger or k > sec) else tao0.
struct ArmGeom {
    int H, W, ch, sec, maxlen, tao0;
    int quirk_right;  // CrossArm.cpp:265 -- the Right pass runs with col = _row: columns j < H, border at H, stride H
};

__device__ __forceinline__ void arm_setup(const ArmGeom& g, int dir, int idx, int& i, int& j, long long& step, int& room,
                                          bool& minrule, bool& valid, long long& store) {
    const int cols = (dir == 1 && g.quirk_right) ? g.H : g.W;
    i = idx / cols;
    j = idx - i * cols;
    valid = i < g.H;
    store = (long long)i * cols + j;  // the Right pass of CrossArmAggregation stores at i*H + j
    if (dir == 0) {
        step = -1; room = j; minrule = j - 1 >= 1;
    } else if (dir == 1) {
        step = 1; room = cols - 1 - j; minrule = j + 1 < cols - 1;
    } else if (dir == 2) {
        step = -(long long)g.W; room = i; minrule = i - 1 >= 1;
    } else {
        step = g.W; room = g.H - 1 - i; minrule = i + 1 < g.H - 1;
    }
}

__device__ __forceinline__ int chan_diff(const uint8_t* a, const uint8_t* b, int ch) {
    int m = abs((int)a[0] - (int)b[0]);
    if (ch == 3) m = max(m, max(abs((int)a[1] - (int)b[1]), abs((int)a[2] - (int)b[2])));
    return m;
}

// pass A: dirs [dir0, dir0+ndir) share one trigger
__global__ void __launch_bounds__(256) arm_trigger_kernel(const uint8_t* __restrict__ img, ArmGeom g, int dir0, int ndir, int per_dir,
                                                          int* __restrict__ trigger) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ndir * per_dir) return;
    const int dir = dir0 + (int)(t / per_dir), idx = (int)(t % per_dir);
    int i, j, room;
    long long step, store;
    bool minrule, valid;
    arm_setup(g, dir, idx, i, j, step, room, minrule, valid, store);
    if (!valid || room < g.sec) return;  // k = sec+1 is entered only if offset sec is inside the image
    const uint8_t* c0 = img + ((long long)i * g.W + j) * g.ch;
    for (int k = 1; k <= g.sec; ++k) {
        if (chan_diff(c0, c0 + step * k * g.ch, g.ch) > g.tao0) return;
    }
    atomicMin(trigger, (int)t);
}

__global__ void __launch_bounds__(256) arm_walk_kernel(const uint8_t* __restrict__ img, ArmGeom g, int dir0, int ndir, int per_dir,
                                                       const int* __restrict__ trigger, int* __restrict__ out0,
                                                       int* __restrict__ out1, int* __restrict__ out2, int* __restrict__ out3) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)ndir * per_dir) return;
    const int dir = dir0 + (int)(t / per_dir), idx = (int)(t % per_dir);
    int i, j, room;
    long long step, store;
    bool minrule, valid;
    arm_setup(g, dir, idx, i, j, step, room, minrule, valid, store);
    if (!valid) return;
    const bool after = t > (long long)*trigger;
    const uint8_t* c0 = img + ((long long)i * g.W + j) * g.ch;
    int saved = 0, k = 0;
    while (k <= room) {
        saved = k;
        ++k;
        if (k > g.sec && k > g.maxlen) break;
        if (k > room) break;
        const int T = (after || k > g.sec) ? 6 : g.tao0;
        if (chan_diff(c0, c0 + step * k * g.ch, g.ch) > T) {
            if (minrule && saved < 1) saved = 1;
            break;
        }
    }
    int* out = dir == 0 ? out0 : dir == 1 ? out1 : dir == 2 ? out2 : out3;
    if (out) out[store] = saved;
}

// variant 0: CrossArmAggregation (one sticky threshold across all four passes, Right-pass quirk);
// variant 1: CBLSM ArmLength* (each call starts again from tao).
static smt_status_t k_arms_gray(smt_ctx* ctx, const uint8_t* img, int ch, int H, int W, int tao, int maxlen, int sec, int variant,
                                int* aL, int* aR, int* aT, int* aB, int* tao_final_host) {
    StageTimer t(ctx, ST_ARMS);
    ArmGeom g{H, W, ch, sec, maxlen, variant == 1 ? (int)(uint8_t)tao : tao, variant == 0 ? 1 : 0};
    const int per_dir = H * W;  // the quirky Right pass enumerates H*H <= H*W walks; the rest are masked by `valid`
    SMT_ALLOC(ctx, trig, 4 * sizeof(int));
    const int init[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    SMT_CUDA(ctx, cudaMemcpyAsync(trig.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    const size_t npix = (size_t)H * W;
    if (variant == 0 && aR) SMT_CUDA(ctx, cudaMemsetAsync(aR, 0, npix * sizeof(int), ctx->stream));  // entries >= H*H stay 0
    if (variant == 0) {
        const unsigned nb = cdiv(4ll * per_dir, 256);
        SMT_LAUNCH(ctx, arm_trigger_kernel, nb, 256, 0, img, g, 0, 4, per_dir, trig.as<int>());
        SMT_LAUNCH(ctx, arm_walk_kernel, nb, 256, 0, img, g, 0, 4, per_dir, trig.as<int>(), aL, aR, aT, aB);
    } else {
        const unsigned nb = cdiv(per_dir, 256);
        for (int dir = 0; dir < 4; ++dir) {
            SMT_LAUNCH(ctx, arm_trigger_kernel, nb, 256, 0, img, g, dir, 1, per_dir, trig.as<int>() + dir);
            SMT_LAUNCH(ctx, arm_walk_kernel, nb, 256, 0, img, g, dir, 1, per_dir, trig.as<int>() + dir, aL, aR, aT, aB);
        }
    }
    if (tao_final_host) {
        SMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, trig.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *tao_final_host = ctx->h_pinned[0] != 0x7fffffff ? 6 : tao;
    }
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- rectangle aggregation
enum { RECT_FLAT_COLMAJOR = 0, RECT_FLAT_EXCLUSIVE = 1, RECT_INBOUNDS = 2, RECT_CELL_EXCLUSIVE = 3 };

// bounds of the flat addresses a variant touches; used to refuse shapes for which the reference itself reads out of
// bounds (memory-safe iff floor((H*H-1)/W)+34 < H-1 for the shipped Right-pass quirk)
__global__ void __launch_bounds__(256) rect_bounds_kernel(const int* __restrict__ aL, const int* __restrict__ aR,
                                                          const int* __restrict__ aT, const int* __restrict__ aB, int H, int W,
                                                          int variant, int D, long long* __restrict__ minmax) {
    // RECT_CELL_EXCLUSIVE: one arm entry per (pixel, d) cell (costAggregationV4's arm volumes), otherwise one per pixel
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nent = (long long)H * W * (variant == RECT_CELL_EXCLUSIVE ? D : 1);
    if (q >= nent) return;
    const long long p = variant == RECT_CELL_EXCLUSIVE ? q / D : q;
    const int i = (int)(p / W), j = (int)(p % W);
    const int ex = (variant == RECT_FLAT_EXCLUSIVE || variant == RECT_CELL_EXCLUSIVE) ? 1 : 0;
    const int l = aL[q], r = aR[q] - ex, up = aT[q], dn = aB[q] - ex;
    if (r < -l || dn < -up) return;  // empty window
    long long lo, hi;
    if (variant == RECT_INBOUNDS) {
        // every row access must stay inside its own row as well
        lo = (i - up < 0 || j - l < 0) ? -1 : 0;
        hi = (i + dn >= H || j + r >= W) ? (long long)H * W : 0;
    } else {
        lo = (long long)(i - up) * W + j - l;
        hi = (long long)(i + dn) * W + j + r;
    }
    atomicMin(minmax, lo);
    atomicMax(minmax + 1, hi);
}

// reference-order sum: one thread per (pixel, d), d fastest
__global__ void __launch_bounds__(256) rect_exact_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const int* __restrict__ aL, const int* __restrict__ aR,
                                                         const int* __restrict__ aT, const int* __restrict__ aB, int H, int W, int D,
                                                         int variant) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)H * W * D) return;
    const size_t pix = cell / D;
    const int d = (int)(cell - pix * D);
    const size_t ai = variant == RECT_CELL_EXCLUSIVE ? cell : pix;
    const int l = aL[ai], r = aR[ai], up = aT[ai], dn = aB[ai];
    const float* base = in + (long long)pix * D + d;
    float value = 0.f;
    int number = 0;
    if (variant == RECT_FLAT_COLMAJOR) {  // CrossArm.cpp:88-95
        for (int a = -l; a <= r; ++a)
            for (int b = -up; b <= dn; ++b) {
                value = __fadd_rn(value, base[((long long)b * W + a) * D]);
                ++number;
            }
    } else if (variant == RECT_FLAT_EXCLUSIVE || variant == RECT_CELL_EXCLUSIVE) {  // CrossArm.cpp:131-138, CBLSM.h:1159-1166
        for (int b = -up; b < dn; ++b)
            for (int a = -l; a < r; ++a) {
                value = __fadd_rn(value, base[((long long)b * W + a) * D]);
                ++number;
            }
    } else {  // CBLSM.h:1211-1218
        for (int b = -up; b <= dn; ++b)
            for (int a = -l; a <= r; ++a) {
                value = __fadd_rn(value, base[((long long)b * W + a) * D]);
                ++number;
            }
    }
    out[cell] = __fdiv_rn(value, (float)number);
}

// ---- fast mode: summed-area table in double-single, small windows summed directly ----
struct dsp {
    float hi, lo;
};
__device__ __forceinline__ dsp dsp_add(dsp a, dsp b) {
    const float s = __fadd_rn(a.hi, b.hi);
    const float bb = __fsub_rn(s, a.hi);
    const float e = __fadd_rn(__fsub_rn(a.hi, __fsub_rn(s, bb)), __fsub_rn(b.hi, bb));
    const float l = __fadd_rn(__fadd_rn(a.lo, b.lo), e);
    dsp r;
    r.hi = __fadd_rn(s, l);
    r.lo = __fsub_rn(l, __fsub_rn(r.hi, s));
    return r;
}
__device__ __forceinline__ dsp dsp_neg(dsp a) { return dsp{-a.hi, -a.lo}; }

// S[(i+1)][(j+1)][d] = sum_{y<=i, x<=j} in[y][x][d], S has (H+1) x (W+1) x D entries, first row/column zero.
// step 1: running sums along x (one thread per (row, d)); step 2: running sums along y (one thread per (column, d)).
__global__ void __launch_bounds__(256) sat_rows_kernel(const float* __restrict__ in, float2* __restrict__ S, int H, int W, int D) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)H * D) return;
    const int i = (int)(t / D), d = (int)(t - (size_t)i * D);
    dsp acc{0.f, 0.f};
    float2* row = S + ((size_t)(i + 1) * (W + 1)) * D + d;
    row[0] = make_float2(0.f, 0.f);
    const float* src = in + (size_t)i * W * D + d;
    for (int j = 0; j < W; ++j) {
        acc = dsp_add(acc, dsp{src[(size_t)j * D], 0.f});
        row[(size_t)(j + 1) * D] = make_float2(acc.hi, acc.lo);
    }
}
__global__ void __launch_bounds__(256) sat_cols_kernel(float2* __restrict__ S, int H, int W, int D) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)(W + 1) * D) return;
    dsp acc{0.f, 0.f};
    S[t] = make_float2(0.f, 0.f);
    for (int i = 1; i <= H; ++i) {
        float2* p = S + (size_t)i * (W + 1) * D + t;
        const float2 v = *p;
        acc = dsp_add(acc, dsp{v.x, v.y});
        *p = make_float2(acc.hi, acc.lo);
    }
}
__device__ __forceinline__ dsp sat_rect(const float2* __restrict__ S, int W, int D, int d, int i0, int i1, int j0, int j1) {
    // rows [i0, i1), columns [j0, j1)
    auto at = [&](int i, int j) {
        const float2 v = S[((size_t)i * (W + 1) + j) * D + d];
        return dsp{v.x, v.y};
    };
    const dsp top = dsp_add(at(i1, j1), dsp_neg(at(i0, j1)));
    const dsp bot = dsp_add(at(i1, j0), dsp_neg(at(i0, j0)));
    return dsp_add(top, dsp_neg(bot));
}
__global__ void __launch_bounds__(256) rect_sat_kernel(const float* __restrict__ in, const float2* __restrict__ S,
                                                       float* __restrict__ out, const int* __restrict__ aL,
                                                       const int* __restrict__ aR, const int* __restrict__ aT,
                                                       const int* __restrict__ aB, int H, int W, int D, int variant) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)H * W * D) return;
    const size_t pix = cell / D;
    const int d = (int)(cell - pix * D);
    const int i = (int)(pix / W), j = (int)(pix - (size_t)i * W);
    const int ex = variant == RECT_FLAT_EXCLUSIVE ? 1 : 0;
    const int l = aL[pix], r = aR[pix] - ex, up = aT[pix], dn = aB[pix] - ex;
    const int nw = r + l + 1, nh = dn + up + 1;
    const int number = (nw > 0 && nh > 0) ? nw * nh : 0;
    float value = 0.f;
    if (number > 0 && number <= 32) {
        // tiny windows: the reference's own order (exact), and no cancellation against a large table entry
        const float* base = in + (long long)pix * D + d;
        if (variant == RECT_FLAT_COLMAJOR) {
            for (int a = -l; a <= r; ++a)
                for (int b = -up; b <= dn; ++b) value = __fadd_rn(value, base[((long long)b * W + a) * D]);
        } else {
            for (int b = -up; b <= dn; ++b)
                for (int a = -l; a <= r; ++a) value = __fadd_rn(value, base[((long long)b * W + a) * D]);
        }
    } else if (number > 0) {
        const int i0 = i - up, i1 = i + dn + 1, j0 = j - l, j1 = j + r + 1;
        dsp s;
        if (j1 <= W) {
            s = sat_rect(S, W, D, d, i0, i1, j0, j1);
        } else {  // flat index: columns past the row end continue at the start of the next row
            s = dsp_add(sat_rect(S, W, D, d, i0, i1, j0, W), sat_rect(S, W, D, d, i0 + 1, i1 + 1, 0, j1 - W));
        }
        value = __fadd_rn(s.hi, s.lo);
    }
    out[cell] = __fdiv_rn(value, (float)number);
}

static smt_status_t k_aggregate_rect(smt_ctx* ctx, const float* in, float* out, const int* aL, const int* aR, const int* aT,
                                     const int* aB, int H, int W, int D, int variant, smt_sum_mode_t mode) {
    StageTimer t(ctx, ST_AGGREGATE);
    const size_t npix = (size_t)H * W, ncell = npix * D;
    {
        SMT_ALLOC(ctx, mm, 2 * sizeof(long long));
        const long long init[2] = {0, 0};
        SMT_CUDA(ctx, cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
        SMT_LAUNCH(ctx, rect_bounds_kernel, cdiv((long long)(variant == RECT_CELL_EXCLUSIVE ? ncell : npix), 256), 256, 0, aL, aR, aT, aB, H, W,
                   variant, D, mm.as<long long>());
        long long* h = reinterpret_cast<long long*>(ctx->h_pinned);
        SMT_CUDA(ctx, cudaMemcpyAsync(h, mm.p, 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        SMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h[0] < 0 || h[1] >= (long long)npix)
            return smt_fail(ctx, SMT_ERR_UNSUPPORTED,
                            "rectangle aggregation: the arm maps address cells outside the %dx%d slice (the reference reads out "
                            "of bounds for this shape; CrossArm.cpp:92 is memory-safe only if floor((H*H-1)/W)+34 < H-1)", W, H);
    }
    const unsigned nb = cdiv((long long)ncell, 256);
    if (mode == SMT_SUM_REFERENCE_ORDER || variant == RECT_CELL_EXCLUSIVE) {  // per-cell arms: reference order only
        SMT_LAUNCH(ctx, rect_exact_kernel, nb, 256, 0, in, out, aL, aR, aT, aB, H, W, D, variant);
        return SMT_OK;
    }
    SMT_ALLOC(ctx, sat, (size_t)(H + 1) * (W + 1) * D * sizeof(float2));
    SMT_LAUNCH(ctx, sat_rows_kernel, cdiv((long long)H * D, 256), 256, 0, in, sat.as<float2>(), H, W, D);
    SMT_LAUNCH(ctx, sat_cols_kernel, cdiv((long long)(W + 1) * D, 256), 256, 0, sat.as<float2>(), H, W, D);
    SMT_LAUNCH(ctx, rect_sat_kernel, nb, 256, 0, in, sat.as<float2>(), out, aL, aR, aT, aB, H, W, D, variant);
    return SMT_OK;
}

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

smt_status_t smt_ad_cost(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols, int32_t D,
                         int32_t view, float* ad, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, left && right && ad && rows > 0 && cols > 0 && D > 0 && (view == 0 || view == 1), SMT_ERR_INVALID_ARG,
                "smt_ad_cost: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* dl = ms.in(left, npix);
    const uint8_t* dr = ms.in(right, npix);
    float* dout = ms.out(ad, npix * D);
    SMT_TRY(ms.check());
    {
        StageTimer t(ctx, ST_COST);
        SMT_LAUNCH(ctx, ad_cost_kernel, cdiv((long long)npix * D, 256), 256, 0, dl, dr, rows, cols, D, view, dout);
    }
    return ms.finish();
}

smt_status_t smt_arms_gray(smt_ctx* ctx, const uint8_t* img, int32_t channels, int32_t rows, int32_t cols, int32_t tao,
                           int32_t max_length, int32_t sec_length, int32_t variant, int32_t* arm_left, int32_t* arm_right,
                           int32_t* arm_top, int32_t* arm_bottom, int32_t* tao_final, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, img && rows > 0 && cols > 0 && (channels == 1 || channels == 3) && (variant == 0 || variant == 1),
                SMT_ERR_INVALID_ARG, "smt_arms_gray: bad argument");
    SMT_REQUIRE(ctx, sec_length >= 0 && max_length >= sec_length, SMT_ERR_UNSUPPORTED,
                "smt_arms_gray: needs 0 <= sec_length <= max_length");
    SMT_REQUIRE(ctx, variant != 0 || rows <= cols, SMT_ERR_UNSUPPORTED,
                "smt_arms_gray: CrossArmAggregation's Right pass indexes columns < rows (CrossArm.cpp:265); rows > cols is out of "
                "bounds in the reference");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const uint8_t* dimg = ms.in(img, npix * channels);
    int32_t* dL = ms.out(arm_left, npix);
    int32_t* dR = ms.out(arm_right, npix);
    int32_t* dT = ms.out(arm_top, npix);
    int32_t* dB = ms.out(arm_bottom, npix);
    SMT_TRY(ms.check());
    SMT_TRY(k_arms_gray(ctx, dimg, channels, rows, cols, tao, max_length, sec_length, variant, dL, dR, dT, dB, tao_final));
    return ms.finish();
}

smt_status_t smt_aggregate_rect(smt_ctx* ctx, const float* vol_in, const int32_t* arm_left, const int32_t* arm_right,
                                const int32_t* arm_top, const int32_t* arm_bottom, int32_t rows, int32_t cols, int32_t D,
                                int32_t variant, smt_sum_mode_t mode, float* vol_out, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, vol_in && vol_out && arm_left && arm_right && arm_top && arm_bottom && rows > 0 && cols > 0 && D > 0,
                SMT_ERR_INVALID_ARG, "smt_aggregate_rect: bad argument");
    SMT_REQUIRE(ctx, variant >= 0 && variant <= 3 && (mode == SMT_SUM_REFERENCE_ORDER || mode == SMT_SUM_FAST), SMT_ERR_INVALID_ARG,
                "smt_aggregate_rect: unknown variant or sum mode");
    SMT_REQUIRE(ctx, vol_in != vol_out, SMT_ERR_INVALID_ARG, "smt_aggregate_rect: in-place aggregation is not supported");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols;
    Marshal ms(ctx, space);
    const float* din = ms.in(vol_in, npix * D);
    const size_t narm = variant == RECT_CELL_EXCLUSIVE ? npix * D : npix;
    const int32_t* dL = ms.in(arm_left, narm);
    const int32_t* dR = ms.in(arm_right, narm);
    const int32_t* dT = ms.in(arm_top, narm);
    const int32_t* dB = ms.in(arm_bottom, narm);
    float* dout = ms.out(vol_out, npix * D);
    SMT_TRY(ms.check());
    SMT_TRY(k_aggregate_rect(ctx, din, dout, dL, dR, dT, dB, rows, cols, D, variant, mode));
    return ms.finish();
}

// CBLSM's main() as one call (CBLSM/CBLSM.cpp:64-67, 101-104, 133-134, 146-153): arm maps of both images, AD volumes of both
// views, costAggregationV5 on (right, right arms), (left, left arms), then a second pass on the left with the left arms
// and on the right with the LEFT arms again (:150, kept), first-minimum disparities of both views.
smt_status_t smt_pipeline_cblsm(smt_ctx* ctx, const uint8_t* gray_left, const uint8_t* gray_right, int32_t rows, int32_t cols, int32_t D,
                                int32_t tao, int32_t max_length, int32_t sec_length, smt_sum_mode_t mode, float* disp_left,
                                float* disp_right, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, gray_left && gray_right && disp_left && disp_right && rows > 0 && cols > 0 && D > 0, SMT_ERR_INVALID_ARG,
                "smt_pipeline_cblsm: bad argument");
    SMT_REQUIRE(ctx, mode == SMT_SUM_REFERENCE_ORDER || mode == SMT_SUM_FAST, SMT_ERR_INVALID_ARG, "smt_pipeline_cblsm: unknown sum mode");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols, nvol = npix * D;
    Marshal ms(ctx, space);
    const uint8_t* gl = ms.in(gray_left, npix);
    const uint8_t* gr = ms.in(gray_right, npix);
    float* dl = ms.out(disp_left, npix);
    float* dr = ms.out(disp_right, npix);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, arms, 8 * npix * sizeof(int));
    SMT_ALLOC(ctx, volA, nvol * sizeof(float));
    SMT_ALLOC(ctx, volB, nvol * sizeof(float));
    int* aL[4];
    int* aR[4];
    for (int k = 0; k < 4; ++k) {
        aL[k] = arms.as<int>() + (size_t)k * npix;
        aR[k] = arms.as<int>() + (size_t)(4 + k) * npix;
    }
    SMT_TRY(k_arms_gray(ctx, gl, 1, rows, cols, tao, max_length, sec_length, 1, aL[0], aL[1], aL[2], aL[3], nullptr));
    SMT_TRY(k_arms_gray(ctx, gr, 1, rows, cols, tao, max_length, sec_length, 1, aR[0], aR[1], aR[2], aR[3], nullptr));
    // right view: ComputeADRight -> V5 (right arms) -> V5 (LEFT arms) -> WTA
    {
        StageTimer t(ctx, ST_COST);
        SMT_LAUNCH(ctx, ad_cost_kernel, cdiv((long long)npix * D, 256), 256, 0, gl, gr, rows, cols, D, 1, volA.as<float>());
    }
    SMT_TRY(k_aggregate_rect(ctx, volA.as<float>(), volB.as<float>(), aR[0], aR[1], aR[2], aR[3], rows, cols, D, RECT_INBOUNDS, mode));
    SMT_TRY(k_aggregate_rect(ctx, volB.as<float>(), volA.as<float>(), aL[0], aL[1], aL[2], aL[3], rows, cols, D, RECT_INBOUNDS, mode));
    SMT_TRY(k_wta(ctx, volA.as<float>(), rows, cols, D, dr));
    // left view: ComputeAD -> V5 -> V5 (left arms) -> WTA
    {
        StageTimer t(ctx, ST_COST);
        SMT_LAUNCH(ctx, ad_cost_kernel, cdiv((long long)npix * D, 256), 256, 0, gl, gr, rows, cols, D, 0, volA.as<float>());
    }
    SMT_TRY(k_aggregate_rect(ctx, volA.as<float>(), volB.as<float>(), aL[0], aL[1], aL[2], aL[3], rows, cols, D, RECT_INBOUNDS, mode));
    SMT_TRY(k_aggregate_rect(ctx, volB.as<float>(), volA.as<float>(), aL[0], aL[1], aL[2], aL[3], rows, cols, D, RECT_INBOUNDS, mode));
    SMT_TRY(k_wta(ctx, volA.as<float>(), rows, cols, D, dl));
    return ms.finish();
}

}  // extern "C"
