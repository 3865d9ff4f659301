// 4-path scanline optimisation.
//
// Replaces ScanlineOptimizer::{ScanLine, ScanLineLeftRight, ScanLineUpDown, WTA} (AD-CensusV1/ScanlineOptimizer.h:104-253,
// :40-64), including the reference's up/down behaviour: the grey pointer advances by one ELEMENT per step (:221,:250),
// the reference grey is the path's start pixel (:210,:232) and the "d-1" term reads prev[d] (:238).
//
// One warp owns one scanline (a row for left/right, a column for up/down).  Lane l keeps DPL consecutive
// disparities d = l*DPL .. l*DPL+DPL-1 of the previous pixel in registers; per step it needs one value from each
// neighbouring lane (two shuffles) and the minimum over all d, which is a single REDUX on an order-preserving uint
// key instead of a five-step shuffle tree -- the recurrence's critical path is ~2 dependent warp ops per pixel.
// Cost rows (and the running sum of the earlier paths) are prefetched PF pixels ahead into registers so the HBM
// latency is off the critical path; every load and store is a contiguous D*4-byte row of the [H][W][D] volume.
// The per-cell arithmetic is exactly the reference's three roundings ((C + min) - m, :180), so with a bit-identical
// input volume the four path volumes, their sum and the WTA disparity are bit-identical.
//
// The four paths are chained through ONE running-sum volume: left -> (left+right) -> (+up) -> (+down), the
// reference's summation order (:124); the last pass can fuse the first-minimum WTA and skip writing the sum.
#include "common.cuh"

__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

// three-input minimum (FMNMX3 on sm_100a)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int N, int VW>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&r)[N], int nvalid) {
    if (VW == 4) {
#pragma unroll
        for (int k = 0; k < N; k += 4)
            if (k < nvalid) {
                float4 v = __ldcs(reinterpret_cast<const float4*>(p + k));
                r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w;
            }
    } else if (VW == 2) {
#pragma unroll
        for (int k = 0; k < N; k += 2)
            if (k < nvalid) {
                float2 v = __ldcs(reinterpret_cast<const float2*>(p + k));
                r[k] = v.x; r[k + 1] = v.y;
            }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k < nvalid) r[k] = __ldcs(p + k);
    }
}
template <int N, int VW>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&r)[N], int nvalid) {
    if (VW == 4) {
#pragma unroll
        for (int k = 0; k < N; k += 4)
            if (k < nvalid) __stcs(reinterpret_cast<float4*>(p + k), make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]));
    } else if (VW == 2) {
#pragma unroll
        for (int k = 0; k < N; k += 2)
            if (k < nvalid) __stcs(reinterpret_cast<float2*>(p + k), make_float2(r[k], r[k + 1]));
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (k < nvalid) __stcs(p + k, r[k]);
    }
}

// VERT=false: line = row, walks x; VERT=true: line = column, walks y.  dir=+1 forward, -1 backward.
template <int DPL, int VW, int PF, bool VERT>
#ifndef SMT_SCAN_VMINB
#define SMT_SCAN_VMINB 1  // minimum resident one-warp blocks per SM asked of the vertical instantiations (register cap)
#endif
__global__ void __launch_bounds__((VERT && SMT_SCAN_VMINB > 1) ? 32 : 64, VERT ? SMT_SCAN_VMINB : 1) scan_path_kernel(const float* __restrict__ cost, const float* __restrict__ gray, int H, int W,
                                                       int D, float p1, float p2init, int dir, float* __restrict__ path_out,
                                                       const float* acc_in, float* acc_out, float* __restrict__ disp, int zero) {
    __shared__ float s_p2[256];  // max(p1, p2Init / (d + 1)) for integer grey differences d
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_p2[i] = fmaxf(p1, __fdiv_rn(p2init, __fadd_rn((float)i, 1.f)));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int line = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nlines = VERT ? W : H;
    if (line >= nlines) return;
    const int N = VERT ? H : W;
    const int d0 = lane * DPL;
    const int nvalid = max(0, min(DPL, D - d0));  // multiple of VW by construction
    const float PAD = 65535.f;                    // 0xffff (:151,:213)

    // pixel index of step s and its stride
    const long long pix0 = VERT ? ((dir > 0) ? (long long)line : (long long)(H - 1) * W + line)
                                : ((dir > 0) ? (long long)line * W : (long long)line * W + (W - 1));
    const long long pstep = VERT ? (long long)dir * W : (long long)dir;
    // grey walk: rows follow the pixel; columns advance by ONE element per step from the start pixel (:221,:250)
    const long long gstep = (long long)dir;

    float ringC[PF][DPL], ringA[PF][DPL], ringG[PF];
    const bool has_acc = acc_in != nullptr;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
        if (u < N) {
            const long long pix = pix0 + pstep * u;
            load_row<DPL, VW>(cost + pix * D + d0, ringC[u], nvalid);
            if (has_acc) load_row<DPL, VW>(acc_in + pix * D + d0, ringA[u], nvalid);
            ringG[u] = __ldg(gray + pix0 + gstep * u);
        }
    }

    const bool p1_nonneg = p1 >= 0.f;
    float prev[DPL];
#pragma unroll
    for (int k = 0; k < DPL; ++k) prev[k] = PAD;
    float m = PAD;
    const float gstart = __ldg(gray + pix0);
    float glast = gstart;

    // ptxas tracks every global load of this kernel on one scoreboard, and a scoreboard is a counter: the first use of ANY
    // loaded value waits for ALL loads issued before it.  So the ring is consumed and refilled in groups of G = PF/2
    // steps: (1) copy the group's rows out of the ring (the one wait, for loads that are at least a group old), (2) refill
    // those slots with the steps PF ahead -- the address depends on a value read in (1), so the loads cannot be hoisted
    // above the wait -- and ask L2 for the rows further ahead, (3) run the group's recurrence steps, which touch no
    // in-flight register.
    constexpr int G = PF / 2 > 0 ? PF / 2 : 1;
#ifndef SMT_SCAN_L2AHEAD
#define SMT_SCAN_L2AHEAD 0
#endif
    constexpr int L2AHEAD = SMT_SCAN_L2AHEAD;  // steps between the register refill and the L2 prefetch running ahead of it (0: none)
    for (int s0 = 0; s0 < N; s0 += PF) {
#pragma unroll
        for (int g0 = 0; g0 < PF; g0 += G) {
            float cC[G][DPL], cA[G][DPL], cG[G];
#pragma unroll
            for (int v = 0; v < G; ++v) {
                cG[v] = ringG[g0 + v];
#pragma unroll
                for (int k = 0; k < DPL; ++k) {
                    cC[v][k] = ringC[g0 + v][k];
                    cA[v][k] = has_acc ? ringA[g0 + v][k] : 0.f;
                }
            }
            const int dep = (__float_as_int(cC[0][0]) & zero) + (has_acc ? (__float_as_int(cA[0][0]) & zero) : 0) +
                            (__float_as_int(cG[0]) & zero);
#pragma unroll
            for (int v = 0; v < G; ++v) {
                const int sn = s0 + g0 + v + PF;
                if (sn < N) {
                    const long long pixn = pix0 + pstep * sn + dep;
                    load_row<DPL, VW>(cost + pixn * D + d0, ringC[g0 + v], nvalid);
                    if (has_acc) load_row<DPL, VW>(acc_in + pixn * D + d0, ringA[g0 + v], nvalid);
                    ringG[g0 + v] = __ldg(gray + pix0 + gstep * sn + dep);
                }
                const int sp = sn + L2AHEAD;
                if (L2AHEAD > 0 && sp < N && nvalid > 0) {
                    const long long pixp = pix0 + pstep * sp + dep;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(cost + pixp * D + d0));
                    if (has_acc) asm volatile("prefetch.global.L2 [%0];" ::"l"(acc_in + pixp * D + d0));
                }
            }
#pragma unroll
            for (int v = 0; v < G; ++v) {
                const int s = s0 + g0 + v;
                if (s < N) {
                    const long long pix = pix0 + pstep * s;
                    float cur[DPL];
                    if (s == 0) {
                        // first pixel of the path: L = C (:153,:215)
#pragma unroll
                        for (int k = 0; k < DPL; ++k) cur[k] = (k < nvalid) ? cC[v][k] : PAD;
                    } else {
                        const float g = cG[v];
                        const float gref = VERT ? gstart : glast;
                        // p2 = max(p1, p2Init / (|g - g_ref| + 1))  (:171,:232).  Grey differences are integers 0..255 for every
                        // image the drivers produce: those quotients come from a table (the IEEE division is a subroutine call,
                        // and a call waits for every load in flight); anything else takes the division itself.
                        const float ad = fabsf(__fsub_rn(g, gref));
                        const int ai = (int)ad;
                        float p2;
                        if ((float)ai == ad && ai < 256)
                            p2 = s_p2[ai];
                        else
                            p2 = fmaxf(p1, __fdiv_rn(p2init, __fadd_rn(ad, 1.f)));
                        const float l4 = __fadd_rn(m, p2);
                        float up_nb = __shfl_up_sync(0xffffffffu, prev[DPL - 1], 1);   // prev[d-1] of my first element
                        float dn_nb = __shfl_down_sync(0xffffffffu, prev[0], 1);       // prev[d+1] of my last element
                        if (lane == 0) up_nb = PAD;
                        if (lane == 31) dn_nb = PAD;
#pragma unroll
                        for (int k = 0; k < DPL; ++k) {
                            const float l1 = prev[k];
                            const float lm = (k == 0) ? up_nb : prev[k - 1];
                            const float lp = (k == DPL - 1) ? dn_nb : prev[k + 1];
                            // min(l1, l2, l3, l4) with l3 = lp + p1, l4 = m + p2 and l2 = lm + p1 (:177), resp. l2 = l1 + p1 for
                            // the up/down paths (:238).  Rounding is monotonic, so min(lm + p1, lp + p1) == min(lm, lp) + p1
                            // bit for bit, and l1 + p1 >= l1 for p1 >= 0 never wins: one add and a three-input minimum.
                            float mn;
                            if (VERT)
                                mn = p1_nonneg ? fmin3(l1, __fadd_rn(lp, p1), l4)
                                               : fmin3(fminf(l1, __fadd_rn(l1, p1)), __fadd_rn(lp, p1), l4);
                            else
                                mn = fmin3(l1, __fadd_rn(fminf(lm, lp), p1), l4);
                            const float vv = __fsub_rn(__fadd_rn(cC[v][k], mn), m);
                            cur[k] = (k < nvalid) ? vv : PAD;
                        }
                        glast = g;
                    }
                    // minimum over d of the new row (capped by the 0xffff pads, :162-166,:169,:182)
                    float lm_ = PAD;
#pragma unroll
                    for (int k = 0; k + 1 < DPL; k += 2) lm_ = fmin3(lm_, cur[k], cur[k + 1]);
                    if (DPL & 1) lm_ = fminf(lm_, cur[DPL - 1]);
                    m = key2f(__reduce_min_sync(0xffffffffu, f2key(lm_)));
#pragma unroll
                    for (int k = 0; k < DPL; ++k) prev[k] = cur[k];

                    if (path_out) store_row<DPL, VW>(path_out + pix * D + d0, cur, nvalid);
                    if (has_acc || acc_out || disp) {
                        float a[DPL];
#pragma unroll
                        for (int k = 0; k < DPL; ++k) a[k] = has_acc ? __fadd_rn(cA[v][k], cur[k]) : cur[k];
                        if (acc_out) store_row<DPL, VW>(acc_out + pix * D + d0, a, nvalid);
                        if (disp) {
                            // first minimum over d: ascending k inside the lane, lowest lane among equals (:51-58)
                            unsigned bk = 0xffffffffu;
                            int bd = 0;
#pragma unroll
                            for (int k = 0; k < DPL; ++k) {
                                const unsigned key = f2key(a[k]);
                                if (k < nvalid && key < bk) {
                                    bk = key;
                                    bd = d0 + k;
                                }
                            }
                            const unsigned gk = __reduce_min_sync(0xffffffffu, bk);
                            const unsigned who = __ballot_sync(0xffffffffu, bk == gk && nvalid > 0);
                            const int src = who ? (__ffs(who) - 1) : 0;
                            const int best = __shfl_sync(0xffffffffu, bd, src);
                            if (lane == 0) disp[pix] = (float)best;
                        }
                    }
                }
            }
        }
    }
}

template <int DPL, int VW, int PF, int PFV = PF>
static smt_status_t launch_paths(smt_ctx* ctx, const float* cost, const float* gray, int H, int W, int D, float p1, float p2,
                                 float* left, float* right, float* up, float* down, float* sum, float* disp, float* scratch) {
    const int WPB = 1;  // one scanline per block: 1080 / 1920 blocks spread evenly over the 148 SMs
    // pass 1: left -> L (scratch unless the caller wants the left volume)
    float* Lbuf = left ? left : scratch;
    SMT_LAUNCH(ctx, (scan_path_kernel<DPL, VW, PF, false>), cdiv(H, WPB), 32 * WPB, 0, cost, gray, H, W, D, p1, p2, +1, Lbuf,
               (const float*)nullptr, (float*)nullptr, (float*)nullptr, 0);
    // pass 2: right; acc = left + right  (in place over scratch when it holds `left`)
    SMT_LAUNCH(ctx, (scan_path_kernel<DPL, VW, PF, false>), cdiv(H, WPB), 32 * WPB, 0, cost, gray, H, W, D, p1, p2, -1, right,
               (const float*)Lbuf, scratch, (float*)nullptr, 0);
    // pass 3: up; acc += up
    SMT_LAUNCH(ctx, (scan_path_kernel<DPL, VW, PFV, true>), cdiv(W, WPB), 32 * WPB, 0, cost, gray, H, W, D, p1, p2, +1, up,
               (const float*)scratch, scratch, (float*)nullptr, 0);
    // pass 4: down; acc += down -> sum (optional) and WTA (optional)
    SMT_LAUNCH(ctx, (scan_path_kernel<DPL, VW, PFV, true>), cdiv(W, WPB), 32 * WPB, 0, cost, gray, H, W, D, p1, p2, -1, down,
               (const float*)scratch, sum, disp, 0);
    return SMT_OK;
}

smt_status_t k_scanline(smt_ctx* ctx, const float* cost, const float* gray, int H, int W, int D, int p1, int p2, float* left,
                        float* right, float* up, float* down, float* sum, float* disp, float* scratch, float* /*unused*/) {
    StageTimer t(ctx, ST_SCANLINE);
    const float fp1 = (float)p1, fp2 = (float)p2;  // float p1 = _p1 (:132-133)
    int dpl = 1;
    while (dpl * 32 < D) dpl *= 2;
    if (dpl > 16) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "scanline: D=%d exceeds 512 disparities per scanline warp", D);
    auto aligned = [](const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    int vw = 1;
    for (int cand : {4, 2}) {
        if (dpl % cand == 0 && D % cand == 0 && aligned(cost, 4 * cand) && aligned(left, 4 * cand) &&
            aligned(right, 4 * cand) && aligned(up, 4 * cand) && aligned(down, 4 * cand) && aligned(sum, 4 * cand) &&
            aligned(scratch, 4 * cand)) {
            vw = cand;
            break;
        }
    }
#ifndef SMT_SCAN_PF8
#define SMT_SCAN_PF8 4  // ring depth (steps) of the DPL = 8, float4 kernel, the 1080p x 256 case: horizontal passes
#endif
#define SMT_SCAN_CASE(DPL_, VW_, PF_) \
    if (dpl == DPL_ && vw == VW_)     \
        return launch_paths<DPL_, VW_, PF_>(ctx, cost, gray, H, W, D, fp1, fp2, left, right, up, down, sum, disp, scratch);
    SMT_SCAN_CASE(16, 4, 2)
#ifndef SMT_SCAN_PFV8
// The vertical passes have 1920 one-warp blocks: at depth 6 (216 registers, 9 warps per SM) they need two waves, at depth 2
// (94 registers) one, and the lighter kernels share the SMs better with the other context's work.  Alone the stage
// measures 4.19 (6, 6), 4.06 (6, 2), 4.13 (4, 2) ms; two contexts: 62.7, 62.9, 63.7 frames/s.
#define SMT_SCAN_PFV8 2
#endif
    if (dpl == 8 && vw == 4)
        return launch_paths<8, 4, SMT_SCAN_PF8, SMT_SCAN_PFV8>(ctx, cost, gray, H, W, D, fp1, fp2, left, right, up, down, sum, disp, scratch);
    SMT_SCAN_CASE(4, 4, 4)
    SMT_SCAN_CASE(16, 2, 2)
    SMT_SCAN_CASE(8, 2, 4)
    SMT_SCAN_CASE(4, 2, 4)
    SMT_SCAN_CASE(2, 2, 4)
    SMT_SCAN_CASE(16, 1, 2)
    SMT_SCAN_CASE(8, 1, 4)
    SMT_SCAN_CASE(4, 1, 4)
    SMT_SCAN_CASE(2, 1, 4)
    SMT_SCAN_CASE(1, 1, 4)
#undef SMT_SCAN_CASE
    return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "scanline: no kernel for D=%d", D);
}

extern "C" smt_status_t smt_scanline(smt_ctx* ctx, const float* cost, const float* gray, int32_t rows, int32_t cols, int32_t D,
                                     int32_t p1, int32_t p2, float* left, float* right, float* up, float* down, float* sum,
                                     float* disp, smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, cost && gray && rows > 0 && cols > 0 && D > 0, SMT_ERR_INVALID_ARG, "smt_scanline: bad argument");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)rows * cols, nvol = npix * D;
    Marshal ms(ctx, space);
    const float* dc = ms.in(cost, nvol);
    const float* dg = ms.in(gray, npix);
    float* dl = ms.out(left, nvol);
    float* dr = ms.out(right, nvol);
    float* du = ms.out(up, nvol);
    float* dd = ms.out(down, nvol);
    float* dsum = ms.out(sum, nvol);
    float* ddisp = ms.out(disp, npix);
    SMT_TRY(ms.check());
    SMT_ALLOC(ctx, scratch, nvol * sizeof(float));
    SMT_TRY(k_scanline(ctx, dc, dg, rows, cols, D, p1, p2, dl, dr, du, dd, dsum, ddisp, scratch.as<float>(), nullptr));
    return ms.finish();
}
