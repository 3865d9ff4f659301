// SMT_SUM_FAST cross aggregation: shared-memory line kernels.
//
// Same passes as CrossAggregator::AggregateInArms (CBLSM/cross_aggregator.cpp:327-394), different evaluation:
//   * a pass along a line (row for the horizontal arm pair, column for the vertical pair) is a variable-width
//     window sum  out[i] = sum_{t=-lo(i)..hi(i)} in[i+t]  -> one exclusive prefix E per line, out[i] = E[i+hi+1]-E[i-lo];
//   * the prefix is kept as an unevaluated float32 pair (hi, lo) ("double-single", error-free TwoSum updates, ~2^-46
//     relative), so E differences are the window sum to ~1e-11 and are rounded to float32 once (the reference rounds
//     after every one of its <=68 sequential adds, so the two differ by the reference's own rounding noise,
//     ~1e-7..1e-6 relative; tests hold the volume to <=1e-5).  float64 was measured first and rejected: on this B200
//     the FP64 pipe (DADD and F2F.F64) sustained ~1.3 lane-ops/clk/SM, 100 ms per frame at 1080p x 256;
//   * consecutive passes in the same direction (the reference alternates H,V | V,H | H,V | V,H, so V,V and H,H pairs
//     occur) are fused in one kernel: 8 passes -> 5 sweeps over the volume.
// A CTA owns a tile of one line: TL consecutive positions (+34 halo per pass each side) x 32 consecutive disparities.
// Lane = disparity (every global access is a 128-byte row of the [H][W][D] volume), warp w owns the w-th eighth of the
// tile positions: it loads them straight into registers (~41-49 independent 128 B loads in flight per warp), the 8
// warps exchange segment totals through shared memory, then each writes its slice of E.  Queries read two float64 per
// output from shared memory; arms and support counts are warp-uniform broadcast loads.
#include <cstdlib>

#include "common.cuh"

namespace {

// ---- double-single arithmetic (no FMA contraction: the library is built with -fmad=false) ----
struct ds {
    float hi, lo;
};
// (a.hi + a.lo) + b, renormalised
__device__ __forceinline__ ds ds_add_f(ds a, float b) {
    const float s = __fadd_rn(a.hi, b);
    const float bb = __fsub_rn(s, a.hi);
    const float e = __fadd_rn(__fsub_rn(a.hi, __fsub_rn(s, bb)), __fsub_rn(b, bb));  // TwoSum error term
    const float l = __fadd_rn(a.lo, e);
    ds r;
    r.hi = __fadd_rn(s, l);
    r.lo = __fsub_rn(l, __fsub_rn(r.hi, s));
    return r;
}
__device__ __forceinline__ ds ds_add(ds a, ds b) {
    const float s = __fadd_rn(a.hi, b.hi);
    const float bb = __fsub_rn(s, a.hi);
    const float e = __fadd_rn(__fsub_rn(a.hi, __fsub_rn(s, bb)), __fsub_rn(b.hi, bb));
    const float l = __fadd_rn(__fadd_rn(a.lo, b.lo), e);
    ds r;
    r.hi = __fadd_rn(s, l);
    r.lo = __fsub_rn(l, __fsub_rn(r.hi, s));
    return r;
}
// float32(b - a) for prefix values b >= a of one chain
__device__ __forceinline__ float ds_diff(float2 b, float2 a) {
    return __fadd_rn(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y));
}

constexpr int AGG_WARPS = 8;
constexpr int AGG_THREADS = AGG_WARPS * 32;
constexpr int ARM_MAX = 34 + 1;  // the kernels assume arms <= 34 (L1 <= 34); checked on the host

template <int NPASS>
struct AggCfg {
    static constexpr int HALO = 34 * NPASS;
    static constexpr int MAXSEG = NPASS == 1 ? 41 : 49;     // positions per warp
    static constexpr int MAXN = MAXSEG * AGG_WARPS;          // loaded positions per tile (328 / 392)
    static constexpr int MAXTL = MAXN - 2 * HALO;            // 260 / 256 output positions per tile
    static constexpr size_t SMEM = sizeof(float2) * ((size_t)(MAXN + 1) * 32 + AGG_WARPS * 32) + sizeof(uint2) * MAXN +
                                   sizeof(float) * MAXN * NPASS;
};

// ---- cascaded compensated prefix for NON-NEGATIVE inputs (Ogita-Rump-Oishi "Sum2" as a running pair) ---------------
// state (s, c) represents s + c: s is the plain float32 running sum, c the running sum of the EXACT rounding errors of
// every s update.  Starting s at a power-of-two bias >= every input keeps |s| >= |v|, so the 3-op FastTwoSum error
// term is exact; c only ever adds numbers of size ~ulp(s), so its own rounding is second order (2^-48 relative to s).
// Two chains over the same data (segment totals vs. segment prefixes) therefore agree to second order, which is what
// lets a window that spans two warps' segments come out exact (e.g. exactly 0 over a run of zero costs).
// 4 FADD per element instead of 10 for the general TwoSum form above.
struct ks {
    float s, c;
};
__device__ __forceinline__ ks ks_add(ks a, float v) {
    ks r;
    r.s = __fadd_rn(a.s, v);
    r.c = __fadd_rn(a.c, __fsub_rn(v, __fsub_rn(r.s, a.s)));
    return r;
}
// float32(P_b - P_a), P = s + c
__device__ __forceinline__ float ks_diff(float2 b, float2 a) {
    return __fadd_rn(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y));
}

// horiz: line index = row, position = column;  else line index = column, position = row.
// NONNEG: inputs are known to be >= 0 and <= the pass's bias: cascaded compensated prefix.
template <int NPASS, bool NONNEG>
__global__ void __launch_bounds__(AGG_THREADS, 2)
    agg_line_kernel(const float* __restrict__ in, float* __restrict__ out, const uchar4* __restrict__ arms,
                    const uint16_t* __restrict__ supA, const uint16_t* __restrict__ supB, int H, int W, int D, int horiz,
                    int TL, int ntile, float bias0, float bias1) {
    using Cfg = AggCfg<NPASS>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* E = reinterpret_cast<float2*>(smem_raw);          // [(MAXN+1)][32] prefix pairs
    float2* tot = E + (size_t)(Cfg::MAXN + 1) * 32;           // [AGG_WARPS][32]
    uint2* win = reinterpret_cast<uint2*>(tot + AGG_WARPS * 32);         // [MAXN] byte offsets of E[j-lo], E[j+hi+1]
    float* supf = reinterpret_cast<float*>(win + Cfg::MAXN);             // [NPASS][MAXN] divisors as float

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x % ntile, chunk = blockIdx.x / ntile;
    const int line = blockIdx.y;
    const int len = horiz ? W : H;
    const int d = chunk * 32 + lane;
    const bool act = d < D;
    const long long pstride = horiz ? 1 : (long long)W;       // pixel-index step along the line
    const long long pbase = horiz ? (long long)line * W : (long long)line;

    const int c0 = tile * TL, c1 = min(len, c0 + TL);         // output positions
    const int l0 = max(0, c0 - Cfg::HALO), l1 = min(len, c1 + Cfg::HALO);  // loaded positions
    const int nload = l1 - l0;
    const int seg = (nload + AGG_WARPS - 1) / AGG_WARPS;      // <= MAXSEG
    const int s0 = min(nload, warp * seg), s1 = min(nload, s0 + seg);     // my local positions [s0, s1)
    const int nseg = s1 - s0;

    // ---- my segment of the input line, straight into registers (independent 128-byte row loads)
    float v[Cfg::MAXSEG];
    {
        const float* src = in + (pbase + pstride * (l0 + s0)) * D + d;
        const long long step = pstride * D;
#pragma unroll
        for (int k = 0; k < Cfg::MAXSEG; ++k) {
            v[k] = 0.f;
            if (k < nseg && act) v[k] = __ldg(src + step * k);
        }
    }
    // ---- window descriptors and divisors of the tile (warp-uniform in the query loops)
    for (int j = threadIdx.x; j < nload; j += AGG_THREADS) {
        const long long pix = pbase + pstride * (l0 + j);
        const uchar4 a = __ldg(arms + pix);
        const int lo = horiz ? a.x : a.z, hi = horiz ? a.y : a.w;
        // clamp to the loaded range: only halo positions whose result is never used can exceed it
        win[j] = make_uint2((unsigned)max(j - lo, 0) * 256u, (unsigned)min(j + hi + 1, nload) * 256u);
        if (supA) supf[j] = (float)(int)__ldg(supA + pix);
        if (NPASS == 2 && supB) supf[Cfg::MAXN + j] = (float)(int)__ldg(supB + pix);
    }

#pragma unroll
    for (int pass = 0; pass < NPASS; ++pass) {
        const bool last = pass == NPASS - 1;
        const float bias = pass == 0 ? bias0 : bias1;
        // ---- exclusive compensated prefix of the line into E (positions outside the valid range hold 0)
        if (NONNEG) {
            ks t = {bias, 0.f};
#pragma unroll
            for (int k = 0; k < Cfg::MAXSEG; ++k) t = ks_add(t, v[k]);
            tot[warp * 32 + lane] = make_float2(__fsub_rn(t.s, bias), t.c);  // exact: t.s >= bias, same binade or above
            __syncthreads();
            ks acc = {bias, 0.f};
            for (int w = 0; w < warp; ++w) {
                const float2 tw = tot[w * 32 + lane];
                // general TwoSum here: a segment total may exceed the running value (<= 7 steps per thread)
                const float sn = __fadd_rn(acc.s, tw.x);
                const float bb = __fsub_rn(sn, acc.s);
                const float er = __fadd_rn(__fsub_rn(acc.s, __fsub_rn(sn, bb)), __fsub_rn(tw.x, bb));
                acc.c = __fadd_rn(__fadd_rn(acc.c, tw.y), er);
                acc.s = sn;
            }
            float2* e = E + (size_t)s0 * 32 + lane;
#pragma unroll
            for (int k = 0; k < Cfg::MAXSEG; ++k) {
                if (k < nseg) e[k * 32] = make_float2(acc.s, acc.c);
                acc = ks_add(acc, v[k]);
            }
            if (warp == AGG_WARPS - 1) E[(size_t)nload * 32 + lane] = make_float2(acc.s, acc.c);
        } else {
            ds t = {0.f, 0.f};
#pragma unroll
            for (int k = 0; k < Cfg::MAXSEG; ++k) t = ds_add_f(t, v[k]);
            tot[warp * 32 + lane] = make_float2(t.hi, t.lo);
            __syncthreads();
            ds acc = {0.f, 0.f};
            for (int w = 0; w < warp; ++w) {
                const float2 tw = tot[w * 32 + lane];
                acc = ds_add(acc, ds{tw.x, tw.y});
            }
            float2* e = E + (size_t)s0 * 32 + lane;
#pragma unroll
            for (int k = 0; k < Cfg::MAXSEG; ++k) {
                if (k < nseg) e[k * 32] = make_float2(acc.hi, acc.lo);
                acc = ds_add_f(acc, v[k]);
            }
            if (warp == AGG_WARPS - 1) E[(size_t)nload * 32 + lane] = make_float2(acc.hi, acc.lo);
        }
        __syncthreads();

        // ---- window sums for the positions the next stage needs (local range [q0, q1))
        const int q0 = last ? (c0 - l0) : max(0, c0 - l0 - 34 * (NPASS - 1 - pass));
        const int q1 = last ? (c1 - l0) : min(nload, c1 - l0 + 34 * (NPASS - 1 - pass));
        const bool has_sup = (pass == 0 ? supA : supB) != nullptr;
        const float* sf = supf + pass * Cfg::MAXN + s0;
        const uint2* wn = win + s0;
        const unsigned char* El = reinterpret_cast<const unsigned char*>(E + lane);
        float* dst = out + (pbase + pstride * (l0 + s0)) * D + d;
        const long long step = pstride * D;
        const int k0 = q0 - s0, k1 = min(q1 - s0, nseg);
#pragma unroll
        for (int k = 0; k < Cfg::MAXSEG; ++k) {
            float r = 0.f;
            if (k >= k0 && k < k1) {
                const uint2 w = wn[k];
                const float2 ea = *reinterpret_cast<const float2*>(El + w.x);
                const float2 eb = *reinterpret_cast<const float2*>(El + w.y);
                r = NONNEG ? ks_diff(eb, ea) : ds_diff(eb, ea);
                if (has_sup) r = __fdiv_rn(r, sf[k]);
                if (last && act) dst[step * k] = r;
            }
            v[k] = r;
        }
        if (!last) __syncthreads();  // everyone is done reading E before the next pass overwrites it
    }
}

struct Pass {
    bool horiz;
    const uint16_t* sup;
};

}  // namespace

smt_status_t k_transpose_maps(smt_ctx* ctx, const uint8_t* arms8, const uint16_t* sup, int H, int W, uint8_t* armsT, uint16_t* supT);
smt_status_t k_agg_seq_pass(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D,
                            int horiz, bool nonneg, float bias, int tmaps);  // cross_agg_seq.cu
smt_status_t k_agg_seq_pair(smt_ctx* ctx, const float* in, float* out, const uint8_t* arms8, const uint16_t* sup_a, int H, int W, int D,
                            int horiz, bool nonneg, float bias_a, float bias_b, int tmaps);
smt_status_t k_agg_seq_pass_wta(smt_ctx* ctx, const float* in, const uint8_t* arms8, const uint16_t* sup, int H, int W, int D, int horiz,
                                float bias, float* disp);

// vmax: an upper bound of the (non-negative) input values, or <= 0 when the sign/range of the input is unknown.
smt_status_t k_aggregate_cross_fast(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms8, const uint16_t* sup0,
                                    const uint16_t* sup1, int H, int W, int D, int iters, float vmax, float** result, float* wta_disp) {
    StageTimer t(ctx, ST_AGGREGATE);
    const uchar4* arms = reinterpret_cast<const uchar4*>(arms8);
    // pass list (cross_aggregator.cpp:102-117): horizontal-first on even iterations; the second pass divides
    std::vector<Pass> passes;
    bool hfirst = true;
    for (int k = 0; k < iters; ++k) {
        passes.push_back({hfirst, nullptr});
        passes.push_back({!hfirst, hfirst ? sup0 : sup1});
        hfirst = !hfirst;
    }
    static bool attr_set_dev[64] = {};  // function attributes are per device
    bool& attr_set = attr_set_dev[ctx->device & 63];
    if (!attr_set) {
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_line_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AggCfg<1>::SMEM));
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_line_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AggCfg<2>::SMEM));
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_line_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AggCfg<1>::SMEM));
        SMT_CUDA(ctx, cudaFuncSetAttribute(agg_line_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AggCfg<2>::SMEM));
        attr_set = true;
    }
    float* cur = vol;
    float* oth = tmp;
    const int nchunk = (D + 31) / 32;
    const bool nonneg = vmax > 0.f && vmax < 1e30f;
    float b_plain = 1.f;  // power of two >= vmax: bias of a pass whose input is a cost (or an average of costs)
    while (nonneg && b_plain < vmax) b_plain *= 2.f;
    const float b_sum = b_plain * 128.f;  // input of a dividing pass: sums of <= 69 such values
    // fused V,V / H,H walks halve the traffic of those passes but need two rings per warp (4 warps per SM); measured
    // 23.7 ms vs 21.0 ms per 1080p x 256 frame, so they are opt-in (SMT_AGG_FUSE=1)
    static const bool fuse_pairs = getenv("SMT_AGG_FUSE") != nullptr;
    static const bool fuse_vertical_only = getenv("SMT_AGG_FUSE_V") != nullptr;
    // default: the streaming one-warp-per-line kernel (cross_agg_seq.cu); SMT_AGG_KERNEL=tile selects the tiled one
    static const bool use_tile = getenv("SMT_AGG_KERNEL") != nullptr && std::string(getenv("SMT_AGG_KERNEL")) == "tile";
    if (!use_tile) {
        // vertical walks read the arm words and sup0 of 32 consecutive rows at a time: give them [W][H] copies
        const size_t npix = (size_t)H * W;
        SMT_ALLOC(ctx, armsT, npix * 4);
        SMT_ALLOC(ctx, sup0T, npix * sizeof(uint16_t));
        SMT_TRY(k_transpose_maps(ctx, arms8, sup0, H, W, armsT.as<uint8_t>(), sup0T.as<uint16_t>()));
        for (size_t k = 0; k < passes.size();) {
            // a dividing pass followed by a non-dividing pass in the same direction (V,V / H,H) runs as one fused walk
            const bool pair = fuse_pairs && k + 1 < passes.size() && passes[k + 1].horiz == passes[k].horiz && passes[k].sup &&
                              !passes[k + 1].sup && (!fuse_vertical_only || !passes[k].horiz);
            // the last pass feeds only a winner-take-all: reduce in the kernel, never write the volume
            if (wta_disp && nonneg && k + 1 == passes.size() && passes[k].sup) {
                SMT_TRY(k_agg_seq_pass_wta(ctx, cur, arms8, passes[k].sup, H, W, D, passes[k].horiz ? 1 : 0, b_sum, wta_disp));
                *result = nullptr;
                return SMT_OK;
            }
            if (pair) {
                if (passes[k].horiz)
                    SMT_TRY(k_agg_seq_pair(ctx, cur, oth, arms8, passes[k].sup, H, W, D, 1, nonneg, b_sum, b_plain, 0));
                else
                    SMT_TRY(k_agg_seq_pair(ctx, cur, oth, armsT.as<uint8_t>(), sup0T.as<uint16_t>(), H, W, D, 0, nonneg, b_sum, b_plain, 1));
                k += 2;
            } else {
                if (passes[k].horiz)
                    SMT_TRY(k_agg_seq_pass(ctx, cur, oth, arms8, passes[k].sup, H, W, D, 1, nonneg, passes[k].sup ? b_sum : b_plain, 0));
                else
                    SMT_TRY(k_agg_seq_pass(ctx, cur, oth, armsT.as<uint8_t>(), passes[k].sup ? sup0T.as<uint16_t>() : nullptr, H, W, D, 0,
                                           nonneg, passes[k].sup ? b_sum : b_plain, 1));
                k += 1;
            }
            float* sw = cur;
            cur = oth;
            oth = sw;
        }
        *result = cur;
        if (wta_disp) SMT_TRY(k_wta(ctx, cur, H, W, D, wta_disp));
        return SMT_OK;
    }
    for (size_t k = 0; k < passes.size();) {
        const bool fuse = fuse_pairs && k + 1 < passes.size() && passes[k + 1].horiz == passes[k].horiz;
        // input of pass k is a plain cost if pass k does not divide (it is the first pass of its iteration)
        const float bias0 = passes[k].sup ? b_sum : b_plain;
        const float bias1 = fuse ? (passes[k + 1].sup ? b_sum : b_plain) : 0.f;
        const int horiz = passes[k].horiz ? 1 : 0;
        const int len = horiz ? W : H, nlines = horiz ? H : W;
        const int maxtl = fuse ? AggCfg<2>::MAXTL : AggCfg<1>::MAXTL;
        const int ntile = (len + maxtl - 1) / maxtl;
        const int TL = (len + ntile - 1) / ntile;
        dim3 grid((unsigned)(ntile * nchunk), (unsigned)nlines);
        if (nlines > 65535) return smt_fail(ctx, SMT_ERR_UNSUPPORTED, "aggregate: more than 65535 lines per direction");
        if (fuse) {
            if (nonneg)
                SMT_LAUNCH(ctx, (agg_line_kernel<2, true>), grid, AGG_THREADS, AggCfg<2>::SMEM, cur, oth, arms, passes[k].sup,
                           passes[k + 1].sup, H, W, D, horiz, TL, ntile, bias0, bias1);
            else
                SMT_LAUNCH(ctx, (agg_line_kernel<2, false>), grid, AGG_THREADS, AggCfg<2>::SMEM, cur, oth, arms, passes[k].sup,
                           passes[k + 1].sup, H, W, D, horiz, TL, ntile, bias0, bias1);
            k += 2;
        } else {
            if (nonneg)
                SMT_LAUNCH(ctx, (agg_line_kernel<1, true>), grid, AGG_THREADS, AggCfg<1>::SMEM, cur, oth, arms, passes[k].sup,
                           (const uint16_t*)nullptr, H, W, D, horiz, TL, ntile, bias0, bias1);
            else
                SMT_LAUNCH(ctx, (agg_line_kernel<1, false>), grid, AGG_THREADS, AggCfg<1>::SMEM, cur, oth, arms, passes[k].sup,
                           (const uint16_t*)nullptr, H, W, D, horiz, TL, ntile, bias0, bias1);
            k += 1;
        }
        float* sw = cur;
        cur = oth;
        oth = sw;
    }
    *result = cur;
    if (wta_disp) SMT_TRY(k_wta(ctx, cur, H, W, D, wta_disp));
    return SMT_OK;
}
