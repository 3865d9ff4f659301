// SMT_SUM_FAST cross aggregation: the sweep schedule.
//
// CrossAggregator::Aggregate (CBLSM/cross_aggregator.cpp:89-118) runs `iters` iterations of two passes each, horizontal
// first on even iterations, the second pass dividing by the support count: H V' | V H' | H V' | V H'.  A dividing pass and
// the non-dividing pass that follows it walk the same direction, so the schedule groups them:
//     H | V'V | H'H | V'V | H'      8 passes, 5 sweeps over the volume.
// Each sweep is one launch of the fixed-point kernels of cross_agg_fx.cu.  Behind it the same passes are queued on the
// compensated float-pair kernel of cross_agg_seq.cu, reading the sweep's redo list: they redo the (line, disparity chunk)
// tasks whose values fell below the fixed-point error bound's domain -- none or a few per cent on AD-Census costs -- or the
// whole sweep once the list outgrew a quarter of the tasks, after which the fixed-point launches return at once.  A
// float-pair redo of a fused pair runs its first pass in place (a line walk reads ahead of what it writes), so both
// routes leave the sweep's result in the same buffer.  Inputs whose sign or range is unknown (vmax <= 0) go straight to
// the float-pair kernels.
#include "common.cuh"

namespace {

struct Pass {
    bool horiz;
    const uint16_t* sup;
};

}  // namespace

// vmax: an upper bound of the (non-negative) input values, or <= 0 when the sign/range of the input is unknown.
// wta_disp (optional): first-minimum winner-take-all of the result.  Reducing inside the last pass (no volume write) was
// measured twice and lost both times: the warp reduction per position costs the pass more than the separate streaming
// WTA kernel (float pairs, round 1: 13.71 against 13.05 + 0.45 ms per frame; fixed point, round 2: 1.53 against
// 0.79 + 0.46 ms for the one pass concerned).
static smt_status_t fast_sweeps(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms8, const uint16_t* sup0, const uint16_t* sup1,
                                int H, int W, int D, int iters, float vmax, float** result) {
    StageTimer t(ctx, ST_AGGREGATE);
    // pass list (cross_aggregator.cpp:102-117): horizontal-first on even iterations; the second pass divides
    std::vector<Pass> passes;
    bool hfirst = true;
    for (int k = 0; k < iters; ++k) {
        passes.push_back({hfirst, nullptr});
        passes.push_back({!hfirst, hfirst ? sup0 : sup1});
        hfirst = !hfirst;
    }
    float* cur = vol;
    float* oth = tmp;
    const size_t npix = (size_t)H * W;
    const bool nonneg = vmax > 0.f && vmax < 1e30f;
    float b_plain = 1.f;  // power of two >= vmax: bias of a float-pair pass whose input is a cost (or an average of costs)
    while (nonneg && b_plain < vmax) b_plain *= 2.f;
    const float b_sum = b_plain * 128.f;  // input of a dividing pass: sums of <= 69 such values
    float inv_q = 0.f, q = 0.f;
    const bool fx = nonneg && fx_quantum(vmax, &inv_q, &q);
    const size_t ntasks = (size_t)(H > W ? H : W) * ((D + 31) / 32);
    int* state = fx ? smt_redo_state(ctx, ntasks) : nullptr;
    if (fx && !state) return smt_fail(ctx, SMT_ERR_OOM, "aggregate: redo list allocation failed");
    // vertical walks read the arm words and the support count of 24 consecutive rows at a time: give them [W][H] copies
    // (a vertical pass only ever divides by sup0: it is the second pass of a horizontal-first iteration)
    SMT_ALLOC(ctx, armsT, npix * 4);
    SMT_ALLOC(ctx, sup0T, npix * sizeof(uint16_t));
    SMT_TRY(k_transpose_maps(ctx, arms8, sup0, H, W, armsT.as<uint8_t>(), sup0T.as<uint16_t>()));
    // a third volume for the float-pair redo of fused pairs: the listed lines' intermediate pass goes there, so that both
    // redo launches can cut their lines into segments (an in-place first pass has to walk whole lines, ~50 us each)
    // (when the device cannot hold a third volume the first pass is redone in place: a line walk reads ahead of what it writes)
    DevBuf spare(ctx, fx && iters > 1 ? npix * (size_t)D * sizeof(float) : 0);
    if (!spare.ok()) cudaGetLastError();  // the failed allocation is not an error of this call

    for (size_t k = 0; k < passes.size();) {
        const int horiz = passes[k].horiz ? 1 : 0;
        const int tmaps = horiz ? 0 : 1;
        const uint8_t* amap = horiz ? arms8 : armsT.as<uint8_t>();
        const uint16_t* smap = !passes[k].sup ? nullptr : (horiz ? passes[k].sup : sup0T.as<uint16_t>());
        const float bias = passes[k].sup ? b_sum : b_plain;
        const bool last = k + 1 == passes.size();
        // a dividing pass followed by a non-dividing pass along the same direction: one fused fixed-point walk
        const bool pair = fx && !last && passes[k].sup && !passes[k + 1].sup && passes[k + 1].horiz == passes[k].horiz;
        if (fx) SMT_TRY(k_fx_state_reset(ctx, state, k == 0, ntasks));
        if (pair) {
            SMT_TRY(k_agg_fx_pair(ctx, cur, oth, amap, smap, H, W, D, horiz, tmaps, inv_q, q, state));
            // float-pair redo of the listed lines: A into the spare volume, then B from there into `oth`
            float* mid = spare.ok() ? spare.as<float>() : cur;
            SMT_TRY(k_agg_seq_pass(ctx, cur, mid, amap, smap, H, W, D, horiz, true, b_sum, tmaps, state));
            SMT_TRY(k_agg_seq_pass(ctx, mid, oth, amap, nullptr, H, W, D, horiz, true, b_plain, tmaps, state));
            k += 2;
        } else {
            if (fx) SMT_TRY(k_agg_fx_pass(ctx, cur, oth, amap, smap, H, W, D, horiz, tmaps, inv_q, q, state));
            SMT_TRY(k_agg_seq_pass(ctx, cur, oth, amap, smap, H, W, D, horiz, nonneg, bias, tmaps, state));
            k += 1;
        }
        float* sw = cur;
        cur = oth;
        oth = sw;
    }
    *result = cur;
    return SMT_OK;
}

smt_status_t k_aggregate_cross_fast(smt_ctx* ctx, float* vol, float* tmp, const uint8_t* arms8, const uint16_t* sup0,
                                    const uint16_t* sup1, int H, int W, int D, int iters, float vmax, float** result, float* wta_disp) {
    SMT_TRY(fast_sweeps(ctx, vol, tmp, arms8, sup0, sup1, H, W, D, iters, vmax, result));
    if (wta_disp) SMT_TRY(k_wta(ctx, *result, H, W, D, wta_disp));  // its own stage (ST_WTA)
    return SMT_OK;
}
