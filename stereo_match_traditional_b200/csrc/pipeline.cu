// P_full: the AD-Census full pipeline as a composition of the reference's entry points:
//   AD_Census::ComputeADcensus / ComputeADcensusRight         (AD-CensusV1/AD-Census.h:271,296; main.cpp:58-60)
//   CrossAggregator::Aggregate(4) on each view                 (CBLSM/cross_aggregator.cpp:89; CBLSM.cpp:138-143)
//   ScanlineOptimizer::ScanLine + WTA on the left view         (ScanlineOptimizer.h:104,40; main.cpp:86-89)
//   first-minimum WTA of the aggregated right volume           (CrossArm.cpp:33; main.cpp:84)
//   LeftRightConsistency, RemoveSpeckles, MedianFilter         (PostProcessing.h:72,250,314; main.cpp:92-94)
// Everything stays on the device; only the two image pairs go in and disparity maps come out.
#include "common.cuh"

extern "C" void smt_adcensus_default_params(smt_adcensus_params* p, int32_t rows, int32_t cols, int32_t D) {
    if (!p) return;
    p->rows = rows;
    p->cols = cols;
    p->D = D;
    p->sigmaC = 10.f;  // main.cpp:25
    p->sigmaS = 30.f;  // main.cpp:26
    p->L1 = 34;        // adcensus_types.h:69
    p->L2 = 17;
    p->t1 = 20;        // adcensus_types.h:70
    p->t2 = 6;
    p->agg_iters = 4;  // CBLSM.cpp:141
    p->p1 = 10;        // main.cpp:28
    p->p2 = 150;       // main.cpp:29
    p->gate = 2;       // main.cpp:30
    p->speckle_diff = 1;
    p->speckle_min_area = 30;
    p->speckle_invalid = 2147483647;  // Invalid_Float bound to `const int&` (main.cpp:93)
    p->median_size = 3;
    p->sum_mode = SMT_SUM_REFERENCE_ORDER;
    p->irv_iters = 0;  // the refinement steps are not part of the reference's pipeline
    p->irv_ts = 20;    // adcensus_types.h:73
    p->irv_th = 0.4f;
    p->do_interpolation = 0;
    p->do_discontinuity_adjustment = 0;
    p->discontinuity_thres = 5.f;
}

extern "C" smt_status_t smt_pipeline_adcensus_full(smt_ctx* ctx, const smt_adcensus_params* p, const uint8_t* bgr_left,
                                                   const uint8_t* bgr_right, const uint8_t* gray_left, const uint8_t* gray_right,
                                                   float* disp_out, float* disp_left_raw, float* disp_right_raw, float* disp_lrc,
                                                   smt_memspace_t space) {
    if (!ctx) return SMT_ERR_INVALID_ARG;
    SMT_REQUIRE(ctx, p && bgr_left && bgr_right && disp_out, SMT_ERR_INVALID_ARG, "smt_pipeline_adcensus_full: null argument");
    const int H = p->rows, W = p->cols, D = p->D;
    SMT_REQUIRE(ctx, H > 0 && W > 0 && D > 0 && p->agg_iters >= 0, SMT_ERR_INVALID_ARG,
                "smt_pipeline_adcensus_full: rows, cols, D must be positive");
    SMT_REQUIRE(ctx, p->sum_mode != SMT_SUM_FAST || p->L1 <= 34, SMT_ERR_UNSUPPORTED,
                "smt_pipeline_adcensus_full: SMT_SUM_FAST tiles assume arms <= 34 (L1 <= 34, the reference default)");
    DeviceGuard dev_guard(ctx->device);
    const size_t npix = (size_t)H * W, nvol = npix * D;

    Marshal ms(ctx, space);
    const uint8_t *bl, *br, *gl, *gr;
    {
        StageTimer tc(ctx, ST_COPY);
        bl = ms.in(bgr_left, npix * 3);
        br = ms.in(bgr_right, npix * 3);
        gl = ms.in(gray_left, npix);
        gr = ms.in(gray_right, npix);
    }
    float* d_out = ms.out(disp_out, npix);
    float* d_lraw = ms.out(disp_left_raw, npix);
    float* d_rraw = ms.out(disp_right_raw, npix);
    float* d_lrc = ms.out(disp_lrc, npix);
    SMT_TRY(ms.check());

    // small per-frame maps
    SMT_ALLOC(ctx, gray8, 2 * npix);
    SMT_ALLOC(ctx, grayf, npix * sizeof(float));
    SMT_ALLOC(ctx, maps, census_maps_bytes(H, W));
    SMT_ALLOC(ctx, armsL, npix * 8);
    SMT_ALLOC(ctx, armsR, npix * 8);
    SMT_ALLOC(ctx, arms_scratch, 2 * npix * sizeof(uint32_t));
    SMT_ALLOC(ctx, dispL, npix * sizeof(float));
    SMT_ALLOC(ctx, dispR, npix * sizeof(float));
    SMT_ALLOC(ctx, counts, 2 * sizeof(int));
    // volumes
    SMT_ALLOC(ctx, volA, nvol * sizeof(float));
    SMT_ALLOC(ctx, volB, nvol * sizeof(float));

    if (!gl) {
        SMT_TRY(k_bgr_to_gray(ctx, bl, npix, gray8.as<uint8_t>()));
        gl = gray8.as<uint8_t>();
    }
    if (!gr) {
        SMT_TRY(k_bgr_to_gray(ctx, br, npix, gray8.as<uint8_t>() + npix));
        gr = gray8.as<uint8_t>() + npix;
    }
    SMT_TRY(k_u8_to_f32(ctx, gl, npix, grayf.as<float>()));
    SMT_TRY(k_ensure_lut(ctx, p->sigmaC, p->sigmaS));


    auto arms_of = [&](DevBuf& b, uint8_t** arms, uint16_t** s0, uint16_t** s1) {
        *arms = b.as<uint8_t>();
        *s0 = reinterpret_cast<uint16_t*>(b.as<uint8_t>() + npix * 4);
        *s1 = reinterpret_cast<uint16_t*>(b.as<uint8_t>() + npix * 6);
    };
    uint8_t *aL, *aR;
    uint16_t *s0L, *s1L, *s0R, *s1R;
    arms_of(armsL, &aL, &s0L, &s1L);
    arms_of(armsR, &aR, &s0R, &s1R);
    // arm construction is image-sized and latency-bound: it runs on the side stream beside the census maps and the cost
    // kernels (right view first: its aggregation comes first); stage timing keeps everything on one stream
    const bool overlap = !ctx->timing;
    if (overlap) {
        SideStream side(ctx);
        smt_status_t st = k_cross_arms(ctx, br, H, W, p->L1, p->L2, p->t1, p->t2, aR, s0R, s1R, arms_scratch.as<uint32_t>());
        side.mark(0);
        if (st == SMT_OK) st = k_cross_arms(ctx, bl, H, W, p->L1, p->L2, p->t1, p->t2, aL, s0L, s1L, arms_scratch.as<uint32_t>() + npix);
        side.mark(1);
        if (st != SMT_OK) return st;
    }

    CensusMaps cm;
    census_maps_carve(maps.p, H, W, &cm);
    SMT_TRY(k_census_maps(ctx, gl, gr, H, W, cm));
    if (!overlap) {
        SMT_TRY(k_cross_arms(ctx, br, H, W, p->L1, p->L2, p->t1, p->t2, aR, s0R, s1R));
        SMT_TRY(k_cross_arms(ctx, bl, H, W, p->L1, p->L2, p->t1, p->t2, aL, s0L, s1L));
    }

    // ---- right view: cost -> aggregate -> WTA
    float* res = nullptr;
    SMT_TRY(k_adcensus_cost(ctx, gl, gr, cm, H, W, D, 1, volA.as<float>(), nullptr, nullptr));
    if (overlap) SideStream::join(ctx, 0);
    // the right view's only consumer is a winner-take-all: the last aggregation pass reduces in the kernel
    SMT_TRY(k_aggregate_cross(ctx, volA.as<float>(), volB.as<float>(), aR, s0R, s1R, H, W, D, p->agg_iters, p->sum_mode,
                              2.0f /* AD-Census cost < 2 (AD-Census.h:287-289) */, &res, dispR.as<float>()));

    // ---- left view: cost -> aggregate -> 4-path scanline with fused WTA
    SMT_TRY(k_adcensus_cost(ctx, gl, gr, cm, H, W, D, 0, volA.as<float>(), nullptr, nullptr));
    if (overlap) SideStream::join(ctx, 1);
    SMT_TRY(k_aggregate_cross(ctx, volA.as<float>(), volB.as<float>(), aL, s0L, s1L, H, W, D, p->agg_iters, p->sum_mode, 2.0f /* AD-Census cost < 2 (AD-Census.h:287-289) */, &res));
    float* scratch = (res == volA.as<float>()) ? volB.as<float>() : volA.as<float>();
    SMT_TRY(k_scanline(ctx, res, grayf.as<float>(), H, W, D, p->p1, p->p2, nullptr, nullptr, nullptr, nullptr, nullptr,
                       dispL.as<float>(), scratch, nullptr));

    // ---- post-processing
    if (d_lraw) SMT_CUDA(ctx, cudaMemcpyAsync(d_lraw, dispL.p, npix * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_rraw) SMT_CUDA(ctx, cudaMemcpyAsync(d_rraw, dispR.p, npix * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    const bool refine = p->irv_iters > 0 || p->do_interpolation || p->do_discontinuity_adjustment;
    SMT_ALLOC(ctx, cls, refine ? npix : 0);
    SMT_TRY(k_lrcheck(ctx, dispL.as<float>(), dispR.as<float>(), W, H, (float)p->gate, false, nullptr, nullptr, nullptr,
                      counts.as<int>(), refine ? cls.as<uint8_t>() : nullptr));
    if (d_lrc) SMT_CUDA(ctx, cudaMemcpyAsync(d_lrc, dispL.p, npix * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    // ---- optional multi-step refinement (fixed specification, see refine.cu): voting over the left view's support
    // regions, interpolation along 16 directions, adjustment at disparity edges with the aggregated left volume
    if (p->irv_iters > 0) SMT_TRY(k_refine_region_voting(ctx, dispL.as<float>(), cls.as<uint8_t>(), aL, H, W, D, p->irv_ts, p->irv_th, p->irv_iters));
    if (p->do_interpolation) SMT_TRY(k_refine_interpolate(ctx, dispL.as<float>(), cls.as<uint8_t>(), bl, H, W, D));
    if (p->do_discontinuity_adjustment) SMT_TRY(k_refine_discontinuity(ctx, dispL.as<float>(), res, H, W, D, p->discontinuity_thres));
    SMT_TRY(k_remove_speckles(ctx, dispL.as<float>(), W, H, p->speckle_diff, p->speckle_min_area, p->speckle_invalid));
    SMT_TRY(k_median(ctx, dispL.as<float>(), d_out, W, H, p->median_size));
    {
        StageTimer tc(ctx, ST_COPY);
        return ms.finish();
    }
}
