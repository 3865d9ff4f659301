/* smt_b200.h -- C ABI of the B200-native stereo-matching engine (libsmt_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The reference
 * (Asherchi/Stereo_Match_Traditional) has no FFI of its own; its "API" is a set of classes and free
 * functions on raw host arrays.  Each entry point below names the reference interface it replaces
 * (file:line under the reference tree); include/smt_b200.hpp re-creates those classes on top of this
 * ABI, and INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - Volumes are float32 [rows][cols][D], D innermost: idx = (i*cols + j)*D + d   (AD-Census.h:87).
 *   - Images are row-major; BGR images are interleaved uint8 [rows][cols][3]      (cross_aggregator.cpp:138).
 *   - Every function returns smt_status_t; nothing throws or aborts across the ABI.
 *   - `space` says where ALL pointer arguments of that call live: SMT_HOST (the reference's convention;
 *     the call copies in, runs on the context's stream, copies out and synchronises) or SMT_DEVICE
 *     (pointers are CUDA device pointers on the context's device; the call only enqueues work on the
 *     context's stream -- use smt_sync()).
 *   - One smt_ctx per (device, stream).  Calls on one context are serialised by the caller; contexts are
 *     independent (one per GPU for batched frame streams).
 *   - There is no CPU fallback: without a CUDA device smt_create() fails with SMT_ERR_CUDA.
 */
#ifndef SMT_B200_H_
#define SMT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SMT_API __declspec(dllexport)
#else
#define SMT_API __attribute__((visibility("default")))
#endif

typedef struct smt_ctx smt_ctx;

typedef int32_t smt_status_t;
enum {
    SMT_OK = 0,
    SMT_ERR_INVALID_ARG = 1,  /* null pointer, non-positive size, ...                                   */
    SMT_ERR_UNSUPPORTED = 2,  /* shape/parameter outside what the reference itself defines (see docs)    */
    SMT_ERR_OOM = 3,          /* device allocation failed                                                */
    SMT_ERR_CUDA = 4,         /* CUDA runtime error; smt_last_error_string() has the text                */
    SMT_ERR_INPUT_RANGE = 5   /* float image is not integer-valued in 0..255 (AD-CensusV1/main.cpp:50-53) */
};

typedef int32_t smt_memspace_t;
enum { SMT_HOST = 0, SMT_DEVICE = 1 };

/* Summation order of the float32 aggregation passes. */
typedef int32_t smt_sum_mode_t;
enum {
    SMT_SUM_REFERENCE_ORDER = 0, /* per-cell sequential adds in the reference's tap order: bit-exact          */
    SMT_SUM_FAST = 1             /* shared-memory line scans, window sums rounded once: <=1e-5 relative    */
};

/* ------------------------------------------------------------------------------------------------ context */
SMT_API int32_t smt_abi_version(void);
/* stream: a cudaStream_t (as void*) to enqueue on, or NULL to let the context create its own. */
SMT_API smt_status_t smt_create(int32_t device, void* stream, smt_ctx** out);
SMT_API void smt_destroy(smt_ctx* ctx);
SMT_API smt_status_t smt_sync(smt_ctx* ctx);
/* The context's cudaStream_t (the one given to smt_create, or its own non-blocking stream): callers that produce or
 * consume SMT_DEVICE buffers on another stream order the two with events on this handle. */
SMT_API void* smt_get_stream(const smt_ctx* ctx);
SMT_API const char* smt_last_error_string(const smt_ctx* ctx);
SMT_API const char* smt_status_string(smt_status_t st);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
SMT_API uint64_t smt_kernel_launches(const smt_ctx* ctx);
/* CUDA-event time (ms) accumulated per pipeline stage since the last reset; stage names via smt_stage_name */
SMT_API int32_t smt_num_stages(void);
SMT_API const char* smt_stage_name(int32_t stage);
SMT_API smt_status_t smt_stage_timing_enable(smt_ctx* ctx, int32_t enable);
SMT_API smt_status_t smt_stage_timing_read(smt_ctx* ctx, double* ms_per_stage, int32_t reset);
/* free cached device workspace */
SMT_API smt_status_t smt_trim(smt_ctx* ctx);
/* plain device memory helpers so a host language needs nothing but this library */
SMT_API smt_status_t smt_device_alloc(smt_ctx* ctx, size_t bytes, void** out);
SMT_API smt_status_t smt_device_free(smt_ctx* ctx, void* p);
SMT_API smt_status_t smt_memcpy_h2d(smt_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
SMT_API smt_status_t smt_memcpy_d2h(smt_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ------------------------------------------------------------------------------------------------ AD-Census cost
 * Replaces AD_Census::{Initialize, ComputeAD[Right], ComputeCensus9x7[Right], ComputeADcensus[Right]}
 * (AD-CensusV1/AD-Census.h:322, :75/:103, :142/:207, :271/:296).
 * left/right: float32 [rows][cols] holding integer grey values 0..255 (main.cpp:50-53).
 * view 0 = left cost (match at j-d), 1 = right cost (match at j+d).
 * cost (required), ad, census (optional, may be NULL): float32 [rows][cols][D].
 */
SMT_API smt_status_t smt_adcensus_cost(smt_ctx* ctx, const float* left, const float* right, int32_t rows,
                                       int32_t cols, int32_t D, float sigmaC, float sigmaS, int32_t view,
                                       float* cost, float* ad, float* census, smt_memspace_t space);

/* The packed 63-bit 9x7 census string of every pixel of one image, in the reference's bit order: the tap
 * (r,c), r=-4..4 outer, c=-3..3 inner, is shifted in MSB-first, so tap (-4,-3) is bit 62 and (4,3) bit 0;
 * out-of-image taps contribute 0 (AD-Census.h:167-191).  img: float32 [rows][cols]; out: uint64 [rows][cols]. */
SMT_API smt_status_t smt_census9x7_u64(smt_ctx* ctx, const float* img, int32_t rows, int32_t cols, uint64_t* out,
                                       smt_memspace_t space);

/* First-minimum winner-take-all over d, disparity stored as float.  Replaces AD_Census::WTA (AD-Census.h:346),
 * CrossArmAggregation::WTA (CrossArm.cpp:33), ScanlineOptimizer::WTA (ScanlineOptimizer.h:40) and
 * ComputeDispOringin (CBLSM/CBLSM.h:383). */
SMT_API smt_status_t smt_wta_first_min(smt_ctx* ctx, const float* vol, int32_t rows, int32_t cols, int32_t D,
                                       float* disp, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ CrossAggregator
 * Replaces CrossAggregator::{Initialize, SetData, SetParams, BuildArms, ComputeSupPixelCount, Aggregate}
 * (CBLSM/cross_aggregator.cpp:19-118, :135-394).
 * bgr: uint8 [rows][cols][3] (arms come from this image only, :138,:206).
 * arms_out (optional): uint8 [rows][cols][4] = {left,right,top,bottom} (cross_aggregator.h:17-20);
 * sup_hfirst / sup_vfirst (optional): uint16 [rows][cols] (vec_sup_count_[0], [1]).
 */
SMT_API smt_status_t smt_cross_arms(smt_ctx* ctx, const uint8_t* bgr, int32_t rows, int32_t cols, int32_t L1,
                                    int32_t L2, int32_t t1, int32_t t2, uint8_t* arms_out, uint16_t* sup_hfirst,
                                    uint16_t* sup_vfirst, smt_memspace_t space);
SMT_API smt_status_t smt_aggregate_cross(smt_ctx* ctx, const uint8_t* bgr, const float* cost_init, int32_t rows,
                                         int32_t cols, int32_t D, int32_t L1, int32_t L2, int32_t t1, int32_t t2,
                                         int32_t iters, smt_sum_mode_t mode, float* cost_out, uint8_t* arms_out,
                                         uint16_t* sup_hfirst, uint16_t* sup_vfirst, smt_memspace_t space);

/* Row-band form of smt_aggregate_cross for one very large frame split over several GPUs
 * (BASELINE.json configs[4]): the 2*iters one-dimensional passes are exposed one by one so the host can refresh the
 * 34-row halos between them (a vertical pass reads 34 rows beyond the band; horizontal passes are row-local).
 * vol_in / vol_out: float32 [rows][cols][D] band buffers INCLUDING their halo rows; arms / sup_*: the same rows of the
 * FULL image's maps (smt_cross_arms on the whole image, then sliced).  Results in halo rows are unspecified.
 * mode: SMT_SUM_REFERENCE_ORDER makes the bands bit-identical to the whole frame (and to the reference); SMT_SUM_FAST
 * keeps every window sum exact up to one rounding, so bands agree with the whole frame to ~1e-6.
 * SMT_SUM_FAST supports arms <= 34 (the reference's cross_L1 default, adcensus_types.h:69): sweep 0 scans the arm map
 * and returns SMT_ERR_UNSUPPORTED when it holds a longer arm.  Sweeps of one aggregation must be issued in order,
 * starting with sweep 0, on the same context.
 * Device memspace only.  smt_cross_sweep_info reports the direction (1 = along rows) and whether the pass divides. */
SMT_API int32_t smt_cross_sweep_count(int32_t iters);
SMT_API smt_status_t smt_cross_sweep_info(int32_t iters, int32_t sweep, int32_t* horizontal, int32_t* divides);
SMT_API smt_status_t smt_aggregate_cross_sweep(smt_ctx* ctx, int32_t iters, int32_t sweep, const float* vol_in,
                                               float* vol_out, const uint8_t* arms, const uint16_t* sup_hfirst,
                                               const uint16_t* sup_vfirst, int32_t rows, int32_t cols, int32_t D,
                                               float vmax, smt_sum_mode_t mode, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ multi-step refinement
 * Iterative region voting, proper interpolation and depth-discontinuity adjustment (Mei et al. 2011, section 2.4), which
 * BASELINE.json's north_star names.  PARITY UNPINNED: the reference implements none of them -- only the option fields
 * irv_ts, irv_th, do_discontinuity_adjustment exist (CBLSM/adcensus_types.h:58-59,65) and nothing reads them -- so the
 * specification is fixed in oracle/port/oracle_port.c (orc_refine_*) and these entry points are bit-identical to that.
 * disp: float32 [rows][cols], +inf = invalid (what LeftRightConsistency writes, PostProcessing.h:93,131), updated in place.
 * cls: uint8 [rows][cols], 0 valid / 1 occlusion / 2 mismatch, updated in place (filled pixels become 0);
 *   smt_invalid_class_map builds it from LeftRightConsistency's two (row, col) lists.
 * smt_refine_region_voting: `iters` synchronous iterations; an invalid pixel takes the first most frequent valid disparity of
 *   its cross support region (arms: CrossAggregator's {left,right,top,bottom} bytes, cross_aggregator.h:17-20) if the region
 *   holds more than ts valid pixels and that disparity's share exceeds th (defaults of the unused fields: ts 20, th 0.4).
 * smt_refine_interpolate: an invalid pixel looks along 16 directions for the first valid disparity (at most max_len-1 steps);
 *   an occlusion takes the smallest one, a mismatch the one whose BGR colour is closest to its own.
 * smt_refine_discontinuity: on edges of the disparity map (3x3 Sobel, |gx|+|gy| > thres) a pixel takes its left or right
 *   neighbour's disparity when the aggregated cost volume cost [rows][cols][D] rates it lower at this pixel. */
SMT_API smt_status_t smt_invalid_class_map(smt_ctx* ctx, const int32_t* occ, int32_t n_occ, const int32_t* mis, int32_t n_mis,
                                           int32_t rows, int32_t cols, uint8_t* cls, smt_memspace_t space);
SMT_API smt_status_t smt_refine_region_voting(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* arms, int32_t rows,
                                              int32_t cols, int32_t D, int32_t ts, float th, int32_t iters, smt_memspace_t space);
SMT_API smt_status_t smt_refine_interpolate(smt_ctx* ctx, float* disp, uint8_t* cls, const uint8_t* bgr, int32_t rows, int32_t cols,
                                            int32_t max_len, smt_memspace_t space);
SMT_API smt_status_t smt_refine_discontinuity(smt_ctx* ctx, float* disp, const float* cost, int32_t rows, int32_t cols, int32_t D,
                                              float thres, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ row-band split
 * One very large frame split into row bands over several GPUs, for the cost and cross-aggregation stages
 * (BASELINE.json configs[4]; the reference has no counterpart: CrossAggregator::Aggregate, CBLSM/cross_aggregator.cpp:89-118,
 * runs on the whole frame and its `int` indices overflow at 7680x4320x512).  Rank g of `world` owns rows
 * [g*ceil(rows/world), ...) and keeps a 34-row halo (the maximum arm length) beyond each interior edge.
 *   smt_band_create      allocates the band: two volumes float32 [top + owned + bottom][cols][D] plus a small header
 *   smt_band_geometry    owned image rows [r0, r1) and halo heights
 *   smt_band_volume      device pointer of volume 0: the caller writes the initial cost of its OWNED rows at row `top`
 *                        (e.g. smt_ad_cost on those image rows) and finds the aggregated cost there afterwards
 *   smt_band_export / smt_band_connect        neighbours in OTHER processes (one process per GPU): exchange the 128-byte
 *                        handles by any host channel, then map the upper / lower neighbour (CUDA IPC, peer access)
 *   smt_band_connect_local                    neighbours in THIS process (contexts on the same or on peer devices)
 *   smt_halo_exchange    one-sided exchange for volume 0: copies the 34 top / bottom owned rows into the neighbours' halo
 *                        rows over NVLink, stores a sequence number in their headers, and queues this context's stream
 *                        behind the neighbours' numbers (a spinning one-warp kernel: no host synchronisation, no collective)
 *   smt_band_aggregate_cross   CrossAggregator::Aggregate(iters) on the band: horizontal passes on the owned rows, one
 *                        exchange before every vertical pass (overlapped with the interior rows of the horizontal pass in
 *                        front of it); arms / sup_*: the band's rows [r0-top, r1+bottom) of the FULL image's maps
 *                        (smt_cross_arms on the whole image, sliced).  SMT_SUM_REFERENCE_ORDER: bit-identical to the whole
 *                        frame and to the reference; SMT_SUM_FAST: within the usual 1e-5, arms <= 34
 *   smt_band_sync        waits for the band's streams; SMT_ERR_CUDA if an exchange timed out; bytes pushed so far
 * All ranks must call smt_band_aggregate_cross with the same iters; calls are asynchronous on the context's stream. */
typedef struct smt_band smt_band;
#define SMT_BAND_HANDLE_BYTES 128
typedef struct smt_band_handle {
    unsigned char bytes[SMT_BAND_HANDLE_BYTES];
} smt_band_handle;
SMT_API smt_status_t smt_band_create(smt_ctx* ctx, int32_t image_rows, int32_t cols, int32_t D, int32_t rank, int32_t world,
                                     smt_band** out);
SMT_API void smt_band_destroy(smt_band* band);
SMT_API smt_status_t smt_band_geometry(const smt_band* band, int32_t* r0, int32_t* r1, int32_t* top, int32_t* bottom);
SMT_API smt_status_t smt_band_volume(smt_band* band, float** vol);
SMT_API smt_status_t smt_band_export(smt_band* band, smt_band_handle* out);
SMT_API smt_status_t smt_band_connect(smt_band* band, const smt_band_handle* upper, const smt_band_handle* lower);
SMT_API smt_status_t smt_band_connect_local(smt_band* band, smt_band* upper, smt_band* lower);
SMT_API smt_status_t smt_halo_exchange(smt_band* band);
SMT_API smt_status_t smt_band_aggregate_cross(smt_band* band, const uint8_t* arms, const uint16_t* sup_hfirst,
                                              const uint16_t* sup_vfirst, int32_t iters, float vmax, smt_sum_mode_t mode);
SMT_API smt_status_t smt_band_sync(smt_band* band, double* pushed_bytes);

/* CBLSM's main() as one call (CBLSM/CBLSM.cpp:64-67,101-104,133-134,146-153): arm maps of both images (tao 25, 34, 17 in the
 * reference), ComputeAD / ComputeADRight, costAggregationV5 on (right volume, right arms) and (left volume, left arms), a
 * second V5 pass on the left volume with the left arms and on the right volume with the LEFT arms again (:150, kept),
 * ComputeDispOringin on both.  gray_*: uint8 [rows][cols]; disp_*: float32 [rows][cols]. */
SMT_API smt_status_t smt_pipeline_cblsm(smt_ctx* ctx, const uint8_t* gray_left, const uint8_t* gray_right, int32_t rows,
                                        int32_t cols, int32_t D, int32_t tao, int32_t max_length, int32_t sec_length,
                                        smt_sum_mode_t mode, float* disp_left, float* disp_right, smt_memspace_t space);

/* SMT_SUM_FAST evaluates window sums as differences of 32-bit wrap-around fixed-point prefixes (quantum = 2^-25 of the
 * power of two above the largest cost).  That is accurate to 2^-20 per pass as long as every non-zero value entering a
 * window is at least 2^19 quanta; lines that hold a smaller value are redone on compensated float32-pair prefixes, and
 * when more than a quarter of a sweep's lines ask for it the rest of the aggregation runs on float pairs altogether.
 * Both routes meet the 1e-5 bound of BASELINE.json; neither exists in the reference, which adds taps one by one
 * (CBLSM/cross_aggregator.cpp:353-389).  This reports what the context's LAST fast aggregation did: the number of
 * (line, 32-disparity chunk) tasks redone on float pairs (0 = pure fixed point), or -1 = float pairs took over.
 * Synchronises the stream. */
SMT_API smt_status_t smt_aggregate_fast_status(smt_ctx* ctx, int32_t* redone_tasks);

/* ------------------------------------------------------------------------------------------------ CrossArm / CBLSM
 * smt_ad_cost replaces CBLSM's ComputeAD / ComputeADRight (CBLSM/CBLSM.h:327, :355) and AD_Census::ComputeAD[Right]
 * (AD-CensusV1/AD-Census.h:75, :103): left/right uint8 [rows][cols], ad float32 [rows][cols][D].
 *
 * smt_arms_gray replaces the four arm-length passes of
 *   variant 0: CrossArmAggregation::Compute{Left,Right,Top,Button}ArmLength (AD-CensusV1/CrossArm.cpp:147-598), run in
 *              main.cpp:69-72 order with ONE member threshold that drops to 6 for good once any arm passes 17 pixels,
 *              and with the Right pass's `col = _row` indexing (:265) -- requires rows <= cols; pass max_length=34,
 *              sec_length=17 (hard-coded in the reference);
 *   variant 1: CBLSM ArmLength{L,R,Up,Down} (CBLSM/CBLSM.h:643, :536, :753, :861): the threshold is a by-value uchar,
 *              so each of the four calls starts again from `tao`.
 * img: uint8 [rows][cols][channels], channels 1 or 3.  Outputs int32 [rows][cols] (any may be NULL).
 * tao_final (optional, HOST pointer in both memspaces): the threshold the reference object is left with (variant 0).
 *
 * smt_aggregate_rect replaces the rectangle-mean aggregations
 *   variant 0: CrossArmAggregation::AggregationVertical (CrossArm.cpp:60)  -- flat slice index, column-major order
 *   variant 1: CrossArmAggregation::Aggregation        (CrossArm.cpp:104) -- flat index, exclusive upper bounds
 *   variant 2: CBLSM costAggregationV5                 (CBLSM/CBLSM.h:1179) -- row-major order
 *   variant 3: CBLSM costAggregationV4                 (CBLSM/CBLSM.h:1128) -- one arm entry per (pixel, d) CELL: the four arm
 *              arrays are int32 [rows][cols][D]; exclusive upper bounds, an empty window gives 0/0 = NaN like the
 *              reference; reference order only (SMT_SUM_FAST falls back to it)
 * CBLSM costAggregationVertical (CBLSM/CBLSM.h:1227) is variant 0 fed with smt_arms_gray variant-1 maps (same loops, same
 * column-major order; those arms never leave the image).
 * Returns SMT_ERR_UNSUPPORTED when the arm maps address cells outside the slice (the reference reads out of bounds
 * there).  SMT_SUM_REFERENCE_ORDER is bit-exact; SMT_SUM_FAST uses a double-single summed-area table.
 */
SMT_API smt_status_t smt_ad_cost(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols,
                                 int32_t D, int32_t view, float* ad, smt_memspace_t space);
SMT_API smt_status_t smt_arms_gray(smt_ctx* ctx, const uint8_t* img, int32_t channels, int32_t rows, int32_t cols,
                                   int32_t tao, int32_t max_length, int32_t sec_length, int32_t variant,
                                   int32_t* arm_left, int32_t* arm_right, int32_t* arm_top, int32_t* arm_bottom,
                                   int32_t* tao_final, smt_memspace_t space);
SMT_API smt_status_t smt_aggregate_rect(smt_ctx* ctx, const float* vol_in, const int32_t* arm_left,
                                        const int32_t* arm_right, const int32_t* arm_top, const int32_t* arm_bottom,
                                        int32_t rows, int32_t cols, int32_t D, int32_t variant, smt_sum_mode_t mode,
                                        float* vol_out, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ ScanlineOptimizer
 * Replaces ScanlineOptimizer::{Initialize, ScanLine, ScanLineLeftRight, ScanLineUpDown, WTA}
 * (AD-CensusV1/ScanlineOptimizer.h:66, :104, :130, :194, :40) including the up/down quirks (:221,:232,:238).
 * cost: float32 [rows][cols][D]; gray: float32 [rows][cols]; p1, p2 are ints as in the reference (:66).
 * Optional outputs (NULL to skip): the four path volumes, their sum (_ProcessedVolume) and the WTA disparity.
 */
SMT_API smt_status_t smt_scanline(smt_ctx* ctx, const float* cost, const float* gray, int32_t rows, int32_t cols,
                                  int32_t D, int32_t p1, int32_t p2, float* left, float* right, float* up,
                                  float* down, float* sum, float* disp, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ post-processing
 * smt_lrcheck replaces LeftRightConsistency (AD-CensusV1/PostProcessing.h:72): disp_left is modified in place
 * (+inf at invalidated pixels); occ / mis receive (row, col) int32 pairs in the reference's row-major emission
 * order (capacity rows*cols pairs each, may be NULL), n_occ / n_mis their counts (host pointers in both spaces).
 * smt_lrcheck_and replaces LeftAndRightConsistency (:10): writes 0 / d into `last`, `>=` float gate.
 */
SMT_API smt_status_t smt_lrcheck(smt_ctx* ctx, float* disp_left, const float* disp_right, int32_t cols,
                                 int32_t rows, int32_t gate, int32_t* occ, int32_t* n_occ, int32_t* mis,
                                 int32_t* n_mis, smt_memspace_t space);
SMT_API smt_status_t smt_lrcheck_and(smt_ctx* ctx, const float* disp_left, const float* disp_right, float* last,
                                     int32_t cols, int32_t rows, float gate, int32_t* occ, int32_t* n_occ,
                                     int32_t* mis, int32_t* n_mis, smt_memspace_t space);
/* Replaces RemoveSpeckles (PostProcessing.h:250).  invalid_val is the reference's *int* parameter; the shipped
 * call passes +inf, which g++ folds to 2147483647 (main.cpp:93) -- pass that to reproduce it. */
SMT_API smt_status_t smt_remove_speckles(smt_ctx* ctx, float* disp, int32_t cols, int32_t rows, int32_t diff,
                                         uint32_t min_area, int32_t invalid_val, smt_memspace_t space);
/* Replaces FillTheHole (PostProcessing.h:156) AS WRITTEN: the function swaps the axes (width = row, height = col), so a
 * list entry (y, x) addresses disp[y*row + x] with y < col, x < row; "invalid" is the value 65535, not +inf.  occ:
 * n_occ (y, x) pairs; mis: in/out, capacity row*col pairs, *n_mis in/out (host pointer in both spaces): the third phase
 * replaces the list by all 65535-valued pixels like the reference's vector assignment (:188).  Entries outside that
 * frame return SMT_ERR_INVALID_ARG (the reference would write out of bounds).
 * variant SMT_FILL_FIXED is the evident intent (no reference implementation exists for it): width = col, height = row,
 * entries are the (row, col) pairs smt_lrcheck emits, invalid = +inf as smt_lrcheck writes; otherwise identical. */
typedef enum { SMT_FILL_AS_WRITTEN = 0, SMT_FILL_FIXED = 1 } smt_fill_variant_t;
SMT_API smt_status_t smt_fill_the_hole(smt_ctx* ctx, float* disp, int32_t row, int32_t col, int32_t disp_range,
                                       const int32_t* occ, int32_t n_occ, int32_t* mis, int32_t* n_mis,
                                       int32_t variant, smt_memspace_t space);
/* Replaces MedianFilter (PostProcessing.h:314): in-bounds taps only, element sorted[n/2]; odd wnd_size <= 9. */
SMT_API smt_status_t smt_median(smt_ctx* ctx, const float* in, float* out, int32_t cols, int32_t rows,
                                int32_t wnd_size, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ window matchers
 * smt_sad replaces GetPointDepthLeft / GetPointDepthRight with sadvalue, OptimalDisparity, GetMinSadIndex
 * (SAD/Sad.h:96, :141, :15, :40, :22).  Inputs are the replicate-padded grey images [rows_padded][cols_padded],
 * pad = winsize+1 per side (SADmain.cpp:47-48); disp: int32 [rows_padded-2*pad][cols_padded-2*pad] (CV_32S);
 * costs (optional): float32 [..][..][D], the per-disparity SAD vector incl. the copy-previous rule (:125-129).
 * subpixel (optional, left view only): float32 [..][..], the parabola-refined disparity OptimalDisparity computes and then
 *   discards (Sad.h:76-84 -- it returns best_disp); 0 wherever the function returns 0.
 * smt_sad_crosscheck replaces CrossCheckDiaparity on the CV_32S maps (Sad.h:184-221): out int32 [rows][cols] (invalid =
 *   2147483647, what (int)Invalid_Float is under g++), occ / mis: (row, col) pairs in raster order, capacity rows*cols each.
 *   The reference indexes right[j - lv] unchecked; the flat element i*cols + j - lv is read whenever it lies inside the map.
 * smt_ncc_shift replaces ncc() (NCC/NCC.h:117-272), the shifted-image NCC with a clipped 11x11 window and offsets 1..79:
 *   depth uint8 [rows][cols] = 3 * the best offset (0 if none); right_view selects its "right" branch, add_constant the
 *   `right += 10` of its last argument.  Needs cols > 79.
 *
 * smt_ncc replaces NCC_algorithem with ComputeCost and WinTakeAll (NCC/NCC.h:69, :15, :53): unpadded grey images
 * [rows][cols]; disp int32 [rows][cols], 0 outside the interior.  costs (optional): float64 [rows][cols][D], the value
 * ComputeCost returns per disparity (double in the reference, :15-49): 255 where the window leaves the image (:88), NaN for
 * a zero-variance window, 0 at border pixels (which the reference never visits).
 *
 * smt_asw replaces AdaptiveSupportWeight / AdaptiveSupportWeightRight with getGausssianMask, getColorMask,
 * bilateralfiterWight, WinTakeAll (ASW/ASW.h:329, :382, :16, :41, :210, :193): replicate-padded grey images, pad =
 * win_size+1 (ASWeight.cpp:54-55); disp float32 [rows_padded-2*pad][cols_padded-2*pad].  In the right view the last
 * win_size+2 output columns are undefined in the reference (it reads costVolume[-1], ASW.h:401,:424); they are 0 here.
 * costs (optional): float32 [..][..][D], the reference's costVolume row of each pixel (ASW.h:358-371, float of the double
 * quotient) incl. the copy-previous rule for disparities whose window leaves the image.
 * smt_asw_crosscheck replaces CrossCheckDiaparity (ASW.h:108): out uint8 [rows][cols].
 */
SMT_API smt_status_t smt_sad(smt_ctx* ctx, const uint8_t* left_padded, const uint8_t* right_padded, int32_t rows_padded,
                             int32_t cols_padded, int32_t D, int32_t winsize, int32_t right_view, int32_t* disp,
                             float* costs, float* subpixel, smt_memspace_t space);
SMT_API smt_status_t smt_sad_crosscheck(smt_ctx* ctx, const int32_t* disp_left, const int32_t* disp_right, int32_t rows,
                                        int32_t cols, int32_t* out, int32_t* occ, int32_t* n_occ, int32_t* mis, int32_t* n_mis,
                                        smt_memspace_t space);
SMT_API smt_status_t smt_ncc_shift(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols,
                                   int32_t right_view, int32_t add_constant, uint8_t* depth, smt_memspace_t space);
SMT_API smt_status_t smt_ncc(smt_ctx* ctx, const uint8_t* left, const uint8_t* right, int32_t rows, int32_t cols,
                             int32_t win_size, int32_t D, int32_t* disp, double* costs, smt_memspace_t space);
SMT_API smt_status_t smt_asw(smt_ctx* ctx, const uint8_t* left_padded, const uint8_t* right_padded, int32_t rows_padded,
                             int32_t cols_padded, int32_t win_size, int32_t D, double space_sigma, double color_sigma,
                             int32_t T, int32_t right_view, float* disp, float* costs, smt_memspace_t space);
SMT_API smt_status_t smt_asw_crosscheck(smt_ctx* ctx, const float* disp_left, const float* disp_right, int32_t rows,
                                        int32_t cols, uint8_t* out, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ pre / post (drivers)
 * What the reference's main()s do with OpenCV around the matchers:
 *   smt_bgr_to_gray          cvtColor(BGR2GRAY) on 8-bit images (fixed point 1868/9617/4899 >> 14)     main.cpp:19-20
 *   smt_replicate_pad        copyMakeBorder(BORDER_REPLICATE); out is [rows+2*pad][cols+2*pad]          SADmain.cpp:47-48
 *   smt_u8_to_f32            the uchar -> float image copies                                            main.cpp:46-55
 *   smt_transform_to_show    TransformToShow (PostProcessing.h:137): (uchar)(int)disp, out-of-int-range floats -> 0
 *   smt_normalize_minmax_u8  normalize(src, dst, 0, 255, NORM_MINMAX) followed by convertTo(CV_8U)     main.cpp:106-111
 */
/* cv::imwrite("*.png", 8-bit image) as the drivers use it for their results (AD-CensusV1/main.cpp:112-114,
 * ASW/ASWeight.cpp:75-92): img uint8 [rows][cols][channels], channels 1 (grey) or 3 (BGR, written as RGB).  The pixel
 * data is exact; the file is a valid PNG with stored deflate blocks, not byte-identical to OpenCV's libpng output. */
SMT_API smt_status_t smt_write_png_u8(smt_ctx* ctx, const char* path, const uint8_t* img, int32_t rows, int32_t cols,
                                      int32_t channels, smt_memspace_t space);
SMT_API smt_status_t smt_bgr_to_gray(smt_ctx* ctx, const uint8_t* bgr, int32_t rows, int32_t cols, uint8_t* gray,
                                     smt_memspace_t space);
SMT_API smt_status_t smt_replicate_pad(smt_ctx* ctx, const uint8_t* img, int32_t rows, int32_t cols, int32_t pad,
                                       uint8_t* out, smt_memspace_t space);
SMT_API smt_status_t smt_u8_to_f32(smt_ctx* ctx, const uint8_t* img, size_t count, float* out, smt_memspace_t space);
SMT_API smt_status_t smt_transform_to_show(smt_ctx* ctx, const float* disp, int32_t rows, int32_t cols, uint8_t* out,
                                           smt_memspace_t space);
SMT_API smt_status_t smt_normalize_minmax_u8(smt_ctx* ctx, const float* src, int32_t rows, int32_t cols, uint8_t* out,
                                             smt_memspace_t space);
/* The ASW driver's display chain on the 8-bit cross-checked map (ASW/ASWeight.cpp:66-78).  normalize, filterSpeckles and
 * medianBlur are OpenCV functions (the reference links a prebuilt OpenCV; no sources in its tree): they follow OpenCV's
 * behaviour and are pinned to cv2 4.13 by tests/golden/display.npz.
 *   smt_normalize_minmax_u8_from_u8   normalize(m, m, 0, 255, NORM_MINMAX) on CV_8U (+ the no-op convertTo)   :69, :72
 *   smt_filter_speckles_u8            cv::filterSpeckles(m, newVal, maxSpeckleSize, maxDiff) on CV_8UC1        :73
 *                                     (4-connected components of |a-b| <= maxDiff among pixels != newVal; <= size -> newVal)
 *   smt_median_blur_u8                cv::medianBlur(m, m, ksize), ksize 3 or 5, BORDER_REPLICATE              :74, :78
 *   smt_fill_image_new                FillImageNew (ASW/ASW.h:434) AS WRITTEN: a zero pixel takes the first non-zero value
 *                                     at columns c-1, c-3, c-6, ... else at c+1, c+3, ...; the reference also reads column
 *                                     == cols (the next row's first pixel; below the last row one byte past the image,
 *                                     which the caller supplies as past_end) and, when a search ends exactly there on a
 *                                     zero, records no value, which shifts the values of all later pixels by one.  The
 *                                     last n_unassigned zero pixels (host pointer, optional) read past the reference's
 *                                     value vector; they are left 0 here.  cols <= 65535.
 *   smt_asw_display_chain             everything after the two matcher calls: CrossCheckDiaparity, the four steps above
 *                                     with the reference's constants (0, 40, 2; 5; 3); outputs are the three images the
 *                                     driver writes ("speckle.png", "filled.png", "medain.png"), the first two optional */
SMT_API smt_status_t smt_normalize_minmax_u8_from_u8(smt_ctx* ctx, const uint8_t* src, int32_t rows, int32_t cols, uint8_t* out,
                                                     smt_memspace_t space);
SMT_API smt_status_t smt_filter_speckles_u8(smt_ctx* ctx, uint8_t* img, int32_t rows, int32_t cols, int32_t new_val,
                                            int32_t max_speckle_size, int32_t max_diff, smt_memspace_t space);
SMT_API smt_status_t smt_median_blur_u8(smt_ctx* ctx, const uint8_t* in, uint8_t* out, int32_t rows, int32_t cols, int32_t ksize,
                                        smt_memspace_t space);
SMT_API smt_status_t smt_fill_image_new(smt_ctx* ctx, uint8_t* disp, int32_t rows, int32_t cols, int32_t past_end,
                                        int32_t* n_unassigned, smt_memspace_t space);
SMT_API smt_status_t smt_asw_display_chain(smt_ctx* ctx, const float* disp_left, const float* disp_right, int32_t rows, int32_t cols,
                                           uint8_t* speckle_img, uint8_t* filled_img, uint8_t* median_img, smt_memspace_t space);

/* ------------------------------------------------------------------------------------------------ full pipeline
 * P_full: the composition of the reference entry points above with the reference's own
 * default parameters (main.cpp:24-30, adcensus_types.h:69-70).
 */
typedef struct smt_adcensus_params {
    int32_t rows, cols, D;
    float sigmaC, sigmaS;          /* main.cpp:25-26,58 : 10, 30          */
    int32_t L1, L2, t1, t2;        /* adcensus_types.h:69-70 : 34,17,20,6 */
    int32_t agg_iters;             /* CBLSM.cpp:141 : 4                   */
    int32_t p1, p2;                /* main.cpp:28-29 : 10, 150            */
    int32_t gate;                  /* main.cpp:30 : 2                     */
    int32_t speckle_diff;          /* main.cpp:93 : 1                     */
    uint32_t speckle_min_area;     /* main.cpp:93 : 30                    */
    int32_t speckle_invalid;       /* main.cpp:93 : +inf -> 2147483647    */
    int32_t median_size;           /* main.cpp:94 : 3                     */
    smt_sum_mode_t sum_mode;
    /* multi-step refinement between the LR check and the speckle filter (see smt_refine_*; NOT part of the reference's
     * pipeline, all off by default): */
    int32_t irv_iters;             /* iterative region voting iterations (0 = off; Mei et al. use 5)         */
    int32_t irv_ts;                /* adcensus_types.h:58 : 20                                               */
    float irv_th;                  /* adcensus_types.h:59 : 0.4                                              */
    int32_t do_interpolation;      /* proper interpolation of what region voting left invalid                */
    int32_t do_discontinuity_adjustment; /* adcensus_types.h:65                                              */
    float discontinuity_thres;     /* Sobel threshold on the disparity map: 5                                */
} smt_adcensus_params;
SMT_API void smt_adcensus_default_params(smt_adcensus_params* p, int32_t rows, int32_t cols, int32_t D);
/* bgr_*: uint8 [rows][cols][3]; gray_*: uint8 [rows][cols].
 * Outputs (all optional except disp_out), float32 [rows][cols]:
 *   disp_out        after LR check (+ the optional refinement steps) + speckle removal + median
 *   disp_left_raw   ScanlineOptimizer::WTA of the left view
 *   disp_right_raw  first-min WTA of the aggregated right volume
 *   disp_lrc        left disparity after LeftRightConsistency (+inf = invalidated)
 */
SMT_API smt_status_t smt_pipeline_adcensus_full(smt_ctx* ctx, const smt_adcensus_params* p, const uint8_t* bgr_left,
                                                const uint8_t* bgr_right, const uint8_t* gray_left,
                                                const uint8_t* gray_right, float* disp_out, float* disp_left_raw,
                                                float* disp_right_raw, float* disp_lrc, smt_memspace_t space);

#ifdef __cplusplus
}
#endif
#endif /* SMT_B200_H_ */
