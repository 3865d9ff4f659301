// ORACLE (test infrastructure, never shipped, never on the product path).
//
// One translation unit that #includes the UNMODIFIED reference sources where they lie
// under /root/reference (include root passed with -I by oracle/Makefile), each app in its
// own namespace because the five apps redefine each other's symbols and define non-inline
// functions in headers.  It exports a flat C API (ref_*) over raw arrays so the parity
// tests, oracle/make_golden.py and bench.py's cpu_baseline leg can drive the reference's own
// code.  Output: oracle/_ref/libsmt_ref.so (git-ignored, travels to the GPU box).
//
// Nothing here re-implements the reference: every ref_* body only marshals arrays into the
// reference's classes / free functions and copies results out.  The single exception is
// ref_sad_costs_*, which replays the reference's d-loop (Sad.h:122-132 / :164-172) around the
// reference's own sadvalue() because the reference never exposes the per-disparity SAD vector.
#include <opencv2/opencv.hpp>
#include <algorithm>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#else
// ASW.h includes <omp.h> unconditionally; provide it via -I oracle/cvshim when OpenMP is off.
#endif

#define private public  // read intermediates (arm maps, per-path volumes, support counts)
#define protected public

namespace rns_adc {
#include "AD-CensusV1/AD-Census.h"
#include "AD-CensusV1/CrossArm.h"
#include "AD-CensusV1/ScanlineOptimizer.h"
#include "AD-CensusV1/PostProcessing.h"
#include "AD-CensusV1/CrossArm.cpp"
}  // namespace rns_adc

namespace rns_cblsm {
#include "CBLSM/adcensus_types.h"
#include "CBLSM/cross_aggregator.h"
#include "CBLSM/CBLSM.h"
#include "CBLSM/cross_aggregator.cpp"
}  // namespace rns_cblsm

namespace rns_sad {
#include "SAD/Sad.h"
}

namespace rns_ncc {
#include "NCC/NCC.h"
}

namespace rns_asw {
#include "ASW/ASW.h"
}

#undef private
#undef protected

namespace {
// The reference prints progress lines from inside its hot loops (Sad.h:136, NCC.h:93, ASW.h:376).
struct CoutSilencer {
    std::streambuf* saved;
    CoutSilencer() : saved(std::cout.rdbuf(nullptr)) {}
    ~CoutSilencer() { std::cout.rdbuf(saved); }
};
cv::Mat wrap_u8(const uint8_t* p, int rows, int cols, int ch = 1) {
    return cv::Mat(rows, cols, ch == 3 ? cv::CV_8UC3 : cv::CV_8UC1, const_cast<uint8_t*>(p));
}
template <typename T> void free_arr(T*& p) { delete[] p; p = nullptr; }
void write_pairs(const std::vector<std::pair<int, int>>& v, int32_t* out, int32_t* n) {
    if (n) *n = (int32_t)v.size();
    if (out) for (size_t k = 0; k < v.size(); ++k) { out[2 * k] = v[k].first; out[2 * k + 1] = v[k].second; }
}
}  // namespace

extern "C" {

int ref_abi_version() { return 1; }
int ref_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------------------------------------------------------- AD_Census (AD-Census.h)
// Any output pointer may be null.  Volumes are [H][W][D] f32.
void ref_adcensus(const float* left, const float* right, int H, int W, int D, float sigmaC, float sigmaS,
                  int do_left, int do_right, float* costL, float* costR, float* adL, float* adR,
                  float* censusL, float* censusR, float* wtaL, float* wtaR) {
    rns_adc::AD_Census adc;
    cv::Mat dummyL, dummyR;
    adc.Initialize(const_cast<float*>(left), const_cast<float*>(right), D, H, W, dummyL, dummyR, sigmaC, sigmaS);
    size_t n = (size_t)H * W * D;
    if (do_left) adc.ComputeADcensus();
    if (do_right) adc.ComputeADcensusRight();
    if (wtaL && wtaR) adc.WTA(wtaL, wtaR);
    if (costL) std::memcpy(costL, adc.GetPtrLeft(), n * 4);
    if (costR) std::memcpy(costR, adc.GetPtrRight(), n * 4);
    if (adL) std::memcpy(adL, adc.ADcostVolum, n * 4);
    if (adR) std::memcpy(adR, adc.ADcostVolumRight, n * 4);
    if (censusL) std::memcpy(censusL, adc.CensusVolum, n * 4);
    if (censusR) std::memcpy(censusR, adc.CensusVolumRight, n * 4);
    // the reference never frees (AD-Census.h:382); the harness does
    free_arr(adc.__leftDisp__); free_arr(adc.__rightDisp__); free_arr(adc.ADcostVolum); free_arr(adc.ADcostVolumRight);
    free_arr(adc.CensusVolum); free_arr(adc.CensusVolumRight); free_arr(adc.costVolume); free_arr(adc.costVolumeRight);
}

// first-minimum WTA (CrossArm.cpp:33 -- same rule as AD-Census.h:346, ScanlineOptimizer.h:40)
void ref_wta(const float* vol, int H, int W, int D, float* disp) {
    rns_adc::CrossArmAggregation ca;
    ca.Initialize(H, W, nullptr, nullptr, 0, D);
    ca.WTA(const_cast<float*>(vol), disp);
    free_arr(ca.leftLength); free_arr(ca.rightLength); free_arr(ca.topLength); free_arr(ca.buttonLenght);
}

// ---------------------------------------------------------------- CrossArmAggregation (CrossArm.cpp)
// Runs Initialize + Left/Right/Top/Button arm passes in main.cpp:68-72 order on an 8UC1 (ch=1) or 8UC3 (ch=3)
// image, optionally AggregationVertical (mode 1) or Aggregation (mode 2) on vol_in.  Returns the final _tao.
int ref_crossarm(const uint8_t* img, int ch, int H, int W, int D, int tao, int32_t* armL, int32_t* armR,
                 int32_t* armT, int32_t* armB, int mode, const float* vol_in, float* vol_out) {
    rns_adc::CrossArmAggregation ca;
    ca.Initialize(H, W, nullptr, nullptr, tao, D);
    cv::Mat im = wrap_u8(img, H, W, ch);
    ca.ComputeLeftArmLength(im);
    ca.ComputeRightArmLength(im);
    ca.ComputeTopArmLength(im);
    ca.ComputeButtonArmLength(im);
    size_t n = (size_t)H * W;
    if (armL) std::memcpy(armL, ca.leftLength, n * 4);
    if (armR) std::memcpy(armR, ca.rightLength, n * 4);
    if (armT) std::memcpy(armT, ca.topLength, n * 4);
    if (armB) std::memcpy(armB, ca.buttonLenght, n * 4);
    if (mode == 1) ca.AggregationVertical(const_cast<float*>(vol_in), vol_out);
    if (mode == 2) ca.Aggregation(const_cast<float*>(vol_in), vol_out);
    int t = ca._tao;
    free_arr(ca.leftLength); free_arr(ca.rightLength); free_arr(ca.topLength); free_arr(ca.buttonLenght);
    return t;
}

// ---------------------------------------------------------------- ScanlineOptimizer (ScanlineOptimizer.h)
void ref_scanline(const float* cost, const float* gray, int H, int W, int D, int p1, int p2, float* left,
                  float* right, float* up, float* down, float* sum, float* disp) {
    rns_adc::ScanlineOptimizer so;
    so.Initialize(H, W, D, const_cast<float*>(cost), p1, p2);
    so.ScanLine(const_cast<float*>(cost), const_cast<float*>(gray));
    size_t n = (size_t)H * W * D;
    if (left) std::memcpy(left, so.leftVolume, n * 4);
    if (right) std::memcpy(right, so.rightVolume, n * 4);
    if (up) std::memcpy(up, so.upVolume, n * 4);
    if (down) std::memcpy(down, so.downVolume, n * 4);
    if (sum) std::memcpy(sum, so._ProcessedVolume, n * 4);
    if (disp) so.WTA(disp);
    free_arr(so.leftVolume); free_arr(so.rightVolume); free_arr(so.upVolume); free_arr(so.downVolume);
    free_arr(so._ProcessedVolume);
}

// ---------------------------------------------------------------- PostProcessing.h
// occ / mis: int32 pairs (row, col) in emission order; capacity H*W pairs each.
void ref_lrcheck(float* dispL, const float* dispR, int W, int H, int gate, int32_t* occ, int32_t* n_occ,
                 int32_t* mis, int32_t* n_mis) {
    std::vector<std::pair<int, int>> o, m;
    rns_adc::LeftRightConsistency(W, H, gate, dispL, const_cast<float*>(dispR), o, m);
    write_pairs(o, occ, n_occ);
    write_pairs(m, mis, n_mis);
}
void ref_lrcheck_and(const float* dispL, const float* dispR, float* last, int W, int H, float gate, int32_t* occ,
                     int32_t* n_occ, int32_t* mis, int32_t* n_mis) {
    std::vector<std::pair<int, int>> o, m;
    rns_adc::LeftAndRightConsistency(const_cast<float*>(dispL), const_cast<float*>(dispR), last, W, H, gate, o, m);
    write_pairs(o, occ, n_occ);
    write_pairs(m, mis, n_mis);
}
// exactly the call in main.cpp:93 -- passes the float +inf constant to a `const int&` parameter
void ref_remove_speckles_inf(float* disp, int W, int H, int diff, unsigned min_area) {
    rns_adc::RemoveSpeckles(disp, W, H, diff, min_area, rns_adc::Invalid_Float);
}
void ref_remove_speckles(float* disp, int W, int H, int diff, unsigned min_area, int invalid_val) {
    rns_adc::RemoveSpeckles(disp, W, H, diff, min_area, invalid_val);
}
void ref_median(const float* in, float* out, int W, int H, int k) { rns_adc::MedianFilter(in, out, W, H, k); }
// FillTheHole (PostProcessing.h:156-248), exactly as written (it swaps width/height and tests 0xffff; never called by the
// reference's main).  occ / mis: (row, col) pairs in, mis is overwritten by the third phase like the reference's vector.
void ref_fill_the_hole(int row, int col, int dispRange, float* disp, const int32_t* occ, int32_t n_occ, int32_t* mis, int32_t* n_mis,
                       int32_t mis_capacity) {
    std::vector<std::pair<int, int>> o, m;
    for (int k = 0; k < n_occ; ++k) o.emplace_back(occ[2 * k], occ[2 * k + 1]);
    for (int k = 0; k < *n_mis; ++k) m.emplace_back(mis[2 * k], mis[2 * k + 1]);
    rns_adc::FillTheHole(row, col, dispRange, disp, o, m);
    *n_mis = (int32_t)m.size();
    for (size_t k = 0; k < m.size() && (int32_t)k < mis_capacity; ++k) { mis[2 * k] = m[k].first; mis[2 * k + 1] = m[k].second; }
}
// TransformToShow (PostProcessing.h:137): float disparity -> 8-bit display image
void ref_transform_to_show(const float* disp, int H, int W, uint8_t* out) {
    cv::Mat t(H, W, cv::CV_8UC1, out);
    rns_adc::TransformToShow(const_cast<float*>(disp), t, H, W);
}

// ---------------------------------------------------------------- CrossAggregator (cross_aggregator.cpp)
// arms: 4 bytes per pixel (left,right,top,bottom); sup0/sup1: uint16 per pixel.
int ref_cross_aggregate(const uint8_t* bgrL, const uint8_t* bgrR, const float* cost_init, int W, int H, int D,
                        int L1, int L2, int t1, int t2, int iters, float* cost_out, uint8_t* arms, uint16_t* sup0,
                        uint16_t* sup1) {
    rns_cblsm::CrossAggregator ag;
    if (!ag.Initialize(W, H, 0, D)) return 0;
    ag.SetData(bgrL, bgrR, cost_init);
    ag.SetParams(L1, L2, t1, t2);
    ag.Aggregate(iters);
    size_t n = (size_t)W * H;
    if (cost_out) std::memcpy(cost_out, ag.get_cost_ptr(), n * D * 4);
    if (arms) std::memcpy(arms, ag.get_arms_ptr(), n * 4);
    if (sup0) std::memcpy(sup0, ag.vec_sup_count_[0].data(), n * 2);
    if (sup1) std::memcpy(sup1, ag.vec_sup_count_[1].data(), n * 2);
    return 1;
}

// ---------------------------------------------------------------- CBLSM.h live functions
void ref_cblsm_ad(const uint8_t* left, const uint8_t* right, int H, int W, int D, float* adL, float* adR) {
    if (adL) rns_cblsm::ComputeAD(W, H, D, const_cast<uint8_t*>(left), const_cast<uint8_t*>(right), adL);
    if (adR) rns_cblsm::ComputeADRight(W, H, D, const_cast<uint8_t*>(left), const_cast<uint8_t*>(right), adR);
}
void ref_cblsm_arms(const uint8_t* img, int ch, int H, int W, int tao, int maxLength, int secLength, int32_t* armL,
                    int32_t* armR, int32_t* armUp, int32_t* armDown) {
    cv::Mat im = wrap_u8(img, H, W, ch);
    rns_cblsm::ArmLengthL(im, (uchar)tao, armL, maxLength, secLength);
    rns_cblsm::ArmLengthR(im, (uchar)tao, armR, maxLength, secLength);
    rns_cblsm::ArmLengthUp(im, (uchar)tao, armUp, maxLength, secLength);
    rns_cblsm::ArmLengthDown(im, (uchar)tao, armDown, maxLength, secLength);
}
void ref_cblsm_aggregate_v5(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                            const int32_t* armUp, const int32_t* armDown, int D, int H, int W) {
    rns_cblsm::costAggregationV5(const_cast<float*>(vol_in), vol_out, const_cast<int*>(armL), const_cast<int*>(armR),
                                 const_cast<int*>(armUp), const_cast<int*>(armDown), D, H, W, 1);
}
// dead-code variants of CBLSM.h kept callable for parity of the "remaining variants" row
void ref_cblsm_aggregate_vertical(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                                  const int32_t* armUp, const int32_t* armDown, int D, int H, int W) {
    rns_cblsm::costAggregationVertical(const_cast<float*>(vol_in), vol_out, const_cast<int*>(armL), const_cast<int*>(armR),
                                       const_cast<int*>(armUp), const_cast<int*>(armDown), D, H, W, 1);
}
// costAggregationV4: the four arm arrays are per (pixel, d) cell, int [H][W][D]
void ref_cblsm_aggregate_v4(const float* vol_in, float* vol_out, const int32_t* armL, const int32_t* armR,
                            const int32_t* armUp, const int32_t* armDown, int D, int H, int W) {
    rns_cblsm::costAggregationV4(const_cast<float*>(vol_in), vol_out, const_cast<int*>(armL), const_cast<int*>(armR),
                                 const_cast<int*>(armUp), const_cast<int*>(armDown), D, H, W, 1);
}
void ref_cblsm_wta(const float* vol, float* disp, int D, int H, int W) {
    rns_cblsm::ComputeDispOringin(const_cast<float*>(vol), disp, D, H, W);
}

// ---------------------------------------------------------------- SAD (Sad.h)
// Inputs are the replicate-padded gray images [Hp][Wp] (pad = winsize+1 per side, SADmain.cpp:47-48);
// disp is CV_32S [Hp-2*pad][Wp-2*pad], zero-initialised by the caller like Mat::zeros (SADmain.cpp:59).
void ref_sad_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp) {
    CoutSilencer q;
    int pad = winsize + 1;
    cv::Mat l = wrap_u8(Lp, Hp, Wp), r = wrap_u8(Rp, Hp, Wp);
    cv::Mat d(Hp - 2 * pad, Wp - 2 * pad, cv::CV_32SC1, disp);
    rns_sad::GetPointDepthLeft(d, l, r, D, winsize);
}
void ref_sad_right(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, int32_t* disp) {
    CoutSilencer q;
    int pad = winsize + 1;
    cv::Mat l = wrap_u8(Lp, Hp, Wp), r = wrap_u8(Rp, Hp, Wp);
    cv::Mat d(Hp - 2 * pad, Wp - 2 * pad, cv::CV_32SC1, disp);
    rns_sad::GetPointDepthRight(d, l, r, D, winsize);
}
// per-disparity SAD vector of the left view, [H][W][D] f32: the d-loop of Sad.h:122-132 around ref sadvalue()
void ref_sad_costs_left(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int D, int winsize, float* costs) {
    int w = winsize + 1, H = Hp - 2 * w, W = Wp - 2 * w;
    cv::Mat l = wrap_u8(Lp, Hp, Wp), r = wrap_u8(Rp, Hp, Wp);
    for (int i = w; i < Hp - w; ++i)
        for (int j = w; j < Wp - w; ++j) {
            cv::Mat lw = l(cv::Range(i - w, i + w + 1), cv::Range(j - w, j + w + 1));
            float* s = costs + ((size_t)(i - w) * W + (j - w)) * D;
            for (int d = 0; d < D; ++d) {
                if (j - w - d < 0) { s[d] = s[d - 1]; continue; }
                cv::Mat rw = r(cv::Range(i - w, i + w + 1), cv::Range(j - w - d, j - d + w + 1));
                s[d] = rns_sad::sadvalue(lw, rw);
            }
        }
    (void)H;
}
float ref_sad_optimal_disparity(const float* sad, int D) {
    std::vector<float> v(sad, sad + D);
    return rns_sad::OptimalDisparity(v, D);
}

// CrossCheckDiaparity on CV_32S maps (Sad.h:184-221).  occ / mis: capacity H*W pairs each; returns the counts.
void ref_sad_crosscheck(const int32_t* dispL, const int32_t* dispR, int H, int W, int32_t* out, int32_t* occ, int32_t* n_occ,
                        int32_t* mis, int32_t* n_mis) {
    cv::Mat l(H, W, cv::CV_32SC1, const_cast<int32_t*>(dispL)), r(H, W, cv::CV_32SC1, const_cast<int32_t*>(dispR));
    cv::Mat o(H, W, cv::CV_32SC1, out);
    std::vector<std::pair<int, int>> vo, vm;
    rns_sad::CrossCheckDiaparity(l, r, o, 0, 0, vo, vm);
    *n_occ = (int32_t)vo.size();
    *n_mis = (int32_t)vm.size();
    for (size_t k = 0; k < vo.size(); ++k) { occ[2 * k] = vo[k].first; occ[2 * k + 1] = vo[k].second; }
    for (size_t k = 0; k < vm.size(); ++k) { mis[2 * k] = vm[k].first; mis[2 * k + 1] = vm[k].second; }
}

// ---------------------------------------------------------------- NCC (NCC.h)
// ncc() (NCC.h:117-272), the shifted-image variant; type 0 = "left", 1 = "right"
void ref_ncc_shift(const uint8_t* L, const uint8_t* R, int H, int W, int type_right, int add_constant, uint8_t* depth) {
    cv::Mat l = wrap_u8(L, H, W).clone(), r = wrap_u8(R, H, W).clone();
    cv::Mat d = rns_ncc::ncc(l, r, type_right ? "right" : "left", add_constant != 0);
    for (int i = 0; i < H; ++i) for (int j = 0; j < W; ++j) depth[(size_t)i * W + j] = d.at<uchar>(i, j);
}

// disp CV_32S [H][W], zero-initialised by the caller (NCC_main.cpp:20)
void ref_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int D, int32_t* disp) {
    CoutSilencer q;
    cv::Mat l = wrap_u8(L, H, W), r = wrap_u8(R, H, W);
    cv::Mat d(H, W, cv::CV_32SC1, disp);
    rns_ncc::NCC_algorithem(l, r, W, H, d, winSize, D);
}
// NCC cost of one (pixel, d) -- for cost-level tolerance checks
double ref_ncc_cost(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int i, int j, int d) {
    cv::Mat l = wrap_u8(L, H, W), r = wrap_u8(R, H, W);
    cv::Mat wl = l(cv::Range(i - winSize, i + winSize + 1), cv::Range(j - winSize, j + winSize + 1));
    cv::Mat wr = r(cv::Range(i - winSize, i + winSize + 1), cv::Range(j - winSize - d, j + winSize - d + 1));
    double c = 0;
    rns_ncc::ComputeCost(wl, wr, c);
    return c;
}

// NCC cost rows of output rows [i0, i1): the loop of NCC_algorithem (NCC.h:73-90) with the reference's ComputeCost doing
// the arithmetic; out: double [i1-i0][W][D], 255 (`invalid`, :7) where the window leaves the image, 0 at border columns.
void ref_ncc_costs(const uint8_t* L, const uint8_t* R, int H, int W, int winSize, int D, int i0, int i1, double* out) {
    cv::Mat l = wrap_u8(L, H, W), r = wrap_u8(R, H, W);
    for (int i = i0; i < i1; ++i)
        for (int j = 0; j < W; ++j) {
            double* row = out + ((size_t)(i - i0) * W + j) * D;
            for (int d = 0; d < D; ++d) row[d] = 0.0;
            if (i < winSize || i >= H - winSize || j < winSize || j >= W - winSize) continue;
            cv::Mat wl = l(cv::Range(i - winSize, i + winSize + 1), cv::Range(j - winSize, j + winSize + 1));
            for (int d = 0; d < D; ++d) {
                if (j - winSize - d >= 0) {
                    cv::Mat wr = r(cv::Range(i - winSize, i + winSize + 1), cv::Range(j - winSize - d, j + winSize - d + 1));
                    rns_ncc::ComputeCost(wl, wr, row[d]);
                } else {
                    row[d] = 255.0;
                }
            }
        }
}

// ---------------------------------------------------------------- ASW (ASW.h)
// costVolume rows of output rows [i0, i1) of the left view: the loop of AdaptiveSupportWeight (ASW.h:336-374) with the
// reference's bilateralfiterWight doing the arithmetic (the function itself keeps costVolume local and only returns the
// winner).  out: float [i1-i0][Wo][D].
void ref_asw_costs(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int winSize, int D, double spaceSigma, double colorSigma,
                   int T, int i0, int i1, float* out) {
    const int wins = winSize + 1, Wo = Wp - 2 * wins;
    cv::Mat l = wrap_u8(Lp, Hp, Wp), r = wrap_u8(Rp, Hp, Wp);
    cv::Mat space;
    std::vector<double> color;
    rns_asw::getGausssianMask(space, cv::Size(winSize * 2 + 3, winSize * 2 + 3), spaceSigma);
    rns_asw::getColorMask(color, colorSigma);
    for (int io = i0; io < i1; ++io)
        for (int jo = 0; jo < Wo; ++jo) {
            const int i = io + wins, j = jo + wins;
            float* row = out + ((size_t)(io - i0) * Wo + jo) * D;
            cv::Mat lw = l(cv::Range(i - wins, i + wins + 1), cv::Range(j - wins, j + wins + 1));
            for (int d = 0; d < D; ++d) {
                if (j - wins - d >= 0) {
                    cv::Mat rw = r(cv::Range(i - wins, i + wins + 1), cv::Range(j - wins - d, j + wins - d + 1));
                    cv::Mat error;
                    cv::absdiff(lw, rw, error);
                    for (int a = 0; a < error.rows; ++a)
                        for (int b = 0; b < error.cols; ++b)
                            if (error.at<uchar>(a, b) > T) error.at<uchar>(a, b) = T;
                    row[d] = rns_asw::bilateralfiterWight(lw, rw, error, winSize, space, color);
                } else {
                    row[d] = row[d - 1];
                }
            }
        }
}

// Lp/Rp: replicate-padded gray [Hp][Wp], pad = winSize+1 (ASWeight.cpp:54-55). disp: CV_32FC1 [H][W] zeroed by caller.
void ref_asw(const uint8_t* Lp, const uint8_t* Rp, int Hp, int Wp, int winSize, int D, double spaceSigma,
             double colorSigma, int T, int right_view, float* disp) {
    CoutSilencer q;
    int pad = winSize + 1;
    cv::Mat l = wrap_u8(Lp, Hp, Wp), r = wrap_u8(Rp, Hp, Wp);
    cv::Mat lab_l = l, lab_r = r;  // the Lab windows are sliced but never read (ASW.h:342,346)
    cv::Mat space;
    std::vector<double> color;
    rns_asw::getGausssianMask(space, cv::Size(winSize * 2 + 3, winSize * 2 + 3), spaceSigma);
    rns_asw::getColorMask(color, colorSigma);
    cv::Mat d(Hp - 2 * pad, Wp - 2 * pad, cv::CV_32FC1, disp);
    if (right_view) rns_asw::AdaptiveSupportWeightRight(d, l, r, winSize, D, lab_l, lab_r, space, color, T);
    else rns_asw::AdaptiveSupportWeight(d, l, r, winSize, D, lab_l, lab_r, space, color, T);
}
void ref_asw_masks(int winSize, double spaceSigma, double colorSigma, double* space_out, double* color_out) {
    cv::Mat space;
    std::vector<double> color;
    int k = winSize * 2 + 3;
    rns_asw::getGausssianMask(space, cv::Size(k, k), spaceSigma);
    rns_asw::getColorMask(color, colorSigma);
    for (int i = 0; i < k; ++i) for (int j = 0; j < k; ++j) space_out[i * k + j] = space.at<double>(i, j);
    for (int i = 0; i < 256; ++i) color_out[i] = color[i];
}
void ref_asw_crosscheck(const float* dispL, const float* dispR, int H, int W, uint8_t* out) {
    cv::Mat l(H, W, cv::CV_32FC1, const_cast<float*>(dispL)), r(H, W, cv::CV_32FC1, const_cast<float*>(dispR));
    cv::Mat o(H, W, cv::CV_8UC1, out);
    rns_asw::CrossCheckDiaparity(l, r, o, 0, 3);
}
// FillImageNew (ASW/ASW.h:434), the hole filling of the ASW driver's display chain (ASWeight.cpp:76).  The function reads
// disp.at<uchar>(row, cols) -- one element past the row, for the last row one byte past the image (:494) -- so the caller
// hands in rows*cols + 1 bytes and decides what that byte holds.  n_values: how many fill values the reference collected
// (when a search ends exactly on that element with a zero in it, :485-500 pushes nothing and every later pixel takes its
// successor's value; the last pixels then read past the end of the value vector, which the caller must not compare).
void ref_fill_image_new(uint8_t* disp_plus1, int H, int W) {
    cv::Mat d(H, W, cv::CV_8UC1, disp_plus1);
    rns_asw::FillImageNew(d);
}

}  // extern "C"
