// smt_b200.hpp -- header-only C++ facade over the C ABI (smt_b200.h) that keeps the reference's per-method entry
// points: same class and function names, same argument order and meaning, caller-owned host arrays.  A call site of
// Asherchi/Stereo_Match_Traditional switches by including this header instead of the reference's and linking
// libsmt_b200.so.  cv::Mat parameters become smt::ImageView (OpenCV is not a dependency);
// define SMT_WITH_OPENCV before including to get cv::Mat overloads.
//
// Differences by design: the classes free what they allocate (the reference never does, e.g. AD-Census.h:382), and a
// failing call throws smt::Error instead of silently continuing.  There is no CPU fallback.
#ifndef SMT_B200_HPP_
#define SMT_B200_HPP_

#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "smt_b200.h"

#ifdef SMT_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace smt {

constexpr float Invalid_Float = std::numeric_limits<float>::infinity();  // PostProcessing.h:7

struct Error : std::runtime_error {
    smt_status_t status;
    Error(smt_status_t st, const std::string& what) : std::runtime_error(what), status(st) {}
};

/// Stand-in for the cv::Mat arguments of the reference: 8-bit image, 1 or 3 interleaved channels.
struct ImageView {
    const uint8_t* data = nullptr;
    int rows = 0, cols = 0, channels = 1;
    size_t step = 0;  // bytes per row; 0 = dense
    ImageView() = default;
    ImageView(const uint8_t* d, int r, int c, int ch = 1, size_t s = 0) : data(d), rows(r), cols(c), channels(ch), step(s) {}
#ifdef SMT_WITH_OPENCV
    ImageView(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), channels(m.channels()), step(m.step) {}
#endif
    /// dense copy when the view has row padding
    std::vector<uint8_t> dense() const {
        if (!data || rows <= 0 || cols <= 0 || channels <= 0) throw Error(SMT_ERR_INVALID_ARG, "ImageView: null or empty image");
        const size_t row_bytes = (size_t)cols * channels;
        std::vector<uint8_t> out((size_t)rows * row_bytes);
        const size_t st = step ? step : row_bytes;
        for (int r = 0; r < rows; ++r) std::copy(data + r * st, data + r * st + row_bytes, out.begin() + r * row_bytes);
        return out;
    }
};

/// One context per (device, stream).  The facade classes share the default one unless given another.
class Context {
public:
    explicit Context(int device = 0, void* stream = nullptr) {
        const smt_status_t st = smt_create(device, stream, &ctx_);
        if (st != SMT_OK) throw Error(st, std::string("smt_create: ") + smt_status_string(st) + " (no CPU fallback)");
    }
    ~Context() { smt_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    smt_ctx* get() const { return ctx_; }
    void check(smt_status_t st) const {
        if (st != SMT_OK) throw Error(st, smt_last_error_string(ctx_));
    }
    static Context& Default() {
        static Context c(0);
        return c;
    }

private:
    smt_ctx* ctx_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ AD-CensusV1/AD-Census.h:9-43
class AD_Census {
public:
    explicit AD_Census(Context& c = Context::Default()) : c_(c) {}
    void Initialize(float* leftImage, float* rightImage, int dispRange, int row, int col, const ImageView& /*LImage*/,
                    const ImageView& /*RImage*/, float sigmaC, float sigmaS) {
        left_ = leftImage; right_ = rightImage; D_ = dispRange; row_ = row; col_ = col; sigmaC_ = sigmaC; sigmaS_ = sigmaS;
        const size_t n = (size_t)row * col * dispRange;
        costVolume.assign(0, 0.f); costVolumeRight.assign(0, 0.f);
        n_ = n;
    }
    void ComputeADcensus() { run(0, costVolume, nullptr, nullptr); }
    void ComputeADcensusRight() { run(1, costVolumeRight, nullptr, nullptr); }
    /// AD-Census.h:75,:103 -- the AD volume alone (kept for API parity; ComputeADcensus does not need it first)
    void ComputeAD() { ad_.resize(n_); tmp_.resize(n_); run(0, tmp_, ad_.data(), nullptr); }
    void ComputeADRight() { adR_.resize(n_); tmp_.resize(n_); run(1, tmp_, adR_.data(), nullptr); }
    /// AD-Census.h:142,:207 -- the Hamming volume alone
    void ComputeCensus9x7(float* leftImage, float* rightImage) { census(leftImage, rightImage, 0, cen_); }
    void ComputeCensus9x7Right(float* leftImage, float* rightImage) { census(leftImage, rightImage, 1, cenR_); }
    float* GetPtrLeft() { return costVolume.empty() ? nullptr : costVolume.data(); }      // :50-60
    float* GetPtrRight() { return costVolumeRight.empty() ? nullptr : costVolumeRight.data(); }
    const std::vector<float>& ADcost(bool right = false) const { return right ? adR_ : ad_; }
    const std::vector<float>& CensusCost(bool right = false) const { return right ? cenR_ : cen_; }
    void WTA(float* leftdisp, float* rightDisp) {  // :346
        c_.check(smt_wta_first_min(c_.get(), costVolume.data(), row_, col_, D_, leftdisp, SMT_HOST));
        c_.check(smt_wta_first_min(c_.get(), costVolumeRight.data(), row_, col_, D_, rightDisp, SMT_HOST));
    }

private:
    void run(int view, std::vector<float>& cost, float* ad, float* cen) {
        cost.resize(n_);
        c_.check(smt_adcensus_cost(c_.get(), left_, right_, row_, col_, D_, sigmaC_, sigmaS_, view, cost.data(), ad, cen, SMT_HOST));
    }
    void census(float* l, float* r, int view, std::vector<float>& out) {
        out.resize(n_); tmp_.resize(n_);
        c_.check(smt_adcensus_cost(c_.get(), l, r, row_, col_, D_, sigmaC_, sigmaS_, view, tmp_.data(), nullptr, out.data(), SMT_HOST));
    }
    Context& c_;
    float *left_ = nullptr, *right_ = nullptr;
    int D_ = 0, row_ = 0, col_ = 0;
    size_t n_ = 0;
    float sigmaC_ = 10.f, sigmaS_ = 30.f;
    std::vector<float> costVolume, costVolumeRight, ad_, adR_, cen_, cenR_, tmp_;
};

// ------------------------------------------------------------------------------------------------ AD-CensusV1/CrossArm.h:9-36
class CrossArmAggregation {
public:
    explicit CrossArmAggregation(Context& c = Context::Default(), smt_sum_mode_t mode = SMT_SUM_REFERENCE_ORDER) : c_(c), mode_(mode) {}
    void Initialize(int row, int col, float* /*leftImage*/, float* /*rightImage*/, int tao, int dispRange) {
        row_ = row; col_ = col; tao_ = tao; D_ = dispRange;
        const size_t n = (size_t)row * col;
        leftLength.assign(n, 0); rightLength.assign(n, 0); topLength.assign(n, 0); buttonLenght.assign(n, 0);
        have_ = false;
    }
    // The four passes share one sticky threshold and must run in main.cpp:69-72 order; they are computed together
    // on the first call with an image and handed out one by one.  A call with a DIFFERENT image (other buffer or
    // geometry) recomputes all four from that image, continuing from the threshold the previous image left behind,
    // like the reference's member _tao does (CrossArm.cpp:168-173); Initialize resets it.
    void ComputeLeftArmLength(const ImageView& Image) { arms(Image); }
    void ComputeRightArmLength(const ImageView& Image) { arms(Image); }
    void ComputeTopArmLength(const ImageView& Image) { arms(Image); }
    void ComputeButtonArmLength(const ImageView& Image) { arms(Image); }
    void Aggregation(float* dispVolume, float* aggregatedCostVolume) { agg(dispVolume, aggregatedCostVolume, 1); }
    void AggregationVertical(float* dispVolume, float* aggregatedCostVolume) { agg(dispVolume, aggregatedCostVolume, 0); }
    void WTA(float* AggredCostVolume, float* disp) {
        c_.check(smt_wta_first_min(c_.get(), AggredCostVolume, row_, col_, D_, disp, SMT_HOST));
    }
    int tao() const { return tao_; }
    std::vector<int32_t> leftLength, rightLength, topLength, buttonLenght;

private:
    void arms(const ImageView& im) {
        if (have_ && im.data == key_.data && im.rows == key_.rows && im.cols == key_.cols && im.channels == key_.channels &&
            im.step == key_.step)
            return;
        key_ = im;
        const std::vector<uint8_t> d = im.dense();
        int32_t tf = tao_;
        c_.check(smt_arms_gray(c_.get(), d.data(), im.channels, row_, col_, tao_, 34, 17, 0, leftLength.data(), rightLength.data(),
                               topLength.data(), buttonLenght.data(), &tf, SMT_HOST));
        tao_ = tf;
        have_ = true;
    }
    void agg(float* in, float* out, int variant) {
        c_.check(smt_aggregate_rect(c_.get(), in, leftLength.data(), rightLength.data(), topLength.data(), buttonLenght.data(), row_,
                                    col_, D_, variant, mode_, out, SMT_HOST));
    }
    Context& c_;
    smt_sum_mode_t mode_;
    int row_ = 0, col_ = 0, tao_ = 0, D_ = 0;
    bool have_ = false;
    ImageView key_;
};

// ------------------------------------------------------------------------------------------------ AD-CensusV1/ScanlineOptimizer.h:8-34
class ScanlineOptimizer {
public:
    explicit ScanlineOptimizer(Context& c = Context::Default()) : c_(c) {}
    void Initialize(int row, int col, int dispRange, float* costVolume, int p1, int p2) {
        row_ = row; col_ = col; D_ = dispRange; cost_ = costVolume; p1_ = p1; p2_ = p2;
        const size_t n = (size_t)row * col * dispRange;
        leftVolume.assign(n, 0.f); rightVolume.assign(n, 0.f); upVolume.assign(n, 0.f); downVolume.assign(n, 0.f);
        _ProcessedVolume.assign(n, 0.f);
        disp_.assign((size_t)row * col, 0.f);
    }
    void ScanLine(float* costVolume, float* Image) {
        c_.check(smt_scanline(c_.get(), costVolume, Image, row_, col_, D_, p1_, p2_, leftVolume.data(), rightVolume.data(),
                              upVolume.data(), downVolume.data(), _ProcessedVolume.data(), disp_.data(), SMT_HOST));
    }
    void WTA(float* disp) { std::copy(disp_.begin(), disp_.end(), disp); }
    std::vector<float> leftVolume, rightVolume, upVolume, downVolume, _ProcessedVolume;

private:
    Context& c_;
    int row_ = 0, col_ = 0, D_ = 0, p1_ = 0, p2_ = 0;
    float* cost_ = nullptr;
    std::vector<float> disp_;
};

// ------------------------------------------------------------------------------------------------ CBLSM/cross_aggregator.h:17-113
struct CrossArm {
    uint8_t left, right, top, bottom;
};

class CrossAggregator {
public:
    explicit CrossAggregator(Context& c = Context::Default(), smt_sum_mode_t mode = SMT_SUM_REFERENCE_ORDER) : c_(c), mode_(mode) {}
    bool Initialize(const int32_t& width, const int32_t& height, const int32_t& min_disparity, const int32_t& max_disparity) {
        width_ = width; height_ = height; D_ = max_disparity - min_disparity;
        is_initialized_ = (int64_t)width * height > 0 && D_ > 0;  // cross_aggregator.cpp:26-31
        if (is_initialized_) {
            arms_.assign((size_t)width * height, CrossArm{0, 0, 0, 0});
            cost_aggr_.assign((size_t)width * height * D_, 0.f);
            sup_[0].assign((size_t)width * height, 0); sup_[1].assign((size_t)width * height, 0);
        }
        return is_initialized_;
    }
    void SetData(const uint8_t* img_left, const uint8_t* img_right, const float* cost_init) {
        img_left_ = img_left; img_right_ = img_right; cost_init_ = cost_init;
    }
    void SetParams(const int32_t& cross_L1, const int32_t& cross_L2, const int32_t& cross_t1, const int32_t& cross_t2) {
        L1_ = cross_L1; L2_ = cross_L2; t1_ = cross_t1; t2_ = cross_t2;
    }
    void Aggregate(const int32_t& num_iters) {
        if (!is_initialized_) return;  // :91-93
        c_.check(smt_aggregate_cross(c_.get(), img_left_, cost_init_, height_, width_, D_, L1_, L2_, t1_, t2_, num_iters, mode_,
                                     cost_aggr_.data(), reinterpret_cast<uint8_t*>(arms_.data()), sup_[0].data(), sup_[1].data(), SMT_HOST));
    }
    CrossArm* get_arms_ptr() { return arms_.data(); }
    float* get_cost_ptr() { return cost_aggr_.empty() ? nullptr : cost_aggr_.data(); }

private:
    Context& c_;
    smt_sum_mode_t mode_;
    int32_t width_ = 0, height_ = 0, D_ = 0, L1_ = 34, L2_ = 17, t1_ = 20, t2_ = 6;
    const uint8_t *img_left_ = nullptr, *img_right_ = nullptr;
    const float* cost_init_ = nullptr;
    bool is_initialized_ = false;
    std::vector<CrossArm> arms_;
    std::vector<float> cost_aggr_;
    std::vector<uint16_t> sup_[2];
};

// ------------------------------------------------------------------------------------------------ AD-CensusV1/PostProcessing.h
namespace detail {
inline void pairs(const std::vector<int32_t>& flat, int n, std::vector<std::pair<int, int>>& out, bool clear) {
    if (clear) out.clear();
    for (int k = 0; k < n; ++k) out.emplace_back(flat[2 * k], flat[2 * k + 1]);
}
}  // namespace detail

inline void LeftRightConsistency(int col, int row, int gate, float* leftDisp, float* rightDisp, std::vector<std::pair<int, int>>& occlusion,
                                 std::vector<std::pair<int, int>>& mismatch, Context& c = Context::Default()) {
    std::vector<int32_t> o((size_t)row * col * 2), m((size_t)row * col * 2);
    int32_t no = 0, nm = 0;
    c.check(smt_lrcheck(c.get(), leftDisp, rightDisp, col, row, gate, o.data(), &no, m.data(), &nm, SMT_HOST));
    detail::pairs(o, no, occlusion, true);  // :82-83 clears both lists
    detail::pairs(m, nm, mismatch, true);
}
inline void LeftAndRightConsistency(float* leftDisp, float* rightDisp, float* lastDisp, int col, int row, float gate,
                                    std::vector<std::pair<int, int>>& occlusion, std::vector<std::pair<int, int>>& mismatch,
                                    Context& c = Context::Default()) {
    std::vector<int32_t> o((size_t)row * col * 2), m((size_t)row * col * 2);
    int32_t no = 0, nm = 0;
    c.check(smt_lrcheck_and(c.get(), leftDisp, rightDisp, lastDisp, col, row, gate, o.data(), &no, m.data(), &nm, SMT_HOST));
    detail::pairs(o, no, occlusion, false);  // appends (:44-66)
    detail::pairs(m, nm, mismatch, false);
}
/// invalid_val is an int in the reference; passing Invalid_Float there becomes 2147483647 under g++ (main.cpp:93).
inline void RemoveSpeckles(float* disparity_map, const int& width, const int& height, const int& diff_insame,
                           const unsigned int& min_speckle_aera, const int& invalid_val, Context& c = Context::Default()) {
    c.check(smt_remove_speckles(c.get(), disparity_map, width, height, diff_insame, min_speckle_aera, invalid_val, SMT_HOST));
}
inline void RemoveSpeckles(float* disparity_map, const int& width, const int& height, const int& diff_insame,
                           const unsigned int& min_speckle_aera, float invalid_val, Context& c = Context::Default()) {
    const int iv = std::isinf(invalid_val) ? 2147483647 : (int)invalid_val;
    c.check(smt_remove_speckles(c.get(), disparity_map, width, height, diff_insame, min_speckle_aera, iv, SMT_HOST));
}
/// As written in the reference (axes swapped, invalid = 0xffff) unless variant = SMT_FILL_FIXED; `mismatch` is replaced by
/// the third phase (:188).
inline void FillTheHole(const int row, const int col, const int dispRange, float* dispLeft, std::vector<std::pair<int, int>>& occlusion,
                        std::vector<std::pair<int, int>>& mismatch, Context& c = Context::Default(), int variant = SMT_FILL_AS_WRITTEN) {
    std::vector<int32_t> o(occlusion.size() * 2), m((size_t)row * col * 2);
    for (size_t k = 0; k < occlusion.size(); ++k) { o[2 * k] = occlusion[k].first; o[2 * k + 1] = occlusion[k].second; }
    for (size_t k = 0; k < mismatch.size() && k < (size_t)row * col; ++k) { m[2 * k] = mismatch[k].first; m[2 * k + 1] = mismatch[k].second; }
    int32_t nm = (int32_t)mismatch.size();
    c.check(smt_fill_the_hole(c.get(), dispLeft, row, col, dispRange, o.data(), (int32_t)occlusion.size(), m.data(), &nm, variant, SMT_HOST));
    detail::pairs(m, nm, mismatch, true);
}
inline void MedianFilter(const float* in, float* out, const int& width, const int& height, const int wnd_size, Context& c = Context::Default()) {
    c.check(smt_median(c.get(), in, out, width, height, wnd_size, SMT_HOST));
}

// ------------------------------------------------------------------------------------------------ CBLSM/CBLSM.h (live part)
inline void ComputeAD(int col, int row, int dispRange, uint8_t* leftImage, uint8_t* rightImage, float* ADcostVolum, Context& c = Context::Default()) {
    c.check(smt_ad_cost(c.get(), leftImage, rightImage, row, col, dispRange, 0, ADcostVolum, SMT_HOST));
}
inline void ComputeADRight(int col, int row, int dispRange, uint8_t* leftImage, uint8_t* rightImage, float* ADcostVolum, Context& c = Context::Default()) {
    c.check(smt_ad_cost(c.get(), leftImage, rightImage, row, col, dispRange, 1, ADcostVolum, SMT_HOST));
}
namespace detail {
inline void cblsm_arm(const ImageView& im, uint8_t tao, int* arm, int maxLength, int secLength, int which, Context& c) {
    const std::vector<uint8_t> d = im.dense();
    int32_t* outs[4] = {nullptr, nullptr, nullptr, nullptr};
    outs[which] = arm;
    c.check(smt_arms_gray(c.get(), d.data(), im.channels, im.rows, im.cols, tao, maxLength, secLength, 1, outs[0], outs[1], outs[2], outs[3],
                          nullptr, SMT_HOST));
}
}  // namespace detail
inline void ArmLengthL(const ImageView& leftImage, uint8_t tao, int* LArm, int maxLength, int secLength, Context& c = Context::Default()) {
    detail::cblsm_arm(leftImage, tao, LArm, maxLength, secLength, 0, c);
}
inline void ArmLengthR(const ImageView& leftImage, uint8_t tao, int* RArm, int maxLength, int secLength, Context& c = Context::Default()) {
    detail::cblsm_arm(leftImage, tao, RArm, maxLength, secLength, 1, c);
}
inline void ArmLengthUp(const ImageView& leftImage, uint8_t tao, int* UpArm, int maxLength, int secLength, Context& c = Context::Default()) {
    detail::cblsm_arm(leftImage, tao, UpArm, maxLength, secLength, 2, c);
}
inline void ArmLengthDown(const ImageView& leftImage, uint8_t tao, int* DownArm, int maxLength, int secLength, Context& c = Context::Default()) {
    detail::cblsm_arm(leftImage, tao, DownArm, maxLength, secLength, 3, c);
}
inline void costAggregationV5(float* dispvolume, float* CostVolume, int* ArmvolumeL, int* ArmvolumeR, int* ArmvolumeUp, int* ArmvolumeDown,
                              int dispRange, int _row_, int _col_, int /*winSize*/, Context& c = Context::Default(),
                              smt_sum_mode_t mode = SMT_SUM_REFERENCE_ORDER) {
    c.check(smt_aggregate_rect(c.get(), dispvolume, ArmvolumeL, ArmvolumeR, ArmvolumeUp, ArmvolumeDown, _row_, _col_, dispRange, 2, mode,
                               CostVolume, SMT_HOST));
}
/// CBLSM.h:1227 -- the V5 rectangle in column-major order
inline void costAggregationVertical(float* dispvolume, float* CostVolume, int* ArmvolumeL, int* ArmvolumeR, int* ArmvolumeUp,
                                    int* ArmvolumeDown, int dispRange, int _row_, int _col_, int /*winSize*/, Context& c = Context::Default()) {
    c.check(smt_aggregate_rect(c.get(), dispvolume, ArmvolumeL, ArmvolumeR, ArmvolumeUp, ArmvolumeDown, _row_, _col_, dispRange, 0,
                               SMT_SUM_REFERENCE_ORDER, CostVolume, SMT_HOST));
}
/// CBLSM.h:1128 -- the four arm arrays are per (pixel, d) cell, int [row][col][dispRange]
inline void costAggregationV4(float* dispvolume, float* CostVolume, int* ArmvolumeL, int* ArmvolumeR, int* ArmvolumeUp, int* ArmvolumeDown,
                              int dispRange, int _row_, int _col_, int /*winSize*/, Context& c = Context::Default()) {
    c.check(smt_aggregate_rect(c.get(), dispvolume, ArmvolumeL, ArmvolumeR, ArmvolumeUp, ArmvolumeDown, _row_, _col_, dispRange, 3,
                               SMT_SUM_REFERENCE_ORDER, CostVolume, SMT_HOST));
}
inline void ComputeDispOringin(float* costVolume, float* disp, int dispRange, int _row_, int _col_, Context& c = Context::Default()) {
    c.check(smt_wta_first_min(c.get(), costVolume, _row_, _col_, dispRange, disp, SMT_HOST));
}

// ------------------------------------------------------------------------------------------------ SAD/Sad.h, NCC/NCC.h, ASW/ASW.h
/// disparity: int32 [rows-2*(winsize+1)][cols-2*(winsize+1)] (CV_32S); images are the replicate-padded grey images.
inline void GetPointDepthLeft(int32_t* disparity, const ImageView& leftimg, const ImageView& rightimg, const int MaxDisparity, const int winsize,
                              Context& c = Context::Default()) {
    const std::vector<uint8_t> l = leftimg.dense(), r = rightimg.dense();
    c.check(smt_sad(c.get(), l.data(), r.data(), leftimg.rows, leftimg.cols, MaxDisparity, winsize, 0, disparity, nullptr, nullptr, SMT_HOST));
}
inline void GetPointDepthRight(int32_t* disparity, const ImageView& leftimg, const ImageView& rightimg, const int MaxDisparity, const int winsize,
                               Context& c = Context::Default()) {
    const std::vector<uint8_t> l = leftimg.dense(), r = rightimg.dense();
    c.check(smt_sad(c.get(), l.data(), r.data(), leftimg.rows, leftimg.cols, MaxDisparity, winsize, 1, disparity, nullptr, nullptr, SMT_HOST));
}
/// disp: int32 [height][width]
inline void NCC_algorithem(const ImageView& leftImage, const ImageView& rigthImage, const int width, const int height, int32_t* disp, int winSize,
                           int dispRange, Context& c = Context::Default()) {
    const std::vector<uint8_t> l = leftImage.dense(), r = rigthImage.dense();
    c.check(smt_ncc(c.get(), l.data(), r.data(), height, width, winSize, dispRange, disp, nullptr, SMT_HOST));
}
/// disp: float [rows-2*(winSize+1)][cols-2*(winSize+1)]; the Lab images and the masks of the reference signature are
/// not needed (the Lab windows are never read, ASW.h:343; the masks are functions of the two sigmas).
inline void AdaptiveSupportWeight(float* disp, const ImageView& leftGray, const ImageView& rightGray, const int winSize, int dispRange,
                                  double spaceSigma, double colorSigma, int T, Context& c = Context::Default()) {
    const std::vector<uint8_t> l = leftGray.dense(), r = rightGray.dense();
    c.check(smt_asw(c.get(), l.data(), r.data(), leftGray.rows, leftGray.cols, winSize, dispRange, spaceSigma, colorSigma, T, 0, disp, nullptr, SMT_HOST));
}
inline void AdaptiveSupportWeightRight(float* disp, const ImageView& leftGray, const ImageView& rightGray, const int winSize, int dispRange,
                                       double spaceSigma, double colorSigma, int T, Context& c = Context::Default()) {
    const std::vector<uint8_t> l = leftGray.dense(), r = rightGray.dense();
    c.check(smt_asw(c.get(), l.data(), r.data(), leftGray.rows, leftGray.cols, winSize, dispRange, spaceSigma, colorSigma, T, 1, disp, nullptr, SMT_HOST));
}
inline void CrossCheckDiaparity(const float* leftdisp, const float* rightdisp, uint8_t* lastdisp, int rows, int cols, Context& c = Context::Default()) {
    c.check(smt_asw_crosscheck(c.get(), leftdisp, rightdisp, rows, cols, lastdisp, SMT_HOST));
}
/// Sad.h:184 -- the CV_32S form; MaxDisparity and winsize are unused by the reference as well
inline void CrossCheckDiaparity(const int32_t* leftdisp, const int32_t* rightdisp, int32_t* lastdisp, int rows, int cols, const int /*MaxDisparity*/,
                                const int /*winsize*/, std::vector<std::pair<int, int>>& occlusion, std::vector<std::pair<int, int>>& mismatch,
                                Context& c = Context::Default()) {
    std::vector<int32_t> o((size_t)rows * cols * 2), m((size_t)rows * cols * 2);
    int32_t no = 0, nm = 0;
    c.check(smt_sad_crosscheck(c.get(), leftdisp, rightdisp, rows, cols, lastdisp, o.data(), &no, m.data(), &nm, SMT_HOST));
    detail::pairs(o, no, occlusion, true);  // the reference clears both lists first (:193-194)
    detail::pairs(m, nm, mismatch, true);
}
/// NCC.h:117 -- depth: uint8 [rows][cols]
inline void ncc(const ImageView& in1, const ImageView& in2, const std::string& type, uint8_t* depth, bool add_constant = false,
                Context& c = Context::Default()) {
    const std::vector<uint8_t> l = in1.dense(), r = in2.dense();
    if (type != "left" && type != "right") throw Error(SMT_ERR_INVALID_ARG, "ncc: type must be \"left\" or \"right\"");
    c.check(smt_ncc_shift(c.get(), l.data(), r.data(), in1.rows, in1.cols, type == "right", add_constant, depth, SMT_HOST));
}
/// ASW.h:434 -- in place on an 8-bit map (rows x cols, dense); the byte the reference reads below the last row counts as 0
inline void FillImageNew(uint8_t* disp, int rows, int cols, Context& c = Context::Default()) {
    c.check(smt_fill_image_new(c.get(), disp, rows, cols, 0, nullptr, SMT_HOST));
}
/// the OpenCV calls of the ASW driver's display chain (ASWeight.cpp:69-78) on dense 8-bit maps
inline void normalize(const uint8_t* src, uint8_t* dst, int rows, int cols, Context& c = Context::Default()) {  // (0, 255, NORM_MINMAX)
    c.check(smt_normalize_minmax_u8_from_u8(c.get(), src, rows, cols, dst, SMT_HOST));
}
inline void filterSpeckles(uint8_t* img, int rows, int cols, double newVal, int maxSpeckleSize, double maxDiff, Context& c = Context::Default()) {
    c.check(smt_filter_speckles_u8(c.get(), img, rows, cols, (int32_t)newVal, maxSpeckleSize, (int32_t)maxDiff, SMT_HOST));
}
inline void medianBlur(const uint8_t* src, uint8_t* dst, int rows, int cols, int ksize, Context& c = Context::Default()) {
    if (src == dst) {  // OpenCV allows in-place use (ASWeight.cpp:74)
        std::vector<uint8_t> tmp(src, src + (size_t)rows * cols);
        c.check(smt_median_blur_u8(c.get(), tmp.data(), dst, rows, cols, ksize, SMT_HOST));
        return;
    }
    c.check(smt_median_blur_u8(c.get(), src, dst, rows, cols, ksize, SMT_HOST));
}
/// cv::imwrite(path, 8-bit image) for .png files (main.cpp:112-114, ASWeight.cpp:75-92)
inline bool imwrite(const std::string& filename, const ImageView& img, Context& c = Context::Default()) {
    const std::vector<uint8_t> d = img.dense();
    c.check(smt_write_png_u8(c.get(), filename.c_str(), d.data(), img.rows, img.cols, img.channels, SMT_HOST));
    return true;
}

}  // namespace smt
#endif  // SMT_B200_HPP_
