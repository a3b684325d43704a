// Host-side C++ mirror of the reference interface for the stereo point-line frontend, over the C ABI of
// include/plf_b200.h.  Same class names, argument meaning and error behaviour as PLI-SLAM:
//   ORB_SLAM3::ORBextractor::operator()           include/ORBextractor.h:61-63, src/ORBextractor.cc:1068-1150
//   ORB_SLAM3::Lineextractor::operator()          include/LineExtractor.h:49-51, src/LineExtractor.cc:31-70
//   ORB_SLAM3::matchNNR / match                   include/LineMatcher.h:59-63, src/LineMatcher.cpp:139-229
//   ORB_SLAM3::ORBmatcher::DescriptorDistance     include/ORBmatcher.h:42, src/ORBmatcher.cc:2495-2511
//   ORB_SLAM3::StereoFrontend                     stands in for Frame::ComputeStereoMatches / _Lines (Frame.h:152,154)
// Header-only.  Without OpenCV headers the image / descriptor arguments are plf::Mat8 views; define PLF_WITH_OPENCV
// (and include <opencv2/core.hpp> first) to get cv::Mat / cv::KeyPoint / KeyLine overloads — the PODs are layout
// identical, so the conversion is a reinterpret_cast.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/plf_b200.h"

namespace plf {

struct Mat8 {                       // non-owning view of an 8-bit single-channel image or an N x 32 descriptor block
    const uint8_t* data = nullptr;
    int rows = 0, cols = 0, step = 0;
    Mat8() {}
    Mat8(const uint8_t* d, int r, int c, int s = 0) : data(d), rows(r), cols(c), step(s ? s : c) {}
    bool empty() const { return !data || rows <= 0 || cols <= 0; }
};

struct Desc {                       // owning N x 32 descriptor matrix (CV_8U rows)
    std::vector<uint8_t> bytes;
    int rows = 0;
    Mat8 view() const { return Mat8(bytes.data(), rows, 32, 32); }
    const uint8_t* row(int i) const { return bytes.data() + (size_t)i * 32; }
};

inline void check(int rc, const char* what) {
    if (rc != PLF_OK) throw std::runtime_error(std::string(what) + ": " + plf_last_error());
}

// One plf_ctx = the two ORBextractor + two Lineextractor objects of a Tracking instance (src/Tracking.cc:87-98,743-749).
class Context {
public:
    explicit Context(const plf_params& p, int device = 0) : params(p) { check(plf_create(&p, device, &ctx_), "plf_create"); }
    ~Context() { plf_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    plf_ctx* get() const { return ctx_; }
    plf_params params;
private:
    plf_ctx* ctx_ = nullptr;
};

}  // namespace plf

namespace ORB_SLAM3 {

using KeyPoint = plf_keypoint;      // == cv::KeyPoint
using KeyLine = plf_keyline;        // == cv::line_descriptor::KeyLine

class ORBextractor {
public:
    // Reference ctor: ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), src/ORBextractor.cc:408.
    // `side` selects the left (0) or right (1) extractor of the shared context.
    ORBextractor(std::shared_ptr<plf::Context> ctx, int side) : ctx_(ctx), side_(side) {
        const int L = ctx->params.n_levels;
        mvScaleFactor.resize(L); mvInvScaleFactor.resize(L); mvLevelSigma2.resize(L); mvInvLevelSigma2.resize(L);
        mnFeaturesPerLevel.resize(L);
        plf::check(plf_get_scale_tables(ctx->get(), mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                        mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "plf_get_scale_tables");
    }
    // int operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray descriptors, vector<int>& vLappingArea)
    // returns monoIndex; -1 on an empty image (src/ORBextractor.cc:1072-1073).  The mask is ignored, as in the reference.
    int operator()(const plf::Mat8& image, const plf::Mat8& /*mask*/, std::vector<KeyPoint>& keypoints,
                   plf::Desc& descriptors, std::vector<int>& vLappingArea) {
        if (image.empty()) return -1;
        const int cap = plf_keypoint_capacity(ctx_->get());
        keypoints.resize(cap);
        descriptors.bytes.resize((size_t)cap * 32);
        int n = 0, mono = 0;
        plf::check(plf_orb_extract(ctx_->get(), side_, image.data, image.cols, image.rows, image.step, vLappingArea.at(0),
                                   vLappingArea.at(1), keypoints.data(), descriptors.bytes.data(), cap, &n, &mono),
                   "plf_orb_extract");
        keypoints.resize(n);
        descriptors.bytes.resize((size_t)n * 32);
        descriptors.rows = n;
        return mono;
    }
    int GetLevels() const { return ctx_->params.n_levels; }
    float GetScaleFactor() const { return ctx_->params.scale_factor; }
    std::vector<float> GetScaleFactors() const { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() const { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() const { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() const { return mvInvLevelSigma2; }
    // mvImagePyramid[level] (include/ORBextractor.h:87): copied out of device memory on demand
    std::vector<uint8_t> ImagePyramidLevel(int level, int* w, int* h) const {
        plf::check(plf_get_pyramid_level(ctx_->get(), side_, level, nullptr, 0, w, h), "plf_get_pyramid_level");
        std::vector<uint8_t> out((size_t)*w * *h);
        plf::check(plf_get_pyramid_level(ctx_->get(), side_, level, out.data(), *w, w, h), "plf_get_pyramid_level");
        return out;
    }
private:
    std::shared_ptr<plf::Context> ctx_;
    int side_;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;
};

class Lineextractor {
public:
    Lineextractor(std::shared_ptr<plf::Context> ctx, int side) : ctx_(ctx), side_(side) {}
    // void operator()(const Mat& image, const Mat& mask, vector<KeyLine>& keylines, Mat& descriptors_line)
    void operator()(const plf::Mat8& image, const plf::Mat8& /*mask*/, std::vector<KeyLine>& keylines,
                    plf::Desc& descriptors_line) {
        keylines.clear();                                   // src/LineExtractor.cc:35
        if (!ctx_->params.has_lines) return;                // Config::hasLines()
        const int cap = plf_keyline_capacity(ctx_->get());
        keylines.resize(cap);
        descriptors_line.bytes.resize((size_t)cap * 32);
        int n = 0;
        plf::check(plf_line_extract(ctx_->get(), side_, image.data, image.cols, image.rows, image.step, keylines.data(),
                                    descriptors_line.bytes.data(), cap, &n), "plf_line_extract");
        keylines.resize(n);
        descriptors_line.bytes.resize((size_t)n * 32);
        descriptors_line.rows = n;
    }
private:
    std::shared_ptr<plf::Context> ctx_;
    int side_;
};

class ORBmatcher {
public:
    static const int TH_LOW = 50, TH_HIGH = 100;            // src/ORBmatcher.cc:36-37
    static int DescriptorDistance(const uint8_t* a, const uint8_t* b) { return plf_hamming256(a, b); }
};

// int matchNNR(const Mat& desc1, const Mat& desc2, float nnr, vector<int>& matches_12)
inline int matchNNR(plf::Context& ctx, const plf::Mat8& desc1, const plf::Mat8& desc2, float nnr, std::vector<int>& matches_12) {
    matches_12.assign(desc1.rows, -1);
    int n = 0;
    plf::check(plf_match_nnr(ctx.get(), desc1.data, desc1.rows, desc2.data, desc2.rows, nnr, matches_12.data(), &n), "plf_match_nnr");
    return n;
}
// int match(const Mat& desc1, const Mat& desc2, float nnr, vector<int>& matches_12) — mutual best when Config::bestLRMatches()
inline int match(plf::Context& ctx, const plf::Mat8& desc1, const plf::Mat8& desc2, float nnr, std::vector<int>& matches_12) {
    matches_12.assign(desc1.rows, -1);
    int n = 0;
    plf::check(plf_match(ctx.get(), desc1.data, desc1.rows, desc2.data, desc2.rows, nnr, ctx.params.best_lr_matches,
                         matches_12.data(), &n), "plf_match");
    return n;
}

// The two stereo members of Frame: fills mvuRight/mvDepth and mvDisparity_l/mvle_l from the state the four extractor
// calls left in the context.
class StereoFrontend {
public:
    explicit StereoFrontend(std::shared_ptr<plf::Context> ctx) : ctx_(ctx) {}
    void ComputeStereoMatches(int N, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
        const int cap = plf_keypoint_capacity(ctx_->get());
        mvuRight.assign(cap, -1.f);
        mvDepth.assign(cap, -1.f);
        plf::check(plf_stereo_match_points(ctx_->get(), mvuRight.data(), mvDepth.data(), cap), "plf_stereo_match_points");
        mvuRight.resize(N);
        mvDepth.resize(N);
    }
    void ComputeStereoMatches_Lines(int N_l, std::vector<std::pair<float, float>>& mvDisparity_l,
                                    std::vector<std::array<double, 3>>& mvle_l) {
        const int cap = plf_keyline_capacity(ctx_->get());
        std::vector<float> disp((size_t)cap * 2, -1.f);
        std::vector<double> le((size_t)cap * 3, 0.0);
        plf::check(plf_stereo_match_lines(ctx_->get(), disp.data(), le.data(), nullptr, cap), "plf_stereo_match_lines");
        mvDisparity_l.resize(N_l);
        mvle_l.resize(N_l);
        for (int i = 0; i < N_l; ++i) {
            mvDisparity_l[i] = {disp[2 * i], disp[2 * i + 1]};
            mvle_l[i] = {le[3 * i], le[3 * i + 1], le[3 * i + 2]};
        }
    }
private:
    std::shared_ptr<plf::Context> ctx_;
};

// The rectification in front of Frame (Examples/Stereo/stereo_euroc.cc:117-118,166-167):
//   cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F, M1, M2)  ->  Rectifier::setMaps(side, M1, M2, src size)
//   cv::remap(im, imRect, M1, M2, cv::INTER_LINEAR)                ->  Rectifier::remap(side, im, imRect)
class Rectifier {
public:
    explicit Rectifier(std::shared_ptr<plf::Context> ctx) : ctx_(ctx) {}
    void setMaps(int side, const float* M1, const float* M2, int srcCols, int srcRows) {
        plf::check(plf_rectify_set_maps(ctx_->get(), side, M1, M2, srcCols, srcRows), "plf_rectify_set_maps");
    }
    void remap(int side, const plf::Mat8& im, std::vector<uint8_t>& imRect) {
        const int w = ctx_->params.width, h = ctx_->params.height;
        imRect.resize((size_t)w * h);
        plf::check(plf_rectify(ctx_->get(), side, im.data, im.step, imRect.data(), w), "plf_rectify");
    }
private:
    std::shared_ptr<plf::Context> ctx_;
};

}  // namespace ORB_SLAM3
