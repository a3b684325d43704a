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
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
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

// Multi-GPU: replicas only (every stereo pair is independent).  Stream s is served by rank / GPU s mod world; the Python twin
// and the per-rank context pool are pli-slam_b200/pool.py.
inline int plf_stream_owner(int stream, int world) { return world > 0 ? stream % world : 0; }

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

// The frame-level pieces after the stereo matchers: Frame::AssignFeaturesToGrid / GetFeaturesInArea
// (src/Frame.cc:451-482, 774-843), Frame::UnprojectStereo / backProjection (:1332-1358), Frame::ComputeBoW (:858-870) and the
// ORBmatcher searches of the tracking / relocalisation / loop-closing threads (src/ORBmatcher.cc:44-130, 2179-2323, 2325-2447,
// 473-704, SearchByBoW 269-471) and the gated line matching of the tracking thread (src/Tracking.cc:3055-3099, 3879-3917).
class FrameTail {
public:
    explicit FrameTail(std::shared_ptr<plf::Context> ctx, int slot = 0) : ctx_(ctx), slot_(slot) {}
    // void Frame::AssignFeaturesToGrid(): keeps mGrid[64][48] as CSR
    void AssignFeaturesToGrid() {
        cellStart_.assign(PLF_GRID_COLS * PLF_GRID_ROWS + 1, 0);
        cellIdx_.assign((size_t)plf_keypoint_capacity(ctx_->get()), -1);
        plf::check(plf_feature_grid(ctx_->get(), slot_, 1, cellStart_.data(), cellIdx_.data(), (int)cellIdx_.size()), "plf_feature_grid");
    }
    // vector<size_t> Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel)
    std::vector<size_t> GetFeaturesInArea(const std::vector<KeyPoint>& mvKeysUn, float x, float y, float r, int minLevel = -1,
                                          int maxLevel = -1) const {
        std::vector<int32_t> idx(mvKeysUn.size() + 1);
        const int n = plf_features_in_area(mvKeysUn.data(), cellStart_.data(), cellIdx_.data(), ctx_->params.width, ctx_->params.height,
                                           x, y, r, minLevel, maxLevel, idx.data(), (int)idx.size());
        return std::vector<size_t>(idx.begin(), idx.begin() + n);
    }
    // cv::Mat Frame::UnprojectStereo(i) for every keypoint and Frame::backProjection for both ends of every line
    void BackProject(const float Rwc[9], const float Ow[3], float fy, float cx, float cy, int N, int N_l, std::vector<float>& x3D,
                     std::vector<double>& lines3D) {
        x3D.assign((size_t)3 * N, 0.f);
        lines3D.assign((size_t)6 * N_l, 0.0);
        plf::check(plf_backproject(ctx_->get(), slot_, 1, Rwc, Ow, fy, cx, cy, N ? x3D.data() : nullptr, N, N_l ? lines3D.data() : nullptr, N_l),
                   "plf_backproject");
    }
    // mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) / the line vocabulary (which = 1)
    void ComputeBoW(int which, int nFeatures, std::vector<std::pair<int32_t, double>>& bowVec,
                    std::vector<std::pair<int32_t, std::vector<int32_t>>>& featVec, int levelsup = 4) {
        const int n = nFeatures;
        std::vector<int32_t> w((size_t)n + 1), nid((size_t)n + 1), bw((size_t)n + 1), fn((size_t)n + 1), fs((size_t)n + 2), ff((size_t)n + 1);
        std::vector<double> v((size_t)n + 1), bv((size_t)n + 1);
        bowVec.clear(); featVec.clear();
        if (n <= 0) return;
        plf::check(plf_bow_transform(ctx_->get(), which, slot_, 1, levelsup, w.data(), v.data(), nid.data(), n), "plf_bow_transform");
        int nNodes = 0;
        const int nWords = plf_bow_build(w.data(), v.data(), nid.data(), n, bw.data(), bv.data(), fn.data(), fs.data(), ff.data(), &nNodes);
        for (int j = 0; j < nWords; ++j) bowVec.emplace_back(bw[j], bv[j]);
        for (int j = 0; j < nNodes; ++j) featVec.emplace_back(fn[j], std::vector<int32_t>(ff.begin() + fs[j], ff.begin() + fs[j + 1]));
    }
    // int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th): match[i] = feature given to map point i or -1
    int SearchByProjection(const std::vector<plf_proj_query>& mapPoints, float th, float mfNNratio, std::vector<uint8_t>& occupied,
                           std::vector<int32_t>& match, int TH_HIGH = 100) {
        match.assign(mapPoints.size(), -1);
        int n = 0;
        plf::check(plf_search_by_projection(ctx_->get(), slot_, mapPoints.data(), (int)mapPoints.size(), th, mfNNratio, TH_HIGH,
                                            occupied.data(), (int)occupied.size(), match.data(), &n), "plf_search_by_projection");
        return n;
    }
    // int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono, match12), projection done by the caller
    int SearchByProjection(const std::vector<plf_frame_query>& lastFramePoints, bool mbCheckOrientation, std::vector<uint8_t>& occupied,
                           std::vector<int32_t>& featQuery, std::vector<int32_t>& match12, int TH_HIGH = 100) {
        featQuery.assign(occupied.size(), -1);
        match12.assign(occupied.size(), -1);
        int n = 0;
        plf::check(plf_search_by_projection_frame(ctx_->get(), slot_, lastFramePoints.data(), (int)lastFramePoints.size(), TH_HIGH,
                                                  mbCheckOrientation ? 1 : 0, occupied.data(), (int)occupied.size(), featQuery.data(),
                                                  match12.data(), &n),
                   "plf_search_by_projection_frame");
        return n;
    }
    // int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)
    // (relocalisation, src/ORBmatcher.cc:2325-2447), projection done by the caller
    int SearchByProjectionReloc(const std::vector<plf_frame_query>& keyFramePoints, int ORBdist, bool mbCheckOrientation,
                                std::vector<uint8_t>& occupied, std::vector<int32_t>& featQuery) {
        featQuery.assign(occupied.size(), -1);
        int n = 0;
        plf::check(plf_search_by_projection_reloc(ctx_->get(), slot_, keyFramePoints.data(), (int)keyFramePoints.size(), ORBdist,
                                                  mbCheckOrientation ? 1 : 0, occupied.data(), (int)occupied.size(), featQuery.data(), &n),
                   "plf_search_by_projection_reloc");
        return n;
    }
    // int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th, ratioHamming) and its vpMatchedKF
    // variant (loop closing, src/ORBmatcher.cc:473-704); the slot plays pKF
    int SearchByProjectionLoop(const std::vector<plf_frame_query>& candidatePoints, float ratioHamming, std::vector<uint8_t>& vpMatchedSet,
                               std::vector<int32_t>& featQuery, int TH_LOW = 50) {
        featQuery.assign(vpMatchedSet.size(), -1);
        int n = 0;
        plf::check(plf_search_by_projection_loop(ctx_->get(), slot_, candidatePoints.data(), (int)candidatePoints.size(), TH_LOW,
                                                 ratioHamming, vpMatchedSet.data(), (int)vpMatchedSet.size(), featQuery.data(), &n),
                   "plf_search_by_projection_loop");
        return n;
    }
    // int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) (src/ORBmatcher.cc:269-471):
    // vpMapPointMatches[f] = index of the keyframe feature whose map point frame feature f received, or -1
    int SearchByBoW(const plf::Mat8& kfDescriptors, const std::vector<float>& kfAngles, const std::vector<int32_t>& kfNodes,
                    const std::vector<uint8_t>& kfHasGoodMapPoint, const std::vector<int32_t>& frameNodes, float mfNNratio,
                    bool mbCheckOrientation, std::vector<int32_t>& vpMapPointMatches, int TH_LOW = 50) {
        vpMapPointMatches.assign(frameNodes.size(), -1);
        int n = 0;
        plf::check(plf_search_by_bow(ctx_->get(), slot_, kfDescriptors.data, kfAngles.data(), kfNodes.data(), kfHasGoodMapPoint.data(),
                                     kfDescriptors.rows, frameNodes.data(), (int)frameNodes.size(), TH_LOW, mfNNratio,
                                     mbCheckOrientation ? 1 : 0, vpMapPointMatches.data(), &n),
                   "plf_search_by_bow");
        return n;
    }
    // match(mLastFrame.mDescriptors_Line, mCurrentFrame.mDescriptors_Line, minRatio12L, matches_12) + the gates of
    // Tracking::TrackWithMotionModel (src/Tracking.cc:3055-3099; mode 0) / match(mvpLocalMapLines_InFrustum, mCurrentFrame, ...) +
    // the gates of Tracking::SearchLocalLines (:3879-3917; mode 1).  Returns the number of map lines attached.
    int MatchLinesTracked(int mode, const plf::Mat8& desc1, const std::vector<plf_track_line>& lines1, const plf::Mat8& desc2,
                          const std::vector<KeyLine>& mvKeysUn_Line, const std::vector<std::pair<float, float>>& mvDisparity_l,
                          const std::vector<uint8_t>* held2, float nnr, float mnMinX, float mnMaxX, float mnMinY, float mnMaxY,
                          std::vector<int>& matches_12, std::vector<int>& assign_12) {
        matches_12.assign(lines1.size(), -1);
        assign_12.assign(lines1.size(), -1);
        int n = 0;
        static_assert(sizeof(std::pair<float, float>) == 8, "mvDisparity_l is passed as float pairs");
        plf::check(plf_match_lines_tracked(ctx_->get(), mode, desc1.data, lines1.data(), (int)lines1.size(), desc2.data, mvKeysUn_Line.data(),
                                           reinterpret_cast<const float*>(mvDisparity_l.data()), held2 ? held2->data() : nullptr,
                                           (int)mvKeysUn_Line.size(), nnr, mnMinX, mnMaxX, mnMinY, mnMaxY, matches_12.data(),
                                           assign_12.data(), &n),
                   "plf_match_lines_tracked");
        return n;
    }
private:
    std::shared_ptr<plf::Context> ctx_;
    int slot_;
    std::vector<int32_t> cellStart_, cellIdx_;
};

}  // namespace ORB_SLAM3
