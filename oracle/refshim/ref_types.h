// TEST INFRASTRUCTURE ONLY — declarations that stand in for include/Frame.h, include/ORBmatcher.h, include/MapPoint.h
// and include/MapLine.h of the reference when its frontend sources are compiled into oracle/_ref
// (oracle/build_ref.py passes -DFRAME_H -DORBMATCHER_H -DMAPPOINT_H -DMAPLINE_H -DKEYFRAME_H and force-includes this
// file).  Those headers drag in DBoW2, g2o, Pangolin-era types and the whole map; the frontend FUNCTION BODIES taken
// from src/Frame.cc:976-1307 and src/ORBmatcher.cc:36-42,2495-2511 only need the members declared here, under the
// reference's own names (include/Frame.h:152-154,218-252,316,373-374; include/ORBmatcher.h:36-42,102-104).  The bodies of
// ORBmatcher::SearchByBoW / ComputeThreeMaxima (src/ORBmatcher.cc:269-471,2449-2490) and the line-match gates of the
// tracking thread (src/Tracking.cc:3055-3099,3879-3919) additionally need the KeyFrame / MapPoint / MapLine members and
// DBoW2::FeatureVector (the reference's own header, Thirdparty/DBoW2/DBoW2/FeatureVector.h) declared below
// (include/KeyFrame.h, include/MapPoint.h, include/MapLine.h).
#pragma once
#include <climits>
#include <list>
#include <map>
#include <set>
#include <thread>
#include <vector>
#include "Auxiliar.h"        // the reference's own header (using-directives for cv / line_descriptor / std / Eigen)
#include "Config.h"          // the reference's own header
#include "ORBextractor.h"    // the reference's own header
#include "LineExtractor.h"   // the reference's own header
#include "gridStructure.h"   // the reference's own header

#define FRAME_GRID_ROWS 48   // include/Frame.h:59
#define FRAME_GRID_COLS 64   // include/Frame.h:60

namespace ORB_SLAM3 {

}  // namespace ORB_SLAM3
#include "DBoW2/BowVector.h"      // the reference's own vendored DBoW2 (Thirdparty/DBoW2/DBoW2; boost::serialization is a stub)
#include "DBoW2/FeatureVector.h"
namespace ORB_SLAM3 {

// a pinhole with fx = fy = 1, cx = cy = 0: the pins choose world points (u z, v z, z) with z a power of two, so the
// projection is exactly the (u, v) the test prescribes
class GeometricCamera {
public:
    cv::Point2f project(const cv::Mat& m) { return cv::Point2f(m.at<float>(0) / m.at<float>(2), m.at<float>(1) / m.at<float>(2)); }
    cv::Point2f project(const cv::Point3f& p) { return cv::Point2f(p.x / p.z, p.y / p.z); }
};
class Frame;
class KeyFrame;
class MapPoint {
public:
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    float GetMinDistanceInvariance() { return mfMinDistance; }
    float GetMaxDistanceInvariance() { return mfMaxDistance; }
    int PredictScale(const float&, Frame*) { return mnPredictedLevel; }      // the level the test prescribes
    int PredictScale(const float&, KeyFrame*) { return mnPredictedLevel; }
    bool mbBad = false, mbTrackInView = false, mbTrackInViewR = false;
    int nObs = 0, mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0, mnPredictedLevel = 0;
    float mTrackDepth = 0, mTrackViewCos = 0, mTrackViewCosR = 0, mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    float mfMinDistance = 0.f, mfMaxDistance = 3.0e38f;
    cv::Mat mDescriptor, mWorldPos, mNormalVector;
};
class KeyFrame {
public:
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeysUn, mvKeys, mvKeysRight;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int NLeft = -1, N = 0;
    float fx = 1, fy = 1, cx = 0, cy = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;                 // include/KeyFrame.h: const int
    std::vector<std::vector<std::vector<size_t> > > mGrid, mGridRight;
    std::vector<float> mvScaleFactors;
};
class MapLine {
public:
    cv::Mat GetDescriptor() { return mLDescriptor.clone(); }
    int Observations() { return nObs; }
    cv::Mat mLDescriptor;
    float mTrackProjsX = 0, mTrackProjsY = 0, mTrackProjeX = 0, mTrackProjeY = 0;
    int nObs = 0;
};
class Frame;

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, std::map<int, int>& match12);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                           float ratioHamming = 1.0);
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3

#include "LineMatcher.h"     // the reference's own header (its Frame.h / MapPoint.h / MapLine.h includes are guarded out)

namespace ORB_SLAM3 {

class Frame {
public:
    void ComputeStereoMatches();
    void ComputeStereoMatches_Lines(bool initial = true);
    double lineSegmentOverlapStereo(double spl_obs, double epl_obs, double spl_proj, double epl_proj);
    void filterLineSegmentDisparity(Vector2d spl, Vector2d epl, Vector2d spr, Vector2d epr, double& disp_s,
                                    double& disp_e);

    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float mbf = 0, mb = 0;
    int N = 0, N_l = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<KeyLine> mvKeys_Line, mvKeysRight_Line;
    cv::Mat mDescriptors_Line, mDescriptorsRight_Line;
    std::vector<std::pair<float, float>> mvDisparity_l;
    std::vector<Vector3d> mvle_l;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    double inv_width = 0, inv_height = 0;
    // members read by SearchByBoW and by the tracking thread's line gates
    DBoW2::FeatureVector mFeatVec;
    int Nleft = -1;
    GeometricCamera* mpCamera2 = nullptr;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<MapLine*> mvpMapLines;
    std::vector<KeyLine> mvKeysUn_Line;
    int n_inliers_ls = 0;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    // members read by the SearchByProjection overloads and the feature grid (include/Frame.h:196-264)
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                          const bool bRight = false) const;
    cv::Mat mTcw;
    static float fx, fy, cx, cy, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    GeometricCamera* mpCamera = nullptr;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};

// the two loops of the tracking thread, generated from src/Tracking.cc by oracle/build_ref.py
int ref_track_gate_f2f(Frame& mCurrentFrame, Frame& mLastFrame, std::vector<int>& matches_out);
void ref_track_gate_local(Frame& mCurrentFrame, std::vector<MapLine*>& mvpLocalMapLines_InFrustum, int nToMatch,
                          std::vector<int>& matches_out);

}  // namespace ORB_SLAM3
