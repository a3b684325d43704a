// TEST INFRASTRUCTURE ONLY — declarations that stand in for include/Frame.h, include/ORBmatcher.h, include/MapPoint.h
// and include/MapLine.h of the reference when its frontend sources are compiled into oracle/_ref
// (oracle/build_ref.py passes -DFRAME_H -DORBMATCHER_H -DMAPPOINT_H -DMAPLINE_H -DKEYFRAME_H and force-includes this
// file).  Those headers drag in DBoW2, g2o, Pangolin-era types and the whole map; the frontend FUNCTION BODIES taken
// from src/Frame.cc:976-1307 and src/ORBmatcher.cc:36-42,2495-2511 only need the members declared here, under the
// reference's own names (include/Frame.h:152-154,218-252,316,373-374; include/ORBmatcher.h:36-42,102-104).
#pragma once
#include <climits>
#include <list>
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

class MapPoint {};
class KeyFrame {};
class MapLine {
public:
    cv::Mat GetDescriptor() { return mLDescriptor.clone(); }
    cv::Mat mLDescriptor;
};

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
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
};

}  // namespace ORB_SLAM3
