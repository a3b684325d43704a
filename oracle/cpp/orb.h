// TEST INFRASTRUCTURE ONLY — CPU oracle, ORB part (restates src/ORBextractor.cc of the reference).
#pragma once
#include "prims.h"
#include "../../include/plf_b200.h"
#include <vector>

namespace plfo {

struct Cand { float x, y, resp; };   // x,y relative to (minBorderX,minBorderY) like vToDistributeKeys

struct OrbConfig {
    int nfeatures = 1200;
    float scaleFactor = 1.2f;
    int nlevels = 8;
    int iniThFAST = 20, minThFAST = 7;
};

struct OrbTables {
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> nPerLevel;
    int umax[16];
};

struct OrbState {
    std::vector<Img8> pyr, blur;
    std::vector<std::vector<Cand>> cands;     // per level, tap
    std::vector<plf_keypoint> kps;            // final row order
    std::vector<uint8_t> desc;                // kps.size() x 32
    int monoIndex = 0;
    bool valid = false;
};

void orb_tables(const OrbConfig& c, OrbTables& t);
// FAST-9/16 + score + strict 3x3 NMS on the window [x0,x1) x [y0,y1) of img (cv::FAST(sub, kps, th, true)).
// Appends (x - x0, y - y0, score) in raster order.
void fast_window(const Img8& img, int x0, int y0, int x1, int y1, int th, std::vector<Cand>& out);
void level_candidates(const Img8& lvl, int iniTh, int minTh, std::vector<Cand>& out);
void distribute_octree(const std::vector<Cand>& in, int minX, int maxX, int minY, int maxY, int N,
                       std::vector<Cand>& out);
float ic_angle(const Img8& lvl, int x, int y, const int* umax);
void orb_descriptor(const Img8& blur, int x, int y, float angleDeg, uint8_t* desc);
int orb_extract(const OrbConfig& c, const OrbTables& t, const uint8_t* img, int w, int h, int stride, int lap0,
                int lap1, OrbState& st);

}  // namespace plfo
