// TEST INFRASTRUCTURE ONLY — C entry points of oracle/_ref/libplf_ref.so: the REFERENCE'S OWN frontend code (compiled
// from /root/reference by oracle/build_ref.py, nothing of it copied into this repository) behind plain-C calls, so that
// tests/test_oracle_ref.py can pin the restated oracle (oracle/cpp) to it.  This file only marshals arrays in and
// out of the reference's classes; it contains no frontend logic of its own.
//
// Heap addresses: DistributeOctTree breaks size ties by ExtractorNode* (src/ORBextractor.cc:682), i.e. by where the
// std::list nodes happen to live.  ref_arena(1) serves exactly the list-node allocations of a call from a bump arena
// (monotonically increasing addresses, the "most recently created node first" rule the oracle declares); ref_arena(0)
// leaves them to malloc (whatever glibc does on this machine).
#include "ref_types.h"
#include "DBoW2/FORB.h"                  // the reference's own vendored DBoW2 (Thirdparty/DBoW2)
#include "DBoW2/TemplatedVocabulary.h"
#include <cstdlib>
#include <new>
#include <string>

using namespace ORB_SLAM3;

#define REF_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------------------------
// bump arena for std::list<ExtractorNode> nodes
namespace {
struct ListNodeProbe { void* a; void* b; ExtractorNode n; };   // layout of std::_List_node<ExtractorNode>
const size_t kNodeSize = sizeof(ListNodeProbe);
const size_t kArenaCap = (size_t)256 << 20;
char* g_arena = nullptr;
size_t g_off = 0;
bool g_arena_enabled = true;
thread_local bool g_in_call = false;
std::string g_err;

struct CallScope {
    CallScope() { g_off = 0; g_in_call = true; }
    ~CallScope() { g_in_call = false; }
};
inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArenaCap; }
}  // namespace

void* operator new(size_t n) {
    if (g_in_call && g_arena_enabled && n == kNodeSize) {
        if (!g_arena) g_arena = (char*)std::malloc(kArenaCap);
        size_t a = (g_off + 15) & ~(size_t)15;
        if (g_arena && a + n <= kArenaCap) { g_off = a + n; return g_arena + a; }
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void* p) noexcept { if (p && !in_arena(p)) std::free(p); }
void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) std::free(p); }

namespace {

struct OrbPeek : public ORBextractor {
    using ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> octree(const std::vector<cv::KeyPoint>& v, int minX, int maxX, int minY, int maxY, int N) {
        return DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    }
    const std::vector<int>& quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int>& umaxTable() const { return umax; }
};

OrbPeek* g_orb[2] = {nullptr, nullptr};
Frame g_frame;

cv::Mat wrap_u8(const uint8_t* p, int w, int h, int stride) { return cv::Mat(h, w, CV_8UC1, (void*)p, (size_t)stride); }
cv::Mat wrap_desc(const uint8_t* p, int n) { return cv::Mat(n, 32, CV_8UC1, (void*)p, 32); }

template <typename F> int guarded(F f) {
    try { return f(); }
    catch (const std::exception& e) { g_err = e.what(); return -1000; }
    catch (...) { g_err = "unknown exception"; return -1001; }
}

}  // namespace

REF_API const char* ref_last_error() { return g_err.c_str(); }
REF_API int ref_arena(int on) { g_arena_enabled = on != 0; return (int)kNodeSize; }

// Config singleton of the reference (include/Config.h, src/Config.cpp): v[] = hasLines, bestLRMatches, matchingSWs,
// minRatio12L, lineSimTh, minDisp, lineHorizTh, stereoOverlapTh, lsMinDispRatio, minRatio12P, lrInParallel,
// lsdNFeatures, lsdRefine, lsdScale, lsdSigmaScale, lsdQuant, lsdAngTh, lsdLogEps, lsdDensityTh, lsdNBins, minLineLength
static void config_read(double* v) {
    v[0] = Config::hasLines(); v[1] = Config::bestLRMatches(); v[2] = Config::matchingSWs(); v[3] = Config::minRatio12L();
    v[4] = Config::lineSimTh(); v[5] = Config::minDisp(); v[6] = Config::lineHorizTh(); v[7] = Config::stereoOverlapTh();
    v[8] = Config::lsMinDispRatio(); v[9] = Config::minRatio12P(); v[10] = Config::lrInParallel();
    v[11] = Config::lsdNFeatures(); v[12] = Config::lsdRefine(); v[13] = Config::lsdScale(); v[14] = Config::lsdSigmaScale();
    v[15] = Config::lsdQuant(); v[16] = Config::lsdAngTh(); v[17] = Config::lsdLogEps(); v[18] = Config::lsdDensityTh();
    v[19] = Config::lsdNBins(); v[20] = Config::minLineLength();
}
REF_API int ref_config_get(double* v21) { return guarded([&] { config_read(v21); return 0; }); }
REF_API int ref_config_load(const char* yaml, double* v21) {
    return guarded([&] { Config::loadFromFile(yaml); config_read(v21); return 0; });
}
REF_API int ref_config_set(const double* v) {
    return guarded([&] {
        Config::hasLines() = v[0] != 0; Config::bestLRMatches() = v[1] != 0; Config::matchingSWs() = (int)v[2];
        Config::minRatio12L() = v[3]; Config::lineSimTh() = v[4]; Config::minDisp() = v[5]; Config::lineHorizTh() = v[6];
        Config::stereoOverlapTh() = v[7]; Config::lsMinDispRatio() = v[8]; Config::minRatio12P() = v[9];
        Config::lrInParallel() = v[10] != 0;
        return 0;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// ORBextractor (src/ORBextractor.cc, whole file)
REF_API int ref_orb_create(int side, int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return guarded([&] {
        delete g_orb[side];
        g_orb[side] = new OrbPeek(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        return 0;
    });
}
REF_API int ref_orb_tables(int side, float* scale, float* invScale, float* sigma2, float* invSigma2, int* quotas, int* umax16) {
    return guarded([&] {
        OrbPeek* e = g_orb[side];
        int n = e->GetLevels();
        std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                           d = e->GetInverseScaleSigmaSquares();
        for (int i = 0; i < n; i++) { scale[i] = a[i]; invScale[i] = b[i]; sigma2[i] = c[i]; invSigma2[i] = d[i];
                                      quotas[i] = e->quotas()[i]; }
        for (int i = 0; i < 16; i++) umax16[i] = e->umaxTable()[i];
        return n;
    });
}
// ORBextractor::operator(): returns monoIndex (or -1 for an empty image); *n = keypoints written.
REF_API int ref_orb_extract(int side, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                            cv::KeyPoint* kp_out, uint8_t* desc_out, int cap, int* n) {
    return guarded([&] {
        CallScope scope;
        cv::Mat im = (img && w > 0 && h > 0) ? wrap_u8(img, w, h, stride) : cv::Mat();
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        std::vector<int> lap = {lap0, lap1};
        int mono = (*g_orb[side])(im, cv::Mat(), kps, desc, lap);
        *n = (int)kps.size();
        if (*n > cap) throw std::runtime_error("ref_orb_extract: capacity");
        for (int i = 0; i < *n; i++) { kp_out[i] = kps[i]; memcpy(desc_out + 32 * i, desc.ptr(i), 32); }
        return mono;
    });
}
REF_API int ref_orb_level(int side, int level, uint8_t* out, int cap, int* w, int* h) {
    return guarded([&] {
        const cv::Mat& m = g_orb[side]->mvImagePyramid[level];
        *w = m.cols; *h = m.rows;
        if (m.cols * m.rows > cap) throw std::runtime_error("ref_orb_level: capacity");
        for (int y = 0; y < m.rows; y++) memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
        return 0;
    });
}
// ORBextractor::DistributeOctTree on a caller-supplied candidate list (x, y, response triples).
REF_API int ref_octree(int side, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out, int cap) {
    return guarded([&] {
        CallScope scope;
        std::vector<cv::KeyPoint> v(n);
        for (int i = 0; i < n; i++) v[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]);
        std::vector<cv::KeyPoint> r = g_orb[side]->octree(v, minX, maxX, minY, maxY, N);
        if ((int)r.size() > cap) throw std::runtime_error("ref_octree: capacity");
        for (size_t i = 0; i < r.size(); i++) { out[3 * i] = r[i].pt.x; out[3 * i + 1] = r[i].pt.y; out[3 * i + 2] = r[i].response; }
        return (int)r.size();
    });
}

// ---------------------------------------------------------------------------------------------------------------
// Lineextractor::operator() (src/LineExtractor.cc whole file -> LSDDetector_custom.cpp whole file -> LBD ranges)
REF_API int ref_line_extract(const uint8_t* img, int w, int h, int stride, int nfeatures, double min_line_length, int refine,
                             double scale, double sigma_scale, double quant, double ang_th, double log_eps,
                             double density_th, int n_bins, KeyLine* kl_out, uint8_t* desc_out, int cap) {
    return guarded([&] {
        Lineextractor ex(nfeatures, min_line_length, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins);
        std::vector<KeyLine> kls;
        cv::Mat desc;
        ex(wrap_u8(img, w, h, stride), cv::Mat(), kls, desc);
        if ((int)kls.size() > cap) throw std::runtime_error("ref_line_extract: capacity");
        for (size_t i = 0; i < kls.size(); i++) { kl_out[i] = kls[i]; memcpy(desc_out + 32 * i, desc.ptr((int)i), 32); }
        return (int)kls.size();
    });
}
// BinaryDescriptor::compute on caller-supplied KeyLines: 72-float LBD (returnFloatDescr) and the 32-byte binary form.
REF_API int ref_lbd(const uint8_t* img, int w, int h, int stride, const KeyLine* kls, int n, float* lbd72, uint8_t* desc) {
    return guarded([&] {
        cv::Ptr<BinaryDescriptor> lbd = BinaryDescriptor::createBinaryDescriptor();
        std::vector<KeyLine> v(kls, kls + n);
        cv::Mat im = wrap_u8(img, w, h, stride), d, f;
        if (desc) { lbd->compute(im, v, d, false); for (int i = 0; i < n; i++) memcpy(desc + 32 * i, d.ptr(i), 32); }
        if (lbd72) { lbd->compute(im, v, f, true); for (int i = 0; i < n; i++) memcpy(lbd72 + 72 * i, f.ptr(i), 72 * 4); }
        return n;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (src/Frame.cc:976-1154) on the pyramids the two extractors hold from their last call.
REF_API int ref_stereo_points(float bf, float fx, const cv::KeyPoint* kpL, int nL, const uint8_t* dL, const cv::KeyPoint* kpR,
                              int nR, const uint8_t* dR, float* uRight, float* depth) {
    return guarded([&] {
        if (nL == 0) return 0;
        Frame& F = g_frame;
        F.mpORBextractorLeft = g_orb[0]; F.mpORBextractorRight = g_orb[1];
        F.mvScaleFactors = g_orb[0]->GetScaleFactors(); F.mvInvScaleFactors = g_orb[0]->GetInverseScaleFactors();
        F.mbf = bf; F.mb = bf / fx;        // src/Frame.cc:1006 reads mb before :197 sets it; declared rule mb := mbf/fx
        F.N = nL;
        F.mvKeys.assign(kpL, kpL + nL); F.mvKeysRight.assign(kpR, kpR + nR);
        F.mDescriptors = wrap_desc(dL, nL).clone(); F.mDescriptorsRight = wrap_desc(dR, nR).clone();
        F.ComputeStereoMatches();
        for (int i = 0; i < nL; i++) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
        return 0;
    });
}

// Frame::ComputeStereoMatches_Lines (src/Frame.cc:1156-1307) -> matchGrid (src/LineMatcher.cpp, whole file) ->
// GridStructure / LineIterator (src/gridStructure.cpp, src/LineIterator.cpp, whole files).  matchGrid's matches_12 is
// a local of the reference function; it is recovered by a second, direct call with the same inputs.
REF_API int ref_stereo_lines(int W, int H, const KeyLine* klL, int nL, const uint8_t* dL, const KeyLine* klR, int nR,
                             const uint8_t* dR, float* disp_se, double* le) {
    return guarded([&] {
        Frame& F = g_frame;
        F.inv_width = FRAME_GRID_COLS / static_cast<double>(W);     // src/Frame.cc:109-110
        F.inv_height = FRAME_GRID_ROWS / static_cast<double>(H);
        F.mvKeys_Line.assign(klL, klL + nL); F.mvKeysRight_Line.assign(klR, klR + nR);
        F.mDescriptors_Line = nL ? wrap_desc(dL, nL).clone() : cv::Mat();
        F.mDescriptorsRight_Line = nR ? wrap_desc(dR, nR).clone() : cv::Mat();
        F.N_l = nL;
        F.ComputeStereoMatches_Lines();
        for (int i = 0; i < nL; i++) {
            disp_se[2 * i] = F.mvDisparity_l[i].first; disp_se[2 * i + 1] = F.mvDisparity_l[i].second;
            for (int k = 0; k < 3; k++) le[3 * i + k] = F.mvle_l[i](k);
        }
        return 0;
    });
}
// matchGrid(lines) called the way src/Frame.cc:1178-1205 calls it (grid of the right lines, endpoint cells of the left).
REF_API int ref_match_grid_lines(int W, int H, const KeyLine* klL, int nL, const uint8_t* dL, const KeyLine* klR, int nR,
                                 const uint8_t* dR, int* m12) {
    return guarded([&] {
        const double inv_width = FRAME_GRID_COLS / static_cast<double>(W), inv_height = FRAME_GRID_ROWS / static_cast<double>(H);
        std::vector<line_2d> coords;
        for (int i = 0; i < nL; i++)
            coords.push_back(std::make_pair(std::make_pair(klL[i].startPointX * inv_width, klL[i].startPointY * inv_height),
                                            std::make_pair(klL[i].endPointX * inv_width, klL[i].endPointY * inv_height)));
        std::list<std::pair<int, int>> line_coords;
        GridStructure grid(FRAME_GRID_ROWS, FRAME_GRID_COLS);
        std::vector<std::pair<double, double>> directions(nR);
        for (int idx = 0; idx < nR; ++idx) {
            const KeyLine& kl = klR[idx];
            directions[idx] = std::make_pair((kl.endPointX - kl.startPointX) * inv_width, (kl.endPointY - kl.startPointY) * inv_height);
            normalize(directions[idx]);
            getLineCoords(kl.startPointX * inv_width, kl.startPointY * inv_height, kl.endPointX * inv_width,
                          kl.endPointY * inv_height, line_coords);
            for (const std::pair<int, int>& p : line_coords) grid.at(p.first, p.second).push_back(idx);
        }
        GridWindow w;
        w.width = std::make_pair(Config::matchingSWs(), 0);
        w.height = std::make_pair(0, 0);
        std::vector<int> matches_12;
        int r = matchGrid(coords, wrap_desc(dL, nL), grid, wrap_desc(dR, nR), directions, w, matches_12);
        for (int i = 0; i < nL; i++) m12[i] = matches_12[i];
        return r;
    });
}
// getLineCoords (src/gridStructure.cpp:33-41 over src/LineIterator.cpp): cells as x,y pairs; returns the count.
REF_API int ref_line_coords(double x1, double y1, double x2, double y2, int* xy, int cap) {
    return guarded([&] {
        std::list<std::pair<int, int>> lc;
        getLineCoords(x1, y1, x2, y2, lc);
        int k = 0;
        for (const auto& p : lc) { if (k < cap) { xy[2 * k] = p.first; xy[2 * k + 1] = p.second; } k++; }
        return k;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// matchNNR / match / distance (src/LineMatcher.cpp:139-159,201-247), ORBmatcher::DescriptorDistance (:2495-2511)
REF_API int ref_match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12) {
    return guarded([&] {
        if (n2 < 2) throw std::runtime_error("undefined in the reference: matches_[idx][1] with fewer than 2 train rows");
        std::vector<int> m;
        int r = matchNNR(wrap_desc(d1, n1), wrap_desc(d2, n2), nnr, m);
        for (int i = 0; i < n1; i++) m12[i] = m[i];
        return r;
    });
}
REF_API int ref_match(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12) {
    return guarded([&] {
        if (n2 < 2 || (Config::bestLRMatches() && n1 < 2)) throw std::runtime_error("undefined in the reference: fewer than 2 train rows");
        std::vector<int> m;
        int r = match(wrap_desc(d1, n1), wrap_desc(d2, n2), nnr, m);
        for (int i = 0; i < n1; i++) m12[i] = m[i];
        return r;
    });
}
REF_API int ref_distance(const uint8_t* a, const uint8_t* b) { return ORB_SLAM3::distance(wrap_desc(a, 1), wrap_desc(b, 1)); }
REF_API int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(wrap_desc(a, 1), wrap_desc(b, 1));
}

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (src/ORBmatcher.cc:269-471): the two FeatureVectors are
// built from per-feature node ids (< 0: the feature is in no entry); kf_valid 0 = no map point, 2 = a bad one.
// match[f] = index of the keyframe feature whose map point frame feature f received, or -1.
REF_API int ref_search_by_bow(const uint8_t* kfDesc, const float* kfAngle, const int* kfNode, const uint8_t* kfValid, int nKF,
                              const uint8_t* fDesc, const float* fAngle, const int* fNode, int nF, float nnratio, int checkOri,
                              int* match) {
    return guarded([&] {
        std::vector<MapPoint> pts(nKF);
        KeyFrame kf;
        kf.mvpMapPoints.assign(nKF, nullptr);
        kf.mvKeysUn.resize(nKF);
        for (int i = 0; i < nKF; i++) {
            kf.mvKeysUn[i].angle = kfAngle[i];
            if (kfValid[i]) { kf.mvpMapPoints[i] = &pts[i]; pts[i].mbBad = kfValid[i] == 2; }
            if (kfNode[i] >= 0) kf.mFeatVec[(DBoW2::NodeId)kfNode[i]].push_back((unsigned)i);
        }
        kf.mDescriptors = wrap_desc(kfDesc, nKF);
        Frame F;
        F.N = nF;
        F.mvKeys.resize(nF);
        for (int i = 0; i < nF; i++) {
            F.mvKeys[i].angle = fAngle[i];
            if (fNode[i] >= 0) F.mFeatVec[(DBoW2::NodeId)fNode[i]].push_back((unsigned)i);
        }
        F.mDescriptors = wrap_desc(fDesc, nF);
        std::vector<MapPoint*> out;
        ORBmatcher matcher(nnratio, checkOri != 0);
        const int n = matcher.SearchByBoW(&kf, F, out);
        for (int f = 0; f < nF; f++) match[f] = out[f] ? (int)(out[f] - pts.data()) : -1;
        return n;
    });
}

// The line half of Tracking::TrackWithMotionModel (src/Tracking.cc:3055-3099): match() with Config::minRatio12L() and the
// orientation / position gates.  has1[i1]: LastFrame.mvpMapLines[i1] != NULL.  assign[i1] = the i2 whose mvpMapLines slot
// received the line of i1, or -1; returns n_inliers_ls; *nnr = the ratio the reference used.
REF_API int ref_track_lines_f2f(const uint8_t* d1, const KeyLine* kl1, const uint8_t* has1, int n1, const uint8_t* d2,
                                const KeyLine* kl2, const float* disp2, int n2, float minX, float maxX, float minY, float maxY,
                                int* matches12, int* assign12, float* nnr) {
    return guarded([&] {
        if (n2 < 2 || n1 < 2) throw std::runtime_error("undefined in the reference: fewer than 2 rows");
        std::vector<MapLine> lines(n1);
        Frame last, cur;
        last.mvpMapLines.assign(n1, nullptr);
        for (int i = 0; i < n1; i++) if (has1[i]) last.mvpMapLines[i] = &lines[i];
        last.mDescriptors_Line = wrap_desc(d1, n1);
        last.mvKeysUn_Line.assign(kl1, kl1 + n1);
        cur.mDescriptors_Line = wrap_desc(d2, n2);
        cur.mvKeysUn_Line.assign(kl2, kl2 + n2);
        cur.mvpMapLines.assign(n2, nullptr);
        cur.mvDisparity_l.resize(n2);
        for (int i = 0; i < n2; i++) cur.mvDisparity_l[i] = std::make_pair(disp2[2 * i], disp2[2 * i + 1]);
        Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
        std::vector<int> m;
        const int n = ref_track_gate_f2f(cur, last, m);
        for (int i = 0; i < n1; i++) { matches12[i] = m[i]; assign12[i] = -1; }
        for (int i2 = 0; i2 < n2; i2++) if (cur.mvpMapLines[i2]) assign12[cur.mvpMapLines[i2] - lines.data()] = i2;
        if (nnr) *nnr = (float)Config::minRatio12L();
        return n;
    });
}

// The matching half of Tracking::SearchLocalLines (src/Tracking.cc:3879-3919): match(local map lines, frame) and the
// position gate against the projections mTrackProjsX .. mTrackProjeY.  obs1[i1]: the local map line has observations;
// held2[i2]: mCurrentFrame.mvpMapLines[i2] is a line with Observations() > 0 (1) or without (2) before the loop.
REF_API int ref_track_lines_local(const uint8_t* d1, const float* proj1 /* n1 x 4: sX sY eX eY */, const uint8_t* obs1, int n1, const uint8_t* d2,
                                  const KeyLine* kl2, const float* disp2, const uint8_t* held2, int n2, float minX, float maxX,
                                  float minY, float maxY, int* matches12, int* assign12, float* nnr) {
    return guarded([&] {
        if (n2 < 2 || n1 < 2) throw std::runtime_error("undefined in the reference: fewer than 2 rows");
        std::vector<MapLine> lines(n1), holders(n2);
        std::vector<MapLine*> local(n1);
        for (int i = 0; i < n1; i++) {
            lines[i].mLDescriptor = wrap_desc(d1 + (size_t)i * 32, 1);
            lines[i].mTrackProjsX = proj1[4 * i]; lines[i].mTrackProjsY = proj1[4 * i + 1];
            lines[i].mTrackProjeX = proj1[4 * i + 2]; lines[i].mTrackProjeY = proj1[4 * i + 3];
            lines[i].nObs = obs1[i] ? 2 : 0;                  // pML->Observations()
            local[i] = &lines[i];
        }
        Frame cur;
        cur.mDescriptors_Line = wrap_desc(d2, n2);
        cur.mvKeysUn_Line.assign(kl2, kl2 + n2);
        cur.mvpMapLines.assign(n2, nullptr);
        cur.mvDisparity_l.resize(n2);
        for (int i = 0; i < n2; i++) {
            cur.mvDisparity_l[i] = std::make_pair(disp2[2 * i], disp2[2 * i + 1]);
            if (held2 && held2[i]) { holders[i].nObs = held2[i] == 1 ? 3 : 0; cur.mvpMapLines[i] = &holders[i]; }
        }
        Frame::mnMinX = minX; Frame::mnMaxX = maxX; Frame::mnMinY = minY; Frame::mnMaxY = maxY;
        std::vector<int> m;
        ref_track_gate_local(cur, local, n1, m);
        int cnt = 0;
        for (int i = 0; i < n1; i++) { matches12[i] = m[i]; assign12[i] = -1; }
        for (int i2 = 0; i2 < n2; i2++) {
            MapLine* q = cur.mvpMapLines[i2];
            if (q && q >= lines.data() && q < lines.data() + n1) { assign12[q - lines.data()] = i2; ++cnt; }
        }
        if (nnr) *nnr = (float)Config::minRatio12L();
        return cnt;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// Frame::ComputeBoW (src/Frame.cc:858-870): mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) by the reference's
// own DBoW2 (TemplatedVocabulary<FORB::TDescriptor, FORB>, include/ORBVocabulary.h) on a vocabulary loaded with its own
// loadFromTextFile (the ORBvoc.txt format).  Outputs: BowVector (word id ascending, value) and FeatureVector as CSR.
REF_API int ref_bow_transform(const char* vocText, const uint8_t* desc, int n, int levelsup, int* bowWord, double* bowValue,
                              int* fvNode, int* fvStart, int* fvFeat, int* nNodes) {
    return guarded([&] {
        typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Vocabulary;
        Vocabulary voc;
        if (!voc.loadFromTextFile(vocText)) throw std::runtime_error("loadFromTextFile failed");
        std::vector<cv::Mat> vCurrentDesc;
        vCurrentDesc.reserve(n);
        for (int i = 0; i < n; i++) vCurrentDesc.push_back(wrap_desc(desc + (size_t)i * 32, 1));   // Converter::toDescriptorVector
        DBoW2::BowVector bv;
        ::DBoW2::FeatureVector fv;
        voc.transform(vCurrentDesc, bv, fv, levelsup);
        int nw = 0;
        for (auto& kv : bv) { bowWord[nw] = (int)kv.first; bowValue[nw] = kv.second; ++nw; }
        int nn = 0, pos = 0;
        for (auto& kv : fv) {
            fvNode[nn] = (int)kv.first; fvStart[nn] = pos;
            for (unsigned f : kv.second) fvFeat[pos++] = (int)f;
            ++nn;
        }
        fvStart[nn] = pos;
        *nNodes = nn;
        return nw;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// The four ORBmatcher::SearchByProjection overloads (src/ORBmatcher.cc:44-214, 473-586, 2179-2323, 2325-2447) on a Frame /
// KeyFrame whose grid was filled by the reference's own AssignFeaturesToGrid.  A query is a plf_frame_query-shaped record
// (u, v, ur, radius, min_level, max_level, skip, has_observations, angle, desc[32], pad) = 18 ints; the map point behind
// it is placed at (u z, v z, z) with z a power of two and the camera is the unit pinhole, poses are identities, so the
// reference's own projection code yields exactly (u, v).
namespace {
struct RefQuery { float u, v, ur, radius; int min_level, max_level, skip, has_observations; float angle; uint8_t desc[32]; int pad; };
static_assert(sizeof(RefQuery) == 72, "same layout as plf_frame_query");

struct RefScene {
    Frame F;
    GeometricCamera cam;
    std::vector<MapPoint> holders;      // what F.mvpMapPoints points at before the call (one per feature)
    void build(const cv::KeyPoint* kp, const uint8_t* desc, const float* uRight, int N, int W, int H, const float* scale, int nLevels,
               const uint8_t* occupied /* 0 none, 1 holder with observations, 2 holder without */) {
        Frame::mnMinX = 0; Frame::mnMaxX = (float)W; Frame::mnMinY = 0; Frame::mnMaxY = (float)H;
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);   // src/Frame.cc:183-184
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        Frame::fx = 1; Frame::fy = 1; Frame::cx = 0; Frame::cy = 0;
        F.N = N;
        F.Nleft = -1;
        F.mvKeys.assign(kp, kp + N);
        F.mvKeysUn = F.mvKeys;
        F.mvuRight.assign(uRight, uRight + N);
        F.mDescriptors = wrap_desc(desc, N);
        F.mvScaleFactors.assign(scale, scale + nLevels);
        F.mpCamera = &cam;
        F.mTcw = cv::Mat::zeros(4, 4, CV_32F);
        for (int i = 0; i < 4; i++) F.mTcw.at<float>(i, i) = 1.f;
        holders.assign(N, MapPoint());
        F.mvpMapPoints.assign(N, nullptr);
        for (int i = 0; i < N; i++) if (occupied && occupied[i]) { holders[i].nObs = occupied[i] == 1 ? 1 : 0; F.mvpMapPoints[i] = &holders[i]; }
        F.AssignFeaturesToGrid();
    }
};
MapPoint make_point(const RefQuery& q, float z) {
    MapPoint p;
    p.mWorldPos = cv::Mat(3, 1, CV_32F);
    p.mWorldPos.at<float>(0) = q.u * z; p.mWorldPos.at<float>(1) = q.v * z; p.mWorldPos.at<float>(2) = z;
    p.mNormalVector = p.mWorldPos.clone();      // PO.dot(Pn) = |PO|^2 >= 0.5 |PO|
    p.mDescriptor = cv::Mat(1, 32, CV_8UC1);
    memcpy(p.mDescriptor.data, q.desc, 32);
    p.nObs = q.has_observations ? 1 : 0;
    return p;
}
}  // namespace

// mGrid as CSR (cell = ix * 48 + iy) and one GetFeaturesInArea lookup, for the pin of plf_feature_grid / plf_features_in_area
REF_API int ref_feature_grid(const cv::KeyPoint* kp, int N, int W, int H, int* cellStart /* 64*48+1 */, int* cellIdx /* N */,
                             float x, float y, float r, int minLevel, int maxLevel, int* area, int areaCap) {
    return guarded([&] {
        RefScene sc;
        std::vector<uint8_t> d((size_t)std::max(N, 1) * 32, 0);
        std::vector<float> ur((size_t)std::max(N, 1), -1.f), scale(8, 1.f);
        sc.build(kp, d.data(), ur.data(), N, W, H, scale.data(), 8, nullptr);
        int n = 0;
        for (int ix = 0; ix < FRAME_GRID_COLS; ix++) for (int iy = 0; iy < FRAME_GRID_ROWS; iy++) {
            cellStart[ix * FRAME_GRID_ROWS + iy] = n;
            for (size_t v : sc.F.mGrid[ix][iy]) cellIdx[n++] = (int)v;
        }
        cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = n;
        const std::vector<size_t> a = sc.F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
        for (size_t i = 0; i < a.size() && (int)i < areaCap; i++) area[i] = (int)a[i];
        return (int)a.size();
    });
}

// SearchByProjection(Frame& F, const vector<MapPoint*>&, th, bFarPoints = false) — the local-map search (:44-214).
// q: level = max_level, view cosine in `angle`, projection (u, v, ur); skip = not in view.  occupied in/out as in the product.
REF_API int ref_sbp_local(const cv::KeyPoint* kp, const uint8_t* desc, const float* uRight, int N, int W, int H, const float* scale,
                          int nLevels, const RefQuery* q, int nq, float th, float nnratio, uint8_t* occupied, int* match) {
    return guarded([&] {
        RefScene sc;
        sc.build(kp, desc, uRight, N, W, H, scale, nLevels, occupied);
        std::vector<MapPoint> pts(nq);
        std::vector<MapPoint*> vp(nq);
        for (int i = 0; i < nq; i++) {
            pts[i] = make_point(q[i], 1.f);
            pts[i].mbTrackInView = !q[i].skip;
            pts[i].mnTrackScaleLevel = q[i].max_level;
            pts[i].mTrackViewCos = q[i].angle;
            pts[i].mTrackProjX = q[i].u; pts[i].mTrackProjY = q[i].v; pts[i].mTrackProjXR = q[i].ur;
            pts[i].nObs = 1;                                   // map points of the local map have observations
            vp[i] = &pts[i];
        }
        ORBmatcher matcher(nnratio, true);
        const int n = matcher.SearchByProjection(sc.F, vp, th);
        for (int i = 0; i < nq; i++) match[i] = -1;
        for (int f = 0; f < N; f++) {
            MapPoint* h = sc.F.mvpMapPoints[f];
            const bool mine = h && h >= pts.data() && h < pts.data() + nq;
            if (mine) match[h - pts.data()] = f;
            occupied[f] = h ? (h->Observations() > 0 ? 1 : 0) : 0;
        }
        return n;
    });
}

// SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono, match12) — TrackWithMotionModel (:2179-2323).
// The last frame carries one map point per query; radius = th * scale[octave] is formed by the reference from
// q.max_level (octave) — direction: 0 around, 1 forward (LastFrame.mTcw z-translation > mb), 2 backward.
REF_API int ref_sbp_frame(const cv::KeyPoint* kp, const uint8_t* desc, const float* uRight, int N, int W, int H, const float* scale,
                          int nLevels, const RefQuery* q, const int* octave, const float* zs, int nq, float th, int direction, float mbf,
                          int checkOri, uint8_t* occupied, int* featQuery, int* match12) {
    return guarded([&] {
        RefScene sc;
        sc.build(kp, desc, uRight, N, W, H, scale, nLevels, occupied);
        sc.F.mb = 0.5f; sc.F.mbf = mbf;
        Frame last;
        last.N = nq;
        last.mTcw = cv::Mat::zeros(4, 4, CV_32F);
        for (int i = 0; i < 4; i++) last.mTcw.at<float>(i, i) = 1.f;
        last.mTcw.at<float>(2, 3) = direction == 1 ? 1.f : (direction == 2 ? -1.f : 0.f);
        std::vector<MapPoint> pts(nq);
        last.mvpMapPoints.assign(nq, nullptr);
        last.mvbOutlier.assign(nq, false);
        last.mvKeys.resize(nq);
        last.mvKeysUn.resize(nq);
        for (int i = 0; i < nq; i++) {
            pts[i] = make_point(q[i], zs[i]);
            if (!q[i].skip) last.mvpMapPoints[i] = &pts[i];
            last.mvKeys[i].octave = octave[i];
            last.mvKeysUn[i].angle = q[i].angle;
        }
        std::map<int, int> m12;
        ORBmatcher matcher(0.9f, checkOri != 0);
        const int n = matcher.SearchByProjection(sc.F, last, th, false, m12);
        for (int f = 0; f < N; f++) {
            MapPoint* h = sc.F.mvpMapPoints[f];
            const bool mine = h && h >= pts.data() && h < pts.data() + nq;
            featQuery[f] = mine ? (int)(h - pts.data()) : -1;
            occupied[f] = h ? (h->Observations() > 0 ? 1 : 0) : 0;
            match12[f] = -1;
        }
        for (auto& kv : m12) match12[kv.first] = kv.second;
        return n;
    });
}

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) — relocalisation (:2325-2447).
// q.max_level - 1 = the predicted level (the window is [pred - 1, pred + 1]); skip 1 = no map point, 2 = already found.
REF_API int ref_sbp_reloc(const cv::KeyPoint* kp, const uint8_t* desc, const float* uRight, int N, int W, int H, const float* scale,
                          int nLevels, const RefQuery* q, const float* zs, int nq, float th, int orbDist, int checkOri, uint8_t* occupied,
                          int* featQuery) {
    return guarded([&] {
        RefScene sc;
        sc.build(kp, desc, uRight, N, W, H, scale, nLevels, occupied);
        KeyFrame kf;
        std::vector<MapPoint> pts(nq);
        kf.mvpMapPoints.assign(nq, nullptr);
        kf.mvKeysUn.resize(nq);
        std::set<MapPoint*> found;
        for (int i = 0; i < nq; i++) {
            pts[i] = make_point(q[i], zs[i]);
            pts[i].mnPredictedLevel = q[i].max_level - 1;
            if (q[i].skip != 1) kf.mvpMapPoints[i] = &pts[i];
            if (q[i].skip == 2) found.insert(&pts[i]);
            if (q[i].skip == 3) pts[i].mbBad = true;
            kf.mvKeysUn[i].angle = q[i].angle;
        }
        ORBmatcher matcher(0.9f, checkOri != 0);
        const int n = matcher.SearchByProjection(sc.F, &kf, found, th, orbDist);
        for (int f = 0; f < N; f++) {
            MapPoint* h = sc.F.mvpMapPoints[f];
            const bool mine = h && h >= pts.data() && h < pts.data() + nq;
            featQuery[f] = mine ? (int)(h - pts.data()) : -1;
            occupied[f] = h ? 1 : 0;
        }
        return n;
    });
}

// SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th, ratioHamming) — loop closing (:473-586); the
// frame's features play the keyframe.  q.max_level = the predicted level; skip 1 = bad point, 2 = already in vpMatched.
REF_API int ref_sbp_loop(const cv::KeyPoint* kp, const uint8_t* desc, int N, int W, int H, const float* scale, int nLevels,
                         const RefQuery* q, const float* zs, int nq, int th, float ratioHamming, uint8_t* occupied, int* featQuery) {
    return guarded([&] {
        RefScene sc;
        std::vector<float> ur((size_t)std::max(N, 1), -1.f);
        sc.build(kp, desc, ur.data(), N, W, H, scale, nLevels, nullptr);
        KeyFrame kf;
        GeometricCamera cam;
        kf.mpCamera = &cam;
        kf.N = N;
        kf.mvKeysUn = sc.F.mvKeysUn;
        kf.mDescriptors = sc.F.mDescriptors;
        kf.mvScaleFactors = sc.F.mvScaleFactors;
        kf.mnMinX = 0; kf.mnMinY = 0; kf.mnMaxX = W; kf.mnMaxY = H;
        kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
        kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t> >(FRAME_GRID_ROWS));       // KeyFrame::KeyFrame copies F.mGrid
        for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) kf.mGrid[i][j] = sc.F.mGrid[i][j];
        std::vector<MapPoint> pts(nq), old(N);
        std::vector<MapPoint*> vp(nq), matched(N, nullptr);
        for (int f = 0; f < N; f++) if (occupied[f]) matched[f] = &old[f];
        for (int i = 0; i < nq; i++) {
            pts[i] = make_point(q[i], zs[i]);
            pts[i].mnPredictedLevel = q[i].max_level;
            pts[i].mbBad = q[i].skip == 1;
            vp[i] = &pts[i];
        }
        // "already found": such a point sits in vpMatched before the call
        int slot = 0;
        for (int i = 0; i < nq; i++) if (q[i].skip == 2) { while (slot < N && !occupied[slot]) ++slot; if (slot < N) matched[slot++] = &pts[i]; }
        cv::Mat Scw = cv::Mat::zeros(4, 4, CV_32F);
        for (int i = 0; i < 4; i++) Scw.at<float>(i, i) = 1.f;
        ORBmatcher matcher(0.75f, true);
        const int n = matcher.SearchByProjection(&kf, Scw, vp, matched, th, ratioHamming);
        for (int f = 0; f < N; f++) {
            MapPoint* h = matched[f];
            const bool mine = h && h >= pts.data() && h < pts.data() + nq && q[h - pts.data()].skip != 2;
            featQuery[f] = mine ? (int)(h - pts.data()) : -1;
            occupied[f] = h ? 1 : 0;
        }
        return n;
    });
}

// ---------------------------------------------------------------------------------------------------------------
// The stereo Frame constructor of the reference as a timed unit (bench.py --impl reference / cpu_baseline): four
// threads per pair — ExtractORB(left), ExtractORB(right), ExtractLine(left), ExtractLine(right), src/Frame.cc:128-135 —
// join, ComputeStereoMatches_Lines, ComputeStereoMatches (:160-163).  One handle per worker; handles share nothing but
// the read-only Config singleton, so several pairs can be in flight on a many-core host.
namespace {
struct RefPipeline {
    OrbPeek* orb[2];
    Lineextractor* line[2];
    Frame frame;
    float bf, fx;
};
}  // namespace

REF_API void* ref_frame_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, int lsd_nfeatures,
                               double min_line_length, int refine, double scale, double sigma_scale, double quant, double ang_th,
                               double log_eps, double density_th, int n_bins, float bf, float fx) {
    RefPipeline* p = new RefPipeline();
    for (int s = 0; s < 2; ++s) {
        p->orb[s] = new OrbPeek(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        p->line[s] = new Lineextractor(lsd_nfeatures, min_line_length, refine, scale, sigma_scale, quant, ang_th, log_eps, density_th, n_bins);
    }
    p->bf = bf; p->fx = fx;
    return p;
}
REF_API void ref_frame_destroy(void* h) {
    RefPipeline* p = (RefPipeline*)h;
    if (!p) return;
    for (int s = 0; s < 2; ++s) { delete p->orb[s]; delete p->line[s]; }
    delete p;
}
// counts[0..5] = N, Nr, N_l, Nr_l, stereo points, stereo lines.  mode bit 0: the reference's four extraction threads;
// bit 1: no ORB extractors / point matcher (line-only isolation, BASELINE config 4: matchNNR on the line descriptors instead);
// bit 2: no line extractors / line matcher (ORB-only isolation, config 3).
REF_API int ref_frame_run(void* h, const uint8_t* left, const uint8_t* right, int w, int hgt, int stride, int mode, int* counts) {
    return guarded([&] {
        const bool use_threads = mode & 1, noPoints = mode & 2, noLines = mode & 4;
        RefPipeline* p = (RefPipeline*)h;
        Frame& F = p->frame;
        cv::Mat im[2] = {wrap_u8(left, w, hgt, stride), wrap_u8(right, w, hgt, stride)};
        std::vector<int> lap = {0, 0};
        cv::Mat none;
        auto orbL = [&] { if (noPoints) { F.mvKeys.clear(); return; } (*p->orb[0])(im[0], none, F.mvKeys, F.mDescriptors, lap); };
        auto orbR = [&] { if (noPoints) { F.mvKeysRight.clear(); return; } (*p->orb[1])(im[1], none, F.mvKeysRight, F.mDescriptorsRight, lap); };
        auto lineL = [&] { if (noLines) { F.mvKeys_Line.clear(); return; } (*p->line[0])(im[0], none, F.mvKeys_Line, F.mDescriptors_Line); };
        auto lineR = [&] { if (noLines) { F.mvKeysRight_Line.clear(); return; } (*p->line[1])(im[1], none, F.mvKeysRight_Line, F.mDescriptorsRight_Line); };
        if (use_threads) {
            std::thread t0(orbL), t1(orbR), t2(lineL), t3(lineR);
            t0.join(); t1.join(); t2.join(); t3.join();
        } else { orbL(); orbR(); lineL(); lineR(); }
        F.N = (int)F.mvKeys.size(); F.N_l = (int)F.mvKeys_Line.size();
        F.mpORBextractorLeft = p->orb[0]; F.mpORBextractorRight = p->orb[1];
        F.mvScaleFactors = p->orb[0]->GetScaleFactors(); F.mvInvScaleFactors = p->orb[0]->GetInverseScaleFactors();
        F.mbf = p->bf; F.mb = p->bf / p->fx;
        F.inv_width = FRAME_GRID_COLS / static_cast<double>(w);
        F.inv_height = FRAME_GRID_ROWS / static_cast<double>(hgt);
        int sp = 0, sl = 0;
        if (noPoints || noLines) {
            if (!noLines && !F.mvKeys_Line.empty()) {
                F.ComputeStereoMatches_Lines();
                for (auto& d : F.mvDisparity_l) sl += d.first >= 0;
                if (F.mDescriptorsRight_Line.rows >= 2) {            // config 4 names matchNNR on the stereo descriptor sets
                    std::vector<int> m12;
                    matchNNR(F.mDescriptors_Line, F.mDescriptorsRight_Line, (float)Config::minRatio12L(), m12);
                }
            }
            if (!noPoints && !F.mvKeys.empty()) { F.ComputeStereoMatches(); for (float u : F.mvuRight) sp += u >= 0; }
        } else if (!F.mvKeys.empty() && !F.mvKeys_Line.empty()) {  // src/Frame.cc:146-149
            F.ComputeStereoMatches_Lines();
            F.ComputeStereoMatches();
            for (float u : F.mvuRight) sp += u >= 0;
            for (auto& d : F.mvDisparity_l) sl += d.first >= 0;
        }
        counts[0] = F.N; counts[1] = (int)F.mvKeysRight.size(); counts[2] = F.N_l; counts[3] = (int)F.mvKeysRight_Line.size();
        counts[4] = sp; counts[5] = sl;
        return 0;
    });
}
