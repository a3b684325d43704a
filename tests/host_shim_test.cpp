// Drives the C++ host mirror (pli-slam_b200/host/plf_frontend.hpp) the way Frame::Frame(stereo) drives the reference
// classes (src/Frame.cc:128-163) and prints a few counters; tests/test_host_shim.py compares them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include "../pli-slam_b200/host/plf_frontend.hpp"

static std::vector<uint8_t> read_file(const char* p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s left.raw right.raw W H [mapx.f32 mapy.f32]\n", argv[0]); return 2; }
    const int W = atoi(argv[3]), H = atoi(argv[4]);
    std::vector<uint8_t> L = read_file(argv[1]), R = read_file(argv[2]);
    if ((int)L.size() != W * H || (int)R.size() != W * H) { fprintf(stderr, "bad raw size\n"); return 2; }
    plf_params p;
    plf_default_params(&p);
    p.width = W; p.height = H;
    try {
        auto ctx = std::make_shared<plf::Context>(p, 0);
        ORB_SLAM3::ORBextractor orbL(ctx, 0), orbR(ctx, 1);
        ORB_SLAM3::Lineextractor lineL(ctx, 0), lineR(ctx, 1);
        std::vector<ORB_SLAM3::KeyPoint> mvKeys, mvKeysRight;
        std::vector<ORB_SLAM3::KeyLine> mvKeys_Line, mvKeysRight_Line;
        plf::Desc mDescriptors, mDescriptorsRight, mDescriptors_Line, mDescriptorsRight_Line;
        std::vector<int> lap = {0, 0};
        plf::Mat8 imL(L.data(), H, W), imR(R.data(), H, W), none;
        int monoLeft = orbL(imL, none, mvKeys, mDescriptors, lap);
        int monoRight = orbR(imR, none, mvKeysRight, mDescriptorsRight, lap);
        lineL(imL, none, mvKeys_Line, mDescriptors_Line);
        lineR(imR, none, mvKeysRight_Line, mDescriptorsRight_Line);
        ORB_SLAM3::StereoFrontend sf(ctx);
        std::vector<float> mvuRight, mvDepth;
        std::vector<std::pair<float, float>> mvDisparity_l;
        std::vector<std::array<double, 3>> mvle_l;
        sf.ComputeStereoMatches_Lines((int)mvKeys_Line.size(), mvDisparity_l, mvle_l);
        sf.ComputeStereoMatches((int)mvKeys.size(), mvuRight, mvDepth);
        std::vector<int> m12;
        int nnr = ORB_SLAM3::matchNNR(*ctx, mDescriptors_Line.view(), mDescriptorsRight_Line.view(), 0.9f, m12);
        int stereoPts = 0, stereoLines = 0;
        double sumU = 0;
        for (size_t i = 0; i < mvuRight.size(); ++i) if (mvuRight[i] >= 0) { ++stereoPts; sumU += mvuRight[i]; }
        for (auto& d : mvDisparity_l) stereoLines += d.first >= 0;
        unsigned long long h = 1469598103934665603ull;
        for (uint8_t b : mDescriptors.bytes) h = (h ^ b) * 1099511628211ull;
        // --- the frame-level tail: grid, area lookup, back-projection (FrameTail), on the state of the frame above
        ORB_SLAM3::FrameTail tail(ctx);
        tail.AssignFeaturesToGrid();
        std::vector<size_t> area = tail.GetFeaturesInArea(mvKeys, 376.f, 240.f, 60.f, 0, 3);
        unsigned long long hArea = 1469598103934665603ull;
        for (size_t i : area) hArea = (hArea ^ (unsigned long long)i) * 1099511628211ull;
        const float Rwc[9] = {0.96f, -0.28f, 0.f, 0.28f, 0.96f, 0.f, 0.f, 0.f, 1.f}, Ow[3] = {0.5f, -1.25f, 2.f};
        std::vector<float> x3D;
        std::vector<double> lines3D;
        tail.BackProject(Rwc, Ow, 435.2047f, 367.4517f, 252.2008f, (int)mvKeys.size(), (int)mvKeys_Line.size(), x3D, lines3D);
        double sumX = 0, sumL = 0;
        for (float v : x3D) sumX += v;
        for (double v : lines3D) sumL += v;
        // --- the gated line matching of the tracking thread (FrameTail::MatchLinesTracked): the frame against a copy of itself
        // shifted by 4 px, every line eligible -> a line is attached iff it has stereo disparities and is its own mutual best
        std::vector<plf_track_line> last(mvKeys_Line.size());
        for (size_t i = 0; i < last.size(); ++i)
            last[i] = {mvKeys_Line[i].startPointX + 4.f, mvKeys_Line[i].startPointY, mvKeys_Line[i].endPointX + 4.f,
                       mvKeys_Line[i].endPointY, mvKeys_Line[i].angle, 1};
        std::vector<int> m12t, a12t;
        const int tracked = tail.MatchLinesTracked(0, mDescriptors_Line.view(), last, mDescriptors_Line.view(), mvKeys_Line, mvDisparity_l,
                                                   nullptr, 0.9f, 0.f, (float)W, 0.f, (float)H, m12t, a12t);
        // --- SearchByBoW of the frame against itself with every feature in one of 8 nodes: each feature finds itself
        std::vector<float> ang(mvKeys.size());
        std::vector<int32_t> nodes(mvKeys.size());
        std::vector<uint8_t> good(mvKeys.size(), 1);
        for (size_t i = 0; i < mvKeys.size(); ++i) { ang[i] = mvKeys[i].angle; nodes[i] = (int32_t)(i % 8); }
        std::vector<int32_t> bowMatches;
        const int nBow = tail.SearchByBoW(mDescriptors.view(), ang, nodes, good, nodes, 0.7f, true, bowMatches);
        int bowSelf = 0;
        for (size_t i = 0; i < bowMatches.size(); ++i) bowSelf += bowMatches[i] == (int32_t)i;
        // --- lapping area {0, 1000} (the monocular constructor, src/Frame.cc:360-361) and a padded-row view of the same image
        std::vector<ORB_SLAM3::KeyPoint> kLap, kPad;
        plf::Desc dLap, dPad;
        std::vector<int> lapMono = {0, 1000};
        const int monoLap = orbL(imL, none, kLap, dLap, lapMono);
        const int S = W + 40;
        std::vector<uint8_t> padded((size_t)S * H, 0x5A);
        for (int y = 0; y < H; ++y) std::copy(L.begin() + (size_t)y * W, L.begin() + (size_t)(y + 1) * W, padded.begin() + (size_t)y * S);
        orbL(plf::Mat8(padded.data(), H, W, S), none, kPad, dPad, lap);
        const int padSame = kPad.size() == mvKeys.size() && dPad.bytes == mDescriptors.bytes;
        const int lapReversed = kLap.size() == mvKeys.size() && !kLap.empty() && kLap.back().x == mvKeys.front().x &&
                                kLap.back().y == mvKeys.front().y && kLap.front().x == mvKeys.back().x;
        // --- rectification in front of the path (Rectifier), when maps are given
        unsigned long long hRect = 0;
        if (argc >= 7) {
            std::vector<uint8_t> mx = read_file(argv[5]), my = read_file(argv[6]);
            if ((int)mx.size() != W * H * 4 || (int)my.size() != W * H * 4) { fprintf(stderr, "bad map size\n"); return 2; }
            ORB_SLAM3::Rectifier rect(ctx);
            rect.setMaps(0, reinterpret_cast<const float*>(mx.data()), reinterpret_cast<const float*>(my.data()), W, H);
            std::vector<uint8_t> imRect;
            rect.remap(0, imL, imRect);
            hRect = 1469598103934665603ull;
            for (uint8_t b : imRect) hRect = (hRect ^ b) * 1099511628211ull;
        }
        printf("{\"N\": %zu, \"Nr\": %zu, \"mono\": [%d, %d], \"Nl\": %zu, \"Nlr\": %zu, \"stereo_pts\": %d, \"sum_u\": %.4f, "
               "\"stereo_lines\": %d, \"nnr\": %d, \"desc_fnv\": %llu, \"hamming01\": %d, \"area_n\": %zu, \"area_fnv\": %llu, "
               "\"sum_x3d\": %.6f, \"sum_l3d\": %.9f, \"mono_lap\": %d, \"lap_reversed\": %d, \"pad_same\": %d, \"rect_fnv\": %llu, "
               "\"tracked\": %d, \"bow\": [%d, %d], \"empty\": %d}\n",
               mvKeys.size(), mvKeysRight.size(), monoLeft, monoRight, mvKeys_Line.size(), mvKeysRight_Line.size(), stereoPts,
               sumU, stereoLines, nnr, h, ORB_SLAM3::ORBmatcher::DescriptorDistance(mDescriptors.row(0), mDescriptors.row(1)),
               area.size(), hArea, sumX, sumL, monoLap, lapReversed, padSame, hRect, tracked, nBow, bowSelf,
               orbL(none, none, mvKeys, mDescriptors, lap));
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
