// Drives the C++ host mirror (pli-slam_b200/host/plf_frontend.hpp) the way Frame::Frame(stereo) drives the reference
// classes (src/Frame.cc:128-163) and prints a few counters; tests/test_host_shim.py compares them with the oracle.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include "../pli-slam_b200/host/plf_frontend.hpp"

static std::vector<uint8_t> read_file(const char* p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s left.raw right.raw W H\n", argv[0]); return 2; }
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
        printf("{\"N\": %zu, \"Nr\": %zu, \"mono\": [%d, %d], \"Nl\": %zu, \"Nlr\": %zu, \"stereo_pts\": %d, \"sum_u\": %.4f, "
               "\"stereo_lines\": %d, \"nnr\": %d, \"desc_fnv\": %llu, \"hamming01\": %d, \"empty\": %d}\n",
               mvKeys.size(), mvKeysRight.size(), monoLeft, monoRight, mvKeys_Line.size(), mvKeysRight_Line.size(), stereoPts,
               sumU, stereoLines, nnr, h, ORB_SLAM3::ORBmatcher::DescriptorDistance(mDescriptors.row(0), mDescriptors.row(1)),
               orbL(none, none, mvKeys, mDescriptors, lap));
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
