// TEST INFRASTRUCTURE ONLY — CPU oracle, line part: LSD (OpenCV imgproc/lsd.cpp as shipped in cv2 4.13, restated from the
// published algorithm and validated black-box), the LSDDetectorC wrapper
// (Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:227-324), Lineextractor::operator()
// (src/LineExtractor.cc:31-70) and LBD (Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp).
#pragma once
#include "prims.h"
#include "../../include/plf_b200.h"
#include <vector>

namespace plfo {

struct LsdConfig {
    int refine = 0;
    double scale = 1.2, sigma_scale = 0.6, quant = 2.0, ang_th = 22.5, log_eps = 1.0, density_th = 0.6;
    int n_bins = 1024;
    // Seed order inside one gradient bin: raster order (== stable sort by bin).  Measured against cv2 4.13 this
    // reproduces its segments exactly (tests/test_oracle_cv2.py::test_lsd_vs_cv2), whereas an unstable std::sort over
    // (x,y,bin) records does not; `false` keeps the std::sort variant only for that comparison.
    bool stable_order = true;
};

struct LsdState {
    Img8 scaled;                       // U = resize(blur(img))
    std::vector<float> angleDeg;       // fastAtan2 output per pixel of U, -1024 = NOTDEF
    std::vector<float> segs;           // x1,y1,x2,y2 per segment, input-image coordinates, detection order
    std::vector<plf_keyline> kls;      // after Lineextractor filtering
    std::vector<uint8_t> desc;         // kls.size() x 32
    std::vector<float> lbd;            // kls.size() x 72
    bool valid = false;
};

// OpenCV getGaussianKernel bit-exact 8.8 fixed-point taps (error-diffusion rounding, sum == 256).
void gaussian_taps_fixed(int ksize, double sigma, std::vector<int>& taps);
void lsd_detect(const LsdConfig& c, const Img8& img, LsdState& st);
// LSDDetectorC::detectImpl post-processing + Lineextractor filter; min_length in pixels.
void lines_to_keylines(const std::vector<float>& segs, int w, int h, double min_length, int nfeatures,
                       std::vector<plf_keyline>& kls);
void lbd_compute(const Img8& img, const std::vector<plf_keyline>& kls, std::vector<float>& lbd72,
                 std::vector<uint8_t>& desc);
bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2);
int line_iterator_count(float x1, float y1, float x2, float y2, int W, int H);

}  // namespace plfo
