// TEST INFRASTRUCTURE ONLY — CPU oracle of the PLI-SLAM stereo point-line frontend (see oracle/README.md).
// Pinned restatements of the OpenCV primitives the reference calls (OpenCV itself is not under /root/reference;
// formulas are the ones SURVEY.md §8c pinned black-box against cv2 4.13.0 and re-checked by tests/test_oracle_cv2.py).
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <cfloat>

namespace plfo {

struct Img8 {
    int w = 0, h = 0;
    std::vector<uint8_t> d;
    Img8() {}
    Img8(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
    uint8_t* row(int y) { return d.data() + (size_t)y * w; }
    const uint8_t* row(int y) const { return d.data() + (size_t)y * w; }
    uint8_t at(int y, int x) const { return d[(size_t)y * w + x]; }
};

struct Img16 {
    int w = 0, h = 0;
    std::vector<int16_t> d;
    Img16() {}
    Img16(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
    int16_t at(int y, int x) const { return d[(size_t)y * w + x]; }
};

// cvRound: round half to even (SSE cvtsd2si semantics), SURVEY Appendix A6.
static inline int cv_round(double v) { return (int)std::nearbyint(v); }
static inline int cv_roundf(float v) { return (int)std::nearbyintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

// BORDER_REFLECT_101 index.
static inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

// cv::fastAtan2 (degrees in [0,360)), call sites src/ORBextractor.cc:101 and OpenCV LSD; SURVEY §8c fact 3.
float fast_atan2(float y, float x);

// cv::resize(8UC1, INTER_LINEAR), call site src/ORBextractor.cc:1165; SURVEY §8c fact 1.
void resize_linear_u8(const Img8& src, Img8& dst, int dw, int dh);
// cv::resize(8UC1, fx=fy=scale, INTER_LINEAR_EXACT) as used inside OpenCV LSD; SURVEY §8c fact 4.
void resize_linear_exact_u8(const Img8& src, Img8& dst, double scale);
// cv::GaussianBlur(8U, REFLECT_101) bit-exact fixed-point path with integer taps summing to 256; fact 2.
// taps has `ksize` entries.
void gaussian_blur_u8(const Img8& src, Img8& dst, const int* taps, int ksize);
// cv::Sobel(8U -> 16S, ksize 3, REFLECT_101), call site binary_descriptor_custom.cpp:395-396.
void sobel3_16s(const Img8& src, Img16& dx, Img16& dy);
// cv::remap(src, dst, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, 0) for 8UC1 with CV_32FC1 maps (the call of
// Examples/Stereo/stereo_euroc.cc:166-167), restated from OpenCV's fixed-point path; dst has the size of the maps.
void remap_linear_u8(const Img8& src, Img8& dst, const float* mapx, const float* mapy, int dw, int dh);

// The reference calls cos / sin / atan2 on FLOAT operands at three places of its own code (src/ORBextractor.cc:112,
// LSDDetector_custom.cpp:297, binary_descriptor_custom.cpp:1134-1135); with GCC >= 6 these bind to the float overloads,
// i.e. to the cosf / sinf / atan2f of whatever libm the binary meets at run time — not correctly rounded before glibc
// 2.41, and CPU-dependent through ifunc variants.  The oracle DECLARES the correctly rounded value (double libm rounded
// to float; what glibc >= 2.41 returns) — mode 0, the only mode the product is compared with.  Mode 1 calls this
// machine's float libm instead: "the reference as built here", which tests/test_oracle_ref.py uses to show that the
// libm rounding is the ONLY difference between the oracle and the reference's own object code.
extern int g_float_libm;
float ref_cosf(float x);
float ref_sinf(float x);
float ref_atan2f(float y, float x);

extern const int TAPS_ORB7[7];   // 7x7 sigma 2   : 18 34 48 56 48 34 18
extern const int TAPS_LBD5[5];   // 5x5 sigma 1   : 14 62 104 62 14
extern const int TAPS_LSD7[7];   // 7x7 sigma 0.6 : 0 1 42 170 42 1 0

}  // namespace plfo
