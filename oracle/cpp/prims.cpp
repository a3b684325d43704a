// TEST INFRASTRUCTURE ONLY — see prims.h.
#include "prims.h"
#include <algorithm>
#include <cstring>

namespace plfo {

int g_float_libm = 0;
float ref_cosf(float x) { return g_float_libm ? ::cosf(x) : (float)std::cos((double)x); }
float ref_sinf(float x) { return g_float_libm ? ::sinf(x) : (float)std::sin((double)x); }
float ref_atan2f(float y, float x) { return g_float_libm ? ::atan2f(y, x) : (float)std::atan2((double)y, (double)x); }


const int TAPS_ORB7[7] = {18, 34, 48, 56, 48, 34, 18};
const int TAPS_LBD5[5] = {14, 62, 104, 62, 14};
const int TAPS_LSD7[7] = {0, 1, 42, 170, 42, 1, 0};

// Degree-7 odd polynomial in float, no FMA (this TU is built with -ffp-contract=off).
float fast_atan2(float y, float x) {
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k;
    const float p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k;
    const float p7 = -0.04432655554792128f * k;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

void resize_linear_u8(const Img8& src, Img8& dst, int dw, int dh) {
    dst = Img8(dw, dh);
    const int sw = src.w, sh = src.h;
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<int> ax0(dw), ax1(dw), ay0(dh), ay1(dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ax0[dx] = cv_roundf((1.f - fx) * 2048.f);
        ax1[dx] = cv_roundf(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        if (sy < 0) { fy = 0; sy = 0; }
        if (sy >= sh - 1) { fy = 0; sy = sh - 1; }
        yofs[dy] = sy;
        ay0[dy] = cv_roundf((1.f - fy) * 2048.f);
        ay1[dy] = cv_roundf(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        const uint8_t* s0 = src.row(yofs[dy]);
        const uint8_t* s1 = src.row(std::min(yofs[dy] + 1, sh - 1));
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
            r0[dx] = s0[sx] * ax0[dx] + s0[sx1] * ax1[dx];
            r1[dx] = s1[sx] * ax0[dx] + s1[sx1] * ax1[dx];
        }
        uint8_t* d = dst.row(dy);
        const int b0 = ay0[dy], b1 = ay1[dy];
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

void resize_linear_exact_u8(const Img8& src, Img8& dst, double scale) {
    const int sw = src.w, sh = src.h;
    const int dw = cv_round(sw * scale), dh = cv_round(sh * scale);
    dst = Img8(dw, dh);
    const double inv = 1.0 / scale;
    auto coeffs = [&](int n_dst, int n_src, std::vector<int>& ofs, std::vector<int>& a1) {
        ofs.resize(n_dst);
        a1.resize(n_dst);
        for (int d = 0; d < n_dst; ++d) {
            double f = inv * (d + 0.5) - 0.5;
            int i = cv_floor(f);
            if (i >= 0 && n_src > 1) {
                if (i < n_src - 1) {
                    ofs[d] = i;
                    a1[d] = cv_round((f - i) * 256.0);
                } else {
                    ofs[d] = n_src - 1;
                    a1[d] = 0;
                }
            } else {
                ofs[d] = 0;
                a1[d] = 0;
            }
        }
    };
    std::vector<int> xo, xa, yo, ya;
    coeffs(dw, sw, xo, xa);
    coeffs(dh, sh, yo, ya);
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        const uint8_t* s0 = src.row(yo[dy]);
        const uint8_t* s1 = src.row(std::min(yo[dy] + 1, sh - 1));
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xo[dx], sx1 = std::min(sx + 1, sw - 1);
            int a = xa[dx];
            r0[dx] = s0[sx] * (256 - a) + s0[sx1] * a;
            r1[dx] = s1[sx] * (256 - a) + s1[sx1] * a;
        }
        uint8_t* d = dst.row(dy);
        const int b = ya[dy];
        for (int dx = 0; dx < dw; ++dx) d[dx] = (uint8_t)((r0[dx] * (256 - b) + r1[dx] * b + 32768) >> 16);
    }
}

void gaussian_blur_u8(const Img8& src, Img8& dst, const int* taps, int ksize) {
    // same arithmetic as before (horizontal sums in 16 bits, vertical in 32, (v + 32768) >> 16), organised for the
    // compiler's vectoriser: a reflect-padded source row, tap-major inner loops over contiguous pixels
    const int w = src.w, h = src.h, r = ksize / 2;
    std::vector<uint16_t> tmp((size_t)w * h);
    std::vector<uint8_t> pad((size_t)w + 2 * r);
    std::vector<uint32_t> acc((size_t)w);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src.row(y);
        for (int x = -r; x < w + r; ++x) pad[x + r] = s[reflect101(x, w)];
        uint16_t* t = tmp.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) acc[x] = 0;
        for (int k = 0; k < ksize; ++k) {
            const uint32_t tk = (uint32_t)taps[k];
            const uint8_t* pk = pad.data() + k;
            for (int x = 0; x < w; ++x) acc[x] += tk * pk[x];
        }
        for (int x = 0; x < w; ++x) t[x] = (uint16_t)acc[x];
    }
    dst = Img8(w, h);
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst.row(y);
        for (int x = 0; x < w; ++x) acc[x] = 0;
        for (int k = 0; k < ksize; ++k) {
            const uint32_t tk = (uint32_t)taps[k];
            const uint16_t* tr = tmp.data() + (size_t)reflect101(y + k - r, h) * w;
            for (int x = 0; x < w; ++x) acc[x] += tk * tr[x];
        }
        for (int x = 0; x < w; ++x) d[x] = (uint8_t)((acc[x] + 32768u) >> 16);
    }
}

void sobel3_16s(const Img8& src, Img16& dx, Img16& dy) {
    const int w = src.w, h = src.h;
    dx = Img16(w, h);
    dy = Img16(w, h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src.row(reflect101(y - 1, h));
        const uint8_t* r1 = src.row(y);
        const uint8_t* r2 = src.row(reflect101(y + 1, h));
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int gx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
            int gy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
            dx.d[(size_t)y * w + x] = (int16_t)gx;
            dy.d[(size_t)y * w + x] = (int16_t)gy;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// cv::remap, 8UC1, INTER_LINEAR, BORDER_CONSTANT(0), planar CV_32F maps (OpenCV imgproc remap: RemapInvoker +
// remapBilinear<FixedPtCast<int, uchar, 15>>; call site Examples/Stereo/stereo_euroc.cc:166-167).
//  * sx = cvRound(mapx * 32), sy = cvRound(mapy * 32) (round half to even); integer part = saturate_cast<short>(s >> 5),
//    fraction index (sy & 31) * 32 + (sx & 31);
//  * weights = saturate_cast<short>(wy * wx * 32768) with wx in {1 - fx/32, fx/32}; all are exact integers
//    32 * (32 - fy or fy) * (32 - fx or fx); the only saturation is the (0, 0) entry, 32768 -> 32767, whose missing unit
//    OpenCV's table fix-up adds to the LAST tap, giving {32767, 0, 0, 1};
//  * out = (sum of tap * weight + 2^14) >> 15; a tap outside the source counts as 0.
void remap_linear_u8(const Img8& src, Img8& dst, const float* mapx, const float* mapy, int dw, int dh) {
    dst = Img8(dw, dh);
    auto sat16 = [](int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); };
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            const size_t i = (size_t)y * dw + x;
            const int fxs = cv_roundf(mapx[i] * 32.f), fys = cv_roundf(mapy[i] * 32.f);
            const int sx = sat16(fxs >> 5), sy = sat16(fys >> 5);
            const int fx = fxs & 31, fy = fys & 31;
            int w[4] = {32 * (32 - fy) * (32 - fx), 32 * (32 - fy) * fx, 32 * fy * (32 - fx), 32 * fy * fx};
            if (fx == 0 && fy == 0) { w[0] = 32767; w[3] = 1; }
            auto tap = [&](int yy, int xx) -> int {
                return (xx >= 0 && yy >= 0 && xx < src.w && yy < src.h) ? src.at(yy, xx) : 0;
            };
            const int sum = tap(sy, sx) * w[0] + tap(sy, sx + 1) * w[1] + tap(sy + 1, sx) * w[2] + tap(sy + 1, sx + 1) * w[3];
            const int v = (sum + (1 << 14)) >> 15;
            dst.row(y)[x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
}

}  // namespace plfo
