// TEST INFRASTRUCTURE ONLY — CPU oracle, LSD + KeyLine construction.  See lsd.h for provenance.
#include "lsd.h"
#include <array>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace plfo {

static const double kPi = 3.14159265358979323846;
static const double kNotDef = -1024.0;
static const double kDegToRad = kPi / 180.0;

void gaussian_taps_fixed(int ksize, double sigma, std::vector<int>& taps) {
    std::vector<double> k(ksize);
    double sum = 0;
    const int r = ksize / 2;
    for (int i = 0; i < ksize; ++i) {
        double x = i - r;
        k[i] = std::exp(-(x * x) / (2 * sigma * sigma));
        sum += k[i];
    }
    for (double& v : k) v /= sum;
    taps.assign(ksize, 0);
    double err = 0;
    int s = 0;
    for (int i = 0; i < r; ++i) {
        double adj = k[i] * 256.0 + err;
        int v0 = cv_round(adj);
        err = adj - v0;
        taps[i] = taps[ksize - 1 - i] = v0;
        s += v0;
    }
    taps[r] = 256 - 2 * s;
}

namespace {
struct RegPt { int x, y; };

struct Lsd {
    const LsdConfig& c;
    int W = 0, H = 0;
    std::vector<double> angles, modgrad;
    std::vector<uint8_t> used;
    struct NormPt { int x, y, norm; };
    std::vector<NormPt> ordered;
    explicit Lsd(const LsdConfig& cc) : c(cc) {}

    void ll_angle(const Img8& U, double threshold, std::vector<float>& angDeg) {
        W = U.w; H = U.h;
        angles.assign((size_t)W * H, kNotDef);
        modgrad.assign((size_t)W * H, 0.0);
        angDeg.assign((size_t)W * H, (float)kNotDef);
        double max_grad = -1;
        for (int y = 0; y < H - 1; ++y) {
            const uint8_t* r0 = U.row(y);
            const uint8_t* r1 = U.row(y + 1);
            for (int x = 0; x < W - 1; ++x) {
                int DA = r1[x + 1] - r0[x];
                int BC = r0[x + 1] - r1[x];
                int gx = DA + BC, gy = DA - BC;
                double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
                modgrad[(size_t)y * W + x] = norm;
                if (norm <= threshold) {
                    angles[(size_t)y * W + x] = kNotDef;
                } else {
                    float deg = fast_atan2((float)gx, (float)-gy);
                    angDeg[(size_t)y * W + x] = deg;
                    angles[(size_t)y * W + x] = deg * kDegToRad;
                    if (norm > max_grad) max_grad = norm;
                }
            }
        }
        const double bin_coef = (max_grad > 0) ? double(c.n_bins - 1) / max_grad : 0;
        ordered.clear();
        ordered.reserve((size_t)(W - 1) * (H - 1));
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x)
                ordered.push_back({x, y, (int)(modgrad[(size_t)y * W + x] * bin_coef)});
        auto cmp = [](const NormPt& a, const NormPt& b) { return a.norm > b.norm; };
        if (c.stable_order) std::stable_sort(ordered.begin(), ordered.end(), cmp);
        else std::sort(ordered.begin(), ordered.end(), cmp);
    }

    bool is_aligned(int x, int y, double theta, double prec) const {
        if (x < 0 || y < 0 || x >= W || y >= H) return false;
        const double a = angles[(size_t)y * W + x];
        if (a == kNotDef) return false;
        double n_theta = theta - a;
        if (n_theta < 0) n_theta = -n_theta;
        if (n_theta > (3 * kPi) / 2) {
            n_theta -= 2 * kPi;
            if (n_theta < 0) n_theta = -n_theta;
        }
        return n_theta <= prec;
    }

    void region_grow(int sx, int sy, std::vector<RegPt>& reg, double& reg_angle, double prec) {
        reg.clear();
        reg.push_back({sx, sy});
        reg_angle = angles[(size_t)sy * W + sx];
        float sumdx = (float)std::cos(reg_angle);
        float sumdy = (float)std::sin(reg_angle);
        used[(size_t)sy * W + sx] = 1;
        for (size_t i = 0; i < reg.size(); ++i) {
            const RegPt rp = reg[i];
            int xx_min = std::max(rp.x - 1, 0), xx_max = std::min(rp.x + 1, W - 1);
            int yy_min = std::max(rp.y - 1, 0), yy_max = std::min(rp.y + 1, H - 1);
            for (int yy = yy_min; yy <= yy_max; ++yy)
                for (int xx = xx_min; xx <= xx_max; ++xx) {
                    uint8_t& u = used[(size_t)yy * W + xx];
                    if (u != 1 && is_aligned(xx, yy, reg_angle, prec)) {
                        const double angle = angles[(size_t)yy * W + xx];
                        u = 1;
                        reg.push_back({xx, yy});
                        // cos(float)/sin(float) of OpenCV's region_grow, taken as correctly rounded (see orb.cpp)
                        sumdx += (float)std::cos((double)(float)angle);
                        sumdy += (float)std::sin((double)(float)angle);
                        reg_angle = fast_atan2(sumdy, sumdx) * kDegToRad;
                    }
                }
        }
    }

    static double angle_diff(double a, double b) {
        double diff = a - b;
        while (diff <= -kPi) diff += 2 * kPi;
        while (diff > kPi) diff -= 2 * kPi;
        if (diff < 0) diff = -diff;
        return diff;
    }

    double get_theta(const std::vector<RegPt>& reg, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0, Iyy = 0, Ixy = 0;
        for (const RegPt& p : reg) {
            const double w = modgrad[(size_t)p.y * W + p.x];
            double dx = (double)p.x - x, dy = (double)p.y - y;
            Ixx += dy * dy * w;
            Iyy += dx * dx * w;
            Ixy -= dx * dy * w;
        }
        double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? (double)fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                                         : (double)fast_atan2((float)Ixy, (float)(lambda - Iyy));
        theta *= kDegToRad;
        if (angle_diff(theta, reg_angle) > prec) theta += kPi;
        return theta;
    }

    struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

    // region2rect (endpoints before the +0.5 shift)
    void region2rect(const std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (const RegPt& q : reg) {
            const double w = modgrad[(size_t)q.y * W + q.x];
            x += (double)q.x * w;
            y += (double)q.y * w;
            sum += w;
        }
        x /= sum;
        y /= sum;
        double theta = get_theta(reg, x, y, reg_angle, prec);
        double dx = std::cos(theta), dy = std::sin(theta);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (const RegPt& q : reg) {
            double rdx = (double)q.x - x, rdy = (double)q.y - y;
            double l = rdx * dx + rdy * dy;
            double w = -rdx * dy + rdy * dx;
            if (l > l_max) l_max = l;
            else if (l < l_min) l_min = l;
            if (w > w_max) w_max = w;
            else if (w < w_min) w_min = w;
        }
        rec.x1 = x + l_min * dx;
        rec.y1 = y + l_min * dy;
        rec.x2 = x + l_max * dx;
        rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min;
        rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    static double dist(double x1, double y1, double x2, double y2) {
        return std::sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
    }
    static double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
    static double angle_diff_signed(double a, double b) {
        double diff = a - b;
        while (diff <= -kPi) diff += 2 * kPi;
        while (diff > kPi) diff -= 2 * kPi;
        return diff;
    }

    // LSD reduce_region_radius (refine >= 1)
    bool reduce_region_radius(std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec, double density,
                              double density_th) {
        double xc = (double)reg[0].x, yc = (double)reg[0].y;
        double radSq1 = dist_sq(xc, yc, rec.x1, rec.y1), radSq2 = dist_sq(xc, yc, rec.x2, rec.y2);
        double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            for (size_t i = 0; i < reg.size(); ++i) {
                if (dist_sq(xc, yc, (double)reg[i].x, (double)reg[i].y) > radSq) {
                    used[(size_t)reg[i].y * W + reg[i].x] = 0;
                    std::swap(reg[i], reg[reg.size() - 1]);
                    reg.pop_back();
                    --i;
                }
            }
            if (reg.size() < 2) return false;
            region2rect(reg, reg_angle, prec, p, rec);
            density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }

    // LSD refine (refine >= 1): density check, re-grow with a tolerance from the local angle spread, radius reduction
    bool refine(std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        double xc = (double)reg[0].x, yc = (double)reg[0].y;
        const double ang_c = angles[(size_t)reg[0].y * W + reg[0].x];
        double sum = 0, s_sum = 0;
        int n = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            used[(size_t)reg[i].y * W + reg[i].x] = 0;
            if (dist(xc, yc, (double)reg[i].x, (double)reg[i].y) < rec.width) {
                const double angle = angles[(size_t)reg[i].y * W + reg[i].x];
                double ang_d = angle_diff_signed(angle, ang_c);
                sum += ang_d;
                s_sum += ang_d * ang_d;
                ++n;
            }
        }
        double mean_angle = sum / (double)n;
        double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / (double)n + mean_angle * mean_angle);
        const int sx = reg[0].x, sy = reg[0].y;
        region_grow(sx, sy, reg, reg_angle, tau);
        if (reg.size() < 2) return false;
        region2rect(reg, reg_angle, prec, p, rec);
        density = (double)reg.size() / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
        return true;
    }
};
}  // namespace

void lsd_detect(const LsdConfig& c, const Img8& img, LsdState& st) {
    st.valid = false;
    st.segs.clear();
    const double prec = kPi * c.ang_th / 180;
    const double p = c.ang_th / 180;
    const double rho = c.quant / std::sin(prec);
    Lsd L(c);
    if (c.scale != 1) {
        const double sigma = (c.scale < 1) ? (c.sigma_scale / c.scale) : c.sigma_scale;
        const double sprec = 3;
        const unsigned h = (unsigned)std::ceil(sigma * std::sqrt(2 * sprec * std::log(10.0)));
        const int ksize = 1 + 2 * (int)h;
        std::vector<int> taps;
        gaussian_taps_fixed(ksize, sigma, taps);
        Img8 g;
        gaussian_blur_u8(img, g, taps.data(), ksize);
        resize_linear_exact_u8(g, st.scaled, c.scale);
    } else {
        st.scaled = img;
    }
    L.ll_angle(st.scaled, rho, st.angleDeg);
    const int W = L.W, H = L.H;
    const double logNT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const size_t min_reg_size = (size_t)(-logNT / std::log10(p));
    L.used.assign((size_t)W * H, 0);
    std::vector<RegPt> reg;
    for (const auto& op : L.ordered) {
        const size_t idx = (size_t)op.y * W + op.x;
        if (L.used[idx] != 0 || L.angles[idx] == kNotDef) continue;
        double reg_angle;
        L.region_grow(op.x, op.y, reg, reg_angle, prec);
        if (reg.size() < min_reg_size) continue;
        Lsd::Rect rec;
        L.region2rect(reg, reg_angle, prec, p, rec);
        if (c.refine >= 1 && !L.refine(reg, reg_angle, prec, p, rec, c.density_th)) continue;
        double r[4] = {rec.x1, rec.y1, rec.x2, rec.y2};
        for (int k = 0; k < 4; ++k) {
            r[k] += 0.5;
            if (c.scale != 1) r[k] /= c.scale;
            st.segs.push_back((float)r[k]);
        }
    }
    st.valid = true;
}

// cv::LineIterator(img, Point2f, Point2f).count for endpoints inside the image.
// cv::clipLine(Size(W, H), pt1, pt2) (OpenCV imgproc/drawing.cpp; restated from the published algorithm and pinned to
// cv2.clipLine in tests/test_oracle_cv2.py): Cohen-Sutherland on 64-bit integers, the intersection offsets truncated
// toward zero, the second end point clipped against the ALREADY MOVED first one.
bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = W - 1, bottom = H - 1;
    if (W <= 0 || H <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// cv::LineIterator(img, Point2f p1, Point2f p2).count, 8-connected (LSDDetector_custom.cpp:295-296): end points rounded
// half-to-even (Point2f -> Point saturate_cast), clipped to the image when one lies outside - which happens: the clamp
// of checkLineExtremes leaves x in [W-0.5, W) and that rounds to W - and 0 when the clipped line is empty.
int line_iterator_count(float x1, float y1, float x2, float y2, int W, int H) {
    long long ax = cv_roundf(x1), ay = cv_roundf(y1), bx = cv_roundf(x2), by = cv_roundf(y2);
    if (ax < 0 || ax >= W || bx < 0 || bx >= W || ay < 0 || ay >= H || by < 0 || by >= H)
        if (!clip_line(W, H, ax, ay, bx, by)) return 0;
    return (int)std::max(std::llabs(bx - ax), std::llabs(by - ay)) + 1;
}

// LSDDetector_custom.cpp:76-102,268-308 then src/LineExtractor.cc:56-65
void lines_to_keylines(const std::vector<float>& segs, int w, int h, double min_length, int nfeatures,
                       std::vector<plf_keyline>& kls) {
    kls.clear();
    int class_counter = -1;
    for (size_t k = 0; k + 3 < segs.size(); k += 4) {
        float e[4] = {segs[k], segs[k + 1], segs[k + 2], segs[k + 3]};
        for (int q = 0; q < 4; q += 2) {
            if (e[q] < 0) e[q] = 0;
            if (e[q] >= w) e[q] = (float)w - 1.0f;
            if (e[q + 1] < 0) e[q + 1] = 0;
            if (e[q + 1] >= h) e[q + 1] = (float)h - 1.0f;
        }
        double length = (float)std::sqrt(std::pow((double)(e[0] - e[2]), 2) + std::pow((double)(e[1] - e[3]), 2));
        if (!(length > min_length)) continue;
        plf_keyline kl;
        kl.startPointX = e[0]; kl.startPointY = e[1]; kl.endPointX = e[2]; kl.endPointY = e[3];
        kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1]; kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
        kl.lineLength = (float)length;
        kl.numOfPixels = line_iterator_count(e[0], e[1], e[2], e[3], w, h);
        kl.angle = ref_atan2f(kl.endPointY - kl.startPointY, kl.endPointX - kl.startPointX);   // atan2 on floats, :297
        kl.class_id = ++class_counter;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, h);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2;
        kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        kls.push_back(kl);
    }
    if ((int)kls.size() > nfeatures && nfeatures != 0) {
        // oracle rule: stable order (response desc, detection index asc) in place of the unstable std::sort
        std::stable_sort(kls.begin(), kls.end(),
                         [](const plf_keyline& a, const plf_keyline& b) { return a.response > b.response; });
        kls.resize(nfeatures);
        for (int i = 0; i < nfeatures; ++i) kls[i].class_id = i;
    }
}

}  // namespace plfo

// oracle-only diagnostics for the design of the CUDA region grower: number of seeds that start a region, total pixels
// claimed, histogram of region sizes
namespace plfo {
void lsd_stats(const LsdConfig& c, const Img8& img, long long out[8]) {
    const double prec = kPi * c.ang_th / 180;
    const double rho = c.quant / std::sin(prec);
    Lsd L(c);
    LsdState st;
    std::vector<int> taps;
    gaussian_taps_fixed(7, 0.6, taps);
    Img8 g;
    gaussian_blur_u8(img, g, taps.data(), 7);
    resize_linear_exact_u8(g, st.scaled, c.scale);
    L.ll_angle(st.scaled, rho, st.angleDeg);
    L.used.assign((size_t)L.W * L.H, 0);
    std::vector<RegPt> reg;
    for (int k = 0; k < 8; ++k) out[k] = 0;
    for (const auto& op : L.ordered) {
        const size_t idx = (size_t)op.y * L.W + op.x;
        if (L.angles[idx] == kNotDef) continue;
        out[0]++;                       // defined pixels
        if (L.used[idx] != 0) continue;
        double ra;
        L.region_grow(op.x, op.y, reg, ra, prec);
        out[1]++;                       // regions
        out[2] += (long long)reg.size();
        if (reg.size() == 1) out[3]++;
        else if (reg.size() < 4) out[4]++;
        else if (reg.size() < 16) out[5]++;
        else { out[6]++; out[7] += (long long)reg.size(); }
    }
}
}  // namespace plfo

// ---------------------------------------------------------------------------------------------------------------------
// Design aid for the CUDA region grower (TEST INFRASTRUCTURE, like everything here): a lock-step emulation of the
// "one lane per region" streaming scheme of csrc/lsd_stream.cu — up to 32 regions of ONE image in flight in the lanes of
// a warp, every lane running the scalar region_grow state machine (one list entry = 8 neighbours per step) against an
// owner map; tickets = seed positions (earlier position = higher priority); in-order commit.  Returns the segments
// (which must equal lsd_detect's: the protocol is exact whatever the heuristics do) and the step /
// occupancy / kill counters the design was tuned with.
//
// Protocol.  O[q]: 0 free, 0xFFFFFFFF committed (or undefined pixel), else tag = seed position + 1 of the in-flight
// region that claims q.  A region with tag T visiting q: committed or mine -> skip; free -> claim if aligned; claimed by a
// LATER ticket -> steal if aligned (the victim is killed); claimed by an EARLIER in-flight ticket E -> skip, and if it
// would have been aligned remember "T depends on E".  Killing a ticket withdraws its claims and kills every ticket that
// depends on it.  A finished region waits; the commit pointer walks the candidate list in order and commits the
// region of each candidate whose pixel is still uncommitted when its turn comes (growing it first if nobody has).
namespace plfo {
struct StreamSimParams { int window; int fifo; double dperp; int heuristic; int lanes; };
void lsd_stream_sim(const LsdConfig& c, const Img8& img, const StreamSimParams& sp, std::vector<float>& segs, long long st[16]) {
    const double prec = kPi * c.ang_th / 180, p = c.ang_th / 180;
    const double rho = c.quant / std::sin(prec);
    Lsd L(c);
    LsdState stt;
    if (c.scale != 1) {
        const double sigma = (c.scale < 1) ? (c.sigma_scale / c.scale) : c.sigma_scale;
        const unsigned h = (unsigned)std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0)));
        std::vector<int> taps;
        gaussian_taps_fixed(1 + 2 * (int)h, sigma, taps);
        Img8 g;
        gaussian_blur_u8(img, g, taps.data(), 1 + 2 * (int)h);
        resize_linear_exact_u8(g, stt.scaled, c.scale);
    } else stt.scaled = img;
    L.ll_angle(stt.scaled, rho, stt.angleDeg);
    const int W = L.W, H = L.H;
    const double logNT = 5 * (std::log10((double)W) + std::log10((double)H)) / 2 + std::log10(11.0);
    const int minReg = (int)(size_t)(-logNT / std::log10(p));
    for (int k = 0; k < 16; ++k) st[k] = 0;
    segs.clear();
    const uint32_t FREE = 0u, COMMITTED = 0xFFFFFFFFu;
    std::vector<uint32_t> O((size_t)W * H, COMMITTED);
    std::vector<int> S;
    for (const auto& op : L.ordered) {
        const size_t idx = (size_t)op.y * W + op.x;
        if (L.angles[idx] == kNotDef) continue;
        O[idx] = FREE;
        S.push_back((int)idx);
    }
    const int ns = (int)S.size();
    const int NL = sp.lanes;
    enum { GROW = 1, DONE = 2 };
    struct Ticket {
        int state = GROW, lane = -1, i = 0; float sumdx = 0, sumdy = 0; double angle = 0; int sx = 0, sy = 0;
        std::vector<int> list; std::vector<uint32_t> deps; bool depAll = false;   // depAll: more than 4 dependencies -> on every earlier ticket
        long long epoch = 0;          // kill counter when the ticket started
    };
    long long killEpoch = 0;
    std::vector<long long> killStamp(4096, -1);     // hashed by tag: epoch of the last kill that hashes there (collisions only invalidate more)
    auto dep_broken = [&](const Ticket& t) {
        if (t.depAll) return killEpoch > t.epoch;
        for (uint32_t d : t.deps) if (killStamp[d & 4095] >= t.epoch) return true;
        return false;
    };
    std::map<uint32_t, Ticket> T;                 // live tickets by tag
    std::vector<uint32_t> laneTag(NL, 0);         // tag growing in each lane (0 = idle)
    std::vector<int> laneDone(NL, 0);             // finished, uncommitted tickets per lane
    std::vector<std::pair<int, uint32_t>> blocked;   // (candidate position, blocker tag): not picked while the blocker grows
    std::vector<int> ready;                          // candidate positions released by a finished / killed blocker or a killed ticket
    const size_t BLOCKED_CAP = 1024, READY_CAP = 1024, TICKET_CAP = 2048;
    int cp = 0, scanPos = 0;
    static const int ddx[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, ddy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
    std::vector<uint32_t> killQueue;
    auto start_ticket = [&](int l, int pos) {
        Ticket t;
        const int q = S[pos];
        t.lane = l; t.sx = q % W; t.sy = q / W; t.angle = L.angles[q];
        t.sumdx = (float)std::cos(t.angle); t.sumdy = (float)std::sin(t.angle);
        t.list.push_back(q);
        const uint32_t tag = (uint32_t)pos + 1u, o = O[q];
        if (o != FREE && o != COMMITTED && o > tag) killQueue.push_back(o);     // a forced ticket steals its seed
        O[q] = tag;
        laneTag[l] = tag;
        t.epoch = killEpoch;
        T[tag] = t;
    };
    auto kill_all = [&]() {          // kill the queued victims and, transitively, every ticket that depends on one
        while (!killQueue.empty()) {
            const uint32_t v = killQueue.back();
            killQueue.pop_back();
            auto it = T.find(v);
            if (it == T.end()) continue;
            Ticket& t = it->second;
            for (int q : t.list) if (O[q] == v) O[q] = FREE;
            st[12] += (long long)t.list.size();
            st[3] += ((long long)t.list.size() + 31) / 32;            // cooperative withdrawal steps
            if (t.state == GROW) laneTag[t.lane] = 0; else laneDone[t.lane]--;
            T.erase(it);
            st[11]++;
            if (ready.size() < READY_CAP) ready.push_back((int)v - 1); else scanPos = std::min(scanPos, (int)v - 1);
            killStamp[v & 4095] = killEpoch++;          // dependents find out lazily: while growing, or at their commit
        }
    };
    long guard = 0;
    while (true) {
        if (++guard > 2000000) { st[15] = -1; break; }
        // ---- 1. commit walk ----
        {
            int walked = 0;
            while (cp < ns) {
                const int q = S[cp];
                const uint32_t o = O[q], tag = (uint32_t)cp + 1u;
                if (o == COMMITTED) { ++cp; ++walked; continue; }
                if (o == tag) {
                    auto it = T.find(tag);
                    if (it == T.end()) { st[15] = -4; break; }
                    if (it->second.state != DONE) break;               // still growing
                    if (dep_broken(it->second)) { killQueue.push_back(tag); kill_all(); st[10]++; continue; }   // a ticket it relied on was killed: grow again
                    Ticket& t = it->second;
                    for (int qq : t.list) O[qq] = COMMITTED;
                    st[3] += ((long long)t.list.size() + 31) / 32;
                    if ((int)t.list.size() >= minReg) {
                        std::vector<RegPt> reg(t.list.size());
                        for (size_t j = 0; j < t.list.size(); ++j) reg[j] = {t.list[j] % W, t.list[j] / W};
                        Lsd::Rect rec;
                        L.region2rect(reg, t.angle, prec, p, rec);
                        double r[4] = {rec.x1, rec.y1, rec.x2, rec.y2};
                        for (int k = 0; k < 4; ++k) { r[k] += 0.5; if (c.scale != 1) r[k] /= c.scale; segs.push_back((float)r[k]); }
                    }
                    laneDone[t.lane]--;
                    if (t.lane == 0) st[2] += 0, st[12] += 0;
                    if ((long long)t.list.size() > st[14]) st[14] = (long long)t.list.size();
                    if (t.lane == 0) st[13] += (long long)t.list.size();
                    T.erase(it);
                    st[4]++;
                    ++cp; ++walked;
                    continue;
                }
                // free, or claimed by a later ticket: this candidate starts its own region now on the reserved lane 0
                if (laneTag[0] == 0) { start_ticket(0, cp); st[5]++; kill_all(); }
                break;
            }
            st[2] += (walked + 31) / 32;
        }
        if (cp >= ns && T.empty()) break;
        // ---- 2. pick ----
        {
            auto lane_free = [&](int l) { return laneTag[l] == 0 && laneDone[l] < sp.fifo; };
            int nIdle = 0;
            for (int l = 1; l < NL; ++l) nIdle += lane_free(l);
            if (scanPos < cp) scanPos = cp;
            auto find_blocker = [&](int q) -> uint32_t {
                if (!sp.heuristic) return 0u;
                const int x = q % W, y = q / W;
                const double a = L.angles[q];
                for (int l = 0; l < NL; ++l) {
                    if (!laneTag[l]) continue;
                    const Ticket& t = T[laneTag[l]];
                    double d = std::fabs(a - t.angle);
                    if (d > 1.5 * kPi) d = std::fabs(d - 2 * kPi);
                    if (d > prec) continue;
                    const double perp = std::fabs(-(x - t.sx) * std::sin(t.angle) + (y - t.sy) * std::cos(t.angle));
                    if (perp <= sp.dperp) return laneTag[l];
                }
                return 0u;
            };
            // released candidates first
            for (size_t r = 0; r < ready.size() && nIdle > 0 && T.size() < TICKET_CAP;) {
                const int pos = ready[r];
                {
                    const uint32_t o = O[S[pos]];
                    if (pos < cp || o == COMMITTED || (o != FREE && o <= (uint32_t)pos + 1u)) { ready[r] = ready.back(); ready.pop_back(); continue; }
                }
                if (pos == cp) { ++r; continue; }
                const uint32_t blocker = find_blocker(S[pos]);
                if (blocker) {
                    if (blocked.size() < BLOCKED_CAP) { blocked.push_back({pos, blocker}); st[6]++; ready[r] = ready.back(); ready.pop_back(); }
                    else ++r;
                    continue;
                }
                for (int l = 1; l < NL; ++l) if (lane_free(l)) { start_ticket(l, pos); --nIdle; st[7]++; break; }
                ready[r] = ready.back(); ready.pop_back();
            }
            int scans = 0;
            while (nIdle > 0 && scanPos < ns && scanPos < cp + sp.window && scans < 2 && blocked.size() + 32 <= BLOCKED_CAP && T.size() + 32 <= TICKET_CAP) {
                ++scans; st[1]++;
                const int end = std::min(ns, scanPos + 32);
                int pos = scanPos;
                for (; pos < end && nIdle > 0; ++pos) {
                    const int q = S[pos];
                    if (O[q] != FREE || pos == cp) continue;
                    const uint32_t blocker = find_blocker(q);
                    if (blocker) { blocked.push_back({pos, blocker}); st[6]++; continue; }
                    for (int l = 1; l < NL; ++l) if (lane_free(l)) { start_ticket(l, pos); --nIdle; st[7]++; break; }
                }
                scanPos = pos;
            }
            kill_all();
        }
        // ---- 3. one lock-step grow step ----
        int active = 0;
        std::vector<std::array<uint32_t, 8>> snap(NL);
        for (int l = 0; l < NL; ++l) {
            if (!laneTag[l]) continue;
            ++active;
            Ticket& t = T[laneTag[l]];
            const int e = t.list[t.i], ex = e % W, ey = e / W;
            for (int k = 0; k < 8; ++k) {
                const int xx = ex + ddx[k], yy = ey + ddy[k];
                snap[l][k] = (xx < 0 || yy < 0 || xx >= W || yy >= H) ? COMMITTED : O[(size_t)yy * W + xx];
            }
        }
        for (int k = 0; k < 8; ++k)
            for (int l = 0; l < NL; ++l) {
                const uint32_t tag = laneTag[l];
                if (!tag) continue;
                Ticket& t = T[tag];
                const uint32_t o = snap[l][k];
                if (o == COMMITTED || o == tag) continue;
                const int e = t.list[t.i];
                const int xx = e % W + ddx[k], yy = e / W + ddy[k];
                if (!L.is_aligned(xx, yy, t.angle, prec)) continue;
                const size_t q = (size_t)yy * W + xx;
                const uint32_t o2 = O[q];                                   // fresh read before the claim
                if (o2 == COMMITTED || o2 == tag) continue;
                if (o2 != FREE && o2 < tag) {                               // an earlier in-flight ticket has it: I depend on that ticket
                    if (!t.depAll && std::find(t.deps.begin(), t.deps.end(), o2) == t.deps.end()) {
                        if (t.deps.size() < 4) t.deps.push_back(o2); else t.depAll = true;
                    }
                    continue;
                }
                if (o2 != FREE) { killQueue.push_back(o2); st[8]++; }       // steal from a later ticket
                O[q] = tag;
                t.list.push_back((int)q);
                const double ang = L.angles[q];
                t.sumdx += (float)std::cos((double)(float)ang);
                t.sumdy += (float)std::sin((double)(float)ang);
                t.angle = fast_atan2(t.sumdy, t.sumdx) * kDegToRad;
            }
        for (int l = 0; l < NL; ++l) {
            const uint32_t tag = laneTag[l];
            if (!tag) continue;
            Ticket& t = T[tag];
            if (++t.i >= (int)t.list.size()) {
                t.state = DONE; laneTag[l] = 0; laneDone[l]++;
                for (size_t b = 0; b < blocked.size();)                     // candidates held back for this region may go now
                    if (blocked[b].second == tag) {
                        if (ready.size() < READY_CAP) ready.push_back(blocked[b].first); else scanPos = std::min(scanPos, blocked[b].first);
                        blocked[b] = blocked.back(); blocked.pop_back();
                    } else ++b;
            }
        }
        st[0]++;
        st[9] += active;
        for (int l = 0; l < NL; ++l) if (laneTag[l] && dep_broken(T[laneTag[l]])) killQueue.push_back(laneTag[l]);
        kill_all();
        // blockers that died release their candidates as well
        for (size_t b = 0; b < blocked.size();)
            if (T.find(blocked[b].second) == T.end()) {
                if (ready.size() < READY_CAP) ready.push_back(blocked[b].first); else scanPos = std::min(scanPos, blocked[b].first);
                blocked[b] = blocked.back(); blocked.pop_back();
            } else ++b;
        st[6] = std::max<long long>(st[6], (long long)T.size());
    }
    (void)minReg;
}
}  // namespace plfo
