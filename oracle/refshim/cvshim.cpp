// TEST INFRASTRUCTURE ONLY — the arithmetic half of the OpenCV stand-in (oracle/refshim/opencv2/core.hpp): every
// primitive forwards to the restatement in oracle/cpp that tests/test_oracle_cv2.py pins bit-exactly to cv2 4.13.
#include <opencv2/core.hpp>
#include <fstream>
#include <sstream>
#include "../cpp/prims.h"
#include "../cpp/orb.h"
#include "../cpp/lsd.h"

namespace cv {

static plfo::Img8 to_img8(const Mat& m) {
    if (m.depth() != CV_8U) throw std::runtime_error("cvshim: 8-bit image expected");
    plfo::Img8 im(m.cols, m.rows);
    for (int y = 0; y < m.rows; y++) memcpy(im.row(y), m.ptr(y), (size_t)m.cols);
    return im;
}
static void from_img8(const plfo::Img8& im, Mat& dst) {
    dst.create(im.h, im.w, CV_8U);
    for (int y = 0; y < im.h; y++) memcpy(dst.ptr(y), im.row(y), (size_t)im.w);
}

float fastAtan2(float y, float x) { return plfo::fast_atan2(y, x); }

// cv::FAST(sub-image, keypoints, threshold, nonmax): KeyPoint(x, y, 7.f, -1, score) in raster order.
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
    if (!nonmaxSuppression) throw std::runtime_error("cvshim: FAST without NMS is not on the path");
    keypoints.clear();
    plfo::Img8 im = to_img8(image);
    std::vector<plfo::Cand> out;
    plfo::fast_window(im, 0, 0, im.w, im.h, threshold, out);
    for (const plfo::Cand& c : out) keypoints.push_back(KeyPoint(c.x, c.y, 7.f, -1.f, c.resp));
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int border) {
    if (ksize.width != ksize.height || (sigmaY != 0 && sigmaY != sigmaX) || (border & ~BORDER_ISOLATED) != BORDER_REFLECT_101)
        throw std::runtime_error("cvshim: GaussianBlur form not on the path");
    std::vector<int> taps;
    plfo::gaussian_taps_fixed(ksize.width, sigmaX, taps);
    plfo::Img8 in = to_img8(src), out;
    plfo::gaussian_blur_u8(in, out, taps.data(), ksize.width);
    from_img8(out, dst);
}

void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
    if (interpolation != INTER_LINEAR || fx != 0 || fy != 0) throw std::runtime_error("cvshim: resize form not on the path");
    plfo::Img8 in = to_img8(src), out;
    plfo::resize_linear_u8(in, out, dsize.width, dsize.height);
    from_img8(out, dst);   // dst keeps its buffer when the size fits (ComputePyramid writes into the bordered block)
}

void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || src.depth() != CV_8U)
        throw std::runtime_error("cvshim: copyMakeBorder form not on the path");
    const int w = src.cols, h = src.rows;
    plfo::Img8 in = to_img8(src);   // src may be the interior view of dst (src/ORBextractor.cc:1167)
    dst.create(h + top + bottom, w + left + right, CV_8U);
    for (int y = 0; y < dst.rows; y++) {
        const uint8_t* s = in.row(plfo::reflect101(y - top, h));
        uchar* d = dst.ptr(y);
        for (int x = 0; x < dst.cols; x++) d[x] = s[plfo::reflect101(x - left, w)];
    }
}

void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize) {
    if (ddepth != CV_16S || ksize != 3 || dx + dy != 1) throw std::runtime_error("cvshim: Sobel form not on the path");
    plfo::Img8 in = to_img8(src);
    plfo::Img16 gx, gy;
    plfo::sobel3_16s(in, gx, gy);
    const plfo::Img16& g = dx ? gx : gy;
    dst.create(g.h, g.w, CV_16S);
    for (int y = 0; y < g.h; y++) memcpy(dst.ptr(y), g.d.data() + (size_t)y * g.w, (size_t)g.w * 2);
}

void cvtColor(const Mat&, Mat&, int) { throw std::runtime_error("cvshim: cvtColor is not on the path"); }
void pyrDown(const Mat&, Mat&, const Size&) { throw std::runtime_error("cvshim: pyrDown is not on the path"); }
void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) {
    throw std::runtime_error("cvshim: KeyPointsFilter is only named by dead code");
}

namespace {
class LsdImpl : public LineSegmentDetector {
public:
    plfo::LsdConfig cfg;
    void detect(const Mat& image, std::vector<Vec4f>& lines) override {
        plfo::Img8 im = to_img8(image);
        plfo::LsdState st;
        plfo::lsd_detect(cfg, im, st);
        lines.clear();
        for (size_t k = 0; k + 3 < st.segs.size(); k += 4)
            lines.push_back(Vec4f(st.segs[k], st.segs[k + 1], st.segs[k + 2], st.segs[k + 3]));
    }
};
}  // namespace

Ptr<LineSegmentDetector> createLineSegmentDetector(int refine, double scale, double sigma_scale, double quant,
                                                   double ang_th, double log_eps, double density_th, int n_bins) {
    auto p = std::make_shared<LsdImpl>();
    p->cfg.refine = refine; p->cfg.scale = scale; p->cfg.sigma_scale = sigma_scale; p->cfg.quant = quant;
    p->cfg.ang_th = ang_th; p->cfg.log_eps = log_eps; p->cfg.density_th = density_th; p->cfg.n_bins = n_bins;
    return p;
}

// 8-connected cv::LineIterator: count = max(|dx|, |dy|) + 1 after cv::clipLine (the cv2-pinned restatement), 0 when the
// clipped line is empty.
LineIterator::LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity, bool) {
    if (connectivity != 8) throw std::runtime_error("cvshim: LineIterator connectivity");
    long long ax = pt1.x, ay = pt1.y, bx = pt2.x, by = pt2.y;
    count = 0;
    if (ax < 0 || ax >= img.cols || bx < 0 || bx >= img.cols || ay < 0 || ay >= img.rows || by < 0 || by >= img.rows)
        if (!plfo::clip_line(img.cols, img.rows, ax, ay, bx, by)) return;
    count = (int)std::max(std::llabs(bx - ax), std::llabs(by - ay)) + 1;
}

// BFMatcher(NORM_HAMMING).knnMatch: k nearest by (distance, train index) — SURVEY 8c fact 6, pinned to cv2 in
// tests/test_oracle_cv2.py.
void BFMatcher::knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const {
    if (norm_ != NORM_HAMMING) throw std::runtime_error("cvshim: BFMatcher norm");
    matches.assign(query.rows, std::vector<DMatch>());
    const int nb = query.cols;
    std::vector<std::pair<int, int>> d(train.rows);
    for (int q = 0; q < query.rows; q++) {
        const uchar* a = query.ptr(q);
        for (int t = 0; t < train.rows; t++) {
            const uchar* b = train.ptr(t);
            int s = 0;
            for (int i = 0; i < nb; i++) s += __builtin_popcount((unsigned)(a[i] ^ b[i]));
            d[t] = std::make_pair(s, t);
        }
        const int kk = std::min(k, train.rows);
        std::partial_sort(d.begin(), d.begin() + kk, d.end());
        for (int i = 0; i < kk; i++) matches[q].push_back(DMatch(q, d[i].second, (float)d[i].first));
    }
}

static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

FileStorage::FileStorage(const std::string& path, int) {
    std::ifstream f(path);
    if (!f) return;
    opened_ = true;
    std::string line;
    while (std::getline(f, line)) {
        size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (line.empty() || line[0] == '%' || line[0] == ' ' || line[0] == '-') continue;
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string key = trim(line.substr(0, c)), val = trim(line.substr(c + 1));
        if (key.empty() || val.empty()) continue;
        if (val[0] == '"' || val[0] == '\'') { kv_[key] = FileNode(FileNode::STR, val.substr(1, val.size() - 2)); continue; }
        char* end = nullptr;
        std::strtod(val.c_str(), &end);
        if (end && *end == 0) {
            bool isint = val.find_first_of(".eE") == std::string::npos;
            kv_[key] = FileNode(isint ? FileNode::INT : FileNode::REAL, val);
        } else kv_[key] = FileNode(FileNode::STR, val);
    }
}
FileNode FileStorage::operator[](const std::string& key) const {
    auto it = kv_.find(key);
    return it == kv_.end() ? FileNode() : it->second;
}

}  // namespace cv
