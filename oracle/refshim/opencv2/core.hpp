// TEST INFRASTRUCTURE ONLY — a one-header stand-in for the slice of the OpenCV C++ API that the reference's
// frontend sources use, so that those sources compile UNMODIFIED from /root/reference into oracle/_ref
// (recipe: oracle/build_ref.py).  OpenCV itself is an un-vendored dependency of the reference (README pins 3.3.1) and
// no OpenCV C++ headers exist in this image.  Containers and glue (Mat, Point_, KeyPoint, InputArray ...) are written
// here from the documented API; the ARITHMETIC primitives (FAST, resize, GaussianBlur, Sobel, LSD, fastAtan2 ...)
// forward to the restatements in oracle/cpp/prims.cpp / orb.cpp / lsd.cpp, which tests/test_oracle_cv2.py pins
// bit-exactly to the real cv2 4.13 wheel.  Nothing in the product includes this file.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cassert>
#include <memory>
#include <vector>
#include <string>
#include <sstream>
#include <map>
#include <stdexcept>
#include <algorithm>
#include <iostream>

#define CV_EXPORTS
#define CV_EXPORTS_W
#define CV_WRAP
#define CV_OUT
#define CV_IN_OUT
#define CV_PI 3.1415926535897932384626433832795

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_16SC1 CV_16S
#define CV_32SC1 CV_32S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

typedef unsigned char uchar;
typedef unsigned short ushort;

// cvRound: lrint semantics (round half to even), cvFloor / cvCeil as in OpenCV's fast_math.hpp.
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

// opencv2/core/base.hpp makes these std names visible inside namespace cv; unqualified max / sqrt / pow / exp / abs
// in the line_descriptor sources (namespace cv::line_descriptor) bind through them.
using std::min; using std::max; using std::abs; using std::swap; using std::sqrt; using std::exp; using std::pow; using std::log;

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
typedef std::string String;

template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline float saturate_cast<float>(double v) { return (float)v; }
template <> inline double saturate_cast<double>(double v) { return v; }

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4,
       BORDER_ISOLATED = 16 };
enum { COLOR_BGR2GRAY = 6 };
enum { LSD_REFINE_NONE = 0, LSD_REFINE_STD = 1, LSD_REFINE_ADV = 2 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x(saturate_cast<T>(p.x)), y(saturate_cast<T>(p.y)) {}
    Point_& operator*=(float s) { x = saturate_cast<T>(x * s); y = saturate_cast<T>(y * s); return *this; }
    Point_& operator+=(const Point_& p) { x += p.x; y += p.y; return *this; }
    bool operator==(const Point_& p) const { return x == p.x && y == p.y; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& s) const { return width == s.width && height == s.height; }
    bool operator!=(const Size_& s) const { return !(*this == s); }
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, "4-element form"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
};
typedef Vec<float, 4> Vec4f;
typedef Vec<int, 4> Vec4i;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int cls = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(cls) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

static inline size_t cvshim_elem_size(int type) {
    switch (type & 7) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2; case CV_64F: return 8;
                        default: return 4; }
}

class _OutputArray;

// Reference-counted 2-D single-channel matrix with ROI views (the subset of cv::Mat the reference touches).
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;   // bytes per row
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), type_(type) {
        step = step_ ? step_ : (size_t)c * cvshim_elem_size(type);
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;   // cv::Mat::create keeps a fitting buffer
        size_t es = cvshim_elem_size(type);
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * c * es + 64);
        rows = r; cols = c; type_ = type; step = (size_t)c * es; data = buf_->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf_.reset(); rows = cols = 0; data = nullptr; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return 1; }
    size_t elemSize() const { return cvshim_elem_size(type_); }
    size_t step1() const { return step / cvshim_elem_size(type_); }
    Size size() const { return Size(cols, rows); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat col(int i) const { return colRange(i, i + 1); }
    // CV_32F small-matrix algebra of the pose code in src/ORBmatcher.cc (3x3 / 3x1 / 4x4): plain float loops.  The pins
    // that go through it use identity rotations and exactly representable operands, so no rounding rule is at stake.
    Mat t() const {
        Mat m(cols, rows, type_);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.at<float>(x, y) = at<float>(y, x);
        return m;
    }
    double dot(const Mat& o) const {
        double acc = 0;
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) acc += (double)at<float>(y, x) * (double)o.at<float>(y, x);
        return acc;
    }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memmove(dst.ptr(y), ptr(y), (size_t)cols * elemSize());
    }
    void copyTo(const _OutputArray& dst) const;
    // 8U -> 16S / 32F and identity are all the reference asks for (src/Frame.cc:1076,1095).
    void convertTo(Mat& dst, int type) const {
        Mat out(rows, cols, type);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++) {
                double v;
                switch (depth()) { case CV_8U: v = at<uchar>(y, x); break; case CV_16S: v = at<short>(y, x); break;
                                   case CV_32F: v = at<float>(y, x); break;
                                   default: throw std::runtime_error("cvshim: convertTo source depth"); }
                switch (type & 7) { case CV_8U: out.at<uchar>(y, x) = (uchar)v; break;
                                    case CV_16S: out.at<short>(y, x) = (short)v; break;
                                    case CV_32F: out.at<float>(y, x) = (float)v; break;
                                    default: throw std::runtime_error("cvshim: convertTo target depth"); }
            }
        dst = out;
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); memset(m.data, 0, (size_t)r * m.step); return m; }
    static Mat ones(int r, int c, int type) {
        Mat m = zeros(r, c, type);
        for (int y = 0; y < r; y++) for (int x = 0; x < c; x++) {
            if ((type & 7) == CV_32F) m.at<float>(y, x) = 1.f; else if ((type & 7) == CV_16S) m.at<short>(y, x) = 1;
            else m.at<uchar>(y, x) = 1;
        }
        return m;
    }
    void reserve(size_t) {}
    void push_back(const Mat& m) {   // appends rows (dead-code callers only: src/LineMatcher.cpp:164-166)
        if (m.empty()) return;
        Mat out(rows + m.rows, m.cols, m.type());
        for (int y = 0; y < rows; y++) memcpy(out.ptr(y), ptr(y), (size_t)cols * elemSize());
        for (int y = 0; y < m.rows; y++) memcpy(out.ptr(rows + y), m.ptr(y), (size_t)m.cols * m.elemSize());
        *this = out;
    }
private:
    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

// Mat - scalar for CV_16S (src/Frame.cc:1077,1096: patch minus its centre value; |result| <= 255, no saturation).
static inline Mat operator-(const Mat& a, double s) {
    Mat out(a.rows, a.cols, a.type());
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) {
        if (a.depth() == CV_16S) {
            double v = (double)a.at<short>(y, x) - s;
            out.at<short>(y, x) = (short)std::max(-32768.0, std::min(32767.0, std::nearbyint(v)));
        } else if (a.depth() == CV_32F) out.at<float>(y, x) = (float)(a.at<float>(y, x) - s);
        else throw std::runtime_error("cvshim: Mat - scalar depth");
    }
    return out;
}

static inline Mat operator*(const Mat& a, const Mat& b) {
    if (a.cols != b.rows || a.depth() != CV_32F || b.depth() != CV_32F) throw std::runtime_error("cvshim: Mat * Mat shape / depth");
    Mat out(a.rows, b.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < b.cols; x++) {
        float acc = 0.f;
        for (int k = 0; k < a.cols; k++) acc += a.at<float>(y, k) * b.at<float>(k, x);
        out.at<float>(y, x) = acc;
    }
    return out;
}
static inline Mat cvshim_zip(const Mat& a, const Mat& b, float sa, float sb) {
    if (a.rows != b.rows || a.cols != b.cols || a.depth() != CV_32F || b.depth() != CV_32F) throw std::runtime_error("cvshim: Mat +- Mat shape / depth");
    Mat out(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) out.at<float>(y, x) = sa * a.at<float>(y, x) + sb * b.at<float>(y, x);
    return out;
}
static inline Mat operator+(const Mat& a, const Mat& b) { return cvshim_zip(a, b, 1.f, 1.f); }
static inline Mat operator-(const Mat& a, const Mat& b) { return cvshim_zip(a, b, 1.f, -1.f); }
static inline Mat operator-(const Mat& a) {
    Mat out(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) out.at<float>(y, x) = -a.at<float>(y, x);
    return out;
}
static inline Mat operator/(const Mat& a, double s) {
    Mat out(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) out.at<float>(y, x) = (float)(a.at<float>(y, x) / s);
    return out;
}
static inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }

// cv::norm(a, b, NORM_L1) for CV_16S / CV_8U / CV_32F (src/Frame.cc:1098).
static inline double norm(const Mat& a, const Mat& b, int normType) {
    if (normType != NORM_L1) throw std::runtime_error("cvshim: norm type");
    double s = 0;
    for (int y = 0; y < a.rows; y++) for (int x = 0; x < a.cols; x++) {
        if (a.depth() == CV_16S) s += std::abs((int)a.at<short>(y, x) - (int)b.at<short>(y, x));
        else if (a.depth() == CV_8U) s += std::abs((int)a.at<uchar>(y, x) - (int)b.at<uchar>(y, x));
        else s += std::fabs((double)a.at<float>(y, x) - (double)b.at<float>(y, x));
    }
    return s;
}

template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(const Mat& m) : Mat(m) {}
};

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray() : m_(nullptr) {}
    _OutputArray(Mat& m) : m_(&m) {}
    _OutputArray(const Mat& m) : own_(m), m_(&own_) {}   // fixed-size view (e.g. descriptors.row(i)): shares the data
    void release() const { if (m_) m_->release(); }
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void create(Size s, int type) const { if (m_) m_->create(s, type); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    Mat* mat() const { return m_; }
private:
    mutable Mat own_;
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline void Mat::copyTo(const _OutputArray& dst) const { if (dst.mat()) copyTo(*dst.mat()); }
static inline const _InputArray& noArray() { static _InputArray a; return a; }

class FileNode {
public:
    enum { NONE = 0, INT = 1, REAL = 2, FLOAT = 2, STR = 3, STRING = 3 };
    FileNode() : type_(NONE) {}
    FileNode(int t, const std::string& s) : type_(t), s_(s) {}
    int type() const { return type_; }
    bool empty() const { return type_ == NONE; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    // sequence access: named by DBoW2's YAML load() (TemplatedVocabulary.h:1660-1720), which oracle/_ref never calls
    size_t size() const { return 0; }
    FileNode operator[](unsigned int) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    operator int() const { return type_ == NONE ? 0 : (int)std::lround(std::atof(s_.c_str())); }
    operator float() const { return type_ == NONE ? 0.f : (float)std::atof(s_.c_str()); }
    operator double() const { return type_ == NONE ? 0.0 : std::atof(s_.c_str()); }
    operator std::string() const { return s_; }
private:
    int type_;
    std::string s_;
};
// Reads the flat "key: value" subset of OpenCV's YAML that Examples/*/Config/*.yaml use (cvshim.cpp).
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string& path, int flags);
    bool isOpened() const { return opened_; }
    FileNode operator[](const std::string& key) const;
    FileNode operator[](const char* key) const { return (*this)[std::string(key)]; }
    void release() {}
    template <typename T> FileStorage& operator<<(const T&) { return *this; }
private:
    bool opened_ = false;
    std::map<std::string, FileNode> kv_;
};

class Algorithm {
public:
    virtual ~Algorithm() {}
    virtual void read(const FileNode&) {}
    virtual void write(FileStorage&) const {}
};

// ---- arithmetic primitives: forwarded to the cv2-pinned restatements (cvshim.cpp) ----
float fastAtan2(float y, float x);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int border = BORDER_DEFAULT);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize = 3);
void cvtColor(const Mat& src, Mat& dst, int code);                       // never reached: inputs are 1-channel
void pyrDown(const Mat& src, Mat& dst, const Size& dstsize = Size());    // never reached: numOctaves == 1

class LineSegmentDetector : public Algorithm {
public:
    virtual void detect(const Mat& image, std::vector<Vec4f>& lines) = 0;
};
Ptr<LineSegmentDetector> createLineSegmentDetector(int refine = LSD_REFINE_STD, double scale = 0.8,
                                                   double sigma_scale = 0.6, double quant = 2.0, double ang_th = 22.5,
                                                   double log_eps = 0, double density_th = 0.7, int n_bins = 1024);

class LineIterator {
public:
    LineIterator(const Mat& img, Point pt1, Point pt2, int connectivity = 8, bool leftToRight = false);
    int count;
};

class BFMatcher {
public:
    BFMatcher(int normType = NORM_L2, bool crossCheck = false) : norm_(normType) { (void)crossCheck; }
    static Ptr<BFMatcher> create(int normType = NORM_L2, bool crossCheck = false) {
        return Ptr<BFMatcher>(new BFMatcher(normType, crossCheck));
    }
    void knnMatch(const Mat& query, const Mat& train, std::vector<std::vector<DMatch>>& matches, int k) const;
private:
    int norm_;
};

struct KeyPointsFilter {   // only named by the dead ComputeKeyPointsOld (src/ORBextractor.cc:880-1057)
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv
