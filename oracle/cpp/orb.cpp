// TEST INFRASTRUCTURE ONLY — CPU oracle, ORB part.  Each function cites the reference lines it restates
// (paths relative to /root/reference).  Built with -ffp-contract=off: float expressions are evaluated exactly
// as written (declared ground truth, SURVEY.md §7 hard part 3).
#include "orb.h"
#include <algorithm>
#include <cstring>
#include <list>

namespace plfo {

static const int kPattern[1024] = {
#include "../../pli-slam_b200/csrc/orb_pattern.inc"
};

static const int EDGE_THRESHOLD = 19;   // src/ORBextractor.cc:72
static const int HALF_PATCH = 15;       // :71
static const int PATCH_SIZE = 31;       // :70

// src/ORBextractor.cc:408-468
void orb_tables(const OrbConfig& c, OrbTables& t) {
    const int n = c.nlevels;
    t.scale.assign(n, 1.f);
    t.sigma2.assign(n, 1.f);
    for (int i = 1; i < n; ++i) {
        t.scale[i] = t.scale[i - 1] * c.scaleFactor;
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    t.invScale.resize(n);
    t.invSigma2.resize(n);
    for (int i = 0; i < n; ++i) {
        t.invScale[i] = 1.0f / t.scale[i];
        t.invSigma2[i] = 1.0f / t.sigma2[i];
    }
    t.nPerLevel.assign(n, 0);
    float factor = 1.0f / c.scaleFactor;
    float nDesired = c.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        t.nPerLevel[l] = cv_round(nDesired);
        sum += t.nPerLevel[l];
        nDesired *= factor;
    }
    t.nPerLevel[n - 1] = std::max(c.nfeatures - sum, 0);
    // umax, :452-467
    int v, v0, vmax = cv_floor(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) t.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
}

// FAST-9/16 score: (largest t for which the pixel is still a corner) — cv::FAST's cornerScore<16>; SURVEY A1.
// Returns M-1 where M = max over the 16 arcs of 9 contiguous circle pixels of min(|d|) with one sign.
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static int fast_score(const Img8& img, int x, int y) {
    int d[16];
    const int p = img.at(y, x);
    for (int k = 0; k < 16; ++k) d[k] = p - img.at(y + kCircle[k][1], x + kCircle[k][0]);
    int best = -1000;
    for (int s = 0; s < 16; ++s) {
        int mn = 1000, mx = -1000;
        for (int k = 0; k < 9; ++k) {
            int v = d[(s + k) & 15];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

// cv::FAST(sub, keypoints, th, true) on the window; call sites src/ORBextractor.cc:808-809,827-828.
void fast_window(const Img8& img, int x0, int y0, int x1, int y1, int th, std::vector<Cand>& out) {
    const int cols = x1 - x0, rows = y1 - y0;
    if (cols < 7 || rows < 7) return;
    const int aw = cols - 6, ah = rows - 6;          // detection area
    static thread_local std::vector<int> sc;
    sc.assign((size_t)aw * ah, 0);
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = kCircle[k][1] * img.w + kCircle[k][0];
    bool any = false;
    for (int i = 0; i < ah; ++i) {
        const uint8_t* r0 = img.row(y0 + 3 + i) + x0 + 3;
        for (int j = 0; j < aw; ++j) {
            // Early reject (the high-speed test every FAST implementation starts with): a 9-arc of the 16-ring holds
            // at least one pixel of each opposite pair (k, k+8), so a corner at threshold th needs, in EVERY pair, a
            // pixel brighter than p+th (bright corner) or in every pair one darker than p-th (dark corner).  A
            // necessary condition only: survivors get the full score and the `>= th` test, so the list is unchanged.
            const uint8_t* c = r0 + j;
            const int p = c[0], hi = p + th, lo = p - th;
            int a = c[off[0]], b = c[off[8]];
            bool bright = a > hi || b > hi, dark = a < lo || b < lo;
            if (!bright && !dark) continue;
            a = c[off[4]]; b = c[off[12]];
            bright = bright && (a > hi || b > hi); dark = dark && (a < lo || b < lo);
            if (!bright && !dark) continue;
            for (int k = 1; k < 8 && (bright || dark); ++k) {
                if (k == 4) continue;
                a = c[off[k]]; b = c[off[k + 8]];
                bright = bright && (a > hi || b > hi);
                dark = dark && (a < lo || b < lo);
            }
            if (!bright && !dark) continue;
            const int s = fast_score(img, x0 + 3 + j, y0 + 3 + i);
            if (s >= th) { sc[(size_t)i * aw + j] = s; any = true; }   // non-corners at this threshold score 0
        }
    }
    if (!any) return;
    for (int i = 0; i < ah; ++i)
        for (int j = 0; j < aw; ++j) {
            int s = sc[(size_t)i * aw + j];
            if (s == 0) continue;
            bool ismax = true;
            for (int dy = -1; dy <= 1 && ismax; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    int ii = i + dy, jj = j + dx;
                    int n = (ii < 0 || jj < 0 || ii >= ah || jj >= aw) ? 0 : sc[(size_t)ii * aw + jj];
                    if (!(s > n)) { ismax = false; break; }
                }
            if (ismax) out.push_back({(float)(j + 3), (float)(i + 3), (float)s});
        }
}

// src/ORBextractor.cc:769-854 (one level)
void level_candidates(const Img8& lvl, int iniTh, int minTh, std::vector<Cand>& out) {
    out.clear();
    const float W = 30;
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = lvl.w - EDGE_THRESHOLD + 3, maxBorderY = lvl.h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<Cand> cell;
    for (int i = 0; i < nRows; ++i) {
        const int iniY = minBorderY + i * hCell;
        int maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = maxBorderY;
        for (int j = 0; j < nCols; ++j) {
            const int iniX = minBorderX + j * wCell;
            int maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = maxBorderX;
            cell.clear();
            fast_window(lvl, iniX, iniY, maxX, maxY, iniTh, cell);
            if (cell.empty()) fast_window(lvl, iniX, iniY, maxX, maxY, minTh, cell);
            for (Cand& c : cell) {
                c.x += j * wCell;
                c.y += i * hCell;
                out.push_back(c);
            }
        }
    }
}

namespace {
struct Node {
    int ulx, uly, brx, bry;          // UL and BR corners (UR.x == BR.x, BL.y == BR.y)
    std::vector<int> keys;           // indices into the candidate array, in vKeys order
    bool noMore = false;
    long seq = 0;                    // creation order: stands in for the heap address in the size tie-break
    std::list<Node>::iterator lit;
};

// ExtractorNode::DivideNode, src/ORBextractor.cc:479-535
void divide(const Node& n, const std::vector<Cand>& pts, Node c[4]) {
    const int halfX = (int)std::ceil((float)(n.brx - n.ulx) / 2);
    const int halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
    const int mx = n.ulx + halfX, my = n.uly + halfY;
    c[0].ulx = n.ulx; c[0].uly = n.uly; c[0].brx = mx;    c[0].bry = my;
    c[1].ulx = mx;    c[1].uly = n.uly; c[1].brx = n.brx; c[1].bry = my;
    c[2].ulx = n.ulx; c[2].uly = my;    c[2].brx = mx;    c[2].bry = n.bry;
    c[3].ulx = mx;    c[3].uly = my;    c[3].brx = n.brx; c[3].bry = n.bry;
    for (int k : n.keys) {
        const Cand& p = pts[k];
        if (p.x < (float)mx) {
            if (p.y < (float)my) c[0].keys.push_back(k);
            else c[2].keys.push_back(k);
        } else if (p.y < (float)my) c[1].keys.push_back(k);
        else c[3].keys.push_back(k);
    }
    for (int q = 0; q < 4; ++q) c[q].noMore = (c[q].keys.size() == 1);
}
}  // namespace

// ORBextractor::DistributeOctTree, src/ORBextractor.cc:537-761.  Declared rule for the reference's
// (count, heap address) sort key: equal counts -> the more recently created node is expanded first.
void distribute_octree(const std::vector<Cand>& in, int minX, int maxX, int minY, int maxY, int N,
                       std::vector<Cand>& out) {
    out.clear();
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return;
    const float hX = (float)(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; ++i) {
        Node n;
        n.ulx = (int)(hX * (float)i);
        n.brx = (int)(hX * (float)(i + 1));
        n.uly = 0;
        n.bry = maxY - minY;
        n.seq = seq++;
        nodes.push_back(n);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < in.size(); ++i) ini[(size_t)(in[i].x / hX)]->keys.push_back((int)i);
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->noMore = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    std::vector<std::pair<int, Node*>> sizeAndNode;
    auto push_children = [&](Node c[4], int* nToExpand) {
        for (int q = 0; q < 4; ++q) {
            if (c[q].keys.empty()) continue;
            c[q].seq = seq++;
            nodes.push_front(c[q]);
            if (c[q].keys.size() > 1) {
                if (nToExpand) ++*nToExpand;
                sizeAndNode.push_back({(int)c[q].keys.size(), &nodes.front()});
                nodes.front().lit = nodes.begin();
            }
        }
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        int nToExpand = 0;
        sizeAndNode.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->noMore) { ++it; continue; }
            Node c[4];
            divide(*it, in, c);
            push_children(c, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                std::vector<std::pair<int, Node*>> prev = sizeAndNode;
                sizeAndNode.clear();
                std::sort(prev.begin(), prev.end(), [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
                    if (a.first != b.first) return a.first < b.first;
                    return a.second->seq < b.second->seq;
                });
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    Node c[4];
                    divide(*prev[j].second, in, c);
                    push_children(c, nullptr);
                    nodes.erase(prev[j].second->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    out.reserve(nodes.size());
    for (const Node& n : nodes) {
        int best = n.keys[0];
        float maxR = in[best].resp;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (in[n.keys[k]].resp > maxR) { best = n.keys[k]; maxR = in[best].resp; }
        out.push_back(in[best]);
    }
}

// IC_Angle, src/ORBextractor.cc:75-102
float ic_angle(const Img8& lvl, int x, int y, const int* umax) {
    int m01 = 0, m10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m10 += u * lvl.at(y, x + u);
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = lvl.at(y + v, x + u), vm = lvl.at(y - v, x + u);
            vsum += vp - vm;
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor, src/ORBextractor.cc:105-145
void orb_descriptor(const Img8& blur, int x, int y, float angleDeg, uint8_t* desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = angleDeg * factorPI;
    // Declared oracle rule: cosf/sinf are taken as correctly rounded, i.e. the double libm result rounded to float.
    // glibc's cosf differs from that by 1 ulp for ~0.9 % of arguments and its ifunc variant depends on the host CPU,
    // which would make the oracle machine-dependent.
    float a = ref_cosf(angle), b = ref_sinf(angle);   // (float)cos(angle) on a float under `using namespace std`, :112
    const int* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int* q = pat + 4 * k;
            int r0 = cv_roundf((float)q[0] * b + (float)q[1] * a), c0 = cv_roundf((float)q[0] * a - (float)q[1] * b);
            int r1 = cv_roundf((float)q[2] * b + (float)q[3] * a), c1 = cv_roundf((float)q[2] * a - (float)q[3] * b);
            int t0 = blur.at(y + r0, x + c0), t1 = blur.at(y + r1, x + c1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ORBextractor::operator(), ComputePyramid, ComputeKeyPointsOctTree: src/ORBextractor.cc:1068-1177, 763-878
int orb_extract(const OrbConfig& c, const OrbTables& t, const uint8_t* img, int w, int h, int stride, int lap0,
                int lap1, OrbState& st) {
    st.valid = false;
    if (!img || w <= 0 || h <= 0) return -1;
    const int nl = c.nlevels;
    st.pyr.assign(nl, Img8());
    st.blur.assign(nl, Img8());
    st.cands.assign(nl, {});
    // ComputePyramid (:1152-1177); the 19-px REFLECT_101 border is never read by the path and is not materialised
    st.pyr[0] = Img8(w, h);
    for (int y = 0; y < h; ++y) std::memcpy(st.pyr[0].row(y), img + (size_t)y * stride, w);
    for (int l = 1; l < nl; ++l) {
        float s = t.invScale[l];
        int lw = cv_roundf((float)w * s), lh = cv_roundf((float)h * s);
        resize_linear_u8(st.pyr[l - 1], st.pyr[l], lw, lh);
    }
    std::vector<std::vector<plf_keypoint>> all(nl);
    for (int l = 0; l < nl; ++l) {
        const Img8& lvl = st.pyr[l];
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = lvl.w - EDGE_THRESHOLD + 3, maxBY = lvl.h - EDGE_THRESHOLD + 3;
        level_candidates(lvl, c.iniThFAST, c.minThFAST, st.cands[l]);
        std::vector<Cand> kept;
        distribute_octree(st.cands[l], minBX, maxBX, minBY, maxBY, t.nPerLevel[l], kept);
        const int scaledPatch = (int)(PATCH_SIZE * t.scale[l]);
        for (const Cand& k : kept) {
            plf_keypoint kp;
            kp.x = k.x + minBX;
            kp.y = k.y + minBY;
            kp.size = (float)scaledPatch;
            kp.response = k.resp;
            kp.octave = l;
            kp.class_id = -1;
            kp.angle = ic_angle(lvl, cv_roundf(kp.x), cv_roundf(kp.y), t.umax);
            all[l].push_back(kp);
        }
    }
    int total = 0;
    for (int l = 0; l < nl; ++l) total += (int)all[l].size();
    st.kps.assign(total, plf_keypoint());
    st.desc.assign((size_t)total * 32, 0);
    int mono = 0, stereo = total - 1;
    for (int l = 0; l < nl; ++l) {
        if (all[l].empty()) continue;
        gaussian_blur_u8(st.pyr[l], st.blur[l], TAPS_ORB7, 7);
        const float scale = t.scale[l];
        for (plf_keypoint kp : all[l]) {
            uint8_t d[32];
            orb_descriptor(st.blur[l], cv_roundf(kp.x), cv_roundf(kp.y), kp.angle, d);
            if (l != 0) { kp.x *= scale; kp.y *= scale; }
            int dst;
            if (kp.x >= lap0 && kp.x <= lap1) dst = stereo--;
            else dst = mono++;
            st.kps[dst] = kp;
            std::memcpy(&st.desc[(size_t)dst * 32], d, 32);
        }
    }
    st.monoIndex = mono;
    st.valid = true;
    return mono;
}

}  // namespace plfo
