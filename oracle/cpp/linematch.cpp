// TEST INFRASTRUCTURE ONLY — CPU oracle, line matching.  Restates matchNNR / match / matchGrid(lines)
// (src/LineMatcher.cpp:139-159,201-229,317-396), GridStructure (src/gridStructure.cpp:33-76), LineIterator
// (src/LineIterator.cpp:34-77) and Frame::ComputeStereoMatches_Lines (src/Frame.cc:1156-1307).
#include "linematch.h"
#include "stereo.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <list>
#include <set>

namespace plfo {

// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): ties ordered by ascending train index.
int match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12) {
    int matches = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    if (n2 < 2) return 0;   // oracle rule: the reference indexes matches_[idx][1] out of range (LineMatcher.cpp:152)
    for (int i = 0; i < n1; ++i) {
        int b0 = INT_MAX, b1 = INT_MAX, i0 = -1;
        for (int j = 0; j < n2; ++j) {
            int d = hamming256(d1 + (size_t)i * 32, d2 + (size_t)j * 32);
            if (d < b0) { b1 = b0; b0 = d; i0 = j; }
            else if (d < b1) b1 = d;
        }
        if ((float)b0 < (float)b1 * nnr) { m12[i] = i0; ++matches; }
    }
    return matches;
}

int match_lr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int* m12) {
    int matches = match_nnr(d1, n1, d2, n2, nnr, m12);
    if (!best_lr) return matches;
    std::vector<int> m21(std::max(n2, 1));
    match_nnr(d2, n2, d1, n1, nnr, m21.data());
    for (int i1 = 0; i1 < n1; ++i1) {
        int& i2 = m12[i1];
        if (i2 >= 0 && m21[i2] != i1) { i2 = -1; --matches; }
    }
    return matches;
}

namespace {
// src/LineIterator.cpp:34-77
struct LineIt {
    double x1, y1, x2, y2, dx, dy, error;
    bool steep;
    int x, y, maxX, ystep;
    LineIt(double x1_, double y1_, double x2_, double y2_)
        : x1(x1_), y1(y1_), x2(x2_), y2(y2_), steep(std::abs(y2_ - y1_) > std::abs(x2_ - x1_)) {
        if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
        if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
        dx = x2 - x1;
        dy = std::abs(y2 - y1);
        error = dx / 2.0;
        ystep = (y1 < y2) ? 1 : -1;
        x = (int)x1;
        y = (int)y1;
        maxX = (int)x2;
    }
    bool next(int& px, int& py) {
        if (x > maxX) return false;
        if (steep) { px = y; py = x; } else { px = x; py = y; }
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
        ++x;
        return true;
    }
};
}  // namespace

void stereo_match_lines(const LineMatchConfig& c, int W, int H, const std::vector<plf_keyline>& klL,
                        const std::vector<uint8_t>& dL, const std::vector<plf_keyline>& klR,
                        const std::vector<uint8_t>& dR, std::vector<float>& disp, std::vector<double>& le,
                        std::vector<int>& m12) {
    const int nL = (int)klL.size(), nR = (int)klR.size();
    disp.assign((size_t)nL * 2, -1.f);
    le.assign((size_t)nL * 3, 0.0);
    m12.assign(nL, -1);
    if (nL == 0 || nR == 0) return;
    const int ROWS = 48, COLS = 64;     // FRAME_GRID_ROWS/COLS, include/Frame.h:59-60
    const double inv_w = COLS / (double)W, inv_h = ROWS / (double)H;   // Frame.cc:109-110
    // coords: line_2d holds ints -> truncation (include/LineMatcher.h:45-46, Frame.cc:1180-1182)
    std::vector<int> sx(nL), sy(nL), ex(nL), ey(nL);
    for (int i = 0; i < nL; ++i) {
        sx[i] = (int)(klL[i].startPointX * inv_w);
        sy[i] = (int)(klL[i].startPointY * inv_h);
        ex[i] = (int)(klL[i].endPointX * inv_w);
        ey[i] = (int)(klL[i].endPointY * inv_h);
    }
    std::vector<std::vector<std::list<int>>> grid(COLS, std::vector<std::list<int>>(ROWS));
    std::vector<std::pair<double, double>> dir(nR);
    for (int idx = 0; idx < nR; ++idx) {
        const plf_keyline& kl = klR[idx];
        double vx = (kl.endPointX - kl.startPointX) * inv_w, vy = (kl.endPointY - kl.startPointY) * inv_h;
        double mag = std::sqrt(vx * vx + vy * vy);
        dir[idx] = {vx / mag, vy / mag};
        LineIt it(kl.startPointX * inv_w, kl.startPointY * inv_h, kl.endPointX * inv_w, kl.endPointY * inv_h);
        int px, py;
        while (it.next(px, py))
            if (px >= 0 && px < COLS && py >= 0 && py < ROWS) grid[px][py].push_back(idx);
    }
    auto grid_get = [&](int x, int y, std::set<int>& out) {
        int min_x = std::max(0, x - c.matching_s_ws), max_x = std::min(COLS, x + 0 + 1);
        int min_y = std::max(0, y - 0), max_y = std::min(ROWS, y + 0 + 1);
        for (int x_ = min_x; x_ < max_x; ++x_)
            for (int y_ = min_y; y_ < max_y; ++y_) out.insert(grid[x_][y_].begin(), grid[x_][y_].end());
    };
    // matchGrid(lines), LineMatcher.cpp:317-396.  Candidate iteration order does not affect the result
    // (SURVEY §7 hard part 6), so an ordered set is used.
    std::vector<int> m21(nR, -1), dists(nR, INT_MAX);
    for (int i1 = 0; i1 < nL; ++i1) {
        int best_d = INT_MAX, best_d2 = INT_MAX, best_idx = -1;
        double vx = ex[i1] - sx[i1], vy = ey[i1] - sy[i1];
        double mag = std::sqrt(vx * vx + vy * vy);
        vx /= mag;
        vy /= mag;
        std::set<int> cand;
        grid_get(sx[i1], sy[i1], cand);
        grid_get(ex[i1], ey[i1], cand);
        if (cand.empty()) continue;
        for (int i2 : cand) {
            if (i2 < 0 || i2 >= nR) continue;
            if (std::abs(vx * dir[i2].first + vy * dir[i2].second) < c.line_sim_th) continue;
            const int d = hamming256(&dL[(size_t)i1 * 32], &dR[(size_t)i2 * 32]);
            if (c.best_lr_matches) {
                if (d < dists[i2]) { dists[i2] = d; m21[i2] = i1; }
                else continue;
            }
            if (d < best_d) { best_d2 = best_d; best_d = d; best_idx = i2; }
            else if (d < best_d2) best_d2 = d;
        }
        if (best_d < best_d2 * c.min_ratio_12_l) m12[i1] = best_idx;
    }
    if (c.best_lr_matches)
        for (int i1 = 0; i1 < nL; ++i1) {
            int& i2 = m12[i1];
            if (i2 >= 0 && m21[i2] != i1) i2 = -1;
        }
    // Frame.cc:1212-1252 with lineSegmentOverlapStereo (:1261-1295) and filterLineSegmentDisparity (:1297-1307)
    for (int i1 = 0; i1 < nL; ++i1) {
        const int i2 = m12[i1];
        if (i2 < 0) continue;
        const double spl[2] = {klL[i1].startPointX, klL[i1].startPointY};
        const double epl[2] = {klL[i1].endPointX, klL[i1].endPointY};
        // le_l = sp_l x ep_l (homogeneous, z = 1), normalised by its first two components
        double l0 = spl[1] * 1.0 - 1.0 * epl[1];
        double l1 = 1.0 * epl[0] - spl[0] * 1.0;
        double l2 = spl[0] * epl[1] - spl[1] * epl[0];
        const double nrm = std::sqrt(l0 * l0 + l1 * l1);
        l0 = l0 / nrm; l1 = l1 / nrm; l2 = l2 / nrm;
        double spr[2] = {klR[i2].startPointX, klR[i2].startPointY};
        double epr[2] = {klR[i2].endPointX, klR[i2].endPointY};
        // overlap
        double overlap = 1.f;
        {
            const double spl_obs = spl[1], epl_obs = epl[1], spl_proj = spr[1], epl_proj = epr[1];
            if (std::fabs(epl_obs - spl_obs) > c.line_horiz_th) {
                double sln = std::min(spl_obs, epl_obs), eln = std::max(spl_obs, epl_obs);
                double spn = std::min(spl_proj, epl_proj), epn = std::max(spl_proj, epl_proj);
                double length = eln - spn;
                if ((epn < sln) || (spn > eln)) overlap = 0.f;
                else if ((epn > eln) && (spn < sln)) overlap = eln - sln;
                else overlap = std::min(eln, epn) - std::max(sln, spn);
                if (length > 0.01f) overlap = overlap / length;
                else overlap = 0.f;
                if (overlap > 1.f) overlap = 1.f;
            }
        }
        // Frame.cc:1228-1229: sp_r is overwritten first and its NEW value feeds the ep_r expression
        const double nsx = (spr[0] * (spl[1] - epr[1]) + epr[0] * (spr[1] - spl[1])) / (spr[1] - epr[1]);
        spr[0] = nsx;
        spr[1] = spl[1];
        const double nex = (spr[0] * (epl[1] - epr[1]) + epr[0] * (spr[1] - epl[1])) / (spr[1] - epr[1]);
        epr[0] = nex;
        epr[1] = epl[1];
        double disp_s = spl[0] - spr[0], disp_e = epl[0] - epr[0];
        if (std::min(disp_s, disp_e) / std::max(disp_s, disp_e) < c.ls_min_disp_ratio) { disp_s = -1.0; disp_e = -1.0; }
        if (disp_s >= c.min_disp && disp_e >= c.min_disp && std::abs(spl[1] - epl[1]) > c.line_horiz_th &&
            std::abs(spr[1] - epr[1]) > c.line_horiz_th && overlap > c.stereo_overlap_th) {
            disp[(size_t)i1 * 2] = (float)disp_s;
            disp[(size_t)i1 * 2 + 1] = (float)disp_e;
            le[(size_t)i1 * 3] = l0;
            le[(size_t)i1 * 3 + 1] = l1;
            le[(size_t)i1 * 3 + 2] = l2;
        }
    }
}

}  // namespace plfo
