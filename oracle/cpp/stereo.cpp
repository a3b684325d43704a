// TEST INFRASTRUCTURE ONLY — CPU oracle, stereo point matching (restates Frame::ComputeStereoMatches,
// src/Frame.cc:976-1154, and ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2495-2511).
#include "stereo.h"
#include <algorithm>
#include <climits>
#include <cmath>

namespace plfo {

int hamming256(const uint8_t* a, const uint8_t* b) { return plf_hamming256(a, b); }

void stereo_match_points(const OrbTables& t, const OrbState& L, const OrbState& R, float mbf, float fx,
                         std::vector<float>& uRight, std::vector<float>& depth) {
    const int N = (int)L.kps.size();
    uRight.assign(N, -1.0f);
    depth.assign(N, -1.0f);
    if (N == 0 || R.kps.empty()) return;
    const int thOrbDist = (100 + 50) / 2;   // (TH_HIGH+TH_LOW)/2, src/ORBmatcher.cc:36-37
    const int nRows = L.pyr[0].h;
    std::vector<std::vector<int>> rowIdx(nRows);
    const int Nr = (int)R.kps.size();
    for (int iR = 0; iR < Nr; ++iR) {
        const plf_keypoint& kp = R.kps[iR];
        const float r = 2.0f * t.scale[kp.octave];
        const int maxr = (int)std::ceil(kp.y + r);
        const int minr = (int)std::floor(kp.y - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);
    }
    // Oracle rule: mb := mbf/fx before matching (the reference reads mb uninitialised, Frame.cc:1006 vs :197)
    const float mb = mbf / fx;
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < N; ++iL) {
        const plf_keypoint& kpL = L.kps[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& cand = rowIdx[row];
        if (cand.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;
        int bestIdxR = 0;
        const uint8_t* dL = &L.desc[(size_t)iL * 32];
        for (int iR : cand) {
            const plf_keypoint& kpR = R.kps[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256(dL, &R.desc[(size_t)iR * 32]);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist >= thOrbDist) continue;
        const float uR0 = R.kps[bestIdxR].x;
        const float sf = t.invScale[kpL.octave];
        const float scaleduL = std::round(kpL.x * sf);
        const float scaledvL = std::round(kpL.y * sf);
        const float scaleduR0 = std::round(uR0 * sf);
        const int w = 5, Lw = 5;
        const Img8& imL = L.pyr[kpL.octave];
        const Img8& imR = R.pyr[kpL.octave];
        const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
        const float iniu = scaleduR0 + Lw - w, endu = scaleduR0 + Lw + w + 1;
        if (iniu < 0 || endu >= imR.w) continue;
        // the reference would throw inside cv::Mat::rowRange here; the oracle skips (cannot happen for kps >= 19 px
        // from the border)
        if (cv - w < 0 || cv + w >= imL.h || cu - w < 0 || cu + w >= imL.w || cr - Lw - w < 0) continue;
        int bestSad = INT_MAX, bestInc = 0;
        float dists[2 * 5 + 1];
        const int cL = imL.at(cv, cu);
        for (int inc = -Lw; inc <= Lw; ++inc) {
            const int cR = imR.at(cv, cr + inc);
            int sad = 0;
            for (int dy = -w; dy <= w; ++dy)
                for (int dx = -w; dx <= w; ++dx) {
                    int a = imL.at(cv + dy, cu + dx) - cL;
                    int b = imR.at(cv + dy, cr + inc + dx) - cR;
                    sad += std::abs(a - b);
                }
            float dist = (float)sad;
            if (dist < bestSad) { bestSad = (int)dist; bestInc = inc; }
            dists[Lw + inc] = dist;
        }
        if (bestInc == -Lw || bestInc == Lw) continue;
        const float d1 = dists[Lw + bestInc - 1], d2 = dists[Lw + bestInc], d3 = dists[Lw + bestInc + 1];
        const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = t.scale[kpL.octave] * ((float)scaleduR0 + (float)bestInc + deltaR);
        float disparity = uL - bestuR;
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = 0.01; bestuR = (float)((double)uL - 0.01); }   // double literals, Frame.cc:1130-1131
            depth[iL] = mbf / disparity;
            uRight[iL] = bestuR;
            distIdx.push_back({bestSad, iL});
        }
    }
    if (distIdx.empty()) return;   // oracle rule: the reference indexes an empty vector here (Frame.cc:1141)
    std::sort(distIdx.begin(), distIdx.end());
    const float median = (float)distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
        if (distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
    }
}

}  // namespace plfo
