// TEST INFRASTRUCTURE ONLY — CPU oracle, line matching.
#pragma once
#include "lsd.h"
namespace plfo {
struct LineMatchConfig {
    int best_lr_matches = 1, matching_s_ws = 10;
    double min_ratio_12_l = 0.9, line_sim_th = 0.75, min_disp = 1.0, line_horiz_th = 0.1, stereo_overlap_th = 0.75,
           ls_min_disp_ratio = 0.7;
};
int match_nnr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int* m12);
int match_lr(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int* m12);
void stereo_match_lines(const LineMatchConfig& c, int W, int H, const std::vector<plf_keyline>& klL,
                        const std::vector<uint8_t>& dL, const std::vector<plf_keyline>& klR,
                        const std::vector<uint8_t>& dR, std::vector<float>& disp, std::vector<double>& le,
                        std::vector<int>& m12);
}
