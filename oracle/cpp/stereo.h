// TEST INFRASTRUCTURE ONLY — CPU oracle, stereo point matching.
#pragma once
#include "orb.h"
namespace plfo {
int hamming256(const uint8_t* a, const uint8_t* b);
void stereo_match_points(const OrbTables& t, const OrbState& L, const OrbState& R, float mbf, float fx,
                         std::vector<float>& uRight, std::vector<float>& depth);
}
