// TEST INFRASTRUCTURE ONLY — CPU oracle, LBD band descriptor.  Restates BinaryDescriptor::computeSobel
// (binary_descriptor_custom.cpp:350-398), ::computeLBD (:1026-1372), ::binaryConversion (:401-412, table :74-107)
// and the Gaussian weight tables of the constructor (:217-259).  Paths relative to
// /root/reference/Thirdparty/line_descriptor/src/.
#include "lsd.h"
#include <cmath>

namespace plfo {

static const int kBands = 9, kBandW = 7;
static const int kComb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                 {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                 {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

void lbd_compute(const Img8& img, const std::vector<plf_keyline>& kls, std::vector<float>& lbd72,
                 std::vector<uint8_t>& desc) {
    const int n = (int)kls.size();
    lbd72.assign((size_t)n * 72, 0.f);
    desc.assign((size_t)n * 32, 0);
    if (n == 0) return;
    Img8 g;
    gaussian_blur_u8(img, g, TAPS_LBD5, 5);
    Img16 dxI, dyI;
    sobel3_16s(g, dxI, dyI);
    // weight tables (:227-258); the integer divisions are the reference's
    double gaussL[kBandW * 3], gaussG[kBands * kBandW];
    {
        double u = (kBandW * 3 - 1) / 2;
        double sigma = (kBandW * 2 + 1) / 2;
        double inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kBandW * 3; ++i) { double d = i - u; gaussL[i] = std::exp(d * d * inv); }
        u = (kBands * kBandW - 1) / 2;
        sigma = u;
        inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < kBands * kBandW; ++i) { double d = i - u; gaussG[i] = std::exp(d * d * inv); }
    }
    const short heightOfLSP = kBandW * kBands;
    const short halfHeight = (heightOfLSP - 1) / 2;
    const short realWidth = (short)img.w, imageWidth = realWidth - 1, imageHeight = (short)(img.h - 1);
    for (int li = 0; li < n; ++li) {
        const plf_keyline& kl = kls[li];
        float band[8][kBands] = {};   // pgdL ngdL pgdL2 ngdL2 pgdO ngdO pgdO2 ngdO2
        const short lengthOfLSP = (short)kl.numOfPixels;
        const short halfWidth = (lengthOfLSP - 1) / 2;
        const float midX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
        const float midY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
        float dL[2], dO[2];
        dL[0] = ref_cosf(kl.angle);   // cos( pSingleLine->direction ) on a float, :1134
        dL[1] = ref_sinf(kl.angle);
        dO[0] = -dL[1];
        dO[1] = dL[0];
        float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + midX;
        float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + midY;
        for (short hID = 0; hID < heightOfLSP; ++hID) {
            float sCorX = sCorX0, sCorY = sCorY0;
            float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
            for (short wID = 0; wID < lengthOfLSP; ++wID) {
                short t = (short)std::round(sCorX);
                short xCor = (t < 0) ? 0 : (t > imageWidth) ? imageWidth : t;
                t = (short)std::round(sCorY);
                short yCor = (t < 0) ? 0 : (t > imageHeight) ? imageHeight : t;
                short dx = dxI.d[(size_t)yCor * realWidth + xCor];
                short dy = dyI.d[(size_t)yCor * realWidth + xCor];
                float gDL = dx * dL[0] + dy * dL[1];
                float gDO = dx * dO[0] + dy * dO[1];
                if (gDL > 0) pgdL += gDL; else ngdL -= gDL;
                if (gDO > 0) pgdO += gDO; else ngdO -= gDO;
                sCorX += dL[0];
                sCorY += dL[1];
            }
            sCorX0 -= dL[1];
            sCorY0 += dL[0];
            float coef = (float)gaussG[hID];
            pgdL = coef * pgdL; ngdL = coef * ngdL;
            float pgdL2 = pgdL * pgdL, ngdL2 = ngdL * ngdL;
            pgdO = coef * pgdO; ngdO = coef * ngdO;
            float pgdO2 = pgdO * pgdO, ngdO2 = ngdO * ngdO;
            auto add = [&](int b, float cf) {
                band[0][b] += cf * pgdL;  band[1][b] += cf * ngdL;
                band[2][b] += cf * cf * pgdL2; band[3][b] += cf * cf * ngdL2;
                band[4][b] += cf * pgdO;  band[5][b] += cf * ngdO;
                band[6][b] += cf * cf * pgdO2; band[7][b] += cf * cf * ngdO2;
            };
            short bandID = (short)(hID / kBandW);
            add(bandID, (float)gaussL[hID % kBandW + kBandW]);
            bandID--;
            if (bandID >= 0) add(bandID, (float)gaussL[hID % kBandW + 2 * kBandW]);
            bandID = bandID + 2;
            if (bandID < kBands) add(bandID, (float)gaussL[hID % kBandW]);
        }
        float* des = &lbd72[(size_t)li * 72];
        const float invN2 = (float)(1.0 / (kBandW * 2.0)), invN3 = (float)(1.0 / (kBandW * 3.0));
        for (int b = 0; b < kBands; ++b) {
            float invN = (b == 0 || b == kBands - 1) ? invN2 : invN3;
            float temp;
            temp = band[0][b] * invN; des[b * 8 + 0] = temp; des[b * 8 + 4] = std::sqrt(band[2][b] * invN - temp * temp);
            temp = band[1][b] * invN; des[b * 8 + 1] = temp; des[b * 8 + 5] = std::sqrt(band[3][b] * invN - temp * temp);
            temp = band[4][b] * invN; des[b * 8 + 2] = temp; des[b * 8 + 6] = std::sqrt(band[6][b] * invN - temp * temp);
            temp = band[5][b] * invN; des[b * 8 + 3] = temp; des[b * 8 + 7] = std::sqrt(band[7][b] * invN - temp * temp);
        }
        float tempM = 0, tempS = 0;
        for (int b = 0; b < kBands; ++b) {
            for (int k = 0; k < 4; ++k) tempM += des[b * 8 + k] * des[b * 8 + k];
            for (int k = 4; k < 8; ++k) tempS += des[b * 8 + k] * des[b * 8 + k];
        }
        tempM = 1 / std::sqrt(tempM);
        tempS = 1 / std::sqrt(tempS);
        for (int b = 0; b < kBands; ++b) {
            for (int k = 0; k < 4; ++k) des[b * 8 + k] = des[b * 8 + k] * tempM;
            for (int k = 4; k < 8; ++k) des[b * 8 + k] = des[b * 8 + k] * tempS;
        }
        for (int i = 0; i < 72; ++i)
            if (des[i] > 0.4) des[i] = (float)0.4;
        float temp = 0;
        for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
        temp = 1 / std::sqrt(temp);
        for (int i = 0; i < 72; ++i) des[i] = des[i] * temp;
        uint8_t* out = &desc[(size_t)li * 32];
        for (int cidx = 0; cidx < 32; ++cidx) {
            const float* f1 = &des[8 * kComb[cidx][0]];
            const float* f2 = &des[8 * kComb[cidx][1]];
            uint8_t r = 0;
            for (int i = 0; i < 8; ++i)
                if (f1[i] > f2[i]) r += (uint8_t)(1 << i);
            out[cidx] = r;
        }
    }
}

}  // namespace plfo
