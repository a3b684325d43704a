// Stereo point matching on sm_100a.  Replaces Frame::ComputeStereoMatches (reference src/Frame.cc:976-1154) and its
// inner ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2495-2511): row-band candidate search with 256-bit Hamming
// distance (__popc over 8 words), 11x11 SAD sliding window at the keypoint's pyramid level, parabola sub-pixel fit,
// then the median-based cull.  Bit-exact against oracle/cpp/stereo.cpp.
#include "plf_ctx.cuh"

namespace {

// One warp per left keypoint.  Lanes stride over the right keypoints; the row-band test replaces vRowIndices
// (Frame.cc:985-1003): iR is a candidate of row (int)vL iff floor(yR - r) <= (int)vL <= ceil(yR + r), r = 2*scale[oct].
// The argmin keeps the lowest iR on ties (the reference scans candidates in ascending iR with a strict '<').
__global__ void __launch_bounds__(256) stereo_points_kernel(PlfGeom g, const uint8_t* pyr, const plf_keypoint* kp,
                                                            const uint8_t* desc, const int* nKp, float* uRight,
                                                            float* depth, int* sadOut, float mbf, float fx,
                                                            int slotFirst) {
    // The 8 warps of a block scan the same right keypoints: their row bands (floor(yR - r), ceil(yR + r)), x and octave are
    // formed once per block into shared memory, 1024 at a time, instead of once per (left, right) pair from the 28-byte
    // keypoint records.
    __shared__ int4 s_band[1024];
    const int slot = slotFirst + blockIdx.y;
    const int imgL = slot * 2, imgR = slot * 2 + 1;
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int N = nKp[imgL], Nr = nKp[imgR];
    float* uR = uRight + (size_t)slot * g.kpCap;
    float* dp = depth + (size_t)slot * g.kpCap;
    int* so = sadOut + (size_t)slot * g.kpCap;
    bool active = iL < N;                           // warp-uniform
    if (active && lane == 0) { uR[iL] = -1.f; dp[iL] = -1.f; so[iL] = -1; }
    const plf_keypoint kpL = kp[(size_t)imgL * g.kpCap + (active ? iL : 0)];
    const plf_keypoint* kR = kp + (size_t)imgR * g.kpCap;
    const uint4* dL4 = reinterpret_cast<const uint4*>(desc + ((size_t)imgL * g.kpCap + (active ? iL : 0)) * 32);
    const uint4 dl0 = dL4[0], dl1 = dL4[1];
    const uint4* dR4 = reinterpret_cast<const uint4*>(desc + (size_t)imgR * g.kpCap * 32);
    // mb := mbf/fx (oracle rule), minZ = mb, maxD = mbf/minZ  (Frame.cc:1006-1008)
    const float mb = __fdiv_rn(mbf, fx);
    const float maxD = __fdiv_rn(mbf, mb), minD = 0.f;
    const float uL = kpL.x, vL = kpL.y;
    const int row = (int)vL;
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    if (maxU < 0) active = false;
    int best = 100, bestIdx = 0x7fffffff;   // TH_HIGH
    const int levelL = kpL.octave;
    for (int c0 = 0; c0 < Nr; c0 += 1024) {
        const int cnt = min(1024, Nr - c0);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += 256) {
            const plf_keypoint k = kR[c0 + j];
            const float r = __fmul_rn(2.0f, g.lv[k.octave].scale);
            const int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));
            s_band[j] = make_int4(minr, maxr, __float_as_int(k.x), k.octave);
        }
        __syncthreads();
        if (active)
            for (int j = lane; j < cnt; j += 32) {
                const int4 k = s_band[j];
                if (row < k.x || row > k.y) continue;
                if (k.w < levelL - 1 || k.w > levelL + 1) continue;
                const float kx = __int_as_float(k.z);
                if (!(kx >= minU && kx <= maxU)) continue;
                const int iR = c0 + j;
                const uint4 a = dR4[iR * 2], b = dR4[iR * 2 + 1];
                const int d = __popc(a.x ^ dl0.x) + __popc(a.y ^ dl0.y) + __popc(a.z ^ dl0.z) + __popc(a.w ^ dl0.w) +
                              __popc(b.x ^ dl1.x) + __popc(b.y ^ dl1.y) + __popc(b.z ^ dl1.z) + __popc(b.w ^ dl1.w);
                if (d < best) { best = d; bestIdx = iR; }   // ascending iR per lane: first minimum kept
            }
    }
    if (!active) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
        if (ob < best || (ob == best && oi < bestIdx)) { best = ob; bestIdx = oi; }
    }
    const int thOrbDist = (100 + 50) / 2;
    if (best >= thOrbDist || bestIdx == 0x7fffffff) return;
    // sub-pixel refinement by correlation (Frame.cc:1062-1137)
    const PlfLevel& lv = g.lv[levelL];
    const float uR0 = kR[bestIdx].x;
    const float sf = lv.invScale;
    const float scaleduL = roundf(__fmul_rn(kpL.x, sf)), scaledvL = roundf(__fmul_rn(kpL.y, sf));
    const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
    const int w = 5, L = 5;
    const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
    if (iniu < 0 || endu >= (float)lv.w) return;
    const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
    if (cv - w < 0 || cv + w >= lv.h || cu - w < 0 || cu + w >= lv.w || cr - L - w < 0) return;
    const uint8_t* imL = pyr + (size_t)imgL * g.pyrBytes + lv.off;
    const uint8_t* imR = pyr + (size_t)imgR * g.pyrBytes + lv.off;
    // 11 shifts x 121 pixels; lane j < 11 owns shift incR = j-5 (L1-cached 21x11 strip of the right image)
    int sad = 0x7fffffff;
    if (lane < 11) {
        const int inc = lane - L;
        const int cL = imL[(size_t)cv * lv.pitch + cu], cR = imR[(size_t)cv * lv.pitch + cr + inc];
        int s = 0;
        for (int dy = -w; dy <= w; ++dy) {
            const uint8_t* rl = imL + (size_t)(cv + dy) * lv.pitch + cu;
            const uint8_t* rr = imR + (size_t)(cv + dy) * lv.pitch + cr + inc;
#pragma unroll
            for (int dx = -w; dx <= w; ++dx) s += abs(((int)rl[dx] - cL) - ((int)rr[dx] - cR));
        }
        sad = s;
    }
    // first minimum in ascending incR (strict '<' in the reference loop)
    int bs = sad, bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int os = __shfl_xor_sync(0xffffffffu, bs, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (os < bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
    }
    const int bestInc = bi - L;
    if (bestInc == -L || bestInc == L) return;
    const float d1 = (float)__shfl_sync(0xffffffffu, sad, bi - 1);
    const float d2 = (float)bs;
    const float d3 = (float)__shfl_sync(0xffffffffu, sad, bi + 1);
    if (lane != 0) return;
    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
    if (deltaR < -1 || deltaR > 1) return;
    float bestuR = __fmul_rn(lv.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestInc), deltaR));
    float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= minD && disparity < maxD) {
        if (disparity <= 0) {
            disparity = 0.01f;
            bestuR = (float)((double)uL - 0.01);
        }
        dp[iL] = __fdiv_rn(mbf, disparity);
        uR[iL] = bestuR;
        so[iL] = bs;
    }
}

// Median cull (Frame.cc:1140-1153): sort (sad, iL) pairs, median = element [size/2], drop sad >= 1.5f*1.4f*median.
// One block per frame; the median is found by rank counting (pairs are distinct, so ranks are a permutation).
__global__ void __launch_bounds__(1024) stereo_cull_kernel(PlfGeom g, const int* nKp, float* uRight, float* depth,
                                                           const int* sadIn, int slotFirst) {
    extern __shared__ int s_sad[];
    __shared__ int s_cnt, s_median;
    const int slot = slotFirst + blockIdx.x;
    const int N = nKp[slot * 2];
    const int* sd = sadIn + (size_t)slot * g.kpCap;
    float* uR = uRight + (size_t)slot * g.kpCap;
    float* dp = depth + (size_t)slot * g.kpCap;
    if (threadIdx.x == 0) { s_cnt = 0; s_median = -1; }
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const int v = sd[i];
        s_sad[i] = v;
        local += v >= 0;
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    const int M = s_cnt;
    if (M == 0) return;   // oracle rule: nothing matched, nothing to cull
    const int target = M / 2;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const int v = s_sad[i];
        if (v < 0) continue;
        int rank = 0;
        for (int j = 0; j < N; ++j) {
            const int u = s_sad[j];
            rank += (u >= 0) && (u < v || (u == v && j < i));
        }
        if (rank == target) s_median = v;
    }
    __syncthreads();
    const float median = (float)s_median;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = threadIdx.x; i < N; i += 1024) {
        const int v = s_sad[i];
        if (v >= 0 && !((float)v < thDist)) { uR[i] = -1.f; dp[i] = -1.f; }
    }
}

}  // namespace

int plf_launch_stereo_points(plf_ctx* c, int slotFirst, int nSlots) {
    const PlfGeom& g = c->g;
    plf_mark(c, "stereo_points");
    stereo_points_kernel<<<dim3((g.kpCap + 7) / 8, nSlots), 256, 0, c->stream>>>(
        g, c->d_pyr, c->d_kp, c->d_desc, c->d_nKp, c->d_uRight, c->d_depth, c->d_sad, c->p.bf, c->p.fx, slotFirst);
    const size_t smem = g.kpCap * sizeof(int);          // bounded by build_geometry (<= 200 KB)
    static size_t s_granted[64] = {};
    if (plf_raise_smem_optin(s_granted, c->device, smem))
        cudaFuncSetAttribute(stereo_cull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stereo_cull_kernel<<<nSlots, 1024, smem, c->stream>>>(g, c->d_nKp, c->d_uRight, c->d_depth,
                                                                          c->d_sad, slotFirst);
    return 2;
}
