// Shared device helpers: REFLECT_101 indexing, the separable fixed-point Gaussian tile and cv::fastAtan2.
#pragma once
#include <cstdint>

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

// K1b  separable 8-bit Gaussian with integer taps summing to 256 (cv::GaussianBlur fixed-point path), REFLECT_101.
// Generic over ksize <= 7.  Tile 32x32 outputs per 256-thread block, staged through shared memory: the (32+6)^2
// source window is read once, the horizontal pass is kept as u16 in shared memory.
struct BlurJob { const uint8_t* src; uint8_t* dst; int w, h, sp, dp; };

template <int K>
__device__ __forceinline__ void blur_tile(const BlurJob& j, const int* taps, int tx0, int ty0) {
    constexpr int R = K / 2, TW = 32, TH = 32, IW = TW + 2 * R, IH = TH + 2 * R;
    __shared__ uint8_t s_in[IH][IW + 2];
    __shared__ uint16_t s_h[IH][TW];
    const int tx = threadIdx.x, ty = threadIdx.y;
    // (32+2R)^2 source window, REFLECT_101 at the image border; block is 32x8 threads
    int gx0 = reflect101(tx0 + tx - R, j.w);
    gx0 = min(max(gx0, 0), j.w - 1);
    int gx1 = reflect101(tx0 + 32 + tx - R, j.w);
    gx1 = min(max(gx1, 0), j.w - 1);
#pragma unroll
    for (int iy = ty; iy < IH; iy += 8) {
        int gy = reflect101(ty0 + iy - R, j.h);
        gy = min(max(gy, 0), j.h - 1);
        const uint8_t* row = j.src + (size_t)gy * j.sp;
        s_in[iy][tx] = row[gx0];
        if (tx < 2 * R) s_in[iy][32 + tx] = row[gx1];
    }
    __syncthreads();
#pragma unroll
    for (int iy = ty; iy < IH; iy += 8) {
        int acc = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) acc += taps[k] * s_in[iy][tx + k];
        s_h[iy][tx] = (uint16_t)acc;
    }
    __syncthreads();
    const int gx = tx0 + tx;
    if (gx < j.w) {
#pragma unroll
        for (int oy = ty; oy < TH; oy += 8) {
            const int gy = ty0 + oy;
            if (gy < j.h) {
                unsigned acc = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) acc += (unsigned)taps[k] * s_h[oy + k][tx];
                j.dst[(size_t)gy * j.dp + gx] = (uint8_t)((acc + 32768u) >> 16);
            }
        }
    }
}


__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // cv::fastAtan2: degree-7 polynomial, float, no FMA
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

