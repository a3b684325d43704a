// Shared device helpers: REFLECT_101 indexing, the separable fixed-point Gaussian tile and cv::fastAtan2.
#pragma once
#include <cstdint>

__device__ __forceinline__ int reflect101(int p, int n) {
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

// K1b  separable 8-bit Gaussian with integer taps summing to 256 (cv::GaussianBlur fixed-point path), REFLECT_101.
// 7 taps (shorter kernels are zero-padded, which leaves the arithmetic unchanged).  128x32 output tile per 256-thread
// block (round 2; 128x8 before: the 6 halo rows made the load and the horizontal pass process 1.75 rows per output row,
// now 1.19), every thread owns 4 horizontally adjacent pixels of 4 rows:
//   load   : the (128+8) x 38 source window as aligned 32-bit words, a warp per row (the row index is reflected once per
//            row; bytes are reflected one by one only in words that straddle the left / right image border)
//   pass 1 : horizontal taps with DP4A on funnel-shifted words; the u16 results of two consecutive rows are packed
//            into one word so that
//   pass 2 : the vertical taps are DP2A dot products over row pairs; out = (acc + 32768) >> 16.
struct BlurJob { const uint8_t* src; uint8_t* dst; int w, h, sp, dp; };

#define BL_TW 128
#define BL_TH PLF_BLUR_TH
#define BL_ROWS (BL_TH + 6)
#define BL_WORDS 34             // (128 + 8) bytes, window starts at x0 - 4

__device__ __forceinline__ void blur_tile7(const BlurJob& j, const int* t, int x0, int y0) {
    __shared__ __align__(16) unsigned s_in[BL_ROWS][BL_WORDS + 2];
    __shared__ __align__(16) unsigned s_h[BL_ROWS / 2][BL_TW];        // (row 2r, row 2r+1) u16 pairs
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int iy = ty; iy < BL_ROWS; iy += 8) {
        int gy = reflect101(y0 - 3 + iy, j.h);
        gy = min(max(gy, 0), j.h - 1);
        const uint8_t* row = j.src + (size_t)gy * j.sp;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int wx = tx + 32 * k;
            if (wx < BL_WORDS) {
                const int gx = x0 - 4 + wx * 4;
                unsigned v;
                if (gx >= 0 && gx + 3 < j.w) {
                    v = *reinterpret_cast<const unsigned*>(row + gx);
                } else {
                    v = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        int x = reflect101(gx + b, j.w);
                        x = min(max(x, 0), j.w - 1);
                        v |= (unsigned)row[x] << (8 * b);
                    }
                }
                s_in[iy][wx] = v;
            }
        }
    }
    __syncthreads();
    const unsigned T0 = (unsigned)t[0] | ((unsigned)t[1] << 8) | ((unsigned)t[2] << 16) | ((unsigned)t[3] << 24);
    const unsigned T1 = (unsigned)t[4] | ((unsigned)t[5] << 8) | ((unsigned)t[6] << 16);
    for (int pr = ty; pr < BL_ROWS / 2; pr += 8) {
        // rows 2*pr and 2*pr+1, my 4 pixels: output j needs window bytes 4*tx + j + 1 .. + 7
        unsigned hrow[2][4];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const unsigned* r = &s_in[2 * pr + rr][tx];
            const unsigned w0 = r[0], w1 = r[1], w2 = r[2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned lo = (q == 3) ? w1 : __funnelshift_r(w0, w1, 8 * (q + 1));
                const unsigned hi = (q == 3) ? w2 : __funnelshift_r(w1, w2, 8 * (q + 1));
                hrow[rr][q] = __dp4a(lo, T0, __dp4a(hi, T1, 0u));
            }
        }
        uint4 o;
        o.x = hrow[0][0] | (hrow[1][0] << 16);
        o.y = hrow[0][1] | (hrow[1][1] << 16);
        o.z = hrow[0][2] | (hrow[1][2] << 16);
        o.w = hrow[0][3] | (hrow[1][3] << 16);
        *reinterpret_cast<uint4*>(&s_h[pr][4 * tx]) = o;
    }
    __syncthreads();
    const int gx = x0 + 4 * tx;
    // an even output row r uses the pairs r/2 .. r/2+3 with taps (t0 t1)(t2 t3)(t4 t5)(t6 -), an odd one (- t0)(t1 t2)(t3 t4)(t5 t6);
    // ty and ty + 8k have the same parity, so the tap words are formed once
    const bool odd = ty & 1;
    // (DP2A multiplies the two u16 halves of its first operand by bytes 0 and 1 of the second)
    const unsigned tp[4] = {odd ? (unsigned)t[0] << 8 : (unsigned)t[0] | ((unsigned)t[1] << 8),
                            odd ? (unsigned)t[1] | ((unsigned)t[2] << 8) : (unsigned)t[2] | ((unsigned)t[3] << 8),
                            odd ? (unsigned)t[3] | ((unsigned)t[4] << 8) : (unsigned)t[4] | ((unsigned)t[5] << 8),
                            odd ? (unsigned)t[5] | ((unsigned)t[6] << 8) : (unsigned)t[6]};
#pragma unroll
    for (int k = 0; k < BL_TH / 8; ++k) {
        const int r = ty + 8 * k, gy = y0 + r;
        if (gx < j.w && gy < j.h) {
            // output row r uses staged rows r .. r+6
            unsigned acc[4] = {0u, 0u, 0u, 0u};
            const int p0 = r >> 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = *reinterpret_cast<const uint4*>(&s_h[p0 + q][4 * tx]);
                acc[0] = __dp2a_lo(v.x, tp[q], acc[0]);
                acc[1] = __dp2a_lo(v.y, tp[q], acc[1]);
                acc[2] = __dp2a_lo(v.z, tp[q], acc[2]);
                acc[3] = __dp2a_lo(v.w, tp[q], acc[3]);
            }
            const unsigned o0 = (acc[0] + 32768u) >> 16, o1 = (acc[1] + 32768u) >> 16;
            const unsigned o2 = (acc[2] + 32768u) >> 16, o3 = (acc[3] + 32768u) >> 16;
            uint8_t* d = j.dst + (size_t)gy * j.dp + gx;
            if (gx + 4 <= j.w && ((j.dp & 3) == 0)) {
                *reinterpret_cast<unsigned*>(d) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
            } else {
                d[0] = (uint8_t)o0;
                if (gx + 1 < j.w) d[1] = (uint8_t)o1;
                if (gx + 2 < j.w) d[2] = (uint8_t)o2;
                if (gx + 3 < j.w) d[3] = (uint8_t)o3;
            }
        }
    }
}


// Bilinear resize of 4 horizontally adjacent output pixels x..x+3 of one output row (x a multiple of 4) from the two
// source rows r0 / r1.  The 4 table entries come as two 16-byte loads (tables start on multiples of 4 entries), the
// source bytes as three aligned words per row starting at the word of the first tap (scale <= 2: the taps span at most
// 12 bytes); each tap pair is picked with one PRMT and weighted with one DP2A.  EXACT = cv::INTER_LINEAR_EXACT (Q8
// weights, one rounding), otherwise cv::INTER_LINEAR's 11-bit two-stage form.  Returns the 4 result bytes.
// The column part (table entries, tap offsets, weights, byte selectors) does not depend on the row: resize_prep forms it
// once, resize_apply uses it for every row a thread produces (4 rows per thread in the kernels).
struct PlfLin;
struct ResizeTaps { unsigned wts[4], sel[4]; int o[4]; int base; };
__device__ __forceinline__ void resize_prep(const void* linX4, int nValid, ResizeTaps& T) {
    const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(linX4)), t1 = __ldg(reinterpret_cast<const uint4*>(linX4) + 1);
    const unsigned e0[4] = {t0.x, t0.z, t1.x, t1.z}, e1[4] = {t0.y, t0.w, t1.y, t1.w};   // (ofs | a0<<16), (a1 | pad<<16)
    T.base = (int)(e0[0] & 0xFFFFu) & ~3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int jj = j < nValid ? j : 0;                                    // entries past the row end are padding
        T.o[j] = (int)(e0[jj] & 0xFFFFu) - T.base;                            // 0 .. 10
        T.wts[j] = __byte_perm(e0[jj], e1[jj], 0x5432);                       // a0 | a1 << 16
        T.sel[j] = (unsigned)(T.o[j] & 3) * 0x11u + 0x10u;                    // bytes (o&3), (o&3)+1 of {lo, hi}
    }
}
template <bool EXACT>
__device__ __forceinline__ unsigned resize_apply(const ResizeTaps& T, const uint8_t* r0, const uint8_t* r1, int ya0, int ya1) {
    const unsigned* w0 = reinterpret_cast<const unsigned*>(r0 + T.base);
    const unsigned* w1 = reinterpret_cast<const unsigned*>(r1 + T.base);
    const unsigned a0 = w0[0], a1 = w0[1], a2 = w0[2], b0 = w1[0], b1 = w1[1], b2 = w1[2];
    unsigned out = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int o = T.o[j];
        const unsigned loA = o < 4 ? a0 : (o < 8 ? a1 : a2), hiA = o < 4 ? a1 : a2;
        const unsigned loB = o < 4 ? b0 : (o < 8 ? b1 : b2), hiB = o < 4 ? b1 : b2;
        const int h0 = (int)__dp2a_lo(T.wts[j], __byte_perm(loA, hiA, T.sel[j]), 0u);
        const int h1 = (int)__dp2a_lo(T.wts[j], __byte_perm(loB, hiB, T.sel[j]), 0u);
        const int v = EXACT ? (h0 * ya0 + h1 * ya1 + 32768) >> 16
                            : (((ya0 * (h0 >> 4)) >> 16) + ((ya1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out |= (unsigned)(v & 0xFF) << (8 * j);
    }
    return out;
}
template <bool EXACT>
__device__ __forceinline__ unsigned resize_quad(const uint8_t* r0, const uint8_t* r1, const void* linX4, int nValid,
                                                int ya0, int ya1) {
    ResizeTaps T;
    resize_prep(linX4, nValid, T);
    return resize_apply<EXACT>(T, r0, r1, ya0, ya1);
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // cv::fastAtan2: degree-7 polynomial, float, no FMA
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    // branch-free form of "if (ax >= ay) c = ay / (ax + eps) else c = ax / (ay + eps), a = 90 - a": same operations
    const bool steep = !(ax >= ay);
    const float c = __fdiv_rn(steep ? ax : ay, __fadd_rn(steep ? ay : ax, (float)2.2204460492503131e-16));
    const float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    if (steep) a = __fsub_rn(90.f, a);
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

