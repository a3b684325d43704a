// ORB extraction on sm_100a: scale pyramid, 7x7 blur, per-cell FAST-9 with threshold retry, quadtree keypoint
// distribution, intensity-centroid orientation and 256-bit rBRIEF.  Replaces ORBextractor::operator()
// (reference src/ORBextractor.cc:1068-1150) and everything it calls.  Integer/byte work, bit-exact against
// oracle/cpp/orb.cpp; float expressions are written with explicit _rn intrinsics (the TU is also built with
// -fmad=false) so that nothing is contracted into FMA.
#include "plf_ctx.cuh"
#include "blur.cuh"

namespace {

// rBRIEF pattern, transposed: [(4*test + coord)*32 + byte].  In global memory on purpose: each lane reads a different
// entry, which the constant cache would serialise 32-way; from global the 32 lanes read one 128-byte line.
__device__ const int d_patternT[1024] = {
#include "orb_pattern_t.inc"
};

// ---------------------------------------------------------------------------------------------------------------
// K1a  pyramid level l from level l-1: cv::resize(INTER_LINEAR) 8U, 11-bit fixed point (src/ORBextractor.cc:1165).
#define PYR_ROWS 4             // output rows per thread (the column taps are prepared once per thread)
// 128x8 tiles, 4 output pixels per thread (resize_quad: aligned word loads of the taps, one 32-bit store); each source
// byte is reused by ~1.4 output pixels in x and y through L1/L2, so the level is read from HBM once.
__global__ void __launch_bounds__(256) pyr_resize_kernel(PlfGeom g, uint8_t* pyr, const PlfLin* lin, int level, int imgFirst) {
    const PlfLevel& d = g.lv[level];
    const PlfLevel& s = g.lv[level - 1];
    const int dx = blockIdx.x * 128 + threadIdx.x * 4;
    if (dx >= d.w) return;
    const int img = imgFirst + blockIdx.z;
    const uint8_t* src = pyr + (size_t)img * g.pyrBytes + s.off;
    uint8_t* dcol = pyr + (size_t)img * g.pyrBytes + d.off + dx;
    const int nValid = min(4, d.w - dx);
    ResizeTaps T;                                                    // column part: once for the 4 rows of this thread
    resize_prep(lin + d.xTab + dx, nValid, T);
#pragma unroll
    for (int k = 0; k < PYR_ROWS; ++k) {
        const int dy = blockIdx.y * (8 * PYR_ROWS) + threadIdx.y + 8 * k;
        if (dy >= d.h) break;
        const PlfLin cy = lin[d.yTab + dy];                            // source index + 11-bit weights, built on the host
        const uint8_t* r0 = src + (size_t)cy.ofs * s.pitch;
        const uint8_t* r1 = src + (size_t)min((int)cy.ofs + 1, s.h - 1) * s.pitch;
        const unsigned v = resize_apply<false>(T, r0, r1, cy.a0, cy.a1);
        uint8_t* dst = dcol + (size_t)dy * d.pitch;
        if (nValid == 4) *reinterpret_cast<unsigned*>(dst) = v;
        else for (int j = 0; j < nValid; ++j) dst[j] = (uint8_t)(v >> (8 * j));
    }
}

// all pyramid levels in one launch: blockIdx.x enumerates the 32x32 tiles of every level
__global__ void __launch_bounds__(256) blur_pyramid_kernel(PlfGeom g, const uint8_t* pyr, uint8_t* blur,
                                                           const PlfTile* tiles, int imgFirst) {
    const PlfTile t = tiles[blockIdx.x];
    const int img = imgFirst + blockIdx.y;
    const PlfLevel& lv = g.lv[t.level];
    BlurJob j;
    j.src = pyr + (size_t)img * g.pyrBytes + lv.off;
    j.dst = blur + (size_t)img * g.pyrBytes + lv.off;
    j.w = lv.w; j.h = lv.h; j.sp = j.dp = lv.pitch;
    const int taps[7] = {18, 34, 48, 56, 48, 34, 18};
    blur_tile7(j, taps, t.x0, t.y0);
}

// ---------------------------------------------------------------------------------------------------------------
// K2a  FAST-9/16 per grid cell with the ORB_SLAM3 retry: cv::FAST(window, iniTh, nms) and, only if the cell came back
// empty, cv::FAST(window, minTh, nms) (src/ORBextractor.cc:787-854).  The corner score s = max{t : still a corner} does
// not depend on the threshold, and for a pixel with s >= t strict 3x3 NMS against thresholded neighbours equals NMS
// against raw scores, so ONE score map serves both thresholds.  Two kernels:
//   fast_score_kernel : pixel-parallel score map of every level (128x8 tiles staged in shared memory as words, all
//                       levels in one launch, 4 pixels per thread); scores below minTh are stored as 0.
//   fast_cells_kernel : warp per cell window: strict NMS restricted to the cell's detection area (neighbours outside
//                       count as 0, exactly like FAST on the sub-image), ini/min threshold choice, ordered compaction.
//
// Corner score by bisection on t with 16-bit arc masks — compare/logic ops only, on purpose: a min/max formulation
// (OpenCV's cornerScore) is fused by ptxas 12.9 — and by the 580 driver's JIT — into VIMNMX3 chains that return wrong
// values on sm_100a when min, max and negation are mixed (reproduced in isolation; ptxas -O0, which emits no VIMNMX3,
// is correct).  tests/test_build.py asserts that the library contains no VIMNMX3.
__device__ __forceinline__ bool fast_run9(unsigned m) {
    const unsigned m2 = m | (m << 16);
    unsigned r = m2 & (m2 >> 1);
    r &= r >> 2;
    r &= r >> 4;
    r &= (m2 >> 8);
    return (r & 0xFFFFu) != 0u;
}
// Exact corner score from the 16 ring pixels packed four to a word (byte j of N[w] = ring pixel 4w+j), or 0 if the
// pixel is not a FAST-9 corner at minTh.  Polarity by majority: a 9-arc needs >= 9 of the 16 ring pixels on one side.
// With e[k] = the ring pixel's distance to the centre on that side (0 on the other side), the score is M-1 with
// M = max over 9-arcs of min(e), found by bit-sliced bisection: plane b holds bit b of the 16 e[k] (gathered from the
// packed bytes with one multiply per word); walking the planes from the top keeps the sets {e > prefix} and
// {e == prefix}, so each of the 8 steps costs a handful of logic ops plus one 9-run test.
__device__ __forceinline__ unsigned fast_gt_const4(unsigned x, unsigned k7f) {
    // per byte: x > t  (t < 128, k7f = (0x7F - t) * 0x01010101); result in bit 7 of every byte
    return (((x & 0x7F7F7F7Fu) + k7f) | x) & 0x80808080u;
}
__device__ __forceinline__ int fast_score_packed(const unsigned* N, unsigned centre, int minTh, unsigned k7f) {
    const unsigned C4 = centre * 0x01010101u;
    unsigned A[4], G[4];
    int nDark = 0, nBright = 0;                    // ring pixels darker / brighter than the centre by more than minTh
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        A[w] = __vabsdiffu4(C4, N[w]);
        G[w] = __vcmpgtu4(C4, N[w]);
        const unsigned far = fast_gt_const4(A[w], k7f);
        nDark += __popc(far & G[w]);
        nBright += __popc(far & ~G[w]);
    }
    if (nDark < 9 && nBright < 9) return 0;
    const unsigned flip = nDark >= 9 ? 0u : 0xFFFFFFFFu;
    unsigned E[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) E[w] = A[w] & (G[w] ^ flip);
    unsigned gt = 0u, eq = 0xFFFFu;
    int M = 0;
#pragma unroll
    for (int b = 7; b >= 0; --b) {
        unsigned P = 0u;
#pragma unroll
        for (int w = 0; w < 4; ++w) P |= ((((E[w] >> b) & 0x01010101u) * 0x01020408u) >> 24) << (4 * w);
        if (fast_run9(gt | (eq & P))) { M |= 1 << b; eq &= P; }
        else { gt |= eq & P; eq &= ~P; }
    }
    return M - 1 >= minTh ? M - 1 : 0;
}

#define FS_TW 128
#define FS_TH PLF_FAST_TH              // tile rows (8 until round 2: 14 staged rows per 8 output rows; now 38 per 32)
#define FS_ROWW 36            // 32-bit words per staged row: (128 + 6 bytes -> 34 words) padded to a multiple of 4
#define FS_IH (FS_TH + 6)
// Packed FAST-9 candidate test: every thread owns 4 horizontally adjacent pixels held as 4 bytes of one register, so
// the 16 circle comparisons of the 4 pixels are byte-SIMD integer ops (VABSDIFF4 + a carry-free byte compare) and the
// "9 contiguous" test is a handful of 3-input ANDs over the 16 per-neighbour flag words — no divergence, the cost
// does not depend on the image content.  The packed test is polarity-blind (|centre - ring| > minTh on 9 contiguous
// ring pixels): a superset of the corners that costs half the instructions of two polarity masks.  Candidates (a few
// per cent of the pixels) are queued and resolved densely: polarity by majority (a 9-arc needs >= 9 of the 16), exact
// score, and the score decides (>= minTh <=> FAST-9 corner at minTh).  In front of it, a four-point compass test
// discards most quads with a quarter of the comparisons; only the survivors (queued, processed densely) run the 16-point test.

__global__ void __launch_bounds__(256) fast_score_kernel(PlfGeom g, const uint8_t* pyr, uint8_t* score,
                                                         const PlfTile* tiles, int imgFirst) {
    __shared__ __align__(16) unsigned s_w[FS_IH][FS_ROWW];
    __shared__ int s_cnt1, s_cnt2;
    __shared__ unsigned short s_quads[FS_TW / 4 * FS_TH];  // phase-1 survivors: quad id (row * 32 + column group) | valid-pixel mask << 10
    __shared__ unsigned short s_queue[FS_TW * FS_TH];      // phase-2 survivors: (row << 7) | column
    // tiles cover [19, w-19) x [19, h-19) of every level: the union of all cell detection areas
    const PlfTile t = tiles[blockIdx.x];
    const PlfLevel& lv = g.lv[t.level];
    const int img = imgFirst + blockIdx.y;
    const uint8_t* src = pyr + (size_t)img * g.pyrBytes + lv.off;
    uint8_t* dst = score + (size_t)img * g.pyrBytes + lv.off + 1;      // score(x,y) lives at byte x+1: word-aligned rows of 4
    const int x0 = t.x0, y0 = t.y0;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    if (tid == 0) { s_cnt1 = 0; s_cnt2 = 0; }
    // the window starts at x0-3 = 16 (mod 128) and level rows are 64-byte aligned: aligned 32-bit loads (the row pitch
    // is padded to 64 B and the pyramid allocation has slack, so the last words never leave the allocation)
    for (int iy = ty; iy < FS_IH; iy += 8) {              // a warp per staged row: 34 words
        const uint8_t* row = src + (size_t)min(y0 - 3 + iy, lv.h - 1) * lv.pitch + (x0 - 3);
        s_w[iy][tx] = *reinterpret_cast<const unsigned*>(row + tx * 4);
        if (tx < 2) s_w[iy][32 + tx] = *reinterpret_cast<const unsigned*>(row + (32 + tx) * 4);
    }
    __syncthreads();
    // neighbour bytes of the 4 pixels of quad (qx, qy) at circle offset (dx, dy): bytes [4qx+3+dx, +4) of staged row qy+3+dy
    auto nb = [&](int qx, int qy, int dx, int dy) -> unsigned {
        const int b = 3 + dx;                      // 0..6, compile-time after unrolling
        const unsigned* r = &s_w[qy + 3 + dy][qx + (b >> 2)];
        return (b & 3) ? __funnelshift_r(r[0], r[1], (b & 3) * 8) : r[0];
    };
    const unsigned k7f = (unsigned)(0x7F - g.minTh) * 0x01010101u;
    const int cdx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int cdy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    // ---- phase 1, every thread: the four compass points.  Nine contiguous ring pixels always contain two compass points
    // that are neighbours on the compass, so a pixel without such a pair (polarity-blind, like the full test below) cannot
    // be a corner; most quads of a frame end here.
#pragma unroll
    for (int rk = 0; rk < FS_TH / 8; ++rk) {
        const int ry = ty + 8 * rk;                            // my row of the tile in this round
        const unsigned C = nb(tx, ry, 0, 0);
        const unsigned f0 = fast_gt_const4(__vabsdiffu4(C, nb(tx, ry, 0, 3)), k7f), f4 = fast_gt_const4(__vabsdiffu4(C, nb(tx, ry, 3, 0)), k7f);
        const unsigned f8 = fast_gt_const4(__vabsdiffu4(C, nb(tx, ry, 0, -3)), k7f), f12 = fast_gt_const4(__vabsdiffu4(C, nb(tx, ry, -3, 0)), k7f);
        unsigned quick = ((f0 | f8) & (f4 | f12)) & 0x80808080u;          // = (f0&f4)|(f4&f8)|(f8&f12)|(f12&f0)
        const int y = y0 + ry;
        const bool rowIn = y < lv.h - PLF_EDGE;
        const int xr = lv.w - PLF_EDGE - (x0 + 4 * tx);        // number of valid pixels from my first one
        if (!rowIn || xr <= 0) quick = 0;
        else if (xr < 4) quick &= (0xFFFFFFFFu >> (8 * (4 - xr)));
        if (rowIn && xr > 0) {
            // everything scores 0 unless the scoring pass below (same block, after the barriers) says otherwise
            unsigned* d4 = reinterpret_cast<unsigned*>(dst + (size_t)y * lv.pitch + x0 + 4 * tx);
            if (xr >= 4) *d4 = 0u;
            else for (int j = 0; j < xr; ++j) dst[(size_t)y * lv.pitch + x0 + 4 * tx + j] = 0;
        }
        const unsigned vote = __ballot_sync(0xffffffffu, quick != 0u);
        if (vote) {                                            // one shared-memory atomic per warp
            int base = 0;
            if (tx == 0) base = atomicAdd(&s_cnt1, __popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (quick) {
                const unsigned m4 = ((quick >> 7) & 1u) | ((quick >> 14) & 2u) | ((quick >> 21) & 4u) | ((quick >> 28) & 8u);
                s_quads[base + __popc(vote & ((1u << tx) - 1u))] = (unsigned short)((ry * 32 + tx) | (m4 << 10));
            }
        }
    }
    __syncthreads();
    // ---- phase 2, dense over the surviving quads: the packed, polarity-blind 16-point test
    const int n1 = s_cnt1;
    for (int i = tid; i < n1; i += 256) {
        const int e = s_quads[i];
        const int qx = e & 31, qy = (e >> 5) & 31;
        const unsigned m4 = (unsigned)e >> 10;
        const unsigned C = nb(qx, qy, 0, 0);
        unsigned F[16];                                // bit 7 of byte j: |centre - neighbour k| of pixel j exceeds minTh
#pragma unroll
        for (int k = 0; k < 16; ++k) F[k] = fast_gt_const4(__vabsdiffu4(C, nb(qx, qy, cdx[k], cdy[k])), k7f);
        // 9 contiguous flags: T3[k] = F[k]&F[k+1]&F[k+2]; run[k] = T3[k]&T3[k+3]&T3[k+6]; any = OR run[k]
        unsigned any = 0;
        {
            unsigned T[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) T[k] = F[k] & F[(k + 1) & 15] & F[(k + 2) & 15];
#pragma unroll
            for (int k = 0; k < 16; ++k) any |= T[k] & T[(k + 3) & 15] & T[(k + 6) & 15];
        }
        const unsigned valid = ((m4 & 1u) << 7) | ((m4 & 2u) << 14) | ((m4 & 4u) << 21) | ((m4 & 8u) << 28);
        const unsigned corner = any & valid;
        const int cnt = __popc(corner);
        if (cnt) {
            int pos = atomicAdd(&s_cnt2, cnt);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (corner & (0x80u << (8 * j))) s_queue[pos++] = (unsigned short)((qy << 7) | (4 * qx + j));
        }
    }
    __syncthreads();
    // ---- phase 3, dense over the candidates: polarity by majority, exact score; the score decides
    const int nq = s_cnt2;
    const uint8_t* sb = reinterpret_cast<const uint8_t*>(&s_w[0][0]);
    constexpr int ST = FS_ROWW * 4;
    for (int i = tid; i < nq; i += 256) {
        const int e = s_queue[i];
        const int py = (e >> 7) & 0xFF, px = e & 0x7F;
        const uint8_t* q = sb + (py + 3) * ST + px + 3;
        unsigned N[4];
#pragma unroll
        for (int w = 0; w < 4; ++w)
            N[w] = (unsigned)q[cdy[4 * w] * ST + cdx[4 * w]] | ((unsigned)q[cdy[4 * w + 1] * ST + cdx[4 * w + 1]] << 8) |
                   ((unsigned)q[cdy[4 * w + 2] * ST + cdx[4 * w + 2]] << 16) | ((unsigned)q[cdy[4 * w + 3] * ST + cdx[4 * w + 3]] << 24);
        const int sc = fast_score_packed(N, *q, g.minTh, k7f);
        if (sc) dst[(size_t)(y0 + py) * lv.pitch + x0 + px] = (uint8_t)sc;
    }
}

#define FC_WARPS 4
#define FC_PITCH 20    // words per staged row: detection areas are <= 66 px wide (window <= 72, checked at plf_create),
#define FC_ROWS 68     // i.e. <= 18 aligned words, plus one zero word / zero row on every side
// One warp per cell.  The detection areas of a level tile it without overlap, so every score byte is read once, as
// aligned 32-bit words whose bytes outside the area are masked to 0 (FAST on the sub-image never sees them).  Scores
// are sparse: only non-zero words do the 3x3 strict-maximum test (byte-SIMD compares against the 8 shifted neighbour
// words), survivors are appended in raster order with two ballots (strict NMS leaves at most 2 per word), and the
// iniTh/minTh choice of src/ORBextractor.cc:806-827 becomes an in-place ordered filter of that short list.
__global__ void __launch_bounds__(32 * FC_WARPS) fast_cells_kernel(PlfGeom g, const uint8_t* score, const PlfCell* cells,
                                                                   int* cellCount, uint32_t* cand, int imgFirst) {
    __shared__ unsigned s_all[FC_WARPS][FC_ROWS * FC_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci = blockIdx.x * FC_WARPS + warp;
    if (ci >= g.nCellsTotal) return;                 // warp-uniform; no block-level barrier below
    const PlfCell c = cells[ci];
    const int img = imgFirst + blockIdx.y;
    const PlfLevel& lv = g.lv[c.level];
    const int aw = c.x1 - c.x0 - 6, ah = c.y1 - c.y0 - 6;     // detection area of cv::FAST on the window
    int* outCount = cellCount + (size_t)img * g.nCellsTotal + ci;
    if (aw <= 0 || ah <= 0) {
        if (lane == 0) *outCount = 0;
        return;
    }
    unsigned* sw = s_all[warp];
    const int bx0 = c.x0 + 4;                        // score(x, y) lives at byte x+1 of its row
    const int wb0 = bx0 >> 2, nw = ((bx0 + aw + 3) >> 2) - wb0, pw = nw + 2;
    const unsigned maskFirst = 0xFFFFFFFFu << (8 * (bx0 & 3));
    const unsigned maskLast = ((bx0 + aw) & 3) ? (0xFFFFFFFFu >> (8 * (4 - ((bx0 + aw) & 3)))) : 0xFFFFFFFFu;
    const unsigned* src = reinterpret_cast<const unsigned*>(score + (size_t)img * g.pyrBytes + lv.off) +
                          (size_t)(c.y0 + 3) * (lv.pitch >> 2) + wb0;
    const int pitchW = lv.pitch >> 2;
    {
        const unsigned inv = (65536u + pw - 1) / pw;          // i / pw == (i * inv) >> 16 for i < 68 * 20
        const int n = (ah + 2) * pw;
        for (int i = lane; i < n; i += 32) {
            const int r = (int)(((unsigned)i * inv) >> 16), wx = i - r * pw;
            unsigned v = 0u;
            if (r >= 1 && r <= ah && wx >= 1 && wx <= nw) {
                v = src[(size_t)(r - 1) * pitchW + (wx - 1)];
                if (wx == 1) v &= maskFirst;
                if (wx == nw) v &= maskLast;
            }
            sw[i] = v;
        }
    }
    __syncwarp();
    uint32_t* out = cand + (size_t)img * g.candCapTotal + c.outBase;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned ini4 = (unsigned)g.iniTh * 0x01010101u;
    const int yBase = c.y0 + 3 - PLF_MINB;
    int total = 0;
    bool anyIni = false;
    {
        const unsigned inv = (65536u + nw - 1) / nw;
        const int n = ah * nw, nIter = (n + 31) >> 5;
        for (int it = 0; it < nIter; ++it) {
            const int i = it * 32 + lane;
            int r = 0, wx = 0;
            unsigned W = 0u;
            if (i < n) {
                r = (int)(((unsigned)i * inv) >> 16); wx = i - r * nw;
                W = sw[(r + 1) * pw + wx + 1];
            }
            if (__ballot_sync(0xffffffffu, W != 0u) == 0u) continue;
            unsigned K = 0u;
            if (W) {
                const unsigned* q = sw + r * pw + wx;             // top-left neighbour word
                unsigned keep = __vcmpgtu4(W, 0u);
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const unsigned a = q[dy * pw], m = q[dy * pw + 1], b = q[dy * pw + 2];
                    keep &= __vcmpgtu4(W, __funnelshift_r(a, m, 24));   // x-1
                    keep &= __vcmpgtu4(W, __funnelshift_r(m, b, 8));    // x+1
                    if (dy != 1) keep &= __vcmpgtu4(W, m);
                }
                K = W & keep;
            }
            const int cnt = __popc(__vcmpgtu4(K, 0u) & 0x01010101u);
            const unsigned b1 = __ballot_sync(0xffffffffu, cnt >= 1), b2 = __ballot_sync(0xffffffffu, cnt >= 2);
            if (cnt) {
                anyIni |= __vcmpgeu4(K, ini4) != 0u;
                int pos = total + __popc(b1 & lt) + __popc(b2 & lt);
                const int x = 4 * (wb0 + wx) - 1 - PLF_MINB;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned sv = (K >> (8 * j)) & 0xFFu;
                    if (sv) out[pos++] = (uint32_t)(x + j) | ((uint32_t)(yBase + r) << 12) | (sv << 24);
                }
            }
            total += __popc(b1) + __popc(b2);
        }
    }
    // scores below minTh were stored as 0, so the list already is the minTh result; iniTh keeps its >= iniTh subset
    if (__any_sync(0xffffffffu, anyIni) && g.iniTh > g.minTh) {
        __syncwarp();
        int wpos = 0;
        for (int base = 0; base < total; base += 32) {
            const int i = base + lane;
            const uint32_t v = i < total ? out[i] : 0u;
            const bool k = (int)(v >> 24) >= g.iniTh;
            const unsigned b = __ballot_sync(0xffffffffu, k);
            if (k) out[wpos + __popc(b & lt)] = v;
            wpos += __popc(b);
            __syncwarp();
        }
        total = wpos;
    }
    if (lane == 0) *outCount = total;
}

// ---------------------------------------------------------------------------------------------------------------
// K2b  quadtree ("octree") keypoint distribution, ORBextractor::DistributeOctTree (src/ORBextractor.cc:537-761).
// One warp per (level, image).  The std::list of nodes is a doubly linked list in shared memory that every lane
// reads uniformly and lane 0 mutates; the per-node key vectors are contiguous slices of two ping-pong point buffers
// and a node is split by a warp-wide stable 4-way partition (ballot + popc).  List order, the expansion order of the
// final phase ((count, creation order) largest first — the declared stand-in for the reference's heap-address tie
// break) and the first-max-response rule are reproduced exactly, so the retained set AND its order are bit-exact.
struct QNode {
    short ulx, uly, brx, bry;
    int start, cnt;
    int seq;
    short next, prev;
    short buf;      // which ping-pong buffer holds the points
    short pad;
};

struct QState {
    int head, tail, size, freeHead, seq;
};

__device__ __forceinline__ int q_alloc(QNode* nodes, QState* st) {
    int n = st->freeHead;
    st->freeHead = nodes[n].next;
    return n;
}
__device__ __forceinline__ void q_free(QNode* nodes, QState* st, int n) {
    nodes[n].next = (short)st->freeHead;
    st->freeHead = n;
}
__device__ __forceinline__ void q_push_front(QNode* nodes, QState* st, int n) {
    nodes[n].prev = -1;
    nodes[n].next = (short)st->head;
    if (st->head >= 0) nodes[st->head].prev = (short)n; else st->tail = n;
    st->head = n;
    st->size++;
}
__device__ __forceinline__ void q_push_back(QNode* nodes, QState* st, int n) {
    nodes[n].next = -1;
    nodes[n].prev = (short)st->tail;
    if (st->tail >= 0) nodes[st->tail].next = (short)n; else st->head = n;
    st->tail = n;
    st->size++;
}
__device__ __forceinline__ void q_unlink(QNode* nodes, QState* st, int n) {
    const int p = nodes[n].prev, q = nodes[n].next;
    if (p >= 0) nodes[p].next = (short)q; else st->head = q;
    if (q >= 0) nodes[q].prev = (short)p; else st->tail = p;
    st->size--;
}

// Splits node n (ExtractorNode::DivideNode, :479-535): partitions its points into the other buffer, pushes the
// non-empty children to the list front in order n1..n4, appends children with >1 points to expandList.
// Executed by the whole warp; returns the number of children with more than one point.
__device__ int q_divide(QNode* nodes, QState* st, int n, uint32_t* buf0, uint32_t* buf1, int* expandList,
                        int* nExpand, int lane) {
    const QNode nd = nodes[n];
    const int halfX = (nd.brx - nd.ulx + 1) >> 1, halfY = (nd.bry - nd.uly + 1) >> 1;   // ceil(w/2)
    const int mx = nd.ulx + halfX, my = nd.uly + halfY;
    const uint32_t* src = (nd.buf ? buf1 : buf0) + nd.start;
    uint32_t* dst = (nd.buf ? buf0 : buf1) + nd.start;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int i = lane; i < ((nd.cnt + 31) & ~31); i += 32) {
        int q = -1;
        if (i < nd.cnt) {
            const uint32_t p = src[i];
            const int x = p & 0xFFF, y = (p >> 12) & 0xFFF;
            q = (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
        }
        c0 += __popc(__ballot_sync(0xffffffffu, q == 0));
        c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
        c2 += __popc(__ballot_sync(0xffffffffu, q == 2));
        c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
    }
    int b0 = 0, b1 = c0, b2 = c0 + c1, b3 = c0 + c1 + c2;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = lane; i < ((nd.cnt + 31) & ~31); i += 32) {
        int q = -1;
        uint32_t p = 0;
        if (i < nd.cnt) {
            p = src[i];
            const int x = p & 0xFFF, y = (p >> 12) & 0xFFF;
            q = (x < mx) ? ((y < my) ? 0 : 2) : ((y < my) ? 1 : 3);
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, q == 0), m1 = __ballot_sync(0xffffffffu, q == 1);
        const unsigned m2 = __ballot_sync(0xffffffffu, q == 2), m3 = __ballot_sync(0xffffffffu, q == 3);
        if (q == 0) dst[b0 + __popc(m0 & lt)] = p;
        else if (q == 1) dst[b1 + __popc(m1 & lt)] = p;
        else if (q == 2) dst[b2 + __popc(m2 & lt)] = p;
        else if (q == 3) dst[b3 + __popc(m3 & lt)] = p;
        b0 += __popc(m0); b1 += __popc(m1); b2 += __popc(m2); b3 += __popc(m3);
    }
    __syncwarp();
    int nBig = 0;
    const int cnts[4] = {c0, c1, c2, c3};
    const int starts[4] = {nd.start, nd.start + c0, nd.start + c0 + c1, nd.start + c0 + c1 + c2};
    if (lane == 0) {
        for (int q = 0; q < 4; ++q) {
            if (cnts[q] == 0) continue;
            const int k = q_alloc(nodes, st);
            QNode& ch = nodes[k];
            ch.ulx = (q & 1) ? (short)mx : nd.ulx;
            ch.brx = (q & 1) ? nd.brx : (short)mx;
            ch.uly = (q & 2) ? (short)my : nd.uly;
            ch.bry = (q & 2) ? nd.bry : (short)my;
            ch.start = starts[q];
            ch.cnt = cnts[q];
            ch.buf = nd.buf ^ 1;
            ch.seq = st->seq++;
            q_push_front(nodes, st, k);
            if (cnts[q] > 1) expandList[(*nExpand)++] = k;
        }
    }
    for (int q = 0; q < 4; ++q) nBig += cnts[q] > 1;
    __syncwarp();
    return nBig;
}

__global__ void __launch_bounds__(32) octree_kernel(PlfGeom g, const PlfCell* cells, const int* cellCount,
                                                    const uint32_t* cand, uint32_t* scratch, uint32_t* lvlKp,
                                                    int* lvlN, int* err, int imgFirst, int poolSize) {
    extern __shared__ unsigned char smem_raw[];
    const int level = blockIdx.x, img = imgFirst + blockIdx.y, lane = threadIdx.x;
    const PlfLevel& lv = g.lv[level];
    QNode* nodes = reinterpret_cast<QNode*>(smem_raw);
    int* listA = reinterpret_cast<int*>(nodes + poolSize);
    int* listB = listA + poolSize;
    __shared__ QState st;
    __shared__ int s_nA, s_nB;
    const int N = lv.quota;
    uint32_t* buf0 = scratch + ((size_t)img * 2 + 0) * g.candCapTotal + lv.candOff;
    uint32_t* buf1 = scratch + ((size_t)img * 2 + 1) * g.candCapTotal + lv.candOff;
    const uint32_t* cnd = cand + (size_t)img * g.candCapTotal;
    const int* cc = cellCount + (size_t)img * g.nCellsTotal + lv.cellFirst;
    const PlfCell* cl = cells + lv.cellFirst;
    uint32_t* outKp = lvlKp + (size_t)img * g.kpLevelCapTotal + lv.kpOff;
    int* outN = lvlN + (size_t)img * g.nLevels + level;
    const unsigned lt = (1u << lane) - 1u;

    // root nodes (:541-561): nIni = round(W'/H'), hX = W'/nIni in float; a point goes to root (int)(x/hX)
    const int Wd = lv.w - 2 * PLF_MINB, Hd = lv.h - 2 * PLF_MINB;
    const int nIni = lv.nIni;   // <= 4, checked at plf_create
    const float hX = __fdiv_rn((float)Wd, (float)nIni);
    if (lane == 0) {
        st.head = st.tail = -1;
        st.size = 0;
        st.seq = 0;
        st.freeHead = 0;
        for (int i = 0; i < poolSize; ++i) nodes[i].next = (short)(i + 1 < poolSize ? i + 1 : -1);
        s_nA = s_nB = 0;
    }
    __syncwarp();
    // stable bucketing of vToDistributeKeys (cell-major, raster inside a cell) into the roots -> buf0
    int rc[4] = {0, 0, 0, 0};
    for (int ci = 0; ci < lv.nCells; ++ci) {
        const int cnt = cc[ci], base = cl[ci].outBase;
        for (int i = lane; i < ((cnt + 31) & ~31); i += 32) {
            int r = -1;
            if (i < cnt) r = (int)__fdiv_rn((float)(cnd[base + i] & 0xFFF), hX);
#pragma unroll
            for (int q = 0; q < 4; ++q) rc[q] += __popc(__ballot_sync(0xffffffffu, r == q));
        }
    }
    int rs[4] = {0, rc[0], rc[0] + rc[1], rc[0] + rc[1] + rc[2]};
    {
        int run[4] = {rs[0], rs[1], rs[2], rs[3]};
        for (int ci = 0; ci < lv.nCells; ++ci) {
            const int cnt = cc[ci], base = cl[ci].outBase;
            for (int i = lane; i < ((cnt + 31) & ~31); i += 32) {
                int r = -1;
                uint32_t p = 0;
                if (i < cnt) {
                    p = cnd[base + i];
                    r = (int)__fdiv_rn((float)(p & 0xFFF), hX);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const unsigned m = __ballot_sync(0xffffffffu, r == q);
                    if (r == q) buf0[run[q] + __popc(m & lt)] = p;
                    run[q] += __popc(m);
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        for (int i = 0; i < nIni; ++i) {
            const int seq = st.seq++;
            if (rc[i] == 0) continue;   // empty roots are erased (:579-580)
            const int k = q_alloc(nodes, &st);
            QNode& nd = nodes[k];
            nd.ulx = (short)(int)__fmul_rn(hX, (float)i);
            nd.brx = (short)(int)__fmul_rn(hX, (float)(i + 1));
            nd.uly = 0;
            nd.bry = (short)Hd;
            nd.start = rs[i];
            nd.cnt = rc[i];
            nd.buf = 0;
            nd.seq = seq;
            q_push_back(nodes, &st, k);
        }
    }
    __syncwarp();

    bool finish = (st.size == 0);
    while (!finish) {
        const int prevSize = st.size;
        int nToExpand = 0;
        if (lane == 0) s_nA = 0;
        __syncwarp();
        int cur = st.head;
        while (cur >= 0) {
            const int nxt = nodes[cur].next;
            if (nodes[cur].cnt > 1) {
                nToExpand += q_divide(nodes, &st, cur, buf0, buf1, listA, &s_nA, lane);
                if (lane == 0) { q_unlink(nodes, &st, cur); q_free(nodes, &st, cur); }
                __syncwarp();
            }
            cur = nxt;
        }
        if (st.size >= N || st.size == prevSize) {
            finish = true;
        } else if (st.size + nToExpand * 3 > N) {
            int* prev = listA;
            int* nextL = listB;
            int* nPrev = &s_nA;
            int* nNext = &s_nB;
            while (!finish) {
                const int prevSize2 = st.size;
                if (lane == 0) *nNext = 0;
                __syncwarp();
                const int m = *nPrev;
                for (int it = 0; it < m; ++it) {
                    // next node to expand: largest (count, creation order) among the unprocessed entries
                    unsigned long long best = 0;
                    int bestIdx = -1;
                    for (int i = lane; i < m; i += 32) {
                        const int k = prev[i];
                        if (k < 0) continue;
                        const unsigned long long key = ((unsigned long long)nodes[k].cnt << 32) | (unsigned)nodes[k].seq;
                        if (bestIdx < 0 || key > best) { best = key; bestIdx = i; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
                        if (oi >= 0 && (bestIdx < 0 || ob > best)) { best = ob; bestIdx = oi; }
                    }
                    const int k = prev[bestIdx];
                    __syncwarp();
                    if (lane == 0) prev[bestIdx] = -1;
                    q_divide(nodes, &st, k, buf0, buf1, nextL, nNext, lane);
                    if (lane == 0) { q_unlink(nodes, &st, k); q_free(nodes, &st, k); }
                    __syncwarp();
                    if (st.size >= N) break;
                }
                if (st.size >= N || st.size == prevSize2) finish = true;
                int* t = prev; prev = nextL; nextL = t;
                int* tn = nPrev; nPrev = nNext; nNext = tn;
            }
        }
    }
    // retain the best point of each node, in list order (:739-758): max response, first wins
    const int nOut = st.size;
    if (lane == 0) {
        int k = 0;
        for (int cur = st.head; cur >= 0; cur = nodes[cur].next) listA[k++] = cur;
        if (nOut > lv.kpCap) atomicOr(err, 1);
        *outN = min(nOut, lv.kpCap);
    }
    __syncwarp();
    for (int k = lane; k < min(nOut, lv.kpCap); k += 32) {
        const QNode& nd = nodes[listA[k]];
        const uint32_t* src = (nd.buf ? buf1 : buf0) + nd.start;
        uint32_t bp = src[0];
        for (int i = 1; i < nd.cnt; ++i) {
            const uint32_t p = src[i];
            if ((p >> 24) > (bp >> 24)) bp = p;
        }
        outKp[k] = bp;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3  orientation (IC_Angle, :75-102) + steered BRIEF (computeOrbDescriptor, :106-145): one warp per keypoint.
// Lanes own the 31 patch rows for the moments (integer sums -> order-free, exact) and one descriptor byte each.
__global__ void __launch_bounds__(256) orient_desc_kernel(PlfGeom g, const uint8_t* pyr, const uint8_t* blur,
                                                          const uint32_t* lvlKp, const int* lvlN,
                                                          plf_keypoint* kpTmp, uint8_t* descTmp, int* nKp,
                                                          int imgFirst) {
    __shared__ __align__(16) unsigned s_patch[8][37 * 11];      // per warp: blurred 37 x 44-byte window
    const int img = imgFirst + blockIdx.y;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int gk = blockIdx.x * 8 + wl;
    const int* ln = lvlN + (size_t)img * g.nLevels;
    int level = 0, base = 0, total = 0;
    bool found = false;
    for (int l = 0; l < g.nLevels; ++l) {
        const int n = ln[l];
        if (!found && gk < total + n) { level = l; base = total; found = true; }
        total += n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) nKp[img] = min(total, g.kpCap);
    if (!found || gk >= g.kpCap) return;
    const PlfLevel& lv = g.lv[level];
    const uint32_t p = lvlKp[(size_t)img * g.kpLevelCapTotal + lv.kpOff + (gk - base)];
    const int x = (int)(p & 0xFFF) + PLF_MINB, y = (int)((p >> 12) & 0xFFF) + PLF_MINB;
    const uint8_t* im = pyr + (size_t)img * g.pyrBytes + lv.off + (size_t)y * lv.pitch + x;
    // stage the blurred 37x37 window (rows/columns -18..18) in shared memory with aligned 32-bit row loads: the 512
    // rotated sample reads then hit shared memory instead of ~25 L1 sectors per warp load
    const uint8_t* blv = blur + (size_t)img * g.pyrBytes + lv.off;
    const int xa = (x - 18) & ~3;                          // aligned start column; x-18 >= 1
    const int sh0 = (x - 18) - xa;                         // 0..3
    unsigned* pw = s_patch[wl];
    {   // 37 rows x 11 words (37 + 3 bytes): all 13 loads of a lane are issued before the first is stored
        unsigned v[13];
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const int i = lane + 32 * k;
            const int r = i / 11, wx = i - r * 11;
            v[k] = (i < 37 * 11) ? *reinterpret_cast<const unsigned*>(blv + (size_t)(y - 18 + r) * lv.pitch + xa + wx * 4) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 13; ++k)
            if (lane + 32 * k < 37 * 11) pw[lane + 32 * k] = v[k];
    }
    const uint8_t* pc = reinterpret_cast<const uint8_t*>(pw) + 18 * 44 + 18 + sh0;      // centre of the staged window
    // moments: lane = column u of the 31x31 patch, loop over rows v (integer sums -> order-free, exact)
    int m10 = 0, m01 = 0;
    {
        const int u = lane - 15, au = u < 0 ? -u : u;
        if (lane < 31) {
#pragma unroll
            for (int v = -15; v <= 15; ++v) {
                const int val = (au <= g.umax[v < 0 ? -v : v]) ? (int)im[v * lv.pitch + u] : 0;
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // descriptor: lane = byte
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ar = __fmul_rn(angle, factorPI);
    // cosf/sinf taken as correctly rounded: the double result rounded to float (declared oracle rule)
    const float a = (float)cos((double)ar), b = (float)sin((double)ar);
    __syncwarp();
    const int* pat = d_patternT + lane;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)__ldg(pat + (4 * k) * 32), y0 = (float)__ldg(pat + (4 * k + 1) * 32);
        const float x1 = (float)__ldg(pat + (4 * k + 2) * 32), y1 = (float)__ldg(pat + (4 * k + 3) * 32);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = pc[r0 * 44 + c0], t1 = pc[r1 * 44 + c1];
        val |= (t0 < t1) << k;
    }
    descTmp[((size_t)img * g.kpCap + gk) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        plf_keypoint kp;
        kp.x = (float)x;
        kp.y = (float)y;
        if (level != 0) {   // :1131-1133
            kp.x = __fmul_rn(kp.x, lv.scale);
            kp.y = __fmul_rn(kp.y, lv.scale);
        }
        kp.size = (float)lv.scaledPatch;
        kp.angle = angle;
        kp.response = (float)(p >> 24);
        kp.octave = level;
        kp.class_id = -1;
        kpTmp[(size_t)img * g.kpCap + gk] = kp;
    }
}

// Row placement of ORBextractor::operator() (:1102-1146): rows outside the lapping area are written front to back,
// rows inside back to front.  One block per image, ordered block scan of the "inside" predicate.
__global__ void __launch_bounds__(1024) place_rows_kernel(PlfGeom g, const plf_keypoint* kpTmp, const uint8_t* descTmp,
                                                          const int* nKp, plf_keypoint* kpOut, uint8_t* descOut,
                                                          int* mono, int lap0, int lap1, int imgFirst) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int img = imgFirst + blockIdx.x;
    const int n = nKp[img];
    const plf_keypoint* src = kpTmp + (size_t)img * g.kpCap;
    const uint4* dsrc = reinterpret_cast<const uint4*>(descTmp + (size_t)img * g.kpCap * 32);
    plf_keypoint* dst = kpOut + (size_t)img * g.kpCap;
    uint4* ddst = reinterpret_cast<uint4*>(descOut + (size_t)img * g.kpCap * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + tid;
        plf_keypoint kp;
        int inside = 0;
        if (i < n) {
            kp = src[i];
            inside = (kp.x >= (float)lap0 && kp.x <= (float)lap1) ? 1 : 0;
        }
        int inc = inside;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int base = s_carry;
        for (int w = 0; w < warp; ++w) base += s_warp[w];
        const int insideBefore = base + inc - inside;   // # inside rows among [0, i)
        if (i < n) {
            const int pos = inside ? (n - 1 - insideBefore) : (i - insideBefore);
            dst[pos] = kp;
            ddst[pos * 2] = dsrc[i * 2];
            ddst[pos * 2 + 1] = dsrc[i * 2 + 1];
        }
        __syncthreads();
        if (tid == 1023) s_carry = base + inc;
        __syncthreads();
    }
    if (tid == 0) mono[img] = n - s_carry;
}

}  // namespace

namespace {
// staging buffer [side][frame][H][stride] -> level 0 of the pyramid block of image frame*2+side (16-byte vectors when
// the geometry allows it)
__global__ void __launch_bounds__(64) unpack_kernel(PlfGeom g, const uint8_t* stage, size_t sideBytes, int stride,
                                                     uint8_t* pyr) {
    const int img = blockIdx.z, frame = img >> 1, side = img & 1;
    const int y = blockIdx.y;
    const uint8_t* src = stage + side * sideBytes + ((size_t)frame * g.H + y) * stride;
    uint8_t* dst = pyr + (size_t)img * g.pyrBytes + g.lv[0].off + (size_t)y * g.lv[0].pitch;
    if ((stride & 15) == 0 && (g.W & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = threadIdx.x; i < g.W / 16; i += 64) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < g.W; i += 64) dst[i] = src[i];
    }
}
}  // namespace

int plf_launch_unpack(plf_ctx* c, const uint8_t* stage, size_t sideBytes, int stride, int batch) {
    unpack_kernel<<<dim3(1, c->g.H, 2 * batch), 64, 0, c->stream>>>(c->g, stage, sideBytes, stride, c->d_pyr);
    return 1;
}

int plf_launch_orb(plf_ctx* c, int imgFirst, int nImg, int lap0, int lap1) {
    const PlfGeom& g = c->g;
    cudaStream_t s = c->stream;
    int launches = 0;
    plf_mark(c, "orb_pyramid");
    for (int l = 1; l < g.nLevels; ++l) {
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 8 * PYR_ROWS - 1) / (8 * PYR_ROWS), nImg);
        pyr_resize_kernel<<<grid, dim3(32, 8), 0, s>>>(g, c->d_pyr, c->d_lin, l, imgFirst);
        ++launches;
    }
    plf_mark(c, "orb_blur");
    blur_pyramid_kernel<<<dim3(c->nTilesBlur, nImg), dim3(32, 8), 0, s>>>(g, c->d_pyr, c->d_blur, c->d_tilesBlur, imgFirst);
    plf_mark(c, "orb_fast");
    fast_score_kernel<<<dim3(c->nTilesFast, nImg), dim3(32, 8), 0, s>>>(g, c->d_pyr, c->d_score, c->d_tilesFast, imgFirst);
    fast_cells_kernel<<<dim3((g.nCellsTotal + FC_WARPS - 1) / FC_WARPS, nImg), 32 * FC_WARPS, 0, s>>>(g, c->d_score, c->d_cells, c->d_cellCount, c->d_cand,
                                                               imgFirst);
    int maxQ = 0;
    for (int l = 0; l < g.nLevels; ++l) maxQ = max(maxQ, max(g.lv[l].quota, 4 * g.lv[l].nIni));
    const int pool = maxQ + 16;
    const size_t smem = (size_t)pool * (sizeof(QNode) + 2 * sizeof(int));
    static size_t s_granted[64] = {};
    if (plf_raise_smem_optin(s_granted, c->device, smem))
        cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    plf_mark(c, "orb_octree");
    octree_kernel<<<dim3(g.nLevels, nImg), 32, smem, s>>>(g, c->d_cells, c->d_cellCount, c->d_cand, c->d_scratch,
                                                         c->d_lvlKp, c->d_lvlN, c->d_err, imgFirst, pool);
    plf_mark(c, "orb_orient_desc");
    orient_desc_kernel<<<dim3((g.kpCap + 7) / 8, nImg), 256, 0, s>>>(g, c->d_pyr, c->d_blur, c->d_lvlKp, c->d_lvlN,
                                                                    c->d_kpTmp, c->d_descTmp, c->d_nKp, imgFirst);
    place_rows_kernel<<<nImg, 1024, 0, s>>>(g, c->d_kpTmp, c->d_descTmp, c->d_nKp, c->d_kp, c->d_desc, c->d_mono,
                                            lap0, lap1, imgFirst);
    return launches + 6;
}
