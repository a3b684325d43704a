// Line matching on sm_100a: brute-force Hamming 2-NN with ratio test (matchNNR / match, reference
// src/LineMatcher.cpp:139-159,201-229) and the stereo line matcher of Frame::ComputeStereoMatches_Lines
// (src/Frame.cc:1156-1307) with matchGrid(lines) (src/LineMatcher.cpp:317-396), GridStructure
// (src/gridStructure.cpp:43-76) and the Bresenham LineIterator (src/LineIterator.cpp:34-77).
// 32-byte descriptors are compared with __popc over 8 words.  Bit-exact against oracle/cpp/linematch.cpp.
#include "plf_ctx.cuh"

namespace {

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4* b) {
    const uint4 b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// matchNNR: one warp per query row; knn k=2 with ties resolved towards the lower train index (cv::BFMatcher order),
// accept iff d0 < d1 * nnr in float.  n2 < 2 -> no match (declared rule for the reference's out-of-range read).
__global__ void __launch_bounds__(256) match_nnr_kernel(const uint8_t* dA, int nA, const uint8_t* dB, int nB, float nnr,
                                                        int* out) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nA) return;
    if (nB < 2) { if (lane == 0) out[i] = -1; return; }
    const uint4* a = reinterpret_cast<const uint4*>(dA + (size_t)i * 32);
    const uint4 a0 = a[0], a1 = a[1];
    int b0 = 0x7fffffff, b1 = 0x7fffffff, i0 = 0x7fffffff;
    for (int j = lane; j < nB; j += 32) {
        const int d = hamming256(a0, a1, reinterpret_cast<const uint4*>(dB + (size_t)j * 32));
        if (d < b0) { b1 = b0; b0 = d; i0 = j; }
        else if (d < b1) b1 = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
        const int oi0 = __shfl_xor_sync(0xffffffffu, i0, o);
        if (ob0 < b0 || (ob0 == b0 && oi0 < i0)) { b1 = min(b0, ob1); b0 = ob0; i0 = oi0; }
        else b1 = min(b1, ob0);
    }
    if (lane == 0) out[i] = (__fmul_rn((float)b1, nnr) > (float)b0) ? i0 : -1;
}

// Right lines -> 64x48 bucket grid.  Instead of per-cell index lists the grid is stored transposed: for every right
// line a 64-bit column mask per grid row, so "is i2 in the window [x-ws, x] x {y}" is one AND.
__global__ void __launch_bounds__(128) line_grid_kernel(PlfGeom g, const plf_keyline* kls, const int* nKl,
                                                        unsigned long long* rowMask, double2* dirR, int slotFirst) {
    const int slot = slotFirst + blockIdx.y;
    const int idx = blockIdx.x * 128 + threadIdx.x;
    const int nR = nKl[slot * 2 + 1];
    if (idx >= nR) return;
    const plf_keyline kl = kls[(size_t)(slot * 2 + 1) * g.klCap + idx];
    const double invW = PLF_GRID_COLS / (double)g.W, invH = PLF_GRID_ROWS / (double)g.H;
    double vx = (double)__fsub_rn(kl.endPointX, kl.startPointX) * invW, vy = (double)__fsub_rn(kl.endPointY, kl.startPointY) * invH;
    const double mag = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
    dirR[(size_t)slot * g.klCap + idx] = make_double2(vx / mag, vy / mag);
    unsigned long long* rm = rowMask + ((size_t)slot * g.klCap + idx) * PLF_GRID_ROWS;
    for (int r = 0; r < PLF_GRID_ROWS; ++r) rm[r] = 0ull;
    // LineIterator (Bresenham on doubles)
    double x1 = kl.startPointX * invW, y1 = kl.startPointY * invH, x2 = kl.endPointX * invW, y2 = kl.endPointY * invH;
    const bool steep = fabs(y2 - y1) > fabs(x2 - x1);
    if (steep) { double t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
    if (x1 > x2) { double t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const double dx = x2 - x1, dy = fabs(y2 - y1);
    double error = dx / 2.0;
    const int ystep = (y1 < y2) ? 1 : -1;
    int x = (int)x1, y = (int)y1;
    const int maxX = (int)x2;
    for (; x <= maxX; ++x) {
        const int px = steep ? y : x, py = steep ? x : y;
        if (px >= 0 && px < PLF_GRID_COLS && py >= 0 && py < PLF_GRID_ROWS) rm[py] |= 1ull << px;
        error -= dy;
        if (error < 0) { y += ystep; error += dx; }
    }
}

__device__ __forceinline__ unsigned long long window_mask(int x, int ws) {
    // GridStructure::get with w.width = (ws, 0): columns [max(0,x-ws), min(cols, x+1))
    const int lo = max(0, x - ws), hi = min(PLF_GRID_COLS, x + 1);
    if (lo >= hi) return 0ull;
    const unsigned long long upTo = (hi >= 64) ? ~0ull : ((1ull << hi) - 1ull);
    return upTo & ~((1ull << lo) - 1ull);
}

// distance matrix of matchGrid(lines): dmat[i1][i2] = Hamming distance if i2 is a grid-window candidate of i1 that
// passes the direction test, else 0xFFFF.
__global__ void __launch_bounds__(128) line_cand_kernel(PlfGeom g, const plf_keyline* kls, const uint8_t* ldesc,
                                                        const int* nKl, const unsigned long long* rowMask,
                                                        const double2* dirR, unsigned short* dmat, int ws, double simTh,
                                                        int slotFirst) {
    const int slot = slotFirst + blockIdx.z;
    const int i1 = blockIdx.y, i2 = blockIdx.x * 128 + threadIdx.x;
    const int nL = nKl[slot * 2], nR = nKl[slot * 2 + 1];
    if (i1 >= nL || i2 >= nR) return;
    const plf_keyline kl = kls[(size_t)(slot * 2) * g.klCap + i1];
    const double invW = PLF_GRID_COLS / (double)g.W, invH = PLF_GRID_ROWS / (double)g.H;
    const int sx = (int)(kl.startPointX * invW), sy = (int)(kl.startPointY * invH);
    const int ex = (int)(kl.endPointX * invW), ey = (int)(kl.endPointY * invH);
    const unsigned long long* rm = rowMask + ((size_t)slot * g.klCap + i2) * PLF_GRID_ROWS;
    bool cand = false;
    if (sy >= 0 && sy < PLF_GRID_ROWS) cand |= (rm[sy] & window_mask(sx, ws)) != 0ull;
    if (ey >= 0 && ey < PLF_GRID_ROWS) cand |= (rm[ey] & window_mask(ex, ws)) != 0ull;
    unsigned short d = 0xFFFF;
    if (cand) {
        double vx = (double)(ex - sx), vy = (double)(ey - sy);
        const double mag = sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
        vx /= mag;
        vy /= mag;
        const double2 dr = dirR[(size_t)slot * g.klCap + i2];
        const double dt = fabs(__dadd_rn(__dmul_rn(vx, dr.x), __dmul_rn(vy, dr.y)));
        if (!(dt < simTh)) {   // NaN (degenerate direction) is NOT skipped, as in the reference
            const uint4* a = reinterpret_cast<const uint4*>(ldesc + ((size_t)(slot * 2) * g.klCap + i1) * 32);
            d = (unsigned short)hamming256(a[0], a[1], reinterpret_cast<const uint4*>(ldesc + ((size_t)(slot * 2 + 1) * g.klCap + i2) * 32));
        }
    }
    dmat[((size_t)slot * g.klCap + i1) * g.klCap + i2] = d;
}

// best_lr bookkeeping (LineMatcher.cpp:364-369): scanning i1 upwards, a pair survives only if d < the running
// minimum of its column; matches_21[i2] = the last survivor.  One thread per column i2.
__global__ void __launch_bounds__(128) line_lr_kernel(PlfGeom g, const int* nKl, unsigned short* dmat, int* m21,
                                                      int slotFirst) {
    const int slot = slotFirst + blockIdx.y;
    const int i2 = blockIdx.x * 128 + threadIdx.x;
    const int nL = nKl[slot * 2], nR = nKl[slot * 2 + 1];
    if (i2 >= nR) return;
    unsigned short* col = dmat + (size_t)slot * g.klCap * g.klCap + i2;
    int best = 0x7fffffff, who = -1;
    for (int i1 = 0; i1 < nL; ++i1) {
        const int d = col[(size_t)i1 * g.klCap];
        if (d == 0xFFFF) continue;
        if (d < best) { best = d; who = i1; }
        else col[(size_t)i1 * g.klCap] = 0xFFFF;
    }
    m21[(size_t)slot * g.klCap + i2] = who;
}

// per left line: best / second best over surviving pairs + ratio test (LineMatcher.cpp:371-382).  Warp per row.
__global__ void __launch_bounds__(256) line_best_kernel(PlfGeom g, const int* nKl, const unsigned short* dmat, int* m12,
                                                        double ratio, int slotFirst) {
    const int slot = slotFirst + blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i1 = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nL = nKl[slot * 2], nR = nKl[slot * 2 + 1];
    if (i1 >= nL) return;
    const unsigned short* row = dmat + ((size_t)slot * g.klCap + i1) * g.klCap;
    int b0 = 0x7fffffff, b1 = 0x7fffffff, i0 = 0x7fffffff;
    for (int j = lane; j < nR; j += 32) {
        const int d = row[j];
        if (d == 0xFFFF) continue;
        if (d < b0) { b1 = b0; b0 = d; i0 = j; }
        else if (d < b1) b1 = d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob0 = __shfl_xor_sync(0xffffffffu, b0, o), ob1 = __shfl_xor_sync(0xffffffffu, b1, o);
        const int oi0 = __shfl_xor_sync(0xffffffffu, i0, o);
        if (ob0 < b0 || (ob0 == b0 && oi0 < i0)) { b1 = min(b0, ob1); b0 = ob0; i0 = oi0; }
        else b1 = min(b1, ob0);
    }
    if (lane == 0) {
        int r = -1;
        if (i0 != 0x7fffffff && (double)b0 < (double)b1 * ratio) r = i0;
        m12[(size_t)slot * g.klCap + i1] = r;
    }
}

// mutual check (LineMatcher.cpp:385-393) + endpoint disparities, overlap and filters (Frame.cc:1212-1252,1261-1307).
__global__ void __launch_bounds__(128) line_geom_kernel(PlfGeom g, const plf_keyline* kls, const int* nKl, int* m12,
                                                        const int* m21, float* disp, double* le, int bestLR,
                                                        double minDisp, double horizTh, double overlapTh,
                                                        double dispRatio, int slotFirst) {
    const int slot = slotFirst + blockIdx.y;
    const int i1 = blockIdx.x * 128 + threadIdx.x;
    const int nL = nKl[slot * 2], nR = nKl[slot * 2 + 1];
    if (i1 >= nL) return;
    float* dp = disp + ((size_t)slot * g.klCap + i1) * 2;
    double* lo = le + ((size_t)slot * g.klCap + i1) * 3;
    dp[0] = -1.f; dp[1] = -1.f;
    lo[0] = 0; lo[1] = 0; lo[2] = 0;
    int i2 = (nR > 0) ? m12[(size_t)slot * g.klCap + i1] : -1;
    if (i2 >= 0 && bestLR && m21[(size_t)slot * g.klCap + i2] != i1) i2 = -1;
    m12[(size_t)slot * g.klCap + i1] = i2;
    if (i2 < 0) return;
    const plf_keyline a = kls[(size_t)(slot * 2) * g.klCap + i1];
    const plf_keyline b = kls[(size_t)(slot * 2 + 1) * g.klCap + i2];
    const double splx = a.startPointX, sply = a.startPointY, eplx = a.endPointX, eply = a.endPointY;
    double l0 = __dsub_rn(sply, eply), l1 = __dsub_rn(eplx, splx);
    double l2 = __dsub_rn(__dmul_rn(splx, eply), __dmul_rn(sply, eplx));
    const double nrm = sqrt(__dadd_rn(__dmul_rn(l0, l0), __dmul_rn(l1, l1)));
    l0 = l0 / nrm; l1 = l1 / nrm; l2 = l2 / nrm;
    double sprx = b.startPointX, spry = b.startPointY, eprx = b.endPointX, epry = b.endPointY;
    double overlap = 1.0;
    if (fabs(eply - sply) > horizTh) {
        const double sln = fmin(sply, eply), eln = fmax(sply, eply);
        const double spn = fmin(spry, epry), epn = fmax(spry, epry);
        const double length = eln - spn;
        if ((epn < sln) || (spn > eln)) overlap = 0.0;
        else if ((epn > eln) && (spn < sln)) overlap = eln - sln;
        else overlap = fmin(eln, epn) - fmax(sln, spn);
        if (length > (double)0.01f) overlap = overlap / length;
        else overlap = 0.0;
        if (overlap > 1.0) overlap = 1.0;
    }
    // Frame.cc:1228-1229 — sp_r is overwritten first, its new value feeds the ep_r expression
    const double nsx = __dadd_rn(__dmul_rn(sprx, __dsub_rn(sply, epry)), __dmul_rn(eprx, __dsub_rn(spry, sply))) / __dsub_rn(spry, epry);
    sprx = nsx;
    spry = sply;
    const double nex = __dadd_rn(__dmul_rn(sprx, __dsub_rn(eply, epry)), __dmul_rn(eprx, __dsub_rn(spry, eply))) / __dsub_rn(spry, epry);
    eprx = nex;
    epry = eply;
    double ds = splx - sprx, de = eplx - eprx;
    if (fmin(ds, de) / fmax(ds, de) < dispRatio) { ds = -1.0; de = -1.0; }
    if (ds >= minDisp && de >= minDisp && fabs(sply - eply) > horizTh && fabs(spry - epry) > horizTh && overlap > overlapTh) {
        dp[0] = (float)ds;
        dp[1] = (float)de;
        lo[0] = l0; lo[1] = l1; lo[2] = l2;
    }
}

}  // namespace

int plf_launch_match_nnr(plf_ctx* c, const uint8_t* dA, int nA, const uint8_t* dB, int nB, float nnr, int* dOut) {
    if (nA <= 0) return 0;
    match_nnr_kernel<<<(nA + 7) / 8, 256, 0, c->stream>>>(dA, nA, dB, nB, nnr, dOut);
    return 1;
}

int plf_launch_stereo_lines(plf_ctx* c, int slotFirst, int nSlots) {
    const PlfGeom& g = c->g;
    cudaStream_t s = c->stream;
    const int kb = (g.klCap + 127) / 128;
    int n = 0;
    plf_mark(c, "stereo_lines");
    line_grid_kernel<<<dim3(kb, nSlots), 128, 0, s>>>(g, c->d_kl, c->d_nKl, c->d_rowMask, c->d_dirR, slotFirst); ++n;
    line_cand_kernel<<<dim3(kb, g.klCap, nSlots), 128, 0, s>>>(g, c->d_kl, c->d_ldesc, c->d_nKl, c->d_rowMask, c->d_dirR,
                                                              c->d_dmat, c->p.matching_s_ws, c->p.line_sim_th, slotFirst); ++n;
    if (c->p.best_lr_matches) { line_lr_kernel<<<dim3(kb, nSlots), 128, 0, s>>>(g, c->d_nKl, c->d_dmat, c->d_m21, slotFirst); ++n; }
    line_best_kernel<<<dim3((g.klCap + 7) / 8, nSlots), 256, 0, s>>>(g, c->d_nKl, c->d_dmat, c->d_m12, c->p.min_ratio_12_l, slotFirst); ++n;
    line_geom_kernel<<<dim3(kb, nSlots), 128, 0, s>>>(g, c->d_kl, c->d_nKl, c->d_m12, c->d_m21, c->d_disp, c->d_le,
                                                     c->p.best_lr_matches, c->p.min_disp, c->p.line_horiz_th,
                                                     c->p.stereo_overlap_th, c->p.ls_min_disp_ratio, slotFirst); ++n;
    return n;
}

