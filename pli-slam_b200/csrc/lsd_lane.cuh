// K4d  lane-per-image region grower: the scalar LSD region growing, one image per LANE (32 images per warp), included
// by lsd.cu after the helpers it shares with the other growers.
//
// Why: region_grow (OpenCV imgproc/lsd.cpp, reached from Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:227-324)
// is a scalar, order-dependent loop.  The warp-per-image kernel spends ~105 warp instructions per accepted pixel on the
// cross-lane protocol that keeps 32 lanes in the scalar order; with many images in flight the whole path is bound by
// issue slots, so what counts is instructions per pixel, not the latency of one image.  Here every lane simply runs the
// scalar loop of its own image — one list entry (its 8 neighbours) per iteration — so one warp instruction serves 32
// images and there is no protocol at all: exactness is by construction.  The price is latency (one launch lasts as long
// as the scalar loop of its slowest image), which wide batches and several contexts in flight hide.
//
// Per lane: seed cursor, region list appended to the image's arena (d_reg; a region that stays below min_reg_size gives
// its space back), BFS frontier in a shared-memory ring (column = lane: conflict-free), used bitmap and records as in the
// other growers.  Finished regions are entered in a per-image table and their rectangles are fitted afterwards by
// lsd_rect_kernel (one warp per region), in region order = segment order.  refine >= 1 keeps the warp-per-image kernel.

#define LN_RING 128                 // ring entries per lane: 32 x 128 ints = 16 KB of shared memory per warp

__device__ __forceinline__ int ln_dx(int k) { return (int)((0x9224u >> (2 * k)) & 3u) - 1; }   // 0 1 2 0 2 0 1 2 (2 bits each) - 1
__device__ __forceinline__ int ln_dy(int k) { return (int)((0xA940u >> (2 * k)) & 3u) - 1; }   // 0 0 0 1 1 2 2 2

__global__ void __launch_bounds__(32) lsd_grow_lane_kernel(PlfGeom g, const float4* lut, const int* gmap, const int* seeds, const int* nSeeds,
                                                           uint32_t* usedAll, int* regAll, int4* rtAll, int* nRegAll, int* err,
                                                           int imgFirst, int nImg) {
    __shared__ int ring[LN_RING][32];
    const int lane = threadIdx.x, li = blockIdx.x * 32 + lane;
    const bool live = li < nImg;
    const int img = imgFirst + (live ? li : nImg - 1);
    const int W = g.Ws, H = g.Hs, PBW = g.Ps >> 5;
    const int* G = gmap + (size_t)img * g.Ps * H;
    uint32_t* used = usedAll + (size_t)img * PBW * H;
    int* R = regAll + (size_t)img * W * H;
    const int* S = seeds + (size_t)img * g.seedCap;
    int4* RT = rtAll + (size_t)img * g.segCap;
    const int ns = live ? nSeeds[img] : 0;
    const AlignTol tol = make_align_tol(g.prec);
    const int minReg = g.minRegSize, segCap = g.segCap;
    const int a0 = (int)((reinterpret_cast<uintptr_t>(S) >> 2) & 3u);

    int sPos = 0, base = 0, n = 0, i = 0, nReg = 0;
    float sumdx = 0.f, sumdy = 0.f, regDeg = 0.f;
    bool fresh = false, done = ns == 0;

    while (__any_sync(0xffffffffu, !done)) {
        if (!done) {
            if (i < n) {
                // ---- expand list entry i: the 8 neighbours in the scalar loop's order (yy outer, xx inner) ------------------
                const int pk = (n - i <= LN_RING) ? ring[i & (LN_RING - 1)][lane] : R[base + i];
                ++i;
                const int x = pk & 0xFFFF, y = pk >> 16;
                const int wi = x >> 5, sh = x & 31;
                unsigned m = 0;       // bit k: neighbour k is unused (unused implies defined)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    unsigned t = 7u;  // used bits of (x-1, x, x+1); outside the image counts as used
                    if (yy >= 0 && yy < H) {
                        const uint32_t* rowp = used + yy * PBW;
                        const uint32_t w0 = rowp[wi];
                        const uint32_t wp = (sh == 31) ? rowp[wi + 1] : 0u;       // x+1 <= W-1 always: the last column is undefined
                        const uint32_t wl = (sh == 0) ? (x > 0 ? rowp[wi - 1] >> 31 : 1u) : (w0 >> (sh - 1));
                        t = (__funnelshift_r(w0, wp, sh) & 3u) << 1 | (wl & 1u);
                    }
                    const unsigned a = ~t & 7u;
                    if (dy == -1) m |= a;
                    else if (dy == 0) m |= ((a & 1u) << 3) | ((a & 4u) << 2);
                    else m |= a << 5;
                }
                while (m) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1;
                    const int xx = x + ln_dx(k), yy = y + ln_dy(k);
                    const float4 r = lut[G[yy * (PBW << 5) + xx]];
                    if (lsd_aligned(regDeg, r.x, tol)) {
                        if (fresh) {          // the seed enters the sums as (float)cos / sin of its double angle
                            double sn, cs;
                            sincos((double)regDeg * kDegToRad, &sn, &cs);
                            sumdx = (float)cs;
                            sumdy = (float)sn;
                            fresh = false;
                        }
                        const int qb = yy * (PBW << 5) + xx;
                        used[qb >> 5] |= 1u << (qb & 31);
                        const int pk2 = (yy << 16) | xx;
                        R[base + n] = pk2;
                        ring[n & (LN_RING - 1)][lane] = pk2;
                        ++n;
                        sumdx = __fadd_rn(sumdx, r.y);
                        sumdy = __fadd_rn(sumdy, r.z);
                        regDeg = fast_atan2_deg(sumdy, sumdx);
                    }
                }
            } else {
                // ---- region finished (or none yet): enter it in the table, then look for the next unused seed ---------------
                if (n >= minReg) {
                    if (nReg < segCap) {
                        RT[nReg] = make_int4(base, n, __float_as_int(regDeg), 0);
                        ++nReg;
                        base += n;
                    } else {
                        atomicOr(err, 2);
                    }
                }
                n = 0;
                i = 0;
                // up to 8 seeds per turn: two 16-byte groups, the first one containing sPos (a0: the groups are aligned in
                // memory, not in the list, so a seed list that starts off a 16-byte boundary is still read with vector loads)
                const int s0 = ((sPos + a0) & ~3) - a0;
                const int4 va = (s0 < ns) ? *reinterpret_cast<const int4*>(S + s0) : make_int4(0, 0, 0, 0);
                const int4 vb = (s0 + 4 < ns) ? *reinterpret_cast<const int4*>(S + s0 + 4) : make_int4(0, 0, 0, 0);
                const int sv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                unsigned freeMask = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int p = s0 + j;
                    if (p >= sPos && p < ns) {
                        const int qb = (sv[j] >> 16) * (PBW << 5) + (sv[j] & 0xFFFF);
                        if (!((used[qb >> 5] >> (qb & 31)) & 1u)) freeMask |= 1u << j;
                    }
                }
                if (freeMask) {
                    const int j = __ffs(freeMask) - 1;
                    int pk0 = sv[0];
#pragma unroll
                    for (int t = 1; t < 8; ++t) pk0 = (j == t) ? sv[t] : pk0;
                    sPos = s0 + j + 1;
                    const int x = pk0 & 0xFFFF, y = pk0 >> 16;
                    const int qb = y * (PBW << 5) + x;
                    used[qb >> 5] |= 1u << (qb & 31);
                    R[base] = pk0;
                    ring[0][lane] = pk0;
                    n = 1;
                    regDeg = lut[G[qb]].x;
                    fresh = true;
                } else {
                    sPos = s0 + 8;
                    done = sPos >= ns;
                }
            }
        }
    }
    if (live) nRegAll[img] = nReg;
}
