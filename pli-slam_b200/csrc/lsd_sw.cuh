// K4c'''  streaming multi-warp region grower for SMALL batches (included by lsd.cu after the helpers it shares with the
// sequential grower): the exact sequential LSD region growing with SW_NW regions of ONE image in flight, one region per
// WARP, no wave barrier — a warp that has finished a region takes the next seeds at once, a commit pointer walks the seed
// list in order behind them.  Replaces the wave-synchronous lsd_grow_mw_kernel on the product path (a wave of that kernel
// lasts as long as its largest region and warp 0 picks / commits serially while 15 warps idle).
//
// Protocol (exact whatever the heuristics do):
//  * Seeds are handed out in CHUNKS of 32 consecutive list positions (shared scan counter); the warp that takes a chunk
//    grows the regions of its free seeds one after the other with grow_region<2>.  The ticket of a region is its seed
//    position; its tag (position + 1) is what it writes into the owner map O.  O[q]: 0 = undefined pixel, PLF_FREE = defined
//    and unclaimed, else the tag of the region that claims q.  An EARLIER position has priority: claims are
//    atomicMin(O[q], tag).
//  * A region with tag T examining pixel q: O[q] < T -> used (an earlier region's, or undefined; if that region is not
//    committed yet, T RELIES on it and remembers its tag); O[q] == T -> mine; O[q] > T (a later region's, or free) -> not
//    used in the sequential order: tested, and claimed if accepted.  A claim that displaces a later tag marks that region
//    ROBBED (one bit per seed position of the window, shared memory); a claim that finds an earlier tag in place (it came
//    between the read and the atomicMin) marks the claimant itself.  A robbed region notices at its next chain round, gives
//    its pixels back (atomicCAS tag -> PLF_FREE), sets its FAILED bit (global, one per seed position) and may try once more.
//  * Giving a pixel back marks the pixel's own seed position DIRTY (position map P[q], one bit per position of the window)
//    if its chunk was handed out already: whoever skipped that seed because the pixel was taken has to look again.
//  * A finished region leaves a record (tag, size, relied-on tags, fitted segment, pixel list) in its warp's buffer and its
//    segment in the warp's segment queue; the chunk is marked done together with a mask of the positions that need a
//    second look anyway (regions that relied on somebody, seeds that were not started or whose region was given up).
//  * Commit (one warp at a time, whoever is free and finds the chunk at the commit pointer done; shared-memory lock).
//    Fast path — no position of the chunk is robbed, dirty or marked: every record is final, the chunk's segments are
//    appended.  Slow path: the positions are visited in order on FRESH owner values.  Seed held by an earlier tag ->
//    nothing to do (every earlier region is final by now).  Seed held by the position's own tag and a record that was
//    neither robbed nor relied on a failed region -> the record is final.  Anything else — a failed record (pixels given
//    back, failed bit set), a seed that was skipped because a since-failed region held it, a seed that was not started —
//    is grown NOW by the committing warp in `final` mode: it is the earliest region alive, so it wins every contested
//    pixel and relies on nobody.
//  * Exactness: when a record is committed every earlier position is final.  Each pixel it accepted was never claimed by
//    an earlier region (that claim would have robbed it); each pixel it skipped as an earlier region's stayed that
//    region's (a region that passes the commit check never gave a pixel back or lost one; one that does not sets its
//    failed bit before any later record is checked); pixels it skipped as committed were final; pixels it examined and
//    rejected for their angle do not depend on the state.  A seed that was skipped stays skipped unless the pixel was
//    given back (dirty bit).  Stale L1 reads of O can only make a region see a pixel as free or as a later region's when
//    an earlier one holds it (the atomicMin then tells) or rely on a claim that was given back (failed / dirty bit).
//  * Heuristic (efficiency only): a seed that lies on the axis of a region another warp is growing right now, with an
//    aligned level-line angle and within reach of it, will most likely be swallowed by it: it is not started (parked);
//    the commit pointer grows it if it is still free when its turn comes.
// refine = 1 (re-growing with a tolerance from the local angle spread, radius reduction): a region that un-marks pixels would
// have to keep the claims of its first growth in place until it commits — an earlier region that takes a given-back pixel could
// otherwise no longer be noticed — so speculative warps only detect that a region is too sparse and leave it to the committing
// warp, which grows AND refines it in final mode (every earlier region final; given-back pixels marked dirty, failed bit set).
// PLF_SW_FLAGS (environment, experiment switches; results are exact with every combination): 1 owner map read through L1,
// 2 parking heuristic on, 4 print the counters of image 0, 16 no second try after a robbery, 32 only warp 0 works, 64 every
// chunk takes the slow commit path.

#define SW_NW PLF_SW_WARPS
#define SW_HDR 16                   // ints in front of a record's pixel list: tag n flags ndep seg[4] deps[8]
#define SW_SEGQ 128                 // segments a warp can leave uncommitted (shared memory)
#define SW_PARK_DIST 2.5f           // px from a growing region's axis
#define SW_CHDEP 8                  // relied-on tags a chunk can keep in shared memory (more: slow path)
#define SW_AHEAD (SW_WIN / 2)       // chunks the scan pointer may run ahead of the commit pointer: a slot (robbed / dirty / failed
                                    // bits) is reused SW_WIN chunks later, so the failed bit of a region outlives every region
                                    // that could have relied on it

struct SwShared {
    int ring[SW_NW][GROW_RING];
    double sum[SW_NW][3][34];
    uint32_t dep[SW_NW][SW_MAXDEP];
    int chStat[SW_WIN];             // 0 not handed out, 1 being scanned / grown, 2 done
    int chRec[SW_WIN];              // warp << 26 | offset of the chunk's first record in that warp's buffer
    int chSeg[SW_WIN];              // segments of the chunk << 16 | first index in that warp's segment queue
    int chN[SW_WIN];                // records of the chunk
    unsigned robbed[SW_WIN];        // per position: the region lost a pixel
    unsigned dirty[SW_WIN];         // per position: the seed pixel was given back after the chunk was handed out
    unsigned slow[SW_WIN];          // per position: needs a look at commit (not started, given up, too many dependencies)
    unsigned slowRec[SW_WIN];       // per position: a RECORD that cannot be checked in shared memory (too many dependencies)
    unsigned failedW[SW_WIN];       // per position: the region gave pixels back (whoever relied on it must be grown again)
    uint32_t depTag[SW_WIN][SW_CHDEP];   // tags the records of the chunk relied on
    int depN[SW_WIN];
    float4 segS[SW_NW][SW_SEGQ];    // segments of the uncommitted records, in record order
    float4 act[SW_NW];              // region being grown by warp w: seed x, y, cos, sin of its level-line angle
    float actDeg[SW_NW];
    int actN[SW_NW];
    uint32_t actTag[SW_NW];         // 0 = none
    int scanChunk, commitChunk, lock, nSeg;
    int panic;                      // watchdog: a warp found nothing to do for seconds (a protocol stall would otherwise hang the GPU)
    int cnt[16];                    // PLF_SW_FLAGS & 4: statistics
    long long clk[8];
};
#define SW_CNT(i) do { if ((flags & 4) && lane == 0) atomicAdd(&sh.cnt[i], 1); } while (0)

__device__ __forceinline__ uint32_t sw_ld_owner(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ bool sw_failed(const SwShared& sh, uint32_t tag) {
    return (*(volatile const unsigned*)&sh.failedW[((tag - 1u) >> 5) & (SW_WIN - 1)] >> ((tag - 1u) & 31u)) & 1u;
}

// gives the claims of a region back (pixels it still holds), marks their seed positions dirty and tells everybody who
// relied on the region
__device__ __forceinline__ void sw_withdraw(SwShared& sh, uint32_t* O, const int* P, const int* list, int n,
                                            uint32_t tag, int PB, int lane) {
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        int pos = -1;
        if (i < n) {
            const int pk = list[i];
            const int q = (pk >> 16) * PB + (pk & 0xFFFF);
            if (atomicCAS(O + q, tag, PLF_FREE) == tag) pos = P[q];
        }
        __threadfence_block();
        if (pos >= 0 && (pos >> 5) < *(volatile int*)&sh.scanChunk)
            atomicOr(&sh.dirty[(pos >> 5) & (SW_WIN - 1)], 1u << (pos & 31));
    }
    if (lane == 0) atomicOr(&sh.failedW[((tag - 1u) >> 5) & (SW_WIN - 1)], 1u << ((tag - 1u) & 31u));
    __syncwarp();
}

__device__ __forceinline__ void sw_segment(const RectFit& rf, double lsdScale, float seg[4]) {
    const double rr[4] = {rf.x1, rf.y1, rf.x2, rf.y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double v = rr[k] + 0.5;
        if (lsdScale != 1) v /= lsdScale;
        seg[k] = (float)v;
    }
}

template <bool REFINE>
__global__ void __launch_bounds__(32 * SW_NW) lsd_grow_sw_kernel(PlfGeom g, const float4* lut, const int* gmap, const int* seeds,
                                                                const int* nSeeds, const uint32_t* usedAll, uint32_t* ownerAll,
                                                                int* regAll, int* posAll, float* segs,
                                                                int* nSegsOut, int* err, int imgFirst, int flags) {
    extern __shared__ __align__(16) unsigned char sw_smem[];
    SwShared& sh = *reinterpret_cast<SwShared*>(sw_smem);
    const int img = imgFirst + blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t npx = (size_t)g.Ws * g.Hs, npb = (size_t)g.Ps * g.Hs;
    const size_t npxA = (npx + 3) & ~(size_t)3;
    const int warpBuf = PLF_SW_WARPBUF;
    const int warpCap = warpBuf;                                          // record headers and pixel lists
    int* const imgReg = regAll + (size_t)img * (npxA + (size_t)SW_NW * PLF_SW_WARPBUF);   // [npxA] commit buffer, then SW_NW buffers of warpBuf
    int* const Rc = imgReg;
    int* const Rw = imgReg + npxA + (size_t)w * warpBuf;
    float4* const Sq = sh.segS[w];
    uint32_t* const O = ownerAll + (size_t)img * npb;      // scratch by IMAGE index: launches for different images may overlap
    int* const P = posAll + (size_t)img * npb;
    const int* S = seeds + (size_t)img * g.seedCap;
    float4* out = reinterpret_cast<float4*>(segs + (size_t)img * g.segCap * 4);
    const int ns = nSeeds[img];
    const int nChunks = (ns + 31) >> 5;
    const double prec = g.prec;
    const AlignTol precTol = make_align_tol(prec);

    GrowCtx c;
    c.W = g.Ws; c.H = g.Hs; c.PB = g.Ps; c.lane = lane;
    c.LUT = lut;
    c.G = gmap + (size_t)img * npb;
    c.used = nullptr;
    c.owner = O;
    c.ring = sh.ring[w];
    c.invalid = nullptr;
    c.robbed = sh.robbed;
    c.deps = sh.dep[w];
    c.actN = &sh.actN[w];
    c.posMap = P;
    c.dirty = sh.dirty;
    c.failedW = sh.failedW;
    c.scanChunk = &sh.scanChunk;
    c.ldcg = (flags & 1) != 0;
    const long long tStart = clock64();
    {
        const int k = lane & 7;
        c.ddx = (k < 3) ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6));
        c.ddy = (k < 3) ? -1 : (k < 5 ? 0 : 1);
    }

    // ---- owner map of a fresh image from the gradient kernel's bitmap (undefined pixels are pre-marked used), the
    //      position of every defined pixel in the seed list, tables ----
    {
        if (!(flags & 128)) {            // (flag 128: lsd_grad_kernel / lsd_order_kernel have filled both maps already)
            const uint32_t* used = usedAll + (size_t)img * (g.Ps >> 5) * g.Hs;
            const int nWords = (int)(npb >> 5);
            for (int wd = w; wd < nWords; wd += SW_NW) {
                const uint32_t bits = used[wd];
                O[(size_t)wd * 32 + lane] = ((bits >> lane) & 1u) ? 0u : PLF_FREE;
            }
            for (int p = threadIdx.x; p < ns; p += 32 * SW_NW) {
                const int sd = S[p];
                P[(sd >> 16) * c.PB + (sd & 0xFFFF)] = p;
            }
        }
        for (int i = threadIdx.x; i < SW_WIN; i += 32 * SW_NW) { sh.chStat[i] = 0; sh.robbed[i] = 0u; sh.dirty[i] = 0u; sh.slow[i] = 0u; sh.failedW[i] = 0u; sh.slowRec[i] = 0u; sh.depN[i] = 0; }
        if (threadIdx.x < SW_NW) sh.actTag[threadIdx.x] = 0u;
        if (threadIdx.x == 0) { sh.scanChunk = 0; sh.commitChunk = 0; sh.lock = 0; sh.nSeg = 0; sh.panic = 0; }
        if (threadIdx.x < 16) sh.cnt[threadIdx.x] = 0;
        if (threadIdx.x < 8) sh.clk[threadIdx.x] = 0;
    }
    __threadfence_block();
    __syncthreads();
    const long long tInit = clock64() - tStart;

    volatile int* vStat = sh.chStat;
    volatile int* vCommit = &sh.commitChunk;
    volatile uint32_t* vActTag = sh.actTag;
    int head = 0, segHead = 0, lastMine = -1, idle = 0;

    // grows the region of tag `tag` from seed pk0 into `dst` (list at dst + SW_HDR); returns n, ndep (-1: given up / robbed)
    int nrelLast = 0;
    auto grow = [&](int* dst, int room, uint32_t tag, int pk0, bool fin, int& ndep, double& regAngle) -> int {
        c.R = dst + SW_HDR;
        c.relTop = dst + room;
        c.tag = tag;
        c.final = fin;
        c.maxN = fin ? 0x7fffffff : room - SW_HDR;
        c.floorTag = (uint32_t)(*vCommit) * 32u + 1u;
        __threadfence_block();
        const float4 r0 = lut[c.G[(pk0 >> 16) * c.PB + (pk0 & 0xFFFF)]];
        if (lane == 0) {
            sh.act[w] = make_float4((float)(pk0 & 0xFFFF), (float)(pk0 >> 16), r0.y, r0.z);
            sh.actDeg[w] = r0.x;
            sh.actN[w] = 1;
            __threadfence_block();
            vActTag[w] = tag;
        }
        __syncwarp();
        const int n = grow_region<2>(c, pk0, 0, precTol, regAngle, &ndep, &nrelLast);
        if (lane == 0) vActTag[w] = 0u;
        __syncwarp();
        return n;
    };

    while (true) {
        int cc = *vCommit;
        if (cc >= nChunks) break;
        if (*(volatile int*)&sh.panic) break;
        // ---- 1. commit: whoever finds the chunk at the commit pointer done, one warp at a time ----
        if (vStat[cc & (SW_WIN - 1)] >= 2) {
            int got = 0;
            if (lane == 0) got = atomicCAS(&sh.lock, 0, 1) == 0;
            got = __shfl_sync(0xffffffffu, got, 0);
            if (got) {
                const long long tc0 = clock64();
                __threadfence_block();
                int nSeg = *(volatile int*)&sh.nSeg;
                while (true) {
                    const long long ta0 = clock64();
                    cc = *vCommit;
                    if (cc >= nChunks) break;
                    const int slot = cc & (SW_WIN - 1);
                    if (vStat[slot] < 2) break;
                    __threadfence_block();
                    if (!(flags & 64)) {
                        // fast path for a RUN of chunks, one per lane: done, nothing robbed / dirty / given up, no relied-on
                        // region failed -> every record of the run is final (a fast commit changes nothing, and no earlier
                        // region is alive that could still touch the later chunks of the run); the segments are appended in
                        // order and the run is published with one fence
                        const int cj = cc + lane, sj = cj & (SW_WIN - 1);
                        bool okj = cj < nChunks && lane < SW_AHEAD && vStat[sj] >= 2;
                        __threadfence_block();
                        if (okj) okj = (*(volatile unsigned*)&sh.robbed[sj] | *(volatile unsigned*)&sh.dirty[sj] | *(volatile unsigned*)&sh.slow[sj]) == 0u;
                        if (okj) {
                            const int dn = *(volatile int*)&sh.depN[sj];
                            for (int k = 0; k < dn; ++k)
                                if (sw_failed(sh, *(volatile uint32_t*)&sh.depTag[sj][k])) okj = false;
                        }
                        const unsigned fm = __ballot_sync(0xffffffffu, okj);
                        const int nFast = (fm == 0xffffffffu) ? 32 : __ffs(~fm) - 1;
                        if (nFast > 0) {
                            const int segPackJ = lane < nFast ? *(volatile int*)&sh.chSeg[sj] : 0;
                            const int recWarpJ = lane < nFast ? (*(volatile int*)&sh.chRec[sj] >> 26) : 0;
                            const int kj = segPackJ >> 16;
                            int incl = kj;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += t;
                            }
                            const int excl = incl - kj, total = __shfl_sync(0xffffffffu, incl, 31);
                            unsigned sm = __ballot_sync(0xffffffffu, kj > 0);
                            while (sm) {
                                const int j = __ffs(sm) - 1;
                                sm &= sm - 1u;
                                const int k = __shfl_sync(0xffffffffu, kj, j), off = __shfl_sync(0xffffffffu, excl, j);
                                const float4* src = sh.segS[__shfl_sync(0xffffffffu, recWarpJ, j)] + (__shfl_sync(0xffffffffu, segPackJ, j) & 0xFFFF);
                                const int dstI = nSeg + off + lane;
                                if (lane < k && dstI < g.segCap) out[dstI] = src[lane];
                            }
                            if (nSeg + total > g.segCap && lane == 0) atomicOr(err, 2);
                            nSeg = min(g.segCap, nSeg + total);
                            if ((flags & 4) && lane == 0) atomicAdd(&sh.cnt[5], nFast);
                            if (lane < nFast) sh.chStat[sj] = 0;
                            __syncwarp();
                            if (lane == 0) {
                                __threadfence_block();
                                *vCommit = cc + nFast;
                            }
                            __syncwarp();
                            continue;
                        }
                    }
                    const long long ta1 = clock64();
                    const int recPack = *(volatile int*)&sh.chRec[slot];
                    const int nRec = *(volatile int*)&sh.chN[slot];
                    const long long tq0 = clock64();
                    bool hardDep = false;
                    unsigned bad = *(volatile unsigned*)&sh.robbed[slot] | *(volatile unsigned*)&sh.dirty[slot] |
                                   *(volatile unsigned*)&sh.slow[slot] | ((flags & 64) ? 1u : 0u);
                    {
                        // did a region the chunk's records relied on give pixels back?
                        const int dn = *(volatile int*)&sh.depN[slot];
                        const bool fd = lane < dn && sw_failed(sh, *(volatile uint32_t*)&sh.depTag[slot][lane]);
                        if (dn > 0 && __any_sync(0xffffffffu, fd)) { bad |= 1u; hardDep = true; SW_CNT(13); }
                    }
                    if (bad != 0u && !(flags & 64) && *(volatile unsigned*)&sh.robbed[slot] == 0u && !hardDep) {
                        // only seeds that were not started / given up / given back: if every one of them belongs to an
                        // earlier region by now there is nothing to grow and the records stand as they are
                        const unsigned quick = *(volatile unsigned*)&sh.dirty[slot] | *(volatile unsigned*)&sh.slow[slot];
                        const int p = cc * 32 + lane;
                        bool open = false;
                        if ((quick >> lane) & 1u) {
                            const int seed = p < ns ? S[p] : -1;
                            open = seed >= 0 && sw_ld_owner(O + (seed >> 16) * c.PB + (seed & 0xFFFF)) >= (uint32_t)p + 1u;
                        }
                        if (!__any_sync(0xffffffffu, open) && !(*(volatile unsigned*)&sh.slowRec[slot])) { bad = 0u; SW_CNT(14); }
                    }
                    const long long tq1 = clock64();
                    if ((flags & 4) && lane == 0 && bad == 0u && tq1 - tq0 > 400) atomicAdd((unsigned long long*)&sh.clk[3], (unsigned long long)(tq1 - tq0));
                    if (bad == 0u) {
                        // fast path: every record of the chunk is final, every other seed is still an earlier region's
                        const int segPack = *(volatile int*)&sh.chSeg[slot];
                        const int k = segPack >> 16;
                        if (k > 0) {
                            const float4* src = sh.segS[recPack >> 26] + (segPack & 0xFFFF);
                            const int room = g.segCap - nSeg;
                            if (lane < k && lane < room) out[nSeg + lane] = src[lane];
                            if (k > room && lane == 0) atomicOr(err, 2);
                            nSeg += min(k, room);
                        }
                        SW_CNT(5);
                    } else {
                        SW_CNT(6);
                        const long long ts0 = clock64();
                        const int p = cc * 32 + lane;
                        const int seed = p < ns ? S[p] : -1;
                        const int pb = seed >= 0 ? (seed >> 16) * c.PB + (seed & 0xFFFF) : 0;
                        const uint32_t myTag = (uint32_t)p + 1u;
                        const int* rec = imgReg + npxA + (size_t)(recPack >> 26) * warpBuf + (recPack & ((1 << 26) - 1));
                        int recIdx = 0, cur = 0;
                        while (true) {
                            const uint32_t o = (lane >= cur && seed >= 0) ? sw_ld_owner(O + pb) : 0u;
                            const unsigned cand = __ballot_sync(0xffffffffu, o >= myTag);      // its own tag, a later one, or free
                            const int l = cand ? __ffs(cand) - 1 : 32;
                            const uint32_t tagStar = cand ? (uint32_t)(cc * 32 + l) + 1u : 0xFFFFFFFFu;
                            // strictly in position order: a record in front of the first candidate lost its seed to an earlier
                            // region; its pixels go back BEFORE anything behind it is looked at (they may free a seed there)
                            if (recIdx < nRec && (uint32_t)rec[0] < tagStar) {
                                const uint32_t dead = (uint32_t)rec[0];
                                sw_withdraw(sh, O, P, rec + SW_HDR, rec[1], dead, c.PB, lane);
                                rec += (SW_HDR + rec[1] + (rec[2] >> 8) + 3) & ~3;
                                ++recIdx;
                                cur = (int)((dead - 1u) & 31u) + 1;
                                continue;
                            }
                            if (!cand) break;
                            const uint32_t oStar = __shfl_sync(0xffffffffu, o, l);
                            const int pkStar = __shfl_sync(0xffffffffu, seed, l);
                            bool done = false;
                            if (recIdx < nRec && (uint32_t)rec[0] == tagStar) {
                                const int n = rec[1], rflags = rec[2], nd = rec[3], nrel = rec[2] >> 8;
                                bool ok = oStar == tagStar && !((*(volatile unsigned*)&sh.robbed[slot] >> l) & 1u);
                                if (ok && nd > 0) {
                                    bool failedDep = nd > SW_MAXDEP;
                                    if (lane < nd && lane < SW_MAXDEP) failedDep = sw_failed(sh, (uint32_t)rec[8 + lane]);
                                    if (__any_sync(0xffffffffu, failedDep)) {
                                        // a region it relied on gave pixels back (or there were too many to remember): what
                                        // counts is whether every pixel it skipped as that region's is an earlier region's
                                        // NOW — all of them are final
                                        ok = (rflags & 2) != 0;
                                        for (int i0 = 0; ok && i0 < nrel; i0 += 32) {
                                            bool open = false;
                                            if (i0 + lane < nrel) open = sw_ld_owner(O + rec[SW_HDR + n + i0 + lane]) >= tagStar;
                                            ok = !__any_sync(0xffffffffu, open);
                                        }
                                        if (ok) SW_CNT(15);
                                    }
                                }
                                if (ok) {
                                    if (rflags & 1) {
                                        if (nSeg < g.segCap) {
                                            if (lane == 0) out[nSeg] = *reinterpret_cast<const float4*>(rec + 4);
                                            ++nSeg;
                                        } else if (lane == 0) atomicOr(err, 2);
                                    }
                                    done = true;
                                } else {
                                    SW_CNT(2);
                                    sw_withdraw(sh, O, P, rec + SW_HDR, n, tagStar, c.PB, lane);
                                }
                                rec += (SW_HDR + n + nrel + 3) & ~3;
                                ++recIdx;
                            }
                            if (!done) {
                                // grown now, as the earliest region alive
                                SW_CNT(3);
                                const long long tg0 = clock64();
                                if (lane == 0) atomicAnd(&sh.robbed[slot], ~(1u << l));
                                __syncwarp();
                                int nd;
                                double regAngle;
                                int n = grow(Rc, 0, tagStar, pkStar, true, nd, regAngle);
                                bool keep = n >= g.minRegSize;
                                RectFit rf;
                                if (keep) {
                                    rect_fit<REFINE>(c, sh.sum[w], n, regAngle, prec, rf);
                                    // refine = 1: re-growing and radius reduction happen here only, where every earlier region
                                    // is final; pixels the region gives back are marked dirty, its failed bit tells whoever relied on it
                                    if (REFINE) keep = lsd_refine<2>(c, sh.sum[w], n, regAngle, prec, g.densityTh, rf);
                                }
                                if (keep) {
                                    float sg[4];
                                    sw_segment(rf, g.lsdScale, sg);
                                    if (nSeg < g.segCap) {
                                        if (lane == 0) out[nSeg] = make_float4(sg[0], sg[1], sg[2], sg[3]);
                                        ++nSeg;
                                    } else if (lane == 0) atomicOr(err, 2);
                                }
                                if ((flags & 4) && lane == 0) { atomicAdd((unsigned long long*)&sh.clk[1], (unsigned long long)(clock64() - tg0)); atomicAdd(&sh.cnt[8], n); }
                            }
                            cur = l + 1;
                        }
                        if ((flags & 4) && lane == 0) atomicAdd((unsigned long long*)&sh.clk[2], (unsigned long long)(clock64() - ts0));
                    }
                    const long long tp0 = clock64();
                    if (lane == 0) {
                        sh.chStat[slot] = 0;
                        __threadfence_block();
                        *vCommit = cc + 1;
                    }
                    __syncwarp();
                    if ((flags & 4) && lane == 0) {
                        atomicAdd((unsigned long long*)&sh.clk[4], (unsigned long long)(ta1 - ta0));
                        atomicAdd((unsigned long long*)&sh.clk[5], (unsigned long long)(tq0 - ta1));
                        atomicAdd((unsigned long long*)&sh.clk[6], (unsigned long long)(clock64() - tp0));
                        atomicAdd((unsigned long long*)&sh.clk[7], (unsigned long long)(tp0 - tq1));
                    }
                }
                if ((flags & 4) && lane == 0) atomicAdd((unsigned long long*)&sh.clk[0], (unsigned long long)(clock64() - tc0));
                if (lane == 0) {
                    *(volatile int*)&sh.nSeg = nSeg;
                    __threadfence_block();
                    atomicExch(&sh.lock, 0);
                }
                __syncwarp();
                idle = 0;
                continue;
            }
        }
        // ---- 2. take the next chunk of seeds ----
        if ((head > 0 || segHead > 0) && cc > lastMine) { head = 0; segHead = 0; }      // everything this warp left behind is committed
        int ch = -1;
        if ((flags & 32) && w > 0) { __nanosleep(1000); continue; }
        if (lane == 0 && head + SW_HDR + 4096 < warpCap && segHead + 32 <= SW_SEGQ) {
            const int s = *(volatile int*)&sh.scanChunk;
            if (s < nChunks && s < *vCommit + SW_AHEAD && atomicCAS(&sh.scanChunk, s, s + 1) == s) ch = s;
        }
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch < 0) {
            // nothing to take (window full, buffer full, or the list is handed out): sleep, longer every time — a spinning
            // warp costs the growing ones of its SM issue slots (a quarter of the kernel's instructions before the back-off)
            SW_CNT(7);
            __nanosleep(idle < 3 ? 250 : (idle < 8 ? 1000 : 3000));
            if (++idle > 6000000) {                // ~18 s without anything to take or commit: give up loudly (flag 8), never hang
                if (lane == 0) { atomicOr(err, 8); atomicExch(&sh.panic, 1); }
                __syncwarp();
            }
            continue;
        }
        idle = 0;
        const int slot = ch & (SW_WIN - 1);
        if (lane == 0) { atomicExch(&sh.robbed[slot], 0u); atomicExch(&sh.dirty[slot], 0u); atomicExch(&sh.failedW[slot], 0u); sh.depN[slot] = 0; sh.chStat[slot] = 1; }
        __syncwarp();
        __threadfence_block();
        const int p = ch * 32 + lane;
        const int seed = p < ns ? S[p] : -1;
        const int pb = seed >= 0 ? (seed >> 16) * c.PB + (seed & 0xFFFF) : 0;
        const uint32_t myTag = (uint32_t)p + 1u;
        const uint32_t o0 = seed >= 0 ? sw_ld_owner(O + pb) : 0u;
        unsigned gm = __ballot_sync(0xffffffffu, seed >= 0 && o0 > myTag);
        const int recStart = head, segStart = segHead;
        int nRec = 0, chDep = 0;
        unsigned slowMask = 0u, slowRecMask = 0u;
        while (gm) {
            const int l = __ffs(gm) - 1;
            gm &= gm - 1u;
            const int pk0 = __shfl_sync(0xffffffffu, seed, l);
            const uint32_t tag = (uint32_t)(ch * 32 + l) + 1u;
            const int qb = (pk0 >> 16) * c.PB + (pk0 & 0xFFFF);
            if (sw_ld_owner(O + qb) < tag) continue;              // taken meanwhile (mostly by the region just grown); a give-back marks it dirty
            if (head + SW_HDR + 4096 >= warpCap) { slowMask |= 1u << l; continue; }      // no room: left to the commit pointer
            // parked?  on the axis of a region another warp is growing, with an aligned angle, within its reach
            if (flags & 2) {
                const float4 r0 = lut[c.G[qb]];
                bool park = false;
                if (lane < SW_NW && lane != w) {
                    const uint32_t at = vActTag[lane];
                    if (at != 0u && at < tag) {
                        const float4 a = sh.act[lane];
                        float d = fabsf(r0.x - sh.actDeg[lane]);
                        if (d > 180.f) d = 360.f - d;
                        const float dx = (float)(pk0 & 0xFFFF) - a.x, dy = (float)(pk0 >> 16) - a.y;
                        const float reach = 24.f + 0.5f * (float)*(volatile int*)&sh.actN[lane];
                        park = d < 22.5f && fabsf(dx * a.w - dy * a.z) < SW_PARK_DIST && fabsf(dx * a.z + dy * a.w) < reach;
                    }
                }
                if (__any_sync(0xffffffffu, park)) { SW_CNT(4); slowMask |= 1u << l; continue; }      // (measured: parking leaves too much to the committing warp; off unless PLF_SW_FLAGS & 2)
            }
            int nd = -1, n = 0;
            double regAngle;
            for (int attempt = 0; attempt < 3; ++attempt) {
                if (attempt > 0) {
                    // robbed while growing (mostly at a junction of two edges): whoever took the pixels needs a few more steps
                    // in this neighbourhood; wait a little, then try again
                    const long long tw = clock64() + (attempt == 1 ? 6000 : 24000);
                    while (clock64() < tw) __nanosleep(200);
                    if (lane == 0) atomicAnd(&sh.robbed[slot], ~(1u << l));
                    __syncwarp();
                    __threadfence_block();
                    if (sw_ld_owner(O + qb) < tag) break;
                    SW_CNT(11);
                }
                n = grow(Rw + head, warpCap - head, tag, pk0, false, nd, regAngle);
                const bool robbedNow = grow_is_invalid<2>(c);
                if (nd >= 0 && !robbedNow) break;
                SW_CNT(1);
                if ((flags & 4) && lane == 0) atomicAdd(&sh.cnt[9], n);
                sw_withdraw(sh, O, P, Rw + head + SW_HDR, n, tag, c.PB, lane);
                nd = -1;
                if ((flags & 16) || !robbedNow) break;        // (not robbed: out of room — left to the committing warp, which has room for any region)
            }
            if (nd < 0) {
                // given up: no record, no claim left; the robbed bit has served (a quick look at the seed is all commit needs)
                if (lane == 0) atomicAnd(&sh.robbed[slot], ~(1u << l));
                slowMask |= 1u << l;
                continue;
            }
            int rflags = 0;
            float sg[4] = {0.f, 0.f, 0.f, 0.f};
            if (n >= g.minRegSize) {
                RectFit rf;
                rect_fit<REFINE>(c, sh.sum[w], n, regAngle, prec, rf);
                if (REFINE && (double)n / __dmul_rn(lsd_dist(rf.x1, rf.y1, rf.x2, rf.y2), rf.width) < g.densityTh) {
                    // too sparse: refine() would re-grow it and give pixels back — only the committing warp does that
                    SW_CNT(1);
                    sw_withdraw(sh, O, P, Rw + head + SW_HDR, n, tag, c.PB, lane);
                    slowMask |= 1u << l;
                    continue;
                }
                sw_segment(rf, g.lsdScale, sg);
                rflags = 1;
                if (lane == 0) Sq[segHead] = make_float4(sg[0], sg[1], sg[2], sg[3]);
                ++segHead;
            }
            // the relied-on pixels move from the top of the free space to the end of the pixel list (n + 2 nrel fits: no overlap)
            const int nrel = (nd > 0 && nrelLast <= SW_MAXREL) ? nrelLast : 0;
            for (int i = lane; i < nrel; i += 32) Rw[head + SW_HDR + n + i] = Rw[warpCap - 1 - i];
            const bool relKnown = nd == 0 || nrelLast <= SW_MAXREL;
            if (nd > 0) {
                SW_CNT(12);
                if (nd <= SW_MAXDEP && chDep + nd <= SW_CHDEP) {
                    if (lane < nd) sh.depTag[slot][chDep + lane] = c.deps[lane];
                    chDep += nd;
                } else { slowMask |= 1u << l; slowRecMask |= 1u << l; }
            }
            int* hdr = Rw + head;
            if (lane == 0) {
                hdr[0] = (int)tag; hdr[1] = n; hdr[2] = rflags | (relKnown ? 2 : 0) | (nrel << 8); hdr[3] = nd;
                hdr[4] = __float_as_int(sg[0]); hdr[5] = __float_as_int(sg[1]); hdr[6] = __float_as_int(sg[2]); hdr[7] = __float_as_int(sg[3]);
            }
            if (lane < SW_MAXDEP) hdr[8 + lane] = (lane < nd) ? (int)c.deps[lane] : 0;
            head += (SW_HDR + n + nrel + 3) & ~3;
            ++nRec;
            SW_CNT(0);
            if ((flags & 4) && lane == 0) atomicAdd(&sh.cnt[10], n);
            __syncwarp();
        }
        if (lane == 0) {
            sh.chRec[slot] = (w << 26) | recStart;
            sh.chSeg[slot] = ((segHead - segStart) << 16) | segStart;
            sh.chN[slot] = nRec;
            sh.slow[slot] = slowMask;
            sh.slowRec[slot] = slowRecMask;
            sh.depN[slot] = chDep;
            __threadfence_block();
            sh.chStat[slot] = 2;
        }
        lastMine = ch;
        __syncwarp();
    }
    __syncthreads();
    if ((flags & 4) && threadIdx.x == 0 && blockIdx.x == 0)
        printf("sw img %d: ns %d chunks %d | recorded %d (%d px, %d with deps) given up %d (%d px, %d retried) failedRec %d (%d more saved by the pixel check) grownAtCommit %d (%d px) parked %d | chunks fast %d (%d after a look at the seeds) slow %d (failed dep %d) | idle spins %d | cycles total %lld commit %lld (regrow %lld, slow path incl. regrow %lld, seed looks %lld) segs %d\n",
               img, ns, nChunks, sh.cnt[0], sh.cnt[10], sh.cnt[12], sh.cnt[1], sh.cnt[9], sh.cnt[11], sh.cnt[2], sh.cnt[15], sh.cnt[3], sh.cnt[8], sh.cnt[4], sh.cnt[5], sh.cnt[14], sh.cnt[6], sh.cnt[13], sh.cnt[7],
               clock64() - tStart, sh.clk[0], sh.clk[1], sh.clk[2], sh.clk[3], sh.nSeg);
    if ((flags & 4) && threadIdx.x == 0 && blockIdx.x == 0)
        printf("sw init (owner map, position map): %lld cycles\n", tInit);
    if ((flags & 4) && threadIdx.x == 0 && blockIdx.x == 0)
        printf("sw commit split: acquire+fence %lld, table reads %lld, body (fast or slow) %lld, publish %lld\n", sh.clk[4], sh.clk[5], sh.clk[7], sh.clk[6]);
    if (threadIdx.x == 0) nSegsOut[img] = min(sh.nSeg, g.segCap);
}
