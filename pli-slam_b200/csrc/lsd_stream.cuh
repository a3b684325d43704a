// K4c''  streaming region grower: the exact sequential LSD region growing with up to 32 regions of ONE image in flight,
// one region per LANE of a warp (included by lsd.cu after the helpers it shares with the sequential grower).
//
// Why: region_grow is a scalar, order-dependent loop (every acceptance moves the region angle the next test uses).  The
// sequential kernel spends a whole warp on one region and finds 4-8 useful lanes per memory round trip; here every lane
// runs that scalar loop for its own region — one list entry (8 neighbours) per step — so a warp instruction serves up to
// 32 regions.  The lock-step emulation of this protocol (oracle/cpp/lsd.cpp: lsd_stream_sim, pinned to cv2 through
// lsd_detect) measured 5.4-8.6 k steps per 752x480 image at 70-90 % lane occupancy against 124-181 k list entries.
//
// Protocol (exact whatever the heuristics do):
//  * Candidates = defined pixels in seed order; the ticket of a candidate is its position + 1; an EARLIER position has
//    priority.  Owner map O[q]: 0 free, 0xFFFFFFFF committed or undefined, else the ticket that claims q.
//  * A region with ticket T visiting neighbour q: committed or mine -> skip.  Free -> claim if aligned.  Claimed by a
//    LATER ticket -> steal if aligned: the victim is killed (claims withdrawn, candidate queued again).  Claimed by an
//    EARLIER in-flight ticket E -> skip, and if it would have been aligned remember "T relies on E".
//  * All lanes decide on the owner values loaded at the top of the step, so two regions can claim one pixel in the same
//    step.  Every lane therefore re-reads, with the loads of its NEXT step, the pixels it claimed: still mine -> fine;
//    an earlier ticket's -> I was robbed, I die; a later ticket's -> I take the pixel back, that ticket dies, and I sit
//    out one step to look at the pixel once more (a third region may have acted on the wrong owner meanwhile).  The
//    earliest ticket in flight therefore never dies.  A region that has expanded its last entry lives one more step
//    for this check; a seed is claimed on a fresh read.
//  * A ticket that relied on a killed ticket is itself killed — found out lazily through hashed kill stamps, while it
//    grows or when it is about to commit (collisions only kill more).
//  * The commit pointer walks the candidates in order.  Committed pixel -> next.  Pixel held by the candidate's own,
//    finished, still valid ticket -> commit (pixels -> committed, list appended to the region arena when it is large
//    enough for a rectangle).  Otherwise the candidate needs its region NOW: lane 0 is reserved for it.  The earliest
//    ticket in flight can neither be robbed nor rely on anybody, so the walk always advances.
//  * Exactness: when a ticket commits, every earlier candidate is resolved; each pixel it accepted was never taken by an
//    earlier region (that would have been a steal), each pixel it skipped as an earlier ticket's stayed that ticket's
//    (or the ticket was killed and so was this one), committed pixels were committed by earlier candidates only.
//  * Heuristic (efficiency only): a candidate that lies on the axis of a growing region with an aligned angle will most
//    likely be swallowed by it; it is parked until that region has finished.
//
// Lists are chains of 32-int chunks (slot 0 = link, 31 pixels) from a per-image arena with a free stack; everything of
// an image is owned by its single warp, so there is no inter-warp synchronisation and no atomic anywhere.  The
// rectangles are fitted afterwards by lsd_rect_kernel, one warp per committed region.
// refine >= 1 (re-growing with feedback from the rectangle) keeps the sequential kernel.

#define ST_FREE 0u
#define ST_COMMITTED 0xFFFFFFFFu
#define ST_SETS 512                 // ticket table (finished, uncommitted regions): ST_SETS x ST_WAYS, global memory
#define ST_WAYS 8
#define ST_TICKETS (ST_SETS * ST_WAYS)
#define ST_TKW 12                   // ints per ticket record: tag first n regDeg epoch ndep dep0..3 - -
#define ST_BLOCKED 1024             // parked candidates (shared memory)
#define ST_READY 1024               // released candidates (shared memory)
#define ST_STAMPS 2048              // hashed kill stamps (shared memory)
#define ST_KILLQ 128
#define ST_RELQ 64
#define ST_FREEC 96                 // shared-memory cache of free chunk ids in front of the global free stack
#define ST_WINDOW 65536             // how far the scan may run ahead of the commit pointer (positions)
#define ST_DPERP2 4.0f              // (2 px)^2: distance to a growing region's axis below which a candidate is parked

struct StreamLayout {               // offsets in ints inside one image's scratch block
    int O, CH, FS, TK, RT, nChunks, total;
};

__host__ __device__ inline StreamLayout stream_layout(int Ps, int Ws, int Hs, int segCap) {
    StreamLayout L;
    int o = 0;
    L.nChunks = ((Ws * Hs + 30) / 31 + ST_TICKETS + 256 + 31) & ~31;
    L.O = o;  o += (Ps * Hs + 31) & ~31;          // every block starts on a 128-byte boundary (int4 / uint4 accesses)
    L.CH = o; o += L.nChunks * 32;
    L.FS = o; o += L.nChunks;
    L.TK = o; o += (ST_TICKETS * ST_TKW + 31) & ~31;
    L.RT = o; o += (segCap * 4 + 31) & ~31;
    L.total = (o + 31) & ~31;
    return L;
}

// owner map of a fresh image: defined pixels free, everything else committed; ticket table cleared
__global__ void __launch_bounds__(256) lsd_stream_init_kernel(PlfGeom g, const int* n2map, int* scratch, StreamLayout L, int imgFirst) {
    const int img = imgFirst + blockIdx.y;
    const int* N2 = n2map + (size_t)img * g.Ps * g.Hs;
    int* base = scratch + (size_t)blockIdx.y * L.total;
    const int nO = g.Ps * g.Hs;
    const int stride = gridDim.x * blockDim.x * 4;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < nO; i += stride) {
        const int4 v = *reinterpret_cast<const int4*>(N2 + i);
        uint4 o;
        o.x = v.x ? ST_FREE : ST_COMMITTED; o.y = v.y ? ST_FREE : ST_COMMITTED;
        o.z = v.z ? ST_FREE : ST_COMMITTED; o.w = v.w ? ST_FREE : ST_COMMITTED;
        *reinterpret_cast<uint4*>(base + L.O + i) = o;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ST_TICKETS * ST_TKW; i += gridDim.x * blockDim.x) base[L.TK + i] = 0;
}

struct StreamShared {
    int ready[ST_READY];
    int2 blocked[ST_BLOCKED];
    int stamp[ST_STAMPS];
    uint32_t killQ[ST_KILLQ];
    uint32_t relQ[ST_RELQ];
    uint32_t victim[9][32];
    int freeC[ST_FREEC];
};

struct StreamCtx {                  // per-image pointers and the warp-uniform bookkeeping
    uint32_t* O; int* CH; int* FS; int* TK; int4* RT; int* RF;
    const float4* LUT; const int* G; const int* S;
    StreamShared* sh;
    int W, H, PB, ns, nChunks, minReg, segCap, lane;
    int cp, scanPos, nBlocked, nReady, bump, stackTop, nFreeC, killEpoch, live, nReg, rfPos, err, nKill, nRel;
};

struct StreamLane {                 // the region in this lane (tag == 0: idle; i >= n: closing, see the header)
    uint32_t tag;
    int curPk, lastPk, claimMask, i, n, first, rChunk, rOff, wChunk, wOff, spare, sx, sy, nBlockedByMe, epoch, ndep;
    uint32_t dep0, dep1, dep2, dep3;
    float sumdx, sumdy, regDeg;
};

__device__ __forceinline__ int st_dx(int k) { return (k < 3) ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6)); }
__device__ __forceinline__ int st_dy(int k) { return (k < 3) ? -1 : (k < 5 ? 0 : 1); }

// ---- chunk arena: shared-memory cache of free ids in front of the global stack, untouched chunks by bump pointer -------
__device__ __forceinline__ void st_free_one(StreamCtx& c, int ch) {      // uniform ch
    if (c.nFreeC == ST_FREEC) {                                          // cache full: spill 32 ids to the global stack
        __syncwarp();
        c.FS[c.stackTop + c.lane] = c.sh->freeC[ST_FREEC - 32 + c.lane];
        c.stackTop += 32; c.nFreeC -= 32;
        __syncwarp();
    }
    if (c.lane == 0) c.sh->freeC[c.nFreeC] = ch;
    c.nFreeC++;
}
// one chunk for every lane of `need`; returns the lane's chunk (or -1)
__device__ __forceinline__ int st_alloc(StreamCtx& c, unsigned need) {
    const int cnt = __popc(need);
    __syncwarp();
    if (c.nFreeC < cnt) {                                                // refill: recycled chunks first (warm), then untouched ones
        const int want = min(32, ST_FREEC - c.nFreeC);
        const int fromStack = min(want, c.stackTop);
        if (c.lane < fromStack) c.sh->freeC[c.nFreeC + c.lane] = c.FS[c.stackTop - 1 - c.lane];
        else if (c.lane < want) c.sh->freeC[c.nFreeC + c.lane] = c.bump + (c.lane - fromStack);
        c.stackTop -= fromStack;
        c.bump += want - fromStack;
        c.nFreeC += want;
        if (c.bump > c.nChunks) c.err |= 16;
        __syncwarp();
    }
    const int rank = __popc(need & ((1u << c.lane) - 1u));
    const int ch = ((need >> c.lane) & 1u) ? c.sh->freeC[c.nFreeC - 1 - rank] : -1;
    c.nFreeC -= cnt;
    __syncwarp();
    return ch;
}

__device__ __forceinline__ void st_ready_push(StreamCtx& c, int pos) {                // uniform pos
    if (c.nReady < ST_READY) { if (c.lane == 0) c.sh->ready[c.nReady] = pos; c.nReady++; }
    else c.scanPos = min(c.scanPos, pos);
}
__device__ __forceinline__ void st_kill_request(StreamCtx& c, uint32_t v) {           // uniform v: killed at the end of the step
    if (c.nKill < ST_KILLQ) { if (c.lane == 0) c.sh->killQ[c.nKill] = v; c.nKill++; }
    else c.err |= 512;
}
__device__ __forceinline__ void st_release_request(StreamCtx& c, uint32_t tag) {      // uniform tag: its parked candidates go at the end of the step
    if (c.nRel < ST_RELQ) { if (c.lane == 0) c.sh->relQ[c.nRel] = tag; c.nRel++; }
    else c.err |= 1024;
}

// candidates parked on a region of the release queue (finished or killed) may go now
__device__ __forceinline__ void st_release_blocked(StreamCtx& c) {
    int kept = 0;
    __syncwarp();
    for (int b0 = 0; b0 < c.nBlocked; b0 += 32) {
        const int b = b0 + c.lane;
        int2 e = make_int2(-1, 0);
        if (b < c.nBlocked) e = c.sh->blocked[b];
        bool rel = false;
        for (int r = 0; r < c.nRel; ++r) rel = rel || (uint32_t)e.y == c.sh->relQ[r];
        rel = rel && b < c.nBlocked;
        const bool keep = b < c.nBlocked && !rel;
        const unsigned rm = __ballot_sync(0xffffffffu, rel), km = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) c.sh->blocked[kept + __popc(km & ((1u << c.lane) - 1u))] = e;      // compaction in place: kept <= b0
        kept += __popc(km);
        for (unsigned m = rm; m; m &= m - 1u) st_ready_push(c, __shfl_sync(0xffffffffu, e.x, __ffs(m) - 1));
        __syncwarp();
    }
    c.nBlocked = kept;
    c.nRel = 0;
}

// walks a chunk chain of n pixels: MODE 0 commit (pixels -> committed; copied to the region arena at dst when dst >= 0),
// MODE 1 withdraw (pixels still holding `tag` -> free).  Chunks go back to the free cache.
template <int MODE>
__device__ __forceinline__ void st_walk(StreamCtx& c, int first, int n, uint32_t tag, int dst) {
    int ch = first;
    __syncwarp();
#pragma unroll 1
    for (int done = 0; done < n; done += 31) {
        int v = 0;
        if (c.lane == 0 || done + c.lane - 1 < n) v = c.CH[ch * 32 + c.lane];
        if (c.lane > 0 && done + c.lane - 1 < n) {
            const int q = (v >> 16) * c.PB + (v & 0xFFFF);
            if (MODE == 0) {
                c.O[q] = ST_COMMITTED;
                if (dst >= 0) c.RF[dst + done + c.lane - 1] = v;
            } else if (c.O[q] == tag) c.O[q] = ST_FREE;
        }
        const int link = __shfl_sync(0xffffffffu, v, 0);
        st_free_one(c, ch);
        ch = link;
    }
    __syncwarp();
}

__device__ __forceinline__ void st_lane_reset(StreamLane& a) {
    a.tag = 0u; a.nBlockedByMe = 0; a.spare = -1; a.claimMask = 0; a.i = 0; a.n = 0;
}

// finished tickets live in the global table: record index of `tag` (uniform) or -1
__device__ __forceinline__ int st_find(const StreamCtx& c, uint32_t tag) {
    const int set = (int)(tag & (ST_SETS - 1)) * ST_WAYS;
    const uint32_t t = c.lane < ST_WAYS ? (uint32_t)c.TK[(set + c.lane) * ST_TKW] : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, c.lane < ST_WAYS && t == tag);
    return m ? set + __ffs(m) - 1 : -1;
}

// has a ticket this one relied on been killed since it started?
__device__ __forceinline__ bool st_dep_broken(const StreamCtx& c, int ndep, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, int epoch) {
    if (ndep > 4) return c.killEpoch > epoch;
    bool b = false;
    if (ndep > 0) b |= c.sh->stamp[d0 & (ST_STAMPS - 1)] >= epoch;
    if (ndep > 1) b |= c.sh->stamp[d1 & (ST_STAMPS - 1)] >= epoch;
    if (ndep > 2) b |= c.sh->stamp[d2 & (ST_STAMPS - 1)] >= epoch;
    if (ndep > 3) b |= c.sh->stamp[d3 & (ST_STAMPS - 1)] >= epoch;
    return b;
}

__global__ void __launch_bounds__(32) lsd_stream_kernel(PlfGeom g, const float4* lut, const int* gmap, const int* seeds, const int* nSeeds, int* scratch,
                                                       StreamLayout L, int* regAll, int* nRegOut, int* err, int imgFirst) {
    __shared__ StreamShared sh;
    const int img = imgFirst + blockIdx.x, lane = threadIdx.x;
    int* base = scratch + (size_t)blockIdx.x * L.total;
    StreamCtx c;
    c.sh = &sh;
    c.O = reinterpret_cast<uint32_t*>(base + L.O); c.CH = base + L.CH; c.FS = base + L.FS; c.TK = base + L.TK;
    c.RT = reinterpret_cast<int4*>(base + L.RT);
    c.RF = regAll + (size_t)img * g.Ws * g.Hs;
    c.LUT = lut;
    c.G = gmap + (size_t)img * g.Ps * g.Hs;
    c.S = seeds + (size_t)img * g.seedCap;
    c.W = g.Ws; c.H = g.Hs; c.PB = g.Ps; c.ns = nSeeds[img]; c.nChunks = L.nChunks; c.minReg = g.minRegSize; c.segCap = g.segCap; c.lane = lane;
    c.cp = 0; c.scanPos = 0; c.nBlocked = 0; c.nReady = 0; c.bump = 0; c.stackTop = 0; c.nFreeC = 0; c.killEpoch = 0; c.live = 0; c.nReg = 0;
    c.rfPos = 0; c.err = 0; c.nKill = 0; c.nRel = 0;
    for (int i = lane; i < ST_STAMPS; i += 32) sh.stamp[i] = -1;
    __syncwarp();
    StreamLane a;
    a.tag = 0u; a.curPk = 0; a.lastPk = 0; a.claimMask = 0; a.i = 0; a.n = 0; a.first = -1; a.rChunk = 0; a.rOff = 0; a.wChunk = 0; a.wOff = 0;
    a.spare = -1; a.sx = 0; a.sy = 0; a.nBlockedByMe = 0; a.epoch = 0; a.ndep = 0; a.dep0 = a.dep1 = a.dep2 = a.dep3 = 0u;
    a.sumdx = a.sumdy = a.regDeg = 0.f;
    const AlignTol tol = make_align_tol(g.prec);
    const int ns = c.ns;
    int cpBase = -1, scBase = -1, cpPk = 0, scPk = 0;      // 32-candidate windows of the seed list held in registers
#ifdef PLF_STREAM_DEBUG
    long long dbg0 = 0;
    const long long t0 = clock64();
#endif
    long guard = 0;
#pragma unroll 1
    for (;; ++guard) {
        if (guard > 400000) {
            c.err |= 32;
#ifdef PLF_STREAM_DEBUG
            if (blockIdx.x < 2) {
                if (lane == 0) printf("STUCK img %d: cp %d scan %d ns %d ready %d blocked %d live %d kills %d nKill %d nFreeC %d stack %d bump %d\n", img, c.cp, c.scanPos, ns,
                                      c.nReady, c.nBlocked, c.live, c.killEpoch, c.nKill, c.nFreeC, c.stackTop, c.bump);
                printf("  lane %d tag %u i %d n %d claim %x ndep %d epoch %d\n", lane, a.tag, a.i, a.n, a.claimMask, a.ndep, a.epoch);
                if (lane == 0 && c.cp < ns) { const int pk = c.S[c.cp]; printf("  O[cp] = %u\n", c.O[(pk >> 16) * c.PB + (pk & 0xFFFF)]); }
            }
#endif
            break;
        }
        // ================= loads of the step, all in flight together =================
        if (c.scanPos < c.cp) c.scanPos = c.cp;
        if (cpBase != (c.cp & ~31)) { cpBase = c.cp & ~31; cpPk = (cpBase + lane < ns) ? c.S[cpBase + lane] : -1; }
        if (scBase != (c.scanPos & ~31)) { scBase = c.scanPos & ~31; scPk = (scBase + lane < ns) ? c.S[scBase + lane] : -1; }
        uint32_t oc = ST_COMMITTED, os = ST_COMMITTED;
        float ds = 0.f;
        if (cpPk >= 0) oc = c.O[(cpPk >> 16) * c.PB + (cpPk & 0xFFFF)];
        if (scPk >= 0) { os = c.O[(scPk >> 16) * c.PB + (scPk & 0xFFFF)]; ds = c.LUT[c.G[(scPk >> 16) * c.PB + (scPk & 0xFFFF)]].x; }
        // released candidates: one per lane from the top of the ready list
        int rdPos = -1, rdPk = 0;
        uint32_t rdO = ST_COMMITTED;
        float rdDeg = 0.f;
        const int rdTake = min(c.nReady, 32), rdTop = c.nReady;
        if (lane < rdTake) {
            rdPos = sh.ready[c.nReady - 1 - lane];
            rdPk = c.S[rdPos];
            rdO = c.O[(rdPk >> 16) * c.PB + (rdPk & 0xFFFF)];
            rdDeg = c.LUT[c.G[(rdPk >> 16) * c.PB + (rdPk & 0xFFFF)]].x;
        }
        // the pixels claimed in the previous step, the neighbours of the entry to expand, the entry after it
        const bool growing = a.tag != 0u && a.i < a.n;
        const uint32_t tagTop = a.tag;               // a lane may be given another region before the grow step
        uint32_t vfy[8], o[8];
        float4 r[8];
        const int lx = a.lastPk & 0xFFFF, ly = a.lastPk >> 16, ex = a.curPk & 0xFFFF, ey = a.curPk >> 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            vfy[k] = a.tag;
            if (a.tag != 0u && ((a.claimMask >> k) & 1)) vfy[k] = c.O[(ly + st_dy(k)) * c.PB + lx + st_dx(k)];
        }
        // bit 8: the seed itself, claimed when the region started (the grow step of that same iteration ran on older owner values)
        uint32_t vfySeed = a.tag;
        if (a.tag != 0u && ((a.claimMask >> 8) & 1)) vfySeed = c.O[ly * c.PB + lx];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int xx = ex + st_dx(k), yy = ey + st_dy(k);
            o[k] = ST_COMMITTED;
            r[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (growing && xx >= 0 && yy >= 0 && xx < c.W && yy < c.H) { o[k] = c.O[yy * c.PB + xx]; r[k] = c.LUT[c.G[yy * c.PB + xx]]; }
        }
        int nxPk = 0;
        bool nxLoaded = false;
        if (growing && a.i + 1 < a.n) {
            int ch = a.rChunk, off = a.rOff + 1;
            if (off == 32) { ch = c.CH[ch * 32]; off = 1; }
            nxPk = c.CH[ch * 32 + off];
            nxLoaded = true;
        }
        // ================= V. broken dependencies; are last step's claims still mine? =================
        // A lane found dead here stays inert for the rest of the step; every kill happens at the end of the step.
        bool dead = a.tag != 0u && st_dep_broken(c, a.ndep, a.dep0, a.dep1, a.dep2, a.dep3, a.epoch);
        bool sitOut = false;
        {
            int keep = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t culprit = 0u;
                if (vfy[k] != a.tag) {
                    // free: a later region claimed it in the same step and has been killed since
                    if (vfy[k] != ST_FREE && (vfy[k] < a.tag || vfy[k] == ST_COMMITTED)) dead = true;     // an earlier region did
                    else { keep |= 1 << k; culprit = vfy[k]; }                                            // a later one did
                }
                sh.victim[k][lane] = culprit;
            }
            uint32_t seedCulprit = 0u;
            if (vfySeed != a.tag) {
                if (vfySeed != ST_FREE && (vfySeed < a.tag || vfySeed == ST_COMMITTED)) dead = true;
                else { keep |= 1 << 8; seedCulprit = vfySeed; }
            }
            if (dead) keep = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if ((keep >> k) & 1) c.O[(ly + st_dy(k)) * c.PB + lx + st_dx(k)] = a.tag;                 // take it back; the later region dies
            if ((keep >> 8) & 1) c.O[ly * c.PB + lx] = a.tag;
            a.claimMask = keep;          // pixels taken back are looked at again next step, and the lane sits this step out
            sitOut = keep != 0;
            sh.victim[8][lane] = dead ? a.tag : seedCulprit;
            __syncwarp();
            if (__ballot_sync(0xffffffffu, keep != 0 || dead)) {
#pragma unroll 1
                for (int k = 0; k < 9; ++k) {
                    const uint32_t v = sh.victim[k][lane];
                    unsigned vm = __ballot_sync(0xffffffffu, v != 0u && (dead ? k == 8 : ((keep >> k) & 1)));
                    while (vm) { const int l = __ffs(vm) - 1; vm &= vm - 1u; st_kill_request(c, __shfl_sync(0xffffffffu, v, l)); }
                }
            }
            __syncwarp();
        }
        // ================= F. regions that expanded their last entry and passed the check go to the table =================
        // (the commit pointer's own candidate stays in its lane: the walk below commits it from the registers)
        const bool closed = a.tag != 0u && !dead && a.i >= a.n && a.claimMask == 0;
        {
            unsigned fm = __ballot_sync(0xffffffffu, closed && (int)a.tag - 1 != c.cp);
            if (fm) {
                int slot = -1;
                if ((fm >> lane) & 1u) {
                    const int set = (int)(a.tag & (ST_SETS - 1)) * ST_WAYS;
                    int t[ST_WAYS];
#pragma unroll
                    for (int w = 0; w < ST_WAYS; ++w) t[w] = c.TK[(set + w) * ST_TKW];
#pragma unroll
                    for (int w = ST_WAYS - 1; w >= 0; --w) if (t[w] == 0) slot = set + w;
                }
                // two lanes finishing into the same way: the lower lane goes, the other stays closing until the next step
                const unsigned same = __match_any_sync(0xffffffffu, slot);
                if (slot >= 0 && (__ffs(same) - 1) != lane) slot = -1;
                if (slot >= 0) {
                    int* t = c.TK + slot * ST_TKW;
                    t[1] = a.first; t[2] = a.n; t[3] = __float_as_int(a.regDeg); t[4] = a.epoch; t[5] = a.ndep;
                    t[6] = (int)a.dep0; t[7] = (int)a.dep1; t[8] = (int)a.dep2; t[9] = (int)a.dep3;
                    t[0] = (int)a.tag;
                }
                unsigned dm = __ballot_sync(0xffffffffu, slot >= 0);
                c.live += __popc(dm);
#pragma unroll 1
                while (dm) {
                    const int l = __ffs(dm) - 1;
                    dm &= dm - 1u;
                    const int spare = __shfl_sync(0xffffffffu, a.spare, l);
                    if (spare >= 0) st_free_one(c, spare);
                    if (__shfl_sync(0xffffffffu, a.nBlockedByMe, l)) st_release_request(c, __shfl_sync(0xffffffffu, a.tag, l));
                }
                if (slot >= 0) st_lane_reset(a);
                __syncwarp();
            }
        }
        // ================= 1. commit walk =================
        int forcedPk = -1;                           // >= 0: the commit pointer's candidate needs its region now (lane 0)
#pragma unroll 1
        for (int round = 0; round < 4 && c.cp < ns; ++round) {
            if (round > 0 || cpBase != (c.cp & ~31)) {                                   // window moved or the owner values are stale: reload
                cpBase = c.cp & ~31;
                cpPk = (cpBase + lane < ns) ? c.S[cpBase + lane] : -1;
                oc = cpPk >= 0 ? c.O[(cpPk >> 16) * c.PB + (cpPk & 0xFFFF)] : ST_COMMITTED;
            }
            const unsigned m = __ballot_sync(0xffffffffu, cpBase + lane >= c.cp && oc != ST_COMMITTED);
            if (!m) { c.cp = min(ns, cpBase + 32); continue; }
            const int j = __ffs(m) - 1;
            c.cp = cpBase + j;
            const uint32_t tag = (uint32_t)c.cp + 1u;
            const uint32_t oj = __shfl_sync(0xffffffffu, oc, j);
            if (oj == tag) {
                // its own ticket: finished in a lane, finished in the table, or still growing
                const unsigned lm = __ballot_sync(0xffffffffu, a.tag == tag);
                int first, n, degBits, rec = -1;
                if (lm) {
                    const int l = __ffs(lm) - 1;
                    if (!__shfl_sync(0xffffffffu, (int)closed, l)) break;                // still growing, closing, or dead
                    first = __shfl_sync(0xffffffffu, a.first, l); n = __shfl_sync(0xffffffffu, a.n, l);
                    degBits = __shfl_sync(0xffffffffu, __float_as_int(a.regDeg), l);
                    const int spare = __shfl_sync(0xffffffffu, a.spare, l);
                    if (spare >= 0) st_free_one(c, spare);
                    if (__shfl_sync(0xffffffffu, a.nBlockedByMe, l)) st_release_request(c, tag);
                    if (lane == l) st_lane_reset(a);
                } else {
                    rec = st_find(c, tag);
                    if (rec < 0) break;                                                  // being killed: the next step sees it free
                    const int* t = c.TK + rec * ST_TKW;
                    first = t[1]; n = t[2]; degBits = t[3];
                    if (st_dep_broken(c, t[5], (uint32_t)t[6], (uint32_t)t[7], (uint32_t)t[8], (uint32_t)t[9], t[4])) {
                        st_kill_request(c, tag);                                         // relied on a killed ticket: grow again
                        break;
                    }
                    if (lane == 0) c.TK[rec * ST_TKW] = 0;
                    c.live--;
                }
                // commit: pixels -> committed, large regions appended to the region arena and the region table
                const bool big = n >= c.minReg;
                if (big && c.nReg >= c.segCap) c.err |= 2;
                const bool keepReg = big && c.nReg < c.segCap;
                st_walk<0>(c, first, n, tag, keepReg ? c.rfPos : -1);
                if (keepReg) {
                    if (lane == 0) c.RT[c.nReg] = make_int4(c.rfPos, n, degBits, 0);
                    c.rfPos += n; c.nReg++;
                }
                c.cp++;
                continue;
            }
            // free, or held by a later region: the candidate grows its own region now on the reserved lane 0
            // (the owner value may be stale after this step's commits: look again)
            {
                const int pk = __shfl_sync(0xffffffffu, cpPk, j);
                const uint32_t fresh = c.O[(pk >> 16) * c.PB + (pk & 0xFFFF)];
                if (fresh == ST_COMMITTED) { c.cp++; continue; }
                if (fresh != ST_FREE && fresh <= tag) break;                             // its own ticket after all: next step
                if (!__shfl_sync(0xffffffffu, (int)a.tag, 0)) forcedPk = pk;
            }
            break;
        }
        if (c.cp >= ns && c.live == 0 && !__ballot_sync(0xffffffffu, a.tag != 0u)) break;
        // ================= 2. pick: the walk's candidate, released candidates, the scan =================
        {
            unsigned idle = __ballot_sync(0xffffffffu, a.tag == 0u) & ~1u;
            if (rdTake) {
                // the taken entries sit below the ones pushed since the top of the step: close the gap
                const int fresh = c.nReady - rdTop;
                __syncwarp();
#pragma unroll 1
                for (int i0 = 0; i0 < fresh; i0 += 32) {
                    int v = 0;
                    if (i0 + lane < fresh) v = sh.ready[rdTop + i0 + lane];
                    __syncwarp();
                    if (i0 + lane < fresh) sh.ready[rdTop - rdTake + i0 + lane] = v;
                    __syncwarp();
                }
                c.nReady -= rdTake;
            }
            const bool scanOk = idle && c.scanPos >= scBase && c.scanPos < scBase + 32 && c.scanPos < ns && c.scanPos < c.cp + ST_WINDOW &&
                                c.nBlocked + 32 <= ST_BLOCKED;
            int scanNext = scanOk ? min(ns, scBase + 32) : c.scanPos;
#pragma unroll 1
            for (int src = 0; src < 3; ++src) {
                // per-lane candidate of this source
                int pos, pk;
                float deg;
                bool have;
                if (src == 0) { pos = c.cp; pk = forcedPk; deg = 0.f; have = lane == 0 && forcedPk >= 0; }
                else if (src == 1) {
                    pos = rdPos; pk = rdPk; deg = rdDeg;
                    have = lane < rdTake && !(rdPos <= c.cp || rdO == ST_COMMITTED || (rdO != ST_FREE && rdO <= (uint32_t)rdPos + 1u));
                } else { pos = scBase + lane; pk = scPk; deg = ds; have = scanOk && scPk >= 0 && pos >= c.scanPos && os == ST_FREE && pos != c.cp; }
                unsigned m = __ballot_sync(0xffffffffu, have);
                bool stall = false;
#pragma unroll 1
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1u;
                    const int pj = __shfl_sync(0xffffffffu, pos, j), pkj = __shfl_sync(0xffffffffu, pk, j);
                    float dj = __shfl_sync(0xffffffffu, deg, j);
                    const uint32_t tag = (uint32_t)pj + 1u;
                    const int x = pkj & 0xFFFF, y = pkj >> 16;
                    bool handled = false;
                    int target = -1;
                    if (stall) {
                        // no lane / no room earlier in this source: keep the candidate
                    } else if (src == 0) {
                        dj = c.LUT[c.G[y * c.PB + x]].x;
                        target = 0;
                    } else if (__ballot_sync(0xffffffffu, a.tag == tag)) {
                        handled = true;                                                  // already in a lane (queued twice)
                    } else {
                        // would a growing region swallow it?  (heuristic: on the region's axis with an aligned angle)
                        bool hit = false;
                        if (a.tag != 0u && a.i < a.n) {
                            float d = fabsf(a.regDeg - dj);
                            if (d > 270.f) d = fabsf(d - 360.f);
                            if (d <= 22.5f) {
                                const float cr = -(float)(x - a.sx) * a.sumdy + (float)(y - a.sy) * a.sumdx;
                                hit = cr * cr <= ST_DPERP2 * (a.sumdx * a.sumdx + a.sumdy * a.sumdy);
                            }
                        }
                        const unsigned hm = __ballot_sync(0xffffffffu, hit);
                        if (hm) {
                            if (c.nBlocked < ST_BLOCKED) {
                                const int bl = __ffs(hm) - 1;
                                const uint32_t bt = __shfl_sync(0xffffffffu, a.tag, bl);
                                if (lane == 0) sh.blocked[c.nBlocked] = make_int2(pj, (int)bt);
                                if (lane == bl) a.nBlockedByMe++;
                                c.nBlocked++;
                                handled = true;
                            }
                        } else if (idle && c.live < ST_TICKETS / 2) target = __ffs(idle) - 1;
                    }
                    if (target >= 0) {
                        // start the region in lane `target`; the seed's owner is read afresh: committed or an earlier ticket's ->
                        // nothing to start; a later ticket's -> that ticket dies
                        const uint32_t o0 = c.O[y * c.PB + x];
                        if (!(o0 == ST_COMMITTED || (o0 != ST_FREE && o0 <= tag))) {
                            const int ch = __shfl_sync(0xffffffffu, st_alloc(c, 1u << target), target);
                            if (lane == target) {
                                double sn, cs;
                                sincos((double)dj * kDegToRad, &sn, &cs);
                                a.tag = tag; a.curPk = pkj; a.lastPk = pkj; a.claimMask = 1 << 8; a.i = 0; a.n = 1; a.first = ch; a.rChunk = ch; a.rOff = 1;
                                a.wChunk = ch; a.wOff = 2; a.spare = -1; a.sx = x; a.sy = y; a.nBlockedByMe = 0; a.epoch = c.killEpoch; a.ndep = 0;
                                a.sumdx = (float)cs; a.sumdy = (float)sn; a.regDeg = dj;
                                c.CH[ch * 32 + 1] = pkj;
                                c.O[y * c.PB + x] = tag;
                            }
                            if (o0 != ST_FREE) st_kill_request(c, o0);
                            idle &= ~(1u << target);
                            __syncwarp();
                        }
                        handled = true;
                    }
                    if (!handled) {
                        stall = true;
                        if (src == 1) st_ready_push(c, pj);
                        else if (src == 2) { scanNext = pj; m = 0u; }
                    }
                }
            }
            c.scanPos = scanNext;
        }
        // ================= 3. one lock-step grow step: every growing lane expands one list entry =================
        {
            const unsigned need = __ballot_sync(0xffffffffu, a.tag != 0u && a.i < a.n && a.wOff > 24 && a.spare < 0);
            if (need) { const int ch = st_alloc(c, need); if ((need >> lane) & 1u) a.spare = ch; }
            // not: dead, given another region in this step, or re-checking a pixel it took back
            const bool live = growing && !dead && a.tag == tagTop && !sitOut;
#ifdef PLF_STREAM_DEBUG
            dbg0 += __popc(__ballot_sync(0xffffffffu, live));
#endif
            int firstApp = 0, claim = 0;
            const int n0 = a.n;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint32_t victim = 0u;
                if (live && o[k] != ST_COMMITTED && o[k] != a.tag && lsd_aligned(a.regDeg, r[k].x, tol)) {
                    if (o[k] != ST_FREE && o[k] < a.tag) {                               // an earlier region has it: I rely on that region
                        const uint32_t d = o[k];
                        const bool known = (a.ndep > 0 && d == a.dep0) || (a.ndep > 1 && d == a.dep1) || (a.ndep > 2 && d == a.dep2) || (a.ndep > 3 && d == a.dep3);
                        if (a.ndep <= 4 && !known) {
                            if (a.ndep == 0) a.dep0 = d; else if (a.ndep == 1) a.dep1 = d; else if (a.ndep == 2) a.dep2 = d; else if (a.ndep == 3) a.dep3 = d;
                            a.ndep++;                                                    // 5 = more than four: relies on every earlier ticket
                        }
                    } else {
                        if (o[k] != ST_FREE) victim = o[k];                              // held by a later region: it loses the pixel and dies
                        const int xx = ex + st_dx(k), yy = ey + st_dy(k);
                        c.O[yy * c.PB + xx] = a.tag;
                        const int pk = (yy << 16) | xx;
                        if (a.wOff == 32) { c.CH[a.wChunk * 32] = a.spare; a.wChunk = a.spare; a.wOff = 1; a.spare = -1; }
                        c.CH[a.wChunk * 32 + a.wOff] = pk;
                        if (a.n == n0) firstApp = pk;
                        a.wOff++;
                        a.n++;
                        claim |= 1 << k;
                        a.sumdx = __fadd_rn(a.sumdx, r[k].y);
                        a.sumdy = __fadd_rn(a.sumdy, r[k].z);
                        a.regDeg = fast_atan2_deg(a.sumdy, a.sumdx);
                    }
                }
                sh.victim[k][lane] = victim;
            }
            // next entry
            if (live) {
                a.lastPk = a.curPk;
                a.claimMask = claim;
                a.i++;
                if (a.i < a.n) {
                    a.rOff++;
                    if (a.rOff == 32) { a.rChunk = c.CH[a.rChunk * 32]; a.rOff = 1; }
                    a.curPk = nxLoaded ? nxPk : firstApp;
                }
            }
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const uint32_t v = sh.victim[k][lane];
                unsigned vm = __ballot_sync(0xffffffffu, v != 0u);
                while (vm) { const int l = __ffs(vm) - 1; vm &= vm - 1u; st_kill_request(c, __shfl_sync(0xffffffffu, v, l)); }
            }
        }
        // ================= 4. the kills of the step, then the parked candidates of finished and killed regions =================
        if (c.nKill) {
            __syncwarp();
#pragma unroll 1
            for (int kq = 0; kq < c.nKill; ++kq) {
                const uint32_t v = sh.killQ[kq];
                const unsigned gm = __ballot_sync(0xffffffffu, a.tag == v);
                int first, n, rec = -1;
                if (gm) {                                   // still in a lane (growing or closing)
                    const int l = __ffs(gm) - 1;
                    first = __shfl_sync(0xffffffffu, a.first, l);
                    n = __shfl_sync(0xffffffffu, a.n, l);
                    const int spare = __shfl_sync(0xffffffffu, a.spare, l);
                    if (spare >= 0) st_free_one(c, spare);
                    if (__shfl_sync(0xffffffffu, a.nBlockedByMe, l)) st_release_request(c, v);
                    if (lane == l) st_lane_reset(a);
                } else {
                    rec = st_find(c, v);
                    if (rec < 0) continue;                  // already gone
                    first = c.TK[rec * ST_TKW + 1];
                    n = c.TK[rec * ST_TKW + 2];
                }
                st_walk<1>(c, first, n, v, -1);
                if (rec >= 0) { if (lane == 0) c.TK[rec * ST_TKW] = 0; c.live--; }
                if (lane == 0) sh.stamp[v & (ST_STAMPS - 1)] = c.killEpoch;
                c.killEpoch++;
                st_ready_push(c, (int)v - 1);
                __syncwarp();
            }
            c.nKill = 0;
        }
        if (c.nRel) st_release_blocked(c);
    }
#ifdef PLF_STREAM_DEBUG
    if (lane == 0 && blockIdx.x < 2)
        printf("stream img %d: ns %d iters %ld kills %d regions %d cycles %lld growLaneSteps %lld err %d\n", img, ns, guard, c.killEpoch, c.nReg,
               clock64() - t0, dbg0, c.err);
#endif
    if (lane == 0) {
        nRegOut[img] = c.nReg;
        if (c.err) atomicOr(err, c.err);
    }
}

// rectangle of every committed region (LSD region2rect), one warp per region; the segments come out in commit order
// (region table: rtBase + blockIdx.y * rtStride, entries {offset in the image's list arena, size, region angle in degrees as float bits, -})
__global__ void __launch_bounds__(128) lsd_rect_kernel(PlfGeom g, const int* n2map, const int4* rtBase, size_t rtStride, int* regAll,
                                                      const int* nRegAll, float* segs, int* nSegsOut, int imgFirst) {
    __shared__ __align__(16) double s_sum[4][3][34];
    const int img = imgFirst + blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nReg = min(nRegAll[img], g.segCap);
    if (blockIdx.x == 0 && threadIdx.x == 0) nSegsOut[img] = nReg;
    for (int ri = blockIdx.x * 4 + w; ri < nReg; ri += gridDim.x * 4) {
        const int4 rt = rtBase[(size_t)blockIdx.y * rtStride + ri];
        GrowCtx c;
        c.W = g.Ws; c.H = g.Hs; c.PB = g.Ps; c.lane = lane;
        c.G = n2map + (size_t)img * g.Ps * g.Hs;
        c.R = regAll + (size_t)img * g.Ws * g.Hs + rt.x;
        RectFit rf;
        rect_fit<false>(c, s_sum[w], rt.y, (double)__int_as_float(rt.z) * kDegToRad, g.prec, rf);
        if (lane == 0) {
            float* out = segs + ((size_t)img * g.segCap + ri) * 4;
            const double rr[4] = {rf.x1, rf.y1, rf.x2, rf.y2};
            for (int q4 = 0; q4 < 4; ++q4) {
                double v = rr[q4] + 0.5;
                if (g.lsdScale != 1) v /= g.lsdScale;
                out[q4] = (float)v;
            }
        }
        __syncwarp();
    }
}
