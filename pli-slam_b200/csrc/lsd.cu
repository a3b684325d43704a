// Line extraction on sm_100a: LSD (pre-blur + x1.2 upscale, level-line field, ordered seeds, region growing,
// rectangle fit) -> KeyLine construction / top-N by response -> LBD band descriptor and its binarisation.
// Replaces Lineextractor::operator() (reference src/LineExtractor.cc:31-70), LSDDetectorC::detectImpl
// (Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:227-324) with the OpenCV LSD it wraps, and
// BinaryDescriptor::compute (Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:524-687,1026-1372).
// The arithmetic follows oracle/cpp/lsd.cpp + lbd.cpp operation for operation (double where OpenCV uses double, float
// with explicit _rn intrinsics where it uses float), including the greedy, order-dependent region growing: seeds are
// visited in (gradient bin descending, raster ascending) order — the order cv2 4.13 was measured to use — and every
// region is grown with the exact sequential acceptance rule, one warp per image.
#include "plf_ctx.cuh"
#include "blur.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

__constant__ float c_gaussL[21];
__constant__ float c_gaussG[63];
__constant__ int c_comb[64];

constexpr double kPi = 3.14159265358979323846;
constexpr double kDegToRad = kPi / 180.0;

// ---------------------------------------------------------------------------------------------------------------
// K4a  pre-blur (cv::GaussianBlur 8U, ksize/sigma from sigma_scale) and x`scale` upscale (INTER_LINEAR_EXACT)
// 7-tap separable blur of a whole image batch (LSD pre-blur; LBD 5x5 blur as 7 taps with zero ends)
__global__ void __launch_bounds__(256) blur_image_kernel(const uint8_t* src, size_t srcImgStride, int sp, uint8_t* dst,
                                                         size_t dstImgStride, int dp, int w, int h, int imgFirst,
                                                         const int t0, const int t1, const int t2, const int t3,
                                                         const int t4, const int t5, const int t6) {
    const int tx = (w + BL_TW - 1) / BL_TW;
    const int img = imgFirst + blockIdx.y;
    BlurJob j;
    j.src = src + (size_t)img * srcImgStride;
    j.dst = dst + (size_t)img * dstImgStride;
    j.w = w; j.h = h; j.sp = sp; j.dp = dp;
    const int taps[7] = {t0, t1, t2, t3, t4, t5, t6};
    blur_tile7(j, taps, (blockIdx.x % tx) * BL_TW, (blockIdx.x / tx) * BL_TH);
}

__global__ void __launch_bounds__(256) lsd_upscale_kernel(PlfGeom g, const uint8_t* src, size_t srcImgStride, int sp,
                                                          uint8_t* dst, const PlfLin* linX, const PlfLin* linY,
                                                          int imgFirst) {
    const int dx = blockIdx.x * 128 + threadIdx.x * 4;
    if (dx >= g.Ws) return;
    const int img = imgFirst + blockIdx.z;
    const uint8_t* s = src + (size_t)img * srcImgStride;
    ResizeTaps T;                                   // column part: once for the 4 rows of this thread
    resize_prep(linX + dx, min(4, g.Ws - dx), T);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dy = blockIdx.y * 32 + threadIdx.y + 8 * k;
        if (dy >= g.Hs) break;
        const PlfLin cy = linY[dy];                     // Q8 weights built on the host (a0 = 256 - a1)
        const uint8_t* r0 = s + (size_t)cy.ofs * sp;
        const uint8_t* r1 = s + (size_t)min((int)cy.ofs + 1, g.H - 1) * sp;
        // 4 pixels per thread; the row pitch Ps is a multiple of 128, so the store is always a whole aligned word
        const unsigned v = resize_apply<true>(T, r0, r1, cy.a0, cy.a1);
        *reinterpret_cast<unsigned*>(dst + (size_t)img * g.Ps * g.Hs + (size_t)dy * g.Ps + dx) = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4b  level-line field (LSD ll_angle): 2x2 gradient, squared norm, fastAtan2 angle (degrees), cosf/sinf of the angle,
// per-image max of the squared norm over defined pixels.
// The record of a defined pixel — angle, cosf, sinf, |g|^2 — is a pure function of the integer gradient (gx, gy), each in
// [-510, 510].  A pixel therefore carries only its 20-bit GRADIENT CODE (gx + 512) | (gy + 512) << 10 (0 = undefined:
// a real code has both fields >= 2) in a 4-byte map, and the records live in one table per device indexed by that code
// (2^20 x 16 B = 16.8 MB, L2-resident; the gradients of a frame touch a few hundred KB of it), filled once by the very
// expression the per-pixel code of the first version used, so every record is unchanged bit for bit.  Against a
// 16-byte record per pixel this takes 8.3 MB per image out of HBM (and out of the gradient kernel's writes and the
// grower's reads) for one dependent, cache-resident load per candidate pixel.
#define LSD_LUT_N (1 << 20)
__device__ __forceinline__ int lsd_code(int gx, int gy) { return (gx + 512) | ((gy + 512) << 10); }
__device__ __forceinline__ int lsd_n2(int code) {          // |g|^2 (x4, as LSD's 2x2 operator gives it) of a DEFINED pixel
    const int gx = (code & 1023) - 512, gy = (code >> 10) - 512;
    return gx * gx + gy * gy;
}
__global__ void __launch_bounds__(256) lsd_lut_kernel(float4* lut) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= LSD_LUT_N) return;
    const int gy = (i >> 10) - 512, gx = (i & 1023) - 512;
    const float a = fast_atan2_deg((float)gx, (float)-gy);
    const float af = (float)((double)a * kDegToRad);
    // cosf/sinf taken as correctly rounded (double result rounded to float), the declared oracle rule
    double sn, cs;
    sincos((double)af, &sn, &cs);
    lut[i] = make_float4(a, (float)cs, (float)sn, __int_as_float(gx * gx + gy * gy));
}

__global__ void __launch_bounds__(256) lsd_grad_kernel(PlfGeom g, const uint8_t* U, int* gmap, uint32_t* used, int* n2max, int imgFirst,
                                                      uint32_t* owner) {
    // 128x8 tiles, 4 horizontally adjacent pixels per thread (aligned 32-bit loads of the u8 image, one 16-byte store of
    // the gradient-code map per thread, one bitmap word per 8 threads): 2x2 gradient and |g|^2; a pixel is defined iff
    // |g|^2 > n2Thresh, the integer image of LSD's "norm > rho" test (exact: host-searched with the same IEEE sqrt).
    const int x = blockIdx.x * 128 + threadIdx.x * 4, y = blockIdx.y * 8 + threadIdx.y;
    const int img = imgFirst + blockIdx.z;
    int best = 0;
    if (y < g.Hs) {                                    // warp-uniform: a warp is one row of the tile
        int gv[4] = {0, 0, 0, 0};
        if (x < g.Ws) {
            const uint8_t* r0 = U + (size_t)img * g.Ps * g.Hs + (size_t)y * g.Ps + x;
            const uint8_t* r1 = r0 + (y + 1 < g.Hs ? g.Ps : 0);
            // bytes x .. x+4 of both rows (the row pitch is padded, the 5th byte is only used when x+4 < Ws)
            const unsigned a0 = *reinterpret_cast<const unsigned*>(r0), b0 = *reinterpret_cast<const unsigned*>(r1);
            const unsigned a1 = (x + 4 < g.Ps) ? *reinterpret_cast<const unsigned*>(r0 + 4) : 0u;
            const unsigned b1 = (x + 4 < g.Ps) ? *reinterpret_cast<const unsigned*>(r1 + 4) : 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (x + j < g.Ws - 1 && y < g.Hs - 1) {
                    const int pa = (a0 >> (8 * j)) & 0xFF;                                             // (x, y)
                    const int pb = j < 3 ? (a0 >> (8 * j + 8)) & 0xFF : a1 & 0xFF;                     // (x+1, y)
                    const int pc = (b0 >> (8 * j)) & 0xFF;                                             // (x, y+1)
                    const int pd = j < 3 ? (b0 >> (8 * j + 8)) & 0xFF : b1 & 0xFF;                     // (x+1, y+1)
                    const int DA = pd - pa, BC = pb - pc;
                    const int gx = DA + BC, gy = DA - BC;
                    const int n2 = gx * gx + gy * gy;
                    if (n2 > g.n2Thresh) {
                        best = max(best, n2);
                        gv[j] = lsd_code(gx, gy);
                    }
                }
            }
        }
        // gradient-code map (0 = undefined) and the grower's bitmap with the undefined pixels pre-marked as used: the
        // grower never looks at the record of an undefined pixel
        *reinterpret_cast<int4*>(gmap + (size_t)img * g.Ps * g.Hs + (size_t)y * g.Ps + x) = make_int4(gv[0], gv[1], gv[2], gv[3]);
        // small launches: the owner map of the streaming grower (0 = undefined, PLF_FREE = unclaimed), so that its blocks need
        // not walk the bitmap first
        if (owner)
            *reinterpret_cast<uint4*>(owner + (size_t)img * g.Ps * g.Hs + (size_t)y * g.Ps + x) =
                make_uint4(gv[0] ? 0xFFFFFFFFu : 0u, gv[1] ? 0xFFFFFFFFu : 0u, gv[2] ? 0xFFFFFFFFu : 0u, gv[3] ? 0xFFFFFFFFu : 0u);
        unsigned w = ((gv[0] == 0) | ((gv[1] == 0) << 1) | ((gv[2] == 0) << 2) | ((gv[3] == 0) << 3)) << (4 * (threadIdx.x & 7));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((threadIdx.x & 7) == 0)
            used[((size_t)img * g.Hs + y) * (g.Ps >> 5) + blockIdx.x * 4 + (threadIdx.x >> 3)] = w;
    }
    // per-image maximum: one atomic per block at most, and none when the image's maximum already covers the block's (a
    // same-address atomic per warp made this reduction the kernel's bottleneck)
    __shared__ int s_best;
    if (threadIdx.x == 0 && threadIdx.y == 0) s_best = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (threadIdx.x == 0 && best > 0) atomicMax(&s_best, best);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && s_best > 0 && s_best > *reinterpret_cast<volatile int*>(n2max + img)) atomicMax(n2max + img, s_best);
}

__device__ __forceinline__ int lsd_bin(int n2, double binCoef) { return (int)(sqrt((double)n2 / 4.0) * binCoef); }
__device__ __forceinline__ double lsd_bin_coef(int n2max, int nBins) {
    const double maxGrad = n2max > 0 ? sqrt((double)n2max / 4.0) : -1.0;
    return maxGrad > 0 ? (double)(nBins - 1) / maxGrad : 0.0;
}

// Ordered seed list: defined pixels sorted by (bin descending, raster index ascending) — a stable counting sort.
// One block of ORD_WARPS warps per image (16 warps x nBins counters = 64 KB of shared memory: three blocks per SM);
// warp w owns the w-th contiguous raster segment.  Pass 1 counts per (warp, bin) in
// shared memory, pass 2 turns the counts into write cursors (bins descending, then warps ascending), pass 3 lets every
// warp walk its segment again and place its pixels: lanes of one 32-pixel step that share a bin are ranked with
// __match_any_sync, so the order inside a bin is raster order by construction and no warp waits for another.
#define ORD_WARPS 16
#define ORD_MLP 8
__global__ void __launch_bounds__(32 * ORD_WARPS) lsd_order_kernel(PlfGeom g, const int* gmap, const int* n2max, int* seeds,
                                                         int* nSeeds, int imgFirst, int* posMap) {
    extern __shared__ int s_cur[];          // [ORD_WARPS][nBins] counters / cursors, then [nBins + 1] bin thresholds
    __shared__ int s_scan[ORD_WARPS];
    __shared__ int s_carry;
    const int img = imgFirst + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nBins = g.nBins;
    const double coef = lsd_bin_coef(n2max[img], nBins);
    // raster walk over the pitched gradient-code map: the padding columns hold 0 (undefined) and cost one compare
    const int npx = g.Ps * g.Hs;
    const int segLen = ((npx + ORD_WARPS - 1) / ORD_WARPS + 31) & ~31;
    const int p0 = warp * segLen, p1 = min(p0 + segLen, npx);
    const int* N2 = gmap + (size_t)img * npx;
    int* mine = s_cur + warp * nBins;
    for (int i = tid; i < ORD_WARPS * nBins; i += 32 * ORD_WARPS) s_cur[i] = 0;
    if (tid == 0) s_carry = 0;
    // The bin of a pixel is (int)(sqrt(|g|^2 / 4.0) * coef) in double: monotone in |g|^2, so it is given by the table
    // T[b] = smallest |g|^2 whose bin is >= b, found per image with the exact expression (a binary search per bin).  A
    // pixel then takes a float estimate of its bin (off by at most one) and corrects it against T: the same bins as the
    // double expression at a quarter of the instructions, two of which were double square roots per defined pixel.
    int* T = s_cur + ORD_WARPS * nBins;
    {
        const int top = n2max[img] + 1;                 // no pixel has a larger |g|^2
        for (int b = tid; b <= nBins; b += 32 * ORD_WARPS) {
            int lo = 0, hi = top + 1;                   // smallest n in [0, top + 1] with bin(n) >= b (top + 1: none)
            if (b == nBins) lo = hi = 0x7fffffff;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (mid <= top && lsd_bin(mid, coef) >= b) hi = mid; else lo = mid + 1;
            }
            T[b] = (b == 0) ? 0 : lo;
        }
    }
    const float coefh = (float)(coef * 0.5);
    auto bin_of = [&](int n2) -> int {
        int b = (int)(__fsqrt_rn((float)n2) * coefh);
        b = min(max(b, 0), nBins - 1);
        b -= (n2 < T[b]);
        b += (n2 >= T[b + 1]);
        return b;
    };
    __syncthreads();
    // both walks keep ORD_MLP independent loads in flight per lane (a single dependent load per step leaves the walk
    // bound by memory latency)
    for (int pb = p0 + lane; pb < p1; pb += 32 * ORD_MLP) {
        int v[ORD_MLP];
#pragma unroll
        for (int u = 0; u < ORD_MLP; ++u) v[u] = (pb + 32 * u < p1) ? N2[pb + 32 * u] : 0;
#pragma unroll
        for (int u = 0; u < ORD_MLP; ++u)
            if (v[u]) atomicAdd(&mine[bin_of(lsd_n2(v[u]))], 1);
    }
    __syncthreads();
    // cursors: for bins in descending order, for warps in ascending order
    for (int b0 = 0; b0 < nBins; b0 += 32 * ORD_WARPS) {
        const int r = b0 + tid;                 // position in descending bin order
        const int b = nBins - 1 - r;
        int tot = 0;
        if (r < nBins)
            for (int w = 0; w < ORD_WARPS; ++w) {
                const int t = s_cur[w * nBins + b];
                s_cur[w * nBins + b] = tot;
                tot += t;
            }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        int start = s_carry;
        for (int w = 0; w < warp; ++w) start += s_scan[w];
        start += inc - tot;
        if (r < nBins)
            for (int w = 0; w < ORD_WARPS; ++w) s_cur[w * nBins + b] += start;
        __syncthreads();
        if (tid == 32 * ORD_WARPS - 1) s_carry = start + tot;
        __syncthreads();
    }
    if (tid == 0) nSeeds[img] = s_carry;
    int* out = seeds + (size_t)img * g.seedCap;
    const unsigned lt = (1u << lane) - 1u;
    const int W = g.Ps;                      // a multiple of 128 and segments start on multiples of 32: my (x, y) advances
    int y = (p0 + lane) / W, x = (p0 + lane) - y * W;     // by 32 columns per step with at most one row wrap
    for (int pc = p0; pc < p1; pc += 32 * ORD_MLP) {
        int vv[ORD_MLP];
#pragma unroll
        for (int u = 0; u < ORD_MLP; ++u) vv[u] = (pc + 32 * u + lane < p1) ? N2[pc + 32 * u + lane] : 0;
#pragma unroll
        for (int u = 0; u < ORD_MLP; ++u) {
            if (pc + 32 * u >= p1) break;
            if (x >= W) { x -= W; ++y; }
            const int v = vv[u];
            const bool def = v != 0;
            const int bin = def ? bin_of(lsd_n2(v)) : 0;
            const unsigned wm = __ballot_sync(0xffffffffu, def);
            if (def) {
                const unsigned grp = __match_any_sync(wm, bin);
                const int b = mine[bin];
                __syncwarp(wm);
                if ((grp & lt) == 0) mine[bin] = b + __popc(grp);
                out[b + __popc(grp & lt)] = (y << 16) | x;             // packed (y<<16 | x)
                if (posMap) posMap[(size_t)img * g.Ps * g.Hs + (size_t)y * W + x] = b + __popc(grp & lt);     // streaming grower: pixel -> position
            }
            __syncwarp();
            x += 32;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4c  region growing + rectangle fit (LSD region_grow / region2rect / get_theta / refine / reduce_region_radius).
// One warp per image; seeds in order; every region is grown with the exact sequential rule: list entries are expanded
// front to back, their 8 neighbours in raster order, a neighbour is accepted iff unused and aligned with the CURRENT
// region angle, which is updated after every acceptance.  The kernel is bound by latency — of its own dependent
// instruction chain and of L2/DRAM round trips (ncu: ~105 warp instructions per accepted pixel, one issue every ~12
// cycles per warp, DRAM at a few per cent) — so the design shortens the critical path, not the byte count:
//  * per-pixel data is one read-only 16-byte record (angle, cosf, sinf, |g|^2) -> one LDG.128 per neighbour, issued
//    together with the bitmap word of that neighbour (one memory round trip per batch);
//  * the `used` map is a bitmap in global memory (74 KB per image, rows of Ps bits); the gradient kernel writes it with
//    the undefined pixels already set, so "unused" implies "defined" and undefined pixels never cost a record load;
//  * list entries are packed (y<<16|x) so no integer division is ever needed; the BFS frontier lives in a small
//    shared-memory ring, the full list also goes to global memory for the rectangle fit;
//  * a batch covers 8 list entries x 8 neighbours = 2 sets of 32 candidates in processing order; each set is resolved
//    in speculative SIMD rounds (grow_chain), usually one round per set;
//  * the rectangle sums keep the scalar loop's summation order (three lanes own one accumulator each).
// Measured and rejected (all parity-green): four images per warp (same instruction count, 4x the latency); 40 / 48
// warps per SM through register caps (same images/s at a full wave); record prefetch at commit (slower: traffic);
// 16 list entries per batch and software-pipelined sets of 4 (frontiers are mostly <= 4 pixels wide).
#define GROW_RING 512
#define GROW_SETS 2

// Sequential (scalar-loop order) accumulation of three quantities over the pixels of a region: the 32 lanes write
// their three products to shared memory, then lanes 0..2 each own one accumulator and add the 32 values in order.
__device__ __forceinline__ void seq_sum3(double (*buf)[34], double a, double b, double c, unsigned cnt, int lane,
                                         double& acc) {
    buf[0][lane] = a;
    buf[1][lane] = b;
    buf[2][lane] = c;
    __syncwarp();
    if (lane < 3) {
        if (cnt == 32) {
            // full chunk: 16-byte shared loads, the 32 additions in list order (rows are 34 doubles: 16-byte aligned)
            const double2* row = reinterpret_cast<const double2*>(buf[lane]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double2 v = row[j];
                acc = __dadd_rn(acc, v.x);
                acc = __dadd_rn(acc, v.y);
            }
        } else {
            for (unsigned j = 0; j < cnt; ++j) acc = __dadd_rn(acc, buf[lane][j]);
        }
    }
    __syncwarp();
}

struct GrowState {
    int nrel;         // streaming mode: entries of the relied-on pixel list (SW_MAXREL + 1: too many)
    uint32_t pend;    // streaming mode, per lane: the value my last claim displaced, looked at one round later (PLF_FREE: nothing)
    int ndep;         // streaming mode: entries of c.deps in use (SW_MAXDEP + 1: too many, the region cannot be verified)
    int n;            // region size
    float sumdx, sumdy;
    float regDeg;     // exact region angle in degrees (fastAtan2 of the sums; the seed's angle for a fresh region) unless `dirty`
    bool dirty;
    bool aborted;     // speculative mode only: a pixel of this region belongs to an earlier region of the wave
};

struct GrowCtx {
    const float4* LUT;   // per-device record table, indexed by gradient code
    uint32_t* used;
    const int* G;    // pitched gradient-code map (0 = undefined)
    int* R;
    int* ring;
    int W, H, PB, lane, ddx, ddy;   // PB: bits per bitmap row (= pitch of G)
    // speculative (several regions of one image in flight) mode only:
    uint32_t* owner;    // per pixel: tag of the region that claims it in the current wave, PLF_FREE if none
    uint32_t tag;       // my tag = slot + 1; a lower tag is an earlier seed
    int* invalid;       // per slot: this region overlaps an earlier one of the wave and must be re-grown
    // streaming (MODE 2, lsd_sw.cuh) only: tag = seed position + 1 (an EARLIER position has the smaller tag)
    unsigned* robbed;   // shared memory, one word per chunk of the window: bit = a region of that chunk lost a pixel
    uint32_t* deps;     // shared memory, SW_MAXDEP tags of earlier, uncommitted regions whose claims this region skipped
    uint32_t floorTag;  // tags below this were committed when the region started: their claims are final
    int maxN;           // room for the pixel list
    const int* posMap;  // seed-list position of every pixel (dirty marks when a pixel is given back)
    unsigned* dirty;    // shared memory, one word per chunk of the window
    unsigned* failedW;  // shared memory, one word per chunk of the window
    const int* scanChunk;   // shared memory: chunks handed out so far
    int* relTop;        // end of the free part of the record buffer: the pixels skipped as an uncommitted earlier region's
                        // are listed downwards from here (with repetitions), SW_MAXREL at most
    int* actN;          // shared memory: current size of the region, for the parking heuristic of the other warps
    bool ldcg;          // experiment switch: owner reads from L2
    bool final;         // grown by the committing warp: every earlier region is final, nobody can take a pixel from it
};
#define PLF_FREE 0xFFFFFFFFu
#define SW_WIN 512                  // chunks (of 32 seed positions) between the commit pointer and the scan pointer
#define SW_MAXDEP 8
#define SW_MAXREL 2048
__device__ __forceinline__ void sw_rob(const GrowCtx& c, uint32_t victimTag) {
    atomicOr(c.robbed + (((victimTag - 1u) >> 5) & (SW_WIN - 1)), 1u << ((victimTag - 1u) & 31u));
}
// what a claim displaced: an earlier tag -> the claimant lost the pixel; a later tag -> that region lost it
__device__ __forceinline__ void sw_settle(const GrowCtx& c, uint32_t old) {
    if (old < c.tag) sw_rob(c, c.tag);
    else if (old != PLF_FREE && old > c.tag) sw_rob(c, old);
}

// q is a BIT index, y * PB + x
__device__ __forceinline__ bool used_bit(const uint32_t* used, int q) {
    // plain (L1-cached) load: the bitmap is only updated by this warp's own SM (stores/atomics keep the SM's L1
    // coherent, __syncwarp orders them), and re-reading L1 hits is what keeps the batch preamble short
    return (used[q >> 5] >> (q & 31)) & 1u;
}

// LSD isAligned(): |theta - a|, folded once around 2*pi, <= tolerance, on radians in double.  Both angles are float
// degrees times one constant, so the decision is taken in float degrees whenever the float difference is farther than
// `kAlignGuard` from the two thresholds (the tolerance and the 270-degree fold) — far more than the float rounding of the
// difference (< 1e-4 below 360) and than the double roundings of the exact form; only the rare pixel inside a guard band
// evaluates the exact double expression.
constexpr float kAlignGuard = 4e-3f;
struct AlignTol {
    double tol;        // radians, the reference's tolerance
    float lo, hi;      // degrees: tol - guard, tol + guard
};
__device__ __forceinline__ AlignTol make_align_tol(double tol) {
    AlignTol t;
    t.tol = tol;
    const double deg = tol / kDegToRad;
    t.lo = (float)(deg - (double)kAlignGuard);
    t.hi = (float)(deg + (double)kAlignGuard);
    return t;
}
__device__ __noinline__ bool lsd_aligned_exact(float thetaDeg, float aDeg, double tol) {
    double nd = fabs(__dsub_rn((double)thetaDeg * kDegToRad, (double)aDeg * kDegToRad));
    const double nw = fabs(__dsub_rn(nd, 2 * kPi));
    nd = (nd > (3 * kPi) / 2) ? nw : nd;
    return nd <= tol;
}
__device__ __forceinline__ bool lsd_aligned(float thetaDeg, float aDeg, const AlignTol& t) {
    float d = fabsf(__fsub_rn(thetaDeg, aDeg));
    const bool nearFold = fabsf(__fsub_rn(d, 270.f)) < kAlignGuard;
    if (d > 270.f) d = fabsf(__fsub_rn(d, 360.f));
    if (nearFold || (d > t.lo && d < t.hi)) return lsd_aligned_exact(thetaDeg, aDeg, t.tol);
    return d <= t.lo;
}

// Resolves one set of 32 candidates (lane order = processing order); `valid` lanes hold pixel q (packed pk) with
// record r.  The sequential rule — a candidate is tested against the region angle left by every acceptance before it
// — is evaluated speculatively, all lanes at once: predict the accepted set A with the current exact angle, let every
// lane rebuild the running sums it would see (the float additions of the A-lanes below it, in lane order: the scalar
// loop's chain, bit for bit), take its own fastAtan2 and test itself.  Up to and including the first lane whose test
// contradicts the prediction every lane has seen the true state, so those decisions are final; they are committed in
// one SIMD step and the rest goes round again.  The angle drifts slowly, so one round usually settles a set:
// ~100 + 9/accepted-pixel warp instructions per round instead of ~95 per accepted pixel.
// speculative mode: has an earlier region of the wave taken one of my pixels?  The flags only ever go 0 -> 1 and are written
// and polled with shared-memory atomics while the regions grow; the authoritative read is after the wave's barrier.
template <int MODE = 1>
__device__ __forceinline__ bool grow_is_invalid(const GrowCtx& c) {
    int f = 0;
    if (MODE == 2) {
        if (c.lane == 0) f = (int)((atomicOr(c.robbed + (((c.tag - 1u) >> 5) & (SW_WIN - 1)), 0u) >> ((c.tag - 1u) & 31u)) & 1u);
    } else {
        if (c.lane == 0) f = atomicOr(c.invalid + (c.tag - 1), 0);
    }
    return __shfl_sync(0xffffffffu, f, 0) != 0;
}

template <int MODE>
__device__ __forceinline__ void grow_chain(GrowState& st, bool valid, int q, int pk, const float4& r, const AlignTol& tol,
                                           const GrowCtx& c, unsigned dupAll, unsigned& accepted) {
    constexpr bool SPEC = MODE != 0;
    unsigned pending = __ballot_sync(0xffffffffu, valid);
    if (!pending) return;
    const unsigned dup = dupAll & pending;          // valid lanes holding my pixel (copies share the used bit)
    const unsigned myBit = 1u << c.lane, lt = myBit - 1u;
    // lanes that are a later copy of a pixel still pending in a lower lane (a copy is never predicted: if the first one
    // is accepted the pixel is used, and copies share the angle); refreshed whenever `pending` shrinks
    unsigned later = __ballot_sync(0xffffffffu, (dup & lt) != 0u);
    while (pending) {
        if (st.dirty) {
            st.regDeg = fast_atan2_deg(st.sumdy, st.sumdx);
            st.dirty = false;
        }
        const bool mine = (pending & myBit) != 0u;
        const unsigned am = __ballot_sync(0xffffffffu, mine && lsd_aligned(st.regDeg, r.x, tol));
        if (!am) break;
        // prediction: the aligned lanes, first instance of every pixel only (a later instance finds it used)
        const unsigned A = am & ~later;
        const bool pred = (A & myBit) != 0u;
        const bool shadowed = (dup & lt & A) != 0u;
        // a lane outside the set (if any) accumulates all of A: the state after the set if the prediction holds
        const unsigned spare = ~pending;
        const int v = spare ? 31 - __clz(spare) : -1;
        float sx = st.sumdx, sy = st.sumdy;
        for (unsigned m = A; m; m &= m - 1u) {
            const int b = __ffs(m) - 1;
            const float cxb = __shfl_sync(0xffffffffu, r.y, b), cyb = __shfl_sync(0xffffffffu, r.z, b);
            if (c.lane > b || c.lane == v) {
                sx = __fadd_rn(sx, cxb);
                sy = __fadd_rn(sy, cyb);
            }
        }
        // lanes below the first A-lane still see the current state: its angle is st.regAngle (for a fresh region
        // that is the seed's angle, not the arctangent of the sums)
        const float ang = ((A & lt) || c.lane == v) ? fast_atan2_deg(sy, sx) : st.regDeg;
        const bool t = mine && !shadowed && lsd_aligned(ang, r.x, tol);
        const unsigned mm = __ballot_sync(0xffffffffu, mine && (t != pred));
        unsigned acc;
        int e;
        if (mm == 0u) {
            acc = A;
            e = 31 - __clz(A);
        } else {
            e = __ffs(mm) - 1;
            const unsigned tb = __ballot_sync(0xffffffffu, t);
            acc = (A & ((1u << e) - 1u)) | (tb & (1u << e));
        }
        if (acc & myBit) {
            const int at = st.n + __popc(acc & lt);
            c.ring[at & (GROW_RING - 1)] = pk;
            c.R[at] = pk;
            if (!SPEC) {
                atomicOr(c.used + (q >> 5), 1u << (q & 31));
            } else if (MODE == 2) {
                // claim with my ticket: the earliest ticket keeps a contested pixel; whoever loses one is marked — when the
                // atomic's result has arrived, i.e. at this lane's next claim or at the end of the region (no round trip
                // on the chain)
                sw_settle(c, st.pend);
                st.pend = atomicMin(c.owner + q, c.tag);
            } else {
                // claim in the wave's owner map: the earlier seed (lower tag) wins a contested pixel, the loser is re-grown
                const uint32_t old = atomicMin(c.owner + q, c.tag);
                if (old < c.tag) atomicOr(c.invalid + (c.tag - 1), 1);
                else if (old != PLF_FREE && old > c.tag) atomicOr(c.invalid + (old - 1), 1);
            }
        }
        if (SPEC && !(MODE == 2 && c.final) && grow_is_invalid<MODE>(c)) { st.aborted = true; st.n += __popc(acc); return; }
        st.n += __popc(acc);
        accepted |= acc;
        if (mm == 0u && v >= 0) {          // the spare lane holds the sums and the angle after all of A
            st.sumdx = __shfl_sync(0xffffffffu, sx, v);
            st.sumdy = __shfl_sync(0xffffffffu, sy, v);
            st.regDeg = __shfl_sync(0xffffffffu, ang, v);
            break;
        }
        float ex = __shfl_sync(0xffffffffu, sx, e), ey = __shfl_sync(0xffffffffu, sy, e);
        if ((acc >> e) & 1u) {
            ex = __fadd_rn(ex, __shfl_sync(0xffffffffu, r.y, e));
            ey = __fadd_rn(ey, __shfl_sync(0xffffffffu, r.z, e));
        }
        st.sumdx = ex;
        st.sumdy = ey;
        st.dirty = true;
        if (mm == 0u) break;
        const unsigned dupAcc = __reduce_or_sync(0xffffffffu, (acc & myBit) ? dup : 0u);
        pending &= ~((2u << e) - 1u) & ~dupAcc;
        later = __ballot_sync(0xffffffffu, (dup & lt & pending) != 0u);
    }
}

// LSD region_grow from the seed (packed pk0, linear index p) with angle tolerance `tol`; returns the region size, the
// pixel list is left in c.R[0..n)
template <int MODE>
__device__ __forceinline__ int grow_region(const GrowCtx& c, int pk0, int p, const AlignTol& tol, double& regAngleOut,
                                           int* ndepOut = nullptr, int* nrelOut = nullptr) {
    constexpr bool SPEC = MODE != 0;
    GrowState st;
    st.n = 1;
    st.ndep = 0;
    st.nrel = 0;
    st.pend = PLF_FREE;
    st.aborted = false;
    if (c.lane == 0) {
        const int pb = (pk0 >> 16) * c.PB + (pk0 & 0xFFFF);
        c.ring[0] = pk0; c.R[0] = pk0;
        if (!SPEC) atomicOr(c.used + (pb >> 5), 1u << (pb & 31));
        else if (MODE == 2) {
            st.pend = atomicMin(c.owner + pb, c.tag);
        } else {
            const uint32_t old = atomicMin(c.owner + pb, c.tag);
            if (old < c.tag) atomicOr(c.invalid + (c.tag - 1), 1);
            else if (old != PLF_FREE && old > c.tag) atomicOr(c.invalid + (old - 1), 1);
        }
    }
    st.regDeg = c.LUT[c.G[(pk0 >> 16) * c.PB + (pk0 & 0xFFFF)]].x;
    st.dirty = false;
    {
        double sn, cs;
        sincos((double)st.regDeg * kDegToRad, &sn, &cs);
        st.sumdx = (float)cs;
        st.sumdy = (float)sn;
    }
    __syncwarp();
    int i = 0;
    while (i < st.n && !(SPEC && st.aborted)) {
        // a batch: up to GROW_SETS sets of 4 list entries x 8 neighbours.  All loads of the batch (ring, bitmap word,
        // record) are issued before the first set is resolved, so a wide frontier pays one memory round trip per
        // 4 * GROW_SETS entries; the sets are then resolved in list order.
        const int nb = min(4 * GROW_SETS, st.n - i);
        const bool inRing = (st.n - i) <= GROW_RING;
        int q[GROW_SETS], pk[GROW_SETS], code[GROW_SETS];
        float4 r[GROW_SETS];
        bool valid[GROW_SETS];
        uint32_t relTag[GROW_SETS];      // streaming mode: the earlier, uncommitted ticket whose claim made me skip this pixel
        if (MODE == 2 && st.n + 2 * st.nrel + 2 * 8 * 4 * GROW_SETS > c.maxN) { st.aborted = true; break; }     // no room: left to the committing warp
        if (MODE == 2 && c.lane == 0) *c.actN = st.n;
#pragma unroll
        for (int s = 0; s < GROW_SETS; ++s) {
            const int e = s * 4 + (c.lane >> 3);
            q[s] = -1 - c.lane;
            pk[s] = 0;
            valid[s] = false;
            code[s] = 0;
            relTag[s] = 0u;
            if (s * 4 >= nb) continue;             // (warp-uniform: a frontier of at most 4 entries skips the second set as a whole)
            if (e < nb) {
                const int rp = inRing ? c.ring[(i + e) & (GROW_RING - 1)] : c.R[i + e];
                const int xx = (rp & 0xFFFF) + c.ddx, yy = (rp >> 16) + c.ddy;
                if (xx >= 0 && yy >= 0 && xx < c.W && yy < c.H) {
                    q[s] = yy * c.PB + xx;
                    pk[s] = (yy << 16) | xx;
                    code[s] = c.G[q[s]];                  // issued together with the bitmap word: one round trip
                    if (MODE == 2) {
                        // owner map only: 0 = undefined, a smaller tag = an earlier region's pixel (used), mine = used,
                        // a larger tag or PLF_FREE = not used in the sequential order (taken from the later region if accepted)
                        const uint32_t o = c.ldcg ? c.owner[q[s]] : __ldcg(c.owner + q[s]);      // L2: the map is written with atomics by every warp
                        valid[s] = o > c.tag;
                        if (!c.final && o < c.tag && o >= c.floorTag) relTag[s] = o;
                    } else {
                        valid[s] = !used_bit(c.used, q[s]);   // unused implies defined: undefined pixels start as used
                        if (SPEC && valid[s] && c.owner[q[s]] == c.tag) valid[s] = false;     // my own pixels count as used
                    }
                }
            }
        }
        // which lanes of a set hold the same pixel: resolved while the loads are in flight
        unsigned dupAll[GROW_SETS];
#pragma unroll
        for (int s = 0; s < GROW_SETS; ++s) dupAll[s] = (s * 4 < nb) ? __match_any_sync(0xffffffffu, q[s]) : 0u;
        // the records of the unused candidates only (a cache-resident table; most candidates are used and load nothing)
#pragma unroll
        for (int s = 0; s < GROW_SETS; ++s) {
            r[s] = make_float4(PLF_NOTDEF, 0.f, 0.f, 0.f);
            if (s * 4 < nb && valid[s]) r[s] = c.LUT[code[s]];
        }
        if (MODE == 2) {
            // remember which uncommitted earlier regions this one relied on (distinct tags; rare)
#pragma unroll
            for (int s = 0; s < GROW_SETS; ++s) {
                unsigned rm = __ballot_sync(0xffffffffu, relTag[s] != 0u);
                if (rm && st.nrel <= SW_MAXREL) {
                    // the pixels themselves, for the commit check of a region whose relied-on region gave pixels back
                    if (st.nrel + __popc(rm) > SW_MAXREL) st.nrel = SW_MAXREL + 1;
                    else {
                        if (relTag[s] != 0u) c.relTop[-1 - st.nrel - __popc(rm & ((1u << c.lane) - 1u))] = q[s];
                        st.nrel += __popc(rm);
                    }
                }
                while (rm) {
                    const uint32_t t = __shfl_sync(0xffffffffu, relTag[s], __ffs(rm) - 1);
                    rm &= ~__ballot_sync(0xffffffffu, relTag[s] == t);
                    if (st.ndep > SW_MAXDEP) continue;
                    const bool have = c.lane < st.ndep && c.deps[c.lane] == t;
                    if (__any_sync(0xffffffffu, have)) continue;
                    if (st.ndep < SW_MAXDEP && c.lane == 0) c.deps[st.ndep] = t;
                    ++st.ndep;
                    __syncwarp();
                }
            }
        }
        unsigned acc[GROW_SETS];
#pragma unroll
        for (int s = 0; s < GROW_SETS; ++s) {
            acc[s] = 0u;
            if (s * 4 < nb) {
                // pixels accepted while resolving the earlier sets are no longer available: compare in registers
                // (the bitmap words were read before those acceptances)
#pragma unroll
                for (int t = 0; t < s; ++t)
                    for (unsigned m = acc[t]; m; m &= m - 1u)
                        if (q[s] == __shfl_sync(0xffffffffu, q[t], __ffs(m) - 1)) valid[s] = false;
                grow_chain<MODE>(st, valid[s], q[s], pk[s], r[s], tol, c, dupAll[s], acc[s]);
                __syncwarp();
            }
        }
        i += nb;
    }
    regAngleOut = (double)(st.dirty ? fast_atan2_deg(st.sumdy, st.sumdx) : st.regDeg) * kDegToRad;
    if (MODE == 2) {
        sw_settle(c, st.pend);
        __syncwarp();
        if (ndepOut) *ndepOut = st.aborted ? -1 : st.ndep;
        if (nrelOut) *nrelOut = st.nrel;
    }
    return st.n;
}

struct RectFit { double x1, y1, x2, y2, width; };

// LSD region2rect + get_theta over c.R[0..n): scalar-loop summation order for the weighted sums
template <bool WIDTH>
__device__ __forceinline__ void rect_fit(const GrowCtx& c, double (*s_sum)[34], int n, double regAngle, double prec, RectFit& rf) {
    const int lane = c.lane;
    double acc3 = 0;      // lane 0: sum x*w, lane 1: sum y*w, lane 2: sum w
    for (int i0 = 0; i0 < n; i0 += 32) {
        const unsigned cnt = min(32, n - i0);
        double wv = 0, xw = 0, yw = 0;
        if (lane < cnt) {
            const int rp = c.R[i0 + lane];
            const int ry = rp >> 16, rx = rp & 0xFFFF;
            wv = sqrt((double)lsd_n2(c.G[ry * c.PB + rx]) / 4.0);
            xw = __dmul_rn((double)rx, wv);
            yw = __dmul_rn((double)ry, wv);
        }
        seq_sum3(s_sum, xw, yw, wv, cnt, lane, acc3);
    }
    const double sw = __shfl_sync(0xffffffffu, acc3, 2);
    const double cxm = __shfl_sync(0xffffffffu, acc3, 0) / sw, cym = __shfl_sync(0xffffffffu, acc3, 1) / sw;
    acc3 = 0;             // lane 0: Ixx, lane 1: Iyy, lane 2: Ixy
    for (int i0 = 0; i0 < n; i0 += 32) {
        const unsigned cnt = min(32, n - i0);
        double vxx = 0, vyy = 0, vxy = 0;
        if (lane < cnt) {
            const int rp = c.R[i0 + lane];
            const int ry = rp >> 16, rx = rp & 0xFFFF;
            const double wv = sqrt((double)lsd_n2(c.G[ry * c.PB + rx]) / 4.0);
            const double dx = __dsub_rn((double)rx, cxm), dy = __dsub_rn((double)ry, cym);
            vxx = __dmul_rn(__dmul_rn(dy, dy), wv);
            vyy = __dmul_rn(__dmul_rn(dx, dx), wv);
            vxy = -__dmul_rn(__dmul_rn(dx, dy), wv);
        }
        seq_sum3(s_sum, vxx, vyy, vxy, cnt, lane, acc3);
    }
    const double Ixx = __shfl_sync(0xffffffffu, acc3, 0), Iyy = __shfl_sync(0xffffffffu, acc3, 1);
    const double Ixy = __shfl_sync(0xffffffffu, acc3, 2);
    const double dI = __dsub_rn(Ixx, Iyy);
    const double lambda = __dmul_rn(0.5, __dsub_rn(__dadd_rn(Ixx, Iyy),
                                                   sqrt(__dadd_rn(__dmul_rn(dI, dI), __dmul_rn(__dmul_rn(4.0, Ixy), Ixy)))));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2_deg((float)__dsub_rn(lambda, Ixx), (float)Ixy)
                                           : (double)fast_atan2_deg((float)Ixy, (float)__dsub_rn(lambda, Iyy));
    theta *= kDegToRad;
    {
        double diff = theta - regAngle;
        while (diff <= -kPi) diff += 2 * kPi;
        while (diff > kPi) diff -= 2 * kPi;
        if (diff < 0) diff = -diff;
        if (diff > prec) theta += kPi;
    }
    double dxr, dyr;
    sincos(theta, &dyr, &dxr);
    double lmin = 0, lmax = 0, wmin = 0, wmax = 0;
    for (int i0 = lane; i0 < n; i0 += 32) {
        const int rp = c.R[i0];
        const int ry = rp >> 16, rx = rp & 0xFFFF;
        const double rdx = __dsub_rn((double)rx, cxm), rdy = __dsub_rn((double)ry, cym);
        const double l = __dadd_rn(__dmul_rn(rdx, dxr), __dmul_rn(rdy, dyr));
        lmax = fmax(lmax, l);
        lmin = fmin(lmin, l);
        if (WIDTH) {      // the rectangle width is only needed by refine()
            const double w = __dadd_rn(__dmul_rn(-rdx, dyr), __dmul_rn(rdy, dxr));
            wmax = fmax(wmax, w);
            wmin = fmin(wmin, w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        if (WIDTH) {
            wmax = fmax(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
            wmin = fmin(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
        }
    }
    rf.x1 = __dadd_rn(cxm, __dmul_rn(lmin, dxr));
    rf.y1 = __dadd_rn(cym, __dmul_rn(lmin, dyr));
    rf.x2 = __dadd_rn(cxm, __dmul_rn(lmax, dxr));
    rf.y2 = __dadd_rn(cym, __dmul_rn(lmax, dyr));
    rf.width = __dsub_rn(wmax, wmin);
    if (rf.width < 1.0) rf.width = 1.0;
}

__device__ __forceinline__ double lsd_dist(double x1, double y1, double x2, double y2) {
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}
__device__ __forceinline__ double lsd_dist_sq(double x1, double y1, double x2, double y2) {
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// LSD refine() for refine = 1 (STANDARD): if the rectangle is too sparse, re-grow with a tolerance taken from the local
// angle spread, then shrink the region radius until it is dense.  Returns false if the region is rejected.
// *(reg[i].used) = NOTUSED: the bitmap bit in the sequential kernel; in the streaming kernel (final mode only) the pixel goes
// back to PLF_FREE and its own seed position is marked dirty — whoever skipped that seed has to look again
template <int MODE>
__device__ __forceinline__ void lsd_unmark(const GrowCtx& c, int qb) {
    if (MODE == 0) {
        atomicAnd(c.used + (qb >> 5), ~(1u << (qb & 31)));
    } else {
        if (atomicCAS(c.owner + qb, c.tag, PLF_FREE) == c.tag) {
            const int pos = c.posMap[qb];
            __threadfence_block();
            if ((pos >> 5) < *(volatile const int*)c.scanChunk) atomicOr(c.dirty + ((pos >> 5) & (SW_WIN - 1)), 1u << (pos & 31));
        }
    }
}

template <int MODE = 0>
__device__ bool lsd_refine(const GrowCtx& c, double (*s_sum)[34], int& n, double& regAngle, double prec, double densityTh,
                           RectFit& rf) {
    const int lane = c.lane, W = c.W;
    double density = (double)n / __dmul_rn(lsd_dist(rf.x1, rf.y1, rf.x2, rf.y2), rf.width);
    if (density >= densityTh) return true;
    const int pk0 = c.R[0];
    const int sy = pk0 >> 16, sx = pk0 & 0xFFFF, p0 = sy * W + sx;
    const double xc = (double)sx, yc = (double)sy;
    const double angC = (double)c.LUT[c.G[sy * c.PB + sx]].x * kDegToRad;
    double acc3 = 0;      // lane 0: sum of signed angle differences, lane 1: sum of their squares
    int cntIn = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const unsigned cnt = min(32, n - i0);
        double v0 = 0, v1 = 0;
        bool in = false;
        if (lane < cnt) {
            const int rp = c.R[i0 + lane];
            const int ry = rp >> 16, rx = rp & 0xFFFF, qb = ry * c.PB + rx;
            lsd_unmark<MODE>(c, qb);                                              // *(reg[i].used) = NOTUSED
            if (lsd_dist(xc, yc, (double)rx, (double)ry) < rf.width) {
                double d = __dsub_rn((double)c.LUT[c.G[qb]].x * kDegToRad, angC);         // angle_diff_signed
                while (d <= -kPi) d += 2 * kPi;
                while (d > kPi) d -= 2 * kPi;
                v0 = d;
                v1 = __dmul_rn(d, d);
                in = true;
            }
        }
        cntIn += __popc(__ballot_sync(0xffffffffu, in));
        seq_sum3(s_sum, v0, v1, 0.0, cnt, lane, acc3);
    }
    const double sum = __shfl_sync(0xffffffffu, acc3, 0), ssum = __shfl_sync(0xffffffffu, acc3, 1);
    const double mean = sum / (double)cntIn;
    const double tau = __dmul_rn(2.0, sqrt(__dadd_rn(__dsub_rn(ssum, __dmul_rn(__dmul_rn(2.0, mean), sum)) / (double)cntIn,
                                                     __dmul_rn(mean, mean))));
    __syncwarp();
    __threadfence_block();
    if (MODE == 2 && lane == 0)       // whoever relied on this region's first growth has to look at the pixels again
        atomicOr(c.failedW + (((c.tag - 1u) >> 5) & (SW_WIN - 1)), 1u << ((c.tag - 1u) & 31u));
    n = grow_region<MODE>(c, pk0, p0, make_align_tol(tau), regAngle);
    if (n < 2) return false;
    rect_fit<true>(c, s_sum, n, regAngle, prec, rf);
    density = (double)n / __dmul_rn(lsd_dist(rf.x1, rf.y1, rf.x2, rf.y2), rf.width);
    if (density >= densityTh) return true;
    // reduce_region_radius: drop the points farther than 75 % of the radius (swap-with-last removal, list order matters
    // for the next rectangle fit, so the removal is done serially by lane 0), refit, until dense
    const double r1 = lsd_dist_sq(xc, yc, rf.x1, rf.y1), r2 = lsd_dist_sq(xc, yc, rf.x2, rf.y2);
    double radSq = r1 > r2 ? r1 : r2;
    while (density < densityTh) {
        radSq = __dmul_rn(radSq, 0.75 * 0.75);
        int sz = n;
        if (lane == 0) {
            for (int i = 0; i < sz; ++i) {
                const int rp = c.R[i];
                const int ry = rp >> 16, rx = rp & 0xFFFF;
                if (lsd_dist_sq(xc, yc, (double)rx, (double)ry) > radSq) {
                    if (MODE == 0) lsd_unmark<MODE>(c, ry * c.PB + rx);
                    c.R[i] = c.R[sz - 1];
                    if (MODE != 0) c.R[sz - 1] = rp;      // the removed pixels collect behind the list, un-marked below by all lanes
                    --sz;
                    --i;
                }
            }
        }
        const int nBefore = n;
        n = __shfl_sync(0xffffffffu, sz, 0);
        __syncwarp();
        if (MODE != 0)
            for (int i = n + lane; i < nBefore; i += 32) {
                const int rp = c.R[i];
                lsd_unmark<MODE>(c, (rp >> 16) * c.PB + (rp & 0xFFFF));
            }
        if (n < 2) return false;
        rect_fit<true>(c, s_sum, n, regAngle, prec, rf);
        density = (double)n / __dmul_rn(lsd_dist(rf.x1, rf.y1, rf.x2, rf.y2), rf.width);
    }
    return true;
}

// (Measured: two images per block with the registers capped at 48 / 40 to hold 40 / 48 warps per SM instead of 32 gives
// the same images/s at a full wave and a slower single warp — the kernel is not occupancy-limited there.)
template <bool REFINE>
__global__ void __launch_bounds__(32) lsd_grow_kernel(PlfGeom g, const float4* lut, const int* gmap, const int* seeds,
                                                      const int* nSeeds, uint32_t* usedAll, int* reg, float* segs,
                                                      int* nSegsOut, int* err, int imgFirst, unsigned long long* imgNs) {
    __shared__ int ring[GROW_RING];
    __shared__ __align__(16) double s_sum[3][34];
    const int img = imgFirst + blockIdx.x, lane = threadIdx.x;
    // stage timing only: nanoseconds this image's warp ran (one warp per image: a launch lasts as long as its slowest image)
    unsigned long long t0 = 0;
    if (imgNs) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    GrowCtx c;
    c.W = g.Ws; c.H = g.Hs; c.PB = g.Ps; c.lane = lane;
    const size_t base = (size_t)img * c.W * c.H;
    c.LUT = lut;
    c.G = gmap + (size_t)img * g.Ps * g.Hs;
    c.used = usedAll + (size_t)img * (g.Ps >> 5) * g.Hs;
    c.R = reg + base;
    c.ring = ring;
    const int k = lane & 7;
    c.ddx = (k < 3) ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6));
    c.ddy = (k < 3) ? -1 : (k < 5 ? 0 : 1);
    const int* S = seeds + (size_t)img * g.seedCap;
    float* out = segs + (size_t)img * g.segCap * 4;
    const int ns = nSeeds[img];
    const double prec = g.prec;
    const AlignTol precTol = make_align_tol(prec);
    int nSeg = 0;
    for (int s0 = 0; s0 < ns; s0 += 32) {
        const int mySeed = (s0 + lane < ns) ? S[s0 + lane] : -1;       // packed (y<<16 | x)
        const int myQ = (mySeed >> 16) * c.W + (mySeed & 0xFFFF), myB = (mySeed >> 16) * c.PB + (mySeed & 0xFFFF);
        const bool myFree = mySeed >= 0 && !used_bit(c.used, myB);
        unsigned fm = __ballot_sync(0xffffffffu, myFree);
        while (fm) {
            const int si = __ffs(fm) - 1;
            fm &= fm - 1;
            const int pk0 = __shfl_sync(0xffffffffu, mySeed, si);
            const int p = __shfl_sync(0xffffffffu, myQ, si);
            if (used_bit(c.used, __shfl_sync(0xffffffffu, myB, si))) continue;   // claimed by a region grown earlier in this chunk
            double regAngle;
            int n = grow_region<0>(c, pk0, p, precTol, regAngle);
            if (n < g.minRegSize) continue;
            RectFit rf;
            rect_fit<REFINE>(c, s_sum, n, regAngle, prec, rf);
            if (REFINE && !lsd_refine(c, s_sum, n, regAngle, prec, g.densityTh, rf)) continue;
            if (lane == 0) {
                if (nSeg < g.segCap) {
                    const double rr[4] = {rf.x1, rf.y1, rf.x2, rf.y2};
                    for (int q4 = 0; q4 < 4; ++q4) {
                        double v = rr[q4] + 0.5;
                        if (g.lsdScale != 1) v /= g.lsdScale;
                        out[nSeg * 4 + q4] = (float)v;
                    }
                } else {
                    atomicOr(err, 2);
                }
            }
            if (nSeg < g.segCap) ++nSeg;
            __syncwarp();
        }
    }
    if (lane == 0) nSegsOut[img] = nSeg;
    if (imgNs && lane == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        imgNs[img] = t1 - t0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4c'  (round 1; superseded on the product path by lsd_grow_sw_kernel, kept behind PLF_LSD_GROWER=mw for comparison)
// the same region growing for SMALL batches: several regions of ONE image in flight (a block of MW warps per
// image) with the sequential result.  A region only depends on the `used` state of the pixels it examines, so the next
// few unused seeds (picked at least MW_DIST pixels apart; seeds on the same edge would only collide) are grown
// speculatively against the committed bitmap, each claiming its pixels in a per-wave owner map with atomicMin(tag):
// the earlier seed wins a contested pixel and the loser is marked invalid.  Then one warp walks the seed list in order
// and commits regions (bitmap bits, segment) until it meets an invalid region or a seed that was skipped for distance
// and that no committed region swallowed; everything after that point is discarded and grown again in the next wave.
// The first unused seed of a wave is always picked and can never lose, so every wave commits at least one region.
// Exactness: a committed region saw exactly the used pixels the sequential order gives it — pixels of earlier regions
// it merely examined and rejected do not matter, pixels it accepted are contested through the owner map.
#define MW PLF_MW_WARPS
#define MW_SCAN 256
#define MW_DIST 12
__global__ void __launch_bounds__(32 * MW) lsd_grow_mw_kernel(PlfGeom g, const float4* lut, const int* gmap, const int* seeds,
                                                            const int* nSeeds, uint32_t* usedAll, uint32_t* ownerAll, int* regAll,
                                                            float* segs, int* nSegsOut, int* err, int imgFirst) {
    __shared__ int ring[MW][GROW_RING];
    __shared__ __align__(16) double s_sum[MW][3][34];
    __shared__ double s_seg[MW][4];
    __shared__ int s_pickPos[MW], s_pickPk[MW], s_n[MW], s_inv[MW], s_hasSeg[MW];
    __shared__ int s_nPick, s_pos, s_scanEnd, s_nSeg, s_segIdx[MW];
    __shared__ unsigned s_committed;
    const int img = imgFirst + blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t npx = (size_t)g.Ws * g.Hs, npb = (size_t)g.Ps * g.Hs;
    GrowCtx c;
    c.W = g.Ws; c.H = g.Hs; c.PB = g.Ps; c.lane = lane;
    c.LUT = lut;
    c.G = gmap + (size_t)img * npb;
    c.used = usedAll + (size_t)img * (g.Ps >> 5) * g.Hs;
    c.owner = ownerAll + (size_t)blockIdx.x * npb;
    c.R = regAll + ((size_t)blockIdx.x * MW + w) * npx;
    c.ring = ring[w];
    c.tag = (uint32_t)w + 1u;
    c.invalid = s_inv;
    {
        const int k = lane & 7;
        c.ddx = (k < 3) ? k - 1 : (k == 3 ? -1 : (k == 4 ? 1 : k - 6));
        c.ddy = (k < 3) ? -1 : (k < 5 ? 0 : 1);
    }
    const int* S = seeds + (size_t)img * g.seedCap;
    float* out = segs + (size_t)img * g.segCap * 4;
    const int ns = nSeeds[img];
    const double prec = g.prec;
    const AlignTol precTol = make_align_tol(prec);
    if (threadIdx.x == 0) { s_pos = 0; s_nSeg = 0; }
    __syncthreads();
    while (s_pos < ns) {
        // ---- 1. pick the seeds of the wave (warp 0) ----
        if (w == 0) {
            int nPick = 0, pos = s_pos, scanned = 0, lastPick = -1;
            int myX = -100000, myY = -100000;                       // lane j < nPick holds pick j
            while (nPick < MW && pos < ns && scanned < MW_SCAN) {
                const int p = pos + lane;
                const int seed = p < ns ? S[p] : -1;
                const bool unused = seed >= 0 && !used_bit(c.used, (seed >> 16) * c.PB + (seed & 0xFFFF));
                unsigned um = __ballot_sync(0xffffffffu, unused);
                while (um && nPick < MW) {
                    const int l = __ffs(um) - 1;
                    um &= um - 1u;
                    const int sd = __shfl_sync(0xffffffffu, seed, l);
                    const int sx = sd & 0xFFFF, sy = sd >> 16;
                    const bool nearPick = lane < nPick && max(abs(sx - myX), abs(sy - myY)) < MW_DIST;
                    if (!__any_sync(0xffffffffu, nearPick)) {
                        if (lane == nPick) { myX = sx; myY = sy; }
                        if (lane == 0) { s_pickPos[nPick] = pos + l; s_pickPk[nPick] = sd; }
                        lastPick = pos + l;
                        ++nPick;
                    }
                }
                pos += 32;
                scanned += 32;
            }
            if (lane < MW) s_inv[lane] = 0;
            if (lane == 0) {
                s_nPick = nPick;
                s_scanEnd = nPick == MW ? lastPick + 1 : min(pos, ns);      // every seed below scanEnd was examined
            }
        }
        __syncthreads();
        // ---- 2. grow the picked regions concurrently ----
        if (w < s_nPick) {
            const int pk0 = s_pickPk[w];
            double regAngle;
            const int n = grow_region<1>(c, pk0, (pk0 >> 16) * c.W + (pk0 & 0xFFFF), precTol, regAngle);
            int hasSeg = 0;
            if (!grow_is_invalid<1>(c) && n >= g.minRegSize) {
                RectFit rf;
                rect_fit<false>(c, s_sum[w], n, regAngle, prec, rf);
                if (lane == 0) { s_seg[w][0] = rf.x1; s_seg[w][1] = rf.y1; s_seg[w][2] = rf.x2; s_seg[w][3] = rf.y2; }
                hasSeg = 1;
            }
            if (lane == 0) { s_n[w] = n; s_hasSeg[w] = hasSeg; }
        }
        __syncthreads();
        // ---- 3. decide, in seed order, which regions of the wave are committed (warp 0; registers only after the loads) ----
        // A seed below scanEnd is skipped if it was used before the wave or lies in a region committed earlier in this
        // wave (its owner tag is that region's: claims are still in place); a picked seed commits its region unless the
        // region is invalid; anything else — a seed skipped for distance that nobody swallowed, an invalid region — ends
        // the wave there.
        if (w == 0) {
            const int nPick = s_nPick, scanEnd = s_scanEnd;
            int pos = s_pos, nSeg = s_nSeg, stopPos = -1;
            unsigned committed = 0u;                 // slots committed so far in this wave
            while (pos < scanEnd && stopPos < 0) {
                const int p = pos + lane;
                const bool inRange = p < scanEnd;
                const int seed = inRange ? S[p] : -1;
                const int bidx = inRange ? (seed >> 16) * c.PB + (seed & 0xFFFF) : 0;
                const bool usedBefore = !inRange || used_bit(c.used, bidx);
                const uint32_t tag = inRange ? c.owner[bidx] : PLF_FREE;
                int pickIdx = -1;
                for (int k = 0; k < nPick; ++k) if (p == s_pickPos[k]) pickIdx = k;
                unsigned open = __ballot_sync(0xffffffffu, !usedBefore);
                while (open) {
                    const int l = __ffs(open) - 1;
                    open &= open - 1u;
                    const uint32_t tl = __shfl_sync(0xffffffffu, tag, l);
                    const int k = __shfl_sync(0xffffffffu, pickIdx, l);
                    if (k >= 0) {
                        // (a pick whose seed was swallowed by an earlier committed region lost that pixel: it is invalid)
                        if (*(volatile int*)(s_inv + k)) {
                            if (tl != PLF_FREE && tl != (uint32_t)k + 1u && ((committed >> (tl - 1u)) & 1u)) continue;   // swallowed
                            stopPos = pos + l;
                            break;
                        }
                        committed |= 1u << k;
                        if (lane == 0) s_segIdx[k] = s_hasSeg[k] ? nSeg : -1;
                        if (s_hasSeg[k]) ++nSeg;
                    } else {
                        if (tl != PLF_FREE && ((committed >> (tl - 1u)) & 1u)) continue;     // swallowed by a committed region
                        stopPos = pos + l;                                                    // would have started its own region here
                        break;
                    }
                }
                if (stopPos < 0) pos += 32;
            }
            if (lane == 0) { s_pos = stopPos >= 0 ? stopPos : scanEnd; s_nSeg = nSeg; s_committed = committed; }
        }
        __syncthreads();
        // ---- 4. committed regions enter the bitmap and write their segment; every region withdraws its claims ----
        if (w < s_nPick) {
            const int n = s_n[w];
            const bool commit = (s_committed >> w) & 1u;
            for (int i = lane; i < n; i += 32) {
                const int pk = c.R[i];
                const int qb = (pk >> 16) * c.PB + (pk & 0xFFFF);
                if (commit) atomicOr(c.used + (qb >> 5), 1u << (qb & 31));
                atomicCAS(c.owner + qb, c.tag, PLF_FREE);
            }
            if (commit && lane == 0 && s_segIdx[w] >= 0) {
                const int si = s_segIdx[w];
                if (si < g.segCap) {
                    for (int q4 = 0; q4 < 4; ++q4) {
                        double v = s_seg[w][q4] + 0.5;
                        if (g.lsdScale != 1) v /= g.lsdScale;
                        out[si * 4 + q4] = (float)v;
                    }
                } else {
                    atomicOr(err, 2);
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) nSegsOut[img] = min(s_nSeg, g.segCap);
}

#include "lsd_sw.cuh"
#include "lsd_stream.cuh"
#include "lsd_lane.cuh"

// cv::LineIterator(img, Point2f, Point2f).count of LSDDetector_custom.cpp:295-296, 8-connected: end points rounded half to
// even, clipped with cv::clipLine when one lies outside the image (the clamp of checkLineExtremes leaves x in
// [W-0.5, W), which rounds to W), count = max(|dx|, |dy|) + 1, 0 when nothing is left.  clipLine is OpenCV's integer
// Cohen-Sutherland: offsets through double, truncated toward zero, the second point clipped against the moved first.
__device__ __forceinline__ int line_iterator_count(float fx1, float fy1, float fx2, float fy2, int W, int H) {
    long long x1 = __float2int_rn(fx1), y1 = __float2int_rn(fy1), x2 = __float2int_rn(fx2), y2 = __float2int_rn(fy2);
    const long long right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    if ((c1 | c2) != 0) return 0;
    const long long dx = x2 > x1 ? x2 - x1 : x1 - x2, dy = y2 > y1 ? y2 - y1 : y1 - y2;
    return (int)(dx > dy ? dx : dy) + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// K4d  KeyLine construction (LSDDetector_custom.cpp:268-308) and top-N by response (src/LineExtractor.cc:56-65;
// declared rule: stable order, response descending then detection index ascending).  One block per image.
__global__ void __launch_bounds__(256) keylines_kernel(PlfGeom g, const float* segs, const int* nSegs, plf_keyline* klAll,
                                                       plf_keyline* klOut, int* nKl, int* err, double minLength,
                                                       int nFeatures, int imgFirst) {
    extern __shared__ float s_resp[];
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    const int img = imgFirst + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = nSegs[img];
    const float* sg = segs + (size_t)img * g.segCap * 4;
    plf_keyline* all = klAll + (size_t)img * g.segCap;
    plf_keyline* outk = klOut + (size_t)img * g.klCap;
    const int W = g.W, H = g.H;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < m; i0 += 256) {
        const int i = i0 + tid;
        bool keep = false;
        plf_keyline kl;
        if (i < m) {
            float ex[4] = {sg[i * 4], sg[i * 4 + 1], sg[i * 4 + 2], sg[i * 4 + 3]};
#pragma unroll
            for (int q = 0; q < 4; q += 2) {
                if (ex[q] < 0) ex[q] = 0;
                if (ex[q] >= W) ex[q] = (float)W - 1.0f;
                if (ex[q + 1] < 0) ex[q + 1] = 0;
                if (ex[q + 1] >= H) ex[q + 1] = (float)H - 1.0f;
            }
            const double dxx = (double)__fsub_rn(ex[0], ex[2]), dyy = (double)__fsub_rn(ex[1], ex[3]);
            const double length = (double)(float)sqrt(__dadd_rn(__dmul_rn(dxx, dxx), __dmul_rn(dyy, dyy)));
            keep = length > minLength;
            kl.startPointX = ex[0]; kl.startPointY = ex[1]; kl.endPointX = ex[2]; kl.endPointY = ex[3];
            kl.sPointInOctaveX = ex[0]; kl.sPointInOctaveY = ex[1]; kl.ePointInOctaveX = ex[2]; kl.ePointInOctaveY = ex[3];
            kl.lineLength = (float)length;
            kl.numOfPixels = line_iterator_count(ex[0], ex[1], ex[2], ex[3], W, H);
            kl.angle = (float)atan2((double)__fsub_rn(ex[3], ex[1]), (double)__fsub_rn(ex[2], ex[0]));
            kl.octave = 0;
            kl.size = __fmul_rn(__fsub_rn(ex[2], ex[0]), __fsub_rn(ex[3], ex[1]));
            kl.response = __fdiv_rn(kl.lineLength, (float)max(W, H));
            kl.pt_x = __fdiv_rn(__fadd_rn(ex[2], ex[0]), 2.f);
            kl.pt_y = __fdiv_rn(__fadd_rn(ex[3], ex[1]), 2.f);
            kl.class_id = 0;
        }
        int inc = keep ? 1 : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int base = s_carry;
        for (int w = 0; w < warp; ++w) base += s_warp[w];
        if (keep) {
            const int pos = base + inc - 1;
            kl.class_id = pos;
            all[pos] = kl;
            s_resp[pos] = kl.response;
        }
        __syncthreads();
        if (tid == 255) s_carry = base + inc;
        __syncthreads();
    }
    const int cnt = s_carry;
    if (cnt > nFeatures && nFeatures != 0) {
        for (int i = tid; i < cnt; i += 256) {
            const float r = s_resp[i];
            int rank = 0;
            for (int j = 0; j < cnt; ++j) {
                const float rj = s_resp[j];
                rank += (rj > r) || (rj == r && j < i);
            }
            if (rank < nFeatures) {
                plf_keyline kl = all[i];
                kl.class_id = rank;
                outk[rank] = kl;
            }
        }
        if (tid == 0) nKl[img] = nFeatures;
    } else {
        if (cnt > g.klCap && tid == 0) atomicOr(err, 4);
        const int c2 = min(cnt, g.klCap);
        for (int i = tid; i < c2; i += 256) outk[i] = all[i];
        if (tid == 0) nKl[img] = c2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4e  LBD: 5x5 sigma-1 blur (blur_image_kernel<5>), 3x3 Sobel to int16 pairs, band descriptor, binarisation.
__global__ void __launch_bounds__(256) sobel_kernel(const uint8_t* src, size_t imgStride, int sp, short2* dst, int w, int h,
                                                    int imgFirst) {
    // 4 horizontally adjacent pixels per thread: three aligned words per row (x-4.., x.., x+4..) give bytes x-1 .. x+4;
    // REFLECT_101 only matters in the first / last column group and the first / last row
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int img = imgFirst + blockIdx.z;
    const uint8_t* s = src + (size_t)img * imgStride;
    const int ys[3] = {reflect101(y - 1, h), y, reflect101(y + 1, h)};
    int v[3][6];        // bytes x-1 .. x+4 of the three rows
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint8_t* row = s + (size_t)ys[r] * sp;
        const unsigned cur = *reinterpret_cast<const unsigned*>(row + x);
        v[r][1] = cur & 0xFF; v[r][2] = (cur >> 8) & 0xFF; v[r][3] = (cur >> 16) & 0xFF; v[r][4] = cur >> 24;
        v[r][0] = row[reflect101(x - 1, w)];
        v[r][5] = row[min(reflect101(x + 4, w), w - 1)];
        // columns beyond the image inside my group of 4 (w not a multiple of 4): reflect them too
        if (x + 3 >= w) {
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (x + j >= w) v[r][1 + j] = row[max(reflect101(x + j, w), 0)];
        }
    }
    short2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        // pixel x+j: left = v[.][j], centre = v[.][j+1], right = v[.][j+2]; at the right image border the "right"
        // neighbour is the reflection of column x+j+1
        int l0 = v[0][j], l1 = v[1][j], l2 = v[2][j], c0 = v[0][j + 1], c2 = v[2][j + 1];
        int r0 = v[0][j + 2], r1 = v[1][j + 2], r2 = v[2][j + 2];
        if (x + j + 1 >= w && x + j < w) {      // last column: reflect101(w) = w-2 = column x+j-1 = the left neighbour
            r0 = l0; r1 = l1; r2 = l2;
        }
        const int gx = (r0 - l0) + 2 * (r1 - l1) + (r2 - l2);
        const int gy = (l2 - l0) + 2 * (c2 - c0) + (r2 - r0);
        o[j] = make_short2((short)gx, (short)gy);
    }
    short2* d = dst + (size_t)img * w * h + (size_t)y * w + x;
    if (x + 3 < w && (w & 3) == 0) {
        *reinterpret_cast<uint4*>(d) = make_uint4(*reinterpret_cast<unsigned*>(&o[0]), *reinterpret_cast<unsigned*>(&o[1]),
                                                  *reinterpret_cast<unsigned*>(&o[2]), *reinterpret_cast<unsigned*>(&o[3]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x + j < w) d[j] = o[j];
    }
}

// One warp per line.  Lane = row hID of the 63-row line support region (two passes of 32): each lane walks its row
// sequentially (sCorX += dL[0] ...) exactly like the scalar loop, so the four row sums are bit-identical; the band
// accumulation is then done in hID order by one lane per band.
__global__ void __launch_bounds__(128) lbd_kernel(PlfGeom g, const short2* sobel, const plf_keyline* kls, const int* nKl,
                                                  float* lbdOut, uint8_t* descOut, int imgFirst) {
    __shared__ float s_row[4][63][4];
    __shared__ float s_des[4][72];
    const int img = imgFirst + blockIdx.y;
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int li = blockIdx.x * 4 + wl;
    const int n = nKl[img];
    if (li >= n) return;
    const plf_keyline kl = kls[(size_t)img * g.klCap + li];
    const short2* sb = sobel + (size_t)img * g.W * g.H;
    const short realWidth = (short)g.W, imageWidth = realWidth - 1, imageHeight = (short)(g.H - 1);
    const short lengthOfLSP = (short)kl.numOfPixels;
    const short halfWidth = (lengthOfLSP - 1) / 2, halfHeight = 31;
    const float midX = (float)(0.5 * (double)__fadd_rn(kl.sPointInOctaveX, kl.ePointInOctaveX));
    const float midY = (float)(0.5 * (double)__fadd_rn(kl.sPointInOctaveY, kl.ePointInOctaveY));
    const float dL0 = (float)cos((double)kl.angle), dL1 = (float)sin((double)kl.angle);
    const float dO0 = -dL1, dO1 = dL0;
    const float sX00 = __fadd_rn(__fadd_rn(__fmul_rn(-dL0, (float)halfWidth), __fmul_rn(dL1, (float)halfHeight)), midX);
    const float sY00 = __fadd_rn(__fsub_rn(__fmul_rn(-dL1, (float)halfWidth), __fmul_rn(dL0, (float)halfHeight)), midY);
    for (int hID = lane; hID < 63; hID += 32) {
        // sCorX0 after hID steps of "sCorX0 -= dL[1]; sCorY0 += dL[0]" (sequential float updates)
        float sX0 = sX00, sY0 = sY00;
        for (int t = 0; t < hID; ++t) { sX0 = __fsub_rn(sX0, dL1); sY0 = __fadd_rn(sY0, dL0); }
        float sX = sX0, sY = sY0;
        float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
        for (short wID = 0; wID < lengthOfLSP; ++wID) {
            short t = (short)roundf(sX);
            const short xCor = (t < 0) ? 0 : (t > imageWidth) ? imageWidth : t;
            t = (short)roundf(sY);
            const short yCor = (t < 0) ? 0 : (t > imageHeight) ? imageHeight : t;
            const short2 d = sb[(int)yCor * realWidth + xCor];
            const float gDL = __fadd_rn(__fmul_rn((float)d.x, dL0), __fmul_rn((float)d.y, dL1));
            const float gDO = __fadd_rn(__fmul_rn((float)d.x, dO0), __fmul_rn((float)d.y, dO1));
            if (gDL > 0) pgdL = __fadd_rn(pgdL, gDL); else ngdL = __fsub_rn(ngdL, gDL);
            if (gDO > 0) pgdO = __fadd_rn(pgdO, gDO); else ngdO = __fsub_rn(ngdO, gDO);
            sX = __fadd_rn(sX, dL0);
            sY = __fadd_rn(sY, dL1);
        }
        const float coef = c_gaussG[hID];
        s_row[wl][hID][0] = __fmul_rn(coef, pgdL);
        s_row[wl][hID][1] = __fmul_rn(coef, ngdL);
        s_row[wl][hID][2] = __fmul_rn(coef, pgdO);
        s_row[wl][hID][3] = __fmul_rn(coef, ngdO);
    }
    __syncwarp();
    if (lane < 9) {
        const int b = lane;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // pgdL ngdL pgdL2 ngdL2 pgdO ngdO pgdO2 ngdO2
        const int h0 = max(0, 7 * (b - 1)), h1 = min(63, 7 * (b + 2));
        for (int hID = h0; hID < h1; ++hID) {
            const int hb = hID / 7;
            // own band: gaussL[hID%7+7]; band above (b == hb-1): gaussL[hID%7+14]; band below (b == hb+1): gaussL[hID%7]
            const float cf = c_gaussL[hID % 7 + (b == hb ? 7 : (b == hb - 1 ? 14 : 0))];
            const float pL = s_row[wl][hID][0], nL = s_row[wl][hID][1], pO = s_row[wl][hID][2], nO = s_row[wl][hID][3];
            const float cf2 = __fmul_rn(cf, cf);
            acc[0] = __fadd_rn(acc[0], __fmul_rn(cf, pL));
            acc[1] = __fadd_rn(acc[1], __fmul_rn(cf, nL));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(cf2, __fmul_rn(pL, pL)));
            acc[3] = __fadd_rn(acc[3], __fmul_rn(cf2, __fmul_rn(nL, nL)));
            acc[4] = __fadd_rn(acc[4], __fmul_rn(cf, pO));
            acc[5] = __fadd_rn(acc[5], __fmul_rn(cf, nO));
            acc[6] = __fadd_rn(acc[6], __fmul_rn(cf2, __fmul_rn(pO, pO)));
            acc[7] = __fadd_rn(acc[7], __fmul_rn(cf2, __fmul_rn(nO, nO)));
        }
        const float invN = (b == 0 || b == 8) ? (float)(1.0 / (7 * 2.0)) : (float)(1.0 / (7 * 3.0));
        float* des = &s_des[wl][b * 8];
        float temp;
        temp = __fmul_rn(acc[0], invN); des[0] = temp; des[4] = sqrtf(__fsub_rn(__fmul_rn(acc[2], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(acc[1], invN); des[1] = temp; des[5] = sqrtf(__fsub_rn(__fmul_rn(acc[3], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(acc[4], invN); des[2] = temp; des[6] = sqrtf(__fsub_rn(__fmul_rn(acc[6], invN), __fmul_rn(temp, temp)));
        temp = __fmul_rn(acc[5], invN); des[3] = temp; des[7] = sqrtf(__fsub_rn(__fmul_rn(acc[7], invN), __fmul_rn(temp, temp)));
    }
    __syncwarp();
    float* des = s_des[wl];
    if (lane == 0) {
        float tempM = 0, tempS = 0;
        for (int b = 0; b < 9; ++b) {
            for (int q = 0; q < 4; ++q) tempM = __fadd_rn(tempM, __fmul_rn(des[b * 8 + q], des[b * 8 + q]));
            for (int q = 4; q < 8; ++q) tempS = __fadd_rn(tempS, __fmul_rn(des[b * 8 + q], des[b * 8 + q]));
        }
        tempM = __fdiv_rn(1.f, sqrtf(tempM));
        tempS = __fdiv_rn(1.f, sqrtf(tempS));
        for (int b = 0; b < 9; ++b) {
            for (int q = 0; q < 4; ++q) des[b * 8 + q] = __fmul_rn(des[b * 8 + q], tempM);
            for (int q = 4; q < 8; ++q) des[b * 8 + q] = __fmul_rn(des[b * 8 + q], tempS);
        }
        for (int i = 0; i < 72; ++i)
            if ((double)des[i] > 0.4) des[i] = (float)0.4;
        float temp = 0;
        for (int i = 0; i < 72; ++i) temp = __fadd_rn(temp, __fmul_rn(des[i], des[i]));
        temp = __fdiv_rn(1.f, sqrtf(temp));
        for (int i = 0; i < 72; ++i) des[i] = __fmul_rn(des[i], temp);
    }
    __syncwarp();
    float* lo = lbdOut + ((size_t)img * g.klCap + li) * 72;
    for (int i = lane; i < 72; i += 32) lo[i] = des[i];
    {
        const float* f1 = &des[8 * c_comb[lane * 2]];
        const float* f2 = &des[8 * c_comb[lane * 2 + 1]];
        int r = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r |= (f1[i] > f2[i]) << i;
        descOut[((size_t)img * g.klCap + li) * 32 + lane] = (uint8_t)r;
    }
}

}  // namespace

static bool s_tablesReady[64] = {};

static void upload_lbd_tables(int device) {
    if (device < 64 && s_tablesReady[device]) return;
    // weight tables of the BinaryDescriptor constructor (binary_descriptor_custom.cpp:227-258), integer divisions kept
    float gl[21], gg[63];
    {
        const int w = 7;
        double u = (w * 3 - 1) / 2, sigma = (w * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 21; ++i) { double d = i - u; gl[i] = (float)std::exp(d * d * inv); }
        u = (9 * w - 1) / 2; sigma = u; inv = -1 / (2 * sigma * sigma);
        for (int i = 0; i < 63; ++i) { double d = i - u; gg[i] = (float)std::exp(d * d * inv); }
    }
    static const int comb[64] = {0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 6, 1, 2, 1, 3, 1, 4, 1, 5, 1, 6, 2, 3, 2, 4, 2, 5, 2, 6, 2, 7,
                                 2, 8, 3, 4, 3, 5, 3, 6, 3, 7, 3, 8, 4, 5, 4, 6, 4, 7, 4, 8, 5, 6, 5, 7, 5, 8, 6, 7, 6, 8, 7, 8};
    cudaMemcpyToSymbol(c_gaussL, gl, sizeof gl);
    cudaMemcpyToSymbol(c_gaussG, gg, sizeof gg);
    cudaMemcpyToSymbol(c_comb, comb, sizeof comb);
    if (device < 64) s_tablesReady[device] = true;
}

// scratch of the small-batch grower, allocated the first time a small launch happens (a context that only ever runs large
// batches never pays for it): owner map (all PLF_FREE) and one region list per wave slot, for up to PLF_MW_MAX_IMG images
// per-device record table of lsd_grad_kernel (process lifetime, like the constant tables)
static float4* s_gradLut[64] = {};
const float4* plf_grad_lut(plf_ctx* c) {
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    const int dev = c->device;
    if (dev < 0 || dev >= 64) return nullptr;
    if (!s_gradLut[dev]) {
        float4* p = nullptr;
        if (cudaMalloc((void**)&p, (size_t)LSD_LUT_N * sizeof(float4)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        lsd_lut_kernel<<<LSD_LUT_N / 256, 256, 0, c->stream>>>(p);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { cudaFree(p); return nullptr; }
        s_gradLut[dev] = p;
    }
    return s_gradLut[dev];
}

static int plf_ensure_stream_buffers(plf_ctx* c) {
    if (c->d_stream) return 0;
    const StreamLayout L = stream_layout(c->g.Ps, c->g.Ws, c->g.Hs, c->g.segCap);
    const size_t nImg = (size_t)c->p.max_batch * 2;
    if (dalloc(&c->d_stream, nImg * (size_t)L.total) != cudaSuccess) { cudaGetLastError(); c->d_stream = nullptr; return 1; }
    return 0;
}

static int plf_ensure_lane_buffers(plf_ctx* c) {
    if (c->d_laneRT) return 0;
    if (dalloc(&c->d_laneRT, (size_t)c->nImgMax * c->g.segCap) != cudaSuccess) { cudaGetLastError(); c->d_laneRT = nullptr; return 1; }
    return 0;
}

static int plf_ensure_mw_buffers(plf_ctx* c) {
    if (c->d_owner && c->d_regMW) return 0;
    const PlfGeom& g = c->g;
    const size_t nLat = std::min<size_t>((size_t)c->nImgMax, PLF_MW_MAX_IMG);
    const size_t npb = (size_t)g.Ps * g.Hs, npx = (size_t)g.Ws * g.Hs;
    if (cudaMalloc((void**)&c->d_owner, nLat * npb * sizeof(uint32_t)) != cudaSuccess) { c->d_owner = nullptr; cudaGetLastError(); return 1; }
    if (cudaMalloc((void**)&c->d_regMW, nLat * PLF_MW_WARPS * npx * sizeof(int)) != cudaSuccess) {
        cudaFree(c->d_owner); c->d_owner = nullptr; c->d_regMW = nullptr; cudaGetLastError(); return 1;      // fall back to the sequential kernel
    }
    cudaMemsetAsync(c->d_owner, 0xFF, nLat * npb * sizeof(uint32_t), c->stream);
    return 0;
}

// scratch of the streaming small-batch grower (owner map, seed-position map, record buffers: 11 MB per 752x480 image), for up
// to PLF_SW_MAX_IMG images, allocated the first time a small launch happens; its shared-memory block (rings, chunk window)
// needs the opt-in limit
static int plf_ensure_sw_buffers(plf_ctx* c) {
    if (!c->d_swOwner) {
        const size_t nLat = std::min<size_t>((size_t)c->nImgMax, PLF_SW_MAX_IMG);
        const size_t npb = (size_t)c->g.Ps * c->g.Hs, npxA = ((size_t)c->g.Ws * c->g.Hs + 3) & ~(size_t)3;
        const size_t perImg = npxA + (size_t)PLF_SW_WARPS * PLF_SW_WARPBUF;
        if (cudaMalloc((void**)&c->d_swOwner, nLat * npb * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc((void**)&c->d_swPos, nLat * npb * sizeof(int)) != cudaSuccess ||
            cudaMalloc((void**)&c->d_swReg, nLat * perImg * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(c->d_swOwner); cudaFree(c->d_swPos); cudaFree(c->d_swReg);
            c->d_swOwner = nullptr; c->d_swPos = nullptr; c->d_swReg = nullptr;
            cudaGetLastError();
            return 1;                      // fall back to the sequential kernel
        }
    }
    static size_t s_granted[64] = {};
    if (plf_raise_smem_optin(s_granted, c->device, sizeof(SwShared)))
    {
        cudaFuncSetAttribute(lsd_grow_sw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SwShared));
        cudaFuncSetAttribute(lsd_grow_sw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SwShared));
    }
    return 0;
}

int plf_launch_lines(plf_ctx* c, int imgFirst, int nImg) {
    const PlfGeom& g = c->g;
    cudaStream_t s = c->stream;
    upload_lbd_tables(c->device);
    const float4* lut = c->d_gradLut;               // per-device table, fetched by plf_create
    const uint8_t* in = c->d_pyr + g.lv[0].off;      // level 0 of the pyramid block is the input image
    const size_t inStride = (size_t)g.pyrBytes;
    const int ip = g.lv[0].pitch;
    const size_t imgBytes = (size_t)ip * g.H;
    const int tiles = ((g.W + BL_TW - 1) / BL_TW) * ((g.H + BL_TH - 1) / BL_TH);
    int launches = 0;
    plf_mark(c, "lsd_prefilter");
    const uint8_t* upSrc = in;
    size_t upStride = inStride;
    if (g.lsdK > 0) {
        const int* t = g.lsdTaps;
        int t7[7] = {0, 0, 0, 0, 0, 0, 0};                 // centre the ksize taps in a 7-tap window
        for (int k = 0; k < g.lsdK; ++k) t7[(7 - g.lsdK) / 2 + k] = t[k];
        blur_image_kernel<<<dim3(tiles, nImg), dim3(32, 8), 0, s>>>(in, inStride, ip, c->d_lsdBlur, imgBytes, ip, g.W, g.H, imgFirst,
                                                                   t7[0], t7[1], t7[2], t7[3], t7[4], t7[5], t7[6]);
        upSrc = c->d_lsdBlur;
        upStride = imgBytes;
        ++launches;
    }
    lsd_upscale_kernel<<<dim3((g.Ws + 127) / 128, (g.Hs + 31) / 32, nImg), dim3(32, 8), 0, s>>>(g, upSrc, upStride, ip, c->d_lsdU, c->d_lin + c->linLsdX, c->d_lin + c->linLsdY, imgFirst);
    plf_mark(c, "lsd_gradient");
    cudaMemsetAsync(c->d_n2max + imgFirst, 0, nImg * sizeof(int), s);
    // will the streaming grower run?  (same condition as below) then the gradient and order kernels fill its owner / position maps
    static const char* s_modeEarly = getenv("PLF_LSD_GROWER");
    const bool swWill = !s_modeEarly && imgFirst + nImg <= std::min(c->nImgMax, PLF_SW_MAX_IMG) && c->growerPolicy != PLF_GROWER_THROUGHPUT &&
                        plf_ensure_sw_buffers(c) == 0;
    lsd_grad_kernel<<<dim3((g.Ws + 127) / 128, (g.Hs + 7) / 8, nImg), dim3(32, 8), 0, s>>>(g, c->d_lsdU, c->d_n2, c->d_used, c->d_n2max, imgFirst,
                                                                                           swWill ? c->d_swOwner : nullptr);
    plf_mark(c, "lsd_order");
    {
        const size_t smem = ((size_t)ORD_WARPS * g.nBins + g.nBins + 1) * sizeof(int);
        static size_t s_granted[64] = {};
        if (plf_raise_smem_optin(s_granted, c->device, smem))
            cudaFuncSetAttribute(lsd_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lsd_order_kernel<<<nImg, 32 * ORD_WARPS, smem, s>>>(g, c->d_n2, c->d_n2max, c->d_seeds, c->d_nSeeds, imgFirst, swWill ? c->d_swPos : nullptr);
    }
    plf_mark(c, "lsd_grow");
    {
        // seq (one warp per image) for wide launches and sw (16 warps per image, streaming: lsd_sw.cuh) for <= 296 images are
        // the product path (plf_set_grower_policy can keep seq for every size); PLF_LSD_GROWER=seq|mw|stream|lane selects one
        // of the growers by hand: mw = the wave-synchronous predecessor of sw, stream / lane = the two structures that were
        // measured and rejected (profiles/r02_grower_experiments.md)
        static const char* s_mode = getenv("PLF_LSD_GROWER");
        // per-image run time of the one-warp-per-image grower, recorded only while stage timing is on (bench.py reports max / mean)
        unsigned long long* growNs = nullptr;
        if (c->stageTiming) {
            if (!c->d_growNs && dalloc(&c->d_growNs, (size_t)c->nImgMax) != cudaSuccess) { cudaGetLastError(); c->d_growNs = nullptr; }
            growNs = c->d_growNs;
            if (growNs) cudaMemsetAsync(growNs + imgFirst, 0, (size_t)nImg * sizeof(unsigned long long), s);
        }
        const bool wantStream = s_mode && !strcmp(s_mode, "stream");
        const bool wantLane = s_mode && !strcmp(s_mode, "lane");
        static const int s_swFlags = getenv("PLF_SW_FLAGS") ? atoi(getenv("PLF_SW_FLAGS")) : 0;      // experiment switches, see lsd_sw.cuh
        if (g.refine >= 1 && !s_mode && imgFirst + nImg <= std::min(c->nImgMax, PLF_SW_MAX_IMG) && c->growerPolicy != PLF_GROWER_THROUGHPUT && plf_ensure_sw_buffers(c) == 0)
            // few images, refine = 1: the streaming grower; regions that need refining are left to its committing warp
            lsd_grow_sw_kernel<true><<<nImg, 32 * SW_NW, sizeof(SwShared), s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_swOwner,
                                                                               c->d_swReg, c->d_swPos, c->d_segs, c->d_nSegs, c->d_err, imgFirst, s_swFlags | (swWill ? 128 : 0));
        else if (g.refine >= 1)
            lsd_grow_kernel<true><<<nImg, 32, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_reg, c->d_segs,
                                                      c->d_nSegs, c->d_err, imgFirst, growNs);
        else if (wantStream && plf_ensure_stream_buffers(c) == 0) {
            // one lane per region: up to 32 regions of each image in flight in its warp, rectangles fitted afterwards
            const StreamLayout L = stream_layout(g.Ps, g.Ws, g.Hs, g.segCap);
            int* scr = c->d_stream + (size_t)imgFirst * L.total;
            lsd_stream_init_kernel<<<dim3(32, nImg), 256, 0, s>>>(g, c->d_n2, scr, L, imgFirst);
            lsd_stream_kernel<<<nImg, 32, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, scr, L, c->d_reg, c->d_nReg, c->d_err, imgFirst);
            lsd_rect_kernel<<<dim3(64, nImg), 128, 0, s>>>(g, c->d_n2, reinterpret_cast<const int4*>(scr + L.RT), (size_t)L.total / 4, c->d_reg,
                                                           c->d_nReg, c->d_segs, c->d_nSegs, imgFirst);
            launches += 2;
        } else if (wantLane && plf_ensure_lane_buffers(c) == 0) {
            // one lane per image: 32 images per warp, the plain scalar loop; rectangles fitted afterwards
            lsd_grow_lane_kernel<<<(nImg + 31) / 32, 32, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_reg, c->d_laneRT, c->d_nReg,
                                                                 c->d_err, imgFirst, nImg);
            lsd_rect_kernel<<<dim3(16, nImg), 128, 0, s>>>(g, c->d_n2, c->d_laneRT + (size_t)imgFirst * g.segCap, (size_t)g.segCap, c->d_reg,
                                                           c->d_nReg, c->d_segs, c->d_nSegs, imgFirst);
            launches += 1;
        } else if (s_mode && !strcmp(s_mode, "seq"))
            lsd_grow_kernel<false><<<nImg, 32, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_reg, c->d_segs,
                                                       c->d_nSegs, c->d_err, imgFirst, growNs);
        else if (nImg <= PLF_MW_MAX_IMG && s_mode && !strcmp(s_mode, "mw") && plf_ensure_mw_buffers(c) == 0) {
            // the wave-synchronous predecessor of the streaming grower (kept for comparison)
            lsd_grow_mw_kernel<<<nImg, 32 * MW, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_owner, c->d_regMW,
                                                        c->d_segs, c->d_nSegs, c->d_err, imgFirst);
        } else if (imgFirst + nImg <= std::min(c->nImgMax, PLF_SW_MAX_IMG) && c->growerPolicy != PLF_GROWER_THROUGHPUT && plf_ensure_sw_buffers(c) == 0) {
            // few images: 16 regions of each image in flight, one per warp, streaming with an in-order commit pointer
            lsd_grow_sw_kernel<false><<<nImg, 32 * SW_NW, sizeof(SwShared), s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_swOwner,
                                                                         c->d_swReg, c->d_swPos, c->d_segs, c->d_nSegs, c->d_err, imgFirst, s_swFlags | (swWill ? 128 : 0));
        } else
            lsd_grow_kernel<false><<<nImg, 32, 0, s>>>(g, lut, c->d_n2, c->d_seeds, c->d_nSeeds, c->d_used, c->d_reg, c->d_segs,
                                                       c->d_nSegs, c->d_err, imgFirst, growNs);
    }
    plf_mark(c, "line_keylines");
    const double minLen = c->p.min_line_length * std::min(g.W, g.H);
    keylines_kernel<<<nImg, 256, g.segCap * sizeof(float), s>>>(g, c->d_segs, c->d_nSegs, c->d_klAll, c->d_kl, c->d_nKl, c->d_err, minLen, c->p.lsd_nfeatures, imgFirst);
    plf_mark(c, "lbd_blur_sobel");
    const int lt[5] = {14, 62, 104, 62, 14};
    blur_image_kernel<<<dim3(tiles, nImg), dim3(32, 8), 0, s>>>(in, inStride, ip, c->d_lbdBlur, imgBytes, ip, g.W, g.H, imgFirst, 0, lt[0], lt[1], lt[2], lt[3], lt[4], 0);
    sobel_kernel<<<dim3((g.W + 127) / 128, (g.H + 7) / 8, nImg), dim3(32, 8), 0, s>>>(c->d_lbdBlur, imgBytes, ip, c->d_sobel, g.W, g.H, imgFirst);
    plf_mark(c, "lbd_descriptor");
    lbd_kernel<<<dim3((g.klCap + 3) / 4, nImg), 128, 0, s>>>(g, c->d_sobel, c->d_kl, c->d_nKl, c->d_lbd, c->d_ldesc, imgFirst);
    return launches + 8;
}
