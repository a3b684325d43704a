// Kernels of the rows SURVEY.md §8(f) marks "next" (see capi_next.cu for the entry points): rectification, feature grid,
// projection-window candidates, bag-of-words descent, landmark back-projection.  Small next to the frontend itself.
#include "plf_ctx.cuh"
#include <algorithm>
#include <cstdint>

// ---------------------------------------------------------------------------------------------------------------
// Rank 2: stereo rectification
namespace {
// cv::remap(raw, M1, M2, INTER_LINEAR), BORDER_CONSTANT 0, from the fixed-point table built by plf_rectify_set_maps,
// written straight into level 0 of the pyramid block (4 output pixels per thread, one 32-bit store).  The 15-bit
// weights are products of the two 5-bit fractions, 32 * (32 - fy | fy) * (32 - fx | fx); only the (0, 0) entry
// saturates in OpenCV's table (32768 -> 32767) and its missing unit goes to the last tap: {32767, 0, 0, 1}.
__global__ void __launch_bounds__(256) rectify_kernel(PlfGeom g, const uint8_t* raw0, const uint8_t* raw1, int rawStride,
                                                      const uint2* map0, const uint2* map1, int sw0, int sh0, int sw1,
                                                      int sh1, uint8_t* pyr, int imgFirst) {
    const int x = blockIdx.x * 128 + threadIdx.x * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const int img = imgFirst + blockIdx.z, side = img & 1, frame = blockIdx.z >> 1;
    const int sw = side ? sw1 : sw0, sh = side ? sh1 : sh0;
    const uint8_t* src = (side ? raw1 : raw0) + (size_t)frame * sh * rawStride;
    const uint2* map = (side ? map1 : map0) + (size_t)y * g.W + x;
    uint8_t* dst = pyr + (size_t)img * g.pyrBytes + g.lv[0].off + (size_t)y * g.lv[0].pitch + x;
    const int nValid = min(4, g.W - x);
    unsigned out = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nValid) break;
        const uint2 m = map[j];
        const int sx = (short)(m.x & 0xFFFFu), sy = (short)(m.x >> 16);
        const int fx = (int)(m.y & 31u), fy = (int)(m.y >> 5);
        int w00 = 32 * (32 - fy) * (32 - fx), w01 = 32 * (32 - fy) * fx, w10 = 32 * fy * (32 - fx), w11 = 32 * fy * fx;
        if (m.y == 0u) { w00 = 32767; w11 = 1; }
        int p00, p01, p10, p11;
        const uint8_t* r0 = src + (size_t)sy * rawStride + sx;
        if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
            p00 = r0[0]; p01 = r0[1]; p10 = r0[rawStride]; p11 = r0[rawStride + 1];
        } else {      // a tap outside the source image counts as 0 (BORDER_CONSTANT)
            const bool x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw;
            const bool y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
            p00 = (x0 && y0) ? r0[0] : 0;
            p01 = (x1 && y0) ? r0[1] : 0;
            p10 = (x0 && y1) ? r0[rawStride] : 0;
            p11 = (x1 && y1) ? r0[rawStride + 1] : 0;
        }
        const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;     // <= 255 by construction
        out |= (unsigned)v << (8 * j);
    }
    if (nValid == 4) *reinterpret_cast<unsigned*>(dst) = out;
    else for (int j = 0; j < nValid; ++j) dst[j] = (uint8_t)(out >> (8 * j));
}
}  // namespace

int plf_launch_rectify(plf_ctx* c, const uint8_t* raw0, const uint8_t* raw1, int rawStride, int imgFirst, int nImg) {
    const PlfGeom& g = c->g;
    rectify_kernel<<<dim3((g.W + 127) / 128, (g.H + 7) / 8, nImg), dim3(32, 8), 0, c->stream>>>(
        g, raw0, raw1, rawStride, c->d_rmap[0], c->d_rmap[1], c->srcW[0], c->srcH[0], c->srcW[1], c->srcH[1], c->d_pyr, imgFirst);
    return 1;
}


// ---------------------------------------------------------------------------------------------------------------
// Rank 4: landmark back-projection; rank 1 (first half): the feature grid
namespace {
// Frame::AssignFeaturesToGrid (src/Frame.cc:451-482) for the left keypoints of one slot per block: the 64 x 48 vectors
// of keypoint indices become CSR (cell = x * 48 + y).  Histogram in shared memory, block scan, then one warp places the
// indices in ascending order (lanes sharing a cell are ranked with __match_any_sync), which is the push_back order.
__global__ void __launch_bounds__(256) feature_grid_kernel(PlfGeom g, const plf_keypoint* kp, const int* nKp, int* cellStart,
                                                           int* cellIdx, float invW, float invH, int slotFirst) {
    constexpr int NC = PLF_GRID_COLS * PLF_GRID_ROWS;
    __shared__ int s_cnt[NC + 1];
    __shared__ int s_part[256];
    const int slot = slotFirst + blockIdx.x, img = slot * 2, tid = threadIdx.x;
    const plf_keypoint* K = kp + (size_t)img * g.kpCap;
    const int N = nKp[img];
    int* outStart = cellStart + (size_t)blockIdx.x * (NC + 1);
    int* outIdx = cellIdx + (size_t)blockIdx.x * g.kpCap;
    auto cell_of = [&](int i) -> int {          // PosInGrid (src/Frame.cc:845-855): round() half away from zero
        const int px = (int)roundf(__fmul_rn(__fsub_rn(K[i].x, 0.0f), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(K[i].y, 0.0f), invH));
        return (px < 0 || px >= PLF_GRID_COLS || py < 0 || py >= PLF_GRID_ROWS) ? -1 : px * PLF_GRID_ROWS + py;
    };
    for (int i = tid; i <= NC; i += 256) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += 256) {
        const int c = cell_of(i);
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan: 12 consecutive cells per thread, then the 256 partial sums
    constexpr int PER = NC / 256;
    int loc[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { loc[k] = sum; sum += s_cnt[tid * PER + k]; }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int v = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    const int base = s_part[tid] - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        s_cnt[tid * PER + k] = base + loc[k];              // becomes the write cursor of the cell
        outStart[tid * PER + k] = base + loc[k];
    }
    if (tid == 255) outStart[NC] = s_part[255];
    __syncthreads();
    if (tid < 32) {
        const unsigned lt = (1u << tid) - 1u;
        for (int b = 0; b < N; b += 32) {
            const int i = b + tid;
            const int c = i < N ? cell_of(i) : -1;
            const unsigned grp = __match_any_sync(0xffffffffu, c >= 0 ? c : -1 - tid);
            int at = 0;
            if (c >= 0) at = s_cnt[c];
            __syncwarp();
            if (c >= 0) {
                if ((grp & lt) == 0u) s_cnt[c] = at + __popc(grp);
                outIdx[at + __popc(grp & lt)] = i;
            }
            __syncwarp();
        }
    }
}
}  // namespace

namespace {
struct BackprojArgs { float fx, fy, cx, cy, invfx, invfy, mb; };
// Frame::UnprojectStereo (src/Frame.cc:1332-1347) per left keypoint and Frame::backProjection (:1349-1358) per line end
// point; thread per keypoint / per line, slot = blockIdx.y.  Every operation is written out (no FMA contraction).
__global__ void __launch_bounds__(256) backproject_kernel(PlfGeom g, const plf_keypoint* kp, const int* nKp, const float* depth,
                                                          const plf_keyline* kl, const int* nKl, const float* disp,
                                                          const float* Rwc, const float* Ow, BackprojArgs a, float* x3d,
                                                          int x3dRows, double* l3d, int l3dRows, int slotFirst) {
    const int s = blockIdx.y, slot = slotFirst + s, img = slot * 2;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const float* R = Rwc + s * 9;
    const float* O = Ow + s * 3;
    if (x3d && i < x3dRows) {
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (i < nKp[img]) {
            const float z = depth[(size_t)slot * g.kpCap + i];
            if (z > 0) {
                const plf_keypoint k = kp[(size_t)img * g.kpCap + i];
                const float x = __fmul_rn(__fmul_rn(__fsub_rn(k.x, a.cx), z), a.invfx);
                const float y = __fmul_rn(__fmul_rn(__fsub_rn(k.y, a.cy), z), a.invfy);
                const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(R[0], x), __fmul_rn(R[1], y)), __fmul_rn(R[2], z));
                const float t1 = __fadd_rn(__fadd_rn(__fmul_rn(R[3], x), __fmul_rn(R[4], y)), __fmul_rn(R[5], z));
                const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(R[6], x), __fmul_rn(R[7], y)), __fmul_rn(R[8], z));
                o0 = (float)__dadd_rn((double)t0, (double)O[0]);
                o1 = (float)__dadd_rn((double)t1, (double)O[1]);
                o2 = (float)__dadd_rn((double)t2, (double)O[2]);
            }
        }
        float* d = x3d + ((size_t)s * x3dRows + i) * 3;
        d[0] = o0; d[1] = o1; d[2] = o2;
    }
    if (l3d && i < l3dRows) {
        double o[6] = {0, 0, 0, 0, 0, 0};
        if (nKl && i < nKl[img]) {
            const float d0 = disp[((size_t)slot * g.klCap + i) * 2], d1 = disp[((size_t)slot * g.klCap + i) * 2 + 1];
            if (d0 > 0 && d1 > 0) {
                const plf_keyline k = kl[(size_t)img * g.klCap + i];
                const float uv[4] = {k.startPointX, k.startPointY, k.endPointX, k.endPointY};
                const float dd[2] = {d0, d1};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double bd = (double)a.mb / (double)dd[e];
                    const double P0 = __dmul_rn(bd, __dsub_rn((double)uv[2 * e], (double)a.cx));
                    const double P1 = __dmul_rn(bd, __dsub_rn((double)uv[2 * e + 1], (double)a.cy));
                    const double P2 = __dmul_rn(bd, (double)a.fx);
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        o[3 * e + r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn((double)R[3 * r], P0), __dmul_rn((double)R[3 * r + 1], P1)),
                                                           __dmul_rn((double)R[3 * r + 2], P2)), (double)O[r]);
                }
            }
        }
        double* d = l3d + ((size_t)s * l3dRows + i) * 6;
#pragma unroll
        for (int k2 = 0; k2 < 6; ++k2) d[k2] = o[k2];
    }
}
}  // namespace

int plf_launch_backproject(plf_ctx* c, int slotFirst, int nSlots, const float* dRwc, const float* dOw, float fy, float cx,
                           float cy, float* dX3d, int x3dRows, double* dL3d, int l3dRows) {
    const PlfGeom& g = c->g;
    BackprojArgs a;
    a.fx = c->p.fx; a.fy = fy; a.cx = cx; a.cy = cy;
    a.invfx = 1.0f / a.fx; a.invfy = 1.0f / fy;          // src/Frame.cc:190-191
    a.mb = c->p.bf / c->p.fx;                             // src/Frame.cc:196 (the declared rule mb := mbf / fx)
    const int rows = max(dX3d ? x3dRows : 0, dL3d ? l3dRows : 0);
    if (rows <= 0) return 0;
    backproject_kernel<<<dim3((rows + 255) / 256, nSlots), 256, 0, c->stream>>>(
        g, c->d_kp, c->d_nKp, c->d_depth, c->d_kl, c->p.has_lines ? c->d_nKl : nullptr, c->d_disp, dRwc, dOw, a, dX3d, x3dRows,
        dL3d, l3dRows, slotFirst);
    return 1;
}

int plf_launch_feature_grid(plf_ctx* c, int slotFirst, int nSlots, int* cellStart, int* cellIdx) {
    const PlfGeom& g = c->g;
    const float invW = (float)PLF_GRID_COLS / ((float)g.W - 0.0f), invH = (float)PLF_GRID_ROWS / ((float)g.H - 0.0f);
    feature_grid_kernel<<<nSlots, 256, 0, c->stream>>>(g, c->d_kp, c->d_nKp, cellStart, cellIdx, invW, invH, slotFirst);
    return 1;
}


// ---------------------------------------------------------------------------------------------------------------
// Bag-of-words descent (DBoW2 TemplatedVocabulary::transform(feature, id, weight, nid, levelsup),
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1230-1270): thread per descriptor; at every level the child with the
// smallest Hamming distance is taken (strict <: the first child wins ties).  The tree is read-only and small next to
// the frame data (ORBvoc: ~1.1 M nodes x 32 B), the features of a frame walk it independently.
namespace {
__global__ void __launch_bounds__(256) bow_kernel(const uint8_t* desc, const int* nFeat, int cap, int imgStride, const int* childFirst,
                                                  const int* childCount, const int* child, const uint8_t* nodeDesc,
                                                  const int* nodeWord, const double* nodeWeight, int levels, int levelsup,
                                                  int* outWord, double* outWeight, int* outNode, int rows, int slotFirst) {
    const int s = blockIdx.y, img = (slotFirst + s) * imgStride;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= rows) return;
    const size_t o = (size_t)s * rows + i;
    if (i >= nFeat[img]) { outWord[o] = -1; outWeight[o] = 0.0; outNode[o] = 0; return; }
    const uint4* f4 = reinterpret_cast<const uint4*>(desc + ((size_t)img * cap + i) * 32);
    const uint4 a = f4[0], b = f4[1];
    const int nidLevel = levels - levelsup;
    int node = 0, level = 0, nid = 0;
    do {
        ++level;
        const int c0 = childFirst[node], nc = childCount[node];
        int best = 0x7fffffff, bestId = node;
        for (int k = 0; k < nc; ++k) {
            const int id = child[c0 + k];
            const uint4* n4 = reinterpret_cast<const uint4*>(nodeDesc + (size_t)id * 32);
            const uint4 p = n4[0], q = n4[1];
            const int d = __popc(a.x ^ p.x) + __popc(a.y ^ p.y) + __popc(a.z ^ p.z) + __popc(a.w ^ p.w) +
                          __popc(b.x ^ q.x) + __popc(b.y ^ q.y) + __popc(b.z ^ q.z) + __popc(b.w ^ q.w);
            if (d < best) { best = d; bestId = id; }
        }
        node = bestId;
        if (level == nidLevel) nid = node;
    } while (childCount[node] > 0);
    outWord[o] = nodeWord[node];
    outWeight[o] = nodeWeight[node];
    outNode[o] = nid;
}
}  // namespace

int plf_launch_bow(plf_ctx* c, int which, int slotFirst, int nSlots, int levelsup, int* dWord, double* dWeight, int* dNode, int rows) {
    const PlfGeom& g = c->g;
    const PlfVocab& v = c->voc[which];
    const uint8_t* desc = which ? c->d_ldesc : c->d_desc;
    const int* nFeat = which ? c->d_nKl : c->d_nKp;
    const int cap = which ? g.klCap : g.kpCap;
    bow_kernel<<<dim3((rows + 255) / 256, nSlots), 256, 0, c->stream>>>(desc, nFeat, cap, 2, v.childFirst, v.childCount, v.child, v.desc,
                                                                       v.word, v.weight, v.levels, levelsup, dWord, dWeight, dNode,
                                                                       rows, slotFirst);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// ORBmatcher::SearchByProjection, device half of both overloads built (src/ORBmatcher.cc:44-130 and :2179-2323): a warp
// per map point walks the grid cells of its window in GetFeaturesInArea order (src/Frame.cc:774-843: cell columns, cell
// rows, insertion order), lanes take the features of a cell 32 at a time, apply the level, window and stereo filters and
// compute the Hamming distances; survivors are compacted in order.  FILL = false only counts (the host sizes the
// candidate pool from the counts), FILL = true writes (feature index, distance | octave << 16).
namespace {
template <bool FILL>
__global__ void __launch_bounds__(256) proj_candidates_kernel(PlfGeom g, const PlfWinQ* qs, int nq, const plf_keypoint* kp,
                                                              const uint8_t* desc, const float* uRight, const int* cellStart,
                                                              const int* cellIdx, int* count, const int* segStart, int2* pool) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const PlfWinQ q = qs[qi];
    int total = 0;
    if (!q.skip) {
        const float rad = q.radius;
        const float invW = (float)PLF_GRID_COLS / ((float)g.W - 0.0f), invH = (float)PLF_GRID_ROWS / ((float)g.H - 0.0f);
        int x0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.x, 0.0f), rad), invW));
        int x1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.x, 0.0f), rad), invW));
        int y0 = (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(q.y, 0.0f), rad), invH));
        int y1 = (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(q.y, 0.0f), rad), invH));
        x0 = max(x0, 0); y0 = max(y0, 0);
        x1 = min(x1, PLF_GRID_COLS - 1); y1 = min(y1, PLF_GRID_ROWS - 1);
        const bool window = x0 < PLF_GRID_COLS && x1 >= 0 && y0 < PLF_GRID_ROWS && y1 >= 0;
        const int minLevel = q.minLevel, maxLevel = q.maxLevel;
        const bool check = minLevel > 0 || maxLevel >= 0;
        const uint4 a = make_uint4(q.desc[0], q.desc[1], q.desc[2], q.desc[3]), b = make_uint4(q.desc[4], q.desc[5], q.desc[6], q.desc[7]);
        const unsigned lt = (1u << lane) - 1u;
        int2* out = FILL ? pool + segStart[qi] : nullptr;
        if (window)
            for (int ix = x0; ix <= x1; ++ix)
                for (int iy = y0; iy <= y1; ++iy) {
                    const int c = ix * PLF_GRID_ROWS + iy;
                    const int j0 = cellStart[c], j1 = cellStart[c + 1];
                    for (int jb = j0; jb < j1; jb += 32) {
                        const int j = jb + lane;
                        bool ok = false;
                        int idx = -1, oct = 0;
                        if (j < j1) {
                            idx = cellIdx[j];
                            const plf_keypoint k = kp[idx];
                            oct = k.octave;
                            ok = !(check && (oct < minLevel || (maxLevel >= 0 && oct > maxLevel)));
                            ok = ok && fabsf(__fsub_rn(k.x, q.x)) < rad && fabsf(__fsub_rn(k.y, q.y)) < rad;
                            if (ok && !(q.pad & 1)) {          // pad bit 0: the overload has no stereo check
                                const float ur = uRight[idx];
                                if (ur > 0 && fabsf(__fsub_rn(q.xr, ur)) > rad) ok = false;
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, ok);
                        if (FILL && ok) {
                            const uint4* f4 = reinterpret_cast<const uint4*>(desc + (size_t)idx * 32);
                            const uint4 p = f4[0], s2 = f4[1];
                            const int d = __popc(a.x ^ p.x) + __popc(a.y ^ p.y) + __popc(a.z ^ p.z) + __popc(a.w ^ p.w) +
                                          __popc(b.x ^ s2.x) + __popc(b.y ^ s2.y) + __popc(b.z ^ s2.z) + __popc(b.w ^ s2.w);
                            out[total + __popc(m & lt)] = make_int2(idx, d | (oct << 16));
                        }
                        total += __popc(m);
                    }
                }
    }
    if (!FILL && lane == 0) count[qi] = total;
}
}  // namespace

int plf_launch_proj_candidates(plf_ctx* c, int slot, const PlfWinQ* dQ, int nq, const int* dCellStart,
                               const int* dCellIdx, int* dCount, const int* dSegStart, int2* dPool, bool fill) {
    const PlfGeom& g = c->g;
    const plf_keypoint* kp = c->d_kp + (size_t)(slot * 2) * g.kpCap;
    const uint8_t* desc = c->d_desc + (size_t)(slot * 2) * g.kpCap * 32;
    const float* ur = c->d_uRight + (size_t)slot * g.kpCap;
    const dim3 grid((nq + 7) / 8);
    if (fill) proj_candidates_kernel<true><<<grid, 256, 0, c->stream>>>(g, dQ, nq, kp, desc, ur, dCellStart, dCellIdx, dCount, dSegStart, dPool);
    else proj_candidates_kernel<false><<<grid, 256, 0, c->stream>>>(g, dQ, nq, kp, desc, ur, dCellStart, dCellIdx, dCount, dSegStart, dPool);
    return 1;
}

// ---- SearchByBoW, device half: a warp per valid keyframe feature; its descriptor against the frame features of the same
// vocabulary node (job = {keyframe feature, begin, end in the node-sorted frame order, first slot of the distance pool})
namespace {
__global__ void __launch_bounds__(256) bow_pairs_kernel(const uint8_t* kfDesc, const int4* jobs, int nJobs, const int* order,
                                                        const uint8_t* desc, int* pool) {
    const int ji = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ji >= nJobs) return;
    const int4 j = jobs[ji];
    const uint4* a4 = reinterpret_cast<const uint4*>(kfDesc + (size_t)j.x * 32);
    const uint4 a = a4[0], b = a4[1];
    for (int k = j.y + lane; k < j.z; k += 32) {
        const uint4* f4 = reinterpret_cast<const uint4*>(desc + (size_t)order[k] * 32);
        const uint4 p = f4[0], s2 = f4[1];
        pool[j.w + (k - j.y)] = __popc(a.x ^ p.x) + __popc(a.y ^ p.y) + __popc(a.z ^ p.z) + __popc(a.w ^ p.w) +
                                __popc(b.x ^ s2.x) + __popc(b.y ^ s2.y) + __popc(b.z ^ s2.z) + __popc(b.w ^ s2.w);
    }
}

// ---- epilogue of the tracking thread's line matching: mutual-best filter of match() (src/LineMatcher.cpp:218-224, mode 0
// only: the MapLine overload of mode 1 returns after the one-way matchNNR, :161-170) and the gates of
// src/Tracking.cc:3062-3098 (mode 0) / :3888-3919 (mode 1), a thread per line of the first set.
// Mode 0 is order-free (a mutual-best i2 belongs to one i1).  In mode 1 several lines may hold the same keyline i2 and the
// reference loop is sequential: a line is looked at only while the keyline's holder has no observations.  With
// pass(i1) = "disparities >= 0 and inside the position gate", the first i1 (in order) of a keyline that passes AND has
// observations blocks everything after it, every earlier one is looked at, and the keyline ends up with the last passing
// line at or before the blocker: an atomicMin, an atomicMax and three small passes instead of a serial loop.
__device__ __forceinline__ bool line_gate_pass(int mode, const plf_track_line& a, const plf_keyline& b, double deltaW, double deltaH) {
    const double kPiD = 3.14159265358979323846;
    if (mode == 0) {
        double theta = (double)__fsub_rn(b.angle, a.angle);
        if (theta < -kPiD) theta += 2 * kPiD;
        else if (theta > kPiD) theta -= 2 * kPiD;
        if (fabs(theta) > kPiD / 8.0) return false;
    }
    return !((double)fabsf(__fsub_rn(b.startPointX, a.sx)) > deltaW || (double)fabsf(__fsub_rn(b.endPointX, a.ex)) > deltaW ||
             (double)fabsf(__fsub_rn(b.startPointY, a.sy)) > deltaH || (double)fabsf(__fsub_rn(b.endPointY, a.ey)) > deltaH);
}

// pass 1.  state[i1]: 0 = not looked at by the gates (no match / not eligible / negative disparity), 1 = fails, 2 = passes.
// Mode 0 finishes here; mode 1 records the blocker of every keyline.
__global__ void __launch_bounds__(128) line_gates_kernel(int mode, const plf_track_line* l1, int n1, const plf_keyline* k2,
                                                         const float2* disp2, double deltaW, double deltaH, int* m12,
                                                         const int* m21, int* assign, int* state, int* blocker) {
    const int i1 = blockIdx.x * 128 + threadIdx.x;
    if (i1 >= n1) return;
    int i2 = m12[i1];
    if (mode == 0 && i2 >= 0 && m21[i2] != i1) i2 = -1;
    int as = -1, st = 0;
    const plf_track_line a = l1[i1];
    if (i2 >= 0 && (mode == 1 || a.eligible)) {
        const float2 d = disp2[i2];
        if (!(d.x < 0 || d.y < 0)) {
            const bool ok = line_gate_pass(mode, a, k2[i2], deltaW, deltaH);
            st = ok ? 2 : 1;
            if (mode == 0) { if (ok) as = i2; else i2 = -1; }
            else if (ok && a.eligible) atomicMin(blocker + i2, i1);
        }
    }
    m12[i1] = i2;
    assign[i1] = as;
    if (mode == 1) state[i1] = st;
}
// pass 2 (mode 1): lines at or before their keyline's blocker are looked at: a failing one loses its match, the last
// passing one is the keyline's final holder
__global__ void __launch_bounds__(128) line_gates_resolve_kernel(int n1, const uint8_t* held2, int* m12, const int* state,
                                                                 const int* blocker, int* last) {
    const int i1 = blockIdx.x * 128 + threadIdx.x;
    if (i1 >= n1) return;
    const int st = state[i1];
    if (st == 0) return;
    const int i2 = m12[i1];
    if ((held2 && held2[i2]) || i1 > blocker[i2]) return;           // the holder has observations: `continue`
    if (st == 1) m12[i1] = -1;
    else atomicMax(last + i2, i1);
}
// pass 3 (mode 1): mCurrentFrame.mvpMapLines[i2] when the loop ends
__global__ void __launch_bounds__(128) line_gates_assign_kernel(int n1, const int* m12, const int* state, const int* last, int* assign) {
    const int i1 = blockIdx.x * 128 + threadIdx.x;
    if (i1 >= n1) return;
    const int i2 = m12[i1];
    assign[i1] = (state[i1] == 2 && i2 >= 0 && last[i2] == i1) ? i2 : -1;
}
}  // namespace

int plf_launch_bow_pairs(plf_ctx* c, int slot, const uint8_t* dKfDesc, const int4* dJobs, int nJobs, const int* dOrder, int* dPool) {
    if (nJobs <= 0) return 0;
    const uint8_t* desc = c->d_desc + (size_t)(slot * 2) * c->g.kpCap * 32;
    bow_pairs_kernel<<<(nJobs + 7) / 8, 256, 0, c->stream>>>(dKfDesc, dJobs, nJobs, dOrder, desc, dPool);
    return 1;
}

int plf_launch_line_gates(plf_ctx* c, int mode, const plf_track_line* dL1, int n1, const plf_keyline* dK2, const float2* dDisp2,
                          const uint8_t* dHeld2, int n2, float minX, float maxX, float minY, float maxY, int* dM12, const int* dM21,
                          int* dAssign, int* dState, int* dBlocker, int* dLast) {
    if (n1 <= 0) return 0;
    const double deltaW = (double)(maxX - minX) * 0.1, deltaH = (double)(maxY - minY) * 0.1;     // float difference times the double 0.1
    const int nb = (n1 + 127) / 128;
    if (mode == 0) {
        line_gates_kernel<<<nb, 128, 0, c->stream>>>(0, dL1, n1, dK2, dDisp2, deltaW, deltaH, dM12, dM21, dAssign, nullptr, nullptr);
        return 1;
    }
    cudaMemsetAsync(dBlocker, 0x7F, (size_t)std::max(n2, 1) * sizeof(int), c->stream);          // "no blocker": larger than any i1
    cudaMemsetAsync(dLast, 0xFF, (size_t)std::max(n2, 1) * sizeof(int), c->stream);             // -1
    line_gates_kernel<<<nb, 128, 0, c->stream>>>(1, dL1, n1, dK2, dDisp2, deltaW, deltaH, dM12, dM21, dAssign, dState, dBlocker);
    line_gates_resolve_kernel<<<nb, 128, 0, c->stream>>>(n1, dHeld2, dM12, dState, dBlocker, dLast);
    line_gates_assign_kernel<<<nb, 128, 0, c->stream>>>(n1, dM12, dState, dLast, dAssign);
    return 3;
}
