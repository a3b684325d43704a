// C ABI, second part: the entry points of the rows SURVEY.md §8(f) marks "next" — the callers and data formats either side
// of the stereo frontend: rectification in front of it (rank 2), the feature grid and the projection-window search
// (rank 1), the bag-of-words transform (rank 3) and the landmark back-projection (rank 4).  Same rules as capi.cu: plain
// C signatures, loud errors, no CPU fallback.
#include "plf_ctx.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

static int fail(int code, const char* msg) { return plf_fail(code, msg); }

extern "C" {

// ---- Frame::AssignFeaturesToGrid (first half) ---------------------------------------------------------
PLF_API int plf_feature_grid(plf_ctx* c, int first_slot, int n_slots, int32_t* cell_start, int32_t* cell_idx, int idx_stride) {
    if (!c || !cell_start || !cell_idx || first_slot < 0 || n_slots < 1 || first_slot + n_slots > c->p.max_batch ||
        idx_stride < c->g.kpCap)
        return fail(PLF_ERR_INVALID, "bad slot range / idx_stride smaller than plf_keypoint_capacity");
    if (!c->orbValid[0]) return fail(PLF_ERR_STATE, "feature_grid before the left keypoints were extracted");
    PLF_CUDA_OK(plf_enter(c));
    constexpr int NC1 = PLF_GRID_COLS * PLF_GRID_ROWS + 1;
    if (!c->d_gridStart) {
        PLF_CUDA_OK(dalloc(&c->d_gridStart, (size_t)c->p.max_batch * NC1));
        PLF_CUDA_OK(dalloc(&c->d_gridIdx, (size_t)c->p.max_batch * c->g.kpCap));
    }
    plf_launch_feature_grid(c, first_slot, n_slots, c->d_gridStart, c->d_gridIdx);
    PLF_CUDA_OK(cudaMemcpyAsync(cell_start, c->d_gridStart, (size_t)n_slots * NC1 * 4, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaMemcpy2DAsync(cell_idx, (size_t)idx_stride * 4, c->d_gridIdx, (size_t)c->g.kpCap * 4, (size_t)c->g.kpCap * 4,
                                  n_slots, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaGetLastError());
    return PLF_OK;
}

PLF_API int plf_get_features_in_area(const plf_keypoint* kps, const int32_t* cell_start, const int32_t* cell_idx, int width,
                                     int height, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    return plf_features_in_area(kps, cell_start, cell_idx, width, height, x, y, r, min_level, max_level, out, cap);
}

// ---- projection-window search (second half) ----------------------------------------------------------
// device half shared by the two overloads: feature grid of the slot, candidate counts, candidate pool (CSR by query)
static int window_candidates(plf_ctx* c, int slot, const std::vector<PlfWinQ>& q, std::vector<int>& start, std::vector<int2>& pool) {
    const int nq = (int)q.size();
    PLF_CUDA_OK(plf_enter(c));
    cudaStream_t s = c->stream;
    constexpr int NC1 = PLF_GRID_COLS * PLF_GRID_ROWS + 1;
    if (!c->d_gridStart) {
        PLF_CUDA_OK(dalloc(&c->d_gridStart, (size_t)c->p.max_batch * NC1));
        PLF_CUDA_OK(dalloc(&c->d_gridIdx, (size_t)c->p.max_batch * c->g.kpCap));
    }
    if (c->projQCap < (size_t)nq) {
        PLF_CUDA_OK(cudaStreamSynchronize(s));
        if (c->d_projQ) { cudaFree(c->d_projQ); cudaFree(c->d_projCount); cudaFree(c->d_projStart); }
        c->d_projQ = nullptr; c->d_projCount = nullptr; c->d_projStart = nullptr; c->projQCap = 0;
        PLF_CUDA_OK(dalloc(&c->d_projQ, (size_t)nq));
        PLF_CUDA_OK(dalloc(&c->d_projCount, (size_t)nq));
        PLF_CUDA_OK(dalloc(&c->d_projStart, (size_t)nq));
        c->projQCap = nq;
    }
    // Frame::AssignFeaturesToGrid for this slot (slot-local CSR at the start of the grid buffers)
    plf_launch_feature_grid(c, slot, 1, c->d_gridStart, c->d_gridIdx);
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_projQ, q.data(), (size_t)nq * sizeof(PlfWinQ), cudaMemcpyHostToDevice, s));
    plf_launch_proj_candidates(c, slot, c->d_projQ, nq, c->d_gridStart, c->d_gridIdx, c->d_projCount, nullptr, nullptr, false);
    std::vector<int> cnt(nq);
    start.assign(nq + 1, 0);
    PLF_CUDA_OK(cudaMemcpyAsync(cnt.data(), c->d_projCount, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    for (int i = 0; i < nq; ++i) start[i + 1] = start[i] + cnt[i];
    const size_t total = (size_t)start[nq];
    pool.resize(total);
    if (total) {
        if (c->projPoolCap < total) {
            if (c->d_projPool) cudaFree(c->d_projPool);
            c->d_projPool = nullptr; c->projPoolCap = 0;
            PLF_CUDA_OK(dalloc(&c->d_projPool, total * 2));
            c->projPoolCap = total * 2;
        }
        PLF_CUDA_OK(cudaMemcpyAsync(c->d_projStart, start.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, s));
        plf_launch_proj_candidates(c, slot, c->d_projQ, nq, c->d_gridStart, c->d_gridIdx, c->d_projCount, c->d_projStart,
                                   c->d_projPool, true);
        PLF_CUDA_OK(cudaMemcpyAsync(pool.data(), c->d_projPool, total * sizeof(int2), cudaMemcpyDeviceToHost, s));
        PLF_CUDA_OK(cudaStreamSynchronize(s));
    }
    PLF_CUDA_OK(cudaGetLastError());
    return PLF_OK;
}

PLF_API int plf_search_by_projection(plf_ctx* c, int slot, const plf_proj_query* queries, int n_queries, float th, float nn_ratio,
                                     int th_high, uint8_t* occupied, int n_features, int32_t* match, int* n_matches) {
    if (!c || slot < 0 || slot >= c->p.max_batch || !queries || n_queries < 0 || !occupied || !match || n_features < 0)
        return fail(PLF_ERR_INVALID, "bad arguments");
    if (!c->orbValid[0] || !c->orbValid[1]) return fail(PLF_ERR_STATE, "search_by_projection before the frame was extracted and stereo-matched");
    if (n_matches) *n_matches = 0;
    {   // occupied[] is indexed by feature: it must cover the slot's left keypoints
        int N = 0;
        PLF_CUDA_OK(plf_enter(c));
        PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
        PLF_CUDA_OK(cudaMemcpy(&N, c->d_nKp + slot * 2, 4, cudaMemcpyDeviceToHost));
        if (n_features < N) return fail(PLF_ERR_INVALID, "occupied[] is shorter than the slot's keypoint count");
    }
    if (n_queries == 0) return PLF_OK;
    std::vector<PlfWinQ> q(n_queries);
    for (int i = 0; i < n_queries; ++i) {
        const plf_proj_query& m = queries[i];
        PlfWinQ& w = q[i];
        w.skip = m.skip || m.level < 0 || m.level >= c->g.nLevels;
        float r = m.view_cos > 0.998f ? 2.5f : 4.0f;            // RadiusByViewingCos, src/ORBmatcher.cc:216-222
        if (th != 1.0f) r *= th;
        w.x = m.proj_x; w.y = m.proj_y; w.xr = m.proj_xr;
        w.radius = w.skip ? 0.f : r * c->scale[m.level];
        w.minLevel = m.level - 1; w.maxLevel = m.level; w.pad = 0;
        memcpy(w.desc, m.desc, 32);
    }
    std::vector<int> start;
    std::vector<int2> pool;
    const int rc = window_candidates(c, slot, q, start, pool);
    if (rc) return rc;
    // the order-dependent half, in query order (src/ORBmatcher.cc:84-129)
    int nm = 0;
    for (int i = 0; i < n_queries; ++i) {
        match[i] = -1;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int j = start[i]; j < start[i + 1]; ++j) {
            const int idx = pool[j].x, dist = pool[j].y & 0xFFFF, oct = pool[j].y >> 16;
            if (occupied[idx]) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = oct; bestDist2 = dist; }
        }
        if (bestDist <= th_high) {
            if (bestLevel == bestLevel2 && (float)bestDist > nn_ratio * (float)bestDist2) continue;
            if (bestLevel != bestLevel2 || (float)bestDist <= nn_ratio * (float)bestDist2) {
                match[i] = bestIdx;
                occupied[bestIdx] = 1;
                ++nm;
            }
        }
    }
    if (n_matches) *n_matches = nm;
    return PLF_OK;
}

}  // extern "C"

// The three overloads that take already projected points (frame-to-frame :2179-2323, relocalisation :2325-2447, loop
// closing :473-704) share everything but four switches.
struct ProjSearchRule {
    bool stereo;          // |ur - mvuRight| <= radius for features with a right match (frame-to-frame only, :2262-2267)
    bool holderByObs;     // a feature is blocked iff its holder has observations (frame-to-frame); else any holder blocks
    bool orientation;     // rotation histogram + ComputeThreeMaxima
    float maxDist;        // accepted iff (float)bestDist <= maxDist (TH_HIGH, ORBdist, or TH_LOW * ratioHamming in float)
};

static int search_projected(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, const ProjSearchRule& rule,
                            uint8_t* occupied, int n_features, int32_t* feat_query, int32_t* match12, int* n_matches) {
    if (!c || slot < 0 || slot >= c->p.max_batch || !queries || n_queries < 0 || !occupied || !feat_query || n_features < 0)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const int check_orientation = rule.orientation;
    if (!c->orbValid[0] || !c->orbValid[1]) return fail(PLF_ERR_STATE, "search_by_projection before the frame was extracted and stereo-matched");
    PLF_CUDA_OK(plf_enter(c));
    // the current frame's keypoint count and angles (rotation histogram)
    const int img = slot * 2;
    int N = 0;
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(&N, c->d_nKp + img, 4, cudaMemcpyDeviceToHost));
    if (n_features < N) return fail(PLF_ERR_INVALID, "occupied[] / feat_query[] are shorter than the slot's keypoint count");
    std::vector<plf_keypoint> kps((size_t)std::max(N, 1));
    if (N > 0) PLF_CUDA_OK(cudaMemcpy(kps.data(), c->d_kp + (size_t)img * c->g.kpCap, (size_t)N * sizeof(plf_keypoint), cudaMemcpyDeviceToHost));
    for (int f = 0; f < N; ++f) { feat_query[f] = -1; if (match12) match12[f] = -1; }
    if (n_matches) *n_matches = 0;
    if (n_queries == 0) return PLF_OK;
    std::vector<PlfWinQ> q(n_queries);
    for (int i = 0; i < n_queries; ++i) {
        const plf_frame_query& m = queries[i];
        PlfWinQ& w = q[i];
        w.skip = m.skip; w.x = m.u; w.y = m.v; w.xr = m.ur; w.radius = m.radius;
        w.minLevel = m.min_level; w.maxLevel = m.max_level; w.pad = rule.stereo ? 0 : 1;      // bit 0: no stereo check
        memcpy(w.desc, m.desc, 32);
    }
    std::vector<int> start;
    std::vector<int2> pool;
    const int rc = window_candidates(c, slot, q, start, pool);
    if (rc) return rc;
    // src/ORBmatcher.cc:2246-2290 in query order, then the rotation consistency of :2293-2317
    constexpr int HISTO = 30;
    std::vector<int> rotHist[HISTO];
    const float factor = 1.0f / HISTO;
    int nm = 0;
    for (int i = 0; i < n_queries; ++i) {
        if (queries[i].skip) continue;
        int bestDist = 256, bestIdx = -1;
        for (int j = start[i]; j < start[i + 1]; ++j) {
            const int idx = pool[j].x, dist = pool[j].y & 0xFFFF;
            if (occupied[idx]) continue;
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestIdx >= 0 && (float)bestDist <= rule.maxDist) {
            feat_query[bestIdx] = i;
            occupied[bestIdx] = (!rule.holderByObs || queries[i].has_observations) ? 1 : 0;     // the holder decides whether later points skip it
            ++nm;
            if (match12 && match12[bestIdx] < 0) match12[bestIdx] = i;    // std::map::insert keeps the first
            if (check_orientation) {
                float rot = queries[i].angle - kps[bestIdx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO) bin = 0;
                if (bin >= 0 && bin < HISTO) rotHist[bin].push_back(bestIdx);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;          // ComputeThreeMaxima, :2449-2490
        for (int i = 0; i < HISTO; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int f : rotHist[i]) {
                    feat_query[f] = -1;
                    occupied[f] = 0;
                    --nm;
                    if (match12) match12[f] = -1;
                }
    }
    if (n_matches) *n_matches = nm;
    return PLF_OK;
}

extern "C" {

PLF_API int plf_search_by_projection_frame(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int th_high,
                                           int check_orientation, uint8_t* occupied, int n_features, int32_t* feat_query,
                                           int32_t* match12, int* n_matches) {
    const ProjSearchRule rule = {true, true, check_orientation != 0, (float)th_high};
    return search_projected(c, slot, queries, n_queries, rule, occupied, n_features, feat_query, match12, n_matches);
}

PLF_API int plf_search_by_projection_reloc(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int orb_dist,
                                           int check_orientation, uint8_t* occupied, int n_features, int32_t* feat_query,
                                           int* n_matches) {
    const ProjSearchRule rule = {false, false, check_orientation != 0, (float)orb_dist};
    return search_projected(c, slot, queries, n_queries, rule, occupied, n_features, feat_query, nullptr, n_matches);
}

PLF_API int plf_search_by_projection_loop(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int th_low,
                                          float ratio_hamming, uint8_t* occupied, int n_features, int32_t* feat_query,
                                          int* n_matches) {
    const ProjSearchRule rule = {false, false, false, (float)th_low * ratio_hamming};
    return search_projected(c, slot, queries, n_queries, rule, occupied, n_features, feat_query, nullptr, n_matches);
}

// grow-only device scratch for the small per-frame searches below
static cudaError_t scratch_reserve(plf_ctx* c, size_t bytes) {
    if (c->scrCap >= bytes) return cudaSuccess;
    cudaStreamSynchronize(c->stream);
    if (c->d_scr) cudaFree(c->d_scr);
    c->d_scr = nullptr;
    c->scrCap = 0;
    const size_t want = std::max(bytes * 2, (size_t)1 << 20);
    cudaError_t e = cudaMalloc((void**)&c->d_scr, want);
    if (e == cudaSuccess) c->scrCap = want;
    return e;
}
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (src/ORBmatcher.cc:269-471) ---------------------------------------
PLF_API int plf_search_by_bow(plf_ctx* c, int slot, const uint8_t* kf_desc, const float* kf_angle, const int32_t* kf_node,
                              const uint8_t* kf_valid, int n_kf, const int32_t* f_node, int n_features, int th_low,
                              float nn_ratio, int check_orientation, int32_t* match, int* n_matches) {
    if (!c || slot < 0 || slot >= c->p.max_batch || n_kf < 0 || n_features < 0 || (n_kf && (!kf_desc || !kf_angle || !kf_node || !kf_valid)) ||
        (n_features && (!f_node || !match)))
        return fail(PLF_ERR_INVALID, "bad arguments");
    if (!c->orbValid[0]) return fail(PLF_ERR_STATE, "search_by_bow before the frame was extracted");
    PLF_CUDA_OK(plf_enter(c));
    cudaStream_t s = c->stream;
    const int img = slot * 2;
    int N = 0;
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    PLF_CUDA_OK(cudaMemcpy(&N, c->d_nKp + img, 4, cudaMemcpyDeviceToHost));
    if (n_features < N) return fail(PLF_ERR_INVALID, "f_node[] / match[] are shorter than the slot's keypoint count");
    for (int f = 0; f < n_features; ++f) match[f] = -1;
    if (n_matches) *n_matches = 0;
    if (n_kf == 0 || N == 0) return PLF_OK;
    // F.mFeatVec as CSR: nodes ascending (std::map order), features of a node ascending (addFeature in feature order)
    std::vector<int> order;
    order.reserve(N);
    for (int f = 0; f < N; ++f) if (f_node[f] >= 0) order.push_back(f);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return f_node[a] < f_node[b]; });
    std::vector<int> kfOrder;
    for (int i = 0; i < n_kf; ++i) if (kf_node[i] >= 0) kfOrder.push_back(i);
    std::stable_sort(kfOrder.begin(), kfOrder.end(), [&](int a, int b) { return kf_node[a] < kf_node[b]; });
    // per valid keyframe feature (in the reference's visiting order): its frame range and its slice of the distance pool
    struct Job { int kf, begin, end, pool; };
    std::vector<Job> jobs;
    size_t pool = 0;
    {
        size_t a = 0, b = 0;
        while (a < kfOrder.size() && b < order.size()) {
            const int na = kf_node[kfOrder[a]], nb = f_node[order[b]];
            if (na < nb) { ++a; continue; }
            if (nb < na) { ++b; continue; }
            size_t b1 = b;
            while (b1 < order.size() && f_node[order[b1]] == na) ++b1;
            for (; a < kfOrder.size() && kf_node[kfOrder[a]] == na; ++a)
                if (kf_valid[kfOrder[a]]) { jobs.push_back({kfOrder[a], (int)b, (int)b1, (int)pool}); pool += b1 - b; }
            b = b1;
        }
    }
    if (jobs.empty()) return PLF_OK;
    const size_t oDesc = 0, oJobs = align256((size_t)n_kf * 32), oOrder = oJobs + align256(jobs.size() * sizeof(Job));
    const size_t oPool = oOrder + align256(order.size() * 4), total = oPool + align256(pool * 4);
    PLF_CUDA_OK(scratch_reserve(c, total));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_scr + oDesc, kf_desc, (size_t)n_kf * 32, cudaMemcpyHostToDevice, s));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_scr + oJobs, jobs.data(), jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, s));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_scr + oOrder, order.data(), order.size() * 4, cudaMemcpyHostToDevice, s));
    c->launches = plf_launch_bow_pairs(c, slot, c->d_scr + oDesc, reinterpret_cast<const int4*>(c->d_scr + oJobs), (int)jobs.size(),
                                       reinterpret_cast<const int*>(c->d_scr + oOrder), reinterpret_cast<int*>(c->d_scr + oPool));
    std::vector<int> dist(pool);
    std::vector<plf_keypoint> kps((size_t)N);
    PLF_CUDA_OK(cudaMemcpyAsync(dist.data(), c->d_scr + oPool, pool * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaMemcpyAsync(kps.data(), c->d_kp + (size_t)img * c->g.kpCap, (size_t)N * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    PLF_CUDA_OK(cudaGetLastError());
    // the order-dependent half (src/ORBmatcher.cc:296-402 for F.Nleft == -1, then :447-467)
    constexpr int HISTO = 30;
    std::vector<int> rotHist[HISTO];
    const float factor = 1.0f / HISTO;
    int nm = 0;
    for (const Job& j : jobs) {
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int k = j.begin; k < j.end; ++k) {
            const int f = order[k];
            if (match[f] >= 0) continue;
            const int d = dist[j.pool + (k - j.begin)];
            if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = f; }
            else if (d < bestDist2) bestDist2 = d;
        }
        if (bestDist1 <= th_low && (float)bestDist1 < nn_ratio * (float)bestDist2) {
            match[bestIdxF] = j.kf;
            if (check_orientation) {
                float rot = kf_angle[j.kf] - kps[bestIdxF].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO) bin = 0;
                if (bin >= 0 && bin < HISTO) rotHist[bin].push_back(bestIdxF);
            }
            ++nm;
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;          // ComputeThreeMaxima, :2449-2490
        for (int i = 0; i < HISTO; ++i) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int f : rotHist[i]) { match[f] = -1; --nm; }
    }
    if (n_matches) *n_matches = nm;
    return PLF_OK;
}

// ---- match() + the tracking thread's gates (src/Tracking.cc:3055-3099, :3879-3917) -------------------------------------
PLF_API int plf_match_lines_tracked(plf_ctx* c, int mode, const uint8_t* desc1, const plf_track_line* lines1, int n1,
                                    const uint8_t* desc2, const plf_keyline* kl2, const float* disp2, const uint8_t* held2, int n2,
                                    float nnr, float min_x, float max_x, float min_y, float max_y, int32_t* matches12,
                                    int32_t* assign12, int* n_assigned) {
    if (!c || mode < 0 || mode > 1 || n1 < 0 || n2 < 0 || (n1 && (!desc1 || !lines1 || !matches12 || !assign12)) ||
        (n2 && (!desc2 || !kl2 || !disp2)))
        return fail(PLF_ERR_INVALID, "bad arguments");
    if (n_assigned) *n_assigned = 0;
    if (n1 == 0) return PLF_OK;
    PLF_CUDA_OK(plf_enter(c));
    cudaStream_t s = c->stream;
    const size_t oD1 = 0, oD2 = oD1 + align256((size_t)n1 * 32), oL1 = oD2 + align256((size_t)n2 * 32);
    const size_t oK2 = oL1 + align256((size_t)n1 * sizeof(plf_track_line)), oDisp = oK2 + align256((size_t)n2 * sizeof(plf_keyline));
    const size_t oHeld = oDisp + align256((size_t)n2 * 8), oM12 = oHeld + align256((size_t)n2);
    const size_t oM21 = oM12 + align256((size_t)n1 * 4), oAsg = oM21 + align256((size_t)n2 * 4), oState = oAsg + align256((size_t)n1 * 4);
    const size_t oBlk = oState + align256((size_t)n1 * 4), oLast = oBlk + align256((size_t)n2 * 4 + 4), total = oLast + align256((size_t)n2 * 4 + 4);
    PLF_CUDA_OK(scratch_reserve(c, total));
    uint8_t* b = c->d_scr;
    PLF_CUDA_OK(cudaMemcpyAsync(b + oD1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    PLF_CUDA_OK(cudaMemcpyAsync(b + oL1, lines1, (size_t)n1 * sizeof(plf_track_line), cudaMemcpyHostToDevice, s));
    if (n2) {
        PLF_CUDA_OK(cudaMemcpyAsync(b + oD2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
        PLF_CUDA_OK(cudaMemcpyAsync(b + oK2, kl2, (size_t)n2 * sizeof(plf_keyline), cudaMemcpyHostToDevice, s));
        PLF_CUDA_OK(cudaMemcpyAsync(b + oDisp, disp2, (size_t)n2 * 8, cudaMemcpyHostToDevice, s));
        if (held2) PLF_CUDA_OK(cudaMemcpyAsync(b + oHeld, held2, (size_t)n2, cudaMemcpyHostToDevice, s));
    }
    int* dM12 = reinterpret_cast<int*>(b + oM12);
    int* dM21 = reinterpret_cast<int*>(b + oM21);
    int* dAsg = reinterpret_cast<int*>(b + oAsg);
    // mode 0: match() = both directions + mutual best; mode 1: the MapLine overload returns after the one-way matchNNR
    c->launches = plf_launch_match_nnr(c, b + oD1, n1, b + oD2, n2, nnr, dM12);
    if (mode == 0) c->launches += plf_launch_match_nnr(c, b + oD2, n2, b + oD1, n1, nnr, dM21);
    c->launches += plf_launch_line_gates(c, mode, reinterpret_cast<const plf_track_line*>(b + oL1), n1,
                                         reinterpret_cast<const plf_keyline*>(b + oK2), reinterpret_cast<const float2*>(b + oDisp),
                                         (mode == 1 && held2) ? b + oHeld : nullptr, n2, min_x, max_x, min_y, max_y, dM12, dM21, dAsg,
                                         reinterpret_cast<int*>(b + oState), reinterpret_cast<int*>(b + oBlk), reinterpret_cast<int*>(b + oLast));
    PLF_CUDA_OK(cudaMemcpyAsync(matches12, dM12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaMemcpyAsync(assign12, dAsg, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    PLF_CUDA_OK(cudaGetLastError());
    int cnt = 0;
    for (int i = 0; i < n1; ++i) cnt += assign12[i] >= 0;
    if (n_assigned) *n_assigned = cnt;
    return PLF_OK;
}

// ---- bag-of-words transform -------------------------------------------------------------------------
PLF_API int plf_bow_set_vocabulary(plf_ctx* c, int which, int n_nodes, int levels, const int32_t* child_first,
                                   const int32_t* child_count, const int32_t* child, const uint8_t* desc, const int32_t* word_id,
                                   const double* weight) {
    if (!c || which < 0 || which > 1 || n_nodes < 2 || levels < 1 || !child_first || !child_count || !child || !desc || !word_id || !weight)
        return fail(PLF_ERR_INVALID, "bad vocabulary");
    // structural check on the host: children in range, every node but the root referenced once, depth <= levels
    long long total = 0;
    for (int i = 0; i < n_nodes; ++i) {
        if (child_count[i] < 0 || child_first[i] < 0) return fail(PLF_ERR_INVALID, "bad vocabulary (child range)");
        total += child_count[i];
    }
    if (total != n_nodes - 1 || child_count[0] == 0) return fail(PLF_ERR_INVALID, "bad vocabulary (not a tree rooted at node 0)");
    for (int i = 0; i < n_nodes; ++i)
        for (int k = 0; k < child_count[i]; ++k) {
            if ((long long)child_first[i] + k >= total) return fail(PLF_ERR_INVALID, "bad vocabulary (child index)");
            const int id = child[child_first[i] + k];
            if (id <= 0 || id >= n_nodes) return fail(PLF_ERR_INVALID, "bad vocabulary (child id)");
        }
    {   // depth of every root-to-leaf path (also rejects cycles: a walk longer than `levels` fails)
        std::vector<int> depth(n_nodes, -1), stack;
        depth[0] = 0; stack.push_back(0);
        while (!stack.empty()) {
            const int i = stack.back(); stack.pop_back();
            for (int k = 0; k < child_count[i]; ++k) {
                const int id = child[child_first[i] + k];
                if (depth[id] >= 0 || depth[i] + 1 > levels) return fail(PLF_ERR_INVALID, "bad vocabulary (cycle or deeper than levels)");
                depth[id] = depth[i] + 1;
                stack.push_back(id);
            }
        }
    }
    PLF_CUDA_OK(plf_enter(c));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PlfVocab& v = c->voc[which];
    void* old[] = {v.childFirst, v.childCount, v.child, v.word, v.desc, v.weight};
    for (void* q : old) if (q) cudaFree(q);
    v = PlfVocab();
    PLF_CUDA_OK(dalloc(&v.childFirst, (size_t)n_nodes));
    PLF_CUDA_OK(dalloc(&v.childCount, (size_t)n_nodes));
    PLF_CUDA_OK(dalloc(&v.child, (size_t)n_nodes));
    PLF_CUDA_OK(dalloc(&v.word, (size_t)n_nodes));
    PLF_CUDA_OK(dalloc(&v.desc, (size_t)n_nodes * 32));
    PLF_CUDA_OK(dalloc(&v.weight, (size_t)n_nodes));
    PLF_CUDA_OK(cudaMemcpy(v.childFirst, child_first, (size_t)n_nodes * 4, cudaMemcpyHostToDevice));
    PLF_CUDA_OK(cudaMemcpy(v.childCount, child_count, (size_t)n_nodes * 4, cudaMemcpyHostToDevice));
    PLF_CUDA_OK(cudaMemcpy(v.child, child, (size_t)(n_nodes - 1) * 4, cudaMemcpyHostToDevice));
    PLF_CUDA_OK(cudaMemcpy(v.word, word_id, (size_t)n_nodes * 4, cudaMemcpyHostToDevice));
    PLF_CUDA_OK(cudaMemcpy(v.desc, desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice));
    PLF_CUDA_OK(cudaMemcpy(v.weight, weight, (size_t)n_nodes * 8, cudaMemcpyHostToDevice));
    v.nNodes = n_nodes; v.levels = levels;
    return PLF_OK;
}

PLF_API int plf_bow_transform(plf_ctx* c, int which, int first_slot, int n_slots, int levelsup, int32_t* word_id, double* weight,
                              int32_t* node_id, int stride) {
    if (!c || which < 0 || which > 1 || !word_id || !weight || !node_id || first_slot < 0 || n_slots < 1 ||
        first_slot + n_slots > c->p.max_batch || stride < 1 || stride > (which ? c->g.klCap : c->g.kpCap))
        return fail(PLF_ERR_INVALID, "bad slot range / stride beyond the descriptor capacity");
    if (!c->voc[which].nNodes) return fail(PLF_ERR_STATE, "bow_transform before bow_set_vocabulary");
    if (which ? !(c->p.has_lines && c->lineValid[0]) : !c->orbValid[0]) return fail(PLF_ERR_STATE, "bow_transform before the descriptors exist");
    PLF_CUDA_OK(plf_enter(c));
    const size_t cap = (size_t)std::max(c->g.kpCap, c->g.klCap);
    if (!c->d_bowWord) {
        PLF_CUDA_OK(dalloc(&c->d_bowWord, (size_t)c->p.max_batch * cap));
        PLF_CUDA_OK(dalloc(&c->d_bowNode, (size_t)c->p.max_batch * cap));
        PLF_CUDA_OK(dalloc(&c->d_bowWeight, (size_t)c->p.max_batch * cap));
    }
    plf_launch_bow(c, which, first_slot, n_slots, levelsup, c->d_bowWord, c->d_bowWeight, c->d_bowNode, stride);
    cudaStream_t s = c->stream;
    PLF_CUDA_OK(cudaMemcpyAsync(word_id, c->d_bowWord, (size_t)n_slots * stride * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaMemcpyAsync(weight, c->d_bowWeight, (size_t)n_slots * stride * 8, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaMemcpyAsync(node_id, c->d_bowNode, (size_t)n_slots * stride * 4, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    PLF_CUDA_OK(cudaGetLastError());
    return PLF_OK;
}

PLF_API int plf_bow_build_vectors(const int32_t* word_id, const double* weight, const int32_t* node_id, int n, int32_t* bow_word,
                                  double* bow_value, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat, int* n_nodes_out) {
    return plf_bow_build(word_id, weight, node_id, n, bow_word, bow_value, fv_node, fv_start, fv_feat, n_nodes_out);
}

// ---- landmark back-projection -------------------------------------------------------------------------
PLF_API int plf_backproject(plf_ctx* c, int first_slot, int n_slots, const float* Rwc, const float* Ow, float fy, float cx,
                            float cy, float* x3d, int x3d_rows, double* l3d, int l3d_rows) {
    if (!c || !Rwc || !Ow || first_slot < 0 || n_slots < 1 || first_slot + n_slots > c->p.max_batch || (!x3d && !l3d) ||
        (x3d && (x3d_rows < 1 || x3d_rows > c->g.kpCap)) || (l3d && (l3d_rows < 1 || l3d_rows > c->g.klCap)) || !(fy > 0))
        return fail(PLF_ERR_INVALID, "bad slot range / rows beyond the keypoint or keyline capacity");
    if (!c->orbValid[0] || !c->orbValid[1]) return fail(PLF_ERR_STATE, "backproject before the stereo matches exist");
    if (l3d && !c->p.has_lines) return fail(PLF_ERR_STATE, "line back-projection on a context without lines");
    PLF_CUDA_OK(plf_enter(c));
    if (!c->d_bpPose) {
        PLF_CUDA_OK(dalloc(&c->d_bpPose, (size_t)c->p.max_batch * 12));
        PLF_CUDA_OK(dalloc(&c->d_bpX, (size_t)c->p.max_batch * c->g.kpCap * 3));
        PLF_CUDA_OK(dalloc(&c->d_bpL, (size_t)c->p.max_batch * c->g.klCap * 6));
    }
    cudaStream_t s = c->stream;
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_bpPose, Rwc, (size_t)n_slots * 36, cudaMemcpyHostToDevice, s));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_bpPose + (size_t)c->p.max_batch * 9, Ow, (size_t)n_slots * 12, cudaMemcpyHostToDevice, s));
    plf_launch_backproject(c, first_slot, n_slots, c->d_bpPose, c->d_bpPose + (size_t)c->p.max_batch * 9, fy, cx, cy,
                           x3d ? c->d_bpX : nullptr, x3d_rows, l3d ? c->d_bpL : nullptr, l3d_rows);
    if (x3d) PLF_CUDA_OK(cudaMemcpyAsync(x3d, c->d_bpX, (size_t)n_slots * x3d_rows * 12, cudaMemcpyDeviceToHost, s));
    if (l3d) PLF_CUDA_OK(cudaMemcpyAsync(l3d, c->d_bpL, (size_t)n_slots * l3d_rows * 48, cudaMemcpyDeviceToHost, s));
    PLF_CUDA_OK(cudaStreamSynchronize(s));
    PLF_CUDA_OK(cudaGetLastError());
    return PLF_OK;
}

// ---- rectification: cv::remap in front of the path ------------------------------------------------
static cudaError_t stage_reserve(plf_ctx* c, size_t bytes) {
    if (c->stageCap >= bytes) return cudaSuccess;
    if (c->d_stage) { cudaStreamSynchronize(c->stream); cudaFree(c->d_stage); }
    c->d_stage = nullptr;
    c->stageCap = 0;
    cudaError_t e = cudaMalloc((void**)&c->d_stage, bytes);
    if (e == cudaSuccess) c->stageCap = bytes;
    return e;
}

PLF_API int plf_rectify_set_maps(plf_ctx* c, int side, const float* mx, const float* my, int src_w, int src_h) {
    if (!c || side < 0 || side > 1 || !mx || !my || src_w < 2 || src_h < 2 || src_w > 32767 || src_h > 32767)
        return fail(PLF_ERR_INVALID, "bad rectification maps");
    PLF_CUDA_OK(plf_enter(c));
    const size_t n = (size_t)c->g.W * c->g.H;
    // cv::remap's own conversion of the float maps: cvRound(map * 32) (round half to even), integer part saturated
    // to int16, 5-bit fractions; done once here instead of once per frame
    std::vector<uint2> t(n);
    auto sat16 = [](int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); };
    for (size_t i = 0; i < n; ++i) {
        const int fxs = (int)std::nearbyintf(mx[i] * 32.f), fys = (int)std::nearbyintf(my[i] * 32.f);
        const int sx = sat16(fxs >> 5), sy = sat16(fys >> 5);
        t[i].x = (unsigned)(sx & 0xFFFF) | ((unsigned)(sy & 0xFFFF) << 16);
        t[i].y = (unsigned)(((fys & 31) << 5) | (fxs & 31));
    }
    if (!c->d_rmap[side]) PLF_CUDA_OK(dalloc(&c->d_rmap[side], n));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(c->d_rmap[side], t.data(), n * sizeof(uint2), cudaMemcpyHostToDevice));
    c->srcW[side] = src_w; c->srcH[side] = src_h;
    return PLF_OK;
}

PLF_API int plf_rectify(plf_ctx* c, int side, const uint8_t* raw, int raw_stride, uint8_t* out, int out_stride) {
    if (!c || side < 0 || side > 1 || !raw || !out) return fail(PLF_ERR_INVALID, "bad arguments");
    if (!c->d_rmap[side]) return fail(PLF_ERR_STATE, "rectify before rectify_set_maps");
    if (raw_stride < c->srcW[side] || out_stride < c->g.W) return fail(PLF_ERR_INVALID, "bad stride");
    PLF_CUDA_OK(plf_enter(c));
    const size_t bytes = (size_t)c->srcH[side] * raw_stride;
    PLF_CUDA_OK(stage_reserve(c, bytes));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_stage, raw, bytes, cudaMemcpyHostToDevice, c->stream));
    plf_launch_rectify(c, c->d_stage, c->d_stage, raw_stride, side, 1);      // slot 0, image index = side
    c->orbValid[side] = c->lineValid[side] = false;                         // level 0 of slot 0 was overwritten
    const PlfLevel& l0 = c->g.lv[0];
    PLF_CUDA_OK(cudaMemcpy2DAsync(out, out_stride, c->d_pyr + (size_t)side * c->g.pyrBytes + l0.off, l0.pitch, c->g.W, c->g.H,
                                  cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaGetLastError());
    return PLF_OK;
}

PLF_API int plf_batch_upload_raw(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int raw_stride) {
    if (!c || !left || !right || batch < 1 || batch > c->p.max_batch) return fail(PLF_ERR_INVALID, "bad batch");
    if (!c->d_rmap[0] || !c->d_rmap[1]) return fail(PLF_ERR_STATE, "batch_upload_raw before rectify_set_maps");
    if (raw_stride < c->srcW[0] || raw_stride < c->srcW[1]) return fail(PLF_ERR_INVALID, "bad stride");
    PLF_CUDA_OK(plf_enter(c));
    c->nMarks = 0;
    plf_mark(c, "h2d");
    const size_t b0 = (size_t)batch * c->srcH[0] * raw_stride, b1 = (size_t)batch * c->srcH[1] * raw_stride;
    const size_t off1 = (b0 + 255) & ~(size_t)255;
    PLF_CUDA_OK(stage_reserve(c, (((size_t)c->p.max_batch * c->srcH[0] * raw_stride + 255) & ~(size_t)255) +
                                     (size_t)c->p.max_batch * c->srcH[1] * raw_stride));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_stage, left, b0, cudaMemcpyHostToDevice, c->stream));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_stage + off1, right, b1, cudaMemcpyHostToDevice, c->stream));
    plf_mark(c, "rectify");
    plf_launch_rectify(c, c->d_stage, c->d_stage + off1, raw_stride, 0, 2 * batch);
    c->batchResident = batch;
    return PLF_OK;
}

}  // extern "C"
