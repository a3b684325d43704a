// TEST INFRASTRUCTURE ONLY — CPU oracle exported through the same C ABI as the product (prefix plf_cpu_).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
#define PLF_ORACLE_BUILD 1
#include "../../include/plf_b200.h"
#include "linematch.h"
#include "lsd.h"
#include "orb.h"
#include "stereo.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <thread>

using namespace plfo;

struct Slot {
    OrbState orb[2];
    LsdState lsd[2];
    Img8 img[2];
    std::vector<float> uRight, depth, disp;
    std::vector<double> le;
    std::vector<int> m12;
};

struct plf_ctx {
    plf_params p;
    OrbConfig oc;
    OrbTables ot;
    LsdConfig lc;
    LineMatchConfig mc;
    std::vector<Slot> slots;
    int kp_cap = 0, kl_cap = 0;
    int threads = 0;         // worker threads for batch calls (0 = hardware concurrency)
    int batch_resident = 0;
    struct Vocab { int levels = 0; std::vector<int> first, count, child, word; std::vector<uint8_t> desc; std::vector<double> weight; } voc[2];
    std::vector<float> mapx[2], mapy[2];   // rectification maps per camera (plf_cpu_rectify_set_maps)
    int srcW[2] = {0, 0}, srcH[2] = {0, 0};
};

static thread_local std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }

extern "C" {

PLF_API const char* plf_cpu_last_error(void) { return g_err.c_str(); }

PLF_API int plf_cpu_default_params(plf_params* p) {
    if (!p) return PLF_ERR_INVALID;
    std::memset(p, 0, sizeof(*p));
    p->width = 752; p->height = 480; p->max_batch = 1;
    p->n_features = 1200; p->scale_factor = 1.2f; p->n_levels = 8; p->ini_th_fast = 20; p->min_th_fast = 7;
    p->has_points = 1;
    p->has_lines = 1; p->lsd_nfeatures = 500; p->lsd_refine = 0; p->lsd_n_bins = 1024;
    p->min_line_length = 0.025; p->lsd_scale = 1.2; p->lsd_sigma_scale = 0.6; p->lsd_quant = 2.0;
    p->lsd_ang_th = 22.5; p->lsd_log_eps = 1.0; p->lsd_density_th = 0.6;
    p->bf = 47.90639384423901f; p->fx = 435.2046959714599f;
    p->best_lr_matches = 1; p->matching_s_ws = 10; p->min_ratio_12_l = 0.9; p->line_sim_th = 0.75;
    p->min_disp = 1.0; p->line_horiz_th = 0.1; p->stereo_overlap_th = 0.75; p->ls_min_disp_ratio = 0.7;
    return PLF_OK;
}

PLF_API int plf_cpu_create(const plf_params* p, int /*device*/, plf_ctx** out) {
    if (!p || !out) return fail(PLF_ERR_INVALID, "null argument");
    if (p->width < 64 || p->height < 64 || p->max_batch < 1 || p->n_levels < 1 || p->n_levels > 16)
        return fail(PLF_ERR_INVALID, "bad image size / batch / levels");
    if (p->lsd_refine < 0 || p->lsd_refine > 1) return fail(PLF_ERR_UNSUPPORTED, "lsd_refine = 2 (ADVANCED: NFA rectangle improvement) is not built");
    plf_ctx* c = new plf_ctx();
    c->p = *p;
    c->oc.nfeatures = p->n_features; c->oc.scaleFactor = p->scale_factor; c->oc.nlevels = p->n_levels;
    c->oc.iniThFAST = p->ini_th_fast; c->oc.minThFAST = p->min_th_fast;
    orb_tables(c->oc, c->ot);
    c->lc.refine = p->lsd_refine; c->lc.scale = p->lsd_scale; c->lc.sigma_scale = p->lsd_sigma_scale;
    c->lc.quant = p->lsd_quant; c->lc.ang_th = p->lsd_ang_th; c->lc.log_eps = p->lsd_log_eps;
    c->lc.density_th = p->lsd_density_th; c->lc.n_bins = p->lsd_n_bins;
    c->mc.best_lr_matches = p->best_lr_matches; c->mc.matching_s_ws = p->matching_s_ws;
    c->mc.min_ratio_12_l = p->min_ratio_12_l; c->mc.line_sim_th = p->line_sim_th; c->mc.min_disp = p->min_disp;
    c->mc.line_horiz_th = p->line_horiz_th; c->mc.stereo_overlap_th = p->stereo_overlap_th;
    c->mc.ls_min_disp_ratio = p->ls_min_disp_ratio;
    c->slots.resize(p->max_batch);
    c->kp_cap = ((p->n_features + 3 * p->n_levels + 31) / 32) * 32;
    c->kl_cap = p->lsd_nfeatures > 0 ? p->lsd_nfeatures : 4096;
    *out = c;
    return PLF_OK;
}

PLF_API int plf_cpu_destroy(plf_ctx* c) { delete c; return PLF_OK; }
PLF_API int plf_cpu_keypoint_capacity(const plf_ctx* c) { return c ? c->kp_cap : 0; }
PLF_API int plf_cpu_keyline_capacity(const plf_ctx* c) { return c ? c->kl_cap : 0; }

PLF_API int plf_cpu_get_scale_tables(const plf_ctx* c, float* s, float* is, float* s2, float* is2, int32_t* n) {
    if (!c) return PLF_ERR_INVALID;
    const int L = c->oc.nlevels;
    if (s) std::memcpy(s, c->ot.scale.data(), L * 4);
    if (is) std::memcpy(is, c->ot.invScale.data(), L * 4);
    if (s2) std::memcpy(s2, c->ot.sigma2.data(), L * 4);
    if (is2) std::memcpy(is2, c->ot.invSigma2.data(), L * 4);
    if (n) std::memcpy(n, c->ot.nPerLevel.data(), L * 4);
    return PLF_OK;
}

// oracle-only knobs
PLF_API int plf_cpu_set_threads(plf_ctx* c, int n) { if (!c) return PLF_ERR_INVALID; c->threads = n; return PLF_OK; }
PLF_API int plf_cpu_set_lsd_stable_order(plf_ctx* c, int on) { if (!c) return PLF_ERR_INVALID; c->lc.stable_order = on != 0; return PLF_OK; }

static int orb_slot(plf_ctx* c, int slot, int side, const uint8_t* img, int w, int h, int stride, int lap0, int lap1) {
    return orb_extract(c->oc, c->ot, img, w, h, stride, lap0, lap1, c->slots[slot].orb[side]);
}

static void line_slot(plf_ctx* c, int slot, int side, const uint8_t* img, int w, int h, int stride) {
    Slot& s = c->slots[slot];
    LsdState& st = s.lsd[side];
    st.kls.clear(); st.desc.clear(); st.lbd.clear(); st.segs.clear();
    if (!c->p.has_lines) { st.valid = true; return; }
    Img8 im(w, h);
    for (int y = 0; y < h; ++y) std::memcpy(im.row(y), img + (size_t)y * stride, w);
    lsd_detect(c->lc, im, st);
    const double min_len = c->p.min_line_length * std::min(w, h);
    lines_to_keylines(st.segs, w, h, min_len, c->p.lsd_nfeatures, st.kls);
    lbd_compute(im, st.kls, st.lbd, st.desc);
}

PLF_API int plf_cpu_orb_extract(plf_ctx* c, int side, const uint8_t* img, int w, int h, int stride, int lap0,
                                int lap1, plf_keypoint* out_kp, uint8_t* out_desc, int cap, int* n, int* mono) {
    if (!c || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad ctx/side");
    if (!img || w <= 0 || h <= 0) return PLF_ERR_EMPTY_IMAGE;
    if (w != c->p.width || h != c->p.height) return fail(PLF_ERR_INVALID, "image size differs from context");
    int m = orb_slot(c, 0, side, img, w, h, stride, lap0, lap1);
    const OrbState& st = c->slots[0].orb[side];
    if ((int)st.kps.size() > cap) return fail(PLF_ERR_INVALID, "keypoint capacity too small");
    if (out_kp) std::memcpy(out_kp, st.kps.data(), st.kps.size() * sizeof(plf_keypoint));
    if (out_desc) std::memcpy(out_desc, st.desc.data(), st.desc.size());
    if (n) *n = (int)st.kps.size();
    if (mono) *mono = m;
    return PLF_OK;
}

static int copy_level(const Img8& im, uint8_t* out, int out_stride, int* w, int* h) {
    if (w) *w = im.w;
    if (h) *h = im.h;
    if (out)
        for (int y = 0; y < im.h; ++y) std::memcpy(out + (size_t)y * out_stride, im.row(y), im.w);
    return PLF_OK;
}

PLF_API int plf_cpu_tap_pyramid_level(plf_ctx* c, int slot, int side, int level, uint8_t* out, int out_stride, int* w, int* h) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1 || level < 0 || level >= c->oc.nlevels) return PLF_ERR_INVALID;
    if (!c->slots[slot].orb[side].valid) return PLF_ERR_STATE;
    return copy_level(c->slots[slot].orb[side].pyr[level], out, out_stride, w, h);
}
PLF_API int plf_cpu_get_pyramid_level(plf_ctx* c, int side, int level, uint8_t* out, int out_stride, int* w, int* h) {
    return plf_cpu_tap_pyramid_level(c, 0, side, level, out, out_stride, w, h);
}
PLF_API int plf_cpu_tap_blurred_level(plf_ctx* c, int slot, int side, int level, uint8_t* out, int out_stride) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1 || level < 0 || level >= c->oc.nlevels) return PLF_ERR_INVALID;
    OrbState& st = c->slots[slot].orb[side];
    if (!st.valid) return PLF_ERR_STATE;
    if (st.blur[level].w == 0) gaussian_blur_u8(st.pyr[level], st.blur[level], TAPS_ORB7, 7);
    return copy_level(st.blur[level], out, out_stride, nullptr, nullptr);
}
PLF_API int plf_cpu_tap_fast_candidates(plf_ctx* c, int slot, int side, int level, float* xyr, int cap, int* n) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1 || level < 0 || level >= c->oc.nlevels) return PLF_ERR_INVALID;
    const OrbState& st = c->slots[slot].orb[side];
    if (!st.valid) return PLF_ERR_STATE;
    const auto& v = st.cands[level];
    if (n) *n = (int)v.size();
    if ((int)v.size() > cap) return PLF_ERR_INVALID;
    for (size_t i = 0; i < v.size(); ++i) { xyr[3 * i] = v[i].x + 16; xyr[3 * i + 1] = v[i].y + 16; xyr[3 * i + 2] = v[i].resp; }
    return PLF_OK;
}

PLF_API int plf_cpu_line_extract(plf_ctx* c, int side, const uint8_t* img, int w, int h, int stride,
                                 plf_keyline* out_kl, uint8_t* out_desc, int cap, int* n) {
    if (!c || side < 0 || side > 1 || !img) return fail(PLF_ERR_INVALID, "bad ctx/side/img");
    if (w != c->p.width || h != c->p.height) return fail(PLF_ERR_INVALID, "image size differs from context");
    line_slot(c, 0, side, img, w, h, stride);
    const LsdState& st = c->slots[0].lsd[side];
    if ((int)st.kls.size() > cap) return fail(PLF_ERR_INVALID, "keyline capacity too small");
    if (out_kl) std::memcpy(out_kl, st.kls.data(), st.kls.size() * sizeof(plf_keyline));
    if (out_desc) std::memcpy(out_desc, st.desc.data(), st.desc.size());
    if (n) *n = (int)st.kls.size();
    return PLF_OK;
}

PLF_API int plf_cpu_tap_lsd_scaled(plf_ctx* c, int slot, int side, uint8_t* out, int out_stride, int* w, int* h) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1) return PLF_ERR_INVALID;
    if (!c->slots[slot].lsd[side].valid) return PLF_ERR_STATE;
    return copy_level(c->slots[slot].lsd[side].scaled, out, out_stride, w, h);
}
PLF_API int plf_cpu_tap_lsd_angles(plf_ctx* c, int slot, int side, float* out, int* w, int* h) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1) return PLF_ERR_INVALID;
    const LsdState& st = c->slots[slot].lsd[side];
    if (!st.valid) return PLF_ERR_STATE;
    if (w) *w = st.scaled.w;
    if (h) *h = st.scaled.h;
    if (out) std::memcpy(out, st.angleDeg.data(), st.angleDeg.size() * 4);
    return PLF_OK;
}
PLF_API int plf_cpu_tap_lsd_segments(plf_ctx* c, int slot, int side, float* xyxy, int cap, int* n) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1) return PLF_ERR_INVALID;
    const LsdState& st = c->slots[slot].lsd[side];
    if (!st.valid) return PLF_ERR_STATE;
    int m = (int)st.segs.size() / 4;
    if (n) *n = m;
    if (m > cap) return PLF_ERR_INVALID;
    if (xyxy) std::memcpy(xyxy, st.segs.data(), st.segs.size() * 4);
    return PLF_OK;
}
PLF_API int plf_cpu_tap_grow_ns(plf_ctx*, unsigned long long* out, int n_images) {      // no such kernel on the CPU: zeros
    for (int i = 0; i < n_images; ++i) out[i] = 0;
    return PLF_OK;
}
PLF_API int plf_cpu_tap_lbd_float(plf_ctx* c, int slot, int side, float* out, int cap, int* n) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || side < 0 || side > 1) return PLF_ERR_INVALID;
    const LsdState& st = c->slots[slot].lsd[side];
    if (!st.valid) return PLF_ERR_STATE;
    int m = (int)st.kls.size();
    if (n) *n = m;
    if (m > cap) return PLF_ERR_INVALID;
    if (out) std::memcpy(out, st.lbd.data(), st.lbd.size() * 4);
    return PLF_OK;
}

static int match_points_slot(plf_ctx* c, int slot) {
    Slot& s = c->slots[slot];
    if (!s.orb[0].valid || !s.orb[1].valid) return PLF_ERR_STATE;
    stereo_match_points(c->ot, s.orb[0], s.orb[1], c->p.bf, c->p.fx, s.uRight, s.depth);
    return PLF_OK;
}
static int match_lines_slot(plf_ctx* c, int slot) {
    Slot& s = c->slots[slot];
    if (!s.lsd[0].valid || !s.lsd[1].valid) return PLF_ERR_STATE;
    stereo_match_lines(c->mc, c->p.width, c->p.height, s.lsd[0].kls, s.lsd[0].desc, s.lsd[1].kls, s.lsd[1].desc,
                       s.disp, s.le, s.m12);
    return PLF_OK;
}

PLF_API int plf_cpu_stereo_match_points(plf_ctx* c, float* u_right, float* depth, int cap) {
    if (!c) return PLF_ERR_INVALID;
    int rc = match_points_slot(c, 0);
    if (rc) return fail(rc, "stereo_match_points before both orb_extract calls");
    Slot& s = c->slots[0];
    if ((int)s.uRight.size() > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (u_right) std::memcpy(u_right, s.uRight.data(), s.uRight.size() * 4);
    if (depth) std::memcpy(depth, s.depth.data(), s.depth.size() * 4);
    return PLF_OK;
}

PLF_API int plf_cpu_stereo_match_lines(plf_ctx* c, float* disp_se, double* le, int32_t* match12, int cap) {
    if (!c) return PLF_ERR_INVALID;
    int rc = match_lines_slot(c, 0);
    if (rc) return fail(rc, "stereo_match_lines before both line_extract calls");
    Slot& s = c->slots[0];
    if ((int)s.m12.size() > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (disp_se) std::memcpy(disp_se, s.disp.data(), s.disp.size() * 4);
    if (le) std::memcpy(le, s.le.data(), s.le.size() * 8);
    if (match12) std::memcpy(match12, s.m12.data(), s.m12.size() * 4);
    return PLF_OK;
}

PLF_API int plf_cpu_match_nnr(plf_ctx*, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int32_t* m12, int* nm) {
    if (n1 < 0 || n2 < 0 || (n1 && !d1) || (n2 && !d2) || (n1 && !m12)) return fail(PLF_ERR_INVALID, "bad descriptors");
    int m = match_nnr(d1, n1, d2, n2, nnr, m12);
    if (nm) *nm = m;
    return PLF_OK;
}
PLF_API int plf_cpu_match(plf_ctx*, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int32_t* m12, int* nm) {
    if (n1 < 0 || n2 < 0 || (n1 && !d1) || (n2 && !d2) || (n1 && !m12)) return fail(PLF_ERR_INVALID, "bad descriptors");
    int m = match_lr(d1, n1, d2, n2, nnr, best_lr, m12);
    if (nm) *nm = m;
    return PLF_OK;
}

// ---- batch -----------------------------------------------------------------------------------------------
PLF_API int plf_cpu_batch_upload(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int stride) {
    if (!c || !left || !right || batch < 1 || batch > (int)c->slots.size()) return fail(PLF_ERR_INVALID, "bad batch");
    const int w = c->p.width, h = c->p.height;
    for (int b = 0; b < batch; ++b)
        for (int s = 0; s < 2; ++s) {
            Img8& im = c->slots[b].img[s];
            im = Img8(w, h);
            const uint8_t* src = (s ? right : left) + (size_t)b * h * stride;
            for (int y = 0; y < h; ++y) std::memcpy(im.row(y), src + (size_t)y * stride, w);
        }
    c->batch_resident = batch;
    return PLF_OK;
}

// ---- rectification --------------------------------------------------------------------
PLF_API int plf_cpu_rectify_set_maps(plf_ctx* c, int side, const float* mx, const float* my, int src_w, int src_h) {
    if (!c || side < 0 || side > 1 || !mx || !my || src_w < 2 || src_h < 2) return fail(PLF_ERR_INVALID, "bad rectification maps");
    const size_t n = (size_t)c->p.width * c->p.height;
    c->mapx[side].assign(mx, mx + n);
    c->mapy[side].assign(my, my + n);
    c->srcW[side] = src_w; c->srcH[side] = src_h;
    return PLF_OK;
}
static void rectify_one(plf_ctx* c, int side, const uint8_t* raw, int stride, Img8& dst) {
    Img8 src(c->srcW[side], c->srcH[side]);
    for (int y = 0; y < src.h; ++y) std::memcpy(src.row(y), raw + (size_t)y * stride, src.w);
    remap_linear_u8(src, dst, c->mapx[side].data(), c->mapy[side].data(), c->p.width, c->p.height);
}
PLF_API int plf_cpu_rectify(plf_ctx* c, int side, const uint8_t* raw, int raw_stride, uint8_t* out, int out_stride) {
    if (!c || side < 0 || side > 1 || !raw || !out) return fail(PLF_ERR_INVALID, "bad arguments");
    if (c->mapx[side].empty()) return fail(PLF_ERR_STATE, "rectify before rectify_set_maps");
    if (raw_stride < c->srcW[side] || out_stride < c->p.width) return fail(PLF_ERR_INVALID, "bad stride");
    Img8 dst;
    rectify_one(c, side, raw, raw_stride, dst);
    for (int y = 0; y < dst.h; ++y) std::memcpy(out + (size_t)y * out_stride, dst.row(y), dst.w);
    return PLF_OK;
}
PLF_API int plf_cpu_batch_upload_raw(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int stride) {
    if (!c || !left || !right || batch < 1 || batch > (int)c->slots.size()) return fail(PLF_ERR_INVALID, "bad batch");
    if (c->mapx[0].empty() || c->mapx[1].empty()) return fail(PLF_ERR_STATE, "batch_upload_raw before rectify_set_maps");
    for (int b = 0; b < batch; ++b)
        for (int s = 0; s < 2; ++s)
            rectify_one(c, s, (s ? right : left) + (size_t)b * c->srcH[s] * stride, stride, c->slots[b].img[s]);
    c->batch_resident = batch;
    return PLF_OK;
}

// ---- Frame::AssignFeaturesToGrid (src/Frame.cc:451-482) + PosInGrid (:845-855), left keypoints, no distortion --------
PLF_API int plf_cpu_feature_grid(plf_ctx* c, int first_slot, int n_slots, int32_t* cell_start, int32_t* cell_idx, int idx_stride) {
    if (!c || !cell_start || !cell_idx || first_slot < 0 || n_slots < 1 || first_slot + n_slots > (int)c->slots.size() ||
        idx_stride < c->kp_cap)
        return fail(PLF_ERR_INVALID, "bad slot range / idx_stride");
    constexpr int NC = PLF_GRID_COLS * PLF_GRID_ROWS;
    const float invW = (float)PLF_GRID_COLS / ((float)c->p.width - 0.0f), invH = (float)PLF_GRID_ROWS / ((float)c->p.height - 0.0f);
    for (int s = 0; s < n_slots; ++s) {
        const std::vector<plf_keypoint>& kps = c->slots[first_slot + s].orb[0].kps;
        std::vector<std::vector<int>> grid(NC);
        for (int i = 0; i < (int)kps.size(); ++i) {
            const int px = (int)std::round((kps[i].x - 0.0f) * invW), py = (int)std::round((kps[i].y - 0.0f) * invH);
            if (px < 0 || px >= PLF_GRID_COLS || py < 0 || py >= PLF_GRID_ROWS) continue;
            grid[px * PLF_GRID_ROWS + py].push_back(i);
        }
        int32_t* st = cell_start + (size_t)s * (NC + 1);
        int32_t* ix = cell_idx + (size_t)s * idx_stride;
        int n = 0;
        for (int cidx = 0; cidx < NC; ++cidx) {
            st[cidx] = n;
            for (int i : grid[cidx]) ix[n++] = i;
        }
        st[NC] = n;
    }
    return PLF_OK;
}

PLF_API int plf_cpu_get_features_in_area(const plf_keypoint* kps, const int32_t* cell_start, const int32_t* cell_idx, int width,
                                         int height, float x, float y, float r, int min_level, int max_level, int32_t* out, int cap) {
    return plf_features_in_area(kps, cell_start, cell_idx, width, height, x, y, r, min_level, max_level, out, cap);
}

// ---- Frame::UnprojectStereo (src/Frame.cc:1332-1347) and Frame::backProjection (:1349-1358) --------------------------
// cv::Mat arithmetic of mRwc*x3Dc+mOw = cv::gemm's 3x3 path (pinned against cv2.gemm: float products summed left to
// right, then (float)((double)t + (double)c)); the Eigen expression of backProjection is plain double arithmetic
// (parity unpinned for that half: no Eigen in this container; evaluated as written, left to right, no contraction).
PLF_API int plf_cpu_backproject(plf_ctx* c, int first_slot, int n_slots, const float* Rwc, const float* Ow, float fy, float cx,
                                float cy, float* x3d, int x3d_rows, double* l3d, int l3d_rows) {
    if (!c || !Rwc || !Ow || first_slot < 0 || n_slots < 1 || first_slot + n_slots > (int)c->slots.size() || (!x3d && !l3d))
        return fail(PLF_ERR_INVALID, "bad arguments");
    const float fx = c->p.fx, invfx = 1.0f / fx, invfy = 1.0f / fy, mb = c->p.bf / c->p.fx;
    for (int s = 0; s < n_slots; ++s) {
        const Slot& sl = c->slots[first_slot + s];
        const float* R = Rwc + s * 9;
        const float* O = Ow + s * 3;
        if (x3d)
            for (int i = 0; i < x3d_rows; ++i) {
                float* d = x3d + ((size_t)s * x3d_rows + i) * 3;
                d[0] = d[1] = d[2] = 0.f;
                if (i >= (int)sl.orb[0].kps.size() || i >= (int)sl.depth.size()) continue;
                const float z = sl.depth[i];
                if (!(z > 0)) continue;
                const float x = (sl.orb[0].kps[i].x - cx) * z * invfx, y = (sl.orb[0].kps[i].y - cy) * z * invfy;
                for (int r = 0; r < 3; ++r) {
                    const float t = R[3 * r] * x + R[3 * r + 1] * y + R[3 * r + 2] * z;
                    d[r] = (float)((double)t + (double)O[r]);
                }
            }
        if (l3d)
            for (int i = 0; i < l3d_rows; ++i) {
                double* d = l3d + ((size_t)s * l3d_rows + i) * 6;
                for (int k = 0; k < 6; ++k) d[k] = 0;
                if (i >= (int)sl.lsd[0].kls.size() || 2 * i + 1 >= (int)sl.disp.size()) continue;
                const float dd[2] = {sl.disp[2 * i], sl.disp[2 * i + 1]};
                if (!(dd[0] > 0 && dd[1] > 0)) continue;
                const plf_keyline& k = sl.lsd[0].kls[i];
                const float uv[4] = {k.startPointX, k.startPointY, k.endPointX, k.endPointY};
                for (int e = 0; e < 2; ++e) {
                    const double bd = (double)mb / (double)dd[e];
                    const double P[3] = {bd * ((double)uv[2 * e] - (double)cx), bd * ((double)uv[2 * e + 1] - (double)cy), bd * (double)fx};
                    for (int r = 0; r < 3; ++r)
                        d[3 * e + r] = (((double)R[3 * r] * P[0] + (double)R[3 * r + 1] * P[1]) + (double)R[3 * r + 2] * P[2]) + (double)O[r];
                }
            }
    }
    return PLF_OK;
}

// ---- DBoW2 transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1270), restated on flat node arrays ----------
// parity unpinned: DBoW2 itself cannot be built here (needs OpenCV C++ headers) and the vocabulary files are not in the
// reference tree; the descent and the vector construction follow the published algorithm and are cross-checked against
// a direct Python restatement in tests/.
PLF_API int plf_cpu_bow_set_vocabulary(plf_ctx* c, int which, int n_nodes, int levels, const int32_t* child_first,
                                       const int32_t* child_count, const int32_t* child, const uint8_t* desc,
                                       const int32_t* word_id, const double* weight) {
    if (!c || which < 0 || which > 1 || n_nodes < 2 || levels < 1 || !child_first || !child_count || !child || !desc || !word_id || !weight)
        return fail(PLF_ERR_INVALID, "bad vocabulary");
    auto& v = c->voc[which];
    v.levels = levels;
    v.first.assign(child_first, child_first + n_nodes);
    v.count.assign(child_count, child_count + n_nodes);
    v.child.assign(child, child + (n_nodes - 1));
    v.word.assign(word_id, word_id + n_nodes);
    v.desc.assign(desc, desc + (size_t)n_nodes * 32);
    v.weight.assign(weight, weight + n_nodes);
    return PLF_OK;
}
PLF_API int plf_cpu_bow_transform(plf_ctx* c, int which, int first_slot, int n_slots, int levelsup, int32_t* word_id, double* weight,
                                  int32_t* node_id, int stride) {
    if (!c || which < 0 || which > 1 || !word_id || !weight || !node_id || first_slot < 0 || n_slots < 1 ||
        first_slot + n_slots > (int)c->slots.size() || stride < 1)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const auto& v = c->voc[which];
    if (v.first.empty()) return fail(PLF_ERR_STATE, "bow_transform before bow_set_vocabulary");
    for (int s = 0; s < n_slots; ++s) {
        const Slot& sl = c->slots[first_slot + s];
        const uint8_t* D = which ? sl.lsd[0].desc.data() : sl.orb[0].desc.data();
        const int n = which ? (int)sl.lsd[0].kls.size() : (int)sl.orb[0].kps.size();
        for (int i = 0; i < stride; ++i) {
            const size_t o = (size_t)s * stride + i;
            if (i >= n) { word_id[o] = -1; weight[o] = 0.0; node_id[o] = 0; continue; }
            const uint8_t* f = D + (size_t)i * 32;
            const int nidLevel = v.levels - levelsup;
            int node = 0, level = 0, nid = 0;
            do {
                ++level;
                int best = 0x7fffffff, bestId = node;
                for (int k = 0; k < v.count[node]; ++k) {
                    const int id = v.child[v.first[node] + k];
                    const int d = plf_hamming256(f, v.desc.data() + (size_t)id * 32);
                    if (d < best) { best = d; bestId = id; }
                }
                node = bestId;
                if (level == nidLevel) nid = node;
            } while (v.count[node] > 0);
            word_id[o] = v.word[node]; weight[o] = v.weight[node]; node_id[o] = nid;
        }
    }
    return PLF_OK;
}
PLF_API int plf_cpu_bow_build_vectors(const int32_t* word_id, const double* weight, const int32_t* node_id, int n, int32_t* bow_word,
                                      double* bow_value, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat, int* n_nodes_out) {
    return plf_bow_build(word_id, weight, node_id, n, bow_word, bow_value, fv_node, fv_start, fv_feat, n_nodes_out);
}

// ---- ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, ...) (src/ORBmatcher.cc:44-130), Nleft == -1 ----
// pinned: tests/test_oracle_ref.py runs the reference's own function (oracle/_ref, src/ORBmatcher.cc:44-214 on a grid filled by
// its AssignFeaturesToGrid) on the same map points and demands equality; the loops follow the reference line by line.
PLF_API int plf_cpu_search_by_projection(plf_ctx* c, int slot, const plf_proj_query* queries, int n_queries, float th, float nn_ratio,
                                         int th_high, uint8_t* occupied, int n_features, int32_t* match, int* n_matches) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !queries || n_queries < 0 || !occupied || !match)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const Slot& sl = c->slots[slot];
    const std::vector<plf_keypoint>& kps = sl.orb[0].kps;
    if (n_features < (int)kps.size()) return fail(PLF_ERR_INVALID, "occupied[] is shorter than the slot's keypoint count");
    const uint8_t* D = sl.orb[0].desc.data();
    const float invW = (float)PLF_GRID_COLS / ((float)c->p.width - 0.0f), invH = (float)PLF_GRID_ROWS / ((float)c->p.height - 0.0f);
    std::vector<std::vector<int>> grid(PLF_GRID_COLS * PLF_GRID_ROWS);           // Frame::AssignFeaturesToGrid
    for (int i = 0; i < (int)kps.size(); ++i) {
        const int px = (int)std::round((kps[i].x - 0.0f) * invW), py = (int)std::round((kps[i].y - 0.0f) * invH);
        if (px < 0 || px >= PLF_GRID_COLS || py < 0 || py >= PLF_GRID_ROWS) continue;
        grid[px * PLF_GRID_ROWS + py].push_back(i);
    }
    const bool bFactor = th != 1.0f;
    int nmatches = 0;
    for (int iMP = 0; iMP < n_queries; ++iMP) {
        const plf_proj_query& q = queries[iMP];
        match[iMP] = -1;
        if (q.skip || q.level < 0 || q.level >= (int)c->ot.scale.size()) continue;
        float r = q.view_cos > 0.998f ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        const float rad = r * c->ot.scale[q.level];
        const int minLevel = q.level - 1, maxLevel = q.level;
        // Frame::GetFeaturesInArea(x, y, rad, minLevel, maxLevel)
        std::vector<int> vIndices;
        {
            const int x0 = std::max(0, (int)std::floor((q.proj_x - 0.0f - rad) * invW));
            const int x1 = std::min(PLF_GRID_COLS - 1, (int)std::ceil((q.proj_x - 0.0f + rad) * invW));
            const int y0 = std::max(0, (int)std::floor((q.proj_y - 0.0f - rad) * invH));
            const int y1 = std::min(PLF_GRID_ROWS - 1, (int)std::ceil((q.proj_y - 0.0f + rad) * invH));
            if (x0 < PLF_GRID_COLS && x1 >= 0 && y0 < PLF_GRID_ROWS && y1 >= 0) {
                const bool check = minLevel > 0 || maxLevel >= 0;
                for (int ix = x0; ix <= x1; ++ix)
                    for (int iy = y0; iy <= y1; ++iy)
                        for (int idx : grid[ix * PLF_GRID_ROWS + iy]) {
                            const plf_keypoint& k = kps[idx];
                            if (check) {
                                if (k.octave < minLevel) continue;
                                if (maxLevel >= 0 && k.octave > maxLevel) continue;
                            }
                            if (std::fabs(k.x - q.proj_x) < rad && std::fabs(k.y - q.proj_y) < rad) vIndices.push_back(idx);
                        }
            }
        }
        if (vIndices.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : vIndices) {
            if (occupied[idx]) continue;
            if (idx < (int)sl.uRight.size() && sl.uRight[idx] > 0) {
                const float er = std::fabs(q.proj_xr - sl.uRight[idx]);
                if (er > rad) continue;
            }
            const int dist = plf_hamming256(q.desc, D + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= th_high) {
            if (bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nn_ratio * bestDist2) {
                match[iMP] = bestIdx;
                occupied[bestIdx] = 1;
                ++nmatches;
            }
        }
    }
    if (n_matches) *n_matches = nmatches;
    return PLF_OK;
}

// ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono, match12) (src/ORBmatcher.cc:2179-2323)
// from the projected points on: window search, best distance, assignment, rotation histogram, ComputeThreeMaxima (:2449-2490).
// pinned like the local-map overload above (oracle/_ref runs src/ORBmatcher.cc:2179-2323 itself, all three motion cases).
PLF_API int plf_cpu_search_by_projection_frame(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int th_high,
                                               int check_orientation, uint8_t* occupied, int n_features, int32_t* feat_query,
                                               int32_t* match12, int* n_matches) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !queries || n_queries < 0 || !occupied || !feat_query)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const Slot& sl = c->slots[slot];
    const std::vector<plf_keypoint>& kps = sl.orb[0].kps;
    const uint8_t* D = sl.orb[0].desc.data();
    const int N = (int)kps.size();
    if (n_features < N) return fail(PLF_ERR_INVALID, "occupied[] / feat_query[] are shorter than the slot's keypoint count");
    const float invW = (float)PLF_GRID_COLS / ((float)c->p.width - 0.0f), invH = (float)PLF_GRID_ROWS / ((float)c->p.height - 0.0f);
    std::vector<std::vector<int>> grid(PLF_GRID_COLS * PLF_GRID_ROWS);
    for (int i = 0; i < N; ++i) {
        const int px = (int)std::round((kps[i].x - 0.0f) * invW), py = (int)std::round((kps[i].y - 0.0f) * invH);
        if (px < 0 || px >= PLF_GRID_COLS || py < 0 || py >= PLF_GRID_ROWS) continue;
        grid[px * PLF_GRID_ROWS + py].push_back(i);
    }
    for (int f = 0; f < N; ++f) { feat_query[f] = -1; if (match12) match12[f] = -1; }
    const int HISTO = 30;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO;
    int nmatches = 0;
    for (int i = 0; i < n_queries; ++i) {
        const plf_frame_query& q = queries[i];
        if (q.skip) continue;
        const float u = q.u, v = q.v, radius = q.radius;
        std::vector<int> vIndices2;                 // CurrentFrame.GetFeaturesInArea(u, v, radius, min_level, max_level)
        {
            const int x0 = std::max(0, (int)std::floor((u - 0.0f - radius) * invW));
            const int x1 = std::min(PLF_GRID_COLS - 1, (int)std::ceil((u - 0.0f + radius) * invW));
            const int y0 = std::max(0, (int)std::floor((v - 0.0f - radius) * invH));
            const int y1 = std::min(PLF_GRID_ROWS - 1, (int)std::ceil((v - 0.0f + radius) * invH));
            if (x0 < PLF_GRID_COLS && x1 >= 0 && y0 < PLF_GRID_ROWS && y1 >= 0) {
                const bool check = q.min_level > 0 || q.max_level >= 0;
                for (int ix = x0; ix <= x1; ++ix)
                    for (int iy = y0; iy <= y1; ++iy)
                        for (int idx : grid[ix * PLF_GRID_ROWS + iy]) {
                            const plf_keypoint& k = kps[idx];
                            if (check) {
                                if (k.octave < q.min_level) continue;
                                if (q.max_level >= 0 && k.octave > q.max_level) continue;
                            }
                            if (std::fabs(k.x - u) < radius && std::fabs(k.y - v) < radius) vIndices2.push_back(idx);
                        }
            }
        }
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (occupied[i2]) continue;
            if (i2 < (int)sl.uRight.size() && sl.uRight[i2] > 0) {
                const float er = std::fabs(q.ur - sl.uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = plf_hamming256(q.desc, D + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestIdx2 >= 0 && bestDist <= th_high) {
            feat_query[bestIdx2] = i;
            occupied[bestIdx2] = q.has_observations ? 1 : 0;
            ++nmatches;
            if (match12 && match12[bestIdx2] < 0) match12[bestIdx2] = i;
            if (check_orientation) {
                float rot = q.angle - kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO) bin = 0;
                if (bin >= 0 && bin < HISTO) rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1, max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); ++j) {
                    const int f = rotHist[i][j];
                    feat_query[f] = -1;
                    occupied[f] = 0;
                    --nmatches;
                    if (match12) match12[f] = -1;
                }
    }
    if (n_matches) *n_matches = nmatches;
    return PLF_OK;
}

// ---- the remaining descriptor searches of SURVEY §8(f) rank 1 and the line-match gates ---------------------------------
namespace {
// Frame::GetFeaturesInArea / KeyFrame::GetFeaturesInArea (src/Frame.cc:774-843, src/KeyFrame.cc:881-925) over mGrid of the slot
struct SlotGrid {
    std::vector<std::vector<int>> cell;
    float invW, invH;
    const std::vector<plf_keypoint>* kps;
    SlotGrid(const plf_ctx* c, const std::vector<plf_keypoint>& k) : cell(PLF_GRID_COLS * PLF_GRID_ROWS), kps(&k) {
        invW = (float)PLF_GRID_COLS / ((float)c->p.width - 0.0f);
        invH = (float)PLF_GRID_ROWS / ((float)c->p.height - 0.0f);
        for (int i = 0; i < (int)k.size(); ++i) {       // AssignFeaturesToGrid + PosInGrid (src/Frame.cc:451-482, :845-855)
            const int px = (int)std::round((k[i].x - 0.0f) * invW), py = (int)std::round((k[i].y - 0.0f) * invH);
            if (px < 0 || px >= PLF_GRID_COLS || py < 0 || py >= PLF_GRID_ROWS) continue;
            cell[px * PLF_GRID_ROWS + py].push_back(i);
        }
    }
    void area(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - 0.0f - r) * invW));
        if (nMinCellX >= PLF_GRID_COLS) return;
        const int nMaxCellX = std::min(PLF_GRID_COLS - 1, (int)std::ceil((x - 0.0f + r) * invW));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - 0.0f - r) * invH));
        if (nMinCellY >= PLF_GRID_ROWS) return;
        const int nMaxCellY = std::min(PLF_GRID_ROWS - 1, (int)std::ceil((y - 0.0f + r) * invH));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
            for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
                for (int idx : cell[ix * PLF_GRID_ROWS + iy]) {
                    const plf_keypoint& kp = (*kps)[idx];
                    if (bCheckLevels) {
                        if (kp.octave < minLevel) continue;
                        if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                    }
                    const float distx = kp.x - x, disty = kp.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(idx);
                }
    }
};

// ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:2449-2490)
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}
}  // namespace

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2325-2447,
// from the projected points on.  Pinned to the reference's own function in tests/test_oracle_ref.py (oracle/_ref).
PLF_API int plf_cpu_search_by_projection_reloc(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int orb_dist,
                                               int check_orientation, uint8_t* occupied, int n_features, int32_t* feat_query,
                                               int* n_matches) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !queries || n_queries < 0 || !occupied || !feat_query)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const Slot& sl = c->slots[slot];
    const std::vector<plf_keypoint>& kps = sl.orb[0].kps;
    const uint8_t* D = sl.orb[0].desc.data();
    const int N = (int)kps.size();
    if (n_features < N) return fail(PLF_ERR_INVALID, "occupied[] / feat_query[] are shorter than the slot's keypoint count");
    const SlotGrid grid(c, kps);
    for (int f = 0; f < N; ++f) feat_query[f] = -1;
    const int HISTO_LENGTH = 30;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    std::vector<int> vIndices2;
    for (int i = 0; i < n_queries; ++i) {
        const plf_frame_query& q = queries[i];
        if (q.skip) continue;
        grid.area(q.u, q.v, q.radius, q.min_level, q.max_level, vIndices2);
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (occupied[i2]) continue;                       // CurrentFrame.mvpMapPoints[i2]
            const int dist = plf_hamming256(q.desc, D + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestIdx2 >= 0 && bestDist <= orb_dist) {
            occupied[bestIdx2] = 1;
            feat_query[bestIdx2] = i;
            nmatches++;
            if (check_orientation) {
                float rot = q.angle - kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin >= 0 && bin < HISTO_LENGTH) rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    occupied[rotHist[i][j]] = 0;              // CurrentFrame.mvpMapPoints[...] = NULL
                    feat_query[rotHist[i][j]] = -1;
                    nmatches--;
                }
    }
    if (n_matches) *n_matches = nmatches;
    return PLF_OK;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th, ratioHamming), src/ORBmatcher.cc:473-586 (and
// :588-704), from the projected points on.  Pinned to the reference's own function (:473-586) in tests/test_oracle_ref.py.
PLF_API int plf_cpu_search_by_projection_loop(plf_ctx* c, int slot, const plf_frame_query* queries, int n_queries, int th_low,
                                              float ratio_hamming, uint8_t* occupied, int n_features, int32_t* feat_query,
                                              int* n_matches) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || !queries || n_queries < 0 || !occupied || !feat_query)
        return fail(PLF_ERR_INVALID, "bad arguments");
    const Slot& sl = c->slots[slot];
    const std::vector<plf_keypoint>& kps = sl.orb[0].kps;
    const uint8_t* D = sl.orb[0].desc.data();
    const int N = (int)kps.size();
    if (n_features < N) return fail(PLF_ERR_INVALID, "occupied[] / feat_query[] are shorter than the slot's keypoint count");
    const SlotGrid grid(c, kps);
    for (int f = 0; f < N; ++f) feat_query[f] = -1;
    int nmatches = 0;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n_queries; ++iMP) {
        const plf_frame_query& q = queries[iMP];
        if (q.skip) continue;
        grid.area(q.u, q.v, q.radius, -1, -1, vIndices);      // pKF->GetFeaturesInArea(u, v, radius): no level filter
        if (vIndices.empty()) continue;
        const int nPredictedLevel = q.max_level;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            if (occupied[idx]) continue;                      // vpMatched[idx]
            const int kpLevel = kps[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int dist = plf_hamming256(q.desc, D + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestIdx >= 0 && bestDist <= th_low * ratio_hamming) {
            occupied[bestIdx] = 1;
            feat_query[bestIdx] = iMP;
            nmatches++;
        }
    }
    if (n_matches) *n_matches = nmatches;
    return PLF_OK;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches), src/ORBmatcher.cc:269-471, F.Nleft == -1.
// The two FeatureVectors are std::map<NodeId, vector<unsigned>>: rebuilt here from the per-feature node ids.
PLF_API int plf_cpu_search_by_bow(plf_ctx* c, int slot, const uint8_t* kf_desc, const float* kf_angle, const int32_t* kf_node,
                                  const uint8_t* kf_valid, int n_kf, const int32_t* f_node, int n_features, int th_low,
                                  float nn_ratio, int check_orientation, int32_t* match, int* n_matches) {
    if (!c || slot < 0 || slot >= (int)c->slots.size() || n_kf < 0 || n_features < 0 ||
        (n_kf && (!kf_desc || !kf_angle || !kf_node || !kf_valid)) || (n_features && (!f_node || !match)))
        return fail(PLF_ERR_INVALID, "bad arguments");
    const Slot& sl = c->slots[slot];
    const std::vector<plf_keypoint>& kps = sl.orb[0].kps;
    const uint8_t* D = sl.orb[0].desc.data();
    const int N = (int)kps.size();
    if (n_features < N) return fail(PLF_ERR_INVALID, "f_node[] / match[] are shorter than the slot's keypoint count");
    std::map<int, std::vector<unsigned>> vFeatVecKF, vFeatVecF;
    for (int i = 0; i < n_kf; ++i) if (kf_node[i] >= 0) vFeatVecKF[kf_node[i]].push_back((unsigned)i);
    for (int i = 0; i < N; ++i) if (f_node[i] >= 0) vFeatVecF[f_node[i]].push_back((unsigned)i);
    for (int f = 0; f < n_features; ++f) match[f] = -1;       // vpMapPointMatches = vector<MapPoint*>(F.N, NULL)
    int nmatches = 0;
    const int HISTO_LENGTH = 30, TH_LOW = th_low;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    auto KFit = vFeatVecKF.begin(), KFend = vFeatVecKF.end();
    auto Fit = vFeatVecF.begin(), Fend = vFeatVecF.end();
    while (KFit != KFend && Fit != Fend) {
        if (KFit->first == Fit->first) {
            const std::vector<unsigned>& vIndicesKF = KFit->second;
            const std::vector<unsigned>& vIndicesF = Fit->second;
            for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
                const unsigned realIdxKF = vIndicesKF[iKF];
                if (!kf_valid[realIdxKF]) continue;           // no map point, or a bad one
                const uint8_t* dKF = kf_desc + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (size_t iF = 0; iF < vIndicesF.size(); iF++) {
                    const unsigned realIdxF = vIndicesF[iF];
                    if (match[realIdxF] >= 0) continue;
                    const int dist = plf_hamming256(dKF, D + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
                    else if (dist < bestDist2) { bestDist2 = dist; }
                }
                if (bestDist1 <= TH_LOW) {
                    if (static_cast<float>(bestDist1) < nn_ratio * static_cast<float>(bestDist2)) {
                        match[bestIdxF] = (int)realIdxKF;
                        if (check_orientation) {
                            float rot = kf_angle[realIdxKF] - kps[bestIdxF].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)std::round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            if (bin >= 0 && bin < HISTO_LENGTH) rotHist[bin].push_back(bestIdxF);
                        }
                        nmatches++;
                    }
                }
            }
            KFit++;
            Fit++;
        } else if (KFit->first < Fit->first) {
            KFit = vFeatVecKF.lower_bound(Fit->first);
        } else {
            Fit = vFeatVecF.lower_bound(KFit->first);
        }
    }
    if (check_orientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    if (n_matches) *n_matches = nmatches;
    return PLF_OK;
}

// match() + the gates of Tracking::TrackWithMotionModel (src/Tracking.cc:3055-3099, mode 0) and Tracking::SearchLocalLines
// (src/Tracking.cc:3879-3917, mode 1), restated loop by loop.
PLF_API int plf_cpu_match_lines_tracked(plf_ctx*, int mode, const uint8_t* desc1, const plf_track_line* lines1, int n1,
                                        const uint8_t* desc2, const plf_keyline* kl2, const float* disp2, const uint8_t* held2,
                                        int n2, float nnr, float min_x, float max_x, float min_y, float max_y, int32_t* matches12,
                                        int32_t* assign12, int* n_assigned) {
    if (mode < 0 || mode > 1 || n1 < 0 || n2 < 0 || (n1 && (!desc1 || !lines1 || !matches12 || !assign12)) ||
        (n2 && (!desc2 || !kl2 || !disp2)))
        return fail(PLF_ERR_INVALID, "bad arguments");
    if (n_assigned) *n_assigned = 0;
    if (n1 == 0) return PLF_OK;
    // mode 0: match(desc1, desc2, ...) = both directions + mutual best (src/LineMatcher.cpp:201-229).  mode 1:
    // match(vpLocalMapLines, CurrentFrame, ...) returns right after matchNNR(desc1, desc2) (src/LineMatcher.cpp:161-170).
    if (mode == 0) match_lr(desc1, n1, desc2, n2, nnr, 1, matches12);
    else match_nnr(desc1, n1, desc2, n2, nnr, matches12);
    const double kPiD = 3.14159265358979323846;
    const double deltaAngle = kPiD / 8.0;
    const double deltaWidth = (max_x - min_x) * 0.1;
    const double deltaHeight = (max_y - min_y) * 0.1;
    int n_inliers_ls = 0;
    // mCurrentFrame.mvpMapLines as the loop sees it: -2 = a line from before the loop with observations, -3 = none,
    // >= 0 = the line of that i1 (attached by this loop)
    std::vector<int> holder(n2 > 0 ? n2 : 1, -3);
    if (mode == 1 && held2) for (int i2 = 0; i2 < n2; ++i2) if (held2[i2]) holder[i2] = -2;
    for (int i1 = 0; i1 < n1; ++i1) assign12[i1] = -1;
    for (int i1 = 0; i1 < n1; ++i1) {
        if (mode == 0 && !lines1[i1].eligible) continue;
        const int i2 = matches12[i1];
        if (i2 < 0) continue;
        if (disp2[i2 * 2] < 0 || disp2[i2 * 2 + 1] < 0) continue;
        if (mode == 1) {                                      // if(mvpMapLines[i2]) if(mvpMapLines[i2]->Observations()>0) continue;
            const int h = holder[i2];
            if (h == -2 || (h >= 0 && lines1[h].eligible)) continue;
        }
        if (mode == 0) {
            double theta = kl2[i2].angle - lines1[i1].angle;
            if (theta < -kPiD) theta += 2 * kPiD;
            else if (theta > kPiD) theta -= 2 * kPiD;
            if (std::fabs(theta) > deltaAngle) { matches12[i1] = -1; continue; }
        }
        const float sX_curr = kl2[i2].startPointX, sX_last = lines1[i1].sx;
        const float sY_curr = kl2[i2].startPointY, sY_last = lines1[i1].sy;
        const float eX_curr = kl2[i2].endPointX, eX_last = lines1[i1].ex;
        const float eY_curr = kl2[i2].endPointY, eY_last = lines1[i1].ey;
        if (std::fabs(sX_curr - sX_last) > deltaWidth || std::fabs(eX_curr - eX_last) > deltaWidth ||
            std::fabs(sY_curr - sY_last) > deltaHeight || std::fabs(eY_curr - eY_last) > deltaHeight) {
            matches12[i1] = -1;
            continue;
        }
        if (holder[i2] >= 0) assign12[holder[i2]] = -1;      // mvpMapLines[i2] is overwritten
        holder[i2] = i1;
        assign12[i1] = i2;
        if (mode == 0) ++n_inliers_ls;
    }
    if (mode == 1) for (int i1 = 0; i1 < n1; ++i1) n_inliers_ls += assign12[i1] >= 0;
    if (n_assigned) *n_assigned = n_inliers_ls;
    return PLF_OK;
}

static void run_pair(plf_ctx* c, int b) {
    Slot& s = c->slots[b];
    const int w = c->p.width, h = c->p.height;
    for (int side = 0; side < 2; ++side) {
        if (c->p.has_points) orb_slot(c, b, side, s.img[side].d.data(), w, h, w, 0, 0);
        else { s.orb[side].kps.clear(); s.orb[side].desc.clear(); s.orb[side].valid = true; }   // line-only isolation (plf_params.has_points)
        line_slot(c, b, side, s.img[side].d.data(), w, h, w);
    }
    // Frame.cc:146-149: both matchers are skipped when there are no keypoints or no lines
    s.uRight.assign(s.orb[0].kps.size(), -1.f);
    s.depth.assign(s.orb[0].kps.size(), -1.f);
    s.disp.assign(s.lsd[0].kls.size() * 2, -1.f);
    s.le.assign(s.lsd[0].kls.size() * 3, 0.0);
    s.m12.assign(s.lsd[0].kls.size(), -1);
    if (c->p.has_points && s.orb[0].kps.empty()) return;
    if (c->p.has_lines && s.lsd[0].kls.empty()) return;
    if (c->p.has_lines) match_lines_slot(c, b);
    if (c->p.has_points) match_points_slot(c, b);
}

PLF_API int plf_cpu_batch_run(plf_ctx* c, int batch) {
    if (!c || batch < 1 || batch > c->batch_resident) return fail(PLF_ERR_INVALID, "bad batch");
    int nt = c->threads > 0 ? c->threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, batch));
    std::atomic<int> next(0);
    auto worker = [&]() { for (int b; (b = next.fetch_add(1)) < batch;) run_pair(c, b); };
    std::vector<std::thread> th;
    for (int i = 1; i < nt; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return PLF_OK;
}

PLF_API int plf_cpu_batch_download(plf_ctx* c, int batch, plf_frame_out* o) {
    if (!c || !o || batch < 1 || batch > c->batch_resident) return fail(PLF_ERR_INVALID, "bad batch");
    const int kc = o->kp_cap, lc = o->kl_cap;
    for (int b = 0; b < batch; ++b) {
        Slot& s = c->slots[b];
        const int nl = (int)s.orb[0].kps.size(), nr = (int)s.orb[1].kps.size();
        const int ll = (int)s.lsd[0].kls.size(), lr = (int)s.lsd[1].kls.size();
        if (nl > kc || nr > kc || ll > lc || lr > lc) return fail(PLF_ERR_INVALID, "output capacity too small");
        if (o->n_kp_left) o->n_kp_left[b] = nl;
        if (o->n_kp_right) o->n_kp_right[b] = nr;
        if (o->n_kl_left) o->n_kl_left[b] = ll;
        if (o->n_kl_right) o->n_kl_right[b] = lr;
        if (o->kp_left) std::memcpy(o->kp_left + (size_t)b * kc, s.orb[0].kps.data(), nl * sizeof(plf_keypoint));
        if (o->kp_right) std::memcpy(o->kp_right + (size_t)b * kc, s.orb[1].kps.data(), nr * sizeof(plf_keypoint));
        if (o->desc_left) std::memcpy(o->desc_left + (size_t)b * kc * 32, s.orb[0].desc.data(), (size_t)nl * 32);
        if (o->desc_right) std::memcpy(o->desc_right + (size_t)b * kc * 32, s.orb[1].desc.data(), (size_t)nr * 32);
        if (o->u_right) std::memcpy(o->u_right + (size_t)b * kc, s.uRight.data(), s.uRight.size() * 4);
        if (o->depth) std::memcpy(o->depth + (size_t)b * kc, s.depth.data(), s.depth.size() * 4);
        if (o->kl_left) std::memcpy(o->kl_left + (size_t)b * lc, s.lsd[0].kls.data(), ll * sizeof(plf_keyline));
        if (o->kl_right) std::memcpy(o->kl_right + (size_t)b * lc, s.lsd[1].kls.data(), lr * sizeof(plf_keyline));
        if (o->ldesc_left) std::memcpy(o->ldesc_left + (size_t)b * lc * 32, s.lsd[0].desc.data(), (size_t)ll * 32);
        if (o->ldesc_right) std::memcpy(o->ldesc_right + (size_t)b * lc * 32, s.lsd[1].desc.data(), (size_t)lr * 32);
        if (o->disp_se) std::memcpy(o->disp_se + (size_t)b * lc * 2, s.disp.data(), s.disp.size() * 4);
        if (o->le) std::memcpy(o->le + (size_t)b * lc * 3, s.le.data(), s.le.size() * 8);
        if (o->line_match12) std::memcpy(o->line_match12 + (size_t)b * lc, s.m12.data(), s.m12.size() * 4);
    }
    return PLF_OK;
}

PLF_API int plf_cpu_frontend_batch(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int stride, plf_frame_out* out) {
    int rc = plf_cpu_batch_upload(c, left, right, batch, stride);
    if (rc) return rc;
    rc = plf_cpu_batch_run(c, batch);
    if (rc) return rc;
    return plf_cpu_batch_download(c, batch, out);
}

PLF_API int plf_cpu_sync(plf_ctx*) { return PLF_OK; }
PLF_API int plf_cpu_batch_io_bytes(const plf_ctx* c, int64_t* a, int64_t* b) {
    if (!c) return PLF_ERR_INVALID;
    if (a) *a = 0;
    if (b) *b = 0;
    return PLF_OK;
}
PLF_API int plf_cpu_last_launch_count(const plf_ctx*) { return 0; }
PLF_API int plf_cpu_set_stage_timing(plf_ctx*, int) { return PLF_OK; }
PLF_API int plf_cpu_set_grower_policy(plf_ctx*, int) { return PLF_OK; }      // one scalar loop here: nothing to choose
PLF_API int plf_cpu_get_stage_ms(plf_ctx*, const char* const** names, const float** ms, int* n) {
    if (names) *names = nullptr;
    if (ms) *ms = nullptr;
    if (n) *n = 0;
    return PLF_OK;
}
PLF_API void* plf_cpu_stream(plf_ctx*) { return nullptr; }

// stage-level entry points used by tests/test_oracle_cv2.py to pin each primitive against cv2
PLF_API float plf_cpu_fast_atan2(float y, float x) { return fast_atan2(y, x); }
PLF_API int plf_cpu_prim_resize_linear(const uint8_t* src, int w, int h, uint8_t* dst, int dw, int dh) {
    Img8 s(w, h), d;
    std::memcpy(s.d.data(), src, (size_t)w * h);
    resize_linear_u8(s, d, dw, dh);
    std::memcpy(dst, d.d.data(), (size_t)dw * dh);
    return PLF_OK;
}
PLF_API int plf_cpu_prim_resize_exact(const uint8_t* src, int w, int h, double scale, uint8_t* dst, int* dw, int* dh) {
    Img8 s(w, h), d;
    std::memcpy(s.d.data(), src, (size_t)w * h);
    resize_linear_exact_u8(s, d, scale);
    if (dw) *dw = d.w;
    if (dh) *dh = d.h;
    if (dst) std::memcpy(dst, d.d.data(), d.d.size());
    return PLF_OK;
}
PLF_API int plf_cpu_prim_gaussian(const uint8_t* src, int w, int h, int ksize, double sigma, uint8_t* dst, int* taps_out) {
    Img8 s(w, h), d;
    std::memcpy(s.d.data(), src, (size_t)w * h);
    std::vector<int> taps;
    gaussian_taps_fixed(ksize, sigma, taps);
    if (taps_out) std::memcpy(taps_out, taps.data(), ksize * 4);
    gaussian_blur_u8(s, d, taps.data(), ksize);
    std::memcpy(dst, d.d.data(), d.d.size());
    return PLF_OK;
}
PLF_API int plf_cpu_prim_sobel(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    Img16 a, b;
    sobel3_16s(s, a, b);
    std::memcpy(dx, a.d.data(), a.d.size() * 2);
    std::memcpy(dy, b.d.data(), b.d.size() * 2);
    return PLF_OK;
}
PLF_API int plf_cpu_prim_fast(const uint8_t* src, int w, int h, int th, float* xyr, int cap, int* n) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    std::vector<Cand> v;
    fast_window(s, 0, 0, w, h, th, v);
    if (n) *n = (int)v.size();
    if ((int)v.size() > cap) return PLF_ERR_INVALID;
    for (size_t i = 0; i < v.size(); ++i) { xyr[3 * i] = v[i].x; xyr[3 * i + 1] = v[i].y; xyr[3 * i + 2] = v[i].resp; }
    return PLF_OK;
}
PLF_API int plf_cpu_prim_lsd(const uint8_t* src, int w, int h, double scale, int stable, float* xyxy, int cap, int* n) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    LsdConfig c;
    c.scale = scale;
    c.refine = stable >> 4;           // bits 4..: refine mode (test hook)
    stable &= 15;
    c.stable_order = stable != 0;
    LsdState st;
    lsd_detect(c, s, st);
    int m = (int)st.segs.size() / 4;
    if (n) *n = m;
    if (m > cap) return PLF_ERR_INVALID;
    std::memcpy(xyxy, st.segs.data(), st.segs.size() * 4);
    return PLF_OK;
}

// function-level entry points used by tests/test_oracle_ref.py to pin the restated reference-owned logic against the
// reference's own code (oracle/_ref)
PLF_API int plf_cpu_prim_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out, int cap, int* nout) {
    std::vector<Cand> in(n), res;
    for (int i = 0; i < n; ++i) in[i] = Cand{xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
    distribute_octree(in, minX, maxX, minY, maxY, N, res);
    if (nout) *nout = (int)res.size();
    if ((int)res.size() > cap) return PLF_ERR_INVALID;
    for (size_t i = 0; i < res.size(); ++i) { out[3 * i] = res[i].x; out[3 * i + 1] = res[i].y; out[3 * i + 2] = res[i].resp; }
    return PLF_OK;
}
PLF_API int plf_cpu_prim_lbd(const uint8_t* src, int w, int h, const plf_keyline* kls, int n, float* lbd72, uint8_t* desc) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    std::vector<plf_keyline> v(kls, kls + n);
    std::vector<float> f;
    std::vector<uint8_t> d;
    lbd_compute(s, v, f, d);
    if (lbd72) std::memcpy(lbd72, f.data(), f.size() * 4);
    if (desc) std::memcpy(desc, d.data(), d.size());
    return PLF_OK;
}
PLF_API int plf_cpu_prim_stereo_lines(const plf_params* p, int W, int H, const plf_keyline* klL, int nL, const uint8_t* dL,
                                      const plf_keyline* klR, int nR, const uint8_t* dR, float* disp_se, double* le, int32_t* m12) {
    LineMatchConfig mc;
    mc.best_lr_matches = p->best_lr_matches; mc.matching_s_ws = p->matching_s_ws; mc.min_ratio_12_l = p->min_ratio_12_l;
    mc.line_sim_th = p->line_sim_th; mc.min_disp = p->min_disp; mc.line_horiz_th = p->line_horiz_th;
    mc.stereo_overlap_th = p->stereo_overlap_th; mc.ls_min_disp_ratio = p->ls_min_disp_ratio;
    std::vector<plf_keyline> a(klL, klL + nL), b(klR, klR + nR);
    std::vector<uint8_t> da(dL, dL + (size_t)nL * 32), db(dR, dR + (size_t)nR * 32);
    std::vector<float> disp;
    std::vector<double> l;
    std::vector<int> m;
    stereo_match_lines(mc, W, H, a, da, b, db, disp, l, m);
    std::memcpy(disp_se, disp.data(), disp.size() * 4);
    std::memcpy(le, l.data(), l.size() * 8);
    for (int i = 0; i < nL; ++i) m12[i] = m[i];
    return PLF_OK;
}
PLF_API int plf_cpu_set_float_libm(int on) { int old = g_float_libm; g_float_libm = on != 0; return old; }
PLF_API int plf_cpu_prim_clip_line(int W, int H, long long* pts4) {
    return clip_line(W, H, pts4[0], pts4[1], pts4[2], pts4[3]) ? 1 : 0;
}
PLF_API int plf_cpu_prim_line_iterator_count(float x1, float y1, float x2, float y2, int W, int H) {
    return line_iterator_count(x1, y1, x2, y2, W, H);
}
PLF_API int plf_cpu_prim_hamming(const uint8_t* a, const uint8_t* b) { return hamming256(a, b); }

}  // extern "C"

namespace plfo { void lsd_stats(const LsdConfig& c, const Img8& img, long long out[8]); }
extern "C" PLF_API int plf_cpu_prim_lsd_stats(const uint8_t* src, int w, int h, long long* out) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    LsdConfig c;
    plfo::lsd_stats(c, s, out);
    return PLF_OK;
}

namespace plfo { struct StreamSimParams { int window; int fifo; double dperp; int heuristic; int lanes; };
void lsd_stream_sim(const LsdConfig& c, const Img8& img, const StreamSimParams& sp, std::vector<float>& segs, long long st[16]); }
extern "C" PLF_API int plf_cpu_prim_lsd_stream_sim(const uint8_t* src, int w, int h, int window, int fifo, double dperp, int heuristic, int lanes,
                                                   float* xyxy, int cap, int* n, long long* st16) {
    Img8 s(w, h);
    std::memcpy(s.d.data(), src, (size_t)w * h);
    LsdConfig c;
    plfo::StreamSimParams sp{window, fifo, dperp, heuristic, lanes};
    std::vector<float> segs;
    plfo::lsd_stream_sim(c, s, sp, segs, st16);
    if (n) *n = (int)segs.size() / 4;
    if ((int)segs.size() / 4 > cap) return PLF_ERR_INVALID;
    std::memcpy(xyxy, segs.data(), segs.size() * 4);
    return PLF_OK;
}
