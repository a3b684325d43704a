// C ABI of the CUDA frontend (include/plf_b200.h): context creation (geometry tables + device buffers), H2D/D2H
// plumbing and the stage launchers.  No CPU fallback anywhere: without a usable sm_100 device plf_create fails.
#include "plf_ctx.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

static thread_local std::string g_err;

int plf_set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    g_err = buf;
    cudaGetLastError();          // reported: do not let a non-sticky error (e.g. out of memory) poison the next call's check
    return PLF_ERR_CUDA;
}
int plf_fail(int code, const char* msg) { g_err = msg; return code; }
static int fail(int code, const char* msg) { return plf_fail(code, msg); }

#define PLF_MAX_MARKS 48
void plf_mark(plf_ctx* c, const char* name) {
    if (!c->stageTiming || c->nMarks >= PLF_MAX_MARKS) return;
    cudaEventRecord(c->ev[c->nMarks], c->stream);
    c->markNames[c->nMarks] = name;
    c->nMarks++;
}

extern "C" {

PLF_API const char* plf_last_error(void) { return g_err.c_str(); }

PLF_API int plf_default_params(plf_params* p) {
    if (!p) return PLF_ERR_INVALID;
    memset(p, 0, sizeof(*p));
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

// Fixed-point Gaussian taps of cv::GaussianBlur's 8-bit path (error-diffusion rounding, sum 256).
static void gaussian_taps_fixed(int ksize, double sigma, int* taps) {
    double k[32], sum = 0;
    const int r = ksize / 2;
    for (int i = 0; i < ksize; ++i) { double x = i - r; k[i] = std::exp(-(x * x) / (2 * sigma * sigma)); sum += k[i]; }
    double err = 0;
    int s = 0;
    for (int i = 0; i < r; ++i) {
        double adj = k[i] / sum * 256.0 + err;
        int v0 = (int)std::nearbyint(adj);
        err = adj - v0;
        taps[i] = taps[ksize - 1 - i] = v0;
        s += v0;
    }
    taps[r] = 256 - 2 * s;
}

static int build_geometry(plf_ctx* c, std::vector<PlfCell>& cells) {
    const plf_params& p = c->p;
    PlfGeom& g = c->g;
    memset(&g, 0, sizeof g);
    const int L = p.n_levels;
    g.nLevels = L; g.W = p.width; g.H = p.height; g.iniTh = p.ini_th_fast; g.minTh = p.min_th_fast;
    // scale tables and per-level quotas: ORBextractor ctor, src/ORBextractor.cc:413-444
    c->scale.assign(L, 1.f); c->sigma2.assign(L, 1.f); c->invScale.resize(L); c->invSigma2.resize(L); c->quota.assign(L, 0);
    for (int i = 1; i < L; ++i) { c->scale[i] = c->scale[i - 1] * p.scale_factor; c->sigma2[i] = c->scale[i] * c->scale[i]; }
    for (int i = 0; i < L; ++i) { c->invScale[i] = 1.0f / c->scale[i]; c->invSigma2[i] = 1.0f / c->sigma2[i]; }
    {
        float factor = 1.0f / p.scale_factor;
        float nDesired = p.n_features * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L));
        int sum = 0;
        for (int l = 0; l < L - 1; ++l) { c->quota[l] = (int)std::nearbyint((double)nDesired); sum += c->quota[l]; nDesired *= factor; }
        c->quota[L - 1] = std::max(p.n_features - sum, 0);
    }
    // umax: src/ORBextractor.cc:452-467
    {
        const int HP = 15;
        int v, v0, vmax = (int)std::floor(HP * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(HP * std::sqrt(2.f) / 2);
        for (v = 0; v <= vmax; ++v) g.umax[v] = (int)std::nearbyint(std::sqrt((double)HP * HP - v * v));
        for (v = HP, v0 = 0; v >= vmin; --v) { while (g.umax[v0] == g.umax[v0 + 1]) ++v0; g.umax[v] = v0; ++v0; }
    }
    long long off = 0;
    int cellFirst = 0, candOff = 0, kpOff = 0;
    cells.clear();
    for (int l = 0; l < L; ++l) {
        PlfLevel& lv = g.lv[l];
        lv.w = l ? (int)std::nearbyintf((float)p.width * c->invScale[l]) : p.width;     // :1156-1157
        lv.h = l ? (int)std::nearbyintf((float)p.height * c->invScale[l]) : p.height;
        lv.pitch = (lv.w + 63) & ~63;
        lv.off = off;
        off += (long long)lv.pitch * lv.h;
        lv.scale = c->scale[l]; lv.invScale = c->invScale[l];
        lv.scaledPatch = (int)(31 * c->scale[l]);
        lv.quota = c->quota[l];
        // cell grid: src/ORBextractor.cc:771-804
        const int minB = PLF_MINB, maxBX = lv.w - PLF_EDGE + 3, maxBY = lv.h - PLF_EDGE + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        if (nCols < 1 || nRows < 1) return fail(PLF_ERR_UNSUPPORTED, "a pyramid level is smaller than one 30-px FAST cell");
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        if (wCell + 6 > 72 || hCell + 6 > 72) return fail(PLF_ERR_UNSUPPORTED, "FAST cell window exceeds 72 px");
        lv.nIni = (int)std::round(width / height);
        if (lv.nIni < 1 || lv.nIni > 4) return fail(PLF_ERR_UNSUPPORTED, "aspect ratio outside [0.5, 4.5): quadtree roots");
        lv.cellFirst = cellFirst;
        lv.candOff = candOff;
        for (int i = 0; i < nRows; ++i) {
            const int iniY = minB + i * hCell;
            int maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = maxBY;
            for (int j = 0; j < nCols; ++j) {
                const int iniX = minB + j * wCell;
                int maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = maxBX;
                PlfCell ce;
                ce.level = (short)l; ce.x0 = (short)iniX; ce.y0 = (short)iniY; ce.x1 = (short)maxX; ce.y1 = (short)maxY;
                const int aw = std::max(maxX - iniX - 6, 0), ah = std::max(maxY - iniY - 6, 0);
                ce.cap = ((aw + 1) / 2) * ((ah + 1) / 2);
                ce.magic = aw > 0 ? ((1 << 20) + aw - 1) / aw : 0;
                for (int i = 0; i < aw * ah; ++i)
                    if ((int)(((unsigned)i * (unsigned)ce.magic) >> 20) != i / aw) return fail(PLF_ERR_UNSUPPORTED, "FAST cell index magic is inexact");
                ce.outBase = candOff;
                candOff += ce.cap;
                cells.push_back(ce);
            }
        }
        lv.nCells = (int)cells.size() - cellFirst;
        cellFirst = (int)cells.size();
        lv.candCap = candOff - lv.candOff;
        lv.kpOff = kpOff;
        lv.kpCap = std::max(lv.quota + 3, 4 * lv.nIni);
        kpOff += lv.kpCap;
        if (lv.w > 4095 + 32 || lv.h > 4095 + 32) return fail(PLF_ERR_UNSUPPORTED, "image larger than 4096 px");
    }
    g.pyrBytes = (off + 255) & ~255LL;
    g.nCellsTotal = (int)cells.size();
    g.candCapTotal = candOff;
    g.kpLevelCapTotal = kpOff;
    g.kpCap = ((p.n_features + 3 * L + 31) / 32) * 32;
    if (g.kpCap < kpOff) g.kpCap = ((kpOff + 31) / 32) * 32;
    g.klCap = p.lsd_nfeatures > 0 ? p.lsd_nfeatures : 4096;
    // shared-memory bounds of the kernels that keep a whole list on chip: refuse loudly here instead of failing at launch
    {
        int maxQ = 0;
        for (int l = 0; l < L; ++l) maxQ = std::max(maxQ, std::max(g.lv[l].quota, 4 * g.lv[l].nIni));
        const size_t octree = (size_t)(maxQ + 16) * 36;      // octree_kernel: QNode (28 B) + two ints per pool entry, 16-bit links
        if (maxQ + 16 > 32000 || octree > 200 * 1024)
            return fail(PLF_ERR_UNSUPPORTED, "n_features too large: the quadtree node pool of one level exceeds shared memory (about 5800 features per level)");
        if ((size_t)g.kpCap * sizeof(int) > 200 * 1024)
            return fail(PLF_ERR_UNSUPPORTED, "n_features too large: the stereo cull keeps one SAD value per keypoint in shared memory (about 51000 features)");
        if (g.klCap > 8192) return fail(PLF_ERR_UNSUPPORTED, "lsd_nfeatures above 8192");
    }
    // LSD constants (OpenCV LineSegmentDetectorImpl::flsd)
    const double kPi = 3.14159265358979323846;
    g.lsdScale = p.lsd_scale;
    g.prec = kPi * p.lsd_ang_th / 180;
    g.rho = p.lsd_quant / std::sin(g.prec);
    g.nBins = p.lsd_n_bins;
    g.refine = p.lsd_refine;
    g.densityTh = p.lsd_density_th;
    g.n2Thresh = 0;
    for (int n2 = 0; n2 <= 2 * 510 * 510; ++n2) {      // exact integer image of LSD's `norm <= threshold` test
        if (std::sqrt((double)n2 / 4.0) <= g.rho) g.n2Thresh = n2; else break;
    }
    if (p.lsd_scale != 1) {
        const double sigma = (p.lsd_scale < 1) ? (p.lsd_sigma_scale / p.lsd_scale) : p.lsd_sigma_scale;
        const unsigned h = (unsigned)std::ceil(sigma * std::sqrt(2 * 3.0 * std::log(10.0)));
        g.lsdK = 1 + 2 * (int)h;
        if (g.lsdK > 7) return fail(PLF_ERR_UNSUPPORTED, "LSD pre-blur wider than 7 taps (sigma_scale too large)");
        gaussian_taps_fixed(g.lsdK, sigma, g.lsdTaps);
        g.Ws = (int)std::nearbyint(p.width * p.lsd_scale);
        g.Hs = (int)std::nearbyint(p.height * p.lsd_scale);
    } else {
        g.lsdK = 0; g.Ws = p.width; g.Hs = p.height;
    }
    g.Ps = (g.Ws + 127) & ~127;
    {
        const double logNT = 5 * (std::log10((double)g.Ws) + std::log10((double)g.Hs)) / 2 + std::log10(11.0);
        g.minRegSize = (int)(size_t)(-logNT / std::log10(p.lsd_ang_th / 180));
    }
    g.segCap = 8192;
    g.seedCap = g.Ws * g.Hs;
    if (g.nBins < 2 || g.nBins > 1536) return fail(PLF_ERR_UNSUPPORTED, "lsd_n_bins outside [2,1536] (32 x n_bins cursors must fit shared memory)");
    if (g.Ws >= 32768 || g.Hs >= 32768) return fail(PLF_ERR_UNSUPPORTED, "scaled LSD image side >= 32768 (packed coordinates)");
    return PLF_OK;
}

static int create_buffers(plf_ctx* c, const plf_params* p, std::vector<PlfCell>& cells);
PLF_API int plf_destroy(plf_ctx* c);

PLF_API int plf_create(const plf_params* p, int device, plf_ctx** out) {
    if (!p || !out) return fail(PLF_ERR_INVALID, "null argument");
    if (p->width < 64 || p->height < 64 || p->max_batch < 1 || p->n_levels < 1 || p->n_levels > PLF_MAX_LEVELS ||
        p->n_features < 1)
        return fail(PLF_ERR_INVALID, "bad image size / batch / levels / features");
    if (!(p->scale_factor > 1.0f && p->scale_factor <= 2.0f) || (p->has_lines && !(p->lsd_scale >= 0.5)))
        return fail(PLF_ERR_UNSUPPORTED, "scale_factor outside (1, 2] or lsd_scale < 0.5 (the resize kernels fetch the taps of 4 pixels from 12 source bytes)");
    if (p->lsd_refine < 0 || p->lsd_refine > 1) return fail(PLF_ERR_UNSUPPORTED, "lsd_refine = 2 (ADVANCED: NFA rectangle improvement) is not built");
    if (p->min_th_fast < 1 || p->min_th_fast > 126 || p->ini_th_fast < p->min_th_fast || p->ini_th_fast > 254)
        return fail(PLF_ERR_UNSUPPORTED, "FAST thresholds outside 1 <= minTh <= 126, minTh <= iniTh <= 254");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return fail(PLF_ERR_NO_DEVICE, "no usable CUDA device (this library has no CPU path)");
    cudaDeviceProp prop;
    PLF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(PLF_ERR_NO_DEVICE, "device is not sm_100 (kernels are built for sm_100a only)");
    PLF_CUDA_OK(cudaSetDevice(device));
    plf_ctx* c = new plf_ctx();
    c->p = *p;
    c->device = device;
    std::vector<PlfCell> cells;
    int rc = build_geometry(c, cells);
    if (!rc) rc = create_buffers(c, p, cells);
    if (rc) {                       // nothing of a half-built context survives (device memory included)
        const std::string why = g_err;
        plf_destroy(c);
        cudaGetLastError();
        g_err = why;
        return rc;
    }
    *out = c;
    return PLF_OK;
}

static int create_buffers(plf_ctx* c, const plf_params* p, std::vector<PlfCell>& cells) {
    PlfGeom& g = c->g;
    const size_t nImg = (size_t)p->max_batch * 2, nSlot = p->max_batch;
    c->nImgMax = (int)nImg;
    PLF_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PLF_CUDA_OK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    PLF_CUDA_OK(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    PLF_CUDA_OK(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    PLF_CUDA_OK(cudaStreamCreateWithFlags(&c->streamSpec[0], cudaStreamNonBlocking));
    PLF_CUDA_OK(cudaStreamCreateWithFlags(&c->streamSpec[1], cudaStreamNonBlocking));
    PLF_CUDA_OK(cudaEventCreateWithFlags(&c->evUp, cudaEventDisableTiming));
    const size_t npx = (size_t)g.Ws * g.Hs;
    PLF_CUDA_OK(dalloc(&c->d_pyr, nImg * g.pyrBytes + 1024));   // +256: the FAST tile loader reads whole 32-bit words
    PLF_CUDA_OK(dalloc(&c->d_blur, nImg * g.pyrBytes));
    PLF_CUDA_OK(dalloc(&c->d_score, nImg * g.pyrBytes + 1024));
    PLF_CUDA_OK(dalloc(&c->d_cells, cells.size()));
    PLF_CUDA_OK(cudaMemcpy(c->d_cells, cells.data(), cells.size() * sizeof(PlfCell), cudaMemcpyHostToDevice));
    {
        // tile tables (one entry per thread block) and bilinear coefficient tables, built once on the host
        std::vector<PlfTile> tb, tf;
        std::vector<PlfLin> lin;
        for (int l = 0; l < g.nLevels; ++l) {
            const PlfLevel& lv = g.lv[l];
            for (int y = 0; y < lv.h; y += PLF_BLUR_TH)
                for (int x = 0; x < lv.w; x += 128) tb.push_back(PlfTile{(short)l, (short)x, (short)y, 0});
            for (int y = PLF_EDGE; y < lv.h - PLF_EDGE; y += PLF_FAST_TH)
                for (int x = PLF_EDGE; x < lv.w - PLF_EDGE; x += 128) tf.push_back(PlfTile{(short)l, (short)x, (short)y, 0});
        }
        // cv::resize(INTER_LINEAR) 8U coefficients of level l from level l-1
        auto lin11 = [&](int nSrc, int nDst) {
            const double scale = (double)nSrc / nDst;
            for (int d = 0; d < nDst; ++d) {
                float f = (float)((d + 0.5) * scale - 0.5);
                int sx = (int)std::floor(f);
                f -= sx;
                if (sx < 0) { f = 0; sx = 0; }
                if (sx >= nSrc - 1) { f = 0; sx = nSrc - 1; }
                lin.push_back(PlfLin{(unsigned short)sx, (short)std::nearbyintf((1.f - f) * 2048.f), (short)std::nearbyintf(f * 2048.f), 0});
            }
        };
        // every table starts on a multiple of 4 entries and is followed by >= 4 padding entries: the resize kernels
        // fetch the x-entries of 4 adjacent output pixels as two 16-byte loads
        auto pad4 = [&] { for (int k = 0; k < 4; ++k) lin.push_back(PlfLin{0, 0, 0, 0}); while (lin.size() & 3) lin.push_back(PlfLin{0, 0, 0, 0}); };
        for (int l = 1; l < g.nLevels; ++l) {
            pad4();
            c->g.lv[l].xTab = (int)lin.size();
            lin11(g.lv[l - 1].w, g.lv[l].w);
            pad4();
            c->g.lv[l].yTab = (int)lin.size();
            lin11(g.lv[l - 1].h, g.lv[l].h);
        }
        // cv::resize(INTER_LINEAR_EXACT) Q8 coefficients of the LSD upscale (fact 4)
        auto lin8 = [&](int nSrc, int nDst) {
            const double inv = 1.0 / g.lsdScale;
            for (int d = 0; d < nDst; ++d) {
                const double f = inv * (d + 0.5) - 0.5;
                const int i = (int)std::floor(f);
                int o = 0, a = 0;
                if (i >= 0 && nSrc > 1) {
                    if (i < nSrc - 1) { o = i; a = (int)std::nearbyint((f - i) * 256.0); } else { o = nSrc - 1; a = 0; }
                }
                lin.push_back(PlfLin{(unsigned short)o, (short)(256 - a), (short)a, 0});
            }
        };
        pad4();
        c->linLsdX = (int)lin.size();
        lin8(g.W, g.Ws);
        pad4();
        c->linLsdY = (int)lin.size();
        lin8(g.H, g.Hs);
        pad4();
        c->nTilesBlur = (int)tb.size();
        c->nTilesFast = (int)tf.size();
        PLF_CUDA_OK(dalloc(&c->d_tilesBlur, tb.size()));
        PLF_CUDA_OK(dalloc(&c->d_tilesFast, tf.size()));
        PLF_CUDA_OK(dalloc(&c->d_lin, lin.size()));
        PLF_CUDA_OK(cudaMemcpy(c->d_tilesBlur, tb.data(), tb.size() * sizeof(PlfTile), cudaMemcpyHostToDevice));
        PLF_CUDA_OK(cudaMemcpy(c->d_tilesFast, tf.data(), tf.size() * sizeof(PlfTile), cudaMemcpyHostToDevice));
        PLF_CUDA_OK(cudaMemcpy(c->d_lin, lin.data(), lin.size() * sizeof(PlfLin), cudaMemcpyHostToDevice));
    }
    PLF_CUDA_OK(dalloc(&c->d_cellCount, nImg * g.nCellsTotal));
    PLF_CUDA_OK(dalloc(&c->d_cand, nImg * g.candCapTotal));
    PLF_CUDA_OK(dalloc(&c->d_scratch, nImg * 2 * g.candCapTotal));
    PLF_CUDA_OK(dalloc(&c->d_lvlKp, nImg * g.kpLevelCapTotal));
    PLF_CUDA_OK(dalloc(&c->d_lvlN, nImg * g.nLevels));
    PLF_CUDA_OK(dalloc(&c->d_kpTmp, nImg * g.kpCap));
    PLF_CUDA_OK(dalloc(&c->d_descTmp, nImg * g.kpCap * 32));
    PLF_CUDA_OK(dalloc(&c->d_kp, nImg * g.kpCap));
    PLF_CUDA_OK(dalloc(&c->d_desc, nImg * g.kpCap * 32));
    PLF_CUDA_OK(dalloc(&c->d_nKp, nImg));
    PLF_CUDA_OK(dalloc(&c->d_mono, nImg));
    PLF_CUDA_OK(dalloc(&c->d_err, 1));
    PLF_CUDA_OK(dalloc(&c->d_uRight, nSlot * g.kpCap));
    PLF_CUDA_OK(dalloc(&c->d_depth, nSlot * g.kpCap));
    PLF_CUDA_OK(dalloc(&c->d_sad, nSlot * g.kpCap));
    if (p->has_lines) {
        PLF_CUDA_OK(dalloc(&c->d_lsdBlur, nImg * (size_t)g.lv[0].pitch * g.H + 64));   // +64: the upscale reads whole words
        PLF_CUDA_OK(dalloc(&c->d_lsdU, nImg * (size_t)g.Ps * g.Hs));
        PLF_CUDA_OK(dalloc(&c->d_n2max, nImg));
        c->d_gradLut = plf_grad_lut(c);
        if (!c->d_gradLut) return fail(PLF_ERR_CUDA, "cannot allocate the LSD gradient record table");
        PLF_CUDA_OK(dalloc(&c->d_seeds, nImg * (size_t)g.seedCap + 8));      // + 8: the lane-per-image grower reads whole 16-byte groups
        PLF_CUDA_OK(dalloc(&c->d_nSeeds, nImg));
        PLF_CUDA_OK(dalloc(&c->d_n2, nImg * (size_t)g.Ps * g.Hs));
        PLF_CUDA_OK(dalloc(&c->d_used, nImg * (size_t)(g.Ps / 32) * g.Hs));
        PLF_CUDA_OK(dalloc(&c->d_reg, nImg * npx));
        PLF_CUDA_OK(dalloc(&c->d_segs, nImg * (size_t)g.segCap * 4));
        PLF_CUDA_OK(dalloc(&c->d_nSegs, nImg));
        PLF_CUDA_OK(dalloc(&c->d_nReg, nImg));
        PLF_CUDA_OK(dalloc(&c->d_klAll, nImg * (size_t)g.segCap));
        PLF_CUDA_OK(dalloc(&c->d_lbdBlur, nImg * (size_t)g.lv[0].pitch * g.H));
        PLF_CUDA_OK(dalloc(&c->d_sobel, nImg * (size_t)g.W * g.H));
    }
    PLF_CUDA_OK(dalloc(&c->d_kl, nImg * g.klCap));
    PLF_CUDA_OK(dalloc(&c->d_nKl, nImg));
    PLF_CUDA_OK(dalloc(&c->d_lbd, nImg * (size_t)g.klCap * 72));
    PLF_CUDA_OK(dalloc(&c->d_ldesc, nImg * (size_t)g.klCap * 32));
    PLF_CUDA_OK(dalloc(&c->d_rowMask, nSlot * (size_t)g.klCap * PLF_GRID_ROWS));
    PLF_CUDA_OK(dalloc(&c->d_dirR, nSlot * (size_t)g.klCap));
    PLF_CUDA_OK(dalloc(&c->d_dmat, nSlot * (size_t)g.klCap * g.klCap));
    PLF_CUDA_OK(dalloc(&c->d_m21, nSlot * (size_t)g.klCap));
    PLF_CUDA_OK(dalloc(&c->d_m12, nSlot * (size_t)g.klCap));
    PLF_CUDA_OK(dalloc(&c->d_disp, nSlot * (size_t)g.klCap * 2));
    PLF_CUDA_OK(dalloc(&c->d_le, nSlot * (size_t)g.klCap * 3));
    PLF_CUDA_OK(cudaMallocHost((void**)&c->h_counts, (nImg * 4 + 16) * sizeof(int)));
    c->ev.assign(PLF_MAX_MARKS, nullptr);
    for (auto& e : c->ev) PLF_CUDA_OK(cudaEventCreate(&e));
    c->markNames.assign(PLF_MAX_MARKS, "");
    c->stageMs.assign(PLF_MAX_MARKS, 0.f);
    return PLF_OK;
}

PLF_API int plf_destroy(plf_ctx* c) {
    if (!c) return PLF_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    void* ptrs[] = {c->d_pyr, c->d_blur, c->d_score, c->d_tilesBlur, c->d_tilesFast, c->d_lin, c->d_cells, c->d_cellCount, c->d_cand, c->d_scratch, c->d_lvlKp, c->d_lvlN,
                    c->d_kpTmp, c->d_descTmp, c->d_kp, c->d_desc, c->d_nKp, c->d_mono, c->d_err, c->d_uRight, c->d_depth,
                    c->d_sad, c->d_lsdBlur, c->d_lsdU, c->d_n2max, c->d_seeds,
                    c->d_nSeeds, c->d_n2, c->d_used, c->d_reg, c->d_owner, c->d_regMW, c->d_swOwner, c->d_swPos, c->d_swReg, c->d_stream, c->d_laneRT, c->d_scr, c->d_growNs, c->d_nReg, c->d_segs, c->d_nSegs, c->d_kl, c->d_klAll, c->d_nKl, c->d_lbdBlur,
                    c->d_sobel, c->d_lbd, c->d_ldesc, c->d_rowMask, c->d_dirR, c->d_dmat, c->d_m21, c->d_m12, c->d_disp,
                    c->d_le, c->d_mA, c->d_mB, c->d_mOut, c->d_mOut2, c->d_stage, c->d_rmap[0], c->d_rmap[1], c->d_gridStart, c->d_gridIdx, c->d_bpPose, c->d_bpX, c->d_bpL, c->d_bowWord, c->d_bowNode, c->d_bowWeight, c->d_projQ, c->d_projCount, c->d_projStart, c->d_projPool,
                    c->voc[0].childFirst, c->voc[0].childCount, c->voc[0].child, c->voc[0].word, c->voc[0].desc, c->voc[0].weight,
                    c->voc[1].childFirst, c->voc[1].childCount, c->voc[1].child, c->voc[1].word, c->voc[1].desc, c->voc[1].weight};
    for (void* q : ptrs) if (q) cudaFree(q);
    if (c->graphExec) cudaGraphExecDestroy(c->graphExec);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) if (c->streamSpec[k]) { cudaStreamSynchronize(c->streamSpec[k]); cudaStreamDestroy(c->streamSpec[k]); }
    if (c->evUp) cudaEventDestroy(c->evUp);
    if (c->d_cmp) cudaFree(c->d_cmp);
    if (c->d_cmpFlag) cudaFree(c->d_cmpFlag);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PLF_OK;
}

PLF_API int plf_keypoint_capacity(const plf_ctx* c) { return c ? c->g.kpCap : 0; }
PLF_API int plf_keyline_capacity(const plf_ctx* c) { return c ? c->g.klCap : 0; }
PLF_API void* plf_stream(plf_ctx* c) { return c ? (void*)c->stream : nullptr; }

PLF_API int plf_get_scale_tables(const plf_ctx* c, float* s, float* is, float* s2, float* is2, int32_t* n) {
    if (!c) return PLF_ERR_INVALID;
    const int L = c->g.nLevels;
    if (s) memcpy(s, c->scale.data(), L * 4);
    if (is) memcpy(is, c->invScale.data(), L * 4);
    if (s2) memcpy(s2, c->sigma2.data(), L * 4);
    if (is2) memcpy(is2, c->invSigma2.data(), L * 4);
    if (n) memcpy(n, c->quota.data(), L * 4);
    return PLF_OK;
}

static int check_device_flags(plf_ctx* c) {
    int e = 0;
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[0], c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    e = c->h_counts[0];
    if (e) {
        cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
        char buf[128];
        snprintf(buf, sizeof buf, "%s, flags=0x%x", (e & 8) ? "streaming region grower stalled (watchdog)" : "device-side capacity overflow", e);
        return fail(PLF_ERR_INVALID, buf);
    }
    return PLF_OK;
}

// ---- speculative line path of the single-image entry points --------------------------------------------------------
// The reference hands one image to ORBextractor::operator() and to Lineextractor::operator() of the same side (four threads,
// src/Frame.cc:128-135); through this ABI the four calls come one after the other.  Once the library has SEEN
// plf_line_extract(side) arrive with the image plf_orb_extract(side) had, plf_orb_extract starts the line path of its image on
// a side stream as well, and plf_line_extract — after comparing its image with the uploaded one, byte for byte, on the device —
// only collects the result: the two line extractions of a pair then overlap each other and the ORB kernels
// (five signatures of one pair: 22.7 -> see DESIGN.md).  Anything else that touches the context first waits for the side streams.
__global__ void __launch_bounds__(256) image_equal_kernel(const uint8_t* a, int ap, const uint8_t* b, int bp, int w, int h, int* differs) {
    const int y = blockIdx.y, x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (y >= h || x >= w) return;
    bool d = false;
    for (int k = 0; k < 4 && x + k < w; ++k) d |= a[(size_t)y * ap + x + k] != b[(size_t)y * bp + x + k];
    if (d) atomicOr(differs, 1);
}

static void plf_spec_join(plf_ctx* c, int side) {
    if (c->specPending[side]) {
        cudaStreamSynchronize(c->streamSpec[side]);
        c->specPending[side] = false;
        if (++c->specMisses >= 2) c->specEnabled = false;      // results nobody collected: stop guessing
    }
}

// every entry point but plf_orb_extract / plf_line_extract / plf_stereo_match_points: device, and nothing speculative in flight
cudaError_t plf_enter(plf_ctx* c) {
    const cudaError_t e = cudaSetDevice(c->device);
    plf_spec_join(c, 0);
    plf_spec_join(c, 1);
    c->orbFresh[0] = c->orbFresh[1] = false;
    return e;
}

// is `img` the image that sits in level 0 of this side's pyramid block?  (exact, on the device; ~50 us)
static int plf_same_image(plf_ctx* c, int side, const uint8_t* img, int stride, bool* same) {
    const PlfGeom& g = c->g;
    if (!c->d_cmp) {
        PLF_CUDA_OK(cudaMalloc((void**)&c->d_cmp, (size_t)g.W * g.H));
        PLF_CUDA_OK(cudaMalloc((void**)&c->d_cmpFlag, sizeof(int)));
    }
    PLF_CUDA_OK(cudaMemcpy2DAsync(c->d_cmp, g.W, img, stride, g.W, g.H, cudaMemcpyHostToDevice, c->stream));
    PLF_CUDA_OK(cudaMemsetAsync(c->d_cmpFlag, 0, sizeof(int), c->stream));
    image_equal_kernel<<<dim3((g.W + 1023) / 1024, g.H), 256, 0, c->stream>>>(c->d_cmp, g.W, c->d_pyr + (size_t)side * g.pyrBytes + g.lv[0].off,
                                                                              g.lv[0].pitch, g.W, g.H, c->d_cmpFlag);
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[3], c->d_cmpFlag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    *same = c->h_counts[3] == 0;
    return PLF_OK;
}

static int upload_image(plf_ctx* c, int img, const uint8_t* src, int stride) {
    const PlfGeom& g = c->g;
    PLF_CUDA_OK(cudaMemcpy2DAsync(c->d_pyr + (size_t)img * g.pyrBytes + g.lv[0].off, g.lv[0].pitch, src, stride, g.W, g.H,
                                  cudaMemcpyHostToDevice, c->stream));
    return PLF_OK;
}

PLF_API int plf_orb_extract(plf_ctx* c, int side, const uint8_t* img, int w, int h, int stride, int lap0, int lap1,
                            plf_keypoint* out_kp, uint8_t* out_desc, int cap, int* n, int* mono) {
    if (!c || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad ctx/side");
    if (!img || w <= 0 || h <= 0) return PLF_ERR_EMPTY_IMAGE;
    if (w != c->g.W || h != c->g.H || stride < w) return fail(PLF_ERR_INVALID, "image size differs from context");
    PLF_CUDA_OK(cudaSetDevice(c->device));
    plf_spec_join(c, side);                   // a line path nobody collected still reads this side's level 0
    int rc = upload_image(c, side, img, stride);
    if (rc) return rc;
    c->orbFresh[side] = true;
    static const bool s_noSpec = getenv("PLF_NO_SPEC") != nullptr || getenv("PLF_LSD_GROWER") != nullptr;
    if (c->specEnabled && !s_noSpec && c->p.has_lines && !c->stageTiming) {
        // the line path of this image, beside the ORB kernels and beside the other side's line path
        cudaStream_t s1 = c->stream;
        cudaEventRecord(c->evUp, s1);
        cudaStreamWaitEvent(c->streamSpec[side], c->evUp, 0);
        c->stream = c->streamSpec[side];
        plf_launch_lines(c, side, 1);
        c->stream = s1;
        c->specPending[side] = true;
        c->lineValid[side] = false;           // this side's line results are being replaced
    }
    c->launches = plf_launch_orb(c, side, 1, lap0, lap1);
    PLF_CUDA_OK(cudaGetLastError());
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[1], c->d_nKp + side, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[2], c->d_mono + side, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_flags(c);
    if (rc) return rc;
    const int nk = c->h_counts[1];
    if (nk > cap) return fail(PLF_ERR_INVALID, "keypoint capacity too small");
    if (out_kp) PLF_CUDA_OK(cudaMemcpyAsync(out_kp, c->d_kp + (size_t)side * c->g.kpCap, nk * sizeof(plf_keypoint), cudaMemcpyDeviceToHost, c->stream));
    if (out_desc) PLF_CUDA_OK(cudaMemcpyAsync(out_desc, c->d_desc + (size_t)side * c->g.kpCap * 32, (size_t)nk * 32, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (n) *n = nk;
    if (mono) *mono = c->h_counts[2];
    c->orbValid[side] = true;
    return PLF_OK;
}

static int copy_level_out(plf_ctx* c, const uint8_t* base, int slot, int side, int level, uint8_t* out, int out_stride, int* w, int* h) {
    if (!c || slot < 0 || slot * 2 + 1 >= c->nImgMax + 0 + (c->nImgMax == 0) || side < 0 || side > 1 || level < 0 || level >= c->g.nLevels)
        return fail(PLF_ERR_INVALID, "bad slot/side/level");
    const PlfLevel& lv = c->g.lv[level];
    if (w) *w = lv.w;
    if (h) *h = lv.h;
    if (!out) return PLF_OK;
    PLF_CUDA_OK(cudaSetDevice(c->device));                 // (reads the point path's pyramid only: a line path on a side stream is left alone)
    PLF_CUDA_OK(cudaMemcpy2DAsync(out, out_stride, base + (size_t)(slot * 2 + side) * c->g.pyrBytes + lv.off, lv.pitch, lv.w, lv.h, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return PLF_OK;
}

PLF_API int plf_tap_pyramid_level(plf_ctx* c, int slot, int side, int level, uint8_t* out, int out_stride, int* w, int* h) {
    return copy_level_out(c, c ? c->d_pyr : nullptr, slot, side, level, out, out_stride, w, h);
}
PLF_API int plf_get_pyramid_level(plf_ctx* c, int side, int level, uint8_t* out, int out_stride, int* w, int* h) {
    return copy_level_out(c, c ? c->d_pyr : nullptr, 0, side, level, out, out_stride, w, h);
}
PLF_API int plf_tap_blurred_level(plf_ctx* c, int slot, int side, int level, uint8_t* out, int out_stride) {
    return copy_level_out(c, c ? c->d_blur : nullptr, slot, side, level, out, out_stride, nullptr, nullptr);
}

PLF_API int plf_tap_fast_candidates(plf_ctx* c, int slot, int side, int level, float* xyr, int cap, int* n) {
    if (!c || slot < 0 || slot * 2 + 1 >= c->nImgMax + 1 || side < 0 || side > 1 || level < 0 || level >= c->g.nLevels) return fail(PLF_ERR_INVALID, "bad slot/side/level");
    PLF_CUDA_OK(plf_enter(c));
    const PlfGeom& g = c->g;
    const PlfLevel& lv = g.lv[level];
    const int img = slot * 2 + side;
    std::vector<int> cnt(lv.nCells);
    std::vector<uint32_t> cd(lv.candCap);
    std::vector<PlfCell> cells(lv.nCells);
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(cnt.data(), c->d_cellCount + (size_t)img * g.nCellsTotal + lv.cellFirst, lv.nCells * 4, cudaMemcpyDeviceToHost));
    PLF_CUDA_OK(cudaMemcpy(cd.data(), c->d_cand + (size_t)img * g.candCapTotal + lv.candOff, (size_t)lv.candCap * 4, cudaMemcpyDeviceToHost));
    PLF_CUDA_OK(cudaMemcpy(cells.data(), c->d_cells + lv.cellFirst, lv.nCells * sizeof(PlfCell), cudaMemcpyDeviceToHost));
    int k = 0;
    for (int ci = 0; ci < lv.nCells; ++ci)
        for (int i = 0; i < cnt[ci]; ++i, ++k) {
            if (k >= cap) continue;
            const uint32_t p = cd[cells[ci].outBase - lv.candOff + i];
            xyr[3 * k] = (float)((p & 0xFFF) + PLF_MINB);
            xyr[3 * k + 1] = (float)(((p >> 12) & 0xFFF) + PLF_MINB);
            xyr[3 * k + 2] = (float)(p >> 24);
        }
    if (n) *n = k;
    return k > cap ? fail(PLF_ERR_INVALID, "capacity too small") : PLF_OK;
}

PLF_API int plf_stereo_match_points(plf_ctx* c, float* u_right, float* depth, int cap) {
    if (!c) return PLF_ERR_INVALID;
    if (!c->orbValid[0] || !c->orbValid[1]) return fail(PLF_ERR_STATE, "stereo_match_points before both orb_extract calls");
    PLF_CUDA_OK(cudaSetDevice(c->device));                 // (point buffers only: a line path in flight on a side stream is left alone)
    c->launches = plf_launch_stereo_points(c, 0, 1);
    PLF_CUDA_OK(cudaGetLastError());
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[1], c->d_nKp, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    const int nk = c->h_counts[1];
    if (nk > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (u_right) PLF_CUDA_OK(cudaMemcpyAsync(u_right, c->d_uRight, nk * 4, cudaMemcpyDeviceToHost, c->stream));
    if (depth) PLF_CUDA_OK(cudaMemcpyAsync(depth, c->d_depth, nk * 4, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return PLF_OK;
}

PLF_API int plf_line_extract(plf_ctx* c, int side, const uint8_t* img, int w, int h, int stride, plf_keyline* out_kl,
                             uint8_t* out_desc, int cap, int* n) {
    if (!c || side < 0 || side > 1 || !img) return fail(PLF_ERR_INVALID, "bad ctx/side/img");
    if (w != c->g.W || h != c->g.H || stride < w) return fail(PLF_ERR_INVALID, "image size differs from context");
    PLF_CUDA_OK(cudaSetDevice(c->device));
    if (!c->p.has_lines) { if (n) *n = 0; c->lineValid[side] = true; cudaMemsetAsync(c->d_nKl + side, 0, sizeof(int), c->stream); return PLF_OK; }
    int rc = PLF_OK;
    bool collected = false;
    if (c->orbFresh[side] && (c->specPending[side] || (!c->specEnabled && c->specMisses < 2))) {
        // level 0 of this side holds the image plf_orb_extract was handed: the same one?
        bool same = false;
        rc = plf_same_image(c, side, img, stride, &same);
        if (rc) return rc;
        if (same && c->specPending[side]) {
            PLF_CUDA_OK(cudaStreamSynchronize(c->streamSpec[side]));      // the line path started by plf_orb_extract: collect it
            c->specPending[side] = false;
            c->specMisses = 0;
            collected = true;
        } else if (same) {
            c->specEnabled = true;                                         // learnt: from the next frame on, start it early
        }
    }
    if (!collected) {
        plf_spec_join(c, side);
        rc = upload_image(c, side, img, stride);
        if (rc) return rc;
        c->orbFresh[side] = false;
        // NOTE: the image lands in level 0 of this side's pyramid block; like the reference's Frame, callers pass the same
        // image to orb_extract and line_extract of one side.
        c->launches = plf_launch_lines(c, side, 1);
    }
    PLF_CUDA_OK(cudaGetLastError());
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[1], c->d_nKl + side, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    rc = check_device_flags(c);
    if (rc) return rc;
    const int nk = c->h_counts[1];
    if (nk > cap) return fail(PLF_ERR_INVALID, "keyline capacity too small");
    if (out_kl) PLF_CUDA_OK(cudaMemcpyAsync(out_kl, c->d_kl + (size_t)side * c->g.klCap, nk * sizeof(plf_keyline), cudaMemcpyDeviceToHost, c->stream));
    if (out_desc) PLF_CUDA_OK(cudaMemcpyAsync(out_desc, c->d_ldesc + (size_t)side * c->g.klCap * 32, (size_t)nk * 32, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (n) *n = nk;
    c->lineValid[side] = true;
    return PLF_OK;
}

PLF_API int plf_stereo_match_lines(plf_ctx* c, float* disp_se, double* le, int32_t* match12, int cap) {
    if (!c) return PLF_ERR_INVALID;
    if (!c->lineValid[0] || !c->lineValid[1]) return fail(PLF_ERR_STATE, "stereo_match_lines before both line_extract calls");
    PLF_CUDA_OK(plf_enter(c));
    c->launches = plf_launch_stereo_lines(c, 0, 1);
    PLF_CUDA_OK(cudaGetLastError());
    PLF_CUDA_OK(cudaMemcpyAsync(&c->h_counts[1], c->d_nKl, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    const int nk = c->h_counts[1];
    if (nk > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (disp_se) PLF_CUDA_OK(cudaMemcpyAsync(disp_se, c->d_disp, (size_t)nk * 8, cudaMemcpyDeviceToHost, c->stream));
    if (le) PLF_CUDA_OK(cudaMemcpyAsync(le, c->d_le, (size_t)nk * 24, cudaMemcpyDeviceToHost, c->stream));
    if (match12) PLF_CUDA_OK(cudaMemcpyAsync(match12, c->d_m12, (size_t)nk * 4, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return PLF_OK;
}

static int ensure_match_scratch(plf_ctx* c, int n) {
    if (n <= c->mCap) return PLF_OK;
    if (c->d_mA) cudaFree(c->d_mA);
    if (c->d_mB) cudaFree(c->d_mB);
    if (c->d_mOut) cudaFree(c->d_mOut);
    if (c->d_mOut2) cudaFree(c->d_mOut2);
    c->d_mA = nullptr; c->d_mB = nullptr; c->d_mOut = nullptr; c->d_mOut2 = nullptr;   // a failed allocation below leaves no stale pointer
    c->mCap = 0;
    const int cap = std::max(n, 1024);
    PLF_CUDA_OK(dalloc(&c->d_mA, (size_t)cap * 32));
    PLF_CUDA_OK(dalloc(&c->d_mB, (size_t)cap * 32));
    PLF_CUDA_OK(dalloc(&c->d_mOut, cap));
    PLF_CUDA_OK(dalloc(&c->d_mOut2, cap));
    c->mCap = cap;
    return PLF_OK;
}

static int match_impl(plf_ctx* c, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr,
                      int32_t* m12, int* nm) {
    if (!c || n1 < 0 || n2 < 0 || (n1 && !d1) || (n2 && !d2) || (n1 && !m12)) return fail(PLF_ERR_INVALID, "bad descriptors");
    PLF_CUDA_OK(plf_enter(c));
    if (nm) *nm = 0;
    if (n1 == 0) return PLF_OK;
    int rc = ensure_match_scratch(c, std::max(n1, n2));
    if (rc) return rc;
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_mA, d1, (size_t)n1 * 32, cudaMemcpyHostToDevice, c->stream));
    if (n2) PLF_CUDA_OK(cudaMemcpyAsync(c->d_mB, d2, (size_t)n2 * 32, cudaMemcpyHostToDevice, c->stream));
    c->launches = plf_launch_match_nnr(c, c->d_mA, n1, c->d_mB, n2, nnr, c->d_mOut);
    std::vector<int> m21;
    if (best_lr && n2 > 0) {
        c->launches += plf_launch_match_nnr(c, c->d_mB, n2, c->d_mA, n1, nnr, c->d_mOut2);
        m21.resize(n2);
        PLF_CUDA_OK(cudaMemcpyAsync(m21.data(), c->d_mOut2, (size_t)n2 * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    PLF_CUDA_OK(cudaGetLastError());
    PLF_CUDA_OK(cudaMemcpyAsync(m12, c->d_mOut, (size_t)n1 * 4, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    int cnt = 0;
    for (int i = 0; i < n1; ++i) {
        if (m12[i] >= 0 && best_lr && m21[m12[i]] != i) m12[i] = -1;   // mutual-best filter, LineMatcher.cpp:218-224
        cnt += m12[i] >= 0;
    }
    if (nm) *nm = cnt;
    return PLF_OK;
}

PLF_API int plf_match_nnr(plf_ctx* c, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int32_t* m12, int* nm) {
    return match_impl(c, d1, n1, d2, n2, nnr, 0, m12, nm);
}
PLF_API int plf_match(plf_ctx* c, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr, int best_lr, int32_t* m12, int* nm) {
    return match_impl(c, d1, n1, d2, n2, nnr, best_lr, m12, nm);
}

// ---- LSD / LBD taps ----------------------------------------------------------------------------------------------
PLF_API int plf_tap_lsd_scaled(plf_ctx* c, int slot, int side, uint8_t* out, int out_stride, int* w, int* h) {
    if (!c || !c->p.has_lines || slot < 0 || slot * 2 + 1 >= c->nImgMax + 1 || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad slot/side");
    if (w) *w = c->g.Ws;
    if (h) *h = c->g.Hs;
    if (!out) return PLF_OK;
    PLF_CUDA_OK(plf_enter(c));
    PLF_CUDA_OK(cudaMemcpy2DAsync(out, out_stride, c->d_lsdU + (size_t)(slot * 2 + side) * c->g.Ps * c->g.Hs, c->g.Ps, c->g.Ws, c->g.Hs, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return PLF_OK;
}
PLF_API int plf_tap_lsd_angles(plf_ctx* c, int slot, int side, float* out, int* w, int* h) {
    if (!c || !c->p.has_lines || slot < 0 || slot * 2 + 1 >= c->nImgMax + 1 || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad slot/side");
    if (w) *w = c->g.Ws;
    if (h) *h = c->g.Hs;
    if (!out) return PLF_OK;
    PLF_CUDA_OK(plf_enter(c));
    const size_t npx = (size_t)c->g.Ws * c->g.Hs;
    // the angle of a defined pixel is the .x of its record in the per-device table, indexed by the pixel's gradient code
    std::vector<int> code(npx);
    std::vector<float> ang((size_t)1 << 20);
    PLF_CUDA_OK(cudaMemcpy2DAsync(code.data(), (size_t)c->g.Ws * 4, c->d_n2 + (size_t)(slot * 2 + side) * c->g.Ps * c->g.Hs, (size_t)c->g.Ps * 4,
                                  (size_t)c->g.Ws * 4, c->g.Hs, cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaMemcpy2DAsync(ang.data(), 4, c->d_gradLut, 16, 4, ang.size(), cudaMemcpyDeviceToHost, c->stream));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < npx; ++i) out[i] = code[i] ? ang[code[i]] : PLF_NOTDEF;
    return PLF_OK;
}
PLF_API int plf_tap_lsd_segments(plf_ctx* c, int slot, int side, float* xyxy, int cap, int* n) {
    if (!c || !c->p.has_lines || slot < 0 || slot * 2 + 1 >= c->nImgMax + 1 || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad slot/side");
    PLF_CUDA_OK(plf_enter(c));
    const int img = slot * 2 + side;
    int m = 0;
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(&m, c->d_nSegs + img, 4, cudaMemcpyDeviceToHost));
    if (n) *n = m;
    if (m > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (xyxy) PLF_CUDA_OK(cudaMemcpy(xyxy, c->d_segs + (size_t)img * c->g.segCap * 4, (size_t)m * 16, cudaMemcpyDeviceToHost));
    return PLF_OK;
}
PLF_API int plf_tap_lbd_float(plf_ctx* c, int slot, int side, float* out, int cap, int* n) {
    if (!c || slot < 0 || slot * 2 + 1 >= c->nImgMax + 1 || side < 0 || side > 1) return fail(PLF_ERR_INVALID, "bad slot/side");
    PLF_CUDA_OK(plf_enter(c));
    const int img = slot * 2 + side;
    int m = 0;
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(&m, c->d_nKl + img, 4, cudaMemcpyDeviceToHost));
    if (n) *n = m;
    if (m > cap) return fail(PLF_ERR_INVALID, "capacity too small");
    if (out) PLF_CUDA_OK(cudaMemcpy(out, c->d_lbd + (size_t)img * c->g.klCap * 72, (size_t)m * 72 * 4, cudaMemcpyDeviceToHost));
    return PLF_OK;
}

// ---- batch ---------------------------------------------------------------------------------------------------------
PLF_API int plf_batch_upload(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int stride) {
    if (!c || !left || !right || batch < 1 || batch > c->p.max_batch || stride < c->g.W) return fail(PLF_ERR_INVALID, "bad batch");
    PLF_CUDA_OK(plf_enter(c));
    const PlfGeom& g = c->g;
    c->nMarks = 0;
    plf_mark(c, "h2d");
    // one bulk H2D copy per side into a staging buffer, then a device-side scatter into level 0 of every pyramid block
    // (2 DMA transfers per call instead of 2*batch strided ones)
    const size_t sideBytes = (size_t)batch * g.H * stride;
    if (c->stageCap < 2 * sideBytes) {
        if (c->d_stage) cudaFree(c->d_stage);
        c->d_stage = nullptr;
        c->stageCap = 0;
        PLF_CUDA_OK(cudaMalloc((void**)&c->d_stage, 2 * (size_t)c->p.max_batch * g.H * stride));
        c->stageCap = 2 * (size_t)c->p.max_batch * g.H * stride;
    }
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_stage, left, sideBytes, cudaMemcpyHostToDevice, c->stream));
    PLF_CUDA_OK(cudaMemcpyAsync(c->d_stage + sideBytes, right, sideBytes, cudaMemcpyHostToDevice, c->stream));
    plf_launch_unpack(c, c->d_stage, sideBytes, stride, batch);
    c->batchResident = batch;
    return PLF_OK;
}

// Frame::Frame(stereo) returns before both matchers when the left image has no keypoints or no keylines
// (`if(mvKeys.empty()) return; if(mvKeys_Line.empty()) return;`, src/Frame.cc:147-150): such a slot keeps the initial values
// of the match arrays (mvuRight / mvDepth -1, no line match, disparities -1, line equations 0).  The matchers have run for
// every slot of the batch; this kernel puts the initial values back where the constructor would not have called them.
__global__ void __launch_bounds__(256) frame_gate_kernel(PlfGeom g, const int* nKp, const int* nKl, int hasPoints, int hasLines,
                                                         float* uRight, float* depth, int* m12, float* disp, double* le) {
    const int slot = blockIdx.x;
    const int nk = nKp[slot * 2], nl = nKl[slot * 2];
    if (!((hasPoints && nk == 0) || (hasLines && nl == 0))) return;
    for (int i = threadIdx.x; i < nk; i += blockDim.x) {
        uRight[(size_t)slot * g.kpCap + i] = -1.f;
        depth[(size_t)slot * g.kpCap + i] = -1.f;
    }
    for (int i = threadIdx.x; i < nl; i += blockDim.x) {
        m12[(size_t)slot * g.klCap + i] = -1;
        disp[((size_t)slot * g.klCap + i) * 2] = -1.f;
        disp[((size_t)slot * g.klCap + i) * 2 + 1] = -1.f;
        le[((size_t)slot * g.klCap + i) * 3] = 0.0;
        le[((size_t)slot * g.klCap + i) * 3 + 1] = 0.0;
        le[((size_t)slot * g.klCap + i) * 3 + 2] = 0.0;
    }
}

PLF_API int plf_batch_run(plf_ctx* c, int batch) {
    if (!c || batch < 1 || batch > c->batchResident) return fail(PLF_ERR_INVALID, "bad batch (upload first)");
    PLF_CUDA_OK(plf_enter(c));
    int n = 0;
    if (c->nMarks && strcmp(c->markNames[0], "h2d") != 0) c->nMarks = 0;   // run without a fresh upload: restart marks
    if (c->nMarks > 1 && !(c->nMarks == 2 && strcmp(c->markNames[1], "rectify") == 0)) c->nMarks = 0;
    // The whole pass is a fixed sequence of ~30 launches and memsets on the context stream with no host round trip, so
    // from the second call with the same batch size on it is replayed as ONE CUDA graph launch (the first call runs
    // eagerly: it performs the lazy allocations and the one-time function attributes that must not happen under capture).
    // Stage timing needs the events between the stages and keeps the eager path; PLF_NO_GRAPH=1 disables the graph.
    static const bool s_noGraph = getenv("PLF_NO_GRAPH") != nullptr;
    const bool wantGraph = !s_noGraph && !c->stageTiming && c->warmBatch == batch;
    if (wantGraph && c->graphExec && c->graphBatch == batch) {
        PLF_CUDA_OK(cudaGraphLaunch(c->graphExec, c->stream));
        c->launches = c->graphLaunches;
        c->orbValid[0] = c->orbValid[1] = c->lineValid[0] = c->lineValid[1] = true;
        return PLF_OK;
    }
    const bool capture = wantGraph && cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    // The point path and the line path share nothing but the input image (level 0 of the pyramid block, read-only for
    // both), so the line path is forked onto a second stream and joined in front of the stereo matchers: inside one call
    // the latency-bound region grower runs beside the issue-bound ORB kernels.  Stage timing (marks between the stages of
    // ONE stream) and PLF_NO_FORK=1 keep the sequential order.
    static const bool s_noFork = getenv("PLF_NO_FORK") != nullptr;
    const bool fork = !s_noFork && !c->stageTiming && c->p.has_points && c->p.has_lines && c->stream2;
    if (fork) {
        cudaStream_t s1 = c->stream;
        cudaEventRecord(c->evFork, s1);
        cudaStreamWaitEvent(c->stream2, c->evFork, 0);
        c->stream = c->stream2;
        n += plf_launch_lines(c, 0, 2 * batch);
        c->stream = s1;
        cudaEventRecord(c->evJoin, c->stream2);
    }
    if (c->p.has_points) n += plf_launch_orb(c, 0, 2 * batch, 0, 0);
    else cudaMemsetAsync(c->d_nKp, 0, (size_t)2 * batch * sizeof(int), c->stream);
    if (fork) cudaStreamWaitEvent(c->stream, c->evJoin, 0);
    else if (c->p.has_lines) n += plf_launch_lines(c, 0, 2 * batch);
    if (c->p.has_lines) n += plf_launch_stereo_lines(c, 0, batch);
    if (c->p.has_points) n += plf_launch_stereo_points(c, 0, batch);
    if (c->p.has_points && c->p.has_lines) {
        frame_gate_kernel<<<batch, 256, 0, c->stream>>>(c->g, c->d_nKp, c->d_nKl, c->p.has_points, c->p.has_lines, c->d_uRight, c->d_depth,
                                                        c->d_m12, c->d_disp, reinterpret_cast<double*>(c->d_le));
        ++n;
    }
    plf_mark(c, "d2h");
    if (capture) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&c->graphExec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) { c->graphExec = nullptr; return plf_set_cuda_error(e, "graph capture of plf_batch_run", __FILE__, __LINE__); }
        c->graphBatch = batch;
        c->graphLaunches = n;
        PLF_CUDA_OK(cudaGraphLaunch(c->graphExec, c->stream));
    }
    PLF_CUDA_OK(cudaGetLastError());
    c->warmBatch = batch;
    c->launches = n;
    c->orbValid[0] = c->orbValid[1] = c->lineValid[0] = c->lineValid[1] = true;
    return PLF_OK;
}

PLF_API int plf_batch_download(plf_ctx* c, int batch, plf_frame_out* o) {
    if (!c || !o || batch < 1 || batch > c->batchResident) return fail(PLF_ERR_INVALID, "bad batch");
    PLF_CUDA_OK(plf_enter(c));
    const PlfGeom& g = c->g;
    if (o->kp_cap < g.kpCap || o->kl_cap < g.klCap) return fail(PLF_ERR_INVALID, "output capacity smaller than plf_keypoint_capacity/plf_keyline_capacity");
    cudaStream_t s = c->stream;
    const size_t kc = o->kp_cap, lc = o->kl_cap;
    // interleaved (left,right) device arrays -> separate host arrays: 2-D copies with a 2-row device pitch
#define D2H_2D(dst, dstRow, src, srcRow, rowBytes, rows) \
    PLF_CUDA_OK(cudaMemcpy2DAsync(dst, dstRow, src, srcRow, rowBytes, rows, cudaMemcpyDeviceToHost, s))
    if (o->n_kp_left) D2H_2D(o->n_kp_left, 4, c->d_nKp, 8, 4, batch);
    if (o->n_kp_right) D2H_2D(o->n_kp_right, 4, c->d_nKp + 1, 8, 4, batch);
    if (o->n_kl_left) D2H_2D(o->n_kl_left, 4, c->d_nKl, 8, 4, batch);
    if (o->n_kl_right) D2H_2D(o->n_kl_right, 4, c->d_nKl + 1, 8, 4, batch);
    const size_t kpRow = (size_t)g.kpCap * sizeof(plf_keypoint), dRow = (size_t)g.kpCap * 32;
    if (o->kp_left) D2H_2D(o->kp_left, kc * sizeof(plf_keypoint), c->d_kp, 2 * kpRow, kpRow, batch);
    if (o->kp_right) D2H_2D(o->kp_right, kc * sizeof(plf_keypoint), c->d_kp + g.kpCap, 2 * kpRow, kpRow, batch);
    if (o->desc_left) D2H_2D(o->desc_left, kc * 32, c->d_desc, 2 * dRow, dRow, batch);
    if (o->desc_right) D2H_2D(o->desc_right, kc * 32, c->d_desc + dRow, 2 * dRow, dRow, batch);
    if (o->u_right) D2H_2D(o->u_right, kc * 4, c->d_uRight, (size_t)g.kpCap * 4, (size_t)g.kpCap * 4, batch);
    if (o->depth) D2H_2D(o->depth, kc * 4, c->d_depth, (size_t)g.kpCap * 4, (size_t)g.kpCap * 4, batch);
    const size_t klRow = (size_t)g.klCap * sizeof(plf_keyline), ldRow = (size_t)g.klCap * 32;
    if (o->kl_left) D2H_2D(o->kl_left, lc * sizeof(plf_keyline), c->d_kl, 2 * klRow, klRow, batch);
    if (o->kl_right) D2H_2D(o->kl_right, lc * sizeof(plf_keyline), c->d_kl + g.klCap, 2 * klRow, klRow, batch);
    if (o->ldesc_left) D2H_2D(o->ldesc_left, lc * 32, c->d_ldesc, 2 * ldRow, ldRow, batch);
    if (o->ldesc_right) D2H_2D(o->ldesc_right, lc * 32, c->d_ldesc + ldRow, 2 * ldRow, ldRow, batch);
    if (o->disp_se) D2H_2D(o->disp_se, lc * 8, c->d_disp, (size_t)g.klCap * 8, (size_t)g.klCap * 8, batch);
    if (o->le) D2H_2D(o->le, lc * 24, c->d_le, (size_t)g.klCap * 24, (size_t)g.klCap * 24, batch);
    if (o->line_match12) D2H_2D(o->line_match12, lc * 4, c->d_m12, (size_t)g.klCap * 4, (size_t)g.klCap * 4, batch);
#undef D2H_2D
    plf_mark(c, "end");
    int rc = check_device_flags(c);   // synchronises the stream
    if (rc) return rc;
    if (c->stageTiming) {
        for (int i = 0; i + 1 < c->nMarks; ++i) cudaEventElapsedTime(&c->stageMs[i], c->ev[i], c->ev[i + 1]);
        if (c->nMarks > 0) c->nMarks--;   // the terminating mark is not a stage
    }
    return PLF_OK;
}

PLF_API int plf_frontend_batch(plf_ctx* c, const uint8_t* left, const uint8_t* right, int batch, int stride, plf_frame_out* out) {
    int rc = plf_batch_upload(c, left, right, batch, stride);
    if (rc) return rc;
    rc = plf_batch_run(c, batch);
    if (rc) return rc;
    return plf_batch_download(c, batch, out);
}

PLF_API int plf_sync(plf_ctx* c) {
    if (!c) return PLF_ERR_INVALID;
    PLF_CUDA_OK(plf_enter(c));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    return PLF_OK;
}

PLF_API int plf_batch_io_bytes(const plf_ctx* c, int64_t* h2d, int64_t* d2h) {
    if (!c) return PLF_ERR_INVALID;
    const PlfGeom& g = c->g;
    if (h2d) *h2d = 2LL * g.W * g.H;
    if (d2h) *d2h = 4 * 4 + 2LL * g.kpCap * (sizeof(plf_keypoint) + 32) + (long long)g.kpCap * 8 +
                    2LL * g.klCap * (sizeof(plf_keyline) + 32) + (long long)g.klCap * (8 + 24 + 4);
    return PLF_OK;
}

PLF_API int plf_last_launch_count(const plf_ctx* c) { return c ? c->launches : 0; }
PLF_API int plf_set_grower_policy(plf_ctx* c, int policy) {
    if (!c || (policy != PLF_GROWER_AUTO && policy != PLF_GROWER_THROUGHPUT)) return fail(PLF_ERR_INVALID, "bad grower policy");
    if (c->growerPolicy != policy) {
        c->growerPolicy = policy;
        if (c->graphExec) { cudaGraphExecDestroy(c->graphExec); c->graphExec = nullptr; }      // the captured pass holds the other kernel
        c->graphBatch = 0;
        c->warmBatch = 0;                       // the next pass runs eagerly (lazy allocations must not happen under capture)
    }
    return PLF_OK;
}
PLF_API int plf_set_stage_timing(plf_ctx* c, int on) { if (!c) return PLF_ERR_INVALID; c->stageTiming = on != 0; return PLF_OK; }
/* Per-image run time (ns) of the one-warp-per-image region grower during the last pass that ran with stage timing on: the
 * launch lasts as long as its slowest image, so max / mean says how much of the stage is tail.  Zeros for images that went
 * through another grower (launches of <= 296 images).  Test / benchmark tap, not part of the reference interface. */
PLF_API int plf_tap_grow_ns(plf_ctx* c, unsigned long long* out, int n_images) {
    if (!c || !out || n_images < 0 || n_images > c->nImgMax) return fail(PLF_ERR_INVALID, "bad image count");
    if (!c->d_growNs) { for (int i = 0; i < n_images; ++i) out[i] = 0; return PLF_OK; }
    PLF_CUDA_OK(plf_enter(c));
    PLF_CUDA_OK(cudaStreamSynchronize(c->stream));
    PLF_CUDA_OK(cudaMemcpy(out, c->d_growNs, (size_t)n_images * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return PLF_OK;
}

PLF_API int plf_get_stage_ms(plf_ctx* c, const char* const** names, const float** ms, int* n) {
    if (!c) return PLF_ERR_INVALID;
    if (names) *names = c->markNames.data();
    if (ms) *ms = c->stageMs.data();
    if (n) *n = c->nMarks;
    return PLF_OK;
}

}  // extern "C"
