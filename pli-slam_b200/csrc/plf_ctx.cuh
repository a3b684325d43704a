// Internal context of the CUDA frontend: geometry tables, device buffers, stream.  All device work of one context
// is issued on ctx->stream; buffers are sized once in plf_create for max_batch stereo pairs (2*max_batch images,
// image index = slot*2 + side) so that a batch run performs no allocation and no host synchronisation.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/plf_b200.h"

#define PLF_MAX_LEVELS 16
#define PLF_EDGE 19               // EDGE_THRESHOLD, src/ORBextractor.cc:72
#define PLF_MINB 16               // minBorder = EDGE_THRESHOLD-3, src/ORBextractor.cc:771
#define PLF_NOTDEF (-1024.0f)
#define PLF_MW_WARPS 16           // warps (= regions in flight) per image of that grower
#define PLF_SW_MAX_IMG 296         // launches of at most this many images use the streaming multi-warp grower (one block per SM: two waves)
#ifndef PLF_SW_WARPS
#define PLF_SW_WARPS 16            // warps (= regions in flight) per image of the streaming grower
#endif
#define PLF_SW_WARPBUF 65536       // ints of uncommitted records (headers + pixel lists + segment queue) per warp of that grower
#define PLF_MW_MAX_IMG 128         // launches of at most this many images use the multi-warp (several regions in flight) grower
#define PLF_FAST_TH 32            // rows of one FAST score tile (orb.cu FS_TH); same for its host-built tile table
#define PLF_BLUR_TH 32            // rows of one blur tile (blur.cuh); the host builds the tile table of the pyramid blur with it
#define PLF_GRID_COLS 64          // FRAME_GRID_COLS, include/Frame.h:60
#define PLF_GRID_ROWS 48          // FRAME_GRID_ROWS, include/Frame.h:59

struct PlfLevel {
    int w, h, pitch;
    long long off;        // byte offset of the level inside one image's pyramid block
    int nCells;           // FAST cells of this level that are actually run
    int cellFirst;        // index of the first cell in the cell table
    int quota;            // mnFeaturesPerLevel[level]
    int nIni;             // root nodes of the quadtree
    int candOff, candCap; // slice of the per-image candidate buffer
    int kpOff, kpCap;     // slice of the per-image per-level keypoint buffer
    float scale, invScale;
    int scaledPatch;      // (int)(31*scale)
    int xTab, yTab;       // offsets into the bilinear coefficient table (levels >= 1)
};

struct PlfTile { short level, x0, y0, pad; };   // one thread-block tile of a per-pyramid-level kernel
struct PlfLin { unsigned short ofs; short a0, a1, pad; };   // source index and the two fixed-point weights of one output coordinate

struct PlfCell {          // one FAST window, src/ORBextractor.cc:787-804
    short level;
    short x0, y0, x1, y1; // window [x0,x1) x [y0,y1) in level coordinates
    int outBase;          // first slot in the per-image candidate buffer
    int cap;
    int magic;            // ceil(2^20 / detection-area width): division-free raster index -> (x, y)
};

struct PlfGeom {          // passed by value to kernels (fits the 4 KB parameter space)
    int nLevels;
    int W, H;
    PlfLevel lv[PLF_MAX_LEVELS];
    int nCellsTotal, candCapTotal, kpLevelCapTotal;
    long long pyrBytes;   // bytes of one image's pyramid block (same layout for the blurred pyramid)
    int kpCap, klCap;     // output capacities per image
    int iniTh, minTh;
    int umax[16];
    // LSD
    int Ws, Hs, Ps;       // scaled image size and pitch (a multiple of 128; also the pitch of the |g|^2 map and, in bits, of the used bitmap)
    int lsdTaps[16], lsdK;
    double lsdScale, rho, prec;
    int nBins, minRegSize;
    int refine;           // LSD refine mode (0 none, 1 standard)
    double densityTh;
    int n2Thresh;         // largest |g|^2 (integer) with sqrt(|g|^2 / 4.0) <= rho: pixel defined iff |g|^2 > n2Thresh
    int segCap;           // max segments kept per image
    int seedCap;
};

struct __align__(16) PlfWinQ {   // one window query of the projection searches, as the device sees it (64 bytes)
    float x, y, xr, radius;      // window centre, right-image coordinate for the stereo check, half side
    int minLevel, maxLevel, skip, pad;
    unsigned desc[8];
};

struct PlfVocab {           // DBoW2 vocabulary tree on the device (plf_bow_set_vocabulary)
    int nNodes = 0, levels = 0;
    int *childFirst = nullptr, *childCount = nullptr, *child = nullptr, *word = nullptr;
    uint8_t* desc = nullptr;
    double* weight = nullptr;
};

struct plf_ctx {
    plf_params p;
    PlfGeom g;
    int device = 0;
    int nImgMax = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;                  // plf_batch_run: the line path runs beside the point path (fork / join)
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    // plf_orb_extract starts the LINE path of the image it was handed on a side stream (speculatively: the reference passes the
    // same image to both extractors of a side, src/Frame.cc:128-135), plf_line_extract collects it after an exact comparison
    cudaStream_t streamSpec[2] = {nullptr, nullptr};
    cudaEvent_t evUp = nullptr;
    bool specPending[2] = {false, false};   // line path of the image in level 0 of this side is running / done on streamSpec[side]
    bool orbFresh[2] = {false, false};      // level 0 of this side still holds the image plf_orb_extract uploaded
    bool specEnabled = false;               // learnt: line_extract(side) was called with the image orb_extract(side) had
    int specMisses = 0;
    uint8_t* d_cmp = nullptr;               // staging copy of the image plf_line_extract was handed (W x H)
    int* d_cmpFlag = nullptr;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
    // ORB device buffers
    uint8_t* d_pyr = nullptr;        // [nImg][pyrBytes]
    uint8_t* d_blur = nullptr;       // [nImg][pyrBytes]
    uint8_t* d_score = nullptr;      // [nImg][pyrBytes] FAST corner score map (0 below minTh), pyramid layout
    PlfCell* d_cells = nullptr;      // [nCellsTotal]
    PlfTile* d_tilesBlur = nullptr;  // 32x32 tiles of every pyramid level
    PlfTile* d_tilesFast = nullptr;  // 32x8 tiles of the FAST detection area of every level
    int nTilesBlur = 0, nTilesFast = 0;
    PlfLin* d_lin = nullptr;         // bilinear tables: pyramid levels (11-bit weights) then the LSD upscale (Q8)
    int linLsdX = 0, linLsdY = 0;    // offsets of the LSD upscale tables inside d_lin
    int* d_cellCount = nullptr;      // [nImg][nCellsTotal]
    uint32_t* d_cand = nullptr;      // [nImg][candCapTotal] packed x | y<<12 | score<<24 (relative to min border)
    uint32_t* d_scratch = nullptr;   // [nImg][2][candCapTotal] quadtree ping-pong point buffers
    uint32_t* d_lvlKp = nullptr;     // [nImg][kpLevelCapTotal] retained keypoints per level (packed, list order)
    int* d_lvlN = nullptr;           // [nImg][nLevels]
    plf_keypoint* d_kpTmp = nullptr; // [nImg][kpCap] level-major
    uint8_t* d_descTmp = nullptr;    // [nImg][kpCap][32]
    plf_keypoint* d_kp = nullptr;    // [nImg][kpCap] final row order
    uint8_t* d_desc = nullptr;       // [nImg][kpCap][32]
    int* d_nKp = nullptr;            // [nImg]
    int* d_mono = nullptr;           // [nImg]
    int* d_err = nullptr;            // [1] device-side overflow / invariant flags
    // stereo points
    float* d_uRight = nullptr;       // [slots][kpCap]
    float* d_depth = nullptr;
    int* d_sad = nullptr;            // [slots][kpCap] best SAD (-1 = unmatched)
    // LSD / LBD
    uint8_t* d_lsdBlur = nullptr;    // [nImg][H][pitch0]
    uint8_t* d_lsdU = nullptr;       // [nImg][Hs][Ps]
    const float4* d_gradLut = nullptr; // per-DEVICE table gradient code -> record (angle in degrees, cosf, sinf, |g|^2 as int bits), shared by every context of the device, never freed
    int* d_n2max = nullptr;          // [nImg]
    int* d_seeds = nullptr;          // [nImg][seedCap] seed pixels (packed y<<16|x) in processing order
    int* d_nSeeds = nullptr;         // [nImg]
    int* d_n2 = nullptr;             // [nImg][Hs][Ps] gradient code (gx + 512) | (gy + 512) << 10 of defined pixels, 0 where undefined:
                                     // index of the pixel's record in d_gradLut; |g|^2 for the seed order and the rectangle weights
    uint32_t* d_used = nullptr;      // [nImg][Hs][Ps/32] used bitmap of the region grower; undefined pixels start as used
    int* d_reg = nullptr;            // [nImg][Hs*Ws] region pixel list, packed y<<16|x (reused per region)
    uint32_t* d_owner = nullptr;     // small-batch grower: [min(nImg, PLF_MW_MAX_IMG)][Hs][Ps] owner tags of the current wave (PLF_FREE = none)
    int* d_regMW = nullptr;          // small-batch grower: [min(nImg, PLF_MW_MAX_IMG)][8][Hs*Ws] region lists, one per wave slot
    // streaming small-batch grower (lsd_grow_sw_kernel), for min(nImg, PLF_SW_MAX_IMG) images, allocated on first use:
    uint32_t* d_swOwner = nullptr;   // [Hs][Ps] owner tags (seed position + 1)
    int* d_swPos = nullptr;          // [Hs][Ps] seed-list position of every defined pixel
    int* d_swReg = nullptr;          // [Hs*Ws rounded to 4] commit buffer + PLF_SW_WARPS x PLF_SW_WARPBUF record buffers
    int* d_stream = nullptr;         // streaming grower: [nImg][StreamLayout.total] owner map, list chunks, ticket table, region table (lazy)
    int4* d_laneRT = nullptr;        // lane-per-image grower: [nImg][segCap] region table {arena offset, size, angle bits, -} (lazy)
    unsigned long long* d_growNs = nullptr;  // [nImg] ns every image spent in lsd_grow_kernel (stage timing only, lazy)
    int* d_nReg = nullptr;           // [nImg] regions entered in the table by the streaming / lane-per-image grower
    float* d_segs = nullptr;         // [nImg][segCap][4]
    int* d_nSegs = nullptr;          // [nImg]
    plf_keyline* d_kl = nullptr;     // [nImg][klCap]
    plf_keyline* d_klAll = nullptr;  // [nImg][segCap] before top-N
    int* d_nKl = nullptr;            // [nImg]
    uint8_t* d_lbdBlur = nullptr;    // [nImg][H][pitch0]
    short2* d_sobel = nullptr;       // [nImg][H][W] (dx,dy)
    float* d_lbd = nullptr;          // [nImg][klCap][72]
    uint8_t* d_ldesc = nullptr;      // [nImg][klCap][32]
    // line stereo matching
    unsigned long long* d_rowMask = nullptr;  // [slots][klCap][48]
    double2* d_dirR = nullptr;       // [slots][klCap]
    unsigned short* d_dmat = nullptr;// [slots][klCap][klCap] Hamming distance or 0xFFFF (not a surviving candidate)
    int* d_m21 = nullptr;            // [slots][klCap]
    int* d_m12 = nullptr;            // [slots][klCap]
    float* d_disp = nullptr;         // [slots][klCap][2]
    double* d_le = nullptr;          // [slots][klCap][3]
    // generic matcher scratch (match_nnr / match)
    uint8_t* d_mA = nullptr; uint8_t* d_mB = nullptr; int* d_mOut = nullptr; int* d_mOut2 = nullptr; int mCap = 0;
    uint8_t* d_stage = nullptr;      // [2][batch][H][stride] bulk H2D landing zone of plf_batch_upload
    size_t stageCap = 0;
    // rectification (cv::remap in front of the path): per camera, per output pixel {x | y << 16 (int16 source
    // position), (fy << 5) | fx (5-bit fractions)} — cv::remap's fixed-point form of the float maps, built once
    uint2* d_rmap[2] = {nullptr, nullptr};
    int* d_gridStart = nullptr;      // [max_batch][64*48+1] CSR of Frame::mGrid (plf_feature_grid), allocated on first use
    int* d_gridIdx = nullptr;        // [max_batch][kpCap]
    PlfVocab voc[2];                 // 0: ORB vocabulary, 1: line vocabulary
    int* d_bowWord = nullptr;        // [max_batch][max(kpCap, klCap)] outputs of plf_bow_transform
    int* d_bowNode = nullptr;
    double* d_bowWeight = nullptr;
    PlfWinQ* d_projQ = nullptr;          // plf_search_by_projection scratch: queries, counts, segment starts, candidate pool
    int* d_projCount = nullptr;
    int* d_projStart = nullptr;
    int2* d_projPool = nullptr;
    size_t projQCap = 0, projPoolCap = 0;
    uint8_t* d_scr = nullptr;        // grow-only scratch of the per-frame searches (plf_search_by_bow, plf_match_lines_tracked)
    size_t scrCap = 0;
    float* d_bpPose = nullptr;       // [max_batch][12] Rwc, Ow of plf_backproject
    float* d_bpX = nullptr;          // [max_batch][kpCap][3]
    double* d_bpL = nullptr;         // [max_batch][klCap][6]
    int srcW[2] = {0, 0}, srcH[2] = {0, 0};
    // pinned host staging for small result reads
    int* h_counts = nullptr;         // pinned
    // state
    int batchResident = 0;
    bool orbValid[2] = {false, false};
    bool lineValid[2] = {false, false};
    int launches = 0;
    cudaGraphExec_t graphExec = nullptr;  // plf_batch_run of `graphBatch` pairs as one graph (captured on the second call with that size)
    int graphBatch = 0, graphLaunches = 0, warmBatch = 0;
    bool stageTiming = false;
    int growerPolicy = 0;            // PLF_GROWER_AUTO / PLF_GROWER_THROUGHPUT (plf_set_grower_policy)
    std::vector<cudaEvent_t> ev;          // pool of timing events (marks)
    std::vector<const char*> markNames;   // name of the stage that STARTS at mark i
    std::vector<float> stageMs;
    int nMarks = 0;
};

// Stage marks: when stage timing is on, every launcher drops a CUDA event on the context stream before each stage;
// the elapsed times between consecutive marks are the ms/stage figures of bench.py.
void plf_mark(plf_ctx* c, const char* name);
extern "C" cudaError_t plf_enter(plf_ctx* c);      // cudaSetDevice + wait for a speculative line path on the side streams (capi.cu)

// --- stage launchers (each returns the number of kernel launches it issued) ---------------------------------------
int plf_launch_orb(plf_ctx* c, int imgFirst, int nImg, int lap0, int lap1);
// Opt-in dynamic shared memory (> 48 KB) is an attribute of (function, device) and the last value set wins, so the
// largest request per device is remembered process-wide and only ever raised.  Returns true if `smem` exceeds it.
#include <mutex>
inline bool plf_raise_smem_optin(size_t (&granted)[64], int device, size_t smem) {
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    if (smem <= 48 * 1024 || device < 0 || device >= 64 || smem <= granted[device]) return device < 0 || device >= 64;
    granted[device] = smem;
    return true;
}

int plf_launch_unpack(plf_ctx* c, const uint8_t* stage, size_t sideBytes, int stride, int batch);
int plf_launch_proj_candidates(plf_ctx* c, int slot, const PlfWinQ* dQ, int nq, const int* dCellStart,
                               const int* dCellIdx, int* dCount, const int* dSegStart, int2* dPool, bool fill);
int plf_launch_bow(plf_ctx* c, int which, int slotFirst, int nSlots, int levelsup, int* dWord, double* dWeight, int* dNode, int rows);
int plf_launch_bow_pairs(plf_ctx* c, int slot, const uint8_t* dKfDesc, const int4* dJobs, int nJobs, const int* dOrder, int* dPool);
int plf_launch_line_gates(plf_ctx* c, int mode, const plf_track_line* dL1, int n1, const plf_keyline* dK2, const float2* dDisp2,
                          const uint8_t* dHeld2, int n2, float minX, float maxX, float minY, float maxY, int* dM12, const int* dM21,
                          int* dAssign, int* dState, int* dBlocker, int* dLast);
int plf_launch_backproject(plf_ctx* c, int slotFirst, int nSlots, const float* dRwc, const float* dOw, float fy, float cx,
                           float cy, float* dX3d, int x3dRows, double* dL3d, int l3dRows);
int plf_launch_feature_grid(plf_ctx* c, int slotFirst, int nSlots, int* cellStart, int* cellIdx);
int plf_launch_rectify(plf_ctx* c, const uint8_t* raw0, const uint8_t* raw1, int rawStride, int imgFirst, int nImg);
int plf_launch_stereo_points(plf_ctx* c, int slotFirst, int nSlots);
int plf_launch_lines(plf_ctx* c, int imgFirst, int nImg);
const float4* plf_grad_lut(plf_ctx* c);          // builds the device's gradient record table on first use; nullptr on failure
int plf_launch_stereo_lines(plf_ctx* c, int slotFirst, int nSlots);
int plf_launch_match_nnr(plf_ctx* c, const uint8_t* dA, int nA, const uint8_t* dB, int nB, float nnr, int* dOut);

#define PLF_CUDA_OK(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return plf_set_cuda_error(_e, #expr, __FILE__, __LINE__);        \
    } while (0)
int plf_set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
int plf_fail(int code, const char* msg);          // sets plf_last_error() and returns `code`

// zero-initialised device allocation of n elements (at least one).  The memset runs on the legacy default stream, which
// does not order with the contexts' non-blocking streams: it is waited for here, otherwise a kernel launched on the
// context stream right after a lazy allocation can be overtaken by the memset and lose what it wrote.
template <typename T>
static inline cudaError_t dalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, (n > 0 ? n : 1) * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, (n > 0 ? n : 1) * sizeof(T));
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    return e;
}
