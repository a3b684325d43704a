/*
 * plf_b200.h — C ABI of the B200-native stereo point-line frontend.
 *
 * This is the drop-in boundary for the hot path of VealFang/PLI-SLAM.  The reference has
 * no FFI layer; its seam is five C++ call signatures.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  Plain pointers and sizes only: no C++
 * types, no torch types.  All functions return a plf_status (0 = OK).
 *
 * Two libraries export this ABI with different prefixes:
 *   libplf_b200.so    plf_*      — the product: hand-written sm_100a CUDA kernels.  No CPU fallback: every
 *                                  call fails with PLF_ERR_CUDA / PLF_ERR_NO_DEVICE if the GPU is unusable.
 *   libplf_oracle.so  plf_cpu_*  — TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference path.
 * The macro PLF_FN(name) selects the prefix so both share these declarations.
 */
#ifndef PLF_B200_H
#define PLF_B200_H

#include <math.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef PLF_ORACLE_BUILD
#define PLF_FN(name) plf_cpu_##name
#else
#define PLF_FN(name) plf_##name
#endif

#if defined(__GNUC__)
#define PLF_API __attribute__((visibility("default")))
#else
#define PLF_API
#endif

typedef enum plf_status {
    PLF_OK = 0,
    PLF_ERR_INVALID = 1,      /* bad argument (null pointer, size mismatch, capacity too small)          */
    PLF_ERR_EMPTY_IMAGE = 2,  /* ORBextractor::operator() returns -1 on an empty image (ORBextractor.cc:1072) */
    PLF_ERR_CUDA = 3,         /* a CUDA call failed; plf_last_error() has the text                         */
    PLF_ERR_NO_DEVICE = 4,    /* no usable sm_100 device: the product path never falls back to the CPU     */
    PLF_ERR_UNSUPPORTED = 5,  /* parameter combination not built (e.g. lsd_refine == 2, n_features too large) */
    PLF_ERR_SIZE_MISMATCH = 6,/* std::runtime_error sites of LineMatcher.cpp:148-149,322-323                */
    PLF_ERR_STATE = 7         /* call order violated (e.g. stereo match before both extractions)           */
} plf_status;

/* Layout-identical to cv::KeyPoint (28 B): pt.x pt.y size angle response octave class_id. */
typedef struct plf_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} plf_keypoint;

/* Layout-identical to cv::line_descriptor::KeyLine (68 B), Thirdparty/line_descriptor/include/
 * line_descriptor/descriptor_custom.hpp:105-181. */
typedef struct plf_keyline {
    float angle;
    int32_t class_id;
    int32_t octave;
    float pt_x, pt_y;
    float response;
    float size;
    float startPointX, startPointY;
    float endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY;
    float ePointInOctaveX, ePointInOctaveY;
    float lineLength;
    int32_t numOfPixels;
} plf_keyline;

/* All tunables of the path as one POD.  Defaults (plf_default_params) are Examples/Stereo/Config/EuRoC.yaml
 * and src/Config.cpp:26-160. */
typedef struct plf_params {
    int32_t width, height;        /* image size the context is built for                                  */
    int32_t max_batch;            /* stereo pairs per batched call (>=1)                                   */
    /* ORBextractor ctor, src/ORBextractor.cc:408-411 */
    int32_t n_features;           /* 1200 */
    float   scale_factor;         /* 1.2  */
    int32_t n_levels;             /* 8    */
    int32_t ini_th_fast;          /* 20   */
    int32_t min_th_fast;          /* 7    */
    /* Lineextractor ctor, include/LineExtractor.h:45-47 */
    int32_t has_lines;            /* Config::hasLines()                                                    */
    int32_t lsd_nfeatures;        /* 500 in EuRoC.yaml:156 (300 = Config default); 0 keeps all             */
    int32_t lsd_refine;           /* 0 none (EuRoC.yaml), 1 standard; 2 (advanced, NFA) -> PLF_ERR_UNSUPPORTED    */
    int32_t lsd_n_bins;           /* 1024 */
    double  min_line_length;      /* 0.025 (relative to min(W,H))                                          */
    double  lsd_scale;            /* 1.2  */
    double  lsd_sigma_scale;      /* 0.6  */
    double  lsd_quant;            /* 2.0  */
    double  lsd_ang_th;           /* 22.5 */
    double  lsd_log_eps;          /* 1.0  */
    double  lsd_density_th;       /* 0.6  */
    /* stereo geometry: mbf and mb of src/Frame.cc:105,197 */
    float   bf;                   /* 47.90639 */
    float   fx;                   /* 435.2047; mb := bf/fx is applied BEFORE matching (oracle rule) */
    /* Config values read by Frame::ComputeStereoMatches_Lines / matchGrid */
    int32_t best_lr_matches;      /* true */
    int32_t matching_s_ws;        /* 10   */
    double  min_ratio_12_l;       /* 0.9  */
    double  line_sim_th;          /* 0.75 */
    double  min_disp;             /* 1.0  */
    double  line_horiz_th;        /* 0.1  */
    double  stereo_overlap_th;    /* 0.75 */
    double  ls_min_disp_ratio;    /* 0.7  */
    /* Isolation switch for the line-only configuration of BASELINE.json (C4): 0 skips the ORB extractors and the stereo
     * point matcher in the BATCHED calls (keypoint counts come back 0; lines and line matching still run).  1 = the
     * reference's path.  Not a reference behaviour: Frame::Frame(stereo) always extracts points. */
    int32_t has_points;           /* 1    */
} plf_params;

typedef struct plf_ctx plf_ctx;

/* Host-side result block of one batched call; the caller owns every array.
 * Strides: keypoint arrays are [batch][kp_cap], keyline arrays [batch][kl_cap]. */
typedef struct plf_frame_out {
    int32_t kp_cap, kl_cap;
    int32_t* n_kp_left;  int32_t* n_kp_right;     /* [batch] == Frame::N, mvKeysRight.size()               */
    int32_t* n_kl_left;  int32_t* n_kl_right;     /* [batch] == mvKeys_Line.size(), mvKeysRight_Line.size()*/
    plf_keypoint* kp_left;  plf_keypoint* kp_right;   /* mvKeys, mvKeysRight                               */
    uint8_t* desc_left;     uint8_t* desc_right;      /* mDescriptors(Right): [batch][kp_cap][32]          */
    float* u_right;         float* depth;             /* mvuRight, mvDepth: [batch][kp_cap]                */
    plf_keyline* kl_left;   plf_keyline* kl_right;    /* mvKeys_Line, mvKeysRight_Line                     */
    uint8_t* ldesc_left;    uint8_t* ldesc_right;     /* mDescriptors_Line(Right): [batch][kl_cap][32]     */
    float* disp_se;                                   /* mvDisparity_l: [batch][kl_cap][2]                 */
    double* le;                                       /* mvle_l: [batch][kl_cap][3]                        */
    int32_t* line_match12;                            /* matches_12 of matchGrid: [batch][kl_cap]          */
} plf_frame_out;

/* ------------------------------------------------------------------------------------------------------ */
/* lifetime                                                                                               */

PLF_API int PLF_FN(default_params)(plf_params* p);
/* Replaces: ORBextractor::ORBextractor (src/ORBextractor.cc:408-468) x2 + Lineextractor ctor x2
 * (src/Tracking.cc:87-98,743-749).  One context = the four extractor objects of one Tracking instance plus
 * device buffers for max_batch stereo pairs and one CUDA stream.  `device` is the CUDA ordinal (ignored by the
 * oracle). */
PLF_API int PLF_FN(create)(const plf_params* p, int device, plf_ctx** out);
PLF_API int PLF_FN(destroy)(plf_ctx* ctx);
PLF_API const char* PLF_FN(last_error)(void);
/* Capacity each out_kp/out_desc row block needs: n_features + 3*n_levels rounded up (DistributeOctTree can
 * overshoot N by <=3 per level, src/ORBextractor.cc:728). */
PLF_API int PLF_FN(keypoint_capacity)(const plf_ctx* ctx);
PLF_API int PLF_FN(keyline_capacity)(const plf_ctx* ctx);
/* ORBextractor getters (include/ORBextractor.h:65-85): scale_factors/inv/sigma2/inv_sigma2, each n_levels floats;
 * features_per_level n_levels ints.  Null pointers are skipped. */
PLF_API int PLF_FN(get_scale_tables)(const plf_ctx* ctx, float* scale, float* inv_scale, float* sigma2,
                                     float* inv_sigma2, int32_t* features_per_level);

/* ------------------------------------------------------------------------------------------------------ */
/* single-frame drop-in calls (batch slot 0).  `side`: 0 = left extractor, 1 = right extractor.           */

/* Replaces: int ORBextractor::operator()(InputArray im, InputArray mask, vector<KeyPoint>&, OutputArray desc,
 * vector<int>& vLappingArea)  (include/ORBextractor.h:61-63, src/ORBextractor.cc:1068-1150).
 * img: 8-bit single channel, `stride` bytes per row.  Writes *n keypoints/descriptor rows in the reference's row
 * order (mono rows front-to-back, lapping rows back-to-front) and *mono_index = the reference's return value. */
PLF_API int PLF_FN(orb_extract)(plf_ctx* ctx, int side, const uint8_t* img, int w, int h, int stride,
                                int lap0, int lap1, plf_keypoint* out_kp, uint8_t* out_desc, int cap,
                                int* n, int* mono_index);
/* Replaces: public member ORBextractor::mvImagePyramid[level] (include/ORBextractor.h:87) read by
 * Frame::ComputeStereoMatches.  Copies the level image (no 19-px border) to `out` with `out_stride`. */
PLF_API int PLF_FN(get_pyramid_level)(plf_ctx* ctx, int side, int level, uint8_t* out, int out_stride,
                                      int* w, int* h);
/* Replaces: void Lineextractor::operator()(const Mat& im, const Mat& mask, vector<KeyLine>&, Mat& desc)
 * (include/LineExtractor.h:49-51, src/LineExtractor.cc:31-70): LSDDetectorC::detect(opts) + top-N by response +
 * BinaryDescriptor::compute. */
PLF_API int PLF_FN(line_extract)(plf_ctx* ctx, int side, const uint8_t* img, int w, int h, int stride,
                                 plf_keyline* out_kl, uint8_t* out_desc, int cap, int* n);
/* Replaces: void Frame::ComputeStereoMatches() (include/Frame.h:152, src/Frame.cc:976-1154).  Uses the keypoints,
 * descriptors and pyramids left in the context by the two orb_extract calls.  u_right/depth: Frame::N floats. */
PLF_API int PLF_FN(stereo_match_points)(plf_ctx* ctx, float* u_right, float* depth, int cap);
/* Replaces: void Frame::ComputeStereoMatches_Lines(bool) (include/Frame.h:154, src/Frame.cc:1156-1307) incl.
 * matchGrid(lines) (src/LineMatcher.cpp:317-396).  disp_se: n x 2 floats (-1,-1 = mono line), le: n x 3 doubles,
 * match12: n ints (matches_12 after the mutual check; may be null). */
PLF_API int PLF_FN(stereo_match_lines)(plf_ctx* ctx, float* disp_se, double* le, int32_t* match12, int cap);
/* Replaces: int matchNNR(const Mat& d1, const Mat& d2, float nnr, vector<int>& m12) (include/LineMatcher.h:59,
 * src/LineMatcher.cpp:139-159).  d1: n1 x 32, d2: n2 x 32, row-major.  n2 < 2 -> no matches (oracle rule). */
PLF_API int PLF_FN(match_nnr)(plf_ctx* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr,
                              int32_t* m12, int* n_matches);
/* Replaces: int match(const Mat&, const Mat&, float nnr, vector<int>&) (src/LineMatcher.cpp:201-229); mutual-best
 * when best_lr != 0 (Config::bestLRMatches()). */
PLF_API int PLF_FN(match)(plf_ctx* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2, float nnr,
                          int best_lr, int32_t* m12, int* n_matches);

/* ------------------------------------------------------------------------------------------------------ */
/* batched calls: `batch` independent stereo pairs per call (<= max_batch); images are [batch][h][stride].  */

/* Replaces: Frame::Frame(stereo) extraction + matching (src/Frame.cc:128-163) for `batch` frames at once.
 * Host buffers; the H2D/D2H copies are part of the call. */
PLF_API int PLF_FN(frontend_batch)(plf_ctx* ctx, const uint8_t* left, const uint8_t* right, int batch,
                                   int stride, plf_frame_out* out);
/* The same split into its three phases so a caller can overlap them or keep inputs resident:
 *   upload  : H2D of `batch` pairs (async on the context stream)
 *   run     : every kernel of the path on the images currently resident in the context (async)
 *   download: D2H of the results + stream synchronise. */
PLF_API int PLF_FN(batch_upload)(plf_ctx* ctx, const uint8_t* left, const uint8_t* right, int batch, int stride);
PLF_API int PLF_FN(batch_run)(plf_ctx* ctx, int batch);
PLF_API int PLF_FN(batch_download)(plf_ctx* ctx, int batch, plf_frame_out* out);
PLF_API int PLF_FN(sync)(plf_ctx* ctx);
/* Bytes moved per stereo pair by upload and download (for the bench's e2e accounting). */
PLF_API int PLF_FN(batch_io_bytes)(const plf_ctx* ctx, int64_t* h2d_per_pair, int64_t* d2h_per_pair);
/* Number of kernel launches issued by the last batch_run (bench "gpu_launches"). 0 for the oracle. */
PLF_API int PLF_FN(last_launch_count)(const plf_ctx* ctx);
/* Device-time per stage of the last batch_run, in ms (CUDA events on the context stream; oracle: wall clock).
 * names: pointer to a static array of `*n_stages` C strings. Requires plf_set_stage_timing(ctx,1) before run. */
PLF_API int PLF_FN(set_stage_timing)(plf_ctx* ctx, int on);
PLF_API int PLF_FN(get_stage_ms)(plf_ctx* ctx, const char* const** names, const float** ms, int* n_stages);
/* Which region grower the line path uses (no counterpart in the reference: its LSD is one scalar loop).  Results are identical.
 *   PLF_GROWER_AUTO (default): launches of at most 296 images run 16 regions of every image concurrently (a single pair in
 *     ~11 ms instead of ~70; 1.9x the instructions), wider launches one warp per image.
 *   PLF_GROWER_THROUGHPUT: always one warp per image — for callers that keep MANY small or medium calls in flight and care
 *     about pairs per second, not about the latency of one call.  The oracle ignores the setting. */
#define PLF_GROWER_AUTO 0
#define PLF_GROWER_THROUGHPUT 1
PLF_API int PLF_FN(set_grower_policy)(plf_ctx* ctx, int policy);
/* Raw CUDA stream (cudaStream_t as void*) so a harness can record events on the launching stream. */
PLF_API void* PLF_FN(stream)(plf_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------ */
/* The step immediately BEFORE the path: stereo rectification of the raw camera frames.  */

/* Replaces: cv::remap(im, imRect, M1, M2, cv::INTER_LINEAR) (Examples/Stereo/stereo_euroc.cc:166-167; the ROS
 * nodes do the same, Examples/ROS/PLI_SLAM2/src/ros_stereo_inertial.cc:275-276) with the CV_32F maps of
 * cv::initUndistortRectifyMap (stereo_euroc.cc:117-118).  8-bit single channel, BORDER_CONSTANT 0, bit-exact
 * to cv::remap's fixed-point path: source coordinate = cvRound(map * 32) (5 fractional bits), four taps weighted
 * with 15-bit products of the two fractions, (sum + 2^14) >> 15.
 * map_x / map_y: height x width floats (the context's image size), row-major, for images of src_w x src_h.   */
PLF_API int PLF_FN(rectify_set_maps)(plf_ctx* ctx, int side, const float* map_x, const float* map_y,
                                     int src_w, int src_h);
/* One raw frame -> one rectified frame (host buffers). */
PLF_API int PLF_FN(rectify)(plf_ctx* ctx, int side, const uint8_t* raw, int raw_stride, uint8_t* out,
                            int out_stride);
/* batch_upload for RAW frames ([batch][src_h][raw_stride]): H2D + rectification on the device, straight into the
 * slots batch_run works on — the rectified image never exists on the host. */
PLF_API int PLF_FN(batch_upload_raw)(plf_ctx* ctx, const uint8_t* left_raw, const uint8_t* right_raw, int batch,
                                     int raw_stride);

/* ------------------------------------------------------------------------------------------------------ */
/* The last step of Frame::Frame(stereo) and the lookup it serves (first half).           */

#define PLF_GRID_COLS 64   /* FRAME_GRID_COLS, include/Frame.h:60 */
#define PLF_GRID_ROWS 48   /* FRAME_GRID_ROWS, include/Frame.h:59 */
/* Replaces: void Frame::AssignFeaturesToGrid() (src/Frame.cc:451-482, called at :204) with PosInGrid (:845-855) for
 * the left keypoints of slots [first_slot, first_slot + n_slots) (undistorted = detected keypoints for rectified
 * stereo; image bounds 0..width, 0..height as ComputeImageBounds sets them without distortion, :967-973).
 * mGrid[x][y] becomes CSR: cell = x * PLF_GRID_ROWS + y, cell_start[n_slots][64*48+1], cell_idx[n_slots][idx_stride]
 * (keypoint indices in ascending order inside a cell = push_back order). */
PLF_API int PLF_FN(feature_grid)(plf_ctx* ctx, int first_slot, int n_slots, int32_t* cell_start, int32_t* cell_idx,
                                 int idx_stride);

/* ------------------------------------------------------------------------------------------------------ */
/* Projection-window descriptor search (second half).                                     */

/* One MapPoint of the local map as ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, ...) reads it
 * (src/ORBmatcher.cc:44-130; rectified stereo, Nleft == -1).  56 bytes. */
typedef struct plf_proj_query {
    float proj_x, proj_y;       /* mTrackProjX, mTrackProjY                                                  */
    float proj_xr;              /* mTrackProjXR (stereo consistency check against mvuRight)                  */
    float view_cos;             /* mTrackViewCos -> RadiusByViewingCos (:216-222): > 0.998 ? 2.5 : 4.0       */
    int32_t level;              /* mnTrackScaleLevel                                                         */
    int32_t skip;               /* != 0: the `continue`s of :52-61 (not in view, too far, bad)               */
    uint8_t desc[32];           /* pMP->GetDescriptor()                                                      */
} plf_proj_query;

/* Replaces: int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th, ...)
 * (src/ORBmatcher.cc:44-130) for the left keypoints of one slot.  The window search, the level / stereo filters and
 * every Hamming distance run on the device (a warp per map point, candidates in GetFeaturesInArea order); the
 * order-dependent part — features already taken are skipped, best / second best with the reference's update rule, the
 * TH_HIGH and mfNNratio tests, the assignment F.mvpMapPoints[bestIdx] = pMP — runs on the host in query order.
 * occupied: one byte per keypoint, in/out: != 0 where F.mvpMapPoints[idx] has Observations() > 0; set for every feature
 * assigned by this call (map points of the local map have observations).  n_features = entries of occupied[]; fewer than
 * the slot's left keypoint count (Frame::N) -> PLF_ERR_INVALID.  match[i] = feature index given to query i or -1.
 * Returns the reference's nmatches in *n_matches. */
PLF_API int PLF_FN(search_by_projection)(plf_ctx* ctx, int slot, const plf_proj_query* queries, int n_queries, float th,
                                         float nn_ratio, int th_high, uint8_t* occupied, int n_features,
                                         int32_t* match, int* n_matches);

/* One map point of LastFrame already projected into CurrentFrame, as the frame-to-frame overload
 * ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono, match12)
 * (src/ORBmatcher.cc:2179-2323, Tracking::TrackWithMotionModel) has it after :2205-2241.  The projection itself
 * (a 3x4 transform per point) stays in the caller.  72 bytes. */
typedef struct plf_frame_query {
    float u, v;                 /* projection into CurrentFrame (:2217-2218)                                 */
    float ur;                   /* u - mbf * invzc (:2262)                                                   */
    float radius;               /* th * CurrentFrame.mvScaleFactors[nLastOctave] (:2229)                     */
    int32_t min_level, max_level; /* GetFeaturesInArea level arguments picked by bForward / bBackward
                                   * (:2233-2238): (oct, -1), (0, oct) or (oct - 1, oct + 1)                 */
    int32_t skip;               /* != 0: no map point, outlier, invzc < 0 or outside the image bounds        */
    int32_t has_observations;   /* pMP->Observations() > 0 (temporal points of UpdateLastFrame have none)    */
    float angle;                /* LastFrame.mvKeysUn[i].angle                                               */
    uint8_t desc[32];           /* pMP->GetDescriptor()                                                      */
    int32_t pad;
} plf_frame_query;

/* Replaces the search loop, the rotation histogram and ComputeThreeMaxima (:2449-2490) of that overload for the left
 * keypoints of one slot.  Device: window search, level / stereo filters, Hamming distances.  Host, in query order:
 * best distance among the features not occupied, TH_HIGH, CurrentFrame.mvpMapPoints[best] = pMP (a later map point
 * may overwrite an earlier one that has no observations, exactly as in the reference), match12.insert, the 30-bin
 * rotation histogram and the removal of everything outside its three maxima.
 * occupied  : in/out, one byte per keypoint: the feature holds a map point with Observations() > 0
 * feat_query: out, per keypoint: index of the query whose map point it holds at the end, or -1
 * match12   : out, per keypoint: the std::map<int,int> of the reference (first insertion wins), or -1; may be NULL
 * n_features: entries of occupied[] / feat_query[] / match12[]; fewer than Frame::N of the slot -> PLF_ERR_INVALID
 * Returns the reference's nmatches in *n_matches. */
PLF_API int PLF_FN(search_by_projection_frame)(plf_ctx* ctx, int slot, const plf_frame_query* queries, int n_queries,
                                               int th_high, int check_orientation, uint8_t* occupied, int n_features,
                                               int32_t* feat_query, int32_t* match12, int* n_matches);

/* Replaces: int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound,
 * const float th, const int ORBdist) (src/ORBmatcher.cc:2325-2447, Tracking::Relocalization) from the projected points
 * on, for the left keypoints of one slot (= CurrentFrame).  One plf_frame_query per keyframe map point i: u, v
 * (:2354), radius = th * mvScaleFactors[nPredictedLevel] (:2377), min_level / max_level = nPredictedLevel -/+ 1 (:2379),
 * angle = pKF->mvKeysUn[i].angle (:2410), desc, skip != 0 for the `continue`s of :2346-2372 (no map point, bad, already
 * found, outside the image bounds or the scale-invariance distances).  `ur` and `has_observations` are not read: this
 * overload has no stereo check and ANY map point held by a feature blocks it (:2393).  Device: window search, level
 * filter, Hamming distances.  Host, in query order: best distance among free features, bestDist <= orb_dist, assignment,
 * the 30-bin rotation histogram and ComputeThreeMaxima.
 * occupied  : in/out, one byte per keypoint: CurrentFrame.mvpMapPoints[idx] != NULL
 * feat_query: out, per keypoint: the query whose map point it received, or -1.  Returns nmatches in *n_matches. */
PLF_API int PLF_FN(search_by_projection_reloc)(plf_ctx* ctx, int slot, const plf_frame_query* queries, int n_queries,
                                               int orb_dist, int check_orientation, uint8_t* occupied, int n_features,
                                               int32_t* feat_query, int* n_matches);

/* Replaces: int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints,
 * vector<MapPoint*>& vpMatched, int th, float ratioHamming) (src/ORBmatcher.cc:473-586) and its vpPointsKFs /
 * vpMatchedKF variant (:588-704; the search is identical, feat_query tells the caller which pKFi to record), the loop
 * closing / place recognition searches, from the projected points on; the slot plays the keyframe.  One plf_frame_query
 * per candidate map point: u, v (:519 / :631-635), radius = th * mvScaleFactors[nPredictedLevel] (:545), min_level /
 * max_level = nPredictedLevel - 1 / nPredictedLevel (the filter of :566), desc, skip != 0 for the `continue`s of
 * :502-541.  No stereo check, no rotation histogram.  Host, in query order: best distance among the features with
 * vpMatched[idx] == NULL, accepted iff (float)bestDist <= (float)th_low * ratio_hamming (:574).
 * occupied: in/out, vpMatched[idx] != NULL.  feat_query: out, query index per keypoint or -1. */
PLF_API int PLF_FN(search_by_projection_loop)(plf_ctx* ctx, int slot, const plf_frame_query* queries, int n_queries,
                                              int th_low, float ratio_hamming, uint8_t* occupied, int n_features,
                                              int32_t* feat_query, int* n_matches);

/* Replaces: int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 * (src/ORBmatcher.cc:269-471; Tracking::TrackReferenceKeyFrame, Relocalization) for rectified stereo (F.Nleft == -1);
 * the slot plays F.  Keyframe side, per keyframe feature: descriptor row, keypoint angle (mvKeysUn), the key of its
 * FeatureVector entry (kf_node: node id at level L - levelsup as plf_bow_transform returns it, < 0 if the feature is in
 * no entry: stopped word) and kf_valid != 0 where vpMapPointsKF[i] is a good map point (:300-306).  Frame side:
 * f_node[n_features] likewise for the slot's left keypoints.  Device: every Hamming distance between a valid keyframe
 * feature and the frame features of the same node (a warp per keyframe feature).  Host, in the reference's order
 * (nodes ascending = std::map order, keyframe features of a node ascending, frame features ascending): features already
 * matched are skipped (:325), best / second best, best <= th_low and (float)best < nn_ratio * (float)second (:388-391),
 * rotation histogram with factor 1/HISTO_LENGTH and ComputeThreeMaxima.
 * match[f]: keyframe feature whose map point frame feature f received, or -1 (vpMapPointMatches); returns nmatches. */
PLF_API int PLF_FN(search_by_bow)(plf_ctx* ctx, int slot, const uint8_t* kf_desc, const float* kf_angle,
                                  const int32_t* kf_node, const uint8_t* kf_valid, int n_kf, const int32_t* f_node,
                                  int n_features, int th_low, float nn_ratio, int check_orientation, int32_t* match,
                                  int* n_matches);

/* One line of the "first" set of the tracking thread's line matching: a keyline of LastFrame (frame-to-frame) or the
 * projection of a local-map line (mTrackProjsX/sY/eX/eY).  24 bytes. */
typedef struct plf_track_line {
    float sx, sy, ex, ey;       /* start / end point in the current image                                    */
    float angle;                /* mvKeysUn_Line[i1].angle (frame-to-frame gate only)                        */
    int32_t eligible;           /* mode 0: LastFrame.mvpMapLines[i1] != NULL (:3064)
                                 * mode 1: pML->Observations() > 0 (decides whether a line attached by this very
                                 *         loop blocks the later ones that matched the same keyline, :3892-3894) */
} plf_track_line;

/* Replaces the line half of Tracking::TrackWithMotionModel (src/Tracking.cc:3055-3099; mode 0) and the matching half of
 * Tracking::SearchLocalLines (src/Tracking.cc:3879-3919; mode 1): the descriptor matching followed by the gates the
 * tracker applies to every match, in one device pass (2-NN, gates, one D2H of the two result arrays).
 *   mode 0 matches with match(desc1, desc2, nnr, matches_12) (src/LineMatcher.cpp:201-229: both directions, mutual best);
 *   mode 1 with match(vpLocalMapLines, CurrentFrame, nnr, matches_12), which RETURNS AFTER matchNNR(desc1, desc2)
 *   (src/LineMatcher.cpp:161-170; the mutual-best code behind the return is dead), so several lines may match one keyline
 *   and the loop is order-dependent: reproduced exactly (the first passing line that has observations keeps the keyline).
 *   lines1[n1]: the first set; kl2[n2]: mCurrentFrame.mvKeysUn_Line; disp2[n2][2]: mCurrentFrame.mvDisparity_l
 *   held2[n2] (mode 1, may be NULL): mCurrentFrame.mvpMapLines[i2] has Observations() > 0 (:3892-3894)
 *   min_x .. max_y: Frame::mnMinX/mnMaxX/mnMinY/mnMaxY (deltaWidth/Height = (max - min) * 0.1, :3060-3061)
 * For every i1 with a match i2: (mode 0) not eligible, a negative disparity (:3067) or (mode 1) a keyline whose current
 * holder has observations -> the match is kept but nothing is assigned; mode 0: |angle2 - angle1| folded to (-pi, pi] >
 * pi/8 -> matches12[i1] = -1 (:3072-3079); both modes: an end point farther than deltaWidth / deltaHeight from its
 * counterpart -> matches12[i1] = -1 (:3080-3093, :3901-3914); otherwise the line is attached to the keyline.
 *   matches12[n1]: out.  assign12[n1]: out, the i2 with mCurrentFrame.mvpMapLines[i2] == line i1 when the loop ends, else -1.
 *   *n_assigned = the reference's n_inliers_ls of mode 0 / the number of keylines that received a line in mode 1. */
PLF_API int PLF_FN(match_lines_tracked)(plf_ctx* ctx, int mode, const uint8_t* desc1, const plf_track_line* lines1, int n1,
                                        const uint8_t* desc2, const plf_keyline* kl2, const float* disp2,
                                        const uint8_t* held2, int n2, float nnr, float min_x, float max_x, float min_y,
                                        float max_y, int32_t* matches12, int32_t* assign12, int* n_assigned);

/* ------------------------------------------------------------------------------------------------------ */
/* Landmark back-projection: the epilogue that turns stereo matches into 3-D landmarks.   */

/* Replaces: cv::Mat Frame::UnprojectStereo(const int& i) (src/Frame.cc:1332-1347; callers src/Tracking.cc:1977, 2866,
 * 3651) for every left keypoint, and Eigen::Vector3d Frame::backProjection(u, v, disp) (src/Frame.cc:1349-1358;
 * callers src/Tracking.cc:1992-1999, 2238-2247, 2904-2911, 3727-3734) for both end points of every left line, of the
 * slots [first_slot, first_slot + n_slots) after batch_run / the stereo matchers.
 *   Rwc: n_slots x 9 floats (row-major mRwc), Ow: n_slots x 3 floats (mOw); fx, bf from plf_params (mb = bf / fx),
 *   fy / cx / cy as given (invfx = 1.0f / fx as in src/Frame.cc:190-191).
 *   x3d: n_slots x x3d_rows x 3 floats; a row is (0,0,0) where mvDepth <= 0 (the reference returns an empty Mat).
 *        Arithmetic of cv::Mat: x = (u-cx)*z*invfx in float; mRwc*x3Dc+mOw = cv::gemm's 3x3 path (float products summed
 *        left to right, then (float)((double)t + (double)mOw)).
 *   l3d: n_slots x l3d_rows x 6 doubles (start xyz, end xyz); zeros unless both disparities are > 0.  Arithmetic of
 *        the Eigen expression in double: bd = mb/disp, P = (bd*(u-cx), bd*(v-cy), bd*fx), R*P+Ow summed left to right.
 * Either output may be NULL. */
PLF_API int PLF_FN(backproject)(plf_ctx* ctx, int first_slot, int n_slots, const float* Rwc, const float* Ow,
                                float fy, float cx, float cy, float* x3d, int x3d_rows, double* l3d, int l3d_rows);

/* ------------------------------------------------------------------------------------------------------ */
/* Bag-of-words transform: Frame::ComputeBoW (src/Frame.cc:858-870).                      */

/* The DBoW2 vocabulary tree (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h, m_nodes; ORBvoc / the line vocabulary:
 * k = 10, L = 6, TF_IDF weighting, L1_NORM scoring) as flat arrays.  Node 0 is the root; the children of node i are
 * child[child_first[i] .. child_first[i] + child_count[i]) in the order of m_nodes[i].children; leaves have
 * child_count 0 and carry word_id / weight.  `which`: 0 = ORB vocabulary, 1 = line (LBD) vocabulary. */
PLF_API int PLF_FN(bow_set_vocabulary)(plf_ctx* ctx, int which, int n_nodes, int levels, const int32_t* child_first,
                                       const int32_t* child_count, const int32_t* child, const uint8_t* desc,
                                       const int32_t* word_id, const double* weight);
/* Replaces the per-feature half of TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
 * (TemplatedVocabulary.h:1139-1205 -> :1230-1270): every left descriptor of the slots (which = 0: ORB rows, 1: LBD
 * rows) is propagated down the tree by smallest Hamming distance (first child wins ties); outputs per feature the
 * word id, the word weight and the node id at level L - levelsup (0 if that level is <= 0).  stride = rows per slot. */
PLF_API int PLF_FN(bow_transform)(plf_ctx* ctx, int which, int first_slot, int n_slots, int levelsup,
                                  int32_t* word_id, double* weight, int32_t* node_id, int stride);

/* ------------------------------------------------------------------------------------------------------ */
/* stage taps for parity tests (slot = batch index, side 0/1).  Not part of the reference interface.       */

PLF_API int PLF_FN(tap_blurred_level)(plf_ctx* ctx, int slot, int side, int level, uint8_t* out, int out_stride);
PLF_API int PLF_FN(tap_pyramid_level)(plf_ctx* ctx, int slot, int side, int level, uint8_t* out, int out_stride,
                                      int* w, int* h);
/* vToDistributeKeys of one level (src/ORBextractor.cc:776-853): x,y (level coords incl. the 16-px min border),
 * response; cell-major then raster order. */
PLF_API int PLF_FN(tap_fast_candidates)(plf_ctx* ctx, int slot, int side, int level, float* xyr, int cap, int* n);
/* LSD internals: the x1.2 scaled image, the level-line angle (degrees, float; -1024 = NOTDEF), segments in
 * detection order (x1,y1,x2,y2 in input-image coordinates). */
PLF_API int PLF_FN(tap_lsd_scaled)(plf_ctx* ctx, int slot, int side, uint8_t* out, int out_stride, int* w, int* h);
PLF_API int PLF_FN(tap_lsd_angles)(plf_ctx* ctx, int slot, int side, float* out, int* w, int* h);
PLF_API int PLF_FN(tap_lsd_segments)(plf_ctx* ctx, int slot, int side, float* xyxy, int cap, int* n);
/* Nanoseconds every image spent in the one-warp-per-image region grower during the last pass run with stage timing on
 * (image index = slot * 2 + side; zeros for launches that used another grower). */
PLF_API int PLF_FN(tap_grow_ns)(plf_ctx* ctx, unsigned long long* out, int n_images);
/* LBD float descriptor (72 floats per line) before binarisation. */
PLF_API int PLF_FN(tap_lbd_float)(plf_ctx* ctx, int slot, int side, float* out, int cap, int* n);

/* Replaces: vector<size_t> Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel, bRight=false)
 * (src/Frame.cc:774-843) on the CSR grid of plf_feature_grid: indices of the keypoints with |dx| < r and |dy| < r
 * whose octave lies in [minLevel, maxLevel] (levels checked iff minLevel > 0 or maxLevel >= 0), in cell-column, cell-
 * row, push_back order.  Pure host inline; returns the count (at most `cap` are written). */
static inline int plf_features_in_area(const plf_keypoint* kps, const int32_t* cell_start, const int32_t* cell_idx,
                                       int width, int height, float x, float y, float r, int min_level,
                                       int max_level, int32_t* out, int cap) {
    const float inv_w = (float)PLF_GRID_COLS / ((float)width - 0.0f), inv_h = (float)PLF_GRID_ROWS / ((float)height - 0.0f);
    int x0 = (int)floorf((x - 0.0f - r) * inv_w), x1 = (int)ceilf((x - 0.0f + r) * inv_w);
    int y0 = (int)floorf((y - 0.0f - r) * inv_h), y1 = (int)ceilf((y - 0.0f + r) * inv_h);
    if (x0 < 0) x0 = 0;
    if (x0 >= PLF_GRID_COLS) return 0;
    if (x1 > PLF_GRID_COLS - 1) x1 = PLF_GRID_COLS - 1;
    if (x1 < 0) return 0;
    if (y0 < 0) y0 = 0;
    if (y0 >= PLF_GRID_ROWS) return 0;
    if (y1 > PLF_GRID_ROWS - 1) y1 = PLF_GRID_ROWS - 1;
    if (y1 < 0) return 0;
    const int check = (min_level > 0) || (max_level >= 0);
    int n = 0;
    for (int ix = x0; ix <= x1; ++ix)
        for (int iy = y0; iy <= y1; ++iy) {
            const int c = ix * PLF_GRID_ROWS + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; ++j) {
                const plf_keypoint* k = kps + cell_idx[j];
                if (check) {
                    if (k->octave < min_level) continue;
                    if (max_level >= 0 && k->octave > max_level) continue;
                }
                if (fabsf(k->x - x) < r && fabsf(k->y - y) < r) {
                    if (n < cap) out[n] = cell_idx[j];
                    ++n;
                }
            }
        }
    return n;
}

/* The same lookup as an exported symbol (for bindings that cannot use a C inline). */
PLF_API int PLF_FN(get_features_in_area)(const plf_keypoint* kps, const int32_t* cell_start, const int32_t* cell_idx,
                                         int width, int height, float x, float y, float r, int min_level,
                                         int max_level, int32_t* out, int cap);

/* The other half of transform(): BowVector (TF_IDF: addWeight in feature order, then L1 normalisation in word order,
 * BowVector.cpp:34-84) and FeatureVector (features grouped by node id, ascending feature index) from the per-feature
 * outputs of plf_bow_transform.  Features with weight <= 0 are skipped ("stopped" words).  Pure host inline.
 * bow_word / bow_value: >= n entries; fv_node / fv_start (>= n + 1) / fv_feat: >= n entries.  Returns the number of
 * words; *n_nodes_out = number of FeatureVector nodes. */
static inline int plf_bow_build(const int32_t* word_id, const double* weight, const int32_t* node_id, int n,
                                int32_t* bow_word, double* bow_value, int32_t* fv_node, int32_t* fv_start,
                                int32_t* fv_feat, int* n_nodes_out) {
    int nw = 0, nn = 0, i, j;
    /* insertion into sorted arrays keeps std::map order; n is ~1000 */
    for (i = 0; i < n; ++i) {
        if (!(weight[i] > 0)) continue;
        for (j = 0; j < nw && bow_word[j] < word_id[i]; ++j) {}
        if (j < nw && bow_word[j] == word_id[i]) bow_value[j] += weight[i];
        else {
            int m;
            for (m = nw; m > j; --m) { bow_word[m] = bow_word[m - 1]; bow_value[m] = bow_value[m - 1]; }
            bow_word[j] = word_id[i]; bow_value[j] = weight[i]; ++nw;
        }
    }
    {
        double norm = 0.0;
        for (j = 0; j < nw; ++j) norm += fabs(bow_value[j]);
        if (norm > 0.0) for (j = 0; j < nw; ++j) bow_value[j] /= norm;
    }
    /* FeatureVector: distinct node ids ascending, then the features of each node in ascending index */
    for (i = 0; i < n; ++i) {
        if (!(weight[i] > 0)) continue;
        for (j = 0; j < nn && fv_node[j] < node_id[i]; ++j) {}
        if (!(j < nn && fv_node[j] == node_id[i])) {
            int m;
            for (m = nn; m > j; --m) fv_node[m] = fv_node[m - 1];
            fv_node[j] = node_id[i]; ++nn;
        }
    }
    {
        int pos = 0;
        for (j = 0; j < nn; ++j) {
            fv_start[j] = pos;
            for (i = 0; i < n; ++i) if (weight[i] > 0 && node_id[i] == fv_node[j]) fv_feat[pos++] = i;
        }
        fv_start[nn] = pos;
    }
    if (n_nodes_out) *n_nodes_out = nn;
    return nw;
}
/* The same as an exported symbol (for bindings that cannot use a C inline). */
PLF_API int PLF_FN(bow_build_vectors)(const int32_t* word_id, const double* weight, const int32_t* node_id, int n,
                              int32_t* bow_word, double* bow_value, int32_t* fv_node, int32_t* fv_start,
                              int32_t* fv_feat, int* n_nodes_out);

/* Replaces: static int ORBmatcher::DescriptorDistance(const Mat&, const Mat&) (include/ORBmatcher.h:42,
 * src/ORBmatcher.cc:2495-2511) and int distance(const Mat&, const Mat&) (src/LineMatcher.cpp:231-247):
 * Hamming distance of two 32-byte rows.  Pure host inline (the reference calls it from several threads). */
static inline int plf_hamming256(const uint8_t* a, const uint8_t* b) {
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t x, y;
        __builtin_memcpy(&x, a + 8 * i, 8);
        __builtin_memcpy(&y, b + 8 * i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

/* Layout contracts of the PODs that cross the boundary (cv::KeyPoint 28 B, KeyLine 68 B, query records 56 / 72 B). */
#if defined(__cplusplus)
static_assert(sizeof(plf_keypoint) == 28, "plf_keypoint must be layout-identical to cv::KeyPoint");
static_assert(sizeof(plf_keyline) == 68, "plf_keyline must be layout-identical to cv::line_descriptor::KeyLine");
static_assert(sizeof(plf_proj_query) == 56, "plf_proj_query is 56 bytes");
static_assert(sizeof(plf_frame_query) == 72, "plf_frame_query is 72 bytes");
static_assert(sizeof(plf_track_line) == 24, "plf_track_line is 24 bytes");
#elif defined(__STDC_VERSION__) && __STDC_VERSION__ >= 201112L
_Static_assert(sizeof(plf_keypoint) == 28, "plf_keypoint must be layout-identical to cv::KeyPoint");
_Static_assert(sizeof(plf_keyline) == 68, "plf_keyline must be layout-identical to cv::line_descriptor::KeyLine");
_Static_assert(sizeof(plf_proj_query) == 56, "plf_proj_query is 56 bytes");
_Static_assert(sizeof(plf_frame_query) == 72, "plf_frame_query is 72 bytes");
_Static_assert(sizeof(plf_track_line) == 24, "plf_track_line is 24 bytes");
#endif

#ifdef __cplusplus
}
#endif
#endif /* PLF_B200_H */
