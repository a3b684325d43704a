"""ctypes binding of the C ABI declared in include/plf_b200.h.

The same declarations serve two libraries: the product `libplf_b200.so` (prefix ``plf_``, hand-written sm_100a
CUDA) and the test-only CPU oracle `oracle/libplf_oracle.so` (prefix ``plf_cpu_``).  This module never falls back
from one to the other: `load_product()` raises if the CUDA library is missing.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.environ.get("PLF_PRODUCT_LIB") or os.path.join(ROOT, "pli-slam_b200", "libplf_b200.so")      # (the override is a developer switch for A/B builds)
ORACLE_LIB = os.path.join(ROOT, "oracle", "libplf_oracle.so")

PLF_OK = 0
STATUS = {0: "OK", 1: "INVALID", 2: "EMPTY_IMAGE", 3: "CUDA", 4: "NO_DEVICE", 5: "UNSUPPORTED", 6: "SIZE_MISMATCH",
          7: "STATE"}


class PlfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("plf status %d (%s): %s" % (code, STATUS.get(code, "?"), msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("max_batch", C.c_int32),
        ("n_features", C.c_int32), ("scale_factor", C.c_float), ("n_levels", C.c_int32),
        ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32),
        ("has_lines", C.c_int32), ("lsd_nfeatures", C.c_int32), ("lsd_refine", C.c_int32), ("lsd_n_bins", C.c_int32),
        ("min_line_length", C.c_double), ("lsd_scale", C.c_double), ("lsd_sigma_scale", C.c_double),
        ("lsd_quant", C.c_double), ("lsd_ang_th", C.c_double), ("lsd_log_eps", C.c_double),
        ("lsd_density_th", C.c_double),
        ("bf", C.c_float), ("fx", C.c_float),
        ("best_lr_matches", C.c_int32), ("matching_s_ws", C.c_int32),
        ("min_ratio_12_l", C.c_double), ("line_sim_th", C.c_double), ("min_disp", C.c_double),
        ("line_horiz_th", C.c_double), ("stereo_overlap_th", C.c_double), ("ls_min_disp_ratio", C.c_double),
        ("has_points", C.c_int32),
    ]


class FrameOut(C.Structure):
    _fields_ = [
        ("kp_cap", C.c_int32), ("kl_cap", C.c_int32),
        ("n_kp_left", C.c_void_p), ("n_kp_right", C.c_void_p), ("n_kl_left", C.c_void_p), ("n_kl_right", C.c_void_p),
        ("kp_left", C.c_void_p), ("kp_right", C.c_void_p), ("desc_left", C.c_void_p), ("desc_right", C.c_void_p),
        ("u_right", C.c_void_p), ("depth", C.c_void_p),
        ("kl_left", C.c_void_p), ("kl_right", C.c_void_p), ("ldesc_left", C.c_void_p), ("ldesc_right", C.c_void_p),
        ("disp_se", C.c_void_p), ("le", C.c_void_p), ("line_match12", C.c_void_p),
    ]


KEYPOINT_DT = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                        ("octave", "i4"), ("class_id", "i4")])
KEYLINE_DT = np.dtype([("angle", "f4"), ("class_id", "i4"), ("octave", "i4"), ("pt_x", "f4"), ("pt_y", "f4"),
                       ("response", "f4"), ("size", "f4"), ("startPointX", "f4"), ("startPointY", "f4"),
                       ("endPointX", "f4"), ("endPointY", "f4"), ("sPointInOctaveX", "f4"), ("sPointInOctaveY", "f4"),
                       ("ePointInOctaveX", "f4"), ("ePointInOctaveY", "f4"), ("lineLength", "f4"),
                       ("numOfPixels", "i4")])
assert KEYPOINT_DT.itemsize == 28 and KEYLINE_DT.itemsize == 68

# every symbol include/plf_b200.h declares (without prefix)
PROJ_QUERY_DT = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                          ("skip", "<i4"), ("desc", "u1", (32,))])       # plf_proj_query, 56 bytes
FRAME_QUERY_DT = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"),
                           ("skip", "<i4"), ("has_observations", "<i4"), ("angle", "<f4"), ("desc", "u1", (32,)), ("pad", "<i4")])
TRACK_LINE_DT = np.dtype([("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"), ("angle", "<f4"), ("eligible", "<i4")])   # plf_track_line
assert PROJ_QUERY_DT.itemsize == 56 and FRAME_QUERY_DT.itemsize == 72 and TRACK_LINE_DT.itemsize == 24
GRID_COLS, GRID_ROWS = 64, 48      # FRAME_GRID_COLS / FRAME_GRID_ROWS (include/Frame.h:59-60)

ABI_SYMBOLS = [
    "default_params", "create", "destroy", "last_error", "keypoint_capacity", "keyline_capacity", "get_scale_tables",
    "orb_extract", "get_pyramid_level", "line_extract", "stereo_match_points", "stereo_match_lines", "match_nnr",
    "match", "frontend_batch", "batch_upload", "batch_run", "batch_download", "sync", "batch_io_bytes",
    "last_launch_count", "set_grower_policy", "set_stage_timing", "get_stage_ms", "stream", "tap_blurred_level", "tap_pyramid_level",
    "tap_fast_candidates", "tap_lsd_scaled", "tap_lsd_angles", "tap_lsd_segments", "tap_lbd_float", "tap_grow_ns",
    "rectify_set_maps", "rectify", "batch_upload_raw", "feature_grid", "get_features_in_area", "backproject", "bow_set_vocabulary", "bow_transform", "bow_build_vectors", "search_by_projection", "search_by_projection_frame",
    "search_by_projection_reloc", "search_by_projection_loop", "search_by_bow", "match_lines_tracked",
]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _image(img):
    """uint8 [H, W] whose rows may be padded (strides (stride, 1)): passed through with its row stride, like a cv::Mat
    ROI; anything else is copied to a dense array."""
    if not (isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
            and img.strides[0] >= img.shape[1]):
        img = np.ascontiguousarray(img, dtype=np.uint8)
    return img


def _image_batch(a):
    """uint8 [B, H, W] with padded rows allowed (strides (H * stride, stride, 1)), as the ABI's [batch][h][stride]."""
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 3 and a.strides[2] == 1
            and a.strides[1] >= a.shape[2] and a.strides[0] == a.shape[1] * a.strides[1]):
        a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class Library:
    """One loaded ABI library (product or oracle)."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s is not built (run `python -c 'import __graft_entry__ as g; g.build()'`)" % path)
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        for s in ABI_SYMBOLS:
            getattr(self.dll, prefix + s)   # raises AttributeError if a declared symbol is not exported
        self.fn("last_error").restype = C.c_char_p
        self.fn("stream").restype = C.c_void_p

    def fn(self, name):
        return getattr(self.dll, self.prefix + name)

    def check(self, rc):
        if rc != PLF_OK:
            raise PlfError(rc, (self.fn("last_error")() or b"").decode())

    def default_params(self, **kw):
        p = Params()
        self.check(self.fn("default_params")(C.byref(p)))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p


_product = None
_oracle = None


def load_product():
    """The CUDA library.  Fails loudly when it is not built; there is no CPU fallback."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_LIB, "plf_")
    return _product


def load_oracle():
    """TEST INFRASTRUCTURE ONLY: the CPU oracle.  Callers: tests/, __graft_entry__.smoke(), bench.py baseline legs."""
    global _oracle
    if _oracle is None:
        _oracle = Library(ORACLE_LIB, "plf_cpu_")
        _oracle.dll.plf_cpu_fast_atan2.restype = C.c_float
        _oracle.dll.plf_cpu_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _oracle


class BatchResult:
    """Host arrays of one batched call; mirrors the Frame members named in plf_frame_out."""

    def __init__(self, batch, kp_cap, kl_cap, pinned=False):
        self.batch, self.kp_cap, self.kl_cap = batch, kp_cap, kl_cap
        self._keep = []

        def alloc(shape, dt):
            if pinned:
                import torch
                n = int(np.prod(shape)) * np.dtype(dt).itemsize
                t = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
                self._keep.append(t)
                return t.numpy()[:n].view(dt).reshape(shape)
            return np.zeros(shape, dt)

        self.n_kp_left = alloc((batch,), "i4"); self.n_kp_right = alloc((batch,), "i4")
        self.n_kl_left = alloc((batch,), "i4"); self.n_kl_right = alloc((batch,), "i4")
        self.kp_left = alloc((batch, kp_cap), KEYPOINT_DT); self.kp_right = alloc((batch, kp_cap), KEYPOINT_DT)
        self.desc_left = alloc((batch, kp_cap, 32), "u1"); self.desc_right = alloc((batch, kp_cap, 32), "u1")
        self.u_right = alloc((batch, kp_cap), "f4"); self.depth = alloc((batch, kp_cap), "f4")
        self.kl_left = alloc((batch, kl_cap), KEYLINE_DT); self.kl_right = alloc((batch, kl_cap), KEYLINE_DT)
        self.ldesc_left = alloc((batch, kl_cap, 32), "u1"); self.ldesc_right = alloc((batch, kl_cap, 32), "u1")
        self.disp_se = alloc((batch, kl_cap, 2), "f4"); self.le = alloc((batch, kl_cap, 3), "f8")
        self.line_match12 = alloc((batch, kl_cap), "i4")
        o = FrameOut()
        o.kp_cap, o.kl_cap = kp_cap, kl_cap
        for name, _ in FrameOut._fields_[2:]:
            setattr(o, name, getattr(self, name).ctypes.data)
        self.c = o

    def nbytes(self):
        return sum(getattr(self, n).nbytes for n, _ in FrameOut._fields_[2:])


class Frontend:
    """Host-side handle over one plf_ctx: the two ORBextractor + two Lineextractor objects of a Tracking instance
    (src/Tracking.cc:87-98,743-749) plus the stereo matchers of Frame (src/Frame.cc:160-163)."""

    def __init__(self, lib, params=None, device=0, **kw):
        self.lib = lib
        self.params = params if params is not None else lib.default_params(**kw)
        self.ctx = C.c_void_p()
        lib.check(lib.fn("create")(C.byref(self.params), int(device), C.byref(self.ctx)))
        self.kp_cap = lib.fn("keypoint_capacity")(self.ctx)
        self.kl_cap = lib.fn("keyline_capacity")(self.ctx)
        self.W, self.H = self.params.width, self.params.height

    def close(self):
        if self.ctx:
            self.lib.fn("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- reference-shaped single-frame calls ---------------------------------------------------------------
    def orb_extract(self, side, img, lapping=(0, 0)):
        """ORBextractor::operator()(im, mask, kps, desc, vLappingArea) -> (monoIndex, keypoints, descriptors)."""
        img = _image(img)
        h, w = img.shape
        kps = np.zeros(self.kp_cap, KEYPOINT_DT)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        self.lib.check(self.lib.fn("orb_extract")(self.ctx, side, _ptr(img), w, h, img.strides[0], int(lapping[0]),
                                                  int(lapping[1]), _ptr(kps), _ptr(desc), self.kp_cap, C.byref(n),
                                                  C.byref(mono)))
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def line_extract(self, side, img):
        """Lineextractor::operator()(im, mask, keylines, desc) -> (keylines, descriptors)."""
        img = _image(img)
        h, w = img.shape
        kls = np.zeros(self.kl_cap, KEYLINE_DT)
        desc = np.zeros((self.kl_cap, 32), np.uint8)
        n = C.c_int(0)
        self.lib.check(self.lib.fn("line_extract")(self.ctx, side, _ptr(img), w, h, img.strides[0], _ptr(kls),
                                                   _ptr(desc), self.kl_cap, C.byref(n)))
        return kls[:n.value].copy(), desc[:n.value].copy()

    def stereo_match_points(self, n):
        """Frame::ComputeStereoMatches -> (mvuRight, mvDepth)."""
        u = np.zeros(self.kp_cap, np.float32)
        d = np.zeros(self.kp_cap, np.float32)
        self.lib.check(self.lib.fn("stereo_match_points")(self.ctx, _ptr(u), _ptr(d), self.kp_cap))
        return u[:n].copy(), d[:n].copy()

    def stereo_match_lines(self, n):
        """Frame::ComputeStereoMatches_Lines -> (mvDisparity_l, mvle_l, matches_12)."""
        disp = np.zeros((self.kl_cap, 2), np.float32)
        le = np.zeros((self.kl_cap, 3), np.float64)
        m12 = np.zeros(self.kl_cap, np.int32)
        self.lib.check(self.lib.fn("stereo_match_lines")(self.ctx, _ptr(disp), _ptr(le), _ptr(m12), self.kl_cap))
        return disp[:n].copy(), le[:n].copy(), m12[:n].copy()

    def match_nnr(self, d1, d2, nnr):
        """matchNNR(desc1, desc2, nnr, matches_12) -> (count, matches_12)."""
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        m = np.full(max(len(d1), 1), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("match_nnr")(self.ctx, _ptr(d1), len(d1), _ptr(d2), len(d2), C.c_float(nnr),
                                                _ptr(m), C.byref(nm)))
        return nm.value, m[:len(d1)]

    def match(self, d1, d2, nnr, best_lr=True):
        """match(desc1, desc2, nnr, matches_12) with Config::bestLRMatches() = best_lr."""
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        m = np.full(max(len(d1), 1), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("match")(self.ctx, _ptr(d1), len(d1), _ptr(d2), len(d2), C.c_float(nnr),
                                            int(bool(best_lr)), _ptr(m), C.byref(nm)))
        return nm.value, m[:len(d1)]

    def pyramid_level(self, side, level, slot=0):
        w, h = C.c_int(0), C.c_int(0)
        self.lib.check(self.lib.fn("tap_pyramid_level")(self.ctx, slot, side, level, None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self.lib.check(self.lib.fn("tap_pyramid_level")(self.ctx, slot, side, level, _ptr(out), w.value, C.byref(w),
                                                        C.byref(h)))
        return out

    def blurred_level(self, side, level, slot=0):
        w, h = C.c_int(0), C.c_int(0)
        self.lib.check(self.lib.fn("tap_pyramid_level")(self.ctx, slot, side, level, None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self.lib.check(self.lib.fn("tap_blurred_level")(self.ctx, slot, side, level, _ptr(out), w.value))
        return out

    def fast_candidates(self, side, level, slot=0, cap=200000):
        out = np.zeros((cap, 3), np.float32)
        n = C.c_int(0)
        self.lib.check(self.lib.fn("tap_fast_candidates")(self.ctx, slot, side, level, _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def lsd_scaled(self, side, slot=0):
        w, h = C.c_int(0), C.c_int(0)
        self.lib.check(self.lib.fn("tap_lsd_scaled")(self.ctx, slot, side, None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self.lib.check(self.lib.fn("tap_lsd_scaled")(self.ctx, slot, side, _ptr(out), w.value, C.byref(w), C.byref(h)))
        return out

    def lsd_angles(self, side, slot=0):
        w, h = C.c_int(0), C.c_int(0)
        self.lib.check(self.lib.fn("tap_lsd_angles")(self.ctx, slot, side, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.float32)
        self.lib.check(self.lib.fn("tap_lsd_angles")(self.ctx, slot, side, _ptr(out), C.byref(w), C.byref(h)))
        return out

    def lsd_segments(self, side, slot=0, cap=20000):
        out = np.zeros((cap, 4), np.float32)
        n = C.c_int(0)
        self.lib.check(self.lib.fn("tap_lsd_segments")(self.ctx, slot, side, _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def lbd_float(self, side, slot=0):
        out = np.zeros((self.kl_cap, 72), np.float32)
        n = C.c_int(0)
        self.lib.check(self.lib.fn("tap_lbd_float")(self.ctx, slot, side, _ptr(out), self.kl_cap, C.byref(n)))
        return out[:n.value].copy()

    def scale_tables(self):
        L = self.params.n_levels
        a = [np.zeros(L, np.float32) for _ in range(4)]
        n = np.zeros(L, np.int32)
        self.lib.check(self.lib.fn("get_scale_tables")(self.ctx, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]),
                                                       _ptr(n)))
        return a[0], a[1], a[2], a[3], n

    # --- batched calls --------------------------------------------------------------------------------------
    def new_result(self, batch, pinned=False):
        return BatchResult(batch, self.kp_cap, self.kl_cap, pinned)

    def frontend_batch(self, left, right, out=None):
        """Frame::Frame(stereo) extraction + matching for left/right arrays of shape [batch, H, W]."""
        left, right = _image_batch(left), _image_batch(right)
        if left.strides != right.strides:
            left, right = np.ascontiguousarray(left), np.ascontiguousarray(right)
        b = left.shape[0]
        out = out or self.new_result(b)
        self.lib.check(self.lib.fn("frontend_batch")(self.ctx, _ptr(left), _ptr(right), b, left.strides[1],
                                                     C.byref(out.c)))
        return out

    def batch_upload(self, left, right):
        self.lib.check(self.lib.fn("batch_upload")(self.ctx, _ptr(left), _ptr(right), left.shape[0],
                                                   left.strides[1]))

    def batch_upload_ptr(self, left_ptr, right_ptr, batch, stride):
        """batch_upload from raw host addresses (pinned torch tensors in bench.py)."""
        self.lib.check(self.lib.fn("batch_upload")(self.ctx, C.c_void_p(left_ptr), C.c_void_p(right_ptr), batch,
                                                   stride))

    # ---- rectification (the cv::remap in front of the path) -------------------------------------
    def rectify_set_maps(self, side, map_x, map_y, src_w=None, src_h=None):
        """Maps of cv::initUndistortRectifyMap(..., CV_32F) for one camera: float32 arrays [H, W]."""
        map_x = np.ascontiguousarray(map_x, np.float32)
        map_y = np.ascontiguousarray(map_y, np.float32)
        assert map_x.shape == map_y.shape == (self.params.height, self.params.width)
        self.lib.check(self.lib.fn("rectify_set_maps")(self.ctx, side, _ptr(map_x), _ptr(map_y),
                                                       int(src_w or self.params.width), int(src_h or self.params.height)))

    def rectify(self, side, raw):
        """cv::remap(raw, M1, M2, INTER_LINEAR) -> rectified uint8 [H, W]."""
        raw = np.ascontiguousarray(raw, np.uint8)
        out = np.zeros((self.params.height, self.params.width), np.uint8)
        self.lib.check(self.lib.fn("rectify")(self.ctx, side, _ptr(raw), raw.strides[0], _ptr(out), out.strides[0]))
        return out

    # ---- Frame::AssignFeaturesToGrid / GetFeaturesInArea (first half) -----------------------------
    def feature_grid(self, first_slot=0, n_slots=1):
        """CSR of Frame::mGrid for the left keypoints of the slots: (cell_start [n, 64*48+1], cell_idx [n, kp_cap])."""
        st = np.zeros((n_slots, GRID_COLS * GRID_ROWS + 1), np.int32)
        ix = np.full((n_slots, self.kp_cap), -1, np.int32)
        self.lib.check(self.lib.fn("feature_grid")(self.ctx, first_slot, n_slots, _ptr(st), _ptr(ix), self.kp_cap))
        return st, ix

    def search_by_projection(self, queries, occupied, th=1.0, nn_ratio=0.8, th_high=100, slot=0):
        """ORBmatcher::SearchByProjection(F, vpMapPoints, th): queries = PROJ_QUERY_DT array, occupied = uint8 per keypoint
        (updated in place) -> (match index per query, nmatches)."""
        queries = np.ascontiguousarray(queries, PROJ_QUERY_DT)
        assert occupied.dtype == np.uint8 and occupied.flags.c_contiguous
        match = np.full(len(queries), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("search_by_projection")(self.ctx, slot, _ptr(queries), len(queries), C.c_float(th),
                                                           C.c_float(nn_ratio), int(th_high), _ptr(occupied), len(occupied),
                                                           _ptr(match), C.byref(nm)))
        return match, nm.value

    def search_by_projection_frame(self, queries, occupied, th_high=100, check_orientation=True, slot=0):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, ...) from the projected points on: queries =
        FRAME_QUERY_DT array, occupied = uint8 per keypoint (updated in place) -> (feat_query, match12, nmatches)."""
        queries = np.ascontiguousarray(queries, FRAME_QUERY_DT)
        assert occupied.dtype == np.uint8 and occupied.flags.c_contiguous
        fq = np.full(len(occupied), -1, np.int32); m12 = np.full(len(occupied), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("search_by_projection_frame")(self.ctx, slot, _ptr(queries), len(queries), int(th_high),
                                                                 int(bool(check_orientation)), _ptr(occupied), len(occupied),
                                                                 _ptr(fq), _ptr(m12), C.byref(nm)))
        return fq, m12, nm.value

    def search_by_projection_reloc(self, queries, occupied, orb_dist=100, check_orientation=True, slot=0):
        """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) from the projected points on
        -> (feat_query, nmatches); occupied (uint8 per keypoint) is updated in place."""
        queries = np.ascontiguousarray(queries, FRAME_QUERY_DT)
        assert occupied.dtype == np.uint8 and occupied.flags.c_contiguous
        fq = np.full(len(occupied), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("search_by_projection_reloc")(self.ctx, slot, _ptr(queries), len(queries), int(orb_dist),
                                                                 int(bool(check_orientation)), _ptr(occupied), len(occupied),
                                                                 _ptr(fq), C.byref(nm)))
        return fq, nm.value

    def search_by_projection_loop(self, queries, occupied, th_low=50, ratio_hamming=1.0, slot=0):
        """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) from the projected points on
        -> (feat_query, nmatches); occupied = vpMatched != NULL, updated in place."""
        queries = np.ascontiguousarray(queries, FRAME_QUERY_DT)
        assert occupied.dtype == np.uint8 and occupied.flags.c_contiguous
        fq = np.full(len(occupied), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("search_by_projection_loop")(self.ctx, slot, _ptr(queries), len(queries), int(th_low),
                                                                C.c_float(ratio_hamming), _ptr(occupied), len(occupied),
                                                                _ptr(fq), C.byref(nm)))
        return fq, nm.value

    def search_by_bow(self, kf_desc, kf_angle, kf_node, kf_valid, f_node, th_low=50, nn_ratio=0.7, check_orientation=True, slot=0):
        """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) -> (match [len(f_node)]: keyframe feature per frame feature
        or -1, nmatches)."""
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 32)
        kf_angle = np.ascontiguousarray(kf_angle, np.float32); kf_node = np.ascontiguousarray(kf_node, np.int32)
        kf_valid = np.ascontiguousarray(kf_valid, np.uint8); f_node = np.ascontiguousarray(f_node, np.int32)
        m = np.full(max(len(f_node), 1), -1, np.int32)
        nm = C.c_int(0)
        self.lib.check(self.lib.fn("search_by_bow")(self.ctx, slot, _ptr(kf_desc), _ptr(kf_angle), _ptr(kf_node), _ptr(kf_valid),
                                                    len(kf_desc), _ptr(f_node), len(f_node), int(th_low), C.c_float(nn_ratio),
                                                    int(bool(check_orientation)), _ptr(m), C.byref(nm)))
        return m[:len(f_node)], nm.value

    def match_lines_tracked(self, mode, desc1, lines1, desc2, kl2, disp2, held2, nnr, bounds):
        """match() + the tracking gates (mode 0: Tracking::TrackWithMotionModel, 1: SearchLocalLines) ->
        (matches12, assign12, n_assigned).  bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        desc1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32); desc2 = np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
        lines1 = np.ascontiguousarray(lines1, TRACK_LINE_DT); kl2 = np.ascontiguousarray(kl2, KEYLINE_DT)
        disp2 = np.ascontiguousarray(disp2, np.float32).reshape(-1, 2)
        held2 = None if held2 is None else np.ascontiguousarray(held2, np.uint8)
        n1, n2 = len(desc1), len(desc2)
        assert len(lines1) == n1 and len(kl2) == n2 and len(disp2) == n2
        m = np.full(max(n1, 1), -1, np.int32); a = np.full(max(n1, 1), -1, np.int32)
        na = C.c_int(0)
        self.lib.check(self.lib.fn("match_lines_tracked")(self.ctx, int(mode), _ptr(desc1), _ptr(lines1), n1, _ptr(desc2), _ptr(kl2),
                                                          _ptr(disp2), _ptr(held2), n2, C.c_float(nnr), C.c_float(bounds[0]),
                                                          C.c_float(bounds[1]), C.c_float(bounds[2]), C.c_float(bounds[3]),
                                                          _ptr(m), _ptr(a), C.byref(na)))
        return m[:n1], a[:n1], na.value

    # ---- bag of words ------------------------------------------------------------------------------
    def bow_set_vocabulary(self, which, voc):
        """voc: dict(levels, child_first, child_count, child, desc [n, 32] uint8, word_id, weight) — the DBoW2 tree, flat."""
        cf = np.ascontiguousarray(voc["child_first"], np.int32); cc = np.ascontiguousarray(voc["child_count"], np.int32)
        ch = np.ascontiguousarray(voc["child"], np.int32); de = np.ascontiguousarray(voc["desc"], np.uint8)
        wi = np.ascontiguousarray(voc["word_id"], np.int32); we = np.ascontiguousarray(voc["weight"], np.float64)
        self.lib.check(self.lib.fn("bow_set_vocabulary")(self.ctx, which, len(cf), int(voc["levels"]), _ptr(cf), _ptr(cc), _ptr(ch),
                                                         _ptr(de), _ptr(wi), _ptr(we)))

    def bow_transform(self, which, n_slots=1, first_slot=0, levelsup=4):
        """Per-feature word id, weight and node id (level L - levelsup) for the left ORB (0) / LBD (1) descriptors."""
        cap = self.kl_cap if which else self.kp_cap
        w = np.zeros((n_slots, cap), np.int32); v = np.zeros((n_slots, cap), np.float64); nd = np.zeros((n_slots, cap), np.int32)
        self.lib.check(self.lib.fn("bow_transform")(self.ctx, which, first_slot, n_slots, levelsup, _ptr(w), _ptr(v), _ptr(nd), cap))
        return w, v, nd

    def bow_build(self, word_id, weight, node_id):
        """BowVector (sorted words, L1-normalised TF-IDF values) and FeatureVector (node -> feature indices)."""
        n = len(word_id)
        word_id = np.ascontiguousarray(word_id, np.int32); weight = np.ascontiguousarray(weight, np.float64)
        node_id = np.ascontiguousarray(node_id, np.int32)
        bw = np.zeros(n + 1, np.int32); bv = np.zeros(n + 1, np.float64)
        fn_ = np.zeros(n + 1, np.int32); fs = np.zeros(n + 2, np.int32); ff = np.zeros(n + 1, np.int32)
        nn = C.c_int(0)
        nw = self.lib.fn("bow_build_vectors")(_ptr(word_id), _ptr(weight), _ptr(node_id), n, _ptr(bw), _ptr(bv), _ptr(fn_), _ptr(fs),
                                              _ptr(ff), C.byref(nn))
        fv = {int(fn_[j]): ff[fs[j]:fs[j + 1]].tolist() for j in range(nn.value)}
        return bw[:nw].copy(), bv[:nw].copy(), fv

    def backproject(self, Rwc, Ow, fy, cx, cy, first_slot=0, lines=True):
        """Frame::UnprojectStereo for every left keypoint and Frame::backProjection for every line end point of the slots.
        Rwc [n, 3, 3] float32, Ow [n, 3] float32 -> (x3d [n, kp_cap, 3] float32, l3d [n, kl_cap, 6] float64 or None)."""
        Rwc = np.ascontiguousarray(Rwc, np.float32).reshape(-1, 9)
        Ow = np.ascontiguousarray(Ow, np.float32).reshape(-1, 3)
        n = Rwc.shape[0]
        x3d = np.zeros((n, self.kp_cap, 3), np.float32)
        l3d = np.zeros((n, self.kl_cap, 6), np.float64) if lines else None
        self.lib.check(self.lib.fn("backproject")(self.ctx, first_slot, n, _ptr(Rwc), _ptr(Ow), C.c_float(fy), C.c_float(cx),
                                                  C.c_float(cy), _ptr(x3d), self.kp_cap, _ptr(l3d), self.kl_cap if lines else 0))
        return x3d, l3d

    def features_in_area(self, kps, cell_start, cell_idx, x, y, r, min_level=-1, max_level=-1):
        """Frame::GetFeaturesInArea on the CSR grid of one slot -> int32 indices."""
        out = np.zeros(len(kps) + 1, np.int32)
        fn = self.lib.fn("get_features_in_area")
        n = fn(_ptr(np.ascontiguousarray(kps)), _ptr(cell_start), _ptr(cell_idx), self.W, self.H, C.c_float(x), C.c_float(y),
               C.c_float(r), int(min_level), int(max_level), _ptr(out), len(out))
        return out[:n].copy()

    def batch_upload_raw_ptr(self, left_ptr, right_ptr, batch, stride):
        self.lib.check(self.lib.fn("batch_upload_raw")(self.ctx, C.c_void_p(left_ptr), C.c_void_p(right_ptr), batch,
                                                       stride))

    def batch_upload_raw(self, left_raw, right_raw):
        """Upload RAW frames [batch, src_h, src_w]; they are rectified on the way into the batch slots."""
        left_raw = np.ascontiguousarray(left_raw, np.uint8)
        right_raw = np.ascontiguousarray(right_raw, np.uint8)
        self.lib.check(self.lib.fn("batch_upload_raw")(self.ctx, _ptr(left_raw), _ptr(right_raw), left_raw.shape[0],
                                                       left_raw.strides[1]))

    def batch_run(self, batch):
        self.lib.check(self.lib.fn("batch_run")(self.ctx, batch))

    def batch_download(self, batch, out):
        self.lib.check(self.lib.fn("batch_download")(self.ctx, batch, C.byref(out.c)))

    def sync(self):
        self.lib.check(self.lib.fn("sync")(self.ctx))

    def io_bytes(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self.lib.check(self.lib.fn("batch_io_bytes")(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return self.lib.fn("last_launch_count")(self.ctx)

    def set_grower_policy(self, policy):
        """0: automatic (streaming multi-warp grower for launches of <= 296 images), 1: throughput (one warp per image always)."""
        self.lib.check(self.lib.fn("set_grower_policy")(self.ctx, int(policy)))

    def set_stage_timing(self, on):
        self.lib.check(self.lib.fn("set_stage_timing")(self.ctx, int(on)))

    def stage_ms(self):
        names = C.POINTER(C.c_char_p)()
        ms = C.POINTER(C.c_float)()
        n = C.c_int(0)
        self.lib.check(self.lib.fn("get_stage_ms")(self.ctx, C.byref(names), C.byref(ms), C.byref(n)))
        return {names[i].decode(): ms[i] for i in range(n.value)}

    def grow_ns(self, n_images):
        """ns per image in the one-warp-per-image region grower during the last stage-timed pass."""
        out = np.zeros(n_images, np.uint64)
        self.lib.check(self.lib.fn("tap_grow_ns")(self.ctx, _ptr(out), n_images))
        return out

    def stream(self):
        return self.lib.fn("stream")(self.ctx)
