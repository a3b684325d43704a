"""Pins the CPU oracle (Tier B, oracle/cpp) against real OpenCV (cv2 4.13) primitive by primitive and stage by stage.

The reference ships no tests for this path, and its arithmetic lives in OpenCV, which is not under
/root/reference; these known-answer checks are what anchors the oracle.  CPU only.
"""
import ctypes as C
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import tier_a  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _rand_img(rng, w, h):
    base = rng.integers(0, 256, size=(h // 8 + 2, w // 8 + 2)).astype(np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(img.astype(np.int32) + rng.integers(-20, 21, size=img.shape), 0, 255).astype(np.uint8)
    return img


def test_fast_atan2_bit_exact(oracle):
    rng = np.random.default_rng(0)
    ys = np.concatenate([rng.normal(0, 1000, 20000), rng.integers(-5000, 5000, 2000), [0, 0, 1, -1, 0]]).astype(np.float32)
    xs = np.concatenate([rng.normal(0, 1000, 20000), rng.integers(-5000, 5000, 2000), [0, 1, 0, 0, -1]]).astype(np.float32)
    bad = 0
    for y, x in zip(ys, xs):
        a = oracle.dll.plf_cpu_fast_atan2(float(y), float(x))
        b = cv2.fastAtan2(float(y), float(x))
        bad += np.float32(a) != np.float32(b)
    assert bad == 0


@pytest.mark.parametrize("w,h", [(752, 480), (1280, 720), (640, 480), (333, 257)])
def test_resize_linear_chain(oracle, w, h):
    rng = np.random.default_rng(w)
    img = _rand_img(rng, w, h)
    ref = tier_a.pyramid(img)
    cur = img
    for l in range(1, 8):
        lh, lw = ref[l].shape
        dst = np.zeros((lh, lw), np.uint8)
        oracle.dll.plf_cpu_prim_resize_linear(_p(cur), cur.shape[1], cur.shape[0], _p(dst), lw, lh)
        assert np.array_equal(dst, ref[l]), "level %d" % l
        cur = dst


@pytest.mark.parametrize("ksize,sigma,taps", [(7, 2.0, [18, 34, 48, 56, 48, 34, 18]), (5, 1.0, [14, 62, 104, 62, 14]),
                                              (7, 0.6, [0, 1, 42, 170, 42, 1, 0])])
def test_gaussian_blur(oracle, ksize, sigma, taps):
    rng = np.random.default_rng(ksize)
    for (w, h) in [(752, 480), (210, 134), (97, 61)]:
        img = rng.integers(0, 256, size=(h, w)).astype(np.uint8)
        dst = np.zeros_like(img)
        t = np.zeros(ksize, np.int32)
        oracle.dll.plf_cpu_prim_gaussian(_p(img), w, h, ksize, C.c_double(sigma), _p(dst), _p(t))
        assert list(t) == taps
        ref = cv2.GaussianBlur(img, (ksize, ksize), sigma, sigma, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(dst, ref)


@pytest.mark.parametrize("w,h", [(752, 480), (1280, 720), (101, 77)])
def test_resize_linear_exact(oracle, w, h):
    rng = np.random.default_rng(h)
    img = _rand_img(rng, w, h)
    ref = cv2.resize(img, None, fx=1.2, fy=1.2, interpolation=cv2.INTER_LINEAR_EXACT)
    dw, dh = C.c_int(0), C.c_int(0)
    dst = np.zeros(ref.shape, np.uint8)
    oracle.dll.plf_cpu_prim_resize_exact(_p(img), w, h, C.c_double(1.2), _p(dst), C.byref(dw), C.byref(dh))
    assert (dh.value, dw.value) == ref.shape
    assert np.array_equal(dst, ref)


def test_sobel(oracle):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(120, 160)).astype(np.uint8)
    dx = np.zeros(img.shape, np.int16)
    dy = np.zeros(img.shape, np.int16)
    oracle.dll.plf_cpu_prim_sobel(_p(img), 160, 120, _p(dx), _p(dy))
    assert np.array_equal(dx, cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(dy, cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3))


@pytest.mark.parametrize("th", [20, 7])
def test_fast_window(oracle, th):
    rng = np.random.default_rng(th)
    for (w, h) in [(44, 40), (36, 36), (120, 90), (7, 7), (9, 12)]:
        img = _rand_img(rng, max(w, 16), max(h, 16))[:h, :w].copy()
        ref = cv2.FastFeatureDetector_create(th, True).detect(img)
        ref = np.array([(k.pt[0], k.pt[1], k.response) for k in ref], np.float32).reshape(-1, 3)
        out = np.zeros((w * h, 3), np.float32)
        n = C.c_int(0)
        assert oracle.dll.plf_cpu_prim_fast(_p(img), w, h, th, _p(out), w * h, C.byref(n)) == 0
        assert n.value == len(ref)
        assert np.array_equal(out[:n.value], ref)


def test_orb_stages_vs_cv2(plf, oracle, pair1):
    """Pyramid bytes, blurred levels and per-cell FAST candidate lists (order included) of the full extractor."""
    L, _ = pair1
    f = plf.Frontend(oracle)
    mono, kps, desc = f.orb_extract(0, L)
    pyr = tier_a.pyramid(L)
    for l in range(8):
        assert np.array_equal(f.pyramid_level(0, l), pyr[l])
        assert np.array_equal(f.blurred_level(0, l),
                              cv2.GaussianBlur(pyr[l], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        assert np.array_equal(f.fast_candidates(0, l), tier_a.level_candidates(pyr[l]))
    assert mono == len(kps)


def test_orb_octree_orientation_descriptor_vs_python(plf, oracle, pair1):
    """Second, independent restatement (python list emulation) of the quadtree, orientation and rBRIEF."""
    _, R = pair1
    f = plf.Frontend(oracle)
    mono, kps, desc = f.orb_extract(1, R)
    pyr = tier_a.pyramid(R)
    sc, _, _, _, nfeat = f.scale_tables()
    pattern = tier_a.load_pattern(os.path.join(ROOT, "pli-slam_b200", "csrc", "orb_pattern.inc"))
    umax = tier_a.umax_table()
    row = 0
    bad_desc = 0
    for l in range(8):
        cand = tier_a.level_candidates(pyr[l])
        rel = cand.copy()
        rel[:, :2] -= 16
        h, w = pyr[l].shape
        keep = tier_a.distribute_octree(rel, 16, w - 16, 16, h - 16, int(nfeat[l]))
        blur = cv2.GaussianBlur(pyr[l], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        for k in keep:
            x, y, r = cand[k]
            kp = kps[row]
            assert kp["octave"] == l and kp["response"] == r
            ex, ey = (np.float32(x), np.float32(y)) if l == 0 else (np.float32(x) * sc[l], np.float32(y) * sc[l])
            assert kp["x"] == ex and kp["y"] == ey
            ang = tier_a.ic_angle(pyr[l], int(x), int(y), umax)
            assert np.float32(ang) == kp["angle"]
            assert kp["size"] == np.float32(int(np.float32(31) * sc[l]))
            d = tier_a.orb_descriptor(blur, int(x), int(y), ang, pattern)
            bad_desc += not np.array_equal(d, desc[row])
            row += 1
    assert row == len(kps)
    # the python restatement rounds cos/sin from double libm, the oracle uses cosf/sinf: allow a vanishing number of
    # descriptor rows to differ, none expected
    assert bad_desc <= 1


def _match_segments(a, b, tol):
    """fraction of segments of `a` that have a segment in `b` with both endpoints within tol (either direction)."""
    if len(a) == 0:
        return 1.0
    hit = 0
    for s in a:
        d1 = np.maximum(np.hypot(b[:, 0] - s[0], b[:, 1] - s[1]), np.hypot(b[:, 2] - s[2], b[:, 3] - s[3]))
        d2 = np.maximum(np.hypot(b[:, 2] - s[0], b[:, 3] - s[1]), np.hypot(b[:, 0] - s[2], b[:, 1] - s[3]))
        hit += (np.minimum(d1, d2).min() <= tol) if len(b) else 0
    return hit / len(a)


@pytest.mark.parametrize("w,h,seed", [(752, 480, 1), (752, 480, 2), (1280, 720, 2000), (640, 480, 9)])
def test_lsd_vs_cv2(oracle, plf, w, h, seed):
    """The LSD restatement (raster seed order inside a gradient bin) reproduces cv2's segments bit for bit, order
    included; the std::sort variant of the seed order does not (kept only to document that)."""
    L, R = plf.synth_pair(w, h, seed)
    for img in (L, R):
        ref = tier_a.lsd_segments(img)
        out = np.zeros((20000, 4), np.float32)
        n = C.c_int(0)
        assert oracle.dll.plf_cpu_prim_lsd(_p(img), w, h, C.c_double(1.2), 1, _p(out), 20000, C.byref(n)) == 0
        assert n.value == len(ref)
        assert np.array_equal(out[:n.value], ref)
        assert oracle.dll.plf_cpu_prim_lsd(_p(img), w, h, C.c_double(1.2), 0, _p(out), 20000, C.byref(n)) == 0
        assert _match_segments(ref, out[:n.value].copy(), 0.5) >= 0.95


@pytest.mark.parametrize("w,h,seed", [(752, 480, 1), (752, 480, 5), (1280, 720, 2000), (640, 480, 9)])
def test_lsd_refine_standard_vs_cv2(oracle, plf, w, h, seed):
    """LSD_REFINE_STD (opts.refine = 1: density check, re-grow with the local angle spread, radius reduction) is
    bit-identical to cv2 as well."""
    L, R = plf.synth_pair(w, h, seed)
    for img in (L, R):
        ref = cv2.createLineSegmentDetector(1, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024).detect(img)[0].reshape(-1, 4)
        out = np.zeros((20000, 4), np.float32)
        n = C.c_int(0)
        assert oracle.dll.plf_cpu_prim_lsd(_p(img), w, h, C.c_double(1.2), 1 | (1 << 4), _p(out), 20000, C.byref(n)) == 0
        assert n.value == len(ref) and np.array_equal(out[:n.value], ref)


# ---- stereo rectification, cv::remap(INTER_LINEAR) with CV_32F maps -------------------------------
def _euroc_maps_cv2(plf, side, w=752, h=480):
    c = plf.EUROC_CALIB[side]
    return cv2.initUndistortRectifyMap(np.array(c["K"], np.float64).reshape(3, 3), np.array(c["D"], np.float64),
                                       np.array(c["R"], np.float64).reshape(3, 3),
                                       np.array(c["P"], np.float64).reshape(3, 4)[:3, :3], (w, h), cv2.CV_32F)


@pytest.mark.parametrize("side", [0, 1])
def test_remap_euroc_vs_cv2(plf, oracle, side):
    """The oracle's remap equals cv2.remap bit for bit on the EuRoC rectification maps (stereo_euroc.cc:117-118,166-167);
    the numpy map generator used by the GPU tests equals cv2.initUndistortRectifyMap."""
    m1, m2 = _euroc_maps_cv2(plf, side)
    a, b = plf.rectify_maps(752, 480, side)
    assert np.array_equal(a, m1) and np.array_equal(b, m2)
    raw = plf.synth_pair(752, 480, 11)[side]
    f = plf.Frontend(oracle, max_batch=1)
    f.rectify_set_maps(side, m1, m2)
    assert np.array_equal(f.rectify(side, raw), cv2.remap(raw, m1, m2, cv2.INTER_LINEAR))


def test_remap_edge_cases_vs_cv2(plf, oracle):
    """Taps outside the source (BORDER_CONSTANT 0), exact integer positions (the saturated 32767/1 table entry),
    half-way roundings of map * 32, a source size different from the output size."""
    rng = np.random.default_rng(0)
    W, H = 752, 480
    f = plf.Frontend(oracle, max_batch=1)
    raw = rng.integers(0, 256, (H, W), dtype=np.uint8)
    mx = rng.uniform(-20, W + 20, (H, W)).astype(np.float32)
    my = rng.uniform(-20, H + 20, (H, W)).astype(np.float32)
    mx[::7, ::5] = np.round(mx[::7, ::5]); my[::7, ::5] = np.round(my[::7, ::5])
    mx[5, :64] = np.arange(64) + 0.5 / 32; my[5, :64] = 10 + 1.5 / 32        # ties of cvRound(map * 32)
    mx[6, :64] = -1 + np.arange(64) / 64.0; my[6, :64] = -1 + np.arange(64) / 64.0
    f.rectify_set_maps(0, mx, my)
    assert np.array_equal(f.rectify(0, raw), cv2.remap(raw, mx, my, cv2.INTER_LINEAR))
    small = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    mx = rng.uniform(-5, 405, (H, W)).astype(np.float32)
    my = rng.uniform(-5, 305, (H, W)).astype(np.float32)
    f.rectify_set_maps(1, mx, my, 400, 300)
    assert np.array_equal(f.rectify(1, small), cv2.remap(small, mx, my, cv2.INTER_LINEAR))


@pytest.mark.parametrize("w,h,seed,refine", [(752, 480, 1, 0), (752, 480, 2, 1), (640, 480, 3, 0)])
def test_lsd_vs_cv2_curved_content(oracle, plf, w, h, seed, refine):
    """Second content type (discs, rings, ellipses, polylines, smooth shading): long chains of slowly turning level-line
    angles and many rejected regions.  The restatement still equals cv2's LSD bit for bit, order included."""
    img = plf.synth_curvy(w, h, seed)
    ref = cv2.createLineSegmentDetector(refine, 1.2, 0.6, 2.0, 22.5, 1.0, 0.6, 1024).detect(img)[0].reshape(-1, 4)
    out = np.zeros((20000, 4), np.float32)
    n = C.c_int(0)
    assert oracle.dll.plf_cpu_prim_lsd(_p(img), w, h, C.c_double(1.2), 1 | (refine << 4), _p(out), 20000, C.byref(n)) == 0
    assert len(ref) > 500 and n.value == len(ref) and np.array_equal(out[:n.value], ref)


def test_clip_line_and_line_iterator_count_vs_cv2(oracle):
    """cv::LineIterator(img, p1, p2).count of LSDDetector_custom.cpp:295-296 = max(|dx|, |dy|) + 1 after cv::clipLine.
    cv2 exposes clipLine; 20 000 random segments around / across / outside several image rectangles, plus the case that
    occurs on the path (an end point clamped into [W-0.5, W) rounds to W)."""
    rng = np.random.default_rng(4)
    pts = np.zeros(4, np.int64)
    for k in range(20000):
        W, H = [(752, 480), (376, 240), (1280, 720), (7, 5)][k % 4]
        if k % 3 == 0:      # just outside on the right / bottom, as the path produces
            p = [int(rng.integers(0, W + 1)), int(rng.integers(0, H + 1)), int(rng.integers(0, W + 1)), int(rng.integers(0, H + 1))]
        else:
            p = [int(v) for v in rng.integers(-2 * max(W, H), 3 * max(W, H), 4)]
        ok, a, b = cv2.clipLine((0, 0, W, H), (p[0], p[1]), (p[2], p[3]))
        pts[:] = p
        got = oracle.dll.plf_cpu_prim_clip_line(W, H, _p(pts))
        assert bool(got) == bool(ok), (W, H, p)
        if ok:
            assert (int(pts[0]), int(pts[1])) == tuple(a) and (int(pts[2]), int(pts[3])) == tuple(b), (W, H, p)
            want = max(abs(b[0] - a[0]), abs(b[1] - a[1])) + 1
        else:
            want = 0
        fn = oracle.dll.plf_cpu_prim_line_iterator_count
        fn.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        assert fn(p[0], p[1], p[2], p[3], W, H) == want
    assert fn(375.6, 10.0, 300.0, 50.0, 376, 240) == 76        # (376,10) is clipped to (375,10)
    assert fn(375.6, 10.0, 375.7, 50.0, 376, 240) == 0         # both ends round to x = W: nothing left
