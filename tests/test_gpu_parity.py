"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.
Bars (BASELINE.json north_star): FAST candidate sets, quadtree-retained keypoints (order included), rBRIEF descriptors
and every match index bit-exact; orientation within 1e-4 rad (observed: bit-exact); LSD/LBD line endpoints within
0.5 px with >= 99 % recall and LBD Hamming distance <= 2 bits (observed: bit-exact)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ANGLE_TOL_DEG = 1e-4 * 180.0 / np.pi        # 1e-4 rad
ENDPOINT_TOL_PX = 0.5
LBD_TOL_BITS = 2


def _check_orb(kg, dg, ko, do):
    assert len(kg) == len(ko)
    for fld in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kg[fld], ko[fld]), fld
    assert np.abs(kg["angle"] - ko["angle"]).max(initial=0) <= ANGLE_TOL_DEG
    assert np.array_equal(dg, do)


def _check_lines(klg, ldg, klo, ldo):
    assert len(klg) == len(klo)
    if len(klo) == 0:
        return
    ends_g = np.stack([klg["startPointX"], klg["startPointY"], klg["endPointX"], klg["endPointY"]], 1)
    ends_o = np.stack([klo["startPointX"], klo["startPointY"], klo["endPointX"], klo["endPointY"]], 1)
    close = np.abs(ends_g - ends_o).max(axis=1) <= ENDPOINT_TOL_PX
    assert close.mean() >= 0.99
    bits = np.unpackbits(ldg ^ ldo, axis=1).sum(axis=1)
    assert bits[close].max(initial=0) <= LBD_TOL_BITS


@pytest.mark.parametrize("seed", [1, 7])
def test_single_frame_calls_stage_by_stage(plf, product, oracle, seed):
    """The five reference entry points, one stereo frame, every intermediate stage compared."""
    L, R = plf.synth_pair(752, 480, seed)
    f, o = plf.Frontend(product), plf.Frontend(oracle)
    for side, img in ((0, L), (1, R)):
        mg, kg, dg = f.orb_extract(side, img)
        mo, ko, do = o.orb_extract(side, img)
        for l in range(8):
            assert np.array_equal(f.pyramid_level(side, l), o.pyramid_level(side, l))
            assert np.array_equal(f.blurred_level(side, l), o.blurred_level(side, l))
            assert np.array_equal(f.fast_candidates(side, l), o.fast_candidates(side, l))
        assert mg == mo
        _check_orb(kg, dg, ko, do)
    ug, dpg = f.stereo_match_points(len(kg))
    uo, dpo = o.stereo_match_points(len(ko))
    nL = None
    for side, img in ((0, L), (1, R)):
        klg, ldg = f.line_extract(side, img)
        klo, ldo = o.line_extract(side, img)
        assert np.array_equal(f.lsd_scaled(side), o.lsd_scaled(side))
        assert np.array_equal(f.lsd_angles(side), o.lsd_angles(side))
        sg, so = f.lsd_segments(side), o.lsd_segments(side)
        assert sg.shape == so.shape and np.abs(sg - so).max(initial=0) <= ENDPOINT_TOL_PX
        _check_lines(klg, ldg, klo, ldo)
        assert np.array_equal(klg, klo) and np.array_equal(ldg, ldo)      # observed: bit-exact
        assert np.array_equal(f.lbd_float(side), o.lbd_float(side))
        nL = len(klg) if side == 0 else nL
    dg2, leg, mg2 = f.stereo_match_lines(nL)
    do2, leo, mo2 = o.stereo_match_lines(nL)
    assert np.array_equal(mg2, mo2) and np.array_equal(dg2, do2)
    assert np.allclose(leg, leo, rtol=1e-12, atol=0)


def test_stereo_points_after_orb(plf, product, oracle, pair1):
    L, R = pair1
    f, o = plf.Frontend(product), plf.Frontend(oracle)
    for fe in (f, o):
        fe.orb_extract(0, L)
        fe.orb_extract(1, R)
    n = len(o.orb_extract(0, L)[1])
    ug, dg = f.stereo_match_points(n)
    uo, do = o.stereo_match_points(n)
    assert np.array_equal(ug, uo) and np.array_equal(dg, do)
    assert (ug >= 0).sum() > 100


def test_state_errors(plf, product):
    f = plf.Frontend(product)
    with pytest.raises(plf.PlfError):
        f.stereo_match_points(10)          # before both extractions: PLF_ERR_STATE
    with pytest.raises(plf.PlfError):
        plf.Frontend(product, lsd_refine=2)
    for bad in (dict(scale_factor=2.5), dict(scale_factor=1.0), dict(lsd_scale=0.4), dict(min_th_fast=0),
                dict(ini_th_fast=5, min_th_fast=7), dict(width=32)):
        with pytest.raises(plf.PlfError):                    # rejected loudly, never approximated
            plf.Frontend(product, **bad)
    with pytest.raises(plf.PlfError):
        f.rectify(0, np.zeros((480, 752), np.uint8))          # before rectify_set_maps: PLF_ERR_STATE
    with pytest.raises(plf.PlfError):
        f.feature_grid(0, 1)                                  # before any extraction: PLF_ERR_STATE


GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_golden_fixtures(plf, product, path):
    """C1 (752x480, 1200 feats + lines), C3 (ORB only, 2000 feats), C4 (1280x720 lines): committed vectors."""
    g = np.load(path)
    W, H, seed = int(g["W"]), int(g["H"]), int(g["seed"])
    kw = dict(n_features=2000, has_lines=0) if "2000feat" in path else {}
    L, R = plf.synth_pair(W, H, seed)
    f = plf.Frontend(product, width=W, height=H, max_batch=1, **kw)
    r = f.frontend_batch(L[None], R[None])
    n, nr = int(r.n_kp_left[0]), int(r.n_kp_right[0])
    _check_orb(r.kp_left[0, :n], r.desc_left[0, :n], g["oracle_kp_left"], g["oracle_desc_left"])
    _check_orb(r.kp_right[0, :nr], r.desc_right[0, :nr], g["oracle_kp_right"], g["oracle_desc_right"])
    assert np.array_equal(r.u_right[0, :n], g["oracle_u_right"]) and np.array_equal(r.depth[0, :n], g["oracle_depth"])
    for l in range(8):
        assert np.array_equal(f.fast_candidates(0, l), g["cv2_cand_L%d" % l])          # real-OpenCV FAST lists
    nl = int(r.n_kl_left[0])
    assert nl == len(g["oracle_kl_left"])
    if nl:
        assert np.abs(f.lsd_segments(0) - g["cv2_lsd_left"]).max() <= ENDPOINT_TOL_PX  # real-OpenCV LSD segments
        assert np.abs(f.lsd_segments(1) - g["cv2_lsd_right"]).max() <= ENDPOINT_TOL_PX
        _check_lines(r.kl_left[0, :nl], r.ldesc_left[0, :nl], g["oracle_kl_left"], g["oracle_ldesc_left"])
        assert np.array_equal(r.line_match12[0, :nl], g["oracle_line_match12"])
        assert np.array_equal(r.disp_se[0, :nl], g["oracle_disp_se"])
        assert np.allclose(r.le[0, :nl], g["oracle_le"], rtol=1e-12, atol=0)


def test_match_nnr_and_match(plf, product, oracle):
    """matchNNR / match (config C4 uses matchNNR on the L/R LBD sets): random + structured descriptors, edge cases."""
    f, o = plf.Frontend(product), plf.Frontend(oracle)
    rng = np.random.default_rng(11)
    for n1, n2 in ((500, 500), (300, 280), (1, 2), (7, 1), (0, 5), (33, 2000)):
        d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        k = min(n1, n2) // 2
        if k:
            d2[:k] = d1[:k] ^ (rng.integers(0, 256, (k, 32), dtype=np.uint8) & 1)
            d2[k // 2] = d2[0]                                    # an exact tie
        for nnr in (0.9, 0.6, 1.0):
            a, b = f.match_nnr(d1, d2, nnr), o.match_nnr(d1, d2, nnr)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])
            a, b = f.match(d1, d2, nnr, True), o.match(d1, d2, nnr, True)
            assert a[0] == b[0] and np.array_equal(a[1], b[1])
    L, R = plf.synth_pair(1280, 720, 2001)
    f4 = plf.Frontend(product, width=1280, height=720)
    o4 = plf.Frontend(oracle, width=1280, height=720)
    kl, dl = f4.line_extract(0, L)
    kr, dr = f4.line_extract(1, R)
    a, b = f4.match_nnr(dl, dr, 0.9), o4.match_nnr(dl, dr, 0.9)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 50


def test_batch64_full_size_properties(plf, product, oracle):
    """BASELINE config C2 at full size (batch 64): a sample of pairs against the oracle, and for the whole batch the
    size-independent properties: run-to-run determinism, batch == single-frame, slot independence (a permuted batch
    gives permuted results), Hamming symmetry of the stereo matches, descriptor self-match."""
    B = 64
    seeds = [1000 + i for i in range(16)]
    Ld, Rd = plf.synth_batch(752, 480, seeds)
    idx = np.arange(B) % 16
    L, R = Ld[idx], Rd[idx]
    f = plf.Frontend(product, max_batch=B, lsd_nfeatures=300)
    r1 = f.frontend_batch(L, R)
    r2 = f.frontend_batch(L, R)
    names = ["n_kp_left", "n_kp_right", "n_kl_left", "n_kl_right", "kp_left", "kp_right", "desc_left", "desc_right",
             "u_right", "depth", "kl_left", "kl_right", "ldesc_left", "ldesc_right", "disp_se", "le", "line_match12"]

    def rows(res, name, b):
        a = getattr(res, name)
        if name.startswith("n_"):
            return a[b]
        n = res.n_kl_left[b] if name in ("kl_left", "ldesc_left", "disp_se", "le", "line_match12") else \
            res.n_kl_right[b] if name in ("kl_right", "ldesc_right") else \
            res.n_kp_right[b] if name in ("kp_right", "desc_right") else res.n_kp_left[b]
        return a[b, :n]

    for name in names:
        for b in range(B):
            assert np.array_equal(rows(r1, name, b), rows(r2, name, b)), ("determinism", name, b)
            assert np.array_equal(rows(r1, name, b), rows(r1, name, b % 16)), ("slot independence", name, b)
    perm = np.random.default_rng(0).permutation(B)
    rp = f.frontend_batch(L[perm], R[perm])
    for name in names:
        for b in range(0, B, 5):
            assert np.array_equal(rows(rp, name, b), rows(r1, name, perm[b])), ("permutation", name, b)
    # oracle on a sample of the pairs
    o = plf.Frontend(oracle, max_batch=4, lsd_nfeatures=300)
    ro = o.frontend_batch(L[:4], R[:4])
    for name in names:
        for b in range(4):
            if name == "le":
                assert np.allclose(rows(r1, name, b), rows(ro, name, b), rtol=1e-12, atol=0)
            else:
                assert np.array_equal(rows(r1, name, b), rows(ro, name, b)), ("oracle", name, b)
    # single-frame API == batch API
    fs = plf.Frontend(product, lsd_nfeatures=300)
    _, k0, d0 = fs.orb_extract(0, L[3])
    assert np.array_equal(k0, rows(r1, "kp_left", 3)) and np.array_equal(d0, rows(r1, "desc_left", 3))
    # descriptor self-match: every row matches itself at distance 0
    n, m = fs.match_nnr(d0, d0, 0.9)
    dup = len(d0) - len(np.unique(d0, axis=0))
    assert n >= len(d0) - 2 * dup
    # every stereo match lies to the left (disparity >= 0) and inside maxD = fx
    u, kx = rows(r1, "u_right", 0), rows(r1, "kp_left", 0)["x"]
    ok = u >= 0
    assert (kx[ok] - u[ok] >= 0).all() and (kx[ok] - u[ok] < 435.3).all()


@pytest.mark.parametrize("W,H,kw", [
    (641, 479, dict()),                                           # odd size: unaligned rows, scalar unpack path
    (320, 240, dict(n_features=500, n_levels=5)),                 # small image, fewer levels
    (1280, 720, dict(lsd_nfeatures=0, n_features=1500)),          # C4 size, keep ALL lines (lsd_nfeatures = 0)
    (752, 480, dict(scale_factor=1.5, n_levels=4, ini_th_fast=30, min_th_fast=10, lsd_nfeatures=100)),
    (752, 480, dict(has_lines=0, n_features=3000)),               # ORB only, many features
    (752, 480, dict(best_lr_matches=0, matching_s_ws=20, min_ratio_12_l=0.8, line_sim_th=0.5)),
    (752, 480, dict(lsd_refine=1)),                               # LSD_REFINE_STD: re-grow + radius reduction
    (1280, 720, dict(lsd_refine=1, lsd_density_th=0.8, lsd_nfeatures=0)),
    (1241, 376, dict(n_features=2000)),                            # Examples/Stereo/Config/KITTI00-02.yaml: 4 quadtree roots
    (1241, 376, dict(n_features=2000, ini_th_fast=12)),            # KITTI04-12.yaml (iniThFAST 12)
])
def test_shapes_and_parameters(plf, product, oracle, W, H, kw):
    """Geometry and parameter sweep (batch of 2 pairs): every output array identical to the oracle."""
    L, R = plf.synth_batch(W, H, [31, 32])
    f = plf.Frontend(product, width=W, height=H, max_batch=2, **kw)
    o = plf.Frontend(oracle, width=W, height=H, max_batch=2, **kw)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for b in range(2):
        for side in ("left", "right"):
            n = int(getattr(ro, "n_kp_" + side)[b])
            assert int(getattr(rg, "n_kp_" + side)[b]) == n and n > 100
            assert np.array_equal(getattr(rg, "kp_" + side)[b, :n], getattr(ro, "kp_" + side)[b, :n])
            assert np.array_equal(getattr(rg, "desc_" + side)[b, :n], getattr(ro, "desc_" + side)[b, :n])
            nl = int(getattr(ro, "n_kl_" + side)[b])
            assert int(getattr(rg, "n_kl_" + side)[b]) == nl
            assert np.array_equal(getattr(rg, "kl_" + side)[b, :nl], getattr(ro, "kl_" + side)[b, :nl])
            assert np.array_equal(getattr(rg, "ldesc_" + side)[b, :nl], getattr(ro, "ldesc_" + side)[b, :nl])
        n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        assert np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n]) and np.array_equal(rg.depth[b, :n], ro.depth[b, :n])
        assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl])
        assert np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl])
        assert np.allclose(rg.le[b, :nl], ro.le[b, :nl], rtol=1e-12, atol=0)


@pytest.mark.parametrize("lap", [(0, 1000), (300, 500), (100, 101), (-5, -1)])
def test_lapping_area_on_device(plf, product, oracle, lap):
    """vLappingArea of ORBextractor::operator() (src/ORBextractor.cc:1135-1146; the monocular constructor passes {0, 1000},
    src/Frame.cc:360-361): rows inside the interval are written back to front, the rest front to back, the return value is
    the number of the rest.  3000 features so that more than 1024 rows fall inside (the placement kernel carries its
    running counts across 1024-thread chunks).  Row order, every field, every descriptor byte, monoIndex: identical."""
    L, R = plf.synth_pair(752, 480, 3)
    f = plf.Frontend(product, n_features=3000, has_lines=0, max_batch=1)
    o = plf.Frontend(oracle, n_features=3000, has_lines=0, max_batch=1)
    for side, img in ((0, L), (1, R)):
        mg, kg, dg = f.orb_extract(side, img, lapping=lap)
        mo, ko, do = o.orb_extract(side, img, lapping=lap)
        assert mg == mo and np.array_equal(kg, ko) and np.array_equal(dg, do)
        inside = int(((ko["x"] >= lap[0]) & (ko["x"] <= lap[1])).sum())
        assert mo == len(ko) - inside
        if lap == (0, 1000):
            assert mo == 0 and inside > 2048
        if lap == (300, 500):
            assert 256 < inside < len(ko)


def test_row_stride_larger_than_width(plf, product, oracle):
    """Images whose rows are padded (stride = width + 40, cv::Mat ROIs): the single-frame calls and the batched upload
    read only the first `width` bytes of each row."""
    W, H, S = 752, 480, 752 + 40
    L, R = plf.synth_batch(W, H, [61, 62, 63])
    rng = np.random.default_rng(9)
    padL = rng.integers(0, 256, (3, H, S), dtype=np.uint8)
    padR = rng.integers(0, 256, (3, H, S), dtype=np.uint8)
    padL[:, :, :W], padR[:, :, :W] = L, R
    vL, vR = padL[:, :, :W], padR[:, :, :W]
    assert vL.strides == (H * S, S, 1)
    f, o = plf.Frontend(product, max_batch=3), plf.Frontend(oracle, max_batch=3)
    for side, view, dense in ((0, vL[1], L[1]), (1, vR[1], R[1])):
        mg, kg, dg = f.orb_extract(side, view)
        mo, ko, do = o.orb_extract(side, dense)
        assert mg == mo and np.array_equal(kg, ko) and np.array_equal(dg, do)
        klg, ldg = f.line_extract(side, view)
        klo, ldo = o.line_extract(side, dense)
        assert np.array_equal(klg, klo) and np.array_equal(ldg, ldo)
    rg, ro = f.frontend_batch(vL, vR), o.frontend_batch(L, R)
    for b in range(3):
        n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        assert int(rg.n_kp_left[b]) == n and int(rg.n_kl_left[b]) == nl
        assert np.array_equal(rg.kp_left[b, :n], ro.kp_left[b, :n]) and np.array_equal(rg.desc_left[b, :n], ro.desc_left[b, :n])
        assert np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n])
        assert np.array_equal(rg.kl_left[b, :nl], ro.kl_left[b, :nl]) and np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl])


def test_limits_are_refused_at_create(plf, product):
    """Feature budgets whose on-chip lists would not fit are refused by plf_create (PLF_ERR_UNSUPPORTED), not at launch."""
    for kw in (dict(n_features=60000), dict(n_features=40000, n_levels=2), dict(lsd_nfeatures=9000)):
        with pytest.raises(plf.PlfError) as e:
            plf.Frontend(product, max_batch=1, **kw)
        assert e.value.code == 5
    f = plf.Frontend(product, n_features=15000, has_lines=0, max_batch=1)      # above the old 12 k stereo-cull limit: works
    L, R = plf.synth_pair(752, 480, 2)
    r = f.frontend_batch(L[None], R[None])
    assert int(r.n_kp_left[0]) > 3000 and int((r.u_right[0] >= 0).sum()) > 300


def test_degenerate_images(plf, product, oracle):
    """Flat, saturated and pure-noise frames: no crash, same (possibly empty) outputs as the oracle."""
    rng = np.random.default_rng(5)
    L = np.stack([np.full((480, 752), 127, np.uint8), np.full((480, 752), 255, np.uint8),
                  rng.integers(0, 256, (480, 752), dtype=np.uint8), np.zeros((480, 752), np.uint8)])
    R = L[::-1].copy()
    f = plf.Frontend(product, max_batch=4)
    o = plf.Frontend(oracle, max_batch=4)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for name in ("n_kp_left", "n_kp_right", "n_kl_left", "n_kl_right"):
        assert np.array_equal(getattr(rg, name), getattr(ro, name)), name
    for b in range(4):
        n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        assert np.array_equal(rg.kp_left[b, :n], ro.kp_left[b, :n])
        assert np.array_equal(rg.desc_left[b, :n], ro.desc_left[b, :n])
        assert np.array_equal(rg.kl_left[b, :nl], ro.kl_left[b, :nl])
        if n and (nl or True):
            pass
    # noise frame: both matchers ran on the GPU; compare where the reference would have run them (non-empty frames)
    b = 2
    n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
    if n and nl:
        assert np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n])
        assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl])


# ---- stereo rectification in front of the path ---------------------------------------------------
def test_rectify_matches_oracle(plf, product, oracle):
    """plf_rectify == the oracle's cv::remap restatement (itself pinned to cv2.remap) bit for bit: EuRoC maps of both
    cameras, random maps with taps outside the source, integer positions and rounding ties, another source size."""
    W, H = 752, 480
    f, o = plf.Frontend(product, max_batch=1), plf.Frontend(oracle, max_batch=1)
    L, R = plf.synth_pair(W, H, 11)
    for side, raw in ((0, L), (1, R)):
        mx, my = plf.rectify_maps(W, H, side)
        f.rectify_set_maps(side, mx, my); o.rectify_set_maps(side, mx, my)
        assert np.array_equal(f.rectify(side, raw), o.rectify(side, raw))
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 256, (H, W), dtype=np.uint8)
    mx = rng.uniform(-20, W + 20, (H, W)).astype(np.float32)
    my = rng.uniform(-20, H + 20, (H, W)).astype(np.float32)
    mx[::7, ::5] = np.round(mx[::7, ::5]); my[::7, ::5] = np.round(my[::7, ::5])
    mx[5, :64] = np.arange(64) + 0.5 / 32; my[5, :64] = 10 + 1.5 / 32
    f.rectify_set_maps(0, mx, my); o.rectify_set_maps(0, mx, my)
    assert np.array_equal(f.rectify(0, raw), o.rectify(0, raw))
    small = rng.integers(0, 256, (300, 400), dtype=np.uint8)
    mx = rng.uniform(-5, 405, (H, W)).astype(np.float32)
    my = rng.uniform(-5, 305, (H, W)).astype(np.float32)
    f.rectify_set_maps(1, mx, my, 400, 300); o.rectify_set_maps(1, mx, my, 400, 300)
    assert np.array_equal(f.rectify(1, small), o.rectify(1, small))
    # odd output width: the tail threads store single bytes
    f2, o2 = plf.Frontend(product, width=641, height=479, max_batch=1), plf.Frontend(oracle, width=641, height=479, max_batch=1)
    mx = rng.uniform(0, 640, (479, 641)).astype(np.float32); my = rng.uniform(0, 478, (479, 641)).astype(np.float32)
    raw = rng.integers(0, 256, (479, 641), dtype=np.uint8)
    f2.rectify_set_maps(0, mx, my); o2.rectify_set_maps(0, mx, my)
    assert np.array_equal(f2.rectify(0, raw), o2.rectify(0, raw))


def test_raw_upload_equals_rectified_upload(plf, product, oracle):
    """batch_upload_raw (H2D + rectification on the device) + batch_run gives exactly the results of uploading the
    rectified frames, and the oracle agrees; calling it before the maps are set is a state error."""
    W, H = 752, 480
    Lr, Rr = plf.synth_batch(W, H, [41, 42, 43])
    f, o = plf.Frontend(product, max_batch=3), plf.Frontend(oracle, max_batch=3)
    with pytest.raises(plf.PlfError):
        f.batch_upload_raw(Lr, Rr)
    for side in (0, 1):
        mx, my = plf.rectify_maps(W, H, side)
        f.rectify_set_maps(side, mx, my); o.rectify_set_maps(side, mx, my)
    Lc = np.stack([o.rectify(0, im) for im in Lr]); Rc = np.stack([o.rectify(1, im) for im in Rr])
    ref = f.frontend_batch(Lc, Rc)
    out = f.new_result(3)
    f.batch_upload_raw(Lr, Rr); f.batch_run(3); f.batch_download(3, out)
    oo = o.new_result(3)
    o.batch_upload_raw(Lr, Rr); o.batch_run(3); o.batch_download(3, oo)
    for b in range(3):
        n, nl = int(ref.n_kp_left[b]), int(ref.n_kl_left[b])
        assert n > 500 and nl > 50
        for r in (out, oo):
            assert int(r.n_kp_left[b]) == n and int(r.n_kl_left[b]) == nl
            assert np.array_equal(r.kp_left[b, :n], ref.kp_left[b, :n]) and np.array_equal(r.desc_left[b, :n], ref.desc_left[b, :n])
            assert np.array_equal(r.kl_left[b, :nl], ref.kl_left[b, :nl]) and np.array_equal(r.ldesc_left[b, :nl], ref.ldesc_left[b, :nl])
            assert np.array_equal(r.u_right[b, :n], ref.u_right[b, :n]) and np.array_equal(r.line_match12[b, :nl], ref.line_match12[b, :nl])


def test_feature_grid_matches_oracle(plf, product, oracle):
    """plf_feature_grid (Frame::AssignFeaturesToGrid as CSR) for a batch: identical to the oracle, every keypoint listed
    once in ascending order per cell, and the area lookup returns the same indices on both libraries."""
    W, H = 752, 480
    L, R = plf.synth_batch(W, H, [51, 52, 53, 54])
    f, o = plf.Frontend(product, max_batch=4), plf.Frontend(oracle, max_batch=4)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    sg, ig = f.feature_grid(0, 4)
    so, io = o.feature_grid(0, 4)
    assert np.array_equal(sg, so)
    for b in range(4):
        n = int(ro.n_kp_left[b])
        tot = int(so[b, -1])
        assert n - 5 <= tot <= n
        assert np.array_equal(ig[b, :tot], io[b, :tot])
        assert len(np.unique(ig[b, :tot])) == tot
        for c in np.nonzero(np.diff(so[b]) > 1)[0][:50]:
            assert np.all(np.diff(ig[b, so[b, c]:so[b, c + 1]]) > 0)
        got = f.features_in_area(rg.kp_left[b, :n], sg[b], ig[b], 300.5, 200.25, 25.0, 0, 3)
        want = o.features_in_area(ro.kp_left[b, :n], so[b], io[b], 300.5, 200.25, 25.0, 0, 3)
        assert np.array_equal(got, want) and len(want) > 0
    s1, i1 = f.feature_grid(2, 2)           # a sub-range of slots
    assert np.array_equal(s1, so[2:]) and np.array_equal(i1[0, :so[2, -1]], io[2, :so[2, -1]])


def test_backproject_matches_oracle(plf, product, oracle):
    """plf_backproject (Frame::UnprojectStereo per keypoint, Frame::backProjection per line end point) for a batch with a
    different pose per slot: bit-identical to the oracle, zeros exactly where the reference has no landmark."""
    W, H = 752, 480
    L, R = plf.synth_batch(W, H, [61, 62, 63])
    f, o = plf.Frontend(product, max_batch=3), plf.Frontend(oracle, max_batch=3)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    rng = np.random.default_rng(4)
    Rwc = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(3)]).astype(np.float32)
    Ow = (rng.normal(size=(3, 3)) * 3).astype(np.float32)
    xg, lg = f.backproject(Rwc, Ow, 435.2047, 367.4517, 252.2009)
    xo, lo = o.backproject(Rwc, Ow, 435.2047, 367.4517, 252.2009)
    assert np.array_equal(xg, xo) and np.array_equal(lg, lo)
    for b in range(3):
        n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        has = ro.depth[b, :n] > 0
        assert has.sum() > 100 and np.all(np.any(xg[b, :n][has] != 0, axis=1)) and not xg[b, :n][~has].any() and not xg[b, n:].any()
        st = (ro.disp_se[b, :nl] > 0).all(axis=1)
        assert st.sum() > 20 and not lg[b, :nl][~st].any()
    x1, _ = f.backproject(Rwc[1:2], Ow[1:2], 435.2047, 367.4517, 252.2009, first_slot=1, lines=False)
    assert np.array_equal(x1[0], xo[1])


@pytest.mark.parametrize("refine", [0, 1])
def test_curved_content_matches_oracle(plf, product, oracle, refine):
    """The whole path on the second content type (curved and slanted edges, smooth shading); right image = left shifted
    by 14 px with its own noise.  Everything identical to the oracle."""
    W, H = 752, 480
    Ls = np.stack([plf.synth_curvy(W, H, s) for s in (5, 6)])
    rng = np.random.default_rng(9)
    Rs = np.clip(np.roll(Ls, -14, axis=2).astype(np.int16) + rng.integers(-2, 3, Ls.shape), 0, 255).astype(np.uint8)
    kw = dict(lsd_refine=refine, lsd_nfeatures=0)
    f, o = plf.Frontend(product, max_batch=2, **kw), plf.Frontend(oracle, max_batch=2, **kw)
    rg, ro = f.frontend_batch(Ls, Rs), o.frontend_batch(Ls, Rs)
    for b in range(2):
        for side in ("left", "right"):
            n, nl = int(getattr(ro, "n_kp_" + side)[b]), int(getattr(ro, "n_kl_" + side)[b])
            assert n > 300 and nl > 200
            assert int(getattr(rg, "n_kp_" + side)[b]) == n and int(getattr(rg, "n_kl_" + side)[b]) == nl
            assert np.array_equal(getattr(rg, "kp_" + side)[b, :n], getattr(ro, "kp_" + side)[b, :n])
            assert np.array_equal(getattr(rg, "desc_" + side)[b, :n], getattr(ro, "desc_" + side)[b, :n])
            assert np.array_equal(getattr(rg, "kl_" + side)[b, :nl], getattr(ro, "kl_" + side)[b, :nl])
            assert np.array_equal(getattr(rg, "ldesc_" + side)[b, :nl], getattr(ro, "ldesc_" + side)[b, :nl])
        n, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        assert (ro.u_right[b, :n] >= 0).sum() > 50
        assert np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n]) and np.array_equal(rg.depth[b, :n], ro.depth[b, :n])
        assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl]) and np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl])


def test_bow_transform_matches_oracle(plf, product, oracle):
    """plf_bow_transform (DBoW2 descent of every left ORB / LBD descriptor) for a batch: word ids, weights and node ids
    identical to the oracle; BowVector / FeatureVector built from them identical too; loud errors without a vocabulary."""
    L, R = plf.synth_batch(752, 480, [71, 72, 73])
    f, o = plf.Frontend(product, max_batch=3), plf.Frontend(oracle, max_batch=3)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    with pytest.raises(plf.PlfError):
        f.bow_transform(0, 3)
    for which, levelsup, voc in ((0, 4, plf.synth_vocabulary(10, 6, seed=1, ragged=0.3)), (1, 2, plf.synth_vocabulary(8, 4, seed=2))):
        f.bow_set_vocabulary(which, voc); o.bow_set_vocabulary(which, voc)
        wg, vg, ng = f.bow_transform(which, 3, 0, levelsup)
        wo, vo, no = o.bow_transform(which, 3, 0, levelsup)
        assert np.array_equal(wg, wo) and np.array_equal(vg, vo) and np.array_equal(ng, no)
        n = int(ro.n_kl_left[1]) if which else int(ro.n_kp_left[1])
        assert n > 100 and (wo[1, :n] >= 0).all()
        bg, bo = f.bow_build(wg[1, :n], vg[1, :n], ng[1, :n]), o.bow_build(wo[1, :n], vo[1, :n], no[1, :n])
        assert np.array_equal(bg[0], bo[0]) and np.array_equal(bg[1], bo[1]) and bg[2] == bo[2]
    w1, v1, n1 = f.bow_transform(0, 1, 2, 4)
    assert np.array_equal(w1[0], wo_first := f.bow_transform(0, 3, 0, 4)[0][2])
    bad = plf.synth_vocabulary(4, 2, seed=5)
    bad["child"] = bad["child"].copy(); bad["child"][0] = 0          # a child pointing back at the root
    with pytest.raises(plf.PlfError):
        f.bow_set_vocabulary(0, bad)


def test_failed_create_releases_device_memory(plf, product):
    """A context that does not fit (cudaMalloc fails half-way through plf_create) reports PLF_ERR_CUDA and leaves no device
    memory behind; the next context works."""
    import torch
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    with pytest.raises(plf.PlfError) as e:
        plf.Frontend(product, max_batch=12000)              # ~300 GB of buffers
    assert e.value.code == 3                                # PLF_ERR_CUDA
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (1 << 30)
    f = plf.Frontend(product, max_batch=1)
    L, R = plf.synth_pair(752, 480, 1)
    assert int(f.frontend_batch(L[None], R[None]).n_kp_left[0]) > 1000


@pytest.mark.parametrize("th", [1.0, 3.0, 10.0])
def test_search_by_projection_matches_oracle(plf, product, oracle, th):
    """plf_search_by_projection (window search + Hamming on the device, order-dependent assignment on the host) against
    the oracle: ~1350 map points against a frame with 10 % of its features already taken; matches, nmatches and the
    updated occupancy identical; th = 10 gives windows of up to +-140 px (hundreds of candidates per map point)."""
    from test_host_logic import _proj_queries
    L, R = plf.synth_batch(752, 480, [81, 82])
    f, o = plf.Frontend(product, max_batch=2), plf.Frontend(oracle, max_batch=2)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for b in (1, 0):
        n = int(ro.n_kp_left[b])
        rng = np.random.default_rng(20 + b)
        q = _proj_queries(plf, ro, b, rng)
        occ0 = (rng.random(n) < 0.1).astype(np.uint8)
        og, oo = occ0.copy(), occ0.copy()
        mg, ng = f.search_by_projection(q, og, th=th, slot=b)
        mo, no = o.search_by_projection(q, oo, th=th, slot=b)
        assert ng == no and np.array_equal(mg, mo) and np.array_equal(og, oo)
        assert no > 300
    with pytest.raises(plf.PlfError):
        plf.Frontend(product, max_batch=1).search_by_projection(q[:4], np.zeros(10, np.uint8))     # before extraction
    for fe in (f, o):                                                                               # occupied[] shorter than Frame::N
        with pytest.raises(plf.PlfError) as e:
            fe.search_by_projection(q[:4], np.zeros(10, np.uint8), slot=0)
        assert e.value.code == 1
        with pytest.raises(plf.PlfError) as e:
            fe.search_by_projection_frame(np.zeros(2, plf.FRAME_QUERY_DT), np.zeros(10, np.uint8), slot=0)
        assert e.value.code == 1


@pytest.mark.parametrize("mode,check", [("around", True), ("forward", True), ("backward", False)])
def test_search_by_projection_frame_matches_oracle(plf, product, oracle, mode, check):
    """plf_search_by_projection_frame (TrackWithMotionModel's search: device candidates, in-order host resolve, rotation
    histogram) against the oracle: final feature -> map point table, the match12 map, nmatches and occupancy identical."""
    from test_host_logic import _frame_queries
    L, R = plf.synth_batch(752, 480, [91, 92])
    f, o = plf.Frontend(product, max_batch=2), plf.Frontend(oracle, max_batch=2)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for b in (0, 1):
        n = int(ro.n_kp_left[b])
        rng = np.random.default_rng(30 + b)
        q = _frame_queries(plf, ro, b, rng, mode)
        occ0 = (rng.random(n) < 0.08).astype(np.uint8)
        og, oo = occ0.copy(), occ0.copy()
        fg, mg, ng = f.search_by_projection_frame(q, og, 100, check, slot=b)
        fo, mo, no = o.search_by_projection_frame(q, oo, 100, check, slot=b)
        assert np.array_equal(rg.kp_left[b, :n], ro.kp_left[b, :n]) and np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n])
        assert ng == no, (ng, no)
        assert np.array_equal(fg, fo), np.nonzero(fg != fo)[0][:10]
        assert np.array_equal(mg, mo) and np.array_equal(og, oo)
        assert no > 400


def test_keyframe_searches_match_oracle(plf, product, oracle):
    """The remaining SearchByProjection overloads (relocalisation src/ORBmatcher.cc:2325-2447, loop closing :473-704) and
    SearchByBoW (:269-471) through the C ABI against the oracle: assignment tables, nmatches and occupancy identical."""
    from test_host_logic import _frame_queries, _bow_case
    L, R = plf.synth_batch(752, 480, [93, 94])
    f, o = plf.Frontend(product, max_batch=2), plf.Frontend(oracle, max_batch=2)
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for b in (1, 0):
        n = int(ro.n_kp_left[b])
        assert np.array_equal(rg.kp_left[b, :n], ro.kp_left[b, :n])
        rng = np.random.default_rng(50 + b)
        q = _frame_queries(plf, ro, b, rng, "around")
        occ0 = (rng.random(n) < 0.1).astype(np.uint8)
        for check in (True, False):
            og, oo = occ0.copy(), occ0.copy()
            fg, ng = f.search_by_projection_reloc(q, og, 64, check, slot=b)
            fo, no = o.search_by_projection_reloc(q, oo, 64, check, slot=b)
            assert ng == no and np.array_equal(fg, fo) and np.array_equal(og, oo) and no > 300
        q2 = _frame_queries(plf, ro, b, rng, "backward")
        q2["min_level"] = q2["max_level"] - 1
        for ratio in (1.0, 0.64):
            og, oo = occ0.copy(), occ0.copy()
            fg, ng = f.search_by_projection_loop(q2, og, 50, ratio, slot=b)
            fo, no = o.search_by_projection_loop(q2, oo, 50, ratio, slot=b)
            assert ng == no and np.array_equal(fg, fo) and np.array_equal(og, oo) and no > 150
        case = _bow_case(plf, ro, rng, b)
        for check, ratio in ((True, 0.7), (False, 0.9)):
            mg, ng = f.search_by_bow(*case, 50, ratio, check, slot=b)
            mo, no = o.search_by_bow(*case, 50, ratio, check, slot=b)
            assert ng == no and np.array_equal(mg, mo) and no > 200
        e, en = f.search_by_bow(case[0][:0], case[1][:0], case[2][:0], case[3][:0], case[4], slot=b)
        assert en == 0 and np.all(e == -1)
    with pytest.raises(plf.PlfError):
        f.search_by_bow(*case[:4], case[4][:10], slot=0)                 # f_node[] shorter than Frame::N


@pytest.mark.parametrize("mode", [0, 1])
def test_match_lines_tracked_matches_oracle(plf, product, oracle, mode):
    """match() + the tracking thread's orientation / position gates (src/Tracking.cc:3055-3099, :3879-3917), fused on the
    device, against the oracle: matches_12, the assignment list and the inlier count identical."""
    from test_host_logic import _track_lines_case
    L, R = plf.synth_batch(752, 480, [95, 96])
    f, o = plf.Frontend(product, max_batch=2, lsd_nfeatures=0), plf.Frontend(oracle, max_batch=2, lsd_nfeatures=0)
    ro = o.frontend_batch(L, R)
    f.frontend_batch(L, R)
    for b in (0, 1):
        sub = type("R", (), {})()
        sub.n_kl_left = ro.n_kl_left[b:b + 1]; sub.kl_left = ro.kl_left[b:b + 1]; sub.ldesc_left = ro.ldesc_left[b:b + 1]
        sub.disp_se = ro.disp_se[b:b + 1]
        case = _track_lines_case(plf, sub, np.random.default_rng(60 + b + 2 * mode), mode)
        for bounds in ((0.0, 752.0, 0.0, 480.0), (0.0, 300.0, 0.0, 200.0)):
            mg, ag, ng = f.match_lines_tracked(mode, *case, 0.9, bounds)
            mo, ao, no = o.match_lines_tracked(mode, *case, 0.9, bounds)
            assert ng == no and np.array_equal(mg, mo) and np.array_equal(ag, ao)
        assert no > 20
    mg, ag, ng = f.match_lines_tracked(mode, case[0][:0], case[1][:0], case[2], case[3], case[4], case[5], 0.9, bounds)
    assert ng == 0 and len(mg) == 0
    mg, ag, ng = f.match_lines_tracked(mode, case[0], case[1], case[2][:1], case[3][:1], case[4][:1], None, 0.9, bounds)
    assert ng == 0 and np.all(mg == -1)                                   # fewer than two train rows: no matches (declared rule)


def test_repeated_calls_replay_the_graph(plf, product, oracle):
    """From the second call with one batch size on, plf_batch_run replays a captured CUDA graph: four calls on one context
    with different frames (and a change of batch size in between, which re-captures) against the oracle, every array."""
    W, H = 752, 480
    f = plf.Frontend(product, max_batch=3, lsd_nfeatures=0)
    o = plf.Frontend(oracle, max_batch=3, lsd_nfeatures=0)
    fields = [("kp_left", "n_kp_left"), ("desc_left", "n_kp_left"), ("kp_right", "n_kp_right"), ("desc_right", "n_kp_right"),
              ("kl_left", "n_kl_left"), ("ldesc_left", "n_kl_left"), ("kl_right", "n_kl_right"), ("ldesc_right", "n_kl_right"),
              ("u_right", "n_kp_left"), ("depth", "n_kp_left"), ("disp_se", "n_kl_left"), ("line_match12", "n_kl_left")]
    for call, seeds in enumerate(([201, 202, 203], [204, 205, 206], [207, 208, 209], [210, 211], [212, 213], [214, 215, 216])):
        L, R = plf.synth_batch(W, H, seeds)
        rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
        for b in range(len(seeds)):
            for name, cnt in fields:
                n = int(getattr(ro, cnt)[b])
                assert int(getattr(rg, cnt)[b]) == n, (call, b, cnt)
                assert np.array_equal(getattr(rg, name)[b, :n], getattr(ro, name)[b, :n]), (call, b, name)
        assert f.launch_count() > 20


def test_large_batch_sequential_grower_matches_oracle(plf, product, oracle):
    """Launches of more than 296 images use the one-warp-per-image region grower (the kernel the benchmark runs); smaller
    ones use the streaming multi-warp grower.  152 pairs (304 images, 8 distinct pairs repeated) against the oracle, exactly, and
    against the same pairs sent through a small batch."""
    W, H = 752, 480
    L8, R8 = plf.synth_batch(W, H, [101, 102, 103, 104, 105, 106, 107, 108])
    idx = np.arange(152) % 8
    f = plf.Frontend(product, max_batch=152, lsd_nfeatures=0)
    o = plf.Frontend(oracle, max_batch=8, lsd_nfeatures=0)
    rg, ro = f.frontend_batch(L8[idx], R8[idx]), o.frontend_batch(L8, R8)
    small = plf.Frontend(product, max_batch=8, lsd_nfeatures=0).frontend_batch(L8, R8)
    for b in range(152):
        r = b % 8
        for side in ("left", "right"):
            nl = int(getattr(ro, "n_kl_" + side)[r])
            assert int(getattr(rg, "n_kl_" + side)[b]) == nl and nl > 300
            assert np.array_equal(getattr(rg, "kl_" + side)[b, :nl], getattr(ro, "kl_" + side)[r, :nl])
            assert np.array_equal(getattr(rg, "ldesc_" + side)[b, :nl], getattr(ro, "ldesc_" + side)[r, :nl])
            if b < 8:
                assert np.array_equal(getattr(small, "kl_" + side)[b, :nl], getattr(ro, "kl_" + side)[r, :nl])
        n, nl = int(ro.n_kp_left[r]), int(ro.n_kl_left[r])
        assert np.array_equal(rg.kp_left[b, :n], ro.kp_left[r, :n]) and np.array_equal(rg.u_right[b, :n], ro.u_right[r, :n])
        assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[r, :nl])


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,batch,refine", [(752, 480, 1, 0), (752, 480, 5, 0), (641, 479, 2, 0), (1241, 376, 3, 0), (752, 480, 2, 1)])
def test_streaming_small_batch_grower_matches_oracle(plf, product, oracle, W, H, batch, refine):
    """Launches of at most 296 images go through lsd_grow_sw_kernel (one region per warp, 16 regions of an image in flight,
    in-order commit pointer): every segment-derived array equal to the oracle with ALL lines kept, over several calls on one
    context (owner map, position map and record buffers are reused from call to call).  refine = 1: regions that need
    refining are grown and refined by the kernel's committing warp."""
    f = plf.Frontend(product, width=W, height=H, max_batch=batch, lsd_nfeatures=0, lsd_refine=refine)
    o = plf.Frontend(oracle, width=W, height=H, max_batch=batch, lsd_nfeatures=0, lsd_refine=refine)
    for call in range(3):
        L, R = plf.synth_batch(W, H, [9100 + 17 * call + b for b in range(batch)])
        rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
        for b in range(batch):
            for side in ("left", "right"):
                nl = int(getattr(ro, "n_kl_" + side)[b])
                assert int(getattr(rg, "n_kl_" + side)[b]) == nl and nl > 100, (call, b, side)
                assert np.array_equal(getattr(rg, "kl_" + side)[b, :nl], getattr(ro, "kl_" + side)[b, :nl]), (call, b, side)
                assert np.array_equal(getattr(rg, "ldesc_" + side)[b, :nl], getattr(ro, "ldesc_" + side)[b, :nl]), (call, b, side)
            nl = int(ro.n_kl_left[b])
            assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl]), (call, b)


def test_grower_policy_switch_gives_identical_results(plf, product):
    """plf_set_grower_policy: the streaming multi-warp grower (automatic choice for small launches) and the one-warp-per-image
    grower (throughput policy) give the same arrays, through the eager pass and the replayed CUDA graph, switching back and forth."""
    W, H = 752, 480
    L, R = plf.synth_batch(W, H, [4301, 4302, 4303])
    f = plf.Frontend(product, max_batch=3, lsd_nfeatures=0)
    ref = None
    for policy in (0, 1, 0, 1):
        f.set_grower_policy(policy)
        for rep in range(3):                      # eager, capture, replay
            r = f.frontend_batch(L, R)
            got = {k: np.array(getattr(r, k), copy=True) for k in ("n_kl_left", "n_kl_right", "kl_left", "kl_right", "ldesc_left", "line_match12", "kp_left", "u_right")}
            if ref is None:
                ref = got
                assert int(ref["n_kl_left"].min()) > 100
            for k, v in got.items():
                assert np.array_equal(v, ref[k]), (policy, rep, k)
    with pytest.raises(Exception):
        f.set_grower_policy(7)


def test_single_image_calls_with_speculative_line_path(plf, product, oracle):
    """plf_orb_extract starts the line path of its image on a side stream once the library has seen plf_line_extract arrive
    with the same image; plf_line_extract compares on the device and collects.  Every call order gives the oracle's arrays:
    the reference's order over several frames (learning frame, then collected results), a DIFFERENT image handed to
    line_extract (miss), ORB-only frames (results nobody collects), lines before ORB, a batched call in between."""
    W, H = 752, 480
    f = plf.Frontend(product, max_batch=2, lsd_nfeatures=0)
    o = plf.Frontend(oracle, max_batch=2, lsd_nfeatures=0)
    imgs = [plf.synth_pair(W, H, 5200 + k) for k in range(9)]
    want = {}

    def lines(fr, side, img):
        kl, ld = fr.line_extract(side, img)
        return np.array(kl, copy=True), np.array(ld, copy=True)

    def expect(k, side):
        if (k, side) not in want:
            want[(k, side)] = lines(o, side, imgs[k][side])
        return want[(k, side)]

    def check(got, k, side, tag):
        kl, ld = expect(k, side)
        assert len(got[0]) == len(kl) and len(kl) > 100, (tag, k, side)
        assert np.array_equal(got[0], kl) and np.array_equal(got[1], ld), (tag, k, side)

    # the reference's order, five frames: ORB left, ORB right, lines left, lines right, both matchers
    for k in range(5):
        L, R = imgs[k]
        kp = f.orb_extract(0, L)[1]; f.orb_extract(1, R)
        gl, gr = lines(f, 0, L), lines(f, 1, R)
        check(gl, k, 0, "ref order"); check(gr, k, 1, "ref order")
        f.stereo_match_points(len(kp)); f.stereo_match_lines(len(gl[0]))
    # a different image handed to line_extract than to orb_extract of that side
    f.orb_extract(0, imgs[5][0]); f.orb_extract(1, imgs[5][1])
    check(lines(f, 0, imgs[6][0]), 6, 0, "miss"); check(lines(f, 1, imgs[5][1]), 5, 1, "after miss")
    # ORB only (whatever was started is never collected), then the reference's order again
    for k in (6, 7):
        f.orb_extract(0, imgs[k][0]); f.orb_extract(1, imgs[k][1])
    check(lines(f, 0, imgs[7][0]), 7, 0, "after orb only"); check(lines(f, 1, imgs[7][1]), 7, 1, "after orb only")
    # lines before ORB, and a batched call in between
    check(lines(f, 0, imgs[8][0]), 8, 0, "lines first"); f.orb_extract(0, imgs[8][0])
    Lb, Rb = plf.synth_batch(W, H, [5200, 5201])
    rb = f.frontend_batch(Lb, Rb)
    assert int(rb.n_kl_left[0]) == len(expect(0, 0)[0]) and np.array_equal(rb.kl_left[0, :len(expect(0, 0)[0])], expect(0, 0)[0])
    f.orb_extract(1, imgs[8][1]); check(lines(f, 1, imgs[8][1]), 8, 1, "after batch")
    for k in range(3):
        L, R = imgs[k]
        f.orb_extract(0, L); f.orb_extract(1, R)
        check(lines(f, 0, L), k, 0, "again"); check(lines(f, 1, R), k, 1, "again")


def test_frame_constructor_returns_before_the_matchers(plf, product, oracle):
    """src/Frame.cc:147-150: no keypoints or no keylines in the LEFT image -> neither stereo matcher runs and the match arrays keep
    their initial values.  A ramp (lines, no corners), isolated dots (corners, no lines), stripes and an ordinary pair in one
    batched call, every array against the oracle."""
    W, H = 752, 480
    yy, xx = np.mgrid[0:H, 0:W]
    dots = np.full((H, W), 60, np.uint8)
    for cy in range(40, H - 40, 57):
        for cx in range(40, W - 40, 61):
            dots[cy - 1:cy + 2, cx - 1:cx + 2] = 250
    ordinary = plf.synth_pair(W, H, 77)
    L = np.stack([((xx * 3) % 256).astype(np.uint8), dots, (((xx // 5) % 2) * 255).astype(np.uint8), ordinary[0]])
    R = np.stack([np.roll(L[0], -3, axis=1), np.roll(dots, -4, axis=1), np.roll(L[2], -3, axis=1), ordinary[1]])
    f = plf.Frontend(product, max_batch=4, lsd_nfeatures=0)
    o = plf.Frontend(oracle, max_batch=4, lsd_nfeatures=0)
    for rep in range(3):                          # eager pass, graph capture, graph replay
        rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
        assert int(ro.n_kp_left[0]) == 0 and int(ro.n_kl_left[0]) > 0          # ramp: lines only
        assert int(ro.n_kp_left[1]) > 0 and int(ro.n_kl_left[1]) == 0          # dots: corners only
        assert int(ro.n_kp_left[3]) > 500 and int(ro.n_kl_left[3]) > 100
        for b in range(4):
            nk, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
            assert int(rg.n_kp_left[b]) == nk and int(rg.n_kl_left[b]) == nl, (rep, b)
            assert np.array_equal(rg.kp_left[b, :nk], ro.kp_left[b, :nk]) and np.array_equal(rg.kl_left[b, :nl], ro.kl_left[b, :nl]), (rep, b)
            assert np.array_equal(rg.u_right[b, :nk], ro.u_right[b, :nk]) and np.array_equal(rg.depth[b, :nk], ro.depth[b, :nk]), (rep, b)
            assert np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl]), (rep, b)
            assert np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl]) and np.allclose(rg.le[b, :nl], ro.le[b, :nl], rtol=1e-12, atol=0), (rep, b)
        assert (ro.u_right[1, :int(ro.n_kp_left[1])] == -1).all() and (ro.line_match12[0, :int(ro.n_kl_left[0])] == -1).all()
