"""The CPU oracle against the committed golden vectors (tests/golden, written by oracle/python/make_golden.py):
cv2_* arrays come from real OpenCV 4.13 primitives, oracle_* arrays freeze the oracle's own outputs.  CPU only."""
import glob
import os
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def case_kwargs(name):
    kw = {}
    if "2000feat" in name:
        kw.update(n_features=2000, has_lines=0)
    return kw


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_reproduces_golden(plf, oracle, path):
    g = np.load(path)
    W, H, seed = int(g["W"]), int(g["H"]), int(g["seed"])
    L, R = plf.synth_pair(W, H, seed)
    assert zlib.crc32(L.tobytes()) == int(g["img_crc"][0]) and zlib.crc32(R.tobytes()) == int(g["img_crc"][1]), \
        "synthetic generator is not reproducible on this machine"
    f = plf.Frontend(oracle, width=W, height=H, max_batch=1, **case_kwargs(path))
    r = f.frontend_batch(L[None], R[None])
    n, nr = int(r.n_kp_left[0]), int(r.n_kp_right[0])
    assert n == len(g["oracle_kp_left"]) and nr == len(g["oracle_kp_right"])
    assert np.array_equal(r.kp_left[0, :n], g["oracle_kp_left"])
    assert np.array_equal(r.desc_left[0, :n], g["oracle_desc_left"])
    assert np.array_equal(r.desc_right[0, :nr], g["oracle_desc_right"])
    assert np.array_equal(r.u_right[0, :n], g["oracle_u_right"])
    assert np.array_equal(r.depth[0, :n], g["oracle_depth"])
    nl = int(r.n_kl_left[0])
    assert nl == len(g["oracle_kl_left"])
    assert np.array_equal(r.kl_left[0, :nl], g["oracle_kl_left"])
    assert np.array_equal(r.ldesc_left[0, :nl], g["oracle_ldesc_left"])
    assert np.array_equal(r.line_match12[0, :nl], g["oracle_line_match12"])
    assert np.array_equal(r.disp_se[0, :nl], g["oracle_disp_se"])
    # cv2-derived vectors: pyramid bytes, per-cell FAST lists, LSD segments
    for l in range(8):
        assert zlib.crc32(f.pyramid_level(0, l, slot=0).tobytes()) == int(g["cv2_pyr_crc"][l])
        assert np.array_equal(f.fast_candidates(0, l), g["cv2_cand_L%d" % l])
    if "cv2_lsd_left" in g:
        assert np.array_equal(f.lsd_segments(0), g["cv2_lsd_left"])
        assert np.array_equal(f.lsd_segments(1), g["cv2_lsd_right"])
