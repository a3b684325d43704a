"""TEST INFRASTRUCTURE ONLY — writes tests/golden/*.npz.

Two kinds of vectors, both produced in this container and committed so that the GPU box (which has neither
/root/reference nor a guarantee about cv2) can check against them:
  * cv2_*   : computed with REAL OpenCV 4.13 primitives through the Tier A restatement (oracle/python/tier_a.py):
              pyramid checksums, per-cell FAST candidate lists, LSD segments.
  * oracle_*: full outputs of the Tier B oracle (oracle/cpp), itself pinned against cv2 by tests/test_oracle_cv2.py.
Run:  python oracle/python/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import plf  # noqa: E402
import tier_a  # noqa: E402

CASES = [("c1_752x480_seed1", 752, 480, 1, dict()),
         ("c3_752x480_seed1000_2000feat", 752, 480, 1000, dict(n_features=2000, has_lines=0)),
         ("c4_1280x720_seed2000", 1280, 720, 2000, dict())]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    orc = plf.load_oracle()
    for name, W, H, seed, kw in CASES:
        L, R = plf.synth_pair(W, H, seed)
        d = {"W": W, "H": H, "seed": seed, "img_crc": np.array([zlib.crc32(L.tobytes()), zlib.crc32(R.tobytes())], np.uint32)}
        pyr = tier_a.pyramid(L)
        d["cv2_pyr_crc"] = np.array([zlib.crc32(np.ascontiguousarray(p).tobytes()) for p in pyr], np.uint32)
        for l in range(8):
            d["cv2_cand_L%d" % l] = tier_a.level_candidates(pyr[l]).astype(np.uint16 if False else np.float32)
        if kw.get("has_lines", 1):
            d["cv2_lsd_left"] = tier_a.lsd_segments(L)
            d["cv2_lsd_right"] = tier_a.lsd_segments(R)
        f = plf.Frontend(orc, width=W, height=H, max_batch=1, **kw)
        r = f.frontend_batch(L[None], R[None])
        nl, nr = int(r.n_kp_left[0]), int(r.n_kp_right[0])
        kl, kr = int(r.n_kl_left[0]), int(r.n_kl_right[0])
        d.update(oracle_kp_left=r.kp_left[0, :nl], oracle_kp_right=r.kp_right[0, :nr],
                 oracle_desc_left=r.desc_left[0, :nl], oracle_desc_right=r.desc_right[0, :nr],
                 oracle_u_right=r.u_right[0, :nl], oracle_depth=r.depth[0, :nl],
                 oracle_kl_left=r.kl_left[0, :kl], oracle_kl_right=r.kl_right[0, :kr],
                 oracle_ldesc_left=r.ldesc_left[0, :kl], oracle_ldesc_right=r.ldesc_right[0, :kr],
                 oracle_disp_se=r.disp_se[0, :kl], oracle_le=r.le[0, :kl], oracle_line_match12=r.line_match12[0, :kl])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "kps", nl, nr, "lines", kl, kr, "stereo pts", int((r.u_right[0, :nl] >= 0).sum()))


if __name__ == "__main__":
    main()
