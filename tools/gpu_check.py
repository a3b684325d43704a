"""Stage-by-stage diff of the CUDA path against the CPU oracle on one synthetic pair (developer tool, GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plf

W, H, seed = 752, 480, 1
if len(sys.argv) > 3:
    W, H, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
L, R = plf.synth_pair(W, H, seed)
prod, orc = plf.load_product(), plf.load_oracle()
kw = dict(width=W, height=H, max_batch=2)
f = plf.Frontend(prod, **kw)
o = plf.Frontend(orc, **kw)


def seg_match(a, b, tol):
    if len(a) == 0:
        return 1.0
    hit = 0
    for s in a:
        d1 = np.maximum(np.hypot(b[:, 0] - s[0], b[:, 1] - s[1]), np.hypot(b[:, 2] - s[2], b[:, 3] - s[3]))
        d2 = np.maximum(np.hypot(b[:, 2] - s[0], b[:, 3] - s[1]), np.hypot(b[:, 0] - s[2], b[:, 1] - s[3]))
        hit += (np.minimum(d1, d2).min() <= tol) if len(b) else 0
    return hit / len(a)


for side, img in ((0, L), (1, R)):
    t = time.time(); mg, kg, dg = f.orb_extract(side, img); tg = time.time() - t
    mo, ko, do = o.orb_extract(side, img)
    print("side", side, "orb gpu %.1f ms" % (tg * 1e3), "n", len(kg), len(ko), "mono", mg, mo)
    for l in range(8):
        pe = np.array_equal(f.pyramid_level(side, l), o.pyramid_level(side, l))
        be = np.array_equal(f.blurred_level(side, l), o.blurred_level(side, l))
        cg, co = f.fast_candidates(side, l), o.fast_candidates(side, l)
        ce = cg.shape == co.shape and np.array_equal(cg, co)
        print("  level", l, "pyr", pe, "blur", be, "cand", ce, len(cg), len(co))
    if len(kg) == len(ko):
        for fld in kg.dtype.names:
            print("  kp", fld, "equal" if np.array_equal(kg[fld], ko[fld]) else "DIFF %d" % (kg[fld] != ko[fld]).sum())
        print("  desc rows differing", (dg != do).any(axis=1).sum())
ug, dg_ = f.stereo_match_points(len(kg))
uo, do_ = o.stereo_match_points(len(ko))
print("stereo points: uRight equal", np.array_equal(ug, uo), "depth equal", np.array_equal(dg_, do_), "matched", (ug >= 0).sum(), (uo >= 0).sum())
for side, img in ((0, L), (1, R)):
    t = time.time(); klg, ldg = f.line_extract(side, img); tg = time.time() - t
    klo, ldo = o.line_extract(side, img)
    print("side", side, "lines gpu %.1f ms" % (tg * 1e3), len(klg), len(klo))
    print("  scaled equal", np.array_equal(f.lsd_scaled(side), o.lsd_scaled(side)))
    ag, ao = f.lsd_angles(side), o.lsd_angles(side)
    print("  angles equal", np.array_equal(ag, ao), "diff px", (ag != ao).sum())
    sg, so = f.lsd_segments(side), o.lsd_segments(side)
    print("  segments", len(sg), len(so), "exact" if sg.shape == so.shape and np.array_equal(sg, so) else
          "recall@0.5 %.4f maxdiff %s" % (seg_match(so, sg, 0.5), np.abs(sg - so).max() if sg.shape == so.shape else "n/a"))
    if len(klg) == len(klo):
        for fld in klg.dtype.names:
            e = np.array_equal(klg[fld], klo[fld])
            print("  kl", fld, "equal" if e else "DIFF %d max %g" % ((klg[fld] != klo[fld]).sum(), np.abs(klg[fld].astype(np.float64) - klo[fld]).max()))
        bits = np.unpackbits(ldg ^ ldo, axis=1).sum(axis=1)
        print("  LBD hamming: max", bits.max(), "rows != 0:", (bits > 0).sum(), "float maxdiff", np.abs(f.lbd_float(side) - o.lbd_float(side)).max())
dg2, leg, mg2 = f.stereo_match_lines(len(klg))
do2, leo, mo2 = o.stereo_match_lines(len(klo))
if len(mg2) == len(mo2):
    print("stereo lines: m12 equal", np.array_equal(mg2, mo2), "disp equal", np.array_equal(dg2, do2), "le maxdiff", np.abs(leg - leo).max(), "matched", (mg2 >= 0).sum())
rng = np.random.default_rng(0)
d1 = rng.integers(0, 256, (300, 32), dtype=np.uint8); d2 = rng.integers(0, 256, (280, 32), dtype=np.uint8)
d2[:100] = d1[:100] ^ (rng.integers(0, 256, (100, 32), dtype=np.uint8) & rng.integers(0, 256, (100, 32), dtype=np.uint8) & 3)
for nnr in (0.9, 0.6):
    a, b = f.match_nnr(d1, d2, nnr), o.match_nnr(d1, d2, nnr)
    print("match_nnr", nnr, a[0], b[0], np.array_equal(a[1], b[1]))
    a, b = f.match(d1, d2, nnr, True), o.match(d1, d2, nnr, True)
    print("match lr ", nnr, a[0], b[0], np.array_equal(a[1], b[1]))
# batch
Lb, Rb = plf.synth_batch(W, H, [seed, seed + 1])
f.set_stage_timing(True)
t = time.time(); rg = f.frontend_batch(Lb, Rb); tg = time.time() - t
ro = o.frontend_batch(Lb, Rb)
print("batch gpu %.1f ms, launches %d, stages %s" % (tg * 1e3, f.launch_count(), f.stage_ms()))
for name in ("n_kp_left", "n_kp_right", "n_kl_left", "n_kl_right"):
    print(" ", name, getattr(rg, name), getattr(ro, name))
for b in range(2):
    n = int(ro.n_kp_left[b]); nl = int(ro.n_kl_left[b])
    print("  pair", b, "kpL", np.array_equal(rg.kp_left[b, :n], ro.kp_left[b, :n]), "descL", np.array_equal(rg.desc_left[b, :n], ro.desc_left[b, :n]),
          "uR", np.array_equal(rg.u_right[b, :n], ro.u_right[b, :n]), "klL", np.array_equal(rg.kl_left[b, :nl], ro.kl_left[b, :nl]),
          "ldescL", np.array_equal(rg.ldesc_left[b, :nl], ro.ldesc_left[b, :nl]), "m12", np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl]),
          "disp", np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl]))
