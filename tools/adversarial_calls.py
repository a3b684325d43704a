"""Degenerate stereo pairs through the single-image entry points (the reference's five signatures) and through line-only /
point-only contexts, every output against the oracle.  python tools/adversarial_calls.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
W, H = 752, 480
rng = np.random.default_rng(11)
yy, xx = np.mgrid[0:H, 0:W]
dots = np.full((H, W), 60, np.uint8)
for cy in range(40, H - 40, 57):
    for cx in range(40, W - 40, 61):
        dots[cy - 1:cy + 2, cx - 1:cx + 2] = 250
imgs = {
    "blank": np.full((H, W), 90, np.uint8),
    "ramp_steep": ((xx * 3) % 256).astype(np.uint8),
    "dots": dots,
    "noise": rng.integers(0, 256, (H, W)).astype(np.uint8),
    "checker31": (((xx // 31 + yy // 31) % 2) * 180 + 40).astype(np.uint8),
    "stripes_v": (((xx // 5) % 2) * 255).astype(np.uint8),
    "stripes_d": ((((xx + yy) // 9) % 2) * 220 + 10).astype(np.uint8),
    "rings": (128 + 120 * np.sin(np.hypot(xx - W / 2, yy - H / 2) / 6.0)).astype(np.uint8),
    "ordinary": plf.synth_pair(W, H, 5)[0],
}
bad = 0
def eq(a, b):
    return len(a) == len(b) and np.array_equal(np.asarray(a), np.asarray(b))
f = plf.Frontend(plf.load_product(), max_batch=1, lsd_nfeatures=0)
o = plf.Frontend(plf.load_oracle(), max_batch=1, lsd_nfeatures=0)
for name, L in imgs.items():
    L = np.ascontiguousarray(L); R = np.ascontiguousarray(np.roll(L, -4, axis=1))
    res = []
    for fr in (f, o):
        try:
            m, k, d = fr.orb_extract(0, L); m2, k2, d2 = fr.orb_extract(1, R)
            kl, ld = fr.line_extract(0, L); klr, ldr = fr.line_extract(1, R)
            u, dep = fr.stereo_match_points(len(k)); disp, le, m12 = fr.stereo_match_lines(len(kl))
            n1, mm1 = fr.match_nnr(ld, ldr, 0.9) if len(ld) and len(ldr) else (0, np.zeros(0, np.int32))
            res.append(dict(k=np.array(k), d=np.array(d), k2=np.array(k2), kl=np.array(kl), ld=np.array(ld), klr=np.array(klr), u=np.array(u), dep=np.array(dep),
                            disp=np.array(disp), m12=np.array(m12), mm1=np.array(mm1), le=np.array(le)))
        except Exception as e:
            res.append(str(e))
    if isinstance(res[0], str) or isinstance(res[1], str):
        agree = isinstance(res[0], str) and isinstance(res[1], str)
        bad += not agree
        print("%-11s product: %s | oracle: %s" % (name, res[0] if isinstance(res[0], str) else "ok", res[1] if isinstance(res[1], str) else "ok"))
        continue
    diffs = [key for key in res[1] if not (np.allclose(res[0][key], res[1][key], rtol=1e-12, atol=0) if key == "le" and res[0][key].shape == res[1][key].shape else eq(res[0][key], res[1][key]))]
    bad += bool(diffs)
    print("%-11s five signatures: %4d kp %4d lines: %s" % (name, len(res[1]["k"]), len(res[1]["kl"]), "equal" if not diffs else "DIFFERENT " + str(diffs)), flush=True)
# line-only and point-only contexts, batched
names = list(imgs)
Lb = np.stack([imgs[n] for n in names]); Rb = np.roll(Lb, -4, axis=2)
for kw in (dict(has_points=0), dict(has_lines=0), dict(n_features=100, n_levels=4, lsd_nfeatures=20)):
    fp = plf.Frontend(plf.load_product(), max_batch=len(names), **kw)
    op = plf.Frontend(plf.load_oracle(), max_batch=len(names), **kw)
    rg, ro = fp.frontend_batch(Lb, Rb), op.frontend_batch(Lb, Rb)
    for b, name in enumerate(names):
        nk, nl = int(ro.n_kp_left[b]), int(ro.n_kl_left[b])
        ok = (int(rg.n_kp_left[b]) == nk and int(rg.n_kl_left[b]) == nl and np.array_equal(rg.kp_left[b, :nk], ro.kp_left[b, :nk]) and
              np.array_equal(rg.u_right[b, :nk], ro.u_right[b, :nk]) and np.array_equal(rg.kl_left[b, :nl], ro.kl_left[b, :nl]) and
              np.array_equal(rg.ldesc_left[b, :nl], ro.ldesc_left[b, :nl]) and np.array_equal(rg.line_match12[b, :nl], ro.line_match12[b, :nl]) and
              np.array_equal(rg.disp_se[b, :nl], ro.disp_se[b, :nl]))
        bad += not ok
        if not ok:
            print("batched %s %-11s DIFFERENT (%d kp, %d lines)" % (kw, name, nk, nl))
    print("batched", kw, "compared")
print("differences:", bad)

# the widened entry points on degenerate batches (no keypoints / no lines / ordinary), against the oracle
names2 = ["blank", "ramp_steep", "dots", "ordinary"]
Lb = np.stack([imgs[n] for n in names2]); Rb = np.roll(Lb, -4, axis=2)
B = len(names2)
outs = []
for lib in (plf.load_product(), plf.load_oracle()):
    fr = plf.Frontend(lib, max_batch=B)
    res = fr.frontend_batch(Lb, Rb)
    st, ix = fr.feature_grid(0, B)
    x3d, l3d = fr.backproject(np.tile(np.eye(3, dtype=np.float32), (B, 1, 1)), np.zeros((B, 3), np.float32), 435.2, 367.4, 252.2)
    fr.bow_set_vocabulary(0, plf.synth_vocabulary(10, 4, seed=1)); fr.bow_set_vocabulary(1, plf.synth_vocabulary(6, 3, seed=2, ragged=0.0))
    bw = fr.bow_transform(0, B); bl = fr.bow_transform(1, B)
    outs.append(dict(st=np.array(st), ix=np.array(ix), x3d=np.array(x3d), l3d=np.array(l3d), bw=[np.array(a) for a in bw], bl=[np.array(a) for a in bl]))
for key in ("st", "ix", "x3d", "l3d"):
    if key == "ix":      # only the first st[b, -1] indices of a slot are defined
        same = all(np.array_equal(outs[0]["ix"][b, :int(outs[1]["st"][b, -1])], outs[1]["ix"][b, :int(outs[1]["st"][b, -1])]) for b in range(B))
    else:
        same = outs[0][key].shape == outs[1][key].shape and np.array_equal(outs[0][key], outs[1][key])
    bad += not same
    print("widened %-4s: %s" % (key, "equal" if same else "DIFFERENT"))
for key in ("bw", "bl"):
    same = len(outs[0][key]) == len(outs[1][key]) and all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(outs[0][key], outs[1][key]))
    bad += not same
    print("widened %-4s: %s" % (key, "equal" if same else "DIFFERENT"))
print("differences (all):", bad)
