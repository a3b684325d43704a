"""Degenerate and adversarial images through the streaming region grower (single-image and small batched calls) against the
oracle: blank, constant ramp (one region larger than a warp's record buffer), noise (thousands of tiny regions), checkerboard
(corners everywhere), stripes, concentric rings, a ramp with noise.  python tools/sw_adversarial.py [refine [W H]]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 0
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (752, 480)
rng = np.random.default_rng(7)
yy, xx = np.mgrid[0:H, 0:W]
imgs = {
    "blank": np.full((H, W), 90, np.uint8),
    "ramp_x": (xx * 255 // (W - 1)).astype(np.uint8),
    "ramp_diag": ((xx + 2 * yy) % 256).astype(np.uint8),
    "ramp_steep": ((xx * 3) % 256).astype(np.uint8),
    "noise": rng.integers(0, 256, (H, W)).astype(np.uint8),
    "noise_soft": np.clip(128 + 20 * rng.standard_normal((H, W)), 0, 255).astype(np.uint8),
    "checker8": (((xx // 8 + yy // 8) % 2) * 200 + 20).astype(np.uint8),
    "checker31": (((xx // 31 + yy // 31) % 2) * 180 + 40).astype(np.uint8),
    "stripes_v": (((xx // 5) % 2) * 255).astype(np.uint8),
    "stripes_d": ((((xx + yy) // 9) % 2) * 220 + 10).astype(np.uint8),
    "rings": (128 + 120 * np.sin(np.hypot(xx - W / 2, yy - H / 2) / 6.0)).astype(np.uint8),
    "ramp_noise": np.clip((xx * 255 // (W - 1)) + 6 * rng.standard_normal((H, W)), 0, 255).astype(np.uint8),
    "one_edge": np.where(xx + 0.37 * yy < 400, 40, 210).astype(np.uint8),
}
names = list(imgs)
f = plf.Frontend(plf.load_product(), width=W, height=H, max_batch=len(names), lsd_nfeatures=0, lsd_refine=refine)
o = plf.Frontend(plf.load_oracle(), width=W, height=H, max_batch=len(names), lsd_nfeatures=0, lsd_refine=refine)
bad = 0
for name in names:
    im = np.ascontiguousarray(imgs[name])
    t0 = time.perf_counter()
    perr = oerr = None
    try:
        kl, ld = f.line_extract(0, im)
    except Exception as e:
        perr = str(e)
    t1 = time.perf_counter()
    try:
        klo, ldo = o.line_extract(0, im)
    except Exception as e:
        oerr = str(e)
    if perr or oerr:
        agree = bool(perr) and bool(oerr)
        bad += not agree
        print("%-12s product: %s | oracle: %s -> %s" % (name, perr or "ok", oerr or "ok", "both refuse" if agree else "DIFFERENT"), flush=True)
        continue
    same = len(kl) == len(klo) and np.array_equal(np.array(kl), np.array(klo)) and np.array_equal(np.array(ld), np.array(ldo))
    bad += not same
    print("%-12s %5d lines, %7.1f ms on the GPU, %s" % (name, len(klo), (t1 - t0) * 1e3, "equal" if same else "DIFFERENT (%d vs %d)" % (len(kl), len(klo))), flush=True)
# all of them in one batched call (left = image, right = image shifted by 3 px)
names = [n for n in names if n != "checker8"]          # (more segments than the context's capacity: both sides refuse it, see above)
L = np.stack([imgs[n] for n in names]); R = np.roll(L, -3, axis=2)
try:
    rg, ro = f.frontend_batch(L, R), o.frontend_batch(L, R)
    for b, name in enumerate(names):
        n = int(ro.n_kl_left[b])
        ok = int(rg.n_kl_left[b]) == n and np.array_equal(rg.kl_left[b, :n], ro.kl_left[b, :n]) and np.array_equal(rg.line_match12[b, :n], ro.line_match12[b, :n])
        nk, nkr = int(ro.n_kp_left[b]), int(ro.n_kp_right[b])
        okp = (int(rg.n_kp_left[b]) == nk and int(rg.n_kp_right[b]) == nkr and np.array_equal(rg.kp_left[b, :nk], ro.kp_left[b, :nk])
               and np.array_equal(rg.kp_right[b, :nkr], ro.kp_right[b, :nkr]) and np.array_equal(rg.desc_left[b, :nk], ro.desc_left[b, :nk])
               and np.array_equal(rg.u_right[b, :nk], ro.u_right[b, :nk]) and np.array_equal(rg.depth[b, :nk], ro.depth[b, :nk])
               and np.array_equal(rg.disp_se[b, :n], ro.disp_se[b, :n]))
        print("batched %-12s %5d keypoints %5d lines: %s" % (name, nk, n, "equal" if ok and okp else "DIFFERENT (lines %s, points %s)" % (ok, okp)), flush=True)
        ok = ok and okp
        bad += not ok
        if not ok:
            print("batched %-12s DIFFERENT" % name)
    print("batched call over all %d images: compared" % len(names))
except Exception as e:
    print("batched call error:", e); bad += 1
print("differences:", bad)
