"""Exact comparison of the CUDA path with the CPU oracle over many synthetic pairs (GPU box; the oracle is the checker).

  python tools/parity_sweep.py [n_pairs] [first_seed] [kind] [lsd_refine] [W] [H] [batch]     kind: rect (default) | curvy; batch <= 4 exercises the small-batch grower
Prints, per output array, the number of pairs on which it differs from the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
KIND = sys.argv[3] if len(sys.argv) > 3 else "rect"
REFINE = int(sys.argv[4]) if len(sys.argv) > 4 else 0
W = int(sys.argv[5]) if len(sys.argv) > 5 else 752
H = int(sys.argv[6]) if len(sys.argv) > 6 else 480
B = int(sys.argv[7]) if len(sys.argv) > 7 else 64
kw = dict(width=W, height=H, max_batch=B, lsd_nfeatures=0, lsd_refine=REFINE)
f = plf.Frontend(plf.load_product(), **kw)
o = plf.Frontend(plf.load_oracle(), **kw)
fields = [("kp_left", "n_kp_left"), ("desc_left", "n_kp_left"), ("kp_right", "n_kp_right"), ("desc_right", "n_kp_right"),
          ("kl_left", "n_kl_left"), ("ldesc_left", "n_kl_left"), ("kl_right", "n_kl_right"), ("ldesc_right", "n_kl_right"),
          ("u_right", "n_kp_left"), ("depth", "n_kp_left"), ("disp_se", "n_kl_left"), ("line_match12", "n_kl_left")]
bad = {k: 0 for k, _ in fields}
bad.update(n_kp_left=0, n_kp_right=0, n_kl_left=0, n_kl_right=0)
bad_pairs = set()
tg = to = 0.0
nkp = nkl = 0
for c0 in range(0, N, B):
    seeds = list(range(S0 + c0, S0 + min(c0 + B, N)))
    if KIND == "curvy":
        L = np.stack([plf.synth_curvy(W, H, s) for s in seeds])
        rng = np.random.default_rng(S0 + c0)
        R = np.clip(np.roll(L, -14, axis=2).astype(np.int16) + rng.integers(-2, 3, L.shape), 0, 255).astype(np.uint8)
    else:
        L, R = plf.synth_batch(W, H, seeds)
    t = time.time(); rg = f.frontend_batch(L, R); tg += time.time() - t
    t = time.time(); ro = o.frontend_batch(L, R); to += time.time() - t
    for b in range(len(seeds)):
        for cnt in ("n_kp_left", "n_kp_right", "n_kl_left", "n_kl_right"):
            if int(getattr(rg, cnt)[b]) != int(getattr(ro, cnt)[b]):
                bad[cnt] += 1; bad_pairs.add(seeds[b])
        for name, cnt in fields:
            n = int(getattr(ro, cnt)[b])
            if not np.array_equal(getattr(rg, name)[b, :n], getattr(ro, name)[b, :n]):
                bad[name] += 1; bad_pairs.add(seeds[b])
        nkp += int(ro.n_kp_left[b]) + int(ro.n_kp_right[b]); nkl += int(ro.n_kl_left[b]) + int(ro.n_kl_right[b])
print("%dx%d refine %d: " % (W, H, REFINE), end="")
print("%d pairs (%s, seeds %d..%d): %d keypoints, %d lines compared; GPU %.1f s, oracle %.1f s" % (N, KIND, S0, S0 + N - 1, nkp, nkl, tg, to))
print("pairs with any difference:", len(bad_pairs), sorted(bad_pairs)[:20])
print({k: v for k, v in bad.items() if v})
