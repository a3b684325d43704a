"""Runs one batch through the CUDA path (developer tool for ncu captures on the GPU box).
  python tools/prof_one.py [pairs] [reps] [W] [H] [n_features] [lsd_nfeatures] [has_points] [has_lines] [distinct]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
a = [int(x) for x in sys.argv[1:]] + [None] * 9
B, reps = a[0] or 4, a[1] or 2
W, H = a[2] or 752, a[3] or 480
nf, nl = (1200 if a[4] is None else a[4]), (300 if a[5] is None else a[5])
hp, hl = (1 if a[6] is None else a[6]), (1 if a[7] is None else a[7])
D = min(B, a[8] or 64)
L, R = plf.synth_batch(W, H, [1000 + i for i in range(D)])
idx = np.arange(B) % D
f = plf.Frontend(plf.load_product(), width=W, height=H, max_batch=B, n_features=nf, lsd_nfeatures=nl, has_points=hp, has_lines=hl)
out = f.new_result(B)
for _ in range(reps):
    f.frontend_batch(L[idx], R[idx], out)
print("ok", out.n_kp_left[:4], out.n_kl_left[:4])
