"""Runs one small batch through the CUDA path (developer tool for ncu captures on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L, R = plf.synth_batch(752, 480, [1000 + i for i in range(min(B, 8))])
idx = np.arange(B) % len(L)
f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300)
out = f.new_result(B)
for _ in range(reps):
    f.frontend_batch(L[idx], R[idx], out)
print("ok", out.n_kp_left[:4], out.n_kl_left[:4])
