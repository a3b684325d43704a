"""Runs the raw-frame upload (H2D + rectification kernel) of one batch (developer tool for ncu captures on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L, R = plf.synth_batch(752, 480, [1000 + i for i in range(8)])
idx = np.arange(B) % 8
f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300)
for side in (0, 1):
    f.rectify_set_maps(side, *plf.rectify_maps(752, 480, side))
for _ in range(3):
    f.batch_upload_raw(L[idx], R[idx])
f.sync()
print("ok")
