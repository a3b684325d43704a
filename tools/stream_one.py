import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, plf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L, R = plf.synth_batch(752, 480, [1 + i for i in range(min(B, 16))])
prod = plf.load_product()
idx = np.arange(B) % min(B, 16)
f = plf.Frontend(prod, max_batch=B, lsd_nfeatures=300)
out = f.new_result(B)
f.set_stage_timing(True)
for _ in range(2):
    f.frontend_batch(L[idx], R[idx], out)
ms = f.stage_ms()
print(B, 'grow ms', ms['lsd_grow'])
