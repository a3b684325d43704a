"""Stage time of the LSD region grower for one context at a given batch size, for the grower selected by
PLF_LSD_GROWER (unset: product default; seq | lane | stream).  python tools/grow_modes.py [pairs] [distinct]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L, R = plf.synth_batch(752, 480, [3000 + i for i in range(D)])
idx = np.arange(B) % D
f = plf.Frontend(plf.load_product(), max_batch=B, lsd_nfeatures=300)
out = f.new_result(B)
f.set_stage_timing(True)
for _ in range(3):
    f.frontend_batch(L[idx], R[idx], out)
ms = f.stage_ms()
print(os.environ.get("PLF_LSD_GROWER", "default"), B, "pairs: grow %.2f ms, keylines %.2f ms; lines %d" % (ms["lsd_grow"], ms["line_keylines"], int(out.n_kl_left[:B].sum())))
