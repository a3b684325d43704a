"""Stage times of one context at increasing batch sizes (how the one-warp-per-image kernels scale with occupancy)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, plf
L, R = plf.synth_batch(752, 480, [1000 + i for i in range(16)])
prod = plf.load_product()
for B in [int(a) for a in sys.argv[1:]] or [64, 256, 1024]:
    idx = np.arange(B) % 16
    f = plf.Frontend(prod, max_batch=B, lsd_nfeatures=300)
    out = f.new_result(B)
    f.set_stage_timing(True)
    for _ in range(2):
        f.frontend_batch(L[idx], R[idx], out)
    ms = f.stage_ms()
    tot = sum(v for k, v in ms.items() if k not in ("h2d", "d2h"))
    print(B, "pairs: total %.1f ms -> %.0f pairs/s | grow %.1f fast %.2f order %.2f grad %.2f orient %.2f blur %.2f lbd %.2f" % (
        tot, B / tot * 1e3, ms["lsd_grow"], ms["orb_fast"], ms["lsd_order"], ms["lsd_gradient"], ms["orb_orient_desc"],
        ms["orb_blur"], ms["lbd_descriptor"] + ms["lbd_blur_sobel"]))
    f.close()
